import os, subprocess, sys
for so in (0, 1, 2, 4, 8):
  env = dict(os.environ, QHBM_SYNC_OPS=str(so))
  out = subprocess.run([sys.executable, "scripts/profile_case.py", "16", "2", "4096", "13", "4", "1", "xxz", "4"],
                       env=env, capture_output=True, text=True)
  print(so, out.stdout.strip()[-60:], out.stderr.strip()[-200:])
