import os, subprocess, sys
for chunk in (37, 74, 148, 296, 592):
  for T in (13, 12):
    env = dict(os.environ, QHBM_CHUNK=str(chunk))
    out = subprocess.run([sys.executable, "scripts/profile_case.py", "16", "2", "4096", str(T), "4", "1", "xxz", "4"],
                         env=env, capture_output=True, text=True)
    # time the 4 reps with a wrapper: use nvidia event timing inside profile_case? simpler: python -X
    print(chunk, T, out.stdout.strip()[-120:], out.stderr.strip()[-200:])
