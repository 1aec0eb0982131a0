"""Per-source-line instruction / stall-sample totals of one launch of an .ncu-rep (needs --import-source on)."""
import csv
import io
import subprocess
import sys

rep, k = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", k, "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []
for r in rows:
  if r and r[0] == "Line No":
    hdr = {h: i for i, h in enumerate(r)}
    continue
  if hdr is None or len(r) < 10 or not r[0]:
    continue
  try:
    ln = int(r[0])
  except ValueError:
    continue
  def I(x):
    try:
      return int(x)
    except ValueError:
      return 0
  lines.append((ln, r[1], I(r[hdr["# Samples"]]), I(r[hdr["Instructions Executed"]]),
                I(r[hdr["L1 Wavefronts Shared"]]), I(r[hdr["L1 Wavefronts Shared Excessive"]])))
TS = sum(x[2] for x in lines) or 1
TI = sum(x[3] for x in lines) or 1
print(f"total samples {TS}  total warp instructions {TI}")
if len(sys.argv) > 4 and sys.argv[4] == "byline":
  lines.sort()
else:
  lines.sort(key=lambda x: -x[2])
  lines = lines[:top]
for ln, src, s, i, w, we in lines:
  print(f"{ln:5d} samp {100*s/TS:5.1f}% inst {100*i/TI:5.1f}% smem_wf {w:>11d} exc {we:>10d} | {src.strip()[:110]}")
