"""Attribute one launch's executed instructions / stall samples to source REGIONS through the inline chain.

ncu's source page charges an inlined instruction to its innermost line (fma2() in a header), which hides
where the time goes.  nvdisasm -gi prints the whole inline chain per instruction, so this script joins
  nvdisasm -gi -c <cubin extracted from the .so that was profiled>      (offset -> chain of source lines)
  ncu -i rep --page source --print-source sass --csv                      (address -> instructions, samples)
on the instruction offset and sums by the chain's outer frames.

usage: ncu_regions.py <rep> <launch index> <lib.so> <mangled kernel substring> [depth]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, k, so, kern = sys.argv[1:5]
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 2
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = ""
for cubin in sorted(f for f in os.listdir(tmp) if f.startswith("sim") and f.endswith(".cubin")):
  dis += subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout

chains = {}   # offset -> tuple of (file, line) outermost first
opc = {}
cur, pending, infn = (), [], False
re_file = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
re_ins = re.compile(r'/\*([0-9a-f]{4,})\*/\s+(.*?);')
for line in dis.splitlines():
  if line.startswith(".text."):
    infn = kern in line
    continue
  if not infn:
    continue
  m = re_file.search(line)
  if m:
    pending.append((os.path.basename(m.group(1)), int(m.group(2))))
    if m.group(3) is None:   # outermost frame closes the chain
      cur = tuple(reversed(pending))
      pending = []
    continue
  m = re_ins.search(line)
  if m:
    off = int(m.group(1), 16)
    chains[off] = cur
    opc[off] = m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1]

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", k, "--launch-count", "1",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
data = []
for r in rows:
  if r and r[0] == "Address":
    hdr = {h: i for i, h in enumerate(r)}
    continue
  if hdr is None or not r or not r[0].startswith("0x"):
    continue
  data.append((int(r[0], 16), int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0), r[1].strip()))
base = min(a for a, _, _, _ in data)
src_cache = {}


def src(f, ln):
  if f not in src_cache:
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "qhbm-library_b200", "csrc", f)
    src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
  L = src_cache[f]
  return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""


agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
TS = TI = 0
missing = 0
for a, s, i, text in data:
  off = a - base
  ch = chains.get(off)
  if ch is None:
    missing += i
    ch = ()
  key = ch[:depth]
  agg[key][0] += s
  agg[key][1] += i
  agg[key][2][opc.get(off, "?").split(".")[0]] += i
  TS += s
  TI += i
print(f"launch {k}: {TI} warp instructions, {TS} samples, unattributed {missing}")
for key, (s, i, oc) in sorted(agg.items(), key=lambda kv: -kv[1][int(os.environ.get("SORTCOL", "0"))]):
  if s < float(os.environ.get("MINFRAC", "0.004")) * TS:
    continue
  label = " > ".join(f"{f.replace('sim_kernels.cuh', '')}:{ln}" for f, ln in key)
  top = ", ".join(f"{o} {100 * v / max(i, 1):.0f}%" for o, v in oc.most_common(5))
  print(f"samp {100 * s / TS:5.1f}%  inst {100 * i / TI:5.1f}%  {label}")
  for f, ln in key:
    print(f"        | {ln}: {src(f, ln)}")
  print(f"        ops: {top}")
