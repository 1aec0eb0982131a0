"""Summarise an .ncu-rep: per-launch key metrics + stall/opcode mix (reads with the ncu CLI)."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
for r in rows[2:]:
  print("--- launch", r[idx["ID"]], r[idx["Kernel Name"]][:50], "grid", r[idx.get("Grid Size", 0)])
  for w in want:
    if w in idx:
      print(f"    {w:75s} {r[idx[w]]} {rows[1][idx[w]]}")
n = len(rows) - 2
for k in range(n):
  src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                       capture_output=True, text=True).stdout
  srows = list(csv.reader(io.StringIO(src)))
  if len(srows) < 3:
    continue
  sh = srows[1]
  si = {h: i for i, h in enumerate(sh)}
  stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
  tot, instr, samp = collections.Counter(), collections.Counter(), collections.Counter()

  def I(x):
    try:
      return int(x)
    except Exception:
      return 0

  for r in srows[2:]:
    if len(r) < len(sh) or r[0] == "Address":
      continue
    toks = r[si["Source"]].split()
    if not toks:
      continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    samp[op] += I(r[si["# Samples"]])
    instr[op] += I(r[si["Instructions Executed"]])
    for s in stalls:
      tot[s] += I(r[si[s]])
  T = sum(tot.values()) or 1
  TI = sum(instr.values()) or 1
  print(f"=== launch {k}: stall mix (samples {T})")
  print("   ", ", ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in tot.most_common(9)))
  print("    opcode mix:", ", ".join(f"{o} {100 * v / TI:.1f}% (stall {100 * samp[o] / T:.1f}%)" for o, v in instr.most_common(14)))
