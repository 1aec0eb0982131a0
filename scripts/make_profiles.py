"""Turns the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/."""
import collections
import csv
import json
import shutil
import subprocess
import sys

rows = list(csv.reader(l for l in open("gpurun_out/bench_launches_r1.csv") if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launch = collections.OrderedDict()
for r in rows[1:]:
  d = launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
  d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
L = list(launch.values())
to_b = lambda v, u: v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
to_ns = lambda v, u: v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
for d in L:
  d["ns"] = to_ns(*d["gpu__time_duration.sum"])
  d["rd"] = to_b(*d["dram__bytes_read.sum"])
  d["wr"] = to_b(*d["dram__bytes_write.sum"])
sw = [d for d in L if "sweep_kernel" in d["name"]]
per_step = 4
step = sw[-per_step:]
tot_ns = sum(d["ns"] for d in step)
tot_b = sum(d["rd"] + d["wr"] for d in step)
labels = ["fwd sweep 1 (init basis, passes, store psi)", "fwd sweep 2",
          "expectation (observable passes + generic x-groups) + bwd sweep 1 (passes + gradient inner products)",
          "bwd sweep 2 (no store)"]
step_ms = json.load(open("gpurun_out/bench_c3.json"))["ms_per_step"]
with open("profiles/r1_bench_launches.md", "w") as f:
  f.write("# Round 1 (final): ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (config c3)\n\n")
  f.write("Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
          "--csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline`\n")
  f.write("(per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes).\n\n")
  f.write("One step = 4096 unique 16-qubit bitstrings in ONE chunk = 4 launches of `qhbm::sweep_kernel<4,true>` "
          "(grid 4096 x 16 tiles = 65536 CTAs, 256 threads, 64 KiB dynamic + 10 KiB static smem). Last step:\n\n")
  f.write("| # | launch | grid | duration us | share | dram read MB | dram write MB |\n|---|---|---|---|---|---|---|\n")
  for i, d in enumerate(step):
    f.write(f"| {i} | {labels[i]} | {d['grid']} | {d['ns'] / 1e3:.1f} | {100 * d['ns'] / tot_ns:.1f}% | {d['rd'] / 1e6:.1f} | "
            f"{d['wr'] / 1e6:.1f} |\n")
  f.write(f"\nSweep-kernel time per step under ncu: {tot_ns / 1e6:.2f} ms (bench.py, CUDA events, not under ncu: see "
          f"r1_bench_c3_n1.json). DRAM traffic per step: {tot_b / 1e9:.2f} GB = {tot_b / 4096 / 2**20:.2f} MiB per bitstring vs "
          f"91 MiB algorithmic (SURVEY 8d) => the state is reused on chip; at the measured step time this is "
          f"{tot_b / 1e9 / (step_ms / 1e3):.0f} GB/s = {100 * tot_b / 1e9 / (step_ms / 1e3) / 6551.7:.0f}% of the measured HBM peak.\n\n")
  tot_all = sum(d["ns"] for d in L)
  names = collections.Counter()
  for d in L:
    if "sweep_kernel" not in d["name"]:
      names[d["name"][:70]] += d["ns"]
  f.write("Share of all GPU time in the run by kernel (the kernel's share of the step agrees with bench.py's "
          "`gpu_launches`: 4 sweep launches + prep + finalize + weighted sum):\n\n| kernel | total ms | share |\n|---|---|---|\n")
  f.write(f"| qhbm::sweep_kernel<4,true> | {sum(d['ns'] for d in sw) / 1e6:.2f} | {100 * sum(d['ns'] for d in sw) / tot_all:.1f}% |\n")
  for n, v in names.most_common(6):
    f.write(f"| {n} | {v / 1e6:.3f} | {100 * v / tot_all:.2f}% |\n")
json.dump({"c3": tot_b, "_note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the 4 sweep_kernel launches of one "
           "bench step (profiles/r1_bench_launches.md); bytes per step"}, open("profiles/dram_traffic.json", "w"), indent=1)
out = subprocess.run([sys.executable, "scripts/ncu_summary.py", "gpurun_out/prof_r1_final.ncu-rep"], capture_output=True, text=True).stdout
open("profiles/r1_ncu_full_final.txt", "w").write(
    "# ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 24 -c 4 python bench.py --steps 1 --warmup 1\n"
    "# four consecutive launches = one step: fwd1, fwd2, expectation+bwd1, bwd2 in step order (config c3, T=12, K=4)\n" + out)
for c in ("c3", "c1", "c2", "c3l7", "c3q", "c4"):
  try:
    shutil.copy(f"gpurun_out/bench_{c}.json", f"profiles/r1_bench_{c}_n1.json")
  except FileNotFoundError:
    pass
shutil.copy("gpurun_out/bench_ref.json", "profiles/r1_bench_reference_arm.json")
shutil.copy("gpurun_out/bench_ebm.log", "profiles/r1_bench_ebm_2p24.jsonl")
for extra, dst in (("bench_api.log", "r1_bench_api_vqt.txt"), ("sanitize.log", "r1_compute_sanitizer.txt")):
  try:
    shutil.copy(f"gpurun_out/{extra}", f"profiles/{dst}")
  except FileNotFoundError:
    pass
print(open("profiles/r1_bench_launches.md").read())
