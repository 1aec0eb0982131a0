"""Turns the raw outputs of scripts/gpu/final_r2.sh (gpurun_out/<tag>/) into the tracked round-2 summaries
under profiles/.  usage: make_profiles_r2.py <tag>"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
D = os.path.join(ROOT, "gpurun_out", tag)
P = os.path.join(ROOT, "profiles")


def last_json(path):
  return [l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1]


# ---- bench lines
have = lambda n: os.path.exists(os.path.join(D, n))  # (a reduced session only re-measures what changed)
for c in ("c1", "c2", "c3q", "c3l7"):
  if have(f"bench_{c}.json"):
    open(os.path.join(P, f"r2_bench_{c}_n1.json"), "w").write(last_json(os.path.join(D, f"bench_{c}.json")) + "\n")
if have("bench_ref.json"):
  open(os.path.join(P, "r2_bench_reference_arm.json"), "w").write(last_json(os.path.join(D, "bench_ref.json")) + "\n")
open(os.path.join(P, "r2_bench_c3_n1.json"), "w").write(last_json(os.path.join(D, "bench_c3.json")) + "\n")
if have("bench_ebm.txt"):
  with open(os.path.join(P, "r2_bench_ebm_2p24.jsonl"), "w") as f:
    f.write("".join(l for l in open(os.path.join(D, "bench_ebm.txt")) if l.startswith("{")))
if have("bench_api_500.txt"):
  with open(os.path.join(P, "r2_bench_api_vqt.txt"), "w") as f:
    for n in ("bench_api_500.txt", "bench_api_100k.txt"):
      f.write("".join(l for l in open(os.path.join(D, n)) if l.startswith("VQT")))
if have("bench_rows.json"):
  shutil.copy(os.path.join(D, "bench_rows.json"), os.path.join(P, "r2_bench_symbol_rows.json"))

# ---- launch list of one bench step
rows = list(csv.reader(l for l in open(os.path.join(D, "launches_c3.csv")) if l.startswith('"')))
ix = {h: i for i, h in enumerate(rows[0])}
launch = collections.OrderedDict()
for r in rows[1:]:
  d = launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
  d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
to_b = lambda v, u: v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
to_ns = lambda v, u: v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
L = list(launch.values())
for d in L:
  d["ns"] = to_ns(*d["gpu__time_duration.sum"])
  d["rd"] = to_b(*d["dram__bytes_read.sum"])
  d["wr"] = to_b(*d["dram__bytes_write.sum"])
sw = [d for d in L if "sweep_kernel" in d["name"]]
# the last FULL-SIZE chunk (the parity block's small launches come after the timed steps)
big = [i for i in range(len(sw) - 3) if sw[i + 2]["grid"].startswith("(65536")
       and sw[i + 1]["grid"].startswith("(65536") and not sw[i]["grid"].startswith("(65536")]
step = sw[big[-1]:big[-1] + 4]
tot_ns = sum(d["ns"] for d in step)
tot_b = sum(d["rd"] + d["wr"] for d in step)
labels = ["fwd sweep 1: basis state, passes, store (dense kernel; one CTA per state, all-zero tiles not stored)",
          "fwd sweep 2 (dense kernel; all-zero tiles not loaded)",
          "observable passes + expectation + bwd sweep 1 (gradient inner products, tasks, descriptor flush)",
          "bwd sweep 2 (no store)"]
line = json.loads(last_json(os.path.join(D, "bench_c3.json")))
ms4096 = line["config"]["ms_per_4096_bitstrings"]
with open(os.path.join(P, "r2_bench_launches.md"), "w") as f:
  f.write("# Round 2 (final): ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (config c3)\n\n")
  f.write("Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
          "-c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n")
  f.write("(per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes).\n\n")
  f.write("A bench step = 32 768 unique 16-qubit bitstrings = 8 chunks of 4096 = 8 x 4 launches of `qhbm::sweep_kernel` "
          "(256 threads; adjoint kernel 64 KiB dynamic + 27 KiB static smem). Last chunk:\n\n")
  f.write("| # | launch | kernel | grid | duration us | share | dram read MB | dram write MB |\n|---|---|---|---|---|---|---|---|\n")
  for i, d in enumerate(step):
    f.write(f"| {i} | {labels[i]} | `{d['name'][:40]}` | {d['grid']} | {d['ns'] / 1e3:.1f} | {100 * d['ns'] / tot_ns:.1f}% | "
            f"{d['rd'] / 1e6:.1f} | {d['wr'] / 1e6:.1f} |\n")
  f.write(f"\nSweep-kernel time per 4096 bitstrings under ncu: {tot_ns / 1e6:.2f} ms (bench.py, CUDA events, not under ncu: "
          f"{ms4096:.2f} ms, `r2_bench_c3_n1.json`). DRAM traffic: {tot_b / 1e9:.2f} GB per 4096 bitstrings = "
          f"{tot_b / 4096 / 2**20:.2f} MiB per bitstring vs 91 MiB algorithmic (SURVEY 8d) => the state is reused on chip; "
          f"at the measured time this is {tot_b / 1e9 / (ms4096 / 1e3):.0f} GB/s = "
          f"{100 * tot_b / 1e9 / (ms4096 / 1e3) / 6551.7:.0f}% of the measured HBM peak.\n\n")
  tot_all = sum(d["ns"] for d in L)
  names = collections.Counter()
  for d in L:
    names[d["name"].split("(")[0][-48:]] += d["ns"]
  f.write("Share of all GPU time in the run by kernel:\n\n| kernel | total ms | share |\n|---|---|---|\n")
  for n, v in names.most_common(8):
    f.write(f"| `{n}` | {v / 1e6:.3f} | {100 * v / tot_all:.2f}% |\n")

# ---- ncu --set full summaries
rep = os.path.join(D, "prof_c3.ncu-rep")
lib = os.path.join(D, "libqhbm_b200.so")
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(os.path.join(P, "r2_ncu_full_final.txt"), "w").write(
    "# ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 96 -c 4 python bench.py --steps 1 --warmup 3\n"
    "# four consecutive launches = one 4096-bitstring chunk of config 3 (T=12, K=4): fwd sweep 1, fwd sweep 2,\n"
    "# observable passes + expectation + bwd sweep 1, bwd sweep 2 (round-2 final build)\n" + out)
for k in (2, 3):
  env = dict(os.environ, MINFRAC="0.004")
  a = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), rep, str(k), lib,
                      "sweep_kernelILi4ELb1ELb0ELb0E", "2"], capture_output=True, text=True, env=env).stdout
  b = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), rep, str(k), lib,
                      "sweep_kernelILi4ELb1ELb0ELb0E", "1"], capture_output=True, text=True, env=env).stdout
  open(os.path.join(P, f"r2_regions_launch{k}.txt"), "w").write(
      f"# scripts/ncu_regions.py <rep> {k} <lib> sweep_kernelILi4ELb1ELb0ELb0E: warp instructions / stall samples of launch {k}\n"
      "# by source region (outer frame > inner frame), through nvdisasm inline chains\n## depth 1\n" + b + "\n## depth 2\n" + a)
subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_counters.py"), rep, "c3", "0", "4", "4096"], check=True,
               stdout=subprocess.DEVNULL)
shutil.copy(os.path.join(D, "pytest_gpu.log"), os.path.join(P, "r2_pytest_gpu.log"))
print(open(os.path.join(P, "r2_bench_launches.md")).read())
