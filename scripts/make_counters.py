"""profiles/kernel_counters.json from an `ncu --set full` capture of one bench step.

usage: make_counters.py <rep> <config> <first launch of the step> <launches per step> <bitstrings per launch>
The counters are tagged with the hash of the kernel sources; bench.py only quotes them (roofline.traffic,
roofline.sm_counters) while that hash matches the sources it runs on."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, config, first, per_step, per_launch = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name):
  v = float(r[ix[name]].replace(",", ""))
  u = units[ix[name]]
  return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)


step = rows[2 + first:2 + first + per_step]
dur = [val(r, "gpu__time_duration.sum") for r in step]
dram = sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in step)
top = max(range(per_step), key=lambda i: dur[i])
r = step[top]
out = {
    "source": os.path.relpath(rep, ROOT) + f" (launches {first}..{first + per_step - 1} = one step of {per_launch} bitstrings)",
    "kernel_src_sha": bench.kernel_source_sha(),
    "dram_bytes_per_4096_bitstrings": dram * 4096.0 / per_launch,
    "launch_ms_under_ncu": [round(1e3 * d, 4) for d in dur],
    "dominant_launch": {
        "index_in_step": top,
        "kernel": r[ix["Kernel Name"]],
        "share_of_step": dur[top] / sum(dur),
        "sm_issue_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "lsu_wavefronts_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "smem_wavefronts_pct": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "dram_pct": val(r, "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": val(r, "launch__registers_per_thread"),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
    },
}
path = os.path.join(ROOT, "profiles", "kernel_counters.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[config] = out
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
