"""CPU-side checks of the bench.py contract: the reference arm prints ONE JSON line with the agreed
keys, and the CUDA arm refuses to run without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_contract_keys():
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.strip()]
  assert len(lines) == 1, out.stdout
  d = json.loads(lines[0])
  assert d["impl"] == "reference"
  for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
    assert key in d, key
  assert d["value"] > 0 and d["unit"] == "bitstrings/s" and d["higher_is_better"] is True
  assert d["vs_baseline"] is None and "workload" in d["config"]
  assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
  assert d["cpu_baseline"]["value"] == d["value"]
  assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_cuda_arm_refuses_to_run_without_a_gpu():
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert out.returncode != 0
  assert "no CPU fallback" in (out.stderr + out.stdout)
  assert not any(l.startswith("{") for l in out.stdout.splitlines())


def test_committed_ncu_counters_belong_to_the_committed_sweep_kernel_sources():
  """bench.py only quotes `roofline.traffic` / `roofline.sm_counters` from profiles/kernel_counters.json while the
  hash stored there equals the hash of the sweep-kernel sources: an edit to those sources without a new ncu
  capture shows up here instead of silently dropping the counters from the bench line."""
  sys.path.insert(0, ROOT)
  import bench
  c = bench.ncu_counters("c3")
  assert c is not None and c["current"], (c and c.get("kernel_src_sha"), bench.kernel_source_sha())
  assert c["dram_bytes_per_4096_bitstrings"] > 0 and 0 < c["dominant_launch"]["share_of_step"] < 1
