#!/usr/bin/env python
"""Headline benchmark: QHBM expectation + adjoint gradient over unique bitstrings.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the TFQ path

Metric (BASELINE.json): unique bitstrings/s for one "step" = forward expectation +
adjoint gradient of every unique bitstring, count-weighted (what one call of
`QHBM.expectation` + `tape.gradient` costs in qhbmlib), at the 16-qubit / 4096-unique
configuration.  One JSON line is printed by rank 0.  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "qhbm-library_b200")):
  if _p not in sys.path:
    sys.path.insert(0, _p)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (n_qubits, layers, unique bitstrings per GPU, hamiltonian, with adjoint gradient)
    "c1": dict(n=4, layers=2, unique=16, ham="tfim", grad=True,
               label="4-qubit 1D TFIM, HEA L=2, all 16 basis states (config 1 shape)"),
    "c2": dict(n=12, layers=2, unique=4096, ham="tfim", grad=True,
               label="12-qubit TFIM, HEA L=2, 4096 unique bitstrings, fwd+adjoint (config 2 shape)"),
    "c3": dict(n=16, layers=2, unique=4096, ham="xxz", grad=True,
               label="16-qubit Heisenberg XXZ ring, HEA L=2, 4096 unique bitstrings, fwd+adjoint (config 3)"),
    "c3l7": dict(n=16, layers=7, unique=4096, ham="xxz", grad=True,
                 label="16-qubit XXZ ring, HEA L=7, 4096 unique bitstrings, fwd+adjoint"),
    "c3q": dict(n=16, layers=2, unique=4096, ham="kobe2", grad=True,
                label="16-qubit QMHL term: data HEA L=2 + model HEA L=2 inverse, 136 KOBE-2 Z-shards, 4096 unique "
                      "bitstrings, fwd+adjoint"),
    "c4": dict(n=20, layers=2, unique=8192, ham="tfim", grad=False,
               label="20-qubit TFIM ring, HEA L=2, 8192 unique bitstrings per GPU, forward (config 4 shard)"),
}


def synth_workload(cfg, rank):
  """Synthetic inputs of SURVEY.md section 8(d): HEA ansatz, phi ~ U(-1,1) seed 11, distinct
  bitstrings (seed 3 + rank), counts = 1 + multinomial(1e6, softmax(-KOBE-2 energy)).
  Returns the C-ABI tables (gate table, Pauli term table) built by the product's own builders."""
  from qhbmlib import architectures as arch
  from qhbmlib import circuits as cq
  from qhbmlib import models
  n, layers, u = cfg["n"], cfg["layers"], cfg["unique"]
  qubits = cq.GridQubit.rect(1, n)
  circuit = arch.get_hardware_efficient_model_unitary(qubits, layers, "q")
  names = sorted(cq.circuit_symbols(circuit))
  if cfg["ham"] == "kobe2":
    # QMHL: <K_model> on data states = data circuit followed by the inverse model circuit, measured on
    # the model energy's Z-string shards (qnn.py:69-72, hamiltonian.py:48-51)
    model = arch.get_hardware_efficient_model_unitary(qubits, layers, "m")
    mnames = sorted(cq.circuit_symbols(model))
    circuit = circuit + model**-1
    names = names + mnames
    shards = models.KOBE(list(range(n)), 2).operator_shards(qubits)
    terms, offs = cq.convert_to_tensor(shards).tables(qubits)
  else:
    ham = arch.xxz_ring(qubits) if cfg["ham"] == "xxz" else arch.tfim_ring(qubits)
    terms, offs = cq.convert_to_tensor([ham]).tables(qubits)
  gates = cq.gate_table(circuit, qubits, names)
  phi = np.random.default_rng(11).uniform(-1, 1, len(names)).astype(np.float32)
  rng = np.random.default_rng(3 + rank)
  u = min(u, 1 << n)
  basis = rng.choice(1 << n, size=u, replace=False).astype(np.int64)
  masks = np.array(models.Parity(list(range(n)), 2).masks(), dtype=np.int64)
  theta = rng.normal(0, 0.1, len(masks))
  x = basis[:, None] & masks[None, :]
  par = np.zeros_like(x)
  for b in range(n):
    par ^= (x >> b) & 1
  energy = ((1 - 2 * par) * theta).sum(1)
  p = np.exp(-energy - (-energy).max())
  counts = (1 + rng.multinomial(1_000_000, p / p.sum())).astype(np.int32)
  return gates, names, phi, (terms, offs), basis, counts


def oracle_ops(terms, offs, n):
  """C-ABI Pauli term table -> the oracle's [(coeff, {qubit: pauli})] lists (CPU baseline leg)."""
  ops = []
  for j in range(len(offs) - 1):
    op = []
    for t in terms[offs[j]:offs[j + 1]]:
      paulis = {}
      for q in range(n):
        bit = 1 << (n - 1 - q)
        xb, zb = int(t["xmask"]) & bit, int(t["zmask"]) & bit
        if xb and zb:
          paulis[q] = "Y"
        elif xb:
          paulis[q] = "X"
        elif zb:
          paulis[q] = "Z"
      op.append((float(t["coeff"]), paulis))
    ops.append(op)
  return ops


def algorithmic_bytes_per_bitstring(cfg):
  """F = L(n-1) fused <=2-qubit blocks; fwd (2F+1) S, fwd+adjoint (6F+2) S."""
  f = cfg["layers"] * (cfg["n"] - 1) * (2 if cfg["ham"] == "kobe2" else 1)
  s = 8 * (1 << cfg["n"])
  return ((6 * f + 2) if cfg["grad"] else (2 * f + 1)) * s


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons during the timed region."""
  FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
            "clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index, self.rows, self.proc, self.thread = index, [], None, None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
           "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
      self.proc = None
      return
    self.thread = threading.Thread(target=self._read, daemon=True)
    self.thread.start()

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, smax, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      try:
        sm.append(float(r[0]))
        smax.append(float(r[1]))
        for k, nm in enumerate(names):
          if r[3 + k].lower().startswith("active"):
            reasons.add(nm)
      except Exception:
        continue
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
            "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
  return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_rate(cfg, sample, threads=0, with_forward_op=True):
  """Bitstrings/s of the CPU restatement of TFQ's ops on `sample` bitstrings of this workload."""
  from oracle import tfq_cpu
  from oracle import qhbm_oracle as orc
  gates, names, phi, (terms, offs), basis, counts = synth_workload(cfg, 0)
  prob = tfq_cpu.Problem(gates.astype(orc.GATE_DTYPE), cfg["n"], phi, oracle_ops(terms, offs, cfg["n"]), "tfq_fd")
  b = basis[:sample]
  n_ops = len(offs) - 1
  w = np.random.default_rng(17).normal(0, 0.1, n_ops).astype(np.float32) if n_ops > 1 else np.ones(1, np.float32)
  dg = ((counts[:sample] / counts.sum()).astype(np.float32)[:, None] * w[None, :]).astype(np.float32)
  t0 = time.perf_counter()
  if cfg["grad"]:
    if with_forward_op:
      prob.expectation(b, threads)  # TfqSimulateExpectation (forward op of the training step)
    prob.adjoint(b, dg, threads)    # TfqAdjointGradient (re-simulates the forward pass itself)
  else:
    prob.expectation(b, threads)
  dt = time.perf_counter() - t0
  return len(b) / dt, dt, (threads or tfq_cpu.max_threads())


def run_reference(args, cfg):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  # every host thread the box offers (torchrun pins OMP_NUM_THREADS=1, so ask explicitly)
  threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
  rate0, _, _ = cpu_reference_rate(cfg, min(2 * threads, cfg["unique"]), threads)
  sample = int(max(8, min(cfg["unique"], rate0 * 4.0)))  # ~4 s of CPU work per step
  for _ in range(args.warmup if args.warmup < 2 else 1):
    cpu_reference_rate(cfg, sample, threads)
  times = []
  for _ in range(args.steps):
    _, dt, _ = cpu_reference_rate(cfg, sample, threads)
    times.append(dt)
  ms = 1e3 * float(np.mean(times))
  value = sample / (ms / 1e3)
  desc = (f"{sample} of {cfg['unique']} unique bitstrings per step; forward op + adjoint op "
          f"(TFQ's adjoint re-simulates the forward); CPU restatement of TFQ 0.6.1's algorithm, not TFQ itself")
  print(json.dumps({
      "impl": "reference", "metric": "unique bitstrings/s (expectation + adjoint gradient)", "value": value,
      "unit": "bitstrings/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64",
      "data": "synthetic", "config": {"workload": cfg["label"], "sample_per_step": sample},
      "cpu_baseline": {"value": value, "unit": "bitstrings/s", "cores": threads, "kind": "port", "sample": desc},
      "e2e": {"value": value, "unit": "bitstrings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }), flush=True)


def run_gpu(args, cfg):
  import torch
  import torch.distributed as dist
  from qhbmlib import engine

  if not torch.cuda.is_available():
    sys.exit("bench.py: no CUDA device. The qhbm_b200 engine has no CPU fallback; "
             "`--impl reference` times the CPU restatement of the reference path.")
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

  gates, names, phi, (terms, offs), basis, counts = synth_workload(cfg, rank)
  n, u, grad = cfg["n"], len(basis), cfg["grad"]
  plan = engine.ExpectationPlan(gates, n, len(names), terms, offs, grad, args.tile_qubits, args.reg_qubits)
  n_ops, n_sym = len(offs) - 1, len(names)
  op_weights = np.random.default_rng(17).normal(0, 0.1, n_ops).astype(np.float32) if n_ops > 1 else np.ones(1, np.float32)

  total_counts = torch.tensor([float(counts.sum())], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(total_counts)
  d_basis = torch.tensor(basis, device=dev)
  d_counts = torch.tensor(counts, device=dev)
  d_phi = torch.tensor(phi, device=dev)
  # upstream gradient of every expectation: count weight x (for shards) the energy parameter theta_j
  dgrad_np = ((counts / float(total_counts.item())).astype(np.float32)[:, None] * op_weights[None, :]).astype(np.float32)
  d_dgrad = torch.tensor(dgrad_np, device=dev)
  packed = torch.zeros(n_ops + 1 + n_sym, dtype=torch.float64, device=dev)

  def step():
    """Resident inputs -> count-weighted expectation [O] and gradient [P] (allreduced)."""
    if grad:
      e, g = plan.forward_adjoint(d_basis, d_phi, d_dgrad, grad_mode=args.grad_mode)
      packed[n_ops + 1:] = g.double()
    else:
      e = plan.forward(d_basis, d_phi)
    packed[:n_ops + 1] = engine.weighted_sum(d_counts, e)
    if world > 1:
      dist.all_reduce(packed)  # the single collective of the path: [sum c<H>, sum c, grad]
    return packed[:n_ops] / packed[n_ops]

  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(max(args.warmup, 3)):
    step()
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  evs = []
  barrier()
  wall0 = time.perf_counter()
  for _ in range(args.steps):
    flush.zero_()  # L2 flush between timed iterations (untimed)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    step()
    b.record()
    evs.append((a, b))
  barrier()
  wall = time.perf_counter() - wall0
  clocks = sampler.stop() if rank == 0 else None
  dev_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
  ms_per_step = float(dev_ms.item()) / args.steps
  value = world * u / (ms_per_step / 1e3)

  # ---- end to end through the host-buffer C-ABI call (what a TF custom-op shim would bind)
  h_basis = torch.tensor(basis).pin_memory().numpy().view(np.uint64)
  h_phi = torch.tensor(phi).pin_memory().numpy()
  h_dgrad = torch.tensor(dgrad_np).pin_memory().numpy()
  h_packed = torch.zeros(n_ops + 1 + n_sym, dtype=torch.float64).pin_memory()

  def step_e2e():
    e, g = plan.run_host(h_basis, h_phi, h_dgrad if grad else None, grad_mode=args.grad_mode)
    h_packed[:n_ops] = torch.from_numpy((counts[:, None] * e.astype(np.float64)).sum(0))
    h_packed[n_ops] = float(counts.sum())
    if grad:
      h_packed[n_ops + 1:] = torch.from_numpy(g.astype(np.float64))
    if world > 1:
      t = h_packed.to(dev, non_blocking=True)
      dist.all_reduce(t)
      h_packed.copy_(t)
    return float(h_packed[0] / h_packed[n_ops])

  for _ in range(2):
    step_e2e()
  barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    loss = step_e2e()
  barrier()
  e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
  e2e_value = world * u * args.steps / float(e2e_t.item())
  h2d = u * 8 + n_sym * 4 + (u * n_ops * 4 if grad else 0)
  d2h = u * n_ops * 4 + (n_sym * 4 if grad else 0)

  if rank == 0:
    peak, peak_src = measured_peak()
    bytes_per = algorithmic_bytes_per_bitstring(cfg)
    chunks = -(-u // plan.info["chunk"])
    sweep_launches = chunks * plan.info["launches"]
    achieved = u * bytes_per / (ms_per_step / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
      with open(tpath) as f:
        traffic = json.load(f).get(args.config)
    out = {
        "metric": "unique bitstrings/s (expectation + adjoint gradient)" if grad else
                  "unique bitstrings/s (expectation)",
        "value": value, "unit": "bitstrings/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex64", "data": "synthetic",
        "config": {"workload": cfg["label"], "unique_per_gpu": u, "symbols": n_sym, "grad_mode": args.grad_mode,
                   "tile_qubits": plan.info["tile_qubits"], "reg_qubits": plan.info["reg_qubits"],
                   "chunk": plan.info["chunk"], "sweeps_fwd": plan.info["sweeps_fwd"],
                   "sweeps_bwd": plan.info["sweeps_bwd"], "parallelism": f"unique bitstrings sharded x{world}",
                   "l2": "256 MiB buffer written between timed steps (untimed); per-step CUDA events summed",
                   "wall_s_timed_region_incl_flush": wall},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_bitstring": bytes_per, "sweep_kernel_launches_per_step": sweep_launches,
                     "note": "HBM-equivalent: the state is smem/L2 resident, so frac > 1 means on-chip reuse, "
                             "not skipped work (see profiles/ for dram bytes and smem throughput)"},
        "e2e": {"value": e2e_value, "unit": "bitstrings/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "last_loss": loss},
        "gpu_launches": sweep_launches + 3,  # + prep (incl. accumulator clears), finalize, weighted sum
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
      nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
      rate0, _, threads = cpu_reference_rate(cfg, min(2 * nthreads, u), nthreads)
      sample = int(max(8, min(u, rate0 * 10.0)))
      rate, dt, threads = cpu_reference_rate(cfg, sample, nthreads)
      out["cpu_baseline"] = {
          "value": rate, "unit": "bitstrings/s", "cores": threads, "kind": "port",
          "sample": f"{sample} of {u} unique bitstrings, {dt:.1f} s; forward op + adjoint op; CPU restatement "
                    "of TFQ 0.6.1's algorithm (oracle/tfq_cpu.c), not TFQ itself"}
    print(json.dumps(out), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
  ap.add_argument("--grad-mode", default="tfq_fd", choices=["exact", "tfq_fd", "tfq_fd_f32"])
  ap.add_argument("--tile-qubits", type=int, default=0)
  ap.add_argument("--reg-qubits", type=int, default=0)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  args = ap.parse_args()
  cfg = CONFIGS[args.config]
  if args.impl == "reference":
    run_reference(args, cfg)
  else:
    run_gpu(args, cfg)


if __name__ == "__main__":
  main()
