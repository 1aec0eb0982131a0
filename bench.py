#!/usr/bin/env python
"""Headline benchmark: QHBM expectation + adjoint gradient over unique bitstrings.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the TFQ path
  python bench.py --config c4 | c5 | ...                   # the other BASELINE.json configurations

Metric (BASELINE.json): unique bitstrings/s for one "step" = forward expectation + adjoint gradient of
every unique bitstring, count-weighted (what one call of `QHBM.expectation` + `tape.gradient` costs in
qhbmlib), on the 16-qubit XXZ / HEA L=2 configuration.  The workload is FIXED (strong scaling): 32 768
distinct bitstrings = eight config-3 batches of 4096, split contiguously over the N ranks, so that at
N = 8 every GPU holds exactly BASELINE config 3 (4096 unique on one B200) and at N = 1 the engine runs
the eight batches back to back (its workspace holds 4096 states per chunk).  One JSON line is printed by
rank 0; it carries a `parity` block: the engine's results on >= 8 bitstrings per rank against committed
oracle values for exactly these inputs (tests/golden/bench_parity_*.npz).  See DESIGN.md "Measurement".
"""
import argparse
import hashlib
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
    # name: n_qubits, layers, TOTAL unique bitstrings (split over the ranks), hamiltonian, adjoint gradient?
    "c1": dict(n=4, layers=2, total=16, ham="tfim", grad=True,
               label="4-qubit 1D TFIM, HEA L=2, all 16 basis states (config 1 shape)"),
    "c2": dict(n=12, layers=2, total=4096, ham="tfim", grad=True,
               label="12-qubit TFIM, HEA L=2, 4096 unique bitstrings, fwd+adjoint (config 2 shape)"),
    "c3": dict(n=16, layers=2, total=32768, ham="xxz", grad=True,
               label="16-qubit Heisenberg XXZ ring, HEA L=2, fwd+adjoint: 32768 unique bitstrings = 8 batches of "
                     "config 3 (4096 unique per batch)"),
    "c3l7": dict(n=16, layers=7, total=32768, ham="xxz", grad=True,
                 label="16-qubit XXZ ring, HEA L=7, 32768 unique bitstrings, fwd+adjoint"),
    "c3q": dict(n=16, layers=2, total=32768, ham="kobe2", grad=True,
                label="16-qubit QMHL term: data HEA L=2 + model HEA L=2 inverse, 136 KOBE-2 Z-shards, 32768 unique "
                      "bitstrings, fwd+adjoint"),
    "c4": dict(n=20, layers=2, total=65536, ham="tfim", grad=False,
               label="20-qubit TFIM ring, HEA L=2, 65536 unique bitstrings, forward expectation (config 4)"),
    "c5": dict(n=24, ebm=True, samples=1_000_000,
               label="AnalyticEnergyInference sweep over all 2^24 bitstrings, MLP energy 24-64-64-1 (tanh), "
                     "logsumexp + entropy + 1e6 categorical samples (config 5)"),
}
PARITY_STATES = 64  # 8 per rank at 8 ranks


# ----------------------------------------------------------------------------------------------
# Synthetic workloads.  Pure numpy + the product's own circuit builders: no GPU.
# ----------------------------------------------------------------------------------------------
def synth_workload(cfg):
  """GLOBAL workload of a state-vector config: HEA ansatz, phi ~ U(-1,1) seed 11, `total` distinct
  bitstrings (seed 3), counts = 1 + multinomial(1e6, softmax(-KOBE-2 energy)).  Returns the C-ABI tables
  (gate table, Pauli term table) built by the product's own builders."""
  from qhbmlib import architectures as arch
  from qhbmlib import circuits as cq
  from qhbmlib import models
  n, layers = cfg["n"], cfg["layers"]
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
  rng = np.random.default_rng(3)
  u = min(cfg["total"], 1 << n)
  basis = rng.choice(1 << n, size=u, replace=False).astype(np.int64)
  masks = np.array(models.Parity(list(range(n)), 2).masks(), dtype=np.int64)
  theta = rng.normal(0, 0.1, len(masks))
  energy = np.zeros(u)
  for lo in range(0, u, 8192):  # chunked: [rows, masks] parity matrix
    x = basis[lo:lo + 8192, None] & masks[None, :]
    par = np.zeros_like(x)
    for b in range(n):
      par ^= (x >> b) & 1
    energy[lo:lo + 8192] = ((1 - 2 * par) * theta).sum(1)
  p = np.exp(-energy - (-energy).max())
  counts = (1 + rng.multinomial(1_000_000, p / p.sum())).astype(np.int32)
  n_ops = len(offs) - 1
  op_weights = (np.random.default_rng(17).normal(0, 0.1, n_ops).astype(np.float32) if n_ops > 1
                else np.ones(1, np.float32))
  return dict(gates=gates, names=names, phi=phi, terms=terms, offs=offs, basis=basis, counts=counts,
              op_weights=op_weights, n_ops=n_ops)


def shard_range(n, rank, world):
  base, rem = divmod(int(n), int(world))
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def parity_positions(total):
  """Positions (into the global bitstring list) of the parity-checked states: 8 at the head of each
  eighth of the list, so every rank of a 1/2/4/8-way split owns at least 8 of them."""
  if total <= PARITY_STATES:
    return np.arange(total)
  eighth = total // 8
  return np.array([r * eighth + j for r in range(8) for j in range(PARITY_STATES // 8)])


def coeff_l1(terms, offs, op_weights):
  """sum_j |w_j| sum_terms |coeff|: the scale of the absolute floor (cancelling Pauli sums)."""
  tot = 0.0
  for j in range(len(offs) - 1):
    tot += abs(float(op_weights[j])) * float(np.abs(terms["coeff"][offs[j]:offs[j + 1]]).sum())
  return tot


def mlp_weights(n):
  """Config 5 energy: Dense(64,tanh) -> Dense(64,tanh) -> Dense(1) on the raw bits, Glorot-uniform seed 4,
  zero bias (family of tests/inference/ebm_utils_test.py:33-47)."""
  rng = np.random.default_rng(4)
  widths = [n, 64, 64, 1]
  ws = []
  for l in range(3):
    lim = np.sqrt(6.0 / (widths[l] + widths[l + 1]))
    ws.append(rng.uniform(-lim, lim, (widths[l], widths[l + 1])).astype(np.float32))
  return widths, ws


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


def kernel_source_sha():
  """Hash of the kernel sources: ncu-derived counters in profiles/ are only quoted for the build they
  were captured on."""
  h = hashlib.sha256()
  csrc = os.path.join(ROOT, "qhbm-library_b200", "csrc")
  # the sources that compile into the state-vector kernels and their plans (the counters are of sweep_kernel);
  # the EBM-side, measurement, collective and coefficient-preparation sources do not change them
  other = {"ebm.cu", "measure.cu", "comm.cu", "philox.h", "prep_kernels.cuh"}
  for f in sorted(os.listdir(csrc)):
    if not os.path.isfile(os.path.join(csrc, f)) or f.startswith(".") or f in other:
      continue
    with open(os.path.join(csrc, f), "rb") as fh:
      h.update(fh.read())
  return h.hexdigest()[:16]


def ncu_counters(config):
  """Per-config counters from the committed ncu capture (profiles/kernel_counters.json): DRAM bytes per
  step and SM utilisation of the dominant launch.  Returned with `current` = whether they were captured
  on exactly this kernel source."""
  path = os.path.join(ROOT, "profiles", "kernel_counters.json")
  if not os.path.exists(path):
    return None
  with open(path) as f:
    data = json.load(f)
  c = data.get(config)
  if c is None:
    return None
  c = dict(c)
  c["current"] = c.get("kernel_src_sha") == kernel_source_sha()
  return c


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


def host_threads():
  return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ----------------------------------------------------------------------------------------------
# CPU arm: the restatement of the two TFQ ops (oracle/tfq_cpu.c), all host threads.
# ----------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, wl, sample, threads=0, with_forward_op=True):
  """Bitstrings/s of the CPU restatement of TFQ's ops on `sample` bitstrings of this workload."""
  from oracle import tfq_cpu
  from oracle import qhbm_oracle as orc
  prob = tfq_cpu.Problem(wl["gates"].astype(orc.GATE_DTYPE), cfg["n"], wl["phi"],
                         oracle_ops(wl["terms"], wl["offs"], cfg["n"]), "tfq_fd")
  b = wl["basis"][:sample]
  dg = ((wl["counts"][:sample] / wl["counts"].sum()).astype(np.float32)[:, None] *
        wl["op_weights"][None, :]).astype(np.float32)
  t0 = time.perf_counter()
  if cfg["grad"]:
    if with_forward_op:
      prob.expectation(b, threads)  # TfqSimulateExpectation (forward op of the training step)
    prob.adjoint(b, dg, threads)    # TfqAdjointGradient (re-simulates the forward pass itself)
  else:
    prob.expectation(b, threads)
  dt = time.perf_counter() - t0
  return len(b) / dt, dt, (threads or tfq_cpu.max_threads())


def cpu_ebm_rate(cfg, rows):
  """Rows/s of a numpy float32 restatement of AnalyticEnergyInference._ready_inference +
  log_partition + entropy (ebm.py:467-485) on `rows` rows of the 2^n enumeration."""
  n = cfg["n"]
  widths, ws = mlp_weights(n)
  t0 = time.perf_counter()
  m, s, t = -np.inf, 0.0, 0.0
  for lo in range(0, rows, 1 << 16):
    idx = np.arange(lo, min(rows, lo + (1 << 16)), dtype=np.int64)
    bits = ((idx[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1).astype(np.float32)
    h = np.tanh(bits @ ws[0])
    h = np.tanh(h @ ws[1])
    logits = -(h @ ws[2])[:, 0]
    m2 = max(m, float(logits.max()))
    e = np.exp(logits.astype(np.float64) - m2)
    s = s * np.exp(m - m2) + e.sum()
    t = t * np.exp(m - m2) + (e * logits).sum()
    m = m2
  dt = time.perf_counter() - t0
  return rows / dt, dt


def run_reference(args, cfg):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  threads = host_threads()  # torchrun pins OMP_NUM_THREADS=1, so ask explicitly
  if cfg.get("ebm"):
    rows = 1 << 20
    cpu_ebm_rate(cfg, 1 << 16)
    times = [cpu_ebm_rate(cfg, rows)[1] for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(times))
    value = rows / (ms / 1e3)
    desc = (f"{rows} of {1 << cfg['n']} rows per step; numpy float32 restatement of "
            "AnalyticEnergyInference._ready_inference + logsumexp + entropy (BLAS threads as numpy uses them)")
    metric, unit = "rows/s (2^n energy sweep + logsumexp + entropy)", "rows/s"
  else:
    wl = synth_workload(cfg)
    rate0, _, _ = cpu_reference_rate(cfg, wl, min(2 * threads, cfg["total"]), threads)
    sample = int(max(8, min(cfg["total"], rate0 * 4.0)))  # ~4 s of CPU work per step
    for _ in range(args.warmup if args.warmup < 2 else 1):
      cpu_reference_rate(cfg, wl, sample, threads)
    times = []
    for _ in range(args.steps):
      _, dt, _ = cpu_reference_rate(cfg, wl, sample, threads)
      times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    desc = (f"{sample} of {cfg['total']} unique bitstrings per step; forward op + adjoint op "
            f"(TFQ's adjoint re-simulates the forward); CPU restatement of TFQ 0.6.1's algorithm "
            f"(oracle/tfq_cpu.c, split re/im AVX loops, OpenMP over bitstrings), not TFQ itself")
    metric = ("unique bitstrings/s (expectation + adjoint gradient)" if cfg["grad"]
              else "unique bitstrings/s (expectation)")
    unit = "bitstrings/s"
  print(json.dumps({
      "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
      "scaling": "strong", "vs_baseline": None, "dtype": "complex64" if not cfg.get("ebm") else "f32",
      "data": "synthetic", "config": {"workload": cfg["label"]},
      "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": desc},
      "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }), flush=True)


# ----------------------------------------------------------------------------------------------
# Parity block: engine vs committed oracle values for the bench's own inputs.
# ----------------------------------------------------------------------------------------------
def load_parity_fixture(config):
  path = os.path.join(ROOT, "tests", "golden", f"bench_parity_{config}.npz")
  if not os.path.exists(path):
    return None
  return dict(np.load(path))


def expectation_parity(plan, wl, cfg, fixture, lo, hi, dev, total_counts):
  """Max errors of expectations and of the count-weighted reduced gradient on this rank's share of the
  fixture's bitstrings, for both gradient modes.  Tolerance: 1e-5 relative, with the absolute floor
  1e-6 * sum|coeff| (cancelling Pauli sums) scaled by the weights for the reduced gradient."""
  import torch
  pos = fixture["positions"]
  mine = np.nonzero((pos >= lo) & (pos < hi))[0]
  if len(mine) == 0:
    return None
  assert np.array_equal(fixture["basis"][mine], wl["basis"][pos[mine]]), "fixture does not match the workload"
  b = torch.tensor(wl["basis"][pos[mine]], device=dev)
  phi = torch.tensor(wl["phi"], device=dev)
  w = wl["op_weights"]
  l1 = coeff_l1(wl["terms"], wl["offs"], np.ones_like(w))
  l1w = coeff_l1(wl["terms"], wl["offs"], w)
  cw = (wl["counts"][pos[mine]] / total_counts).astype(np.float64)
  out = {"states_checked": int(len(mine)), "rel_tol": 1e-5, "abs_floor": "1e-6 * sum|coeff|"}
  e_ref = fixture["expectations"][mine]
  if cfg["grad"]:
    dg = torch.tensor(np.tile(w[None, :], (len(mine), 1)).astype(np.float32), device=dev)
    for mode in ("exact", "tfq_fd"):
      e, g = plan.forward_adjoint(b, phi, dg, per_state=True, grad_mode=mode)
      g = g.double().cpu().numpy()
      g_ref = fixture[f"grad_{mode}"][mine]
      red, red_ref = (cw[:, None] * g).sum(0), (cw[:, None] * g_ref).sum(0)
      floor = 1e-6 * l1w * cw.sum()
      err = np.abs(red - red_ref)
      out[f"grad_{mode}_max_abs_err"] = float(err.max())
      out[f"grad_{mode}_max_rel_err"] = float((err / np.maximum(np.abs(red_ref), floor / 1e-5)).max())
      out[f"grad_{mode}_norm_rel_err"] = float(err.max() / np.abs(red_ref).max())
      perr = np.abs(g - g_ref)
      out[f"grad_{mode}_per_state_norm_rel_err"] = float((perr.max(1) / np.abs(g_ref).max(1)).max())
  else:
    e = plan.forward(b, phi)
  e = e.double().cpu().numpy()
  err = np.abs(e - e_ref)
  floor = 1e-6 * l1
  out["exp_max_abs_err"] = float(err.max())
  out["exp_max_rel_err"] = float((err / np.maximum(np.abs(e_ref), floor / 1e-5)).max())
  out["exp_abs_floor"] = floor
  worst = max([out["exp_max_rel_err"]] + [out[k] for k in out if k.startswith("grad_") and k.endswith("max_rel_err")])
  out["max_rel_err"] = float(worst)
  out["pass"] = bool(worst <= 1e-5)
  return out


# ----------------------------------------------------------------------------------------------
# State-vector arm (configs c1..c4)
# ----------------------------------------------------------------------------------------------
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
  if cfg.get("ebm"):
    return run_gpu_ebm(args, cfg, world, rank, local_rank, dev)

  wl = synth_workload(cfg)
  total = len(wl["basis"])
  lo, hi = shard_range(total, rank, world)
  basis, counts = wl["basis"][lo:hi], wl["counts"][lo:hi]
  n, u, grad = cfg["n"], hi - lo, cfg["grad"]
  n_ops, n_sym = wl["n_ops"], len(wl["names"])
  plan = engine.ExpectationPlan(wl["gates"], n, n_sym, wl["terms"], wl["offs"], grad, args.tile_qubits,
                                args.reg_qubits)
  total_counts = float(wl["counts"].sum())
  d_basis = torch.tensor(basis, device=dev)
  d_counts = torch.tensor(counts, device=dev)
  d_phi = torch.tensor(wl["phi"], device=dev)
  # upstream gradient of every expectation: count weight x (for shards) the energy parameter theta_j
  dgrad_np = ((counts / total_counts).astype(np.float32)[:, None] * wl["op_weights"][None, :]).astype(np.float32)
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

  warmup = max(args.warmup, 3)
  for _ in range(warmup):
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
  value = total / (ms_per_step / 1e3)
  device_loss = float((packed[:n_ops] / packed[n_ops])[0].item())

  # ---- end to end through the host-buffer C-ABI call (what a TF custom-op shim would bind)
  h_basis = torch.tensor(basis).pin_memory().numpy().view(np.uint64)
  h_phi = torch.tensor(wl["phi"]).pin_memory().numpy()
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
  e2e_value = total * args.steps / float(e2e_t.item())
  h2d = u * 8 + n_sym * 4 + (u * n_ops * 4 if grad else 0)
  d2h = u * n_ops * 4 + (n_sym * 4 if grad else 0)

  # ---- parity (untimed): this rank's share of the committed oracle values
  fixture = load_parity_fixture(args.config)
  parity = None
  if fixture is not None:
    parity = expectation_parity(plan, wl, cfg, fixture, lo, hi, dev, total_counts)
    if "loss" in fixture:
      pass
  if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, parity)
  else:
    gathered = [parity]

  if rank == 0:
    peak, peak_src = measured_peak()
    bytes_per = algorithmic_bytes_per_bitstring(cfg)
    chunks = -(-u // plan.info["chunk"])
    sweep_launches = chunks * plan.info["launches"]
    achieved = total * bytes_per / (ms_per_step / 1e3) / 1e9 / world  # per GPU
    counters = ncu_counters(args.config)
    dom = (counters or {}).get("dominant_launch", {}) if (counters or {}).get("current") else {}
    par = merge_parity([g for g in gathered if g is not None])
    if par is not None and fixture is not None and "loss" in fixture:
      # the whole-job count-weighted mean against the oracle's value over ALL bitstrings (float64)
      par["loss_device"], par["loss_oracle_all_bitstrings"] = device_loss, float(fixture["loss"])
      par["loss_rel_err"] = abs(device_loss - float(fixture["loss"])) / max(abs(float(fixture["loss"])), 1e-30)
    out = {
        "metric": "unique bitstrings/s (expectation + adjoint gradient)" if grad else
                  "unique bitstrings/s (expectation)",
        "value": value, "unit": "bitstrings/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "complex64", "data": "synthetic",
        "config": {"workload": cfg["label"], "total_unique": total, "unique_per_gpu": u, "symbols": n_sym,
                   "ms_per_4096_bitstrings": ms_per_step * 4096.0 / u,
                   "grad_mode": args.grad_mode, "tile_qubits": plan.info["tile_qubits"],
                   "reg_qubits": plan.info["reg_qubits"], "chunk": plan.info["chunk"],
                   "sweeps_fwd": plan.info["sweeps_fwd"], "sweeps_bwd": plan.info["sweeps_bwd"],
                   "parallelism": f"unique bitstrings sharded x{world}, one all-reduce of [sum c<H>, sum c, grad]",
                   "l2": "256 MiB buffer written between timed steps (untimed); per-step CUDA events summed",
                   "wall_s_timed_region_incl_flush": wall},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (counters or {}).get("dram_bytes_per_4096_bitstrings") if (counters or {}).get("current") else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_bitstring": bytes_per,
                     "sweep_kernel_launches_per_step": sweep_launches,
                     "actual_limiter": "on-chip latency at the occupancy the register file allows (16 warps/SM, 128 "
                                       "registers): no unit is saturated (see sm_issue_pct / fma_pipe_pct / smem_pct); "
                                       "the state stays in shared memory / L2, so the HBM-equivalent frac exceeds 1 "
                                       "by on-chip reuse, not skipped work",
                     "sm_counters": counters,
                     # the dominant launch's utilisation, flat (None unless the capture matches this build)
                     "sm_issue_pct": dom.get("sm_issue_pct"), "fma_pipe_pct": dom.get("fma_pipe_pct"),
                     "smem_pct": dom.get("smem_wavefronts_pct"), "lsu_pct": dom.get("lsu_wavefronts_pct"),
                     "dram_pct": dom.get("dram_pct"),
                     "note": "achieved = algorithmic (6F+2) x 8 x 2^n bytes per bitstring / device time, per GPU; "
                             "sm_counters / traffic come from the ncu capture named in sm_counters.source and are "
                             "only quoted when sm_counters.current (same kernel source hash)"},
        "e2e": {"value": e2e_value, "unit": "bitstrings/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "last_loss": loss},
        "parity": par,
        "gpu_launches": sweep_launches + 3,  # + prep (incl. accumulator clears), finalize, weighted sum
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
      nthreads = host_threads()
      rate0, _, threads = cpu_reference_rate(cfg, wl, min(2 * nthreads, total), nthreads)
      sample = int(max(8, min(total, rate0 * 10.0)))
      rate, dt, threads = cpu_reference_rate(cfg, wl, sample, nthreads)
      out["cpu_baseline"] = {
          "value": rate, "unit": "bitstrings/s", "cores": threads, "kind": "port",
          "sample": f"{sample} of {total} unique bitstrings, {dt:.1f} s; forward op + adjoint op; CPU restatement "
                    "of TFQ 0.6.1's algorithm (oracle/tfq_cpu.c, split re/im AVX loops), not TFQ itself"}
    print(json.dumps(out), flush=True)
    failed = par is not None and not par["pass"]
  else:
    failed = False
  if world > 1:
    dist.destroy_process_group()
  if failed and not args.no_parity_fail:
    sys.exit("bench.py: parity block failed (see the `parity` object of the JSON line)")


def merge_parity(blocks):
  """Worst case over the ranks."""
  if not blocks:
    return None
  out = dict(blocks[0])
  for b in blocks[1:]:
    for k, v in b.items():
      if k == "states_checked":
        out[k] += v
      elif k == "pass":
        out[k] = out[k] and v
      elif isinstance(v, float):
        out[k] = max(out[k], v)
  out["ranks_reporting"] = len(blocks)
  return out


# ----------------------------------------------------------------------------------------------
# EBM arm (config c5): 2^n energy sweep + logsumexp + entropy + categorical samples, row range sharded.
# ----------------------------------------------------------------------------------------------
def run_gpu_ebm(args, cfg, world, rank, local_rank, dev):
  import torch
  import torch.distributed as dist
  from qhbmlib import _native as nat
  from qhbmlib import distributed as qd
  from qhbmlib import engine

  n, n_samples = cfg["n"], cfg["samples"]
  rows = 1 << n
  widths, ws = mlp_weights(n)
  acts = ["tanh", "tanh", "linear"]
  h_ws = [torch.tensor(w).pin_memory() for w in ws]
  d_ws = [w.to(dev) for w in h_ws]
  d_bs = [torch.zeros(widths[l + 1], device=dev) for l in range(3)]
  desc = engine.EnergyDescriptor(nat.ENERGY_MLP, n, layers=[(d_ws[l], d_bs[l], acts[l]) for l in range(3)])
  lo, hi = shard_range(rows, rank, world)
  state = {}

  def step(seed):
    """_ready_inference (logits of this rank's rows + (m,s,t)) -> merged log Z / entropy -> this rank's
    share of the samples.  Collective: one all-gather of 3 doubles per rank."""
    logits, _, log_z, entropy, masses = qd.sharded_ebm_sweep(desc, n, device=dev)
    split = qd.split_samples(n_samples, masses, seed)
    first = int(split[:rank].sum())
    samples = engine.categorical_sample(logits, int(split[rank]), seed, first_sample=first, row_offset=lo)
    state.update(logits=logits, log_z=log_z, entropy=entropy, samples=samples, split=split)
    return log_z

  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  warmup = max(args.warmup, 3)
  for i in range(warmup):
    step((3, 4 + i))
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  evs, sweep_evs = [], []
  barrier()
  for i in range(args.steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    step((3, 4))
    b.record()
    evs.append((a, b))
  barrier()
  # the dominant kernel alone (the dense-stack sweep), for the roofline object
  for i in range(args.steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    desc.sweep(lo, hi, want_logits=True, device=dev)
    b.record()
    sweep_evs.append((a, b))
  barrier()
  # the sampling part alone (max, block + sub-block sums, scan, draw of this rank's share)
  samp_evs = []
  for i in range(args.steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    engine.categorical_sample(state["logits"], int(state["split"][rank]), (3, 4), first_sample=0, row_offset=lo)
    b.record()
    samp_evs.append((a, b))
  barrier()
  clocks = sampler.stop() if rank == 0 else None
  dev_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs), sum(a.elapsed_time(b) for a, b in sweep_evs),
                         sum(a.elapsed_time(b) for a, b in samp_evs)], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
  ms_per_step, ms_sweep, ms_sampling = (dev_ms / args.steps).tolist()
  value = rows / (ms_per_step / 1e3)

  # ---- end to end: weights from pinned host memory in, samples + log Z out, every step
  h_samples = torch.empty(n_samples, dtype=torch.int64).pin_memory()

  def step_e2e():
    for l in range(3):
      d_ws[l].copy_(h_ws[l], non_blocking=True)
    log_z = step((3, 4))
    k = state["samples"].shape[0]
    h_samples[:k].copy_(state["samples"], non_blocking=True)
    torch.cuda.synchronize()
    return log_z

  for _ in range(2):
    step_e2e()
  barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step_e2e()
  barrier()
  e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
  e2e_value = rows * args.steps / float(e2e_t.item())

  # ---- parity vs the committed float64 oracle values (tests/golden/bench_parity_c5.npz)
  fixture = load_parity_fixture(args.config)
  parity = None
  hist = torch.bincount((state["samples"] >> (n - 6)), minlength=64).double()  # 64 coarse bins of the row index
  if world > 1:
    dist.all_reduce(hist)
  if fixture is not None:
    pos = fixture["logit_rows"]
    mine = (pos >= lo) & (pos < hi)
    got = state["logits"][torch.tensor(pos[mine] - lo, device=dev)].double().cpu().numpy()
    ref = fixture["logits"][mine]
    lerr = float(np.abs(got - ref).max()) if mine.any() else 0.0
    lscaled = float((np.abs(got - ref) / (1e-5 * np.abs(ref) + 2e-6)).max()) if mine.any() else 0.0
    lerr_t = torch.tensor([lerr, lscaled], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(lerr_t, op=dist.ReduceOp.MAX)
    expect = fixture["bin_probabilities"] * n_samples
    chi2 = float((((hist.cpu().numpy() - expect)**2) / expect).sum())
    parity = {
        "log_z": state["log_z"], "log_z_oracle": float(fixture["log_z"]),
        "log_z_rel_err": abs(state["log_z"] - float(fixture["log_z"])) / abs(float(fixture["log_z"])),
        "entropy": state["entropy"], "entropy_oracle": float(fixture["entropy"]),
        "entropy_rel_err": abs(state["entropy"] - float(fixture["entropy"])) / abs(float(fixture["entropy"])),
        "logits_checked": int(len(pos)), "logits_max_abs_err": float(lerr_t[0].item()),
        "logits_max_err_over_tol": float(lerr_t[1].item()), "logits_tol": "1e-5 * |ref| + 2e-6",
        "samples": int(hist.sum().item()), "chi2_64_bins": chi2, "chi2_limit": 2.0 * 64,
    }
    parity["max_rel_err"] = max(parity["log_z_rel_err"], parity["entropy_rel_err"])
    parity["pass"] = bool(parity["max_rel_err"] <= 1e-5 and parity["logits_max_err_over_tol"] <= 1.0 and
                          chi2 < 2.0 * 64 and parity["samples"] == n_samples)
  if rank == 0:
    flops = rows * 2 * sum(widths[l] * widths[l + 1] for l in range(3))
    peak_tf = 148 * 128 * 2 * 1.965e9 / 1e12  # nominal fp32 FMA peak of a B200 (no measured fp32 figure exists)
    achieved = flops / world / (ms_sweep / 1e3) / 1e12
    out = {
        "metric": "rows/s (2^n energy sweep + logsumexp + entropy + 1e6 categorical samples)", "value": value,
        "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["label"], "rows": rows, "rows_per_gpu": hi - lo, "samples": n_samples,
                   "parallelism": f"row range sharded x{world}; all-gather of (max, sum exp, sum exp*l) per rank; "
                                  "rank-level multinomial split of the samples from the shared seed",
                   "l2": "256 MiB buffer written between timed steps (untimed)", "ms_sweep_kernel": ms_sweep,
                   "ms_sampling": ms_sampling,
                   "sampling_logits_GBps": (hi - lo) * 4 * 2 / (ms_sampling / 1e3) / 1e9},
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf, "traffic": None,
                     "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json has no fp32 "
                                    "non-tensor figure)",
                     "flops_per_row": flops // rows,
                     "note": "dominant kernel = ebm_mlp_sweep_kernel timed alone over this rank's rows; the dense "
                             "stack is FP32-FMA bound (11 392 flop per row against 4 B of logits written)"},
        "e2e": {"value": e2e_value, "unit": "rows/s",
                "h2d_bytes_per_step": int(sum(w.numel() for w in h_ws) * 4), "d2h_bytes_per_step": n_samples * 8 // world},
        "parity": parity, "gpu_launches": 6, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
      cpu_ebm_rate(cfg, 1 << 16)
      rate, dt = cpu_ebm_rate(cfg, 1 << 21)
      out["cpu_baseline"] = {"value": rate, "unit": "rows/s", "cores": host_threads(), "kind": "port",
                             "sample": f"{1 << 21} of {rows} rows, {dt:.1f} s; numpy float32 restatement of "
                                       "ebm.py:467-485 (sweep + logsumexp + entropy, no sampling)"}
    print(json.dumps(out), flush=True)
    failed = parity is not None and not parity["pass"]
  else:
    failed = False
  if world > 1:
    dist.destroy_process_group()
  if failed and not args.no_parity_fail:
    sys.exit("bench.py: parity block failed (see the `parity` object of the JSON line)")


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
  ap.add_argument("--no-parity-fail", action="store_true", help="report a failed parity block without exiting 1")
  args = ap.parse_args()
  cfg = CONFIGS[args.config]
  if args.impl == "reference":
    run_reference(args, cfg)
  else:
    run_gpu(args, cfg)


if __name__ == "__main__":
  main()
