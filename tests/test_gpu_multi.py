"""Two-GPU NCCL test of the sharded paths: skipped on a single-GPU box.

Every rank runs the CUDA engine on its own GPU; the sharded count-weighted expectation + gradient
and the sharded 2^n EBM sweep (log Z, entropy, rank-level sample split, local sampling) must equal
the single-GPU results of the same engine and the oracle."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qhbm_oracle as orc
import helpers as hp

pytestmark = pytest.mark.gpu


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _inputs():
  n = 13
  gates, names = orc.hea_circuit(n, 2)
  rng = np.random.default_rng(5)
  phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
  basis = rng.choice(1 << n, 301, replace=False).astype(np.int64)
  counts = rng.integers(1, 40, 301).astype(np.int32)
  ops = [orc.tfim_ring(n), orc.xxz_ring(n)]
  n_bits = 14
  thetas = rng.normal(0, 0.3, len(orc.parity_indices(n_bits, 2))).astype(np.float32)
  return n, gates, names, phi, basis, counts, ops, n_bits, thetas


def _worker(rank, world_size, port, out_dir):
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  for p in (root, os.path.join(root, "qhbm-library_b200"), os.path.join(root, "tests")):
    if p not in sys.path:
      sys.path.insert(0, p)
  from qhbmlib import distributed as qd
  from qhbmlib import engine
  from qhbmlib import _native as nat
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
  try:
    dev = torch.device("cuda", rank)
    n, gates, names, phi, basis, counts, ops, n_bits, thetas = _inputs()
    terms, offs = hp.ops_to_tables(ops, n)
    plan = engine.ExpectationPlan(gates, n, len(names), terms, offs, True)
    sharded = qd.ShardedExpectation(plan)
    avg, total, grad = sharded(torch.tensor(basis, device=dev), torch.tensor(counts, device=dev),
                               torch.tensor(phi, device=dev), grad_mode="exact")
    masks = [sum(1 << (n_bits - 1 - i) for i in grp) for grp in orc.parity_indices(n_bits, 2)]
    desc = engine.EnergyDescriptor(nat.ENERGY_KOBE, n_bits, torch.tensor(masks, dtype=torch.int32, device=dev),
                                   torch.tensor(thetas, device=dev))
    logits, (lo, hi), log_z, entropy, masses = qd.sharded_ebm_sweep(desc, n_bits, device=dev)
    split = qd.split_samples(200000, masses, (7, 8))
    rows = engine.categorical_sample(logits, int(split[rank]), (7, 8 + rank), row_offset=lo)
    hist = torch.bincount(rows, minlength=1 << n_bits).double()
    dist.all_reduce(hist)
    # the library's own collective (qhbm_comm_* / qhbm_allreduce) against torch.distributed's
    comm = qd.NativeComm()
    probe64 = torch.arange(97, dtype=torch.float64, device=dev) * (rank + 1) + 0.25 * rank
    probe32 = torch.arange(5, dtype=torch.float32, device=dev) - rank
    want64, want32 = probe64.clone(), probe32.clone()
    dist.all_reduce(want64), dist.all_reduce(want32)
    comm.all_reduce_(probe64), comm.all_reduce_(probe32)
    native_ok = bool(torch.equal(probe64, want64) and torch.equal(probe32, want32) and comm.world_size == world_size)
    os.environ["QHBM_NATIVE_ALLREDUCE"] = "1"
    avg_n, total_n, grad_n = sharded(torch.tensor(basis, device=dev), torch.tensor(counts, device=dev),
                                     torch.tensor(phi, device=dev), grad_mode="exact")
    del os.environ["QHBM_NATIVE_ALLREDUCE"]
    torch.cuda.synchronize()
    comm.close()
    qd.close_native_comms()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), avg=avg.cpu().numpy(), total=float(total), grad=grad.cpu().numpy(),
             log_z=log_z, entropy=entropy, split=split, hist=hist.cpu().numpy(), lo=lo, hi=hi,
             rows_min=int(rows.min()), rows_max=int(rows.max()), native_ok=native_ok, avg_n=avg_n.cpu().numpy(),
             grad_n=grad_n.cpu().numpy(), total_n=float(total_n))
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_expectation_and_ebm_sweep(tmp_path):
  world_size = 2
  mp.spawn(_worker, args=(world_size, _free_port(), str(tmp_path)), nprocs=world_size, join=True)
  r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
  for key in ("avg", "total", "grad", "log_z", "entropy", "split", "hist"):
    np.testing.assert_array_equal(r0[key], r1[key])  # identical on every rank after the collectives
  # qhbm_allreduce (the library's own NCCL communicator) gives torch.distributed's sums, alone and behind the
  # sharded expectation (the per-rank partial sums may differ in their last float64 bit between two runs: atomics)
  assert bool(r0["native_ok"]) and bool(r1["native_ok"])
  np.testing.assert_allclose(r0["avg_n"], r0["avg"], rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(r0["grad_n"], r0["grad"], rtol=1e-6, atol=1e-7)
  np.testing.assert_array_equal(r0["avg_n"], r1["avg_n"])
  assert float(r0["total_n"]) == float(r0["total"])
  n, gates, names, phi, basis, counts, ops, n_bits, thetas = _inputs()
  dg = np.tile((counts / counts.sum())[:, None], (1, 2))
  e, g = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg)
  np.testing.assert_allclose(r0["avg"], orc.weighted_average(counts, e), rtol=1e-5, atol=1e-5)
  assert r0["total"] == counts.sum()
  np.testing.assert_allclose(r0["grad"], g.sum(0), rtol=1e-5, atol=2e-5)
  energies = orc.kobe_energy(orc.all_bitstrings(n_bits), 2, thetas.astype(np.float64))
  np.testing.assert_allclose(float(r0["log_z"]), orc.analytic_log_partition(energies), rtol=1e-6)
  np.testing.assert_allclose(float(r0["entropy"]), orc.analytic_entropy(energies), rtol=1e-5)
  # every rank sampled inside its own row range; together the samples follow p(x)
  half = 1 << (n_bits - 1)
  assert (int(r0["lo"]), int(r0["hi"])) == (0, half) and (int(r1["lo"]), int(r1["hi"])) == (half, 2 * half)
  assert 0 <= int(r0["rows_min"]) and int(r0["rows_max"]) < half <= int(r1["rows_min"]) and int(r1["rows_max"]) < 2 * half
  assert int(r0["split"].sum()) == 200000 and r0["hist"].sum() == 200000
  p = orc.analytic_probabilities(energies)
  p0 = p[:half].sum()
  assert abs(r0["split"][0] / 200000 - p0) < 5 * math.sqrt(p0 * (1 - p0) / 200000)
  # coarse chi-square over 64 bins of the row index
  obs = r0["hist"].reshape(64, -1).sum(1)
  exp = p.reshape(64, -1).sum(1) * 200000
  assert np.sum((obs - exp)**2 / exp) < 2.5 * 64


# ------------------------------------------------------------------------------------------------
# Sharding behind the API: vqt / qmhl losses, their gradients and the analytic sampler must be the same at
# world size 1 and 2 (north_star: "the API stays intact").
# ------------------------------------------------------------------------------------------------
def _api_results(dev):
  """VQT and QMHL losses + gradients and a sample draw, built from fixed seeds on device `dev`.
  Under an initialised process group the inference engines shard by themselves."""
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  for p in (root, os.path.join(root, "qhbm-library_b200")):
    if p not in sys.path:
      sys.path.insert(0, p)
  from qhbmlib import architectures as arch
  from qhbmlib import circuits as cq
  from qhbmlib import data as qdata
  from qhbmlib import distributed as qd
  from qhbmlib import inference
  from qhbmlib import models
  from qhbmlib.models import energy_utils
  n, num_samples = 12, 3000
  qubits = cq.GridQubit.rect(1, n)
  out = {}

  def make_qhbm(tag, seed):
    energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.3, seed))
    e_inf = inference.AnalyticEnergyInference(energy, num_samples, initial_seed=[seed, seed + 1])
    circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 2, tag),
                                       energy_utils.RandomUniform(-1, 1, seed + 2))
    return inference.QHBM(e_inf, inference.AnalyticQuantumInference(circ, grad_mode="exact")), energy, circ

  # --- VQT
  qhbm, energy, circ = make_qhbm("v", 21)
  loss = inference.vqt(qhbm, cq.convert_to_tensor([arch.tfim_ring(qubits)]), torch.tensor(0.7, device=dev))
  params = [energy.post_process[0].kernel, circ.trainable_variables[0]]
  loss.backward()
  qd.sync_gradients(params)
  out["vqt_loss"] = np.array([float(loss)])
  out["vqt_g_theta"], out["vqt_g_phi"] = [p.grad.detach().double().cpu().numpy() for p in params]
  out["samples"] = qhbm.e_inference.sample(20000).to(torch.int64).sum(1).cpu().numpy()  # popcount per sample
  out["entropy"] = np.array([float(qhbm.e_inference.entropy())])
  # --- QMHL: data QHBM (fixed) against a model QHBM
  data_qhbm, _, _ = make_qhbm("d", 31)
  model, m_energy, m_circ = make_qhbm("m", 41)
  loss = inference.qmhl(qdata.QHBMData(data_qhbm), model)
  params = [m_energy.post_process[0].kernel, m_circ.trainable_variables[0]]
  loss.backward()
  qd.sync_gradients(params)
  out["qmhl_loss"] = np.array([float(loss)])
  out["qmhl_g_theta"], out["qmhl_g_phi"] = [p.grad.detach().double().cpu().numpy() for p in params]
  return out


def _api_worker(rank, world_size, port, out_dir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
  try:
    np.savez(os.path.join(out_dir, f"api{rank}.npz"), **_api_results(torch.device("cuda", rank)))
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_vqt_and_qmhl_equal_single_gpu(tmp_path):
  world_size = 2
  mp.spawn(_api_worker, args=(world_size, _free_port(), str(tmp_path)), nprocs=world_size, join=True)
  r0, r1 = np.load(tmp_path / "api0.npz"), np.load(tmp_path / "api1.npz")
  torch.cuda.set_device(0)
  single = _api_results(torch.device("cuda", 0))
  for key in single:
    np.testing.assert_array_equal(r0[key], r1[key])              # all ranks agree exactly
  # the sharded analytic sampler reproduces the single-GPU draw (same seed, any number of ranks)
  np.testing.assert_array_equal(r0["samples"], single["samples"])  # (shards sweep bit-identical logits)
  np.testing.assert_allclose(r0["entropy"], single["entropy"], rtol=1e-6)
  for key in ("vqt_loss", "qmhl_loss"):
    np.testing.assert_allclose(r0[key], single[key], rtol=2e-6, atol=2e-6)
  for key in ("vqt_g_theta", "vqt_g_phi", "qmhl_g_theta", "qmhl_g_phi"):
    scale = np.abs(single[key]).max()
    np.testing.assert_allclose(r0[key], single[key], rtol=1e-5, atol=2e-6 * max(scale, 1.0))
    assert scale > 1e-3
