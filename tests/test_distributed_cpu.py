"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, the single packed
all-reduce, log-partition merging and the rank-level sample split.  The per-shard compute is
stood in by the oracle here; the CUDA path itself is covered by the `-m gpu` tests and bench.py."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qhbm_oracle as orc


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  for p in (root, os.path.join(root, "qhbm-library_b200")):
    if p not in sys.path:
      sys.path.insert(0, p)
  from qhbmlib import distributed as qd
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world_size)
  try:
    n = 5
    gates, names = orc.hea_circuit(n, 2)
    phi = np.random.default_rng(0).uniform(-1, 1, len(names))
    rng = np.random.default_rng(1)
    basis = rng.choice(1 << n, 21, replace=False)
    counts = rng.integers(1, 50, 21)
    ops = [orc.tfim_ring(n), orc.xxz_ring(n)]
    lo, hi = qd.shard_range(len(basis), rank, world_size)
    dg = np.tile((counts[lo:hi] / counts.sum())[:, None], (1, 2))
    e, g = orc.batch_expectation_and_gradient(gates, n, phi, basis[lo:hi], ops, dg)
    packed = qd.pack(torch.tensor((counts[lo:hi, None] * e).sum(0)), torch.tensor([float(counts[lo:hi].sum())]),
                     torch.tensor(g.sum(0)))
    qd.all_reduce_packed(packed)
    avg, total, grad = qd.unpack(packed, 2)
    # EBM sweep merge
    thetas = np.random.default_rng(2).normal(0, 0.5, len(orc.parity_indices(n, 2)))
    energies = orc.kobe_energy(orc.all_bitstrings(n), 2, thetas)
    rlo, rhi = qd.shard_range(1 << n, rank, world_size)
    l = -energies[rlo:rhi]
    m = l.max()
    stats = torch.tensor([m, np.exp(l - m).sum(), (np.exp(l - m) * l).sum()], dtype=torch.float64)
    triples = qd.all_gather_stats(stats)
    mm, ss, tt = qd.merge_log_stats(triples)
    masses = [t[0] + math.log(t[1]) for t in triples]
    split = qd.split_samples(100000, masses, (3, 4))
    np.save(os.path.join(out_dir, f"r{rank}.npy"),
            np.concatenate([avg.numpy(), [float(total)], grad.numpy(), [mm + math.log(ss), mm + math.log(ss) - tt / ss],
                            split.astype(np.float64), [lo, hi]]))
  finally:
    dist.destroy_process_group()


def test_two_rank_packed_allreduce_and_sweep_merge(tmp_path):
  world_size = 2
  mp.spawn(_worker, args=(world_size, _free_port(), str(tmp_path)), nprocs=world_size, join=True)
  r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
  np.testing.assert_array_equal(r0[:-2], r1[:-2])  # every rank ends with the same reduced values
  assert (r0[-2], r0[-1]) == (0, 11) and (r1[-2], r1[-1]) == (11, 21)
  n = 5
  gates, names = orc.hea_circuit(n, 2)
  phi = np.random.default_rng(0).uniform(-1, 1, len(names))
  rng = np.random.default_rng(1)
  basis = rng.choice(1 << n, 21, replace=False)
  counts = rng.integers(1, 50, 21)
  ops = [orc.tfim_ring(n), orc.xxz_ring(n)]
  dg = np.tile((counts / counts.sum())[:, None], (1, 2))
  e, g = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg)
  nsym = len(names)
  np.testing.assert_allclose(r0[:2], orc.weighted_average(counts, e), rtol=1e-12)
  assert r0[2] == counts.sum()
  np.testing.assert_allclose(r0[3:3 + nsym], g.sum(0), rtol=1e-10, atol=1e-12)
  thetas = np.random.default_rng(2).normal(0, 0.5, len(orc.parity_indices(n, 2)))
  energies = orc.kobe_energy(orc.all_bitstrings(n), 2, thetas)
  np.testing.assert_allclose(r0[3 + nsym], orc.analytic_log_partition(energies), rtol=1e-12)
  np.testing.assert_allclose(r0[4 + nsym], orc.analytic_entropy(energies), rtol=1e-10)
  split = r0[5 + nsym:7 + nsym]
  assert split.sum() == 100000
  p = orc.analytic_probabilities(energies)
  p0 = p[:16].sum()
  assert abs(split[0] / 100000 - p0) < 5 * math.sqrt(p0 * (1 - p0) / 100000)


def test_shard_range_covers_everything():
  from qhbmlib import distributed as qd
  for n in (0, 1, 7, 4096, 65537):
    for ws in (1, 2, 3, 8):
      spans = [qd.shard_range(n, r, ws) for r in range(ws)]
      assert spans[0][0] == 0 and spans[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      sizes = [b - a for a, b in spans]
      assert max(sizes) - min(sizes) <= 1


def test_merge_log_stats_matches_logsumexp():
  from qhbmlib import distributed as qd
  rng = np.random.default_rng(0)
  l = rng.normal(0, 5, 1000)
  parts = np.array_split(l, 7)
  triples = [(p.max(), np.exp(p - p.max()).sum(), (np.exp(p - p.max()) * p).sum()) for p in parts]
  m, s, t = qd.merge_log_stats(triples)
  np.testing.assert_allclose(m + math.log(s), orc.logsumexp(l), rtol=1e-13)
  pr = np.exp(l - orc.logsumexp(l))
  np.testing.assert_allclose(m + math.log(s) - t / s, -(pr * np.log(pr)).sum(), rtol=1e-10)


# ------------------------------------------------------------------------------------------------
# Sharding behind the API (host logic on CPU): EnergyInference._expectation / _log_partition with the
# differentiable collectives of qhbmlib.distributed.  The sampler is replaced by fixed unique bitstrings so
# that no CUDA kernel is needed; values and gradients at world size 2 must equal the single-process ones.
# ------------------------------------------------------------------------------------------------
def _make_inference():
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  for p in (root, os.path.join(root, "qhbm-library_b200")):
    if p not in sys.path:
      sys.path.insert(0, p)
  from qhbmlib import models
  from qhbmlib.inference import ebm
  from qhbmlib.models import energy_utils

  class FixedSamples(ebm.EnergyInference):
    """EnergyInference whose `unique_samples` returns a fixed dedup result (identical on every rank)."""

    def __init__(self, energy):
      super().__init__(energy, 100, initial_seed=5)
      rng = np.random.default_rng(9)
      rows = rng.choice(1 << 5, 13, replace=False)
      self.bits = torch.tensor(((rows[:, None] >> (4 - np.arange(5))[None, :]) & 1).astype(np.int8))
      self.counts = torch.tensor(rng.integers(1, 30, 13).astype(np.int32))

    def unique_samples(self, num_samples):
      return self.bits, None, self.counts

    def _ready_inference(self):
      pass

    def _call(self, inputs, *a, **k):
      raise NotImplementedError

    def _sample(self, n):
      raise NotImplementedError

    def _log_partition_forward_pass(self):
      return torch.tensor(0.25)

  energy = models.KOBE(list(range(5)), 2, energy_utils.RandomNormal(0.0, 0.5, 3))
  inf = FixedSamples(energy)
  w = torch.nn.Parameter(torch.tensor([0.3, -0.7, 1.1, 0.2, -0.4], dtype=torch.float32))

  def function(bits):
    x = bits.to(torch.float32)
    return {"a": torch.stack([x @ w, (x * x) @ (w * w)], 1), "b": [torch.sin(x @ w)]}

  return inf, energy, w, function


def _loss_and_grads():
  from qhbmlib import distributed as qd
  inf, energy, w, function = _make_inference()
  out = inf.expectation(function)
  loss = (out["a"] * torch.tensor([1.0, -2.0])).sum() + 3.0 * out["b"][0] + 0.5 * inf.log_partition()
  loss = loss + 0.1 * (w * w).sum()  # a replicated (unsharded) term must survive the gradient averaging
  loss.backward()
  params = [w] + list(energy.parameters())
  qd.sync_gradients(params)
  return np.concatenate([[float(loss)]] + [p.grad.detach().double().reshape(-1).numpy() for p in params])


def _api_worker(rank, world_size, port, out_dir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world_size)
  try:
    np.save(os.path.join(out_dir, f"api{rank}.npy"), _loss_and_grads())
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 3])
def test_sharded_expectation_and_log_partition_match_single_process(tmp_path, world_size):
  single = _loss_and_grads()
  mp.spawn(_api_worker, args=(world_size, _free_port(), str(tmp_path)), nprocs=world_size, join=True)
  outs = [np.load(tmp_path / f"api{r}.npy") for r in range(world_size)]
  for o in outs[1:]:
    np.testing.assert_array_equal(outs[0], o)           # every rank holds the same loss and gradients
  np.testing.assert_allclose(outs[0], single, rtol=2e-6, atol=2e-7)


def _collective_worker(rank, world_size, port, out_dir):
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  for p in (root, os.path.join(root, "qhbm-library_b200")):
    if p not in sys.path:
      sys.path.insert(0, p)
  from qhbmlib import distributed as qd
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world_size)
  try:
    x = torch.nn.Parameter(torch.arange(7, dtype=torch.float64) + 1.0)
    lo, hi = qd.shard_range(7, rank, world_size)
    rows = (x[lo:hi] ** 2).unsqueeze(1)                 # this rank's rows of a [7, 1] result
    full = qd.all_gather_rows(rows, 7)                  # [7, 1] on every rank
    total = qd.all_reduce_sum((x[lo:hi] ** 3).sum().reshape(1))
    with qd.local_shard():
      assert not qd.active()
    assert qd.active()
    loss = (full[:, 0] * torch.arange(7, dtype=torch.float64)).sum() + 2.0 * total[0]
    loss.backward()
    qd.sync_gradients([x])
    np.save(os.path.join(out_dir, f"c{rank}.npy"), np.concatenate([[float(loss)], x.grad.numpy()]))
  finally:
    dist.destroy_process_group()


def test_differentiable_collectives_give_the_single_process_gradient(tmp_path):
  mp.spawn(_collective_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
  r0, r1 = np.load(tmp_path / "c0.npy"), np.load(tmp_path / "c1.npy")
  np.testing.assert_array_equal(r0, r1)
  x = np.arange(7) + 1.0
  np.testing.assert_allclose(r0[0], (x**2 * np.arange(7)).sum() + 2 * (x**3).sum(), rtol=1e-14)
  np.testing.assert_allclose(r0[1:], 2 * x * np.arange(7) + 6 * x**2, rtol=1e-14)
