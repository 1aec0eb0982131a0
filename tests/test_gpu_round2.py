"""GPU tests of the round-2 additions: fused score-function gradient, two-step categorical
sampler and its sharded mode, the 2^24-row MLP sweep at BASELINE config 5's own size, the
content-keyed plan cache, stream ordering of one plan used from two streams, the hash-unique sentinel key,
and a deep circuit whose passes have to be split to fit the program staging buffers."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import qhbm_oracle as orc
from qhbmlib import _native as nat
from qhbmlib import architectures as arch
from qhbmlib import circuits as cq
from qhbmlib import engine
from qhbmlib import inference
from qhbmlib import models
from qhbmlib import utils
from qhbmlib.models import energy_utils
import helpers as hp

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------- f1
def test_score_gradient_kernel_against_oracle():
  """qhbm_score_gradient == E[c]E[dE] - E[c dE] of ebm.py:282-325 with the +-1 parity Jacobian."""
  rng = np.random.default_rng(0)
  n, order, u, width = 9, 2, 300, 3
  groups = orc.parity_indices(n, order)
  masks = np.array([sum(1 << (n - 1 - i) for i in g) for g in groups], dtype=np.int64)
  keys = rng.choice(1 << n, u, replace=False).astype(np.int64)
  counts = rng.integers(1, 50, u).astype(np.int32)
  vals = rng.normal(0, 2, (u, width)).astype(np.float32)
  upstream = rng.normal(0, 1, width).astype(np.float32)
  bits = ((keys[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1).astype(np.int8)
  jac = orc.parity_features(bits, order)
  avg = orc.weighted_average(counts, vals.astype(np.float64))
  ref = orc.expectation_score_gradient(counts, vals, jac, np.zeros(len(groups)), upstream)
  for lo, hi, scale in [(0, u, 1.0), (0, 120, 2.0), (120, u, 2.0)]:
    got = engine.score_gradient(
        torch.tensor(keys[lo:hi], device=DEV), torch.tensor(counts[lo:hi], device=DEV),
        torch.tensor(vals[lo:hi], device=DEV), torch.tensor(upstream, device=DEV),
        torch.tensor(avg.astype(np.float32), device=DEV), torch.tensor(masks, device=DEV).to(torch.int32),
        torch.tensor([float(counts.sum())], dtype=torch.float64, device=DEV), scale).double().cpu().numpy()
    if (lo, hi) == (0, u):
      np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6)
    elif lo == 0:
      part = got
    else:  # two shards with scale = world size average to the full gradient (sync_gradients convention)
      np.testing.assert_allclose((part + got) / 2.0, ref, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("kind", ["kobe", "bernoulli"])
def test_fused_score_term_equals_torch_surrogate(kind, monkeypatch):
  """EnergyInference._expectation with the fused kernel == the generic autograd surrogate, for a nested
  result structure and parameters shared between the energy and the function."""
  n = 6
  if kind == "kobe":
    energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.4, 7))
    make = lambda: inference.AnalyticEnergyInference(energy, 3000, initial_seed=[5, 6])
  else:
    energy = models.BernoulliEnergy(list(range(n)), energy_utils.RandomNormal(0.0, 0.7, 8))
    make = lambda: inference.BernoulliEnergyInference(energy, 3000, initial_seed=[5, 6])
  w = torch.nn.Parameter(torch.linspace(-1, 1, n, device=DEV))
  theta = energy.post_process[0].kernel

  def function(bits):
    x = bits.to(torch.float32)
    return {"lin": torch.stack([x @ w, energy(bits).detach() * (x @ w)], 1), "sq": [(x @ w) ** 2]}

  def loss_and_grads(fused):
    inf = make()
    if not fused:
      monkeypatch.setattr(inf, "_parity_feature_tables", lambda: None)
    out = inf.expectation(function)
    loss = (out["lin"] * torch.tensor([0.7, -1.3], device=DEV)).sum() + 0.4 * out["sq"][0]
    g = torch.autograd.grad(loss, (theta, w))
    return float(loss), g[0].double().cpu().numpy(), g[1].double().cpu().numpy()

  l1, gt1, gw1 = loss_and_grads(True)
  l0, gt0, gw0 = loss_and_grads(False)
  assert l1 == l0
  np.testing.assert_allclose(gw1, gw0, rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(gt1, gt0, rtol=2e-5, atol=2e-6)
  assert np.abs(gt0).max() > 1e-3


# ------------------------------------------------------------------------------- categorical sampler
def test_sampler_prepare_once_draw_many_and_sharded_draw_equals_single_draw():
  rng = np.random.default_rng(3)
  rows = 1 << 17
  logits_np = rng.normal(0, 2.5, rows).astype(np.float32)
  logits = torch.tensor(logits_np, device=DEV)
  gmax = float(logits_np.max())
  n_samples, seed = 300_000, (11, 12)
  ref = engine.categorical_sample(logits, n_samples, seed)
  smp = engine.CategoricalSampler(logits, given_max=gmax)          # max from the sweep statistics
  assert torch.equal(smp.draw(n_samples, seed), ref)
  assert torch.equal(smp.draw(n_samples, seed), ref)              # prepared state is reusable
  assert not torch.equal(smp.draw(n_samples, (11, 13)), ref)
  assert torch.equal(smp.draw(1000, seed, first_sample=5000), ref[5000:6000])
  np.testing.assert_allclose(float(smp.local_mass()), np.exp(logits_np.astype(np.float64) - gmax).sum(), rtol=1e-12)
  for world in (2, 3, 8):
    bounds = [(r * rows // world, (r + 1) * rows // world) for r in range(world)]
    shards = [engine.CategoricalSampler(logits[lo:hi].contiguous(), given_max=gmax) for lo, hi in bounds]
    cum = [0.0]
    for s in shards:
      cum.append(cum[-1] + float(s.local_mass()))
    total = torch.zeros(n_samples, dtype=torch.int64, device=DEV)
    written = torch.zeros(n_samples, dtype=torch.int64, device=DEV)
    for r, (s, (lo, hi)) in enumerate(zip(shards, bounds)):
      out = torch.full((n_samples,), -1, dtype=torch.int64, device=DEV)
      end = cum[r + 1] if r + 1 < world else math.inf
      s.draw(n_samples, seed, row_offset=lo, mass_interval=(cum[r], end, cum[-1]), out=out)
      mine = out >= 0
      assert bool(((out[mine] >= lo) & (out[mine] < hi)).all())
      written += mine.to(torch.int64)
      total += torch.where(mine, out, torch.zeros_like(out))
    assert bool((written == 1).all())                              # every sample drawn by exactly one rank
    # identical to the single-GPU draw up to float64 rounding of the total mass (a sample can move to a
    # neighbouring row only if its uniform lands within ~1e-16 of a boundary)
    assert int((total != ref).sum()) <= 2


def test_unique_handles_the_all_ones_key():
  """the key equal to the hash table's empty marker (int64 -1) has its own slot."""
  keys = torch.tensor([5, -1, 7, -1, 5, -1, 0, 7], dtype=torch.int64, device=DEV)
  uniq, idx, count = engine.unique_with_counts(keys)
  assert uniq.tolist() == [5, -1, 7, 0] and count.tolist() == [2, 3, 2, 1]
  assert idx.tolist() == [0, 1, 2, 1, 0, 1, 3, 2]
  big = torch.randint(-3, 3, (100_000,), dtype=torch.int64, device=DEV)
  uniq, idx, count = engine.unique_with_counts(big)
  ref_u, ref_c = np.unique(big.cpu().numpy(), return_counts=True)
  assert sorted(uniq.tolist()) == ref_u.tolist() and int(count.sum()) == 100_000
  assert torch.equal(uniq[idx.long()], big)
  assert dict(zip(uniq.tolist(), count.tolist())) == dict(zip(ref_u.tolist(), ref_c.tolist()))


# -------------------------------------------------------------------------- config 5 at its own size
def test_mlp_sweep_over_2p24_rows_against_float64_values():
  """BASELINE config 5: AnalyticEnergyInference over all 2^24 bitstrings with the 24-64-64-1 tanh stack.
  Reference values: float64 numpy evaluation of the same stack (tests/golden/make_bench_parity.py c5):
  log Z, entropy, 4108 logits spread over the range (including both ends and the 8-way shard boundaries)
  and the probability mass of 64 coarse bins for the sample histogram."""
  sys.path.insert(0, ROOT)
  import bench
  fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "bench_parity_c5.npz")))
  n = 24
  widths, ws = bench.mlp_weights(n)
  lin = [FloatLinear(widths[0], widths[1])] + [torch.nn.Linear(widths[l], widths[l + 1]) for l in (1, 2)]
  with torch.no_grad():
    for l in range(3):
      lin[l].weight.copy_(torch.tensor(ws[l]).t())
      lin[l].bias.zero_()
  energy = models.BitstringEnergy(list(range(n)), [lin[0], torch.nn.Tanh(), lin[1], torch.nn.Tanh(), lin[2],
                                                   utils.Squeeze(-1)])
  inf = inference.AnalyticEnergyInference(energy, 1000, initial_seed=[3, 4])
  from qhbmlib.inference import ebm
  assert ebm.energy_descriptor(energy) is not None                      # the CUDA dense-stack sweep is used
  with torch.no_grad():
    log_z, entropy = float(inf.log_partition()), float(inf.entropy())
  np.testing.assert_allclose(log_z, float(fx["log_z"]), rtol=1e-6)      # float32 scalars returned by the API
  np.testing.assert_allclose(entropy, float(fx["entropy"]), rtol=1e-6)
  m, s, t = inf._stats.tolist()                                          # the float64 statistics behind them
  np.testing.assert_allclose(m + math.log(s), float(fx["log_z"]), rtol=1e-8)
  np.testing.assert_allclose(m + math.log(s) - t / s, float(fx["entropy"]), rtol=1e-8)
  got = inf.distribution.logits_parameter()[torch.tensor(fx["logit_rows"], device=DEV)].double().cpu().numpy()
  np.testing.assert_allclose(got, fx["logits"], rtol=1e-5, atol=2e-6)
  n_samples = 1_000_000
  bits = inf.sample(n_samples)
  assert bits.shape == (n_samples, n) and bits.dtype == torch.int8
  rows = engine.pack_bits(bits, utils._natural_shifts(n))
  hist = torch.bincount(rows >> (n - 6), minlength=64).double().cpu().numpy()
  expect = fx["bin_probabilities"] * n_samples
  assert ((hist - expect) ** 2 / expect).sum() < 2.0 * 64             # chi-square, 63 degrees of freedom
  # fixed seed: the draw repeats exactly; a new preface step changes it (ebm_test.py:280-297)
  assert torch.equal(inf.sample(n_samples), bits)
  inf.seed = None
  assert not torch.equal(inf.sample(n_samples), bits)


class FloatLinear(torch.nn.Linear):
  """Dense layer on raw int8 bits (the Keras Dense of the reference casts its input the same way)."""

  def forward(self, x):
    return super().forward(x.to(torch.float32))


# ----------------------------------------------------------------------------------- plan cache / streams
def test_plan_cache_reuses_plans_for_rebuilt_observables_and_stays_bounded():
  """an observable rebuilt every training step must not compile (and keep) a new plan."""
  n = 5
  qubits = cq.GridQubit.rect(1, n)
  circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 1, "p"))
  q = inference.AnalyticQuantumInference(circ)
  states = torch.tensor([[0, 1, 0, 1, 1], [1, 1, 0, 0, 0]], dtype=torch.int8, device=DEV)
  first = q.expectation(states, cq.convert_to_tensor([arch.tfim_ring(qubits)]))
  for _ in range(5):
    again = q.expectation(states, cq.convert_to_tensor([arch.tfim_ring(qubits)]))  # fresh object, same content
    assert torch.equal(first, again)
  assert len(q._plans) == 1
  for k in range(12):
    q.expectation(states, cq.convert_to_tensor([arch.tfim_ring(qubits, bias=0.1 * (k + 1))]))
  assert len(q._plans) <= 8


def test_one_plan_used_from_two_streams_is_ordered():
  """a plan's scratch buffers are shared by its calls; calls arriving on different streams are
  serialised on the device by the per-plan event instead of racing."""
  n = 12
  gates, names = orc.hea_circuit(n, 2)
  ops = [orc.xxz_ring(n)]
  terms, offs = hp.ops_to_tables(ops, n)
  plan = engine.ExpectationPlan(gates, n, len(names), terms, offs, True)
  rng = np.random.default_rng(1)
  phis = [torch.tensor(rng.uniform(-1, 1, len(names)).astype(np.float32), device=DEV) for _ in range(2)]
  basis = torch.tensor(rng.choice(1 << n, 2048, replace=False).astype(np.int64), device=DEV)
  dg = torch.ones((2048, 1), device=DEV)
  ref = [plan.forward_adjoint(basis, p, dg) for p in phis]
  torch.cuda.synchronize()
  streams = [torch.cuda.Stream(), torch.cuda.Stream()]
  for _ in range(3):
    outs = []
    for s, p in zip(streams, phis):
      s.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(s):
        outs.append(plan.forward_adjoint(basis, p, dg))
    torch.cuda.synchronize()
    for (e, g), (e0, g0) in zip(outs, ref):
      assert torch.equal(e, e0)
      np.testing.assert_allclose(g.cpu().numpy(), g0.cpu().numpy(), rtol=1e-5, atol=1e-4)


# ----------------------------------------------------------------------- program staging limits
@pytest.mark.parametrize("n,layers,T,K", [(9, 9, 0, 4), (12, 6, 10, 4), (13, 5, 0, 5)])
def test_deep_circuits_split_passes_to_fit_the_staging_buffers(n, layers, T, K):
  """Every pass's program is staged in a fixed shared-memory buffer and prefetched with cp.async; deep
  circuits give the scheduler long runs of ready gates, which it must cut into passes that fit."""
  rng = np.random.default_rng(n * 10 + layers)
  gates, names = orc.hea_circuit(n, layers)
  ops = [orc.xxz_ring(n), orc.tfim_ring(n)]
  terms, offs = hp.ops_to_tables(ops, n)
  plan = engine.ExpectationPlan(gates, n, len(names), terms, offs, True, T, K)
  phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
  basis = rng.choice(1 << n, 3, replace=False).astype(np.int64)
  dg = rng.uniform(-1, 1, (3, 2)).astype(np.float32)
  e_ref, g_ref = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg)
  e, g = plan.forward_adjoint(torch.tensor(basis, device=DEV), torch.tensor(phi, device=DEV),
                              torch.tensor(dg, device=DEV), per_state=True)
  scale = np.array([sum(abs(c) for c, _ in op) for op in ops])
  np.testing.assert_allclose(e.cpu().numpy(), e_ref, rtol=1e-5, atol=1e-6 * scale.max())
  floor = 1e-6 * (np.abs(dg) * scale[None, :]).sum(1)[:, None] * math.sqrt(layers)
  assert (np.abs(g.cpu().numpy() - g_ref) <= 1e-5 * np.abs(g_ref) + floor).all()


# ------------------------------------------------------------------ shards sweep bit-identical logits
@pytest.mark.parametrize("kind", ["kobe", "mlp"])
def test_shard_sweeps_reproduce_the_full_sweep_bit_for_bit(kind):
  """A rank that sweeps a shard of the 2^n rows must produce the logits a single sweep of the whole range
  produces, bit for bit (the kernel is chosen from the problem size, not from the shard's row count):
  only then does the rank-sharded sampler draw exactly the single-GPU samples.  Covers aligned 8-way
  shards and an unaligned split, and the merged (m, s, t) statistics."""
  from qhbmlib import distributed as qd
  from qhbmlib.inference import ebm
  n = 12
  if kind == "kobe":
    energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.3, 21))
  else:
    torch.manual_seed(3)
    energy = models.BitstringEnergy(list(range(n)), [FloatLinear(n, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32),
                                                     torch.nn.Tanh(), torch.nn.Linear(32, 1), utils.Squeeze(-1)])
  inf = inference.AnalyticEnergyInference(energy, 10, initial_seed=[1, 2])
  with torch.no_grad():
    log_z = float(inf.log_partition())
  desc = ebm.energy_descriptor(inf.energy)
  assert desc is not None
  full, stats = desc.sweep(0, 1 << n, device=DEV)
  assert torch.equal(full, inf._logits)
  for bounds in ([i << 9 for i in range(9)], [0, 100, 1357, 1358, 2048, 4000, 4096]):
    triples = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
      part, st = desc.sweep(lo, hi, device=DEV)
      assert torch.equal(part, full[lo:hi]), (kind, lo, hi)
      triples.append(st.cpu().tolist())
    m, s, _ = qd.merge_log_stats(triples)
    np.testing.assert_allclose(m + math.log(s), log_z, rtol=1e-6)


def test_native_communicator_of_one_rank():
  """qhbm_comm_create / qhbm_allreduce on a one-rank communicator: NCCL binds at run time, the in-place sum of one
  rank is the identity for both dtypes, on torch's current stream."""
  from qhbmlib import distributed as qd
  comm = qd.NativeComm(rank=0, world_size=1)
  assert (comm.rank, comm.world_size) == (0, 1) and comm.nccl_version >= 22000
  x64 = torch.linspace(-3, 3, 97, dtype=torch.float64, device="cuda")
  x32 = torch.arange(1000, dtype=torch.float32, device="cuda")
  w64, w32 = x64.clone(), x32.clone()
  side = torch.cuda.Stream()
  side.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(side):
    comm.all_reduce_(x64)
  comm.all_reduce_(x32)
  side.synchronize()
  torch.cuda.synchronize()
  assert torch.equal(x64, w64) and torch.equal(x32, w32)
  with pytest.raises(TypeError, match="float32 or float64"):
    comm.all_reduce_(torch.zeros(3, dtype=torch.int32, device="cuda"))
  comm.close()
