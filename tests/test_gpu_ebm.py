"""GPU parity tests of the classical (EBM) kernels through the C ABI: bit packing,
first-occurrence unique (bit-exact), energies, the 2^n logits/logsumexp/entropy sweep,
categorical and Bernoulli sampling (fixed-seed repeatability + count distributions)."""
import numpy as np
import pytest
import torch

from oracle import qhbm_oracle as orc

pytestmark = pytest.mark.gpu


def _eng():
  from qhbmlib import engine
  return engine


def test_pack_unpack_and_reference_bit_order():
  eng = _eng()
  rng = np.random.default_rng(0)
  for n in (1, 3, 10, 12, 20):
    bits = rng.integers(0, 2, size=(257, n)).astype(np.int8)
    pi = orc.bit_column_to_qubit(n)
    shifts = [n - 1 - pi[j] for j in range(n)]
    keys = eng.pack_bits(torch.tensor(bits, device="cuda"), shifts)
    np.testing.assert_array_equal(keys.cpu().numpy(), orc.bitstrings_to_index(bits))
    back = eng.unpack_bits(keys, n, shifts).cpu().numpy()
    np.testing.assert_array_equal(back, bits)


@pytest.mark.parametrize("n_rows,n_bits", [(8, 3), (1, 5), (1000, 4), (5000, 12), (200000, 16),
                                            (1 << 20, 24)])
def test_unique_with_counts_first_occurrence_bit_exact(n_rows, n_bits):
  eng = _eng()
  rng = np.random.default_rng(n_rows)
  if n_rows == 8:  # reference golden vector, tests/utils_test.py:151-186
    keys = np.array([5, 7, 3, 5, 7, 3, 5, 5], dtype=np.int64)
  else:
    p = rng.dirichlet(np.ones(min(1 << n_bits, 4096)) * 0.3)
    support = rng.choice(1 << n_bits, size=len(p), replace=False)
    keys = support[rng.choice(len(p), size=n_rows, p=p)].astype(np.int64)
  uq, idx, cnt = eng.unique_with_counts(torch.tensor(keys, device="cuda"))
  _, first = np.unique(keys, return_index=True)
  order = np.sort(first)
  exp_u = keys[order]
  pos = {int(k): i for i, k in enumerate(exp_u)}
  exp_idx = np.array([pos[int(k)] for k in keys], dtype=np.int32)
  exp_cnt = np.bincount(exp_idx, minlength=len(exp_u)).astype(np.int32)
  np.testing.assert_array_equal(uq.cpu().numpy(), exp_u)
  np.testing.assert_array_equal(idx.cpu().numpy(), exp_idx)
  np.testing.assert_array_equal(cnt.cpu().numpy(), exp_cnt)
  if n_rows == 8:
    np.testing.assert_array_equal(exp_idx, [0, 1, 2, 0, 1, 2, 0, 0])
    np.testing.assert_array_equal(exp_cnt, [4, 2, 2])


def test_unique_empty():
  eng = _eng()
  uq, idx, cnt = eng.unique_with_counts(torch.zeros(0, dtype=torch.int64, device="cuda"))
  assert uq.numel() == 0 and idx.numel() == 0 and cnt.numel() == 0


def test_segment_sum_and_weighted_sum():
  eng = _eng()
  rng = np.random.default_rng(1)
  vals = rng.normal(size=(1000, 3)).astype(np.float32)
  idx = rng.integers(0, 17, size=1000).astype(np.int32)
  out = eng.segment_sum(torch.tensor(vals, device="cuda"), torch.tensor(idx, device="cuda"), 17)
  ref = np.zeros((17, 3))
  np.add.at(ref, idx, vals)
  np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
  counts = rng.integers(1, 50, size=1000).astype(np.int32)
  ws = eng.weighted_sum(torch.tensor(counts, device="cuda"), torch.tensor(vals, device="cuda")).cpu().numpy()
  np.testing.assert_allclose(ws[:3] / ws[3], orc.weighted_average(counts, vals), rtol=1e-12)
  assert ws[3] == counts.sum()


def _kobe_desc(n, order, thetas):
  eng = _eng()
  from qhbmlib import _native as nat
  idx = orc.parity_indices(n, order)
  masks = np.array([sum(1 << (n - 1 - q) for q in c) for c in idx], dtype=np.int32)
  return eng.EnergyDescriptor(nat.ENERGY_KOBE, n, torch.tensor(masks, device="cuda"),
                              torch.tensor(np.asarray(thetas, dtype=np.float32), device="cuda"))


def test_kobe_golden_and_sweep_stats():
  """G4: KOBE [1.5, 2.7, -4.0] (energy_test.py:233-249, ebm_test.py:517-559)."""
  d = _kobe_desc(2, 2, [1.5, 2.7, -4.0])
  logits, stats = d.sweep(0, 4)
  np.testing.assert_allclose(-logits.cpu().numpy(), [0.2, 2.8, 5.2, -8.2], rtol=1e-6)
  m, s, t = stats.cpu().numpy()
  np.testing.assert_allclose(m + np.log(s), np.log(3641.8353), rtol=1e-6)
  np.testing.assert_allclose(m + np.log(s) - t / s, 0.00233551808, rtol=2e-4)


@pytest.mark.parametrize("n,order", [(12, 2), (16, 3), (5, 5)])
def test_kobe_sweep_against_oracle(n, order):
  rng = np.random.default_rng(n)
  thetas = rng.normal(0, 0.3, len(orc.parity_indices(n, order))).astype(np.float32)
  d = _kobe_desc(n, order, thetas)
  logits, stats = d.sweep(0, 1 << n)
  e_ref = orc.kobe_energy(orc.all_bitstrings(n), order, thetas)
  # fp32 sum of len(thetas) terms (the reference's tf.reduce_sum is fp32 too): floor 2e-6 * sum|theta|
  atol = 2e-6 * float(np.abs(thetas).sum()) + 1e-6
  np.testing.assert_allclose(-logits.cpu().numpy(), e_ref, rtol=1e-5, atol=atol)
  m, s, t = stats.cpu().numpy()
  np.testing.assert_allclose(m + np.log(s), orc.analytic_log_partition(e_ref), rtol=1e-6, atol=atol)
  np.testing.assert_allclose(m + np.log(s) - t / s, orc.analytic_entropy(e_ref), rtol=1e-5, atol=atol)
  # partial ranges merge like the multi-GPU split does
  half = 1 << (n - 1)
  _, s0 = d.sweep(0, half)
  _, s1 = d.sweep(half, 1 << n)
  (m0, a0, t0), (m1, a1, t1) = s0.cpu().numpy(), s1.cpu().numpy()
  mm = max(m0, m1)
  ss = a0 * np.exp(m0 - mm) + a1 * np.exp(m1 - mm)
  np.testing.assert_allclose(mm + np.log(ss), orc.analytic_log_partition(e_ref), rtol=1e-6, atol=atol)
  # explicit rows
  keys = torch.tensor(rng.integers(0, 1 << n, size=1000), device="cuda")
  er = d.energies(keys).cpu().numpy()
  np.testing.assert_allclose(er, e_ref[keys.cpu().numpy()], rtol=1e-5, atol=atol)


@pytest.mark.parametrize("n,order,lo,hi", [(14, 3, 1000, 13001), (13, 1, 77, 8000), (16, 2, 3 * 4096 + 5, 5 * 4096 + 3),
                                           (12, 12, 0, 4096)])
def test_parity_sweep_tiled_kernel_edge_cases(n, order, lo, hi):
  """The Walsh-Hadamard tiled sweep (>= 4096 rows) on ranges that do not align with its 256-row tiles,
  for single-Z (Bernoulli-like), order-3 and full-order term sets; repeated runs are bit-identical."""
  rng = np.random.default_rng(n * 10 + order)
  thetas = rng.normal(0, 0.3, len(orc.parity_indices(n, order))).astype(np.float32)
  if order == 12:
    thetas[rng.random(len(thetas)) < 0.9] = 0.0  # keep the fp32 sum of 4095 terms well conditioned
  d = _kobe_desc(n, order, thetas)
  logits, stats = d.sweep(lo, hi)
  e_ref = orc.kobe_energy(orc.all_bitstrings(n)[lo:hi], order, thetas)
  atol = 2e-6 * float(np.abs(thetas).sum()) + 1e-6
  np.testing.assert_allclose(-logits.cpu().numpy(), e_ref, rtol=1e-5, atol=atol)
  m, s, t = stats.cpu().numpy()
  np.testing.assert_allclose(m + np.log(s), orc.analytic_log_partition(e_ref), rtol=1e-6, atol=atol)
  np.testing.assert_allclose(m + np.log(s) - t / s, orc.analytic_entropy(e_ref), rtol=1e-5, atol=atol)
  logits2, stats2 = d.sweep(lo, hi)
  assert torch.equal(logits, logits2)
  # rows evaluated one by one (per-row kernel) agree with the tiled sweep
  keys = torch.arange(lo, min(hi, lo + 3000), device="cuda")
  np.testing.assert_allclose(d.energies(keys).cpu().numpy(), -logits[:len(keys)].cpu().numpy(), rtol=1e-5, atol=atol)


def test_mlp_energy_sweep_against_oracle():
  """Dense(64,tanh)->Dense(64,tanh)->Dense(1) on raw bits (ebm_utils_test.py:33-47 family)."""
  eng = _eng()
  from qhbmlib import _native as nat
  rng = np.random.default_rng(4)
  n = 14
  widths = [n, 64, 64, 1]
  acts = ["tanh", "tanh", "linear"]
  layers = []
  for l in range(3):
    lim = np.sqrt(6.0 / (widths[l] + widths[l + 1]))
    layers.append((rng.uniform(-lim, lim, (widths[l], widths[l + 1])).astype(np.float32),
                   rng.normal(0, 0.1, widths[l + 1]).astype(np.float32), acts[l]))
  d = eng.EnergyDescriptor(nat.ENERGY_MLP, n, layers=[(torch.tensor(w, device="cuda"),
                                                       torch.tensor(b, device="cuda"), a)
                                                      for w, b, a in layers])
  logits, stats = d.sweep(0, 1 << n)
  e_ref = orc.mlp_energy(orc.all_bitstrings(n), layers)
  np.testing.assert_allclose(-logits.cpu().numpy(), e_ref, rtol=1e-5, atol=2e-6)
  m, s, t = stats.cpu().numpy()
  np.testing.assert_allclose(m + np.log(s), orc.analytic_log_partition(e_ref), rtol=1e-6)


@pytest.mark.parametrize("n,widths,acts,lo,hi", [
    (14, [14, 64, 64, 1], ["tanh", "tanh", "linear"], 1000, 13001),       # ragged ends inside tiles
    (13, [13, 37, 5, 1], ["relu", "tanh", "linear"], 0, 1 << 13),         # widths that are not multiples of 8
    (13, [13, 20, 1], ["tanh", "linear"], 77, 8000),                      # two layers: bit-select + output layer
    (13, [13, 64, 64, 64, 1], ["tanh", "relu", "tanh", "tanh"], 64, 4200),  # three hidden layers, bounded output
    (16, [16, 8, 1], ["linear", "linear"], 3 * 4096 + 5, 5 * 4096 + 3),   # range away from row 0
])
def test_mlp_sweep_tiled_kernel_edge_cases(n, widths, acts, lo, hi):
  """The register-tiled sweep kernel (>= 4096 rows) on row ranges that do not align with its 128-row
  tiles and on layer shapes that exercise its output padding; statistics cover [lo, hi) only."""
  eng = _eng()
  from qhbmlib import _native as nat
  rng = np.random.default_rng(n * 100 + len(widths))
  layers = []
  for l in range(len(widths) - 1):
    lim = np.sqrt(6.0 / (widths[l] + widths[l + 1]))
    layers.append((rng.uniform(-lim, lim, (widths[l], widths[l + 1])).astype(np.float32),
                   rng.normal(0, 0.1, widths[l + 1]).astype(np.float32), acts[l]))
  d = eng.EnergyDescriptor(nat.ENERGY_MLP, n, layers=[(torch.tensor(w, device="cuda"),
                                                       torch.tensor(b, device="cuda"), a)
                                                      for w, b, a in layers])
  assert hi - lo >= 4096
  logits, stats = d.sweep(lo, hi)
  assert logits.shape == (hi - lo,)
  e_ref = orc.mlp_energy(orc.all_bitstrings(n)[lo:hi], layers)
  np.testing.assert_allclose(-logits.cpu().numpy(), e_ref, rtol=1e-5, atol=2e-6)
  m, s, t = stats.cpu().numpy()
  np.testing.assert_allclose(m + np.log(s), orc.analytic_log_partition(e_ref), rtol=1e-6)
  np.testing.assert_allclose(m + np.log(s) - t / s, orc.analytic_entropy(e_ref), rtol=1e-5)
  _, stats_only = d.sweep(lo, hi, want_logits=False)
  np.testing.assert_allclose(stats_only.cpu().numpy(), stats.cpu().numpy(), rtol=1e-12)


def test_categorical_sampling_distribution_and_seeding():
  eng = _eng()
  rng = np.random.default_rng(9)
  n_rows = 5000
  logits = rng.normal(0, 2.0, n_rows).astype(np.float32)
  d_logits = torch.tensor(logits, device="cuda")
  n_samples = 2_000_000
  s1 = eng.categorical_sample(d_logits, n_samples, (3, 4)).cpu().numpy()
  s2 = eng.categorical_sample(d_logits, n_samples, (3, 4)).cpu().numpy()
  s3 = eng.categorical_sample(d_logits, n_samples, (3, 5)).cpu().numpy()
  np.testing.assert_array_equal(s1, s2)           # same seed repeats exactly (ebm_test.py:280-297)
  assert (s1 != s3).mean() > 0.5                  # a different seed differs
  assert s1.min() >= 0 and s1.max() < n_rows
  p = orc.analytic_probabilities(-logits.astype(np.float64))
  counts = np.bincount(s1, minlength=n_rows)
  big = p * n_samples > 50
  z = (counts[big] - n_samples * p[big]) / np.sqrt(n_samples * p[big] * (1 - p[big]))
  assert np.abs(z).max() < 6.0
  assert abs(z.std() - 1.0) < 0.1
  # split streams: samples [0, N) == concat of [0, N/2) and [N/2, N)
  a = eng.categorical_sample(d_logits, 1000, (3, 4), first_sample=0).cpu().numpy()
  b = eng.categorical_sample(d_logits, 500, (3, 4), first_sample=500).cpu().numpy()
  np.testing.assert_array_equal(a[500:], b)
  # degenerate distribution
  one = torch.full((300,), -1e30, device="cuda")
  one[123] = 0.0
  assert set(eng.categorical_sample(one, 100, (1, 2)).cpu().numpy().tolist()) == {123}


def test_bernoulli_sampling_distribution_and_seeding():
  eng = _eng()
  n = 12
  thetas = np.linspace(-1.5, 1.5, n).astype(np.float32)
  logits = torch.tensor(2 * thetas, device="cuda")
  pi = orc.bit_column_to_qubit(n)
  shifts = [n - 1 - pi[j] for j in range(n)]
  n_samples = 1_000_000
  k1 = eng.bernoulli_sample(logits, shifts, n_samples, (5, 6))
  k2 = eng.bernoulli_sample(logits, shifts, n_samples, (5, 6))
  assert torch.equal(k1, k2)
  bits = eng.unpack_bits(k1, n, shifts).cpu().numpy()
  p1 = 1 / (1 + np.exp(-2 * thetas.astype(np.float64)))
  z = (bits.mean(0) - p1) / np.sqrt(p1 * (1 - p1) / n_samples)
  assert np.abs(z).max() < 5.0
  # independence between columns: correlation ~ 0
  c = np.corrcoef(bits[:200000].T.astype(np.float64))
  assert np.abs(c - np.eye(n)).max() < 0.02
