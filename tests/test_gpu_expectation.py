"""GPU parity tests of the state-vector path, through the C ABI (libqhbm_b200.so).

Oracle: oracle/qhbm_oracle.py (complex128).  Tolerance: north_star asks 1e-5 relative for
complex64 results; every comparison uses rtol 1e-5 with the absolute floor SURVEY section 7 item 6
prescribes for sums of cancelling Pauli terms, 1e-6 * sum|coeff| (times the upstream weights for
gradients).  `_check` prints the worst error / tolerance ratio it saw, so the margin is visible in
`pytest -s` output and in the per-round pytest log."""
import numpy as np
import pytest
import torch

from oracle import qhbm_oracle as orc
import helpers as hp

pytestmark = pytest.mark.gpu

RTOL = 1e-5
FLOOR = 1e-6  # x sum|coeff| (x upstream weight)
WORST = {"ratio": 0.0}


def _check(got, ref, floor, what):
  """|got - ref| <= RTOL |ref| + floor element-wise (floor broadcastable)."""
  got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
  ratio = np.abs(got - ref) / (RTOL * np.abs(ref) + floor)
  worst = float(ratio.max()) if ratio.size else 0.0
  WORST["ratio"] = max(WORST["ratio"], worst)
  assert worst <= 1.0, f"{what}: error / tolerance = {worst:.3g} (worst so far {WORST['ratio']:.3g})"


def _plan(gates, n, nsym, ops, grad=True, T=0, K=0):
  from qhbmlib import engine
  terms, offs = hp.ops_to_tables(ops, n)
  return engine.ExpectationPlan(gates, n, nsym, terms, offs, grad, T, K)


def _scale(ops):
  return np.array([sum(abs(c) for c, _ in op) for op in ops])


def _compare(gates, n, nsym, ops, rng, n_states, T=0, K=0, mode="exact", check_state=True):
  plan = _plan(gates, n, nsym, ops, True, T, K)
  phi = rng.uniform(-1, 1, max(nsym, 1)).astype(np.float32)[:nsym]
  basis = rng.choice(1 << n, size=min(n_states, 1 << n), replace=False).astype(np.int64)
  dg = rng.uniform(-1, 1, (len(basis), len(ops))).astype(np.float32)
  d_phi = torch.tensor(phi, device="cuda")
  d_basis = torch.tensor(basis, device="cuda")
  d_dg = torch.tensor(dg, device="cuda")
  e_ref, g_ref = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg, mode)
  scale = _scale(ops)  # sum|coeff| per observable
  # forward only
  e_fwd = plan.forward(d_basis, d_phi).cpu().numpy()
  _check(e_fwd, e_ref, FLOOR * scale[None, :], "forward expectations")
  # forward + adjoint, reduced gradient
  e, g = plan.forward_adjoint(d_basis, d_phi, d_dg, grad_mode=mode)
  _check(e.cpu().numpy(), e_ref, FLOOR * scale[None, :], "expectations of the adjoint call")
  # a gradient that vanishes analytically still carries complex64 rounding of O(|upstream| sum|coeff|)
  gfloor_state = FLOOR * (np.abs(dg) * scale[None, :]).sum(1)  # per state
  _check(g.cpu().numpy(), g_ref.sum(0), gfloor_state.sum(), "reduced gradient")
  # un-reduced gradient (the TFQ op's own output shape)
  _, gp = plan.forward_adjoint(d_basis, d_phi, d_dg, per_state=True, grad_mode=mode)
  _check(gp.cpu().numpy(), g_ref, gfloor_state[:, None], "per-state gradient")
  if check_state:
    st = plan.state(int(basis[0]), d_phi).cpu().numpy()
    np.testing.assert_allclose(st, orc.simulate(gates, n, phi, basis[0]), atol=3e-6)
  return plan


@pytest.mark.parametrize("n,layers,T,K", [(3, 2, 0, 4), (4, 2, 0, 5), (6, 3, 0, 4), (10, 2, 10, 5),
                                           (11, 2, 9, 4), (12, 2, 10, 5), (12, 3, 9, 4), (12, 2, 0, 0),
                                           (13, 2, 0, 0), (13, 2, 0, 5)])
def test_hea_against_oracle(n, layers, T, K):
  rng = np.random.default_rng(100 + n)
  gates, names = orc.hea_circuit(n, layers)
  ops = [orc.tfim_ring(n), orc.xxz_ring(n)] + orc.kobe_shards(n, 2)[:5]
  _compare(gates, n, len(names), ops, rng, 5, T, K)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("n,T,K", [(2, 0, 4), (5, 0, 5), (11, 9, 4), (11, 10, 5)])
def test_random_circuits_all_gate_types(seed, n, T, K):
  rng = np.random.default_rng(1000 * n + seed)
  gates = hp.random_circuit(n, 30, 6, rng)
  ops = hp.random_ops(n, 3, rng)
  _compare(gates, n, 6, ops, rng, 4, T, K)


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("n,T,K", [(3, 0, 4), (6, 0, 4), (10, 9, 4), (11, 10, 5), (13, 12, 4), (14, 12, 4)])
def test_single_observable_passes(seed, n, T, K):
  """One observable of 1- and 2-local X/Y strings with Z tails: in-tile flips run as observable passes
  (OP_HX / OP_HD, adjoint plans), odd-Y strings and cross-tile flips through the generic tables."""
  rng = np.random.default_rng(1000 * n + seed)
  nsym = 4
  gates = hp.random_circuit(n, 14, nsym, rng)
  terms = []
  for _ in range(int(rng.integers(3, 12))):
    paulis = {}
    for q in rng.choice(n, int(rng.integers(0, min(n, 2) + 1)), replace=False):
      paulis[int(q)] = str(rng.choice(["X", "Y"]))
    for q in rng.choice(n, int(rng.integers(0, min(n, 4) + 1)), replace=False):
      paulis.setdefault(int(q), "Z")
    terms.append((float(rng.uniform(-2, 2)), paulis))
  _compare(gates, n, nsym, [terms], rng, 4, T, K, check_state=False)


def test_observable_passes_match_generic_tables(monkeypatch):
  """Same plan compiled with and without observable passes (QHBM_NO_HPASS) on the headline shapes."""
  from qhbmlib import engine
  for n, ham in ((12, orc.xxz_ring), (14, orc.tfim_ring), (16, orc.xxz_ring)):
    rng = np.random.default_rng(n)
    gates, names = orc.hea_circuit(n, 2)
    terms, offs = hp.ops_to_tables([ham(n)], n)
    phi = torch.tensor(rng.uniform(-1, 1, len(names)).astype(np.float32), device="cuda")
    basis = torch.tensor(rng.choice(1 << n, 64, replace=False).astype(np.int64), device="cuda")
    dg = torch.tensor(rng.uniform(-1, 1, (64, 1)).astype(np.float32), device="cuda")
    monkeypatch.delenv("QHBM_NO_HPASS", raising=False)
    plan_h = engine.ExpectationPlan(gates, n, len(names), terms, offs, True)
    monkeypatch.setenv("QHBM_NO_HPASS", "1")
    plan_g = engine.ExpectationPlan(gates, n, len(names), terms, offs, True)
    assert plan_h.info["passes"] > plan_g.info["passes"]
    e_h, g_h = plan_h.forward_adjoint(basis, phi, dg)
    e_g, g_g = plan_g.forward_adjoint(basis, phi, dg)
    scale = float(sum(abs(c) for c, _ in ham(n)))
    np.testing.assert_allclose(e_h.cpu().numpy(), e_g.cpu().numpy(), rtol=RTOL, atol=RTOL * scale)
    np.testing.assert_allclose(g_h.cpu().numpy(), g_g.cpu().numpy(), rtol=RTOL,
                               atol=RTOL * float(g_g.abs().max()) * 3)


@pytest.mark.parametrize("n,T,K,ham", [(15, 0, 0, "tfim"), (15, 12, 5, "xxz"), (12, 10, 5, "tfim")])
def test_forward_only_expectation_stages_and_passes(n, T, K, ham, monkeypatch):
  """The optional forward-only paths (extra expectation stages with their own tile maps, observable
  passes in the K = 5 kernel) stay correct although they are off by default."""
  from qhbmlib import engine
  monkeypatch.setenv("QHBM_EXPECT_STAGES", "1")
  monkeypatch.setenv("QHBM_HPASS_FORWARD", "1")
  rng = np.random.default_rng(n)
  gates, names = orc.hea_circuit(n, 2)
  ops = [orc.tfim_ring(n) if ham == "tfim" else orc.xxz_ring(n)]
  terms, offs = hp.ops_to_tables(ops, n)
  plan = engine.ExpectationPlan(gates, n, len(names), terms, offs, False, T, K)
  assert plan.info["launches"] > plan.info["sweeps_fwd"] + 1
  phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
  basis = rng.choice(1 << n, 6, replace=False).astype(np.int64)
  e = plan.forward(torch.tensor(basis, device="cuda"), torch.tensor(phi, device="cuda")).cpu().numpy()
  e_ref = orc.expectations(gates, n, phi, basis, ops)
  _check(e, e_ref, FLOOR * _scale(ops)[None, :], "expectations")


@pytest.mark.parametrize("n,T,K", [(4, 0, 4), (11, 9, 4)])
def test_tfq_fd_mode(n, T, K):
  rng = np.random.default_rng(7)
  gates = hp.random_circuit(n, 25, 5, rng)
  ops = hp.random_ops(n, 2, rng)
  _compare(gates, n, 5, ops, rng, 3, T, K, mode="tfq_fd")


@pytest.mark.parametrize("n,T,K,grad", [(12, 0, 0, True), (12, 0, 5, True), (14, 12, 4, True), (13, 9, 4, True),
                                          (12, 0, 0, False), (15, 13, 5, False)])
def test_many_diagonal_shards_walsh_hadamard_path(n, T, K, grad):
  """>= 32 diagonal terms switch the expectation phase to the WHT evaluation: KOBE-2 Z-shards (the
  modular-Hamiltonian observables of hamiltonian.py:48-51) mixed with TFIM / XXZ sums."""
  rng = np.random.default_rng(40 + n)
  gates, names = orc.hea_circuit(n, 2)
  ops = orc.kobe_shards(n, 2) + [orc.tfim_ring(n), orc.xxz_ring(n), [(0.7, {}), (1.1, {0: "Z", n - 1: "Z", 3: "Z"})]]
  phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
  basis = rng.choice(1 << n, 4, replace=False).astype(np.int64)
  plan = _plan(gates, n, len(names), ops, grad, T, K)
  d_phi, d_basis = torch.tensor(phi, device="cuda"), torch.tensor(basis, device="cuda")
  scale = _scale(ops)[None, :]
  if not grad:
    e = plan.forward(d_basis, d_phi).cpu().numpy()
    e_ref = orc.expectations(gates, n, phi, basis, ops)
    _check(e, e_ref, FLOOR * scale, "WHT forward expectations")
    return
  dg = rng.uniform(-1, 1, (len(basis), len(ops))).astype(np.float32)
  e_ref, g_ref = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg)
  e, g = plan.forward_adjoint(d_basis, d_phi, torch.tensor(dg, device="cuda"), grad_mode="exact")
  _check(e.cpu().numpy(), e_ref, FLOOR * scale, "WHT expectations")
  _check(g.cpu().numpy(), g_ref.sum(0), FLOOR * (np.abs(dg) * scale).sum(), "WHT reduced gradient")
  e2 = plan.forward(d_basis, d_phi).cpu().numpy()   # forward-only call on an adjoint plan
  _check(e2, e_ref, FLOOR * scale, "WHT forward-only call on an adjoint plan")


def test_config1_4q_tfim_bernoulli_samples():
  """BASELINE config 1: 4-qubit TFIM, 2-layer HEA, 1k Bernoulli samples -> unique -> weighted mean."""
  rng = np.random.default_rng(5)
  n = 4
  gates, names = orc.hea_circuit(n, 2)
  phi = np.random.default_rng(11).uniform(-1, 1, len(names)).astype(np.float32)
  thetas = rng.uniform(-1, 1, n)
  p1 = 1 / (1 + np.exp(-2 * thetas))
  samples = (rng.random((1000, n)) < p1).astype(np.int8)
  y, idx, counts = orc.unique_bitstrings_with_counts(samples)
  ops = [orc.tfim_ring(n)]
  ref_avg, ref_vals = orc.qhbm_expectation(gates, n, phi, y, counts, ops)
  plan = _plan(gates, n, len(names), ops)
  basis = torch.tensor(orc.bitstrings_to_index(y), device="cuda")
  vals = plan.forward(basis, torch.tensor(phi, device="cuda")).cpu().numpy()
  np.testing.assert_allclose(vals, ref_vals, rtol=RTOL, atol=FLOOR * 8)
  np.testing.assert_allclose(orc.weighted_average(counts, vals), ref_avg, rtol=RTOL, atol=FLOOR * 8)


def test_config3_16q_xxz_adjoint_sample_against_oracle():
  """BASELINE config 3 (headline): 16-qubit XXZ, HEA L=2, adjoint gradient; oracle on a sample."""
  rng = np.random.default_rng(3)
  n = 16
  gates, names = orc.hea_circuit(n, 2)
  _compare(gates, n, len(names), [orc.xxz_ring(n)], rng, 3, check_state=True)


def test_config3_full_size_properties():
  """4096 unique 16-qubit bitstrings: size-independent checks (no oracle at this size):
  the identity observable gives exactly 1 (norm), expectations are linear in the
  observable, the gradient is linear in dgrad, and chunking does not change results."""
  rng = np.random.default_rng(33)
  n, u = 16, 4096
  gates, names = orc.hea_circuit(n, 2)
  h1, h2 = orc.xxz_ring(n), orc.tfim_ring(n)
  ops = [[(1.0, {})], h1, h2, h1 + h2]
  plan = _plan(gates, n, len(names), ops)
  phi = torch.tensor(rng.uniform(-1, 1, len(names)).astype(np.float32), device="cuda")
  basis = torch.tensor(rng.choice(1 << n, u, replace=False).astype(np.int64), device="cuda")
  dg = torch.tensor(rng.uniform(0, 1, (u, 4)).astype(np.float32), device="cuda")
  e, g = plan.forward_adjoint(basis, phi, dg)
  e = e.cpu().numpy().astype(np.float64)
  np.testing.assert_allclose(e[:, 0], 1.0, atol=2e-6)
  np.testing.assert_allclose(e[:, 3], e[:, 1] + e[:, 2], atol=2e-4)
  _, g2 = plan.forward_adjoint(basis, phi, 2 * dg)
  np.testing.assert_allclose(g2.cpu().numpy(), 2 * g.cpu().numpy(), rtol=1e-5, atol=1e-3)
  # first half + second half == whole
  _, ga = plan.forward_adjoint(basis[:u // 2], phi, dg[:u // 2])
  _, gb = plan.forward_adjoint(basis[u // 2:], phi, dg[u // 2:])
  np.testing.assert_allclose((ga + gb).cpu().numpy(), g.cpu().numpy(), rtol=1e-5,
                             atol=1e-5 * float(g.abs().max()))
  # identity observable has zero gradient
  dg0 = torch.zeros_like(dg)
  dg0[:, 0] = 1.0
  _, g0 = plan.forward_adjoint(basis, phi, dg0)
  assert float(g0.abs().max()) < 2e-3 * 1e-2 * u


def test_config4_20q_tfim_forward_sample_and_properties():
  """BASELINE config 4: 20-qubit TFIM expectation (forward).  Oracle on two bitstrings; for a larger batch
  the identity observable must give 1 and sharding the batch must not change any value."""
  rng = np.random.default_rng(4)
  n = 20
  gates, names = orc.hea_circuit(n, 2)
  ops = [orc.tfim_ring(n), [(1.0, {})]]
  plan = _plan(gates, n, len(names), ops, grad=False)
  assert plan.info["sweeps_fwd"] >= 2 and plan.info["tile_qubits"] == 13
  phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
  d_phi = torch.tensor(phi, device="cuda")
  basis = rng.choice(1 << n, 64, replace=False).astype(np.int64)
  e = plan.forward(torch.tensor(basis, device="cuda"), d_phi).cpu().numpy()
  e_ref = orc.expectations(gates, n, phi, basis[:2], ops)
  np.testing.assert_allclose(e[:2], e_ref, rtol=RTOL, atol=FLOOR * 2 * n)
  np.testing.assert_allclose(e[:, 1], 1.0, atol=3e-6)
  e_a = plan.forward(torch.tensor(basis[:23], device="cuda"), d_phi).cpu().numpy()
  e_b = plan.forward(torch.tensor(basis[23:], device="cuda"), d_phi).cpu().numpy()
  np.testing.assert_allclose(np.concatenate([e_a, e_b]), e, rtol=0, atol=1e-6)   # rows are independent


def test_24q_forward_is_schedule_independent():
  """Size-independent check at n = 24 (128 MiB per state, 2048+ tiles): two different tilings of the
  same circuit must agree, the identity observable must give 1, and <Z_0> must lie in [-1, 1]."""
  rng = np.random.default_rng(24)
  n = 24
  gates, names = orc.hea_circuit(n, 2)
  ops = [orc.tfim_ring(n), [(1.0, {})], [(1.0, {0: "Z"})]]
  phi = torch.tensor(rng.uniform(-1, 1, len(names)).astype(np.float32), device="cuda")
  basis = torch.tensor(rng.choice(1 << n, 3, replace=False).astype(np.int64), device="cuda")
  e1 = _plan(gates, n, len(names), ops, False, 13, 5).forward(basis, phi).cpu().numpy()
  e2 = _plan(gates, n, len(names), ops, False, 12, 4).forward(basis, phi).cpu().numpy()
  np.testing.assert_allclose(e1, e2, rtol=1e-5, atol=1e-5 * 2 * n)
  np.testing.assert_allclose(e1[:, 1], 1.0, atol=5e-6)
  assert np.all(np.abs(e1[:, 2]) <= 1.0 + 1e-5)


def test_adjoint_is_schedule_independent_16q():
  """Same for the gradient at n = 16: tile 2^12/K=4, 2^13/K=4 and 2^13/K=5 give the same numbers."""
  rng = np.random.default_rng(16)
  n = 16
  gates, names = orc.hea_circuit(n, 3)
  ops = [orc.xxz_ring(n)] + orc.kobe_shards(n, 2)[:3]
  phi = torch.tensor(rng.uniform(-1, 1, len(names)).astype(np.float32), device="cuda")
  basis = torch.tensor(rng.choice(1 << n, 37, replace=False).astype(np.int64), device="cuda")
  dg = torch.tensor(rng.uniform(-1, 1, (37, len(ops))).astype(np.float32), device="cuda")
  outs = [_plan(gates, n, len(names), ops, True, T, K).forward_adjoint(basis, phi, dg) for T, K in
          [(12, 4), (13, 4), (13, 5)]]
  for e, g in outs[1:]:
    np.testing.assert_allclose(e.cpu().numpy(), outs[0][0].cpu().numpy(), rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(g.cpu().numpy(), outs[0][1].cpu().numpy(), rtol=1e-5,
                               atol=1e-5 * float(outs[0][1].abs().max()))


def test_18q_adjoint_sample_against_oracle():
  """Multi-tile adjoint with 64 tiles per state (n = 18, tile 2^12)."""
  rng = np.random.default_rng(18)
  n = 18
  gates, names = orc.hea_circuit(n, 2)
  _compare(gates, n, len(names), [orc.xxz_ring(n), orc.tfim_ring(n)], rng, 2, check_state=False)


def test_trace_property_all_basis_states():
  """sum over ALL basis states of <x|U^dag H U|x> = Tr H = 0 for a traceless H (n=12, 4096 rows)."""
  n = 12
  gates, names = orc.hea_circuit(n, 2)
  ops = [orc.xxz_ring(n), [(1.0, {})]]
  plan = _plan(gates, n, len(names), ops, grad=False)
  rng = np.random.default_rng(8)
  phi = torch.tensor(rng.uniform(-1, 1, len(names)).astype(np.float32), device="cuda")
  basis = torch.arange(1 << n, device="cuda", dtype=torch.int64)
  e = plan.forward(basis, phi).double()
  assert abs(float(e[:, 0].sum())) < 1e-2
  np.testing.assert_allclose(e[:, 1].cpu().numpy(), 1.0, atol=2e-6)


def test_host_buffer_entry_point():
  rng = np.random.default_rng(12)
  n = 6
  gates, names = orc.hea_circuit(n, 2)
  ops = [orc.tfim_ring(n), orc.xxz_ring(n)]
  plan = _plan(gates, n, len(names), ops)
  phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
  basis = rng.choice(1 << n, 7, replace=False).astype(np.uint64)
  dg = rng.uniform(-1, 1, (7, 2)).astype(np.float32)
  e, g = plan.run_host(basis, phi, dg)
  e_ref, g_ref = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg)
  scale = _scale(ops)[None, :]
  _check(e, e_ref, FLOOR * scale, "host-buffer expectations")
  _check(g, g_ref.sum(0), FLOOR * (np.abs(dg) * scale).sum(), "host-buffer reduced gradient")
  e2, g2 = plan.run_host(basis, phi)
  assert g2 is None
  _check(e2, e_ref, FLOOR * scale, "host-buffer forward expectations")


def test_errors_are_reported():
  from qhbmlib import _native as nat
  from qhbmlib import engine
  gates, names = orc.hea_circuit(3, 1)
  terms, offs = hp.ops_to_tables([orc.tfim_ring(3)], 3)
  bad = gates.copy()
  bad["q0"][0] = 7
  with pytest.raises(nat.NativeError, match="q0 out of range"):
    engine.ExpectationPlan(bad, 3, len(names), terms, offs)
  plan = engine.ExpectationPlan(gates, 3, len(names), terms, offs, with_gradient=False)
  with pytest.raises(nat.NativeError, match="without with_gradient"):
    plan.forward_adjoint(torch.zeros(1, dtype=torch.int64, device="cuda"),
                         torch.zeros(len(names), device="cuda"),
                         torch.zeros((1, 1), device="cuda"))
  with pytest.raises(TypeError):
    plan.forward(torch.zeros(1, dtype=torch.int64), torch.zeros(len(names)))


def test_empty_batch_and_empty_circuit():
  n = 5
  gates = np.zeros(0, dtype=orc.GATE_DTYPE)
  ops = [[(1.5, {}), (0.5, {0: "Z"})], [(2.0, {n - 1: "X"})]]
  plan = _plan(gates, n, 0, ops)
  basis = torch.tensor([0, 16, 31], device="cuda", dtype=torch.int64)
  e = plan.forward(basis, torch.zeros(0, device="cuda")).cpu().numpy()
  np.testing.assert_allclose(e, [[2.0, 0.0], [1.0, 0.0], [1.0, 0.0]], atol=1e-6)
  e0 = plan.forward(basis[:0], torch.zeros(0, device="cuda"))
  assert tuple(e0.shape) == (0, 2)


def test_plan_lifecycle_does_not_leak_device_memory():
  """Creating, running and destroying many plans returns the device memory they took (handles own
  their workspace; nothing is cached per call)."""
  import gc
  from qhbmlib import engine
  n = 13
  gates, names = orc.hea_circuit(n, 2)
  terms, offs = hp.ops_to_tables([orc.tfim_ring(n)], n)
  phi = torch.zeros(len(names), device="cuda")
  basis = torch.arange(64, dtype=torch.int64, device="cuda")
  dg = torch.ones((64, 1), device="cuda")

  def cycle():
    plan = engine.ExpectationPlan(gates, n, len(names), terms, offs, True)
    plan.forward_adjoint(basis, phi, dg)
    plan.final_states(basis[:4], phi)
    del plan

  cycle()
  gc.collect()
  torch.cuda.synchronize()
  free0, _ = torch.cuda.mem_get_info()
  for _ in range(40):
    cycle()
  gc.collect()
  torch.cuda.synchronize()
  free1, _ = torch.cuda.mem_get_info()
  assert free0 - free1 < 8 << 20, (free0, free1)


@pytest.mark.parametrize("n,ham", [(14, "xxz"), (13, "tfim")])
def test_adjoint_in_small_chunks_matches_one_chunk(n, ham, monkeypatch):
  """Multi-tile adjoint plan (expectation fused with the first backward sweep, psi ping-pong) when
  the batch is split into several workspace chunks: same values and gradients as one chunk."""
  from qhbmlib import engine
  rng = np.random.default_rng(n)
  gates, names = orc.hea_circuit(n, 2)
  ops = [orc.xxz_ring(n) if ham == "xxz" else orc.tfim_ring(n)]
  terms, offs = hp.ops_to_tables(ops, n)
  phi = torch.tensor(rng.uniform(-1, 1, len(names)).astype(np.float32), device="cuda")
  basis = torch.tensor(rng.choice(1 << n, 11, replace=False).astype(np.int64), device="cuda")
  dg = torch.tensor(rng.uniform(-1, 1, (11, 1)).astype(np.float32), device="cuda")
  monkeypatch.delenv("QHBM_CHUNK", raising=False)
  whole = engine.ExpectationPlan(gates, n, len(names), terms, offs, True, 12, 4)
  assert whole.info["chunk"] >= 11
  monkeypatch.setenv("QHBM_CHUNK", "3")
  split = engine.ExpectationPlan(gates, n, len(names), terms, offs, True, 12, 4)
  assert split.info["chunk"] == 3
  e0, g0 = whole.forward_adjoint(basis, phi, dg, per_state=True)
  e1, g1 = split.forward_adjoint(basis, phi, dg, per_state=True)
  assert torch.equal(e0, e1) and torch.equal(g0, g1)  # per-state results do not depend on the chunking
  _, gr0 = whole.forward_adjoint(basis, phi, dg)
  _, gr1 = split.forward_adjoint(basis, phi, dg)
  np.testing.assert_allclose(gr1.cpu().numpy(), gr0.cpu().numpy(), rtol=1e-5, atol=1e-6)
  e_ref, g_ref = orc.batch_expectation_and_gradient(gates, n, phi.cpu().numpy(), basis.cpu().numpy(), ops,
                                                    dg.cpu().numpy())
  _check(e1.cpu().numpy(), e_ref, FLOOR * _scale(ops)[None, :], "chunked expectations")
  _check(g1.cpu().numpy(), g_ref, FLOOR * np.abs(dg.cpu().numpy()) * _scale(ops).max(), "chunked per-state gradient")


# ------------------------------------------------------------------ one row of symbol values per state
def _compare_rows(gates, n, nsym, ops, rng, n_states, T=0, K=0, mode="exact"):
  """symbols f32[U, P] (tfq_simulate_expectation / tfq_adjoint_gradient's general form): state u is simulated
  with row u; oracle = one single-state call per row."""
  plan = _plan(gates, n, nsym, ops, True, T, K)
  basis = rng.choice(1 << n, size=min(n_states, 1 << n), replace=False).astype(np.int64)
  u = len(basis)
  phi = rng.uniform(-1, 1, (u, nsym)).astype(np.float32)
  dg = rng.uniform(-1, 1, (u, len(ops))).astype(np.float32)
  e_ref = np.zeros((u, len(ops)))
  g_ref = np.zeros((u, nsym))
  for i in range(u):
    e1, g1 = orc.batch_expectation_and_gradient(gates, n, phi[i], basis[i:i + 1], ops, dg[i:i + 1], mode)
    e_ref[i], g_ref[i] = e1[0], g1[0]
  d_phi, d_basis, d_dg = (torch.tensor(a, device="cuda") for a in (phi, basis, dg))
  scale = _scale(ops)
  gfloor_state = FLOOR * (np.abs(dg) * scale[None, :]).sum(1)
  _check(plan.forward(d_basis, d_phi).cpu().numpy(), e_ref, FLOOR * scale[None, :], "forward, symbol rows")
  e, gp = plan.forward_adjoint(d_basis, d_phi, d_dg, per_state=True, grad_mode=mode)
  _check(e.cpu().numpy(), e_ref, FLOOR * scale[None, :], "expectations, symbol rows")
  _check(gp.cpu().numpy(), g_ref, gfloor_state[:, None], "per-state gradient, symbol rows")
  _, g = plan.forward_adjoint(d_basis, d_phi, d_dg, grad_mode=mode)
  _check(g.cpu().numpy(), g_ref.sum(0), gfloor_state.sum(), "reduced gradient, symbol rows")
  return plan, d_basis, d_phi, d_dg


@pytest.mark.parametrize("n,layers,T,K,mode", [(4, 2, 0, 4, "exact"), (10, 2, 10, 5, "tfq_fd"), (12, 2, 9, 4, "exact"),
                                                (13, 2, 0, 0, "exact"), (16, 2, 0, 0, "tfq_fd")])
def test_symbol_rows_hea(n, layers, T, K, mode):
  rng = np.random.default_rng(300 + n)
  gates, names = orc.hea_circuit(n, layers)
  ops = [orc.xxz_ring(n), orc.tfim_ring(n)] if n < 16 else [orc.xxz_ring(n)]
  _compare_rows(gates, n, len(names), ops, rng, 6 if n < 16 else 3, T, K, mode)


@pytest.mark.parametrize("seed", range(3))
@pytest.mark.parametrize("n,T,K", [(5, 0, 5), (11, 9, 4), (13, 12, 4)])
def test_symbol_rows_all_gate_types(seed, n, T, K):
  rng = np.random.default_rng(7000 * n + seed)
  gates = hp.random_circuit(n, 24, 5, rng)
  _compare_rows(gates, n, 5, hp.random_ops(n, 2, rng), rng, 5, T, K)


def test_symbol_rows_equal_rows_match_shared_row(monkeypatch):
  """U identical rows give the results of the shared-row call (same tables, same sweeps), also when
  the call is split into several chunks of tables (QHBM_CHUNK)."""
  from qhbmlib import engine
  n = 14
  rng = np.random.default_rng(14)
  gates, names = orc.hea_circuit(n, 2)
  terms, offs = hp.ops_to_tables([orc.xxz_ring(n)], n)
  phi = torch.tensor(rng.uniform(-1, 1, len(names)).astype(np.float32), device="cuda")
  basis = torch.tensor(rng.choice(1 << n, 37, replace=False).astype(np.int64), device="cuda")
  dg = torch.tensor(rng.uniform(-1, 1, (37, 1)).astype(np.float32), device="cuda")
  rows = phi[None, :].repeat(37, 1).contiguous()
  for chunk in (None, "8"):
    if chunk:
      monkeypatch.setenv("QHBM_CHUNK", chunk)
    plan = engine.ExpectationPlan(gates, n, len(names), terms, offs, True)
    def same(x, y):  # (float64 atomics of the tiles of a state arrive in any order: last-bit differences)
      np.testing.assert_allclose(x.cpu().numpy(), y.cpu().numpy(), rtol=2e-6, atol=2e-7)
    e0, g0 = plan.forward_adjoint(basis, phi, dg, per_state=True)
    e1, g1 = plan.forward_adjoint(basis, rows, dg, per_state=True)
    same(e0, e1), same(g0, g1)
    same(plan.forward(basis, phi), plan.forward(basis, rows))
    # and the shared-row path still works on the plan after its table buffer grew
    e2, g2 = plan.forward_adjoint(basis, phi, dg, per_state=True)
    same(e0, e2), same(g0, g2)
  with pytest.raises(ValueError, match="per-state symbols"):
    plan.forward(basis, rows[:5])
