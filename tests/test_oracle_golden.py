"""Pins oracle/qhbm_oracle.py against the reference's own known-answer tests.

Each test cites the reference test (file:line under /root/reference) whose closed
form or golden vector it reproduces.  Closed forms
are exact, so they are checked far tighter than the reference's tolerances.
"""
import itertools
import math

import numpy as np
import pytest

from oracle import qhbm_oracle as orc


def _gates(rows):
  g = np.zeros(len(rows), dtype=orc.GATE_DTYPE)
  for i, r in enumerate(rows):
    g[i] = r
  return g


# ---- G1: X^p|s>  (tests/inference/qnn_test.py:83-180) ----------------------
@pytest.mark.parametrize("mode,tol", [("exact", 1e-9), ("tfq_fd", 1e-4)])
def test_g1_xpow_closed_form(mode, tol):
  n, p = 3, 0.3718
  gates = _gates([orc._gate(orc.GATE_XPOW, q, sym=(0, -1, -1), scalar=(1, 0, 0))
                  for q in range(n)])
  for bits in itertools.product([0, 1], repeat=n):
    idx = orc.bitstrings_to_index([bits])[0]
    for pauli, val, dval in [
        ("X", lambda s: 0.0, lambda s: 0.0),
        ("Y", lambda s: -((-1.0)**s) * math.sin(math.pi * p),
         lambda s: -((-1.0)**s) * math.pi * math.cos(math.pi * p)),
        ("Z", lambda s: ((-1.0)**s) * math.cos(math.pi * p),
         lambda s: -((-1.0)**s) * math.pi * math.sin(math.pi * p)),
    ]:
      ops = [[(1.0, {q: pauli})] for q in range(n)]
      for j in range(n):
        dg = np.zeros(n)
        dg[j] = 1.0
        e, g = orc.adjoint_gradient(gates, n, [p], idx, ops, dg, mode)
        np.testing.assert_allclose(e, [val(s) for s in bits], atol=1e-12)
        np.testing.assert_allclose(g[0], dval(bits[j]), atol=tol * math.pi)


def test_tfq_fd_is_sinc_scaled_exact():
  """for gap-1, shift-0 eigen-gates the FD derivative equals the
  exact one times sin(pi eps)/(pi eps)."""
  gates, names = orc.hea_circuit(3, 2)
  rng = np.random.default_rng(0)
  phi = rng.uniform(-1, 1, len(names))
  ops = [orc.tfim_ring(3)]
  _, ge = orc.adjoint_gradient(gates, 3, phi, 5, ops, [1.0], "exact")
  _, gf = orc.adjoint_gradient(gates, 3, phi, 5, ops, [1.0], "tfq_fd")
  f = math.sin(math.pi * orc.TFQ_GRAD_EPS) / (math.pi * orc.TFQ_GRAD_EPS)
  np.testing.assert_allclose(gf, ge * f, rtol=1e-9, atol=1e-12)
  assert abs(1 - f - 4.11e-5) < 1e-7


def test_adjoint_matches_numeric_stencil_all_gate_types():
  """Same role as tests/test_util.py:210-309 (five-point stencil oracle)."""
  rng = np.random.default_rng(1)
  n = 4
  rows = [
      orc._gate(orc.GATE_HPOW, 0, sym=(0, -1, -1), scalar=(0.7, 0, 0), gshift=-0.5),
      orc._gate(orc.GATE_YPOW, 1, sym=(1, -1, -1), scalar=(1, 0, 0), cnst=(0.2, 0, 0)),
      orc._gate(orc.GATE_XPOW, 2, sym=(2, -1, -1), scalar=(1 / math.pi, 0, 0), gshift=-0.5),
      orc._gate(orc.GATE_ZPOW, 3, sym=(3, -1, -1), scalar=(1, 0, 0)),
      orc._gate(orc.GATE_CNOTPOW, 0, 1, sym=(4, -1, -1), scalar=(1, 0, 0)),
      orc._gate(orc.GATE_SWAPPOW, 1, 2, sym=(5, -1, -1), scalar=(1, 0, 0)),
      orc._gate(orc.GATE_ISWAPPOW, 2, 3, sym=(6, -1, -1), scalar=(1, 0, 0)),
      orc._gate(orc.GATE_XXPOW, 3, 0, sym=(7, -1, -1), scalar=(1, 0, 0), gshift=-0.5),
      orc._gate(orc.GATE_YYPOW, 0, 2, sym=(8, -1, -1), scalar=(-1, 0, 0)),
      orc._gate(orc.GATE_ZZPOW, 1, 3, sym=(9, -1, -1), scalar=(1, 0, 0)),
      orc._gate(orc.GATE_CZPOW, 2, 1, sym=(10, -1, -1), scalar=(1, 0, 0)),
      orc._gate(orc.GATE_PHASEDXPOW, 0, sym=(11, 12, -1), scalar=(1, 1, 0), nparams=2),
      orc._gate(orc.GATE_FSIM, 1, 2, sym=(13, 14, -1), scalar=(1, 1, 0), nparams=2),
      orc._gate(orc.GATE_PHASEDISWAPPOW, 3, 2, sym=(15, 0, -1), scalar=(1, 0.5, 0),
                nparams=2),
  ]
  gates = _gates(rows)
  phi = rng.uniform(-1, 1, 16)
  ops = [orc.xxz_ring(n), orc.tfim_ring(n), [(0.3, {0: "Y", 2: "X"}), (1.5, {})]]
  dg = np.array([0.7, -1.3, 2.0])
  for idx in (0, 6, 15):
    _, g = orc.adjoint_gradient(gates, n, phi, idx, ops, dg, "exact")
    gn = orc.numeric_gradient(gates, n, phi, idx, ops, dg)
    np.testing.assert_allclose(g, gn, atol=1e-8)


def test_gate_matrices_unitary_and_known_values():
  rng = np.random.default_rng(2)
  for t in range(15):
    p = rng.uniform(-1, 1, 3)
    m = orc.gate_matrix(t, p, rng.uniform(-1, 1))
    np.testing.assert_allclose(m.conj().T @ m, np.eye(m.shape[0]), atol=1e-12)
  np.testing.assert_allclose(orc.gate_matrix(orc.GATE_XPOW, [1.0]), orc._X, atol=1e-12)
  np.testing.assert_allclose(orc.gate_matrix(orc.GATE_HPOW, [1.0]), orc._H, atol=1e-12)
  np.testing.assert_allclose(orc.gate_matrix(orc.GATE_CZPOW, [1.0]),
                             np.diag([1, 1, 1, -1]), atol=1e-12)
  np.testing.assert_allclose(
      orc.gate_matrix(orc.GATE_CNOTPOW, [1.0]),
      [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], atol=1e-12)
  np.testing.assert_allclose(orc.gate_matrix(orc.GATE_SWAPPOW, [1.0]), orc._SWAP, atol=1e-12)
  np.testing.assert_allclose(
      orc.gate_matrix(orc.GATE_ISWAPPOW, [1.0]),
      [[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], atol=1e-12)
  th = 0.83  # rx(theta) = XPow(theta/pi, shift=-1/2), vqt_loss_test.py:156-157
  rx = orc.gate_matrix(orc.GATE_XPOW, [th / math.pi], -0.5)
  np.testing.assert_allclose(
      rx, [[math.cos(th / 2), -1j * math.sin(th / 2)],
           [-1j * math.sin(th / 2), math.cos(th / 2)]], atol=1e-12)


# ---- G2: rx + Bernoulli VQT  (tests/inference/vqt_loss_test.py:132-205) ----
def test_g2_vqt_rx_bernoulli_closed_form():
  rng = np.random.default_rng(3)
  for n in (1, 2, 3, 4):
    thetas = rng.uniform(-2, 2, n)
    phis = rng.uniform(-1, 1, n)
    beta = 1.7
    gates = _gates([orc._gate(orc.GATE_XPOW, q, sym=(q, -1, -1),
                              scalar=(1 / math.pi, 0, 0), gshift=-0.5) for q in range(n)])
    ops = [[(1.0, {q: "Y"}) for q in range(n)]]
    bits = orc.all_bitstrings(n)
    probs = orc.analytic_probabilities(orc.bernoulli_energy(bits, thetas))
    idx = orc.bitstrings_to_index(bits)
    e, g = orc.batch_expectation_and_gradient(
        gates, n, phis, idx, ops, beta * probs[:, None], "exact")
    expectation = float(probs @ e[:, 0])
    np.testing.assert_allclose(expectation, np.sum(np.tanh(thetas) * np.sin(phis)),
                               atol=2e-7)
    entropy = orc.bernoulli_entropy(thetas)
    np.testing.assert_allclose(
        entropy, np.sum(-thetas * np.tanh(thetas) + np.log(2 * np.cosh(thetas))),
        atol=1e-12)
    np.testing.assert_allclose(orc.analytic_entropy(orc.bernoulli_energy(bits, thetas)),
                               entropy, atol=1e-12)
    # d loss / d phi = beta tanh(theta) cos(phi)
    # (the 1/pi exponent scalar is stored as float32 in the gate table, as in TFQ's proto)
    np.testing.assert_allclose(g.sum(0), beta * np.tanh(thetas) * np.cos(phis), atol=2e-7)
    # d loss / d theta via the score-function formula of ebm.py:282-325 with exact
    # probabilities standing in for counts: f = beta <H> - E (E stop-gradient).
    energies = orc.bernoulli_energy(bits, thetas)
    f = beta * e[:, 0] - energies
    jac = orc.spins_from_bitstrings(bits)
    gth = orc.expectation_score_gradient(probs, f, jac, np.zeros(n), 1.0)
    np.testing.assert_allclose(
        gth, (1 - np.tanh(thetas)**2) * (beta * np.sin(phis) + thetas), atol=2e-7)


# ---- G3: QMHL rx / ry  (tests/inference/qmhl_loss_test.py:136-272) ---------
def test_g3_qmhl_rx_ry_closed_form():
  rng = np.random.default_rng(4)
  for n in (1, 2, 3):
    thetas = rng.uniform(0.25, 1.0, n)
    phis = rng.uniform(math.pi / 4, math.pi, n)
    alphas = rng.uniform(-math.pi, math.pi, n)
    data_probs = rng.uniform(0, 1, n)  # prob of bit = 0 is data_probs (samples ~ Bernoulli(1-p))
    ry = [orc._gate(orc.GATE_YPOW, q, cnst=(alphas[q] / math.pi, 0, 0), gshift=-0.5)
          for q in range(n)]
    rx = _gates([orc._gate(orc.GATE_XPOW, q, sym=(q, -1, -1), scalar=(1 / math.pi, 0, 0),
                           gshift=-0.5) for q in range(n)])
    total = orc.concat_circuits(_gates(ry), 0, orc.inverse_circuit(rx))
    bits = orc.all_bitstrings(n)
    pb = np.prod(np.where(bits == 1, 1 - data_probs, data_probs), axis=1)
    val, per_row = orc.modular_hamiltonian_expectation(
        total, n, phis, bits, pb, orc.bernoulli_shards(n), thetas)
    expected = np.sum(thetas * (2 * data_probs - 1) * np.cos(alphas) * np.cos(phis))
    np.testing.assert_allclose(val, expected, atol=2e-7)
    np.testing.assert_allclose(orc.bernoulli_log_partition(thetas),
                               np.sum(np.log(2 * np.cosh(thetas))), atol=1e-12)
    # phi gradient through the dagger circuit
    idx = orc.bitstrings_to_index(bits)
    dg = pb[:, None] * thetas[None, :]
    _, g = orc.batch_expectation_and_gradient(total, n, phis, idx,
                                              orc.bernoulli_shards(n), dg, "exact")
    np.testing.assert_allclose(
        g.sum(0), -thetas * (2 * data_probs - 1) * np.cos(alphas) * np.sin(phis), atol=2e-7)
    # log-partition gradient: -E[dE/dtheta] = tanh(theta)
    mp = orc.analytic_probabilities(orc.bernoulli_energy(bits, thetas))
    np.testing.assert_allclose(
        orc.log_partition_gradient(mp, orc.spins_from_bitstrings(bits)), np.tanh(thetas),
        atol=1e-12)


# ---- G4: KOBE constants (energy_test.py:233-249, ebm_test.py:517-559) ------
def test_g4_kobe_constants():
  thetas = [1.5, 2.7, -4.0]
  bits = [[0, 0], [0, 1], [1, 0], [1, 1]]
  e = orc.kobe_energy(bits, 2, thetas)
  np.testing.assert_allclose(e, [0.2, 2.8, 5.2, -8.2], atol=1e-12)
  np.testing.assert_allclose(orc.analytic_log_partition(e), math.log(3641.8353), rtol=1e-7)
  np.testing.assert_allclose(orc.analytic_entropy(e), 0.00233551808, rtol=5e-5)  # reference constant is float32


# ---- G5: Bernoulli energies / Jacobian / entropy ---------------------------
def test_g5_bernoulli():
  v = np.array([1.0, 1.7, -2.8])  # energy_test.py:121-145
  b = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 1]])
  np.testing.assert_allclose(orc.bernoulli_energy(b, v),
                             [v[0] + v[1] + v[2], -v[0] + v[1] + v[2], v[0] - v[1] - v[2]])
  np.testing.assert_array_equal(orc.spins_from_bitstrings(b), 1 - 2 * b)
  th = np.array([-1.5, 0.6, 2.1])  # ebm_test.py:736-762
  p = np.exp(2 * th) / (1 + np.exp(2 * th))
  allp = [np.prod([p[i] if bit else 1 - p[i] for i, bit in enumerate(bits)])
          for bits in itertools.product([0, 1], repeat=3)]
  np.testing.assert_allclose(orc.bernoulli_entropy(th), -np.sum(allp * np.log(allp)),
                             rtol=1e-12)
  np.testing.assert_allclose(
      orc.analytic_probabilities(orc.bernoulli_energy(orc.all_bitstrings(3), th)), allp,
      rtol=1e-12)


# ---- G6: Parity (energy_utils_test.py:86-110) ------------------------------
def test_g6_parity():
  assert orc.parity_indices(4, 3) == [
      (0,), (1,), (2,), (3,), (0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3), (0, 1, 2),
      (0, 1, 3), (0, 2, 3), (1, 2, 3)]
  bits = [[1, 0, 1, 1]]  # spins [-1, 1, -1, -1]
  np.testing.assert_array_equal(
      orc.parity_features(bits, 3),
      [[-1, 1, -1, -1] + [-1, 1, 1, -1, -1, 1] + [1, 1, -1, 1]])


# ---- G7: unique / counts / weighted average (utils_test.py:48-186) ---------
def test_g7_unique_and_weighted_average():
  b = np.array([[1, 0, 1], [1, 1, 1], [0, 1, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1],
                [1, 0, 1], [1, 0, 1]], dtype=np.int8)
  y, idx, c = orc.unique_bitstrings_with_counts(b)
  np.testing.assert_array_equal(y, [[1, 0, 1], [1, 1, 1], [0, 1, 1]])
  np.testing.assert_array_equal(idx, [0, 1, 2, 0, 1, 2, 0, 0])
  np.testing.assert_array_equal(c, [4, 2, 2])
  np.testing.assert_array_equal(orc.expand_unique_results(y, idx), b)
  b1 = np.array([[1], [0], [0], [1], [1], [0], [1], [1]], dtype=np.int8)
  y, idx, c = orc.unique_bitstrings_with_counts(b1)
  np.testing.assert_array_equal(y, [[1], [0]])
  np.testing.assert_array_equal(idx, [0, 1, 1, 0, 0, 1, 0, 0])
  np.testing.assert_array_equal(c, [5, 3])
  counts = np.array([37, 5])  # utils_test.py:48-73 structure
  vals = np.array([[2.7, -5.9], [0.5, 3.0]])
  np.testing.assert_allclose(orc.weighted_average(counts, vals),
                             (37 * vals[0] + 5 * vals[1]) / 42)


# ---- G8: all_bitstrings order (ebm_test.py:183-186) ------------------------
def test_g8_all_bitstrings_order():
  np.testing.assert_array_equal(
      orc.all_bitstrings(3),
      [[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0],
       [1, 1, 1]])
  np.testing.assert_array_equal(orc.all_bitstrings(5),
                                list(itertools.product([0, 1], repeat=5)))


# ---- G9: all-ones energy (ebm_test.py:300-453) -----------------------------
def test_g9_all_ones_energy_score_gradient():
  """E(x) = -theta if x == 1..1 else 0 wait-free restatement: p* = e^theta/(2^n-1+e^theta);
  f(x) = mu [x == 1..1]:  E[f] = mu p*, d/dtheta = mu p*(1-p*) with E = -theta 1[x=ones].
  The reference (ebm_test.py:300-453) states it with the opposite energy sign as
  d/dtheta = mu p*(p*-1); both are the same formula E[c]E[dE]-E[c dE]."""
  n, theta, mu = 3, 0.8, 1.9
  bits = orc.all_bitstrings(n)
  ones = (bits.sum(1) == n).astype(np.float64)
  for sign in (+1.0, -1.0):
    energies = sign * theta * ones
    p = orc.analytic_probabilities(energies)
    pstar = p[-1]
    f = mu * ones
    jac = (sign * ones)[:, None]
    g = orc.expectation_score_gradient(p, f, jac, np.zeros(1), 1.0)
    np.testing.assert_allclose(float(p @ f), mu * pstar, atol=1e-12)
    np.testing.assert_allclose(g[0], sign * mu * pstar * (pstar - 1), atol=1e-12)


# ---- G11: QHBM.expectation == weighted average of per-state expectations ---
def test_g11_qhbm_expectation_self_consistency():
  """tests/inference/qhbm_test.py:150-203."""
  n = 3
  gates, names = orc.hea_circuit(n, 2)
  rng = np.random.default_rng(5)
  phi = rng.uniform(-1, 1, len(names))
  samples = rng.integers(0, 2, size=(200, n)).astype(np.int8)
  y, idx, c = orc.unique_bitstrings_with_counts(samples)
  ops = [orc.tfim_ring(n), orc.xxz_ring(n)]
  avg, vals = orc.qhbm_expectation(gates, n, phi, y, c, ops)
  full = orc.expectations(gates, n, phi, orc.bitstrings_to_index(samples), ops)
  np.testing.assert_allclose(avg, full.mean(0), atol=1e-12)
  np.testing.assert_allclose(orc.expand_unique_results(vals, idx), full, atol=1e-12)


# ---- G12: operator_expectation on a basis state equals the energy ----------
def test_g12_operator_expectation_equals_energy():
  """tests/models/energy_test.py:194-226, 269-302."""
  n = 3
  thetas = np.array([100.0, -200.0, 300.0, 10, -20, 30])
  bits = np.array([[0, 0, 1]])
  shards = orc.kobe_shards(n, 2)
  vals = orc.expectations(np.zeros(0, dtype=orc.GATE_DTYPE), n, [],
                          orc.bitstrings_to_index(bits), shards)
  np.testing.assert_allclose(vals[0] @ thetas, orc.kobe_energy(bits, 2, thetas)[0])
  tb = np.array([0.3, -0.7, 1.1])
  vals = orc.expectations(np.zeros(0, dtype=orc.GATE_DTYPE), n, [],
                          orc.bitstrings_to_index(bits), orc.bernoulli_shards(n))
  np.testing.assert_allclose(vals[0] @ tb, orc.bernoulli_energy(bits, tb)[0])


# ---- G13: self-QMHL optimum (qmhl_loss_test.py:48-80) ----------------------
def test_g13_self_qmhl_is_entropy_with_zero_gradient():
  n = 3
  gates, names = orc.hea_circuit(n, 1)
  rng = np.random.default_rng(6)
  phi = rng.uniform(-1, 1, len(names))
  kthetas = rng.uniform(-1, 1, len(orc.parity_indices(n, n)))
  bits = orc.all_bitstrings(n)
  energies = orc.kobe_energy(bits, n, kthetas)
  p = orc.analytic_probabilities(energies)
  # data = model: circuit + circuit^-1 is the identity, so <K> = E_p[E].
  total = orc.concat_circuits(gates, len(names), orc.inverse_circuit(gates))
  val, _ = orc.modular_hamiltonian_expectation(
      total, n, np.concatenate([phi, phi]), bits, p, orc.kobe_shards(n, n), kthetas)
  loss = val + orc.analytic_log_partition(energies)
  np.testing.assert_allclose(loss, orc.analytic_entropy(energies), atol=1e-10)
  feats = orc.parity_features(bits, n)
  gtheta = p @ feats + orc.log_partition_gradient(p, feats)
  np.testing.assert_allclose(gtheta, 0, atol=1e-12)


def test_bit_order_quirk():
  """lexicographic sort of bit_circuit_bit_{k}."""
  assert orc.bit_column_to_qubit(10) == list(range(10))
  assert orc.bit_column_to_qubit(12) == [0, 1, 10, 11, 2, 3, 4, 5, 6, 7, 8, 9]
  b = np.zeros((1, 12), dtype=np.int8)
  b[0, 2] = 1  # column 2 drives qubit 10 -> bit (12-1-10) = 1
  assert orc.bitstrings_to_index(b)[0] == 2
  assert orc.bitstrings_to_index(b, ref_bit_order=False)[0] == 1 << 9
