"""CPU tests of the host compiler (plan.cpp): the compiled sweep/pass/op program, run by
the TEST-ONLY scalar interpreter in tests/native/, must reproduce the oracle.

This covers the scheduling logic (tiles, register qubits, diagonal-run tables, gradient
slots) and the coefficient jobs without a GPU.  The CUDA kernels execute the same program;
their parity tests are the `-m gpu` ones."""
import numpy as np
import pytest

from oracle import qhbm_oracle as orc
import helpers as hp


def _check(gates, n, nsym, ops, rng, T, K, mode="exact", atol=2e-5):
  phi = rng.uniform(-1, 1, max(nsym, 1)).astype(np.float32)[:nsym]
  dg = rng.uniform(-1, 1, len(ops)).astype(np.float32)
  basis = int(rng.integers(0, 1 << n))
  e, g, state, info = hp.verify_run(gates, n, nsym, ops, phi, basis, dg, True, T, K,
                                    {"exact": 0, "tfq_fd": 1}[mode])
  ref_state = orc.simulate(gates, n, phi, basis)
  np.testing.assert_allclose(state, ref_state, atol=atol)
  e_ref, g_ref = orc.adjoint_gradient(gates, n, phi, basis, ops, dg, mode)
  np.testing.assert_allclose(e, e_ref, atol=atol * 10)
  np.testing.assert_allclose(g, g_ref, atol=atol * 20)
  return info


@pytest.mark.parametrize("n,layers,T,K", [(3, 2, 0, 4), (4, 2, 0, 5), (6, 3, 0, 4), (10, 2, 10, 5),
                                           (11, 2, 9, 4), (12, 2, 10, 5), (12, 3, 9, 4)])
def test_hea_tfim_xxz(n, layers, T, K):
  rng = np.random.default_rng(100 + n)
  gates, names = orc.hea_circuit(n, layers)
  ops = [orc.tfim_ring(n), orc.xxz_ring(n)] + orc.kobe_shards(n, 2)[:5]
  info = _check(gates, n, len(names), ops, rng, T, K)
  n_eff = max(n, K + 5)
  assert info[4] == (min(T, n_eff) if T else (n_eff if n_eff <= 13 else 12))


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("n,T,K", [(2, 0, 4), (5, 0, 5), (11, 9, 4), (11, 10, 5)])
def test_random_circuits_all_gate_types(seed, n, T, K):
  rng = np.random.default_rng(1000 * n + seed)
  nsym = 6
  gates = hp.random_circuit(n, 30, nsym, rng)
  ops = hp.random_ops(n, 3, rng)
  _check(gates, n, nsym, ops, rng, T, K)


@pytest.mark.parametrize("n,T,K", [(4, 0, 4), (11, 9, 4)])
def test_tfq_fd_mode(n, T, K):
  rng = np.random.default_rng(7)
  gates = hp.random_circuit(n, 25, 5, rng)
  ops = hp.random_ops(n, 2, rng)
  _check(gates, n, 5, ops, rng, T, K, mode="tfq_fd")


@pytest.mark.parametrize("n,T,K", [(5, 0, 4), (11, 9, 4)])
def test_many_diagonal_terms_table(n, T, K):
  """>= 32 diagonal terms are routed to the WHT term table (checked here through the interpreter)."""
  rng = np.random.default_rng(21)
  gates, names = orc.hea_circuit(n, 1)
  ops = orc.kobe_shards(n, 3)[:40] + [orc.tfim_ring(n), [(0.7, {}), (1.1, {0: "Z", n - 1: "Z"})]]
  _check(gates, n, len(names), ops, rng, T, K)


def test_empty_circuit_and_identity_terms():
  rng = np.random.default_rng(3)
  for n, T, K in [(3, 0, 4), (11, 9, 4)]:
    gates = np.zeros(0, dtype=orc.GATE_DTYPE)
    ops = [[(1.5, {}), (0.5, {0: "Z"})], [(2.0, {n - 1: "X"})]]
    _check(gates, n, 0, ops, rng, T, K)


def test_qmhl_style_circuit_plus_dagger():
  """U_data + U_model^-1 with Z-shard observables (qnn.py:69-72, hamiltonian.py:48-51)."""
  rng = np.random.default_rng(11)
  n = 11
  g1, n1 = orc.hea_circuit(n, 1, "data")
  g2, n2 = orc.hea_circuit(n, 1, "model")
  total = orc.concat_circuits(g1, len(n1), orc.inverse_circuit(g2))
  ops = orc.kobe_shards(n, 2)[:12]
  _check(total, n, len(n1) + len(n2), ops, rng, 9, 4)


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("n,T,K", [(3, 0, 4), (6, 0, 4), (10, 9, 4), (11, 10, 5)])
def test_single_observable_passes(seed, n, T, K):
  """One observable made of 1- and 2-local X/Y strings with arbitrary Z tails: the strings whose flips
  fit the tile run as observable passes (OP_HX / OP_HD), odd-Y strings and cross-tile flips stay in the
  generic tables; both must add up to the oracle's value and adjoint gradient."""
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
  _check(gates, n, nsym, [terms], rng, T, K)


@pytest.mark.parametrize("n,T,K,kind", [(11, 10, 5, "tfim"), (12, 10, 5, "xxz"), (13, 10, 5, "random"),
                                        (12, 9, 4, "random"), (14, 10, 5, "tfim"), (6, 0, 5, "random")])
def test_forward_only_plans_with_expectation_stages(n, T, K, kind, monkeypatch):
  """Forward-only plans of multi-tile states can evaluate x-groups that flip out-of-tile qubits in extra
  expectation stages with their own tile maps (QHBM_EXPECT_STAGES; off by default because it measured
  neutral on B200) and run a single observable as observable passes (QHBM_HPASS_FORWARD); values must
  match the oracle for every observable."""
  monkeypatch.setenv("QHBM_EXPECT_STAGES", "1")
  monkeypatch.setenv("QHBM_HPASS_FORWARD", "1")
  rng = np.random.default_rng(n * 7 + T)
  gates, names = orc.hea_circuit(n, 2)
  nsym = len(names)
  if kind == "tfim":
    ops = [orc.tfim_ring(n)]
  elif kind == "xxz":
    ops = [orc.xxz_ring(n), orc.tfim_ring(n)]
  else:
    ops = hp.random_ops(n, 3, rng, max_terms=6)
  phi = rng.uniform(-1, 1, nsym).astype(np.float32)
  basis = int(rng.integers(0, 1 << n))
  e, _, state, info = hp.verify_run(gates, n, nsym, ops, phi, basis, np.zeros(len(ops), np.float32), False, T, K)
  np.testing.assert_allclose(state, orc.simulate(gates, n, phi, basis), atol=2e-5)
  e_ref = orc.expectations(gates, n, phi, [basis], ops)[0]
  np.testing.assert_allclose(e, e_ref, atol=2e-4)
  if n > max(T, K + 5) and kind != "random":
    assert info[6] > info[0] + 1  # launches > forward sweeps + one expectation launch


def test_deep_circuit_several_flush_windows():
  """A single-tile plan with 18 gradient passes in one launch: the reduced sums of all its passes exceed one
  flush window (kGaccFloats), so the device program must flush mid-launch and rebase later passes."""
  rng = np.random.default_rng(5)
  n, layers = 10, 14
  gates, names = orc.hea_circuit(n, layers)
  _check(gates, n, len(names), [orc.xxz_ring(n)], rng, 10, 4)


@pytest.mark.parametrize("block", range(6))
def test_fuzz_compiler_over_sizes_tiles_and_modes(block):
  """Seeded fuzz of the host compiler: random qubit counts (2..13), tile / register qubits (defaults and forced,
  single- and multi-tile), HEA or random circuits over every gate type (also empty), mixed observables (random
  strings, TFIM ring, Z-shards up to the WHT table threshold), both gradient modes; the compiled program run by
  the scalar interpreter must reproduce the oracle's state, expectations and gradient (worst errors seen over the
  300 cases: state 1.3e-7, expectation 2.1e-7 x sum|coeff|, gradient 7.4e-7 x sum|coeff|)."""
  for s in range(50 * block, 50 * block + 50):
    rng = np.random.default_rng(90000 + s)
    n = int(rng.integers(2, 14))
    K = int(rng.choice([0, 4, 5]))
    T = int(rng.choice([0] + list(range((K if K else 4) + 5, n + 1))))
    kind = int(rng.integers(0, 3))
    if kind == 0:
      gates, names = orc.hea_circuit(n, int(rng.integers(1, 4)))
      nsym = len(names)
    else:
      n_gates = int(rng.integers(0, 40))
      nsym = int(rng.integers(1, 7)) if n_gates else 0
      gates = hp.random_circuit(n, n_gates, nsym, rng) if n_gates else np.zeros(0, dtype=orc.GATE_DTYPE)
    ops = hp.random_ops(n, int(rng.integers(1, 4)), rng)
    if rng.random() < 0.3:
      ops = ops + [orc.tfim_ring(n)]
    if rng.random() < 0.3 and n >= 3:
      ops = ops + orc.kobe_shards(n, 2)[:int(rng.integers(1, 40))]
    mode = ("exact", "tfq_fd")[int(rng.integers(0, 2))]
    phi = rng.uniform(-1, 1, max(nsym, 1)).astype(np.float32)[:nsym]
    dg = rng.uniform(-1, 1, len(ops)).astype(np.float32)
    basis = int(rng.integers(0, 1 << n))
    e, g, state, _ = hp.verify_run(gates, n, nsym, ops, phi, basis, dg, True, T, K, {"exact": 0, "tfq_fd": 1}[mode])
    scale = max(1.0, max(sum(abs(c) for c, _ in op) for op in ops))
    what = f"seed {s}: n={n} T={T} K={K} kind={kind} mode={mode}"
    np.testing.assert_allclose(state, orc.simulate(gates, n, phi, basis), atol=2e-6, err_msg=what)
    e_ref, g_ref = orc.adjoint_gradient(gates, n, phi, basis, ops, dg, mode)
    np.testing.assert_allclose(e, e_ref, atol=3e-6 * scale, err_msg=what)
    np.testing.assert_allclose(g, g_ref, atol=1e-5 * scale, err_msg=what)
