"""GPU tests of the shot-based path: batched final states, shot sampling, binomial shot noise,
SampledQuantumInference (reference qnn.py:142-292) and the dense metrics unitary / density_matrix /
fidelity (reference qnn_utils.py, qhbm_utils.py).  Statistical tolerances are the reference's
(tests/inference/qnn_test.py:50-51: atol 2e-2 at 1e6 shots) unless noted."""
import itertools
import math

import numpy as np
import pytest
import torch

import helpers as hp
from oracle import qhbm_oracle as orc
from qhbmlib import circuits as cq
from qhbmlib import engine
from qhbmlib import inference
from qhbmlib import models
from qhbmlib import utils

pytestmark = pytest.mark.gpu
DEV = "cuda"
SEED = (11, 22)


def _bits(rows):
  return torch.tensor(rows, dtype=torch.int8, device=DEV)


def _plan(gates, n, n_sym, grad):
  terms, offs = hp.ops_to_tables([[(1.0, {0: "Z"})]], n)
  return engine.ExpectationPlan(gates, n, n_sym, terms, offs, grad)


# ---------------------------------------------------------------- final states
@pytest.mark.parametrize("n,grad", [(1, False), (3, True), (6, False), (9, True), (12, False), (12, True),
                                    (14, False), (14, True)])
def test_final_states_match_oracle(n, grad):
  """Padded (n < 9), single-tile and multi-tile plans, forward-only and adjoint kernels."""
  rng = np.random.default_rng(100 + n)
  n_sym = 5
  gates = hp.random_circuit(n, 30, n_sym, rng)
  phi = rng.uniform(-1, 1, n_sym).astype(np.float32)
  plan = _plan(gates, n, n_sym, grad)
  idx = rng.choice(1 << n, size=min(1 << n, 5), replace=False).astype(np.int64)
  states = plan.final_states(torch.tensor(idx, device=DEV), torch.tensor(phi, device=DEV)).cpu().numpy()
  assert states.shape == (len(idx), 1 << n)
  for row, k in zip(states, idx):
    ref = orc.simulate(gates, n, phi, int(k))
    np.testing.assert_allclose(row, ref, atol=3e-6)
  single = plan.state(int(idx[0]), torch.tensor(phi, device=DEV)).cpu().numpy()
  np.testing.assert_array_equal(single, states[0])


def test_final_states_chunked_batch(monkeypatch):
  """More states than one workspace chunk holds (multi-tile plan, chunk forced to 3)."""
  monkeypatch.setenv("QHBM_CHUNK", "3")
  n, n_sym = 14, 4
  rng = np.random.default_rng(7)
  gates = hp.random_circuit(n, 24, n_sym, rng)
  phi = rng.uniform(-1, 1, n_sym).astype(np.float32)
  plan = _plan(gates, n, n_sym, False)
  idx = rng.choice(1 << n, size=8, replace=False).astype(np.int64)
  states = plan.final_states(torch.tensor(idx, device=DEV), torch.tensor(phi, device=DEV)).cpu().numpy()
  for row, k in zip(states, idx):
    np.testing.assert_allclose(row, orc.simulate(gates, n, phi, int(k)), atol=3e-6)


# ---------------------------------------------------------------- shot sampling kernel
def test_sample_states_distribution_and_seeding():
  n = 5
  rng = np.random.default_rng(3)
  amps = (rng.normal(size=(3, 1 << n)) + 1j * rng.normal(size=(3, 1 << n))).astype(np.complex64)
  amps[1, 7:] = 0  # zero-probability tail must never be drawn
  amps /= np.linalg.norm(amps, axis=1, keepdims=True)
  states = torch.tensor(amps, device=DEV)
  counts = torch.tensor([200000, 50000, 0], device=DEV)
  keys, offsets = engine.sample_states(states, counts, SEED)
  assert offsets.tolist() == [0, 200000, 250000, 250000]
  keys_again, _ = engine.sample_states(states, counts, SEED)
  assert torch.equal(keys, keys_again)
  other, _ = engine.sample_states(states, counts, (11, 23))
  assert not torch.equal(keys, other)
  k = keys.cpu().numpy()
  for u, (lo, hi) in enumerate([(0, 200000), (200000, 250000)]):
    hist = np.bincount(k[lo:hi], minlength=1 << n)
    p = np.abs(amps[u])**2
    assert hist[p == 0].sum() == 0
    expected = p * (hi - lo)
    mask = expected > 0
    chi2 = np.sum((hist[mask] - expected[mask])**2 / expected[mask])
    assert chi2 < 2.5 * mask.sum(), chi2  # dof ~ 31 (6): far below a biased sampler's chi2


def test_sample_states_large_state_and_many_slices():
  """n = 16 (256-amplitude chunks) with enough shots for several sample slices per state."""
  n = 16
  rng = np.random.default_rng(5)
  p = rng.dirichlet(np.full(64, 0.5))
  support = rng.choice(1 << n, size=64, replace=False)
  amps = np.zeros((2, 1 << n), dtype=np.complex64)
  amps[0, support] = np.sqrt(p) * np.exp(1j * rng.uniform(0, 6.28, 64))
  amps[1, (1 << n) - 1] = 1.0
  shots = 300000
  keys, _ = engine.sample_states(torch.tensor(amps, device=DEV), torch.tensor([shots, 5000], device=DEV), SEED)
  k = keys.cpu().numpy()
  assert np.all(k[shots:] == (1 << n) - 1)
  hist = np.bincount(k[:shots], minlength=1 << n)
  assert hist.sum() == hist[support].sum()
  np.testing.assert_allclose(hist[support] / shots, p, atol=5 * np.sqrt(p.max() / shots) + 1e-3)


# ---------------------------------------------------------------- binomial shot noise
@pytest.mark.parametrize("shots", [1, 7, 1000, 1000000])
def test_binomial_shots_moments(shots):
  exact = torch.tensor([-1.0, 1.0, 0.0, 0.3, -0.9999, 0.99, 1e-3], device=DEV)
  reps = 20000
  tiled = exact.repeat(reps, 1).contiguous()
  noisy = engine.binomial_shots(tiled, shots, SEED)
  assert torch.equal(noisy, engine.binomial_shots(tiled, shots, SEED))
  x = noisy.double().cpu().numpy()
  e = exact.double().cpu().numpy()
  assert np.all(x[:, 0] == -1.0) and np.all(x[:, 1] == 1.0)
  assert np.all(np.abs(x) <= 1.0)
  # every value is (2k - shots)/shots for an integer k
  k = (x * shots + shots) / 2
  np.testing.assert_allclose(k, np.round(k), atol=1e-3 * max(1, shots * 1e-6) + 1e-6 * shots)
  var = (1 - e**2) / shots
  assert np.all(np.abs(x.mean(0) - e) <= 5 * np.sqrt(var / reps) + 1e-7), (x.mean(0), e)
  # the variance estimate of a rare outcome rests on few events: widen its band accordingly
  events = reps * shots * 0.5 * (1 - np.abs(e))
  band = 5 * np.sqrt(1 / np.maximum(events, 1) + 2 / reps) * var
  assert np.all(np.abs(x.var(0) - var) <= band + 1e-12), (x.var(0), var)


# ---------------------------------------------------------------- SampledQuantumInference
def _p_qnn(num_bits, value):
  qubits = cq.GridQubit.rect(1, num_bits)
  p = cq.Symbol("p")
  circuit = cq.Circuit(cq.X(q)**p for q in qubits)
  qnn = models.DirectQuantumCircuit(circuit, initializer=lambda shape: torch.full(shape, value), name="p_qnn")
  return qubits, qnn


def test_sampled_init():
  """tests/inference/qnn_test.py:66-81."""
  _, qnn = _p_qnn(3, 0.3)
  actual = inference.SampledQuantumInference(qnn, 41827, name="test_qnn_name")
  assert actual.name == "test_qnn_name"
  assert actual._expectation_samples == 41827
  assert actual.circuit is qnn


def test_sampled_expectation_xpow():
  """tests/inference/qnn_test.py:83-181 (G1): X^p|s>, <X>=0, <Y>=-(-1)^s sin(pi p), <Z>=(-1)^s cos(pi p)
  and d/dp, at the reference's sampled tolerance."""
  num_bits, p_val = 3, 0.37
  qubits, qnn = _p_qnn(num_bits, p_val)
  q_infer = inference.SampledQuantumInference(qnn, int(1e6), initial_seed=5)
  ops = cq.convert_to_tensor([1.0 * cq.X(q) for q in qubits] + [1.0 * cq.Y(q) for q in qubits] +
                             [1.0 * cq.Z(q) for q in qubits])
  bitstrings = _bits(list(itertools.product([0, 1], repeat=num_bits)))
  out = q_infer.expectation(bitstrings, ops)
  assert out.shape == (8, 9)
  signs = 1.0 - 2.0 * bitstrings.float().cpu().numpy()
  expected = np.concatenate([np.zeros((8, 3)), -signs * math.sin(math.pi * p_val), signs * math.cos(math.pi * p_val)], 1)
  np.testing.assert_allclose(out.detach().cpu().numpy(), expected, atol=2e-2)
  assert not np.array_equal(out.detach().cpu().numpy(), expected.astype(np.float32))  # it IS shot noise
  # d/dp of sum of <Y_k> and <Z_k> over all inputs, weighted to avoid cancellation
  w = torch.tensor(np.concatenate([np.zeros((8, 3)), -signs, signs], 1), dtype=torch.float32, device=DEV)
  (out * w).sum().backward()
  grad = qnn.trainable_variables[0].grad.cpu().numpy()
  exact = 24 * math.pi * (math.cos(math.pi * p_val) - math.sin(math.pi * p_val))
  np.testing.assert_allclose(grad, [exact], atol=24 * 2e-2)


def _random_direct_circuit(qubits, names, rng, layers=2):
  syms = [cq.Symbol(s) for s in names]
  ops, k = [], 0
  for _ in range(layers):
    for q in qubits:
      gate = [cq.X, cq.Y, cq.Z, cq.H][int(rng.integers(4))]
      ops.append(gate(q)**(syms[k % len(syms)] * float(rng.choice([1.0, -0.5, 2.0]))))
      k += 1
    for a, b in zip(qubits[:-1], qubits[1:]):
      gate = [cq.CZ, cq.CNOT, cq.ZZ, cq.XX][int(rng.integers(4))]
      ops.append(gate(a, b)**syms[k % len(syms)])
      k += 1
  return cq.Circuit(ops)


def test_sampled_matches_analytic_modular_hamiltonian():
  """tests/inference/qnn_test.py:266-369: expectation of a PauliMixin modular Hamiltonian and its
  derivatives w.r.t. the state circuit, the Hamiltonian circuit and the energy variables."""
  rng = np.random.default_rng(17)
  qubits = cq.GridQubit.rect(1, 3)
  state = models.DirectQuantumCircuit(_random_direct_circuit(qubits, ["a0", "a1", "a2"], rng),
                                      initializer=lambda s: torch.tensor(rng.uniform(0.25, 0.75, s), dtype=torch.float32))
  ham_circuit = models.DirectQuantumCircuit(_random_direct_circuit(qubits, ["b0", "b1"], rng),
                                            initializer=lambda s: torch.tensor(rng.uniform(0.25, 0.75, s),
                                                                                dtype=torch.float32))
  energy = models.KOBE(list(range(3)), 2, initializer=lambda s: torch.tensor(rng.uniform(-1, 1, s), dtype=torch.float32))
  energy.to(DEV)
  hamiltonian = models.Hamiltonian(energy, ham_circuit)
  bitstrings = _bits([[0, 1, 1], [1, 0, 0], [0, 1, 1], [1, 1, 1]])
  variables = state.trainable_variables + hamiltonian.trainable_variables

  def run(q_infer):
    for v in variables:
      v.grad = None
    out = q_infer.expectation(bitstrings, hamiltonian)
    weights = torch.tensor([[1.0], [-0.5], [0.25], [2.0]], device=DEV)
    (out * weights).sum().backward()
    return out.detach().cpu().numpy(), [v.grad.detach().cpu().numpy().copy() for v in variables]

  exact_out, exact_grads = run(inference.AnalyticQuantumInference(state, grad_mode="exact"))
  sampled_out, sampled_grads = run(inference.SampledQuantumInference(state, int(1e6), initial_seed=9))
  assert sampled_out.shape == (4, 1)
  np.testing.assert_allclose(sampled_out, exact_out, atol=2e-2)
  assert len(sampled_grads) == len(exact_grads) == 3
  assert max(np.abs(g).max() for g in exact_grads) > 0.1
  for s, e in zip(sampled_grads, exact_grads):
    np.testing.assert_allclose(s, e, atol=6e-2)  # sums of up to ~10 shifted terms at 2e-2 each


def test_sampled_expectation_bitstring_energy():
  """tests/inference/qnn_test.py:372-550: a Hamiltonian whose diagonal is a general BitstringEnergy
  (dense stack, no Pauli form).  Exact reference: sum_x |<x|V|s>|^2 E(x) with V from `unitary`."""
  rng = np.random.default_rng(23)
  n = 3
  qubits = cq.GridQubit.rect(1, n)
  init = lambda s: torch.tensor(rng.uniform(0.25, 0.75, s), dtype=torch.float32)
  state = models.DirectQuantumCircuit(_random_direct_circuit(qubits, ["s0", "s1"], rng), initializer=init)
  ham_circuit = models.DirectQuantumCircuit(_random_direct_circuit(qubits, ["h0", "h1"], rng), initializer=init)
  torch.manual_seed(3)
  layers = [torch.nn.Linear(n, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1), utils.Squeeze(-1)]

  class _Float(torch.nn.Module):

    def forward(self, x):
      return x.float()

  energy = models.BitstringEnergy(list(range(n)), [_Float()] + layers).to(DEV)
  hamiltonian = models.Hamiltonian(energy, ham_circuit)
  with pytest.raises(TypeError):
    inference.AnalyticQuantumInference(state).expectation(_bits([[0, 0, 0]]), hamiltonian)
  q_infer = inference.SampledQuantumInference(state, int(1e6), initial_seed=2)
  bitstrings = _bits([[0, 0, 1], [1, 1, 0], [0, 0, 1]])
  weights = torch.tensor([[1.0], [-2.0], [0.5]], device=DEV)

  def exact(state_vals, ham_vals):
    """Exact values through the dense unitaries; differentiable w.r.t. the energy variables."""
    with torch.no_grad():
      old_s, old_h = state.trainable_variables[0].clone(), ham_circuit.trainable_variables[0].clone()
      state.trainable_variables[0].copy_(state_vals)
      ham_circuit.trainable_variables[0].copy_(ham_vals)
      total = inference.unitary(hamiltonian.circuit_dagger) @ inference.unitary(state)
      state.trainable_variables[0].copy_(old_s)
      ham_circuit.trainable_variables[0].copy_(old_h)
    all_bits = _bits(list(itertools.product([0, 1], repeat=n)))
    energies = energy(all_bits)
    idx = (bitstrings.long() * torch.tensor([4, 2, 1], device=DEV)).sum(1)
    probs = (total[:, idx].abs()**2).transpose(0, 1)  # [U, 2^n]
    return (probs.float() @ energies).unsqueeze(1)

  s0 = state.trainable_variables[0].detach().clone()
  h0 = ham_circuit.trainable_variables[0].detach().clone()
  for v in energy.parameters():
    v.grad = None
  ex = exact(s0, h0)
  (ex * weights).sum().backward()
  exact_energy_grads = [v.grad.detach().clone() for v in energy.parameters()]

  def fd(which, k, eps=1e-3):
    d = torch.zeros_like(s0 if which == "s" else h0)
    d[k] = eps
    hi = exact(s0 + d, h0) if which == "s" else exact(s0, h0 + d)
    lo = exact(s0 - d, h0) if which == "s" else exact(s0, h0 - d)
    return float(((hi - lo) * weights).sum().detach() / (2 * eps))

  exact_state_grad = [fd("s", k) for k in range(2)]
  exact_ham_grad = [fd("h", k) for k in range(2)]

  for v in list(energy.parameters()) + state.trainable_variables + ham_circuit.trainable_variables:
    v.grad = None
  out = q_infer.expectation(bitstrings, hamiltonian)
  assert out.shape == (3, 1)
  np.testing.assert_allclose(out.detach().cpu().numpy(), ex.detach().cpu().numpy(), atol=2e-2)
  (out * weights).sum().backward()
  np.testing.assert_allclose(state.trainable_variables[0].grad.cpu().numpy(), exact_state_grad, atol=6e-2)
  np.testing.assert_allclose(ham_circuit.trainable_variables[0].grad.cpu().numpy(), exact_ham_grad, atol=6e-2)
  assert max(abs(x) for x in exact_state_grad + exact_ham_grad) > 0.05
  for v, g in zip(energy.parameters(), exact_energy_grads):
    np.testing.assert_allclose(v.grad.cpu().numpy(), g.cpu().numpy(), atol=2e-2)


def test_sampled_gradient_rejects_three_eigenvalue_gates():
  qubits = cq.GridQubit.rect(1, 2)
  circuit = cq.Circuit(cq.ISWAP(qubits[0], qubits[1])**cq.Symbol("t"), cq.X(qubits[0])**cq.Symbol("u"))
  qnn = models.DirectQuantumCircuit(circuit, initializer=lambda s: torch.full(s, 0.4))
  q_infer = inference.SampledQuantumInference(qnn, 1000, initial_seed=1)
  out = q_infer.expectation(_bits([[0, 1]]), cq.convert_to_tensor([1.0 * cq.Z(qubits[0])]))
  assert out.shape == (1, 1)
  with pytest.raises(NotImplementedError, match="two-eigenvalue"):
    out.sum().backward()


def test_sample_basic():
  """tests/inference/qnn_test.py:552-604: identity, bit flip and GHZ circuits."""
  num_bits = 3
  qubits = cq.GridQubit.rect(1, num_bits)
  rows = list(itertools.product([0, 1], repeat=num_bits))
  bitstrings = _bits(rows)
  counts = torch.randint(100, 1000, (len(rows),), device=DEV, generator=torch.Generator(DEV).manual_seed(1))
  ident = models.DirectQuantumCircuit(cq.Circuit(cq.I(q) for q in qubits), name="identity")
  samples = inference.SampledQuantumInference(ident, 10)._sample(bitstrings, counts)
  assert len(samples) == len(rows)
  for i, (b, c) in enumerate(zip(rows, counts.tolist())):
    assert samples[i].shape == (c, num_bits)
    assert torch.all(samples[i] == _bits([b]))
  flip = models.DirectQuantumCircuit(cq.Circuit(cq.X(q) for q in qubits), name="flip")
  samples = inference.SampledQuantumInference(flip, 10)._sample(bitstrings, counts)
  for i, (b, c) in enumerate(zip(rows, counts.tolist())):
    assert samples[i].shape == (c, num_bits)
    assert torch.all(samples[i] == 1 - _bits([b]))
  ghz_circuit = cq.Circuit(cq.X(qubits[0])**cq.Symbol("ghz")) + cq.Circuit(
      cq.CNOT(a, b) for a, b in zip(qubits, qubits[1:]))
  ghz = models.DirectQuantumCircuit(ghz_circuit, initializer=lambda s: torch.full(s, 0.5), name="ghz")
  samples = inference.SampledQuantumInference(ghz, 10)._sample(_bits([[0] * num_bits]), counts[:1])[0]
  seen = {tuple(r) for r in samples.cpu().tolist()}
  assert seen == {(0, 0, 0), (1, 1, 1)}


def test_sample_uneven():
  """tests/inference/qnn_test.py:606-621: H|0> with different shot counts per row."""
  max_counts = int(1e7)
  counts = torch.tensor([max_counts // 2, max_counts], device=DEV)
  qnn = models.DirectQuantumCircuit(cq.Circuit(cq.H(cq.GridQubit(0, 0))))
  samples = inference.SampledQuantumInference(qnn, 10)._sample(_bits([[0], [0]]), counts)
  assert samples.row_lengths().tolist() == [max_counts // 2, max_counts]
  for i, c in enumerate(counts.tolist()):
    ones = int(samples[i].sum().item())
    assert abs(ones - c / 2) < c // 1000


# ---------------------------------------------------------------- dense metrics
def test_unitary_matches_oracle():
  """tests/inference/qnn_utils_test.py:41-...: unitary of a random parameterised circuit."""
  rng = np.random.default_rng(31)
  for n in (1, 3, 6, 10):
    qubits = cq.GridQubit.rect(1, n)
    if n == 1:
      circuit = cq.Circuit(cq.X(qubits[0])**cq.Symbol("u0"), cq.Z(qubits[0])**cq.Symbol("u1"))
    else:
      circuit = _random_direct_circuit(qubits, ["u0", "u1", "u2"], rng)
    qnn = models.DirectQuantumCircuit(circuit, initializer=lambda s: torch.tensor(rng.uniform(-1, 1, s),
                                                                                dtype=torch.float32))
    actual = inference.unitary(qnn).cpu().numpy()
    gates = qnn.gate_table()
    phi = qnn.symbol_values.detach().cpu().numpy()
    expected = np.stack([orc.simulate(gates, n, phi, k) for k in range(1 << n)], axis=1)
    np.testing.assert_allclose(actual, expected, atol=3e-6)
    np.testing.assert_allclose(actual.conj().T @ actual, np.eye(1 << n), atol=2e-5)


def _bell_model(theta):
  qubits = cq.GridQubit.rect(1, 2)
  circuit = cq.Circuit(cq.H(qubits[0]), cq.CNOT(qubits[0], qubits[1]))
  qnn = models.DirectQuantumCircuit(circuit)
  energy = models.BernoulliEnergy([0, 1], initializer=lambda s: torch.tensor(theta, dtype=torch.float32))
  energy.to(DEV)
  return models.Hamiltonian(energy, qnn)


def test_density_matrix_bell():
  """tests/inference/qhbm_utils_test.py:29-61 (G10): strongly polarised spins -> |00>, H then CNOT
  gives the Bell state (|00>+|11>)/sqrt 2."""
  model = _bell_model([-10.0, -10.0])  # E(b) = sum (1-2b) theta: theta << 0 favours b = 0
  rho = inference.density_matrix(model).cpu().numpy()
  expected = np.zeros((4, 4), dtype=np.complex64)
  expected[0, 0] = expected[0, 3] = expected[3, 0] = expected[3, 3] = 0.5
  np.testing.assert_allclose(rho, expected, atol=1e-6)


def test_fidelity_self_and_random():
  """tests/inference/qhbm_utils_test.py:63-...: F(rho, rho) = 1; random sigma vs the textbook
  formula evaluated with dense eigendecompositions in float64."""
  rng = np.random.default_rng(41)
  n = 3
  qubits = cq.GridQubit.rect(1, n)
  qnn = models.DirectQuantumCircuit(_random_direct_circuit(qubits, ["f0", "f1", "f2"], rng),
                                    initializer=lambda s: torch.tensor(rng.uniform(-1, 1, s), dtype=torch.float32))
  energy = models.KOBE(list(range(n)), 2, initializer=lambda s: torch.tensor(rng.uniform(-1, 1, s), dtype=torch.float32))
  energy.to(DEV)
  model = models.Hamiltonian(energy, qnn)
  rho = inference.density_matrix(model)
  np.testing.assert_allclose(float(inference.fidelity(model, rho)), 1.0, atol=1e-4)
  a = rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8))
  sigma = a @ a.conj().T
  sigma /= np.trace(sigma)
  r = rho.cpu().numpy().astype(np.complex128)
  w, v = np.linalg.eigh(r)
  sqrt_r = (v * np.sqrt(np.clip(w, 0, None))) @ v.conj().T
  ev = np.linalg.eigvalsh(sqrt_r @ sigma @ sqrt_r)
  expected = np.sum(np.sqrt(np.clip(ev, 0, None)))**2
  actual = float(inference.fidelity(model, torch.tensor(sigma, dtype=torch.complex64)))
  np.testing.assert_allclose(actual, expected, rtol=1e-4)
