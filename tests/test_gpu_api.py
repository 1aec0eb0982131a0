"""GPU tests of the reference-shaped Python API (qhbmlib.inference / models / utils on the CUDA
engine).  Each test mirrors a reference test (file:line under /root/reference/tests) and checks
the same closed form; tolerances are the reference's unless tighter is noted."""
import itertools
import math

import numpy as np
import pytest
import torch

from oracle import qhbm_oracle as orc
from qhbmlib import architectures as arch
from qhbmlib import circuits as cq
from qhbmlib import data
from qhbmlib import inference
from qhbmlib import models
from qhbmlib import utils
from qhbmlib.models import energy_utils

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bits(rows):
  return torch.tensor(rows, dtype=torch.int8, device=DEV)


# ---------------------------------------------------------------- utils_test.py
def test_utils_unique_and_expand():
  """tests/utils_test.py:107-186."""
  b = _bits([[1, 0, 1], [1, 1, 1], [0, 1, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1], [1, 0, 1], [1, 0, 1]])
  y, idx, count = utils.unique_bitstrings_with_counts(b)
  assert y.tolist() == [[1, 0, 1], [1, 1, 1], [0, 1, 1]]
  assert idx.tolist() == [0, 1, 2, 0, 1, 2, 0, 0]
  assert count.tolist() == [4, 2, 2]
  assert torch.equal(utils.expand_unique_results(y, idx), b)
  vals = torch.rand(3, device=DEV)
  assert torch.equal(utils.expand_unique_results(vals, idx), vals[idx.long()])
  b1 = _bits([[1], [0], [0], [1], [1], [0], [1], [1]])
  y, idx, count = utils.unique_bitstrings_with_counts(b1)
  assert y.tolist() == [[1], [0]] and idx.tolist() == [0, 1, 1, 0, 0, 1, 0, 0] and count.tolist() == [5, 3]


def test_utils_weighted_average_and_gradients():
  """tests/utils_test.py:48-104."""
  counts = torch.tensor([37, 5], dtype=torch.int32, device=DEV)
  vals = torch.tensor([[2.7, -5.9], [0.5, 3.0]], device=DEV, requires_grad=True)
  out = utils.weighted_average(counts, vals)
  exp = (37 * vals[0] + 5 * vals[1]) / 42
  np.testing.assert_allclose(out.detach().cpu(), exp.detach().cpu(), rtol=1e-6)
  out.sum().backward()
  np.testing.assert_allclose(vals.grad.cpu(), [[37 / 42] * 2, [5 / 42] * 2], rtol=1e-6)
  # gradient through expand (segment-sum kernel)
  y = torch.rand((3, 2), device=DEV, requires_grad=True)
  idx = torch.tensor([0, 1, 2, 0, 1, 2, 0, 0], dtype=torch.int32, device=DEV)
  w = torch.arange(16, device=DEV, dtype=torch.float32).reshape(8, 2)
  (utils.expand_unique_results(y, idx) * w).sum().backward()
  ref = torch.zeros(3, 2, device=DEV).index_add_(0, idx.long(), w)
  np.testing.assert_allclose(y.grad.cpu(), ref.cpu(), rtol=1e-6)


# ---------------------------------------------------------------- qnn_test.py
def _p_qnn(num_bits, seed=11):
  qubits = cq.GridQubit.rect(1, num_bits)
  p = cq.Symbol("p")
  circuit = cq.Circuit(cq.X(q)**p for q in qubits)
  return qubits, models.DirectQuantumCircuit(circuit, energy_utils.RandomUniform(-1.0, 1.0, seed), name="p_qnn")


@pytest.mark.parametrize("grad_mode,gtol", [("exact", 1e-4), ("tfq_fd", 2e-3)])
def test_qnn_expectation_xpow_closed_form(grad_mode, gtol):
  """tests/inference/qnn_test.py:83-180 (reference atol 2e-3)."""
  num_bits = 3
  qubits, p_qnn = _p_qnn(num_bits)
  states = 5 * list(itertools.product([0, 1], repeat=num_bits))
  initial_states = _bits(states)
  qnn = inference.AnalyticQuantumInference(p_qnn, grad_mode=grad_mode)
  p = float(p_qnn.symbol_values[0].detach())
  sin_p, cos_p = math.sin(math.pi * p), math.cos(math.pi * p)
  expected = {
      "X": ([[0.0] * 3 for _ in states], [[0.0] * 3 for _ in states]),
      "Y": ([[-((-1.0)**s) * sin_p for s in b] for b in states],
            [[-((-1.0)**s) * math.pi * cos_p for s in b] for b in states]),
      "Z": ([[((-1.0)**s) * cos_p for s in b] for b in states],
            [[-((-1.0)**s) * math.pi * sin_p for s in b] for b in states]),
  }
  for name, gate in (("X", cq.X), ("Y", cq.Y), ("Z", cq.Z)):
    ops = cq.convert_to_tensor([1 * gate(q) for q in qubits])
    var = p_qnn.trainable_variables[0]
    out = qnn.expectation(initial_states, ops)
    assert tuple(out.shape) == (len(states), num_bits)
    np.testing.assert_allclose(out.detach().cpu(), expected[name][0], atol=1e-5)
    jac = torch.stack([torch.autograd.grad(out[i, j], var, retain_graph=True)[0].reshape(())
                       for i in range(0, len(states), 7) for j in range(num_bits)])
    ref = np.array([expected[name][1][i][j] for i in range(0, len(states), 7) for j in range(num_bits)])
    np.testing.assert_allclose(jac.cpu(), ref, atol=gtol)


def test_qnn_expectation_random_circuit_vs_oracle():
  """tests/inference/qnn_test.py:183-264 (there against cirq.Simulator, atol 2e-3)."""
  n = 5
  qubits = cq.GridQubit.rect(1, n)
  circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 3, "r"),
                                     energy_utils.RandomUniform(-1, 1, 3))
  qnn = inference.AnalyticQuantumInference(circ, grad_mode="exact")
  ops_list = [arch.tfim_ring(qubits), arch.xxz_ring(qubits), cq.PauliSum.from_pauli_strings(cq.Y(qubits[2]))]
  rng = np.random.default_rng(0)
  bits = rng.integers(0, 2, (40, n)).astype(np.int8)
  out = qnn.expectation(_bits(bits), cq.convert_to_tensor(ops_list))
  phi = circ.symbol_values.detach().cpu().numpy()
  gates = circ.gate_table().astype(orc.GATE_DTYPE)
  oracle_ops = [[(t.coefficient.real, {qubits.index(q): p for q, p in t.paulis.items()}) for t in s.terms]
                for s in ops_list]
  ref = orc.expectations(gates, n, phi, orc.bitstrings_to_index(bits), oracle_ops)
  np.testing.assert_allclose(out.detach().cpu(), ref, rtol=1e-5, atol=2e-5)
  up = torch.tensor(rng.uniform(-1, 1, ref.shape).astype(np.float32), device=DEV)
  (g,) = torch.autograd.grad((out * up).sum(), circ.trainable_variables)
  _, g_ref = orc.batch_expectation_and_gradient(gates, n, phi, orc.bitstrings_to_index(bits), oracle_ops,
                                                up.cpu().numpy())
  np.testing.assert_allclose(g.cpu(), g_ref.sum(0), rtol=1e-4, atol=1e-4)


def test_qnn_expectation_modular_hamiltonian():
  """tests/inference/qnn_test.py:266-369: circuit + H.circuit_dagger, Z shards, energy post-process."""
  n = 4
  qubits = cq.GridQubit.rect(1, n)
  data_circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 2, "d"),
                                          energy_utils.RandomUniform(-1, 1, 1), name="d")
  model_circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 1, "m"),
                                           energy_utils.RandomUniform(-1, 1, 2), name="m")
  energy = models.KOBE(list(range(n)), 2, energy_utils.RandomUniform(-1, 1, 3))
  ham = models.Hamiltonian(energy, model_circ)
  qnn = inference.AnalyticQuantumInference(data_circ, grad_mode="exact")
  bits = np.array(list(itertools.product([0, 1], repeat=n)), dtype=np.int8)
  out = qnn.expectation(_bits(bits), ham)
  assert tuple(out.shape) == (16, 1)
  g1 = data_circ.gate_table().astype(orc.GATE_DTYPE)
  g2 = model_circ.gate_table().astype(orc.GATE_DTYPE)
  total = orc.concat_circuits(g1, len(data_circ.symbol_names), orc.inverse_circuit(g2))
  phi = np.concatenate([data_circ.symbol_values.detach().cpu().numpy(),
                        model_circ.symbol_values.detach().cpu().numpy()])
  thetas = energy.post_process[0].kernel.detach().cpu().numpy()
  _, per_row = orc.modular_hamiltonian_expectation(total, n, phi, bits, np.ones(16), orc.kobe_shards(n, 2), thetas)
  np.testing.assert_allclose(out[:, 0].detach().cpu(), per_row, rtol=1e-5, atol=2e-5)
  grads = torch.autograd.grad(out.sum(), ham.trainable_variables + data_circ.trainable_variables)
  shards = orc.expectations(total, n, phi, orc.bitstrings_to_index(bits), orc.kobe_shards(n, 2))
  np.testing.assert_allclose(grads[0].cpu(), shards.sum(0), rtol=1e-4, atol=1e-4)  # d/dtheta = <Z shards>
  dg = np.tile(thetas[None, :], (16, 1))
  _, g_ref = orc.batch_expectation_and_gradient(total, n, phi, orc.bitstrings_to_index(bits),
                                                orc.kobe_shards(n, 2), dg)
  g_ref = g_ref.sum(0)
  np.testing.assert_allclose(grads[1].cpu(), g_ref[len(data_circ.symbol_names):], rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(grads[2].cpu(), g_ref[:len(data_circ.symbol_names)], rtol=1e-4, atol=1e-4)
  with pytest.raises(TypeError, match="General Hamiltonians not accepted"):
    bad = models.Hamiltonian(models.BitstringEnergy(list(range(n)), [torch.nn.Linear(n, 1)]), model_circ)
    qnn.expectation(_bits(bits), bad)


# ---------------------------------------------------------------- ebm_test.py
def test_analytic_init_and_all_bitstrings():
  """tests/inference/ebm_test.py:176-197."""
  energy = models.KOBE([0, 1, 3], 2)
  layer = inference.AnalyticEnergyInference(energy, 1000, [44, 22], "name")
  assert layer.name == "name" and layer.seed.tolist() == [44, 22]
  assert layer.all_bitstrings.tolist() == [list(b) for b in itertools.product([0, 1], repeat=3)]
  np.testing.assert_allclose(layer.all_energies.detach().cpu(), energy(layer.all_bitstrings).detach().cpu())


def test_analytic_sampling_follows_energy_and_seeding():
  """tests/inference/ebm_test.py:199-297."""
  n_samples = 1_000_000
  one_bit = models.KOBE([0], 1, energy_utils.Constant(0.0))
  layer = inference.AnalyticEnergyInference(one_bit, n_samples, initial_seed=[5, 6])
  s = layer.sample(n_samples)
  assert s.dtype == torch.int8 and tuple(s.shape) == (n_samples, 1)
  assert abs(float(s.float().mean()) - 0.5) < 5e-3
  one_bit.set_weights([torch.tensor([math.log(3.0) / 2])])   # p(1)/p(0) = e^{2 theta} = 3
  s = layer.sample(n_samples)
  assert abs(float(s.float().mean()) - 0.75) < 5e-3
  three = models.KOBE([0, 1, 2], 3, energy_utils.Constant(0.0))
  layer3 = inference.AnalyticEnergyInference(three, n_samples, initial_seed=[5, 6])
  three.set_weights([torch.tensor([100.0, 0.0, 0.0, -100.0, 0.0, 100.0, 0.0])])
  y, _, _ = utils.unique_bitstrings_with_counts(layer3.sample(n_samples))
  assert y.tolist() == [[1, 1, 0]]
  five = models.KOBE(list(range(5)), 2)
  l5 = inference.AnalyticEnergyInference(five, 1000, initial_seed=[5, 6])
  assert torch.equal(l5.sample(1000), l5.sample(1000))
  l5.seed = None
  assert not torch.equal(l5.sample(1000), l5.sample(1000))


def test_analytic_log_partition_entropy_and_gradient():
  """tests/inference/ebm_test.py:514-559."""
  energy = models.KOBE([0, 1], 2)
  layer = inference.AnalyticEnergyInference(energy, 1_000_000, initial_seed=[1, 2])
  energy.set_weights([torch.tensor([1.5, 2.7, -4.0])])
  lp = layer.log_partition()
  np.testing.assert_allclose(float(lp), math.log(3641.8353), rtol=1e-6)
  np.testing.assert_allclose(float(layer.entropy()), 0.00233551808, rtol=2e-4)
  (g,) = torch.autograd.grad(lp, energy.trainable_variables)
  bits = orc.all_bitstrings(2)
  th = np.array([1.5, 2.7, -4.0])
  p = orc.analytic_probabilities(orc.kobe_energy(bits, 2, th))
  np.testing.assert_allclose(g.cpu(), -(p @ orc.parity_features(bits, 2)), rtol=2e-2, atol=2e-3)


class AllOnes(torch.nn.Module):
  """prefactor * [x == 1...1] (tests/inference/ebm_test.py:357-372)."""

  def __init__(self, ones_prefactor):
    super().__init__()
    self.ones_prefactor = ones_prefactor

  def forward(self, inputs):
    return self.ones_prefactor * torch.prod(inputs.to(torch.float32), 1)


def test_expectation_explicit_all_ones_energy():
  """tests/inference/ebm_test.py:300-453: E[f] = mu p*, d/dtheta = mu p*(p*-1), d/dmu = p*; and the
  shared-variable case g = theta [x = 1]: d/dtheta = theta p*(p*-1) + p*.  (reference rtol 1e-2)"""
  num_bits = 3
  theta = torch.nn.Parameter(torch.tensor(-2.4, device=DEV))
  energy = models.BitstringEnergy(list(range(num_bits)), [AllOnes(theta)])
  pstar = float(torch.exp(-theta) / (2**num_bits - 1 + torch.exp(-theta)))
  e_infer = inference.AnalyticEnergyInference(energy, 2_000_000, initial_seed=[5, 6])
  mu = torch.nn.Parameter(torch.tensor(1.3, device=DEV))
  avg = e_infer.expectation(AllOnes(mu))
  g_theta, g_mu = torch.autograd.grad(avg, (theta, mu))
  np.testing.assert_allclose(float(avg), 1.3 * pstar, rtol=1e-2)
  np.testing.assert_allclose(float(g_theta), 1.3 * pstar * (pstar - 1), rtol=1e-2)
  np.testing.assert_allclose(float(g_mu), pstar, rtol=1e-2)
  avg = e_infer.expectation(AllOnes(theta))
  (g_theta,) = torch.autograd.grad(avg, (theta,))
  np.testing.assert_allclose(float(g_theta), -2.4 * pstar * (pstar - 1) + pstar, rtol=1e-2)
  # nested structure of outputs
  nested = e_infer.expectation(lambda x: {"a": AllOnes(mu)(x), "b": [torch.stack([AllOnes(mu)(x)] * 2, 1)]})
  np.testing.assert_allclose(float(nested["a"]), 1.3 * pstar, rtol=1e-2)
  assert tuple(nested["b"][0].shape) == (2,)


def test_variable_change_triggers_ready_inference():
  """reference ebm.py:142-162: stale logits must not survive a parameter update."""
  energy = models.KOBE([0, 1, 2], 2, energy_utils.Constant(0.0))
  layer = inference.AnalyticEnergyInference(energy, 1000, initial_seed=[3, 4])
  np.testing.assert_allclose(float(layer.log_partition()), 3 * math.log(2), rtol=1e-6)
  with torch.no_grad():
    energy.post_process[0].kernel.add_(0.7)
  th = np.full(6, 0.7)
  ref = orc.analytic_log_partition(orc.kobe_energy(orc.all_bitstrings(3), 2, th))
  np.testing.assert_allclose(float(layer.log_partition()), ref, rtol=1e-6)
  energy.set_weights([torch.zeros(6)])
  np.testing.assert_allclose(float(layer.entropy()), 3 * math.log(2), rtol=1e-6)


def test_generic_energy_stack_and_mlp_fast_path():
  """AnalyticEnergyInference with a dense stack on raw bits (ebm_utils_test.py:33-47 family)."""
  n = 10
  torch.manual_seed(0)
  lin = [torch.nn.Linear(n, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1),
         utils.Squeeze(-1)]

  class Cast(torch.nn.Module):
    def forward(self, x):
      return x.to(torch.float32)

  fast = models.BitstringEnergy(list(range(n)), [Cast()] + lin)      # Cast is unknown -> generic path
  layer_generic = inference.AnalyticEnergyInference(fast, 1000, initial_seed=[1, 1])
  lp_generic = float(layer_generic.log_partition())

  class FloatLinear(torch.nn.Linear):
    def forward(self, x):
      return super().forward(x.to(torch.float32))

  first = FloatLinear(n, 64)
  first.load_state_dict(lin[0].state_dict())
  mlp = models.BitstringEnergy(list(range(n)), [first] + lin[1:])    # recognised dense stack -> CUDA sweep
  layer_fast = inference.AnalyticEnergyInference(mlp, 1000, initial_seed=[1, 1])
  from qhbmlib.inference import ebm
  assert ebm.energy_descriptor(mlp) is not None and ebm.energy_descriptor(fast) is None
  np.testing.assert_allclose(float(layer_fast.log_partition()), lp_generic, rtol=1e-6)
  np.testing.assert_allclose(layer_fast.distribution.logits_parameter().cpu(),
                             layer_generic.distribution.logits_parameter().cpu(), rtol=1e-5, atol=1e-5)
  probs = inference.probabilities(mlp)
  np.testing.assert_allclose(float(probs.sum()), 1.0, rtol=1e-5)


def test_bernoulli_inference():
  """tests/inference/ebm_test.py:600-790."""
  energy = models.BernoulliEnergy([0, 1, 2])
  layer = inference.BernoulliEnergyInference(energy, 1_000_000, initial_seed=[7, 8])
  th = torch.tensor([-1.5, 0.6, 2.1])
  energy.set_weights([th])
  np.testing.assert_allclose(float(layer.entropy()), orc.bernoulli_entropy(th.numpy()), rtol=1e-5)
  np.testing.assert_allclose(float(layer.log_partition()), orc.bernoulli_log_partition(th.numpy()), rtol=1e-6)
  s = layer.sample(1_000_000)
  assert torch.equal(s, layer.sample(1_000_000))
  p1 = torch.sigmoid(2 * th)
  np.testing.assert_allclose(s.float().mean(0).cpu(), p1, atol=3e-3)
  lp = layer.log_partition()
  (g,) = torch.autograd.grad(lp, energy.trainable_variables)
  np.testing.assert_allclose(g.cpu(), np.tanh(th.numpy()), rtol=2e-2, atol=3e-3)
  layer.seed = None
  assert not torch.equal(layer.sample(1000), layer.sample(1000))
  assert isinstance(layer(None), type(layer.distribution)) and tuple(layer(5).shape) == (5, 3)


# ---------------------------------------------------------------- qhbm_test.py
def _random_qhbm(n, layers, seed, num_samples, ebm_seed=None):
  """tests/test_util.py:70-95: KOBE of order n + HEA + analytic engines."""
  qubits = cq.GridQubit.rect(1, n)
  energy = models.KOBE(list(range(n)), n, energy_utils.RandomUniform(-1.0, 1.0, seed))
  e_infer = inference.AnalyticEnergyInference(energy, num_samples, initial_seed=ebm_seed)
  circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, layers, f"id{seed}"),
                                     energy_utils.RandomUniform(-1.0, 1.0, seed + 1))
  q_infer = inference.AnalyticQuantumInference(circ)
  qhbm = inference.QHBM(e_infer, q_infer)
  return qubits, qhbm


def test_qhbm_expectation_pauli_self_consistency():
  """tests/inference/qhbm_test.py:150-203: QHBM.expectation == weighted average over sampled states,
  same fixed seed, rtol 1e-6."""
  n, num_samples = 4, 100_000
  qubits, qhbm = _random_qhbm(n, 2, 3, num_samples, ebm_seed=[5, 6])
  ops = cq.convert_to_tensor([arch.tfim_ring(qubits), arch.xxz_ring(qubits),
                              cq.PauliSum.from_pauli_strings(cq.Z(qubits[1]) * cq.X(qubits[3]))])
  actual = qhbm.expectation(ops)
  samples = qhbm.e_inference.sample(num_samples)
  bitstrings, _, counts = utils.unique_bitstrings_with_counts(samples)
  exps = qhbm.q_inference.expectation(bitstrings, ops)
  expected = utils.weighted_average(counts, exps)
  np.testing.assert_allclose(actual.detach().cpu(), expected.detach().cpu(), rtol=1e-6, atol=1e-7)
  states, cnt = qhbm.circuits(num_samples)
  assert int(cnt.sum()) == num_samples and len(states) == cnt.shape[0]
  # exact value with exact probabilities (large-sample limit), loose tolerance
  bits = orc.all_bitstrings(n)
  th = qhbm.e_inference.energy.post_process[0].kernel.detach().cpu().numpy()
  p = orc.analytic_probabilities(orc.kobe_energy(bits, n, th))
  circ = qhbm.q_inference.circuit
  oracle_ops = [[(t.coefficient.real, {qubits.index(q): pp for q, pp in t.paulis.items()}) for t in s.terms]
                for s in ops.pauli_sums]
  ref = p @ orc.expectations(circ.gate_table().astype(orc.GATE_DTYPE), n,
                             circ.symbol_values.detach().cpu().numpy(), orc.bitstrings_to_index(bits), oracle_ops)
  np.testing.assert_allclose(actual.detach().cpu(), ref, atol=3e-2)


def test_qhbm_expectation_modular_hamiltonian():
  """tests/inference/qhbm_test.py:210-246."""
  n, num_samples = 3, 100_000
  _, qhbm = _random_qhbm(n, 2, 5, num_samples, ebm_seed=[5, 6])
  _, other = _random_qhbm(n, 1, 9, num_samples, ebm_seed=[7, 8])
  h = other.modular_hamiltonian
  actual = qhbm.expectation(h)
  samples = qhbm.e_inference.sample(num_samples)
  bitstrings, _, counts = utils.unique_bitstrings_with_counts(samples)
  expected = utils.weighted_average(counts, qhbm.q_inference.expectation(bitstrings, h))
  np.testing.assert_allclose(actual.detach().cpu(), expected.detach().cpu(), rtol=1e-6, atol=1e-7)
  assert tuple(actual.shape) == (1,)


# ---------------------------------------------------------------- vqt_loss_test.py / qmhl_loss_test.py
@pytest.mark.parametrize("num_qubits", [1, 2, 3, 4])
def test_vqt_loss_x_rot_closed_form(num_qubits):
  """tests/inference/vqt_loss_test.py:132-205 (1e7 samples, rtol 3e-2 there; 4e6 here)."""
  num_samples = 4_000_000
  energy = models.BernoulliEnergy(list(range(num_qubits)), energy_utils.RandomUniform(-2.0, 2.0, 11))
  e_infer = inference.BernoulliEnergyInference(energy, num_samples, initial_seed=[5, 6])
  qubits = cq.GridQubit.rect(1, num_qubits)
  r_symbols = [cq.Symbol(f"phi_{k}") for k in range(num_qubits)]
  r_circuit = cq.Circuit(cq.rx(s)(q) for s, q in zip(r_symbols, qubits))
  circ = models.DirectQuantumCircuit(r_circuit, energy_utils.RandomUniform(-1, 1, 12))
  qhbm = inference.QHBM(e_infer, inference.AnalyticQuantumInference(circ))
  model_h = qhbm.modular_hamiltonian
  test_h = cq.convert_to_tensor([cq.PauliSum.from_pauli_strings([cq.Y(q) for q in qubits])])
  beta = torch.tensor(1.7, device=DEV)
  thetas = model_h.energy.trainable_variables[0]
  phis = model_h.circuit.trainable_variables[0]
  expected_expectation = torch.sum(torch.tanh(thetas) * torch.sin(phis))
  np.testing.assert_allclose(float(qhbm.expectation(test_h)[0]), float(expected_expectation), rtol=3e-2, atol=2e-3)
  expected_entropy = torch.sum(-thetas * torch.tanh(thetas) + torch.log(2 * torch.cosh(thetas)))
  np.testing.assert_allclose(float(qhbm.e_inference.entropy()), float(expected_entropy), rtol=1e-5)
  loss = inference.vqt(qhbm, test_h, beta)
  np.testing.assert_allclose(float(loss), float(beta * expected_expectation - expected_entropy), rtol=3e-2,
                             atol=3e-3)
  g_thetas, g_phis = torch.autograd.grad(loss, (thetas, phis))
  e_thetas = (1 - torch.tanh(thetas)**2) * (beta * torch.sin(phis) + thetas)
  e_phis = beta * torch.tanh(thetas) * torch.cos(phis)
  np.testing.assert_allclose(g_thetas.cpu(), e_thetas.detach().cpu(), rtol=3e-2, atol=4e-3)
  np.testing.assert_allclose(g_phis.cpu(), e_phis.detach().cpu(), rtol=3e-2, atol=4e-3)


def test_self_vqt_and_self_qmhl_optimum():
  """tests/inference/vqt_loss_test.py:46-83 and qmhl_loss_test.py:48-80: a model against a data QHBM
  with identical weights: the VQT loss is -log Z and QMHL is the entropy; gradients vanish."""
  n, num_samples = 3, 2_000_000
  _, data_qhbm = _random_qhbm(n, 1, 21, num_samples, ebm_seed=[5, 6])
  _, model_qhbm = _random_qhbm(n, 1, 22, num_samples, ebm_seed=[5, 6])
  data_h, model_h = data_qhbm.modular_hamiltonian, model_qhbm.modular_hamiltonian
  with torch.no_grad():
    for pd, pm in zip(data_h.trainable_variables, model_h.trainable_variables):
      pd.copy_(pm)
  loss = inference.vqt(model_qhbm, data_h, torch.tensor(1.0, device=DEV))
  np.testing.assert_allclose(float(loss), -float(data_qhbm.e_inference.log_partition()), rtol=2e-2, atol=3e-3)
  grads = torch.autograd.grad(loss, model_h.trainable_variables, allow_unused=True)
  for g in grads:
    assert g is None or float(g.abs().max()) < 2e-2
  loss = inference.qmhl(data.QHBMData(data_qhbm), model_qhbm)
  np.testing.assert_allclose(float(loss), float(model_qhbm.e_inference.entropy()), rtol=2e-2, atol=3e-3)
  grads = torch.autograd.grad(loss, model_h.trainable_variables, allow_unused=True)
  for g in grads:
    assert g is None or float(g.abs().max()) < 2e-2


@pytest.mark.parametrize("num_qubits", [1, 2, 3])
def test_qmhl_loss_x_rot_closed_form(num_qubits):
  """tests/inference/qmhl_loss_test.py:136-272 (1e6 samples, rtol 2e-2)."""
  num_samples = 1_000_000
  energy = models.BernoulliEnergy(list(range(num_qubits)), energy_utils.RandomUniform(0.25, 1.0, 31))
  e_infer = inference.BernoulliEnergyInference(energy, num_samples, initial_seed=[5, 6])
  qubits = cq.GridQubit.rect(1, num_qubits)
  r_circuit = cq.Circuit(cq.rx(cq.Symbol(f"phi_{k}"))(q) for k, q in enumerate(qubits))
  circ = models.DirectQuantumCircuit(r_circuit, energy_utils.RandomUniform(math.pi / 4, math.pi, 32))
  qhbm = inference.QHBM(e_infer, inference.AnalyticQuantumInference(circ))
  model = qhbm.modular_hamiltonian
  thetas, phis = model.energy.trainable_variables[0], model.circuit.trainable_variables[0]
  rng = np.random.default_rng(33)
  alphas = rng.uniform(-math.pi, math.pi, num_qubits)
  y_rot = cq.Circuit(cq.ry(float(a))(q) for a, q in zip(alphas, qubits))
  data_q = inference.AnalyticQuantumInference(models.DirectQuantumCircuit(y_rot))
  data_probs = rng.uniform(0, 1, num_qubits)
  samples = torch.tensor((rng.random((num_samples, num_qubits)) < (1 - data_probs)).astype(np.int8), device=DEV)

  class FixedData(data.QuantumData):

    def expectation(self, observable):
      return torch.mean(data_q.expectation(samples, observable))

  loss = inference.qmhl(FixedData(), qhbm)
  t, p = thetas.detach().cpu().numpy(), phis.detach().cpu().numpy()
  exp_expect = np.sum(t * (2 * data_probs - 1) * np.cos(alphas) * np.cos(p))
  exp_lp = np.sum(np.log(2 * np.cosh(t)))
  np.testing.assert_allclose(float(loss), exp_expect + exp_lp, rtol=2e-2, atol=3e-3)
  g_t, g_p = torch.autograd.grad(loss, (thetas, phis))
  np.testing.assert_allclose(g_t.cpu(), (2 * data_probs - 1) * np.cos(alphas) * np.cos(p) + np.tanh(t), rtol=2e-2,
                             atol=4e-3)
  np.testing.assert_allclose(g_p.cpu(), -t * (2 * data_probs - 1) * np.cos(alphas) * np.sin(p), rtol=2e-2,
                             atol=4e-3)


def test_config2_12q_vqt_kobe2_analytic():
  """BASELINE config 2: 12-qubit TFIM VQT loss + gradient, KOBE order 2, AnalyticEnergyInference.
  Checked against the oracle evaluated on the very bitstrings/counts the engine sampled (n >= 11,
  so the reference's bit-column permutation is exercised)."""
  n, num_samples = 12, 500
  qubits = cq.GridQubit.rect(1, n)
  energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.1, 4))
  e_infer = inference.AnalyticEnergyInference(energy, num_samples, initial_seed=[3, 4])
  circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 2, "c2"),
                                     energy_utils.RandomUniform(-1, 1, 11))
  qhbm = inference.QHBM(e_infer, inference.AnalyticQuantumInference(circ, grad_mode="exact"))
  h = cq.convert_to_tensor([arch.tfim_ring(qubits)])
  beta = torch.tensor(0.8, device=DEV)
  loss = inference.vqt(qhbm, h, beta)
  g_theta, g_phi = torch.autograd.grad(loss, (energy.post_process[0].kernel, circ.trainable_variables[0]))
  # reproduce: with a fixed seed, the two preface calls inside vqt() draw the same samples
  samples = e_infer.sample(num_samples).cpu().numpy()
  y, _, counts = orc.unique_bitstrings_with_counts(samples)
  th = energy.post_process[0].kernel.detach().cpu().numpy()
  phi = circ.symbol_values.detach().cpu().numpy()
  gates = circ.gate_table().astype(orc.GATE_DTYPE)
  idx = orc.bitstrings_to_index(y)  # includes the n >= 11 column permutation
  w = counts / counts.sum()
  e_h, g_h = orc.batch_expectation_and_gradient(gates, n, phi, idx, [orc.tfim_ring(n)], (0.8 * w)[:, None])
  energies = orc.kobe_energy(y, 2, th)
  f = 0.8 * e_h[:, 0] - energies
  all_e = orc.kobe_energy(orc.all_bitstrings(n), 2, th)
  ref_loss = float(w @ f) - orc.analytic_log_partition(all_e)
  # tolerances: 1e-5 relative with the absolute floor 1e-6 * (scale of the summed terms): the TFIM ring has
  # sum|coeff| = 2n = 24, weighted by beta = 0.8
  floor = 1e-6 * 0.8 * 24
  np.testing.assert_allclose(float(loss), ref_loss, rtol=1e-5, atol=floor)
  np.testing.assert_allclose(g_phi.cpu(), g_h.sum(0), rtol=1e-5, atol=floor)
  ref_gth = orc.expectation_score_gradient(counts, f, orc.parity_features(y, 2), np.zeros(len(th)), 1.0)
  np.testing.assert_allclose(g_theta.cpu(), ref_gth, rtol=1e-5, atol=floor + 1e-6 * np.abs(energies).max())


def test_single_observable_jacobian_in_forward_matches_resimulation():
  """One observable: the autograd forward computes d<H>_u/d phi per state (fused run) and the backward
  is a contraction; it must equal the re-simulating backward used for several observables, and a
  no-grad call must not pay for it."""
  from qhbmlib.inference import qnn as qnn_mod
  n = 6
  qubits = cq.GridQubit.rect(1, n)
  circuit = arch.get_hardware_efficient_model_unitary(qubits, 2, "jac")
  qnn = models.DirectQuantumCircuit(circuit, initializer=energy_utils.RandomUniform(-1, 1, seed=3))
  q_infer = inference.AnalyticQuantumInference(qnn, grad_mode="exact")
  ham = cq.convert_to_tensor([arch.xxz_ring(qubits)])
  rng = np.random.default_rng(4)
  bitstrings = _bits(rng.integers(0, 2, (40, n)).tolist())
  weights = torch.tensor(rng.uniform(-1, 1, (40, 1)), dtype=torch.float32, device=DEV)
  grads, outs = [], []
  for flag in (True, False):
    qnn_mod._PlanHolder.jacobian_in_forward = flag
    try:
      for p in qnn.parameters():
        p.grad = None
      out = q_infer.expectation(bitstrings, ham)
      (out * weights).sum().backward()
      outs.append(out.detach().cpu().numpy())
      grads.append(qnn.trainable_variables[0].grad.cpu().numpy().copy())
    finally:
      qnn_mod._PlanHolder.jacobian_in_forward = True
  np.testing.assert_allclose(outs[0], outs[1], rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(grads[0], grads[1], rtol=1e-4, atol=1e-4)
  assert np.abs(grads[0]).max() > 0.1
  with torch.no_grad():
    quiet = q_infer.expectation(bitstrings, ham)
  np.testing.assert_allclose(quiet.cpu().numpy(), outs[0], rtol=1e-5, atol=1e-5)


def test_expectation_nested_structure_and_energy_gradient():
  """tests/inference/ebm_test.py:test_expectation_finite_difference: a function with nested outputs
  (scalar, vector, a list holding a tensor that itself depends on the energy variable); value and
  d/d theta of the summed outputs against the exact distribution (the reference compares with
  sampled finite differences)."""
  num_bits, order, num_samples = 3, 2, int(2e6)
  energy = models.KOBE(list(range(num_bits)), order, energy_utils.RandomUniform(1, 2, seed=7))
  energy.to(DEV)
  e_infer = inference.AnalyticEnergyInference(energy, num_samples, initial_seed=[5, 6])
  theta = energy.trainable_variables[0]
  gen = torch.Generator().manual_seed(3)
  scalar_var = torch.nn.Parameter((1 + torch.rand((), generator=gen)).to(DEV))
  dense = torch.nn.Linear(num_bits, 5).to(DEV)
  with torch.no_grad():
    dense.weight.copy_((1 + torch.rand(5, num_bits, generator=gen)).to(DEV))
    dense.bias.copy_((1 + torch.rand(5, generator=gen)).to(DEV))

  def f(bitstrings):
    reduced = bitstrings.float().sum(1)
    return [scalar_var * reduced, dense(bitstrings.float()), [torch.einsum("i,j->ij", reduced, theta)]]

  actual = e_infer.expectation(f)
  assert tuple(actual[0].shape) == () and tuple(actual[1].shape) == (5,) and tuple(actual[2][0].shape) == (6,)
  total = actual[0] + actual[1].sum() + actual[2][0].sum()
  (g_actual,) = torch.autograd.grad(total, theta)

  all_bits = _bits(list(itertools.product([0, 1], repeat=num_bits)))
  probs = torch.softmax(-energy(all_bits).double(), 0)
  exact = [torch.tensordot(probs, v.double(), dims=([0], [0])) for v in (f(all_bits)[0], f(all_bits)[1], f(all_bits)[2][0])]
  exact_total = exact[0] + exact[1].sum() + exact[2].sum()
  (g_exact,) = torch.autograd.grad(exact_total, theta)
  for a, e in zip((actual[0], actual[1], actual[2][0]), exact):
    assert float(e.abs().min()) > 1e-3
    np.testing.assert_allclose(a.detach().cpu().numpy(), e.detach().cpu().numpy(), rtol=5e-3)
  assert float(g_exact.abs().min()) > 1e-3
  np.testing.assert_allclose(g_actual.cpu().numpy(), g_exact.cpu().numpy(), rtol=3e-2, atol=5e-3)


def test_qhbm_circuit_param_update():
  """tests/inference/qhbm_test.py:test_circuit_param_update: expectations follow in-place updates of
  the circuit variables (compiled plans keep no symbol values)."""
  n = 3
  qubits, qhbm = _random_qhbm(n, 2, 11, 50_000, ebm_seed=[3, 4])
  ops = cq.convert_to_tensor([arch.tfim_ring(qubits)])
  circ = qhbm.q_inference.circuit
  bits = orc.all_bitstrings(n)
  oracle_ops = [[(t.coefficient.real, {qubits.index(q): pp for q, pp in t.paulis.items()}) for t in s.terms]
                for s in ops.pauli_sums]
  bitstrings = _bits(bits.tolist())

  def reference():
    return orc.expectations(circ.gate_table().astype(orc.GATE_DTYPE), n, circ.symbol_values.detach().cpu().numpy(),
                            orc.bitstrings_to_index(bits), oracle_ops)

  before = qhbm.q_inference.expectation(bitstrings, ops).detach().cpu().numpy()
  np.testing.assert_allclose(before, reference(), rtol=1e-5, atol=1e-5)
  with torch.no_grad():
    for p in circ.parameters():
      p.add_(0.37)
  after = qhbm.q_inference.expectation(bitstrings, ops).detach().cpu().numpy()
  np.testing.assert_allclose(after, reference(), rtol=1e-5, atol=1e-5)
  assert np.abs(after - before).max() > 1e-2
  # and the QHBM-level estimate moves with it
  est = qhbm.expectation(ops).detach().cpu().numpy()
  th = qhbm.e_inference.energy.post_process[0].kernel.detach().cpu().numpy()
  p = orc.analytic_probabilities(orc.kobe_energy(bits, n, th))
  np.testing.assert_allclose(est, p @ reference(), atol=3e-2)


def test_vqt_training_approaches_free_energy():
  """End-to-end training loop (the reference's baselines/train.py in miniature): minimising the VQT loss
  of a 5-qubit TFIM at beta = 1 with Adam drives it towards the exact free energy -log tr exp(-beta H),
  which bounds it from below."""
  n, beta_val = 5, 1.0
  qubits = cq.GridQubit.rect(1, n)
  energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.1, 4))
  e_infer = inference.AnalyticEnergyInference(energy, 20_000, initial_seed=[1, 2])
  circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 3, "t"),
                                     energy_utils.RandomUniform(-0.2, 0.2, 5))
  qhbm = inference.QHBM(e_infer, inference.AnalyticQuantumInference(circ))
  ham = arch.tfim_ring(qubits)
  h = cq.convert_to_tensor([ham])
  beta = torch.tensor(beta_val, device=DEV)
  # exact free energy from the dense Hamiltonian (float64 on the host)
  dim = 1 << n
  pauli = {"X": np.array([[0, 1], [1, 0]], complex), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1.0 + 0j, -1.0])}
  dense = np.zeros((dim, dim), complex)
  for t in ham.terms:
    m = np.array([[1.0 + 0j]])
    for q in qubits:
      m = np.kron(m, pauli[t.paulis[q]] if q in t.paulis else np.eye(2))
    dense += t.coefficient.real * m
  free_energy = -np.log(np.exp(-beta_val * np.linalg.eigvalsh(dense)).sum())
  opt = torch.optim.Adam(qhbm.trainable_variables, lr=0.05)
  losses = []
  for _ in range(150):
    opt.zero_grad()
    loss = inference.vqt(qhbm, h, beta)
    loss.backward()
    opt.step()
    losses.append(float(loss.detach()))
  first, last = np.mean(losses[:5]), np.mean(losses[-10:])
  assert last < first - 1.0, (first, last)
  assert last > free_energy - 0.05            # variational bound (sampling noise allowance)
  assert last - free_energy < 0.35 * (first - free_energy), (first, last, free_energy)


def test_qmhl_training_raises_fidelity_with_target_state():
  """Quantum modular Hamiltonian learning in miniature (reference qmhl_loss.py + qhbm_utils.py): a model
  QHBM trained on samples of a target QHBM with the QMHL loss ends up with a thermal state close to the
  target's (fidelity from the dense metrics)."""
  n, num_samples = 3, 50_000
  _, target = _random_qhbm(n, 2, 31, num_samples, ebm_seed=[9, 10])
  _, model = _random_qhbm(n, 2, 32, num_samples, ebm_seed=[11, 12])
  # the model must be able to express the target: same ansatz family, different starting point
  sigma = inference.density_matrix(target.modular_hamiltonian)
  f0 = float(inference.fidelity(model.modular_hamiltonian, sigma))
  target_data = data.QHBMData(target)
  opt = torch.optim.Adam(model.trainable_variables, lr=0.05)
  losses = []
  for _ in range(200):
    opt.zero_grad()
    loss = inference.qmhl(target_data, model)
    loss.backward()
    opt.step()
    losses.append(float(loss.detach()))
  f1 = float(inference.fidelity(model.modular_hamiltonian, sigma))
  assert np.mean(losses[-10:]) < np.mean(losses[:5])
  assert f1 > max(f0 + 0.05, 0.9), (f0, f1)
