"""Circuit ansaetze and lattice Hamiltonians used by the reference's tests and driver.

These are spec restatements, written against `qhbmlib.circuits`:
  * hardware-efficient ansatz: /root/reference/tests/test_util.py:25-67 (== baselines/pqc.py:21-63)
  * 1-D transverse-field Ising ring: /root/reference/baselines/train.py:46-58
  * XXZ ring: synthetic benchmark Hamiltonian
"""
from qhbmlib import circuits as cq


def get_xz_rotation(q, a, b):
  return cq.Circuit(cq.X(q)**a, cq.Z(q)**b)


def get_cz_exp(q0, q1, a):
  return cq.Circuit(cq.CZPowGate(exponent=a)(q0, q1))


def get_xz_rotation_layer(qubits, layer_num, name):
  layer = cq.Circuit()
  for k, q in enumerate(qubits):
    layer += get_xz_rotation(q, cq.Symbol(f"sx_{name}_{layer_num}_{k}"), cq.Symbol(f"sz_{name}_{layer_num}_{k}"))
  return layer


def get_cz_exp_layer(qubits, layer_num, name):
  layer = cq.Circuit()
  for k, (q0, q1) in enumerate(zip(qubits[::2], qubits[1::2])):
    layer += get_cz_exp(q0, q1, cq.Symbol(f"sc_{name}_{layer_num}_{2 * k}"))
  shifted = qubits[1:]
  for k, (q0, q1) in enumerate(zip(shifted[::2], shifted[1::2])):
    layer += get_cz_exp(q0, q1, cq.Symbol(f"sc_{name}_{layer_num}_{2 * k + 1}"))
  return layer


def get_hardware_efficient_model_unitary(qubits, num_layers, name):
  """X^sx Z^sz on every qubit, then CZ^sc on even and odd nearest-neighbour pairs, per layer."""
  circuit = cq.Circuit()
  for layer in range(num_layers):
    circuit += get_xz_rotation_layer(qubits, layer, name)
    if len(qubits) > 1:
      circuit += get_cz_exp_layer(qubits, layer, name)
  return circuit


def tfim_ring(qubits, bias=1.0):
  """H = -sum_i Z_i Z_{i+1 mod n} - bias sum_i X_i."""
  n = len(qubits)
  h = cq.PauliSum()
  for i in range(n):
    h -= bias * cq.X(qubits[i])
  for i in range(n):
    h -= cq.Z(qubits[i]) * cq.Z(qubits[(i + 1) % n])
  return h


def xxz_ring(qubits, delta=0.5):
  """H = sum_i X_i X_{i+1} + Y_i Y_{i+1} + delta Z_i Z_{i+1} on a ring."""
  n = len(qubits)
  h = cq.PauliSum()
  for i in range(n):
    a, b = qubits[i], qubits[(i + 1) % n]
    h += cq.X(a) * cq.X(b)
    h += cq.Y(a) * cq.Y(b)
    h += delta * (cq.Z(a) * cq.Z(b))
  return h
