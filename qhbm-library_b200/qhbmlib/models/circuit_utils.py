"""Helpers for circuit models (mirror of /root/reference/qhbmlib/models/circuit_utils.py)."""
from qhbmlib import circuits as cq


def bit_circuit(qubits, name="bit_circuit"):
  """X**bit on every qubit, with one symbol per qubit: the basis-state injector."""
  circuit = cq.Circuit()
  for k, q in enumerate(qubits):
    circuit += cq.X(q)**cq.Symbol(f"{name}_bit_{k}")
  return circuit
