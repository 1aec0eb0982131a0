"""Utilities for metrics on QuantumCircuit (mirror of reference inference/qnn_utils.py)."""
import torch

from qhbmlib import circuits as cq
from qhbmlib import engine


def unitary(input_circuit):
  """complex64 [2^n, 2^n] unitary of `input_circuit.pqc` at the current symbol values
  (reference qnn_utils.py:23-33, tfq.layers.Unitary): column k is the final state of basis
  input k, all 2^n columns simulated in one batch by the sweep kernels."""
  qubits = input_circuit.qubits
  n = len(qubits)
  if n > 14:
    raise ValueError("unitary() materialises 4^n amplitudes; n must be <= 14")
  cache = input_circuit.__dict__.setdefault("_unitary_plan", {})
  if "plan" not in cache:
    terms, offsets = cq.convert_to_tensor([cq.PauliSum.from_pauli_strings(cq.Z(qubits[0]))]).tables(qubits)
    cache["plan"] = engine.ExpectationPlan(input_circuit.gate_table(), n, len(input_circuit.symbol_names), terms,
                                           offsets, False)
  values = input_circuit.symbol_values.detach().float()
  if not values.is_cuda:
    values = values.to("cuda")
  basis = torch.arange(1 << n, dtype=torch.int64, device=values.device)
  states = cache["plan"].final_states(basis, values.contiguous())
  return states.transpose(0, 1).contiguous()
