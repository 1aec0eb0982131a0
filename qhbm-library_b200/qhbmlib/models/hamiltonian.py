"""Spectral (modular) Hamiltonian K = U diag(E) U^dagger (mirror of reference hamiltonian.py)."""
import torch

from qhbmlib import circuits as cq
from qhbmlib.models import energy as energy_lib


class Hamiltonian(torch.nn.Module):
  """Eigenvalues from a BitstringEnergy, eigenvectors from a QuantumCircuit."""

  def __init__(self, input_energy, input_circuit, name=None):
    super().__init__()
    self.name = name
    if input_energy.num_bits != len(input_circuit.qubits):
      raise ValueError("`input_energy` and `input_circuit` must act on the same number of bits.")
    self.energy = input_energy
    self.circuit = input_circuit
    self.circuit_dagger = input_circuit**-1
    self.operator_shards = None
    if isinstance(self.energy, energy_lib.PauliMixin):
      self.operator_shards = cq.convert_to_tensor(self.energy.operator_shards(self.circuit.qubits))

  @property
  def trainable_variables(self):
    seen, out = set(), []
    for p in list(self.energy.parameters()) + list(self.circuit.parameters()):
      if id(p) not in seen and p.requires_grad:
        seen.add(id(p))
        out.append(p)
    return out
