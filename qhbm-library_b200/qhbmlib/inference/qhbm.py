"""QHBM = EBM eigenvalues + QNN eigenvectors (mirror of /root/reference/qhbmlib/inference/qhbm.py)."""
import functools

import torch

from qhbmlib import utils
from qhbmlib.models import hamiltonian


class QHBM(torch.nn.Module):
  """Inference on rho = sum_x p_theta(x) U_phi|x><x|U_phi^dagger through its canonical ensemble."""

  def __init__(self, input_ebm, input_qnn, name=None):
    super().__init__()
    self.name = name
    self._e_inference = input_ebm
    self._q_inference = input_qnn
    self._modular_hamiltonian = hamiltonian.Hamiltonian(self.e_inference.energy, self.q_inference.circuit)

  @property
  def e_inference(self):
    return self._e_inference

  @property
  def q_inference(self):
    return self._q_inference

  @property
  def modular_hamiltonian(self):
    return self._modular_hamiltonian

  @property
  def trainable_variables(self):
    return self._modular_hamiltonian.trainable_variables

  def circuits(self, num_samples):
    """(CircuitBatch of U_phi|x_i> for the unique sampled x_i, int32 counts)."""
    if hasattr(self.e_inference, "unique_samples"):
      bitstrings, _, counts = self.e_inference.unique_samples(num_samples)
    else:
      bitstrings, _, counts = utils.unique_bitstrings_with_counts(self.e_inference.sample(num_samples))
    states = self.q_inference.circuit(bitstrings)
    return states, counts

  def expectation(self, observables):
    """[n_ops] sample-averaged expectation of each observable against rho (qhbm.py:124-147)."""
    return self.e_inference.expectation(functools.partial(self.q_inference.expectation, observables=observables))
