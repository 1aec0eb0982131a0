"""Utilities for metrics on Hamiltonian (mirror of reference inference/qhbm_utils.py).

These are dense 2^n x 2^n contractions for logging small models; they run as library GEMMs /
eigensolvers on the unitary assembled by the sweep kernels."""
import torch

from qhbmlib.inference import ebm_utils
from qhbmlib.inference import qnn_utils


def density_matrix(model):
  """rho = U diag(p) U^dagger of a modular Hamiltonian (reference qhbm_utils.py:24-61)."""
  unitary_matrix = qnn_utils.unitary(model.circuit)
  probabilities = ebm_utils.probabilities(model.energy).detach().to(unitary_matrix.device)
  return (unitary_matrix * probabilities.to(torch.complex64).unsqueeze(0)) @ unitary_matrix.conj().transpose(0, 1)


def fidelity(model, sigma):
  """(tr sqrt(sqrt(rho) sigma sqrt(rho)))^2 with rho the thermal state of `model`: the
  eigenvalues of omega = sqrt(P) U^dagger sigma U sqrt(P) are found with a Hermitian solver
  (reference qhbm_utils.py:64-116)."""
  u_phi = qnn_utils.unitary(model.circuit)
  sigma = torch.as_tensor(sigma).to(device=u_phi.device, dtype=torch.complex64)
  k_theta = ebm_utils.probabilities(model.energy).detach().to(u_phi.device)
  sqrt_k = torch.sqrt(k_theta).to(torch.complex64)
  omega = sqrt_k.unsqueeze(1) * (u_phi.conj().transpose(0, 1) @ sigma @ u_phi) * sqrt_k.unsqueeze(0)
  omega = 0.5 * (omega + omega.conj().transpose(0, 1))
  d_omega = torch.linalg.eigvalsh(omega).clamp_min(0.0)
  return torch.sum(torch.sqrt(d_omega))**2
