"""VQT loss (mirror of /root/reference/qhbmlib/inference/vqt_loss.py)."""
import torch


def vqt(input_qhbm, target_hamiltonian, beta):
  """beta <H>_rho - S(rho), differentiable w.r.t. the QHBM's variables under torch autograd."""

  def f_vqt(bitstrings):
    h_expectations = torch.squeeze(input_qhbm.q_inference.expectation(bitstrings, target_hamiltonian), 1)
    beta_h_expectations = beta * h_expectations
    energies = input_qhbm.e_inference.energy(bitstrings).detach()
    return beta_h_expectations - energies

  average_expectation = input_qhbm.e_inference.expectation(f_vqt)
  current_partition = input_qhbm.e_inference.log_partition().detach()
  return average_expectation - current_partition
