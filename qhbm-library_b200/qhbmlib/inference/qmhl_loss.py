"""QMHL loss (mirror of /root/reference/qhbmlib/inference/qmhl_loss.py)."""


def qmhl(data, input_qhbm):
  """Quantum cross entropy between the data state and the model: <K_model>_data + log Z_model."""
  return data.expectation(input_qhbm.modular_hamiltonian) + input_qhbm.e_inference.log_partition()
