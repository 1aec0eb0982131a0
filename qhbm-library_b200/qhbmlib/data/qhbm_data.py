"""Quantum data defined by a QHBM (mirror of reference data/qhbm_data.py)."""
import torch

from qhbmlib.data import quantum_data


class QHBMData(quantum_data.QuantumData):

  def __init__(self, input_qhbm):
    self.qhbm = input_qhbm

  def expectation(self, observable):
    return torch.squeeze(self.qhbm.expectation(observable), 0)
