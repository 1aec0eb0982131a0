"""Utilities for metrics on BitstringEnergy (mirror of reference inference/ebm_utils.py)."""
import torch

from qhbmlib import engine
from qhbmlib import utils


def probabilities(input_energy):
  """Exact p(x) for every bitstring, rows in big-endian counting order (2^n softmax)."""
  n = input_energy.num_bits
  dev = next(input_energy.parameters()).device if list(input_energy.parameters()) else torch.device("cuda")
  rows = torch.arange(1 << n, dtype=torch.int64, device=dev)
  all_bitstrings = engine.unpack_bits(rows, n, utils._natural_shifts(n))
  return torch.softmax(-input_energy(all_bitstrings), 0)
