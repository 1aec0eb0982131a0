"""Building blocks of energy functions (mirror of /root/reference/qhbmlib/models/energy_utils.py)."""
import itertools

import torch


def check_bits(bits):
  """Bit labels must be pairwise distinct."""
  if len(set(bits)) != len(bits):
    raise ValueError("All entries of `bits` must be unique.")
  return bits


def check_order(order):
  """A parity order is a positive integer."""
  if not isinstance(order, int):
    raise TypeError("`order` must be an integer.")
  if order <= 0:
    raise ValueError("`order` must be greater than zero.")
  return order


class RandomUniform:
  """Keras-style initializer: U(minval, maxval) with an optional seed."""

  def __init__(self, minval=-0.05, maxval=0.05, seed=None):
    self.minval, self.maxval, self.seed = minval, maxval, seed

  def __call__(self, shape):
    gen = None
    if self.seed is not None:
      gen = torch.Generator().manual_seed(int(self.seed))
    return torch.rand(tuple(shape), generator=gen) * (self.maxval - self.minval) + self.minval


class Constant:

  def __init__(self, value=0.0):
    self.value = value

  def __call__(self, shape):
    return torch.full(tuple(shape), float(self.value))


class RandomNormal:

  def __init__(self, mean=0.0, stddev=0.05, seed=None):
    self.mean, self.stddev, self.seed = mean, stddev, seed

  def __call__(self, shape):
    gen = None
    if self.seed is not None:
      gen = torch.Generator().manual_seed(int(self.seed))
    return torch.randn(tuple(shape), generator=gen) * self.stddev + self.mean


class SpinsFromBitstrings(torch.nn.Module):
  """bit 0 -> spin +1, bit 1 -> spin -1 (reference energy_utils.py:39-52)."""

  def forward(self, inputs):
    return (1 - 2 * inputs).to(torch.float32)


class VariableDot(torch.nn.Module):
  """Dot product of the last axis with a trainable vector `kernel` (energy_utils.py:55-81).
  The vector is created on first use (Keras `build`) or by `build(input_shape)`."""

  def __init__(self, initializer=None):
    super().__init__()
    self._initializer = initializer if initializer is not None else RandomUniform()
    self.kernel = None

  def build(self, input_shape, device=None):
    if self.kernel is None:
      self.kernel = torch.nn.Parameter(self._initializer((int(input_shape[-1]),)).to(torch.float32).to(
          device if device is not None else "cpu"))

  def forward(self, inputs):
    self.build(inputs.shape, inputs.device)
    if self.kernel.device != inputs.device:
      self.kernel.data = self.kernel.data.to(inputs.device)
    return torch.sum(inputs * self.kernel, -1)


class Parity(torch.nn.Module):
  """Products of spins over every index group of size 1..order, groups in
  `itertools.combinations` order (energy_utils.py:84-110)."""

  def __init__(self, bits, order):
    super().__init__()
    bits = check_bits(bits)
    order = check_order(order)
    groups = []
    for size in range(1, order + 1):
      groups.extend(itertools.combinations(range(len(bits)), size))
    self.indices = [list(g) for g in groups]
    self.num_terms = len(groups)
    self._num_bits = len(bits)
    member = torch.zeros((self.num_terms, len(bits)), dtype=torch.float32)
    for t, g in enumerate(groups):
      member[t, list(g)] = 1.0
    self.register_buffer("_member", member, persistent=False)

  def masks(self):
    """Index-bit mask of each group for the CUDA energy kernels (column j <-> bit n-1-j).  The groups never
    change after construction, so the list is built once (it is asked for on every inference call)."""
    cached = self.__dict__.get("_masks_cache")
    if cached is None:
      n = self._num_bits
      cached = [sum(1 << (n - 1 - j) for j in g) for g in self.indices]
      self.__dict__["_masks_cache"] = cached
    return cached

  def forward(self, inputs):
    """Inputs are spins (+-1, any numeric dtype).  prod over a group = (-1)^{#(-1) in the group},
    evaluated as one [N, n] x [n, terms] product instead of the reference's per-term scatter loop."""
    neg = (inputs < 0).to(torch.float32)
    odd = torch.remainder(neg @ self._member.to(inputs.device).t(), 2.0)
    return 1.0 - 2.0 * odd
