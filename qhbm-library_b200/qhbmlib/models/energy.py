"""Energy functions over bitstrings (mirror of /root/reference/qhbmlib/models/energy.py)."""
import abc

import torch

from qhbmlib import circuits as cq
from qhbmlib.models import energy_utils


class BitstringEnergy(torch.nn.Module):
  """E(x): a stack of layers mapping int8 bitstrings [N, n] to scalars [N]; defines the EBM
  p(x) = exp(-E(x)) / Z (reference energy.py:26-87)."""

  def __init__(self, bits, energy_layers, name=None):
    super().__init__()
    self.name = name
    self._bits = energy_utils.check_bits(bits)
    self._energy_layers = torch.nn.ModuleList(energy_layers)

  @property
  def num_bits(self):
    return len(self._bits)

  @property
  def bits(self):
    return self._bits

  @property
  def energy_layers(self):
    return list(self._energy_layers)

  def build(self, input_shape, device=None):
    """Creates lazily-shaped variables (Keras `build`): runs a dummy batch through the stack."""
    dev = device
    if dev is None:
      params = list(self.parameters())
      dev = params[0].device if params else "cpu"
    with torch.no_grad():
      self.forward(torch.zeros((1, int(input_shape[-1])), dtype=torch.int8, device=dev))

  @property
  def variables(self):
    return list(self.parameters())

  @property
  def trainable_variables(self):
    return [p for p in self.parameters() if p.requires_grad]

  def set_weights(self, weights):
    params = list(self.parameters())
    if len(params) != len(weights):
      raise ValueError(f"expected {len(params)} weight tensors, got {len(weights)}")
    with torch.no_grad():
      for p, w in zip(params, weights):
        p.copy_(torch.as_tensor(w, dtype=p.dtype).reshape(p.shape))

  def forward(self, inputs):
    x = inputs
    for layer in self._energy_layers:
      x = layer(x)
    return x


class PauliMixin(abc.ABC):
  """Adds a Pauli-Z representation: E = post_process(<Z-string shards>) (energy.py:90-120)."""

  @property
  @abc.abstractmethod
  def post_process(self):
    raise NotImplementedError()

  @abc.abstractmethod
  def operator_shards(self, qubits):
    raise NotImplementedError()

  def operator_expectation(self, expectation_shards):
    x = expectation_shards
    for layer in self.post_process:
      x = layer(x)
    return x


class BernoulliEnergy(BitstringEnergy, PauliMixin):
  """Independent spins in a field: E(b) = sum_i (1 - 2 b_i) theta_i (energy.py:123-167)."""

  def __init__(self, bits, initializer=None, name=None):
    post = [energy_utils.VariableDot(initializer=initializer)]
    super().__init__(bits, [energy_utils.SpinsFromBitstrings()] + post, name)
    self._post_process = post
    post[0].build([None, len(bits)])

  @property
  def logits(self):
    """log p(1)/p(0) per bit = 2 theta."""
    return 2 * self.post_process[0].kernel

  @property
  def post_process(self):
    return self._post_process

  def operator_shards(self, qubits):
    return [cq.PauliSum.from_pauli_strings(cq.Z(q)) for q in qubits]

  def kernel_descriptor(self):
    """(kind, masks, theta) for the CUDA energy kernels."""
    n = self.num_bits
    return "bernoulli", [1 << (n - 1 - j) for j in range(n)], self.post_process[0].kernel


class KOBE(BitstringEnergy, PauliMixin):
  """K-th order binary energy: E = sum_t theta_t prod_{i in t} (1 - 2 b_i) (energy.py:170-209)."""

  def __init__(self, bits, order, initializer=None, name=None):
    parity = energy_utils.Parity(bits, order)
    post = [energy_utils.VariableDot(initializer=initializer)]
    super().__init__(bits, [energy_utils.SpinsFromBitstrings(), parity] + post, name)
    self._num_terms = parity.num_terms
    self._indices = parity.indices
    self._parity = parity
    self._post_process = post
    post[0].build([None, parity.num_terms])

  @property
  def post_process(self):
    return self._post_process

  def operator_shards(self, qubits):
    ops = []
    for group in self._indices:
      string = cq.PauliString(1.0, {qubits[loc]: "Z" for loc in group})
      ops.append(cq.PauliSum.from_pauli_strings(string))
    return ops

  def kernel_descriptor(self):
    return "kobe", self._parity.masks(), self.post_process[0].kernel
