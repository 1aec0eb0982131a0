"""Inference on energy functions (mirror of /root/reference/qhbmlib/inference/ebm.py).

Same template as the reference: every public method first runs `_preface_inference` (first-call
initialisation, seed advance unless the user fixed it, re-`_ready_inference` when a tracked
variable changed).  The data-parallel parts run in libqhbm_b200.so: the exhaustive 2^n
logits / logsumexp / entropy sweep, categorical and Bernoulli sampling, first-occurrence
dedup and the count-weighted reductions.  Gradients follow the reference's estimators
(ebm.py:282-325 and 396-415) expressed as surrogate terms for torch autograd.
"""
import abc
import functools
import math
import secrets

import torch

from qhbmlib import _native as nat
from qhbmlib import engine
from qhbmlib import utils


def preface_inference(f):
  """Decorator: run `self._preface_inference()` before the wrapped method."""

  @functools.wraps(f)
  def wrapper(self, *args, **kwargs):
    self._preface_inference()  # pylint: disable=protected-access
    return f(self, *args, **kwargs)

  return wrapper


def map_structure(fn, *structures):
  """tf.nest.map_structure for tensors nested in lists / tuples / dicts."""
  first = structures[0]
  if isinstance(first, dict):
    return {k: map_structure(fn, *[s[k] for s in structures]) for k in first}
  if isinstance(first, (list, tuple)):
    out = [map_structure(fn, *items) for items in zip(*structures)]
    return type(first)(out) if not hasattr(first, "_fields") else type(first)(*out)
  return fn(*structures)


_MASK64 = (1 << 64) - 1


def _splitmix(x):
  x = (x + 0x9E3779B97F4A7C15) & _MASK64
  z = x
  z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
  z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
  return z ^ (z >> 31)


def sanitize_seed(seed):
  """[2] int64 tensor: the given pair, or a fresh random one (tfp.random.sanitize_seed)."""
  if seed is None:
    return torch.tensor([secrets.randbits(31), secrets.randbits(31)], dtype=torch.int64)
  t = torch.as_tensor(seed).reshape(-1).to(torch.int64).cpu()
  if t.numel() == 1:
    t = torch.stack([torch.zeros((), dtype=torch.int64), t[0]])
  if t.numel() != 2:
    raise ValueError("seed must hold one or two integers")
  return t.clone()


def split_seed(seed):
  """Two statistically independent child seeds (counterpart of tfp.random.split_seed)."""
  a, b = int(seed[0]) & _MASK64, int(seed[1]) & _MASK64
  h = _splitmix(a ^ _splitmix(b))
  c0, c1 = _splitmix(h), _splitmix(h ^ 0xD1B54A32D192ED03)
  to31 = lambda v: torch.tensor([(v >> 33) & 0x7FFFFFFF, v & 0x7FFFFFFF], dtype=torch.int64)
  return to31(c0), to31(c1)


class EnergyInferenceBase(torch.nn.Module, abc.ABC):
  """Interface for inference on the EBM p(x) ~ exp(-E(x)) of a BitstringEnergy."""

  def __init__(self, input_energy, initial_seed=None, name=None):
    super().__init__()
    self.name = name
    self._energy = input_energy
    if torch.cuda.is_available():
      self._energy.to("cuda")
    self._energy.build([None, self._energy.num_bits])
    self._tracked_variables = input_energy.variables
    if len(self._tracked_variables) == 0:
      self._checkpoint = False
    else:
      self._tracked_variables_checkpoint = [v.detach().clone() for v in self._tracked_variables]
      self._tracked_versions = [None for _ in self._tracked_variables]
      self._checkpoint = True
    self._update_seed = initial_seed is None
    self._seed = sanitize_seed(initial_seed)
    self._first_inference = True

  @property
  def device(self):
    params = list(self._energy.parameters())
    if params:
      return params[0].device
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")

  @property
  def energy(self):
    return self._energy

  @property
  def seed(self):
    """Seed used by the next `sample`; advanced after every inference call unless fixed."""
    return self._seed

  @seed.setter
  def seed(self, initial_seed):
    self._update_seed = initial_seed is None
    self._seed = sanitize_seed(initial_seed)

  @property
  def variables_updated(self):
    """True iff some tracked variable differs from its checkpointed value."""
    if not self._checkpoint:
      return False
    changed = False
    for i, (v, vc) in enumerate(zip(self._tracked_variables, self._tracked_variables_checkpoint)):
      stamp = (v.data_ptr(), v._version)  # in-place updates bump the version: skip the compare otherwise
      if self._tracked_versions[i] == stamp:
        continue
      if vc.device != v.device or vc.shape != v.shape or not torch.equal(v.detach(), vc):
        changed = True
      else:
        self._tracked_versions[i] = stamp
    return changed

  def _checkpoint_variables(self):
    if self._checkpoint:
      self._tracked_variables_checkpoint = [v.detach().clone() for v in self._tracked_variables]
      self._tracked_versions = [(v.data_ptr(), v._version) for v in self._tracked_variables]

  def _preface_inference(self):
    if self._first_inference:
      self._checkpoint_variables()
      self._ready_inference()
      self._first_inference = False
    if self._update_seed:
      new_seed, _ = split_seed(self._seed)
      self._seed = new_seed
    if self.variables_updated:
      self._checkpoint_variables()
      self._ready_inference()

  @abc.abstractmethod
  def _ready_inference(self):
    """Recomputes whatever depends on the energy's variables."""

  @preface_inference
  def forward(self, inputs, *args, **kwargs):
    return self._call(inputs, *args, **kwargs)

  def call(self, inputs, *args, **kwargs):
    return self.forward(inputs, *args, **kwargs)

  @preface_inference
  def entropy(self):
    return self._entropy()

  @preface_inference
  def expectation(self, function):
    """Estimate of E_{x~p}[function(x)]; `function` maps int8 bitstrings [U, n] to a (nested
    structure of) float tensor(s) with leading dimension U."""
    return self._expectation(function)

  @preface_inference
  def log_partition(self):
    return self._log_partition()

  @preface_inference
  def sample(self, num_samples):
    return self._sample(num_samples)

  @preface_inference
  def unique_samples(self, num_samples):
    """(unique bitstrings int8 [U, n] in first-occurrence order, idx int32 [N], counts int32 [U]) of
    `num_samples` fresh samples -- the same draw `sample` would return (same seed step).  Engines
    that sample packed keys (`_sample_keys`) never materialise the int8 [N, n] tensor here."""
    keys = self._sample_keys(num_samples)
    if keys is None:
      return utils.unique_bitstrings_with_counts(self._sample(num_samples).detach())
    n = self.energy.num_bits
    uniq, idx, count = engine.unique_with_counts(keys)
    return engine.unpack_bits(uniq, n, utils._natural_shifts(n)), idx, count

  def _sample_keys(self, num_samples):
    """Packed uint64 keys (bit n-1-j of the key = column j) of `num_samples` samples, or None."""
    del num_samples
    return None

  @abc.abstractmethod
  def _call(self, inputs, *args, **kwargs):
    raise NotImplementedError()

  @abc.abstractmethod
  def _entropy(self):
    raise NotImplementedError()

  @abc.abstractmethod
  def _expectation(self, function):
    raise NotImplementedError()

  @abc.abstractmethod
  def _log_partition(self):
    raise NotImplementedError()

  @abc.abstractmethod
  def _sample(self, num_samples):
    raise NotImplementedError()


class EnergyInference(EnergyInferenceBase):
  """Default estimators: sample averages with score-function gradients."""

  def __init__(self, input_energy, num_expectation_samples, initial_seed=None, name=None):
    super().__init__(input_energy, initial_seed, name)
    self.num_expectation_samples = num_expectation_samples

  def _entropy(self):
    return self.expectation(self.energy) + self.log_partition()

  def _energy_needs_grad(self):
    return torch.is_grad_enabled() and any(p.requires_grad for p in self.energy.parameters())

  def _expectation(self, function):
    """Sample average; d/dtheta = E[c]E[dE] - E[c dE] + E[d function] (reference ebm.py:282-325),
    realised by adding a zero-valued surrogate whose gradient is the covariance term."""
    bitstrings, _, counts = self.unique_samples(self.num_expectation_samples)
    values = function(bitstrings)
    average_of_values = map_structure(lambda x: utils.weighted_average(counts, x), values)
    if not self._energy_needs_grad():
      return average_of_values
    energies = self.energy(bitstrings)
    weights = counts.to(energies.dtype) / counts.sum().to(energies.dtype)
    weighted_energies = weights * energies

    def add_score_term(avg, val):
      centered = (avg.detach().unsqueeze(0) - val.detach()).to(weighted_energies.dtype)
      surrogate = torch.tensordot(weighted_energies, centered, dims=([0], [0]))
      return avg + (surrogate - surrogate.detach()).to(avg.dtype)

    return map_structure(add_score_term, average_of_values, values)

  def _log_partition(self):
    """Forward value from the subclass; gradient -E_{x~p}[dE/dtheta] from fresh samples
    (reference ebm.py:331-343, 396-415)."""
    result = self._log_partition_forward_pass().detach()
    if not self._energy_needs_grad():
      return result
    unique_samples, _, counts = self.unique_samples(self.num_expectation_samples)
    unique_energies = self.energy(unique_samples)
    surrogate = -utils.weighted_average(counts, unique_energies)
    return result + (surrogate - surrogate.detach())

  def _log_partition_forward_pass(self):
    """Monte-Carlo estimate with uniform samples: n log 2 - log N_s + logsumexp(-E(x_i))
    (reference ebm.py:345-394)."""
    n = self.energy.num_bits
    n_s = self.num_expectation_samples
    seed, _ = split_seed(self._seed)
    keys = engine.bernoulli_sample(torch.zeros(n, device=self.device), utils._natural_shifts(n), n_s, seed)
    samples = engine.unpack_bits(keys, n, utils._natural_shifts(n))
    energies = self.energy(samples)
    return n * math.log(2.0) - math.log(float(n_s)) + torch.logsumexp(-1.0 * energies, 0)


class Categorical:
  """The explicit distribution of AnalyticEnergyInference (stands in for tfd.Categorical)."""

  def __init__(self, owner):
    self._owner = owner

  def logits_parameter(self):
    return self._owner._logits

  def probs_parameter(self):
    return torch.softmax(self._owner._logits.double(), 0).float()

  def entropy(self):
    m, s, t = self._owner._stats.tolist()
    return torch.tensor(m + math.log(s) - t / s, dtype=torch.float32, device=self._owner._logits.device)

  def sample(self, num_samples, seed=None):
    seed = self._owner.seed if seed is None else sanitize_seed(seed)
    return engine.categorical_sample(self._owner._logits, int(num_samples), seed)


def _mlp_layers(energy):
  """[(W[in,out], b[out], act)] if the energy is a Linear/tanh/relu stack on raw bits, else None."""
  layers, pending = [], None
  mods = list(energy.energy_layers)
  for m in mods:
    if isinstance(m, torch.nn.Linear):
      if pending is not None:
        layers.append(pending + ("linear",))
      bias = m.bias if m.bias is not None else torch.zeros(m.out_features, device=m.weight.device)
      pending = (m.weight.detach().t().contiguous().float(), bias.detach().contiguous().float())
    elif isinstance(m, torch.nn.Tanh) and pending is not None:
      layers.append(pending + ("tanh",))
      pending = None
    elif isinstance(m, torch.nn.ReLU) and pending is not None:
      layers.append(pending + ("relu",))
      pending = None
    elif isinstance(m, (utils.Squeeze, torch.nn.Flatten, torch.nn.Identity)):
      continue
    else:
      return None
  if pending is not None:
    layers.append(pending + ("linear",))
  if not layers or layers[-1][0].shape[1] != 1 or layers[0][0].shape[0] != energy.num_bits:
    return None
  if any(w.shape[1] > 64 or w.shape[0] > 64 for w, _, _ in layers) or len(layers) > 8:
    return None
  return layers


def energy_descriptor(energy):
  """CUDA-kernel description of a recognised energy (Bernoulli, KOBE, dense stack), else None."""
  if hasattr(energy, "kernel_descriptor"):
    kind, masks, theta = energy.kernel_descriptor()
    dev = theta.device
    masks_t = torch.tensor(masks, dtype=torch.int64, device=dev).to(torch.int32)
    code = nat.ENERGY_BERNOULLI if kind == "bernoulli" else nat.ENERGY_KOBE
    return engine.EnergyDescriptor(code, energy.num_bits, masks_t, theta.detach().contiguous().float())
  layers = _mlp_layers(energy)
  if layers is not None:
    return engine.EnergyDescriptor(nat.ENERGY_MLP, energy.num_bits, layers=layers)
  return None


class AnalyticEnergyInference(EnergyInference):
  """Exact inference by enumerating all 2^n bitstrings (reference ebm.py:418-492).

  The reference materialises an int8 [2^n, n] tensor and evaluates the Keras energy on it; here
  row r IS the big-endian bitstring of r, the energy is evaluated by the CUDA sweep kernel straight
  from the row index (Bernoulli / KOBE / dense-stack energies; other stacks run through their own
  torch layers in chunks), and logsumexp / entropy statistics come out of the same sweep.
  """

  def __init__(self, input_energy, num_expectation_samples, initial_seed=None, name=None):
    super().__init__(input_energy, num_expectation_samples, initial_seed, name)
    if input_energy.num_bits > 30:
      raise ValueError("AnalyticEnergyInference enumerates 2^n rows; n must be <= 30")
    self._all_bitstrings = None
    self._logits = None
    self._stats = None
    self._distribution = Categorical(self)

  @property
  def all_bitstrings(self):
    """int8 [2^n, n]: row r is the big-endian binary of r (itertools.product order)."""
    if self._all_bitstrings is None:
      n = self.energy.num_bits
      rows = torch.arange(1 << n, dtype=torch.int64, device=self.device)
      self._all_bitstrings = engine.unpack_bits(rows, n, utils._natural_shifts(n))
    return self._all_bitstrings

  @property
  def all_energies(self):
    return self.energy(self.all_bitstrings)

  @property
  def distribution(self):
    return self._distribution

  def _ready_inference(self):
    n = self.energy.num_bits
    desc = energy_descriptor(self.energy)
    if desc is not None:
      self._logits, self._stats = desc.sweep(0, 1 << n, device=self.device)
      self._stats = self._stats.cpu()
      return
    # arbitrary layer stack: the user's own torch layers, evaluated in chunks of rows
    chunks = []
    with torch.no_grad():
      step = 1 << 20
      for lo in range(0, 1 << n, step):
        rows = torch.arange(lo, min(lo + step, 1 << n), dtype=torch.int64, device=self.device)
        bits = engine.unpack_bits(rows, n, utils._natural_shifts(n))
        chunks.append(-self.energy(bits).float())
    self._logits = torch.cat(chunks)
    l64 = self._logits.double()
    m = l64.max()
    w = torch.exp(l64 - m)
    self._stats = torch.stack([m, w.sum(), (w * l64).sum()]).cpu()

  def _call(self, inputs, *args, **kwargs):
    if inputs is None:
      return self.distribution
    return self.sample(inputs)

  def _entropy(self):
    return self.distribution.entropy()

  def _log_partition_forward_pass(self):
    m, s, _ = self._stats.tolist()
    return torch.tensor(m + math.log(s), dtype=torch.float32, device=self.device)

  def _sample_keys(self, num_samples):
    return engine.categorical_sample(self._logits, int(num_samples), self.seed)  # row index IS the key

  def _sample(self, num_samples):
    n = self.energy.num_bits
    return engine.unpack_bits(self._sample_keys(num_samples), n, utils._natural_shifts(n))


class Bernoulli:
  """Product of independent bits (stands in for tfd.Bernoulli(logits, dtype=int8))."""

  def __init__(self, owner):
    self._owner = owner

  def logits_parameter(self):
    return self._owner._logits

  def probs_parameter(self):
    return torch.sigmoid(self._owner._logits)

  def entropy(self):
    l = self._owner._logits.double()
    p = torch.sigmoid(l)
    return (torch.nn.functional.softplus(l) - p * l).float()  # -p log p - (1-p) log(1-p)

  def sample(self, num_samples, seed=None):
    return self._owner._sample(num_samples) if seed is None else self._owner._sample_with(num_samples, seed)


class BernoulliEnergyInference(EnergyInference):
  """Exact inference for independent spins (reference ebm.py:495-561)."""

  def __init__(self, input_energy, num_expectation_samples, initial_seed=None, name=None):
    super().__init__(input_energy, num_expectation_samples, initial_seed, name)
    self._logits = input_energy.logits.detach().clone()
    self._distribution = Bernoulli(self)

  @property
  def distribution(self):
    return self._distribution

  def _ready_inference(self):
    self._logits = self.energy.logits.detach().clone().float()

  def _call(self, inputs, *args, **kwargs):
    if inputs is None:
      return self.distribution
    return self.sample(inputs)

  def _entropy(self):
    return torch.sum(self.distribution.entropy())

  def _log_partition_forward_pass(self):
    thetas = 0.5 * self._logits
    return torch.sum(torch.log(torch.exp(thetas) + torch.exp(-thetas)))

  def _keys_with(self, num_samples, seed):
    shifts = utils._natural_shifts(self.energy.num_bits)
    return engine.bernoulli_sample(self._logits.contiguous(), shifts, int(num_samples), sanitize_seed(seed))

  def _sample_keys(self, num_samples):
    return self._keys_with(num_samples, self.seed)

  def _sample_with(self, num_samples, seed):
    n = self.energy.num_bits
    return engine.unpack_bits(self._keys_with(num_samples, seed), n, utils._natural_shifts(n))

  def _sample(self, num_samples):
    return self._sample_with(num_samples, self.seed)
