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
from qhbmlib import distributed as qd
from qhbmlib import engine
from qhbmlib import utils


def preface_inference(f):
  """Decorator: run `self._preface_inference()` before the wrapped method."""

  @functools.wraps(f)
  def wrapper(self, *args, **kwargs):
    self._preface_inference()  # pylint: disable=protected-access
    return f(self, *args, **kwargs)

  return wrapper


def _leaves(structure):
  out = []
  map_structure(out.append, structure)
  return out


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
    # Values are compared on every call, as the reference does (ebm.py:125-140): writes through `.data`
    # change neither the storage pointer nor the version counter, so no cheaper test is safe.  All
    # variables are compared on the device and ONE flag crosses to the host.
    flags = []
    for v, vc in zip(self._tracked_variables, self._tracked_variables_checkpoint):
      if vc.device != v.device or vc.shape != v.shape:
        return True
    if len(self._tracked_variables) == 1:  # (KOBE, Bernoulli: one fused compare-and-reduce launch)
      return not torch.equal(self._tracked_variables[0].detach(), self._tracked_variables_checkpoint[0])
    for v, vc in zip(self._tracked_variables, self._tracked_variables_checkpoint):
      flags.append((v.detach() != vc).any())
    return bool(torch.stack(flags).any().item()) if flags else False

  def _checkpoint_variables(self):
    if self._checkpoint:
      self._tracked_variables_checkpoint = [v.detach().clone() for v in self._tracked_variables]

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
    return utils.mark_unique_rows(engine.unpack_bits(uniq, n, utils._natural_shifts(n))), idx, count

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


class _LogPartitionGrad(torch.autograd.Function):
  """log Z with the reference's gradient estimator attached LAZILY, as `tf.custom_gradient` does it
  (ebm.py:331-343 forward, 396-415 grad_fn): the samples behind -E_{x~p}[dE/dtheta] are drawn in backward,
  i.e. only when somebody differentiates through log Z.  `vqt` detaches it (vqt_loss.py:52-54), so a VQT step
  pays no sampling, dedup or energy evaluation for a gradient nobody takes; `qmhl` takes it."""

  @staticmethod
  def forward(ctx, owner, result, sharded, *params):
    ctx.owner, ctx.sharded, ctx.params = owner, sharded, params  # (leaf variables: no version check wanted)
    return result.detach().clone()

  @staticmethod
  def backward(ctx, upstream):
    params = ctx.params
    with torch.enable_grad():
      surrogate = ctx.owner._log_partition_surrogate(ctx.sharded)  # pylint: disable=protected-access
      grads = torch.autograd.grad(surrogate, params, upstream.to(surrogate.dtype).reshape(surrogate.shape),
                                  allow_unused=True)
    return (None, None, None) + tuple(grads)


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
    realised by adding a zero-valued surrogate whose gradient is the covariance term.

    Under torch.distributed (world size > 1) every rank draws the SAME samples (same seed) and owns a
    contiguous shard of the unique bitstrings: `function` only sees that shard, and the count-weighted
    partial sums of all leaves of its result travel in ONE all-reduce.  Gradients follow
    the convention of `qhbmlib.distributed` (call `sync_gradients` after `backward`)."""
    bitstrings, _, counts = self.unique_samples(self.num_expectation_samples)
    sharded = qd.active()
    if sharded:
      rank, world = qd.world()
      lo, hi = qd.shard_range(bitstrings.shape[0], rank, world)
      total = counts.sum()
      bitstrings, counts = utils.mark_unique_rows(bitstrings[lo:hi].contiguous()), counts[lo:hi].contiguous()
      with qd.local_shard():
        values = function(bitstrings)
      leaves = []
      map_structure(leaves.append, values)
      partial = [utils.weighted_sum(counts, x).reshape(-1) for x in leaves]  # sum_u c_u v_u of this shard
      packed = qd.all_reduce_sum(torch.cat([p.double() for p in partial])) / total.double()
      it, off = iter(leaves), [0]

      def take(_):
        x = next(it)
        n = x[0].numel() if x.dim() > 1 else 1
        out = packed[off[0]:off[0] + n].to(x.dtype).reshape(x.shape[1:])
        off[0] += n
        return out

      average_of_values = map_structure(take, values)
      weight_norm = total
    else:
      values = function(bitstrings)
      average_of_values = map_structure(lambda x: utils.weighted_average(counts, x), values)
      weight_norm = counts.sum()
    if not self._energy_needs_grad():
      return average_of_values
    fused = self._parity_feature_tables()
    if fused is not None and bitstrings.is_cuda and all(v.is_cuda and v.dtype == torch.float32 for v in _leaves(values)):
      # f1: on-device score-function glue.  For parity-feature energies (Bernoulli, KOBE) the Jacobian
      # dE/dtheta is the +-1 feature matrix, so the whole backward term is one fused kernel pair over the
      # packed keys -- no energy evaluation, no Jacobian, no chain of small torch ops.
      theta, masks = fused
      n = self.energy.num_bits
      keys = engine.pack_bits(bitstrings, utils._natural_shifts(n))
      total = weight_norm.double().reshape(1)
      scale = float(qd.world()[1]) if sharded else 1.0
      return map_structure(
          lambda avg, val: utils.score_function_term(avg, theta, val, keys, counts, total, masks, scale),
          average_of_values, values)
    energies = self.energy(bitstrings)
    weights = counts.to(energies.dtype) / weight_norm.to(energies.dtype)
    weighted_energies = weights * energies

    def add_score_term(avg, val):
      centered = (avg.detach().unsqueeze(0) - val.detach()).to(weighted_energies.dtype)
      surrogate = torch.tensordot(weighted_energies, centered, dims=([0], [0]))
      if sharded:
        surrogate = qd.shard_term(surrogate)  # its value is zero anyway; its gradient is one shard of the sum
      return avg + (surrogate - surrogate.detach()).to(avg.dtype)

    return map_structure(add_score_term, average_of_values, values)

  def _parity_feature_tables(self):
    """(theta parameter, int32 masks on the device) when the energy is sum_t theta_t x parity feature
    (BernoulliEnergy, KOBE) and theta is trainable; else None."""
    if not hasattr(self.energy, "kernel_descriptor"):
      return None
    _, masks, theta = self.energy.kernel_descriptor()
    if not (isinstance(theta, torch.nn.Parameter) and theta.requires_grad and theta.is_cuda and
            theta.dtype == torch.float32):
      return None
    cache = self.__dict__.setdefault("_mask_cache", {})
    key = (tuple(masks), theta.device)
    if key not in cache:
      cache.clear()
      cache[key] = torch.tensor(masks, dtype=torch.int64, device=theta.device).to(torch.int32)
    return theta, cache[key]

  def _log_partition(self):
    """Forward value from the subclass; gradient -E_{x~p}[dE/dtheta] from fresh samples
    (reference ebm.py:331-343, 396-415).  Sharded like `_expectation` when torch.distributed is active."""
    result = self._log_partition_forward_pass().detach()
    if not self._energy_needs_grad():
      return result
    params = [p for p in self.energy.parameters() if p.requires_grad]
    return _LogPartitionGrad.apply(self, result, qd.active(), *params)

  def _log_partition_surrogate(self, sharded):
    """-E_{x~p}[E(x)] over fresh samples as a differentiable function of the energy's variables: its gradient is
    the reference's log-partition gradient estimator (ebm.py:396-415)."""
    unique_samples, _, counts = self.unique_samples(self.num_expectation_samples)
    if sharded:
      rank, world = qd.world()
      lo, hi = qd.shard_range(unique_samples.shape[0], rank, world)
      total = counts.sum()
      e = self.energy(unique_samples[lo:hi].contiguous())
      w = counts[lo:hi].to(e.dtype) / total.to(e.dtype)
      return qd.shard_term(-(w * e).sum())
    return -utils.weighted_average(counts, self.energy(unique_samples))

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
    return self._owner._full_logits()

  def probs_parameter(self):
    return torch.softmax(self._owner._full_logits().double(), 0).float()

  def entropy(self):
    m, s, t = self._owner._stats.tolist()
    return torch.tensor(m + math.log(s) - t / s, dtype=torch.float32, device=self._owner._logits.device)

  def sample(self, num_samples, seed=None):
    seed = self._owner.seed if seed is None else sanitize_seed(seed)
    return self._owner._draw_keys(int(num_samples), seed)


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
    # the masks never change: one host -> device copy per (energy, device), not one per parameter update
    cache = energy.__dict__.setdefault("_masks_dev_cache", {})
    key = (str(dev), len(masks))
    masks_t = cache.get(key)
    if masks_t is None:
      masks_t = torch.tensor(masks, dtype=torch.int64, device=dev).to(torch.int32)
      cache.clear()
      cache[key] = masks_t
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
    self._sampler = None
    self._row_range = (0, 1 << input_energy.num_bits)
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
    """logits = -E(all rows), plus (max, sum exp, sum exp * l) for log Z and the entropy (reference
    ebm.py:467-469).  Under torch.distributed every rank sweeps its contiguous share of the 2^n rows and
    keeps only those logits; the statistics merge through one all-gather of three doubles per rank."""
    n = self.energy.num_bits
    desc = energy_descriptor(self.energy)
    self._sampler = None
    self._row_range = (0, 1 << n)
    rank, world = qd.world()
    if desc is not None and world > 1:
      lo, hi = qd.shard_range(1 << n, rank, world)
      self._logits, stats = desc.sweep(lo, hi, device=self.device)
      m, sm, t = qd.merge_log_stats(qd.all_gather_stats(stats))
      self._stats = torch.tensor([m, sm, t], dtype=torch.float64)
      self._row_range = (lo, hi)
      return
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

  def _sharded(self):
    return self._row_range != (0, 1 << self.energy.num_bits)

  def _full_logits(self):
    """f32[2^n] on every rank (the local shard otherwise stays where it was computed)."""
    if not self._sharded():
      return self._logits
    rank, world = qd.world()
    n = self.energy.num_bits
    sizes = [qd.shard_range(1 << n, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(width, dtype=torch.float32, device=self._logits.device)
    pad[:self._logits.shape[0]] = self._logits
    out = [torch.empty_like(pad) for _ in range(world)]
    torch.distributed.all_gather(out, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)])

  def _draw_keys(self, num_samples, seed):
    """Row indices (= packed bitstring keys) of `num_samples` samples.  The prefix sums over the logits
    are prepared once per parameter change, with the maximum taken from the sweep statistics.  Sharded:
    sample k is drawn by the rank whose interval of the global cumulative mass holds its uniform, the
    others leave a zero, and one all-reduce(sum) of the int64 sample vector gives every rank the full
    draw -- the samples a single GPU would have drawn, whatever the number of ranks."""
    if self._sampler is None:
      self._sampler = engine.CategoricalSampler(self._logits, given_max=float(self._stats[0]))
      self._mass_interval = None
      if self._sharded():
        rank, world = qd.world()
        mass = self._sampler.local_mass().clone()
        masses = [torch.zeros_like(mass) for _ in range(world)]
        torch.distributed.all_gather(masses, mass)
        cum = [0.0]
        for m in masses:
          cum.append(cum[-1] + float(m.item()))
        self._mass_interval = (cum[rank], cum[rank + 1] if rank + 1 < world else math.inf, cum[-1])
    if not self._sharded():
      return self._sampler.draw(num_samples, seed)
    out = torch.zeros((num_samples,), dtype=torch.int64, device=self._logits.device)
    self._sampler.draw(num_samples, seed, row_offset=self._row_range[0], mass_interval=self._mass_interval, out=out)
    torch.distributed.all_reduce(out)
    return out

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
    return self._draw_keys(int(num_samples), self.seed)  # row index IS the key

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
