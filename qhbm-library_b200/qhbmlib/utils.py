"""Helpers shared by the inference engines (mirror of /root/reference/qhbmlib/utils.py).

The integer work (row dedup in first-occurrence order) and the count-weighted reduction run
in libqhbm_b200.so; torch supplies tensors and autograd glue only.
"""
import torch

from qhbmlib import engine


class Squeeze(torch.nn.Module):
  """`tf.squeeze` as a layer (reference utils.py:20-40)."""

  def __init__(self, axis=None):
    super().__init__()
    self._axis = [] if axis is None else axis

  def forward(self, inputs):
    if self._axis == [] or self._axis is None:
      return torch.squeeze(inputs)
    axes = self._axis if isinstance(self._axis, (list, tuple)) else [self._axis]
    out = inputs
    for a in sorted((a % inputs.dim() for a in axes), reverse=True):
      out = torch.squeeze(out, a)
    return out


class _WeightedSum(torch.autograd.Function):
  """sum_u counts[u] * values[u, ...] / sum_u counts[u] on the GPU kernel (float64 accumulate)."""

  @staticmethod
  def forward(ctx, counts, values):
    flat = values.reshape(values.shape[0], -1).contiguous().float()
    acc = engine.weighted_sum(counts, flat)
    total = acc[-1]
    ctx.save_for_backward(counts)
    ctx.total = total
    ctx.shape = values.shape
    return (acc[:-1] / total).to(values.dtype).reshape(values.shape[1:])

  @staticmethod
  def backward(ctx, grad_out):
    (counts,) = ctx.saved_tensors
    w = (counts.double() / ctx.total).to(grad_out.dtype)
    g = w.reshape((-1,) + (1,) * (len(ctx.shape) - 1)) * grad_out.unsqueeze(0)
    return None, g


def weighted_average(counts, values):
  """Mean of `values` over axis 0 weighted by integer `counts` (reference utils.py:43-58).
  float32 values on the GPU go through the float64-accumulating kernel; float64 values keep their
  precision on the torch path."""
  if values.is_cuda and values.shape[0] > 0 and values.dtype == torch.float32:
    return _WeightedSum.apply(counts.to(torch.int32), values)
  fc = counts.to(values.dtype)
  return torch.tensordot(fc, values, dims=([0], [0])) / fc.sum()


class _WeightedPartialSum(torch.autograd.Function):
  """sum_u counts[u] * values[u, ...] (no normalisation): one rank's share of a weighted average."""

  @staticmethod
  def forward(ctx, counts, values):
    flat = values.reshape(values.shape[0], -1).contiguous().float()
    acc = engine.weighted_sum(counts, flat)
    ctx.save_for_backward(counts)
    ctx.shape = values.shape
    return acc[:-1].reshape(values.shape[1:])  # float64

  @staticmethod
  def backward(ctx, grad_out):
    (counts,) = ctx.saved_tensors
    g = counts.double().reshape((-1,) + (1,) * (len(ctx.shape) - 1)) * grad_out.unsqueeze(0)
    return None, g.float()


def weighted_sum(counts, values):
  """float64 sum over axis 0 of counts[u] * values[u, ...] (differentiable w.r.t. values)."""
  if values.is_cuda and values.shape[0] > 0 and values.dtype == torch.float32:
    return _WeightedPartialSum.apply(counts.to(torch.int32), values)
  return torch.tensordot(counts.double(), values.double(), dims=([0], [0]))


class _ScoreTerm(torch.autograd.Function):
  """Identity on `average` whose backward adds the score-function gradient of
  EnergyInference._expectation w.r.t. the parameters theta of a parity-feature energy
  (reference ebm.py:282-325): d/dtheta_t = E[c] E[f_t] - E[c f_t], computed by one fused kernel pair
  from the packed keys of the unique bitstrings instead of an energy Jacobian."""

  @staticmethod
  def forward(ctx, average, theta, values, keys, counts, total, masks, scale):
    ctx.save_for_backward(average.detach(), values.detach(), keys, counts, total, masks)
    ctx.scale, ctx.theta_shape, ctx.theta_dtype = scale, theta.shape, theta.dtype
    return average.view_as(average)

  @staticmethod
  def backward(ctx, grad):
    average, values, keys, counts, total, masks = ctx.saved_tensors
    flat = values.reshape(values.shape[0], -1).contiguous().float()
    g_theta = engine.score_gradient(keys, counts, flat, grad.reshape(-1).contiguous().float(),
                                    average.reshape(-1).contiguous().float(), masks, total, ctx.scale)
    return grad, g_theta.reshape(ctx.theta_shape).to(ctx.theta_dtype), None, None, None, None, None, None


def score_function_term(average, theta, values, keys, counts, total, masks, scale=1.0):
  """`average` with the score-function gradient path to `theta` attached (see _ScoreTerm)."""
  return _ScoreTerm.apply(average, theta, values, keys, counts.to(torch.int32), total, masks, scale)


def _natural_shifts(n):
  return [n - 1 - j for j in range(n)]


def mark_unique_rows(bitstrings):
  """Tags a tensor whose rows are known to be distinct (the output of a dedup): `QuantumInference.expectation`
  then skips its own dedup (reference qnn.py:68), which would be the identity on it.  The tag lives on the
  tensor OBJECT: any op on it (slice, cast, clone) yields an untagged tensor, which is deduplicated as usual."""
  bitstrings._qhbm_unique_rows = True
  return bitstrings


def rows_known_unique(bitstrings):
  return getattr(bitstrings, "_qhbm_unique_rows", False)


def unique_bitstrings_with_counts(bitstrings, out_idx=torch.int32):
  """Unique rows in first-occurrence order, inverse index and counts (reference utils.py:61-78,
  tf.raw_ops.UniqueWithCountsV2 on axis 0)."""
  if bitstrings.dim() != 2:
    raise ValueError("bitstrings must be 2-D")
  if not bitstrings.is_cuda:
    raise TypeError("bitstrings must live on the GPU (no CPU path)")
  n = bitstrings.shape[1]
  shifts = _natural_shifts(n)
  keys = engine.pack_bits(bitstrings.to(torch.int8), shifts)
  uniq, idx, count = engine.unique_with_counts(keys)
  y = engine.unpack_bits(uniq, n, shifts).to(bitstrings.dtype)
  return mark_unique_rows(y), idx.to(out_idx), count.to(out_idx)


class _Expand(torch.autograd.Function):

  @staticmethod
  def forward(ctx, y, idx):
    ctx.save_for_backward(idx)
    ctx.n_unique = y.shape[0]
    return y.index_select(0, idx.long())

  @staticmethod
  def backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    g = engine.segment_sum(grad.contiguous().float(), idx.to(torch.int32), ctx.n_unique)
    return g.to(grad.dtype), None


def expand_unique_results(y, idx):
  """Inverse of the dedup: expanded[i] = y[idx[i]] (reference utils.py:81-92); the backward is
  a segment-sum kernel."""
  if y.is_cuda and y.is_floating_point() and y.requires_grad:
    return _Expand.apply(y, idx)
  return y.index_select(0, idx.long())
