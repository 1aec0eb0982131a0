"""Multi-GPU plumbing for the hot path: one process per GPU, `torch.distributed` (NCCL on the B200s,
gloo in the CPU tests).  The reference is single-process; the sharding below
follows SURVEY section 8(e):

  * unique bitstrings are independent -> contiguous shards, no data-path collective, and ONE
    all-reduce of the packed vector [sum_u c_u <H_j>_u (O) | sum_u c_u | gradient (P)] per step;
  * the 2^n EBM sweep shards the row range; per-rank (max, sum exp, sum exp*l) triples merge by
    rebasing to the common max; sampling first splits the sample count over ranks with a
    multinomial drawn identically on every rank from the shared seed.
"""
import contextlib
import math
import threading

import numpy as np
import torch
import torch.distributed as dist

_local = threading.local()


@contextlib.contextmanager
def local_shard():
  """Inside this context the inference engines do NOT shard again: the caller already handed them
  this rank's share (EnergyInference._expectation shards the unique bitstrings and then calls the
  user's function, which usually ends in QuantumInference.expectation)."""
  prev = getattr(_local, "depth", 0)
  _local.depth = prev + 1
  try:
    yield
  finally:
    _local.depth = prev


def active(group=None):
  """True when the engines should shard their data-parallel work over the ranks of `group`."""
  return (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and
          getattr(_local, "depth", 0) == 0)


def world(group=None):
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(group), dist.get_world_size(group)
  return 0, 1


def shard_range(n, rank, world_size):
  """Contiguous [lo, hi) of n items owned by `rank` (sizes differ by at most one)."""
  base, rem = divmod(int(n), int(world_size))
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def pack(weighted_sums, total_count, grad=None):
  """[sum c<H> (O) | sum c | grad (P)] as one float64 vector (the only thing ever all-reduced)."""
  parts = [weighted_sums.reshape(-1).double(), total_count.reshape(1).double()]
  if grad is not None:
    parts.append(grad.reshape(-1).double())
  return torch.cat(parts)


def unpack(packed, n_ops):
  sums, total, grad = packed[:n_ops], packed[n_ops], packed[n_ops + 1:]
  return sums / total, total, grad


class NativeComm:
  """The C-ABI collective (`qhbm_comm_*`, `qhbm_allreduce`; include/qhbm_b200.h): the communicator a host
  without torch would use.  Rank 0 draws the NCCL unique id, `torch.distributed` (any backend) hands it to the
  other ranks, every rank joins on its current CUDA device.  `all_reduce_` sums a float32 / float64 CUDA
  tensor in place over the ranks on torch's current stream."""

  def __init__(self, group=None, rank=None, world_size=None, unique_id=None):
    import ctypes
    from qhbmlib import _native as nat
    self._nat, self._handle = nat, ctypes.c_void_p()
    if rank is None:
      rank, world_size = world(group)
    if unique_id is None:
      ident = np.zeros(nat.COMM_ID_BYTES, dtype=np.uint8)
      if rank == 0:
        nat.check(nat.lib().qhbm_comm_unique_id(ident.ctypes.data, ident.size))
      if world_size > 1:
        box = [ident.tobytes()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = np.frombuffer(box[0], dtype=np.uint8).copy()
      unique_id = ident
    unique_id = np.ascontiguousarray(unique_id, dtype=np.uint8)
    if unique_id.size != nat.COMM_ID_BYTES:
      raise ValueError(f"unique_id must hold {nat.COMM_ID_BYTES} bytes")
    nat.check(nat.lib().qhbm_comm_create(unique_id.ctypes.data, int(rank), int(world_size), ctypes.byref(self._handle)))
    self.rank, self.world_size = int(rank), int(world_size)

  @property
  def nccl_version(self):
    import ctypes
    v = ctypes.c_int32()
    self._nat.check(self._nat.lib().qhbm_comm_info(self._handle, None, None, ctypes.byref(v)))
    return int(v.value)

  def all_reduce_(self, t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
      raise TypeError("all_reduce_ needs a contiguous CUDA tensor")
    if t.dtype not in (torch.float32, torch.float64):
      raise TypeError("all_reduce_ sums float32 or float64 tensors")
    from qhbmlib.engine import _stream
    self._nat.check(self._nat.lib().qhbm_allreduce(self._handle, t.data_ptr(), t.numel(),
                                                   0 if t.dtype == torch.float32 else 1, _stream()))
    return t

  def close(self):
    if self._handle:
      self._nat.lib().qhbm_comm_destroy(self._handle)
      self._handle = None

  def __del__(self):
    import sys
    if sys.is_finalizing():  # the CUDA context may already be gone: leave the communicator to process exit
      return
    try:
      self.close()
    except Exception:
      pass


_native_comms = {}


def close_native_comms():
  """Destroys the communicators `native_comm` created (call before `destroy_process_group`)."""
  while _native_comms:
    _native_comms.popitem()[1].close()


def native_comm(group=None):
  """The process's NativeComm of `group` (created collectively on first use)."""
  key = id(group) if group is not None else None
  if key not in _native_comms:
    _native_comms[key] = NativeComm(group)
  return _native_comms[key]


def all_reduce_packed(packed, group=None):
  """The single collective of an expectation(+gradient) step.  QHBM_NATIVE_ALLREDUCE=1 routes it through the
  library's own communicator (`qhbm_allreduce`) instead of torch.distributed's; the sums are the same."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    import os
    if packed.is_cuda and os.environ.get("QHBM_NATIVE_ALLREDUCE") == "1":
      native_comm(group).all_reduce_(packed)  # (packed comes from torch.cat: contiguous)
    else:
      dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
  return packed


def merge_log_stats(triples):
  """Merges per-shard (m, s, t) = (max logit, sum exp(l-m), sum exp(l-m) l) into the global triple;
  logZ = m + log s, entropy = logZ - t/s."""
  m = max(float(t[0]) for t in triples)
  s = sum(float(t[1]) * math.exp(float(t[0]) - m) for t in triples)
  tt = sum(float(t[2]) * math.exp(float(t[0]) - m) for t in triples)
  return m, s, tt


def all_gather_stats(stats, group=None):
  """stats: float64[3] tensor of this rank -> list of triples of every rank."""
  rank, ws = world(group)
  if ws == 1:
    return [stats.tolist()]
  out = [torch.zeros_like(stats) for _ in range(ws)]
  dist.all_gather(out, stats, group=group)
  return [o.tolist() for o in out]


def split_samples(num_samples, shard_log_masses, seed):
  """How many of `num_samples` each shard draws: multinomial over softmax(shard log masses), drawn
  with a generator seeded identically on every rank (so all ranks agree without communication)."""
  lm = np.asarray(shard_log_masses, dtype=np.float64)
  p = np.exp(lm - lm.max())
  p /= p.sum()
  rng = np.random.default_rng([int(seed[0]) & 0xFFFFFFFF, int(seed[1]) & 0xFFFFFFFF, 0x5EED])
  return rng.multinomial(int(num_samples), p)


class ShardedExpectation:
  """Count-weighted expectation (+ adjoint gradient) of ALL unique bitstrings, sharded over ranks.

  Every rank passes the same global arrays; each computes its contiguous shard on its own GPU and the
  packed partial sums are all-reduced once."""

  def __init__(self, plan, group=None):
    self.plan, self.group = plan, group

  def __call__(self, basis_idx, counts, symbols, with_gradient=True, grad_mode="tfq_fd"):
    from qhbmlib import engine
    rank, ws = world(self.group)
    lo, hi = shard_range(basis_idx.shape[0], rank, ws)
    total = counts.sum().double()
    b, c = basis_idx[lo:hi].contiguous(), counts[lo:hi].contiguous()
    if with_gradient:
      dgrad = (c.double() / total).float().unsqueeze(1).expand(-1, self.plan.n_ops).contiguous()
      e, g = self.plan.forward_adjoint(b, symbols, dgrad, grad_mode=grad_mode)
    else:
      e, g = self.plan.forward(b, symbols), None
    ws_sum = engine.weighted_sum(c.to(torch.int32), e)
    packed = pack(ws_sum[:-1], ws_sum[-1:], g)
    all_reduce_packed(packed, self.group)
    return unpack(packed, self.plan.n_ops)


def sharded_ebm_sweep(descriptor, n_bits, group=None, want_logits=True, device="cuda"):
  """Each rank sweeps rows [lo, hi) of the 2^n enumeration.  Returns (local logits, (lo, hi),
  global log Z, global entropy, per-rank log masses)."""
  rank, ws = world(group)
  lo, hi = shard_range(1 << n_bits, rank, ws)
  logits, stats = descriptor.sweep(lo, hi, want_logits=want_logits, device=device)
  triples = all_gather_stats(stats, group)
  m, s, t = merge_log_stats(triples)
  log_z = m + math.log(s)
  masses = [tr[0] + math.log(tr[1]) if tr[1] > 0 else -math.inf for tr in triples]
  return logits, (lo, hi), log_z, log_z - t / s, masses


# ----------------------------------------------------------------------------------------------
# Autograd glue.  Convention ("average", the one DistributedDataParallel uses): every rank computes the
# SAME scalar loss from all-reduced / all-gathered values, each rank's backward pass only sees its own
# shard, and `sync_gradients` AVERAGES parameter gradients over the ranks afterwards.  For that average to
# equal the true gradient, the backward of every collective multiplies by the world size; quantities that
# are computed identically on all ranks (regularisers, closed-form log-partition terms) then come out
# right as well, because their gradients are already identical on every rank.
# ----------------------------------------------------------------------------------------------
class _AllReduceSum(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, group):
    ctx.world = dist.get_world_size(group)
    y = x.detach().clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y

  @staticmethod
  def backward(ctx, grad):
    return grad * ctx.world, None


def all_reduce_sum(x, group=None):
  """Differentiable sum over ranks of a tensor of per-rank partial sums (ONE collective)."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return x
  return _AllReduceSum.apply(x, group)


class _ScaleGrad(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, factor):
    ctx.factor = factor
    return x.view_as(x)

  @staticmethod
  def backward(ctx, grad):
    return grad * ctx.factor, None


def shard_term(x, group=None):
  """Marks a per-rank partial term whose VALUE needs no communication (e.g. the zero-valued
  score-function surrogates) but whose gradient is one shard of a sum over ranks."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return x
  return _ScaleGrad.apply(x, float(dist.get_world_size(group)))


class _AllGatherRows(torch.autograd.Function):
  """Concatenates the ranks' row blocks (sizes from shard_range) on every rank."""

  @staticmethod
  def forward(ctx, x, n_total, group):
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    ctx.world, ctx.range = ws, shard_range(n_total, rank, ws)
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[:x.shape[0]] = x.detach()
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)

  @staticmethod
  def backward(ctx, grad):
    lo, hi = ctx.range
    return grad[lo:hi] * ctx.world, None, None


def all_gather_rows(x, n_total, group=None):
  """Differentiable all-gather of contiguous row shards (x = rows shard_range(n_total, rank, world))."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return x
  return _AllGatherRows.apply(x, n_total, group)


def sync_gradients(parameters, group=None):
  """Averages `.grad` of the given parameters over the ranks with one all-reduce of a flat buffer
  (call after `loss.backward()`; DistributedDataParallel does the same thing).  Parameters without a
  gradient on this rank contribute zeros."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return
  params = [p for p in parameters if p.requires_grad]
  if not params:
    return
  flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).double() for p in params])
  dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  flat /= dist.get_world_size(group)
  off = 0
  for p in params:
    n = p.numel()
    g = flat[off:off + n].reshape(p.shape).to(p.dtype)
    if p.grad is None:
      p.grad = g.clone()
    else:
      p.grad.copy_(g)
    off += n
