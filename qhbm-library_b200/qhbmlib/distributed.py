"""Multi-GPU plumbing for the hot path: one process per GPU, `torch.distributed` (NCCL on the B200s,
gloo in the CPU tests).  The reference is single-process; the sharding below
follows SURVEY section 8(e):

  * unique bitstrings are independent -> contiguous shards, no data-path collective, and ONE
    all-reduce of the packed vector [sum_u c_u <H_j>_u (O) | sum_u c_u | gradient (P)] per step;
  * the 2^n EBM sweep shards the row range; per-rank (max, sum exp, sum exp*l) triples merge by
    rebasing to the common max; sampling first splits the sample count over ranks with a
    multinomial drawn identically on every rank from the shared seed.
"""
import math

import numpy as np
import torch
import torch.distributed as dist


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


def all_reduce_packed(packed, group=None):
  """The single collective of an expectation(+gradient) step."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
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
