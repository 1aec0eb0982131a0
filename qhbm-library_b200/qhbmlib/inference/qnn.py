"""Inference on parameterised circuits (mirror of /root/reference/qhbmlib/inference/qnn.py).

`AnalyticQuantumInference.expectation` is the entry point of the B200 hot path: unique initial
states -> basis indices -> one compiled plan per (circuit, observables) -> CUDA expectation kernels,
with the adjoint-gradient kernels wired into torch autograd (the reference wires TFQ's adjoint
differentiator into TF autodiff, qnn.py:87-139).
"""
import abc
import collections
import math
import os

import numpy as np
import torch

from qhbmlib import _native as nat
from qhbmlib import circuits as cq
from qhbmlib import distributed as qd
from qhbmlib import engine
from qhbmlib import utils
from qhbmlib.models import energy as energy_lib
from qhbmlib.models import hamiltonian as hamiltonian_lib


class _LRU:
  """Small least-recently-used map.  Compiled plans own device tables and grow-only workspaces, so the
  caches that hold them are keyed by CONTENT (an observable or circuit rebuilt inside a training loop hits
  the same entry) and bounded (an evicted plan frees its device memory when it is collected)."""

  def __init__(self, maxsize=8):
    self.maxsize = maxsize
    self._d = collections.OrderedDict()

  def get(self, key, make):
    if key in self._d:
      self._d.move_to_end(key)
      return self._d[key]
    value = make()
    self._d[key] = value
    while len(self._d) > self.maxsize:
      self._d.popitem(last=False)
    return value

  def __len__(self):
    return len(self._d)


class _ExpectationOp(torch.autograd.Function):
  """f32[U, O] expectations; backward = adjoint gradient w.r.t. the symbol values.

  Several observables: like TFQ, the backward pass re-simulates the forward circuit (the upstream
  gradient is needed before the reverse sweep can start).  A single observable does not need it:
  d<H>_u/d phi is computed per state in the forward call (one fused forward + adjoint run) and the
  backward pass is a [U] x [U, P] contraction."""

  @staticmethod
  def forward(ctx, symbol_values, basis_idx, holder):
    values = symbol_values.detach().contiguous().float()
    ctx.holder = holder
    ctx.jacobian = None
    single = holder.plan.n_ops == 1 and symbol_values.numel() > 0 and basis_idx.shape[0] > 0
    if single and ctx.needs_input_grad[0] and holder.jacobian_in_forward:
      ones = torch.ones((basis_idx.shape[0], 1), dtype=torch.float32, device=basis_idx.device)
      vals, ctx.jacobian = holder.plan.forward_adjoint(basis_idx, values, ones, per_state=True,
                                                       grad_mode=holder.grad_mode)
      ctx.save_for_backward(symbol_values)
      return vals
    ctx.save_for_backward(symbol_values, basis_idx)
    return holder.forward_plan.forward(basis_idx, values)

  @staticmethod
  def backward(ctx, grad_out):
    symbol_values = ctx.saved_tensors[0]
    if symbol_values.numel() == 0:
      return torch.zeros_like(symbol_values), None, None
    if ctx.jacobian is not None:
      grad = grad_out.reshape(-1).float() @ ctx.jacobian
      return grad.to(symbol_values.dtype), None, None
    basis_idx = ctx.saved_tensors[1]
    holder = ctx.holder
    _, grad = holder.plan.forward_adjoint(basis_idx, symbol_values.detach().contiguous().float(),
                                          grad_out.contiguous().float(), per_state=False,
                                          grad_mode=holder.grad_mode)
    return grad.to(symbol_values.dtype), None, None


class _PlanHolder:
  """The adjoint plan, plus a forward-only plan (wider register blocking, no lambda tile) compiled
  lazily for the forward pass of autograd."""

  jacobian_in_forward = True

  def __init__(self, plan, grad_mode, make_forward_plan):
    self.plan, self.grad_mode = plan, grad_mode
    self._make_forward_plan = make_forward_plan
    self._forward_plan = None

  @property
  def forward_plan(self):
    if self._forward_plan is None:
      self._forward_plan = self._make_forward_plan()
    return self._forward_plan


class QuantumInference(torch.nn.Module, abc.ABC):
  """Interface: expectation values of observables against U(phi)|initial state>."""

  def __init__(self, input_circuit, name=None):
    super().__init__()
    self.name = name
    input_circuit.build([])
    if torch.cuda.is_available():
      input_circuit.to("cuda")  # the engine has no CPU path; parameters live next to the kernels
    self._circuit = input_circuit

  @property
  def circuit(self):
    return self._circuit

  def expectation(self, initial_states, observables):
    """[batch, n_ops] un-averaged <op_j> on circuit|initial_states[i]>.

    initial_states: int8 [batch, num_qubits] on the GPU.  observables: an OperatorTensor
    (`circuits.convert_to_tensor([PauliSum, ...])`) or a `Hamiltonian`, in which case the
    circuit is extended by the Hamiltonian's inverse eigenvector circuit and its Z-string shards are
    measured (reference qnn.py:50-80)."""
    if utils.rows_known_unique(initial_states):
      # the rows come straight from a dedup (EnergyInference._expectation hands its unique samples to the
      # user's function): deduplicating them again would be the identity, six launches and a host sync
      unique_states, idx = initial_states, None
    else:
      unique_states, idx, _ = utils.unique_bitstrings_with_counts(initial_states)
    if isinstance(observables, cq.OperatorTensor):
      total_circuit = self.circuit
    else:
      total_circuit = self._total_circuit(observables)
    if qd.active():
      # one process per GPU: this rank simulates its contiguous share of the unique states; the rows are
      # all-gathered so that every rank returns the full [batch, n_ops] result
      rank, world = qd.world()
      n_unique = unique_states.shape[0]
      lo, hi = qd.shard_range(n_unique, rank, world)
      circuits = total_circuit(unique_states[lo:hi].contiguous())
      local = self._expectation(circuits, total_circuit.symbol_names, total_circuit.symbol_values, observables)
      unique_expectations = qd.all_gather_rows(local, n_unique)
    else:
      circuits = total_circuit(unique_states)
      unique_expectations = self._expectation(circuits, total_circuit.symbol_names, total_circuit.symbol_values,
                                              observables)
    return unique_expectations if idx is None else utils.expand_unique_results(unique_expectations, idx)

  def _total_circuit(self, observables):
    """circuit + observables.circuit_dagger (reference qnn.py:69-72), kept for the few Hamiltonians in
    use (the entries hold their key objects alive, so ids cannot be recycled while cached)."""
    cache = self.__dict__.setdefault("_total_cache", _LRU(4))
    return cache.get(id(observables), lambda: (observables, self.circuit + observables.circuit_dagger))[1]

  @abc.abstractmethod
  def _expectation(self, circuits, symbol_names, symbol_values, observables):
    raise NotImplementedError()


class AnalyticQuantumInference(QuantumInference):
  """Exact expectation values with adjoint-method gradients on the B200 engine."""

  def __init__(self, input_circuit, name=None, grad_mode="tfq_fd", tile_qubits=0, reg_qubits=0):
    """grad_mode: "tfq_fd" reproduces TFQ 0.6.1's finite-difference gate derivative (SURVEY App.
    A.6, default for drop-in parity), "exact" the analytic derivative."""
    super().__init__(input_circuit, name)
    self.grad_mode = grad_mode
    self._tile_qubits, self._reg_qubits = tile_qubits, reg_qubits
    self._plans = _LRU(8)

  def _plan_for(self, circuit, ops_tensor):
    """Compiled plan for (gate table, Pauli tables), cached by content."""

    def make():
      terms, offsets = ops_tensor.tables(circuit.qubits)
      args = (circuit.gate_table(), len(circuit.qubits), len(circuit.symbol_names), terms, offsets)
      plan = engine.ExpectationPlan(*args, True, self._tile_qubits, self._reg_qubits)
      return _PlanHolder(plan, self.grad_mode, lambda: engine.ExpectationPlan(*args, False, 0, 0))

    key = (circuit.gate_table_digest(), ops_tensor.tables_digest(circuit.qubits))
    holder = self._plans.get(key, make)
    holder.grad_mode = self.grad_mode
    return holder

  def _expectation(self, circuits, symbol_names, symbol_values, observables):
    del symbol_names
    if isinstance(observables, cq.OperatorTensor):
      ops = observables
      post_process = lambda x: x
    elif isinstance(observables, hamiltonian_lib.Hamiltonian) and isinstance(observables.energy,
                                                                          energy_lib.PauliMixin):
      ops = observables.operator_shards
      post_process = lambda y: observables.energy.operator_expectation(y).unsqueeze(-1)
    else:
      raise TypeError("General Hamiltonians not accepted.  Please use `SampledQuantumInference` instead.")
    holder = self._plan_for(circuits.circuit, ops)
    values = symbol_values
    if not values.is_cuda:
      values = values.to(circuits.basis_idx.device)
    expectations = _ExpectationOp.apply(values, circuits.basis_idx, holder)
    return post_process(expectations)


# ------------------------------------------------------------------------------------------------
# Shot-based inference (reference qnn.py:142-292)

# gates whose exponent generator has exactly two distinct eigenvalues (gap 1): the two-term
# parameter-shift rule  d f/d t = (pi/2) [f(t + 1/2) - f(t - 1/2)]  holds for the exponent t
_TWO_EIGENVALUE_GATES = (1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12)  # X Y Z H CZ CNOT SWAP XX YY ZZ PhasedX (exponent)


class _Occurrences:
  """Gate table in which every shiftable symbolic exponent is its own symbol.

  TFQ's ParameterShift differentiator shifts each gate occurrence of a symbol separately
  (`get_gradient_circuits`, used at reference qnn.py:192-195).  Here the table keeps the original
  symbols at [0, P) and appends one symbol per shiftable occurrence whose value is the gate's
  whole exponent scalar * phi[sym] + const; shifting occurrence g is adding +-1/2 to entry P + g."""

  def __init__(self, gates, n_symbols):
    self.n_symbols = int(n_symbols)
    table = np.array(gates, dtype=nat.GATE_DTYPE, copy=True)
    sym, scalar, const, blocked = [], [], [], []
    for i in range(len(table)):
      for k in range(int(table[i]["nparams"])):
        s = int(table[i]["sym"][k])
        if s < 0:
          continue
        if k == 0 and int(table[i]["type"]) in _TWO_EIGENVALUE_GATES:
          sym.append(s)
          scalar.append(float(table[i]["scalar"][k]))
          const.append(float(table[i]["cnst"][k]))
          table[i]["sym"][k] = self.n_symbols + len(sym) - 1
          table[i]["scalar"][k] = 1.0
          table[i]["cnst"][k] = 0.0
        else:
          blocked.append(s)
    self.table = table
    self.sym = torch.tensor(sym, dtype=torch.int64)
    self.scalar = torch.tensor(scalar, dtype=torch.float32)
    self.const = torch.tensor(const, dtype=torch.float32)
    self.blocked = sorted(set(blocked))
    self.total_symbols = self.n_symbols + len(sym)

  def values(self, symbol_values):
    """phi [P] -> [P + G] values of the occurrence table."""
    dev = symbol_values.device
    phi = symbol_values.detach().float()
    occ = self.scalar.to(dev) * phi[self.sym.to(dev)] + self.const.to(dev)
    return torch.cat([phi, occ]).contiguous()


class _ParameterShiftOp(torch.autograd.Function):
  """Zero in the forward pass; in the backward pass the parameter-shift estimate of
  d estimator / d symbol_values contracted with the upstream gradient, every shifted circuit
  measured with fresh shots (reference qnn.py:188-228)."""

  @staticmethod
  def forward(ctx, symbol_values, occ, estimator, shape):
    ctx.occ, ctx.estimator = occ, estimator
    ctx.save_for_backward(symbol_values)
    return torch.zeros(shape, dtype=torch.float32, device=symbol_values.device)

  @staticmethod
  def backward(ctx, grad_out):
    (symbol_values,) = ctx.saved_tensors
    occ = ctx.occ
    if occ.blocked:
      raise NotImplementedError(
          "parameter-shift gradients need two-eigenvalue gates; decompose the gates carrying symbol index "
          f"{occ.blocked} (ISwapPow / FSim / PhasedISwapPow / phase exponents) first")
    grad = torch.zeros_like(symbol_values, dtype=torch.float32)
    base = occ.values(symbol_values)
    with torch.no_grad():
      for g in range(len(occ.sym)):
        shifted = base.clone()
        shifted[occ.n_symbols + g] += 0.5
        plus = ctx.estimator(shifted)
        shifted[occ.n_symbols + g] -= 1.0
        minus = ctx.estimator(shifted)
        weight = 0.5 * math.pi * float(occ.scalar[g])
        grad[int(occ.sym[g])] += weight * torch.sum(grad_out * (plus - minus))
    return grad.to(symbol_values.dtype), None, None, None


class RaggedBitstrings:
  """Rows of int8 bitstrings grouped by source state (stands in for the tf.RaggedTensor returned
  by the reference `_sample`): `ragged[i]` is int8 [counts[i], num_qubits]."""

  def __init__(self, flat_values, row_splits):
    self.flat_values = flat_values
    self.row_splits = row_splits
    self._splits = row_splits.tolist()

  def __len__(self):
    return len(self._splits) - 1

  def __getitem__(self, i):
    return self.flat_values[self._splits[i]:self._splits[i + 1]]

  def row_lengths(self):
    return self.row_splits[1:] - self.row_splits[:-1]


class SampledQuantumInference(QuantumInference):
  """Expectation values estimated from measurement shots, differentiated by parameter shift.

  Stands where the reference drives `tfq.layers.Sample`, `tfq.layers.SampledExpectation` and
  `tfq.differentiators.ParameterShift` (qnn.py:142-292).  Final states come from the same sweep
  kernels as the analytic path; shots are drawn on the GPU (`qhbm_sample_states`), and Pauli-term
  estimates are exact binomial draws around the simulated term expectation (`qhbm_binomial_shots`),
  which is the distribution of the mean of `expectation_samples` independent +-1 outcomes."""

  _STATE_BUDGET = 1 << 27  # amplitudes per chunk of final states (1 GiB of complex64)

  def __init__(self, input_circuit, expectation_samples, name=None, initial_seed=None):
    super().__init__(input_circuit, name)
    self._expectation_samples = int(expectation_samples)
    if initial_seed is None:
      initial_seed = int.from_bytes(os.urandom(8), "little")
    self._seed_state = int(initial_seed) & ((1 << 64) - 1)
    self._plans = _LRU(8)

  def _next_seed(self):
    """A fresh (seed0, seed1) per measurement (splitmix64 stream)."""
    out = []
    for _ in range(2):
      self._seed_state = (self._seed_state + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
      z = self._seed_state
      z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
      z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & ((1 << 64) - 1)
      out.append(z ^ (z >> 31))
    return tuple(out)

  # ------------------------------------------------------------------ plans
  def _compiled(self, circuit, key_obj, term_ops):
    """(occurrence table, plan) for `circuit`; the plan measures `term_ops` (an OperatorTensor with
    one Pauli string per entry) or, when None, is used for final states only."""
    qubits = circuit.qubits
    if term_ops is None:
      term_ops = cq.convert_to_tensor([cq.PauliSum.from_pauli_strings(cq.Z(qubits[0]))])
    del key_obj

    def make():
      occ = _Occurrences(circuit.gate_table(), len(circuit.symbol_names))
      terms, offsets = term_ops.tables(qubits)
      return occ, engine.ExpectationPlan(occ.table, len(qubits), occ.total_symbols, terms, offsets, False)

    return self._plans.get((circuit.gate_table_digest(), term_ops.tables_digest(qubits)), make)

  @staticmethod
  def _split_terms(ops, qubits):
    """Observables [O] -> (one unit-coefficient OperatorTensor entry per non-identity Pauli string,
    mixing matrix f32[T, O], identity offsets f32[O])."""
    strings, rows, offsets = [], [], [0.0] * len(ops)
    for j, pauli_sum in enumerate(ops.pauli_sums):
      for t in pauli_sum.terms:
        if abs(t.coefficient.imag) > 1e-12 * max(1.0, abs(t.coefficient)):
          raise ValueError("PauliSum coefficients must be real (Hermitian observables)")
        if not t.paulis:
          offsets[j] += t.coefficient.real
          continue
        strings.append(cq.PauliSum.from_pauli_strings(cq.PauliString(1.0, dict(t.paulis))))
        rows.append((len(strings) - 1, j, t.coefficient.real))
    mix = torch.zeros((max(len(strings), 1), len(ops)), dtype=torch.float32)
    for t, j, c in rows:
      mix[t, j] = c
    if not strings:  # only identity terms: measure a dummy Z with zero weight
      strings.append(cq.PauliSum.from_pauli_strings(cq.Z(qubits[0])))
    return cq.convert_to_tensor(strings), mix, torch.tensor(offsets, dtype=torch.float32)

  # ------------------------------------------------------------------ estimators
  def _pauli_estimator(self, circuits, ops):
    qubits = circuits.circuit.qubits
    cache = self.__dict__.setdefault("_split_cache", _LRU(8))
    term_ops, mix, offsets = cache.get((ops.tables_digest(qubits), len(qubits)),
                                       lambda: self._split_terms(ops, qubits))
    occ, plan = self._compiled(circuits.circuit, ops, term_ops)
    basis_idx = circuits.basis_idx
    dev = basis_idx.device
    mix_d, offsets_d = mix.to(dev), offsets.to(dev)

    def estimate(values):
      exact_terms = plan.forward(basis_idx, values)
      noisy = engine.binomial_shots(exact_terms, self._expectation_samples, self._next_seed())
      return noisy @ mix_d + offsets_d

    return occ, estimate

  def _final_state_chunks(self, plan, basis_idx, values):
    n = plan.n_qubits
    step = max(1, self._STATE_BUDGET >> n)
    for lo in range(0, basis_idx.shape[0], step):
      yield lo, plan.final_states(basis_idx[lo:lo + step].contiguous(), values)

  def _bitstring_estimator(self, circuits, observables):
    occ, plan = self._compiled(circuits.circuit, observables, None)
    basis_idx = circuits.basis_idx
    n = plan.n_qubits
    shots = self._expectation_samples
    shifts = utils._natural_shifts(n)

    def estimate(values):
      """mean_k E(x_k) over `shots` measured bitstrings of every state; E is evaluated once per
      distinct (state, bitstring) pair and differentiable w.r.t. the energy's variables."""
      pieces = []
      for lo, states in self._final_state_chunks(plan, basis_idx, values):
        u = states.shape[0]
        counts = torch.full((u,), shots, dtype=torch.int64, device=states.device)
        keys, _ = engine.sample_states(states, counts, self._next_seed())
        owner = torch.arange(u, dtype=torch.int64, device=states.device).repeat_interleave(shots)
        uniq, _, cnt = engine.unique_with_counts((owner << n) | keys)
        energies = observables.energy(engine.unpack_bits(uniq & ((1 << n) - 1), n, shifts)).float()
        weighted = energies * (cnt.to(torch.float32) / float(shots))
        pieces.append(torch.zeros(u, dtype=torch.float32, device=states.device).index_add(0, uniq >> n, weighted))
      return torch.cat(pieces).unsqueeze(1)

    return occ, estimate

  def _expectation(self, circuits, symbol_names, symbol_values, observables):
    del symbol_names
    if isinstance(observables, cq.OperatorTensor):
      occ, estimate = self._pauli_estimator(circuits, observables)
      post_process = lambda x: x
    elif isinstance(observables, hamiltonian_lib.Hamiltonian) and isinstance(observables.energy,
                                                                          energy_lib.PauliMixin):
      occ, estimate = self._pauli_estimator(circuits, observables.operator_shards)
      post_process = lambda y: observables.energy.operator_expectation(y).unsqueeze(-1)
    else:
      occ, estimate = self._bitstring_estimator(circuits, observables)
      post_process = lambda x: x
    values = symbol_values if symbol_values.is_cuda else symbol_values.to(circuits.basis_idx.device)
    forward_pass = estimate(occ.values(values))
    if torch.is_grad_enabled() and values.requires_grad:
      forward_pass = forward_pass + _ParameterShiftOp.apply(values, occ, estimate, tuple(forward_pass.shape))
    return post_process(forward_pass)

  def _sample(self, initial_states, counts):
    """`ragged[i]` holds `counts[i]` bitstrings measured on circuit|initial_states[i]>
    (reference qnn.py:262-292)."""
    circuits = self.circuit(initial_states)
    occ, plan = self._compiled(self.circuit, None, None)
    values = self.circuit.symbol_values
    if not values.is_cuda:
      values = values.to(circuits.basis_idx.device)
    values = occ.values(values)
    n = plan.n_qubits
    counts = counts.to(device=circuits.basis_idx.device, dtype=torch.int64)
    pieces = []
    for lo, states in self._final_state_chunks(plan, circuits.basis_idx, values):
      keys, _ = engine.sample_states(states, counts[lo:lo + states.shape[0]], self._next_seed())
      pieces.append(keys)
    keys = torch.cat(pieces) if pieces else torch.zeros(0, dtype=torch.int64, device=counts.device)
    splits = torch.zeros(counts.shape[0] + 1, dtype=torch.int64, device=counts.device)
    splits[1:] = torch.cumsum(counts, 0)
    return RaggedBitstrings(engine.unpack_bits(keys, n, utils._natural_shifts(n)), splits)
