"""Inference on parameterised circuits (mirror of /root/reference/qhbmlib/inference/qnn.py).

`AnalyticQuantumInference.expectation` is the entry point of the B200 hot path: unique initial
states -> basis indices -> one compiled plan per (circuit, observables) -> CUDA expectation kernels,
with the adjoint-gradient kernels wired into torch autograd (the reference wires TFQ's adjoint
differentiator into TF autodiff, qnn.py:87-139).
"""
import abc

import torch

from qhbmlib import circuits as cq
from qhbmlib import engine
from qhbmlib import utils
from qhbmlib.models import energy as energy_lib
from qhbmlib.models import hamiltonian as hamiltonian_lib


class _ExpectationOp(torch.autograd.Function):
  """f32[U, O] expectations; backward = adjoint gradient w.r.t. the symbol values.
  Like TFQ, the backward pass re-simulates the forward circuit."""

  @staticmethod
  def forward(ctx, symbol_values, basis_idx, holder):
    vals = holder.forward_plan.forward(basis_idx, symbol_values.detach().contiguous().float())
    ctx.holder = holder
    ctx.save_for_backward(symbol_values, basis_idx)
    return vals

  @staticmethod
  def backward(ctx, grad_out):
    symbol_values, basis_idx = ctx.saved_tensors
    holder = ctx.holder
    if symbol_values.numel() == 0:
      return torch.zeros_like(symbol_values), None, None
    _, grad = holder.plan.forward_adjoint(basis_idx, symbol_values.detach().contiguous().float(),
                                          grad_out.contiguous().float(), per_state=False,
                                          grad_mode=holder.grad_mode)
    return grad.to(symbol_values.dtype), None, None


class _PlanHolder:
  """The adjoint plan, plus a forward-only plan (wider register blocking, no lambda tile) compiled
  lazily for the forward pass of autograd."""

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
    unique_states, idx, _ = utils.unique_bitstrings_with_counts(initial_states)
    if isinstance(observables, cq.OperatorTensor):
      total_circuit = self.circuit
    else:
      total_circuit = self._total_circuit(observables)
    circuits = total_circuit(unique_states)
    unique_expectations = self._expectation(circuits, total_circuit.symbol_names, total_circuit.symbol_values,
                                            observables)
    return utils.expand_unique_results(unique_expectations, idx)

  def _total_circuit(self, observables):
    cache = self.__dict__.setdefault("_total_cache", {})
    key = id(observables)
    if key not in cache:
      cache[key] = (observables, self.circuit + observables.circuit_dagger)
    return cache[key][1]

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
    self._plans = {}

  def _plan_for(self, circuit, ops_tensor):
    key = (id(circuit), id(ops_tensor))
    hit = self._plans.get(key)
    if hit is None:
      terms, offsets = ops_tensor.tables(circuit.qubits)
      args = (circuit.gate_table(), len(circuit.qubits), len(circuit.symbol_names), terms, offsets)
      plan = engine.ExpectationPlan(*args, True, self._tile_qubits, self._reg_qubits)
      make_fwd = lambda: engine.ExpectationPlan(*args, False, 0, 0)
      hit = (circuit, ops_tensor, _PlanHolder(plan, self.grad_mode, make_fwd))  # keep the keys alive
      self._plans[key] = hit
    hit[2].grad_mode = self.grad_mode
    return hit[2]

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
