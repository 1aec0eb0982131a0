"""Executes the TensorFlow and cirq adapters against minimal stand-in modules.

TensorFlow, TFQ and cirq cannot be installed in this image, so the binding code could not run at
all; these tests inject tiny fake `tensorflow` / `cirq` modules into sys.modules that implement exactly the
calls the adapters make (DLPack hand-over, tf.custom_gradient, tf.identity; the cirq gate classes and
Circuit.all_operations), and check that the adapters wire them correctly.  They prove the code parses,
runs and routes shapes / gradients properly -- not that a real TensorFlow accepts the DLPack capsules."""
import importlib
import sys
import types

import numpy as np
import pytest
import sympy
import torch

from qhbmlib import circuits as cq


# ------------------------------------------------------------------------------------------ fake tensorflow
class _FakeTFTensor:
  """Owns a torch tensor; DLPack round trips go through torch."""

  def __init__(self, t):
    self.t = t
    self.shape = tuple(t.shape)


def _fake_tensorflow():
  tf = types.ModuleType("tensorflow")
  calls = {"to_dlpack": 0, "from_dlpack": 0, "identity": 0}

  def to_dlpack(x):
    calls["to_dlpack"] += 1
    return torch.utils.dlpack.to_dlpack(x.t)

  def from_dlpack(capsule):
    calls["from_dlpack"] += 1
    return _FakeTFTensor(torch.utils.dlpack.from_dlpack(capsule))

  def identity(x):
    calls["identity"] += 1
    return _FakeTFTensor(x.t.clone())

  def custom_gradient(f):
    def wrapped(*args):
      out, grad_fn = f(*args)
      out.grad_fn = grad_fn  # what tf.GradientTape would call with the upstream gradient
      return out
    return wrapped

  tf.experimental = types.SimpleNamespace(dlpack=types.SimpleNamespace(to_dlpack=to_dlpack, from_dlpack=from_dlpack))
  tf.identity = identity
  tf.custom_gradient = custom_gradient
  tf.float32 = "float32"
  tf._calls = calls
  return tf


class _FakePlan:
  """Duck-typed ExpectationPlan on CPU tensors: <op_j>_u = sum_p phi_p * (u + 1) * (j + 1)."""
  n_ops = 2

  def forward(self, basis, values):
    u = basis.to(torch.float32) + 1.0
    j = torch.arange(1, self.n_ops + 1, dtype=torch.float32)
    return values.sum() * u[:, None] * j[None, :]

  def forward_adjoint(self, basis, values, upstream, per_state=False, grad_mode="exact"):
    assert grad_mode in ("exact", "tfq_fd") and per_state == (values.dim() == 2)
    u = basis.to(torch.float32) + 1.0
    j = torch.arange(1, self.n_ops + 1, dtype=torch.float32)
    if per_state:  # rows of symbol values: <op_j>_u = sum_p phi_up (u + 1)(j + 1)
      g = (upstream * u[:, None] * j[None, :]).sum(1, keepdim=True) * torch.ones_like(values)
      return values.sum(1)[:, None] * u[:, None] * j[None, :], g
    g = (upstream * u[:, None] * j[None, :]).sum() * torch.ones_like(values)
    return self.forward(basis, values), g


def test_tf_adapter_runs_against_a_fake_tensorflow(monkeypatch):
  fake = _fake_tensorflow()
  monkeypatch.setitem(sys.modules, "tensorflow", fake)
  import qhbmlib.tf_adapter as adapter
  adapter = importlib.reload(adapter)
  try:
    assert adapter.available()
    basis = _FakeTFTensor(torch.tensor([0, 1, 2], dtype=torch.int64))
    phi = _FakeTFTensor(torch.tensor([0.5, -0.25, 1.0]))
    out = adapter.expectation(_FakePlan(), basis, phi, grad_mode="tfq_fd")
    assert isinstance(out, _FakeTFTensor) and out.shape == (3, 2)
    np.testing.assert_allclose(out.t.numpy(), 1.25 * np.outer([1, 2, 3], [1, 2]))
    upstream = _FakeTFTensor(torch.ones(3, 2))
    g = out.grad_fn(upstream)
    assert isinstance(g, _FakeTFTensor) and g.shape == (3,)
    np.testing.assert_allclose(g.t.numpy(), np.full(3, 18.0))  # sum_u (u+1) sum_j (j+1) = 6 * 3
    # inputs crossed as consumed DLPack capsules, outputs were allocated on the adapter's side
    assert fake._calls["to_dlpack"] == 3 and fake._calls["from_dlpack"] == 2 and fake._calls["identity"] == 1
    # one row of symbol values per state: the gradient keeps the [U, P] shape of the TFQ op
    rows = _FakeTFTensor(torch.tensor([[0.5, -0.25, 1.0]]).repeat(3, 1))
    out_r = adapter.expectation(_FakePlan(), basis, rows, grad_mode="tfq_fd")
    g_r = out_r.grad_fn(upstream)
    assert g_r.shape == (3, 3)
    np.testing.assert_allclose(g_r.t.numpy(), np.outer([3.0, 6.0, 9.0], np.ones(3)))
  finally:
    monkeypatch.delitem(sys.modules, "tensorflow")
    importlib.reload(adapter)


def test_tf_adapter_without_tensorflow_raises():
  import qhbmlib.tf_adapter as adapter
  if adapter.available():
    pytest.skip("TensorFlow is installed")
  with pytest.raises(ImportError, match="TensorFlow is required"):
    adapter.expectation(_FakePlan(), None, None)


# ------------------------------------------------------------------------------------------------ fake cirq
def _fake_cirq():
  cirq = types.ModuleType("cirq")

  class GridQubit:
    def __init__(self, row, col):
      self.row, self.col = row, col

  class _Eigen:
    def __init__(self, exponent=1.0, global_shift=0.0):
      self.exponent, self.global_shift = exponent, global_shift

  names = ["XPowGate", "YPowGate", "ZPowGate", "HPowGate", "CZPowGate", "CNotPowGate", "SwapPowGate",
           "ISwapPowGate", "XXPowGate", "YYPowGate", "ZZPowGate"]
  for nm in names:
    setattr(cirq, nm, type(nm, (_Eigen,), {}))

  class IdentityGate:
    pass

  class PhasedXPowGate:
    def __init__(self, phase_exponent, exponent=1.0, global_shift=0.0):
      self.phase_exponent, self.exponent, self.global_shift = phase_exponent, exponent, global_shift

  class FSimGate:
    def __init__(self, theta, phi):
      self.theta, self.phi = theta, phi

  class PhasedISwapPowGate:
    def __init__(self, phase_exponent, exponent=1.0):
      self.phase_exponent, self.exponent = phase_exponent, exponent

  class ControlledGate:
    def __init__(self, sub_gate, num_controls=1, control_values=((1,),)):
      self.sub_gate, self._n, self.control_values = sub_gate, num_controls, control_values

    def num_controls(self):
      return self._n

  class Operation:
    def __init__(self, gate, qubits):
      self.gate, self.qubits = gate, qubits

  class Circuit:
    def __init__(self, ops):
      self._ops = list(ops)

    def all_operations(self):
      return iter(self._ops)

  class PauliSum(list):
    pass

  class PauliString:
    def __init__(self, coefficient, paulis):
      self.coefficient, self._p = coefficient, paulis

    def items(self):
      return self._p.items()

  for cls in (GridQubit, IdentityGate, PhasedXPowGate, FSimGate, PhasedISwapPowGate, ControlledGate, Operation,
              Circuit, PauliSum, PauliString):
    setattr(cirq, cls.__name__, cls)
  return cirq


def test_from_cirq_converts_the_tfq_gate_set(monkeypatch):
  cirq = _fake_cirq()
  monkeypatch.setitem(sys.modules, "cirq", cirq)
  q = [cirq.GridQubit(0, k) for k in range(3)]
  s = sympy.Symbol("a")
  circuit = cirq.Circuit([
      cirq.Operation(cirq.XPowGate(exponent=s), (q[0],)),
      cirq.Operation(cirq.ZPowGate(exponent=0.25, global_shift=-0.5), (q[1],)),
      cirq.Operation(cirq.CZPowGate(exponent=2 * s), (q[0], q[1])),
      cirq.Operation(cirq.PhasedXPowGate(0.3, 0.7), (q[2],)),
      cirq.Operation(cirq.FSimGate(0.1, 0.2), (q[1], q[2])),
      cirq.Operation(cirq.PhasedISwapPowGate(0.4, 0.5), (q[0], q[2])),
      cirq.Operation(cirq.IdentityGate(), (q[2],)),
      cirq.Operation(cirq.ControlledGate(cirq.XPowGate(exponent=0.5)), (q[0], q[1])),   # -> CNOT**0.5
      cirq.Operation(cirq.ControlledGate(cirq.ZPowGate(exponent=s)), (q[1], q[2])),     # -> CZ**a
  ])
  out = cq.from_cirq(circuit)
  qubits = sorted(out.all_qubits())
  table = cq.gate_table(out, qubits, ["a"])
  assert [int(t) for t in table["type"]] == [1, 3, 5, 12, 13, 14, 0, 6, 5]
  assert table["sym"][0][0] == 0 and table["scalar"][0][0] == 1.0            # X**a
  assert table["sym"][2][0] == 0 and table["scalar"][2][0] == 2.0            # CZ**(2a)
  assert table["sym"][1][0] == -1 and abs(table["cnst"][1][0] - 0.25) < 1e-7 and table["gshift"][1] == -0.5
  assert (int(table["q0"][7]), int(table["q1"][7])) == (0, 1) and abs(table["cnst"][7][0] - 0.5) < 1e-7
  assert table["sym"][8][0] == 0
  ps = cirq.PauliSum([cirq.PauliString(1.5, {q[0]: "X", q[2]: "Z"}), cirq.PauliString(-0.5, {q[1]: "Y"})])
  conv = cq.from_cirq(ps)
  assert isinstance(conv, cq.PauliSum) and len(conv) == 2


def test_from_cirq_rejects_controlled_gates_outside_the_gate_table(monkeypatch):
  """The gate table has no control_qubits field (TFQ's serializer has): anything that is not CNOT**t / CZ**t
  is refused with a message that asks for a decomposition."""
  cirq = _fake_cirq()
  monkeypatch.setitem(sys.modules, "cirq", cirq)
  q = [cirq.GridQubit(0, k) for k in range(3)]
  for gate in (cirq.ControlledGate(cirq.YPowGate(exponent=0.5)),                        # controlled-Y
               cirq.ControlledGate(cirq.XPowGate(exponent=0.5), num_controls=2),         # Toffoli-like
               cirq.ControlledGate(cirq.XPowGate(exponent=0.5), control_values=((0,),)),  # control on |0>
               cirq.ControlledGate(cirq.XPowGate(exponent=0.5, global_shift=-0.5))):     # controlled phase
    with pytest.raises(ValueError, match="decompose"):
      cq.from_cirq(cirq.Circuit([cirq.Operation(gate, tuple(q[:2 if gate.num_controls() == 1 else 3]))]))

  class Weird:
    pass
  with pytest.raises(ValueError, match="outside the TFQ-serialisable set"):
    cq.from_cirq(cirq.Circuit([cirq.Operation(Weird(), (q[0],))]))
