"""Circuit and Pauli-operator construction without cirq / TFQ.

The reference builds `cirq.Circuit`s with `sympy` symbols and hands them to TensorFlow
Quantum as serialized protos (qhbmlib/models/circuit.py:30-63).  Neither cirq nor TFQ exist in
this environment, and the B200 engine consumes a flat gate table instead of protos, so this
module offers the small cirq-shaped surface qhbmlib code actually uses -- qubits, the
TFQ-serialisable gate set, `Circuit`, `PauliString`/`PauliSum` -- and lowers
it to the C ABI's tables (include/qhbm_b200.h).  `from_cirq` converts real cirq objects when
cirq is importable.
"""
import functools
import math
import numbers

import hashlib

import numpy as np

from qhbmlib import _native as nat

try:  # sympy symbols are accepted wherever a gate parameter is expected
  import sympy as _sympy
except Exception:  # pragma: no cover
  _sympy = None

# ----------------------------------------------------------------------------- qubits


@functools.total_ordering
class GridQubit:
  """A qubit at integer (row, col); sorts row-major like cirq.GridQubit."""
  __slots__ = ("row", "col")

  def __init__(self, row, col):
    self.row, self.col = int(row), int(col)

  @staticmethod
  def rect(rows, cols, top=0, left=0):
    return [GridQubit(r, c) for r in range(top, top + rows) for c in range(left, left + cols)]

  def _key(self):
    return (self.row, self.col)

  def __eq__(self, other):
    return isinstance(other, GridQubit) and self._key() == other._key()

  def __lt__(self, other):
    return self._key() < other._key()

  def __hash__(self):
    return hash(("GridQubit",) + self._key())

  def __repr__(self):
    return f"GridQubit({self.row}, {self.col})"


def LineQubit(x):  # pylint: disable=invalid-name
  return GridQubit(0, x)


LineQubit.range = lambda n: [GridQubit(0, i) for i in range(n)]

# ----------------------------------------------------------------------------- parameters


class Symbol:
  """Free circuit parameter; arithmetic gives `scalar * symbol + const` expressions."""
  __slots__ = ("name",)

  def __init__(self, name):
    self.name = str(name)

  def __repr__(self):
    return self.name

  def __hash__(self):
    return hash(self.name)

  def __eq__(self, other):
    return isinstance(other, Symbol) and other.name == self.name

  def _lin(self):
    return Linear(self.name, 1.0, 0.0)

  def __mul__(self, k):
    return self._lin() * k

  __rmul__ = __mul__

  def __truediv__(self, k):
    return self._lin() / k

  def __neg__(self):
    return self._lin() * -1.0

  def __add__(self, k):
    return self._lin() + k

  __radd__ = __add__

  def __sub__(self, k):
    return self._lin() + (-k)


class Linear:
  """scalar * symbol + const (TFQ: `exponent_scalar`, `exponent`)."""
  __slots__ = ("symbol", "scalar", "const")

  def __init__(self, symbol, scalar, const):
    self.symbol, self.scalar, self.const = symbol, float(scalar), float(const)

  def __mul__(self, k):
    return Linear(self.symbol, self.scalar * float(k), self.const * float(k))

  __rmul__ = __mul__

  def __truediv__(self, k):
    return self * (1.0 / float(k))

  def __neg__(self):
    return self * -1.0

  def __add__(self, k):
    return Linear(self.symbol, self.scalar, self.const + float(k))

  __radd__ = __add__

  def __sub__(self, k):
    return self + (-float(k))

  def __repr__(self):
    return f"{self.scalar}*{self.symbol}+{self.const}"


def symbols(names):
  """`sympy.symbols`-like helper: "a b c" -> tuple of Symbols."""
  out = tuple(Symbol(s) for s in str(names).replace(",", " ").split())
  return out[0] if len(out) == 1 else out


def as_param(value):
  """Normalises a gate parameter to Linear (symbol may be None for constants)."""
  if isinstance(value, Linear):
    return value
  if isinstance(value, Symbol):
    return value._lin()
  if isinstance(value, numbers.Real):
    return Linear(None, 0.0, float(value))
  if _sympy is not None and isinstance(value, _sympy.Basic):
    free = sorted(value.free_symbols, key=str)
    if not free:
      return Linear(None, 0.0, float(value))
    if len(free) > 1:
      raise ValueError("gate parameters must depend on at most one symbol (TFQ restriction)")
    s = free[0]
    scalar = value.coeff(s)
    const = value.subs(s, 0)
    if _sympy.simplify(value - (scalar * s + const)) != 0 or scalar.free_symbols:
      raise ValueError(f"gate parameter {value} is not of the form scalar*symbol + const")
    return Linear(str(s), float(scalar), float(const))
  if hasattr(value, "item"):
    return Linear(None, 0.0, float(value.item()))
  raise TypeError(f"unsupported gate parameter {value!r}")


# ----------------------------------------------------------------------------- gates

_TWO_QUBIT = {nat_t for nat_t in (5, 6, 7, 8, 9, 10, 11, 13, 14)}
_NAMES = {0: "I", 1: "X", 2: "Y", 3: "Z", 4: "H", 5: "CZ", 6: "CNOT", 7: "SWAP", 8: "ISWAP", 9: "XX",
          10: "YY", 11: "ZZ", 12: "PhasedX", 13: "FSim", 14: "PhasedISwap"}


class Gate:
  """One member of the TFQ-serialisable gate set.  `params` are Linear expressions:
  eigen-gates (exponent,), PhasedX / PhasedISwap (exponent, phase_exponent), FSim (theta, phi)."""

  def __init__(self, kind, params=(), global_shift=0.0):
    self.kind = int(kind)
    self.params = tuple(as_param(p) for p in params)
    self.global_shift = float(global_shift)

  @property
  def num_qubits(self):
    return 2 if self.kind in _TWO_QUBIT else 1

  def on(self, *qubits):
    if len(qubits) != self.num_qubits:
      raise ValueError(f"{_NAMES[self.kind]} acts on {self.num_qubits} qubit(s), got {len(qubits)}")
    return Operation(self, tuple(qubits))

  __call__ = on

  def __pow__(self, exponent):
    if self.kind == 0:
      return self
    if self.kind == 13:  # FSim: only the inverse is meaningful
      if exponent == -1:
        return Gate(13, (-1.0 * self.params[0], -1.0 * self.params[1]))
      raise ValueError("FSimGate only supports exponent -1")
    e = as_param(exponent)
    base = self.params[0]
    if base.symbol is not None and e.symbol is not None:
      raise ValueError("cannot raise a symbolic gate to a symbolic power")
    if e.symbol is None:
      new = base * e.const
    else:
      new = e * base.const
    return Gate(self.kind, (new,) + self.params[1:], self.global_shift)

  def __repr__(self):
    return f"{_NAMES[self.kind]}({', '.join(map(repr, self.params))}, shift={self.global_shift})"


def XPowGate(exponent=1.0, global_shift=0.0):  # pylint: disable=invalid-name
  return Gate(nat_type("XPOW"), (exponent,), global_shift)


def nat_type(name):
  return {"I": 0, "XPOW": 1, "YPOW": 2, "ZPOW": 3, "HPOW": 4, "CZPOW": 5, "CNOTPOW": 6, "SWAPPOW": 7,
          "ISWAPPOW": 8, "XXPOW": 9, "YYPOW": 10, "ZZPOW": 11, "PHASEDXPOW": 12, "FSIM": 13,
          "PHASEDISWAPPOW": 14}[name]


def _eigen(kind):
  def make(exponent=1.0, global_shift=0.0):
    return Gate(kind, (exponent,), global_shift)
  return make


YPowGate, ZPowGate, HPowGate = _eigen(2), _eigen(3), _eigen(4)
CZPowGate, CNotPowGate, SwapPowGate, ISwapPowGate = _eigen(5), _eigen(6), _eigen(7), _eigen(8)
XXPowGate, YYPowGate, ZZPowGate = _eigen(9), _eigen(10), _eigen(11)


def PhasedXPowGate(phase_exponent, exponent=1.0, global_shift=0.0):  # pylint: disable=invalid-name
  return Gate(12, (exponent, phase_exponent), global_shift)


def FSimGate(theta, phi):  # pylint: disable=invalid-name
  return Gate(13, (theta, phi))


def PhasedISwapPowGate(phase_exponent=0.25, exponent=1.0):  # pylint: disable=invalid-name
  return Gate(14, (exponent, phase_exponent))


def rx(rads):
  """cirq.rx: XPowGate(exponent=rads/pi, global_shift=-0.5)."""
  return Gate(1, (as_param(rads) * (1.0 / math.pi),), -0.5)


def ry(rads):
  return Gate(2, (as_param(rads) * (1.0 / math.pi),), -0.5)


def rz(rads):
  return Gate(3, (as_param(rads) * (1.0 / math.pi),), -0.5)


I, X, Y, Z, H = Gate(0), Gate(1, (1.0,)), Gate(2, (1.0,)), Gate(3, (1.0,)), Gate(4, (1.0,))
CZ, CNOT, SWAP, ISWAP = Gate(5, (1.0,)), Gate(6, (1.0,)), Gate(7, (1.0,)), Gate(8, (1.0,))
XX, YY, ZZ = Gate(9, (1.0,)), Gate(10, (1.0,)), Gate(11, (1.0,))


class Operation:
  """A gate applied to qubits.  Pauli operations also act as PauliStrings in products."""

  def __init__(self, gate, qubits):
    self.gate, self.qubits = gate, tuple(qubits)

  def __pow__(self, exponent):
    return Operation(self.gate**exponent, self.qubits)

  def _pauli(self):
    g = self.gate
    if g.kind in (1, 2, 3) and g.params[0].symbol is None and g.params[0].const == 1.0 and g.global_shift == 0.0:
      return PauliString(1.0, {self.qubits[0]: "XYZ"[g.kind - 1]})
    if g.kind == 0:
      return PauliString(1.0, {})
    raise TypeError("only plain X, Y, Z operations form Pauli strings")

  def __mul__(self, other):
    return self._pauli() * other

  def __rmul__(self, other):
    return self._pauli().__rmul__(other)

  def __add__(self, other):
    return self._pauli() + other

  def __radd__(self, other):
    return self._pauli().__radd__(other)

  def __sub__(self, other):
    return self._pauli() - other

  def __neg__(self):
    return -self._pauli()

  def __repr__(self):
    return f"{self.gate!r}.on{self.qubits}"


def _flatten_ops(items):
  for it in items:
    if isinstance(it, Operation):
      yield it
    elif isinstance(it, Circuit):
      yield from it.operations
    elif isinstance(it, (list, tuple)) or hasattr(it, "__iter__"):
      yield from _flatten_ops(it)
    else:
      raise TypeError(f"cannot add {it!r} to a Circuit")


class Circuit:
  """Ordered list of operations (moment structure is irrelevant to a state-vector engine)."""

  def __init__(self, *items):
    self.operations = list(_flatten_ops(items))

  def append(self, item):
    self.operations.extend(_flatten_ops([item]))

  def __add__(self, other):
    return Circuit(self, other)

  def __iadd__(self, other):
    self.append(other)
    return self

  def __pow__(self, exponent):
    if exponent != -1:
      raise ValueError("only the inverse of a circuit is defined")
    return Circuit([op**-1 for op in reversed(self.operations)])

  def all_qubits(self):
    return frozenset(q for op in self.operations for q in op.qubits)

  def all_operations(self):
    return iter(self.operations)

  def __len__(self):
    return len(self.operations)

  def __eq__(self, other):
    return isinstance(other, Circuit) and repr(self) == repr(other)

  def __repr__(self):
    return "Circuit(" + ", ".join(map(repr, self.operations)) + ")"


def circuit_symbols(circuit):
  """Names of the free symbols (tfq.util.get_circuit_symbols)."""
  out = set()
  for op in circuit.operations:
    for p in op.gate.params:
      if p.symbol is not None:
        out.add(p.symbol)
  return out


def gate_table(circuit, qubits, symbol_names):
  """Lowers a Circuit to qhbm_gate_t rows.  `qubits` is the sorted qubit list (position k is
  bit n-1-k of the basis index); `symbol_names` fixes the symbol order."""
  qpos = {q: i for i, q in enumerate(qubits)}
  spos = {s: i for i, s in enumerate(symbol_names)}
  rows = np.zeros(len(circuit.operations), dtype=nat.GATE_DTYPE)
  for i, op in enumerate(circuit.operations):
    g = op.gate
    rows[i]["type"] = g.kind
    rows[i]["q0"] = qpos[op.qubits[0]]
    rows[i]["q1"] = qpos[op.qubits[1]] if len(op.qubits) == 2 else -1
    rows[i]["nparams"] = len(g.params)
    rows[i]["gshift"] = g.global_shift
    sym, scalar, cnst = [-1, -1, -1], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    for k, p in enumerate(g.params):
      if p.symbol is not None:
        if p.symbol not in spos:
          raise ValueError(f"symbol {p.symbol} is not in symbol_names")
        sym[k], scalar[k] = spos[p.symbol], p.scalar
      cnst[k] = p.const
    rows[i]["sym"], rows[i]["scalar"], rows[i]["cnst"] = sym, scalar, cnst
  return rows


# ----------------------------------------------------------------------------- Pauli algebra

_PAULI_PRODUCT = {("X", "Y"): (1j, "Z"), ("Y", "X"): (-1j, "Z"), ("Y", "Z"): (1j, "X"),
                  ("Z", "Y"): (-1j, "X"), ("Z", "X"): (1j, "Y"), ("X", "Z"): (-1j, "Y")}


class PauliString:
  """coefficient * prod_q sigma_q."""

  def __init__(self, coefficient=1.0, paulis=None):
    self.coefficient = complex(coefficient)
    self.paulis = dict(paulis or {})

  def __mul__(self, other):
    if isinstance(other, Operation):
      other = other._pauli()
    if isinstance(other, numbers.Number):
      return PauliString(self.coefficient * other, self.paulis)
    if isinstance(other, PauliString):
      coeff = self.coefficient * other.coefficient
      out = dict(self.paulis)
      for q, p in other.paulis.items():
        if q not in out:
          out[q] = p
        elif out[q] == p:
          del out[q]
        else:
          ph, r = _PAULI_PRODUCT[(out[q], p)]
          coeff *= ph
          out[q] = r
      return PauliString(coeff, out)
    return NotImplemented

  def __rmul__(self, other):
    if isinstance(other, numbers.Number):
      return PauliString(self.coefficient * other, self.paulis)
    return NotImplemented

  def __neg__(self):
    return PauliString(-self.coefficient, self.paulis)

  def __add__(self, other):
    return PauliSum([self]) + other

  def __radd__(self, other):
    return PauliSum([self]).__radd__(other)

  def __sub__(self, other):
    return PauliSum([self]) - other

  def _key(self):
    return tuple(sorted((q, p) for q, p in self.paulis.items()))

  def __eq__(self, other):
    return isinstance(other, PauliString) and self._key() == other._key() and self.coefficient == other.coefficient

  def __hash__(self):
    return hash((self._key(), self.coefficient))

  def __repr__(self):
    body = "*".join(f"{p}({q})" for q, p in sorted(self.paulis.items())) or "I"
    return f"{self.coefficient}*{body}"


class PauliSum:
  """Sum of PauliStrings; like terms are combined."""

  def __init__(self, terms=()):
    acc = {}
    for t in terms:
      if isinstance(t, Operation):
        t = t._pauli()
      k = t._key()
      acc[k] = acc.get(k, 0) + t.coefficient
    self.terms = [PauliString(c, dict(k)) for k, c in acc.items() if c != 0]

  @staticmethod
  def from_pauli_strings(strings):
    if isinstance(strings, (PauliString, Operation)):
      strings = [strings]
    return PauliSum(list(strings))

  def _coerce(self, other):
    if isinstance(other, PauliSum):
      return other.terms
    if isinstance(other, (PauliString, Operation)):
      return [other]
    if isinstance(other, numbers.Number):
      return [PauliString(other, {})]
    raise TypeError(f"cannot combine PauliSum with {other!r}")

  def __add__(self, other):
    return PauliSum(self.terms + list(self._coerce(other)))

  __radd__ = __add__

  def __sub__(self, other):
    neg = [(-t if isinstance(t, PauliString) else -t._pauli()) for t in self._coerce(other)]
    return PauliSum(self.terms + neg)

  def __iadd__(self, other):
    return self + other

  def __isub__(self, other):
    return self - other

  def __mul__(self, k):
    if isinstance(k, numbers.Number):
      return PauliSum([t * k for t in self.terms])
    return NotImplemented

  __rmul__ = __mul__

  def __neg__(self):
    return self * -1.0

  def __iter__(self):
    return iter(self.terms)

  def __len__(self):
    return len(self.terms)

  def __eq__(self, other):
    return isinstance(other, PauliSum) and {t._key(): t.coefficient for t in self.terms} == \
        {t._key(): t.coefficient for t in other.terms}

  def __repr__(self):
    return " + ".join(map(repr, self.terms)) or "0"

  def qubits(self):
    return frozenset(q for t in self.terms for q in t.paulis)


class OperatorTensor:
  """Stand-in for `tfq.convert_to_tensor([PauliSum, ...])`: the 1-D "tensor" of observables that
  qhbmlib passes around as `tf.string` (qhbmlib/inference/qnn.py:50-63)."""

  def __init__(self, pauli_sums):
    out = []
    for s in pauli_sums:
      if isinstance(s, (PauliString, Operation)):
        s = PauliSum.from_pauli_strings(s)
      if not isinstance(s, PauliSum):
        raise TypeError("observables must be PauliSums")
      out.append(s)
    self.pauli_sums = out

  @property
  def shape(self):
    return (len(self.pauli_sums),)

  def __len__(self):
    return len(self.pauli_sums)

  def tables(self, qubits):
    """(terms TERM_DTYPE[], offsets int32[O+1]) over the sorted qubit list (memoised per qubit list:
    like a tf.string tensor, the observables do not change after conversion)."""
    memo = self.__dict__.setdefault("_tables_memo", {})
    key = tuple(qubits)
    if key not in memo:
      terms, offsets = self._build_tables(qubits)
      digest = hashlib.sha1(terms.tobytes() + offsets.tobytes()).hexdigest()
      memo.clear()
      memo[key] = (terms, offsets, digest)
    return memo[key][0], memo[key][1]

  def tables_digest(self, qubits):
    """Content hash of `tables(qubits)`: the key under which compiled plans are cached."""
    self.tables(qubits)
    return self._tables_memo[tuple(qubits)][2]

  def _build_tables(self, qubits):
    n = len(qubits)
    qpos = {q: i for i, q in enumerate(qubits)}
    rows, offsets = [], [0]
    for s in self.pauli_sums:
      for t in s.terms:
        if abs(t.coefficient.imag) > 1e-12 * max(1.0, abs(t.coefficient)):
          raise ValueError("PauliSum coefficients must be real (Hermitian observables)")
        x = z = 0
        for q, p in t.paulis.items():
          if q not in qpos:
            raise ValueError(f"observable acts on {q}, which is not a circuit qubit")
          bit = 1 << (n - 1 - qpos[q])
          if p in ("X", "Y"):
            x |= bit
          if p in ("Z", "Y"):
            z |= bit
        rows.append((t.coefficient.real, x, z))
      offsets.append(len(rows))
    terms = np.zeros(len(rows), dtype=nat.TERM_DTYPE)
    for i, r in enumerate(rows):
      terms[i] = r
    return terms, np.asarray(offsets, dtype=np.int32)


def convert_to_tensor(pauli_sums):
  return OperatorTensor(pauli_sums)


# ----------------------------------------------------------------------------- cirq adapter


def from_cirq(obj):
  """Converts a cirq.Circuit / cirq.PauliSum into this module's types (needs cirq).

  cirq is not installable in this image: tests/test_adapters_fake_modules.py runs this function against
  a minimal stand-in `cirq` module.  Gate set = what TFQ 0.6.1 serialises.  Controlled
  operations (`op.controlled_by(...)`, TFQ's control_qubits / control_values) are accepted where they are
  one of the native two-qubit gates -- a singly-controlled X**t is CNOT**t and a singly-controlled Z**t is
  CZ**t (control value 1, zero global shift); any other controlled gate is rejected with a ValueError that
  asks for a decomposition, because the gate table has no control field."""
  import cirq  # pylint: disable=import-outside-toplevel

  def qubit(q):
    return GridQubit(q.row, q.col) if hasattr(q, "row") else GridQubit(0, q.x)

  if isinstance(obj, cirq.PauliSum):
    return PauliSum([PauliString(ps.coefficient, {qubit(q): str(p) for q, p in ps.items()}) for ps in obj])
  kinds = [(cirq.XPowGate, 1), (cirq.YPowGate, 2), (cirq.ZPowGate, 3), (cirq.HPowGate, 4),
           (cirq.CZPowGate, 5), (cirq.CNotPowGate, 6), (cirq.SwapPowGate, 7), (cirq.ISwapPowGate, 8),
           (cirq.XXPowGate, 9), (cirq.YYPowGate, 10), (cirq.ZZPowGate, 11)]
  controlled_gate = getattr(cirq, "ControlledGate", ())
  out = Circuit()
  for op in obj.all_operations():
    g = op.gate
    qs = [qubit(q) for q in op.qubits]
    if controlled_gate and isinstance(g, controlled_gate):
      sub = g.sub_gate
      values = [tuple(v) for v in getattr(g, "control_values", [(1,)])]
      simple = g.num_controls() == 1 and values == [(1,)] and getattr(sub, "global_shift", 0) == 0
      if simple and isinstance(sub, cirq.XPowGate):
        out.append(Gate(6, (sub.exponent,), 0.0).on(*qs))  # CNotPow: control first
        continue
      if simple and isinstance(sub, cirq.ZPowGate):
        out.append(Gate(5, (sub.exponent,), 0.0).on(*qs))  # CZPow
        continue
      raise ValueError(f"controlled gate {g!r} has no entry in the gate table (no control_qubits field): "
                       "decompose it into the TFQ-serialisable one- and two-qubit gates first")
    if isinstance(g, cirq.IdentityGate):
      out.append(I.on(qs[0]))
      continue
    if isinstance(g, cirq.PhasedXPowGate):
      out.append(PhasedXPowGate(g.phase_exponent, g.exponent, g.global_shift).on(*qs))
      continue
    if isinstance(g, cirq.FSimGate):
      out.append(FSimGate(g.theta, g.phi).on(*qs))
      continue
    if isinstance(g, cirq.PhasedISwapPowGate):
      out.append(PhasedISwapPowGate(g.phase_exponent, g.exponent).on(*qs))
      continue
    for cls, kind in kinds:
      if isinstance(g, cls):
        out.append(Gate(kind, (g.exponent,), g.global_shift).on(*qs))
        break
    else:
      raise ValueError(f"gate {g!r} is outside the TFQ-serialisable set")
  return out
