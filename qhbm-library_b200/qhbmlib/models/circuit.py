"""Parameterised quantum circuit models (mirror of /root/reference/qhbmlib/models/circuit.py).

A `QuantumCircuit` holds a `qhbmlib.circuits.Circuit`, its sorted qubits, the ordered symbol
names and the layers that produce the symbol values.  Where the reference serialises one
circuit per bitstring (`tfq.resolve_parameters` + `tfq.append_circuit`, circuit.py:129-136), this
class emits the gate table once and a uint64 basis index per bitstring.
"""
import torch

from qhbmlib import circuits as cq
from qhbmlib import engine
from qhbmlib.models import circuit_utils
from qhbmlib.models import energy_utils


class CircuitBatch:
  """What `QuantumCircuit.call` returns: one circuit applied to many basis states.
  Replaces the `tf.string` tensor of serialized circuits."""

  def __init__(self, circuit, basis_idx):
    self.circuit = circuit
    self.basis_idx = basis_idx

  @property
  def shape(self):
    return (int(self.basis_idx.shape[0]),)

  def __len__(self):
    return int(self.basis_idx.shape[0])


class QuantumCircuit(torch.nn.Module):
  """Circuit + symbol names + the (trainable) map producing their values."""

  def __init__(self, pqc, qubits, symbol_names, value_layers_inputs, value_layers, name=None):
    super().__init__()
    self.name = name if name is not None else "quantum_circuit"
    self._pqc = pqc
    self._qubits = sorted(qubits)
    self._symbol_names = [str(s) for s in symbol_names]
    self._value_layers_inputs = value_layers_inputs
    self._value_layers = value_layers
    # register parameters / sub-layers so that .parameters() sees them (shared, not copied)
    flat_inputs = []
    for inp in value_layers_inputs:
      flat_inputs.extend(inp if isinstance(inp, (list, tuple)) else [inp])
    self._registered_inputs = torch.nn.ParameterList([p for p in flat_inputs if isinstance(p, torch.nn.Parameter)])
    self._registered_layers = torch.nn.ModuleList(
        [l for layers in value_layers for l in layers if isinstance(l, torch.nn.Module)])

    raw_bit_circuit = circuit_utils.bit_circuit(self._qubits)
    # reference circuit.py:59-63: the injector symbols are sorted as STRINGS; column j of a
    # bitstring drives the j-th sorted name, i.e. qubit pi(j).
    self._bit_symbol_names = sorted(cq.circuit_symbols(raw_bit_circuit))
    self._bit_circuit = raw_bit_circuit
    n = len(self._qubits)
    column_to_qubit = [int(s.rsplit("_", 1)[1]) for s in self._bit_symbol_names]
    self._bit_shifts = [n - 1 - q for q in column_to_qubit]
    self._gate_table = None

  @property
  def qubits(self):
    return self._qubits

  @property
  def symbol_names(self):
    return self._symbol_names

  @property
  def value_layers_inputs(self):
    return self._value_layers_inputs

  @property
  def value_layers(self):
    return self._value_layers

  @property
  def symbol_values(self):
    """1-D float tensor: current value of every symbol, in `symbol_names` order."""
    pieces = []
    for inputs, layers in zip(self._value_layers_inputs, self._value_layers):
      x = inputs
      for layer in layers:
        x = layer(x)
      pieces.append(x.reshape(-1))
    if not pieces:
      dev = next(self.parameters()).device if list(self.parameters()) else "cpu"
      return torch.zeros((0,), dtype=torch.float32, device=dev)
    devices = [p.device for p in pieces if p.is_cuda]
    if devices:  # summed circuits may mix parameter devices; the engine wants them on the GPU
      pieces = [p.to(devices[0]) for p in pieces]
    return torch.cat(pieces, 0)

  @property
  def pqc(self):
    return self._pqc

  @property
  def trainable_variables(self):
    return [p for p in self.parameters() if p.requires_grad]

  @property
  def variables(self):
    return list(self.parameters())

  def build(self, input_shape):
    del input_shape

  def gate_table(self):
    """qhbm_gate_t rows of `pqc` over the sorted qubits / ordered symbols (cached)."""
    if self._gate_table is None:
      self._gate_table = cq.gate_table(self._pqc, self._qubits, self._symbol_names)
    return self._gate_table

  def gate_table_digest(self):
    """Content hash of the gate table (+ qubit and symbol counts): the key under which compiled plans
    are cached, so that an equal circuit built again reuses the plan instead of compiling a new one."""
    if getattr(self, "_gate_digest", None) is None:
      import hashlib  # pylint: disable=import-outside-toplevel
      h = hashlib.sha1(self.gate_table().tobytes())
      h.update(f"|{len(self._qubits)}|{len(self._symbol_names)}".encode())
      self._gate_digest = h.hexdigest()
    return self._gate_digest

  def basis_indices(self, bitstrings):
    """int8 [N, n] -> int64 basis index per row, with the reference's column -> qubit map."""
    if bitstrings.shape[1] != len(self._qubits):
      raise ValueError("bitstrings must have one column per qubit")
    return engine.pack_bits(bitstrings.to(torch.int8).contiguous(), self._bit_shifts)

  def forward(self, inputs):
    """Bitstrings prepended as initial basis states to `pqc`."""
    return CircuitBatch(self, self.basis_indices(inputs))

  def __add__(self, other):
    """`self.pqc` followed by `other.pqc`; variables are shared with both operands."""
    if not isinstance(other, QuantumCircuit):
      raise TypeError
    if set(self.symbol_names) & set(other.symbol_names):
      raise ValueError("Circuits to be summed must not have symbols in common.")
    return QuantumCircuit(
        self.pqc + other.pqc, list(set(self.qubits + other.qubits)), self.symbol_names + other.symbol_names,
        self.value_layers_inputs + other.value_layers_inputs, self.value_layers + other.value_layers,
        self.name + "_" + other.name)

  def __pow__(self, exponent):
    """Inverse circuit sharing this circuit's variables."""
    if exponent != -1:
      raise ValueError("Only the inverse (exponent == -1) is supported.")
    inverse = self.pqc**-1
    return QuantumCircuit(inverse, self.qubits, self.symbol_names, self.value_layers_inputs, self.value_layers,
                          self.name + "_inverse")


class DirectQuantumCircuit(QuantumCircuit):
  """One trainable vector holds the symbol values directly, in lexicographic symbol order
  (reference circuit.py:181-208; default initializer U(0, 2): parameters are exponents)."""

  def __init__(self, pqc, initializer=None, name=None):
    if initializer is None:
      initializer = energy_utils.RandomUniform(0, 2)
    names = sorted(cq.circuit_symbols(pqc))
    values = [torch.nn.Parameter(initializer((len(names),)).to(torch.float32))]
    super().__init__(pqc, pqc.all_qubits(), names, values, [[]], name)


class _Lambda(torch.nn.Module):

  def __init__(self, fn):
    super().__init__()
    self._fn = fn

  def forward(self, inputs):
    return self._fn(inputs)


class QAIA(QuantumCircuit):
  """Quantum adiabatic-inspired ansatz (reference circuit.py:211-292): per layer, the exponential
  of every `quantum_h_terms` entry (symbols gamma_{l}_{k}) followed by the exponential of every
  `classical_h_terms` entry (symbols eta_{l}_{k}).  Variables: etas [L], thetas [C], gammas [L, Q];
  the value map ties eta_l * theta_k across layers exactly as the reference's `embed_params`
  (including its value ordering)."""

  def __init__(self, quantum_h_terms, classical_h_terms, num_layers, initializer=None, name=None):
    import math
    if initializer is None:
      initializer = energy_utils.RandomUniform(0, 2 * math.pi)
    quantum_symbols, classical_symbols = [], []
    for j in range(num_layers):
      quantum_symbols.append([f"gamma_{j}_{k}" for k, _ in enumerate(quantum_h_terms)])
      classical_symbols.append([f"eta_{j}_{k}" for k, _ in enumerate(classical_h_terms)])
    pqc = cq.Circuit()
    flat_symbols = []
    for q_symb, c_symb in zip(quantum_symbols, classical_symbols):
      pqc += exponential(quantum_h_terms, q_symb)
      pqc += exponential(classical_h_terms, c_symb)
      flat_symbols.extend(q_symb + c_symb)
    inputs = [[
        torch.nn.Parameter(initializer((num_layers,)).to(torch.float32)),                       # etas
        torch.nn.Parameter(initializer((len(classical_h_terms),)).to(torch.float32)),           # thetas
        torch.nn.Parameter(initializer((num_layers, len(quantum_h_terms))).to(torch.float32)),  # gammas
    ]]

    def embed_params(x):
      classical = x[0].unsqueeze(1) * x[1].unsqueeze(0)
      return torch.cat([classical, x[2]], 1).reshape(-1)

    qubits = set()
    for term in list(quantum_h_terms) + list(classical_h_terms):
      qubits |= set(term.qubits())
    super().__init__(pqc, sorted(qubits | set(pqc.all_qubits())), flat_symbols, inputs, [[_Lambda(embed_params)]],
                     name)


def exponential(operators, coefficients):
  """Circuit for prod_k exp(-i c_k O_k) (tfq.util.exponential): O_k PauliSums whose strings
  commute, c_k floats or symbol names."""
  out = cq.Circuit()
  for op, c in zip(operators, coefficients):
    if isinstance(op, (cq.PauliString, cq.Operation)):
      op = cq.PauliSum.from_pauli_strings(op)
    out += _exponential(op, cq.Symbol(c) if isinstance(c, str) else c)
  return out


def _exponential(pauli_sum, symbol):
  """exp(-i symbol * P) for every Pauli string P of the sum (they must commute): a string on
  qubits q1..qk conjugated to Z..Z becomes a CNOT ladder + rz; 1- and 2-local strings map to
  native power gates (as tfq.util.exponential does)."""
  out = cq.Circuit()
  import math
  for term in pauli_sum.terms:
    coeff = term.coefficient.real
    items = sorted(term.paulis.items())
    if not items:
      continue
    expo = cq.as_param(symbol) * (2.0 * coeff / math.pi)  # exp(-i c s P) = P**(2 c s/pi), global shift -1/2
    if len(items) == 1:
      q, p = items[0]
      out += cq.Gate({"X": 1, "Y": 2, "Z": 3}[p], (expo,), -0.5).on(q)
    elif len(items) == 2 and items[0][1] == items[1][1]:
      (q0, p), (q1, _) = items
      out += cq.Gate({"X": 9, "Y": 10, "Z": 11}[p], (expo,), -0.5).on(q0, q1)
    else:
      pre = cq.Circuit()
      for q, p in items:
        if p == "X":
          pre += cq.H(q)
        elif p == "Y":
          pre += cq.rx(math.pi / 2).on(q)
      ladder = cq.Circuit(cq.CNOT(a[0], b[0]) for a, b in zip(items[:-1], items[1:]))
      out += pre
      out += ladder
      out += cq.Gate(3, (expo,), -0.5).on(items[-1][0])
      out += ladder**-1
      out += pre**-1
  return out
