"""CPU oracle for the QHBM expectation / gradient hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy (complex128) restatement of the algorithm the reference
runs for the path named in BASELINE.json.  Nothing in the product package may
import it: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` do, and only as the checker.

Where the algorithm lives.  qhbmlib itself (Python, `/root/reference/qhbmlib`)
contains the orchestration; the arithmetic is inside third-party wheels that are
NOT vendored under `/root/reference` and cannot be installed here:
  * tensorflow-quantum == 0.6.1 (pyproject.toml:32)  -- qsim state-vector
    simulation, `TfqSimulateExpectation`, `TfqAdjointGradient`
  * cirq-core == 0.14.1 (poetry.lock:122-123)         -- gate matrices
  * tensorflow-probability == 0.15.0 (pyproject.toml:31) -- samplers
Their published algorithms are restated below; every function cites the
reference call site (file:line under /root/reference) it stands in for.

Pinning status.  The oracle is pinned (tests/test_oracle_golden.py) against every
closed-form known answer the reference's own tests hold for this path
(SURVEY.md section 8c, G1-G13), to 1e-9 or better instead of the reference's
2e-3..3e-2.  It is NOT pinned against outputs of TFQ itself (not installable):
TFQ-specific behaviours that the reference's tests never exercise -- the
finite-difference gate derivative (`grad_mode="tfq_fd"`) and the n>=11
bit-column permutation (`ref_bit_order`) -- are restated from the published
TFQ/qhbmlib sources and remain "parity unpinned" at the 1e-5 level.

Conventions: qubit k of the sorted qubit list is bit
(n-1-k) of the basis index (big-endian, cirq/TFQ); spin of bit b is 1-2b.
"""

import itertools
import math

import numpy as np

# --------------------------------------------------------------------------
# Gate table.  Same numeric layout as include/qhbm_b200.h (qhbm_gate_t), kept
# here as an independent definition so the oracle does not import the product.
# --------------------------------------------------------------------------
GATE_I, GATE_XPOW, GATE_YPOW, GATE_ZPOW, GATE_HPOW = 0, 1, 2, 3, 4
GATE_CZPOW, GATE_CNOTPOW, GATE_SWAPPOW, GATE_ISWAPPOW = 5, 6, 7, 8
GATE_XXPOW, GATE_YYPOW, GATE_ZZPOW = 9, 10, 11
GATE_PHASEDXPOW, GATE_FSIM, GATE_PHASEDISWAPPOW = 12, 13, 14

GATE_DTYPE = np.dtype([
    ("type", np.int32), ("q0", np.int32), ("q1", np.int32), ("nparams", np.int32),
    ("sym", np.int32, (3,)), ("scalar", np.float32, (3,)), ("cnst", np.float32, (3,)),
    ("gshift", np.float32),
])

TWO_QUBIT = {GATE_CZPOW, GATE_CNOTPOW, GATE_SWAPPOW, GATE_ISWAPPOW, GATE_XXPOW,
             GATE_YYPOW, GATE_ZZPOW, GATE_FSIM, GATE_PHASEDISWAPPOW}

_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2.0)
_I2 = np.eye(2, dtype=np.complex128)
_I4 = np.eye(4, dtype=np.complex128)
_P1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]],
                 dtype=np.complex128)
PAULI = {"I": _I2, "X": _X, "Y": _Y, "Z": _Z}


def _two_level(p1, t, g):
  """cirq EigenGate with eigen-exponents {0,1}: e^{i pi t g}(I + (e^{i pi t}-1) P1)."""
  d = p1.shape[0]
  return np.exp(1j * math.pi * t * g) * (
      np.eye(d, dtype=np.complex128) + (np.exp(1j * math.pi * t) - 1.0) * p1)


def gate_matrix(gtype, p, g=0.0):
  """Unitary of one gate (cirq 0.14.1 definitions).

  p: parameter values (exponent, phase_exponent / theta, phi), g: global shift.
  Two-qubit matrices are big-endian in (q0, q1): row index = 2*b(q0) + b(q1).
  """
  if gtype == GATE_I:
    return _I2.copy()
  if gtype == GATE_XPOW:
    return _two_level((_I2 - _X) / 2, p[0], g)
  if gtype == GATE_YPOW:
    return _two_level((_I2 - _Y) / 2, p[0], g)
  if gtype == GATE_ZPOW:
    return _two_level(_P1, p[0], g)
  if gtype == GATE_HPOW:
    return _two_level((_I2 - _H) / 2, p[0], g)
  if gtype == GATE_CZPOW:
    return _two_level(np.kron(_P1, _P1), p[0], g)
  if gtype == GATE_CNOTPOW:
    return _two_level(np.kron(_P1, (_I2 - _X) / 2), p[0], g)
  if gtype == GATE_SWAPPOW:
    return _two_level((_I4 - _SWAP) / 2, p[0], g)
  if gtype == GATE_XXPOW:
    return _two_level((_I4 - np.kron(_X, _X)) / 2, p[0], g)
  if gtype == GATE_YYPOW:
    return _two_level((_I4 - np.kron(_Y, _Y)) / 2, p[0], g)
  if gtype == GATE_ZZPOW:
    return _two_level((_I4 - np.kron(_Z, _Z)) / 2, p[0], g)
  if gtype == GATE_ISWAPPOW:
    c, s = math.cos(math.pi * p[0] / 2), math.sin(math.pi * p[0] / 2)
    m = np.array([[1, 0, 0, 0], [0, c, 1j * s, 0], [0, 1j * s, c, 0],
                  [0, 0, 0, 1]], dtype=np.complex128)
    return np.exp(1j * math.pi * p[0] * g) * m
  if gtype == GATE_PHASEDXPOW:
    t, ph = p[0], p[1]
    zp = np.diag([1.0, np.exp(1j * math.pi * ph)])
    return zp @ _two_level((_I2 - _X) / 2, t, g) @ zp.conj().T
  if gtype == GATE_FSIM:
    th, phi = p[0], p[1]
    c, s = math.cos(th), math.sin(th)
    return np.array([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0],
                     [0, 0, 0, np.exp(-1j * phi)]], dtype=np.complex128)
  if gtype == GATE_PHASEDISWAPPOW:
    t, ph = p[0], p[1]
    c, s = math.cos(math.pi * t / 2), math.sin(math.pi * t / 2)
    f = np.exp(2j * math.pi * ph)
    return np.array([[1, 0, 0, 0], [0, c, 1j * s * f, 0],
                     [0, 1j * s * np.conj(f), c, 0], [0, 0, 0, 1]],
                    dtype=np.complex128)
  raise ValueError(f"unknown gate type {gtype}")


def gate_params(gate, symbol_values):
  """Resolved parameter values: cnst + scalar * symbol (TFQ: exponent_scalar)."""
  out = []
  for k in range(int(gate["nparams"])):
    v = float(gate["cnst"][k])
    s = int(gate["sym"][k])
    if s >= 0:
      v += float(gate["scalar"][k]) * float(symbol_values[s])
    out.append(v)
  return out


def gate_qubits(gate):
  if int(gate["type"]) in TWO_QUBIT:
    return (int(gate["q0"]), int(gate["q1"]))
  return (int(gate["q0"]),)


# --------------------------------------------------------------------------
# State-vector simulation (stands in for qsim inside TfqSimulateExpectation,
# reached from qhbmlib/inference/qnn.py:134-138).
# --------------------------------------------------------------------------
def apply_matrix(state, n, qubits, m):
  """Applies a 2^k x 2^k matrix to `qubits` (big-endian) of an n-qubit state."""
  k = len(qubits)
  psi = state.reshape([2] * n)
  axes = list(qubits)
  m = m.reshape([2] * (2 * k))
  psi = np.tensordot(m, psi, axes=(list(range(k, 2 * k)), axes))
  psi = np.moveaxis(psi, list(range(k)), axes)
  return psi.reshape(-1)


def basis_state(n, index):
  s = np.zeros(1 << n, dtype=np.complex128)
  s[int(index)] = 1.0
  return s


def simulate(gates, n, symbol_values, index):
  """U(phi)|index>.  The X**bit injector circuit of qhbmlib/models/circuit.py:129-136
  (+ circuit_utils.py:23-29) applied to |0..0> is exactly the basis state."""
  state = basis_state(n, index)
  for gate in gates:
    m = gate_matrix(int(gate["type"]), gate_params(gate, symbol_values),
                    float(gate["gshift"]))
    state = apply_matrix(state, n, gate_qubits(gate), m)
  return state


# --------------------------------------------------------------------------
# Bitstrings -> basis index, including the reference's column permutation.
# --------------------------------------------------------------------------
def bit_column_to_qubit(n, ref_bit_order=True):
  """Which sorted-qubit each bitstring column drives.

  qhbmlib/models/circuit.py:59-63 sorts the injector symbol names
  "bit_circuit_bit_{k}" lexicographically and circuit.py:132-134 binds column j
  to the j-th sorted name, so for n >= 11 column j drives qubit pi(j) with pi the
  string-sorted order of 0..n-1.
  """
  if not ref_bit_order:
    return list(range(n))
  return sorted(range(n), key=lambda k: f"bit_circuit_bit_{k}")


def bitstrings_to_index(bitstrings, ref_bit_order=True):
  b = np.asarray(bitstrings).astype(np.int64)
  n = b.shape[1]
  pi = bit_column_to_qubit(n, ref_bit_order)
  idx = np.zeros(b.shape[0], dtype=np.int64)
  for j in range(n):
    idx |= b[:, j] << (n - 1 - pi[j])
  return idx


# --------------------------------------------------------------------------
# PauliSum expectation (TfqSimulateExpectation).
# A PauliSum is a list of terms (coeff, {qubit: "X"|"Y"|"Z"}).
# --------------------------------------------------------------------------
def apply_pauli_sum(state, n, terms):
  out = np.zeros_like(state)
  for coeff, paulis in terms:
    t = state
    for q, p in paulis.items():
      t = apply_matrix(t, n, (q,), PAULI[p])
    out = out + float(np.real(coeff)) * t
  return out


def expectation(state, n, terms):
  return float(np.real(np.vdot(state, apply_pauli_sum(state, n, terms))))


def expectations(gates, n, symbol_values, indices, ops):
  """f32[U,O] of TfqSimulateExpectation: <index_u|U^dag H_j U|index_u>."""
  out = np.zeros((len(indices), len(ops)))
  for u, idx in enumerate(indices):
    psi = simulate(gates, n, symbol_values, idx)
    for j, terms in enumerate(ops):
      out[u, j] = expectation(psi, n, terms)
  return out


# --------------------------------------------------------------------------
# Adjoint gradient (TfqAdjointGradient, the default differentiator behind
# tfq.layers.Expectation() at qnn.py:112).
# --------------------------------------------------------------------------
TFQ_GRAD_EPS = 5e-3


def _shifted_params(gate, symbol_values, k, delta):
  p = gate_params(gate, symbol_values)
  p[k] += float(gate["scalar"][k]) * delta
  return p


def gate_derivative(gate, symbol_values, k, grad_mode, fd_float32=False):
  """d(gate matrix)/d(symbol of parameter k).

  "exact": analytic limit, evaluated by a 4th-order central stencil in float64
    with h = 1e-3 (truncation ~1e-11 relative, far below every tolerance used).
  "tfq_fd": TFQ 0.6.1 adj_util.cc -- central difference of the gate matrix with
    eps = 5e-3 on the symbol value; `fd_float32` additionally rounds the two
    matrices to complex64 first, as the C++ does.
  """
  gt, g = int(gate["type"]), float(gate["gshift"])
  if grad_mode == "tfq_fd":
    a = gate_matrix(gt, _shifted_params(gate, symbol_values, k, +TFQ_GRAD_EPS), g)
    b = gate_matrix(gt, _shifted_params(gate, symbol_values, k, -TFQ_GRAD_EPS), g)
    if fd_float32:
      a = a.astype(np.complex64)
      b = b.astype(np.complex64)
      return ((a - b) * np.float32(0.5 / TFQ_GRAD_EPS)).astype(np.complex128)
    return (a - b) / (2 * TFQ_GRAD_EPS)
  if grad_mode == "exact":
    h = 1e-3
    f = lambda d: gate_matrix(gt, _shifted_params(gate, symbol_values, k, d), g)
    return (-f(2 * h) + 8 * f(h) - 8 * f(-h) + f(-2 * h)) / (12 * h)
  raise ValueError(grad_mode)


def adjoint_gradient(gates, n, symbol_values, index, ops, dgrad,
                     grad_mode="exact", fd_float32=False):
  """Returns (expectations f64[O], grad f64[P]) for one basis state.

  grad[s] = sum_j dgrad[j] d<H_j>/d symbol_s, accumulated over every gate
  parameter bound to symbol s, computed the way TfqAdjointGradient does:
  lambda = sum_j dgrad_j H_j psi; walk the gates in reverse un-applying each on
  psi and lambda, adding 2 Re <lambda_k| dG_k |psi_{k-1}> at parameterised gates.
  """
  nsym = len(symbol_values)
  psi = simulate(gates, n, symbol_values, index)
  exps = np.array([expectation(psi, n, t) for t in ops])
  lam = np.zeros_like(psi)
  for j, terms in enumerate(ops):
    if dgrad[j] != 0.0:
      lam = lam + float(dgrad[j]) * apply_pauli_sum(psi, n, terms)
  grad = np.zeros(nsym)
  for gate in reversed(list(gates)):
    qs = gate_qubits(gate)
    m = gate_matrix(int(gate["type"]), gate_params(gate, symbol_values),
                    float(gate["gshift"]))
    psi = apply_matrix(psi, n, qs, m.conj().T)
    for k in range(int(gate["nparams"])):
      s = int(gate["sym"][k])
      if s < 0:
        continue
      dm = gate_derivative(gate, symbol_values, k, grad_mode, fd_float32)
      grad[s] += 2.0 * float(np.real(np.vdot(lam, apply_matrix(psi, n, qs, dm))))
    lam = apply_matrix(lam, n, qs, m.conj().T)
  return exps, grad


def batch_expectation_and_gradient(gates, n, symbol_values, indices, ops, dgrads,
                                   grad_mode="exact", fd_float32=False):
  """f64[U,O], f64[U,P] -- the two TFQ op outputs for a batch of basis states."""
  nsym = len(symbol_values)
  e = np.zeros((len(indices), len(ops)))
  g = np.zeros((len(indices), nsym))
  for u, idx in enumerate(indices):
    e[u], g[u] = adjoint_gradient(gates, n, symbol_values, idx, ops, dgrads[u],
                                  grad_mode, fd_float32)
  return e, g


def numeric_gradient(gates, n, symbol_values, index, ops, dgrad, h=1e-3):
  """4th-order stencil on the expectation itself (tests/test_util.py:210-309 uses a
  five-point stencil with delta=0.1 the same way)."""
  symbol_values = np.asarray(symbol_values, dtype=np.float64)
  grad = np.zeros(len(symbol_values))

  def f(vals):
    psi = simulate(gates, n, vals, index)
    return sum(float(dgrad[j]) * expectation(psi, n, t) for j, t in enumerate(ops))

  for s in range(len(symbol_values)):
    e = np.zeros_like(symbol_values)
    e[s] = h
    grad[s] = (-f(symbol_values + 2 * e) + 8 * f(symbol_values + e) -
               8 * f(symbol_values - e) + f(symbol_values - 2 * e)) / (12 * h)
  return grad


# --------------------------------------------------------------------------
# Circuit / Hamiltonian builders (specs: tests/test_util.py:25-67 == baselines/pqc.py:21-63;
# baselines/train.py:46-58).
# --------------------------------------------------------------------------
def _gate(gtype, q0, q1=-1, sym=(-1, -1, -1), scalar=(0, 0, 0), cnst=(0, 0, 0),
          gshift=0.0, nparams=1):
  g = np.zeros((), dtype=GATE_DTYPE)
  g["type"], g["q0"], g["q1"], g["nparams"] = gtype, q0, q1, nparams
  g["sym"], g["scalar"], g["cnst"], g["gshift"] = sym, scalar, cnst, gshift
  return g


def hea_circuit(n, num_layers, name="q"):
  """Hardware-efficient ansatz of tests/test_util.py:25-67.

  Returns (gates, symbol_names) with symbol_names sorted lexicographically the way
  DirectQuantumCircuit does (models/circuit.py:201-203); gate symbol indices
  point into that sorted list.
  """
  raw = []  # (type, q0, q1, symbol name)
  for layer in range(num_layers):
    for q in range(n):
      raw.append((GATE_XPOW, q, -1, f"sx_{name}_{layer}_{q}"))
      raw.append((GATE_ZPOW, q, -1, f"sz_{name}_{layer}_{q}"))
    if n > 1:
      for k, q0 in enumerate(range(0, n - 1, 2)):
        raw.append((GATE_CZPOW, q0, q0 + 1, f"sc_{name}_{layer}_{2 * k}"))
      for k, q0 in enumerate(range(1, n - 1, 2)):
        raw.append((GATE_CZPOW, q0, q0 + 1, f"sc_{name}_{layer}_{2 * k + 1}"))
  names = sorted({r[3] for r in raw})
  pos = {s: i for i, s in enumerate(names)}
  gates = np.zeros(len(raw), dtype=GATE_DTYPE)
  for i, (t, q0, q1, s) in enumerate(raw):
    gates[i] = _gate(t, q0, q1, sym=(pos[s], -1, -1), scalar=(1, 0, 0))
  return gates, names


def inverse_circuit(gates):
  """cirq `circuit**-1` as used by models/circuit.py:171-176: reversed order,
  exponents negated, same symbols.  (Only eigen-gates; FSim/phased gates negate
  their first parameter too -- theta, resp. exponent -- and FSim also phi.)"""
  out = gates[::-1].copy()
  for g in out:
    t = int(g["type"])
    g["scalar"][0] = -g["scalar"][0]
    g["cnst"][0] = -g["cnst"][0]
    if t == GATE_FSIM:
      g["scalar"][1] = -g["scalar"][1]
      g["cnst"][1] = -g["cnst"][1]
  return out


def concat_circuits(gates_a, nsym_a, gates_b):
  """`QuantumCircuit.__add__` (models/circuit.py:138-162): symbols of b follow a's."""
  b = gates_b.copy()
  for g in b:
    for k in range(3):
      if g["sym"][k] >= 0:
        g["sym"][k] += nsym_a
  return np.concatenate([gates_a, b])


def tfim_ring(n, bias=1.0):
  """H = -sum Z_i Z_{i+1 mod n} - bias sum X_i (baselines/train.py:46-58, 1D)."""
  terms = []
  for i in range(n):
    terms.append((-bias, {i: "X"}))
  for i in range(n):
    j = (i + 1) % n
    if i == j:
      terms.append((-1.0, {}))
    else:
      terms.append((-1.0, {i: "Z", j: "Z"}))
  return terms


def xxz_ring(n, delta=0.5):
  """H = sum_i X_iX_{i+1} + Y_iY_{i+1} + delta Z_iZ_{i+1} (synthetic)."""
  terms = []
  for i in range(n):
    j = (i + 1) % n
    terms.append((1.0, {i: "X", j: "X"}))
    terms.append((1.0, {i: "Y", j: "Y"}))
    terms.append((delta, {i: "Z", j: "Z"}))
  return terms


# --------------------------------------------------------------------------
# Energy functions (qhbmlib/models/energy.py, energy_utils.py).
# --------------------------------------------------------------------------
def spins_from_bitstrings(bitstrings):
  """energy_utils.py:46-52: 0 -> +1, 1 -> -1."""
  return (1 - 2 * np.asarray(bitstrings)).astype(np.float64)


def parity_indices(num_bits, order):
  """energy_utils.py:97-100: itertools.combinations order, i = 1..order."""
  out = []
  for i in range(1, order + 1):
    out.extend(itertools.combinations(range(num_bits), i))
  return [tuple(c) for c in out]


def parity_features(bitstrings, order):
  """energy_utils.py:104-110: products of spins over each index group."""
  s = spins_from_bitstrings(bitstrings)
  idx = parity_indices(s.shape[1], order)
  return np.stack([np.prod(s[:, list(c)], axis=1) for c in idx], axis=1)


def bernoulli_energy(bitstrings, thetas):
  """energy.py:123-167: E(b) = sum_i (1-2 b_i) theta_i."""
  return spins_from_bitstrings(bitstrings) @ np.asarray(thetas, dtype=np.float64)


def kobe_energy(bitstrings, order, thetas):
  """energy.py:170-209: E(b) = parity features . theta."""
  return parity_features(bitstrings, order) @ np.asarray(thetas, dtype=np.float64)


def mlp_energy(bitstrings, layers):
  """Generic BitstringEnergy stack on raw bits (energy.py:82-87) for the Dense/tanh
  family of tests/inference/ebm_utils_test.py:33-47.  layers: [(W[in,out], b[out], act)]."""
  x = np.asarray(bitstrings).astype(np.float64)
  for w, b, act in layers:
    x = x @ np.asarray(w, dtype=np.float64) + np.asarray(b, dtype=np.float64)
    if act == "tanh":
      x = np.tanh(x)
    elif act == "relu":
      x = np.maximum(x, 0.0)
    elif act not in ("linear", None):
      raise ValueError(act)
  return x.reshape(x.shape[0]) if x.ndim == 2 and x.shape[1] == 1 else x


def kobe_shards(num_bits, order):
  """energy.py:200-209: one Z-string PauliSum per parity group."""
  return [[(1.0, {q: "Z" for q in c})] for c in parity_indices(num_bits, order)]


def bernoulli_shards(num_bits):
  """energy.py:165-167."""
  return [[(1.0, {q: "Z"})] for q in range(num_bits)]


# --------------------------------------------------------------------------
# utils.py
# --------------------------------------------------------------------------
def unique_bitstrings_with_counts(bitstrings):
  """utils.py:61-78 (UniqueWithCountsV2, axis 0): first-occurrence order."""
  b = np.asarray(bitstrings)
  seen = {}
  y, idx, count = [], np.zeros(b.shape[0], dtype=np.int32), []
  for i, row in enumerate(map(bytes, np.ascontiguousarray(b))):
    j = seen.get(row)
    if j is None:
      j = len(y)
      seen[row] = j
      y.append(b[i])
      count.append(0)
    idx[i] = j
    count[j] += 1
  y = np.stack(y) if y else b[:0]
  return y, idx, np.asarray(count, dtype=np.int32)


def weighted_average(counts, values):
  """utils.py:43-58 (float32 counts in the reference; float64 here)."""
  c = np.asarray(counts, dtype=np.float64)
  v = np.asarray(values, dtype=np.float64)
  return np.tensordot(c, v, axes=(0, 0)) / c.sum()


def expand_unique_results(y, idx):
  """utils.py:81-92."""
  return np.asarray(y)[np.asarray(idx)]


# --------------------------------------------------------------------------
# EBM inference (qhbmlib/inference/ebm.py).
# --------------------------------------------------------------------------
def all_bitstrings(n):
  """ebm.py:445-447: itertools.product order == big-endian counting."""
  r = np.arange(1 << n, dtype=np.int64)
  return ((r[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1).astype(np.int8)


def logsumexp(x):
  x = np.asarray(x, dtype=np.float64)
  m = x.max()
  return float(m + np.log(np.exp(x - m).sum()))


def analytic_log_partition(energies):
  """ebm.py:482-485."""
  return logsumexp(-np.asarray(energies, dtype=np.float64))


def analytic_entropy(energies):
  """ebm.py:478-480: entropy of Categorical(logits=-E)."""
  logits = -np.asarray(energies, dtype=np.float64)
  logp = logits - logsumexp(logits)
  return float(-(np.exp(logp) * logp).sum())


def analytic_probabilities(energies):
  logits = -np.asarray(energies, dtype=np.float64)
  return np.exp(logits - logsumexp(logits))


def bernoulli_log_partition(thetas):
  """ebm.py:546-557."""
  t = np.asarray(thetas, dtype=np.float64)
  return float(np.log(np.exp(t) + np.exp(-t)).sum())


def bernoulli_entropy(thetas):
  """ebm.py:537-544: sum of entropies of Bernoulli(logits=2 theta)."""
  t = np.asarray(thetas, dtype=np.float64)
  p = 1.0 / (1.0 + np.exp(-2.0 * t))
  h = -(p * np.log(p) + (1 - p) * np.log1p(-p))
  return float(h.sum())


def expectation_score_gradient(counts, values, energy_jacobian, function_grad,
                               upstream):
  """grad_fn of EnergyInference._expectation (ebm.py:282-325) for one flat value tensor.

  counts int[U]; values f[U,...]; upstream f[...]; energy_jacobian f[U,T] =
  dE(x_u)/dtheta; function_grad f[T] = gradient of the weighted average through
  `function` itself.  Returns E[c]E[dE] - E[c dE] + function_grad, c_u = sum(upstream*values_u).
  """
  v = np.asarray(values, dtype=np.float64)
  c = (v * np.asarray(upstream, dtype=np.float64)).reshape(v.shape[0], -1).sum(1)
  jac = np.asarray(energy_jacobian, dtype=np.float64)
  avg_c = weighted_average(counts, c)
  avg_j = weighted_average(counts, jac)
  avg_cj = weighted_average(counts, jac * c[:, None])
  return avg_c * avg_j - avg_cj + np.asarray(function_grad, dtype=np.float64)


def log_partition_gradient(counts, energy_jacobian, upstream=1.0):
  """ebm.py:396-415: -upstream * E_{x~p}[dE/dtheta] from (unique) samples."""
  return -upstream * weighted_average(counts, energy_jacobian)


# --------------------------------------------------------------------------
# Compositions: QHBM.expectation (qhbm.py:124-147), vqt (vqt_loss.py:25-55),
# qmhl (qmhl_loss.py:21-34).  They take the *sampled* unique bitstrings and
# counts as inputs (the sampler's random stream is not reproducible; App. A.8).
# --------------------------------------------------------------------------
def qhbm_expectation(gates, n, symbol_values, unique_bitstrings, counts, ops,
                     ref_bit_order=True):
  idx = bitstrings_to_index(unique_bitstrings, ref_bit_order)
  vals = expectations(gates, n, symbol_values, idx, ops)
  return weighted_average(counts, vals), vals


def modular_hamiltonian_expectation(gates_total, n, symbol_values, unique_bitstrings,
                                    counts, shards, thetas, ref_bit_order=True):
  """qnn.py:120-139 with a Hamiltonian observable: sum_k theta_k <Z-string_k>."""
  avg, vals = qhbm_expectation(gates_total, n, symbol_values, unique_bitstrings,
                               counts, shards, ref_bit_order)
  th = np.asarray(thetas, dtype=np.float64)
  return float(avg @ th), vals @ th
