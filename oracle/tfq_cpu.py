"""ctypes wrapper + problem builder for oracle/tfq_cpu.c (CPU restatement of TFQ 0.6.1's
expectation / adjoint algorithm).  TEST / BASELINE INFRASTRUCTURE ONLY -- see the C header.

The shared object is compiled on first use for the machine it runs on (gcc -O3
-march=native -fopenmp) into oracle/_build/ (git-ignored)."""
import ctypes
import hashlib
import os
import platform
import subprocess

import numpy as np

from oracle import qhbm_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class _Problem(ctypes.Structure):
  _fields_ = [("n", ctypes.c_int), ("n_blocks", ctypes.c_int), ("n_gates", ctypes.c_int),
              ("n_ops", ctypes.c_int), ("n_sym", ctypes.c_int)] + [
                  (k, ctypes.c_void_p) for k in
                  ("bq0", "bq1", "bmat", "gq0", "gq1", "gdag", "grad_off", "grad_sym", "grad_mat",
                   "t_coeff", "t_x", "t_z", "t_off")]


def _cpu_tag():
  try:
    with open("/proc/cpuinfo") as f:
      flags = [l for l in f if l.startswith("flags")][0]
  except Exception:
    flags = platform.processor()
  return hashlib.sha1(flags.encode()).hexdigest()[:10]


def lib():
  global _lib
  if _lib is None:
    src = os.path.join(_HERE, "tfq_cpu.c")
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libtfq_cpu_{_cpu_tag()}.so")
    if not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
      subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-fno-math-errno", "-fno-trapping-math",
                             "-ffp-contract=fast", "-shared", "-fPIC", "-std=c11", src, "-o", out, "-lm"])
    _lib = ctypes.CDLL(out)
    _lib.tfq_cpu_max_threads.restype = ctypes.c_int
  return _lib


def max_threads():
  return int(lib().tfq_cpu_max_threads())


def _pad16(m):
  out = np.zeros(16, dtype=np.complex64)
  out[:m.size] = m.reshape(-1)
  return out


def fuse_blocks(gates, symbol_values):
  """<=2-qubit gate fusion in the spirit of qsim's BasicGateFuser: one-qubit gates are absorbed
  into the next (or, at the end, the previous) two-qubit block on their qubit; consecutive
  two-qubit gates on the same pair merge.  Returns [(q0, q1, matrix)] with q1 = -1 for 1q."""
  blocks = []          # [q0, q1, matrix(complex128)]
  pending = {}         # qubit -> 2x2 waiting for a block
  last_block = {}      # qubit -> index of the last block touching it
  for g in gates:
    t = int(g["type"])
    if t == orc.GATE_I:
      continue
    m = orc.gate_matrix(t, orc.gate_params(g, symbol_values), float(g["gshift"]))
    qs = orc.gate_qubits(g)
    if len(qs) == 1:
      q = qs[0]
      pending[q] = m @ pending.get(q, np.eye(2, dtype=np.complex128))
      continue
    q0, q1 = qs
    full = m @ np.kron(pending.pop(q0, np.eye(2)), pending.pop(q1, np.eye(2)))
    b0, b1 = last_block.get(q0), last_block.get(q1)
    if b0 is not None and b0 == b1 and blocks[b0][1] >= 0:
      bq0, bq1, bm = blocks[b0]
      if (bq0, bq1) == (q0, q1):
        blocks[b0][2] = full @ bm
        continue
      if (bq0, bq1) == (q1, q0):
        perm = [0, 2, 1, 3]
        blocks[b0][2] = full[np.ix_(perm, perm)] @ bm
        continue
    blocks.append([q0, q1, full])
    last_block[q0] = last_block[q1] = len(blocks) - 1
  for q, m in pending.items():
    b = last_block.get(q)
    if b is not None and blocks[b][1] >= 0:
      bq0, bq1, bm = blocks[b]
      ext = np.kron(m, np.eye(2)) if q == bq0 else np.kron(np.eye(2), m)
      blocks[b][2] = ext @ bm
    else:
      blocks.append([q, -1, m])
      last_block[q] = len(blocks) - 1
  return [(b[0], b[1], b[2]) for b in blocks]


class Problem:
  """Everything the C code needs for one (circuit, symbol values, observables) triple."""

  def __init__(self, gates, n, symbol_values, ops, grad_mode="tfq_fd"):
    symbol_values = np.asarray(symbol_values, dtype=np.float64)
    blocks = fuse_blocks(gates, symbol_values)
    self.n, self.n_sym, self.n_ops = n, len(symbol_values), len(ops)
    self.n_blocks = len(blocks)
    self.bq0 = np.array([b[0] for b in blocks] or [0], dtype=np.int32)
    self.bq1 = np.array([b[1] for b in blocks] or [0], dtype=np.int32)
    self.bmat = np.concatenate([_pad16(b[2]) for b in blocks] or [np.zeros(16, np.complex64)])
    live = [g for g in gates if int(g["type"]) != orc.GATE_I]
    self.n_gates = len(live)
    gq0, gq1, gdag, goff, gsym, gmat = [], [], [], [0], [], []
    for g in live:
      qs = orc.gate_qubits(g)
      m = orc.gate_matrix(int(g["type"]), orc.gate_params(g, symbol_values), float(g["gshift"]))
      gq0.append(qs[0])
      gq1.append(qs[1] if len(qs) == 2 else -1)
      gdag.append(_pad16(m.conj().T))
      for k in range(int(g["nparams"])):
        if int(g["sym"][k]) >= 0:
          gsym.append(int(g["sym"][k]))
          gmat.append(_pad16(orc.gate_derivative(g, symbol_values, k, grad_mode, fd_float32=(grad_mode == "tfq_fd"))))
      goff.append(len(gsym))
    self.gq0 = np.array(gq0 or [0], dtype=np.int32)
    self.gq1 = np.array(gq1 or [0], dtype=np.int32)
    self.gdag = np.concatenate(gdag or [np.zeros(16, np.complex64)])
    self.grad_off = np.array(goff, dtype=np.int32)
    self.grad_sym = np.array(gsym or [0], dtype=np.int32)
    self.grad_mat = np.concatenate(gmat or [np.zeros(16, np.complex64)])
    tc, tx, tz, toff = [], [], [], [0]
    for op in ops:
      for coeff, paulis in op:
        x = z = 0
        for q, p in paulis.items():
          b = 1 << (n - 1 - q)
          if p in ("X", "Y"):
            x |= b
          if p in ("Z", "Y"):
            z |= b
        tc.append(coeff)
        tx.append(x)
        tz.append(z)
      toff.append(len(tc))
    self.t_coeff = np.array(tc or [0], dtype=np.float32)
    self.t_x = np.array(tx or [0], dtype=np.uint32)
    self.t_z = np.array(tz or [0], dtype=np.uint32)
    self.t_off = np.array(toff, dtype=np.int32)
    self._c = _Problem(n, self.n_blocks, self.n_gates, self.n_ops, self.n_sym,
                       *[getattr(self, k).ctypes.data for k in
                         ("bq0", "bq1", "bmat", "gq0", "gq1", "gdag", "grad_off", "grad_sym", "grad_mat",
                          "t_coeff", "t_x", "t_z", "t_off")])

  def expectation(self, basis, threads=0):
    basis = np.ascontiguousarray(basis, dtype=np.uint64)
    out = np.zeros((len(basis), self.n_ops), dtype=np.float32)
    rc = lib().tfq_cpu_expectation(ctypes.byref(self._c), ctypes.c_void_p(basis.ctypes.data),
                                   ctypes.c_int64(len(basis)), ctypes.c_void_p(out.ctypes.data), int(threads))
    assert rc == 0
    return out

  def adjoint(self, basis, dgrad, threads=0):
    basis = np.ascontiguousarray(basis, dtype=np.uint64)
    dgrad = np.ascontiguousarray(dgrad, dtype=np.float32)
    out = np.zeros((len(basis), self.n_ops), dtype=np.float32)
    grad = np.zeros((len(basis), max(self.n_sym, 1)), dtype=np.float32)
    rc = lib().tfq_cpu_adjoint(ctypes.byref(self._c), ctypes.c_void_p(basis.ctypes.data),
                               ctypes.c_int64(len(basis)), ctypes.c_void_p(dgrad.ctypes.data),
                               ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(grad.ctypes.data), int(threads))
    assert rc == 0
    return out, grad[:, :self.n_sym]
