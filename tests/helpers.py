"""Shared helpers for the tests: random circuits in gate-table form, Pauli tables."""
import ctypes
import math
import os
import subprocess
import sys

import numpy as np

from oracle import qhbm_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TERM_DTYPE = np.dtype([("coeff", np.float32), ("xmask", np.uint32), ("zmask", np.uint32)])


def gates_array(rows):
  g = np.zeros(len(rows), dtype=orc.GATE_DTYPE)
  for i, r in enumerate(rows):
    g[i] = r
  return g


def random_circuit(n, n_gates, n_symbols, rng, two_qubit=True):
  """Random gate table over every supported gate type, with shared symbols."""
  one_q = [orc.GATE_XPOW, orc.GATE_YPOW, orc.GATE_ZPOW, orc.GATE_HPOW, orc.GATE_PHASEDXPOW,
           orc.GATE_I]
  two_q = [orc.GATE_CZPOW, orc.GATE_CNOTPOW, orc.GATE_SWAPPOW, orc.GATE_ISWAPPOW, orc.GATE_XXPOW,
           orc.GATE_YYPOW, orc.GATE_ZZPOW, orc.GATE_FSIM, orc.GATE_PHASEDISWAPPOW]
  rows = []
  for _ in range(n_gates):
    if two_qubit and n > 1 and rng.random() < 0.45:
      t = int(rng.choice(two_q))
      q0, q1 = (int(x) for x in rng.choice(n, 2, replace=False))
    else:
      t = int(rng.choice(one_q))
      q0, q1 = int(rng.integers(n)), -1
    npar = 0 if t == orc.GATE_I else (2 if t in (orc.GATE_PHASEDXPOW, orc.GATE_FSIM,
                                                 orc.GATE_PHASEDISWAPPOW) else 1)
    sym, scalar, cnst = [-1, -1, -1], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    for k in range(npar):
      if n_symbols > 0 and rng.random() < 0.7:
        sym[k] = int(rng.integers(n_symbols))
        scalar[k] = float(rng.choice([1.0, -1.0, 0.5, 1 / math.pi, 2.0]))
      if rng.random() < 0.5:
        cnst[k] = float(rng.uniform(-1, 1))
    gshift = float(rng.choice([0.0, 0.0, -0.5, 0.25]))
    rows.append(orc._gate(t, q0, q1, sym=tuple(sym), scalar=tuple(scalar), cnst=tuple(cnst),
                          gshift=gshift, nparams=npar))
  return gates_array(rows)


def random_ops(n, n_ops, rng, max_terms=5):
  ops = []
  for _ in range(n_ops):
    terms = []
    for _ in range(int(rng.integers(1, max_terms + 1))):
      k = int(rng.integers(0, min(n, 4) + 1))
      qs = rng.choice(n, k, replace=False) if k else []
      terms.append((float(rng.uniform(-2, 2)), {int(q): str(rng.choice(["X", "Y", "Z"])) for q in qs}))
    ops.append(terms)
  return ops


def ops_to_tables(ops, n):
  """Oracle-style PauliSums -> (terms TERM_DTYPE[], offsets int32[O+1])."""
  terms, offsets = [], [0]
  for op in ops:
    for coeff, paulis in op:
      x = z = 0
      for q, p in paulis.items():
        b = 1 << (n - 1 - q)
        if p in ("X", "Y"):
          x |= b
        if p in ("Z", "Y"):
          z |= b
      terms.append((coeff, x, z))
    offsets.append(len(terms))
  t = np.zeros(len(terms), dtype=TERM_DTYPE)
  for i, r in enumerate(terms):
    t[i] = r
  return t, np.asarray(offsets, dtype=np.int32)


_verify = None


def verify_lib():
  """Builds (g++, host only) and loads the TEST-ONLY schedule interpreter."""
  global _verify
  if _verify is None:
    src = os.path.join(ROOT, "tests", "native", "verify_plan.cpp")
    plan = os.path.join(ROOT, "qhbm-library_b200", "csrc", "plan.cpp")
    out = os.path.join(ROOT, "tests", "native", "libverify_plan.so")
    deps = [src, plan] + [os.path.join(ROOT, "qhbm-library_b200", "csrc", f)
                          for f in ("plan.h", "program.h", "gate_math.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
      subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC",
                             "-I/usr/local/cuda/include", src, plan, "-o", out])
    _verify = ctypes.CDLL(out)
    _verify.verify_last_error.restype = ctypes.c_char_p
  return _verify


def verify_run(gates, n, nsym, ops, symbols, basis, dgrad, with_grad=True, T=0, K=0, mode=0):
  """Runs the compiled program on the host interpreter; returns (E[O], grad[P], state, info)."""
  lib = verify_lib()
  terms, offs = ops_to_tables(ops, n)
  gates = np.ascontiguousarray(gates)
  symbols = np.ascontiguousarray(symbols, dtype=np.float32)
  dgrad = np.ascontiguousarray(dgrad, dtype=np.float32)
  n_eff = max(n, (K or (4 if with_grad else 5)) + 5)
  e = np.zeros(len(ops))
  g = np.zeros(max(nsym, 1))
  st = np.zeros(2 << n_eff)
  info = np.zeros(8, dtype=np.int64)
  rc = lib.verify_run(
      ctypes.c_void_p(gates.ctypes.data), len(gates), n, nsym, ctypes.c_void_p(terms.ctypes.data),
      ctypes.c_void_p(offs.ctypes.data), len(ops), int(with_grad), T, K, mode,
      ctypes.c_void_p(symbols.ctypes.data), ctypes.c_uint64(int(basis)),
      ctypes.c_void_p(dgrad.ctypes.data), ctypes.c_void_p(e.ctypes.data),
      ctypes.c_void_p(g.ctypes.data), ctypes.c_void_p(st.ctypes.data),
      ctypes.c_void_p(info.ctypes.data))
  if rc != 0:
    raise RuntimeError(lib.verify_last_error().decode())
  state = (st[0::2] + 1j * st[1::2])[:1 << n]
  return e, g[:nsym], state, info


def dump_plan(gates, n, nsym, ops, with_grad=True, T=0, K=0):
  """Text dump of the compiled plan (launches, passes, op histograms) from the TEST-ONLY verifier library;
  the C code prints to stdout, which is captured through a temporary file descriptor."""
  import tempfile
  lib = verify_lib()
  terms, offs = ops_to_tables(ops, n)
  gates = np.ascontiguousarray(gates)
  sys.stdout.flush()
  with tempfile.TemporaryFile(mode="w+b") as tmp:
    saved = os.dup(1)
    try:
      os.dup2(tmp.fileno(), 1)
      rc = lib.verify_dump(ctypes.c_void_p(gates.ctypes.data), len(gates), n, nsym, ctypes.c_void_p(terms.ctypes.data),
                           ctypes.c_void_p(offs.ctypes.data), len(ops), int(with_grad), T, K)
      libc = ctypes.CDLL(None)
      libc.fflush(None)
    finally:
      os.dup2(saved, 1)
      os.close(saved)
    tmp.seek(0)
    text = tmp.read().decode()
  if rc != 0:
    raise RuntimeError(lib.verify_last_error().decode())
  return text
