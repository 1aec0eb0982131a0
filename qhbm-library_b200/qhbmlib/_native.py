"""ctypes binding of libqhbm_b200.so (C ABI: include/qhbm_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a) and
sits next to this package.  There is no CPU fallback: if the library is missing or no
CUDA device is present, the calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# QHBM_B200_LIB: development switch (A/B timing of two builds of the same C ABI); default = the in-tree build
LIB_PATH = os.environ.get("QHBM_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libqhbm_b200.so")

GATE_DTYPE = np.dtype([
    ("type", np.int32), ("q0", np.int32), ("q1", np.int32), ("nparams", np.int32),
    ("sym", np.int32, (3,)), ("scalar", np.float32, (3,)), ("cnst", np.float32, (3,)),
    ("gshift", np.float32),
])
TERM_DTYPE = np.dtype([("coeff", np.float32), ("xmask", np.uint32), ("zmask", np.uint32)])

GRAD_EXACT, GRAD_TFQ_FD, GRAD_TFQ_FD_F32 = 0, 1, 2
ENERGY_BERNOULLI, ENERGY_KOBE, ENERGY_MLP = 0, 1, 2
COMM_ID_BYTES = 128  # QHBM_COMM_ID_BYTES


class EnergyDesc(ctypes.Structure):
  _fields_ = [
      ("kind", ctypes.c_int32), ("n_bits", ctypes.c_int32), ("n_terms", ctypes.c_int32),
      ("d_masks", ctypes.c_void_p), ("d_theta", ctypes.c_void_p),
      ("n_layers", ctypes.c_int32), ("widths", ctypes.c_int32 * 9), ("act", ctypes.c_int32 * 8),
      ("d_weights", ctypes.c_void_p * 8), ("d_bias", ctypes.c_void_p * 8),
  ]


class NativeError(RuntimeError):
  pass


_lib = None

_VP = ctypes.c_void_p
_I32, _I64, _U64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64

# name -> (restype, argtypes); every symbol declared in include/qhbm_b200.h
SIGNATURES = {
    "qhbm_last_error": (ctypes.c_char_p, []),
    "qhbm_version": (ctypes.c_int, []),
    "qhbm_circuit_create": (ctypes.c_int, [_VP, _I32, _I32, _I32, ctypes.POINTER(_VP)]),
    "qhbm_circuit_destroy": (None, [_VP]),
    "qhbm_ops_create": (ctypes.c_int, [_VP, _VP, _I32, _I32, ctypes.POINTER(_VP)]),
    "qhbm_ops_destroy": (None, [_VP]),
    "qhbm_plan_create": (ctypes.c_int, [_VP, _VP, _I32, _I32, _I32, ctypes.POINTER(_VP)]),
    "qhbm_plan_destroy": (None, [_VP]),
    "qhbm_plan_info": (ctypes.c_int, [_VP, ctypes.POINTER(_I64)]),
    "qhbm_expectation_forward": (ctypes.c_int, [_VP, _VP, _I64, _VP, _VP, _VP]),
    "qhbm_expectation_adjoint": (ctypes.c_int, [_VP, _VP, _I64, _VP, _VP, _VP, _VP, _I32, _I32, _VP]),
    "qhbm_expectation_forward_rows": (ctypes.c_int, [_VP, _VP, _I64, _VP, _VP, _VP]),
    "qhbm_expectation_adjoint_rows": (ctypes.c_int, [_VP, _VP, _I64, _VP, _VP, _VP, _VP, _I32, _I32, _VP]),
    "qhbm_expectation_host": (ctypes.c_int, [_VP, _VP, _I64, _VP, _VP, _VP, _VP, _I32, _VP]),
    "qhbm_debug_state": (ctypes.c_int, [_VP, _U64, _VP, _VP, _VP]),
    "qhbm_final_states": (ctypes.c_int, [_VP, _VP, _I64, _VP, _VP, _VP]),
    "qhbm_sample_states": (ctypes.c_int, [_VP, _I64, _I32, _VP, _I64, _U64, _U64, _VP, _VP]),
    "qhbm_binomial_shots": (ctypes.c_int, [_VP, _I64, _I64, _U64, _U64, _VP, _VP]),
    "qhbm_pack_bits": (ctypes.c_int, [_VP, _I64, _I32, _VP, _VP, _VP]),
    "qhbm_unpack_bits": (ctypes.c_int, [_VP, _I64, _I32, _VP, _VP, _VP]),
    "qhbm_unique_workspace_bytes": (_I64, [_I64]),
    "qhbm_unique_with_counts": (ctypes.c_int, [_VP, _I64, _VP, _VP, _VP, _VP, _VP, _VP]),
    "qhbm_segment_sum": (ctypes.c_int, [_VP, _VP, _I64, _I32, _VP, _I64, _VP]),
    "qhbm_energy_rows": (ctypes.c_int, [ctypes.POINTER(EnergyDesc), _VP, _I64, _VP, _VP]),
    "qhbm_ebm_sweep": (ctypes.c_int, [ctypes.POINTER(EnergyDesc), _U64, _U64, _VP, _VP, _VP]),
    "qhbm_sample_workspace_bytes": (_I64, [_I64]),
    "qhbm_categorical_sample": (ctypes.c_int, [_VP, _I64, _U64, _U64, _U64, _U64, _I64, _VP, _VP, _VP]),
    "qhbm_score_gradient": (ctypes.c_int, [_VP, _VP, _I64, _VP, _I32, _VP, _VP, _VP, _I32, _VP, ctypes.c_float, _VP, _VP, _VP]),
    "qhbm_categorical_prepare": (ctypes.c_int, [_VP, _I64, _I32, ctypes.c_float, _VP, _VP]),
    "qhbm_categorical_draw": (ctypes.c_int, [_VP, _I64, _U64, _VP, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                             _U64, _U64, _U64, _I64, _VP, _VP]),
    "qhbm_bernoulli_sample": (ctypes.c_int, [_VP, _I32, _VP, _U64, _U64, _U64, _I64, _VP, _VP]),
    "qhbm_weighted_sum": (ctypes.c_int, [_VP, _VP, _I64, _I32, _VP, _VP]),
    "qhbm_comm_unique_id": (ctypes.c_int, [_VP, _I32]),
    "qhbm_comm_create": (ctypes.c_int, [_VP, _I32, _I32, ctypes.POINTER(_VP)]),
    "qhbm_comm_adopt": (ctypes.c_int, [_VP, ctypes.POINTER(_VP)]),
    "qhbm_comm_info": (ctypes.c_int, [_VP, ctypes.POINTER(_I32), ctypes.POINTER(_I32), ctypes.POINTER(_I32)]),
    "qhbm_comm_destroy": (None, [_VP]),
    "qhbm_allreduce": (ctypes.c_int, [_VP, _VP, _I64, _I32, _VP]),
}


def lib():
  """Loads the shared library once; raises NativeError if it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise NativeError(
          f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
          "(nvcc, sm_100a).  The qhbm_b200 engine has no CPU fallback.")
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
      fn = getattr(handle, name)
      fn.restype = res
      fn.argtypes = args
    _lib = handle
  return _lib


def check(status):
  if status != 0:
    raise NativeError(lib().qhbm_last_error().decode("utf-8", "replace"))


def ptr(t):
  """Raw device/host pointer of a torch tensor or numpy array (None -> NULL)."""
  if t is None:
    return None
  if isinstance(t, np.ndarray):
    return t.ctypes.data
  return t.data_ptr()
