"""TensorFlow binding of the engine (import-guarded: TensorFlow is not installable in this image, so
this module is exercised only where TF >= 2.7 with GPU support exists).

`expectation(plan, basis_idx, symbol_values)` is the drop-in for the `tfq.layers.Expectation()` call at
/root/reference/qhbmlib/inference/qnn.py:134-138: a `tf.custom_gradient` whose forward is
`qhbm_expectation_forward` and whose backward is `qhbm_expectation_adjoint`, with tensors handed over
as raw device pointers through DLPack.
"""
import ctypes

from qhbmlib import _native as nat

try:
  import tensorflow as tf  # pylint: disable=import-error
except Exception:  # pragma: no cover
  tf = None


def available():
  return tf is not None


def _dev_ptr(t):  # pragma: no cover - needs TensorFlow
  cap = tf.experimental.dlpack.to_dlpack(t)
  get = ctypes.pythonapi.PyCapsule_GetPointer
  get.restype = ctypes.c_void_p
  get.argtypes = [ctypes.py_object, ctypes.c_char_p]
  managed = get(cap, b"dltensor")
  return ctypes.cast(managed, ctypes.POINTER(ctypes.c_void_p))[0]  # DLTensor.data is the first field


def expectation(plan, basis_idx, symbol_values, grad_mode="tfq_fd"):  # pragma: no cover - needs TensorFlow
  """f32[U, O] expectations, differentiable w.r.t. `symbol_values` (f32[P]) under tf.GradientTape.
  `plan` is a qhbmlib.engine.ExpectationPlan; `basis_idx` an int64 GPU tensor [U]."""
  if tf is None:
    raise ImportError("TensorFlow is required for qhbmlib.tf_adapter")
  from qhbmlib import engine
  lib = nat.lib()
  mode = engine.GRAD_MODES[grad_mode]

  @tf.custom_gradient
  def op(values):
    u = int(basis_idx.shape[0])
    out = tf.zeros([u, plan.n_ops], tf.float32)
    nat.check(lib.qhbm_expectation_forward(plan._plan, _dev_ptr(basis_idx), u, _dev_ptr(values), _dev_ptr(out),
                                           None))

    def grad(upstream):
      g = tf.zeros_like(values)
      e = tf.zeros([u, plan.n_ops], tf.float32)
      nat.check(lib.qhbm_expectation_adjoint(plan._plan, _dev_ptr(basis_idx), u, _dev_ptr(values),
                                             _dev_ptr(tf.identity(upstream)), _dev_ptr(e), _dev_ptr(g), 0, mode, None))
      return g

    return out, grad

  return op(symbol_values)
