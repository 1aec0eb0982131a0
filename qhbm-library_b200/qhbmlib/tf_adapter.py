"""TensorFlow binding of the engine.  EXPERIMENTAL: TensorFlow is not installable in this image, so the
module has never run under a real TensorFlow; tests/test_adapters_fake_modules.py executes every line of
it against a minimal stand-in `tensorflow` module (DLPack hand-over, custom_gradient wiring, shapes), and
INTEGRATION.md describes what a maintainer has to check on a TF >= 2.7 GPU install.

`expectation(plan, basis_idx, symbol_values)` is the drop-in for the `tfq.layers.Expectation()` call at
/root/reference/qhbmlib/inference/qnn.py:134-138: a `tf.custom_gradient` whose forward is
`qhbm_expectation_forward` and whose backward is `qhbm_expectation_adjoint`.

Memory and ordering rules (the points a raw-pointer binding gets wrong):
  * inputs cross as DLPack capsules that are CONSUMED (`torch.utils.dlpack.from_dlpack`), so TF's buffers
    stay alive for the duration of the call and are only read;
  * outputs are allocated by this side (never written into a TF tensor, which may be a shared constant) and
    handed to TF with `tf.experimental.dlpack.from_dlpack`;
  * the engine runs on torch's current stream; before TF sees an output the stream is synchronised, because
    TF's compute stream is not reachable from Python and no cross-framework event exists.
"""
import torch
from torch.utils import dlpack as torch_dlpack

try:
  import tensorflow as tf  # pylint: disable=import-error
except Exception:  # TensorFlow absent: `available()` is False and `expectation` raises
  tf = None


def available():
  return tf is not None


def _to_torch(t):
  """TF tensor -> torch view of the same device memory (the capsule is consumed)."""
  return torch_dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(t))


def _to_tf(t):
  """torch tensor -> TF tensor owning a reference to the same memory, after the producing stream finished."""
  if t.is_cuda:
    torch.cuda.current_stream(t.device).synchronize()
  return tf.experimental.dlpack.from_dlpack(torch_dlpack.to_dlpack(t))


def expectation(plan, basis_idx, symbol_values, grad_mode="tfq_fd"):
  """f32[U, O] expectations, differentiable w.r.t. `symbol_values` (f32[P], or f32[U, P] with one row per
  state) under tf.GradientTape.
  `plan` is a qhbmlib.engine.ExpectationPlan; `basis_idx` an int64 GPU tensor [U]."""
  if tf is None:
    raise ImportError("TensorFlow is required for qhbmlib.tf_adapter")
  basis = _to_torch(basis_idx)

  @tf.custom_gradient
  def op(values):
    vals = _to_torch(values).to(torch.float32).contiguous()
    out = plan.forward(basis, vals)

    def grad(upstream):
      up = _to_torch(tf.identity(upstream)).to(torch.float32).contiguous()
      # symbol_values f32[U, P] (one row per state, the TFQ op's own signature): gradient f32[U, P] as well
      _, g = plan.forward_adjoint(basis, vals, up, per_state=vals.dim() == 2, grad_mode=grad_mode)
      return _to_tf(g)

    return _to_tf(out), grad

  return op(symbol_values)
