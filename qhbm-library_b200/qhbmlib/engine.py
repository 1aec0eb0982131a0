"""Thin Python layer over the C ABI: plans for circuit x observables and bitstring utilities.

All tensors are torch CUDA tensors; torch is used only for device memory and streams
(`tensor.data_ptr()`, `torch.cuda.current_stream()`); every kernel is in libqhbm_b200.so.
"""
import ctypes

import numpy as np
import torch

from qhbmlib import _native as nat

GRAD_MODES = {"exact": nat.GRAD_EXACT, "tfq_fd": nat.GRAD_TFQ_FD, "tfq_fd_f32": nat.GRAD_TFQ_FD_F32}


def _stream():
  """torch's current CUDA stream as a raw handle.  `torch.cuda.current_stream().cuda_stream` costs ~20 us of
  Python per call (18 calls in one VQT step); the private raw accessor it wraps is a single C call."""
  raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
  if raw is not None:
    return ctypes.c_void_p(raw(torch.cuda.current_device()))
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, name, dtype=None):
  if not isinstance(t, torch.Tensor) or not t.is_cuda:
    raise TypeError(f"{name} must be a CUDA torch tensor (the engine has no CPU path)")
  if dtype is not None and t.dtype != dtype:
    raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
  return t.contiguous()


def terms_from_pauli_sums(pauli_sums, n_qubits):
  """[(coeff, {qubit_index: 'X'|'Y'|'Z'}), ...] per observable -> (TERM_DTYPE[], offsets)."""
  rows, offsets = [], [0]
  for op in pauli_sums:
    for coeff, paulis in op:
      x = z = 0
      for q, p in paulis.items():
        bit = 1 << (n_qubits - 1 - int(q))
        if p in ("X", "Y"):
          x |= bit
        if p in ("Z", "Y"):
          z |= bit
      rows.append((float(np.real(coeff)), x, z))
    offsets.append(len(rows))
  terms = np.zeros(len(rows), dtype=nat.TERM_DTYPE)
  for i, r in enumerate(rows):
    terms[i] = r
  return terms, np.asarray(offsets, dtype=np.int32)


class ExpectationPlan:
  """Compiled (circuit, observables) pair.  Stands where qhbmlib hands serialized circuits
  and PauliSums to `tfq.layers.Expectation()` (qhbmlib/inference/qnn.py:112,134-138)."""

  def __init__(self, gates, n_qubits, n_symbols, terms, offsets, with_gradient=True,
               tile_qubits=0, reg_qubits=0):
    lib = nat.lib()
    gates = np.ascontiguousarray(gates, dtype=nat.GATE_DTYPE)
    terms = np.ascontiguousarray(terms, dtype=nat.TERM_DTYPE)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    self.n_qubits, self.n_symbols = int(n_qubits), int(n_symbols)
    self.n_ops = len(offsets) - 1
    self.with_gradient = bool(with_gradient)
    self._circuit = ctypes.c_void_p()
    self._ops = ctypes.c_void_p()
    self._plan = ctypes.c_void_p()
    nat.check(lib.qhbm_circuit_create(gates.ctypes.data, len(gates), self.n_qubits, self.n_symbols,
                                      ctypes.byref(self._circuit)))
    nat.check(lib.qhbm_ops_create(terms.ctypes.data, offsets.ctypes.data, self.n_ops, self.n_qubits,
                                  ctypes.byref(self._ops)))
    nat.check(lib.qhbm_plan_create(self._circuit, self._ops, int(self.with_gradient), int(tile_qubits),
                                   int(reg_qubits), ctypes.byref(self._plan)))
    info = (ctypes.c_int64 * 8)()
    nat.check(lib.qhbm_plan_info(self._plan, info))
    self.info = dict(zip(("sweeps_fwd", "sweeps_bwd", "passes", "ops", "tile_qubits", "reg_qubits",
                          "launches", "chunk"), list(info)))
    self.n_eff = max(self.n_qubits, self.info["reg_qubits"] + 5)

  def __del__(self):
    try:
      lib = nat.lib()
      if self._plan:
        lib.qhbm_plan_destroy(self._plan)
      if self._ops:
        lib.qhbm_ops_destroy(self._ops)
      if self._circuit:
        lib.qhbm_circuit_destroy(self._circuit)
    except Exception:  # interpreter shutdown
      pass

  def _basis(self, basis_idx):
    # int64 storage is reinterpreted as the uint64 the ABI names (indices are < 2^30)
    return _require_cuda(basis_idx, "basis_idx", torch.int64)

  def _symbols(self, symbols, u):
    """symbols f32[P] (one row shared by all states, the reference's case: qnn.py:74-76) or f32[U, P]
    (one row per state, the TFQ op's general form).  Returns (tensor, has_rows)."""
    symbols = _require_cuda(symbols, "symbols", torch.float32)
    if symbols.dim() == 2:
      if tuple(symbols.shape) != (u, self.n_symbols):
        raise ValueError(f"per-state symbols must have shape {(u, self.n_symbols)}, got {tuple(symbols.shape)}")
      return symbols, True
    if symbols.numel() < self.n_symbols:
      raise ValueError(f"symbols must have {self.n_symbols} entries, got {symbols.numel()}")
    return symbols, False

  def forward(self, basis_idx, symbols):
    """f32[U, O] expectation values (TfqSimulateExpectation)."""
    basis_idx = self._basis(basis_idx)
    u = basis_idx.shape[0]
    symbols, rows = self._symbols(symbols, u)
    out = torch.empty((u, self.n_ops), dtype=torch.float32, device=basis_idx.device)
    fn = nat.lib().qhbm_expectation_forward_rows if rows else nat.lib().qhbm_expectation_forward
    nat.check(fn(self._plan, nat.ptr(basis_idx), u, nat.ptr(symbols), nat.ptr(out), _stream()))
    return out

  def forward_adjoint(self, basis_idx, symbols, dgrad, per_state=False, grad_mode="exact"):
    """(f32[U,O], f32[P] or f32[U,P]): expectations and adjoint gradient (TfqAdjointGradient)."""
    basis_idx = self._basis(basis_idx)
    dgrad = _require_cuda(dgrad, "dgrad", torch.float32)
    u = basis_idx.shape[0]
    symbols, rows = self._symbols(symbols, u)
    if tuple(dgrad.shape) != (u, self.n_ops):
      raise ValueError(f"dgrad must have shape {(u, self.n_ops)}")
    out = torch.empty((u, self.n_ops), dtype=torch.float32, device=basis_idx.device)
    gshape = (u, self.n_symbols) if per_state else (self.n_symbols,)
    grad = torch.zeros(gshape, dtype=torch.float32, device=basis_idx.device)
    fn = nat.lib().qhbm_expectation_adjoint_rows if rows else nat.lib().qhbm_expectation_adjoint
    nat.check(fn(self._plan, nat.ptr(basis_idx), u, nat.ptr(symbols), nat.ptr(dgrad), nat.ptr(out), nat.ptr(grad),
                 int(per_state), GRAD_MODES[grad_mode], _stream()))
    return out, grad

  def run_host(self, basis_idx, symbols, dgrad=None, grad_mode="exact", stream=None):
    """Same computation from HOST numpy buffers through `qhbm_expectation_host`."""
    basis_idx = np.ascontiguousarray(basis_idx, dtype=np.uint64)
    symbols = np.ascontiguousarray(symbols, dtype=np.float32)
    u = basis_idx.shape[0]
    out = np.empty((u, self.n_ops), dtype=np.float32)
    grad = None
    if dgrad is not None:
      dgrad = np.ascontiguousarray(dgrad, dtype=np.float32)
      grad = np.zeros(self.n_symbols, dtype=np.float32)
    nat.check(nat.lib().qhbm_expectation_host(
        self._plan, basis_idx.ctypes.data, u, symbols.ctypes.data,
        None if dgrad is None else dgrad.ctypes.data, out.ctypes.data,
        None if grad is None else grad.ctypes.data, GRAD_MODES[grad_mode],
        ctypes.c_void_p(stream) if stream else None))
    return out, grad

  def state(self, basis_index, symbols):
    """complex64[2^n] final state U|basis> (debug / parity)."""
    symbols = _require_cuda(symbols, "symbols", torch.float32)
    out = torch.zeros((1 << self.n_qubits, 2), dtype=torch.float32, device=symbols.device)
    nat.check(nat.lib().qhbm_debug_state(self._plan, ctypes.c_uint64(int(basis_index)), nat.ptr(symbols),
                                         nat.ptr(out), _stream()))
    return torch.view_as_complex(out)

  def final_states(self, basis_idx, symbols):
    """complex64[U, 2^n]: row u = U|basis_u>, big-endian amplitudes (tfq.layers.State)."""
    basis_idx = self._basis(basis_idx)
    symbols = _require_cuda(symbols, "symbols", torch.float32)
    u = basis_idx.shape[0]
    out = torch.zeros((u, 1 << self.n_qubits, 2), dtype=torch.float32, device=basis_idx.device)
    nat.check(nat.lib().qhbm_final_states(self._plan, nat.ptr(basis_idx), u, nat.ptr(symbols), nat.ptr(out),
                                          _stream()))
    return torch.view_as_complex(out)


def sample_states(states, counts, seed):
  """Measurement shots of many states.  states complex64[U, 2^n]; counts int[U] shots per state.
  Returns (int64[sum counts] basis indices grouped by state, int64[U+1] offsets)."""
  states = _require_cuda(states, "states", torch.complex64)
  u, dim = states.shape
  n = int(dim).bit_length() - 1
  if (1 << n) != dim:
    raise ValueError("states must have 2^n columns")
  counts = counts.to(device=states.device, dtype=torch.int64)
  offsets = torch.zeros(u + 1, dtype=torch.int64, device=states.device)
  offsets[1:] = torch.cumsum(counts, 0)
  total = int(offsets[-1].item())
  out = torch.empty(total, dtype=torch.int64, device=states.device)
  flat = torch.view_as_real(states)
  nat.check(nat.lib().qhbm_sample_states(nat.ptr(flat), u, n, nat.ptr(offsets), total, int(seed[0]), int(seed[1]),
                                         nat.ptr(out), _stream()))
  return out, offsets


def binomial_shots(exact, shots, seed):
  """Shot-noise estimate of +-1 valued measurements with exact means `exact` (any shape):
  (2 Binomial(shots, (1 + exact) / 2) - shots) / shots, element-wise."""
  exact = _require_cuda(exact, "exact", torch.float32)
  out = torch.empty_like(exact)
  nat.check(nat.lib().qhbm_binomial_shots(nat.ptr(exact), exact.numel(), int(shots), int(seed[0]), int(seed[1]),
                                          nat.ptr(out), _stream()))
  return out


def _shift_array(shifts):
  return np.ascontiguousarray(shifts, dtype=np.int32)


def pack_bits(bits, shifts):
  """int8[N, n] CUDA -> uint64-as-int64[N] keys, key = sum_j bits[:, j] << shifts[j]."""
  bits = _require_cuda(bits, "bits", torch.int8)
  n_rows, n_bits = bits.shape
  keys = torch.empty((n_rows,), dtype=torch.int64, device=bits.device)
  sh = _shift_array(shifts)
  nat.check(nat.lib().qhbm_pack_bits(nat.ptr(bits), n_rows, n_bits, sh.ctypes.data, nat.ptr(keys), _stream()))
  return keys


def unpack_bits(keys, n_bits, shifts):
  keys = _require_cuda(keys, "keys", torch.int64)
  bits = torch.empty((keys.shape[0], n_bits), dtype=torch.int8, device=keys.device)
  sh = _shift_array(shifts)
  nat.check(nat.lib().qhbm_unpack_bits(nat.ptr(keys), keys.shape[0], n_bits, sh.ctypes.data, nat.ptr(bits),
                                       _stream()))
  return bits


def unique_with_counts(keys):
  """First-occurrence unique of int64 keys: (unique[U], idx int32[N], count int32[U])."""
  keys = _require_cuda(keys, "keys", torch.int64)
  n = keys.shape[0]
  dev = keys.device
  uniq = torch.empty((n,), dtype=torch.int64, device=dev)
  idx = torch.empty((n,), dtype=torch.int32, device=dev)
  count = torch.empty((n,), dtype=torch.int32, device=dev)
  n_unique = torch.zeros((1,), dtype=torch.int64, device=dev)
  ws_bytes = nat.lib().qhbm_unique_workspace_bytes(n)
  ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=dev)
  nat.check(nat.lib().qhbm_unique_with_counts(nat.ptr(keys), n, nat.ptr(uniq), nat.ptr(idx), nat.ptr(count),
                                              nat.ptr(n_unique), nat.ptr(ws), _stream()))
  u = int(n_unique.item())
  return uniq[:u], idx, count[:u]


def segment_sum(vals, idx, n_unique):
  vals = _require_cuda(vals, "vals", torch.float32)
  idx = _require_cuda(idx, "idx", torch.int32)
  flat = vals.reshape(vals.shape[0], -1)
  out = torch.empty((n_unique, flat.shape[1]), dtype=torch.float32, device=vals.device)
  nat.check(nat.lib().qhbm_segment_sum(nat.ptr(flat), nat.ptr(idx), flat.shape[0], flat.shape[1],
                                       nat.ptr(out), n_unique, _stream()))
  return out.reshape((n_unique,) + tuple(vals.shape[1:]))


def weighted_sum(counts, vals):
  """float64[width + 1]: sum_u counts[u] * vals[u, :], then sum_u counts[u]."""
  counts = _require_cuda(counts, "counts", torch.int32)
  vals = _require_cuda(vals, "vals", torch.float32)
  flat = vals.reshape(vals.shape[0], -1)
  out = torch.empty((flat.shape[1] + 1,), dtype=torch.float64, device=vals.device)
  nat.check(nat.lib().qhbm_weighted_sum(nat.ptr(counts), nat.ptr(flat), flat.shape[0], flat.shape[1],
                                        nat.ptr(out), _stream()))
  return out


def score_gradient(keys, counts, vals, upstream, average, masks, total_count, scale=1.0):
  """f32[T]: E[c] E[f_t] - E[c f_t] over the given unique rows (see qhbm_score_gradient)."""
  keys = _require_cuda(keys, "keys", torch.int64)
  counts = _require_cuda(counts, "counts", torch.int32)
  vals = _require_cuda(vals, "vals", torch.float32)
  upstream = _require_cuda(upstream, "upstream", torch.float32)
  average = _require_cuda(average, "average", torch.float32)
  masks = _require_cuda(masks, "masks", torch.int32)
  total_count = _require_cuda(total_count, "total_count", torch.float64)
  n_rows, width = vals.shape
  out = torch.empty((masks.shape[0],), dtype=torch.float32, device=keys.device)
  ws = torch.empty((n_rows + 4,), dtype=torch.float32, device=keys.device)
  nat.check(nat.lib().qhbm_score_gradient(nat.ptr(keys), nat.ptr(counts), n_rows, nat.ptr(vals), width,
                                          nat.ptr(upstream), nat.ptr(average), nat.ptr(masks), masks.shape[0],
                                          nat.ptr(total_count), float(scale), nat.ptr(out), nat.ptr(ws), _stream()))
  return out


class EnergyDescriptor:
  """Device-side description of an energy function for the EBM kernels."""

  def __init__(self, kind, n_bits, masks=None, theta=None, layers=None):
    self.desc = nat.EnergyDesc()
    self.desc.kind = kind
    self.desc.n_bits = n_bits
    self._keep = []
    if kind in (nat.ENERGY_BERNOULLI, nat.ENERGY_KOBE):
      masks = _require_cuda(masks, "masks", torch.int32)
      theta = _require_cuda(theta, "theta", torch.float32)
      self._keep += [masks, theta]
      self.desc.n_terms = masks.shape[0]
      self.desc.d_masks = masks.data_ptr()
      self.desc.d_theta = theta.data_ptr()
    else:
      acts = {"linear": 0, None: 0, "tanh": 1, "relu": 2}
      self.desc.n_layers = len(layers)
      self.desc.widths[0] = n_bits
      for l, (w, b, act) in enumerate(layers):
        w = _require_cuda(w, "weight", torch.float32)
        b = _require_cuda(b, "bias", torch.float32)
        self._keep += [w, b]
        self.desc.widths[l + 1] = w.shape[1]
        self.desc.act[l] = acts[act]
        self.desc.d_weights[l] = w.data_ptr()
        self.desc.d_bias[l] = b.data_ptr()

  def energies(self, keys):
    keys = _require_cuda(keys, "keys", torch.int64)
    out = torch.empty((keys.shape[0],), dtype=torch.float32, device=keys.device)
    nat.check(nat.lib().qhbm_energy_rows(ctypes.byref(self.desc), nat.ptr(keys), keys.shape[0], nat.ptr(out),
                                         _stream()))
    return out

  def sweep(self, lo, hi, want_logits=True, device="cuda"):
    """(logits f32[hi-lo] or None, stats f64[3] = (max, sum exp(l-max), sum exp(l-max) l))."""
    logits = torch.empty((hi - lo,), dtype=torch.float32, device=device) if want_logits else None
    stats = torch.empty((3,), dtype=torch.float64, device=device)
    nat.check(nat.lib().qhbm_ebm_sweep(ctypes.byref(self.desc), lo, hi, nat.ptr(logits), nat.ptr(stats),
                                       _stream()))
    return logits, stats


def categorical_sample(logits, n_samples, seed, first_sample=0, row_offset=0):
  logits = _require_cuda(logits, "logits", torch.float32)
  out = torch.empty((n_samples,), dtype=torch.int64, device=logits.device)
  ws = torch.empty((nat.lib().qhbm_sample_workspace_bytes(logits.shape[0]),), dtype=torch.uint8,
                   device=logits.device)
  nat.check(nat.lib().qhbm_categorical_sample(nat.ptr(logits), logits.shape[0], row_offset, int(seed[0]),
                                              int(seed[1]), first_sample, n_samples, nat.ptr(out), nat.ptr(ws),
                                              _stream()))
  return out


class CategoricalSampler:
  """Prefix sums over fixed logits, prepared once and reused by every draw (the reference rebuilds the
  tfd.Categorical only when the energy variables change, ebm.py:467-469).  `given_max`: the global
  maximum when the logits are one shard of a range split over ranks."""

  def __init__(self, logits, given_max=None):
    self.logits = _require_cuda(logits, "logits", torch.float32)
    n = self.logits.shape[0]
    self.ws = torch.empty((nat.lib().qhbm_sample_workspace_bytes(n),), dtype=torch.uint8, device=logits.device)
    nat.check(nat.lib().qhbm_categorical_prepare(nat.ptr(self.logits), n, int(given_max is not None),
                                                 float(given_max or 0.0), nat.ptr(self.ws), _stream()))
    self._nblocks = (n + 255) // 256

  def local_mass(self):
    """float64 device scalar: sum_rows exp(logit - max)."""
    off = 256 + 8 * self._nblocks
    return self.ws[off:off + 8].view(torch.float64)

  def draw(self, n_samples, seed, first_sample=0, row_offset=0, mass_interval=None, out=None):
    """int64[n_samples] row indices.  mass_interval = (begin, end, total): only the samples whose point of
    the global cumulative mass falls in [begin, end) are written (others keep the value in `out`)."""
    if out is None:
      out = torch.zeros((n_samples,), dtype=torch.int64, device=self.logits.device)
    b, e, t = mass_interval if mass_interval is not None else (0.0, 0.0, 0.0)
    nat.check(nat.lib().qhbm_categorical_draw(nat.ptr(self.logits), self.logits.shape[0], int(row_offset),
                                              nat.ptr(self.ws), float(b), float(e), float(t), int(seed[0]),
                                              int(seed[1]), int(first_sample), int(n_samples), nat.ptr(out), _stream()))
    return out


def bernoulli_sample(logits, shifts, n_samples, seed, first_sample=0):
  logits = _require_cuda(logits, "logits", torch.float32)
  out = torch.empty((n_samples,), dtype=torch.int64, device=logits.device)
  sh = _shift_array(shifts)
  nat.check(nat.lib().qhbm_bernoulli_sample(nat.ptr(logits), logits.shape[0], sh.ctypes.data, int(seed[0]),
                                            int(seed[1]), first_sample, n_samples, nat.ptr(out), _stream()))
  return out
