/* qhbm_b200.h -- C ABI of the B200-native QHBM hot-path engine.
 *
 * Drop-in boundary for the one data-parallel path of google/qhbm-library
 * (reference citations are file:line under /root/reference):
 *
 *   - per-unique-bitstring  basis state -> QNN circuit -> PauliSum expectation,
 *     plus its adjoint gradient, i.e. what qhbmlib reaches through
 *     `tfq.layers.Expectation()` (qhbmlib/inference/qnn.py:112,134-138): the TFQ
 *     custom ops `TfqSimulateExpectation` and `TfqAdjointGradient`, together with
 *     the circuit-string plumbing that feeds them (`tfq.resolve_parameters`,
 *     `tfq.append_circuit`; qhbmlib/models/circuit.py:129-136);
 *   - AnalyticEnergyInference's exhaustive 2^n energy / logsumexp / entropy /
 *     categorical sampling sweep (qhbmlib/inference/ebm.py:445-492) and
 *     BernoulliEnergyInference's sampler (ebm.py:559-561);
 *   - `unique_bitstrings_with_counts` (qhbmlib/utils.py:61-78,
 *     tf.raw_ops.UniqueWithCountsV2) in first-occurrence order;
 *   - (widened) the simulation and measurement halves of
 *     `tfq.layers.Sample`, `tfq.layers.SampledExpectation` and `tfq.layers.Unitary`
 *     as used by SampledQuantumInference and the dense metrics
 *     (qhbmlib/inference/qnn.py:142-292, qnn_utils.py:23-33);
 *   - the TFQ ops' general symbol signature, `symbol_values f32[U,P]` with one
 *     row per circuit (qhbm_expectation_*_rows), and the one collective of a
 *     sharded step: the sum over ranks of the packed count-weighted partial sums
 *     behind `weighted_average` (qhbmlib/utils.py:43-58; qhbm_comm_*, qhbm_allreduce).
 *
 * Conventions: every function returns 0 on success, non-zero on error; the
 * message is available from qhbm_last_error() (thread-local).  All `d_` pointers
 * are DEVICE pointers owned by the caller; the library never frees or retains
 * them past the call.  Work is enqueued on `stream` (a cudaStream_t passed as
 * void*); no call synchronises the device unless documented.  Handles are
 * immutable after creation and may be used from several streams, except that
 * one plan owns one workspace: the library orders the device work of calls on
 * the SAME plan (a call that arrives on another stream than the previous one
 * waits for it through an event), so they never race; use several plans for
 * real concurrency.
 *
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef QHBM_B200_H_
#define QHBM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- gate table (replaces the serialized cirq/TFQ circuit protos) ---------- */
/* Gate set = what TFQ 0.6.1 can serialise.  Matrices follow
 * cirq 0.14.1.  Parameter k of a gate has value  cnst[k] + scalar[k]*symbols[sym[k]]
 * (sym[k] < 0: constant) -- TFQ's `exponent` / `exponent_scalar` pair. */
enum qhbm_gate_type {
  QHBM_GATE_I = 0,
  QHBM_GATE_XPOW = 1,   /* params: exponent; uses gshift                      */
  QHBM_GATE_YPOW = 2,
  QHBM_GATE_ZPOW = 3,
  QHBM_GATE_HPOW = 4,
  QHBM_GATE_CZPOW = 5,  /* two-qubit; q0 is the more significant matrix index */
  QHBM_GATE_CNOTPOW = 6,
  QHBM_GATE_SWAPPOW = 7,
  QHBM_GATE_ISWAPPOW = 8,
  QHBM_GATE_XXPOW = 9,
  QHBM_GATE_YYPOW = 10,
  QHBM_GATE_ZZPOW = 11,
  QHBM_GATE_PHASEDXPOW = 12,     /* params: exponent, phase_exponent; gshift   */
  QHBM_GATE_FSIM = 13,           /* params: theta, phi (radians)               */
  QHBM_GATE_PHASEDISWAPPOW = 14, /* params: exponent, phase_exponent           */
  QHBM_GATE_NUM_TYPES = 15
};

typedef struct qhbm_gate {
  int32_t type;      /* enum qhbm_gate_type                                     */
  int32_t q0, q1;    /* index into the SORTED qubit list (circuit.py:54); q1=-1  */
  int32_t nparams;   /* 0..3                                                    */
  int32_t sym[3];    /* symbol index per parameter, -1 = constant               */
  float scalar[3];
  float cnst[3];
  float gshift;      /* cirq EigenGate global_shift                             */
} qhbm_gate_t;

/* ---- Pauli sums (replace serialized cirq.PauliSum protos) ------------------ */
/* One term = coeff * prod_q sigma_q.  xmask/zmask are over BASIS-INDEX bits
 * (qubit k <-> bit n-1-k): X -> x bit, Z -> z bit, Y -> both. */
typedef struct qhbm_pauli_term {
  float coeff;
  uint32_t xmask;
  uint32_t zmask;
} qhbm_pauli_term_t;

typedef struct qhbm_circuit qhbm_circuit_t;
typedef struct qhbm_ops qhbm_ops_t;
typedef struct qhbm_plan qhbm_plan_t;

enum qhbm_grad_mode {
  QHBM_GRAD_EXACT = 0,       /* analytic gate derivative                         */
  QHBM_GRAD_TFQ_FD = 1,      /* TFQ 0.6.1 adj_util.cc: central difference of the  */
                             /* gate matrix, eps = 5e-3 on the symbol value      */
  QHBM_GRAD_TFQ_FD_F32 = 2   /* same, matrices rounded to float32 before diffing */
};

const char* qhbm_last_error(void);
int qhbm_version(void);

/* Replaces: building `tf.string` circuits (models/circuit.py:129-162). */
int qhbm_circuit_create(const qhbm_gate_t* gates, int32_t n_gates, int32_t n_qubits,
                        int32_t n_symbols, qhbm_circuit_t** out);
void qhbm_circuit_destroy(qhbm_circuit_t* c);

/* Replaces: `tfq.convert_to_tensor([PauliSum...])` operands of the expectation op
 * (qnn.py:120-133).  term_offsets has n_ops+1 entries into `terms`. */
int qhbm_ops_create(const qhbm_pauli_term_t* terms, const int32_t* term_offsets,
                    int32_t n_ops, int32_t n_qubits, qhbm_ops_t** out);
void qhbm_ops_destroy(qhbm_ops_t* o);

/* Compiles circuit x observables into the sweep/pass program for this GPU.
 * with_gradient != 0 also compiles the reverse (adjoint) program.
 * tile_qubits / reg_qubits: 0 = library default. */
int qhbm_plan_create(const qhbm_circuit_t* c, const qhbm_ops_t* o, int32_t with_gradient,
                     int32_t tile_qubits, int32_t reg_qubits, qhbm_plan_t** out);
void qhbm_plan_destroy(qhbm_plan_t* p);
/* Fills: [0]=n_sweeps_forward [1]=n_sweeps_backward [2]=n_passes [3]=n_ops
 *        [4]=tile_qubits [5]=reg_qubits [6]=n_launches_per_chunk [7]=chunk size */
int qhbm_plan_info(const qhbm_plan_t* p, int64_t* out8);

/* Replaces: TfqSimulateExpectation (forward of qnn.py:134-138).
 *   d_basis_idx  u64[U]   basis index of each unique bitstring
 *   d_symbols    f32[P]   symbol values shared by all bitstrings (qnn.py:74-76 tiles one row)
 *   d_out        f32[U,O] <basis_u| U^dag H_j U |basis_u> */
int qhbm_expectation_forward(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                             const float* d_symbols, float* d_out, void* stream);

/* Replaces: TfqSimulateExpectation + TfqAdjointGradient in one pass.
 *   d_dgrad      f32[U,O] upstream gradient of every expectation
 *   d_grad_out   f32[P]   sum_u sum_j dgrad[u,j] d<H_j>_u / d symbol   (per_state == 0)
 *                f32[U,P] the un-reduced TFQ output                     (per_state != 0) */
int qhbm_expectation_adjoint(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                             const float* d_symbols, const float* d_dgrad, float* d_out,
                             float* d_grad_out, int32_t per_state, int32_t grad_mode,
                             void* stream);

/* The same two operators with ONE ROW OF SYMBOL VALUES PER STATE: the general form of the TFQ ops
 * (tfq_simulate_expectation / tfq_adjoint_gradient take symbol_values f32[U,P]; the reference always
 * tiles one row, qnn.py:74-76, which is the f32[P] form above and the fast path).
 *   d_symbol_rows f32[U,P]   row u holds the symbol values of state u
 * The gate coefficient tables are built per state (chunks of at most 1 GiB of tables); everything else,
 * including per_state / grad_mode and the outputs, is as in the shared-row entry points. */
int qhbm_expectation_forward_rows(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                                  const float* d_symbol_rows, float* d_out, void* stream);
int qhbm_expectation_adjoint_rows(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                                  const float* d_symbol_rows, const float* d_dgrad, float* d_out,
                                  float* d_grad_out, int32_t per_state, int32_t grad_mode,
                                  void* stream);

/* Same two entry points with HOST buffers (pageable or pinned): copies in, runs,
 * copies out and synchronises `stream`.  This is the call a TF custom-op shim binds. */
int qhbm_expectation_host(qhbm_plan_t* p, const uint64_t* h_basis_idx, int64_t n_states,
                          const float* h_symbols, const float* h_dgrad, float* h_out,
                          float* h_grad_out, int32_t grad_mode, void* stream);

/* Replaces: tfq.layers.State / tfq.layers.Unitary as used by qnn_utils.py:23-33 (the unitary is the
 * batch of final states of all 2^n basis inputs) and the simulation half of tfq.layers.Sample
 * (qnn.py:166, 283-289).
 *   d_states_out  complex64[U, 2^n] interleaved (re, im), row u = U|basis_u>, index big-endian */
int qhbm_final_states(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                      const float* d_symbols, float* d_states_out, void* stream);

/* Replaces: the measurement half of tfq.layers.Sample (qnn.py:262-292 `_sample`).
 *   d_states   complex64[U, 2^n] as written by qhbm_final_states
 *   d_offsets  i64[U+1]  shots of state u are d_out[d_offsets[u] .. d_offsets[u+1])
 *   d_out      u64[total_samples]  measured basis index (qubit k <-> bit n-1-k)
 * Philox4x32-10 counter = global shot index: a fixed seed repeats exactly. */
int qhbm_sample_states(const float* d_states, int64_t n_states, int32_t n_qubits,
                       const int64_t* d_offsets, int64_t total_samples, uint64_t seed0,
                       uint64_t seed1, uint64_t* d_out, void* stream);

/* Replaces: the shot noise of tfq.layers.SampledExpectation (qnn.py:255-260), which measures each
 * Pauli term with its own `shots` repetitions: d_out[i] = (2 Binomial(shots, (1+d_exact[i])/2)
 * - shots) / shots, drawn exactly (geometric gaps / BTRS rejection). */
int qhbm_binomial_shots(const float* d_exact, int64_t n, int64_t shots, uint64_t seed0,
                        uint64_t seed1, float* d_out, void* stream);

/* Debug / parity: final state of one bitstring, complex64[2^n] interleaved. */
int qhbm_debug_state(qhbm_plan_t* p, uint64_t basis_idx, const float* d_symbols,
                     float* d_state_out, void* stream);

/* ---- bitstring utilities ---------------------------------------------------- */
/* int8[N,n] rows -> u64 keys, key = sum_j b[i,j] << shift[j].  With
 * shift[j] = n-1-pi(j) this is the basis index incl. the reference's column
 * permutation (circuit.py:59-63,132-134). */
int qhbm_pack_bits(const int8_t* d_bits, int64_t n_rows, int32_t n_bits,
                   const int32_t* h_shift, uint64_t* d_keys, void* stream);
int qhbm_unpack_bits(const uint64_t* d_keys, int64_t n_rows, int32_t n_bits,
                     const int32_t* h_shift, int8_t* d_bits, void* stream);

/* Replaces: tf.raw_ops.UniqueWithCountsV2(axis=0) (utils.py:76-77) on packed keys.
 * Outputs are in FIRST-OCCURRENCE order.  d_unique/d_count need room for N rows;
 * the number of unique rows is written to d_n_unique (device int64).
 * d_workspace: at least qhbm_unique_workspace_bytes(N) bytes. */
int64_t qhbm_unique_workspace_bytes(int64_t n_rows);
int qhbm_unique_with_counts(const uint64_t* d_keys, int64_t n_rows, uint64_t* d_unique,
                            int32_t* d_idx, int32_t* d_count, int64_t* d_n_unique,
                            void* d_workspace, void* stream);

/* Replaces: tf.gather backward (segment-sum) of utils.expand_unique_results. */
int qhbm_segment_sum(const float* d_vals, const int32_t* d_idx, int64_t n_rows, int32_t width,
                     float* d_out, int64_t n_unique, void* stream);

/* ---- energy-based-model sweep ---------------------------------------------- */
enum qhbm_energy_kind {
  QHBM_ENERGY_BERNOULLI = 0,  /* E = sum_i (1-2b_i) theta_i       (energy.py:123-167) */
  QHBM_ENERGY_KOBE = 1,       /* E = sum_t theta_t prod_{i in t} (1-2b_i) (energy.py:170-209) */
  QHBM_ENERGY_MLP = 2         /* dense stack on raw bits           (energy.py:82-87)   */
};

typedef struct qhbm_energy_desc {
  int32_t kind;
  int32_t n_bits;
  /* BERNOULLI / KOBE: n_terms masks over the ROW INDEX bits (bit column j of the
   * bitstring <-> index bit n-1-j) and n_terms parameters. */
  int32_t n_terms;
  const uint32_t* d_masks;
  const float* d_theta;
  /* MLP: n_layers dense layers; layer l has weights f32[in_l, out_l] (row-major),
   * bias f32[out_l], activation act[l] (0 linear, 1 tanh, 2 relu); out of the last = 1. */
  int32_t n_layers;
  int32_t widths[9];          /* widths[0] = n_bits, widths[l+1] = out_l           */
  int32_t act[8];
  const float* d_weights[8];
  const float* d_bias[8];
} qhbm_energy_desc_t;

/* Energies of explicit rows given as packed index keys: f32[N]. */
int qhbm_energy_rows(const qhbm_energy_desc_t* e, const uint64_t* d_keys, int64_t n_rows,
                     float* d_energy, void* stream);

/* Replaces AnalyticEnergyInference._ready_inference + log_partition + entropy
 * (ebm.py:467-485) for rows [lo, hi) of the big-endian enumeration (ebm.py:445-447):
 *   d_logits  f32[hi-lo] = -E(row)        (may be NULL)
 *   d_stats   f64[3]: max logit m, s = sum exp(l-m), t = sum exp(l-m)*l
 *             => logZ = m + log s, entropy = logZ - t/s.  Partial stats of several
 *             ranks merge by rebasing to the common max: s = sum_r s_r e^{m_r - m}
 *             (one all-gather of the triples, or allreduce-max then allreduce-sum). */
int qhbm_ebm_sweep(const qhbm_energy_desc_t* e, uint64_t lo, uint64_t hi, float* d_logits,
                   double* d_stats, void* stream);

/* Replaces tfd.Categorical(logits).sample(N, seed) + gather (ebm.py:487-492) over
 * the local logits f32[n_rows]: inverse-CDF sampling with a counter-based Philox
 * stream keyed by (seed0, seed1, sample number).  Writes row indices (+ row_offset).
 * qhbm_ebm_sweep chooses its kernel from 2^n_bits, never from hi - lo, so a sweep of a shard
 * [lo, hi) writes the same logits, bit for bit, as a sweep of the whole range: rank-sharded
 * sampling (below) reproduces the single-GPU draw exactly.
 * d_workspace: at least qhbm_sample_workspace_bytes(n_rows). */
int64_t qhbm_sample_workspace_bytes(int64_t n_rows);
int qhbm_categorical_sample(const float* d_logits, int64_t n_rows, uint64_t row_offset,
                            uint64_t seed0, uint64_t seed1, uint64_t first_sample,
                            int64_t n_samples, uint64_t* d_samples, void* d_workspace,
                            void* stream);

/* The same sampler in two steps, for (a) many draws from unchanged logits and (b) a row range that is
 * sharded over ranks.
 *   qhbm_categorical_prepare: prefix sums of exp(l - max) over blocks of 256 rows, plus the sums of their
 *     eight 32-row sub-blocks, into d_workspace (a draw then scans 8 sums and at most 32 rows).
 *     use_given_max != 0:
 *     `given_max` (the GLOBAL maximum from the merged sweep statistics) replaces the local maximum, which
 *     also saves the max pass over the logits.  The local mass sum_rows exp(l - max) is the float64 at
 *     byte offset 256 + 8 * ceil(n_rows / 256) of the workspace.
 *   qhbm_categorical_draw: mass_total <= 0: every sample k in [first_sample, first_sample + n_samples) is
 *     drawn from the local rows, as qhbm_categorical_sample does.  mass_total > 0: sample k is drawn here
 *     iff u_k * mass_total lies in [mass_begin, mass_end) -- this rank's interval of the global cumulative
 *     mass -- and d_samples[k] is left untouched otherwise, so that the ranks' outputs (zero-initialised)
 *     add up to exactly the samples a single GPU would draw from the whole range
 *     (ebm.py:487-492 semantics, independent of the number of ranks). */
int qhbm_categorical_prepare(const float* d_logits, int64_t n_rows, int32_t use_given_max, float given_max,
                             void* d_workspace, void* stream);
int qhbm_categorical_draw(const float* d_logits, int64_t n_rows, uint64_t row_offset, const void* d_workspace,
                          double mass_begin, double mass_end, double mass_total, uint64_t seed0,
                          uint64_t seed1, uint64_t first_sample, int64_t n_samples, uint64_t* d_samples,
                          void* stream);

/* Replaces tfd.Bernoulli(logits, int8).sample(N, seed) (ebm.py:559-561): packed keys,
 * bit column j lands at key bit h_shift[j]. */
int qhbm_bernoulli_sample(const float* d_logits, int32_t n_bits, const int32_t* h_shift,
                          uint64_t seed0, uint64_t seed1, uint64_t first_sample,
                          int64_t n_samples, uint64_t* d_samples, void* stream);

/* Count-weighted reductions of utils.weighted_average (utils.py:43-58):
 *   d_out[w] = sum_u count[u]*vals[u,w]   (f64[width]),  d_out[width] = sum_u count[u]. */
int qhbm_weighted_sum(const int32_t* d_counts, const float* d_vals, int64_t n_rows,
                      int32_t width, double* d_out, void* stream);

/* Backward of EnergyInference._expectation (ebm.py:282-325) w.r.t. the parameters theta of a
 * parity-feature energy E(x) = sum_t theta_t (-1)^{parity(x & mask_t)} (BernoulliEnergy, KOBE;
 * Jacobian of energy_utils.py:97-110), over the unique rows handed in:
 *   d_grad_theta[t] = scale * sum_u (count_u / total) f_t(x_u) (E[c] - c_u),
 *   c_u = sum_j d_upstream[j] d_vals[u, j],  E[c] = sum_j d_upstream[j] d_average[j]
 * i.e. E[c] E[dE/dtheta_t] - E[c dE/dtheta_t].  *d_total_count: sum of counts over ALL rows (the
 * rows may be one rank's shard; the partial results then add up over ranks).
 * d_workspace: n_rows + 4 floats. */
int qhbm_score_gradient(const uint64_t* d_keys, const int32_t* d_counts, int64_t n_rows,
                        const float* d_vals, int32_t width, const float* d_upstream,
                        const float* d_average, const int32_t* d_masks, int32_t n_terms,
                        const double* d_total_count, float scale, float* d_grad_theta,
                        float* d_workspace, void* stream);

/* ---------------------------------------------------------------- collective
 * Replaces: the cross-replica sum behind utils.weighted_average (utils.py:43-58) when the unique
 * bitstrings, or the 2^n energy sweep, are sharded one process per GPU.  The only data ever exchanged
 * on the path is ONE small packed vector per step ([sum c<H_j> (O) | sum c | gradient (P) | score-function
 * partials]), summed in place over the ranks of the communicator with ncclAllReduce on
 * the caller's stream.  NCCL is bound at run time (libnccl.so.2 already in the process, else the system
 * copy; QHBM_NCCL_LIB overrides), so single-GPU users never load it.
 *   qhbm_comm_unique_id  rank 0 fills QHBM_COMM_ID_BYTES bytes and hands them to the other ranks by any
 *                        out-of-band means (the Python side broadcasts them through torch.distributed)
 *   qhbm_comm_create     collective over all ranks: communicator on the CURRENT device of the calling thread
 *   qhbm_comm_adopt      wraps an ncclComm_t the caller already owns (never destroyed by the library)
 *   qhbm_allreduce       d_buf[count] <- sum over ranks, in place, dtype QHBM_F32 or QHBM_F64 */
#define QHBM_COMM_ID_BYTES 128
enum { QHBM_F32 = 0, QHBM_F64 = 1 };
typedef struct qhbm_comm qhbm_comm_t;
int qhbm_comm_unique_id(uint8_t* out, int32_t out_bytes);
int qhbm_comm_create(const uint8_t* id, int32_t rank, int32_t nranks, qhbm_comm_t** out);
int qhbm_comm_adopt(void* nccl_comm, qhbm_comm_t** out);
int qhbm_comm_info(const qhbm_comm_t* c, int32_t* rank, int32_t* nranks, int32_t* nccl_version);
void qhbm_comm_destroy(qhbm_comm_t* c);
int qhbm_allreduce(qhbm_comm_t* c, void* d_buf, int64_t count, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QHBM_B200_H_ */
