// sim.cu -- C-ABI entry points of the state-vector path (plan handles, expectation
// forward / adjoint).  Declarations and the reference interfaces they replace are in
// include/qhbm_b200.h.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"
#include "plan.h"
#include "sim_launch.h"

namespace qhbm {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  void upload(const std::vector<T>& v) {
    reserve(std::max<size_t>(v.size(), 1));
    if (!v.empty()) QHBM_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) QHBM_CUDA(cudaFree(p));
    p = nullptr;
    cap = 0;  // a failed cudaMalloc below must not leave a stale capacity behind
    QHBM_CUDA(cudaMalloc(&p, n * sizeof(T)));
    cap = n;
  }
};

}  // namespace qhbm

using namespace qhbm;

struct qhbm_circuit {
  CircuitIR ir;
};
struct qhbm_ops {
  OpsIR ir;
};

struct qhbm_plan {
  HostPlan hp;
  DevBuf<DevPass> d_passes;
  DevBuf<PackedOp> d_ops;
  DevBuf<DevGradDesc> d_gdescs;
  DevBuf<PrepJob> d_jobs;
  DevBuf<int32_t> d_lists;
  DevBuf<qhbm_gate_t> d_gates;
  DevBuf<DevTerm> d_terms;
  DevBuf<DevTermGroup> d_groups;
  DevBuf<DevOpRange> d_opranges;
  DevBuf<DevDiagTerm> d_dterms;
  DevBuf<float> d_coef;
  // workspace (grown on demand)
  DevBuf<float2> d_psi, d_psi_alt, d_lam;
  DevBuf<double> d_eacc, d_gacc;
  // staging for the host-buffer entry point
  DevBuf<uint64_t> d_basis;
  DevBuf<float> d_sym, d_dgrad, d_out, d_gout;
  int chunk = 1;
  int sm_count = 148;
  std::mutex mu;
  // The scratch buffers above are per plan, so device work of two calls on the same plan must not
  // overlap: every call records `done` on its stream and the next call, if it arrives on another
  // stream, waits for it there (calls on one stream are ordered anyway).
  cudaEvent_t done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool used = false;
  ~qhbm_plan() { if (done) cudaEventDestroy(done); }
};

namespace qhbm {
void launch_sweep_gen(int K, bool adj, bool dense, const KernelArgs& ka, unsigned grid, int threads, size_t smem,
                      cudaStream_t s) {
  launch_sweep_impl<true>(K, adj, dense, ka, grid, threads, smem, s);
}
void allow_large_smem_gen() { QHBM_CUDA(allow_large_smem_impl<true>()); }
}  // namespace qhbm

namespace {
// Holds the plan's host lock and orders this call's device work after the previous call's.
struct PlanUse {
  qhbm_plan* p;
  cudaStream_t s;
  std::lock_guard<std::mutex> lk;
  PlanUse(qhbm_plan* plan, cudaStream_t stream) : p(plan), s(stream), lk(plan->mu) {
    if (!p->done) QHBM_CUDA(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
    if (p->used && p->last_stream != s) QHBM_CUDA(cudaStreamWaitEvent(s, p->done, 0));
  }
  ~PlanUse() {
    cudaEventRecord(p->done, s);
    p->last_stream = s;
    p->used = true;
  }
};
}  // namespace

namespace {

// Opt in to large dynamic shared memory once per plan creation (the attribute is per device and
// function and sticky; launches then only pass the size they need).
void allow_large_smem() {
  allow_large_smem_gen();
  allow_large_smem_lean();
}

void launch_any(const qhbm_plan* p, bool adj, const KernelArgs& ka, int n_states, cudaStream_t s) {
  const HostPlan& hp = p->hp;
  const int threads = 1 << (hp.T - hp.K);
  static const bool no_lean = std::getenv("QHBM_NO_LEAN") != nullptr;
  const auto launch = (hp.lean && !no_lean) ? launch_sweep_lean : launch_sweep_gen;
  // (the sparse first forward sweep only launches the one tile per state that holds its basis index)
  const unsigned grid = (ka.L.flags & LF_SPARSE_OUT) ? (unsigned)n_states : (unsigned)(n_states * hp.tiles());
  // forward sweeps of an adjoint plan (psi only, no expectation phase, no backward passes) run on the dense
  // forward kernel: one tile of shared memory and ~80 registers instead of two tiles and 128
  const bool psi_only = !(ka.L.flags & (LF_EXPECT | LF_LOAD_LAM | LF_STORE_LAM | LF_WRITE_STATE)) &&
                        ka.L.pass_b_end == ka.L.pass_b_begin;
  static const bool no_dense = std::getenv("QHBM_NO_DENSE_FWD") != nullptr;
  if (adj && psi_only && hp.K == 4 && threads <= 256 && hp.tiles() > 1 && !no_dense) {
    launch(4, false, true, ka, grid, threads, (size_t)8u * (1u << hp.T), s);
    QHBM_CUDA(cudaGetLastError());
    return;
  }
  // psi tile (+ lambda tile for the adjoint kernel; forward-only WHT needs a float scratch tile)
  const size_t smem = (size_t)(adj ? 2 : 1) * 8u * (1u << hp.T) + ((!adj && !hp.dterms.empty()) ? 4u * (1u << hp.T) : 0u);
  launch(hp.K, adj, false, ka, grid, threads, smem, s);
  QHBM_CUDA(cudaGetLastError());
}

int default_chunk(const qhbm_plan* p) {
  const HostPlan& hp = p->hp;
  if (const char* e = std::getenv("QHBM_CHUNK")) {
    int v = std::atoi(e);
    if (v > 0) return v;
  }
  const int tiles = hp.tiles();
  if (tiles == 1) return 1 << 20;  // single launch, no workspace
  // Multi-tile states round-trip through global memory between sweeps.  The sweeps are compute
  // bound (DRAM traffic is a few % of peak), so L2 residency does not matter; what matters is that
  // each launch is many waves deep.  Workspace budget: 4 GiB (of 180 GB).
  const size_t state_bytes = (size_t)8 << hp.n_eff;
  const size_t per_state = state_bytes * (hp.grad ? 2 : 1);
  size_t budget = (size_t)4 << 30;
  if (const char* e = std::getenv("QHBM_WORKSPACE_MB")) {
    long v = std::atol(e);
    if (v > 0) budget = (size_t)v << 20;
  }
  return (int)std::max<size_t>(1, budget / per_state);
}

void fill_common(const qhbm_plan* p, KernelArgs& ka) {
  const HostPlan& hp = p->hp;
  std::memset(&ka, 0, sizeof(ka));
  ka.passes = p->d_passes.p;
  ka.ops = p->d_ops.p;
  ka.coef = p->d_coef.p;
  ka.gdescs = p->d_gdescs.p;
  ka.terms = p->d_terms.p;
  ka.groups = p->d_groups.p;
  ka.opranges = p->d_opranges.p;
  ka.dterms = p->d_dterms.p;
  ka.n_dterms = (int)hp.dterms.size();
  ka.n_groups = (int)hp.groups.size();
  ka.n_terms = (int)hp.terms.size();
  ka.n = hp.n_eff;
  ka.T = hp.T;
  ka.O = hp.O;
  ka.P = hp.P;
  ka.phase_coef = hp.phase_coef;
  // cp.async tile loads: measured 17.38 -> 17.04 ms per 4096 bitstrings on config 3
  // (profiles/r2_tile_copy_experiment.md); QHBM_SYNC_TILE=1 restores the register-staged copy
  static const bool sync_tile = std::getenv("QHBM_SYNC_TILE") != nullptr;
  ka.async_tile = sync_tile ? 0 : 1;
  // pass programs through the bulk-copy engine (profiles/r2_bulk_stage_experiment.md)
  static const bool no_bulk = std::getenv("QHBM_NO_BULK_STAGE") != nullptr;
  ka.bulk_stage = no_bulk ? 0 : 1;
}

// Coefficient jobs + clearing of the call's float64 accumulators in one launch.
// rows > 1: one coefficient table per symbol row (symbols f32[rows, P], tables coef_stride floats apart).
void run_prep(qhbm_plan* p, const float* d_symbols, int mode, cudaStream_t s, double* zero_a = nullptr,
              int64_t n_a = 0, double* zero_b = nullptr, int64_t n_b = 0, int rows = 1, uint32_t coef_stride = 0) {
  const HostPlan& hp = p->hp;
  const int n_jobs = (int)hp.jobs.size();
  const int64_t nz = n_a + n_b;
  const int zero_blocks = nz > 0 ? (int)std::min<int64_t>((nz + kPrepThreads - 1) / kPrepThreads, 148 * 4) : 0;
  if (n_jobs + zero_blocks == 0) return;
  const dim3 grid((unsigned)(n_jobs + zero_blocks), (unsigned)((n_jobs > 0 && rows > 1) ? (rows + kPrepRows - 1) / kPrepRows : 1));
  prep_kernel<<<grid, kPrepThreads, 0, s>>>(p->d_jobs.p, n_jobs, p->d_lists.p, p->d_gates.p, d_symbols, p->d_coef.p, mode,
                                            zero_a, n_a, zero_b, n_b, rows, rows > 1 ? (uint32_t)hp.P : 0u,
                                            rows > 1 ? coef_stride : 0u);
  QHBM_CUDA(cudaGetLastError());
}

// Core driver shared by the forward and adjoint entry points.
// sym_rows: d_symbols is f32[U, P], one row of symbol values per state (the TFQ op's general form); the
// coefficient tables are then built per state, chunk by chunk, and the sweeps read table u of the chunk.
void run_expectation(qhbm_plan* p, const uint64_t* d_basis, int64_t U, const float* d_symbols,
                     const float* d_dgrad, float* d_out, float* d_grad_out, int per_state, int mode,
                     bool adjoint, cudaStream_t s, bool sym_rows = false) {
  const HostPlan& hp = p->hp;
  if (adjoint && !hp.grad) throw std::runtime_error("plan was created without with_gradient");
  if (U < 0) throw std::runtime_error("n_states must be >= 0");
  if (U > (int64_t)1 << 30) throw std::runtime_error("n_states too large");
  const int tiles = hp.tiles();
  const bool multi = tiles > 1;
  const int64_t grows = adjoint ? (per_state ? U : 1) : 0;
  if (U == 0) {
    if (adjoint && !per_state && hp.P > 0) QHBM_CUDA(cudaMemsetAsync(d_grad_out, 0, sizeof(float) * hp.P, s));
    return;
  }
  p->d_eacc.reserve((size_t)U * hp.O);
  if (grows) p->d_gacc.reserve((size_t)std::max<int64_t>(1, grows * hp.P));
  // equal chunks no larger than the workspace allows (one chunk when everything fits)
  // (per-state tables: a chunk holds at most 1 GiB of them and fits the prep grid's y extent)
  const uint32_t coef_stride = sym_rows ? (uint32_t)((std::max(hp.ncoef, 4) + 3) & ~3) : 0u;
  const int64_t max_chunk = sym_rows ? std::min<int64_t>({(int64_t)p->chunk, 65535, std::max<int64_t>(1, ((int64_t)1 << 28) / coef_stride)})
                                     : (int64_t)p->chunk;
  const int64_t n_chunks = (U + max_chunk - 1) / max_chunk;
  const int chunk = (int)((U + n_chunks - 1) / n_chunks);
  if (sym_rows) p->d_coef.reserve((size_t)chunk * coef_stride);
  bool pingpong = false;
  for (const LaunchDesc& L : hp.launches) pingpong = pingpong || (L.flags & LF_PSI_ALT);
  if (multi) {
    p->d_psi.reserve((size_t)chunk << hp.n_eff);
    if (adjoint) p->d_lam.reserve((size_t)chunk << hp.n_eff);
    if (adjoint && pingpong) p->d_psi_alt.reserve((size_t)chunk << hp.n_eff);
  }
  if (!sym_rows)
    run_prep(p, d_symbols, mode, s, p->d_eacc.p, U * hp.O, p->d_gacc.p, (grows && hp.P > 0) ? grows * hp.P : 0);

  KernelArgs ka;
  fill_common(p, ka);
  ka.coef_stride = coef_stride;
  ka.per_state = per_state;
  ka.gacc = p->d_gacc.p;
  // forward-only runs on an adjoint plan stop after the expectation launch
  for (int64_t u0 = 0; u0 < U; u0 += chunk) {
    const int c = (int)std::min<int64_t>(chunk, U - u0);
    if (sym_rows) {  // this chunk's tables (the first chunk's launch also clears the call's accumulators)
      const bool z = u0 == 0;
      run_prep(p, d_symbols + u0 * hp.P, mode, s, z ? p->d_eacc.p : nullptr, z ? U * hp.O : 0, z ? p->d_gacc.p : nullptr,
               (z && grows && hp.P > 0) ? grows * hp.P : 0, c, coef_stride);
    }
    ka.basis = d_basis + u0;
    ka.dgrad = (adjoint && d_dgrad) ? d_dgrad + u0 * hp.O : nullptr;
    ka.eacc = p->d_eacc.p + u0 * hp.O;
    ka.grow0 = (int)u0;
    float2* cur_psi = multi ? p->d_psi.p : nullptr;
    ka.lam = (multi && adjoint) ? p->d_lam.p : nullptr;
    for (size_t li = 0; li < hp.launches.size(); ++li) {
      ka.L = hp.launches[li];
      if (!adjoint) {
        if (ka.L.pass_b_end > ka.L.pass_b_begin && !(ka.L.flags & LF_EXPECT)) break;  // backward sweeps
        ka.L.pass_b_begin = ka.L.pass_b_end = 0;
        ka.L.flags &= ~(uint32_t)(LF_STORE_LAM | LF_PSI_ALT);
        if (ka.L.flags & LF_EXPECT) ka.L.flags &= ~(uint32_t)(LF_STORE_PSI);
      }
      ka.psi = cur_psi;
      ka.psi_out = cur_psi;
      if (multi && (ka.L.flags & LF_PSI_ALT)) {
        ka.psi_out = cur_psi == p->d_psi.p ? p->d_psi_alt.p : p->d_psi.p;
        cur_psi = ka.psi_out;
      }
      launch_any(p, hp.grad, ka, c, s);
    }
  }
  {
    const int64_t ne = U * hp.O, ng = (grows && hp.P > 0) ? grows * hp.P : 0;
    if (ne + ng > 0) {
      finalize_kernel<<<(unsigned)((ne + ng + 255) / 256), 256, 0, s>>>(p->d_eacc.p, d_out, ne, p->d_gacc.p, d_grad_out, ng);
      QHBM_CUDA(cudaGetLastError());
    }
  }
}

// Final states U|basis_u> as dense complex64 [U, 2^n] rows (global phase restored): the forward
// sweeps without the expectation phase, then one write-out launch per chunk.
void run_states(qhbm_plan* p, const uint64_t* d_basis, int64_t U, const float* d_symbols, float2* d_out,
                cudaStream_t s) {
  const HostPlan& hp = p->hp;
  if (U < 0) throw std::runtime_error("n_states must be >= 0");
  if (U == 0) return;
  const bool multi = hp.tiles() > 1;
  const bool padded = hp.n_eff != hp.n;  // small circuits are simulated with idle high qubits
  const size_t dim_eff = (size_t)1 << hp.n_eff, dim = (size_t)1 << hp.n;
  int64_t chunk = multi ? p->chunk : (int64_t)1 << 16;
  if (padded) chunk = std::min<int64_t>(chunk, std::max<int64_t>(1, ((int64_t)1 << 28) >> hp.n_eff));
  chunk = std::min<int64_t>(chunk, U);
  p->d_eacc.reserve(std::max(hp.O, 1));
  if (multi) p->d_psi.reserve((size_t)chunk << hp.n_eff);
  if (padded || (multi && hp.grad)) p->d_lam.reserve((size_t)chunk << hp.n_eff);  // never both
  run_prep(p, d_symbols, QHBM_GRAD_EXACT, s);
  KernelArgs ka;
  fill_common(p, ka);
  ka.eacc = p->d_eacc.p;
  ka.psi = multi ? p->d_psi.p : nullptr;
  ka.psi_out = ka.psi;
  ka.lam = (multi && hp.grad) ? p->d_lam.p : nullptr;
  for (int64_t u0 = 0; u0 < U; u0 += chunk) {
    const int c = (int)std::min<int64_t>(chunk, U - u0);
    ka.basis = d_basis + u0;
    ka.state_out = padded ? p->d_lam.p : d_out + (size_t)u0 * dim;
    for (int li = 0; li < hp.n_fwd_launches; ++li) {
      ka.L = hp.launches[li];
      ka.L.pass_b_begin = ka.L.pass_b_end = 0;
      ka.L.flags &= ~(uint32_t)(LF_EXPECT | LF_STORE_LAM);
      if (!multi) ka.L.flags |= LF_WRITE_STATE;
      launch_any(p, hp.grad, ka, c, s);
    }
    if (multi) {
      // one more launch: load the stored state and write it out with the dropped global phase
      ka.L = hp.launches[hp.n_fwd_launches];  // the expectation launch has the contiguous tile map
      ka.L.pass_a_begin = ka.L.pass_a_end = ka.L.pass_b_begin = ka.L.pass_b_end = 0;
      ka.L.flags = LF_LOAD_PSI | LF_WRITE_STATE;
      launch_any(p, hp.grad, ka, c, s);
    }
    if (padded)
      QHBM_CUDA(cudaMemcpy2DAsync(d_out + (size_t)u0 * dim, dim * sizeof(float2), p->d_lam.p,
                                  dim_eff * sizeof(float2), dim * sizeof(float2), (size_t)c,
                                  cudaMemcpyDeviceToDevice, s));
  }
}

}  // namespace

extern "C" {

const char* qhbm_last_error(void) { return g_last_error.c_str(); }
int qhbm_version(void) { return 100; }

int qhbm_circuit_create(const qhbm_gate_t* gates, int32_t n_gates, int32_t n_qubits, int32_t n_symbols,
                        qhbm_circuit_t** out) {
  return guarded([&] {
    if (!out) throw std::runtime_error("out is null");
    if (n_gates < 0 || (n_gates > 0 && !gates)) throw std::runtime_error("bad gate table");
    auto* c = new qhbm_circuit;
    c->ir.n_qubits = n_qubits;
    c->ir.n_symbols = n_symbols;
    c->ir.gates.assign(gates, gates + n_gates);
    try {
      validate_circuit(c->ir);
    } catch (...) {
      delete c;
      throw;
    }
    *out = c;
  });
}
void qhbm_circuit_destroy(qhbm_circuit_t* c) { delete c; }

int qhbm_ops_create(const qhbm_pauli_term_t* terms, const int32_t* term_offsets, int32_t n_ops,
                    int32_t n_qubits, qhbm_ops_t** out) {
  return guarded([&] {
    if (!out || !term_offsets || n_ops < 0) throw std::runtime_error("bad arguments");
    auto* o = new qhbm_ops;
    o->ir.n_qubits = n_qubits;
    o->ir.offsets.assign(term_offsets, term_offsets + n_ops + 1);
    const int nt = o->ir.offsets.back();
    if (nt < 0 || (nt > 0 && !terms)) { delete o; throw std::runtime_error("bad term table"); }
    o->ir.terms.assign(terms, terms + nt);
    try {
      if (n_qubits < 1 || n_qubits > kMaxQubits) throw std::runtime_error("n_qubits out of range");
      validate_ops(o->ir);
    } catch (...) {
      delete o;
      throw;
    }
    *out = o;
  });
}
void qhbm_ops_destroy(qhbm_ops_t* o) { delete o; }

int qhbm_plan_create(const qhbm_circuit_t* c, const qhbm_ops_t* o, int32_t with_gradient, int32_t tile_qubits,
                     int32_t reg_qubits, qhbm_plan_t** out) {
  return guarded([&] {
    if (!c || !o || !out) throw std::runtime_error("null handle");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw std::runtime_error("no CUDA device: the qhbm_b200 engine has no CPU fallback");
    if (tile_qubits == 0) if (const char* e = std::getenv("QHBM_TILE_QUBITS")) tile_qubits = std::atoi(e);
    if (reg_qubits == 0) if (const char* e = std::getenv("QHBM_REG_QUBITS")) reg_qubits = std::atoi(e);
    auto* p = new qhbm_plan;
    try {
      p->hp = compile_plan(c->ir, o->ir, with_gradient != 0, tile_qubits, reg_qubits);
      const HostPlan& hp = p->hp;
      p->d_passes.upload(hp.dev_passes);
      p->d_ops.upload(hp.dev_ops);
      p->d_gdescs.upload(hp.gdescs);
      p->d_jobs.upload(hp.jobs);
      p->d_lists.upload(hp.lists);
      p->d_gates.upload(hp.gates);
      p->d_terms.upload(hp.terms);
      p->d_groups.upload(hp.groups);
      p->d_opranges.upload(hp.opranges);
      p->d_dterms.upload(hp.dterms);
      p->d_coef.reserve(std::max(hp.ncoef, 4));
      int dev = 0;
      QHBM_CUDA(cudaGetDevice(&dev));
      QHBM_CUDA(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev));
      p->chunk = default_chunk(p);
      allow_large_smem();
    } catch (...) {
      delete p;
      throw;
    }
    *out = p;
  });
}
void qhbm_plan_destroy(qhbm_plan_t* p) { delete p; }

int qhbm_plan_info(const qhbm_plan_t* p, int64_t* out8) {
  return guarded([&] {
    if (!p || !out8) throw std::runtime_error("null argument");
    out8[0] = p->hp.n_sweeps_fwd;
    out8[1] = p->hp.n_sweeps_bwd;
    out8[2] = (int64_t)p->hp.passes.size();
    out8[3] = (int64_t)p->hp.ops.size();
    out8[4] = p->hp.T;
    out8[5] = p->hp.K;
    out8[6] = (int64_t)p->hp.launches.size();
    out8[7] = p->chunk;
  });
}

int qhbm_expectation_forward(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                             const float* d_symbols, float* d_out, void* stream) {
  return guarded([&] {
    if (!p) throw std::runtime_error("null plan");
    PlanUse use(p, (cudaStream_t)stream);
    run_expectation(p, d_basis_idx, n_states, d_symbols, nullptr, d_out, nullptr, 0, QHBM_GRAD_EXACT, false,
                    (cudaStream_t)stream);
  });
}

int qhbm_expectation_adjoint(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                             const float* d_symbols, const float* d_dgrad, float* d_out, float* d_grad_out,
                             int32_t per_state, int32_t grad_mode, void* stream) {
  return guarded([&] {
    if (!p) throw std::runtime_error("null plan");
    if (grad_mode < 0 || grad_mode > 2) throw std::runtime_error("bad grad_mode");
    if (!d_dgrad) throw std::runtime_error("d_dgrad is null");
    PlanUse use(p, (cudaStream_t)stream);
    run_expectation(p, d_basis_idx, n_states, d_symbols, d_dgrad, d_out, d_grad_out, per_state, grad_mode, true,
                    (cudaStream_t)stream);
  });
}

int qhbm_expectation_forward_rows(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                                  const float* d_symbol_rows, float* d_out, void* stream) {
  return guarded([&] {
    if (!p) throw std::runtime_error("null plan");
    PlanUse use(p, (cudaStream_t)stream);
    run_expectation(p, d_basis_idx, n_states, d_symbol_rows, nullptr, d_out, nullptr, 0, QHBM_GRAD_EXACT, false,
                    (cudaStream_t)stream, true);
  });
}

int qhbm_expectation_adjoint_rows(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states,
                                  const float* d_symbol_rows, const float* d_dgrad, float* d_out,
                                  float* d_grad_out, int32_t per_state, int32_t grad_mode, void* stream) {
  return guarded([&] {
    if (!p) throw std::runtime_error("null plan");
    if (grad_mode < 0 || grad_mode > 2) throw std::runtime_error("bad grad_mode");
    if (!d_dgrad) throw std::runtime_error("d_dgrad is null");
    PlanUse use(p, (cudaStream_t)stream);
    run_expectation(p, d_basis_idx, n_states, d_symbol_rows, d_dgrad, d_out, d_grad_out, per_state, grad_mode, true,
                    (cudaStream_t)stream, true);
  });
}

int qhbm_expectation_host(qhbm_plan_t* p, const uint64_t* h_basis_idx, int64_t n_states, const float* h_symbols,
                          const float* h_dgrad, float* h_out, float* h_grad_out, int32_t grad_mode,
                          void* stream) {
  return guarded([&] {
    if (!p) throw std::runtime_error("null plan");
    PlanUse use(p, (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    const HostPlan& hp = p->hp;
    const int64_t U = n_states;
    const bool adjoint = h_dgrad != nullptr && h_grad_out != nullptr;
    p->d_basis.reserve(std::max<int64_t>(U, 1));
    p->d_sym.reserve(std::max(hp.P, 1));
    p->d_out.reserve(std::max<int64_t>(U * hp.O, 1));
    QHBM_CUDA(cudaMemcpyAsync(p->d_basis.p, h_basis_idx, sizeof(uint64_t) * U, cudaMemcpyHostToDevice, s));
    if (hp.P) QHBM_CUDA(cudaMemcpyAsync(p->d_sym.p, h_symbols, sizeof(float) * hp.P, cudaMemcpyHostToDevice, s));
    if (adjoint) {
      p->d_dgrad.reserve(std::max<int64_t>(U * hp.O, 1));
      p->d_gout.reserve(std::max(hp.P, 1));
      QHBM_CUDA(cudaMemcpyAsync(p->d_dgrad.p, h_dgrad, sizeof(float) * U * hp.O, cudaMemcpyHostToDevice, s));
    }
    run_expectation(p, p->d_basis.p, U, p->d_sym.p, adjoint ? p->d_dgrad.p : nullptr, p->d_out.p,
                    adjoint ? p->d_gout.p : nullptr, 0, grad_mode, adjoint, s);
    QHBM_CUDA(cudaMemcpyAsync(h_out, p->d_out.p, sizeof(float) * U * hp.O, cudaMemcpyDeviceToHost, s));
    if (adjoint && hp.P)
      QHBM_CUDA(cudaMemcpyAsync(h_grad_out, p->d_gout.p, sizeof(float) * hp.P, cudaMemcpyDeviceToHost, s));
    QHBM_CUDA(cudaStreamSynchronize(s));
  });
}

int qhbm_final_states(qhbm_plan_t* p, const uint64_t* d_basis_idx, int64_t n_states, const float* d_symbols,
                      float* d_states_out, void* stream) {
  return guarded([&] {
    if (!p) throw std::runtime_error("null plan");
    PlanUse use(p, (cudaStream_t)stream);
    run_states(p, d_basis_idx, n_states, d_symbols, reinterpret_cast<float2*>(d_states_out), (cudaStream_t)stream);
  });
}

int qhbm_debug_state(qhbm_plan_t* p, uint64_t basis_idx, const float* d_symbols, float* d_state_out,
                     void* stream) {
  return guarded([&] {
    if (!p) throw std::runtime_error("null plan");
    PlanUse use(p, (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    p->d_basis.reserve(1);
    QHBM_CUDA(cudaMemcpyAsync(p->d_basis.p, &basis_idx, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    QHBM_CUDA(cudaStreamSynchronize(s));
    run_states(p, p->d_basis.p, 1, d_symbols, reinterpret_cast<float2*>(d_state_out), s);
  });
}

}  // extern "C"
