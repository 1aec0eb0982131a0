// verify_plan.cpp -- TEST-ONLY scalar interpreter of the compiled sweep/pass/op program.
//
// It executes exactly the program that plan.cpp emits (same sweeps, tiles, passes, ops,
// coefficient jobs) with plain loops on the host, so that the scheduler and the op
// semantics can be checked against oracle/ without a GPU.  It is NOT part of the
// product library (libqhbm_b200.so does not contain it) and is never a fallback.
#include <cmath>
#include <complex>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../qhbm-library_b200/csrc/gate_math.h"
#include "../../qhbm-library_b200/csrc/plan.h"

using namespace qhbm;
typedef std::complex<double> cplx;

static std::string g_err;

static uint32_t scatter(uint32_t l, const BitRun* runs, int nr) {
  uint32_t g = 0;
  for (int i = 0; i < nr; ++i) g |= ((l >> runs[i].local_start) & ((1u << runs[i].len) - 1u)) << runs[i].global_start;
  return g;
}

static void prep(const HostPlan& hp, const float* symbols, int mode, std::vector<float>& coef) {
  coef.assign(hp.ncoef + 4, 0.f);
  auto wr = [&](int off, int i, cd v) { coef[off + 2 * i] = (float)v.re; coef[off + 2 * i + 1] = (float)v.im; };
  for (const PrepJob& job : hp.jobs) {
    const int32_t* list = hp.lists.data() + job.list_off;
    cd m[16], t[16], w[16];
    switch (job.kind) {
      case PJ_MAT1: {
        cd acc[4] = {mk(1, 0), mk(0, 0), mk(0, 0), mk(1, 0)};
        for (int i = 0; i < job.list_len; ++i) {
          gate_matrix_of(hp.gates[list[i]], symbols, m);
          matmul(m, acc, 2, t);
          for (int k = 0; k < 4; ++k) acc[k] = t[k];
        }
        if (job.a) { dagger(acc, 2, t); for (int k = 0; k < 4; ++k) acc[k] = t[k]; }
        for (int k = 0; k < 4; ++k) wr(job.out, k, acc[k]);
      } break;
      case PJ_MAT2: {
        gate_matrix_of(hp.gates[list[0]], symbols, m);
        if (job.b) { swap_qubits(m, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
        if (job.a) { dagger(m, 4, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
        for (int k = 0; k < 16; ++k) wr(job.out, k, m[k]);
      } break;
      case PJ_ROT: {
        if (job.list_len == 0) { coef[job.out] = 1.f; coef[job.out + 1] = 0.f; coef[job.out + 2] = 0.f; break; }
        double pv[3];
        gate_param_values(hp.gates[list[0]], symbols, pv);
        const cd e = expipi(0.5 * pv[0]);
        coef[job.out] = (float)e.re;
        coef[job.out + 1] = (float)(job.a ? -e.im : e.im);
      } break;
      case PJ_NONE: break;
      case PJ_CONST:
        for (int i = 0; i < job.list_len; ++i) std::memcpy(&coef[job.out + i], &list[i], sizeof(float));
        break;
      case PJ_ROTF: {
        const int K = job.d;
        double c[8], sn[8], kap[8];
        bool fast = true;
        for (int P = 0; P < K; ++P) {
          c[P] = 1.0; sn[P] = 0.0; kap[P] = 0.0;
          if (list[P] < 0) continue;
          const qhbm_gate_t g = hp.gates[list[P]];
          double pv[3];
          gate_param_values(g, symbols, pv);
          const cd e = expipi(0.5 * pv[0]);
          c[P] = e.re; sn[P] = job.a ? -e.im : e.im;
          if (job.b & (1 << P)) {
            const int dim = gate_matrix_of(g, symbols, m);
            gate_derivative(g, symbols, 0, mode, t);
            dagger(m, dim, w);
            matmul(t, w, dim, m);
            kap[P] = -2.0 * m[1].im;
          }
          if (P < K - 1 && fabs(c[P]) < 0.05) fast = false;
        }
        double scale = 1.0;
        for (int P = 0; P < K; ++P) {
          float* o = &coef[job.out + 4 * P];
          if (fast && P < K - 1) { o[0] = (float)(sn[P] / c[P]); o[1] = 0.f; o[2] = (float)(kap[P] * scale * scale); scale *= c[P]; }
          else if (fast) { o[0] = (float)(c[P] * scale); o[1] = (float)(sn[P] * scale); o[2] = (float)(kap[P] * scale * scale); }
          else { o[0] = (float)c[P]; o[1] = (float)sn[P]; o[2] = (float)kap[P]; }
          o[3] = fast ? 1.f : 0.f;
        }
      } break;
      case PJ_PHASE: {
        cd acc = mk(1, 0);
        for (int i = 0; i < job.list_len; ++i) {
          const qhbm_gate_t g = hp.gates[list[i]];
          double pv[3];
          gate_param_values(g, symbols, pv);
          acc = acc * expipi(pv[0] * ((double)g.gshift + 0.5));
        }
        wr(job.out, 0, acc);
      } break;
      case PJ_GRAD1: case PJ_GRAD2: case PJ_GDIAG: case PJ_KAPPA: {
        const qhbm_gate_t g = hp.gates[list[0]];
        const int dim = gate_matrix_of(g, symbols, m);
        gate_derivative(g, symbols, job.c, mode, t);
        dagger(m, dim, w);
        matmul(t, w, dim, m);
        if (dim == 4 && job.b && job.kind != PJ_KAPPA) { swap_qubits(m, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
        if (job.kind == PJ_GDIAG) for (int k = 0; k < 4; ++k) wr(job.out, k, k < dim ? m[k * dim + k] : mk(0, 0));
        else if (job.kind == PJ_KAPPA) coef[job.out] = (float)(job.b == 0 ? -2.0 * m[1].im : -2.0 * m[1].re);
        else for (int k = 0; k < dim * dim; ++k) wr(job.out, k, m[k]);
      } break;
      case PJ_DPAIR: {
        const int dim = gate_matrix_of(hp.gates[list[0]], symbols, m);
        if (dim == 4 && job.b) { swap_qubits(m, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
        for (int k = 0; k < 4; ++k) { cd v = k < dim ? m[k * dim + k] : mk(1, 0); wr(job.out, k, job.a ? conj(v) : v); }
      } break;
      case PJ_DTAB: {
        const int entries = 1 << job.d;
        for (int v = 0; v < entries; ++v) {
          cd acc = mk(1, 0);
          for (int q = 0; q < job.list_len / 3; ++q) {
            const int dim = gate_matrix_of(hp.gates[list[3 * q]], symbols, m);
            int sel = (v >> list[3 * q + 1]) & 1;
            if (list[3 * q + 2] >= 0) sel = 2 * sel + ((v >> list[3 * q + 2]) & 1);
            cd d = m[sel * dim + sel];
            acc = acc * (job.a ? conj(d) : d);
          }
          if (job.b) {  // register phase table: 4 floats per entry (re, im, -im, im)
            coef[job.out + 4 * v + 0] = (float)acc.re;
            coef[job.out + 4 * v + 1] = (float)acc.im;
            coef[job.out + 4 * v + 2] = -(float)acc.im;
            coef[job.out + 4 * v + 3] = (float)acc.im;
          } else {
            wr(job.out, v, acc);
          }
        }
      } break;
    }
  }
}

struct Ctx {
  const HostPlan* hp;
  std::vector<float> coef;
  std::vector<cplx> psi, lam;  // global state
  std::vector<double> eacc, gacc;
  std::vector<double> gwin = std::vector<double>(kGaccFloats, 0.0);  // reduced sums of the current flush window
  const float* dgrad;
};

static cplx cf(const Ctx& c, int off, int i) { return cplx(c.coef[off + 2 * i], c.coef[off + 2 * i + 1]); }

// Unpacks a device op (PackedOp) back into the DevOp fields the interpreter reads.
static DevOp unpack_op(const PackedOp& q) {
  DevOp o;
  std::memset(&o, 0, sizeof(o));
  o.type = (int32_t)(q.w0 & 0xffu);
  o.p0 = (int32_t)((q.w0 >> 8) & 0xffu);
  o.p1 = (int32_t)((q.w0 >> 16) & 0xffu);
  o.gslot = (int32_t)(q.w0 >> 24);
  if (o.p0 == 255) o.p0 = -1;
  if (o.p1 == 255) o.p1 = -1;
  o.coef = q.coef;
  o.aux0 = q.aux0;
  o.aux1 = q.aux1;
  return o;
}

// Runs DEVICE pass `pi` (HostPlan::dev_passes / dev_ops) the way the kernel does, including the gradient
// machinery: per-thread scratch units, the reduction tasks behind the pass's ops, and the descriptors
// of a flush window.
static void run_pass(Ctx& c, const LaunchDesc& L, int pi, std::vector<cplx>& tp, std::vector<cplx>& tl,
                     uint32_t goff, bool both) {
  const HostPlan& hp = *c.hp;
  const DevPass& ps = hp.dev_passes[pi];
  const int K = hp.K, R = 1 << K, nthr = 1 << (hp.T - K);
  std::vector<double> scratch((size_t)4 * R * nthr, 0.0);  // 4 * 2^K units of one float per thread
  for (int tid = 0; tid < nthr; ++tid) {
    uint32_t base = tid;
    for (int j = 0; j < K; ++j) { int sp = ps.sorted[j]; base = ((base >> sp) << (sp + 1)) | (base & ((1u << sp) - 1u)); }
    const uint32_t gbase = goff | scatter(base, L.runs, L.n_runs);
    std::vector<uint32_t> idx(R);
    for (int r = 0; r < R; ++r) { uint32_t dep = 0; for (int j = 0; j < K; ++j) if ((r >> j) & 1) dep |= 1u << ps.regbit[j]; idx[r] = base | dep; }
    std::vector<cplx> a(R), b(R);
    for (int r = 0; r < R; ++r) { a[r] = tp[idx[r]]; if (both) b[r] = tl[idx[r]]; }
    cplx F(1, 0);
    auto mat1 = [&](std::vector<cplx>& v, int p, int off) {
      for (int r = 0; r < R; ++r) if (!(r & (1 << p))) {
        cplx x0 = v[r], x1 = v[r | (1 << p)];
        v[r] = cf(c, off, 0) * x0 + cf(c, off, 1) * x1;
        v[r | (1 << p)] = cf(c, off, 2) * x0 + cf(c, off, 3) * x1;
      }
    };
    auto mat2 = [&](std::vector<cplx>& v, int lo, int off, const std::vector<cplx>* bb, double* gout) {
      double s = 0;
      for (int r = 0; r < R; ++r) if (!(r & (3 << lo))) {
        cplx x[4], y[4];
        for (int j = 0; j < 4; ++j) x[j] = v[r | (j << lo)];
        for (int i = 0; i < 4; ++i) { y[i] = 0; for (int j = 0; j < 4; ++j) y[i] += cf(c, off, 4 * i + j) * x[j]; }
        if (bb) for (int i = 0; i < 4; ++i) s += (std::conj((*bb)[r | (i << lo)]) * y[i]).real();
        else for (int i = 0; i < 4; ++i) v[r | (i << lo)] = y[i];
      }
      if (gout) *gout = 2 * s;
    };
    for (int oi = ps.op_begin; oi < ps.exec_end; ++oi) {
      const DevOp op = unpack_op(hp.dev_ops[oi]);
      switch (op.type) {
        case OP_MAT1: mat1(a, op.p0, op.coef); if (both) mat1(b, op.p0, op.coef); break;
        case OP_XROT: case OP_YROT: {
          const double cc = c.coef[op.coef], ss = c.coef[op.coef + 1];
          const cplx m01 = op.type == OP_XROT ? cplx(0, -ss) : cplx(-ss, 0);
          const cplx m10 = op.type == OP_XROT ? cplx(0, -ss) : cplx(ss, 0);
          if (both && op.gslot >= 0) {
            double sacc = 0;
            for (int r = 0; r < R; ++r) if (!(r & (1 << op.p0))) {
              const int q = r | (1 << op.p0);
              cplx y0, y1;
              if (op.type == OP_XROT) { y0 = a[q]; y1 = a[r]; }
              else { y0 = cplx(0, -1) * a[q]; y1 = cplx(0, 1) * a[r]; }
              sacc += (std::conj(b[r]) * y0 + std::conj(b[q]) * y1).imag();
            }
            scratch[(size_t)op.gslot * nthr + tid] = c.coef[op.coef + 2] * sacc;
          }
          for (int which = 0; which < (both ? 2 : 1); ++which) {
            std::vector<cplx>& v = which ? b : a;
            for (int r = 0; r < R; ++r) if (!(r & (1 << op.p0))) {
              cplx x0 = v[r], x1 = v[r | (1 << op.p0)];
              v[r] = cc * x0 + m01 * x1;
              v[r | (1 << op.p0)] = m10 * x0 + cc * x1;
            }
          }
        } break;
        case OP_XROTM: case OP_YROTM: case OP_XROTF: {
          for (int P = 0; P < K; ++P) if (((op.p0 | (both ? op.aux0 : 0)) & (1 << P)) || op.type == OP_XROTF) {
            const bool rotate = (op.p0 & (1 << P)) || op.type == OP_XROTF;  // else: gradient only
            double cc = c.coef[op.coef + 4 * P], ss = c.coef[op.coef + 4 * P + 1];
            const double kap = c.coef[op.coef + 4 * P + 2];
            if (op.type == OP_XROTF && c.coef[op.coef + 3] != 0.f && P < K - 1) { ss = cc; cc = 1.0; }  // (I - i t X)
            const bool isx = op.type != OP_YROTM;
            const cplx m01 = isx ? cplx(0, -ss) : cplx(-ss, 0);
            const cplx m10 = isx ? cplx(0, -ss) : cplx(ss, 0);
            if (both && (op.aux0 & (1 << P))) {
              double sacc = 0;
              for (int r = 0; r < R; ++r) if (!(r & (1 << P))) {
                const int q = r | (1 << P);
                cplx y0, y1;
                if (isx) { y0 = a[q]; y1 = a[r]; } else { y0 = cplx(0, -1) * a[q]; y1 = cplx(0, 1) * a[r]; }
                sacc += (std::conj(b[r]) * y0 + std::conj(b[q]) * y1).imag();
              }
              scratch[(size_t)(P < 4 ? ((op.aux1 >> (8 * P)) & 0xff) : op.p1) * nthr + tid] = kap * sacc;
            }
            for (int which = 0; rotate && which < (both ? 2 : 1); ++which) {
              std::vector<cplx>& v = which ? b : a;
              for (int r = 0; r < R; ++r) if (!(r & (1 << P))) {
                cplx x0 = v[r], x1 = v[r | (1 << P)];
                v[r] = cc * x0 + m01 * x1;
                v[r | (1 << P)] = m10 * x0 + cc * x1;
              }
            }
          }
        } break;
        case OP_GRAD_X: case OP_GRAD_Y: {
          double sacc = 0;
          for (int r = 0; r < R; ++r) if (!(r & (1 << op.p0))) {
            const int q = r | (1 << op.p0);
            cplx y0, y1;  // (A a) on the pair
            if (op.type == OP_GRAD_X) { y0 = a[q]; y1 = a[r]; }
            else { y0 = cplx(0, -1) * a[q]; y1 = cplx(0, 1) * a[r]; }
            sacc += (std::conj(b[r]) * y0 + std::conj(b[q]) * y1).imag();
          }
          scratch[(size_t)op.gslot * nthr + tid] = c.coef[op.coef] * sacc;
        } break;
        case OP_MAT2: mat2(a, op.p0 ? 2 : 0, op.coef, nullptr, nullptr); if (both) mat2(b, op.p0 ? 2 : 0, op.coef, nullptr, nullptr); break;
        case OP_DCONST_TAB: F *= cf(c, op.coef, (gbase >> op.aux0) & op.aux1); break;
        case OP_DCONST_PAIR: { int sel = (gbase >> op.aux0) & 1; if (op.aux1 >= 0) sel = 2 * sel + ((gbase >> op.aux1) & 1); F *= cf(c, op.coef, sel); } break;
        case OP_DREG_TAB: for (int r = 0; r < R; ++r) { cplx k = F * cf(c, op.coef, 2 * r); a[r] *= k; if (both) b[r] *= k; } F = 1; break;
        case OP_DAPPLY: for (int r = 0; r < R; ++r) { a[r] *= F; if (both) b[r] *= F; } F = 1; break;
        case OP_DCROSS: { int cb = (gbase >> op.aux0) & 1; for (int r = 0; r < R; ++r) { cplx k = cf(c, op.coef, 2 * cb + ((r >> op.p0) & 1)); a[r] *= k; if (both) b[r] *= k; } } break;
        case OP_GRAD_MAT1: {
          double s = 0;
          for (int r = 0; r < R; ++r) if (!(r & (1 << op.p0))) {
            cplx x0 = a[r], x1 = a[r | (1 << op.p0)];
            s += (std::conj(b[r]) * (cf(c, op.coef, 0) * x0 + cf(c, op.coef, 1) * x1)).real();
            s += (std::conj(b[r | (1 << op.p0)]) * (cf(c, op.coef, 2) * x0 + cf(c, op.coef, 3) * x1)).real();
          }
          scratch[(size_t)op.gslot * nthr + tid] = 2 * s;
        } break;
        case OP_GRAD_MAT2: { double g = 0; mat2(a, op.p0 ? 2 : 0, op.coef, &b, &g); scratch[(size_t)op.gslot * nthr + tid] = g; } break;
        case OP_GD_BEGIN: {
          // marginal vectors of w = conj(b) a = u + i v, stored as (u, -v): T, S[p], SS[pi] in mask order
          const uint32_t vmask = (uint32_t)op.coef;
          int vec = 0;
          auto store = [&](cplx w) {
            const size_t at = (size_t)(op.gslot + 2 * vec) * nthr + 2 * (size_t)tid;
            scratch[at] = w.real();
            scratch[at + 1] = -w.imag();
            ++vec;
          };
          cplx T = 0;
          for (int r = 0; r < R; ++r) T += std::conj(b[r]) * a[r];
          store(T);
          for (int p = 0; p < K; ++p) {
            if (!((vmask >> (1 + p)) & 1)) continue;
            cplx S = 0;
            for (int r = 0; r < R; ++r) if ((r >> p) & 1) S += std::conj(b[r]) * a[r];
            store(S);
          }
          for (int ph = 1; ph < K; ++ph) for (int pl = 0; pl < ph; ++pl) {
            if (!((vmask >> (1 + K + ph * (ph - 1) / 2 + pl)) & 1)) continue;
            cplx S = 0;
            for (int r = 0; r < R; ++r) if (((r >> ph) & 1) && ((r >> pl) & 1)) S += std::conj(b[r]) * a[r];
            store(S);
          }
          oi += op.aux0;  // the run's descriptors are host-side information only
        } break;
        default: throw std::runtime_error("verify: unknown op");
      }
    }
    for (int r = 0; r < R; ++r) { tp[idx[r]] = a[r]; if (both) tl[idx[r]] = b[r]; }
  }
  if (!both) return;
  for (int t = ps.exec_end; t < ps.op_end; ++t) {  // reduction tasks
    const DevOp op = unpack_op(hp.dev_ops[t]);
    if (op.type == OP_TASK_F) {
      double sum = 0;
      for (int i = 0; i < nthr; ++i) sum += scratch[(size_t)op.p0 * nthr + i];
      c.gwin.at(op.coef) = sum;
    } else if (op.type == OP_TASK_C) {
      double sx = 0, sy = 0;
      for (int i = 0; i < nthr; ++i) {
        if (((uint32_t)i & (uint32_t)op.aux0) != (uint32_t)op.aux0) continue;
        sx += scratch[(size_t)op.p0 * nthr + 2 * (size_t)i];
        sy += scratch[(size_t)op.p0 * nthr + 2 * (size_t)i + 1];
      }
      c.gwin.at(op.coef) = sx;
      c.gwin.at(op.coef + 1) = sy;
    } else {
      throw std::runtime_error("verify: unexpected op among the reduction tasks");
    }
  }
  for (int g = ps.gd_flush_begin; g < ps.gd_flush_end; ++g) {  // end of a flush window
    const DevGradDesc& d = hp.gdescs.at(g);
    double val;
    if (d.kind == 0) {
      val = c.gwin.at(d.i_tot);
    } else {
      const bool ok_a = d.cond_a < 0 || ((goff >> d.cond_a) & 1u), ok_b = d.cond_b < 0 || ((goff >> d.cond_b) & 1u);
      auto G2 = [&](int i) { return std::pair<double, double>(c.gwin.at(i), c.gwin.at(i + 1)); };
      const auto tot = G2(d.i_tot);
      const auto A = ok_a ? G2(d.i_a) : std::pair<double, double>(0, 0);
      const float* m = &c.coef[d.coef];
      if (d.kind == 1) {
        val = 2 * (m[0] * (tot.first - A.first) + m[1] * (tot.second - A.second) + m[2] * A.first + m[3] * A.second);
      } else {
        const auto B = ok_b ? G2(d.i_b) : std::pair<double, double>(0, 0);
        const auto AB = (ok_a && ok_b) ? G2(d.i_ab) : std::pair<double, double>(0, 0);
        const double u10x = A.first - AB.first, u10y = A.second - AB.second;
        const double u01x = B.first - AB.first, u01y = B.second - AB.second;
        const double u00x = tot.first - A.first - u01x, u00y = tot.second - A.second - u01y;
        val = 2 * (m[0] * u00x + m[1] * u00y + m[2] * u01x + m[3] * u01y + m[4] * u10x + m[5] * u10y +
                   m[6] * AB.first + m[7] * AB.second);
      }
    }
    c.gacc[d.sym] += val;
  }
}

// Observable pass (OP_HX / OP_HD): th += H_part tp for the strings this pass owns.
static void run_hpass(Ctx& c, const LaunchDesc& L, const DevPass& ps, const std::vector<cplx>& tp,
                      std::vector<cplx>& th, uint32_t goff) {
  const HostPlan& hp = *c.hp;
  const int K = hp.K, R = 1 << K, nthr = 1 << (hp.T - K);
  for (int tid = 0; tid < nthr; ++tid) {
    uint32_t base = tid;
    for (int j = 0; j < K; ++j) { int sp = ps.sorted[j]; base = ((base >> sp) << (sp + 1)) | (base & ((1u << sp) - 1u)); }
    const uint32_t gbase = goff | scatter(base, L.runs, L.n_runs);
    std::vector<uint32_t> idx(R);
    for (int r = 0; r < R; ++r) { uint32_t dep = 0; for (int j = 0; j < K; ++j) if ((r >> j) & 1) dep |= 1u << ps.regbit[j]; idx[r] = base | dep; }
    for (int oi = ps.op_begin; oi < ps.op_end; ++oi) {
      const DevOp op = unpack_op(hp.dev_ops[oi]);
      if (op.type != OP_HX && op.type != OP_HD) throw std::runtime_error("unexpected op in an observable pass");
      const double sg = (__builtin_popcount(gbase & (uint32_t)op.aux0) & 1) ? -1.0 : 1.0;
      const int xr = op.type == OP_HD ? 0 : op.p0;
      for (int r = 0; r < R; ++r) th[idx[r]] += sg * (double)c.coef[op.coef + r] * tp[idx[r ^ xr]];
    }
  }
}

static void run_launch(Ctx& c, LaunchDesc L, uint32_t basis, bool adjoint) {
  const HostPlan& hp = *c.hp;
  const int tiles = hp.tiles(), tsz = 1 << hp.T;
  std::vector<std::vector<cplx>> new_psi, new_lam;
  // process all tiles reading the OLD global arrays for cross-tile gathers (launch semantics)
  std::vector<cplx> psi_next = c.psi, lam_next = c.lam;
  for (int tile = 0; tile < tiles; ++tile) {
    const uint32_t goff = scatter(tile, L.oruns, L.n_oruns);
    std::vector<cplx> tp(tsz, 0), tl(tsz, 0);
    bool active = true;
    if (L.flags & LF_INIT_BASIS) {
      active = (basis & ~L.tile_mask) == goff;
      if (!active && (L.flags & LF_SPARSE_OUT)) continue;  // the CTA exits: nothing is stored for this tile
      for (int l = 0; l < tsz; ++l) if (active && (goff | scatter(l, L.runs, L.n_runs)) == basis) tp[l] = 1;
    } else if (L.flags & LF_LOAD_PSI) {
      for (int l = 0; l < tsz; ++l) {
        const uint32_t gi = goff | scatter(l, L.runs, L.n_runs);
        const bool zero = (L.flags & LF_SPARSE_IN) && ((gi ^ (uint32_t)basis) & L.sparse_mask);
        tp[l] = zero ? cplx(0, 0) : c.psi[gi];
      }
    }
    if (L.flags & LF_LOAD_LAM) for (int l = 0; l < tsz; ++l) tl[l] = c.lam[goff | scatter(l, L.runs, L.n_runs)];
    if (active) for (int p = L.pass_a_begin; p < L.pass_a_end; ++p) run_pass(c, L, p, tp, tl, goff, false);
    if (L.flags & LF_EXPECT) {
      for (const DevDiagTerm& d : hp.dterms) {  // WHT-path terms: same mathematics, evaluated directly here
        if (L.expect_stage != 0) break;
        const double gj = (adjoint && c.dgrad) ? c.dgrad[d.op] : 0.0;
        for (int l = 0; l < tsz; ++l) {
          const uint32_t gi = goff | scatter(l, L.runs, L.n_runs);
          const double sg = (__builtin_popcount(gi & d.z) & 1) ? -1.0 : 1.0;
          c.eacc[d.op] += sg * d.coeff * std::norm(tp[l]);
          tl[l] += gj * sg * d.coeff * tp[l];
        }
      }
      if (L.pass_h_end > L.pass_h_begin) {  // single observable: its in-register strings
        std::vector<cplx> th(tsz, 0);
        for (int p = L.pass_h_begin; p < L.pass_h_end; ++p) run_hpass(c, L, hp.dev_passes[p], tp, th, goff);
        const double g0 = (adjoint && c.dgrad) ? c.dgrad[0] : 0.0;
        for (int l = 0; l < tsz; ++l) {
          c.eacc[0] += (std::conj(tp[l]) * th[l]).real();
          tl[l] += g0 * th[l];
        }
      }
      for (int ri = L.rng_begin; ri < L.rng_end; ++ri) {
        const DevOpRange& orng = hp.opranges[ri];
        const int j = orng.op;
        const double gj = (adjoint && c.dgrad) ? c.dgrad[j] : 0.0;
        double ej = 0;
        for (int g = orng.group_begin; g < orng.group_end; ++g) {
          if (g < L.grp_begin || g >= L.grp_end) throw std::runtime_error("group outside the launch's stage slice");
          const DevTermGroup& G = hp.groups[g];
          for (int l = 0; l < tsz; ++l) {
            const uint32_t gi = goff | scatter(l, L.runs, L.n_runs);
            cplx k(G.k0r, G.k0i);
            for (int t = G.term_begin; t < G.term_end; ++t) {
              const DevTerm& T = hp.terms[t];
              const double sg = (__builtin_popcount(gi & T.z) & 1) ? -1.0 : 1.0;
              k += sg * cplx(T.kr, T.ki);
            }
            const cplx pv = G.xl >= 0 ? tp[l ^ (uint32_t)G.xl] : c.psi[gi ^ G.x];
            const cplx h = k * pv;
            ej += (std::conj(tp[l]) * h).real();
            tl[l] += gj * h;
          }
        }
        c.eacc[j] += ej;
      }
    }
    for (int p = L.pass_b_begin; p < L.pass_b_end; ++p) run_pass(c, L, p, tp, tl, goff, true);
    for (int l = 0; l < tsz; ++l) {
      const uint32_t gi = goff | scatter(l, L.runs, L.n_runs);
      if ((L.flags & LF_STORE_PSI) || tiles == 1) psi_next[gi] = tp[l];
      if ((L.flags & LF_STORE_LAM) || tiles == 1) lam_next[gi] = tl[l];
    }
  }
  c.psi.swap(psi_next);
  c.lam.swap(lam_next);
}

extern "C" {

const char* verify_last_error() { return g_err.c_str(); }

// Runs the compiled program for ONE basis state on the host.  state_out (2*2^n_eff doubles)
// receives U|basis> (state after the forward launches).  info_out[8] as qhbm_plan_info.
int verify_run(const qhbm_gate_t* gates, int n_gates, int n, int P, const qhbm_pauli_term_t* terms,
               const int32_t* offs, int O, int with_grad, int T, int K, int mode, const float* symbols,
               uint64_t basis, const float* dgrad, double* e_out, double* g_out, double* state_out,
               int64_t* info_out) {
  try {
    CircuitIR c; c.n_qubits = n; c.n_symbols = P; c.gates.assign(gates, gates + n_gates);
    OpsIR o; o.n_qubits = n; o.offsets.assign(offs, offs + O + 1); o.terms.assign(terms, terms + offs[O]);
    HostPlan hp = compile_plan(c, o, with_grad != 0, T, K);
    Ctx ctx; ctx.hp = &hp; ctx.dgrad = dgrad;
    prep(hp, symbols, mode, ctx.coef);
    // (the global state starts as NaN: a sweep that reads what an LF_SPARSE_OUT sweep did not store shows up)
    ctx.psi.assign((size_t)1 << hp.n_eff, hp.tiles() > 1 ? cplx(NAN, NAN) : cplx(0, 0));
    ctx.lam.assign((size_t)1 << hp.n_eff, 0);
    ctx.eacc.assign(O, 0); ctx.gacc.assign(std::max(P, 1), 0);
    for (size_t li = 0; li < hp.launches.size(); ++li) {
      LaunchDesc L = hp.launches[li];
      if (hp.tiles() == 1 && state_out) {
        // single launch: capture the forward state by running the forward part alone first
        LaunchDesc F = L; F.flags &= ~(uint32_t)LF_EXPECT; F.pass_b_begin = F.pass_b_end = 0;
        Ctx tmp = ctx; run_launch(tmp, F, (uint32_t)basis, false);
        const cplx gph = hp.phase_coef >= 0 ? cplx(ctx.coef[hp.phase_coef], ctx.coef[hp.phase_coef + 1]) : cplx(1, 0);
        for (size_t i = 0; i < tmp.psi.size(); ++i) { const cplx v = tmp.psi[i] * gph; state_out[2 * i] = v.real(); state_out[2 * i + 1] = v.imag(); }
      }
      run_launch(ctx, L, (uint32_t)basis, with_grad != 0);
      if (hp.tiles() > 1 && state_out && (int)li == hp.n_fwd_launches - 1)
      {
        const cplx gph = hp.phase_coef >= 0 ? cplx(ctx.coef[hp.phase_coef], ctx.coef[hp.phase_coef + 1]) : cplx(1, 0);
        for (size_t i = 0; i < ctx.psi.size(); ++i) { const cplx v = ctx.psi[i] * gph; state_out[2 * i] = v.real(); state_out[2 * i + 1] = v.imag(); }
      }
    }
    for (int j = 0; j < O; ++j) e_out[j] = ctx.eacc[j];
    if (with_grad) for (int s = 0; s < P; ++s) g_out[s] = ctx.gacc[s];
    if (info_out) {
      info_out[0] = hp.n_sweeps_fwd; info_out[1] = hp.n_sweeps_bwd; info_out[2] = hp.passes.size();
      info_out[3] = hp.ops.size(); info_out[4] = hp.T; info_out[5] = hp.K; info_out[6] = hp.launches.size(); info_out[7] = hp.n_eff;
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
}

#include <cstdio>
extern "C" int verify_dump(const qhbm_gate_t* gates, int n_gates, int n, int P, const qhbm_pauli_term_t* terms,
                           const int32_t* offs, int O, int with_grad, int T, int K) {
  try {
    CircuitIR c; c.n_qubits = n; c.n_symbols = P; c.gates.assign(gates, gates + n_gates);
    OpsIR o; o.n_qubits = n; o.offsets.assign(offs, offs + O + 1); o.terms.assign(terms, terms + offs[O]);
    HostPlan hp = compile_plan(c, o, with_grad != 0, T, K);
    static const char* names[] = {"NOP", "MAT1", "MAT2", "DCONST_TAB", "DCONST_PAIR", "DREG_TAB", "DAPPLY", "DCROSS",
                                  "XROT", "YROT", "GRAD_MAT1", "GRAD_MAT2", "XROTM", "YROTM", "XROTF", "GRAD_X", "GRAD_Y", "GD_BEGIN",
                                  "GD_CONST", "GD_REG1", "GD_REG2", "GD_MIX", "HX", "HD"};
    printf("n_eff=%d T=%d K=%d ncoef=%d jobs=%zu terms=%zu groups=%zu\n", hp.n_eff, hp.T, hp.K, hp.ncoef, hp.jobs.size(),
           hp.terms.size(), hp.groups.size());
    for (size_t li = 0; li < hp.launches.size(); ++li) {
      const LaunchDesc& L = hp.launches[li];
      printf("launch %zu flags=0x%x tile_mask=0x%x passA=[%d,%d) passH=[%d,%d) passB=[%d,%d)\n", li, L.flags, L.tile_mask,
             L.pass_a_begin, L.pass_a_end, L.pass_h_begin, L.pass_h_end, L.pass_b_begin, L.pass_b_end);
      for (int pass = 0; pass < 3; ++pass) {
        int b = pass == 2 ? L.pass_b_begin : (pass ? L.pass_h_begin : L.pass_a_begin);
        int e = pass == 2 ? L.pass_b_end : (pass ? L.pass_h_end : L.pass_a_end);
        for (int p = b; p < e; ++p) {
          const DevPass& ps = hp.passes[p];
          int hist[24] = {0};
          for (int oi = ps.op_begin; oi < ps.op_end; ++oi) hist[hp.ops[oi].type]++;
          printf("   pass %d regbits=[", p);
          for (int j = 0; j < hp.K; ++j) printf("%d ", ps.regbit[j]);
          printf("] ngrad=%d ops:", ps.ngrad);
          for (int t = 0; t < 24; ++t) if (hist[t]) printf(" %s=%d", names[t], hist[t]);
          printf("\n");
        }
      }
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
