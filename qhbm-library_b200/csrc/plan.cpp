// plan.cpp -- compiles a gate table + Pauli sums into the sweep/pass/op program.
//
// What the reference does instead: qhbmlib builds one serialized circuit proto per
// unique bitstring (qhbmlib/models/circuit.py:129-136) and TFQ's C++ op parses,
// resolves and fuses it again for every row (TfqSimulateExpectation / TfqAdjointGradient,
// reached from qhbmlib/inference/qnn.py:134-138).  Here the circuit is compiled ONCE
// per (circuit, observables) into a schedule that is independent of the symbol
// values and of the bitstrings; only a small coefficient buffer is recomputed per call.
#include "plan.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <utility>

#include "gate_math.h"

namespace qhbm {

static std::string gate_err(int i, const char* what) {
  return "gate " + std::to_string(i) + ": " + what;
}

void validate_circuit(const CircuitIR& c) {
  if (c.n_qubits < 1 || c.n_qubits > kMaxQubits)
    throw std::runtime_error("n_qubits must be in [1, " + std::to_string(kMaxQubits) + "]");
  if (c.n_symbols < 0) throw std::runtime_error("n_symbols must be >= 0");
  for (size_t i = 0; i < c.gates.size(); ++i) {
    const qhbm_gate_t& g = c.gates[i];
    if (g.type < 0 || g.type >= QHBM_GATE_NUM_TYPES) throw std::runtime_error(gate_err(i, "unknown type"));
    if (g.q0 < 0 || g.q0 >= c.n_qubits) throw std::runtime_error(gate_err(i, "q0 out of range"));
    if (gate_is_two_qubit(g.type)) {
      if (g.q1 < 0 || g.q1 >= c.n_qubits) throw std::runtime_error(gate_err(i, "q1 out of range"));
      if (g.q1 == g.q0) throw std::runtime_error(gate_err(i, "q0 == q1"));
    }
    if (g.nparams != gate_num_params(g.type)) throw std::runtime_error(gate_err(i, "wrong nparams for type"));
    for (int k = 0; k < g.nparams; ++k)
      if (g.sym[k] >= c.n_symbols) throw std::runtime_error(gate_err(i, "symbol index out of range"));
  }
}

void validate_ops(const OpsIR& o) {
  if (o.offsets.empty() || o.offsets[0] != 0) throw std::runtime_error("term_offsets must start at 0");
  for (size_t j = 0; j + 1 < o.offsets.size(); ++j)
    if (o.offsets[j + 1] < o.offsets[j]) throw std::runtime_error("term_offsets must be non-decreasing");
  if (o.offsets.back() != (int)o.terms.size()) throw std::runtime_error("term_offsets/terms size mismatch");
  const uint32_t lim = o.n_qubits >= 32 ? 0xffffffffu : ((1u << o.n_qubits) - 1u);
  for (const auto& t : o.terms)
    if ((t.xmask & ~lim) || (t.zmask & ~lim)) throw std::runtime_error("Pauli term acts outside the qubits");
}

namespace {

struct Atom {
  bool tail = false;  // backward order: no later atom touches its qubits, so un-applying it serves nothing
  bool diag = false;
  int nq = 1;
  int bit[2] = {-1, -1};
  std::vector<int> gates;
};

struct SweepOut {
  std::vector<int> tile_bits;  // ascending state-index bits; local bit j = tile_bits[j]
  int pass_begin = 0, pass_end = 0;
};

struct DiagRun {
  std::vector<int32_t> reg_list;
  std::vector<std::vector<int32_t>> grp_list;
  std::vector<DevOp> pair_ops, cross_ops, gd_ops;
  bool any = false, any_const = false;
  uint32_t bits = 0;   // state-index bits some pending diagonal gate acts on
  uint32_t vmask = 0;  // marginal vectors the run's gradient ops read (OP_GD_BEGIN)
};

DevOp make_op(int type) {
  DevOp o;
  std::memset(&o, 0, sizeof(o));
  o.type = type;
  o.p0 = o.p1 = -1;
  o.coef = -1;
  o.gslot = -1;
  o.aux0 = o.aux1 = -1;
  return o;
}

class Compiler {
 public:
  explicit Compiler(HostPlan& hp, bool defer_tails = false) : hp_(hp), defer_tails_(defer_tails) {}

  int bit_of(int q) const { return hp_.n - 1 - q; }

  int32_t alloc_coef(int nfloats) {
    int32_t off = hp_.ncoef;
    hp_.ncoef += (nfloats + 3) & ~3;
    return off;
  }
  void add_job(int kind, int32_t out, int a, int b, int c, int d, const std::vector<int32_t>& list) {
    PrepJob j;
    j.kind = kind; j.out = out; j.a = a; j.b = b; j.c = c; j.d = d;
    j.list_off = (int32_t)hp_.lists.size();
    j.list_len = (int32_t)list.size();
    hp_.lists.insert(hp_.lists.end(), list.begin(), list.end());
    hp_.jobs.push_back(j);
  }

  std::vector<Atom> forward_atoms() const {
    std::vector<Atom> atoms;
    std::vector<int> last(hp_.n_eff, -1);  // last atom touching each bit
    for (int gi = 0; gi < (int)hp_.gates.size(); ++gi) {
      const qhbm_gate_t& g = hp_.gates[gi];
      if (g.type == QHBM_GATE_I) continue;
      const bool two = gate_is_two_qubit(g.type);
      const bool diag = gate_is_diagonal(g.type);
      if (!two) {
        int b = bit_of(g.q0);
        int la = last[b];
        // consecutive 1-qubit gates of the same class fuse (non-diagonal: one 2x2 product;
        // diagonal: one phase-table entry).  Mixed classes stay apart so that X/Y powers keep
        // their cheap rotation form and diagonal gates fold into the pass's phase tables.
        if (la >= 0 && atoms[la].nq == 1 && atoms[la].diag == diag) {
          atoms[la].gates.push_back(gi);
          continue;
        }
        Atom a;
        a.diag = diag; a.nq = 1; a.bit[0] = b; a.gates.push_back(gi);
        last[b] = (int)atoms.size();
        atoms.push_back(a);
      } else {
        Atom a;
        a.diag = diag; a.nq = 2; a.bit[0] = bit_of(g.q0); a.bit[1] = bit_of(g.q1);
        a.gates.push_back(gi);
        last[a.bit[0]] = last[a.bit[1]] = (int)atoms.size();
        atoms.push_back(a);
      }
    }
    return atoms;
  }

  std::vector<Atom> backward_atoms() const {
    std::vector<Atom> atoms;
    for (int gi = (int)hp_.gates.size() - 1; gi >= 0; --gi) {
      const qhbm_gate_t& g = hp_.gates[gi];
      if (g.type == QHBM_GATE_I) continue;
      Atom a;
      a.diag = gate_is_diagonal(g.type);
      a.nq = gate_is_two_qubit(g.type) ? 2 : 1;
      a.bit[0] = bit_of(g.q0);
      if (a.nq == 2) a.bit[1] = bit_of(g.q1);
      a.gates.push_back(gi);
      atoms.push_back(a);
    }
    // The first gates of the circuit come last here.  Nothing is un-applied after them on their qubits and
    // every later gradient inner product <lam| dG |psi> sits on other qubits, which they commute with:
    // only their own gradient is needed, never the un-application (a third of a rotation's arithmetic).
    std::vector<char> seen(hp_.n_eff, 0);
    for (size_t ai = atoms.size(); ai-- > 0;) {
      Atom& a = atoms[ai];
      a.tail = true;
      for (int i = 0; i < a.nq; ++i) a.tail = a.tail && !seen[a.bit[i]];
      for (int i = 0; i < a.nq; ++i) seen[a.bit[i]] = 1;
    }
    return atoms;
  }

  // ---- readiness bookkeeping -------------------------------------------------
  struct Block {
    std::vector<char> nd, d;
    explicit Block(int n) : nd(n, 0), d(n, 0) {}
    bool ready(const Atom& a) const {
      for (int i = 0; i < a.nq; ++i) {
        if (nd[a.bit[i]]) return false;
        if (!a.diag && d[a.bit[i]]) return false;
      }
      return true;
    }
    void block(const Atom& a) {
      for (int i = 0; i < a.nq; ++i) (a.diag ? d : nd)[a.bit[i]] = 1;
    }
  };

  std::vector<int> choose_tile(const std::vector<Atom>& atoms, const std::vector<char>& done) const {
    const int n = hp_.n_eff, T = hp_.T;
    std::vector<char> in(n, 0);
    int cnt = 0;
    if (n <= T) {
      std::vector<int> all(n);
      for (int i = 0; i < n; ++i) all[i] = i;
      return all;
    }
    for (int b = 0; b < 5; ++b) { in[b] = 1; ++cnt; }  // 256-byte contiguous global segments
    Block blk(n);
    std::vector<int> later;
    for (size_t ai = 0; ai < atoms.size(); ++ai) {
      if (done[ai]) continue;
      const Atom& a = atoms[ai];
      if (!blk.ready(a)) { blk.block(a); if (!a.diag) later.push_back((int)ai); continue; }
      if (a.diag) continue;
      int need = 0;
      for (int i = 0; i < a.nq; ++i) if (!in[a.bit[i]]) ++need;
      if (a.nq == 2 && a.bit[0] == a.bit[1]) need = std::min(need, 1);
      if (cnt + need <= T) {
        for (int i = 0; i < a.nq; ++i) if (!in[a.bit[i]]) { in[a.bit[i]] = 1; ++cnt; }
      } else {
        blk.block(a);
        later.push_back((int)ai);
      }
    }
    for (int ai : later) {
      const Atom& a = atoms[ai];
      int need = 0;
      for (int i = 0; i < a.nq; ++i) if (!in[a.bit[i]]) ++need;
      if (cnt + need <= T)
        for (int i = 0; i < a.nq; ++i) if (!in[a.bit[i]]) { in[a.bit[i]] = 1; ++cnt; }
    }
    for (int b = n - 1; b >= 0 && cnt < T; --b) if (!in[b]) { in[b] = 1; ++cnt; }
    std::vector<int> out;
    for (int b = 0; b < n; ++b) if (in[b]) out.push_back(b);
    return out;
  }

  // Register-position allocator: positions (1,0) and (3,2) can host a 2-qubit block.
  struct RegAlloc {
    int K;
    int pos_bit[kMaxRegQubits];  // position -> tile-local bit, -1 free
    explicit RegAlloc(int k) : K(k) { for (int i = 0; i < kMaxRegQubits; ++i) pos_bit[i] = -1; }
    int pos_of(int lb) const { for (int i = 0; i < K; ++i) if (pos_bit[i] == lb) return i; return -1; }
    int nfree() const { int c = 0; for (int i = 0; i < K; ++i) c += pos_bit[i] < 0; return c; }
    bool place1(int lb) {
      if (pos_of(lb) >= 0) return true;
      for (int i = K - 1; i >= 0; --i) if (pos_bit[i] < 0) { pos_bit[i] = lb; return true; }
      return false;
    }
    bool place2(int la, int lb) {
      int pa = pos_of(la), pb = pos_of(lb);
      if (pa >= 0 && pb >= 0) return (pa ^ pb) == 1 && std::max(pa, pb) < 4;
      if (pa >= 0 || pb >= 0) {
        int p = pa >= 0 ? pa : pb;
        int other = pa >= 0 ? lb : la;
        if (p >= 4 || (p ^ 1) >= K || pos_bit[p ^ 1] >= 0) return false;
        pos_bit[p ^ 1] = other;
        return true;
      }
      for (int base = 0; base + 1 < std::min(K, 4); base += 2)
        if (pos_bit[base] < 0 && pos_bit[base + 1] < 0) { pos_bit[base + 1] = la; pos_bit[base] = lb; return true; }
      return false;
    }
  };

  // Schedules `atoms` (already in execution order) into sweeps and passes.
  void schedule(const std::vector<Atom>& atoms, bool backward, std::vector<SweepOut>& sweeps) {
    const int n = hp_.n_eff, K = hp_.K;
    std::vector<char> done(atoms.size(), 0);
    size_t remaining = atoms.size();
    // Gradient scratch = the two dead smem tiles = 4 * 2^K units of one float per thread.  A rotation /
    // matrix-block gradient takes one unit, every marginal vector of a diagonal run two (complex).
    const int max_units = std::min(4 * (1 << K), 255);
    while (remaining > 0) {
      SweepOut sw;
      sw.tile_bits = choose_tile(atoms, done);
      if (backward && sweeps.empty() && n > hp_.T && std::getenv("QHBM_NO_FUSE") == nullptr) {
        // The expectation phase runs on the contiguous tile map and shares a launch with the first
        // backward sweep when that sweep uses the same map: prefer it unless it would run fewer of the
        // gates that are ready now.
        std::vector<int> contiguous;
        for (int b = 0; b < hp_.T; ++b) contiguous.push_back(b);
        auto ready_inside = [&](const std::vector<int>& bits) {
          uint32_t mask = 0;
          for (int b : bits) mask |= 1u << b;
          Block blk(n);
          int count = 0;
          for (size_t ai = 0; ai < atoms.size(); ++ai) {
            if (done[ai]) continue;
            const Atom& a = atoms[ai];
            if (!blk.ready(a)) { blk.block(a); continue; }
            bool inside = true;
            for (int i = 0; i < a.nq; ++i) inside = inside && ((mask >> a.bit[i]) & 1);
            if (inside && !a.diag) ++count;
            else if (!inside) blk.block(a);
          }
          return count;
        };
        if (ready_inside(contiguous) >= ready_inside(sw.tile_bits)) sw.tile_bits = contiguous;
      }
      sw.pass_begin = (int)hp_.passes.size();
      std::vector<int> local_of(n, -1);
      for (size_t j = 0; j < sw.tile_bits.size(); ++j) local_of[sw.tile_bits[j]] = (int)j;
      const int Tloc = (int)sw.tile_bits.size();
      size_t executed_in_sweep = 0;
      while (remaining > 0) {
        // -- tentative scan: pick the register qubits
        RegAlloc ra(K);
        bool placed_only_tails = true, outside_left = false;
        {
          Block blk(n);
          for (size_t ai = 0; ai < atoms.size(); ++ai) {
            if (done[ai]) continue;
            const Atom& a = atoms[ai];
            if (!a.diag)
              for (int i = 0; i < a.nq; ++i) outside_left = outside_left || local_of[a.bit[i]] < 0;
            if (!blk.ready(a)) { blk.block(a); continue; }
            if (a.diag) continue;
            bool ok = true;
            for (int i = 0; i < a.nq; ++i) ok = ok && local_of[a.bit[i]] >= 0;
            if (ok) {
              RegAlloc trial = ra;
              ok = a.nq == 1 ? trial.place1(local_of[a.bit[0]])
                             : trial.place2(local_of[a.bit[0]], local_of[a.bit[1]]);
              if (ok) { ra = trial; placed_only_tails = placed_only_tails && a.tail; }
            }
            if (!ok) blk.block(a);
          }
        }
        const bool have_nondiag = ra.nfree() < K;
        // A further pass of this sweep that would only take gradients of the circuit's first gates (they are
        // not un-applied, so nothing waits for them) is not worth its shared-memory round trip when another
        // sweep follows anyway: those gates join that sweep's passes instead.
        if (defer_tails_ && backward && (int)hp_.passes.size() > sw.pass_begin && have_nondiag && placed_only_tails &&
            outside_left)
          break;
        // pad the register set from the top of the tile
        for (int lb = Tloc - 1; lb >= 0 && ra.nfree() > 0; --lb)
          if (ra.pos_of(lb) < 0) ra.place1(lb);
        // -- final scan: emit
        DevPass ps;
        std::memset(&ps, 0, sizeof(ps));
        for (int j = 0; j < kMaxRegQubits; ++j) { ps.regbit[j] = 0; ps.sorted[j] = 0; }
        for (int j = 0; j < K; ++j) ps.regbit[j] = ra.pos_bit[j];
        {
          std::vector<int> s(ps.regbit, ps.regbit + K);
          std::sort(s.begin(), s.end());
          for (int j = 0; j < K; ++j) ps.sorted[j] = s[j];
        }
        for (int r = 0; r < (1 << kMaxRegQubits); ++r) {
          uint32_t dep = 0;
          for (int j = 0; j < K; ++j) if ((r >> j) & 1) dep |= 1u << ps.regbit[j];
          const uint32_t sw = (dep & ~15u) | ((dep ^ (dep >> 4) ^ (dep >> 8) ^ (dep >> 12)) & 15u);
          ps.eoff8[r] = r < (1 << K) ? 8u * sw : 0u;
        }
        ps.op_begin = (int)hp_.ops.size();
        ps.gsym_off = (int)hp_.gsym.size();
        ps.ngrad = 0;
        ps.coef_begin = hp_.ncoef;
        std::vector<int> regpos(n, -1);  // state bit -> register position
        for (int j = 0; j < K; ++j) regpos[sw.tile_bits[ra.pos_bit[j]]] = j;

        DiagRun run;
        run.grp_list.assign((n + kConstGroupBits - 1) / kConstGroupBits, {});
        size_t executed = 0;
        Block blk(n);
        units_ = 0;
        tasks_bound_ = 0;
        // marginal vectors the gradient of diagonal atom x reads (see OP_GD_BEGIN), 0 if it has no symbol
        auto atom_vmask = [&](const Atom& x) -> uint32_t {
          if (!backward || !x.diag) return 0u;
          uint32_t m = 0;
          for (int gi : x.gates) {
            const qhbm_gate_t& g = hp_.gates[gi];
            bool any = false;
            for (int k = 0; k < g.nparams; ++k) any = any || g.sym[k] >= 0;
            if (!any) continue;
            const int pa = regpos[x.bit[0]], pb = x.nq == 2 ? regpos[x.bit[1]] : -1;
            m |= 1u;
            if (pa >= 0) m |= 2u << pa;
            if (pb >= 0) m |= 2u << pb;
            if (pa >= 0 && pb >= 0) {
              const int ph = std::max(pa, pb), pl = std::min(pa, pb);
              m |= 1u << (1 + K + ph * (ph - 1) / 2 + pl);
            }
          }
          return m;
        };
        auto units_with = [&](uint32_t extra_vmask, int extra_float) {
          return units_ + 2 * __builtin_popcount(run.vmask | extra_vmask) + extra_float;
        };
        // A pass's program (op descriptors + coefficients) is staged in a fixed shared-memory buffer: the
        // pass is closed before the next atom could overflow it (headroom = the largest single atom, the
        // tables the pending diagonal run will emit when flushed, and the merged-rotation tables).
        int n_rot = 0;
        const int stage_ops = backward ? kStageOpsAdj : kStageOps;
        auto fits = [&](const Atom& a) {
          int rops = 0, rcoef = 0;
          if (run.any) {
            if (!run.gd_ops.empty()) rops += 1 + (int)run.gd_ops.size();
            for (const auto& g : run.grp_list) if (!g.empty()) { ++rops; rcoef += 2 << kConstGroupBits; }
            rops += (int)run.pair_ops.size() + (int)run.cross_ops.size() + 1;
            rcoef += 4 << K;
          }
          const int ng = (int)a.gates.size();
          const int ops_now = (int)hp_.ops.size() - ps.op_begin + rops;
          const int coef_now = hp_.ncoef - ps.coef_begin + rcoef + n_rot * (4 * K + 4);
          // (a gradient pass also stages its reduction tasks: <= 1 per float slot, <= 3 + 1 per diagonal op)
          // coefficient floats the atom itself may add (its tables are in the constant headroom below):
          // X/Y rotation 4 + a merged table; diagonal gate 8 per gradient; 2x2 block 8 + 8 per gradient;
          // 4x4 block 32 + 32 per gradient
          int own = 0;
          for (int gi : a.gates) {
            const qhbm_gate_t& g = hp_.gates[gi];
            int nsym = 0;
            for (int k = 0; k < g.nparams; ++k) nsym += g.sym[k] >= 0;
            if (a.diag) own += 8 * nsym;
            else if (a.nq == 2) own += 32 + 32 * nsym + 8;
            else if (g.type == QHBM_GATE_XPOW || g.type == QHBM_GATE_YPOW) own += 4 + 4 * K + 4;
            else own += 8 + 8 * nsym + 8;
          }
          return ops_now + 6 * ng + 4 + (backward ? tasks_bound_ + 4 * ng + 1 : 0) <= stage_ops &&
                 coef_now + own + (2 << kConstGroupBits) + (4 << K) + (4 * K + 4) <= kStageCoef;
        };
        for (size_t ai = 0; ai < atoms.size(); ++ai) {
          if (done[ai]) continue;
          const Atom& a = atoms[ai];
          if (!blk.ready(a)) { blk.block(a); continue; }
          if (!fits(a)) break;
          bool ok = true;
          int ngrads = 0;
          if (backward) {
            const qhbm_gate_t& g = hp_.gates[a.gates[0]];
            for (int k = 0; k < g.nparams; ++k) ngrads += g.sym[k] >= 0;
            if (ps.ngrad + ngrads > 255) ok = false;
            if (a.diag ? units_with(atom_vmask(a), 0) > max_units : units_with(0, ngrads) > max_units) ok = false;
          }
          if (ok && !a.diag) {
            for (int i = 0; i < a.nq; ++i) ok = ok && regpos[a.bit[i]] >= 0;
            if (ok && a.nq == 2) {
              int pa = regpos[a.bit[0]], pb = regpos[a.bit[1]];
              ok = (pa ^ pb) == 1 && std::max(pa, pb) < 4;
            }
          }
          if (!ok) { blk.block(a); continue; }
          if (a.diag) emit_diag(a, backward, regpos, ps, run);
          else {
            // a non-diagonal block commutes with every pending diagonal gate on OTHER qubits, so the
            // pending run stays open (and keeps growing into one table op) unless it touches this block
            uint32_t abits = 0;
            for (int i = 0; i < a.nq; ++i) abits |= 1u << a.bit[i];
            if (run.bits & abits) {
              // before closing the run, pull in every later diagonal gate that commutes with all the
              // blocks still ahead of it (e.g. the Z^-s of the other qubits of this layer)
              Block b2 = blk;
              for (size_t aj = ai; aj < atoms.size(); ++aj) {
                if (done[aj]) continue;
                const Atom& x = atoms[aj];
                if (!x.diag || !b2.ready(x)) { b2.block(x); continue; }
                int ng = 0;
                if (backward) {
                  const qhbm_gate_t& g = hp_.gates[x.gates[0]];
                  for (int k = 0; k < g.nparams; ++k) ng += g.sym[k] >= 0;
                }
                // (the block that closes the run, atom `a`, still needs its own `ngrads` units)
                if (ps.ngrad + ng + ngrads > 255 || units_with(atom_vmask(x), ngrads) > max_units || !fits(x)) {
                  b2.block(x);
                  continue;
                }
                emit_diag(x, backward, regpos, ps, run);
                done[aj] = 1;
                --remaining;
                ++executed;
              }
              flush_diag(run, backward);
            }
            emit_nondiag(a, backward, regpos, ps);
            ++n_rot;
          }
          done[ai] = 1;
          --remaining;
          ++executed;
        }
        flush_diag(run, backward);
        if (executed == 0) break;  // nothing runnable with this tile: next sweep
        merge_rotations(ps);
        ps.op_end = (int)hp_.ops.size();
        ps.coef_end = hp_.ncoef;
        if (ps.op_end - ps.op_begin > stage_ops || ps.coef_end - ps.coef_begin > kStageCoef)
          throw std::runtime_error("internal: a pass program exceeds the staging buffer");
        if (units_ > max_units) throw std::runtime_error("internal: a pass needs more gradient scratch than the tiles hold");
        hp_.passes.push_back(ps);
        executed_in_sweep += executed;
        if (!have_nondiag && remaining > 0) {
          // only diagonal work was possible; a different tile is needed for the rest
          break;
        }
      }
      sw.pass_end = (int)hp_.passes.size();
      if (executed_in_sweep == 0)
        throw std::runtime_error("internal: scheduler made no progress");
      sweeps.push_back(sw);
    }
  }

  // Peephole: consecutive X (or Y) rotations on distinct register positions become ONE op whose
  // handler is straight-line code over the positions (no per-gate dispatch).
  void merge_rotations(const DevPass& ps) {
    std::vector<DevOp> out;
    const int K = hp_.K;
    size_t i = (size_t)ps.op_begin;
    const size_t end = hp_.ops.size();
    while (i < end) {
      const DevOp& o = hp_.ops[i];
      if (o.type != OP_XROT && o.type != OP_YROT) { out.push_back(o); ++i; continue; }
      size_t j = i;
      int mask = 0, used = 0;  // positions that are rotated / positions taken by this op (rotated or gradient-only)
      while (j < end && hp_.ops[j].type == o.type && !(used & (1 << hp_.ops[j].p0))) {
        used |= 1 << hp_.ops[j].p0;
        if (hp_.ops[j].aux0 != 1) mask |= 1 << hp_.ops[j].p0;
        ++j;
      }
      DevOp m = make_op(o.type == OP_XROT ? OP_XROTM : OP_YROTM);
      m.p0 = mask;
      m.aux0 = 0;
      m.coef = alloc_coef(4 * K);
      // per-position gradient slots, 8 bits each: positions 0..3 in aux1, position 4 in p1
      m.aux1 = 0;
      m.p1 = 0;
      for (size_t q = i; q < j; ++q) {
        const DevOp& r = hp_.ops[q];
        for (auto& job : hp_.jobs) {  // retarget the coefficient jobs of this rotation
          if ((job.kind == PJ_ROT && job.out == r.coef) || (job.kind == PJ_KAPPA && job.out == r.coef + 2))
            job.out = m.coef + 4 * r.p0 + (job.kind == PJ_KAPPA ? 2 : 0);
        }
        if (r.gslot >= 0) {
          m.aux0 |= 1 << r.p0;
          if (r.p0 < 4) m.aux1 |= (r.gslot & 0xff) << (8 * r.p0);
          else m.p1 = r.gslot;
        }
      }
      if (m.type == OP_XROTM && mask == used && __builtin_popcount(mask) >= K - 1) {
        // one job computes all K rotations of the op together (it needs the product of the cosines)
        m.type = OP_XROTF;
        std::vector<int32_t> per_pos(K, -1);
        int dag = 0;
        for (auto& job : hp_.jobs) {
          if (job.kind == PJ_ROT && job.out >= m.coef && job.out < m.coef + 4 * K) {
            per_pos[(job.out - m.coef) / 4] = hp_.lists[job.list_off];
            dag = job.a;
            job.kind = PJ_NONE;
          } else if (job.kind == PJ_KAPPA && job.out >= m.coef && job.out < m.coef + 4 * K) {
            job.kind = PJ_NONE;
          }
        }
        add_job(PJ_ROTF, m.coef, dag, m.aux0, 0, K, per_pos);
      }
      out.push_back(m);
      i = j;
    }
    hp_.ops.resize(ps.op_begin);
    hp_.ops.insert(hp_.ops.end(), out.begin(), out.end());
  }

  void emit_nondiag(const Atom& a, bool backward, const std::vector<int>& regpos, DevPass& ps) {
    if (a.nq == 1 && a.gates.size() == 1 &&
        (hp_.gates[a.gates[0]].type == QHBM_GATE_XPOW || hp_.gates[a.gates[0]].type == QHBM_GATE_YPOW)) {
      // two-level X/Y power: rotation form, global phase dropped (tracked for debug output)
      const int p = regpos[a.bit[0]];
      const int gi = a.gates[0];
      const qhbm_gate_t& g = hp_.gates[gi];
      const bool is_y = g.type == QHBM_GATE_YPOW;
      const bool grad_only = backward && a.tail && std::getenv("QHBM_NO_TAIL_SKIP") == nullptr;
      if (grad_only && g.sym[0] < 0) return;  // nothing depends on this un-rotation
      DevOp o = make_op(is_y ? OP_YROT : OP_XROT);
      o.p0 = p;
      o.aux0 = grad_only ? 1 : 0;  // 1: gradient inner product only (merge_rotations keeps it out of the rotation mask)
      o.coef = alloc_coef(4);  // (c, s, kappa, -)
      add_job(PJ_ROT, o.coef, backward ? 1 : 0, 0, 0, 0, {gi});
      if (backward) {
        if (g.sym[0] >= 0) {  // gradient inner product fused into the un-rotation (gslot >= 0)
          o.gslot = ps.ngrad++;
          ++units_;
          ++tasks_bound_;
          hp_.gsym.push_back(g.sym[0]);
          add_job(PJ_KAPPA, o.coef + 2, 0, is_y ? 1 : 0, 0, 0, {gi});
        }
      } else {
        phase_gates_.push_back(gi);
      }
      hp_.ops.push_back(o);
    } else if (a.nq == 1) {
      const int p = regpos[a.bit[0]];
      if (backward) {
        const int gi = a.gates[0];
        const qhbm_gate_t& g = hp_.gates[gi];
        for (int k = 0; k < g.nparams; ++k) {
          if (g.sym[k] < 0) continue;
          DevOp o = make_op(OP_GRAD_MAT1);
          o.p0 = p;
          o.coef = alloc_coef(8);
          o.gslot = ps.ngrad++;
          ++units_;
          ++tasks_bound_;
          hp_.gsym.push_back(g.sym[k]);
          add_job(PJ_GRAD1, o.coef, 0, 0, k, 0, {gi});
          hp_.ops.push_back(o);
        }
      }
      DevOp o = make_op(OP_MAT1);
      o.p0 = p;
      o.coef = alloc_coef(8);
      add_job(PJ_MAT1, o.coef, backward ? 1 : 0, 0, 0, 0, std::vector<int32_t>(a.gates.begin(), a.gates.end()));
      hp_.ops.push_back(o);
    } else {
      const int pa = regpos[a.bit[0]], pb = regpos[a.bit[1]];
      const int pair = std::max(pa, pb) == 1 ? 0 : 1;
      const int swap = pa < pb ? 1 : 0;  // matrix index = 2*bit(hi position) + bit(lo position)
      const int gi = a.gates[0];
      if (backward) {
        const qhbm_gate_t& g = hp_.gates[gi];
        for (int k = 0; k < g.nparams; ++k) {
          if (g.sym[k] < 0) continue;
          DevOp o = make_op(OP_GRAD_MAT2);
          o.p0 = pair;
          o.coef = alloc_coef(32);
          o.gslot = ps.ngrad++;
          ++units_;
          ++tasks_bound_;
          hp_.gsym.push_back(g.sym[k]);
          add_job(PJ_GRAD2, o.coef, 0, swap, k, 0, {gi});
          hp_.ops.push_back(o);
        }
      }
      DevOp o = make_op(OP_MAT2);
      o.p0 = pair;
      o.coef = alloc_coef(32);
      add_job(PJ_MAT2, o.coef, backward ? 1 : 0, swap, 0, 0, {gi});
      hp_.ops.push_back(o);
    }
  }

  void emit_diag(const Atom& a, bool backward, const std::vector<int>& regpos, DevPass& ps, DiagRun& run) {
    run.any = true;
    for (int i = 0; i < a.nq; ++i) run.bits |= 1u << a.bit[i];
    const int dag = backward ? 1 : 0;
    const int pa = regpos[a.bit[0]];
    const int pb = a.nq == 2 ? regpos[a.bit[1]] : -1;
    const bool reg_a = pa >= 0, reg_b = a.nq == 2 && pb >= 0;
    const bool all_reg = a.nq == 1 ? reg_a : (reg_a && reg_b);
    const bool all_const = a.nq == 1 ? !reg_a : (!reg_a && !reg_b);
    for (int gi : a.gates) {
      const qhbm_gate_t& g = hp_.gates[gi];
      // ---- gradient inner products (before the un-apply; they commute with the whole run)
      if (backward) {
        for (int k = 0; k < g.nparams; ++k) {
          if (g.sym[k] < 0) continue;
          DevOp o = make_op(OP_GD_CONST);
          int swap = 0;
          if (all_const) {
            o.type = OP_GD_CONST;
            o.aux0 = a.bit[0];
            o.aux1 = a.nq == 2 ? a.bit[1] : -1;
          } else if (all_reg && a.nq == 1) {
            o.type = OP_GD_REG1;
            o.p0 = pa;
          } else if (all_reg) {
            o.type = OP_GD_REG2;  // kernel wants p0 > p1
            if (pa > pb) { o.p0 = pa; o.p1 = pb; } else { o.p0 = pb; o.p1 = pa; swap = 1; }
          } else {
            o.type = OP_GD_MIX;  // entries indexed [2*bit(const) + regbit]
            if (reg_a) { o.p0 = pa; o.aux0 = a.bit[1]; swap = 1; } else { o.p0 = pb; o.aux0 = a.bit[0]; }
          }
          o.coef = alloc_coef(8);
          o.gslot = ps.ngrad++;
          hp_.gsym.push_back(g.sym[k]);
          add_job(PJ_GDIAG, o.coef, 0, swap, k, 0, {gi});
          tasks_bound_ += run.gd_ops.empty() ? 4 : 3;
          run.vmask |= 1u;
          if (reg_a) run.vmask |= 2u << pa;
          if (reg_b) run.vmask |= 2u << pb;
          if (reg_a && reg_b) {
            const int ph = std::max(pa, pb), pl = std::min(pa, pb);
            run.vmask |= 1u << (1 + hp_.K + ph * (ph - 1) / 2 + pl);
          }
          run.gd_ops.push_back(o);
        }
      }
      // ---- the phase itself
      if (all_reg) {
        run.reg_list.push_back(gi);
        run.reg_list.push_back(pa);
        run.reg_list.push_back(a.nq == 2 ? pb : -1);
      } else if (all_const) {
        run.any_const = true;
        const int ga = a.bit[0] / kConstGroupBits;
        const int gb = a.nq == 2 ? a.bit[1] / kConstGroupBits : ga;
        if (ga == gb) {
          run.grp_list[ga].push_back(gi);
          run.grp_list[ga].push_back(a.bit[0] - ga * kConstGroupBits);
          run.grp_list[ga].push_back(a.nq == 2 ? a.bit[1] - ga * kConstGroupBits : -1);
        } else {
          DevOp o = make_op(OP_DCONST_PAIR);
          o.aux0 = a.bit[0];
          o.aux1 = a.bit[1];
          o.coef = alloc_coef(8);
          add_job(PJ_DPAIR, o.coef, dag, 0, 0, 0, {gi});
          run.pair_ops.push_back(o);
        }
      } else {
        DevOp o = make_op(OP_DCROSS);  // entries indexed [2*bit(const) + regbit]
        int swap;
        if (reg_a) { o.p0 = pa; o.aux0 = a.bit[1]; swap = 1; } else { o.p0 = pb; o.aux0 = a.bit[0]; swap = 0; }
        o.coef = alloc_coef(8);
        add_job(PJ_DPAIR, o.coef, dag, swap, 0, 0, {gi});
        run.cross_ops.push_back(o);
      }
    }
  }

  void flush_diag(DiagRun& run, bool backward) {
    if (!run.any) return;
    const int dag = backward ? 1 : 0;
    if (!run.gd_ops.empty()) {
      auto rank = [](const DevOp& o) { return o.type == OP_GD_CONST ? 0 : (o.type == OP_GD_REG2 ? 2 : 1); };
      std::stable_sort(run.gd_ops.begin(), run.gd_ops.end(),
                       [&](const DevOp& x, const DevOp& y) { return rank(x) < rank(y); });
      DevOp b = make_op(OP_GD_BEGIN);
      b.aux0 = (int)run.gd_ops.size();
      b.aux1 = 0;
      b.p0 = b.p1 = 0;
      for (auto& o : run.gd_ops) {
        if (o.type == OP_GD_REG2) b.aux1 = 1;
        if (rank(o) == 0) ++b.p0;
        if (rank(o) == 1) ++b.p1;
      }
      hp_.ops.push_back(b);
      for (auto& o : run.gd_ops) hp_.ops.push_back(o);
    }
    for (size_t g = 0; g < run.grp_list.size(); ++g) {
      if (run.grp_list[g].empty()) continue;
      DevOp o = make_op(OP_DCONST_TAB);
      o.aux0 = (int)g * kConstGroupBits;
      o.aux1 = (1 << kConstGroupBits) - 1;
      o.coef = alloc_coef(2 << kConstGroupBits);
      add_job(PJ_DTAB, o.coef, dag, 0, 0, kConstGroupBits, run.grp_list[g]);
      hp_.ops.push_back(o);
    }
    for (auto& o : run.pair_ops) hp_.ops.push_back(o);
    if (!run.reg_list.empty()) {
      DevOp o = make_op(OP_DREG_TAB);
      o.aux0 = run.any_const ? 1 : 0;  // 0: no thread-constant factor is pending (F == 1)
      o.coef = alloc_coef(4 << hp_.K);  // entries (re, im, -im, im): see OP_DREG_TAB
      add_job(PJ_DTAB, o.coef, dag, 1, 0, hp_.K, run.reg_list);
      hp_.ops.push_back(o);
    } else if (run.any_const) {
      hp_.ops.push_back(make_op(OP_DAPPLY));
    }
    for (auto& o : run.cross_ops) hp_.ops.push_back(o);
    units_ += 2 * __builtin_popcount(run.vmask);
    const size_t ng = run.grp_list.size();
    run = DiagRun();
    run.grp_list.assign(ng, {});
  }

  void fill_runs(const std::vector<int>& tile_bits, LaunchDesc& L) const {
    const int n = hp_.n_eff;
    std::vector<char> in(n, 0);
    for (int b : tile_bits) in[b] = 1;
    L.tile_mask = 0;
    for (int b : tile_bits) L.tile_mask |= 1u << b;
    auto build = [&](const std::vector<int>& bits, BitRun* runs, int32_t& nr) {
      nr = 0;
      size_t j = 0;
      while (j < bits.size()) {
        size_t e = j + 1;
        while (e < bits.size() && bits[e] == bits[e - 1] + 1) ++e;
        if (nr >= kMaxRuns) throw std::runtime_error("internal: too many bit runs in a tile map");
        runs[nr].local_start = (int8_t)j;
        runs[nr].global_start = (int8_t)bits[j];
        runs[nr].len = (int8_t)(e - j);
        runs[nr].pad = 0;
        ++nr;
        j = e;
      }
    };
    std::vector<int> obits;
    for (int b = 0; b < n; ++b) if (!in[b]) obits.push_back(b);
    build(tile_bits, L.runs, L.n_runs);
    build(obits, L.oruns, L.n_oruns);
    const int mshift = (int)tile_bits.size() - hp_.K;
    for (int m = 0; m < (1 << kMaxRegQubits); ++m) {
      L.moff[m] = 0;
      L.soff[m] = 0;
      if (m >= (1 << hp_.K) || mshift < 0) continue;
      const uint32_t l = (uint32_t)m << mshift;
      uint32_t g = 0;
      for (size_t j = 0; j < tile_bits.size(); ++j) if ((l >> j) & 1) g |= 1u << tile_bits[j];
      L.moff[m] = g;
      L.soff[m] = (uint16_t)((l & ~15u) | ((l ^ (l >> 4) ^ (l >> 8) ^ (l >> 12)) & 15u));
    }
  }

  // Expectation stages.  Stage 0 runs in the expectation launch (contiguous tile map) and owns every
  // x-group whose flips stay inside that tile, all diagonal terms and -- as a fallback -- groups that
  // fit no tile (their partner amplitudes are read across tiles through L2).  Forward-only plans on
  // multi-tile states add stages with other tile maps so that the remaining groups also find their
  // partners in shared memory: one extra read of the state instead of one per out-of-tile qubit.
  struct RawGroup {
    int op = 0;
    uint32_t x = 0;
    std::vector<int> term_idx;
    int stage = 0;
  };

  void build_terms(const OpsIR& o) {
    // k = coeff * (-i)^{ny}; sign from parity(i & z) of the OUTPUT index i (derivation in DESIGN.md).
    const int Tl = std::min(hp_.T, hp_.n_eff);
    const bool multi = hp_.n_eff > hp_.T;
    // Many diagonal terms (e.g. the K Z-string shards of a modular Hamiltonian, hamiltonian.py:48-51):
    // evaluate them all at once from one Walsh-Hadamard transform of |psi|^2 per tile.
    int n_diag = 0;
    for (const auto& t : o.terms) n_diag += t.xmask == 0;
    const bool use_wht = n_diag >= kWhtMinTerms;
    // A single observable whose strings flip at most two in-tile qubits runs as observable passes
    // (OP_HX / OP_HD); everything else stays in the generic x-group tables below.
    // Adjoint plans only: measured on B200, the forward-only kernel is faster with the generic tables
    // (the passes' barriers and staging cost more than the per-amplitude sign arithmetic they save).
    const bool hpass_ok = (hp_.grad || std::getenv("QHBM_HPASS_FORWARD")) && o.n_ops() == 1 && !use_wht && !no_hpass_;
    std::vector<RawGroup> raw;
    for (int j = 0; j < o.n_ops(); ++j) {
      std::vector<int> idx;
      for (int t = o.offsets[j]; t < o.offsets[j + 1]; ++t) idx.push_back(t);
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return o.terms[a].xmask < o.terms[b].xmask; });
      size_t i = 0;
      while (i < idx.size()) {
        const uint32_t x = o.terms[idx[i]].xmask;
        size_t e = i;
        while (e < idx.size() && o.terms[idx[e]].xmask == x) ++e;
        if (x == 0 && use_wht) {
          for (; i < e; ++i) {
            DevDiagTerm d;
            d.coeff = o.terms[idx[i]].coeff;
            d.z = o.terms[idx[i]].zmask;
            d.op = j;
            d.pad = 0;
            hp_.dterms.push_back(d);
          }
          continue;
        }
        RawGroup g;
        g.op = j;
        g.x = x;
        g.term_idx.assign(idx.begin() + (long)i, idx.begin() + (long)e);
        raw.push_back(g);
        i = e;
      }
    }
    // ---- stage tile maps
    stage_bits_.clear();
    {
      std::vector<int> contiguous;
      for (int b = 0; b < Tl; ++b) contiguous.push_back(b);
      stage_bits_.push_back(contiguous);
    }
    auto mask_of = [](const std::vector<int>& bits) {
      uint32_t m = 0;
      for (int b : bits) m |= 1u << b;
      return m;
    };
    const uint32_t mask0 = multi ? mask_of(stage_bits_[0]) : 0xffffffffu;
    // Extra stages are OFF by default: on B200 the cross-tile partner reads are served by L2 and the
    // extra launch costs as much as it saves (n = 20 TFIM forward: 48.0 vs 47.1 ms; profiles/).  The
    // switch keeps the path testable.
    const bool extra_stages = multi && !hp_.grad && std::getenv("QHBM_EXPECT_STAGES") != nullptr;
    for (RawGroup& g : raw) g.stage = (g.x & ~mask0) ? -1 : 0;
    if (extra_stages) {
      for (;;) {
        // next map: bits 0..4 (coalescing) + the flips of as many unplaced groups as fit
        uint32_t want = 0x1fu;
        bool any = false;
        for (const RawGroup& g : raw) {
          if (g.stage >= 0) continue;
          if (__builtin_popcount(want | g.x) <= Tl) { want |= g.x; any = true; }
        }
        if (!any) break;
        for (int b = 0; b < hp_.n_eff && __builtin_popcount(want) < Tl; ++b) want |= 1u << b;
        std::vector<int> bits;
        for (int b = 0; b < hp_.n_eff; ++b) if ((want >> b) & 1) bits.push_back(b);
        const int st = (int)stage_bits_.size();
        stage_bits_.push_back(bits);
        for (RawGroup& g : raw) if (g.stage < 0 && !(g.x & ~want)) g.stage = st;
      }
    }
    for (RawGroup& g : raw) if (g.stage < 0) g.stage = 0;  // partner read across tiles
    // ---- emit per stage
    stage_ranges_.clear();
    for (int st = 0; st < (int)stage_bits_.size(); ++st) {
      const std::vector<int>& bits = stage_bits_[st];
      const uint32_t smask = multi ? mask_of(bits) : 0xffffffffu;
      std::vector<int> local_of(32, -1);
      for (size_t j = 0; j < bits.size(); ++j) local_of[bits[j]] = (int)j;
      const int mshift = (int)bits.size() - hp_.K;  // the thread's m-th amplitude: local index m << mshift | tid
      std::vector<uint32_t> mstate(1 << hp_.K, 0);  // its state-index contribution
      for (int m = 0; m < (1 << hp_.K); ++m)
        for (size_t j = 0; j < bits.size(); ++j)
          if ((((uint32_t)m << mshift) >> j) & 1) mstate[m] |= 1u << bits[j];
      std::vector<HCand> hcands;
      StageRange sr;
      sr.grp_begin = (int32_t)hp_.groups.size();
      sr.term_begin = (int32_t)hp_.terms.size();
      sr.rng_begin = (int32_t)hp_.opranges.size();
      for (int j = 0; j < o.n_ops(); ++j) {
        DevOpRange r;
        r.group_begin = (int32_t)hp_.groups.size();
        for (const RawGroup& rg : raw) {
          if (rg.op != j || rg.stage != st) continue;
          const uint32_t x = rg.x;
          if (hpass_ok && !(x & ~smask) && __builtin_popcount(x) <= 2) {
            bool real = true;
            for (int ti : rg.term_idx) real = real && !(__builtin_popcount(o.terms[ti].xmask & o.terms[ti].zmask) & 1);
            if (real) {
              HCand hc;
              hc.x = x;
              for (int ti : rg.term_idx) {
                const qhbm_pauli_term_t& t = o.terms[ti];
                const int ny = __builtin_popcount(t.xmask & t.zmask) & 3;
                hc.terms.push_back({ny == 0 ? t.coeff : -t.coeff, t.zmask});
              }
              hcands.push_back(hc);
              continue;
            }
          }
          DevTermGroup g;
          std::memset(&g, 0, sizeof(g));
          g.x = x;
          g.xl = -1;
          if (!(x & ~smask)) {
            g.xl = 0;
            for (int b = 0; b < 32; ++b) if ((x >> b) & 1) g.xl |= 1 << (multi ? local_of[b] : b);
          }
          g.term_begin = (int32_t)hp_.terms.size();
          for (int ti : rg.term_idx) {
            const qhbm_pauli_term_t& t = o.terms[ti];
            const int ny = __builtin_popcount(t.xmask & t.zmask) & 3;
            DevTerm d;
            d.kr = ny == 0 ? t.coeff : (ny == 2 ? -t.coeff : 0.f);
            d.ki = ny == 1 ? -t.coeff : (ny == 3 ? t.coeff : 0.f);
            d.z = t.zmask;
            d.mword = 0;
            for (int m = 0; m < (1 << hp_.K); ++m)
              if (__builtin_popcount(mstate[m] & t.zmask) & 1) d.mword |= 1u << m;
            if (d.ki != 0.f) g.is_complex = 1;
            if (t.zmask == 0) { g.k0r += d.kr; g.k0i += d.ki; }
            else hp_.terms.push_back(d);
          }
          g.term_end = (int32_t)hp_.terms.size();
          hp_.groups.push_back(g);
        }
        r.group_end = (int32_t)hp_.groups.size();
        r.op = j;
        r.pad = 0;
        // observables without generic groups cost nothing in the kernel; the single observable of a
        // stage with observable passes keeps its (possibly empty) range: the phase finishes E and lambda
        if (r.group_end > r.group_begin || !hcands.empty()) hp_.opranges.push_back(r);
      }
      sr.rng_end = (int32_t)hp_.opranges.size();
      sr.grp_end = (int32_t)hp_.groups.size();
      sr.term_end = (int32_t)hp_.terms.size();
      stage_ranges_.push_back(sr);
      build_hpasses(st, hcands);
    }
  }

  struct StageRange {
    int32_t grp_begin = 0, grp_end = 0, term_begin = 0, term_end = 0, rng_begin = 0, rng_end = 0;
  };

  struct HCand {
    uint32_t x = 0;                                  // state-index xor mask (0: diagonal terms)
    std::vector<std::pair<float, uint32_t>> terms;   // (real coefficient incl. the Y phases, z-mask)
  };

  // Observable passes of one expectation stage.  Off-diagonal candidates are packed greedily into sets
  // of four register qubits (chains of neighbouring pairs share qubits); with K = 5 a fifth, flip-free
  // register qubit pads the set.  Every diagonal term goes to the pass whose registers cover most of
  // its z-mask.  The part of a z-mask outside the registers becomes a per-thread sign (DevOp::aux0).
  // Register qubits are STATE bits here; DevPass stores them as tile-local bits of the stage's map.
  void build_hpasses(int stage, const std::vector<HCand>& cands) {
    if ((int)h_ranges_.size() <= stage) h_ranges_.resize(stage + 1, {0, 0});
    h_ranges_[stage] = {(int)hp_.passes.size(), (int)hp_.passes.size()};
    if (cands.empty()) return;
    const int K = hp_.K, R = 1 << K, Kx = 4;
    const std::vector<int>& tile_bits = stage_bits_[stage];
    std::vector<int> local_of(32, -1);
    for (size_t j = 0; j < tile_bits.size(); ++j) local_of[tile_bits[j]] = (int)j;
    struct HP {
      std::vector<int> regs;  // state bits; positions 0..3 may carry flips
      std::vector<int> offdiag;
    };
    std::vector<HP> hps;
    auto bits_of = [](uint32_t x) {
      std::vector<int> b;
      for (int i = 0; i < 32; ++i) if ((x >> i) & 1) b.push_back(i);
      return b;
    };
    std::vector<int> pending;
    for (size_t c = 0; c < cands.size(); ++c) if (cands[c].x != 0) pending.push_back((int)c);
    while (!pending.empty()) {
      HP p;
      auto missing = [&](int c) {
        int m = 0;
        for (int b : bits_of(cands[c].x)) m += std::find(p.regs.begin(), p.regs.end(), b) == p.regs.end();
        return m;
      };
      auto take = [&](size_t pi) {
        const int c = pending[pi];
        for (int b : bits_of(cands[c].x))
          if (std::find(p.regs.begin(), p.regs.end(), b) == p.regs.end()) p.regs.push_back(b);
        p.offdiag.push_back(c);
        pending.erase(pending.begin() + (long)pi);
      };
      size_t seed = 0;
      for (size_t pi = 0; pi < pending.size(); ++pi)
        if (__builtin_popcount(cands[pending[pi]].x) == 2) { seed = pi; break; }
      take(seed);
      for (;;) {
        // best = fewest new registers; among equals prefer one that shares a qubit with the set
        int best = -1, best_missing = 99, best_shared = -1;
        for (size_t pi = 0; pi < pending.size(); ++pi) {
          const int m = missing(pending[pi]);
          if ((int)p.regs.size() + m > Kx) continue;
          const int shared = __builtin_popcount(cands[pending[pi]].x) - m;
          if (m < best_missing || (m == best_missing && shared > best_shared)) {
            best = (int)pi; best_missing = m; best_shared = shared;
          }
        }
        if (best < 0) break;
        take((size_t)best);
      }
      hps.push_back(p);
    }
    int diag = -1;
    for (size_t c = 0; c < cands.size(); ++c) if (cands[c].x == 0) diag = (int)c;
    if (hps.empty()) hps.push_back(HP());
    for (HP& p : hps)  // pad the register set from the top of the tile
      for (int j = (int)tile_bits.size() - 1; j >= 0 && (int)p.regs.size() < K; --j)
        if (std::find(p.regs.begin(), p.regs.end(), tile_bits[j]) == p.regs.end()) p.regs.push_back(tile_bits[j]);
    std::vector<std::vector<std::pair<float, uint32_t>>> diag_of(hps.size());
    if (diag >= 0) {
      for (const auto& t : cands[diag].terms) {
        size_t best = 0;
        int cover = -1;
        for (size_t h = 0; h < hps.size(); ++h) {
          uint32_t rm = 0;
          for (int b : hps[h].regs) rm |= 1u << b;
          const int cv = __builtin_popcount(t.second & rm);
          if (cv > cover) { cover = cv; best = h; }
        }
        diag_of[best].push_back(t);
      }
    }
    for (size_t h = 0; h < hps.size(); ++h) {
      const HP& p = hps[h];
      DevPass ps;
      std::memset(&ps, 0, sizeof(ps));
      for (int j = 0; j < K; ++j) ps.regbit[j] = local_of[p.regs[j]];
      {
        std::vector<int> srt(ps.regbit, ps.regbit + K);
        std::sort(srt.begin(), srt.end());
        for (int j = 0; j < K; ++j) ps.sorted[j] = srt[j];
      }
      uint32_t regmask = 0;
      std::vector<uint32_t> rbits(R, 0);  // state-index bits set in register amplitude r
      for (int j = 0; j < K; ++j) regmask |= 1u << p.regs[j];
      for (int r = 0; r < (1 << kMaxRegQubits); ++r) {
        uint32_t dep = 0, sdep = 0;
        for (int j = 0; j < K; ++j) if ((r >> j) & 1) { dep |= 1u << ps.regbit[j]; sdep |= 1u << p.regs[j]; }
        if (r < R) rbits[r] = sdep;
        const uint32_t sw = (dep & ~15u) | ((dep ^ (dep >> 4) ^ (dep >> 8) ^ (dep >> 12)) & 15u);
        ps.eoff8[r] = r < R ? 8u * sw : 0u;
      }
      ps.op_begin = (int)hp_.ops.size();
      ps.gsym_off = 0;
      ps.ngrad = -1;
      ps.coef_begin = hp_.ncoef;
      // one table per (xor mask, out-of-register z part)
      auto emit = [&](int type, int xr, const std::vector<std::pair<float, uint32_t>>& terms) {
        std::vector<uint32_t> zouts;
        for (const auto& t : terms) {
          const uint32_t zo = t.second & ~regmask;
          if (std::find(zouts.begin(), zouts.end(), zo) == zouts.end()) zouts.push_back(zo);
        }
        for (uint32_t zo : zouts) {
          if ((int)hp_.ops.size() - ps.op_begin + 1 > kStageOps || hp_.ncoef - ps.coef_begin + R > kStageCoef) {
            // the staging buffer is full: continue in another pass over the same register set
            ps.op_end = (int)hp_.ops.size();
            ps.coef_end = hp_.ncoef;
            hp_.passes.push_back(ps);
            ps.op_begin = (int)hp_.ops.size();
            ps.coef_begin = hp_.ncoef;
          }
          std::vector<int32_t> tab(R, 0);
          for (int r = 0; r < R; ++r) {
            float v = 0.f;
            for (const auto& t : terms) {
              if ((t.second & ~regmask) != zo) continue;
              v += (__builtin_popcount(rbits[r] & t.second) & 1) ? -t.first : t.first;
            }
            std::memcpy(&tab[r], &v, sizeof(float));
          }
          DevOp op = make_op(type);
          op.p0 = xr;
          op.aux0 = (int32_t)zo;
          // table shape (see hx_apply): 1 = uniform, 2 = zero on even flip parity and uniform elsewhere
          op.aux1 = 0;
          {
            std::vector<float> tv(R);
            std::memcpy(tv.data(), tab.data(), sizeof(float) * R);
            bool uniform = true, flip = xr != 0;
            const float odd = tv[xr & -xr];
            for (int r = 0; r < R; ++r) {
              uniform = uniform && tv[r] == tv[0];
              const bool oddp = __builtin_popcount(r & xr) & 1;
              flip = flip && (oddp ? tv[r] == odd : tv[r] == 0.f);
            }
            if (uniform) op.aux1 = 1;
            else if (flip) op.aux1 = 2;
          }
          op.coef = alloc_coef(R);
          add_job(PJ_CONST, op.coef, 0, 0, 0, 0, tab);
          hp_.ops.push_back(op);
        }
      };
      for (int c : p.offdiag) {
        int xr = 0;
        for (int j = 0; j < Kx; ++j) if ((cands[c].x >> p.regs[j]) & 1) xr |= 1 << j;
        emit(OP_HX, xr, cands[c].terms);
      }
      if (!diag_of[h].empty()) emit(OP_HD, 0, diag_of[h]);
      ps.op_end = (int)hp_.ops.size();
      ps.coef_end = hp_.ncoef;
      if (ps.op_end > ps.op_begin) hp_.passes.push_back(ps);
    }
    h_ranges_[stage].second = (int)hp_.passes.size();
  }

  void compile(const CircuitIR& c, const OpsIR& o) {
    std::vector<Atom> fwd = forward_atoms();
    std::vector<SweepOut> fs, bs;
    schedule(fwd, false, fs);
    if (hp_.grad) {
      std::vector<Atom> bwd = backward_atoms();
      schedule(bwd, true, bs);
    }
    hp_.n_sweeps_fwd = (int)fs.size();
    hp_.n_sweeps_bwd = (int)bs.size();
    build_terms(o);
    hp_.phase_coef = -1;
    if (!phase_gates_.empty()) {
      hp_.phase_coef = alloc_coef(4);
      add_job(PJ_PHASE, hp_.phase_coef, 0, 0, 0, 0, std::vector<int32_t>(phase_gates_.begin(), phase_gates_.end()));
    }
    std::vector<int> contiguous;
    for (int b = 0; b < std::min(hp_.T, hp_.n_eff); ++b) contiguous.push_back(b);

    auto blank = [&]() {
      LaunchDesc L;
      std::memset(&L, 0, sizeof(L));
      return L;
    };
    auto set_stage = [&](LaunchDesc& L, int st) {
      L.expect_stage = st;
      L.pass_h_begin = h_ranges_[st].first;
      L.pass_h_end = h_ranges_[st].second;
      L.grp_begin = stage_ranges_[st].grp_begin;
      L.grp_end = stage_ranges_[st].grp_end;
      L.term_begin = stage_ranges_[st].term_begin;
      L.term_end = stage_ranges_[st].term_end;
      L.rng_begin = stage_ranges_[st].rng_begin;
      L.rng_end = stage_ranges_[st].rng_end;
    };
    if (hp_.n_eff <= hp_.T) {
      // whole state in one tile: one launch does forward, expectation and backward
      LaunchDesc L = blank();
      L.flags = LF_INIT_BASIS | LF_EXPECT;
      set_stage(L, 0);
      fill_runs(contiguous, L);
      L.pass_a_begin = fs.empty() ? 0 : fs.front().pass_begin;
      L.pass_a_end = fs.empty() ? 0 : fs.back().pass_end;
      L.pass_b_begin = bs.empty() ? 0 : bs.front().pass_begin;
      L.pass_b_end = bs.empty() ? 0 : bs.back().pass_end;
      hp_.launches.push_back(L);
      hp_.n_fwd_launches = 1;
      return;
    }
    if (fs.empty()) {
      SweepOut sw;
      sw.tile_bits = contiguous;
      sw.pass_begin = sw.pass_end = 0;
      fs.push_back(sw);
    }
    for (size_t i = 0; i < fs.size(); ++i) {
      LaunchDesc L = blank();
      L.flags = (i == 0 ? LF_INIT_BASIS : LF_LOAD_PSI) | LF_STORE_PSI;
      fill_runs(fs[i].tile_bits, L);
      L.pass_a_begin = fs[i].pass_begin;
      L.pass_a_end = fs[i].pass_end;
      hp_.launches.push_back(L);
    }
    hp_.n_fwd_launches = (int)fs.size();
    if (fs.size() >= 2 && std::getenv("QHBM_NO_SPARSE_INIT") == nullptr) {
      LaunchDesc& first = hp_.launches[hp_.launches.size() - fs.size()];
      LaunchDesc& second = hp_.launches[hp_.launches.size() - fs.size() + 1];
      first.flags |= LF_SPARSE_OUT;
      second.flags |= LF_SPARSE_IN;
      second.sparse_mask = ~first.tile_mask & (hp_.n_eff >= 32 ? 0xffffffffu : ((1u << hp_.n_eff) - 1u));
    }
    // The expectation phase needs the contiguous tile map.  When the first backward sweep uses the same
    // map it runs in the same launch: psi and lambda stay in shared memory instead of a round trip
    // through global memory.  psi then goes to the alternate buffer, because other tiles of this launch
    // may still read the final state across tiles (x-groups that flip out-of-tile qubits).
    const bool fuse = !bs.empty() && bs[0].tile_bits == contiguous && std::getenv("QHBM_NO_FUSE") == nullptr;
    {
      LaunchDesc L = blank();
      L.flags = LF_LOAD_PSI | LF_EXPECT | (hp_.grad ? LF_STORE_LAM : 0);
      set_stage(L, 0);
      fill_runs(contiguous, L);
      if (fuse) {
        L.pass_b_begin = bs[0].pass_begin;
        L.pass_b_end = bs[0].pass_end;
        L.flags = LF_LOAD_PSI | LF_EXPECT;
        if (bs.size() > 1) L.flags |= LF_STORE_PSI | LF_STORE_LAM | LF_PSI_ALT;
      }
      hp_.launches.push_back(L);
    }
    for (int st = 1; st < (int)stage_bits_.size(); ++st) {  // forward-only plans: the other tile maps
      LaunchDesc L = blank();
      L.flags = LF_LOAD_PSI | LF_EXPECT;
      set_stage(L, st);
      fill_runs(stage_bits_[st], L);
      hp_.launches.push_back(L);
    }
    for (size_t i = fuse ? 1 : 0; i < bs.size(); ++i) {
      LaunchDesc L = blank();
      L.flags = LF_LOAD_PSI | LF_LOAD_LAM;
      if (i + 1 < bs.size()) L.flags |= LF_STORE_PSI | LF_STORE_LAM;
      fill_runs(bs[i].tile_bits, L);
      L.pass_b_begin = bs[i].pass_begin;
      L.pass_b_end = bs[i].pass_end;
      hp_.launches.push_back(L);
    }
    (void)c;
  }

 private:
  HostPlan& hp_;
  std::vector<int> phase_gates_;  // forward X/Y powers whose global phase was dropped
  std::vector<std::pair<int, int>> h_ranges_;  // observable passes of every stage
  std::vector<std::vector<int>> stage_bits_;  // tile map of every expectation stage
  std::vector<StageRange> stage_ranges_;
  bool no_hpass_ = std::getenv("QHBM_NO_HPASS") != nullptr;  // development switch: generic tables only
  bool defer_tails_ = false;  // end a backward sweep instead of opening a pass of first-gate gradients only
  int units_ = 0;        // gradient scratch units committed by the pass being scheduled (flushed runs + float slots)
  int tasks_bound_ = 0;  // upper bound on the reduction tasks that pass will stage
};

}  // namespace


// ---------------------------------------------------------------------------------
// Device program.  The kernels execute `dev_passes` / `dev_ops`: the schedule above with the gradient
// bookkeeping of a pass turned into (i) scratch UNITS (one float per thread each) that the per-thread
// ops write, (ii) reduction TASKS staged behind the pass's ops, which sum those units over the CTA
// (complex marginal vectors: also over the subset of threads whose index has given bits set), and
// (iii) DESCRIPTORS, evaluated once per CTA at the end of a flush window, that combine the reduced sums
// into one float64 atomic per gradient slot.  A diagonal gate therefore costs no per-thread work
// beyond the run's shared marginal vectors: its 2 or 4 selected sums come from inclusion-exclusion
// over the task results (DevGradDesc).
// ---------------------------------------------------------------------------------
namespace {

struct BitRef {
  int kind;  // 0: register position, 1: thread-index bit, 2: state-index bit outside the tile
  int v;
};
struct Marg {
  int vec;        // marginal vector: 0 = T, 1 + p = S[p], 1 + K + pi = SS[pi]
  uint32_t mask;  // thread-index bits that must be set
  int cond_a, cond_b;
};

void lower_device_program(HostPlan& hp) {
  const int K = hp.K;
  const int npass = (int)hp.passes.size();
  std::vector<int> launch_of(npass, -1);
  for (size_t li = 0; li < hp.launches.size(); ++li)
    for (int p = hp.launches[li].pass_b_begin; p < hp.launches[li].pass_b_end; ++p) launch_of[p] = (int)li;
  hp.dev_passes = hp.passes;
  hp.dev_ops.clear();
  hp.gdescs.clear();
  std::vector<int> pass_floats(npass, 0);                 // reduced floats the pass's tasks produce
  std::vector<std::pair<int, int>> pass_descs(npass, {0, 0});
  for (int p = 0; p < npass; ++p) {
    const DevPass& ps = hp.passes[p];
    DevPass& dp = hp.dev_passes[p];
    dp.op_begin = (int32_t)hp.dev_ops.size();
    dp.gd_flush_begin = dp.gd_flush_end = 0;
    pass_descs[p] = {(int)hp.gdescs.size(), (int)hp.gdescs.size()};
    if (ps.ngrad <= 0 || launch_of[p] < 0) {
      for (int i = ps.op_begin; i < ps.op_end; ++i) hp.dev_ops.push_back(pack_op(hp.ops[i]));
      dp.exec_end = dp.op_end = (int32_t)hp.dev_ops.size();
      continue;
    }
    const LaunchDesc& L = hp.launches[launch_of[p]];
    std::vector<int> local_of(32, -1);
    for (int r = 0; r < L.n_runs; ++r)
      for (int i = 0; i < L.runs[r].len; ++i) local_of[L.runs[r].global_start + i] = L.runs[r].local_start + i;
    uint32_t regmask = 0;
    for (int j = 0; j < K; ++j) regmask |= 1u << ps.regbit[j];
    auto ref_of = [&](int state_bit) -> BitRef {
      const int l = local_of[state_bit];
      if (l < 0) return {2, state_bit};
      if ((regmask >> l) & 1) throw std::runtime_error("internal: a thread-constant bit is a register bit");
      return {1, l - __builtin_popcount(regmask & ((1u << l) - 1u))};
    };
    auto marg1 = [&](const BitRef& a) -> Marg {
      if (a.kind == 0) return {1 + a.v, 0u, -1, -1};
      if (a.kind == 1) return {0, 1u << a.v, -1, -1};
      return {0, 0u, a.v, -1};
    };
    auto marg2 = [&](const BitRef& a, const BitRef& b) -> Marg {
      Marg m{0, 0u, -1, -1};
      if (a.kind == 0 && b.kind == 0) {
        const int ph = std::max(a.v, b.v), pl = std::min(a.v, b.v);
        m.vec = 1 + K + ph * (ph - 1) / 2 + pl;
      } else if (a.kind == 0) m.vec = 1 + a.v;
      else if (b.kind == 0) m.vec = 1 + b.v;
      if (a.kind == 1) m.mask |= 1u << a.v;
      if (b.kind == 1) m.mask |= 1u << b.v;
      if (a.kind == 2) m.cond_a = a.v;
      if (b.kind == 2) m.cond_b = b.v;
      return m;
    };
    int units = 0, floats = 0;
    std::vector<PackedOp> tasks;
    auto float_task = [&](int sym) {  // one float per thread in the next unit -> gacc[floats]
      const int unit = units++;
      DevOp t = make_op(OP_TASK_F);
      t.p0 = unit;
      t.coef = floats;
      t.aux0 = t.aux1 = 0;
      tasks.push_back(pack_op(t));
      DevGradDesc d;
      std::memset(&d, 0, sizeof(d));
      d.kind = 0;
      d.sym = sym;
      d.coef = -1;
      d.i_tot = (int16_t)floats;
      d.i_a = d.i_b = d.i_ab = 0;
      d.cond_a = d.cond_b = -1;
      hp.gdescs.push_back(d);
      floats += 1;
      return unit;
    };
    for (int i = ps.op_begin; i < ps.op_end; ++i) {
      DevOp o = hp.ops[i];
      switch (o.type) {
        case OP_XROTM: case OP_YROTM: case OP_XROTF: {
          int aux1 = 0, p1 = o.p1;
          for (int P = 0; P < K; ++P) {
            if (!(o.aux0 & (1 << P))) continue;
            const int slot = P < 4 ? ((o.aux1 >> (8 * P)) & 0xff) : o.p1;
            const int unit = float_task(hp.gsym[ps.gsym_off + slot]);
            if (P < 4) aux1 |= unit << (8 * P);
            else p1 = unit;
          }
          if (o.aux0) { o.aux1 = aux1; o.p1 = p1; }
          hp.dev_ops.push_back(pack_op(o));
        } break;
        case OP_XROT: case OP_YROT: case OP_GRAD_X: case OP_GRAD_Y:
          throw std::runtime_error("internal: unmerged rotation op in a device program");
        case OP_GRAD_MAT1: case OP_GRAD_MAT2:
          o.gslot = float_task(hp.gsym[ps.gsym_off + o.gslot]);
          hp.dev_ops.push_back(pack_op(o));
          break;
        case OP_GD_BEGIN: {
          const int count = o.aux0;
          uint32_t vmask = 1u;
          std::vector<std::pair<int, std::pair<BitRef, BitRef>>> gates;  // (kind, (A, B))
          for (int q = 1; q <= count; ++q) {
            const DevOp& g = hp.ops[i + q];
            BitRef a{0, 0}, b{0, 0};
            int kind = 2;
            switch (g.type) {
              case OP_GD_CONST:
                a = ref_of(g.aux0);
                if (g.aux1 >= 0) b = ref_of(g.aux1); else kind = 1;
                break;
              case OP_GD_REG1: a = {0, g.p0}; kind = 1; break;
              case OP_GD_REG2: a = {0, g.p0}; b = {0, g.p1}; break;
              case OP_GD_MIX: a = ref_of(g.aux0); b = {0, g.p0}; break;
              default: throw std::runtime_error("internal: malformed diagonal-gradient run");
            }
            gates.push_back({kind, {a, b}});
            vmask |= 1u << marg1(a).vec;
            if (kind == 2) { vmask |= 1u << marg1(b).vec; vmask |= 1u << marg2(a, b).vec; }
          }
          const int unit0 = units;
          units += 2 * __builtin_popcount(vmask);
          std::vector<std::pair<std::pair<int, uint32_t>, int>> seen;  // (vector, mask) -> gacc offset
          auto task_of = [&](const Marg& m) {
            for (const auto& s : seen) if (s.first.first == m.vec && s.first.second == m.mask) return s.second;
            DevOp t = make_op(OP_TASK_C);
            t.p0 = unit0 + 2 * __builtin_popcount(vmask & ((1u << m.vec) - 1u));
            t.coef = floats;
            t.aux0 = (int32_t)m.mask;
            t.aux1 = 0;
            tasks.push_back(pack_op(t));
            seen.push_back({{m.vec, m.mask}, floats});
            floats += 2;
            return floats - 2;
          };
          for (int q = 1; q <= count; ++q) {
            const DevOp& g = hp.ops[i + q];
            const auto& gt = gates[q - 1];
            DevGradDesc d;
            std::memset(&d, 0, sizeof(d));
            d.kind = gt.first;
            d.sym = hp.gsym[ps.gsym_off + g.gslot];
            d.coef = g.coef;
            d.i_tot = (int16_t)task_of(Marg{0, 0u, -1, -1});
            const Marg ma = marg1(gt.second.first);
            d.i_a = (int16_t)task_of(ma);
            d.cond_a = (int8_t)ma.cond_a;
            d.cond_b = -1;
            if (gt.first == 2) {
              const Marg mb = marg1(gt.second.second);
              const Marg mab = marg2(gt.second.first, gt.second.second);
              d.i_b = (int16_t)task_of(mb);
              d.i_ab = (int16_t)task_of(mab);
              d.cond_b = (int8_t)mb.cond_a;
            }
            hp.gdescs.push_back(d);
          }
          o.coef = (int32_t)vmask;
          o.gslot = unit0;
          hp.dev_ops.push_back(pack_op(o));
          for (int q = 1; q <= count; ++q) hp.dev_ops.push_back(pack_op(hp.ops[i + q]));  // skipped by the kernel
          i += count;
        } break;
        default:
          hp.dev_ops.push_back(pack_op(o));
          break;
      }
    }
    if (units > std::min(4 * (1 << K), 255))
      throw std::runtime_error("internal: a pass needs more gradient scratch than the tiles hold");
    dp.exec_end = (int32_t)hp.dev_ops.size();
    hp.dev_ops.insert(hp.dev_ops.end(), tasks.begin(), tasks.end());
    dp.op_end = (int32_t)hp.dev_ops.size();
    if (dp.op_end - dp.op_begin > kStageOpsAdj)
      throw std::runtime_error("internal: a gradient pass program exceeds the staging buffer");
    if (floats > kGaccFloats) throw std::runtime_error("internal: a pass reduces more sums than the flush window holds");
    pass_floats[p] = floats;
    pass_descs[p].second = (int)hp.gdescs.size();
  }
  for (int p = 0; p < npass; ++p) {
    const bool has_next = p + 1 < npass;
    hp.dev_passes[p].next_op_end = has_next ? hp.dev_passes[p + 1].op_end : hp.dev_passes[p].op_end;
  }
  hp.lean = true;
  for (const PackedOp& q : hp.dev_ops) {
    const int t = (int)(q.w0 & 0xffu);
    if (t == OP_MAT1 || t == OP_MAT2 || t == OP_GRAD_MAT1 || t == OP_GRAD_MAT2 || t == OP_YROTM) hp.lean = false;
  }
  // Flush windows: consecutive gradient passes of a launch share the shared-memory sums until they would
  // overflow; the last pass of a window evaluates the window's descriptors.
  for (const LaunchDesc& L : hp.launches) {
    int p = L.pass_b_begin;
    while (p < L.pass_b_end) {
      int e = p, used = 0;
      while (e < L.pass_b_end && used + pass_floats[e] <= kGaccFloats) {
        // rebase pass e's task outputs and descriptor indices into the window
        DevPass& dp = hp.dev_passes[e];
        for (int t = dp.exec_end; t < dp.op_end; ++t) hp.dev_ops[t].coef += used;
        for (int d = pass_descs[e].first; d < pass_descs[e].second; ++d) {
          DevGradDesc& g = hp.gdescs[d];
          g.i_tot = (int16_t)(g.i_tot + used);
          if (g.kind >= 1) g.i_a = (int16_t)(g.i_a + used);
          if (g.kind == 2) { g.i_b = (int16_t)(g.i_b + used); g.i_ab = (int16_t)(g.i_ab + used); }
        }
        used += pass_floats[e];
        ++e;
      }
      DevPass& last = hp.dev_passes[e - 1];
      last.gd_flush_begin = pass_descs[p].first;
      last.gd_flush_end = pass_descs[e - 1].second;
      p = e;
    }
  }
}

}  // namespace

static HostPlan compile_plan_variant(const CircuitIR& c, const OpsIR& o, bool with_gradient, int tile_qubits,
                                     int reg_qubits, bool defer_tails);

// The scheduler has one heuristic whose benefit depends on the circuit: deferring passes that would only take
// gradients of the circuit's first gates to the next backward sweep (Compiler::defer_tails_).  It saves a pass
// when that sweep has register positions to spare (HEA(16, 2): 15 -> 14 passes) and costs a whole extra sweep
// when it has not (HEA + inverse HEA).  Scheduling is milliseconds of host work, so both variants are
// compiled and the one with fewer launches, then fewer passes, is kept.
HostPlan compile_plan(const CircuitIR& c, const OpsIR& o, bool with_gradient, int tile_qubits,
                      int reg_qubits) {
  HostPlan plain = compile_plan_variant(c, o, with_gradient, tile_qubits, reg_qubits, false);
  if (!with_gradient || plain.tiles() == 1 || std::getenv("QHBM_NO_TAIL_DEFER") != nullptr) return plain;
  HostPlan deferred = compile_plan_variant(c, o, with_gradient, tile_qubits, reg_qubits, true);
  const bool better = deferred.launches.size() < plain.launches.size() ||
                      (deferred.launches.size() == plain.launches.size() && deferred.passes.size() < plain.passes.size());
  return better ? deferred : plain;
}

static HostPlan compile_plan_variant(const CircuitIR& c, const OpsIR& o, bool with_gradient, int tile_qubits,
                                     int reg_qubits, bool defer_tails) {
  validate_circuit(c);
  validate_ops(o);
  if (c.n_qubits != o.n_qubits) throw std::runtime_error("circuit and observables act on different qubit counts");
  HostPlan hp;
  hp.n = c.n_qubits;
  hp.P = c.n_symbols;
  hp.O = o.n_ops();
  hp.grad = with_gradient;
  hp.K = reg_qubits > 0 ? reg_qubits : (with_gradient ? 4 : 5);
  if (hp.K < 2 || hp.K > kMaxRegQubits) throw std::runtime_error("reg_qubits must be in [2, 5]");
  hp.n_eff = std::max(hp.n, hp.K + 5);
  // Defaults from measurement on B200 (profiles/): a state that fits one tile stays in shared memory
  // for the whole computation; larger states use 2^12-amplitude tiles for the adjoint (psi + lambda =
  // 64 KiB, two CTAs per SM) and 2^13 for forward-only sweeps (64 KiB, three CTAs per SM).
  const int t_max = with_gradient ? 13 : 14;  // 2 tiles (psi, lambda) of 8*2^T bytes must fit 227 KB
  int T = tile_qubits;
  if (T <= 0) T = hp.n_eff <= std::min(t_max, hp.K + 9) ? hp.n_eff : (with_gradient ? 12 : 13);
  if (T > t_max) throw std::runtime_error("tile_qubits too large for shared memory");
  if (T < hp.K + 5) throw std::runtime_error("tile_qubits must be >= reg_qubits + 5");
  if (T - hp.K > 9) throw std::runtime_error("tile_qubits - reg_qubits must be <= 9 (512 threads per CTA)");
  if (hp.K != 4 && hp.K != 5) throw std::runtime_error("reg_qubits must be 4 or 5");
  hp.T = std::min(T, hp.n_eff);
  hp.gates = c.gates;
  Compiler comp(hp, defer_tails);
  comp.compile(c, o);
  // Prefetch links and gradient-slot ranges.  Inside a launch's pass range the programs are consecutive
  // (pass i + 1 starts where pass i ends), which is what lets the kernel prefetch the next program from
  // the staged descriptor of the current one.
  for (size_t i = 0; i < hp.passes.size(); ++i) {
    const bool has_next = i + 1 < hp.passes.size();
    hp.passes[i].next_op_end = has_next ? hp.passes[i + 1].op_end : hp.passes[i].op_end;
    hp.passes[i].next_coef_end = has_next ? hp.passes[i + 1].coef_end : hp.passes[i].coef_end;
  }
  auto check_range = [&](int b, int e) {
    for (int i = b; i + 1 < e; ++i)
      if (hp.passes[i + 1].op_begin != hp.passes[i].op_end || hp.passes[i + 1].coef_begin != hp.passes[i].coef_end)
        throw std::runtime_error("internal: pass programs of a launch range are not consecutive");
  };
  for (LaunchDesc& L : hp.launches) {
    check_range(L.pass_a_begin, L.pass_a_end);
    check_range(L.pass_h_begin, L.pass_h_end);
    check_range(L.pass_b_begin, L.pass_b_end);
    L.gslot_begin = L.gslot_count = 0;
    if (L.pass_b_end > L.pass_b_begin) {
      const DevPass& first = hp.passes[L.pass_b_begin];
      const DevPass& last = hp.passes[L.pass_b_end - 1];
      L.gslot_begin = first.gsym_off;
      L.gslot_count = last.gsym_off + std::max(last.ngrad, 0) - first.gsym_off;
    }
  }
  lower_device_program(hp);
  return hp;
}

}  // namespace qhbm
