// Host-side symbolic analysis for the sparse AC path (structure-aware
// factorisation with symbolic reuse across the batch).
//
// Every point of an AC batch shares one sparsity pattern and, in practice, one pivot
// sequence.  A *pilot* factorisation of one representative point — the reference's own
// algorithm (lib/math/solveComplex.ts:15-53: partial pivoting, first maximum wins, row
// swaps) run once on the host — yields that sequence.  With the sequence fixed, the fill
// pattern is purely symbolic, so the whole elimination and back-substitution becomes a
// straight-line program over value slots:
//
//   PIVOT  k: candidate operands (rows with a structural non-zero in column k, in the
//             reference's logical row order), which of them the pilot chose, and the slot
//             that receives 1/pivot
//   ELIM   i: a_ij <- a_ij - f * a_kj for the structural non-zeros j of the pivot row
//   BSUB   i: x_i = (b_i - sum U_ij x_j) / u_ii
//
// The device executes the program with one thread per system and VERIFIES the pivot of
// every step against the reference's rule on that system's own numbers; a system whose
// pivot choice differs from the pilot's is handed to the dense pivoting kernel.  Results
// are therefore those of partial pivoting for every system, never of a frozen ordering.
//
// Two optimisations keep the per-thread workspace small (it is what the kernel's HBM
// traffic consists of besides the results):
//  * lazy stamping — an operand that still holds its stamped value is encoded as a
//    *pristine entry* (negative operand) and recomputed on the device from the entry's
//    per-topology constants instead of being written to and re-read from the workspace;
//  * liveness-based slot allocation — the program is first emitted over single-assignment
//    virtual slots, then physical slots are assigned by a linear scan with a LIFO free list,
//    so short-lived values overwrite each other in cache and only the values that really
//    survive until the back-substitution (1/u_kk, the eliminated right-hand side, U) occupy
//    memory.
#pragma once
#include <climits>
#include <cmath>
#include <complex>
#include <cstdint>
#include <vector>

namespace spicey {

enum SparseOp { SOP_PIVOT = 1, SOP_ELIM = 2, SOP_BSUB = 3, SOP_END = 4 };
constexpr int kNoOperand = INT_MIN;  // "structurally zero" operand

struct SparseProgram {
  bool ok = false;
  int n = 0;
  int n_stamp = 0;   // gather-plan entries (pristine operands index these)
  int n_slots = 0;   // physical workspace slots per system
  int n_virtual = 0; // single-assignment values before allocation (statistics)
  long long n_fma = 0, n_div = 0;  // executed complex FMAs / reciprocals per system
  std::vector<int> code;     // program words, then nothing else
  std::vector<int> x_slot;   // [n] physical slot of x_i at the end of the program
  // per stamped entry: value = (alpha + j*aim0) + j*(omega*beta - gamma/omega)
  std::vector<double> ent_alpha, ent_beta, ent_gamma, ent_jre, ent_jim;
  // per AC element (R,C,L,V order): current = Y_e * (v1 - v2), Y_e = ya + j*(omega*yb - yg/omega)
  std::vector<double> el_a, el_b, el_g;
  std::vector<double> ind_L;  // inductances, for the per-point guards of simulateAC.ts:47-51
};

struct PilotInput {
  int n = 0;                                   // nvar
  const std::vector<int>* row_ptr = nullptr;   // gather plan (AC)
  const std::vector<int>* ent_col = nullptr;
  std::vector<std::complex<double>> ent_val;   // numeric pilot matrix entries, plan order
};

namespace sparse_detail {

// Operand of the intermediate program: >= 0 virtual slot, < 0 pristine entry ~idx, kNoOperand none.
struct Update { int dst_old, dst_new, src; };
struct IrOp {
  int kind = 0;
  std::vector<int> reads;       // PIVOT: candidates; ELIM: {a_ik}; BSUB: {b, rcp, a_0, x_0, a_1, x_1, ...}
  int pidx = 0;                 // PIVOT
  int def = -1;                 // PIVOT: 1/pivot; BSUB: x_i
  int var = 0;                  // BSUB: i
  std::vector<Update> upd;      // ELIM
};

}  // namespace sparse_detail

// Builds the program from the pilot point.  Returns ok=false when the pilot itself is
// singular / hits the divide guard (the dense kernel then reports the exact status).
inline void build_sparse_program(const PilotInput& in, SparseProgram& sp) {
  using namespace sparse_detail;
  typedef std::complex<double> cd;
  const int n = in.n, ld = n + 1;
  const double EPS = 1e-15;
  sp.ok = false;
  sp.n = n;
  sp.code.clear();
  std::vector<cd> M((size_t)n * ld, cd(0, 0));
  // cur[r][c]: operand currently holding A(r,c): kNoOperand = structural zero
  std::vector<int> cur((size_t)n * ld, kNoOperand);
  int n_ent = 0;
  for (int r = 0; r < n; ++r)
    for (int en = (*in.row_ptr)[r]; en < (*in.row_ptr)[r + 1]; ++en) {
      int c = (*in.ent_col)[en];
      M[(size_t)r * ld + c] = in.ent_val[en];
      cur[(size_t)r * ld + c] = ~en;  // pristine
      ++n_ent;
    }
  sp.n_stamp = n_ent;
  int nv = 0;  // virtual slots
  std::vector<int> rows(n);  // logical position -> physical row (the reference's swapped row array)
  for (int i = 0; i < n; ++i) rows[i] = i;
  std::vector<int> rcp_slot(n, -1);
  std::vector<IrOp> ir;
  sp.n_fma = sp.n_div = 0;

  for (int k = 0; k < n; ++k) {
    // numeric pivot of the pilot, reference rule (solveComplex.ts:18-28)
    int imax = k;
    double vmax = std::hypot(M[(size_t)rows[k] * ld + k].real(), M[(size_t)rows[k] * ld + k].imag());
    for (int i = k + 1; i < n; ++i) {
      const cd& z = M[(size_t)rows[i] * ld + k];
      double v = std::hypot(z.real(), z.imag());
      if (v > vmax) { vmax = v; imax = i; }
    }
    if (vmax < EPS) return;                                         // pilot singular
    if (cur[(size_t)rows[imax] * ld + k] == kNoOperand) return;     // cannot happen
    // candidates in logical order BEFORE the swap (the order the reference scans)
    std::vector<int> cand_row;
    IrOp pv;
    pv.kind = SOP_PIVOT;
    for (int i = k; i < n; ++i)
      if (cur[(size_t)rows[i] * ld + k] != kNoOperand) {
        if (i == imax) pv.pidx = (int)cand_row.size();
        cand_row.push_back(rows[i]);
        pv.reads.push_back(cur[(size_t)rows[i] * ld + k]);
      }
    std::swap(rows[k], rows[imax]);
    const int p = rows[k];
    const cd pivot = M[(size_t)p * ld + k];
    if (std::norm(pivot) < EPS) return;                             // Complex.div guard on the pilot
    pv.def = nv++;
    rcp_slot[k] = pv.def;
    ir.push_back(pv);
    sp.n_div++;
    std::vector<int> pcols;  // structural non-zeros of the pivot row right of k (incl. rhs column n)
    for (int j = k + 1; j <= n; ++j)
      if (cur[(size_t)p * ld + j] != kNoOperand) pcols.push_back(j);
    for (int r : cand_row) {
      if (r == p) continue;
      IrOp el;
      el.kind = SOP_ELIM;
      el.reads.push_back(cur[(size_t)r * ld + k]);
      const cd f = M[(size_t)r * ld + k] / pivot;
      const bool act = !(std::abs(f) < EPS);
      for (int j : pcols) {
        Update u;
        u.dst_old = cur[(size_t)r * ld + j];
        u.dst_new = nv++;
        u.src = cur[(size_t)p * ld + j];
        cur[(size_t)r * ld + j] = u.dst_new;
        el.upd.push_back(u);
        if (act) M[(size_t)r * ld + j] -= f * M[(size_t)p * ld + j];
        sp.n_fma++;
      }
      cur[(size_t)r * ld + k] = kNoOperand;  // eliminated
      ir.push_back(el);
    }
  }
  std::vector<int> x_virtual(n, -1);
  for (int i = n - 1; i >= 0; --i) {
    const int r = rows[i];
    IrOp bs;
    bs.kind = SOP_BSUB;
    bs.var = i;
    bs.reads.push_back(cur[(size_t)r * ld + n]);  // rhs (may be kNoOperand)
    bs.reads.push_back(rcp_slot[i]);
    for (int j = i + 1; j < n; ++j)
      if (cur[(size_t)r * ld + j] != kNoOperand) {
        bs.reads.push_back(cur[(size_t)r * ld + j]);
        bs.reads.push_back(x_virtual[j]);
        sp.n_fma++;
      }
    bs.def = nv++;
    x_virtual[i] = bs.def;
    ir.push_back(bs);
  }
  sp.n_virtual = nv;

  // ---- liveness: last reader of every virtual slot (micro-op granularity) ----
  // Micro-op index: PIVOT = 1, ELIM = 1 (header) + one per update, BSUB = 1.
  std::vector<int> last(nv, -1);
  {
    int t = 0;
    auto use = [&](int o, int when) { if (o >= 0) last[o] = when; };
    for (const IrOp& op : ir) {
      if (op.kind == SOP_ELIM) {
        use(op.reads[0], t);
        ++t;
        for (const Update& u : op.upd) { use(u.dst_old, t); use(u.src, t); ++t; }
      } else {
        for (int o : op.reads) use(o, t);
        ++t;
      }
    }
    for (int i = 0; i < n; ++i) last[x_virtual[i]] = INT_MAX;  // x is read by the unpack phase
  }
  // ---- linear-scan allocation with a LIFO free list ----
  std::vector<int> phys(nv, -1), free_list;
  int high = 0;
  auto alloc = [&]() { if (!free_list.empty()) { int s = free_list.back(); free_list.pop_back(); return s; } return high++; };
  auto release = [&](int o, int now) {
    if (o >= 0 && last[o] == now && phys[o] >= 0) { free_list.push_back(phys[o]); last[o] = -2; }
  };
  auto P = [&](int o) { return o >= 0 ? phys[o] : o; };  // operand -> physical encoding
  {
    int t = 0;
    for (const IrOp& op : ir) {
      if (op.kind == SOP_PIVOT) {
        sp.code.push_back(SOP_PIVOT);
        sp.code.push_back((int)op.reads.size());
        sp.code.push_back(op.pidx);
        for (int o : op.reads) sp.code.push_back(P(o));
        for (int o : op.reads) release(o, t);
        phys[op.def] = alloc();
        sp.code.push_back(phys[op.def]);
        ++t;
      } else if (op.kind == SOP_ELIM) {
        sp.code.push_back(SOP_ELIM);
        sp.code.push_back(P(op.reads[0]));
        sp.code.push_back((int)op.upd.size());
        release(op.reads[0], t);
        ++t;
        for (const Update& u : op.upd) {
          const int eo = P(u.dst_old), es = P(u.src);
          release(u.dst_old, t);   // in-place update when the old version dies here
          release(u.src, t);
          phys[u.dst_new] = alloc();
          sp.code.push_back(eo);
          sp.code.push_back(es);
          sp.code.push_back(phys[u.dst_new]);
          ++t;
        }
      } else {
        sp.code.push_back(SOP_BSUB);
        sp.code.push_back(op.var);
        sp.code.push_back(P(op.reads[0]));
        sp.code.push_back(P(op.reads[1]));
        sp.code.push_back((int)(op.reads.size() - 2) / 2);
        for (size_t q = 2; q < op.reads.size(); ++q) sp.code.push_back(P(op.reads[q]));
        for (int o : op.reads) release(o, t);
        phys[op.def] = alloc();
        sp.code.push_back(phys[op.def]);
        ++t;
      }
    }
  }
  sp.code.push_back(SOP_END);
  sp.n_slots = high;
  sp.x_slot.resize(n);
  for (int i = 0; i < n; ++i) sp.x_slot[i] = phys[x_virtual[i]];
  sp.ok = true;
}

}  // namespace spicey
