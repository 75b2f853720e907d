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

enum SparseOp { SOP_PIVOT = 1, SOP_ELIM = 2, SOP_BSUB = 3, SOP_END = 4 };  // intermediate program

// Device micro-ops: one 16-byte word {hdr, A, B, C} each (a single uniform 128-bit load, prefetched one
// micro-op ahead).  hdr = opcode | kindA << 4 | kindB << 6 | flag << 8 | kindC << 10; operand kinds:
// 1 global workspace slot, 2 pristine stamped entry (lazy stamping), 3 fast slot (shared memory; one
// fast slot permanently holds 0 and stands for structural zeros).  A/B/C hold slot numbers (the
// uploader turns them into pool-tagged, stride-scaled offsets) or entry indices.
enum MicroOp {
  MOP_END = 0,
  MOP_PIVHEAD = 1,   // A = pilot's pivot candidate:           ap = A, mp = |ap|^2
  MOP_CAND = 2,      // A = another candidate, flag = scanned before the pivot:  ok &= flag ? m < mp : !(m > mp)
  MOP_PIVEND = 3,    // C = slot of 1/pivot:                   verify, r = 1/ap, W[C] = r
  MOP_ELIM = 4,      // A = a_ik:                              f = A * r (zeroed when |f| < EPS)
  MOP_UPD = 5,       // A = a_ij (old), B = a_kj, C = a_ij:    W[C] = A - f * B
  MOP_BHEAD = 6,     // A = b_i, B = 1/u_ii:                   acc = A, rc = B
  MOP_BTERM = 7,     // A = u_ij, B = x_j:                     acc -= A * B
  MOP_BEND = 8       // C = slot of x_i, A = i:                W[C] = acc * rc, result[i] = W[C]
};
struct MicroWord { int hdr, a, b, c; };
constexpr int kNoOperand = INT_MIN;  // "structurally zero" operand

namespace sparse_detail {

// Operand of the intermediate program: >= 0 virtual slot, < 0 pristine entry ~idx, kNoOperand none.
struct Update { int dst_old, dst_new, src; int col = 0; };   // col: column of the updated entry (n = right-hand side)
struct IrOp {
  int kind = 0;
  std::vector<int> reads;       // PIVOT: candidates; ELIM: {a_ik}; BSUB: {b, rcp, a_0, x_0, a_1, x_1, ...}
  int pidx = 0;                 // PIVOT
  int def = -1;                 // PIVOT: 1/pivot; BSUB: x_i
  int var = 0;                  // BSUB: i
  std::vector<Update> upd;      // ELIM
  bool active = true;           // ELIM: false when the pilot skipped the row (|f| < EPS, solveComplex.ts:46)
};

}  // namespace sparse_detail

struct SparseProgram {
  bool ok = false;
  int n = 0;
  int n_stamp = 0;   // gather-plan entries (pristine operands index these)
  int n_slots = 0;   // global workspace slots per system
  int n_fast = 0;    // fast (shared-memory) slots per system, the first n_const of them hold constants
  int n_const = 0;   // distinct pristine values materialised once per system (0 = computed where used)
  std::vector<int> const_entry;  // [n_const] a stamped entry whose value the constant slot holds
  int n_virtual = 0; // single-assignment values before allocation (statistics)
  long long n_fma = 0, n_div = 0;  // executed complex FMAs / reciprocals per system
  std::vector<MicroWord> code;  // micro-ops, terminated by MOP_END (+ one pad word for the prefetch)
  std::vector<sparse_detail::IrOp> ir;  // single-assignment intermediate program (input of the code generator)
  std::vector<int> x_virtual;           // [n] virtual slot of x_i
  std::vector<int> x_slot;      // [n] global slot of x_i at the end of the program
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


// Builds the program from the pilot point.  Returns ok=false when the pilot itself is
// singular / hits the divide guard (the dense kernel then reports the exact status).
inline void build_sparse_program(const PilotInput& in, SparseProgram& sp, int fast_slots = 12,
                                 const std::vector<int>* entry_class = nullptr, int n_class = 0) {
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
      el.active = act;
      for (int j : pcols) {
        Update u;
        u.dst_old = cur[(size_t)r * ld + j];
        u.dst_new = nv++;
        u.src = cur[(size_t)p * ld + j];
        u.col = j;
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
  sp.ir = ir;
  sp.x_virtual = x_virtual;

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
  // ---- definition time of every virtual slot (for lifetimes) ----
  std::vector<int> deft(nv, 0);
  {
    int t = 0;
    for (const IrOp& op : ir) {
      if (op.kind == SOP_ELIM) {
        ++t;
        for (const Update& u : op.upd) { deft[u.dst_new] = t; ++t; }
      } else {
        deft[op.def] = t;
        ++t;
      }
    }
  }
  // ---- two-pool linear-scan allocation, LIFO free lists ----
  // Fast pool (shared memory, `fast_slots` per system): constants first (distinct pristine values, when
  // few enough), then values whose lifetime is short.  Slow pool: global workspace.
  // x_i always lives in the global pool (the unpack phase gathers it from there).
  const int kWindow = 96;
  int n_const = 0;
  if (entry_class && n_class > 0 && n_class <= (fast_slots - 1) / 2) n_const = n_class;
  sp.n_const = n_const;
  sp.const_entry.assign(n_const, -1);
  if (n_const)
    for (int en = 0; en < n_ent; ++en)
      if (sp.const_entry[(*entry_class)[en]] < 0) sp.const_entry[(*entry_class)[en]] = en;
  std::vector<int> phys(nv, -1), pool(nv, 0), free_fast, free_slow;
  const int zero_slot = n_const;            // fast slot that always holds 0 (structural zeros)
  int high_fast = n_const + 1, high_slow = 0;
  std::vector<char> is_x(nv, 0);
  for (int i = 0; i < n; ++i) is_x[x_virtual[i]] = 1;
  auto alloc = [&](int v) {
    const bool want_fast = !is_x[v] && last[v] >= 0 && last[v] != INT_MAX && last[v] - deft[v] <= kWindow;
    if (want_fast) {
      if (!free_fast.empty()) { phys[v] = free_fast.back(); free_fast.pop_back(); pool[v] = 3; return; }
      if (high_fast < fast_slots) { phys[v] = high_fast++; pool[v] = 3; return; }
    }
    pool[v] = 1;
    if (!free_slow.empty()) { phys[v] = free_slow.back(); free_slow.pop_back(); return; }
    phys[v] = high_slow++;
  };
  auto release = [&](int o, int now) {
    if (o >= 0 && last[o] == now && phys[o] >= 0) {
      (pool[o] == 3 ? free_fast : free_slow).push_back(phys[o]);
      last[o] = -2;
    }
  };
  // operand -> (kind, value): 1 global slot, 3 fast slot, 2 pristine entry (or its constant slot), 0 zero
  auto K = [&](int o) {
    if (o >= 0) return pool[o];
    if (o == kNoOperand) return 3;          // the zero slot
    return n_const ? 3 : 2;
  };
  auto V = [&](int o) {
    if (o >= 0) return phys[o];
    if (o == kNoOperand) return zero_slot;
    return n_const ? (*entry_class)[~o] : ~o;
  };
  auto emit = [&](int opc, int oa, int ob, int def, int flag) {
    MicroWord w;
    w.hdr = opc | (K(oa) << 4) | (K(ob) << 6) | (flag << 8) | ((def >= 0 ? pool[def] : 0) << 10);
    w.a = V(oa); w.b = V(ob); w.c = def >= 0 ? phys[def] : 0;
    sp.code.push_back(w);
  };
  {
    int t = 0;
    for (const IrOp& op : ir) {
      if (op.kind == SOP_PIVOT) {
        emit(MOP_PIVHEAD, op.reads[op.pidx], kNoOperand, -1, 0);
        for (int c = 0; c < (int)op.reads.size(); ++c)
          if (c != op.pidx) emit(MOP_CAND, op.reads[c], kNoOperand, -1, c < op.pidx ? 1 : 0);
        for (int o : op.reads) release(o, t);
        alloc(op.def);
        emit(MOP_PIVEND, kNoOperand, kNoOperand, op.def, 0);
        ++t;
      } else if (op.kind == SOP_ELIM) {
        emit(MOP_ELIM, op.reads[0], kNoOperand, -1, 0);
        release(op.reads[0], t);
        ++t;
        for (const Update& u : op.upd) {
          const int ka = K(u.dst_old), va = V(u.dst_old), kb = K(u.src), vb = V(u.src);
          release(u.dst_old, t);   // in-place update when the old version dies here
          release(u.src, t);
          alloc(u.dst_new);
          MicroWord w;
          w.hdr = MOP_UPD | (ka << 4) | (kb << 6) | (pool[u.dst_new] << 10);
          w.a = va; w.b = vb; w.c = phys[u.dst_new];
          sp.code.push_back(w);
          ++t;
        }
      } else {
        emit(MOP_BHEAD, op.reads[0], op.reads[1], -1, 0);
        for (size_t q = 2; q + 1 < op.reads.size(); q += 2) emit(MOP_BTERM, op.reads[q], op.reads[q + 1], -1, 0);
        for (int o : op.reads) release(o, t);
        alloc(op.def);
        emit(MOP_BEND, kNoOperand, kNoOperand, op.def, 0);
        sp.code.back().a = op.var;  // variable index: the device also stores x_i straight into the result
        sp.code.back().hdr &= ~(3 << 4);  // A is a plain integer here, not an operand
        ++t;
      }
    }
  }
  const int high = high_slow;
  sp.n_fast = high_fast;
  { MicroWord e = {MOP_END, 0, 0, 0}; for (int q = 0; q < 4; ++q) sp.code.push_back(e); }  // END + prefetch padding
  sp.n_slots = high;
  sp.x_slot.resize(n);
  for (int i = 0; i < n; ++i) sp.x_slot[i] = phys[x_virtual[i]];
  sp.ok = true;
}

}  // namespace spicey
