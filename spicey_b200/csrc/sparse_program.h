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
//   PIVOT  k: candidate slots (rows with a structural non-zero in column k, in the
//             reference's logical row order) and which of them the pilot chose
//   ELIM   i: a_ij -= f * a_kj for the structural non-zeros j of the pivot row
//   BSUB   i: x_i = (b_i - sum U_ij x_j) / u_ii
//
// The device executes the program with one thread per system and VERIFIES the pivot of
// every step against the reference's rule on that system's own numbers; a system whose
// pivot choice differs from the pilot's is handed to the dense pivoting kernel.  Results
// are therefore those of partial pivoting for every system, never of a frozen ordering.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <vector>

namespace spicey {

enum SparseOp { SOP_PIVOT = 1, SOP_ELIM = 2, SOP_BSUB = 3, SOP_END = 4 };

struct SparseProgram {
  bool ok = false;
  int n = 0;
  int n_stamp = 0;   // slots [0, n_stamp) are the gather-plan entries, in plan order
  int n_slots = 0;   // matrix slots incl. fill (x lives in [n_slots, n_slots+n))
  long long n_fma = 0, n_div = 0;  // executed complex FMAs / reciprocals per system
  std::vector<int> code;
  // per stamped entry: value = alpha + j*(omega*beta - gamma/omega) (+ phasor for rhs entries)
  std::vector<double> ent_alpha, ent_beta, ent_gamma, ent_jre, ent_jim;
  // per AC element (R,C,L,V order): current = Y_e * (v1 - v2), Y_e = ya + j*(omega*yb - yg/omega)
  std::vector<double> el_a, el_b, el_g;
  std::vector<double> ind_L;  // inductances, for the per-point guards of simulateAC.ts:47-51
};

struct PilotInput {
  int n = 0;                              // nvar
  const std::vector<int>* row_ptr = nullptr;   // gather plan (AC)
  const std::vector<int>* ent_col = nullptr;
  // numeric pilot matrix entries, one per plan entry, in plan order
  std::vector<std::complex<double>> ent_val;
};

// Builds the program from the pilot point.  Returns ok=false when the pilot itself is
// singular / hits the divide guard (the dense kernel then reports the exact status).
inline void build_sparse_program(const PilotInput& in, SparseProgram& sp) {
  typedef std::complex<double> cd;
  const int n = in.n, ld = n + 1;
  const double EPS = 1e-15;
  sp.ok = false;
  sp.n = n;
  sp.code.clear();
  std::vector<cd> M((size_t)n * ld, cd(0, 0));
  std::vector<int> slot((size_t)n * ld, -1);
  int ns = 0;
  for (int r = 0; r < n; ++r)
    for (int en = (*in.row_ptr)[r]; en < (*in.row_ptr)[r + 1]; ++en) {
      int c = (*in.ent_col)[en];
      M[(size_t)r * ld + c] = in.ent_val[en];
      slot[(size_t)r * ld + c] = en;
      ++ns;
    }
  sp.n_stamp = ns;
  std::vector<int> rows(n);  // logical position -> physical row (the reference's swapped row array)
  for (int i = 0; i < n; ++i) rows[i] = i;
  std::vector<int> piv_slot(n, -1);
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
    if (vmax < EPS) return;                                  // pilot singular
    if (slot[(size_t)rows[imax] * ld + k] < 0) return;       // cannot happen (non-zero value has a slot)
    // candidates in logical order BEFORE the swap (that is the order the reference scans)
    std::vector<int> cand_row;
    int pidx = -1;
    for (int i = k; i < n; ++i)
      if (slot[(size_t)rows[i] * ld + k] >= 0) {
        if (i == imax) pidx = (int)cand_row.size();
        cand_row.push_back(rows[i]);
      }
    sp.code.push_back(SOP_PIVOT);
    sp.code.push_back((int)cand_row.size());
    sp.code.push_back(pidx);
    for (int r : cand_row) sp.code.push_back(slot[(size_t)r * ld + k]);
    std::swap(rows[k], rows[imax]);
    const int p = rows[k];
    const cd pivot = M[(size_t)p * ld + k];
    if (std::norm(pivot) < EPS) return;                      // Complex.div guard on the pilot
    piv_slot[k] = slot[(size_t)p * ld + k];
    sp.n_div++;
    // structural non-zeros of the pivot row right of k (incl. rhs column n)
    std::vector<int> pcols;
    for (int j = k + 1; j <= n; ++j)
      if (slot[(size_t)p * ld + j] >= 0) pcols.push_back(j);
    for (int r : cand_row) {
      if (r == p) continue;
      sp.code.push_back(SOP_ELIM);
      sp.code.push_back(slot[(size_t)r * ld + k]);
      sp.code.push_back((int)pcols.size());
      const cd f = M[(size_t)r * ld + k] / pivot;
      const bool act = !(std::abs(f) < EPS);
      for (int j : pcols) {
        int& s = slot[(size_t)r * ld + j];
        int fresh = 0;
        if (s < 0) { s = ns++; fresh = 1; }
        sp.code.push_back((s << 1) | fresh);
        sp.code.push_back(slot[(size_t)p * ld + j]);
        if (act) M[(size_t)r * ld + j] -= f * M[(size_t)p * ld + j];
        sp.n_fma++;
      }
    }
  }
  sp.n_slots = ns;
  for (int i = n - 1; i >= 0; --i) {
    const int r = rows[i];
    sp.code.push_back(SOP_BSUB);
    sp.code.push_back(i);
    sp.code.push_back(slot[(size_t)r * ld + n]);  // rhs slot or -1
    sp.code.push_back(piv_slot[i]);
    int cnt_at = (int)sp.code.size();
    sp.code.push_back(0);
    int cnt = 0;
    for (int j = i + 1; j < n; ++j)
      if (slot[(size_t)r * ld + j] >= 0) {
        sp.code.push_back(slot[(size_t)r * ld + j]);
        sp.code.push_back(j);
        ++cnt;
        sp.n_fma++;
      }
    sp.code[cnt_at] = cnt;
  }
  sp.code.push_back(SOP_END);
  sp.ok = true;
}

}  // namespace spicey
