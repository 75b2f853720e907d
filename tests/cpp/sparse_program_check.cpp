// CPU check of the host-side sparse program builder (spicey_b200/csrc/sparse_program.h): builds the
// micro-op program for random sparse complex systems and runs it with a reference interpreter that
// follows the device semantics (fast / global pools, constants, zero slot, pristine operands, pivot
// verification), comparing with dense Gaussian elimination with partial pivoting.
//   usage: sparse_program_check <seed> <n> <density%> <distinct_values (0 = all different)>
// Prints "OK <max rel err> slots=<..> fast=<..> const=<..> microops=<..>" or "FAIL ...".
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <map>
#include <random>
#include "../../spicey_b200/csrc/sparse_program.h"
#include "../../spicey_b200/csrc/warp_program.h"
using namespace spicey;
typedef std::complex<double> cd;

static bool dense_solve(int n, std::vector<cd> A, std::vector<cd> b, std::vector<cd>& x) {
  std::vector<int> rows(n);
  for (int i = 0; i < n; ++i) rows[i] = i;
  for (int k = 0; k < n; ++k) {
    int imax = k; double vmax = std::abs(A[rows[k] * n + k]);
    for (int i = k + 1; i < n; ++i) { double v = std::abs(A[rows[i] * n + k]); if (v > vmax) { vmax = v; imax = i; } }
    if (vmax < 1e-15) return false;
    std::swap(rows[k], rows[imax]);
    for (int i = k + 1; i < n; ++i) {
      cd f = A[rows[i] * n + k] / A[rows[k] * n + k];
      if (std::abs(f) < 1e-15) continue;
      for (int j = k; j < n; ++j) A[rows[i] * n + j] -= f * A[rows[k] * n + j];
      b[rows[i]] -= f * b[rows[k]];
    }
  }
  x.assign(n, cd(0, 0));
  for (int i = n - 1; i >= 0; --i) {
    cd s = b[rows[i]];
    for (int j = i + 1; j < n; ++j) s -= A[rows[i] * n + j] * x[j];
    x[i] = s / A[rows[i] * n + i];
  }
  return true;
}

int main(int argc, char** argv) {
  const unsigned seed = argc > 1 ? atoi(argv[1]) : 1;
  const int n = argc > 2 ? atoi(argv[2]) : 12;
  const double dens = (argc > 3 ? atoi(argv[3]) : 30) / 100.0;
  const int distinct = argc > 4 ? atoi(argv[4]) : 0;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> U(-1, 1), U01(0, 1);
  std::vector<cd> palette;
  for (int i = 0; i < distinct; ++i) palette.push_back(cd(U(rng), U(rng)));
  auto val = [&]() { return distinct ? palette[rng() % distinct] : cd(U(rng), U(rng)); };
  // random sparse matrix with a guaranteed transversal (a shuffled diagonal) and a dense-ish rhs
  std::map<std::pair<int, int>, cd> ent;
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  std::shuffle(perm.begin(), perm.end(), rng);
  for (int i = 0; i < n; ++i) ent[{i, perm[i]}] = val() + cd(2, 0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (U01(rng) < dens) ent[{i, j}] = val();
  for (int i = 0; i < n; ++i)
    if (U01(rng) < 0.5 || i == 0) ent[{i, n}] = val();
  std::vector<int> row_ptr(n + 1, 0), col;
  PilotInput in;
  in.n = n;
  int row = 0;
  for (auto& kv : ent) {
    while (row < kv.first.first) row_ptr[++row] = (int)col.size();
    col.push_back(kv.first.second);
    in.ent_val.push_back(kv.second);
  }
  while (row < n) row_ptr[++row] = (int)col.size();
  in.row_ptr = &row_ptr;
  in.ent_col = &col;
  const int n_ent = (int)col.size();
  std::vector<int> cls(n_ent);
  std::map<std::pair<double, double>, int> seen;
  int nc = 0;
  for (int i = 0; i < n_ent; ++i) {
    auto k = std::make_pair(in.ent_val[i].real(), in.ent_val[i].imag());
    if (!seen.count(k)) seen[k] = nc++;
    cls[i] = seen[k];
  }
  SparseProgram sp;
  build_sparse_program(in, sp, 12, &cls, nc);
  std::vector<cd> A((size_t)n * n, cd(0, 0)), b(n, cd(0, 0)), xref;
  for (int r = 0; r < n; ++r)
    for (int en = row_ptr[r]; en < row_ptr[r + 1]; ++en)
      (col[en] == n ? b[r] : A[(size_t)r * n + col[en]]) = in.ent_val[en];
  const bool ref_ok = dense_solve(n, A, b, xref);
  if (!sp.ok) { printf(ref_ok ? "FAIL builder refused a solvable pilot\n" : "OK singular pilot refused\n"); return ref_ok; }
  if (!ref_ok) { printf("FAIL builder accepted a singular pilot\n"); return 1; }
  // ---- reference interpreter (device semantics) ----
  std::vector<cd> W(sp.n_slots + 1, cd(1e300, 1e300)), F(sp.n_fast + 1, cd(1e300, 1e300)), xout(n, cd(0, 0));
  for (int c = 0; c < sp.n_const; ++c) F[c] = in.ent_val[sp.const_entry[c]];
  F[sp.n_const] = cd(0, 0);
  auto fetch = [&](int kind, int v) -> cd {
    if (kind == 1) return W[v];
    if (kind == 3) return F[v];
    if (kind == 2) return in.ent_val[v];
    printf("FAIL operand kind 0\n"); exit(1);
  };
  cd r, fm, ap, acc, rc;
  double mp = 0;
  bool ok = true;
  for (size_t pc = 0; pc < sp.code.size(); ++pc) {
    const MicroWord u = sp.code[pc];
    const int op = u.hdr & 15, ka = (u.hdr >> 4) & 3, kb = (u.hdr >> 6) & 3, kc = (u.hdr >> 10) & 3;
    auto put = [&](cd v) { (kc == 3 ? F[u.c] : W[u.c]) = v; if (kc != 1 && kc != 3) { printf("FAIL dst kind\n"); exit(1); } };
    if (op == MOP_END) break;
    switch (op) {
      case MOP_UPD: put(fetch(ka, u.a) - fm * fetch(kb, u.b)); break;
      case MOP_BTERM: acc -= fetch(ka, u.a) * fetch(kb, u.b); break;
      case MOP_CAND: { double m = std::norm(fetch(ka, u.a)); ok = ok && (((u.hdr >> 8) & 1) ? m < mp : !(m > mp)); break; }
      case MOP_ELIM: fm = fetch(ka, u.a) * r; if (std::norm(fm) < 1e-30) fm = 0; break;
      case MOP_PIVHEAD: ap = fetch(ka, u.a); mp = std::norm(ap); ok = true; break;
      case MOP_PIVEND: if (!ok) { printf("FAIL pivot verification rejected the pilot itself\n"); return 1; } r = cd(1, 0) / ap; put(r); break;
      case MOP_BHEAD: acc = fetch(ka, u.a); rc = fetch(kb, u.b); break;
      case MOP_BEND: put(acc * rc); xout[u.a] = acc * rc; break;
      default: printf("FAIL opcode %d\n", op); return 1;
    }
  }
  double err = 0, scale = 0;
  for (int i = 0; i < n; ++i) { err = std::max(err, std::abs(xout[i] - xref[i])); scale = std::max(scale, std::abs(xref[i])); }
  for (int i = 0; i < n; ++i)
    if (std::abs(W[sp.x_slot[i]] - xout[i]) != 0) { printf("FAIL x_slot mismatch\n"); return 1; }
  const double rel = err / std::max(scale, 1e-300);
  // ---- warp program (warp_program.h): emulate the kernel's schedule — levels in chunks of 32 operations,
  //      every chunk loads all its operands before it stores any result ----
  WarpProgram wp;
  // per-entry constants as the library fills them (here: frequency independent)
  sp.ent_alpha.assign(n_ent, 0.0); sp.ent_jre.assign(n_ent, 0.0); sp.ent_jim.assign(n_ent, 0.0);
  sp.ent_beta.assign(n_ent, 0.0); sp.ent_gamma.assign(n_ent, 0.0);
  for (int en = 0; en < n_ent; ++en) { sp.ent_alpha[en] = in.ent_val[en].real(); sp.ent_jim[en] = in.ent_val[en].imag(); }
  build_warp_program(sp, 1 << 20, wp);
  if (!wp.ok) { printf("OK %.3e (warp program not applicable: pool %d)\n", rel, wp.n_pool); return rel < 1e-9 ? 0 : 1; }
  double relw = 0;
  {
    // executes the PACKED records exactly as ac_warp.cuh parses them
    std::vector<cd> pool(wp.n_pool, cd(1e300, 1e300)), G(wp.n_gslots, cd(1e300, 1e300)), Fm(std::max(1, wp.max_elim));
    pool[0] = cd(0, 0);
    auto fb = [&](int e) -> cd { return e == kWarpZero ? cd(0, 0) : (e >= 0 ? G[e] : in.ent_val[~e]); };
    for (int s = 0; s < n; ++s) {
      const int* rec = wp.stream.data() + 4 * (size_t)wp.fwd_tab[2 * s];
      if (wp.fwd_tab[2 * s + 1] > wp.max_rec16) { printf("FAIL warp: record larger than max_rec16\n"); return 1; }
      const int n_cand = rec[0], pidx = rec[1], rcp_g = rec[2], n_elim = rec[3], n_cols = rec[4], n_stamp = rec[5];
      const int* stamp = rec + 8;
      const int* cand = stamp + 12 * n_stamp;
      const int* elim = cand + n_cand;
      const int* src = elim + n_elim;
      const int* ops = rec + ((8 + 12 * n_stamp + n_cand + n_elim + n_cols + 3) & ~3);
      const int* opg = ops + ((n_elim * n_cols + 3) & ~3);
      for (int q = 0; q < n_stamp; ++q) {
        double c[4];
        memcpy(c, stamp + 12 * q + 4, sizeof c);
        pool[stamp[12 * q]] = cd(c[0], c[1]);
      }
      const cd apv = pool[cand[pidx]];
      const double mpv = std::norm(apv);
      for (int c = 0; c < n_cand; ++c) {
        const double m = std::norm(pool[cand[c]]);
        if ((c < pidx && !(m < mpv)) || (c > pidx && (m > mpv))) { printf("FAIL warp: pivot verification rejected the pilot\n"); return 1; }
      }
      const cd rv = cd(1, 0) / apv;
      if (rcp_g >= 0) G[rcp_g] = rv;
      for (int e = 0; e < n_elim; ++e) { cd f = pool[elim[e]] * rv; if (std::norm(f) < 1e-30) f = 0; Fm[e] = f; }
      for (int c0 = 0; c0 < n_cols; c0 += 32) {      // a pass: pivot-row entries first, then row by row
        const int c1 = std::min(n_cols, c0 + 32);
        std::vector<cd> sv(c1 - c0), val(c1 - c0);
        for (int c = c0; c < c1; ++c) sv[c - c0] = pool[src[c]];
        for (int e = 0; e < n_elim; ++e) {
          for (int c = c0; c < c1; ++c) val[c - c0] = pool[((unsigned)ops[e * n_cols + c] & 0xfff0) / 16] - Fm[e] * sv[c - c0];
          for (int c = c0; c < c1; ++c) {
            const unsigned w = (unsigned)ops[e * n_cols + c];
            const unsigned dst = w >> 16;
            if (dst != 0xfff0) pool[dst / 16] = val[c - c0];
            if (w & 1) G[opg[e * n_cols + c]] = val[c - c0];
            else if (opg[e * n_cols + c] >= 0) { printf("FAIL warp: global copy not flagged\n"); return 1; }
          }
        }
      }
    }
    std::vector<cd> accv(n);
    for (int i = 0; i < n; ++i) accv[i] = fb(wp.rhs_init[i]);
    int expect_first = wp.g_first0, expect_count = wp.g_count0;
    for (int gi = 0; gi < wp.n_groups; ++gi) {
      const int* rec = wp.stream.data() + 4 * (size_t)wp.back_tab[2 * gi];
      const int nc = rec[0];
      int mn = INT_MAX, mx = -1;
      for (int ci = 0; ci < nc; ++ci) {
        const int* c = rec + 4 + 4 * ci;   // {rcp_g, ent_begin, count, j}
        const cd xj = accv[c[3]] * G[c[0]];
        mn = std::min(mn, c[0]); mx = std::max(mx, c[0]);
        for (int q = 0; q < c[2]; ++q) {
          const int row = rec[c[1] + 2 * q], ue = rec[c[1] + 2 * q + 1];
          if (ue >= 0) { mn = std::min(mn, ue); mx = std::max(mx, ue); }
          accv[row] -= fb(ue) * xj;
        }
        accv[c[3]] = xj;
      }
      if (mx >= 0 && (mn < expect_first || mx >= expect_first + expect_count)) { printf("FAIL warp: prefetch range does not cover group %d\n", gi); return 1; }
      expect_first = rec[1]; expect_count = rec[2];
    }
    double errw = 0;
    for (int i = 0; i < n; ++i) errw = std::max(errw, std::abs(accv[i] - xref[i]));
    relw = errw / std::max(scale, 1e-300);
  }
  const bool good = rel < 1e-9 && relw < 1e-9;
  printf("%s %.3e warp %.3e slots=%d fast=%d const=%d microops=%zu virtual=%d pool=%d gslots=%d\n", good ? "OK" : "FAIL", rel, relw, sp.n_slots,
         sp.n_fast, sp.n_const, sp.code.size(), sp.n_virtual, wp.n_pool, wp.n_gslots);
  return good ? 0 : 1;
}
