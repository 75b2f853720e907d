// Host side of the banded + bordered AC tier (band_kernel.cuh): ordering, pilot factorisation, stamp tables.
//
// The reference eliminates the unknowns in netlist order (node ids in order of first appearance, then the V-source
// branches: NodeIndex.ts:28-31, parseNetlist.ts:455-460) with partial pivoting (solveComplex.ts:15-53).  A sparse
// direct solver is free to renumber the unknowns symmetrically first: the result of partial pivoting on the
// renumbered system is the same solution to rounding (SURVEY.md 8 c, hazard H7: <= 4.6e-12 component-wise on the
// ladder and the mesh against an unrelated elimination order), and a small bandwidth is what lets one system live
// in the registers of a few lanes.  This file
//   1. builds the node graph of the AC matrix and a reverse Cuthill-McKee ordering of it (the netlist order is kept
//      when it is already as narrow: a ladder numbered end to end has half-bandwidth 1);
//   2. runs the reference's algorithm once on a pilot point of the renumbered system — pivot metric hypot, first
//      maximum wins, row swaps, the |f| < EPS row skip — and records the pivot row of every step and, for the tie rule,
//      where every row stood in the scan order of every step;
//   3. checks that the permuted matrix is a band of half-width W <= 32 around the pivot rows plus the NB = nV
//      border rows / columns, and lays out the per-step tables of stamp constants the kernel consumes: every
//      stamped entry is delivered exactly once, at the step where the register window first has a slot for it
//      (see "deliveries" below).
// No device code; included by spicey_native.cu and by tests/cpp/band_plan_check.cpp (which replays the kernel's
// schedule on the CPU against dense pivoted elimination).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

namespace spicey {

constexpr int kBandMaxW = 32;
constexpr int kBandMaxNB = 4;

struct BandRecipe { double alpha_jre, jim, beta, gamma; };   // value = alpha_jre + j (w beta - gamma / w + jim)

struct BandInput {   // the AC gather plan with per-entry constants (spicey_native.cu: build_sparse_host)
  int n = 0, nn = 0, nV = 0;
  const std::vector<int>* row_ptr = nullptr;   // [n + 1]
  const std::vector<int>* ent_col = nullptr;   // column, n = right-hand side
  const std::vector<double>*ent_alpha = nullptr, *ent_beta = nullptr, *ent_gamma = nullptr, *ent_jre = nullptr, *ent_jim = nullptr;
  double pilot_w = 1.0;                        // angular frequency of the pilot point
};

struct BandPlan {
  bool ok = false;
  int n = 0, nb = 0, NB = 0, W = 0, L = 0, RPL = 0;   // W = L * RPL >= measured half-bandwidth; n, nb include the padding
  int n_orig = 0;                                      // unknowns of the circuit (n = n_orig + padding rows)
  int bandwidth = 0;                                   // measured half-bandwidth around the pivot rows
  bool renumbered = false;                             // false: netlist order kept
  unsigned abmask = 0;                                 // border columns with structural non-zeros in band rows
  std::vector<int> newvar, oldvar;                     // original variable <-> elimination-order index
  std::vector<int> prow;                               // [n] pivot row (an ORIGINAL row index) of step k
  std::vector<BandRecipe> tab;
  std::vector<unsigned> flags;                         // [n][2]
  int o_init = 0, o_initb = 0, o_brd0 = 0, o_bb0 = 0, o_step = 0;   // table offsets (entries)
  int step_stride = 0;                                 // entries per step record (band_kernel.cuh: BST_*)
  bool rc_only = false;                                // no entry has a gamma (inductor) or Im J term: (alpha, beta) tables
  long long g_stride = 0;                              // workspace per system, complex values
  long long n_cfma = 0;                                // complex FMAs the kernel executes per system (dense band)
};

namespace band_detail {

// Reverse Cuthill-McKee over the node graph; every component from a pseudo-peripheral start node.
inline std::vector<int> rcm_order(int nn, const std::vector<std::vector<int>>& adj) {
  std::vector<int> order, level(nn), seen(nn, 0);
  order.reserve(nn);
  auto bfs_far = [&](int start, std::vector<int>& comp) {   // returns a node of the last level with minimum degree
    comp.clear();
    std::vector<int> q = {start};
    std::fill(level.begin(), level.end(), -1);
    level[start] = 0;
    for (size_t h = 0; h < q.size(); ++h) {
      const int u = q[h];
      for (int v : adj[u]) if (level[v] < 0) { level[v] = level[u] + 1; q.push_back(v); }
    }
    comp = q;
    const int last = level[q.back()];
    int best = q.back();
    for (int u : q) if (level[u] == last && adj[u].size() < adj[best].size()) best = u;
    return std::make_pair(best, last);
  };
  for (int root = 0; root < nn; ++root) {
    if (seen[root]) continue;
    std::vector<int> comp;
    int start = root, ecc = -1;
    {  // minimum-degree node of the component first, then walk to a pseudo-peripheral one
      bfs_far(root, comp);
      for (int u : comp) if (adj[u].size() < adj[start].size()) start = u;
    }
    for (int it = 0; it < 8; ++it) {
      auto fe = bfs_far(start, comp);
      if (fe.second <= ecc) break;
      ecc = fe.second;
      start = fe.first;
    }
    // Cuthill-McKee from `start`: neighbours in order of increasing degree
    std::vector<int> q = {start};
    seen[start] = 1;
    for (size_t h = 0; h < q.size(); ++h) {
      const int u = q[h];
      std::vector<int> nb;
      for (int v : adj[u]) if (!seen[v]) { seen[v] = 1; nb.push_back(v); }
      std::stable_sort(nb.begin(), nb.end(), [&](int x, int y) { return adj[x].size() < adj[y].size(); });
      for (int v : nb) q.push_back(v);
    }
    for (int u : q) order.push_back(u);
  }
  std::reverse(order.begin(), order.end());
  return order;   // order[new index] = old node
}

struct Pilot {
  bool ok = false;
  std::vector<int> prow;          // [n] pivot row (initial position index) of step k
  std::vector<int> where;         // [n][n] position of row r in the scan order at the start of step k
  std::vector<int> pivpos;        // [n] position the pivot was found at
};

// The reference's elimination (solveComplex.ts:15-53) on the renumbered pilot matrix M[n][n+1] (row-major).
inline void run_pilot(int n, std::vector<std::complex<double>>& M, Pilot& P) {
  typedef std::complex<double> cd;
  const double EPS = 1e-15;
  const int ld = n + 1;
  std::vector<int> rows(n), pos(n);
  for (int i = 0; i < n; ++i) rows[i] = pos[i] = i;
  P.ok = false;
  P.prow.assign(n, 0); P.pivpos.assign(n, 0);
  P.where.assign((size_t)n * n, 0);
  for (int k = 0; k < n; ++k) {
    for (int i = 0; i < n; ++i) P.where[(size_t)k * n + rows[i]] = i;
    int imax = k;
    double vmax = std::hypot(M[(size_t)rows[k] * ld + k].real(), M[(size_t)rows[k] * ld + k].imag());
    for (int i = k + 1; i < n; ++i) {
      const cd& z = M[(size_t)rows[i] * ld + k];
      if (z.real() == 0.0 && z.imag() == 0.0) continue;
      const double v = std::hypot(z.real(), z.imag());
      if (v > vmax) { vmax = v; imax = i; }
    }
    if (vmax < EPS) return;
    P.pivpos[k] = imax;
    std::swap(rows[k], rows[imax]);
    const int p = rows[k];
    P.prow[k] = p;
    const cd pivot = M[(size_t)p * ld + k];
    if (std::norm(pivot) < EPS) return;   // Complex.div guard
    for (int i = k + 1; i < n; ++i) {
      const int r = rows[i];
      const cd z = M[(size_t)r * ld + k];
      if (z.real() == 0.0 && z.imag() == 0.0) continue;
      const cd f = z / pivot;
      M[(size_t)r * ld + k] = cd(0, 0);
      if (std::abs(f) < EPS) continue;
      for (int j = k + 1; j <= n; ++j) {
        const cd u = M[(size_t)p * ld + j];
        if (u.real() != 0.0 || u.imag() != 0.0) M[(size_t)r * ld + j] -= f * u;
      }
    }
  }
  P.ok = true;
}

// The reference's whole solve (solveComplex.ts:15-72: elimination as run_pilot, back-substitution :55-72) of the dense
// augmented matrix M[n][n+1]; false when a pivot trips one of its guards.
inline bool solve_like_reference(int n, std::vector<std::complex<double>>& M, std::vector<std::complex<double>>& x) {
  typedef std::complex<double> cd;
  const double EPS = 1e-15;
  const int ld = n + 1;
  for (int k = 0; k < n; ++k) {
    int imax = k;
    double vmax = std::hypot(M[(size_t)k * ld + k].real(), M[(size_t)k * ld + k].imag());
    for (int i = k + 1; i < n; ++i) {
      const cd& z = M[(size_t)i * ld + k];
      const double v = std::hypot(z.real(), z.imag());
      if (v > vmax) { vmax = v; imax = i; }
    }
    if (vmax < EPS) return false;
    if (imax != k) for (int j = 0; j <= n; ++j) std::swap(M[(size_t)k * ld + j], M[(size_t)imax * ld + j]);
    const cd pivot = M[(size_t)k * ld + k];
    if (std::norm(pivot) < EPS) return false;
    for (int i = k + 1; i < n; ++i) {
      const cd z = M[(size_t)i * ld + k];
      if (z.real() == 0.0 && z.imag() == 0.0) continue;
      const cd f = z / pivot;
      if (std::abs(f) < EPS) continue;
      for (int j = k; j <= n; ++j) {
        const cd u = M[(size_t)k * ld + j];
        if (u.real() != 0.0 || u.imag() != 0.0) M[(size_t)i * ld + j] -= f * u;
      }
    }
  }
  x.assign(n, cd(0, 0));
  for (int i = n - 1; i >= 0; --i) {
    cd acc = M[(size_t)i * ld + n];
    for (int j = i + 1; j < n; ++j) acc -= M[(size_t)i * ld + j] * x[j];
    if (std::norm(M[(size_t)i * ld + i]) < EPS) return false;
    x[i] = acc / M[(size_t)i * ld + i];
  }
  return true;
}

}  // namespace band_detail

// (L, RPL) for a measured half-bandwidth: W = L * RPL is a power of two >= max(2, bandwidth).
inline bool band_shape(int bandwidth, int& L, int& RPL) {
  if (bandwidth <= 2) { L = 2; RPL = 1; }
  else if (bandwidth <= 4) { L = 4; RPL = 1; }
  else if (bandwidth <= 8) { L = 8; RPL = 1; }
  else if (bandwidth <= 16) { L = 8; RPL = 2; }
  else if (bandwidth <= 32) { L = 32; RPL = 1; }
  else return false;
  return true;
}

// Builds the plan for one ordering (order[new] = old node).  Returns bp.ok = false when the matrix is not banded +
// bordered under this ordering, or when the pilot is singular.
inline void build_band_plan_for_order(const BandInput& in, const std::vector<int>& order, BandPlan& bp, int force_L = 0, int force_RPL = 0) {
  using namespace band_detail;
  typedef std::complex<double> cd;
  const int n = in.n, nn = in.nn, NB = in.nV, nb = nn, ld = n + 1;
  bp = BandPlan();
  bp.n = n; bp.nb = nb; bp.NB = NB;
  if (NB > kBandMaxNB || nb < 2) return;
  bp.oldvar.resize(n); bp.newvar.resize(n);
  for (int i = 0; i < nb; ++i) bp.oldvar[i] = order[i];
  for (int b = 0; b < NB; ++b) bp.oldvar[nb + b] = nn + b;
  for (int i = 0; i < n; ++i) bp.newvar[bp.oldvar[i]] = i;
  // entry lookup in the renumbered system: E[row][col] = gather-plan entry or -1 (col n = rhs)
  std::vector<int> E((size_t)n * ld, -1);
  std::vector<cd> M((size_t)n * ld, cd(0, 0));
  const double w = in.pilot_w;
  for (int r = 0; r < n; ++r)
    for (int en = (*in.row_ptr)[r]; en < (*in.row_ptr)[r + 1]; ++en) {
      const int c = (*in.ent_col)[en];
      const int pr = bp.newvar[r], pc = c < n ? bp.newvar[c] : n;
      E[(size_t)pr * ld + pc] = en;
      M[(size_t)pr * ld + pc] = cd((*in.ent_alpha)[en] + (*in.ent_jre)[en],
                                   w * (*in.ent_beta)[en] - (*in.ent_gamma)[en] / w + (*in.ent_jim)[en]);
    }
  Pilot P;
  run_pilot(n, M, P);
  if (!P.ok) return;
  bp.prow = P.prow;   // positions in the renumbered system; converted to original rows at the end
  // band check around the pivot rows
  int bw = 1;
  unsigned abmask = 0;
  for (int k = 0; k < nb; ++k) {
    const int r = P.prow[k];
    for (int c = 0; c < n; ++c) {
      if (E[(size_t)r * ld + c] < 0) continue;
      if (c >= nb) abmask |= 1u << (c - nb);
      else bw = std::max(bw, std::abs(c - k));
    }
  }
  bp.bandwidth = bw;
  bp.abmask = abmask;
  int L = 0, RPL = 0;
  if (!band_shape(bw, L, RPL)) return;
  if (force_L > 0 && force_RPL > 0 && force_L * force_RPL >= std::max(2, bw)) { L = force_L; RPL = force_RPL; }
  const int W = L * RPL;
  bp.L = L; bp.RPL = RPL; bp.W = W;

  // The band part is padded to a multiple of W with identity rows (diagonal 1, nothing else, right-hand side 0: the
  // extra unknowns are exact zeros): the kernel's loops, unrolled W times, then run whole blocks without a tail test.
  // From here on indices are PADDED: band 0 .. nbp-1, border nbp .. nbp+NB-1, right-hand side np = nbp + NB.
  const int nbp = (nb + W - 1) / W * W, np = nbp + NB;
  bp.n_orig = n;
  bp.n = np; bp.nb = nbp;
  for (int b = 0; b < NB; ++b) bp.newvar[nn + b] = nbp + b;
  bp.oldvar.assign(np, -1);
  for (int v = 0; v < n; ++v) bp.oldvar[bp.newvar[v]] = v;
  auto ucol = [&](int c) -> int {   // padded column -> column of the renumbered (unpadded) system, -1 = padding
    if (c < nb) return c;
    if (c < nbp) return -1;
    return c == np ? n : nb + (c - nbp);
  };

  // ---- deliveries ----
  // rec(i, c): stamped entry of band row i (pivot row of step i) at padded column c (np = rhs); zero outside.
  const BandRecipe zero = {0.0, 0.0, 0.0, 0.0}, one = {1.0, 0.0, 0.0, 0.0};
  auto rec_row = [&](int prow_pos, int c) -> BandRecipe {
    const int uc = ucol(c);
    if (uc < 0) return zero;
    const int en = E[(size_t)prow_pos * ld + uc];
    if (en < 0) return zero;
    BandRecipe q = {(*in.ent_alpha)[en] + (*in.ent_jre)[en], (*in.ent_jim)[en], (*in.ent_beta)[en], (*in.ent_gamma)[en]};
    return q;
  };
  auto rec = [&](int i, int c) -> BandRecipe {
    if (i < 0 || i >= nbp || c < 0 || c > np) return zero;
    if (i >= nb) return c == i ? one : zero;   // padding row
    return rec_row(P.prow[i], c);
  };
  auto recb = [&](int b, int c) -> BandRecipe { return (c < 0 || c > np) ? zero : rec_row(P.prow[nb + b], c); };
  auto bcol = [&](int j) { return j < NB ? nbp + j : np; };   // border column j, or the right-hand side (j = NB)
  auto row_at = [&](int t, int lo) { return lo + (((t - lo) % W) + W) % W; };   // the row = t (mod W) in [lo, lo + W)
  std::vector<BandRecipe>& T = bp.tab;
  T.clear();
  bp.o_init = (int)T.size();     // rows 0 .. W-1: column 0, and the entries above the diagonal of columns 1 .. W-1
  for (int i = 0; i < W; ++i)
    for (int c = 0; c < W; ++c) T.push_back((c == 0 || i < c) ? rec(i, c) : zero);
  bp.o_initb = (int)T.size();
  for (int i = 0; i < W; ++i)
    for (int j = 0; j <= NB; ++j) T.push_back(rec(i, bcol(j)));
  bp.o_brd0 = (int)T.size();
  for (int b = 0; b < NB; ++b)
    for (int c = 0; c < W; ++c) T.push_back(recb(b, c));
  bp.o_bb0 = (int)T.size();
  for (int b = 0; b < NB; ++b)
    for (int j = 0; j <= NB; ++j) T.push_back(recb(b, bcol(j)));
  // One record per step (what step k delivers, contiguous: the kernel stages a record two steps ahead):
  //   [0, W)        position t: entry of column k + W in the row = t (mod W) of k .. k+W-1 (above the diagonal)
  //   [W, 2W)       position t: entry of column k + 1 in the row = t (mod W) of k+1 .. k+W (diagonal and below)
  //   2W            entry (k + W, k) of the entering row
  //   2W+1 ..       border columns / rhs of the entering row [NB + 1], then the border rows' entries of column k + W [NB]
  //   2W+2NB+2      the tie-rule masks of the step
  // nbp + 3 records: the publisher of step k reads record k + 1, the staging reaches record k + 2.
  while (T.size() % 8) T.push_back(zero);
  bp.o_step = (int)T.size();
  bp.step_stride = (2 * W + 2 * NB + 3 + 7) / 8 * 8;
  const size_t first_record = T.size();
  for (int k = 0; k < nbp + 3; ++k) {
    const size_t base = T.size();
    for (int t = 0; t < W; ++t) T.push_back(k + W < nbp ? rec(row_at(t, k), k + W) : zero);
    for (int t = 0; t < W; ++t) T.push_back(k + 1 < nbp ? rec(row_at(t, k + 1), k + 1) : zero);
    T.push_back(k < nbp ? rec(k + W, k) : zero);
    for (int j = 0; j <= NB; ++j) T.push_back(rec(k + W, bcol(j)));
    for (int b = 0; b < NB; ++b) T.push_back(k + W < nbp ? recb(b, k + W) : zero);
    T.push_back(zero);   // the tie-rule masks of the step (filled in below)
    while (T.size() < base + (size_t)bp.step_stride) T.push_back(zero);
  }
  bp.rc_only = true;
  for (const BandRecipe& q : T) if (q.jim != 0.0 || q.gamma != 0.0) { bp.rc_only = false; break; }

  // ---- tie rule: is the candidate scanned before the pilot's pivot? (solveComplex.ts:18-28, strict '>') ----
  // (padding rows are never candidates: their columns hold nothing but their own diagonal)
  bp.flags.assign((size_t)2 * np, 0u);
  for (int ku = 0; ku < n; ++ku) {   // ku: step of the unpadded pilot
    const int* wh = &P.where[(size_t)ku * n];
    const int ppos = P.pivpos[ku];
    unsigned fx = 0, fy = 0;
    int k = ku;                      // padded step
    if (ku < nb) {
      for (int t = 0; t < W; ++t) {
        const int j = row_at(t, ku + 1);
        if (j < nb && wh[P.prow[j]] < ppos) fx |= 1u << t;
      }
      for (int b = 0; b < NB; ++b) if (wh[P.prow[nb + b]] < ppos) fy |= 1u << b;
    } else {
      k = nbp + (ku - nb);
      for (int b = ku - nb + 1; b < NB; ++b) if (wh[P.prow[nb + b]] < ppos) fy |= 1u << b;
    }
    bp.flags[2 * k] = fx; bp.flags[2 * k + 1] = fy;
    if (ku < nb) {   // the kernel reads the masks of a band step from the step's record
      const unsigned long long bits = (unsigned long long)fx | ((unsigned long long)fy << 32);
      double asd;
      static_assert(sizeof asd == sizeof bits, "double is 64 bits");
      memcpy(&asd, &bits, sizeof asd);
      T[first_record + (size_t)k * bp.step_stride + 2 * W + 2 * NB + 2].alpha_jre = asd;
    }
  }
  bp.g_stride = ((long long)(nbp + W) * W + (long long)nbp * (NB + 2) + 7) / 8 * 8;
  int nbc = 1;
  for (int j = 0; j < NB; ++j) nbc += (abmask >> j) & 1;
  bp.n_cfma = (long long)nbp * ((long long)W * (W + nbc) + (long long)NB * (W + NB + 1)) + (long long)nbp * (W + nbc);
  // pivot rows as ORIGINAL row indices (the dense fallback and the tests speak that language)
  for (int k = 0; k < n; ++k) { const int pr = bp.prow[k]; bp.prow[k] = pr < nb ? order[pr] : nn + (pr - nb); }
  bp.ok = true;
}

// Chooses between the netlist order and reverse Cuthill-McKee (the narrower band; the netlist order on a tie).
inline void build_band_plan(const BandInput& in, BandPlan& bp, int force_L = 0, int force_RPL = 0) {
  const int n = in.n, nn = in.nn;
  bp = BandPlan();
  if (nn < 2 || in.nV > kBandMaxNB) return;
  std::vector<int> ident(nn);
  for (int i = 0; i < nn; ++i) ident[i] = i;
  // node graph: off-diagonal node-node entries; the two nodes of a V source count as neighbours (the source's row
  // replaces one of their rows in the band, with entries in both columns)
  std::vector<std::vector<int>> adj(nn);
  for (int r = 0; r < n; ++r) {
    std::vector<int> nodes;
    for (int en = (*in.row_ptr)[r]; en < (*in.row_ptr)[r + 1]; ++en) {
      const int c = (*in.ent_col)[en];
      if (c < nn) nodes.push_back(c);
    }
    if (r < nn) { for (int c : nodes) if (c != r) adj[r].push_back(c); }
    else for (int u : nodes) for (int v : nodes) if (u != v) adj[u].push_back(v);
  }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  auto graph_bw = [&](const std::vector<int>& order) {
    std::vector<int> inv(nn);
    for (int i = 0; i < nn; ++i) inv[order[i]] = i;
    int bw = 0;
    for (int u = 0; u < nn; ++u) for (int v : adj[u]) bw = std::max(bw, std::abs(inv[u] - inv[v]));
    return bw;
  };
  const std::vector<int> rcm = band_detail::rcm_order(nn, adj);
  const int bw_id = graph_bw(ident), bw_rcm = graph_bw(rcm);
  int Li = 0, Ri = 0, Lr = 0, Rr = 0;
  const bool id_fits = band_shape(std::max(1, bw_id), Li, Ri), rcm_fits = band_shape(std::max(1, bw_rcm), Lr, Rr);
  // the netlist order unless the renumbering buys a smaller window
  if (id_fits && (!rcm_fits || Li * Ri <= Lr * Rr)) {
    build_band_plan_for_order(in, ident, bp, force_L, force_RPL);
    if (bp.ok) return;
  }
  if (rcm_fits) {
    build_band_plan_for_order(in, rcm, bp, force_L, force_RPL);
    if (bp.ok) bp.renumbered = true;
  }
}

// What a renumbering costs in per-entry agreement with the reference.  A symmetric renumbering changes the elimination
// order, not the exact solution; in floating point both orders are backward stable, i.e. equally good relative to the
// LARGEST unknown, but the small unknowns of a strongly attenuating network (a node voltage 1e-6 of the source's) keep
// their relative accuracy only along some orders.  The parity bar is per entry against the reference's order (1e-9), so
// a plan that renumbers is checked before it is used: the reference's algorithm on the host in both orders at angular
// frequency w, returns max_i |x_plan[i] - x_netlist[i]| / max(|x_netlist[i]|, 1e-12 max|x_netlist|), or a negative value
// when either solve trips a guard.  (cfg 4's mesh: a few 1e-11; a random RC tree of 100 nodes with chords: 1e-6.)
inline double band_order_deviation(const BandInput& in, const BandPlan& bp, double w) {
  typedef std::complex<double> cd;
  const int n = in.n, ld = n + 1;
  if (!bp.ok || (int)bp.newvar.size() < n) return -1.0;
  // compact permutation: rank of every original variable in the plan's (padded) elimination order
  std::vector<int> by_new(n), rank(n);
  for (int v = 0; v < n; ++v) by_new[v] = v;
  std::sort(by_new.begin(), by_new.end(), [&](int a, int b) { return bp.newvar[a] < bp.newvar[b]; });
  for (int i = 0; i < n; ++i) rank[by_new[i]] = i;
  std::vector<cd> A((size_t)n * ld, cd(0, 0)), B((size_t)n * ld, cd(0, 0));
  for (int r = 0; r < n; ++r)
    for (int en = (*in.row_ptr)[r]; en < (*in.row_ptr)[r + 1]; ++en) {
      const int c = (*in.ent_col)[en];
      const cd v((*in.ent_alpha)[en] + (*in.ent_jre)[en], w * (*in.ent_beta)[en] - (*in.ent_gamma)[en] / w + (*in.ent_jim)[en]);
      A[(size_t)r * ld + c] += v;
      B[(size_t)rank[r] * ld + (c < n ? rank[c] : n)] += v;
    }
  std::vector<cd> xa, xb;
  if (!band_detail::solve_like_reference(n, A, xa) || !band_detail::solve_like_reference(n, B, xb)) return -1.0;
  double big = 0.0;
  for (int i = 0; i < n; ++i) big = std::max(big, std::abs(xa[i]));
  double dev = 0.0;
  for (int i = 0; i < n; ++i) dev = std::max(dev, std::abs(xb[rank[i]] - xa[i]) / std::max(std::abs(xa[i]), 1e-12 * big));
  return dev;
}

}  // namespace spicey
