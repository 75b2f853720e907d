// CPU check of the banded + bordered AC tier's host plan (spicey_b200/csrc/band_plan.h) and of the schedule the
// kernel (spicey_b200/csrc/band_kernel.cuh) executes: an emulator with the kernel's data placement — L lanes per
// system, RPL rows per lane, column slots c mod W, the published pivot record, the per-step deliveries, the
// column-major U workspace, the column-oriented back-substitution — runs one system on the CPU, lane by lane,
// and the result is compared with dense Gaussian elimination with partial pivoting in the ORIGINAL unknown order
// (the reference's algorithm, solveComplex.ts:4-73).
//   usage: band_plan_check <kind> <seed> <size> [nV] [L RPL]
//     kind 0: side x side RC mesh, V source at a corner, node ids in order of first appearance (cfg 4's numbering)
//     kind 1: random banded RLC-like system, nodes shuffled, nV sources between random nearby nodes
//     kind 2: RC ladder of <size> nodes
//     kind 3: as 1 with inductors (pivot order depends on the frequency: most points are flagged for the fallback)
// Prints "OK err=<max rel err> W=<..> L=<..> RPL=<..> bw=<..> renumbered=<..> bad=<..>" or "FAIL ...".
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include "../../spicey_b200/csrc/band_plan.h"
using namespace spicey;
typedef std::complex<double> cd;

struct Sys {
  int n = 0, nn = 0, nV = 0;
  std::map<std::pair<int, int>, BandRecipe> ent;   // (row, col) -> constants; col n = rhs
  void add(int r, int c, double a, double b, double g, double jre = 0, double jim = 0) {
    BandRecipe& q = ent[{r, c}];
    q.alpha_jre += a + jre; q.jim += jim; q.beta += b; q.gamma += g;
  }
  void admittance(int n1, int n2, double a, double b, double g) {   // node ids, 0 = ground
    const int i1 = n1 - 1, i2 = n2 - 1;
    if (i1 >= 0) add(i1, i1, a, b, g);
    if (i2 >= 0) add(i2, i2, a, b, g);
    if (i1 >= 0 && i2 >= 0) { add(i1, i2, -a, -b, -g); add(i2, i1, -a, -b, -g); }
  }
  void vsource(int n1, int n2, int k, double re, double im) {
    const int j = nn + k, i1 = n1 - 1, i2 = n2 - 1;
    if (i1 >= 0) { add(i1, j, 1, 0, 0); add(j, i1, 1, 0, 0); }
    if (i2 >= 0) { add(i2, j, -1, 0, 0); add(j, i2, -1, 0, 0); }
    add(j, n, 0, 0, 0, re, im);
  }
};

static cd value(const BandRecipe& q, double w) { return cd(q.alpha_jre, w * q.beta - q.gamma / w + q.jim); }

static bool dense_solve(int n, std::vector<cd> A, std::vector<cd> b, std::vector<cd>& x) {
  std::vector<int> rows(n);
  for (int i = 0; i < n; ++i) rows[i] = i;
  for (int k = 0; k < n; ++k) {
    int imax = k; double vmax = std::abs(A[(size_t)rows[k] * n + k]);
    for (int i = k + 1; i < n; ++i) { double v = std::abs(A[(size_t)rows[i] * n + k]); if (v > vmax) { vmax = v; imax = i; } }
    if (vmax < 1e-15) return false;
    std::swap(rows[k], rows[imax]);
    for (int i = k + 1; i < n; ++i) {
      cd f = A[(size_t)rows[i] * n + k] / A[(size_t)rows[k] * n + k];
      if (std::abs(f) < 1e-15) continue;
      for (int j = k; j < n; ++j) A[(size_t)rows[i] * n + j] -= f * A[(size_t)rows[k] * n + j];
      b[rows[i]] -= f * b[rows[k]];
    }
  }
  x.assign(n, cd(0, 0));
  for (int i = n - 1; i >= 0; --i) {
    cd s = b[rows[i]];
    for (int j = i + 1; j < n; ++j) s -= A[(size_t)rows[i] * n + j] * x[j];
    x[i] = s / A[(size_t)rows[i] * n + i];
  }
  return true;
}

// The same elimination in extended precision: tells a rounding-level difference from a wrong schedule.
typedef std::complex<long double> cld;
static void dense_solve_ld(int n, std::vector<cld> A, std::vector<cld> b, std::vector<cld>& x) {
  std::vector<int> rows(n);
  for (int i = 0; i < n; ++i) rows[i] = i;
  for (int k = 0; k < n; ++k) {
    int imax = k; long double vmax = std::abs(A[(size_t)rows[k] * n + k]);
    for (int i = k + 1; i < n; ++i) { long double v = std::abs(A[(size_t)rows[i] * n + k]); if (v > vmax) { vmax = v; imax = i; } }
    std::swap(rows[k], rows[imax]);
    for (int i = k + 1; i < n; ++i) {
      cld f = A[(size_t)rows[i] * n + k] / A[(size_t)rows[k] * n + k];
      if (f == cld(0, 0)) continue;
      for (int j = k; j < n; ++j) A[(size_t)rows[i] * n + j] -= f * A[(size_t)rows[k] * n + j];
      b[rows[i]] -= f * b[rows[k]];
    }
  }
  x.assign(n, cld(0, 0));
  for (int i = n - 1; i >= 0; --i) {
    cld s = b[rows[i]];
    for (int j = i + 1; j < n; ++j) s -= A[(size_t)rows[i] * n + j] * x[j];
    x[i] = s / A[(size_t)rows[i] * n + i];
  }
}

// The kernel's schedule, lane by lane.  Returns the solution in elimination order; `bad` = a verification failed.
static std::vector<cd> emulate(const BandPlan& bp, double w, bool& bad) {
  const int L = bp.L, RPL = bp.RPL, W = bp.W, NB = bp.NB, n = bp.n, nb = bp.nb, PS = W + NB + 2;
  const double EPS = 1e-15, THR = 1e-30;
  auto REC = [&](int idx) { return value(bp.tab[idx], w); };
  auto mag = [](cd z) { return z.real() * z.real() + z.imag() * z.imag(); };
  auto act = [&](int j) { return j == NB || ((bp.abmask >> j) & 1u); };
  std::vector<cd> A((size_t)L * RPL * W), AB((size_t)L * RPL * (NB + 1)), BR((size_t)L * std::max(1, NB) * RPL), BB((size_t)std::max(1, NB) * (NB + 1));
  auto a = [&](int l, int q, int t) -> cd& { return A[((size_t)l * RPL + q) * W + t]; };
  auto ab = [&](int l, int q, int j) -> cd& { return AB[((size_t)l * RPL + q) * (NB + 1) + j]; };
  auto br = [&](int l, int b, int q) -> cd& { return BR[((size_t)l * std::max(1, NB) + b) * RPL + q]; };
  auto bb = [&](int b, int j) -> cd& { return BB[(size_t)b * (NB + 1) + j]; };
  std::vector<cd> P(2 * PS, cd(0, 0)), xs(n, cd(0, 0));
  std::vector<cd> Gu((size_t)(nb + W) * W, cd(1e300, 1e300)), Gb((size_t)nb * (NB + 1), cd(1e300, 1e300)), Gr(nb, cd(1e300, 1e300));
  for (int q = 0; q < W * W; ++q) Gu[q] = cd(0, 0);
  bad = false;
  for (int l = 0; l < L; ++l)
    for (int q = 0; q < RPL; ++q) {
      const int i = l + L * q;
      for (int c = 0; c < W; ++c) a(l, q, c) = REC(bp.o_init + i * W + c);
      for (int j = 0; j <= NB; ++j) if (act(j)) ab(l, q, j) = REC(bp.o_initb + i * (NB + 1) + j);
      for (int b = 0; b < NB; ++b) br(l, b, q) = REC(bp.o_brd0 + b * W + l + L * q);
    }
  for (int b = 0; b < NB; ++b) for (int j = 0; j <= NB; ++j) bb(b, j) = REC(bp.o_bb0 + b * (NB + 1) + j);
  P[W + NB + 1] = a(0, 0, 0);
  for (int t = 1; t < W; ++t) P[t] = a(0, 0, t);
  const int ST = bp.step_stride, o_nc = 0, o_lc = W, o_e0 = 2 * W, o_erb = 2 * W + 1, o_brd = 2 * W + 1 + NB + 1;
  P[0] = REC(bp.o_step + o_nc + 0);
  for (int j = 0; j <= NB; ++j) if (act(j)) P[W + j] = ab(0, 0, j);
  for (int k = 0; k < nb; ++k) {
    const int s = k % W, pl = s % L, rs = s / L, s1 = (s + 1) % W, pl1 = s1 % L, rs1 = s1 / L;
    cd* Pc = &P[(k & 1) * PS];
    cd* Pn = &P[((k + 1) & 1) * PS];
    const int rk = bp.o_step + k * ST;
    unsigned long long fbits;   // the kernel reads a band step's tie-rule masks from the step's record
    memcpy(&fbits, &bp.tab[rk + 2 * W + 2 * NB + 2].alpha_jre, sizeof fbits);
    const unsigned fx = (unsigned)fbits, fy = (unsigned)(fbits >> 32);
    if (fx != bp.flags[2 * k] || fy != bp.flags[2 * k + 1]) { bad = true; printf("FAIL flags of step %d differ between record and array\n", k); exit(1); }
    const cd dg = Pc[W + NB + 1];
    const double mp = mag(dg);
    bad = bad || !(mp >= EPS);
    const cd r = cd(dg.real() / mp, -dg.imag() / mp);
    std::vector<cd> F((size_t)L * RPL), FB(std::max(1, NB));
    for (int l = 0; l < L; ++l)
      for (int q = 0; q < RPL; ++q) {
        cd aik = a(l, q, s);
        a(l, q, s) = REC(rk + o_nc + l + L * q);
        if (q == rs && l == pl) {
          aik = REC(rk + o_e0);
          for (int t = 0; t < W; ++t) a(l, q, t) = cd(0, 0);
          for (int j = 0; j <= NB; ++j) if (act(j)) ab(l, q, j) = REC(rk + o_erb + j);
        }
        a(l, q, s1) += REC(rk + o_lc + l + L * q);
        const double m = mag(aik);
        const bool strict = (fx >> (l + L * q)) & 1u;
        bad = bad || (strict ? !(m < mp) : (m > mp));
        cd f = aik * r;
        if (mag(f) < THR) f = cd(0, 0);
        F[(size_t)l * RPL + q] = f;
      }
    for (int b = 0; b < NB; ++b) {
      const cd bk = br(pl, b, rs);
      const double m = mag(bk);
      const bool strict = (fy >> b) & 1u;
      bad = bad || (strict ? !(m < mp) : (m > mp));
      br(pl, b, rs) = REC(rk + o_brd + b);
      cd f = bk * r;
      if (mag(f) < THR) f = cd(0, 0);
      FB[b] = f;
    }
    for (int l = 0; l < L; ++l)
      for (int q = 0; q < RPL; ++q) {
        const cd f = F[(size_t)l * RPL + q];
        for (int t = 0; t < W; ++t) a(l, q, t) -= f * Pc[t];
        for (int j = 0; j <= NB; ++j) if (act(j)) ab(l, q, j) -= f * Pc[W + j];
      }
    for (int j = 0; j <= NB; ++j) if (act(j)) for (int b = 0; b < NB; ++b) bb(b, j) -= FB[b] * Pc[W + j];
    for (int l = 0; l < L; ++l)
      for (int q = 0; q < RPL; ++q) {
        const int slot = l + L * q;
        const cd pt = Pc[slot];
        for (int b = 0; b < NB; ++b) br(l, b, q) -= FB[b] * pt;
        const int c = k + 1 + ((slot - s - 1) & (W - 1));
        Gu[(size_t)c * W + s] = pt;
      }
    Gr[k] = r;
    for (int j = 0; j <= NB; ++j) if (act(j)) Gb[(size_t)k * (NB + 1) + j] = Pc[W + j];
    Pn[W + NB + 1] = a(pl1, rs1, s1);
    for (int t = 0; t < W; ++t) if (t != s1) Pn[t] = a(pl1, rs1, t);
    Pn[s1] = REC(rk + ST + o_nc + s1);
    for (int j = 0; j <= NB; ++j) if (act(j)) Pn[W + j] = ab(pl1, rs1, j);
  }
  std::vector<cd> RB(std::max(1, NB)), XB(std::max(1, NB));
  for (int b = 0; b < NB; ++b) {
    const unsigned fy = bp.flags[2 * (nb + b) + 1];
    const double mp = mag(bb(b, b));
    bad = bad || !(mp >= EPS);
    RB[b] = cd(bb(b, b).real() / mp, -bb(b, b).imag() / mp);
    for (int b2 = b + 1; b2 < NB; ++b2) {
      const double m = mag(bb(b2, b));
      const bool strict = (fy >> b2) & 1u;
      bad = bad || (strict ? !(m < mp) : (m > mp));
      cd f = bb(b2, b) * RB[b];
      if (mag(f) < THR) f = cd(0, 0);
      for (int j = b + 1; j <= NB; ++j) bb(b2, j) -= f * bb(b, j);
    }
  }
  for (int b = NB - 1; b >= 0; --b) {
    cd acc = bb(b, NB);
    for (int j = b + 1; j < NB; ++j) acc -= bb(b, j) * XB[j];
    XB[b] = acc * RB[b];
    xs[nb + b] = XB[b];
  }
  auto row_rhs = [&](int i) {
    cd acc = Gb[(size_t)i * (NB + 1) + NB];
    for (int j = 0; j < NB; ++j) if ((bp.abmask >> j) & 1u) acc -= Gb[(size_t)i * (NB + 1) + j] * XB[j];
    return acc;
  };
  std::vector<cd> ACC((size_t)L * RPL);
  for (int l = 0; l < L; ++l)
    for (int q = 0; q < RPL; ++q) {
      const int slot = l + L * q;
      const int i = nb - 1 - ((nb - 1 - slot) & (W - 1));
      ACC[(size_t)l * RPL + q] = i >= 0 ? row_rhs(i) : cd(0, 0);
    }
  for (int jb = (nb - 1) / W * W; jb >= 0; jb -= W)
    for (int s = W - 1; s >= 0; --s) {
      const int j = jb + s;
      if (j >= nb) continue;
      const int pl = s % L, rs = s / L;
      const cd xj = ACC[(size_t)pl * RPL + rs] * Gr[j];
      xs[j] = xj;
      ACC[(size_t)pl * RPL + rs] = j - W >= 0 ? row_rhs(j - W) : cd(0, 0);
      for (int l = 0; l < L; ++l)
        for (int q = 0; q < RPL; ++q) ACC[(size_t)l * RPL + q] -= Gu[(size_t)j * W + l + L * q] * xj;
    }
  return xs;
}

int main(int argc, char** argv) {
  const int kind = argc > 1 ? atoi(argv[1]) : 0;
  const unsigned seed = argc > 2 ? atoi(argv[2]) : 1;
  const int size = argc > 3 ? atoi(argv[3]) : 16;
  const int nVarg = argc > 4 ? atoi(argv[4]) : 1;
  const int fL = argc > 6 ? atoi(argv[5]) : 0, fR = argc > 6 ? atoi(argv[6]) : 0;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> U(0.5, 1.5);
  Sys S;
  if (kind == 0) {   // mesh, node ids in order of first appearance (spicey_b200/workloads.py: rc_mesh)
    std::map<std::pair<int, int>, int> id;
    auto node = [&](int r, int c) { auto it = id.find({r, c}); if (it == id.end()) it = id.insert({{r, c}, (int)id.size() + 1}).first; return it->second; };
    node(0, 0);
    std::vector<std::pair<int, int>> res;
    for (int r = 0; r < size; ++r)
      for (int c = 0; c < size; ++c) {
        if (c + 1 < size) { int u = node(r, c), v = node(r, c + 1); res.push_back({u, v}); }
        if (r + 1 < size) { int u = node(r, c), v = node(r + 1, c); res.push_back({u, v}); }
      }
    S.nn = size * size; S.nV = 1; S.n = S.nn + 1;
    for (auto& e : res) S.admittance(e.first, e.second, 1e-3 * U(rng), 0, 0);
    for (int r = 0; r < size; ++r) for (int c = 0; c < size; ++c) if (r || c) S.admittance(node(r, c), 0, 0, 1e-9 * U(rng), 0);
    S.vsource(node(0, 0), 0, 0, 1.0, 0.0);
  } else if (kind == 1 || kind == 3) {   // random banded network with shuffled node ids (kind 3: with inductors)
    const bool rlc = kind == 3;
    const int nn = size, bwid = 6;
    S.nn = nn; S.nV = nVarg; S.n = nn + S.nV;
    std::vector<int> shuf(nn);
    for (int i = 0; i < nn; ++i) shuf[i] = i + 1;
    std::shuffle(shuf.begin(), shuf.end(), rng);
    for (int i = 0; i < nn; ++i) {
      S.admittance(shuf[i], 0, 1e-4 * U(rng), 1e-9 * U(rng), 0);
      for (int d = 1; d <= bwid && i + d < nn; ++d)
        if (d == 1 || rng() % 3 == 0) {
          const int ty = rlc ? rng() % 3 : rng() % 2;
          S.admittance(shuf[i], shuf[i + d], ty == 0 ? 1e-3 * U(rng) : 0, ty == 1 ? 1e-9 * U(rng) : 0, ty == 2 ? 1.0 / (1e-3 * U(rng)) : 0);
        }
    }
    for (int k = 0; k < S.nV; ++k) {
      const int i = rng() % nn, d = rng() % 3;
      const int n2 = (d == 0 || i + d >= nn || getenv("GROUNDED")) ? 0 : shuf[i + d];
      S.vsource(shuf[i], n2, k, U(rng), 0.3 * U(rng));
    }
  } else {   // ladder
    S.nn = size; S.nV = 1; S.n = size + 1;
    for (int k = 1; k < size; ++k) { S.admittance(k, k + 1, 1e-3, 0, 0); S.admittance(k + 1, 0, 0, 1e-9, 0); }
    S.vsource(1, 0, 0, 1.0, 0.0);
  }
  const int n = S.n;
  std::vector<int> row_ptr(n + 1, 0), col;
  std::vector<double> al, be, ga, jr, ji;
  int row = 0;
  for (auto& kv : S.ent) {
    while (row < kv.first.first) row_ptr[++row] = (int)col.size();
    col.push_back(kv.first.second);
    al.push_back(kv.second.alpha_jre); be.push_back(kv.second.beta); ga.push_back(kv.second.gamma); jr.push_back(0.0); ji.push_back(kv.second.jim);
  }
  while (row < n) row_ptr[++row] = (int)col.size();
  BandInput in;
  in.n = n; in.nn = S.nn; in.nV = S.nV; in.row_ptr = &row_ptr; in.ent_col = &col;
  in.ent_alpha = &al; in.ent_beta = &be; in.ent_gamma = &ga; in.ent_jre = &jr; in.ent_jim = &ji;
  in.pilot_w = 2 * M_PI * 300.0;
  BandPlan bp;
  build_band_plan(in, bp, fL, fR);
  if (!bp.ok) { printf("NOPLAN bw=%d\n", bp.bandwidth); return 2; }
  double worst = 0, worst_ref = 0;
  int nbad = 0, ngood = 0;
  const double fs[5] = {300.0, 1.0, 37.0, 4321.0, 1e5};   // the pilot's own frequency first: never flagged
  for (double f : fs) {
    const double w = 2 * M_PI * f;
    std::vector<cd> A((size_t)n * n, cd(0, 0)), b(n, cd(0, 0)), x;
    for (auto& kv : S.ent) {
      if (kv.first.second == n) b[kv.first.first] = value(kv.second, w);
      else A[(size_t)kv.first.first * n + kv.first.second] = value(kv.second, w);
    }
    if (!dense_solve(n, A, b, x)) { printf("FAIL dense singular\n"); return 1; }
    bool bad = false;
    std::vector<cd> xs = emulate(bp, w, bad);
    nbad += bad;
    if (bad && f == 300.0) { printf("FAIL the pilot's own point was flagged\n"); return 1; }
    if (bad) continue;   // the device hands such a system to the dense kernel
    ++ngood;
    std::vector<cld> Al((size_t)n * n), bl(n), xl;
    for (size_t q = 0; q < A.size(); ++q) Al[q] = cld(A[q].real(), A[q].imag());
    for (int i = 0; i < n; ++i) bl[i] = cld(b[i].real(), b[i].imag());
    dense_solve_ld(n, Al, bl, xl);
    double xmax = 0;
    for (int i = 0; i < n; ++i) xmax = std::max(xmax, std::abs(x[i]));
    for (int i = 0; i < n; ++i) {
      // against the extended-precision solution; the reference's own double-precision error is the yardstick
      const cd xt((double)xl[i].real(), (double)xl[i].imag());
      const double den = std::max(std::abs(xt), 1e-9 * xmax);
      const double e = std::abs(xs[bp.newvar[i]] - xt) / den, eref = std::abs(x[i] - xt) / den;
      if (!(e <= worst)) worst = e;
      if (!(eref <= worst_ref)) worst_ref = eref;
    }
  }
  // pass: within 1e-9 of the exact solution, or no worse than 4x the reference order's own rounding error
  const bool ok = (worst < 1e-9 || worst <= 4 * worst_ref) && ngood > 0;
  printf("%s err=%.3e ref_err=%.3e W=%d L=%d RPL=%d bw=%d renumbered=%d bad=%d cfma=%lld\n", ok ? "OK" : "FAIL", worst, worst_ref, bp.W, bp.L, bp.RPL,
         bp.bandwidth, (int)bp.renumbered, nbad, bp.n_cfma);
  return ok ? 0 : 1;
}
