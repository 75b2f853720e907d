// Banded + bordered complex LU, register-blocked: AC tier 8 (SPICEY_TIER_BAND).
//
// Replaces the same reference code as the other AC tiers — buildLinearSystemForAC (simulateAC.ts:24-60), solveComplex
// (lib/math/solveComplex.ts:4-73: elimination :15-53, back-substitution :55-72), unpack (simulateAC.ts:85-126) — for
// circuits whose MNA matrix, after a bandwidth-reducing renumbering of the nodes (band_plan.h), is a band of
// half-width <= W plus NB border rows / columns (the V-source branches): meshes, long ladders, transmission lines.
// cfg 4 (16 x 16 RC mesh, Nvar 257): W = 16, NB = 1.
//
// This file is a self-contained translation unit: the host embeds its text and compiles it with NVRTC once per
// (L, RPL, NB, ...) combination (spicey_native.cu: ensure_band_jit), and `nvcc -DBAND_L=8 ...` compiles it stand-alone
// for register / SASS inspection.  Parameters (macros):
//   BAND_L        lanes per system (power of two <= 32); 32 / BAND_L systems share a warp
//   BAND_RPL      window rows per lane; W = BAND_L * BAND_RPL is the window (a power of two, >= 2)
//   BAND_NB       border rows = border columns (0..4)
//   BAND_ABMASK   bit j set: border column j has structural non-zeros in band rows (the right-hand side always has)
//   BAND_IELEM    1: element currents are computed and stored
//   BAND_WARPS    warps per CTA,  BAND_MINB  CTAs per SM (launch bounds)
//
// Mapping.  One group of L lanes owns one system.  Pivot step k (column k, pivot row k of the pilot's order) touches
// the W rows k+1 .. k+W: row i lives in lane i mod L, row slot (i / L) mod RPL, as W band entries (column c in
// register slot c mod W), the border-column entries and the right-hand side — all in registers for the whole
// elimination.  The loop over k is unrolled W times so that every slot index is a compile-time constant; the window
// slides without moving a register: column k's slot is reused for column k + W, the pivot row's registers for the
// entering row k + W.  The pivot row is published to a small double-buffered shared-memory record (one STS.128 per
// entry by its owner, one broadcast LDS.128 per entry by everybody) — the only communication of a step besides the
// shuffle that broadcasts the border rows' multipliers.  Border rows are distributed column-wise over the lanes.
// U (the pivot rows), 1/u_kk and the eliminated right-hand side go to a per-group global workspace, written once,
// column-major in the band so that the column-oriented back-substitution reads one contiguous 16*W bytes per
// unknown.  Stamped values (simulateAC.ts:24-60) are delivered from per-topology tables of (alpha, Im J, beta, gamma)
// constants, value = alpha + j (w beta - gamma / w + Im J), where the window first needs them.
//
// Pivoting.  The pilot (band_plan.h) ran the reference's rule (solveComplex.ts:18-28: largest |a_ik|, first maximum
// wins, row swap) on one representative point; every system re-checks every step on its own numbers — every
// candidate row of the window and the border against the pilot's choice, strict or non-strict according to the
// candidate's position in the reference's scan order — and a system that disagrees anywhere, or trips a guard
// (|pivot|^2 < EPS: singular / Complex.div, inductor guards), is appended to the fallback list and re-solved by the
// dense pivoting kernel in the same stream, which also reports the exact status.
#ifndef BAND_L
#error "define BAND_L, BAND_RPL, BAND_NB, BAND_ABMASK, BAND_IELEM, BAND_WARPS, BAND_MINB"
#endif

#define BW (BAND_L * BAND_RPL)
#define BNB BAND_NB
#define BPS (BW + BNB + 2)          /* pivot record: W band entries | NB border columns | rhs | diagonal */
#define BGPW (32 / BAND_L)           /* systems per warp */
#define B_EPS 1e-15
#define B_THR 1e-30
#define B_TWO_PI 6.283185307179586

struct BandArgs {   // must match BandArgs in spicey_native.cu
  const double* freqs; long long p_count;
  double2* x; double2* ielem; int* status; long long series_ld;
  long long* fb_list; int* fb_count;
  double2* G; long long g_stride;       // per-group workspace, double2 units
  const double2* tab;                   // recipe tables, two double2 per entry: (alpha, Im J), (beta, gamma)
  const uint2* flags;                   // [n] strict-compare masks: .x window positions, .y border rows
  const int* newvar;                    // [n] original variable -> index in the elimination order
  const int2* el_idx;                   // [n_ac_elem] elimination-order indices of the element's nodes (-1 = ground)
  const double *el_a, *el_b, *el_g;     // element admittance constants
  const double* ind_L;
  int n, nb, n_ac_elem, v_first, n_ind;
  int o_init, o_initb, o_nc, o_lc, o_e0, o_erb, o_brd0, o_brdnc, o_bb0;   // table offsets (entries)
};

typedef double2 bcplx;

__device__ __forceinline__ double band_rcp(double a) {   // MUFU seed + two Newton steps (<= 1 ulp), no slow path
  double y, e;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  return y;
}
__device__ __forceinline__ bcplx band_mul(bcplx a, bcplx b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a - f * p
__device__ __forceinline__ bcplx band_submul(bcplx a, bcplx f, bcplx p) {
  return make_double2(fma(-f.x, p.x, fma(f.y, p.y, a.x)), fma(-f.x, p.y, fma(-f.y, p.x, a.y)));
}
__device__ __forceinline__ double band_mag(bcplx a) { return fma(a.x, a.x, a.y * a.y); }

// stamped value of a table entry at angular frequency w (simulateAC.ts:36-57)
__device__ __forceinline__ bcplx band_rec(const double2* t, int idx, double w, double iw) {
  const double2 c0 = __ldg(t + 2 * idx), c1 = __ldg(t + 2 * idx + 1);
  return make_double2(c0.x, fma(w, c1.x, -c1.y * iw) + c0.y);
}

extern __shared__ double2 band_sm[];

extern "C" __global__ void __launch_bounds__(BAND_WARPS * 32, BAND_MINB) spicey_band_jit(BandArgs a) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int l = lane & (BAND_L - 1);          // lane within the group
  const int g = lane / BAND_L;                // group within the warp
  const int n = a.n, nb = a.nb;
  // shared memory of this group: x in elimination order | two pivot records
  const int sys_stride = n + 2 * BPS;
  double2* xs = band_sm + (size_t)(wib * BGPW + g) * sys_stride;
  double2* Pb = xs + n;
  const long long n_groups = (long long)gridDim.x * BAND_WARPS * BGPW;
  const long long grp = ((long long)blockIdx.x * BAND_WARPS + wib) * BGPW + g;
  double2* Gu = a.G + grp * a.g_stride;                 // [(nb + W) columns][W]
  double2* Gb = Gu + (size_t)(nb + BW) * BW;            // [nb][NB + 1]
  double2* Gr = Gb + (size_t)nb * (BNB + 1);            // [nb]
  const double2* tab = a.tab;

  // Rows above the matrix (the first W unknowns have fewer than W rows above them): their U entries are read as
  // zeros by the back-substitution and never written by anybody, so one fill per launch is enough.
  for (int q = l; q < BW * BW; q += BAND_L) Gu[q] = make_double2(0.0, 0.0);
  for (int q = l; q < 2 * BPS; q += BAND_L) Pb[q] = make_double2(0.0, 0.0);
  __syncwarp();

  for (long long base = ((long long)blockIdx.x * BAND_WARPS + wib) * BGPW; base < a.p_count; base += n_groups) {
    const bool valid = base + g < a.p_count;
    const long long p = valid ? base + g : a.p_count - 1;
    const double w = B_TWO_PI * a.freqs[p];
    const double iw = 1.0 / w;
#define REC(idx) band_rec(tab, (idx), w, iw)
    bool bad = false;
    for (int q = l; q < a.n_ind; q += BAND_L) {   // inductor guards of simulateAC.ts:47-51: the dense kernel decides
      const double d = w * a.ind_L[q];
      bad = bad || (fabs(d) < B_EPS || d * d < B_EPS);
    }

    bcplx A[BAND_RPL][BW];            // band entries of my rows, slot = column mod W
    bcplx AB[BAND_RPL][BNB + 1];      // border columns, right-hand side
    bcplx BR[BNB > 0 ? BNB : 1][BAND_RPL];       // border rows: my columns (column c: lane c mod L, slot (c / L) mod RPL)
    bcplx BB[BNB > 0 ? BNB : 1][BNB + 1];        // border rows x (border columns | rhs), replicated in every lane
    bcplx XB[BNB > 0 ? BNB : 1];

    // ---- prologue: rows 0 .. W-1, columns 0 .. W-1 ----
#pragma unroll
    for (int q = 0; q < BAND_RPL; ++q) {
      const int i = l + BAND_L * q;
#pragma unroll
      for (int c = 0; c < BW; ++c) A[q][c] = REC(a.o_init + i * BW + c);
#pragma unroll
      for (int j = 0; j <= BNB; ++j)
        if (j == BNB || ((BAND_ABMASK >> j) & 1)) AB[q][j] = REC(a.o_initb + i * (BNB + 1) + j);
    }
#pragma unroll
    for (int b = 0; b < BNB; ++b) {
#pragma unroll
      for (int q = 0; q < BAND_RPL; ++q) BR[b][q] = REC(a.o_brd0 + b * BW + l + BAND_L * q);
#pragma unroll
      for (int j = 0; j <= BNB; ++j) BB[b][j] = REC(a.o_bb0 + b * (BNB + 1) + j);
    }
    __syncwarp();   // the previous system's readers of the pivot records and of xs are done
    if (l == 0) {   // row 0 is the first pivot row
      double2* P0 = Pb;
      P0[BW + BNB + 1] = A[0][0];
#pragma unroll
      for (int t = 1; t < BW; ++t) P0[t] = A[0][t];
      P0[0] = REC(a.o_nc + 0);          // a[0][W]: column W enters with step 0
#pragma unroll
      for (int j = 0; j <= BNB; ++j)
        if (j == BNB || ((BAND_ABMASK >> j) & 1)) P0[BW + j] = AB[0][j];
    }
    __syncwarp();

    // ---- elimination of the band columns ----
    for (int kb = 0; kb < nb; kb += BW) {
#pragma unroll
      for (int s = 0; s < BW; ++s) {
        const int k = kb + s;
        if (k < nb) {
          const int pl = s % BAND_L, rs = s / BAND_L;              // owner lane / row slot of rows = s (mod W)
          const int s1 = (s + 1) % BW, pl1 = s1 % BAND_L, rs1 = s1 / BAND_L;
          const double2* Pc = Pb + (s & 1) * BPS;                  // W is even: the parity of k is the parity of s
          double2* Pn = Pb + ((s + 1) & 1) * BPS;
          const uint2 fl = __ldg(a.flags + k);
          const bcplx dg = Pc[BW + BNB + 1];
          const double mp = band_mag(dg);
          bad = bad || !(mp >= B_EPS);            // singular / Complex.div guard / NaN
          const double inv = band_rcp(mp);
          const bcplx r = make_double2(dg.x * inv, -dg.y * inv);
          // multipliers of my rows; the new column k + W takes over the slot of column k
          bcplx F[BAND_RPL];
#pragma unroll
          for (int q = 0; q < BAND_RPL; ++q) {
            bcplx aik = A[q][s];
            A[q][s] = REC(a.o_nc + k * BW + l + BAND_L * q);
            if (q == rs && l == pl) {   // the pivot row's registers become the entering row k + W (zero so far)
              aik = REC(a.o_e0 + k);
#pragma unroll
              for (int t = 0; t < BW; ++t) A[q][t] = make_double2(0.0, 0.0);
#pragma unroll
              for (int j = 0; j <= BNB; ++j)
                if (j == BNB || ((BAND_ABMASK >> j) & 1)) AB[q][j] = REC(a.o_erb + k * (BNB + 1) + j);
            }
            // column k + 1: diagonal and lower entries of the rows k+1 .. k+W arrive one step before they are read
            { const bcplx lc = REC(a.o_lc + k * BW + l + BAND_L * q); A[q][s1].x += lc.x; A[q][s1].y += lc.y; }
            const double m = band_mag(aik);
            const bool strict = (fl.x >> (l + BAND_L * q)) & 1u;
            bad = bad || (strict ? !(m < mp) : (m > mp));          // solveComplex.ts:18-28: first maximum wins
            bcplx f = band_mul(aik, r);
            if (band_mag(f) < B_THR) f = make_double2(0.0, 0.0);   // solveComplex.ts:46
            F[q] = f;
          }
          // border rows: column k's entry sits in lane pl, slot rs
          bcplx FB[BNB > 0 ? BNB : 1];
#pragma unroll
          for (int b = 0; b < BNB; ++b) {
            const bcplx bk = BR[b][rs];
            const double m = band_mag(bk);
            if (l == pl) {
              const bool strict = (fl.y >> b) & 1u;
              bad = bad || (strict ? !(m < mp) : (m > mp));
              BR[b][rs] = REC(a.o_brdnc + k * BNB + b);
            }
            bcplx f = band_mul(bk, r);
            f.x = __shfl_sync(FULL, f.x, pl, BAND_L);
            f.y = __shfl_sync(FULL, f.y, pl, BAND_L);
            if (band_mag(f) < B_THR) f = make_double2(0.0, 0.0);
            FB[b] = f;
          }
          // the update: a_ij -= f_i * u_kj
#pragma unroll
          for (int t = 0; t < BW; ++t) {
            const bcplx pt = Pc[t];
#pragma unroll
            for (int q = 0; q < BAND_RPL; ++q) A[q][t] = band_submul(A[q][t], F[q], pt);
          }
#pragma unroll
          for (int j = 0; j <= BNB; ++j)
            if (j == BNB || ((BAND_ABMASK >> j) & 1)) {
              const bcplx pt = Pc[BW + j];
#pragma unroll
              for (int q = 0; q < BAND_RPL; ++q) AB[q][j] = band_submul(AB[q][j], F[q], pt);
#pragma unroll
              for (int b = 0; b < BNB; ++b) BB[b][j] = band_submul(BB[b][j], FB[b], pt);
            }
          // my columns of the pivot row: border rows' update, and U leaves for the workspace (column-major)
#pragma unroll
          for (int q = 0; q < BAND_RPL; ++q) {
            const int slot = l + BAND_L * q;
            const bcplx pt = Pc[slot];
#pragma unroll
            for (int b = 0; b < BNB; ++b) BR[b][q] = band_submul(BR[b][q], FB[b], pt);
            const int c = k + 1 + ((slot - s - 1) & (BW - 1));
            Gu[(size_t)c * BW + s] = pt;
          }
          if (l == pl) {
            Gr[k] = r;
#pragma unroll
            for (int j = 0; j <= BNB; ++j)
              if (j == BNB || ((BAND_ABMASK >> j) & 1)) Gb[(size_t)k * (BNB + 1) + j] = Pc[BW + j];
          }
          // row k + 1 is final: its owner publishes it as the next pivot record
          if (l == pl1) {
            Pn[BW + BNB + 1] = A[rs1][s1];
#pragma unroll
            for (int t = 0; t < BW; ++t)
              if (t != s1) Pn[t] = A[rs1][t];
            Pn[s1] = REC(a.o_nc + (k + 1) * BW + s1);    // a[k+1][k+1+W]
#pragma unroll
            for (int j = 0; j <= BNB; ++j)
              if (j == BNB || ((BAND_ABMASK >> j) & 1)) Pn[BW + j] = AB[rs1][j];
          }
          __syncwarp();
        }
      }
    }

    // ---- the border block: NB x NB, replicated; same verification ----
    bcplx RB[BNB > 0 ? BNB : 1];
#pragma unroll
    for (int b = 0; b < BNB; ++b) {
      const uint2 fl = __ldg(a.flags + nb + b);
      const double mp = band_mag(BB[b][b]);
      bad = bad || !(mp >= B_EPS);
      const double inv = band_rcp(mp);
      RB[b] = make_double2(BB[b][b].x * inv, -BB[b][b].y * inv);
#pragma unroll
      for (int b2 = b + 1; b2 < BNB; ++b2) {
        const double m = band_mag(BB[b2][b]);
        const bool strict = (fl.y >> b2) & 1u;
        bad = bad || (strict ? !(m < mp) : (m > mp));
        bcplx f = band_mul(BB[b2][b], RB[b]);
        if (band_mag(f) < B_THR) f = make_double2(0.0, 0.0);
#pragma unroll
        for (int j = b + 1; j <= BNB; ++j) BB[b2][j] = band_submul(BB[b2][j], f, BB[b][j]);
      }
    }
#pragma unroll
    for (int b = BNB - 1; b >= 0; --b) {
      bcplx acc = BB[b][BNB];
#pragma unroll
      for (int j = b + 1; j < BNB; ++j) acc = band_submul(acc, BB[b][j], XB[j]);
      XB[b] = band_mul(acc, RB[b]);
      if (l == 0) xs[nb + b] = XB[b];
    }
    __syncwarp();   // the workspace written by other lanes of the group is visible

    // ---- back-substitution, column oriented: x_j by its owner, then every row of column j takes its term ----
    // row i: lane i mod L, slot (i / L) mod RPL, accumulator = b_i - sum over the border columns - sum_j u_ij x_j
    auto row_rhs = [&](int i) -> bcplx {
      bcplx acc = Gb[(size_t)i * (BNB + 1) + BNB];
#pragma unroll
      for (int j = 0; j < BNB; ++j)
        if ((BAND_ABMASK >> j) & 1) acc = band_submul(acc, Gb[(size_t)i * (BNB + 1) + j], XB[j]);
      return acc;
    };
    bcplx ACC[BAND_RPL];
#pragma unroll
    for (int q = 0; q < BAND_RPL; ++q) {
      const int slot = l + BAND_L * q;
      const int i = nb - 1 - ((nb - 1 - slot) & (BW - 1));    // the row = slot (mod W) among nb-W .. nb-1
      ACC[q] = i >= 0 ? row_rhs(i) : make_double2(0.0, 0.0);
    }
    for (int jb = (nb - 1) / BW * BW; jb >= 0; jb -= BW) {
      // One batch of loads in front of BW dependent steps: the block's U columns, and — for the steps this lane owns
      // (s = l + L q) — the reciprocal of the pivot and the accumulator of the row that enters there.
      bcplx U[BW][BAND_RPL];
      bcplx RJ[BAND_RPL], ENT[BAND_RPL];
#pragma unroll
      for (int s = 0; s < BW; ++s) {
#pragma unroll
        for (int q = 0; q < BAND_RPL; ++q) U[s][q] = Gu[(size_t)(jb + s) * BW + l + BAND_L * q];   // jb + s < nb + W
      }
#pragma unroll
      for (int q = 0; q < BAND_RPL; ++q) {
        const int j = jb + l + BAND_L * q;
        RJ[q] = j < nb ? Gr[j] : make_double2(0.0, 0.0);
        ENT[q] = (j < nb && j - BW >= 0) ? row_rhs(j - BW) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int s = BW - 1; s >= 0; --s) {
        const int j = jb + s;
        if (j < nb) {
          const int pl = s % BAND_L, rs = s / BAND_L;
          bcplx xj = band_mul(ACC[rs], RJ[rs]);
          xj.x = __shfl_sync(FULL, xj.x, pl, BAND_L);
          xj.y = __shfl_sync(FULL, xj.y, pl, BAND_L);
          if (l == pl) {
            xs[j] = xj;
            ACC[rs] = ENT[rs];   // row j - W enters
          }
#pragma unroll
          for (int q = 0; q < BAND_RPL; ++q) ACC[q] = band_submul(ACC[q], U[s][q], xj);
        }
      }
    }
    __syncwarp();

    // ---- status, results (simulateAC.ts:85-126) ----
    const unsigned vote = __ballot_sync(FULL, bad);
    const unsigned gmask = BAND_L == 32 ? FULL : (((1u << BAND_L) - 1u) << (g * BAND_L));
    const bool good = (vote & gmask) == 0u;
    if (valid) {
      if (l == 0) {
        a.status[p] = good ? 0 : -1;
        if (!good) a.fb_list[atomicAdd(a.fb_count, 1)] = p;
      }
      const long long xst = a.series_ld ? a.series_ld : 1;
      double2* xo = a.series_ld ? a.x + p : a.x + p * n;
      for (int i = l; i < n; i += BAND_L) xo[(long long)i * xst] = xs[__ldg(a.newvar + i)];
#if BAND_IELEM
      double2* io = a.series_ld ? a.ielem + p : a.ielem + p * a.n_ac_elem;
      for (int e = l; e < a.n_ac_elem; e += BAND_L) {
        bcplx cur;
        if (e >= a.v_first) cur = xs[nb + e - a.v_first];
        else {
          const int2 en = __ldg(a.el_idx + e);
          const bcplx v1 = en.x >= 0 ? xs[en.x] : make_double2(0.0, 0.0);
          const bcplx v2 = en.y >= 0 ? xs[en.y] : make_double2(0.0, 0.0);
          const bcplx Y = make_double2(__ldg(a.el_a + e), fma(w, __ldg(a.el_b + e), -__ldg(a.el_g + e) * iw));
          cur = band_mul(Y, make_double2(v1.x - v2.x, v1.y - v2.y));
        }
        io[(long long)e * xst] = cur;
      }
#endif
    }
#undef REC
  }
}
