// Dense batched LU with partial pivoting, the matrix of a system RESIDENT IN REGISTERS (tier 9, SPICEY_TIER_TILE).
//
// Replaces the per-frequency body of simulateAC (lib/analysis/simulateAC.ts:80-127) for circuits whose MNA matrix is
// not sparse enough for the program tiers, and every large batch run with SPICEY_FLAG_DENSE:
//   buildLinearSystemForAC :24-60  -> TL_CONST: every matrix entry from per-topology constants (alpha, Im J, beta, gamma),
//                                     value = alpha + j (w beta - gamma / w + Im J), read straight into the registers that
//                                     hold it; else (per-instance values) element admittances + gather stamping, one
//                                     thread per matrix ENTRY, through shared memory
//   solveComplex (lib/math/solveComplex.ts:15-53: pivot search, row swap, elimination) -> tl_search / the step loop
//   back-substitution :56-71       -> one warp, column oriented, from the factor written back to shared memory
//   unpack simulateAC.ts:85-126    -> x and element currents straight to HBM
// This text is compiled by NVRTC once per (Nvar, tile shape, variant) with the TL_* macros in front of it
// (spicey_native.cu: tile_source) and cached on disk (jit_runtime.h); band_kernel.cuh travels the same way.
//
// Why registers.  The one-thread-per-row kernel (lu_rowthread.cuh) keeps [A b] in shared memory: every complex FMA of
// the trailing update loads its operand and stores its result there (32 bytes of shared-memory traffic per 4 DFMA), so
// the SM's 128 B/clk shared-memory port caps it at a quarter of the FP64 pipe before any latency is counted (measured:
// 12.6 % of the dense roofline at Nvar = 65 with its structural-zero skipping, 0.7 % on a matrix without zeros).  Here the
// (Nvar) x (Nvar + 1) augmented matrix is distributed 2-D cyclically over a TR x TC grid of threads, each holding an
// MR x MC tile in registers (row i -> thread row i mod TR, local row i / TR; column j -> thread column j mod TC, local
// column j / TC).  A pivot step moves MR raw column entries and MC pivot-row entries per thread through shared memory
// (one 16-byte load each) for MR * MC complex FMAs.
//
// One CTA = one system at a time (persistent, strided over the points).  The reference swaps rows physically (:30-34);
// so does the kernel (row k and the pivot row exchange through shared memory when they differ), which keeps the live
// rows the trailing ones: the unrolled code of a segment of steps with the same (k / TR, k / TC) touches live local
// rows and columns only.
//
// The critical path of a step is the pivot search, and a lone warp runs ~6 cycles per dependent instruction (first
// version: search + reciprocal + multipliers in the owning warp, 300 instructions = 2,000 of the 2,300 cycles of a step,
// everybody else at the barrier).  Now: a thread column lives inside one warp; that warp publishes the raw column k,
// finds the pivot with three redux.sync over all its lanes (|a|^2 as the IEEE bit pattern, high word, low word, lowest
// row index among equals = the reference's first maximum) and publishes (p, 1/a_pk); every thread forms the multipliers
// of its own rows from the raw column (f_i = a_ik / a_pk, zeroed when |f| < EPS: :46).  And the search of step k + 1 is
// done DURING step k by the warp that owns column k + 1, in the same basic block as its share of the update (it first
// brings its column k + 1 up to date in temporaries), so that the search's dependent chain fills with the update's
// independent DFMAs instead of stalling the CTA.
// Arithmetic is the default policy of common.cuh (FMA contraction, |a|^2 pivot metric, reciprocal multiply; parity
// 1e-9); SPICEY_FLAG_STRICT calls stay with lu_rowthread.cuh.
//
// Shared memory per CTA: the column-major image of [A b] (U for the back-substitution; the stamping target without
// TL_CONST) | [element admittances, source phasors] | raw column (double-buffered) | pivot row, old row k | 1/u_kk | x |
// winner records.

typedef double2 tcplx;

#define TL_NC (TL_N + 1)
#define TL_MR ((TL_N + TL_TR - 1) / TL_TR)
#define TL_MC ((TL_NC + TL_TC - 1) / TL_TC)
#define TL_TPW (32 / TL_TR)            /* thread columns per warp */
#define TL_LD (TL_N | 1)               /* rows per column of the shared-memory image */
#define TL_NP (TL_MR * TL_TR)          /* padded rows */
#define TL_NCP (TL_MC * TL_TC)         /* padded columns */
#define TL_THREADS (TL_WARPS * 32)
#define TL_EPS 1e-15
#define TL_PI 3.141592653589793
#define TL_FULL 0xffffffffu

struct TileArgs {   // must match TileArgs in spicey_native.cu
  const double* freqs; long long n_freq, p_begin, p_count;
  double2* x; double2* ielem; int* status; long long series_ld;
  // per-instance stamping (TL_CONST 0)
  const int4* ends; const int2* meta; const double* values; const int* var_of_slot; const double* var_values; long long n_inst;
  const int* ent_rc;       // [n_ent] row | column << 16 of every structurally non-zero entry (column Nvar = right-hand side)
  const int* ent_ptr;      // [n_ent + 1] -> contrib
  const int* contrib;      // idx << 3 | src << 1 | neg (common.cuh)
  // constants of a plain frequency sweep (TL_CONST 1)
  const double2* ctab;     // [TL_MR * TL_MC][TL_THREADS] in tile order; TL_RC: (alpha, beta), else (alpha + Re J, Im J), (beta, gamma)
  const double4* el_rec;   // [n_ac_elem] {bits of (i1, i2), ya, yb, yg}: current = Y (x[i1] - x[i2]), Y = ya + j (w yb - yg / w);
                           // index Nvar = the ground node; V: (branch, Nvar, 1, 0, 0)
  const double* ind_L;     // inductances: the guards of simulateAC.ts:47-51 are checked per point, a point that trips one
  long long* fb_list; int* fb_count;   // ... is left to the one-thread-per-row kernel, which also reports the exact status
  const long long* plist; const int* pcount;   // optional: solve the launch-local points plist[0 .. *pcount) only (the
                                               // fallback list of a program tier), adding their number to fb_total[0 .. 1]
  unsigned long long* fb_total;
  int n_ind, n_ent, nn, nV, n_elem, n_ac_elem, off_v, off_v_end, off_i;
};

struct __align__(16) TileWin { int p; int pad0, pad1, pad2; };   // winner record of a step: the pivot row (0x7fffffff: the column is all zeros)

__device__ __forceinline__ tcplx tl_mul(tcplx a, tcplx b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ tcplx tl_submul(tcplx a, tcplx f, tcplx p) {
  return make_double2(fma(-f.x, p.x, fma(f.y, p.y, a.x)), fma(-f.x, p.y, fma(-f.y, p.x, a.y)));
}
__device__ __forceinline__ double tl_rcp(double a) {   // MUFU seed + two Newton steps (<= 1 ulp), no slow path
  double y, e;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  return y;
}
// Shared-memory stores of the look-ahead search: no "memory" clobber, so that the compiler may keep scheduling the
// update's loads and DFMAs around them (nothing in this thread reads these locations before the next barrier).
__device__ __forceinline__ void tl_sts(tcplx* p, tcplx v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v.x), "d"(v.y));
}
template <int OFF>   // ... with a compile-time byte offset folded into the instruction
__device__ __forceinline__ void tl_sts_at(unsigned base, tcplx v) {
  asm volatile("st.shared.v2.f64 [%0 + %3], {%1, %2};" ::"r"(base), "d"(v.x), "d"(v.y), "n"(OFF));
}
__device__ __forceinline__ void tl_sts1(void* p, int a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(a));
}

#if !TL_CONST
__device__ __forceinline__ double tl_value(const TileArgs& a, int slot, long long inst) {
  const int v = __ldg(a.var_of_slot + slot);
  return v < 0 ? __ldg(a.values + slot) : __ldg(a.var_values + (long long)v * a.n_inst + inst);
}
// Element admittance at frequency f (simulateAC.ts:36-52) and source phasor (:54-57); returns the status.
__device__ __forceinline__ int tl_element(const TileArgs& a, int type, int vidx, long long inst, double f, tcplx& Y, tcplx& J) {
  const double twoPi = 2 * TL_PI;
  Y = make_double2(0.0, 0.0);
  J = make_double2(0.0, 0.0);
  if (type == 0) {          // R
    const double R = tl_value(a, vidx, inst);
    if (R <= 0) return 3;   // :37
    Y.x = 1 / R;
  } else if (type == 1) {   // C: twoPi * f * c.C  :43
    Y.y = __dmul_rn(__dmul_rn(twoPi, f), tl_value(a, vidx, inst));
  } else if (type == 2) {   // L
    const double d = __dmul_rn(__dmul_rn(twoPi, f), tl_value(a, vidx, inst));
    if (fabs(d) < TL_EPS) return 0;     // denom.abs() < EPS -> Y = 0  :49
    const double dd = __dmul_rn(d, d);
    if (dd < TL_EPS) return 2;          // Complex.div guard (Complex.ts:41-42)
    Y.x = 0.0 / dd;
    Y.y = (0.0 - d) / dd;
  } else if (type == 3 || type == 6) {  // V, I: phasor fromPolar(acMag, acPhaseDeg)  Complex.ts:16-19
    const double mag = tl_value(a, vidx + 1, inst), deg = tl_value(a, vidx + 2, inst);
    const double ph = (deg * TL_PI) / 180;
    double s, c;
    sincos(ph, &s, &c);
    J.x = mag * c;
    J.y = mag * s;
  }
  return 0;
}
#endif

struct TileCtx {   // what the step functions need besides the tile
  tcplx *Cb, *Pb, *Kb, *Rd;
  TileWin* Wn;
  int tr, tc, lane, warp;
  bool act;
};

// Pivot search of step k (solveComplex.ts:17-29) on the raw column `col` (local rows KR..), executed by all 32 lanes of
// the warp that owns column k; `owner`: this lane holds entries of the column.  Publishes the raw column (everybody's
// reciprocal and multipliers come from it) and the pivot row p.  The dependent chain is what the whole CTA waits for, so
// it carries nothing else: |a|^2 per candidate, a first-maximum scan in ascending row order (a NaN wins only in place,
// as `v > vmax` never holds for it in JavaScript), redux.sync.max over the high words of the IEEE bit patterns and, when a
// single lane holds that maximum (the usual case), one shuffle of its row index; ties go through the low words and the
// lowest row index.
template <int KR, int M>
__device__ __forceinline__ void tl_search_rows(const TileCtx& t, int k, const tcplx (&col)[TL_MR], bool owner, unsigned cb,
                                               double& bm, int& bi, bool& nanp) {
  if (M >= TL_MR) return;
  constexpr int MM = M < TL_MR ? M : TL_MR - 1;
  const int i = MM * TL_TR + t.tr;
  if (owner) tl_sts_at<MM * TL_TR * 16>(cb, col[MM]);
  const tcplx v = col[MM];
  const double mt = fma(v.x, v.x, v.y * v.y);
  const bool in = owner && i >= k && i < TL_N;
  if (in && mt > bm) { bm = mt; bi = i; }          // ascending i, strict >: the first maximum (a NaN never passes)
  nanp = nanp || (in && i == k && mt != mt);       // ... except in place: JS keeps a NaN vmax (v > NaN never holds)
  tl_search_rows<KR, (M < TL_MR ? M + 1 : M)>(t, k, col, owner, cb, bm, bi, nanp);
}

template <int KR>
__device__ __forceinline__ void tl_search(const TileCtx& t, int k, const tcplx (&col)[TL_MR], bool owner) {
  const int par = k & 1;
  const unsigned cb = (unsigned)__cvta_generic_to_shared(t.Cb + par * (TL_NP + 1) + t.tr);
  double bm = -1.0;
  int bi = 0x7fffffff;
  bool nanp = false;
  tl_search_rows<KR, KR>(t, k, col, owner, cb, bm, bi, nanp);
  unsigned long long key = bi == 0x7fffffff ? 0ull : (unsigned long long)__double_as_longlong(bm);
  if (nanp) { key = ~0ull; bi = k; }
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(TL_FULL, hi);
  const unsigned tie = __ballot_sync(TL_FULL, hi == mh);
  int p;
  if (__popc(tie) == 1) {
    p = __shfl_sync(TL_FULL, bi, __ffs(tie) - 1);
  } else {
    const unsigned ml = __reduce_max_sync(TL_FULL, hi == mh ? lo : 0u);
    p = __reduce_min_sync(TL_FULL, (hi == mh && lo == ml) ? bi : 0x7fffffff);
  }
  if (t.lane == 0) tl_sts1(&t.Wn[par].p, p);
}

// The update a_ij -= f_i * u_kj (:47-52) of step k on the live part of the tile.  `own`: this warp owns column k + 1; it
// first brings that column up to date in temporaries and runs the search of step k + 1 on them (the update below then
// recomputes the same values in place: one code path for every warp, so the tile's registers need no reconciling
// moves where paths would merge).
template <int KR, int KC>
__device__ __forceinline__ void tl_update(const TileCtx& t, int k, int p, bool own, tcplx (&A)[TL_MR][TL_MC], const tcplx (&F)[TL_MR]) {
  const int trk = k - KR * TL_TR;
  const bool patch = p != k && t.tr == trk;      // position k takes the pivot row (F of row k is 0)
  if (own) {   // (column k + 1 is local column KC of its thread column: the first column of a block is searched at the block's start)
    const tcplx pv1 = t.Pb[KC * TL_TC + t.tc];
    tcplx col1[TL_MR];
#pragma unroll
    for (int m = KR; m < TL_MR; ++m) col1[m] = tl_submul(A[m][KC], F[m], pv1);
    tl_search<KR>(t, k + 1, col1, t.act && t.tc == k + 1 - KC * TL_TC);
  }
#pragma unroll
  for (int c = KC; c < TL_MC; ++c) {
    const tcplx pv = t.Pb[c * TL_TC + t.tc];
    if (patch) A[KR][c] = pv;
#pragma unroll
    for (int m = KR; m < TL_MR; ++m) A[m][c] = tl_submul(A[m][c], F[m], pv);
  }
}

// Local row mp (uniform across the CTA) of the tile, columns KC.., to / from a shared-memory row: a chain of uniform
// branches, one taken, instead of a select per register.
template <int KR, int KC, int M>
__device__ __forceinline__ void tl_row_out(const tcplx (&A)[TL_MR][TL_MC], int mp, bool mine, tcplx* row, int tc) {
  if (M >= TL_MR) return;
  constexpr int MM = M < TL_MR ? M : TL_MR - 1;
  if (mp == M) {
    if (mine) {
#pragma unroll
      for (int c = KC; c < TL_MC; ++c) row[c * TL_TC + tc] = A[MM][c];
    }
  } else {
    tl_row_out<KR, KC, (M < TL_MR ? M + 1 : M)>(A, mp, mine, row, tc);
  }
}
template <int KR, int KC, int M>
__device__ __forceinline__ void tl_row_in(tcplx (&A)[TL_MR][TL_MC], int mp, bool mine, const tcplx* row, int tc) {
  if (M >= TL_MR) return;
  constexpr int MM = M < TL_MR ? M : TL_MR - 1;
  if (mp == M) {
    if (mine) {
#pragma unroll
      for (int c = KC; c < TL_MC; ++c) A[MM][c] = row[c * TL_TC + tc];
    }
  } else {
    tl_row_in<KR, KC, (M < TL_MR ? M + 1 : M)>(A, mp, mine, row, tc);
  }
}

// The steps whose row k is local row KR of its thread row and whose column k is local column KC of its thread column:
// one rolled loop, unrolled over the live part of the tile.
template <int KR, int KC>
__device__ __forceinline__ void tl_segment(const TileCtx& t, tcplx (&A)[TL_MR][TL_MC], int& status) {
  constexpr int k0 = (KR * TL_TR > KC * TL_TC) ? KR * TL_TR : KC * TL_TC;
  constexpr int k1a = ((KR + 1) * TL_TR < (KC + 1) * TL_TC) ? (KR + 1) * TL_TR : (KC + 1) * TL_TC;
  constexpr int k1 = k1a > TL_N ? TL_N : k1a;
  if (k0 >= k1) return;
  const int tr = t.tr, tc = t.tc, warp = t.warp;
  tcplx *Cb = t.Cb, *Pb = t.Pb, *Kb = t.Kb, *Rd = t.Rd;
  if (k0 == KC * TL_TC && status == 0) {   // column k0 opens local column KC: nobody searched it during step k0 - 1
    if (warp == 0) {
      tcplx col0[TL_MR];
#pragma unroll
      for (int m = KR; m < TL_MR; ++m) col0[m] = A[m][KC];
      tl_search<KR>(t, k0, col0, t.act && tc == 0);
    }
  }
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    if (status != 0) break;
    const int trk = k - KR * TL_TR, par = k & 1;
    __syncthreads();   // the search of step k is published; everybody has finished step k - 1
    const tcplx* Ck = Cb + par * (TL_NP + 1);
    // warp-uniform by construction (one shared-memory word): telling the compiler so turns the row selections below
    // into uniform branches instead of MR-way select chains
    const int p = __reduce_min_sync(TL_FULL, t.Wn[par].p);
    if (p == 0x7fffffff) { status = 1; break; }          // no candidate above zero: singular (:29)
    // -- 1 / a_pk by everybody (:45), and the reference's two guards on the pivot --
    const tcplx apk = Ck[p];
    const double vmax = fma(apk.x, apk.x, apk.y * apk.y);   // |a_pk|^2
    if (vmax < TL_EPS * TL_EPS) { status = 1; break; }    // |a_pk| < EPS: singular (:29)
    if (vmax < TL_EPS) { status = 2; break; }             // Complex.div by this pivot throws (Complex.ts:41-42)
    const double inv = tl_rcp(vmax);
    const tcplx rk = make_double2(apk.x * inv, -apk.y * inv);
    const int trp = p % TL_TR, mp = p / TL_TR;
    if (t.lane == 0 && t.warp == 0) Rd[k] = rk;
    // -- the pivot row and (when they differ) the old row k through shared memory: the swap of :30-34 --
    tl_row_out<KR, KC, KR>(A, mp, tr == trp, Pb, tc);
    if (p != k) {
      if (tr == trk) {
#pragma unroll
        for (int c = KC; c < TL_MC; ++c) Kb[c * TL_TC + tc] = A[KR][c];
      }
    }
    __syncthreads();
    // -- my rows' multipliers f_i = a_ik / a_pk from the raw column (:45-46): position p holds the old row k, finished
    //    rows and padding read the zero behind the column --
    tcplx F[TL_MR];
#pragma unroll
    for (int m = KR; m < TL_MR; ++m) {
      const int i = m * TL_TR + tr;
      int idx = (i == p) ? k : i;
      if (i <= k || i >= TL_N) idx = TL_NP;
      tcplx fm = tl_mul(Ck[idx], rk);
      if (fma(fm.x, fm.x, fm.y * fm.y) < TL_EPS * TL_EPS) fm = make_double2(0.0, 0.0);   // :46
      F[m] = fm;
    }
    if (p != k) tl_row_in<KR, KC, KR>(A, mp, tr == trp, Kb, tc);
    // column k + 1 belongs to thread column (k + 1) - KC TC while it is local column KC; the warp that owns it searches it
    // during this step (the first column of the next block is searched at that block's start: it sits in other registers)
    const int tck1 = k + 1 - KC * TL_TC;
    tl_update<KR, KC>(t, k, p, k + 1 < TL_N && tck1 < TL_TC && warp == tck1 / TL_TPW, A, F);
  }
}

// compile-time walk over the (KR, KC) segments in step order
template <int KR, int KC>
struct TlSeg {
  static __device__ __forceinline__ void run(const TileCtx& t, tcplx (&A)[TL_MR][TL_MC], int& status) {
    tl_segment<KR, KC>(t, A, status);
    // the next segment starts where this one ends: at a row-block boundary, a column-block boundary, or both
    constexpr int er = (KR + 1) * TL_TR, ec = (KC + 1) * TL_TC;
    constexpr int NR = er <= ec ? KR + 1 : KR, NCc = ec <= er ? KC + 1 : KC;
    TlSeg<(NR < TL_MR && NCc < TL_MC) ? NR : TL_MR, (NR < TL_MR && NCc < TL_MC) ? NCc : TL_MC>::run(t, A, status);
  }
};
template <>
struct TlSeg<TL_MR, TL_MC> {
  static __device__ __forceinline__ void run(const TileCtx&, tcplx (&)[TL_MR][TL_MC], int&) {}
};

extern __shared__ __align__(16) unsigned char tl_smem[];

extern "C" __global__ void __launch_bounds__(TL_THREADS, TL_MINB) spicey_tile_jit(const TileArgs a) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TileCtx t;
  t.lane = lane; t.warp = warp;
  t.tr = lane % TL_TR;
  const int tcl = lane / TL_TR;
  t.act = tcl < TL_TPW && warp * TL_TPW + tcl < TL_TC;
  // lanes without a thread column of their own mirror the last one: same registers, same (duplicate) shared-memory
  // stores, no part in the pivot search and no global stores
  t.tc = min(warp * TL_TPW + tcl, TL_TC - 1);
  const int tr = t.tr, tc = t.tc;

  tcplx* Aimg = (tcplx*)tl_smem;                       // [TL_NC][TL_LD]
#if TL_CONST
  tcplx* Cb = Aimg + (size_t)TL_NC * TL_LD;            // [2][TL_NP + 1]
#else
  const int n_src = a.nV + (a.n_elem - a.off_i);
  tcplx* Yv = Aimg + (size_t)TL_NC * TL_LD;            // [n_elem]
  tcplx* Jv = Yv + a.n_elem;                           // [n_src]
  tcplx* Cb = Jv + (n_src > 0 ? n_src : 1);            // [2][TL_NP + 1]
#endif
  tcplx* Pb = Cb + 2 * (TL_NP + 1);                    // [TL_NCP]   (Cb: [2][TL_NP + 1], the last entry of each a zero)
  tcplx* Kb = Pb + TL_NCP;                             // [TL_NCP]
  tcplx* Rd = Kb + TL_NCP;                             // [TL_N]
  tcplx* xs = Rd + TL_N;                               // [TL_N + 1]  (xs[TL_N] = 0: the ground node)
  TileWin* Wn = (TileWin*)(xs + TL_N + 1);             // [2]
#if !TL_CONST
  int* s_status = (int*)(Wn + 2);
#endif
  t.Cb = Cb; t.Pb = Pb; t.Kb = Kb; t.Rd = Rd; t.Wn = Wn;
  if (tid < 2) Cb[tid * (TL_NP + 1) + TL_NP] = make_double2(0.0, 0.0);   // (visible after the first barrier of the point loop)

  const long long work = a.plist ? (long long)*a.pcount : a.p_count;
  if (a.plist && a.fb_total && blockIdx.x == 0 && tid == 0 && work > 0) {
    atomicAdd(a.fb_total, (unsigned long long)work);
    atomicAdd(a.fb_total + 1, (unsigned long long)work);
  }
  for (long long qi = blockIdx.x; qi < work; qi += gridDim.x) {
    const long long q = a.plist ? a.plist[qi] : qi;
    const long long p_abs = a.p_begin + q;
    const long long inst = p_abs / a.n_freq;
    const double f = a.freqs[p_abs - inst * a.n_freq];
    tcplx A[TL_MR][TL_MC];
    int status = 0;
#if TL_CONST
    const double w = (2 * TL_PI) * f;
    const double iw = 1.0 / w;
    if (a.n_ind > 0) {   // inductor guards of simulateAC.ts:47-51: the one-thread-per-row kernel decides
      int bad = 0;
      for (int qi = tid; qi < a.n_ind; qi += TL_THREADS) {
        const double d = w * a.ind_L[qi];
        bad |= (int)(fabs(d) < TL_EPS) | (int)(d * d < TL_EPS);
      }
      if (__syncthreads_or(bad)) {
        if (tid == 0) a.fb_list[atomicAdd(a.fb_count, 1)] = q;
        continue;
      }
    }
    __syncthreads();   // the previous point's readers of xs / Aimg are done
    // ---- my tile, straight from the constants of the topology ----
#pragma unroll
    for (int m = 0; m < TL_MR; ++m) {
      // one local row's loads in flight at a time (all MR * MC at once would take 8 registers each on top of the tile)
      asm volatile("" ::: "memory");
#pragma unroll
      for (int c = 0; c < TL_MC; ++c) {
        const size_t o = (size_t)(m * TL_MC + c) * TL_THREADS + tid;
#if TL_RC
        const double2 c0 = __ldg(a.ctab + o);
        A[m][c] = make_double2(c0.x, w * c0.y);
#else
        const double2 c0 = __ldg(a.ctab + 2 * o), c1 = __ldg(a.ctab + 2 * o + 1);
        A[m][c] = make_double2(c0.x, fma(w, c1.x, -c1.y * iw) + c0.y);
#endif
      }
    }
#else
    if (tid == 0) *s_status = 0;
    __syncthreads();   // the previous point's readers of xs / Yv / Aimg are done
    // ---- element admittances, image cleared ----
    for (int e = tid; e < a.n_elem; e += TL_THREADS) {
      const int2 mt = __ldg(a.meta + e);
      tcplx Y, J;
      const int st = tl_element(a, mt.x, mt.y, inst, f, Y, J);
      Yv[e] = Y;
      if (mt.x == 3) Jv[e - a.off_v] = J;
      else if (mt.x == 6) Jv[a.nV + (e - a.off_i)] = J;
      if (st) atomicMax(s_status, st);   // R <= 0 (3) outranks the inductor's divide guard (2): the R loop runs first
    }
    for (int o = tid; o < TL_NC * TL_LD; o += TL_THREADS) Aimg[o] = make_double2(0.0, 0.0);
    __syncthreads();
    status = *s_status;
    if (status == 0) {
      // ---- gather stamping, one thread per entry, contributions summed in the reference's stamping order ----
      for (int en = tid; en < a.n_ent; en += TL_THREADS) {
        const int rc = __ldg(a.ent_rc + en);
        const int c_end = __ldg(a.ent_ptr + en + 1);
        tcplx acc = make_double2(0.0, 0.0);
        for (int c = __ldg(a.ent_ptr + en); c < c_end; ++c) {
          const int wd = __ldg(a.contrib + c);
          const int src = (wd >> 1) & 3, idx = wd >> 3;
          tcplx v;
          if (src == 0) v = Yv[idx];
          else if (src == 1) v = Jv[idx < a.off_v_end ? idx - a.off_v : a.nV + (idx - a.off_i)];
          else v = make_double2(1.0, 0.0);
          if (wd & 1) { acc.x -= v.x; acc.y -= v.y; } else { acc.x += v.x; acc.y += v.y; }
        }
        Aimg[(size_t)(rc >> 16) * TL_LD + (rc & 0xffff)] = acc;
      }
      __syncthreads();
      // ---- my tile ----
#pragma unroll
      for (int m = 0; m < TL_MR; ++m)
#pragma unroll
        for (int c = 0; c < TL_MC; ++c) {
          const int i = m * TL_TR + tr, j = c * TL_TC + tc;
          A[m][c] = (i < TL_N && j < TL_NC) ? Aimg[(size_t)j * TL_LD + i] : make_double2(0.0, 0.0);
        }
    }
#endif

    if (status == 0) {
      // ---- elimination (solveComplex.ts:15-53) ----
      TlSeg<0, 0>::run(t, A, status);
    }

    if (status == 0) {
      // ---- U (rows in pivot order) and the eliminated right-hand side to the image ----
      __syncthreads();
#pragma unroll
      for (int m = 0; m < TL_MR; ++m)
#pragma unroll
        for (int c = 0; c < TL_MC; ++c) {
          const int i = m * TL_TR + tr, j = c * TL_TC + tc;
          if (i < TL_N && j < TL_NC && j >= i) Aimg[(size_t)j * TL_LD + i] = A[m][c];
        }
      __syncthreads();
      // ---- back-substitution (:56-71), column oriented, warp 0: lane l owns rows l, l + 32, ... ----
      if (warp == 0) {
        constexpr int RQ = (TL_N + 31) / 32;
        tcplx b[RQ];
#pragma unroll
        for (int qq = 0; qq < RQ; ++qq) {
          const int r = qq * 32 + lane;
          b[qq] = r < TL_N ? Aimg[(size_t)TL_N * TL_LD + r] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int qi = RQ - 1; qi >= 0; --qi) {
          const int ihi = (qi * 32 + 31 < TL_N - 1) ? qi * 32 + 31 : TL_N - 1;
          // column i of U for my rows, loaded one step ahead of the dependent chain
          tcplx un[RQ];
#pragma unroll
          for (int qq = 0; qq <= qi; ++qq) {
            const int r = qq * 32 + lane;
            un[qq] = r < TL_N ? Aimg[(size_t)ihi * TL_LD + r] : make_double2(0.0, 0.0);
          }
          tcplx rn = Rd[ihi];
#pragma unroll 4
          for (int i = ihi; i >= qi * 32; --i) {
            tcplx u[RQ];
#pragma unroll
            for (int qq = 0; qq <= qi; ++qq) u[qq] = un[qq];
            const tcplx ri = rn;
            if (i > qi * 32) {
#pragma unroll
              for (int qq = 0; qq <= qi; ++qq) {
                const int r = qq * 32 + lane;
                un[qq] = r < TL_N ? Aimg[(size_t)(i - 1) * TL_LD + r] : make_double2(0.0, 0.0);
              }
              rn = Rd[i - 1];
            }
            tcplx xi = tl_mul(b[qi], ri);
            xi.x = __shfl_sync(TL_FULL, xi.x, i & 31);
            xi.y = __shfl_sync(TL_FULL, xi.y, i & 31);
            if (lane == (i & 31)) xs[i] = xi;
#pragma unroll
            for (int qq = 0; qq <= qi; ++qq) {
              const int r = qq * 32 + lane;
              if (r < i) b[qq] = tl_submul(b[qq], u[qq], xi);
            }
          }
        }
        if (lane == 0) xs[TL_N] = make_double2(0.0, 0.0);
      }
      __syncthreads();
    }

    // ---- unpack (simulateAC.ts:85-126) ----
    const long long sld = a.series_ld;
    const long long xst = sld ? sld : 1;
    tcplx* xo = sld ? a.x + q : a.x + (size_t)q * TL_N;
    tcplx* io = a.ielem ? (sld ? a.ielem + q : a.ielem + (size_t)q * a.n_ac_elem) : nullptr;
    if (status == 0) {
      for (int e = tid; e < TL_N; e += TL_THREADS) xo[e * xst] = xs[e];
#if TL_IELEM
      if (io)
        for (int e = tid; e < a.n_ac_elem; e += TL_THREADS) {
#if TL_CONST
          const double2 lo = __ldg((const double2*)(a.el_rec + e)), hi = __ldg((const double2*)(a.el_rec + e) + 1);
          const long long ij = __double_as_longlong(lo.x);
          const tcplx v1 = xs[(int)(ij & 0xffffffffll)], v2 = xs[(int)(ij >> 32)];
          const tcplx Y = make_double2(lo.y, fma(w, hi.x, -hi.y * iw));
          io[e * xst] = tl_mul(Y, make_double2(v1.x - v2.x, v1.y - v2.y));
#else
          tcplx cur;
          if (e >= a.off_v) {
            cur = xs[a.nn + (e - a.off_v)];
          } else {
            const int4 en = __ldg(a.ends + e);
            const tcplx v1 = xs[en.x == 0 ? TL_N : en.x - 1], v2 = xs[en.y == 0 ? TL_N : en.y - 1];
            cur = tl_mul(Yv[e], make_double2(v1.x - v2.x, v1.y - v2.y));
          }
          io[e * xst] = cur;
#endif
        }
#endif
    } else {
      const double qn = __longlong_as_double(0x7ff8000000000000ll);
      const tcplx nanv = make_double2(qn, qn);
      for (int e = tid; e < TL_N; e += TL_THREADS) xo[e * xst] = nanv;
      if (io)
        for (int e = tid; e < a.n_ac_elem; e += TL_THREADS) io[e * xst] = nanv;
    }
    if (tid == 0) a.status[q] = status;
  }
}
