// Dense batched complex LU with partial pivoting, ONE WARP PER SYSTEM, for Nvar <= 32 (tier 9, plain frequency sweeps).
//
// North star piece (2): "one warp per system for n <= 32 ... __shfl_sync pivot-search reductions".  Replaces, like
// tile_kernel.cuh, the per-frequency body of simulateAC (lib/analysis/simulateAC.ts:80-127): buildLinearSystemForAC
// :24-60 (every entry from per-topology constants, value = alpha + j (w beta - gamma / w + Im J)), solveComplex
// (lib/math/solveComplex.ts:15-53 elimination, :56-71 back-substitution), unpack :85-126.  Compiled by NVRTC per Nvar
// with the WL_* macros in front (spicey_native.cu: warp_lu_source).
//
// Mapping.  Lane i holds row i of [A b] — Nvar + 1 complex values in registers — for the whole solve.  Rows never move
// (implicit pivoting): every lane keeps whether its row has been a pivot (`done`) and the logical position the
// reference's row swaps (:30-34) would have given it (`pos`), which reproduces the first-maximum tie rule of :18-28
// exactly as lu_rowthread.cuh does.  The loop over the pivot steps is fully unrolled, so every column index is a
// constant and a step touches live columns only.  Per step: |a_ik|^2 per lane, one redux.sync.max over the high words of
// the IEEE bit patterns (ties: low words, then the lowest logical position), the pivot lane publishes 1 / a_pk and its
// row through a double-buffered shared-memory record (one __syncwarp per step, no CTA barrier on the data path), every
// lane forms its multiplier (zeroed when |f| < EPS: :46) and updates its row with one broadcast LDS.128 per complex FMA.
// The generic tile kernel spends ~530 warp instructions per step of a 32-unknown system on barriers, row exchange and
// multiplier traffic; this one ~75 + 6 per live column.  The warps of a CTA meet at a barrier every WL_SYNC steps only to
// stay on the same instruction-cache lines (the unrolled body is ~80 KB).
//
// Statuses: a singular pivot (|a| < EPS) or a pivot that trips Complex.div's guard (|a|^2 < EPS) marks the system
// (NaN results, status 1 / 2) and the warp runs on; inductor guards (simulateAC.ts:47-51) send the point to the
// one-thread-per-row kernel through the fallback list.  Arithmetic: the default policy of common.cuh (FMA contraction,
// |a|^2 metric, reciprocal multiply; parity 1e-9).

typedef double2 wcplx;

#define WL_NC (WL_N + 1)
#define WL_THREADS (WL_WARPS * 32)
#define WL_EPS 1e-15
#define WL_PI 3.141592653589793
#define WL_FULL 0xffffffffu
#define WL_REC (WL_NC + 2)             /* record of a step: 1 / a_pk | (|a_pk|^2, logical position of the pivot row) | the pivot row */
#ifndef WL_SYNC
#define WL_SYNC 4
#endif
#ifndef WL_CONST
#define WL_CONST 1                     /* 1: entries from per-topology constants (plain sweep); 0: per-instance values, stamped from the element table */
#endif

struct TileArgs {   // must match TileArgs in spicey_native.cu (shared with tile_kernel.cuh)
  const double* freqs; long long n_freq, p_begin, p_count;
  double2* x; double2* ielem; int* status; long long series_ld;
  const int4* ends; const int2* meta; const double* values; const int* var_of_slot; const double* var_values; long long n_inst;
  const int* ent_rc; const int* ent_ptr; const int* contrib;
  const double2* ctab;     // here: [WL_NC columns][32 lanes]; WL_RC: (alpha, beta), else (alpha + Re J, Im J), (beta, gamma)
  const double4* el_rec;   // [n_ac_elem] {bits of (i1, i2), ya, yb, yg}; index Nvar = the ground node
  const double* ind_L;
  long long* fb_list; int* fb_count;
  const long long* plist; const int* pcount;   // optional: solve the launch-local points plist[0 .. *pcount) only
  unsigned long long* fb_total;
  int n_ind, n_ent, nn, nV, n_elem, n_ac_elem, off_v, off_v_end, off_i;
};

__device__ __forceinline__ wcplx wl_mul(wcplx a, wcplx b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ wcplx wl_submul(wcplx a, wcplx f, wcplx p) {
  return make_double2(fma(-f.x, p.x, fma(f.y, p.y, a.x)), fma(-f.x, p.y, fma(-f.y, p.x, a.y)));
}
#if !WL_CONST
__device__ __forceinline__ double wl_value(const TileArgs& a, int slot, long long inst) {
  const int v = __ldg(a.var_of_slot + slot);
  return v < 0 ? __ldg(a.values + slot) : __ldg(a.var_values + (long long)v * a.n_inst + inst);
}
// Element admittance at frequency f (simulateAC.ts:36-52) and source phasor (:54-57); returns the status.
__device__ __forceinline__ int wl_element(const TileArgs& a, int type, int vidx, long long inst, double f, wcplx& Y, wcplx& J) {
  const double twoPi = 2 * WL_PI;
  Y = make_double2(0.0, 0.0);
  J = make_double2(0.0, 0.0);
  if (type == 0) {          // R
    const double R = wl_value(a, vidx, inst);
    if (R <= 0) return 3;   // :37
    Y.x = 1 / R;
  } else if (type == 1) {   // C: twoPi * f * c.C  :43
    Y.y = __dmul_rn(__dmul_rn(twoPi, f), wl_value(a, vidx, inst));
  } else if (type == 2) {   // L
    const double d = __dmul_rn(__dmul_rn(twoPi, f), wl_value(a, vidx, inst));
    if (fabs(d) < WL_EPS) return 0;     // denom.abs() < EPS -> Y = 0  :49
    const double dd = __dmul_rn(d, d);
    if (dd < WL_EPS) return 2;          // Complex.div guard (Complex.ts:41-42)
    Y.x = 0.0 / dd;
    Y.y = (0.0 - d) / dd;
  } else if (type == 3 || type == 6) {  // V, I: phasor fromPolar(acMag, acPhaseDeg)  Complex.ts:16-19
    const double mag = wl_value(a, vidx + 1, inst), deg = wl_value(a, vidx + 2, inst);
    const double ph = (deg * WL_PI) / 180;
    double sn, cs;
    sincos(ph, &sn, &cs);
    J.x = mag * cs;
    J.y = mag * sn;
  }
  return 0;
}
#endif

__device__ __forceinline__ double wl_rcp(double a) {   // MUFU seed + two Newton steps (<= 1 ulp), no slow path
  double y, e;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  return y;
}

extern __shared__ __align__(16) unsigned char wl_smem[];

extern "C" __global__ void __launch_bounds__(WL_THREADS, WL_MINB) spicey_warp_lu_jit(const TileArgs a) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per warp: two step records | x (Nvar + 1 entries, the last one the ground node's zero)
  //           [| element admittances | source phasors | the stamped image [column][lane = row]: per-instance values only]
#if WL_CONST
  wcplx* rec = (wcplx*)wl_smem + (size_t)wib * (2 * WL_REC + WL_NC);
  wcplx* xs = rec + 2 * WL_REC;
#else
  const int n_src = a.nV + (a.n_elem - a.off_i);
  wcplx* rec = (wcplx*)wl_smem + (size_t)wib * (2 * WL_REC + WL_NC + a.n_elem + (n_src > 0 ? n_src : 1) + WL_NC * 32);
  wcplx* xs = rec + 2 * WL_REC;
  wcplx* Yv = xs + WL_NC;
  wcplx* Jv = Yv + a.n_elem;
  wcplx* img = Jv + (n_src > 0 ? n_src : 1);
#endif
  const long long n_warps = (long long)gridDim.x * WL_WARPS;
  const bool row = lane < WL_N;

  // CTA-uniform trip count (the warps of a CTA meet at barriers inside): warps past the end solve the last point again
  // and store nothing
  const long long work = a.plist ? (long long)*a.pcount : a.p_count;
  if (a.plist && a.fb_total && blockIdx.x == 0 && threadIdx.x == 0 && work > 0) {
    atomicAdd(a.fb_total, (unsigned long long)work);
    atomicAdd(a.fb_total + 1, (unsigned long long)work);
  }
  for (long long cta_base = (long long)blockIdx.x * WL_WARPS; cta_base < work; cta_base += n_warps) {
    const long long q0 = cta_base + wib;
    bool valid = q0 < work;
    const long long qi = valid ? q0 : work - 1;
    const long long q = a.plist ? a.plist[qi] : qi;
#if WL_CONST
    const double w = (2 * WL_PI) * a.freqs[a.p_begin + q];
    const double iw = 1.0 / w;
#else
    const long long p_abs = a.p_begin + q;
    const long long inst = p_abs / a.n_freq;
    const double fq = a.freqs[p_abs - inst * a.n_freq];
#endif
    if (WL_CONST && a.n_ind > 0) {   // inductor guards of simulateAC.ts:47-51: the one-thread-per-row kernel decides
      int bad = 0;
      for (int li = lane; li < a.n_ind; li += 32) {
#if WL_CONST
        const double d = w * a.ind_L[li];
#else
        const double d = 1.0;
#endif
        bad |= (int)(fabs(d) < WL_EPS) | (int)(d * d < WL_EPS);
      }
      if (__any_sync(WL_FULL, bad)) {
        if (valid && lane == 0) a.fb_list[atomicAdd(a.fb_count, 1)] = q;
        valid = false;
      }
    }
    // ---- my row, straight from the constants of the topology (simulateAC.ts:24-60) ----
    wcplx A[WL_NC];
    int st = 0;
#if WL_CONST
#pragma unroll
    for (int j = 0; j < WL_NC; ++j) {
      const size_t o = (size_t)j * 32 + lane;
#if WL_RC
      const double2 c0 = __ldg(a.ctab + o);
      A[j] = make_double2(c0.x, w * c0.y);
#else
      const double2 c0 = __ldg(a.ctab + 2 * o), c1 = __ldg(a.ctab + 2 * o + 1);
      A[j] = make_double2(c0.x, fma(w, c1.x, -c1.y * iw) + c0.y);
#endif
    }
#else
    // element admittances of this instance, then gather stamping with one lane per matrix entry (the contributions of
    // an entry summed in the reference's stamping order), through the warp's image in shared memory
    __syncwarp();   // the previous point's readers of Yv / xs are done
    {
      int st_el = 0;
      for (int e = lane; e < a.n_elem; e += 32) {
        const int2 mt = __ldg(a.meta + e);
        wcplx Y, J;
        const int s1 = wl_element(a, mt.x, mt.y, inst, fq, Y, J);
        Yv[e] = Y;
        if (mt.x == 3) Jv[e - a.off_v] = J;
        else if (mt.x == 6) Jv[a.nV + (e - a.off_i)] = J;
        st_el = s1 > st_el ? s1 : st_el;
      }
      st = __reduce_max_sync(WL_FULL, st_el);   // R <= 0 (3) outranks the inductor's divide guard (2)
#pragma unroll
      for (int j = 0; j < WL_NC; ++j) img[j * 32 + lane] = make_double2(0.0, 0.0);
      __syncwarp();
      for (int en = lane; en < a.n_ent; en += 32) {
        const int rc = __ldg(a.ent_rc + en);
        const int c_end = __ldg(a.ent_ptr + en + 1);
        wcplx acc = make_double2(0.0, 0.0);
        for (int c = __ldg(a.ent_ptr + en); c < c_end; ++c) {
          const int wd = __ldg(a.contrib + c);
          const int src = (wd >> 1) & 3, idx = wd >> 3;
          wcplx v;
          if (src == 0) v = Yv[idx];
          else if (src == 1) v = Jv[idx < a.off_v_end ? idx - a.off_v : a.nV + (idx - a.off_i)];
          else v = make_double2(1.0, 0.0);
          if (wd & 1) { acc.x -= v.x; acc.y -= v.y; } else { acc.x += v.x; acc.y += v.y; }
        }
        img[(rc >> 16) * 32 + (rc & 0xffff)] = acc;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < WL_NC; ++j) A[j] = img[j * 32 + lane];
    }
#endif
    bool done = !row;
    int pos = lane;
    wcplx rdiag = make_double2(0.0, 0.0);

    // ---- elimination (solveComplex.ts:15-53), fully unrolled ----
#pragma unroll
    for (int k = 0; k < WL_N; ++k) {
      if (WL_SYNC > 0 && k % WL_SYNC == 0) __syncthreads();   // instruction-cache locality only
      wcplx* R = rec + (k & 1) * WL_REC;
      const wcplx aik = A[k];
      const double m = fma(aik.x, aik.x, aik.y * aik.y);
      const bool cand = !done;
      unsigned long long key = (unsigned long long)__double_as_longlong(m);
      if (m != m) key = (pos == k) ? ~0ull : 0ull;   // NaN only wins in place (JS: v > vmax is false)
      if (!cand) key = 0ull;
      const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
      const unsigned mh = __reduce_max_sync(WL_FULL, hi);
      unsigned top = __ballot_sync(WL_FULL, cand && hi == mh);
      if (__popc(top) != 1) {   // equal high words: the low words, then the first in the reference's scan order
        const bool t1 = cand && hi == mh;
        const unsigned ml = __reduce_max_sync(WL_FULL, t1 ? lo : 0u);
        const bool t2 = t1 && lo == ml;
        const int mp = __reduce_min_sync(WL_FULL, t2 ? pos : 0x7fffffff);
        top = __ballot_sync(WL_FULL, t2 && pos == mp);
      }
      const int pl = __ffs(top) - 1;
      if (lane == pl) {   // the pivot row: 1 / a_pk, its metric and position, its live entries
        const double inv = wl_rcp(m);
        R[0] = make_double2(aik.x * inv, -aik.y * inv);
        R[1] = make_double2(m, __longlong_as_double((long long)pos));
#pragma unroll
        for (int j = k + 1; j < WL_NC; ++j) R[2 + j] = A[j];
      }
      __syncwarp();
      const wcplx r = R[0];
      const wcplx mp_ = R[1];
      const double vmax = mp_.x;
      const int bpos = (int)__double_as_longlong(mp_.y);
      if (st == 0) {
        if (vmax < WL_EPS * WL_EPS) st = 1;        // singular (:29); the metric is |a|^2
        else if (vmax < WL_EPS) st = 2;            // Complex.div by this pivot throws (Complex.ts:41-42)
      }
      wcplx f = make_double2(0.0, 0.0);
      if (lane == pl) {
        done = true;
        pos = k;
        rdiag = r;
      } else if (!done) {
        if (pos == k) pos = bpos;                    // the reference's row swap (:30-34)
        f = wl_mul(aik, r);                          // :45
        if (fma(f.x, f.x, f.y * f.y) < WL_EPS * WL_EPS) f = make_double2(0.0, 0.0);   // :46
      }
      // the update a_ij -= f_i * u_kj (:47-52); rows that are done keep f = 0
#pragma unroll
      for (int j = k + 1; j < WL_NC; ++j) A[j] = wl_submul(A[j], f, R[2 + j]);
    }

    // ---- back-substitution (:56-71), column oriented: the row at logical position i yields x_i ----
    wcplx b = A[WL_N];
    wcplx myx = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = WL_N - 1; i >= 0; --i) {
      const int ol = __ffs(__ballot_sync(WL_FULL, row && pos == i)) - 1;
      wcplx xi = wl_mul(b, rdiag);
      xi.x = __shfl_sync(WL_FULL, xi.x, ol);
      xi.y = __shfl_sync(WL_FULL, xi.y, ol);
      if (lane == i) myx = xi;
      if (row && pos < i) b = wl_submul(b, A[i], xi);
    }

    // ---- unpack (simulateAC.ts:85-126) ----
    __syncwarp();   // the previous point's readers of xs are done
    if (lane <= WL_N) xs[lane] = row ? myx : make_double2(0.0, 0.0);
#if WL_N == 32
    if (lane == 0) xs[WL_N] = make_double2(0.0, 0.0);
#endif
    __syncwarp();
    if (valid) {
      const long long sld = a.series_ld;
      const long long xst = sld ? sld : 1;
      wcplx* xo = sld ? a.x + q : a.x + (size_t)q * WL_N;
      const double qn = __longlong_as_double(0x7ff8000000000000ll);
      const wcplx nanv = make_double2(qn, qn);
      if (row) xo[lane * xst] = st == 0 ? myx : nanv;
#if WL_IELEM
      if (a.ielem) {
        wcplx* io = sld ? a.ielem + q : a.ielem + (size_t)q * a.n_ac_elem;
        for (int e = lane; e < a.n_ac_elem; e += 32) {
#if WL_CONST
          const double2 lo2 = __ldg((const double2*)(a.el_rec + e)), hi2 = __ldg((const double2*)(a.el_rec + e) + 1);
          const long long ij = __double_as_longlong(lo2.x);
          const wcplx v1 = xs[(int)(ij & 0xffffffffll)], v2 = xs[(int)(ij >> 32)];
          const wcplx Y = make_double2(lo2.y, fma(w, hi2.x, -hi2.y * iw));
          io[e * xst] = st == 0 ? wl_mul(Y, make_double2(v1.x - v2.x, v1.y - v2.y)) : nanv;
#else
          wcplx cur;
          if (e >= a.off_v) {
            cur = xs[a.nn + (e - a.off_v)];
          } else {
            const int4 en = __ldg(a.ends + e);
            const wcplx v1 = xs[en.x == 0 ? WL_N : en.x - 1], v2 = xs[en.y == 0 ? WL_N : en.y - 1];
            cur = wl_mul(Yv[e], make_double2(v1.x - v2.x, v1.y - v2.y));
          }
          io[e * xst] = st == 0 ? cur : nanv;
#endif
        }
      }
#endif
      if (lane == 0) a.status[q] = st;
    }
  }
}
