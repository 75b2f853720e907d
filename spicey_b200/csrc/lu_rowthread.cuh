// Batched dense LU with partial pivoting, one CTA per system, one THREAD per matrix row.
//
// Replaces lib/math/solveComplex.ts:4-73 and lib/math/solveReal.ts:3-73 (Gaussian
// elimination on the augmented matrix [A b], first-max pivot rule, |f| < EPS multiplier
// skip, back-substitution) for one system held in shared memory (or an L2-resident
// global scratch for systems too large for one SM).
//
// B200 mapping
//  * The matrix is column-major, A(r, j) at A[j*ldr + r]: the 32 threads of a warp own
//    32 consecutive rows, so every access of "my row, column j" is a contiguous 512-byte
//    (complex) shared-memory request without bank conflicts, and every access to the
//    pivot row is a same-address broadcast.
//  * Pivoting is implicit.  A row never moves; each thread keeps in registers whether
//    its row was already chosen (done) and the logical position the reference's
//    row-swapping would have given it (pos), which reproduces the reference's
//    "first maximum wins" tie rule (solveComplex.ts:20-28,30-34) without moving data.
//  * The pivot search is a redux.sync warp arg-max on the IEEE bit pattern of the pivot
//    metric followed by ONE __syncthreads per elimination step (double-buffered
//    per-warp partials).  Every thread speculatively computes the reciprocal of its own
//    candidate while the reduction is in flight; the winner's reciprocal travels with
//    the partial, so no thread waits for a divide after the barrier.
//  * After the barrier each remaining row updates itself: nothing else synchronises,
//    because a thread only ever writes its own row and only reads rows that are done.
//  * The reference skips a row whose multiplier is ~0 (solveComplex.ts:46).  The kernel
//    extends that to columns: every row carries the bit mask of its structural
//    non-zeros (updated by OR-ing in the pivot row's mask), and the row update walks
//    only the set bits of the pivot row's mask.  Structural zeros stay exact zeros, so
//    results equal the dense computation.
#pragma once
#include "common.cuh"

namespace spicey {

struct PivotPartial {      // 32 bytes, one per warp per buffer
  unsigned long long key;  // bit pattern of the metric (0 = no candidate)
  int pos;                 // logical position of the candidate row
  int row;                 // physical row (= owning thread)
  double2 aux;             // reciprocal of the candidate (fast mode)
};

template <typename T> __device__ __forceinline__ double2 to_aux(T v);
template <> __device__ __forceinline__ double2 to_aux<double>(double v) { return make_double2(v, 0.0); }
template <> __device__ __forceinline__ double2 to_aux<cplx>(cplx v) { return v; }
template <typename T> __device__ __forceinline__ T from_aux(double2 v);
template <> __device__ __forceinline__ double from_aux<double>(double2 v) { return v.x; }
template <> __device__ __forceinline__ cplx from_aux<cplx>(double2 v) { return v; }

// Solves the n x n system stored augmented (column n = rhs) in A; writes x to xs[0..n).
// All threads of the CTA must call it (threads >= n only take part in the reductions).
// mask: [n][MW] structural masks, already initialised by the caller for this system.
// Returns the status (uniform across the CTA).
template <typename T, bool STRICT>
__device__ int lu_solve_rowthread(T* __restrict__ A, int ldr, int n, unsigned* __restrict__ mask,
                                  int MW, T* __restrict__ xs, PivotPartial* __restrict__ red) {
  typedef Num<T> N;
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  const unsigned full = 0xffffffffu;

  bool done = (t >= n);
  int pos = t;
  T rdiag = N::zero();
  int status = ST_OK;
  const double thr = N::template thresh<STRICT>();

  for (int k = 0; k < n; ++k) {
    // ---- pivot candidate of my row (solveComplex.ts:18-28) ----
    T a = N::zero();
    unsigned long long key = 0ull;
    int cpos = 0x7fffffff;
    T rc = N::zero();
    if (!done) {
      a = A[(size_t)k * ldr + t];
      double m = N::template metric<STRICT>(a);
      key = (unsigned long long)__double_as_longlong(m);
      if (m != m) key = (pos == k) ? ~0ull : 0ull;  // NaN only wins in place (JS: v > vmax is false)
      cpos = pos;
      if (!STRICT) rc = N::recip(a);  // speculative: overlaps the reduction
    }
    unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    unsigned mh = __reduce_max_sync(full, hi);
    unsigned ml = __reduce_max_sync(full, hi == mh ? lo : 0u);
    bool top = (hi == mh) && (lo == ml);
    int mp = __reduce_min_sync(full, top ? cpos : 0x7fffffff);
    unsigned winners = __ballot_sync(full, top && cpos == mp);
    PivotPartial* buf = red + (k & 1) * nwarps;
    if (lane == __ffs(winners) - 1) {
      PivotPartial pp;
      pp.key = key; pp.pos = cpos; pp.row = t; pp.aux = to_aux<T>(rc);
      buf[warp] = pp;
    }
    __syncthreads();
    unsigned long long bkey = 0ull;
    int bpos = 0x7fffffff, bw = 0;
    for (int w = 0; w < nwarps; ++w) {
      unsigned long long kw = buf[w].key;
      int pw = buf[w].pos;
      if (kw > bkey || (kw == bkey && pw < bpos)) { bkey = kw; bpos = pw; bw = w; }
    }
    const double vmax = __longlong_as_double((long long)bkey);
    if (vmax < thr) { status = ST_SINGULAR; break; }            // solveComplex.ts:29
    const int p = buf[bw].row;
    // Complex.ts:41-42: every division by this pivot throws when re^2+im^2 < EPS.
    if (N::template div_guard<STRICT>(vmax, A[(size_t)k * ldr + p])) { status = ST_CDIV; break; }
    if (t == p) {
      done = true;
      pos = k;
      rdiag = STRICT ? a : from_aux<T>(buf[bw].aux);
      continue;
    }
    if (done) continue;
    if (pos == k) pos = bpos;                                   // the reference's row swap (:30-34)
    T f;
    if (STRICT) f = N::div_strict(a, A[(size_t)k * ldr + p]);   // :45
    else f = N::mul(a, from_aux<T>(buf[bw].aux));
    if (N::template metric<STRICT>(f) < thr) continue;          // :46
    const unsigned* pm = mask + (size_t)p * MW;
    unsigned* om = mask + (size_t)t * MW;
    for (int w = (k + 1) >> 5; w < MW; ++w) {
      unsigned pbits = pm[w];
      unsigned own = om[w];
      if ((own | pbits) != own) om[w] = own | pbits;
      if (w == ((k + 1) >> 5)) pbits &= (0xffffffffu << ((k + 1) & 31));
      while (pbits) {
        int j = (w << 5) + __ffs(pbits) - 1;
        pbits &= pbits - 1;
        size_t o = (size_t)j * ldr;
        A[o + t] = N::template submul<STRICT>(A[o + t], f, A[o + p]);  // :47-52
      }
    }
  }
  // A break above is uniform (every thread sees the same partials).
  if (status != ST_OK) return status;

  // ---- back-substitution (solveComplex.ts:56-71), column oriented ----
  T b = N::zero();
  if (t < n) b = A[(size_t)n * ldr + t];
  if (!STRICT) {
    for (int i = n - 1; i >= 0; --i) {
      if (pos == i && t < n) xs[i] = N::mul(b, rdiag);
      __syncthreads();
      if (t < n && pos < i && ((mask[(size_t)t * MW + (i >> 5)] >> (i & 31)) & 1u))
        b = N::template submul<false>(b, A[(size_t)i * ldr + t], xs[i]);
    }
  } else {
    // reference order: s = b_i - sum_{j>i} U_ij x_j with j ascending, then s / U_ii
    for (int i = n - 1; i >= 0; --i) {
      if (pos == i && t < n) {
        T s = b;
        for (int j = i + 1; j < n; ++j)
          if ((mask[(size_t)t * MW + (j >> 5)] >> (j & 31)) & 1u)
            s = N::template submul<true>(s, A[(size_t)j * ldr + t], xs[j]);
        xs[i] = N::div_strict(s, rdiag);
      }
      __syncthreads();
    }
  }
  __syncthreads();
  return ST_OK;
}

}  // namespace spicey
