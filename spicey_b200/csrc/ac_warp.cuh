// Sparse AC kernel, one WARP per frequency point: executes the level-scheduled program of warp_program.h.
//
// Replaces the same reference code as ac_sparse.cuh (simulateAC.ts:24-60 stamp, solveComplex.ts:4-73 solve,
// simulateAC.ts:85-126 unpack) for systems whose factorisation is too large for one thread's share of the chip
// (cfg 4: Nvar = 257, 4,000 U entries).  Same contract: the pilot's pivot sequence is verified on every
// system's own numbers with the reference's rule (first maximum wins); a system that disagrees, or that trips a
// guard, goes to the dense pivoting kernel in the same stream, which also reports the exact status.
//
// B200 mapping
//  * the elimination's working set (the rows a pivot step touches, ~300 values for the mesh) is a per-warp pool
//    in shared memory; the 32 lanes execute 32 independent updates a_ij <- a_ij - f_i * a_kj at a time;
//  * the program is the same for every system, so the warps of a CTA run it in lock step and the CTA stages it
//    through shared memory one record ahead with cp.async (double buffer): no lane ever waits on L2 for an
//    operation word, and the program is fetched once per CTA instead of once per warp;
//  * what only the back-substitution reads (U, 1/u_kk, the eliminated right-hand side) is written once to a
//    per-warp global workspace whose slots are numbered in the order the back-substitution reads them, so each
//    group of columns reads one contiguous range, prefetched into L1 a group ahead;
//  * the back-substitution is column oriented — x_j = acc_j / u_jj by all lanes, then one lane per row of
//    column j updates its accumulator in shared memory — so the dependent chain per unknown is one shared-memory
//    round trip and two complex operations, with no reduction across lanes;
//  * results leave from the accumulator array: lanes = unknowns / elements.
#pragma once
#include "ac_kernels.cuh"
#include "warp_program.h"

namespace spicey {

struct WarpArgs {
  const int4* stream;     // packed records (warp_program.h)
  const int2* fwd_tab;    // [n] (offset, length) in 16-byte units
  const int2* back_tab;   // [n_groups]
  const int* rhs_init;    // [n]
  const double2 *ent_c0, *ent_c1;   // stamped entries: (alpha + Re J, Im J), (beta, gamma)
  const double *el_a, *el_b, *el_g; // element admittance constants
  const int4* el_ends;              // element table rows: (n1, n2, .., ..) node ids, 0 = ground
  const double* ind_L;
  int n_ind;
  int n, nn, n_ac_elem, v_first, n_pool, n_gslots, max_elim, n_groups, max_rec16, g_first0, g_count0;
  const double* freqs;
  long long p_count;
  double2* G;             // [resident warps][n_gslots]
  double2* x;
  double2* ielem;
  int* status;
  long long series_ld;
  long long* fb_list;
  int* fb_count;
};

__device__ __forceinline__ double fast_rcp(double a) {  // MUFU seed + two Newton steps (<= 1 ulp)
  double y, e;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  return y;
}

// CTA-wide asynchronous copy of one record into a staging buffer (L2 -> shared memory, no registers).
__device__ __forceinline__ void stage_record(int4* dst, const int4* src, int len16, int tid, int nthreads) {
  for (int i = tid; i < len16; i += nthreads) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ac_warp_kernel(WarpArgs a) {
  typedef Num<cplx> N;
  extern __shared__ __align__(16) double2 wsm[];   // [2 staging buffers][per-warp: pool (later: accumulators) | multipliers]
  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  // The accumulators of the back-substitution take over the pool's memory: the pool is dead once the
  // elimination is done (what the back-substitution reads comes from the global workspace).
  const int work = max(a.n_pool, a.n), per_warp = work + a.max_elim;
  // all shared-memory pointers are offsets from the __shared__ symbol, so that every access is an LDS/STS
  double2* pool = wsm + 2 * (size_t)a.max_rec16 + (size_t)wib * per_warp;
  double2* acc = pool;
  double2* Fm = pool + work;
  const long long gw = (long long)blockIdx.x * WARPS + wib;
  double2* G = a.G + gw * a.n_gslots;
  const double thr = kEps * kEps;
  const int n = a.n;

  // CTA-uniform trip count: the warps past the end solve the last point again and store nothing
  for (long long base = (long long)blockIdx.x * WARPS; base < a.p_count; base += (long long)gridDim.x * WARPS) {
    const bool valid = base + wib < a.p_count;
    const long long p = valid ? base + wib : a.p_count - 1;
    const double w = (2 * kPi) * a.freqs[p];
    const double iw = 1.0 / w;
    auto pristine = [&](int en) -> cplx {
      const double2 c0 = __ldg(a.ent_c0 + en), c1 = __ldg(a.ent_c1 + en);
      return make_double2(c0.x, fma(w, c1.x, -c1.y * iw) + c0.y);
    };
    bool ok = true;
    for (int k = 0; k < a.n_ind; ++k) {  // inductor guards of simulateAC.ts:47-51: the dense kernel decides
      const double d = w * a.ind_L[k];
      ok = ok && !(fabs(d) < kEps || d * d < kEps);
    }
    if (lane == 0) pool[0] = make_double2(0.0, 0.0);
    // ---- elimination: record s in staging buffer s & 1, record s + 1 on its way ----
    __syncthreads();   // the previous point's last group is done with both buffers
    {
      const int2 t0 = __ldg(a.fwd_tab);
      stage_record((int4*)wsm, a.stream + t0.x, t0.y, tid, WARPS * 32);
      stage_wait();
    }
    int2 tnext = n > 1 ? __ldg(a.fwd_tab + 1) : make_int2(0, 0);
    for (int s = 0; s < n; ++s) {
      __syncthreads();   // record s is complete and visible; everyone is done with record s - 1
      if (s + 1 < n) stage_record((int4*)wsm + ((s + 1) & 1) * a.max_rec16, a.stream + tnext.x, tnext.y, tid, WARPS * 32);
      if (s + 2 < n) tnext = __ldg(a.fwd_tab + s + 2);
      const int* rec = (const int*)(wsm + (s & 1) * a.max_rec16);
      const int n_cand = rec[0], pidx = rec[1], rcp_g = rec[2], n_elim = rec[3], n_cols = rec[4], n_stamp = rec[5];
      const int* stamp = rec + 8;
      const int* cand = stamp + 12 * n_stamp;
      const int* elim = cand + n_cand;
      const int* src = elim + n_elim;
      const int* ops = rec + ((8 + 12 * n_stamp + n_cand + n_elim + n_cols + 3) & ~3);
      const int* opg = ops + ((n_elim * n_cols + 3) & ~3);
      // stamped entries this step reads (simulateAC.ts:24-60): alpha + j(w*beta - gamma/w), constants in the record
      for (int q = lane; q < n_stamp; q += 32) {
        const double2 c0 = *(const double2*)(stamp + 12 * q + 4), c1 = *(const double2*)(stamp + 12 * q + 8);
        pool[stamp[12 * q]] = make_double2(c0.x, fma(w, c1.x, -c1.y * iw) + c0.y);
      }
      __syncwarp();
      const cplx ap = pool[cand[pidx]];
      const double mp = fma(ap.x, ap.x, ap.y * ap.y);
      for (int c = lane; c < n_cand; c += 32) {  // solveComplex.ts:18-28: first maximum wins
        if (c == pidx) continue;
        const cplx z = pool[cand[c]];
        const double m = fma(z.x, z.x, z.y * z.y);
        ok = ok && (c < pidx ? m < mp : !(m > mp));
      }
      ok = ok && (mp >= kEps);   // singular / Complex.div guard (or NaN)
      const double inv = fast_rcp(mp);
      const cplx r = make_double2(ap.x * inv, -ap.y * inv);
      if (lane == 0 && rcp_g >= 0) G[rcp_g] = r;
      for (int e = lane; e < n_elim; e += 32) {
        cplx fm = N::mul(pool[elim[e]], r);
        if (fma(fm.x, fm.x, fm.y * fm.y) < thr) fm = make_double2(0.0, 0.0);   // solveComplex.ts:46
        Fm[e] = fm;
      }
      __syncwarp();
      // Updates: a dense (rows to eliminate) x (columns of the pivot row) block.  Lanes = columns, so the pivot-row
      // entry of a lane's column is loaded once per pass and the multiplier of a row is one broadcast; the loop
      // runs over the rows, each operation word = pool byte offsets (old | flag) | dst << 16.  The loads of
      // row e + 1 are issued before row e is computed and stored (no operation of a step reads what another one
      // writes, and a slot is never re-written before its last reader in this execution order).
      for (int c0 = 0; c0 < n_cols; c0 += 32) {
        const int cl = c0 + lane;
        const bool act = cl < n_cols;
        const cplx sv = pool[act ? src[cl] : 0];
        const int* opc = ops + (act ? cl : 0);
        const int* gpc = opg + (act ? cl : 0);
        const char* pb = (const char*)pool;
        const unsigned idle = 0xfff00000u;   // zero slot, no destination
        unsigned opn = (act && n_elim > 0) ? (unsigned)opc[0] : idle;
        cplx on = *(const cplx*)(pb + (opn & 0xfff0u));
        for (int e = 0; e < n_elim; ++e) {
          const unsigned op = opn;
          const cplx o = on;
          const cplx f = Fm[e];
          if (e + 1 < n_elim) {
            opn = act ? (unsigned)opc[(e + 1) * n_cols] : idle;
            on = *(const cplx*)(pb + (opn & 0xfff0u));
          }
          const cplx v = N::submul<false>(o, f, sv);
          __syncwarp();   // every operand of the row has been read before any slot is overwritten
          const unsigned dst = op >> 16;
          if (dst != 0xfff0u) *(cplx*)((char*)pool + dst) = v;
          if (op & 1u) G[gpc[e * n_cols]] = v;
        }
      }
      stage_wait();
    }
    // ---- back-substitution, column oriented; group records staged the same way ----
    __syncthreads();
    {
      const int2 t0 = __ldg(a.back_tab);
      stage_record((int4*)wsm, a.stream + t0.x, t0.y, tid, WARPS * 32);
    }
    auto fb = [&](int e) -> cplx {
      if (e == kWarpZero) return make_double2(0.0, 0.0);
      return e >= 0 ? G[e] : pristine(~e);
    };
    // the first group's U range: into L1 while the accumulators are initialised
    for (int q = lane * 8; q < a.g_count0; q += 256) asm volatile("prefetch.global.L1 [%0];" ::"l"(G + a.g_first0 + q));
    for (int i = lane; i < n; i += 32) acc[i] = fb(__ldg(a.rhs_init + i));
    stage_wait();
    tnext = a.n_groups > 1 ? __ldg(a.back_tab + 1) : make_int2(0, 0);
    for (int gi = 0; gi < a.n_groups; ++gi) {
      __syncthreads();
      if (gi + 1 < a.n_groups) stage_record((int4*)wsm + ((gi + 1) & 1) * a.max_rec16, a.stream + tnext.x, tnext.y, tid, WARPS * 32);
      if (gi + 2 < a.n_groups) tnext = __ldg(a.back_tab + gi + 2);
      const int* rec = (const int*)(wsm + (gi & 1) * a.max_rec16);
      const int nc = rec[0], nf = rec[1], ncnt = rec[2];
      for (int q = lane * 8; q < ncnt; q += 256) asm volatile("prefetch.global.L1 [%0];" ::"l"(G + nf + q));
      const int4* cols = (const int4*)(rec + 4);
      for (int ci = 0; ci < nc; ++ci) {
        const int4 c = cols[ci];   // {rcp_g, ent_begin, count, j}
        const int2* ents = (const int2*)(rec + c.y);
        int2 ce = make_int2(0, kWarpZero);
        cplx uv = make_double2(0.0, 0.0);
        if (lane < c.z) { ce = ents[lane]; uv = fb(ce.y); }
        const cplx xj = N::mul(acc[c.w], G[c.x]);
        __syncwarp();
        if (lane < c.z) acc[ce.x] = N::submul<false>(acc[ce.x], uv, xj);
        for (int t = lane + 32; t < c.z; t += 32) {
          const int2 e2 = ents[t];
          acc[e2.x] = N::submul<false>(acc[e2.x], fb(e2.y), xj);
        }
        if (lane == 0) acc[c.w] = xj;
        __syncwarp();
      }
      stage_wait();
    }
    // ---- status, results ----
    const bool good = __all_sync(FULL, ok);
    if (lane == 0 && valid) {
      a.status[p] = good ? ST_OK : -1;
      if (!good) a.fb_list[atomicAdd(a.fb_count, 1)] = p;
    }
    if (valid) {
      const long long xst = a.series_ld ? a.series_ld : 1;
      double2* xo = a.series_ld ? a.x + p : a.x + p * n;
      for (int i = lane; i < n; i += 32) xo[(long long)i * xst] = acc[i];
      if (a.ielem) {  // simulateAC.ts:94-126
        double2* io = a.series_ld ? a.ielem + p : a.ielem + p * a.n_ac_elem;
        for (int e = lane; e < a.n_ac_elem; e += 32) {
          cplx cur;
          if (e >= a.v_first) cur = acc[a.nn + e - a.v_first];
          else {
            const int4 en = __ldg(a.el_ends + e);
            const cplx v1 = en.x ? acc[en.x - 1] : make_double2(0.0, 0.0);
            const cplx v2 = en.y ? acc[en.y - 1] : make_double2(0.0, 0.0);
            const cplx Y = make_double2(__ldg(a.el_a + e), fma(w, __ldg(a.el_b + e), -__ldg(a.el_g + e) * iw));
            cur = N::mul(Y, csub(v1, v2));
          }
          io[(long long)e * xst] = cur;
        }
      }
    }
  }
}

}  // namespace spicey
