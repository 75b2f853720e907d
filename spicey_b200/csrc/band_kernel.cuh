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
//   BAND_RC       1: tables hold (alpha, beta) only (no inductors, real source phasors)
//   BAND_SYNC     n > 0: the warps of a CTA meet at a barrier n times per W steps (instruction-cache locality)
//   BAND_UMODE    how the pivot rows (U) leave for the workspace:
//                 0  one STG.128 per entry from the lane that owns the column: a warp's 32 lanes write 32 different
//                    128-byte lines per instruction (the workspace is column-major for the back-substitution)
//                 2  (W >= 4) the pivot record itself is the source of a TMA tensor store (cp.async.bulk.tensor.3d, one
//                    instruction per warp and step, issued by lane 0): the record of a warp's systems is laid out
//                    [column][system] in shared memory, the tensor map (second kernel parameter) describes the
//                    workspace as (row slot, system, column), box = one row slot x the warp's systems x W columns.
//                    No load / store unit cycle is spent on U.
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
// constants, value = alpha + j (w beta - gamma / w + Im J), where the window first needs them: one record per step,
// staged into a per-warp ring in shared memory with cp.async two steps ahead (the tables are the same for every system
// but stream through: nothing would keep them in L1), so that no instruction of a step waits on L2.
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

#ifndef BAND_RC
#define BAND_RC 0
#endif
#ifndef BAND_SYNC
#define BAND_SYNC 1
#endif
#ifndef BAND_UMODE
#define BAND_UMODE 0
#endif
#define BAND_TMA (BAND_UMODE == 2)
#define BW (BAND_L * BAND_RPL)
#define BNB BAND_NB
#define BPS (BW + BNB + 2)          /* pivot record: W band entries | NB border columns | rhs | diagonal */
#define BPX (BNB + 2)               /* the part of it that stays per system under BAND_TMA */
#define BPBUF (BAND_TMA ? 4 : 2)    /* pivot records alive: the tensor store reads a record up to 2 steps later */
#if BAND_UMODE != 0 && BAND_UMODE != 2
#error "BAND_UMODE is 0 or 2"
#endif
#if BAND_TMA && BW < 4
#error "BAND_UMODE 2 needs W >= 4 (the record buffer of step k is k mod 4 = s mod 4)"
#endif
// per-step record of the stamp tables (entries): new column [W] | column k+1 [W] | entering row: entry (k+W, k),
// border columns + rhs [NB+1] | border rows' new column [NB]; padded to a multiple of 8 entries
#define BST_NC 0
#define BST_LC BW
#define BST_E0 (2 * BW)
#define BST_ERB (2 * BW + 1)
#define BST_BRD (2 * BW + 1 + BNB + 1)
#define BST_FLAGS (2 * BW + 2 * BNB + 2)   /* tie-rule masks of the step, as the bits of the entry's first double */
#define BST_STRIDE ((2 * BW + 2 * BNB + 3 + 7) / 8 * 8)
#define BRING 4                       /* records in flight per warp */
#define BREC (BAND_RC ? 1 : 2)      /* double2 per table entry */
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
  const double4* el_rec;                // [n_ac_elem] {bits of (i1, i2), ya, yb, yg}: current = Y (x[i1] - x[i2]),
                                        // Y = ya + j (w yb - yg / w); i = n: the zero slot (ground); V: (branch, n, 1, 0, 0)
  const double* ind_L;
  int n, nb, n_out, n_ac_elem, n_ind;           // n = nb + NB unknowns incl. padding (nb: a multiple of W), n_out: the circuit's own
  int o_init, o_initb, o_brd0, o_bb0, o_step;   // table offsets (entries); o_step: record of step 0
  unsigned long long* prof;                     // BAND_PROF: cycles of [stamping | elimination | back-substitution | results], summed over warps
};
#ifndef BAND_PROF
#define BAND_PROF 0
#endif

#if BAND_PROF
#define BAND_TICK(i) { const long long t_ = clock64(); if (lane == 0) atomicAdd(a.prof + (i), (unsigned long long)(t_ - tick)); tick = t_; }
#else
#define BAND_TICK(i)
#endif

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
// x = 0 in the lanes where p holds: one predicated 64-bit move, no branch (the compiler's own choice for a
// conditional block of 32 assignments was a divergent region of ~100 moves)
__device__ __forceinline__ void band_zero_if(bool p, double& x) {
  asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %1, 0;\n\t@q mov.f64 %0, 0d0000000000000000;\n\t}" : "+d"(x) : "r"((int)p));
}

// an L2 load the compiler must leave where it is written (the software pipeline of the back-substitution)
__device__ __forceinline__ bcplx band_ldcg_now(const double2* ptr) {
  bcplx v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(ptr) : "memory");
  return v;
}

// stamped value of a table entry at angular frequency w (simulateAC.ts:36-57)
// BAND_RC: every entry of the circuit is (alpha, w beta) — no inductors, real source phasors — and the tables hold
// one double2 (alpha, beta) per entry
__device__ __forceinline__ bcplx band_rec(const double2* t, int idx, double w, double iw) {
#if BAND_RC
  const double2 c = __ldg(t + idx);
  return make_double2(c.x, w * c.y);
#else
  const double2 c0 = __ldg(t + 2 * idx), c1 = __ldg(t + 2 * idx + 1);
  return make_double2(c0.x, fma(w, c1.x, -c1.y * iw) + c0.y);
#endif
}
// the same from a staged record in shared memory
__device__ __forceinline__ bcplx band_rec_sm(const double2* t, int idx, double w, double iw) {
#if BAND_RC
  const double2 c = t[idx];
  return make_double2(c.x, w * c.y);
#else
  const double2 c0 = t[2 * idx], c1 = t[2 * idx + 1];
  return make_double2(c0.x, fma(w, c1.x, -c1.y * iw) + c0.y);
#endif
}
// one step record, global -> this warp's ring slot (16 bytes per lane and instruction), as one cp.async group
__device__ __forceinline__ void band_stage(double2* dst, const double2* src, int lane) {
#pragma unroll
  for (int i = 0; i < BST_STRIDE * BREC; i += 32)
    if (i + lane < BST_STRIDE * BREC) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i + lane);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i + lane) : "memory");
    }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

extern __shared__ __align__(128) double2 band_sm[];

#if BAND_TMA
struct __align__(64) BandTmap { unsigned long long opaque[16]; };   // a CUtensorMap (cuTensorMapEncodeTiled, spicey_native.cu)
extern "C" __global__ void __launch_bounds__(BAND_WARPS * 32, BAND_MINB) spicey_band_jit(BandArgs a, const __grid_constant__ BandTmap tm) {
#else
extern "C" __global__ void __launch_bounds__(BAND_WARPS * 32, BAND_MINB) spicey_band_jit(BandArgs a) {
#endif
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // The groups of a warp are interleaved: lane = l * (32 / L) + g.  The owners of one row in the 32 / L systems of a
  // warp are then neighbouring lanes, and the store that publishes a pivot-row entry is one shared-memory wavefront
  // instead of one per group (a 128-bit access is processed a quarter-warp at a time).
  const int g = lane % BGPW;                  // group (system) within the warp
  const int l = lane / BGPW;                  // lane within the group
  const int n = a.n, nb = a.nb;
  // shared memory: [BAND_TMA: per warp BPBUF pivot records [column][system], the TMA sources, 128-byte aligned |]
  // per warp a ring of BRING step records; per group x in elimination order | the pivot records
#if BAND_TMA
  double2* PRw = band_sm + (size_t)wib * (BPBUF * BW * BGPW);   // W BGPW 16 = 512 RPL bytes per record
  double2* PR = PRw + g;
  double2* band_sm1 = band_sm + (size_t)BAND_WARPS * (BPBUF * BW * BGPW);
  const int grp0 = (blockIdx.x * BAND_WARPS + wib) * BGPW;      // first system of this warp: the box's system coordinate
#else
  double2* band_sm1 = band_sm;
#endif
  double2* ring = band_sm1 + (size_t)wib * (BRING * BST_STRIDE * BREC);
  // x[n] is a zero (the ground node of the element records).  The stride is odd (in 16-byte units): the systems of a
  // warp read and write the same offsets of their areas in one instruction, and an odd stride puts up to eight of
  // them on eight different 16-byte bank groups (an even stride of 296 units made every access a 4-way conflict)
  const int sys_stride = (n + 1 + (BAND_TMA ? 2 * BPX : BPBUF * BPS)) | 1;
  double2* xs = band_sm1 + (size_t)BAND_WARPS * (BRING * BST_STRIDE * BREC) + (size_t)(wib * BGPW + g) * sys_stride;
  double2* Pb = xs + n + 1;
  // The pivot record of the step at unroll position s_ (s_ = k mod W; W is a multiple of BPBUF, so the buffer index is
  // a constant of the unrolled code): P_BAND(s_, t) = the entry in column slot t, P_EXT(s_, j) = border column j |
  // rhs (j = NB) | diagonal (j = NB + 1).  Under BAND_TMA the band part is stored by column, (t - s_ - 1) mod W =
  // the column's distance from k + 1, which is the order the tensor store wants and still a constant.
#if BAND_TMA
#define P_BAND(s_, t) PR[((((s_) & (BPBUF - 1)) * BW) + (((t) - (s_) - 1) & (BW - 1))) * BGPW]
#define P_EXT(s_, j) Pb[((s_) & 1) * BPX + (j)]
#else
#define P_BAND(s_, t) Pb[((s_) & (BPBUF - 1)) * BPS + (t)]
#define P_EXT(s_, j) Pb[((s_) & (BPBUF - 1)) * BPS + BW + (j)]
#endif
  const long long n_groups = (long long)gridDim.x * BAND_WARPS * BGPW;
  const long long grp = ((long long)blockIdx.x * BAND_WARPS + wib) * BGPW + g;
  double2* Gu = a.G + grp * a.g_stride;                 // [(nb + W) columns][W]
  double2* Gb = Gu + (size_t)(nb + BW) * BW;            // [nb][NB + 1]
  double2* Gr = Gb + (size_t)nb * (BNB + 1);            // [nb]
  const double2* tab = a.tab;

  // Rows above the matrix (the first W unknowns have fewer than W rows above them): their U entries are read as
  // zeros by the back-substitution and never written by anybody, so one fill per launch is enough.
  for (int q = l; q < BW * BW; q += BAND_L) Gu[q] = make_double2(0.0, 0.0);
#if BAND_TMA
  for (int q = l; q < 2 * BPX; q += BAND_L) Pb[q] = make_double2(0.0, 0.0);
  for (int q = lane; q < BPBUF * BW * BGPW; q += 32) PRw[q] = make_double2(0.0, 0.0);
  asm volatile("fence.proxy.async;" ::: "memory");   // the fill above precedes the tensor stores to the same lines
#else
  for (int q = l; q < BPBUF * BPS; q += BAND_L) Pb[q] = make_double2(0.0, 0.0);
#endif
  if (l == 0) xs[n] = make_double2(0.0, 0.0);
  __syncwarp();

  // CTA-uniform trip count (the warps of a CTA meet at a barrier inside): groups past the end solve the last point
  // again and store nothing
  for (long long cta_base = (long long)blockIdx.x * BAND_WARPS * BGPW; cta_base < a.p_count; cta_base += n_groups) {
    const long long base = cta_base + wib * BGPW;
    const bool valid = base + g < a.p_count;
    const long long p = valid ? base + g : a.p_count - 1;
    const double w = B_TWO_PI * a.freqs[p];
    const double iw = 1.0 / w;
#if BAND_PROF
    long long tick = clock64();
#endif
#define REC(idx) band_rec(tab, (idx), w, iw)
    const double2* steps = tab + (size_t)a.o_step * BREC;
    __syncwarp();   // the previous system is done with the ring
    band_stage(ring, steps, lane);
    band_stage(ring + BST_STRIDE * BREC, steps + (size_t)BST_STRIDE * BREC, lane);
    unsigned bad = 0u;   // any verification of this lane failed (kept branch-free: OR of predicates)
    for (int q = l; q < a.n_ind; q += BAND_L) {   // inductor guards of simulateAC.ts:47-51: the dense kernel decides
      const double d = w * a.ind_L[q];
      bad |= (unsigned)(fabs(d) < B_EPS) | (unsigned)(d * d < B_EPS);
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
#if BAND_TMA
    asm volatile("fence.proxy.async;" ::: "memory");   // the previous system's reads of the workspace precede this one's tensor stores
#endif
    if (l == 0) {   // row 0 is the first pivot row
      P_EXT(0, BNB + 1) = A[0][0];
#pragma unroll
      for (int t = 1; t < BW; ++t) P_BAND(0, t) = A[0][t];
      P_BAND(0, 0) = REC(a.o_step + BST_NC + 0);          // a[0][W]: column W enters with step 0
#pragma unroll
      for (int j = 0; j <= BNB; ++j)
        if (j == BNB || ((BAND_ABMASK >> j) & 1)) P_EXT(0, j) = AB[0][j];
    }
    __syncwarp();

    BAND_TICK(0)
    // ---- elimination of the band columns ----
    for (int kb = 0; kb < nb; kb += BW) {
#pragma unroll
      for (int s = 0; s < BW; ++s) {
        const int k = kb + s;
#if BAND_SYNC
        // The unrolled body is several times the instruction cache: warps that run different parts of it evict each
        // other's lines (25 % of the stall samples were instruction fetch).  Meeting BAND_SYNC times per W steps keeps
        // them on the same lines without putting them in lock step.
        if (s % (BW / (BAND_SYNC < BW ? BAND_SYNC : BW)) == 0) __syncthreads();
#endif
        // (always true — nb is a multiple of W: band_plan.h pads with identity rows — but the branch keeps the W steps
        //  separate basic blocks: scheduled as one, the compiler hoists the next step's loads over this step's updates
        //  and spills)
        if (k < nb) {
          const int pl = s % BAND_L, rs = s / BAND_L;              // owner lane / row slot of rows = s (mod W)
          const int s1 = (s + 1) % BW, pl1 = s1 % BAND_L, rs1 = s1 / BAND_L;
          // records k and k + 1 have landed (all but the newest cp.async group), record k + 2 leaves now; the barrier
          // also publishes the pivot record written at the end of the previous step
          band_stage(ring + ((k + 2) & (BRING - 1)) * (BST_STRIDE * BREC), steps + (size_t)(k + 2) * (BST_STRIDE * BREC), lane);
          asm volatile("cp.async.wait_group 1;" ::: "memory");
#if BAND_TMA
          // the owners' stores of record k become visible to the async proxy; the tensor store of record k - 3 has read
          // its buffer, which record k + 1 takes at the end of this step
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
#endif
          __syncwarp();
#if BAND_TMA
          if (lane == 0) {   // U row k: columns k + 1 .. k + W of row slot s, for the BGPW systems of this warp
            const unsigned src = (unsigned)__cvta_generic_to_shared(PRw + (s & (BPBUF - 1)) * (BW * BGPW));
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"((unsigned long long)&tm), "r"(2 * s), "r"(grp0), "r"(k + 1), "r"(src) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
#endif
          const double2* rk = ring + (k & (BRING - 1)) * (BST_STRIDE * BREC);
          const double2* rk1 = ring + ((k + 1) & (BRING - 1)) * (BST_STRIDE * BREC);
#define RECS(off) band_rec_sm(rk, (off), w, iw)
          const bcplx dg = P_EXT(s, BNB + 1);
          const uint2 fl = *(const uint2*)(rk + BST_FLAGS * BREC);
          const double mp = band_mag(dg);
          bad |= (unsigned)!(mp >= B_EPS);            // singular / Complex.div guard / NaN
          const double inv = band_rcp(mp);
          const bcplx r = make_double2(dg.x * inv, -dg.y * inv);
          const double thr_mp = B_THR * mp;           // |f|^2 = |a_ik|^2 / |pivot|^2 < EPS^2  (solveComplex.ts:46)
          // multipliers of my rows; the new column k + W takes over the slot of column k
          bcplx F[BAND_RPL];
#pragma unroll
          for (int q = 0; q < BAND_RPL; ++q) {
            bcplx aik = A[q][s];
            A[q][s] = RECS(BST_NC + l + BAND_L * q);
            if (q == rs) {   // in lane pl the pivot row's registers become the entering row k + W (zero so far)
              const bool own = l == pl;
              const bcplx e0 = RECS(BST_E0);
              aik = own ? e0 : aik;
#pragma unroll
              for (int t = 0; t < BW; ++t) { band_zero_if(own, A[q][t].x); band_zero_if(own, A[q][t].y); }
#pragma unroll
              for (int j = 0; j <= BNB; ++j)
                if (j == BNB || ((BAND_ABMASK >> j) & 1)) { const bcplx eb = RECS(BST_ERB + j); AB[q][j] = own ? eb : AB[q][j]; }
            }
            // column k + 1: diagonal and lower entries of the rows k+1 .. k+W arrive one step before they are read
            { const bcplx lc = RECS(BST_LC + l + BAND_L * q); A[q][s1].x += lc.x; A[q][s1].y += lc.y; }
            const double m = band_mag(aik);
            const bool strict = (fl.x >> (l + BAND_L * q)) & 1u;
            bad |= strict ? (unsigned)!(m < mp) : (unsigned)(m > mp);          // solveComplex.ts:18-28: first maximum wins
            const bcplx f = band_mul(aik, r);
            F[q] = m < thr_mp ? make_double2(0.0, 0.0) : f;
          }
          // border rows: column k's entry sits in lane pl, slot rs
          bcplx FB[BNB > 0 ? BNB : 1];
#pragma unroll
          for (int b = 0; b < BNB; ++b) {
            const bcplx bk = BR[b][rs];
            const double m = band_mag(bk);
            if (l == pl) {
              const bool strict = (fl.y >> b) & 1u;
              bad |= strict ? (unsigned)!(m < mp) : (unsigned)(m > mp);
              BR[b][rs] = RECS(BST_BRD + b);
            }
            bcplx f = band_mul(bk, r);
            f = m < thr_mp ? make_double2(0.0, 0.0) : f;
            f.x = __shfl_sync(FULL, f.x, pl * BGPW + g);
            f.y = __shfl_sync(FULL, f.y, pl * BGPW + g);
            FB[b] = f;
          }
          // the update: a_ij -= f_i * u_kj
#pragma unroll
          for (int t = 0; t < BW; ++t) {
            const bcplx pt = P_BAND(s, t);
#pragma unroll
            for (int q = 0; q < BAND_RPL; ++q) A[q][t] = band_submul(A[q][t], F[q], pt);
          }
#pragma unroll
          for (int j = 0; j <= BNB; ++j)
            if (j == BNB || ((BAND_ABMASK >> j) & 1)) {
              const bcplx pt = P_EXT(s, j);
#pragma unroll
              for (int q = 0; q < BAND_RPL; ++q) AB[q][j] = band_submul(AB[q][j], F[q], pt);
#pragma unroll
              for (int b = 0; b < BNB; ++b) BB[b][j] = band_submul(BB[b][j], FB[b], pt);
              if (l == pl) __stcg(Gb + (size_t)k * (BNB + 1) + j, pt);
            }
          // my columns of the pivot row: border rows' update, and U leaves for the workspace (column-major)
#pragma unroll
          for (int q = 0; q < BAND_RPL; ++q) {
            const int slot = l + BAND_L * q;
#if BNB > 0 || BAND_UMODE != 2
            const bcplx pt = P_BAND(s, slot);
#endif
#pragma unroll
            for (int b = 0; b < BNB; ++b) BR[b][q] = band_submul(BR[b][q], FB[b], pt);
#if BAND_UMODE == 0
            const int c = k + 1 + ((slot - s - 1) & (BW - 1));
            __stcg(Gu + (size_t)c * BW + s, pt);
#endif
          }
          if (l == pl) __stcg(Gr + k, r);
          // row k + 1 is final: its owner publishes it as the next pivot record
          if (l == pl1) {
            P_EXT(s + 1, BNB + 1) = A[rs1][s1];
#pragma unroll
            for (int t = 0; t < BW; ++t)
              if (t != s1) P_BAND(s + 1, t) = A[rs1][t];
            P_BAND(s + 1, s1) = band_rec_sm(rk1, BST_NC + s1, w, iw);    // a[k+1][k+1+W]
#pragma unroll
            for (int j = 0; j <= BNB; ++j)
              if (j == BNB || ((BAND_ABMASK >> j) & 1)) P_EXT(s + 1, j) = AB[rs1][j];
          }
#undef RECS
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    BAND_TICK(1)
#if BAND_TMA
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every U row is in the workspace
#endif

    // ---- the border block: NB x NB, replicated; same verification ----
    bcplx RB[BNB > 0 ? BNB : 1];
#pragma unroll
    for (int b = 0; b < BNB; ++b) {
      const uint2 fl = __ldg(a.flags + nb + b);
      const double mp = band_mag(BB[b][b]);
      bad |= (unsigned)!(mp >= B_EPS);
      const double inv = band_rcp(mp);
      RB[b] = make_double2(BB[b][b].x * inv, -BB[b][b].y * inv);
#pragma unroll
      for (int b2 = b + 1; b2 < BNB; ++b2) {
        const double m = band_mag(BB[b2][b]);
        const bool strict = (fl.y >> b2) & 1u;
        bad |= strict ? (unsigned)!(m < mp) : (unsigned)(m > mp);
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
    __syncwarp();   // the workspace written by other lanes of the group (BAND_TMA: by lane 0's tensor stores) is visible
#if BAND_TMA
    asm volatile("fence.proxy.async;" ::: "memory");
#endif

    // ---- back-substitution, column oriented: x_j by its owner, then every row of column j takes its term ----
    // row i: lane i mod L, slot (i / L) mod RPL, accumulator = b_i - sum over the border columns - sum_j u_ij x_j
    auto row_rhs = [&](int i) -> bcplx {
      bcplx acc = __ldcg(Gb + (size_t)i * (BNB + 1) + BNB);
#pragma unroll
      for (int j = 0; j < BNB; ++j)
        if ((BAND_ABMASK >> j) & 1) acc = band_submul(acc, __ldcg(Gb + (size_t)i * (BNB + 1) + j), XB[j]);
      return acc;
    };
    bcplx ACC[BAND_RPL];
#pragma unroll
    for (int q = 0; q < BAND_RPL; ++q) {
      const int slot = l + BAND_L * q;
      const int i = nb - 1 - ((nb - 1 - slot) & (BW - 1));    // the row = slot (mod W) among nb-W .. nb-1
      ACC[q] = i >= 0 ? row_rhs(i) : make_double2(0.0, 0.0);
    }
    // Software pipeline over half blocks of BH = W / 2 steps: the U columns of the next half block (and, once per
    // block, the reciprocals and entering-row accumulators of the steps this lane owns, s = l + L q) are loaded while
    // the current half block runs its dependent chain, so the chain never waits for the workspace (which sits in L2 or DRAM).
    {
      constexpr int BH = BW / 2;
      bcplx Ua[BH][BAND_RPL], Ub[BH][BAND_RPL];
      bcplx RJ[BAND_RPL], ENT[BAND_RPL], RJn[BAND_RPL], ENTn[BAND_RPL];
      // (volatile: the compiler otherwise sinks a half block's loads to their first use — the r3c capture has a
      //  ~2,900-cycle stall per block on a DFMA thirty instructions behind its LDG — which undoes the pipeline)
#define BAND_LOAD_U(U, j0)                                                                                   \
      _Pragma("unroll") for (int s_ = 0; s_ < BH; ++s_)                                                      \
        _Pragma("unroll") for (int q = 0; q < BAND_RPL; ++q) U[s_][q] = band_ldcg_now(Gu + (size_t)((j0) + s_) * BW + l + BAND_L * q);
#define BAND_LOAD_RE(R, E, jb_)                                                                              \
      _Pragma("unroll") for (int q = 0; q < BAND_RPL; ++q) {                                                 \
        const int j = (jb_) + l + BAND_L * q;                                                                \
        R[q] = __ldcg(Gr + j);                                                                               \
        E[q] = j - BW >= 0 ? row_rhs(j - BW) : make_double2(0.0, 0.0);                                       \
      }
#define BAND_BSTEP(s, uv)                                                                                    \
      {                                                                                                      \
        const int j = jb + (s);                                                                              \
        {                                                                                                    \
          const int pl = (s) % BAND_L, rs = (s) / BAND_L;                                                    \
          bcplx xj = band_mul(ACC[rs], RJ[rs]);                                                              \
          xj.x = __shfl_sync(FULL, xj.x, pl * BGPW + g);                                                     \
          xj.y = __shfl_sync(FULL, xj.y, pl * BGPW + g);                                                     \
          if (l == pl) {                                                                                     \
            xs[j] = xj;                                                                                      \
            ACC[rs] = ENT[rs];   /* row j - W enters */                                                      \
          }                                                                                                  \
          _Pragma("unroll") for (int q = 0; q < BAND_RPL; ++q) ACC[q] = band_submul(ACC[q], uv[q], xj);      \
        }                                                                                                    \
      }
      int jb = nb - BW;   // nb is a multiple of W
      BAND_LOAD_U(Ua, jb + BH)   // jb + BH + s < nb + W: inside the workspace
      BAND_LOAD_RE(RJ, ENT, jb)
      for (; jb >= 0; jb -= BW) {
        if (jb >= 2 * BW) {   // the block after next, from DRAM into L2: 16 W^2 bytes of U columns, one 128-byte line per lane and pass
          const char* nx = (const char*)(Gu + (size_t)(jb - 2 * BW) * BW);
#pragma unroll
          for (int ln = 0; ln < (BW * BW * 16 + 127) / 128; ln += BAND_L)
            if (ln + l < (BW * BW * 16 + 127) / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (ln + l) * 128));
        }
        BAND_LOAD_U(Ub, jb)
        __syncwarp();   // ptxas does not move the loads below a warp barrier (it sank them into the second half block)
#pragma unroll
        for (int s = BW - 1; s >= BH; --s) BAND_BSTEP(s, Ua[s - BH])
        if (jb >= BW) {
          BAND_LOAD_U(Ua, jb - BW + BH)
          BAND_LOAD_RE(RJn, ENTn, jb - BW)
        }
#pragma unroll
        for (int s = BH - 1; s >= 0; --s) BAND_BSTEP(s, Ub[s])
#pragma unroll
        for (int q = 0; q < BAND_RPL; ++q) { RJ[q] = RJn[q]; ENT[q] = ENTn[q]; }
      }
#undef BAND_LOAD_U
#undef BAND_LOAD_RE
#undef BAND_BSTEP
    }
    __syncwarp();
    BAND_TICK(2)

    // ---- status, results (simulateAC.ts:85-126) ----
    const unsigned vote = __ballot_sync(FULL, bad != 0u);
    unsigned gmask = 0u;   // the lanes of my group: g, g + 32 / L, ...
#pragma unroll
    for (int q = 0; q < BAND_L; ++q) gmask |= 1u << (q * BGPW + g);
    const bool good = (vote & gmask) == 0u;
    if (valid) {
      if (l == 0) {
        a.status[p] = good ? 0 : -1;
        if (!good) a.fb_list[atomicAdd(a.fb_count, 1)] = p;
      }
      // four independent loads in flight per lane: the index / constant tables are read through L2 behind the
      // workspace traffic, and one dependent load per iteration was 5 % of the kernel
      const long long xst = a.series_ld ? a.series_ld : 1;
      const int n_out = a.n_out;
      double2* xo = a.series_ld ? a.x + p : a.x + p * n_out;
      for (int i0 = l; i0 < n_out; i0 += 4 * BAND_L) {
        int idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) idx[u] = i0 + u * BAND_L < n_out ? __ldg(a.newvar + i0 + u * BAND_L) : n;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i0 + u * BAND_L < n_out) xo[(long long)(i0 + u * BAND_L) * xst] = xs[idx[u]];
      }
#if BAND_IELEM
      double2* io = a.series_ld ? a.ielem + p : a.ielem + p * a.n_ac_elem;
      const int ne = a.n_ac_elem;
      for (int e0 = l; e0 < ne; e0 += 4 * BAND_L) {
        double4 er[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = min(e0 + u * BAND_L, ne - 1);
          const double2 lo = __ldg((const double2*)(a.el_rec + e)), hi = __ldg((const double2*)(a.el_rec + e) + 1);
          er[u] = make_double4(lo.x, lo.y, hi.x, hi.y);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * BAND_L;
          const long long ij = __double_as_longlong(er[u].x);
          const bcplx v1 = xs[(int)(ij & 0xffffffffll)], v2 = xs[(int)(ij >> 32)];
          const bcplx Y = make_double2(er[u].y, fma(w, er[u].z, -er[u].w * iw));
          const bcplx cur = band_mul(Y, make_double2(v1.x - v2.x, v1.y - v2.y));
          if (e < ne) io[(long long)e * xst] = cur;
        }
      }
#endif
    }
    BAND_TICK(3)
#undef REC
  }
}
