// Register-resident persistent transient kernel for small systems (Nvar <= 6): one thread per
// Monte-Carlo / sweep instance, the whole time loop of simulateTRAN (lib/analysis/
// simulateTRAN.ts:146-238) in one launch.
//
// What differs from the generic thread tier (tran_kernels.cuh):
//  * NV (padded system size) and NT (threads per CTA) are template parameters: the LU factors, the
//    right-hand side and the solution live in REGISTERS, every loop over matrix indices is unrolled
//    without run-time guards (a system smaller than NV is padded with identity rows, which never win
//    a pivot search), pivoting is a chain of selects, and every shared-memory address is
//    base + compile-time stride.
//  * Stamping follows the reference literally — per-type element loops in the order R, C, L, S,
//    V, D (simulateTRAN.ts:35-101) scattering into a per-thread [entry][thread] shared-memory
//    image with one extra row/column that absorbs ground stamps, so the loops are branch-free
//    and the floating-point summation order is the reference's.  Node ids are converted once per
//    CTA into byte offsets of that image.
//  * The reference re-stamps and re-factors the matrix at every step (:152-157).  The matrix
//    only changes when a switch toggles or a diode is re-linearised, and Gaussian elimination is
//    deterministic, so the kernel factors once and afterwards only replays the recorded row
//    swaps and multipliers on the new right-hand side: the same operations on the same operands
//    in the same order, i.e. the reference's numbers, at O(n^2) instead of O(n^3) per step.
//    (Circuits with diodes re-factor every solve; circuits with switches when a state changed.)
#pragma once
#include "tran_kernels.cuh"

namespace spicey {

// Per-element record in shared memory: byte offsets into the per-thread arrays.
struct SmallElem {
  int x1, x2;    // offsets of x[n1], x[n2] (also of b[n1], b[n2]); ground -> slot NV
  int c1, c2;    // switch control nodes
  int st;        // offset of the element's state slot (or of a scratch slot)
  int ec;        // offset of the element's 4 constants
  int a11, a22, a12, a21;  // offsets of A(n1,n1), A(n2,n2), A(n1,n2), A(n2,n1)
};

struct TranSmallSmem {
  size_t total, elem_off;
  int o_ac, o_as, o_b, o_x, o_st, o_ec, per_thread;
  __host__ __device__ TranSmallSmem(int nv, int n_elem, int n_state, bool dyn, int nt) {
    const int n1 = nv + 1;
    int o = 0;
    o_ac = o; o += n1 * n1;
    o_as = o; o += dyn ? n1 * n1 : 0;
    o_b = o; o += n1;
    o_x = o; o += n1;
    o_st = o; o += n_state + 1;
    o_ec = o; o += 4 * n_elem;
    per_thread = o;
    size_t bytes = sizeof(double) * (size_t)o * nt;
    bytes = (bytes + 15) & ~(size_t)15;
    elem_off = bytes;
    bytes += sizeof(SmallElem) * (size_t)(n_elem > 0 ? n_elem : 1);
    total = (bytes + 15) & ~(size_t)15;
  }
};

template <int NV, bool STRICT>
struct SmallLU {
  double f[NV][NV];  // upper: U (diagonal holds the pivot, or its reciprocal in fast mode); lower: multipliers
  int perm[NV];      // row chosen at step k (solveReal.ts:16-26)

  // solveReal.ts:14-54 on the matrix part; returns status.
  __device__ __forceinline__ int factor() {
    int status = ST_OK;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int imax = k;
      double vmax = fabs(f[k][k]);
#pragma unroll
      for (int i = k + 1; i < NV; ++i) {
        double v = fabs(f[i][k]);
        if (v > vmax) { vmax = v; imax = i; }
      }
      if (vmax < kEps) status = ST_SINGULAR;   // :28 (the caller stops; the arithmetic below is harmless)
      perm[k] = imax;
#pragma unroll
      for (int i = k + 1; i < NV; ++i) {
        const bool sw = (imax == i);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const double u = f[k][j], w = f[i][j];
          f[k][j] = sw ? w : u;
          f[i][j] = sw ? u : w;
        }
      }
      const double pivot = f[k][k];
      const double rp = 1.0 / pivot;
#pragma unroll
      for (int i = k + 1; i < NV; ++i) {
        double m = STRICT ? __ddiv_rn(f[i][k], pivot) : f[i][k] * rp;
        m = (fabs(m) < kEps) ? 0.0 : m;  // :45 skip == zero multiplier
        f[i][k] = m;
#pragma unroll
        for (int j = k + 1; j < NV; ++j) f[i][j] = Num<double>::submul<STRICT>(f[i][j], m, f[k][j]);
      }
      if (!STRICT) f[k][k] = rp;
    }
    return status;
  }

  // Replays the elimination on b (the augmented column of solveReal.ts) and back-substitutes (:56-71).
  // factor() swapped whole rows, stored multipliers included (as LAPACK does), so the multipliers are in
  // FINAL row order: every interchange is applied to b first, then the eliminations.  Each b entry still
  // meets the same multipliers in the same order as in the reference's augmented elimination.
  __device__ __forceinline__ void solve(double (&b)[NV]) const {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int i = k + 1; i < NV; ++i) {
        const bool sw = (perm[k] == i);
        const double u = b[k], w = b[i];
        b[k] = sw ? w : u;
        b[i] = sw ? u : w;
      }
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int i = k + 1; i < NV; ++i) b[i] = Num<double>::submul<STRICT>(b[i], f[i][k], b[k]);
#pragma unroll
    for (int i = NV - 1; i >= 0; --i) {
      double s = b[i];
#pragma unroll
      for (int j = i + 1; j < NV; ++j) s = Num<double>::submul<STRICT>(s, f[i][j], b[j]);
      b[i] = STRICT ? __ddiv_rn(s, f[i][i]) : s * f[i][i];
    }
  }
};

#define SM_D(base, off) (*(double*)((base) + (off)))

template <int NV, int NT, bool STRICT>
__global__ void __launch_bounds__(NT) tran_small_kernel(DevPlan P, TranArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int N1 = NV + 1;        // row/column NV absorbs ground stamps
  constexpr int SB = NT * 8;        // byte stride between consecutive per-thread array entries
  const int t = threadIdx.x;
  const int nvar = P.nvar, nn = P.nn, ne = P.n_elem, ns = P.n_state;
  const int oC = P.off[ELEM_C], oL = P.off[ELEM_L], oV = P.off[ELEM_V], oS = P.off[ELEM_S], oD = P.off[ELEM_D],
            oE = P.off[ELEM_D + 1];
  const bool dyn = oE > oS, has_diode = oE > oD;
  const TranSmallSmem L(NV, ne, ns, dyn, NT);
  unsigned char* tb = smem + t * 8;
  unsigned char* Ac = tb + (size_t)L.o_ac * SB;
  unsigned char* As = tb + (size_t)L.o_as * SB;
  unsigned char* bs = tb + (size_t)L.o_b * SB;
  unsigned char* xs = tb + (size_t)L.o_x * SB;
  unsigned char* st = tb + (size_t)L.o_st * SB;
  unsigned char* ec = tb + (size_t)L.o_ec * SB;
  SmallElem* el = (SmallElem*)(smem + L.elem_off);
  for (int e = t; e < ne; e += NT) {
    const int4 q = P.ends[e];
    const int i1 = q.x ? q.x - 1 : NV, i2 = q.y ? q.y - 1 : NV;
    SmallElem r;
    r.x1 = i1 * SB; r.x2 = i2 * SB;
    r.c1 = (q.z ? q.z - 1 : NV) * SB; r.c2 = (q.w ? q.w - 1 : NV) * SB;
    const int s = P.state_idx[e];
    r.st = (s >= 0 ? s : ns) * SB;
    r.ec = 4 * e * SB;
    r.a11 = (i1 * N1 + i1) * SB; r.a22 = (i2 * N1 + i2) * SB;
    r.a12 = (i1 * N1 + i2) * SB; r.a21 = (i2 * N1 + i1) * SB;
    el[e] = r;
  }
  __syncthreads();
  const long long li = (long long)blockIdx.x * NT + t;
  if (li >= a.n_local) return;
  const long long inst = a.inst0 + li, NL = a.n_local, S1 = a.steps + 1;
  const double dtc = fmax(a.dt, kEps);

  for (int e = 0; e < ne; ++e) {
    double c4[4];
    element_constants(P, P.meta[e].x, P.meta[e].y, inst, dtc, c4);
    SM_D(ec, (4 * e + 0) * SB) = c4[0]; SM_D(ec, (4 * e + 1) * SB) = c4[1];
    SM_D(ec, (4 * e + 2) * SB) = c4[2]; SM_D(ec, (4 * e + 3) * SB) = c4[3];
  }
  for (int s = 0; s < ns; ++s) SM_D(st, s * SB) = a.state0 ? a.state0[(long long)s * P.n_inst + inst] : 0.0;
  // Constant part of A: R, C, L admittances then V incidence (V entries never overlap an admittance),
  // identity on the padding rows nvar..NV-1.
  for (int i = 0; i < N1 * N1; ++i) SM_D(Ac, i * SB) = 0.0;
  for (int p = nvar; p < NV; ++p) SM_D(Ac, (p * N1 + p) * SB) = 1.0;
#define STAMP_Y(M, r, g)                                      \
  do {  /* stampAdmittanceReal.ts:3-29, same order */         \
    SM_D(M, (r).a11) += (g);                                  \
    SM_D(M, (r).a22) += (g);                                  \
    SM_D(M, (r).a12) -= (g);                                  \
    SM_D(M, (r).a21) -= (g);                                  \
  } while (0)
  #pragma unroll 1
  for (int e = 0; e < oV; ++e) {
    const SmallElem r = el[e];
    STAMP_Y(Ac, r, SM_D(ec, r.ec));
  }
  #pragma unroll 1
  for (int e = oV; e < oS; ++e) {  // stampVoltageSourceReal.ts:4-32
    const SmallElem r = el[e];
    const int j = nn + (e - oV);
    SM_D(Ac, (r.x1 / SB * N1 + j) * SB) += 1.0;
    SM_D(Ac, (r.x2 / SB * N1 + j) * SB) -= 1.0;
    SM_D(Ac, j * N1 * SB + r.x1) += 1.0;
    SM_D(Ac, j * N1 * SB + r.x2) -= 1.0;
  }
  if (dyn)
    for (int i = 0; i < N1 * N1; ++i) SM_D(As, i * SB) = SM_D(Ac, i * SB);
  SM_D(xs, NV * SB) = 0.0;  // ground
#pragma unroll
  for (int i = 0; i < NV; ++i) SM_D(xs, i * SB) = 0.0;

  SmallLU<NV, STRICT> lu;
  bool factored = false;
  int status = ST_OK;
  long long step = 0;
  double x[NV];
  double* vo = a.v + li;
  double* io = a.ielem ? a.ielem + li : nullptr;
  int* ito = a.iters ? a.iters + li : nullptr;
  const long long v_stride = (long long)nn * NL, i_stride = (long long)ne * NL;
  const double* vsrc = a.vsrc;
  const unsigned vmask = a.vmask_bits, wmask = a.wmask_bits;   // the host routes nV > 32 to the generic tiers
  for (; step < S1; ++step) {
    // :149 zeroes x every step; nothing reads x before the step's first solve (the diode uses vdPrev at
    // iteration 0, :85), so the zeroing is not materialised.
    int it = 0;
    for (; it < 20; ++it) {                                                  // :151
      // ---- right-hand side, reference order C, L, V, D (:41-53, :66-69, :98-100) ----
#pragma unroll
      for (int i = 0; i < N1; ++i) SM_D(bs, i * SB) = 0.0;
      #pragma unroll 1
      for (int e = oC; e < oL; ++e) {
        const SmallElem r = el[e];
        const double ieq = t_mul<STRICT>(-SM_D(ec, r.ec), SM_D(st, r.st));
        SM_D(bs, r.x1) -= ieq;
        SM_D(bs, r.x2) += ieq;
      }
      #pragma unroll 1
      for (int e = oL; e < oV; ++e) {
        const SmallElem r = el[e];
        const double ip = SM_D(st, r.st);
        SM_D(bs, r.x1) -= ip;
        SM_D(bs, r.x2) += ip;
      }
      #pragma unroll 1
      for (int e = oV; e < oS; ++e) {
        const int k = e - oV;
        double vs;
        if ((wmask >> k) & 1u) vs = wave_value(P, a.waves[k], inst, __dmul_rn((double)step, a.dt));   // :147 t = step*dt
        else vs = ((vmask >> k) & 1u) ? __ldg(vsrc + (long long)k * S1 + step) : SM_D(ec, 4 * e * SB);
        SM_D(bs, (nn + k) * SB) += vs;
      }
      // ---- dynamic part of A: switches then diodes (:56-63, :72-101) ----
      if (dyn) {
        // switches only: `factored` is cleared where a switch toggles; diodes: re-linearised every solve
        const bool dirty = !factored || has_diode;
        if (dirty) {
#pragma unroll
          for (int i = 0; i < N1 * N1; ++i) SM_D(As, i * SB) = SM_D(Ac, i * SB);
          #pragma unroll 1
          for (int e = oS; e < oD; ++e) {
            const SmallElem r = el[e];
            const double g = 1 / (SM_D(st, r.st) != 0.0 ? SM_D(ec, r.ec) : SM_D(ec, r.ec + SB));
            STAMP_Y(As, r, g);
          }
        }
        #pragma unroll 1
        for (int e = oD; e < oE; ++e) {
          const SmallElem r = el[e];
          const double vd = it == 0 ? SM_D(st, r.st) : SM_D(xs, r.x1) - SM_D(xs, r.x2);   // :85
          double gd, ieq;
          diode_companion<STRICT>(vd, SM_D(ec, r.ec), SM_D(ec, r.ec + SB), SM_D(ec, r.ec + 2 * SB), SM_D(ec, r.ec + 3 * SB), gd, ieq);
          STAMP_Y(As, r, gd);
          SM_D(bs, r.x1) -= ieq;
          SM_D(bs, r.x2) += ieq;
        }
        if (dirty) {
#pragma unroll
          for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) lu.f[i][j] = SM_D(As, (i * N1 + j) * SB);
          status = lu.factor();
          factored = true;
        }
      } else if (!factored) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int j = 0; j < NV; ++j) lu.f[i][j] = SM_D(Ac, (i * N1 + j) * SB);
        status = lu.factor();
        factored = true;
      }
      if (status != ST_OK) break;
#pragma unroll
      for (int i = 0; i < NV; ++i) x[i] = SM_D(bs, i * SB);
      lu.solve(x);
#pragma unroll
      for (int i = 0; i < NV; ++i) SM_D(xs, i * SB) = x[i];
      bool switched = false;                                                 // :108-128
      #pragma unroll 1
      for (int e = oS; e < oD; ++e) {
        const SmallElem r = el[e];
        const double vctrl = SM_D(xs, r.c1) - SM_D(xs, r.c2);
        const bool on = SM_D(st, r.st) != 0.0;
        bool nxt = on;
        if (on) { if (vctrl < SM_D(ec, r.ec + 3 * SB)) nxt = false; }
        else if (vctrl > SM_D(ec, r.ec + 2 * SB)) nxt = true;
        if (nxt != on) { SM_D(st, r.st) = nxt ? 1.0 : 0.0; switched = true; }
      }
      if (switched) factored = false;  // conductances changed: re-factor at the next solve
      if (!switched) break;
    }
    if (status != ST_OK) break;
    if (ito) { *ito = it < 20 ? it + 1 : 20; ito += NL; }
    // ---- recording (:164-219) + state update (:221-237), per-type loops in table order ----
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nn) vo[(long long)i * NL] = x[i];
    vo += v_stride;
    double* ie = io;
    #pragma unroll 1
    for (int e = 0; e < oC; ++e, ie += NL) {
      const SmallElem r = el[e];
      const double d = SM_D(xs, r.x1) - SM_D(xs, r.x2);
      const double cur = STRICT ? __ddiv_rn(d, SM_D(ec, r.ec + SB)) : d * SM_D(ec, r.ec);
      if (io) *ie = cur;
    }
    #pragma unroll 1
    for (int e = oC; e < oL; ++e, ie += NL) {
      const SmallElem r = el[e];
      const double d = SM_D(xs, r.x1) - SM_D(xs, r.x2);
      const double dv = d - SM_D(st, r.st);
      const double cur = STRICT ? __ddiv_rn(__dmul_rn(SM_D(ec, r.ec + SB), dv), dtc) : SM_D(ec, r.ec) * dv;
      SM_D(st, r.st) = d;
      if (io) *ie = cur;
    }
    #pragma unroll 1
    for (int e = oL; e < oV; ++e, ie += NL) {
      const SmallElem r = el[e];
      const double d = SM_D(xs, r.x1) - SM_D(xs, r.x2);
      const double cur = t_add<STRICT>(t_mul<STRICT>(SM_D(ec, r.ec), d), SM_D(st, r.st));
      SM_D(st, r.st) = cur;
      if (io) *ie = cur;
    }
    #pragma unroll 1
    for (int e = oV; e < oS; ++e, ie += NL)
      if (io) *ie = SM_D(xs, (nn + e - oV) * SB);
    #pragma unroll 1
    for (int e = oS; e < oD; ++e, ie += NL) {
      const SmallElem r = el[e];
      const double d = SM_D(xs, r.x1) - SM_D(xs, r.x2);
      if (io) *ie = d / (SM_D(st, r.st) != 0.0 ? SM_D(ec, r.ec) : SM_D(ec, r.ec + SB));
    }
    #pragma unroll 1
    for (int e = oD; e < oE; ++e, ie += NL) {
      const SmallElem r = el[e];
      const double d = SM_D(xs, r.x1) - SM_D(xs, r.x2);
      const double ex = exp(STRICT ? __ddiv_rn(d, SM_D(ec, r.ec + SB)) : d * SM_D(ec, r.ec + 3 * SB));
      if (io) *ie = t_mul<STRICT>(SM_D(ec, r.ec), t_sub<STRICT>(ex, 1.0));  // unclamped vd (H6)
      SM_D(st, r.st) = d;
    }
    if (io) io += i_stride;
  }
#undef STAMP_Y
  if (status != ST_OK) {
    for (; step < S1; ++step) {
      for (int i = 0; i < nn; ++i) a.v[(step * nn + i) * NL + li] = CUDART_NAN;
      if (a.ielem) for (int e = 0; e < ne; ++e) a.ielem[(step * ne + e) * NL + li] = CUDART_NAN;
      if (a.iters) a.iters[step * NL + li] = 0;
    }
  }
  if (a.state_out) for (int s = 0; s < ns; ++s) a.state_out[(long long)s * NL + li] = SM_D(st, s * SB);
  a.status[li] = status;
}

#undef SM_D

}  // namespace spicey
