// Register-resident persistent transient kernel for small systems (Nvar <= 6): one thread per
// Monte-Carlo / sweep instance, the whole time loop of simulateTRAN (lib/analysis/
// simulateTRAN.ts:146-238) in one launch.
//
// What differs from the generic thread tier (tran_kernels.cuh):
//  * NV is a template parameter: the LU factors, the right-hand side and the solution live in
//    REGISTERS, every loop over matrix indices is unrolled, pivoting is a chain of selects.
//  * Stamping follows the reference literally — per-type element loops in the order R, C, L, S,
//    V, D (simulateTRAN.ts:35-101) scattering into a per-thread [entry][thread] shared-memory
//    image with one extra row/column that absorbs ground stamps, so the loops are branch-free
//    and the floating-point summation order is the reference's.
//  * The reference re-stamps and re-factors the matrix at every step (:152-157).  The matrix
//    only changes when a switch toggles or a diode is re-linearised, and Gaussian elimination is
//    deterministic, so the kernel factors once and afterwards only replays the recorded row
//    swaps and multipliers on the new right-hand side: the same operations on the same operands
//    in the same order, i.e. the reference's numbers, at O(n^2) instead of O(n^3) per step.
//    (Circuits with diodes re-factor every solve; circuits with switches when a state changed.)
//  * x, the companion state and per-element constants stay in shared memory because element
//    loops index them with run-time node ids.
#pragma once
#include "tran_kernels.cuh"

namespace spicey {

struct TranSmallSmem {
  size_t per_thread_doubles, ends_off, total;
  int o_ac, o_as, o_b, o_x, o_st, o_ec;  // per-thread array offsets in doubles (each scaled by NT)
  __host__ __device__ TranSmallSmem(int nv, int n_elem, int n_state, bool dyn, int nt) {
    const int n1 = nv + 1;
    int o = 0;
    o_ac = o; o += n1 * n1;
    o_as = o; o += dyn ? n1 * n1 : 0;
    o_b = o; o += n1;
    o_x = o; o += n1;
    o_st = o; o += n_state;
    o_ec = o; o += 4 * n_elem;
    per_thread_doubles = (size_t)o;
    size_t bytes = sizeof(double) * per_thread_doubles * nt;
    bytes = (bytes + 15) & ~(size_t)15;
    ends_off = bytes;
    bytes += sizeof(int4) * n_elem + sizeof(int) * n_elem;
    total = (bytes + 15) & ~(size_t)15;
  }
};

template <int NV, bool STRICT>
struct SmallLU {
  double f[NV][NV];  // upper: U (diagonal holds the pivot, or its reciprocal in fast mode); lower: multipliers
  int perm[NV];      // row chosen at step k (solveReal.ts:16-26)

  // solveReal.ts:14-54 on the matrix part; returns status.
  __device__ __forceinline__ int factor(int n) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (k < n) {
        int imax = k;
        double vmax = fabs(f[k][k]);
#pragma unroll
        for (int i = k + 1; i < NV; ++i)
          if (i < n) {
            double v = fabs(f[i][k]);
            if (v > vmax) { vmax = v; imax = i; }
          }
        if (vmax < kEps) return ST_SINGULAR;
        perm[k] = imax;
#pragma unroll
        for (int i = k + 1; i < NV; ++i)
          if (imax == i) {
#pragma unroll
            for (int j = 0; j < NV; ++j) { double t = f[k][j]; f[k][j] = f[i][j]; f[i][j] = t; }
          }
        const double pivot = f[k][k];
        const double rp = 1.0 / pivot;
#pragma unroll
        for (int i = k + 1; i < NV; ++i)
          if (i < n) {
            double m = STRICT ? __ddiv_rn(f[i][k], pivot) : f[i][k] * rp;
            if (fabs(m) < kEps) m = 0.0;  // :45 skip == zero multiplier
            f[i][k] = m;
#pragma unroll
            for (int j = k + 1; j < NV; ++j) f[i][j] = Num<double>::submul<STRICT>(f[i][j], m, f[k][j]);
          }
        if (!STRICT) f[k][k] = rp;
      }
    }
    return ST_OK;
  }

  // Replays the elimination on b (the augmented column of solveReal.ts) and back-substitutes (:56-71).
  __device__ __forceinline__ void solve(int n, double (&b)[NV]) const {
    // factor() swapped whole rows, stored multipliers included (as LAPACK does), so the multipliers are
    // in FINAL row order: apply every interchange to b first, then eliminate.  Each b entry still meets
    // the same multipliers in the same order as in the reference's augmented elimination.
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (k < n) {
#pragma unroll
        for (int i = k + 1; i < NV; ++i)
          if (perm[k] == i) { double t = b[k]; b[k] = b[i]; b[i] = t; }
      }
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (k < n) {
#pragma unroll
        for (int i = k + 1; i < NV; ++i)
          if (i < n && f[i][k] != 0.0) b[i] = Num<double>::submul<STRICT>(b[i], f[i][k], b[k]);
      }
#pragma unroll
    for (int i = NV - 1; i >= 0; --i)
      if (i < n) {
        double s = b[i];
#pragma unroll
        for (int j = i + 1; j < NV; ++j)
          if (j < n) s = Num<double>::submul<STRICT>(s, f[i][j], b[j]);
        b[i] = STRICT ? __ddiv_rn(s, f[i][i]) : s * f[i][i];
      }
  }
};

template <int NV, bool STRICT>
__global__ void __launch_bounds__(128) tran_small_kernel(DevPlan P, TranArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int t = threadIdx.x, NT = blockDim.x;
  const int nvar = P.nvar, nn = P.nn, ne = P.n_elem, ns = P.n_state;
  const bool dyn = P.off[ELEM_S + 1] > P.off[ELEM_S] || P.off[ELEM_D + 1] > P.off[ELEM_D];
  const bool has_diode = P.off[ELEM_D + 1] > P.off[ELEM_D];
  const TranSmallSmem L(NV, ne, ns, dyn, NT);
  constexpr int N1 = NV + 1;  // row/column NV absorbs ground stamps
  double* base = (double*)smem + t;
  double* Ac = base + (size_t)L.o_ac * NT;
  double* As = base + (size_t)L.o_as * NT;
  double* bs = base + (size_t)L.o_b * NT;
  double* xs = base + (size_t)L.o_x * NT;
  double* st = base + (size_t)L.o_st * NT;
  double* ec = base + (size_t)L.o_ec * NT;
  int4* ends = (int4*)(smem + L.ends_off);   // node ids mapped to matrix indices, ground -> NV
  int* sidx = (int*)(ends + ne);
  for (int e = t; e < ne; e += NT) {
    int4 q = P.ends[e];
    ends[e] = make_int4(q.x ? q.x - 1 : NV, q.y ? q.y - 1 : NV, q.z ? q.z - 1 : NV, q.w ? q.w - 1 : NV);
    sidx[e] = P.state_idx[e];
  }
  __syncthreads();
  const long long li = (long long)blockIdx.x * NT + t;
  if (li >= a.n_local) return;
  const long long inst = a.inst0 + li, NL = a.n_local, S1 = a.steps + 1;
  const double dtc = fmax(a.dt, kEps);
  const int oR = 0, oC = P.off[ELEM_C], oL = P.off[ELEM_L], oV = P.off[ELEM_V], oS = P.off[ELEM_S],
            oD = P.off[ELEM_D], oE = P.off[ELEM_D + 1];

  for (int e = 0; e < ne; ++e) {
    double c4[4];
    element_constants(P, P.meta[e].x, P.meta[e].y, inst, dtc, c4);
    for (int q = 0; q < 4; ++q) ec[(4 * e + q) * NT] = c4[q];
  }
  for (int s = 0; s < ns; ++s) st[s * NT] = a.state0 ? a.state0[(long long)s * P.n_inst + inst] : 0.0;
  // Constant part of A: R, C, L admittances then V incidence (the entries of V never overlap an admittance).
  for (int i = 0; i < N1 * N1; ++i) Ac[i * NT] = 0.0;
#define STAMP_Y(M, i1, i2, g)                                   \
  do {                                                          \
    M[((i1) * N1 + (i1)) * NT] += (g);                          \
    M[((i2) * N1 + (i2)) * NT] += (g);                          \
    M[((i1) * N1 + (i2)) * NT] -= (g);                          \
    M[((i2) * N1 + (i1)) * NT] -= (g);                          \
  } while (0)
  for (int e = oR; e < oS; ++e) {
    const int4 q = ends[e];
    if (e < oV) {
      // stampAdmittanceReal.ts:3-29 — when n1 == n2 the reference also adds and subtracts in this order
      STAMP_Y(Ac, q.x, q.y, ec[(4 * e) * NT]);
    } else {  // stampVoltageSourceReal.ts:4-32
      const int j = nn + (e - oV);
      Ac[(q.x * N1 + j) * NT] += 1.0;
      Ac[(q.y * N1 + j) * NT] -= 1.0;
      Ac[(j * N1 + q.x) * NT] += 1.0;
      Ac[(j * N1 + q.y) * NT] -= 1.0;
    }
  }

  xs[NV * NT] = 0.0;  // ground
  SmallLU<NV, STRICT> lu;
  bool factored = false;
  int status = ST_OK;
  long long step = 0;
  double x[NV];
  for (; step < S1; ++step) {
    // :149 zeroes x every step; nothing reads x before the step's first solve (the diode uses vdPrev at
    // iteration 0, :85), so the zeroing is not materialised.
    int it = 0;
    for (; it < 20; ++it) {                                                  // :151
      // ---- right-hand side, reference order C, L, V, D (:41-53, :66-69, :98-100) ----
      for (int i = 0; i < N1; ++i) bs[i * NT] = 0.0;
      for (int e = oC; e < oL; ++e) {
        const int4 q = ends[e];
        const double ieq = t_mul<STRICT>(-ec[(4 * e) * NT], st[sidx[e] * NT]);
        bs[q.x * NT] -= ieq;
        bs[q.y * NT] += ieq;
      }
      for (int e = oL; e < oV; ++e) {
        const int4 q = ends[e];
        const double ip = st[sidx[e] * NT];
        bs[q.x * NT] -= ip;
        bs[q.y * NT] += ip;
      }
      for (int e = oV; e < oS; ++e) {
        const int k = e - oV;
        bs[(nn + k) * NT] += a.vsrc_mask[k] ? a.vsrc[(long long)k * S1 + step] : ec[(4 * e) * NT];
      }
      // ---- dynamic part of A: switches then diodes (:56-63, :72-101) ----
      if (dyn) {
        // switches only: `factored` is cleared where a switch toggles; diodes: re-linearised every solve
        const bool dirty = !factored || has_diode;
        if (dirty) {
          for (int i = 0; i < N1 * N1; ++i) As[i * NT] = Ac[i * NT];
          for (int e = oS; e < oD; ++e) {
            const int4 q = ends[e];
            const double g = 1 / (st[sidx[e] * NT] != 0.0 ? ec[(4 * e) * NT] : ec[(4 * e + 1) * NT]);
            STAMP_Y(As, q.x, q.y, g);
          }
        }
        for (int e = oD; e < oE; ++e) {
          const int4 q = ends[e];
          const double vd = it == 0 ? st[sidx[e] * NT] : xs[q.x * NT] - xs[q.y * NT];   // :85
          double gd, ieq;
          diode_companion<STRICT>(vd, ec[(4 * e) * NT], ec[(4 * e + 1) * NT], gd, ieq);
          STAMP_Y(As, q.x, q.y, gd);
          bs[q.x * NT] -= ieq;
          bs[q.y * NT] += ieq;
        }
        if (dirty) {
#pragma unroll
          for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) lu.f[i][j] = As[(i * N1 + j) * NT];
          status = lu.factor(nvar);
          factored = true;
        }
      } else if (!factored) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int j = 0; j < NV; ++j) lu.f[i][j] = Ac[(i * N1 + j) * NT];
        status = lu.factor(nvar);
        factored = true;
      }
      if (status != ST_OK) break;
#pragma unroll
      for (int i = 0; i < NV; ++i) x[i] = bs[i * NT];
      lu.solve(nvar, x);
#pragma unroll
      for (int i = 0; i < NV; ++i) xs[i * NT] = x[i];
      bool switched = false;                                                 // :108-128
      for (int e = oS; e < oD; ++e) {
        const int4 q = ends[e];
        const double vctrl = xs[q.z * NT] - xs[q.w * NT];
        const bool on = st[sidx[e] * NT] != 0.0;
        bool nxt = on;
        if (on) { if (vctrl < ec[(4 * e + 3) * NT]) nxt = false; }
        else if (vctrl > ec[(4 * e + 2) * NT]) nxt = true;
        if (nxt != on) { st[sidx[e] * NT] = nxt ? 1.0 : 0.0; switched = true; }
      }
      if (switched) factored = false;  // conductances changed: re-factor at the next solve
      if (!switched) break;
    }
    if (status != ST_OK) break;
    if (a.iters) a.iters[step * NL + li] = it < 20 ? it + 1 : 20;
    // ---- recording (:164-219) + state update (:221-237), per-type loops in table order ----
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nn) a.v[(step * nn + i) * NL + li] = x[i];
    double* io = a.ielem ? a.ielem + step * ne * NL + li : nullptr;
    for (int e = oR; e < oC; ++e) {
      const int4 q = ends[e];
      const double d = xs[q.x * NT] - xs[q.y * NT];
      const double cur = STRICT ? __ddiv_rn(d, ec[(4 * e + 1) * NT]) : d * ec[(4 * e) * NT];
      if (io) io[(long long)e * NL] = cur;
    }
    for (int e = oC; e < oL; ++e) {
      const int4 q = ends[e];
      const double d = xs[q.x * NT] - xs[q.y * NT];
      const double dv = d - st[sidx[e] * NT];
      const double cur = STRICT ? __ddiv_rn(__dmul_rn(ec[(4 * e + 1) * NT], dv), dtc) : ec[(4 * e) * NT] * dv;
      st[sidx[e] * NT] = d;
      if (io) io[(long long)e * NL] = cur;
    }
    for (int e = oL; e < oV; ++e) {
      const int4 q = ends[e];
      const double d = xs[q.x * NT] - xs[q.y * NT];
      const double cur = t_add<STRICT>(t_mul<STRICT>(ec[(4 * e) * NT], d), st[sidx[e] * NT]);
      st[sidx[e] * NT] = cur;
      if (io) io[(long long)e * NL] = cur;
    }
    for (int e = oV; e < oS; ++e)
      if (io) io[(long long)e * NL] = xs[(nn + e - oV) * NT];
    for (int e = oS; e < oD; ++e) {
      const int4 q = ends[e];
      const double d = xs[q.x * NT] - xs[q.y * NT];
      if (io) io[(long long)e * NL] = d / (st[sidx[e] * NT] != 0.0 ? ec[(4 * e) * NT] : ec[(4 * e + 1) * NT]);
    }
    for (int e = oD; e < oE; ++e) {
      const int4 q = ends[e];
      const double d = xs[q.x * NT] - xs[q.y * NT];
      if (io) io[(long long)e * NL] = t_mul<STRICT>(ec[(4 * e) * NT], t_sub<STRICT>(exp(d / ec[(4 * e + 1) * NT]), 1.0));
      st[sidx[e] * NT] = d;
    }
  }
#undef STAMP_Y
  if (status != ST_OK) {
    for (; step < S1; ++step) {
      for (int i = 0; i < nn; ++i) a.v[(step * nn + i) * NL + li] = CUDART_NAN;
      if (a.ielem) for (int e = 0; e < ne; ++e) a.ielem[(step * ne + e) * NL + li] = CUDART_NAN;
      if (a.iters) a.iters[step * NL + li] = 0;
    }
  }
  if (a.state_out) for (int s = 0; s < ns; ++s) a.state_out[(long long)s * NL + li] = st[s * NT];
  a.status[li] = status;
}

}  // namespace spicey
