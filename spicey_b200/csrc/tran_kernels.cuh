// Persistent transient kernels: the whole fixed-step backward-Euler loop of
// simulateTRAN (lib/analysis/simulateTRAN.ts:146-238) runs on the device, one launch
// per batch, no per-step launches or host round-trips.
//
//   stampAllElementsAtTime  :25-102  -> per-element companion values + gather stamping
//   solveReal               :157     -> per-thread LU (thread tier) / lu_solve_rowthread<double>
//   updateSwitchStates...   :108-128 -> per-instance toggle flag = the re-solve mask
//   recording               :164-219 -> v[step][node][inst], ielem[step][elem][inst] (coalesced)
//   state update            :221-237 -> registers / shared memory, written back once at the end
//
// Iteration policy is the reference's (hazards H3, H4): x is zeroed every step, another
// solve happens only while a switch toggled (at most 20), a diode is linearised about
// vdPrev at iteration 0 and about the previous iterate afterwards.
//
// Two tiers:
//  * thread tier (small Nvar): one thread per instance, the instance's matrix in shared
//    memory laid out [entry][thread] (conflict-free), instance values and state read with
//    coalesced loads from [slot][inst] arrays; lanes of a warp are 32 Monte-Carlo
//    instances, per-instance convergence = per-lane loop exit.
//  * CTA tier (large Nvar): one CTA per instance, one thread per matrix row.
#pragma once
#include "lu_rowthread.cuh"

namespace spicey {

struct TranArgs {
  double dt;
  long long steps;
  const double* vsrc;       // [nV][steps+1] or null: rows of the sources of kind WAVE_TABLE
  const int4* waves;        // [nV] {kind, first value slot of the parameters, PWL pair count, 0}
  unsigned vmask_bits;      // sources of kind WAVE_TABLE as bits (first 32 sources), for the compiled kernel
  unsigned wmask_bits;      // sources evaluated on the device (WAVE_PULSE / WAVE_PWL) as bits
  const double* state0;     // [n_state][n_inst] or null
  long long inst0;          // global index of local instance 0 (for sweep values / state0)
  long long n_local;        // instances handled by this launch
  double* v;                // [steps+1][nn][n_local]
  double* ielem;            // [steps+1][n_elem][n_local] or null
  double* state_out;        // [n_state][n_local] or null
  int* iters;               // [steps+1][n_local] or null
  int* status;              // [n_local]
};

// Source waveforms (simulateTRAN.ts:66-69: V = waveform ? waveform(t) : dc).
enum { WAVE_DC = 0, WAVE_TABLE = 1, WAVE_PULSE = 2, WAVE_PWL = 3 };

// pulseValue (lib/parsing/pulseValue.ts:4-22) for one instance: the eight parameters v1, v2, td, tr, tf, ton,
// period, ncycles are value slots like any element value, so a sweep can vary them per instance.  Every
// operation is a separately rounded IEEE operation (no FMA contraction): the branch conditions compare
// times that differ in the last bit at a pulse edge, and the host's pre-sampled table is what the result
// must equal bit for bit.  period == 0 and ncycles == Infinity behave as in JavaScript (IEEE division).
__device__ __forceinline__ double pulse_value(double v1, double v2, double td, double tr, double tf, double ton,
                                              double period, double ncycles, double t) {
  if (t < td) return v1;
  const double tt = __dsub_rn(t, td);
  const double cycles = floor(__ddiv_rn(tt, period));
  if (cycles >= ncycles) return v1;
  const double tc = __dsub_rn(tt, __dmul_rn(cycles, period));
  if (tc < tr) return __dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), __ddiv_rn(tc, fmax(tr, kEps))));
  const double t1 = __dadd_rn(tr, ton);
  if (tc < t1) return v2;
  if (tc < __dadd_rn(t1, tf)) return __dadd_rn(v2, __dmul_rn(__dsub_rn(v1, v2), __ddiv_rn(__dsub_rn(tc, t1), fmax(tf, kEps))));
  return v1;
}

// One segment of pwlValue (lib/parsing/pwlValue.ts:9-13).
__device__ __forceinline__ double pwl_segment(double pt, double pv, double ct, double cv, double t) {
  const double a = __ddiv_rn(__dsub_rn(t, pt), fmax(__dsub_rn(ct, pt), kEps));
  return __dadd_rn(pv, __dmul_rn(__dsub_rn(cv, pv), a));
}

// Value of a device-evaluated source at t = step * dt (simulateTRAN.ts:147) for instance `inst`.
__device__ __noinline__ double wave_value(const DevPlan& P, int4 w, long long inst, double t) {
  const int s = w.y;
  if (w.x == WAVE_PULSE)
    return pulse_value(inst_value(P, s, inst), inst_value(P, s + 1, inst), inst_value(P, s + 2, inst),
                       inst_value(P, s + 3, inst), inst_value(P, s + 4, inst), inst_value(P, s + 5, inst),
                       inst_value(P, s + 6, inst), inst_value(P, s + 7, inst), t);
  // pwlValue.ts:3-16
  if (w.z <= 0) return 0.0;
  double pt = inst_value(P, s, inst), pv = inst_value(P, s + 1, inst);
  if (t <= pt) return pv;
  for (int i = 1; i < w.z; ++i) {
    const double ct = inst_value(P, s + 2 * i, inst), cv = inst_value(P, s + 2 * i + 1, inst);
    if (t <= ct) return pwl_segment(pt, pv, ct, cv, t);
    pt = ct; pv = cv;
  }
  return pv;
}

// The k-th V element's value at this step: dc, its pre-sampled row, or the waveform evaluated here.
__device__ __forceinline__ double source_value(const DevPlan& P, const TranArgs& a, int k, long long step,
                                               long long inst, double dc) {
  const int4 w = a.waves[k];
  if (w.x == WAVE_DC) return dc;
  if (w.x == WAVE_TABLE) return a.vsrc[(long long)k * (a.steps + 1) + step];
  return wave_value(P, w, inst, __dmul_rn((double)step, a.dt));
}

template <bool STRICT> __device__ __forceinline__ double t_mul(double a, double b) {
  return STRICT ? __dmul_rn(a, b) : a * b;
}
template <bool STRICT> __device__ __forceinline__ double t_sub(double a, double b) {
  return STRICT ? __dsub_rn(a, b) : a - b;
}
template <bool STRICT> __device__ __forceinline__ double t_add(double a, double b) {
  return STRICT ? __dadd_rn(a, b) : a + b;
}

// Diode companion (simulateTRAN.ts:87-97): clamp, exp, gd floor, ieq.  is_over_vth = Is / vth is the
// reference's own sub-expression, computed once per instance; the fast path multiplies by 1/vth.
template <bool STRICT>
__device__ __forceinline__ void diode_companion(double vd, double Is, double vth, double is_over_vth,
                                                double inv_vth, double& gd, double& ieq) {
  double vlim = vd;
  if (vd > 0.8) vlim = 0.8;
  if (vd < -1.0) vlim = -1.0;
  double e = exp(STRICT ? __ddiv_rn(vlim, vth) : vlim * inv_vth);
  double id = t_mul<STRICT>(Is, t_sub<STRICT>(e, 1.0));
  gd = fmax(t_mul<STRICT>(is_over_vth, e), 1e-12);
  ieq = t_sub<STRICT>(id, t_mul<STRICT>(gd, vlim));
}

// Per-element constants, 4 doubles per element (derived once per instance):
//  R: G=1/R, R        C: Gc=C/dtc, C       L: Gl=dtc/L       V: dc
//  S: Ron', Roff' (clamped), Von, Voff     D: Is, vth=N*VT, Is/vth, 1/vth
__device__ __forceinline__ void element_constants(const DevPlan& P, int type, int vidx, long long inst,
                                                  double dtc, double* c4) {
  c4[0] = c4[1] = c4[2] = c4[3] = 0.0;
  if (type == ELEM_R) {
    double R = inst_value(P, vidx, inst);
    c4[0] = 1 / R; c4[1] = R;
  } else if (type == ELEM_C) {
    double C = inst_value(P, vidx, inst);
    c4[0] = C / dtc; c4[1] = C;
  } else if (type == ELEM_L) {
    c4[0] = dtc / inst_value(P, vidx, inst);
  } else if (type == ELEM_V || type == ELEM_I) {   // dc value (I: constant current, stampCurrentReal.ts:3-14)
    c4[0] = inst_value(P, vidx, inst);
  } else if (type == ELEM_S) {
    c4[0] = fmax(fabs(inst_value(P, vidx, inst)), kEps);          // Rclamped :60-61
    c4[1] = fmax(fabs(inst_value(P, vidx + 1, inst)), kEps);
    c4[2] = inst_value(P, vidx + 2, inst);
    c4[3] = inst_value(P, vidx + 3, inst);
  } else {  // ELEM_D
    double Is = inst_value(P, vidx, inst), N = inst_value(P, vidx + 1, inst);
    c4[0] = Is; c4[1] = N * kVt300; c4[2] = Is / c4[1]; c4[3] = 1.0 / c4[1];
  }
}

// ------------------------------------------------------------------------------------
// Thread tier.  Per-thread arrays live in shared memory with stride NT = blockDim.x:
//   A[nvar*(nvar+1)], x[nvar], st[n_state], ec[4*n_elem], g[n_elem], jj[n_elem]
// Block-shared: element table and the gather plan.
struct TranThreadSmem {
  size_t per_thread_doubles, shared_off, ends_off, meta_off, sidx_off, rowptr_off, entcol_off, entptr_off,
      contrib_off, total;
  __host__ __device__ TranThreadSmem(int nvar, int n_elem, int n_state, int n_ent, int n_con, int nt) {
    per_thread_doubles = (size_t)nvar * (nvar + 1) + nvar + n_state + 6 * (size_t)n_elem;
    size_t o = sizeof(double) * per_thread_doubles * nt;
    o = (o + 15) & ~(size_t)15;
    shared_off = o;
    ends_off = o; o += sizeof(int4) * n_elem;
    meta_off = o; o += sizeof(int2) * n_elem;
    sidx_off = o; o += sizeof(int) * n_elem;
    rowptr_off = o; o += sizeof(int) * (nvar + 1);
    entcol_off = o; o += sizeof(int) * n_ent;
    entptr_off = o; o += sizeof(int) * (n_ent + 1);
    contrib_off = o; o += sizeof(int) * n_con;
    total = (o + 15) & ~(size_t)15;
  }
};

// solveReal.ts:3-73 for one thread; M(r,c) at M[(r*(n+1)+c)*S]; x written to x[i*S].
template <bool STRICT>
__device__ __forceinline__ int solve_real_thread(double* M, int n, int S, double* x) {
  const int ld = n + 1;
  for (int k = 0; k < n; ++k) {
    int imax = k;
    double vmax = fabs(M[(k * ld + k) * S]);
    for (int i = k + 1; i < n; ++i) {
      double v = fabs(M[(i * ld + k) * S]);
      if (v > vmax) { vmax = v; imax = i; }
    }
    if (vmax < kEps) return ST_SINGULAR;
    if (imax != k)
      for (int j = k; j <= n; ++j) {
        double a = M[(k * ld + j) * S], b = M[(imax * ld + j) * S];
        M[(k * ld + j) * S] = b;
        M[(imax * ld + j) * S] = a;
      }
    const double pivot = M[(k * ld + k) * S];
    const double rp = 1.0 / pivot;
    for (int i = k + 1; i < n; ++i) {
      double aik = M[(i * ld + k) * S];
      double f = STRICT ? __ddiv_rn(aik, pivot) : aik * rp;
      if (fabs(f) < kEps) continue;
      for (int j = k + 1; j <= n; ++j)
        M[(i * ld + j) * S] = Num<double>::submul<STRICT>(M[(i * ld + j) * S], f, M[(k * ld + j) * S]);
    }
    if (!STRICT) M[(k * ld + k) * S] = rp;  // keep the reciprocal for the back-substitution
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = M[(i * ld + n) * S];
    for (int j = i + 1; j < n; ++j) s = Num<double>::submul<STRICT>(s, M[(i * ld + j) * S], x[j * S]);
    x[i * S] = STRICT ? __ddiv_rn(s, M[(i * ld + i) * S]) : s * M[(i * ld + i) * S];
  }
  return ST_OK;
}

template <bool STRICT>
__global__ void tran_thread_kernel(DevPlan P, TranArgs a, int n_ent, int n_con) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int t = threadIdx.x, NT = blockDim.x;
  const int nvar = P.nvar, nn = P.nn, ne = P.n_elem, ns = P.n_state;
  const TranThreadSmem L(nvar, ne, ns, n_ent, n_con, NT);
  double* base = (double*)smem + t;
  double* A = base;
  double* x = A + (size_t)nvar * (nvar + 1) * NT;
  double* st = x + (size_t)nvar * NT;
  double* ec = st + (size_t)ns * NT;
  double* g = ec + (size_t)4 * ne * NT;
  double* jj = g + (size_t)ne * NT;
  int4* ends = (int4*)(smem + L.ends_off);
  int2* meta = (int2*)(smem + L.meta_off);
  int* sidx = (int*)(smem + L.sidx_off);
  int* row_ptr = (int*)(smem + L.rowptr_off);
  int* ent_col = (int*)(smem + L.entcol_off);
  int* ent_ptr = (int*)(smem + L.entptr_off);
  int* contrib = (int*)(smem + L.contrib_off);
  const GatherPlan& G = P.tran;
  for (int e = t; e < ne; e += NT) { ends[e] = P.ends[e]; meta[e] = P.meta[e]; sidx[e] = P.state_idx[e]; }
  for (int i = t; i <= nvar; i += NT) row_ptr[i] = G.row_ptr[i];
  for (int i = t; i < n_ent; i += NT) ent_col[i] = G.ent_col[i];
  for (int i = t; i <= n_ent; i += NT) ent_ptr[i] = G.ent_ptr[i];
  for (int i = t; i < n_con; i += NT) contrib[i] = G.contrib[i];
  __syncthreads();

  const long long li = (long long)blockIdx.x * NT + t;
  if (li >= a.n_local) return;
  const long long inst = a.inst0 + li;
  const long long NL = a.n_local;
  const long long S1 = a.steps + 1;
  const double dtc = fmax(a.dt, kEps);
  const int ld = nvar + 1;

  for (int e = 0; e < ne; ++e) {
    double c4[4];
    element_constants(P, meta[e].x, meta[e].y, inst, dtc, c4);
    for (int q = 0; q < 4; ++q) ec[(4 * e + q) * NT] = c4[q];
    g[e * NT] = c4[0];
    jj[e * NT] = meta[e].x == ELEM_I ? c4[0] : 0.0;   // a current source's right-hand-side term never changes
  }
  for (int s = 0; s < ns; ++s) st[s * NT] = a.state0 ? a.state0[(long long)s * P.n_inst + inst] : 0.0;

#define VOLT(n) ((n) == 0 ? 0.0 : x[((n) - 1) * NT])
  int status = ST_OK;
  long long step = 0;
  for (; step < S1; ++step) {
    for (int i = 0; i < nvar; ++i) x[i * NT] = 0.0;                       // :149
    for (int e = P.off[ELEM_V]; e < P.off[ELEM_V + 1]; ++e) {             // :66-69
      int k = e - P.off[ELEM_V];
      jj[e * NT] = source_value(P, a, k, step, inst, ec[(4 * e) * NT]);
    }
    int it = 0;
    for (; it < 20; ++it) {                                               // :151
      // companion values (:35-101)
      for (int e = P.off[ELEM_C]; e < P.off[ELEM_C + 1]; ++e)
        jj[e * NT] = t_mul<STRICT>(-ec[(4 * e) * NT], st[sidx[e] * NT]);  // Ieq = -Gc*vPrev
      for (int e = P.off[ELEM_L]; e < P.off[ELEM_L + 1]; ++e) jj[e * NT] = st[sidx[e] * NT];
      for (int e = P.off[ELEM_S]; e < P.off[ELEM_S + 1]; ++e)
        g[e * NT] = 1 / (st[sidx[e] * NT] != 0.0 ? ec[(4 * e) * NT] : ec[(4 * e + 1) * NT]);  // :62
      for (int e = P.off[ELEM_D]; e < P.off[ELEM_D + 1]; ++e) {
        int4 en = ends[e];
        double vd = it == 0 ? st[sidx[e] * NT] : VOLT(en.x) - VOLT(en.y);  // :85
        double gd, ieq;
        diode_companion<STRICT>(vd, ec[(4 * e) * NT], ec[(4 * e + 1) * NT], ec[(4 * e + 2) * NT], ec[(4 * e + 3) * NT], gd, ieq);
        g[e * NT] = gd;
        jj[e * NT] = ieq;
      }
      // gather-stamp (zero + ordered sums)
      for (int i = 0; i < nvar * ld; ++i) A[i * NT] = 0.0;
      for (int r = 0; r < nvar; ++r)
        for (int en = row_ptr[r]; en < row_ptr[r + 1]; ++en) {
          double acc = 0.0;
          for (int c = ent_ptr[en]; c < ent_ptr[en + 1]; ++c) {
            int w = contrib[c];
            int src = (w >> 1) & 3, idx = w >> 3;
            double v = src == SRC_Y ? g[idx * NT] : (src == SRC_J ? jj[idx * NT] : 1.0);
            acc = (w & 1) ? acc - v : acc + v;
          }
          A[(r * ld + ent_col[en]) * NT] = acc;
        }
      status = solve_real_thread<STRICT>(A, nvar, NT, x);
      if (status != ST_OK) break;
      bool switched = false;                                              // :108-128
      for (int e = P.off[ELEM_S]; e < P.off[ELEM_S + 1]; ++e) {
        int4 en = ends[e];
        double vctrl = VOLT(en.z) - VOLT(en.w);
        bool on = st[sidx[e] * NT] != 0.0, nxt = on;
        if (on) { if (vctrl < ec[(4 * e + 3) * NT]) nxt = false; }
        else if (vctrl > ec[(4 * e + 2) * NT]) nxt = true;
        if (nxt != on) { st[sidx[e] * NT] = nxt ? 1.0 : 0.0; switched = true; }
      }
      if (!switched) break;
    }
    if (status != ST_OK) break;
    if (a.iters) a.iters[step * NL + li] = it < 20 ? it + 1 : 20;
    // recording (:164-219) and state update (:221-237)
    for (int i = 0; i < nn; ++i) a.v[(step * nn + i) * NL + li] = x[i * NT];
    for (int e = 0; e < ne; ++e) {
      int4 en = ends[e];
      int type = meta[e].x;
      double d = VOLT(en.x) - VOLT(en.y);
      double cur;
      if (type == ELEM_R) {
        cur = STRICT ? __ddiv_rn(d, ec[(4 * e + 1) * NT]) : d * ec[(4 * e) * NT];
      } else if (type == ELEM_C) {
        double dv = d - st[sidx[e] * NT];
        cur = STRICT ? __ddiv_rn(__dmul_rn(ec[(4 * e + 1) * NT], dv), dtc) : ec[(4 * e) * NT] * dv;
        st[sidx[e] * NT] = d;
      } else if (type == ELEM_L) {
        cur = t_add<STRICT>(t_mul<STRICT>(ec[(4 * e) * NT], d), st[sidx[e] * NT]);
        st[sidx[e] * NT] = cur;
      } else if (type == ELEM_V) {
        cur = x[(nn + e - P.off[ELEM_V]) * NT];
      } else if (type == ELEM_S) {
        cur = d / (st[sidx[e] * NT] != 0.0 ? ec[(4 * e) * NT] : ec[(4 * e + 1) * NT]);  // post-toggle state :196-204
      } else if (type == ELEM_I) {
        cur = ec[(4 * e) * NT];
      } else {
        cur = t_mul<STRICT>(ec[(4 * e) * NT], t_sub<STRICT>(exp(d / ec[(4 * e + 1) * NT]), 1.0));  // unclamped vd (H6)
        st[sidx[e] * NT] = d;
      }
      if (a.ielem) a.ielem[(step * ne + e) * NL + li] = cur;
    }
  }
#undef VOLT
  if (status != ST_OK) {  // the reference would have thrown here: poison the rest of this instance
    for (; step < S1; ++step) {
      for (int i = 0; i < nn; ++i) a.v[(step * nn + i) * NL + li] = CUDART_NAN;
      if (a.ielem) for (int e = 0; e < ne; ++e) a.ielem[(step * ne + e) * NL + li] = CUDART_NAN;
      if (a.iters) a.iters[step * NL + li] = 0;
    }
  }
  if (a.state_out) for (int s = 0; s < ns; ++s) a.state_out[(long long)s * NL + li] = st[s * NT];
  a.status[li] = status;
}

// ------------------------------------------------------------------------------------
// CTA tier: one CTA per instance, one thread per matrix row (lu_solve_rowthread<double>).
struct TranCtaSmem {
  size_t a_off, xs_off, st_off, ec_off, g_off, j_off, red_off, ends_off, meta_off, sidx_off, mask_off, total;
  __host__ __device__ TranCtaSmem(int nvar, int n_elem, int n_state, int MW, int nwarps, bool gmem) {
    size_t o = 0;
    a_off = o; o += gmem ? 0 : sizeof(double) * (size_t)nvar * (nvar + 1);
    xs_off = o; o += sizeof(double) * nvar;
    st_off = o; o += sizeof(double) * (n_state + 1);
    ec_off = o; o += sizeof(double) * 4 * n_elem;
    g_off = o; o += sizeof(double) * n_elem;
    j_off = o; o += sizeof(double) * n_elem;
    o = (o + 31) & ~(size_t)31;
    red_off = o; o += sizeof(PivotPartial) * 2 * nwarps;
    ends_off = o; o += sizeof(int4) * n_elem;
    meta_off = o; o += sizeof(int2) * n_elem;
    sidx_off = o; o += sizeof(int) * n_elem;
    mask_off = o; o += gmem ? 0 : sizeof(unsigned) * (size_t)nvar * MW;
    total = (o + 15) & ~(size_t)15;
  }
  static __host__ __device__ size_t scratch_bytes(int nvar, int MW) {
    return ((sizeof(double) * (size_t)nvar * (nvar + 1) + sizeof(unsigned) * (size_t)nvar * MW) + 15) & ~(size_t)15;
  }
};

template <bool STRICT, bool GMEM>
__global__ void tran_cta_kernel(DevPlan P, TranArgs a, double* scratch) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int t = threadIdx.x, NT = blockDim.x;
  const int nvar = P.nvar, nn = P.nn, ne = P.n_elem, ns = P.n_state, MW = P.MW;
  const int nwarps = (NT + 31) >> 5;
  const TranCtaSmem L(nvar, ne, ns, MW, nwarps, GMEM);
  unsigned char* gscr = GMEM ? (unsigned char*)scratch + (size_t)blockIdx.x * TranCtaSmem::scratch_bytes(nvar, MW) : nullptr;
  double* A = GMEM ? (double*)gscr : (double*)(smem + L.a_off);
  double* xs = (double*)(smem + L.xs_off);
  double* st = (double*)(smem + L.st_off);
  double* ec = (double*)(smem + L.ec_off);
  double* g = (double*)(smem + L.g_off);
  double* jj = (double*)(smem + L.j_off);
  PivotPartial* red = (PivotPartial*)(smem + L.red_off);
  int4* ends = (int4*)(smem + L.ends_off);
  int2* meta = (int2*)(smem + L.meta_off);
  int* sidx = (int*)(smem + L.sidx_off);
  unsigned* mask = GMEM ? (unsigned*)(gscr + sizeof(double) * (size_t)nvar * (nvar + 1)) : (unsigned*)(smem + L.mask_off);
  const GatherPlan& G = P.tran;
  const int ldr = nvar;
  const long long S1 = a.steps + 1, NL = a.n_local;
  const double dtc = fmax(a.dt, kEps);
  for (int e = t; e < ne; e += NT) { ends[e] = P.ends[e]; meta[e] = P.meta[e]; sidx[e] = P.state_idx[e]; }
  __syncthreads();

#define VOLT(n) ((n) == 0 ? 0.0 : xs[(n) - 1])
  for (long long li = blockIdx.x; li < NL; li += gridDim.x) {
    const long long inst = a.inst0 + li;
    for (int e = t; e < ne; e += NT) {
      double c4[4];
      element_constants(P, meta[e].x, meta[e].y, inst, dtc, c4);
      for (int q = 0; q < 4; ++q) ec[4 * e + q] = c4[q];
      g[e] = c4[0];
      jj[e] = meta[e].x == ELEM_I ? c4[0] : 0.0;
    }
    for (int s = t; s < ns; s += NT) st[s] = a.state0 ? a.state0[(long long)s * P.n_inst + inst] : 0.0;
    __syncthreads();
    int status = ST_OK;
    long long step = 0;
    for (; step < S1; ++step) {
      if (t < nvar) xs[t] = 0.0;                                          // :149
      int it = 0;
      for (; it < 20; ++it) {
        __syncthreads();  // xs / st of the previous iteration are complete
        for (int e = t; e < ne; e += NT) {
          int type = meta[e].x;
          if (type == ELEM_C) jj[e] = t_mul<STRICT>(-ec[4 * e], st[sidx[e]]);
          else if (type == ELEM_L) jj[e] = st[sidx[e]];
          else if (type == ELEM_S) g[e] = 1 / (st[sidx[e]] != 0.0 ? ec[4 * e] : ec[4 * e + 1]);
          else if (type == ELEM_V) {
            int k = e - P.off[ELEM_V];
            jj[e] = source_value(P, a, k, step, inst, ec[4 * e]);
          } else if (type == ELEM_D) {
            int4 en = ends[e];
            double vd = it == 0 ? st[sidx[e]] : VOLT(en.x) - VOLT(en.y);
            double gd, ieq;
            diode_companion<STRICT>(vd, ec[4 * e], ec[4 * e + 1], ec[4 * e + 2], ec[4 * e + 3], gd, ieq);
            g[e] = gd;
            jj[e] = ieq;
          }
        }
        if (t < nvar) {
          for (int j = 0; j <= nvar; ++j) A[(size_t)j * ldr + t] = 0.0;
          for (int w = 0; w < MW; ++w) mask[t * MW + w] = G.rowmask[t * MW + w];
        }
        __syncthreads();
        if (t < nvar)
          for (int en = G.row_ptr[t]; en < G.row_ptr[t + 1]; ++en) {
            double acc = 0.0;
            for (int c = G.ent_ptr[en]; c < G.ent_ptr[en + 1]; ++c) {
              int w = G.contrib[c];
              int src = (w >> 1) & 3, idx = w >> 3;
              double v = src == SRC_Y ? g[idx] : (src == SRC_J ? jj[idx] : 1.0);
              acc = (w & 1) ? acc - v : acc + v;
            }
            A[(size_t)G.ent_col[en] * ldr + t] = acc;
          }
        __syncthreads();
        status = lu_solve_rowthread<double, STRICT>(A, ldr, nvar, mask, MW, xs, red);
        if (status != ST_OK) break;
        int switched = 0;
        for (int e = P.off[ELEM_S] + t; e < P.off[ELEM_S + 1]; e += NT) {
          int4 en = ends[e];
          double vctrl = VOLT(en.z) - VOLT(en.w);
          bool on = st[sidx[e]] != 0.0, nxt = on;
          if (on) { if (vctrl < ec[4 * e + 3]) nxt = false; }
          else if (vctrl > ec[4 * e + 2]) nxt = true;
          if (nxt != on) { st[sidx[e]] = nxt ? 1.0 : 0.0; switched = 1; }
        }
        if (!__syncthreads_or(switched)) break;
      }
      if (status != ST_OK) break;
      if (t == 0 && a.iters) a.iters[step * NL + li] = it < 20 ? it + 1 : 20;
      if (t < nn) a.v[(step * nn + t) * NL + li] = xs[t];
      for (int e = t; e < ne; e += NT) {
        int4 en = ends[e];
        int type = meta[e].x;
        double d = VOLT(en.x) - VOLT(en.y);
        double cur;
        if (type == ELEM_R) cur = STRICT ? __ddiv_rn(d, ec[4 * e + 1]) : d * ec[4 * e];
        else if (type == ELEM_C) {
          double dv = d - st[sidx[e]];
          cur = STRICT ? __ddiv_rn(__dmul_rn(ec[4 * e + 1], dv), dtc) : ec[4 * e] * dv;
          st[sidx[e]] = d;
        } else if (type == ELEM_L) {
          cur = t_add<STRICT>(t_mul<STRICT>(ec[4 * e], d), st[sidx[e]]);
          st[sidx[e]] = cur;
        } else if (type == ELEM_V) cur = xs[nn + e - P.off[ELEM_V]];
        else if (type == ELEM_S) {
          cur = d / (st[sidx[e]] != 0.0 ? ec[4 * e] : ec[4 * e + 1]);
        } else if (type == ELEM_I) cur = ec[4 * e];
        else {
          cur = t_mul<STRICT>(ec[4 * e], t_sub<STRICT>(exp(d / ec[4 * e + 1]), 1.0));
          st[sidx[e]] = d;
        }
        if (a.ielem) a.ielem[(step * ne + e) * NL + li] = cur;
      }
      __syncthreads();  // state updates and xs reads done before the next step zeroes xs
    }
    if (status != ST_OK) {
      for (; step < S1; ++step) {
        if (t < nn) a.v[(step * nn + t) * NL + li] = CUDART_NAN;
        if (a.ielem) for (int e = t; e < ne; e += NT) a.ielem[(step * ne + e) * NL + li] = CUDART_NAN;
        if (t == 0 && a.iters) a.iters[step * NL + li] = 0;
      }
    }
    __syncthreads();
    if (a.state_out) for (int s = t; s < ns; s += NT) a.state_out[(long long)s * NL + li] = st[s];
    if (t == 0) a.status[li] = status;
    __syncthreads();
  }
#undef VOLT
}

}  // namespace spicey
