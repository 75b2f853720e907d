// AC small-signal batch kernel: stamp + complex LU + unpack, fused, one CTA per point.
//
// Replaces the per-frequency body of simulateAC (lib/analysis/simulateAC.ts:80-127):
//   buildLinearSystemForAC :24-60  -> element admittances + gather stamping (shared memory)
//   solveComplex           :83     -> lu_solve_rowthread<cplx>
//   unpack                 :85-126 -> x and element currents written straight to HBM
// The matrix never exists in HBM (except in the global-scratch tier for systems that do
// not fit one SM): it is generated from the element table and the frequency inside the
// CTA.  Algorithmic HBM traffic per point is 8 B in, 16*(Nvar + nAc) + 4 B out.
//
// Batch axis: point p = inst * n_freq + k.  Persistent CTAs stride over the points.
#pragma once
#include "lu_rowthread.cuh"

namespace spicey {

struct AcArgs {
  const double* freqs;  // [n_freq]
  long long n_freq;
  long long p_begin;    // first point handled by this launch (index into the full batch)
  long long p_count;    // points in this launch
  double2* x;           // [p_count][nvar]   (indexed from p_begin)
  double2* ielem;       // [p_count][n_ac_elem] or null
  int* status;          // [p_count]
  double2* scratch;     // global-scratch tier: gridDim.x * nvar*(nvar+1)
  const long long* plist;  // optional: launch-local point indices to solve (fallback of the sparse path)
  const int* pcount;       // optional: number of entries of plist (device-resident)
  unsigned long long* fb_total;  // optional: [0] running total of this call's fallback solves (statistics), [1] since the program was built
  long long series_ld;     // 0: x[q][var], ielem[q][e] (point-major); else x[var][series_ld], ielem[e][series_ld]
};

// Shared-memory carve-up, identical on host (sizing) and device.
struct AcSmem {
  size_t a_off, y_off, j_off, xs_off, mask_off, red_off, ends_off, meta_off, total;
  __host__ __device__ AcSmem(int nvar, int n_elem, int nV, int MW, int nwarps, bool gmem) {
    size_t o = 0;
    a_off = o; o += gmem ? 0 : sizeof(double2) * (size_t)nvar * (nvar + 1);
    y_off = o; o += sizeof(double2) * n_elem;
    j_off = o; o += sizeof(double2) * (nV > 0 ? nV : 1);
    xs_off = o; o += sizeof(double2) * nvar;
    red_off = o; o += sizeof(PivotPartial) * 2 * nwarps;
    ends_off = o; o += sizeof(int4) * n_elem;
    meta_off = o; o += sizeof(int2) * n_elem;
    mask_off = o; o += gmem ? 0 : sizeof(unsigned) * (size_t)nvar * MW;  // global tier: masks follow the matrix in the scratch
    total = (o + 15) & ~(size_t)15;
  }
  // bytes of global scratch per CTA in the global tier: matrix + row masks
  static __host__ __device__ size_t scratch_bytes(int nvar, int MW) {
    return ((sizeof(double2) * (size_t)nvar * (nvar + 1) + sizeof(unsigned) * (size_t)nvar * MW) + 15) & ~(size_t)15;
  }
};

// Element admittance at frequency f (simulateAC.ts:36-52) and source phasor (:54-57).
template <bool STRICT>
__device__ __forceinline__ int ac_element_values(const DevPlan& P, int type, int vidx, long long inst,
                                                 double f, cplx& Y, cplx& J) {
  const double twoPi = 2 * kPi;
  Y = make_double2(0.0, 0.0);
  J = make_double2(0.0, 0.0);
  if (type == ELEM_R) {
    double R = inst_value(P, vidx, inst);
    if (R <= 0) return ST_R_NONPOS;                       // :37
    Y.x = 1 / R;
  } else if (type == ELEM_C) {
    Y.y = __dmul_rn(__dmul_rn(twoPi, f), inst_value(P, vidx, inst));   // twoPi * f * c.C  :43
  } else if (type == ELEM_L) {
    double d = __dmul_rn(__dmul_rn(twoPi, f), inst_value(P, vidx, inst));
    if (fabs(d) < kEps) return ST_OK;                     // denom.abs() < EPS -> Y = 0  :49
    double dd = __dmul_rn(d, d);
    if (dd < kEps) return ST_CDIV;                        // Complex.div guard (H5)
    Y.x = 0.0 / dd;                                       // (1*0 + 0*d)/dd
    Y.y = (0.0 - d) / dd;                                 // (0*0 - 1*d)/dd
  } else if (type == ELEM_V || type == ELEM_I) {   // source phasor: V rows' rhs, I elements' KCL terms
    double mag = inst_value(P, vidx + 1, inst), deg = inst_value(P, vidx + 2, inst);
    double ph = (deg * kPi) / 180;                        // Complex.ts:16-19
    double s, c;
    sincos(ph, &s, &c);
    J.x = mag * c;
    J.y = mag * s;
  }
  return ST_OK;
}

// Slot of a source element's phasor: V elements first, then I elements.
__device__ __forceinline__ int ac_source_slot(const DevPlan& P, int e) {
  return e < P.off[ELEM_V + 1] ? e - P.off[ELEM_V] : P.nV + (e - P.off[ELEM_I]);
}

template <bool STRICT, bool GMEM>
__global__ void ac_cta_kernel(DevPlan P, AcArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int t = threadIdx.x;
  const int nvar = P.nvar, ne = P.n_elem, MW = P.MW;
  const int nwarps = (blockDim.x + 31) >> 5;
  const AcSmem L(nvar, ne, P.nV + (P.off[ELEM_I + 1] - P.off[ELEM_I]), MW, nwarps, GMEM);
  unsigned char* gscr = GMEM ? (unsigned char*)a.scratch + (size_t)blockIdx.x * AcSmem::scratch_bytes(nvar, MW) : nullptr;
  cplx* A = GMEM ? (cplx*)gscr : (cplx*)(smem + L.a_off);
  cplx* Yv = (cplx*)(smem + L.y_off);
  cplx* Jv = (cplx*)(smem + L.j_off);
  cplx* xs = (cplx*)(smem + L.xs_off);
  unsigned* mask = GMEM ? (unsigned*)(gscr + sizeof(double2) * (size_t)nvar * (nvar + 1)) : (unsigned*)(smem + L.mask_off);
  PivotPartial* red = (PivotPartial*)(smem + L.red_off);
  int4* ends = (int4*)(smem + L.ends_off);
  int2* meta = (int2*)(smem + L.meta_off);
  __shared__ int s_status;

  // Element table -> shared memory once per CTA (vectorised, coalesced).
  for (int e = t; e < ne; e += blockDim.x) { ends[e] = P.ends[e]; meta[e] = P.meta[e]; }
  const int ldr = nvar;
  const GatherPlan& G = P.ac;

  const long long work = a.plist ? (long long)*a.pcount : a.p_count;
  if (a.fb_total && blockIdx.x == 0 && t == 0 && work > 0) {   // [0] this call's fallbacks, [1] the program's lifetime total
    atomicAdd(a.fb_total, (unsigned long long)work);
    atomicAdd(a.fb_total + 1, (unsigned long long)work);
  }
  for (long long qi = blockIdx.x; qi < work; qi += gridDim.x) {
    const long long q = a.plist ? a.plist[qi] : qi;
    const long long p = a.p_begin + q;
    const long long inst = p / a.n_freq;
    const double f = a.freqs[p - inst * a.n_freq];
    if (t == 0) s_status = ST_OK;
    __syncthreads();  // also fences the previous point's readers of xs/Yv

    // Phase 1: per-element admittances / phasors.
    for (int e = t; e < ne; e += blockDim.x) {
      cplx Y, J;
      int st = ac_element_values<STRICT>(P, meta[e].x, meta[e].y, inst, f, Y, J);
      Yv[e] = Y;
      if (meta[e].x == ELEM_V || meta[e].x == ELEM_I) Jv[ac_source_slot(P, e)] = J;
      if (st != ST_OK) atomicMax(&s_status, st);  // R<=0 (3) outranks the L divide guard (2): R loop runs first
    }
    // Phase 2a: clear my row and reset its structural mask.
    if (t < nvar) {
      for (int j = 0; j <= nvar; ++j) A[(size_t)j * ldr + t] = make_double2(0.0, 0.0);
      for (int w = 0; w < MW; ++w) mask[t * MW + w] = G.rowmask[t * MW + w];
    }
    __syncthreads();
    int status = s_status;
    if (status == ST_OK) {
      // Phase 2b: gather-stamp my row in the reference's stamping order.
      if (t < nvar) {
        for (int en = G.row_ptr[t]; en < G.row_ptr[t + 1]; ++en) {
          cplx acc = make_double2(0.0, 0.0);
          for (int c = G.ent_ptr[en]; c < G.ent_ptr[en + 1]; ++c) {
            int w = G.contrib[c];
            int src = (w >> 1) & 3, idx = w >> 3;
            cplx v = src == SRC_Y ? Yv[idx] : (src == SRC_J ? Jv[ac_source_slot(P, idx)] : make_double2(1.0, 0.0));
            if (w & 1) { acc.x -= v.x; acc.y -= v.y; } else { acc.x += v.x; acc.y += v.y; }
          }
          A[(size_t)G.ent_col[en] * ldr + t] = acc;
        }
      }
      __syncthreads();
      status = lu_solve_rowthread<cplx, STRICT>(A, ldr, nvar, mask, MW, xs, red);
    }

    // Phase 4: unpack (simulateAC.ts:85-126).
    const long long sld = a.series_ld;
    cplx* xo = sld ? a.x + q : a.x + (size_t)q * nvar;
    const long long xst = sld ? sld : 1;   // stride between consecutive variables / elements of one point
    if (status == ST_OK) {
      if (t < nvar) xo[t * xst] = xs[t];
      if (a.ielem) {
        cplx* io = sld ? a.ielem + q : a.ielem + (size_t)q * P.n_ac_elem;
        for (int e = t; e < P.n_ac_elem; e += blockDim.x) {
          int4 en = ends[e];
          cplx cur;
          if (meta[e].x == ELEM_V) {
            cur = xs[P.nn + (e - P.off[ELEM_V])];
          } else {
            cplx v1 = en.x == 0 ? make_double2(0.0, 0.0) : xs[en.x - 1];
            cplx v2 = en.y == 0 ? make_double2(0.0, 0.0) : xs[en.y - 1];
            cplx d = csub(v1, v2);
            cur = STRICT ? Num<cplx>::mul_strict(Yv[e], d) : Num<cplx>::mul(Yv[e], d);
          }
          io[e * xst] = cur;
        }
      }
    } else {
      const cplx qn = Num<cplx>::nan();
      if (t < nvar) xo[t * xst] = qn;
      if (a.ielem) {
        cplx* io = sld ? a.ielem + q : a.ielem + (size_t)q * P.n_ac_elem;
        for (int e = t; e < P.n_ac_elem; e += blockDim.x) io[e * xst] = qn;
      }
    }
    if (t == 0) a.status[q] = status;
    __syncthreads();  // s_status / xs / Yv are reused by the next point
  }
}

}  // namespace spicey
