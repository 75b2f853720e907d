// Shared device-side definitions of the batched MNA engine (sm_100a).
//
// Data model: one *plan* per netlist topology, built on the host
// from the flat element table and resident in HBM for the call:
//   - the element table as int4 {n1,n2,nc1,nc2} + int2 {type,value_idx} rows, read with
//     vectorised loads and staged in shared memory once per CTA;
//   - a gather-form stamp plan: for every matrix row, its structurally non-zero entries
//     and, per entry, the ordered list of signed contributions.  One thread owns one
//     matrix row and sums its contributions in the reference's stamping order
//     (lib/stamping/*.ts via simulateAC.ts:36-57 / simulateTRAN.ts:35-101), so stamping
//     needs no atomics and reproduces the reference's summation order;
//   - the structural bit mask of every row, which the LU carries along so that the
//     reference's "skip zero multiplier" shortcut (solveComplex.ts:46) extends to
//     structurally zero columns.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace spicey {

constexpr double kEps = 1e-15;      // lib/constants/EPS.ts:1
constexpr double kVt300 = 0.02585;  // lib/constants/physics.ts:1
constexpr double kPi = 3.141592653589793;

enum ElemType { ELEM_R = 0, ELEM_C = 1, ELEM_L = 2, ELEM_V = 3, ELEM_S = 4, ELEM_D = 5, ELEM_I = 6 };
constexpr int kElemKinds = 7;
enum Status { ST_OK = 0, ST_SINGULAR = 1, ST_CDIV = 2, ST_R_NONPOS = 3 };

// Contribution word of the gather plan: idx << 3 | src << 1 | neg.
enum ContribSrc { SRC_Y = 0, SRC_J = 1, SRC_ONE = 2 };

struct GatherPlan {
  const int* row_ptr;        // [nvar+1] -> entries
  const int* ent_col;        // [n_ent]  column (nvar = right-hand side)
  const int* ent_ptr;        // [n_ent+1] -> contributions
  const int* contrib;        // [n_con]
  const unsigned* rowmask;   // [nvar][MW] structural non-zeros incl. bit nvar (rhs)
};

struct DevPlan {
  int nn, nV, nvar, n_elem, n_values, n_ac_elem, n_state;
  int off[8];                // element group offsets R,C,L,V,S,D,I,end
  int MW;                    // mask words per row = ceil((nvar+1)/32)
  int n_var;                 // swept value slots
  long long n_inst;
  const int4* ends;          // [n_elem] n1,n2,nc1,nc2
  const int2* meta;          // [n_elem] type,value_idx
  const int* state_idx;      // [n_elem] slot in the state vector or -1
  const double* values;      // [n_values] nominal
  const int* var_of_slot;    // [n_values] row of var_values or -1
  const double* var_values;  // [n_var][n_inst]
  GatherPlan ac, tran;
};

__device__ __forceinline__ double inst_value(const DevPlan& P, int slot, long long inst) {
  int v = P.var_of_slot[slot];
  return v < 0 ? P.values[slot] : P.var_values[(long long)v * P.n_inst + inst];
}

// ---- scalar policies -------------------------------------------------------------
// STRICT reproduces the reference's operation sequence with unfused IEEE operations
// (Complex.ts:33-47, solveComplex.ts:45-52); the default uses FMA contraction,
// squared-magnitude pivot metrics and reciprocal multiplies (1e-9 parity, not bitwise).

typedef double2 cplx;

template <typename T> struct Num;

template <> struct Num<double> {
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ double one() { return 1.0; }
  static __device__ __forceinline__ double nan() { return CUDART_NAN; }
  template <bool STRICT> static __device__ __forceinline__ double metric(double a) { return fabs(a); }
  template <bool STRICT> static __device__ __forceinline__ double thresh() { return kEps; }
  // |a|^2-style guard of Complex.div does not exist for reals.
  template <bool STRICT> static __device__ __forceinline__ bool div_guard(double, double) { return false; }
  static __device__ __forceinline__ double recip(double p) { return 1.0 / p; }
  static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
  static __device__ __forceinline__ double div_strict(double a, double p) { return __ddiv_rn(a, p); }
  template <bool STRICT> static __device__ __forceinline__ double submul(double a, double f, double p) {
    if (STRICT) return __dsub_rn(a, __dmul_rn(f, p));
    return fma(-f, p, a);
  }
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
  static __device__ __forceinline__ double neg(double a) { return -a; }
};

template <> struct Num<cplx> {
  static __device__ __forceinline__ cplx zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ cplx one() { return make_double2(1.0, 0.0); }
  static __device__ __forceinline__ cplx nan() { return make_double2(CUDART_NAN, CUDART_NAN); }
  template <bool STRICT> static __device__ __forceinline__ double metric(cplx a) {
    if (STRICT) return hypot(a.x, a.y);            // Complex.ts:55-57
    return fma(a.x, a.x, a.y * a.y);               // monotone in |a|
  }
  template <bool STRICT> static __device__ __forceinline__ double thresh() {
    return STRICT ? kEps : kEps * kEps;
  }
  // Complex.ts:41-42: throws when re^2+im^2 < EPS.  m is the pivot metric.
  template <bool STRICT> static __device__ __forceinline__ bool div_guard(double m, cplx pv) {
    if (STRICT) return __dadd_rn(__dmul_rn(pv.x, pv.x), __dmul_rn(pv.y, pv.y)) < kEps;
    return m < kEps;  // fast mode: the metric already is re^2+im^2
  }
  static __device__ __forceinline__ cplx recip(cplx p) {
    double inv = 1.0 / fma(p.x, p.x, p.y * p.y);
    return make_double2(p.x * inv, -p.y * inv);
  }
  static __device__ __forceinline__ cplx mul(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
  }
  static __device__ __forceinline__ cplx mul_strict(cplx a, cplx b) {  // Complex.ts:33-38
    return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                        __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
  }
  static __device__ __forceinline__ cplx div_strict(cplx a, cplx b) {  // Complex.ts:40-47
    double d = __dadd_rn(__dmul_rn(b.x, b.x), __dmul_rn(b.y, b.y));
    return make_double2(__ddiv_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), d),
                        __ddiv_rn(__dsub_rn(__dmul_rn(a.y, b.x), __dmul_rn(a.x, b.y)), d));
  }
  template <bool STRICT> static __device__ __forceinline__ cplx submul(cplx a, cplx f, cplx p) {
    if (STRICT) {  // row[j] = target.sub(f.mul(source))  solveComplex.ts:51
      cplx m = mul_strict(f, p);
      return make_double2(__dsub_rn(a.x, m.x), __dsub_rn(a.y, m.y));
    }
    return make_double2(fma(-f.x, p.x, fma(f.y, p.y, a.x)), fma(-f.x, p.y, fma(-f.y, p.x, a.y)));
  }
  static __device__ __forceinline__ cplx add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
  static __device__ __forceinline__ cplx neg(cplx a) { return make_double2(-a.x, -a.y); }
};

__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }

}  // namespace spicey
