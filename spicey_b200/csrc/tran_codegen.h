// Straight-line CUDA source of the persistent transient kernel for ONE netlist topology (host side).
//
// tran_small.cuh runs any small circuit: element records, node offsets and constants sit in shared memory
// and every step walks per-type element loops — ~330 warp instructions per recorded step for a three-element
// tank, which makes the kernel issue-bound long before the 48 B/step of results load HBM.  For a batch worth
// compiling for, the same time loop (simulateTRAN.ts:146-238, same stamping order R,C,L,S,V,D, same
// iteration policy: re-solve only while a switch toggled, diode linearised about vdPrev at iteration 0) is
// written out for the netlist at hand: every element is a handful of named scalars in registers, every
// matrix entry an ordered sum of named conductances, the LU a fully unrolled NV x NV elimination with
// select-based pivoting (solveReal.ts:14-71 on the same operands in the same order).  NVRTC compiles it once
// per topology and handle (spicey_native.cu); STRICT mode and large systems stay on the generic kernels.
//
//  * factor reuse: no switch and no diode -> the matrix is constant, factored once before the time loop;
//    switches only -> re-factored when a switch toggled; diodes -> every solve (as tran_small.cuh);
//  * diode exponentials: the companion model needs exp(clamp(vdPrev)/vth) and the recorded current of the
//    previous step already computed exp(vdPrev/vth) — the same bits whenever the clamp is inactive, and a
//    per-instance constant when it is active — so a step costs one exp instead of two.
#pragma once
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace spicey {

struct TranCodegenInput {
  int nn = 0, nV = 0, nvar = 0, n_elem = 0, n_state = 0;
  const int* off = nullptr;          // [7] first element of each type R,C,L,V,S,D (+ end)
  const int* n1 = nullptr;           // [n_elem] node ids, 0 = ground
  const int* n2 = nullptr;
  const int* nc1 = nullptr;          // switch control nodes
  const int* nc2 = nullptr;
  const int* value_idx = nullptr;    // [n_elem] first value slot
  const int* state_idx = nullptr;    // [n_elem] state slot or -1
  const double* values = nullptr;    // [n_values] nominal values
  const int* var_of_slot = nullptr;  // [n_values] sweep variable of a slot or -1
  bool with_ielem = true;
  int block = 64;
  // Device-evaluated source waveforms (null: every source is dc or a pre-sampled row, chosen by a.vmask):
  const int* wave_kind = nullptr;    // [nV] 0 dc, 1 pre-sampled row, 2 PULSE, 3 PWL
  const int* wave_vidx = nullptr;    // [nV] first value slot of the parameters
  const int* wave_npairs = nullptr;  // [nV] PWL pair count
};
constexpr int kTranJitMaxPwlPairs = 32;

inline const char* tran_jit_prelude() {
  return R"SRC(
struct TranJitArgs {
  const double* var_values; long long n_inst;
  double dt; long long steps;
  const double* vsrc; unsigned vmask;
  const double* state0; long long inst0; long long n_local;
  double* v; double* ielem; double* state_out; int* iters; int* status;
};
#define EPS 1e-15
#define VT300 0.02585
// 1/a: MUFU seed + two Newton steps (<= 1 ulp from the correctly rounded quotient; no slow-path branch and
// 4 instead of ~10 FP64-pipe instructions — the diode circuits are bound by that pipe)
__device__ __forceinline__ double rcp_nr(double a) {
  double y, e;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  return y;
}
// pulseValue.ts:4-22 and one segment of pwlValue.ts:9-13, every operation separately rounded (the branch
// conditions sit on pulse edges; the result must equal the host's pre-sampled table bit for bit).
__device__ __forceinline__ double pulse_value(double v1, double v2, double td, double tr, double tf, double ton,
                                              double period, double ncycles, double t) {
  if (t < td) return v1;
  const double tt = __dsub_rn(t, td);
  const double cycles = floor(__ddiv_rn(tt, period));
  if (cycles >= ncycles) return v1;
  const double tc = __dsub_rn(tt, __dmul_rn(cycles, period));
  if (tc < tr) return __dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), __ddiv_rn(tc, fmax(tr, EPS))));
  const double t1 = __dadd_rn(tr, ton);
  if (tc < t1) return v2;
  if (tc < __dadd_rn(t1, tf)) return __dadd_rn(v2, __dmul_rn(__dsub_rn(v1, v2), __ddiv_rn(__dsub_rn(tc, t1), fmax(tf, EPS))));
  return v1;
}
__device__ __forceinline__ double pwl_segment(double pt, double pv, double ct, double cv, double t) {
  const double a = __ddiv_rn(__dsub_rn(t, pt), fmax(__dsub_rn(ct, pt), EPS));
  return __dadd_rn(pv, __dmul_rn(__dsub_rn(cv, pv), a));
}
// solveReal.ts:14-71 for an NV x NV system held in registers: factor() once per matrix, solve() per
// right-hand side (replays the recorded interchanges and multipliers: the same operations on the same
// operands as the reference's augmented elimination).
template <int NV>
struct SmallLU {
  double f[NV][NV];
  int perm[NV];
  __device__ __forceinline__ int factor() {
    int status = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int imax = k;
      double vmax = fabs(f[k][k]);
#pragma unroll
      for (int i = k + 1; i < NV; ++i) {
        const double v = fabs(f[i][k]);
        if (v > vmax) { vmax = v; imax = i; }
      }
      if (vmax < EPS) status = 1;
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
      const double rp = rcp_nr(f[k][k]);
#pragma unroll
      for (int i = k + 1; i < NV; ++i) {
        double m = f[i][k] * rp;
        m = (fabs(m) < EPS) ? 0.0 : m;
        f[i][k] = m;
#pragma unroll
        for (int j = k + 1; j < NV; ++j) f[i][j] = fma(-m, f[k][j], f[i][j]);
      }
      f[k][k] = rp;
    }
    return status;
  }
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
      for (int i = k + 1; i < NV; ++i) b[i] = fma(-f[i][k], b[k], b[i]);
#pragma unroll
    for (int i = NV - 1; i >= 0; --i) {
      double s = b[i];
#pragma unroll
      for (int j = i + 1; j < NV; ++j) s = fma(-f[i][j], b[j], s);
      b[i] = s * f[i][i];
    }
  }
};
)SRC";
}

inline std::string generate_tran_kernel_source(const TranCodegenInput& in) {
  const int nn = in.nn, nvar = in.nvar, ne = in.n_elem;
  const int oR = in.off[0], oC = in.off[1], oL = in.off[2], oV = in.off[3], oS = in.off[4], oD = in.off[5], oE = in.off[6];
  (void)oR;
  const bool has_sw = oD > oS, has_d = oE > oD, dyn = has_sw || has_d;
  auto N = [](int v) { return std::to_string(v); };
  auto hexlit = [](double v) {
    char buf[64];
    if (v == 0.0) return std::string("0.0");
    if (v != v || v - v != 0.0) {   // NaN / +-Infinity (PULSE ncycles defaults to Infinity): no literal form
      unsigned long long bits;
      memcpy(&bits, &v, sizeof bits);
      snprintf(buf, sizeof buf, "__longlong_as_double(0x%llxll)", bits);
      return std::string(buf);
    }
    snprintf(buf, sizeof buf, "%a", v);
    return v < 0 ? std::string("(") + buf + ")" : std::string(buf);
  };
  auto val = [&](int slot) -> std::string {   // value of a slot for this instance
    const int v = in.var_of_slot[slot];
    if (v < 0) return hexlit(in.values[slot]);
    return "a.var_values[" + N(v) + "ll * a.n_inst + inst]";
  };
  auto xs = [&](int node) -> std::string { return node == 0 ? std::string("0.0") : "x[" + N(node - 1) + "]"; };
  auto diff = [&](int a, int b) -> std::string {
    if (a == 0 && b == 0) return "0.0";
    if (b == 0) return xs(a);
    if (a == 0) return "(0.0 - " + xs(b) + ")";
    return "(" + xs(a) + " - " + xs(b) + ")";
  };

  std::string s = tran_jit_prelude();
  s.reserve(1 << 16);
  s += "#define BLOCK " + N(in.block) + "\n#define NV " + N(nvar) + "\n";
  s += "extern \"C\" __global__ void __launch_bounds__(BLOCK) spicey_tran_jit(TranJitArgs a) {\n";
  s += "  const long long li = (long long)blockIdx.x * BLOCK + threadIdx.x;\n  if (li >= a.n_local) return;\n";
  s += "  const long long inst = a.inst0 + li, NL = a.n_local, S1 = a.steps + 1;\n";
  s += "  const double dtc = fmax(a.dt, EPS);\n";
  // ---- per-instance element constants (tran_kernels.cuh element_constants) ----
  for (int e = 0; e < ne; ++e) {
    const std::string E = N(e);
    const int vi = in.value_idx[e];
    if (e < oC) s += "  const double g" + E + " = 1 / " + val(vi) + ";\n";                       // R :36-38
    else if (e < oL) s += "  const double g" + E + " = " + val(vi) + " / dtc;\n";                 // C :41-43
    else if (e < oV) s += "  const double g" + E + " = dtc / " + val(vi) + ";\n";                 // L :47-49
    else if (e < oS) s += "  const double dc" + E + " = " + val(vi) + ";\n";                      // V :66-69
    else if (e < oD) {                                                                             // S :56-63
      s += "  const double ron" + E + " = fmax(fabs(" + val(vi) + "), EPS), roff" + E + " = fmax(fabs(" + val(vi + 1) + "), EPS);\n";
      s += "  const double von" + E + " = " + val(vi + 2) + ", voff" + E + " = " + val(vi + 3) + ";\n";
    } else {                                                                                       // D :72-101
      s += "  const double is" + E + " = " + val(vi) + ", vth" + E + " = " + val(vi + 1) + " * VT300;\n";
      s += "  const double isv" + E + " = is" + E + " / vth" + E + ", ivth" + E + " = 1.0 / vth" + E + ";\n";
      s += "  const double elo" + E + " = exp(-1.0 * ivth" + E + "), ehi" + E + " = exp(0.8 * ivth" + E + ");\n";
    }
  }
  // ---- per-instance waveform parameters (value slots like any other: a sweep may vary them) ----
  auto wkind = [&](int k) { return in.wave_kind ? in.wave_kind[k] : -1; };
  for (int e = oV; e < oS; ++e) {
    const int k = e - oV, ws = in.wave_vidx ? in.wave_vidx[k] : 0;
    const std::string E = N(e);
    if (wkind(k) == 2) {
      const char* nm[8] = {"v1", "v2", "td", "tr", "tf", "ton", "per", "ncy"};
      for (int q = 0; q < 8; ++q) s += "  const double pw" + E + nm[q] + " = " + val(ws + q) + ";\n";
    } else if (wkind(k) == 3) {
      for (int q = 0; q < in.wave_npairs[k]; ++q)
        s += "  const double pl" + E + "t" + N(q) + " = " + val(ws + 2 * q) + ", pl" + E + "v" + N(q) + " = " + val(ws + 2 * q + 1) + ";\n";
    }
  }
  for (int e = 0; e < ne; ++e)
    if (in.state_idx[e] >= 0)
      s += "  double st" + N(in.state_idx[e]) + " = a.state0 ? a.state0[" + N(in.state_idx[e]) + "ll * a.n_inst + inst] : 0.0;\n";
  for (int e = oD; e < oE; ++e)   // exp(vdPrev / vth), kept current by the recording step
    s += "  double ex" + N(e) + " = exp(st" + N(in.state_idx[e]) + " * ivth" + N(e) + ");\n";

  // ---- matrix entries as ordered sums (stampAdmittanceReal.ts:3-29, stampVoltageSourceReal.ts:4-32) ----
  // static part: R, C, L then V; dynamic part appended per solve: S then D (V never shares an entry with them)
  std::map<std::pair<int, int>, std::string> stat, dynm;
  auto acc = [&](std::map<std::pair<int, int>, std::string>& M, int i, int j, const std::string& term, bool neg) {
    std::string& cur = M[std::make_pair(i, j)];
    if (cur.empty()) cur = "0.0";
    cur = "(" + cur + (neg ? " - " : " + ") + term + ")";
  };
  auto admittance = [&](std::map<std::pair<int, int>, std::string>& M, int e, const std::string& g) {
    const int i1 = in.n1[e] - 1, i2 = in.n2[e] - 1;
    if (i1 >= 0) acc(M, i1, i1, g, false);
    if (i2 >= 0) acc(M, i2, i2, g, false);
    if (i1 >= 0 && i2 >= 0) { acc(M, i1, i2, g, true); acc(M, i2, i1, g, true); }
  };
  for (int e = 0; e < oV; ++e) admittance(stat, e, "g" + N(e));
  for (int e = oV; e < oS; ++e) {
    const int i1 = in.n1[e] - 1, i2 = in.n2[e] - 1, j = nn + (e - oV);
    if (i1 >= 0) acc(stat, i1, j, "1.0", false);
    if (i2 >= 0) acc(stat, i2, j, "1.0", true);
    if (i1 >= 0) acc(stat, j, i1, "1.0", false);
    if (i2 >= 0) acc(stat, j, i2, "1.0", true);
  }
  for (auto& kv : stat) s += "  const double ac_" + N(kv.first.first) + "_" + N(kv.first.second) + " = " + kv.second + ";\n";
  for (int e = oS; e < oD; ++e) admittance(dynm, e, "gs" + N(e));
  for (int e = oD; e < oE; ++e) admittance(dynm, e, "gd" + N(e));
  auto entry = [&](int i, int j) -> std::string {   // full expression of A(i,j) for this solve
    const auto key = std::make_pair(i, j);
    std::string base = stat.count(key) ? "ac_" + N(i) + "_" + N(j) : std::string("0.0");
    if (!dynm.count(key)) return base;
    std::string d = dynm[key];      // "(...((0.0 + a) + b)...)": continue the static sum instead of 0.0
    const size_t z = d.find("0.0");
    d.replace(z, 3, base);
    return d;
  };
  auto load_matrix = [&](const std::string& ind) {
    std::string t;
    for (int i = 0; i < nvar; ++i)
      for (int j = 0; j < nvar; ++j) t += ind + "lu.f[" + N(i) + "][" + N(j) + "] = " + entry(i, j) + ";\n";
    return t;
  };

  s += "  SmallLU<NV> lu;\n  int status = 0;\n  double x[NV];\n";
  for (int i = 0; i < nvar; ++i) s += "  x[" + N(i) + "] = 0.0;\n";
  if (!dyn) s += load_matrix("  ") + "  status = lu.factor();\n";
  if (has_sw && !has_d) s += "  bool factored = false;\n";
  s += "  char* vo = (char*)(a.v + li);\n";
  if (in.with_ielem) s += "  char* io = (char*)(a.ielem + li);\n";
  s += "  const unsigned pitch = (unsigned)NL * 8u;   // bytes between consecutive rows (n_local < 2^29 checked by the host)\n";
  s += "  int* ito = a.iters ? a.iters + li : nullptr;\n";
  s += "  const size_t v_stride = (size_t)pitch * " + N(nn) + "u, i_stride = (size_t)pitch * " + N(ne) + "u;\n";
  s += "  long long step = 0;\n  int pat = -1;   // pivot sequence of the previous factorisation (diode circuits)\n";
  s += "  for (; step < S1 && status == 0; ++step) {\n";
  bool any_dev_wave = false;
  for (int e = oV; e < oS; ++e) any_dev_wave |= wkind(e - oV) >= 2;
  if (any_dev_wave) s += "    const double tnow = __dmul_rn((double)step, a.dt);   // t = step * dt (simulateTRAN.ts:147)\n";
  for (int e = oV; e < oS; ++e) {
    const int k = e - oV;
    const std::string E = N(e);
    if (wkind(k) == 2) {
      s += "    const double vs" + E + " = pulse_value(pw" + E + "v1, pw" + E + "v2, pw" + E + "td, pw" + E + "tr, pw" + E + "tf, pw" + E +
           "ton, pw" + E + "per, pw" + E + "ncy, tnow);\n";
    } else if (wkind(k) == 3) {   // pwlValue.ts:3-16, the pair loop written out
      const int np = in.wave_npairs[k];
      if (np <= 0) { s += "    const double vs" + E + " = 0.0;\n"; continue; }
      s += "    double vs" + E + " = pl" + E + "v" + N(np - 1) + ";\n";
      s += "    if (tnow <= pl" + E + "t0) vs" + E + " = pl" + E + "v0;\n";
      for (int q = 1; q < np; ++q)
        s += "    else if (tnow <= pl" + E + "t" + N(q) + ") vs" + E + " = pwl_segment(pl" + E + "t" + N(q - 1) + ", pl" + E + "v" + N(q - 1) +
             ", pl" + E + "t" + N(q) + ", pl" + E + "v" + N(q) + ", tnow);\n";
    } else {
      s += "    const double vs" + E + " = ((a.vmask >> " + N(k) + ") & 1u) ? __ldg(a.vsrc + " + N(k) + "ll * S1 + step) : dc" + E + ";\n";
    }
  }
  s += "    int it = 0;\n";
  if (has_sw) s += "    for (; it < 20; ++it) {\n";
  else s += "    {\n";
  // right-hand side: C, L, V, then D (:41-53, :66-69, :98-100)
  for (int i = 0; i < nvar; ++i) s += "      double b" + N(i) + " = 0.0;\n";
  auto rhs = [&](int e, const std::string& cur) {   // b[n+] -= cur; b[n-] += cur  (stampCurrentReal.ts:3-14)
    if (in.n1[e] > 0) s += "      b" + N(in.n1[e] - 1) + " -= " + cur + ";\n";
    if (in.n2[e] > 0) s += "      b" + N(in.n2[e] - 1) + " += " + cur + ";\n";
  };
  for (int e = oC; e < oL; ++e) {
    s += "      const double ieq" + N(e) + " = -g" + N(e) + " * st" + N(in.state_idx[e]) + ";\n";
    rhs(e, "ieq" + N(e));
  }
  for (int e = oL; e < oV; ++e) rhs(e, "st" + N(in.state_idx[e]));
  for (int e = oV; e < oS; ++e) s += "      b" + N(nn + e - oV) + " += vs" + N(e) + ";\n";
  for (int e = oS; e < oD; ++e)
    s += "      const double gs" + N(e) + " = 1 / (st" + N(in.state_idx[e]) + " != 0.0 ? ron" + N(e) + " : roff" + N(e) + ");\n";
  for (int e = oD; e < oE; ++e) {
    const std::string E = N(e), ST = "st" + N(in.state_idx[e]);
    // :85 vd = iter == 0 ? vdPrev : x[+] - x[-];  :87-97 clamp, exp, gd floor, ieq
    s += "      double gd" + E + ", jd" + E + ";\n      {\n";
    s += "        const double vd = " + (has_sw ? "it == 0 ? " + ST + " : " + diff(in.n1[e], in.n2[e]) : ST) + ";\n";
    s += "        const double vlim = vd > 0.8 ? 0.8 : (vd < -1.0 ? -1.0 : vd);\n";
    if (has_sw) s += "        const double ee = it == 0 ? (vd > 0.8 ? ehi" + E + " : (vd < -1.0 ? elo" + E + " : ex" + E + ")) : exp(vlim * ivth" + E + ");\n";
    else s += "        const double ee = vd > 0.8 ? ehi" + E + " : (vd < -1.0 ? elo" + E + " : ex" + E + ");\n";
    s += "        const double id = is" + E + " * (ee - 1.0);\n";
    s += "        gd" + E + " = fmax(isv" + E + " * ee, 1e-12);\n";
    s += "        jd" + E + " = id - gd" + E + " * vlim;\n      }\n";
    rhs(e, "jd" + E);
  }
  const bool fast_piv = has_d && nvar <= 4;
  if (fast_piv) {
    // Diode circuits re-factor at every solve, and the select-based pivoting is most of a step's instructions
    // (63 of 248 for cfg 5).  The pivot sequence rarely changes from one solve to the next, so each of the NV!
    // possible sequences gets its own straight-line elimination of [A | b] on named scalars (no selects, and
    // the entries that are structurally 0 or 1 fold away); the case of the previous solve's sequence runs
    // first and checks every pivot it assumes against the reference's rule (first maximum wins,
    // solveReal.ts:16-26).  If one check fails the generic code below decides, as it does for the first solve.
    // Same operations on the same operands in the same order either way.
    for (int i = 0; i < nvar; ++i)
      for (int j = 0; j < nvar; ++j) s += "      const double m_" + N(i) + "_" + N(j) + " = " + entry(i, j) + ";\n";
    s += "      bool redo = true;\n      switch (pat) {\n";
    std::vector<int> perm(nvar, 0);
    for (int k = 0; k < nvar; ++k) perm[k] = k;
    for (;;) {
      int id = 0, mul = 1;
      for (int k = 0; k < nvar; ++k) { id += perm[k] * mul; mul *= nvar; }
      s += "        case " + N(id) + ": {\n";
      std::vector<int> L(nvar);
      for (int i = 0; i < nvar; ++i) L[i] = i;
      auto A = [&](int r, int c) { return "a" + N(r) + "_" + N(c); };
      for (int r = 0; r < nvar; ++r) {
        for (int c = 0; c < nvar; ++c) s += "          double " + A(r, c) + " = m_" + N(r) + "_" + N(c) + ";\n";
        s += "          double " + A(r, nvar) + " = b" + N(r) + ";\n";
      }
      s += "          bool okp = true;\n";
      for (int k = 0; k < nvar; ++k) {
        const int j = perm[k], prow = L[j];
        s += "          { const double vj = fabs(" + A(prow, k) + "); okp = okp && (vj >= EPS);\n";
        for (int q = k; q < nvar; ++q) {
          if (q == j) continue;
          s += "            okp = okp && " + std::string(q < j ? "(fabs(" + A(L[q], k) + ") < vj)" : "!(fabs(" + A(L[q], k) + ") > vj)") + ";\n";
        }
        s += "          }\n";
        std::swap(L[k], L[j]);
        s += "          const double rp" + N(k) + " = rcp_nr(" + A(prow, k) + ");\n";
        for (int q = k + 1; q < nvar; ++q) {
          const int r = L[q];
          s += "          { double mm = " + A(r, k) + " * rp" + N(k) + "; mm = (fabs(mm) < EPS) ? 0.0 : mm;\n";
          for (int c = k + 1; c <= nvar; ++c) s += "            " + A(r, c) + " = fma(-mm, " + A(prow, c) + ", " + A(r, c) + ");\n";
          s += "          }\n";
        }
      }
      for (int i = nvar - 1; i >= 0; --i) {
        const int r = L[i];
        s += "          double xx" + N(i) + " = " + A(r, nvar) + ";\n";
        for (int c = i + 1; c < nvar; ++c) s += "          xx" + N(i) + " = fma(-" + A(r, c) + ", xx" + N(c) + ", xx" + N(i) + ");\n";
        s += "          xx" + N(i) + " *= rp" + N(i) + ";\n";
      }
      s += "          if (okp) {";
      for (int i = 0; i < nvar; ++i) s += " x[" + N(i) + "] = xx" + N(i) + ";";
      s += " redo = false; }\n        } break;\n";
      // next pivot sequence: perm[k] in [k, nvar)
      int k = nvar - 1;
      while (k >= 0 && perm[k] == nvar - 1) { perm[k] = k; --k; }
      if (k < 0) break;
      ++perm[k];
    }
    s += "        default: break;\n      }\n      if (redo) {\n";
    for (int i = 0; i < nvar; ++i)
      for (int j = 0; j < nvar; ++j) s += "        lu.f[" + N(i) + "][" + N(j) + "] = m_" + N(i) + "_" + N(j) + ";\n";
    s += "        status = lu.factor();\n        if (status != 0) break;\n";
    for (int i = 0; i < nvar; ++i) s += "        x[" + N(i) + "] = b" + N(i) + ";\n";
    s += "        lu.solve(x);\n        pat = 0;\n";
    s += "        { int mul = 1; for (int k = 0; k < NV; ++k) { pat += lu.perm[k] * mul; mul *= NV; } }\n      }\n";
  } else {
    if (has_d) s += load_matrix("      ") + "      status = lu.factor();\n";
    else if (has_sw) s += "      if (!factored) {\n" + load_matrix("        ") + "        status = lu.factor();\n        factored = true;\n      }\n";
    if (dyn) s += "      if (status != 0) break;\n";
    for (int i = 0; i < nvar; ++i) s += "      x[" + N(i) + "] = b" + N(i) + ";\n";
    s += "      lu.solve(x);\n";
  }
  if (has_sw) {                                                     // :108-128
    s += "      bool switched = false;\n";
    for (int e = oS; e < oD; ++e) {
      const std::string E = N(e), ST = "st" + N(in.state_idx[e]);
      s += "      {\n        const double vctrl = " + diff(in.nc1[e], in.nc2[e]) + ";\n";
      s += "        const bool on = " + ST + " != 0.0;\n        bool nxt = on;\n";
      s += "        if (on) { if (vctrl < voff" + E + ") nxt = false; } else if (vctrl > von" + E + ") nxt = true;\n";
      s += "        if (nxt != on) { " + ST + " = nxt ? 1.0 : 0.0; switched = true; }\n      }\n";
    }
    if (!has_d) s += "      if (switched) factored = false;\n";
    s += "      if (!switched) break;\n";
  }
  s += "    }\n";
  if (dyn) s += "    if (status != 0) break;\n";
  s += "    if (ito) { *ito = it < 20 ? it + 1 : 20; ito += NL; }\n";
  // ---- recording (:164-219) and state update (:221-237), table order ----
  for (int i = 0; i < nn; ++i) s += "    *(double*)(vo + (size_t)pitch * " + N(i) + "u) = x[" + N(i) + "];\n";
  s += "    vo += v_stride;\n";
  for (int e = 0; e < ne; ++e) {
    const std::string E = N(e), ST = in.state_idx[e] >= 0 ? "st" + N(in.state_idx[e]) : std::string();
    std::string cur;
    s += "    {\n";
    if (e >= oV && e < oS) cur = "x[" + N(nn + e - oV) + "]";
    else {
      s += "      const double d = " + diff(in.n1[e], in.n2[e]) + ";\n";
      if (e < oC) cur = "d * g" + E;
      else if (e < oL) { s += "      const double cur = g" + E + " * (d - " + ST + ");\n      " + ST + " = d;\n"; cur = "cur"; }
      else if (e < oV) { s += "      const double cur = g" + E + " * d + " + ST + ";\n      " + ST + " = cur;\n"; cur = "cur"; }
      else if (e < oD) cur = "d / (" + ST + " != 0.0 ? ron" + E + " : roff" + E + ")";
      else {  // unclamped vd (:213-217)
        s += "      ex" + E + " = exp(d * ivth" + E + ");\n      " + ST + " = d;\n";
        cur = "is" + E + " * (ex" + E + " - 1.0)";
      }
    }
    if (in.with_ielem) s += "      *(double*)(io + (size_t)pitch * " + E + "u) = " + cur + ";\n";
    s += "    }\n";
  }
  if (in.with_ielem) s += "    io += i_stride;\n";
  s += "  }\n";
  // a failed instance: NaN rows from the failing step on (as the generic kernels)
  s += "  if (status != 0) {\n    const double qn = __longlong_as_double(0x7ff8000000000000ll);\n";
  s += "    for (; step < S1; ++step) {\n";
  s += "      for (int i = 0; i < " + N(nn) + "; ++i) a.v[(step * " + N(nn) + " + i) * NL + li] = qn;\n";
  if (in.with_ielem) s += "      for (int e = 0; e < " + N(ne) + "; ++e) a.ielem[(step * " + N(ne) + " + e) * NL + li] = qn;\n";
  s += "      if (a.iters) a.iters[step * NL + li] = 0;\n    }\n  }\n";
  s += "  if (a.state_out) {\n";
  for (int e = 0; e < ne; ++e)
    if (in.state_idx[e] >= 0) s += "    a.state_out[" + N(in.state_idx[e]) + "ll * NL + li] = st" + N(in.state_idx[e]) + ";\n";
  s += "  }\n  a.status[li] = status;\n}\n";
  return s;
}

}  // namespace spicey
