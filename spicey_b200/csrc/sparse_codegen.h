// Straight-line CUDA source for one topology's sparse LU program (host side).
//
// The interpreter in ac_sparse.cuh spends ~50 instructions of decode and operand addressing on every
// micro-op that does 4 DFMAs of work.  For programs small enough to compile in seconds, the same
// single-assignment program (sparse_program.h: pilot pivot sequence, symbolic fill, per-system pivot
// verification) is instead *written out* as one straight-line sm_100a kernel: every value is a local
// `double2`, every operand a name, every stamped entry a literal expression in w = 2*pi*f — the CUDA
// compiler then does what the interpreter's host-side liveness pass approximates (register allocation,
// spilling only what survives until the back-substitution) and no decode work remains.  NVRTC compiles
// the source once per topology (cached per handle); semantics, verification and the dense fallback are
// exactly those of the interpreter, which stays the path for large programs and for component sweeps.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "sparse_program.h"

namespace spicey {

struct CodegenInput {
  const SparseProgram* sp = nullptr;
  int nn = 0, n_ac_elem = 0, v_first = 0;
  const int* n1 = nullptr;   // [n_ac_elem] node ids (0 = ground)
  const int* n2 = nullptr;
};

namespace codegen_detail {

inline std::string lit(double v) {  // exact hexadecimal floating literal
  char buf[64];
  if (v == 0.0) return "0.0";
  snprintf(buf, sizeof buf, "%a", v);
  return buf;
}

}  // namespace codegen_detail

// Kernel ABI of the generated source (must match JitArgs in spicey_native.cu).
inline const char* sparse_jit_prelude() {
  return R"SRC(
struct JitArgs {
  const double* freqs; long long p_count;
  double2* x; double2* ielem; int* status; long long series_ld;
  long long* fb_list; int* fb_count; int n; int n_ac_elem;
};
#define EPS 1e-15
#define THR 1e-30
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 submul(double2 a, double2 f, double2 p) {
  return make_double2(fma(-f.x, p.x, fma(f.y, p.y, a.x)), fma(-f.x, p.y, fma(-f.y, p.x, a.y)));
}
__device__ __forceinline__ double nrm(double2 a) { return fma(a.x, a.x, a.y * a.y); }
)SRC";
}

inline std::string generate_sparse_kernel_source(const CodegenInput& in) {
  using namespace codegen_detail;
  using namespace sparse_detail;
  const SparseProgram& sp = *in.sp;
  std::string s = sparse_jit_prelude();
  s.reserve(1 << 20);
  auto entry_expr = [&](int en) {
    // (alpha + Re J) + j*(w*beta - gamma/w + Im J), constants summed on the host in stamping order
    std::string im;
    const double b = sp.ent_beta[en], g = sp.ent_gamma[en], ji = sp.ent_jim[en];
    if (b != 0.0 && g != 0.0) im = "fma(w, " + lit(b) + ", -(" + lit(g) + ") * iw)";
    else if (b != 0.0) im = "w * " + lit(b);
    else if (g != 0.0) im = "-(" + lit(g) + ") * iw";
    else im = "0.0";
    if (ji != 0.0) im += " + " + lit(ji);
    return "make_double2(" + lit(sp.ent_alpha[en] + sp.ent_jre[en]) + ", " + im + ")";
  };
  auto opnd = [&](int o) -> std::string {
    if (o >= 0) return "v" + std::to_string(o);
    if (o == kNoOperand) return "Z";
    return "e" + std::to_string(~o);
  };
  s += "#ifndef MINB\n#define MINB 4\n#endif\n";
  s += "extern \"C\" __global__ void __launch_bounds__(128, MINB) spicey_sparse_jit(JitArgs a) {\n";
  s += "  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n";
  s += "  const long long nthreads = (long long)gridDim.x * blockDim.x;\n";
  s += "  const double2 Z = make_double2(0.0, 0.0);\n";
  s += "  for (long long p = tid; p < a.p_count; p += nthreads) {\n";
  s += "    const double w = 6.283185307179586 * a.freqs[p];\n    const double iw = 1.0 / w;\n";
  s += "    const long long xst = a.series_ld ? a.series_ld : 1;\n";
  s += "    double2* __restrict__ xout = a.series_ld ? a.x + p : a.x + p * a.n;\n";
  s += "    bool ok = true, bad = false;\n    int status = 0;\n    double mp, m;\n    double2 ap, r, fm, acc;\n";
  for (double L : sp.ind_L)  // inductor guards of simulateAC.ts:47-51 are value dependent: dense kernel decides
    s += "    { const double d = w * " + lit(L) + "; bad = bad || fabs(d) < EPS || d * d < EPS; }\n";
  // stamped entries used by the program, one expression each (the compiler merges identical ones)
  {
    std::vector<char> used(sp.n_stamp, 0);
    auto mark = [&](int o) { if (o < 0 && o != kNoOperand) used[~o] = 1; };
    for (const IrOp& op : sp.ir) {
      for (int o : op.reads) mark(o);
      for (const Update& u : op.upd) { mark(u.dst_old); mark(u.src); }
    }
    for (int en = 0; en < sp.n_stamp; ++en)
      if (used[en]) s += "    const double2 e" + std::to_string(en) + " = " + entry_expr(en) + ";\n";
  }
  for (const IrOp& op : sp.ir) {
    if (op.kind == SOP_PIVOT) {
      s += "    ap = " + opnd(op.reads[op.pidx]) + "; mp = nrm(ap); ok = ok && (mp == mp);\n";
      for (int c = 0; c < (int)op.reads.size(); ++c) {
        if (c == op.pidx) continue;
        s += "    m = nrm(" + opnd(op.reads[c]) + "); ok = ok && " + (c < op.pidx ? "(m < mp)" : "!(m > mp)") + ";\n";
      }
      s += "    if (status == 0) status = mp < THR ? 1 : (mp < EPS ? 2 : 0);\n";
      s += "    { const double inv = 1.0 / mp; r = make_double2(ap.x * inv, -ap.y * inv); }\n";
      s += "    const double2 v" + std::to_string(op.def) + " = r;\n";
    } else if (op.kind == SOP_ELIM) {
      s += "    fm = cmul(" + opnd(op.reads[0]) + ", r); if (nrm(fm) < THR) fm = Z;\n";
      for (const Update& u : op.upd)
        s += "    const double2 v" + std::to_string(u.dst_new) + " = submul(" + opnd(u.dst_old) + ", fm, " + opnd(u.src) + ");\n";
    } else {
      s += "    acc = " + opnd(op.reads[0]) + ";\n";
      for (size_t q = 2; q + 1 < op.reads.size(); q += 2)
        s += "    acc = submul(acc, " + opnd(op.reads[q]) + ", " + opnd(op.reads[q + 1]) + ");\n";
      s += "    const double2 v" + std::to_string(op.def) + " = cmul(acc, " + opnd(op.reads[1]) + ");\n";
      s += "    xout[" + std::to_string(op.var) + " * xst] = v" + std::to_string(op.def) + ";\n";
    }
  }
  s += "    if (bad || !ok) { a.status[p] = -1; a.fb_list[atomicAdd(a.fb_count, 1)] = p; continue; }\n";
  s += "    double2* __restrict__ io = a.ielem ? (a.series_ld ? a.ielem + p : a.ielem + p * a.n_ac_elem) : nullptr;\n";
  s += "    if (status != 0) {\n      const double qn = __longlong_as_double(0x7ff8000000000000ll);\n";
  s += "      for (int i = 0; i < a.n; ++i) xout[i * xst] = make_double2(qn, qn);\n";
  s += "      if (io) for (int e = 0; e < a.n_ac_elem; ++e) io[e * xst] = make_double2(qn, qn);\n";
  s += "      a.status[p] = status;\n      continue;\n    }\n";
  s += "    if (io) {\n";
  auto xv = [&](int node) { return node == 0 ? std::string("Z") : "v" + std::to_string(sp.x_virtual[node - 1]); };
  for (int e = 0; e < in.n_ac_elem; ++e) {
    if (e >= in.v_first) {
      s += "      io[" + std::to_string(e) + " * xst] = v" + std::to_string(sp.x_virtual[in.nn + e - in.v_first]) + ";\n";
      continue;
    }
    std::string y = "make_double2(" + lit(sp.el_a[e]) + ", ";
    if (sp.el_b[e] != 0.0) y += "w * " + lit(sp.el_b[e]);
    else if (sp.el_g[e] != 0.0) y += "-(" + lit(sp.el_g[e]) + ") * iw";
    else y += "0.0";
    y += ")";
    s += "      io[" + std::to_string(e) + " * xst] = cmul(" + y + ", csub(" + xv(in.n1[e]) + ", " + xv(in.n2[e]) + "));\n";
  }
  s += "    }\n    a.status[p] = 0;\n  }\n}\n";
  return s;
}

}  // namespace spicey
