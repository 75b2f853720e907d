// Straight-line CUDA source for one topology's sparse LU program (host side).
//
// The interpreter in ac_sparse.cuh spends ~50 instructions of decode and operand addressing on every
// micro-op that does 4 DFMAs of work.  For programs small enough to compile in seconds, the same
// single-assignment program (sparse_program.h: pilot pivot sequence, symbolic fill, per-system pivot
// verification) is instead *written out* as one straight-line sm_100a kernel that NVRTC compiles once
// per topology (cached per handle).  Semantics, verification and the dense fallback are exactly those
// of the interpreter, which stays the path for large programs and for component sweeps.
//
// What the generator decides (the compiler cannot):
//  * Where the factorisation lives between the two phases.  One thread owns one system; the values the
//    forward elimination produces for the back-substitution (1/u_kk, eliminated rhs, modified U) are
//    ~2 KB per system.  Left to the register allocator they spill to local memory and the dependent
//    chain of the back-substitution then waits on L2 for every one of them.  Here the values with the
//    longest def->use distance are placed in a per-thread column of shared memory ([slot][thread], one
//    conflict-free 16-byte access per lane) and re-loaded by name in the back phase; the short-lived
//    rest stays in registers.  Nothing of the matrix touches HBM.
//  * Element currents are emitted as soon as both node voltages exist, so an x_i dies a few
//    instructions after it is produced instead of surviving until an unpack loop.
//  * Arithmetic is specialised on what the operand structurally is: a stamped entry that is purely real
//    (1/R sums), purely imaginary (w*C - 1/(w*L)) or +-1 (source rows) costs half the DFMAs of a
//    complex multiply, with bit-identical results (the dropped terms are exact zeros).
//  * The reference's `|f| < EPS -> skip row` (solveComplex.ts:46) zeroes the multiplier (branch-free); 1/|pivot|^2
//    is a MUFU seed + two Newton steps, so the whole elimination is one basic block; real constants are
//    operands from a __constant__ table instead of 64-bit immediates.
//  * A bulk-copy (cp.async.bulk) epilogue staging the results in freed factor slots was built and measured
//    slower (the issuing warp pays ~90 cycles per copy, tools/micro/bulk_store.cu); results leave as plain
//    predicated 16-byte stores, 512 contiguous bytes per warp and series.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "sparse_program.h"

namespace spicey {

struct CodegenInput {
  const SparseProgram* sp = nullptr;
  int nn = 0, n_ac_elem = 0, v_first = 0;
  const int* n1 = nullptr;   // [n_ac_elem] node ids (0 = ground)
  const int* n2 = nullptr;
  // Per-instance stamping (component sweeps / Monte-Carlo on the AC axis): the entries are sums of element
  // admittances that differ from instance to instance, so the kernel builds them from the element table.
  bool eager = false;
  const int* ent_ptr = nullptr;      // gather plan: entry en sums contrib[ent_ptr[en] .. ent_ptr[en + 1])
  const int* contrib = nullptr;      // (element << 3) | (source << 1) | negate; source 0 = admittance, 1 = phasor, 2 = one
  const int* el_type = nullptr;      // [n_ac_elem] 0 R, 1 C, 2 L, 3 V
  const int* el_vidx = nullptr;      // [n_ac_elem] first value slot
  const int* var_of_slot = nullptr;  // [n_values] sweep variable of a slot or -1
  const double* values = nullptr;    // [n_values] nominal values
};

struct CodegenOptions {
  int block = 128;        // threads per CTA
  int min_blocks = 2;     // __launch_bounds__ second argument
  int smem_slots = 48;    // shared-memory double2 slots per thread for the factor values
  bool with_ielem = true; // false: the caller passed ielem = NULL, no current is computed
  int sync_every = 0;     // > 0: __syncthreads() every that many pivots / back-substitution rows (instruction-cache locality)
  int stagger_ns = 0;     // > 0: CTAs start in four phases this many ns apart, so that the SMs are not all storing at once
  int prefetch_steps = 8; // per-instance stamping: element values are loaded this many pivots / rows ahead of their use
  int antiphase_ns = 0;     // min_blocks >= 2: start delay of the k-th wave of CTAs (k * antiphase_ns)
  int reg_values = -1;      // >= 0: cross-phase values that may stay in registers beside the shared-memory slots; the
                            // longest-lived of the rest go to a [slot][thread] column of global memory (JitArgs.work)
  int gmem_ahead = 6;       // back-substitution rows between the load of a global-column value and its use
  bool l2_policy = true;    // global column: L2 evict_last, result stores: evict_first (the column of the resident grid is
                            // rewritten by every point and fits L2 for mid-size programs; the results only pass through)
};

struct CodegenStats {
  int n_saved = 0;        // values crossing from the elimination into the back-substitution
  int smem_slots = 0;     // of those, placed in shared memory
  int gmem_slots = 0;     // ... and in the global column (double2 per thread each)
  size_t smem_bytes = 0;  // dynamic shared memory per CTA
  int n_classes = 0;      // distinct stamped values
};

namespace codegen_detail {

inline std::string hexlit(double v) {  // exact hexadecimal floating literal
  char buf[64];
  if (v == 0.0) return "0.0";
  snprintf(buf, sizeof buf, "%a", v);
  return v < 0 ? std::string("(") + buf + ")" : std::string(buf);
}

// A complex operand as two scalar expressions plus what is structurally known about it.
struct Opnd {
  std::string re, im;
  bool re0 = false, im0 = false;   // component is an exact structural zero
  double re_c = 0;                 // value of re when it is a literal
  bool re_lit = false;
};

}  // namespace codegen_detail

// Kernel ABI of the generated source (must match JitArgs in spicey_native.cu).
inline const char* sparse_jit_prelude() {
  return R"SRC(
struct JitArgs {
  const double* freqs; long long p_count;
  double2* x; double2* ielem; int* status; long long series_ld;
  long long* fb_list; int* fb_count; int n; int n_ac_elem;
  const double* var_values; long long n_inst; long long n_freq; long long p_begin;   // per-instance stamping
  double2* work;   // [gmem slot][gridDim.x * BLOCK]: the factor values that fit neither registers nor shared memory
};
#define EPS 1e-15
#define THR 1e-30
#define D2(a, b) make_double2((a), (b))
// 1/a for a in [1e-15, huge): MUFU seed + two Newton steps (<= 1 ulp off the correctly rounded quotient,
// no slow-path branch, so the whole elimination stays one basic block)
__device__ __forceinline__ double rcp_nr(double a) {
  double y, e;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  e = fma(-a, y, 1.0); y = fma(y, e, y);
  return y;
}
extern __shared__ double2 sm[];
// Shared-memory column of this thread, addressed with immediate offsets.  Inline PTX on purpose: with plain
// C++ accesses the compiler forwards every stored value to its re-load and keeps it in a register (or in
// local memory) across the whole elimination, which is exactly what the placement is meant to avoid.
#define SMST(off, v) asm volatile("st.shared.v2.f64 [%0+" #off "], {%1, %2};" :: "r"(sbase), "d"((v).x), "d"((v).y) : "memory")
#define SMLD(v, off) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+" #off "];" : "=d"((v).x), "=d"((v).y) : "r"(sbase) : "memory")
// Global column of the factor values (programs past a thread's registers + shared memory) with an L2 evict_last policy,
// result stores with evict_first: the column is rewritten by every point and should stay in L2, the results only pass.
#define GST(p, v) asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(p), "d"((v).x), "d"((v).y), "l"(pol_keep) : "memory")
#define GLD(v, p) asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"((v).x), "=d"((v).y) : "l"(p), "l"(pol_keep) : "memory")
#define RST(p, re, im) asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(p), "d"((double)(re)), "d"((double)(im)), "l"(pol_stream) : "memory")
)SRC";
}

// Values the elimination hands to the back-substitution (1/u_kk, eliminated rhs, modified U): what has to fit the
// registers and shared memory of one thread.  The launcher checks it before it asks for a compile.
inline int count_cross_phase_values(const SparseProgram& sp) {
  using namespace sparse_detail;
  const int n_ir = (int)sp.ir.size();
  int B = n_ir;
  for (int t = 0; t < n_ir; ++t) if (sp.ir[t].kind == SOP_BSUB) { B = t; break; }
  std::vector<char> fwd_def(sp.n_virtual, 0), back_use(sp.n_virtual, 0);
  for (int t = 0; t < B; ++t) {
    if (sp.ir[t].def >= 0) fwd_def[sp.ir[t].def] = 1;
    for (const Update& u : sp.ir[t].upd) fwd_def[u.dst_new] = 1;
  }
  for (int t = B; t < n_ir; ++t)
    for (int o : sp.ir[t].reads) if (o >= 0) back_use[o] = 1;
  int n = 0;
  for (int v = 0; v < sp.n_virtual; ++v) n += fwd_def[v] && back_use[v];
  return n;
}

inline std::string generate_sparse_kernel_source(const CodegenInput& in, const CodegenOptions& opt = CodegenOptions(),
                                                 CodegenStats* stats_out = nullptr) {
  using namespace codegen_detail;
  using namespace sparse_detail;
  const SparseProgram& sp = *in.sp;
  const std::vector<IrOp>& ir = sp.ir;
  const int n_ir = (int)ir.size();
  CodegenStats st;
  // Real constants live in a __constant__ table: a DFMA/DMUL takes c[bank][offset] as an operand for free,
  // whereas a 64-bit immediate costs two extra instructions every time it is used.
  std::vector<double> ktab;
  std::map<unsigned long long, int> kidx;
  auto lit = [&](double v) -> std::string {
    if (v == 0.0) return "0.0";
    unsigned long long bits;
    memcpy(&bits, &v, 8);
    auto it = kidx.find(bits);
    if (it == kidx.end()) { it = kidx.insert(std::make_pair(bits, (int)ktab.size())).first; ktab.push_back(v); }
    return "KC[" + std::to_string(it->second) + "]";
  };

  // ---- stamped entries -> classes of identical constants ----
  std::vector<int> cls(sp.n_stamp, -1);
  std::vector<int> cls_rep;
  {
    std::map<std::vector<double>, int> seen;
    for (int en = 0; en < sp.n_stamp; ++en) {
      std::vector<double> key = {sp.ent_alpha[en] + sp.ent_jre[en], sp.ent_jim[en], sp.ent_beta[en], sp.ent_gamma[en]};
      auto it = seen.find(key);
      if (it == seen.end()) { it = seen.insert(std::make_pair(key, (int)cls_rep.size())).first; cls_rep.push_back(en); }
      cls[en] = it->second;
    }
  }
  st.n_classes = (int)cls_rep.size();
  const bool named_classes = st.n_classes <= 32;  // otherwise expressions are written where they are used
  bool need_iw = false;
  for (int en = 0; en < sp.n_stamp; ++en) need_iw = need_iw || sp.ent_gamma[en] != 0.0;
  for (int e = 0; e < in.n_ac_elem; ++e) need_iw = need_iw || sp.el_g[e] != 0.0;

  auto im_expr = [&](double b, double g, double ji) -> std::string {  // w*b - g/w + ji
    std::string im;
    if (b != 0.0 && g != 0.0) im = "fma(w, " + lit(b) + ", -" + lit(g) + " * iw)";
    else if (b != 0.0) im = "w * " + lit(b);
    else if (g != 0.0) im = "-" + lit(g) + " * iw";
    else return ji != 0.0 ? lit(ji) : "0.0";
    if (ji != 0.0) im = "(" + im + " + " + lit(ji) + ")";
    return im;
  };
  auto entry_opnd = [&](int en) -> Opnd {
    Opnd o;
    const double re = sp.ent_alpha[en] + sp.ent_jre[en];
    const double b = sp.ent_beta[en], g = sp.ent_gamma[en], ji = sp.ent_jim[en];
    o.re = lit(re); o.re0 = re == 0.0; o.re_lit = true; o.re_c = re;
    o.im0 = b == 0.0 && g == 0.0 && ji == 0.0;
    if (o.im0) o.im = "0.0";
    else if (named_classes) o.im = "q" + std::to_string(cls[en]);
    else o.im = "(" + im_expr(b, g, ji) + ")";
    return o;
  };

  // ---- positions: op index; boundary B = first BSUB ----
  int B = n_ir;
  for (int t = 0; t < n_ir; ++t) if (ir[t].kind == SOP_BSUB) { B = t; break; }
  std::vector<int> deft(sp.n_virtual, -1), last(sp.n_virtual, -1);
  for (int t = 0; t < n_ir; ++t) {
    const IrOp& op = ir[t];
    auto use = [&](int o) { if (o >= 0) last[o] = std::max(last[o], t); };
    for (int o : op.reads) use(o);
    for (const Update& u : op.upd) { use(u.dst_old); use(u.src); deft[u.dst_new] = t; }
    if (op.def >= 0) deft[op.def] = t;
  }
  // element currents: emitted after the BSUB that produces the later of the two node voltages
  std::vector<std::vector<int>> cur_at(n_ir);
  auto xv_of = [&](int node) { return node == 0 ? -1 : sp.x_virtual[node - 1]; };
  for (int e = 0; e < in.n_ac_elem; ++e) {
    int when = -1;
    if (e >= in.v_first) when = deft[sp.x_virtual[in.nn + e - in.v_first]];
    else {
      const int a = xv_of(in.n1[e]), b = xv_of(in.n2[e]);
      if (a >= 0) when = std::max(when, deft[a]);
      if (b >= 0) when = std::max(when, deft[b]);
    }
    if (when < 0) when = n_ir - 1;  // both ends grounded: current is zero, emit at the end
    cur_at[when].push_back(e);
    if (e < in.v_first) {
      const int a = xv_of(in.n1[e]), b = xv_of(in.n2[e]);
      if (a >= 0) last[a] = std::max(last[a], when);
      if (b >= 0) last[b] = std::max(last[b], when);
    }
  }
  // ---- shared-memory placement of the cross-phase values: the longest-lived go to shared memory, registers
  //      keep what the back-substitution consumes first ----
  // Larger programs (a 400-node ladder hands 1,600 values to its back-substitution) would leave the rest to the register
  // allocator, i.e. to kilobytes of local-memory spills: with opt.reg_values >= 0 the longest-lived values beyond
  // registers + shared memory go to a per-thread column of GLOBAL memory instead ([slot][thread]: one coalesced 512-byte
  // store per warp when the value is produced, one such load a few rows before the back-substitution needs it).
  std::vector<int> slot_of(sp.n_virtual, -1), gslot_of(sp.n_virtual, -1);
  int ns = 0, ng = 0;
  {
    std::vector<std::pair<int, int>> saved;  // (-lifetime, v)
    for (int v = 0; v < sp.n_virtual; ++v)
      if (deft[v] >= 0 && deft[v] < B && last[v] >= B) saved.push_back(std::make_pair(-(last[v] - deft[v]), v));
    std::sort(saved.begin(), saved.end());
    st.n_saved = (int)saved.size();
    ns = std::min<int>(std::max(0, opt.smem_slots), (int)saved.size());
    if (opt.reg_values >= 0) ng = std::max(0, (int)saved.size() - ns - opt.reg_values);
    for (int i = 0; i < ng; ++i) gslot_of[saved[i].second] = i;
    for (int i = 0; i < ns; ++i) slot_of[saved[ng + i].second] = i;
    st.smem_slots = ns;
    st.gmem_slots = ng;
    st.smem_bytes = (size_t)ns * opt.block * 16;
  }
  auto soff = [&](int slot) { return std::to_string((long long)slot * opt.block * 16); };
  auto gst = [&](int v, const std::string& name) -> std::string {   // store of a value that lives in the global column
    if (gslot_of[v] < 0) return std::string();
    if (opt.l2_policy) return "    GST(gwp + " + std::to_string(gslot_of[v]) + " * gT, " + name + ");\n";
    return "    __stcg(gwp + " + std::to_string(gslot_of[v]) + " * gT, " + name + ");\n";
  };

  std::string s;
  s.reserve(1 << 20);
  // ---- per-instance stamping: raw element values are loaded a few pivots ahead of their use, element
  //      admittances (simulateAC.ts:36-57) and the ordered entry sums (stamp*Complex.ts) are formed where first
  //      used; the back-substitution phase re-loads what the element currents need ----
  char phase = 'f';
  std::map<std::string, char> defined;   // name -> phase it was last defined in
  auto fresh = [&](const std::string& nm) -> bool {
    auto it = defined.find(nm);
    if (it != defined.end() && it->second == phase) return false;
    defined[nm] = phase;
    return true;
  };
  auto ph = [&]() { return std::string(1, phase); };
  auto raw = [&](int slot) -> std::string {   // value of a slot for this instance
    const int v = in.var_of_slot[slot];
    if (v < 0) return lit(in.values[slot]);
    const std::string nm = "raw" + std::to_string(slot) + ph();
    if (fresh(nm)) s += "    const double " + nm + " = vv[" + std::to_string(v) + "u * vs];\n";
    return nm;
  };
  auto need_elem = [&](int e) {   // defines yr<e><ph> / yi<e><ph> (V: jr / ji)
    const std::string E = std::to_string(e) + ph();
    if (!fresh("el" + E)) return;
    const int ty = in.el_type[e], vi = in.el_vidx[e];
    if (ty == 0) {          // R: Y = 1/R, R <= 0 is the reference's "must be > 0" (:37): dense kernel reports it
      const std::string R = raw(vi);
      if (in.var_of_slot[vi] < 0) s += "    const double yr" + E + " = " + lit(1 / in.values[vi]) + ";\n";
      else s += "    bad = bad || !(" + R + " > 0.0);\n    const double yr" + E + " = rcp_nr(" + R + ");\n";
    } else if (ty == 1) {   // C: Y = j (2 pi f) C  (:43)
      s += "    const double yi" + E + " = w * " + raw(vi) + ";\n";
    } else if (ty == 2) {   // L: Y = 1 / (j (2 pi f) L), zero when |denominator| < EPS (:47-51)
      const std::string L = raw(vi);
      s += "    double yi" + E + ";\n    { const double d = w * " + L + ", dd = d * d; bad = bad || (!(fabs(d) < EPS) && dd < EPS);\n";
      s += "      yi" + E + " = fabs(d) < EPS ? 0.0 : -d * rcp_nr(dd); }\n";
    } else {                // V: phasor acMag * exp(j acPhase)  (Complex.ts:16-19)
      if (in.var_of_slot[vi + 1] < 0 && in.var_of_slot[vi + 2] < 0) {
        const double phs = (in.values[vi + 2] * 3.141592653589793) / 180;
        s += "    const double jr" + E + " = " + lit(in.values[vi + 1] * cos(phs)) + ", ji" + E + " = " + lit(in.values[vi + 1] * sin(phs)) + ";\n";
      } else {
        const std::string mag = raw(vi + 1), deg = raw(vi + 2);
        s += "    double jr" + E + ", ji" + E + ";\n    { double sn, cs; sincos((" + deg + " * 3.141592653589793) / 180, &sn, &cs); jr" + E +
             " = " + mag + " * cs; ji" + E + " = " + mag + " * sn; }\n";
      }
    }
  };
  auto entry_opnd_eager = [&](int en) -> Opnd {
    Opnd o;
    const std::string EN = std::to_string(en) + ph();
    std::string re, im;
    // structure first (no code emitted): which parts exist
    bool has_re = false, has_im = false;
    for (int c = in.ent_ptr[en]; c < in.ent_ptr[en + 1]; ++c) {
      const int cw = in.contrib[c], src = (cw >> 1) & 3, idx = cw >> 3;
      if (src == 2) has_re = true;
      else if (src == 1) { has_re = true; has_im = true; }
      else if (in.el_type[idx] == 0) has_re = true;
      else has_im = true;
    }
    o.re0 = !has_re; o.im0 = !has_im;
    o.re = has_re ? "er" + EN : std::string("0.0");
    o.im = has_im ? "ei" + EN : std::string("0.0");
    if (!fresh("en" + EN)) return o;
    for (int c = in.ent_ptr[en]; c < in.ent_ptr[en + 1]; ++c) {
      const int cw = in.contrib[c], src = (cw >> 1) & 3, idx = cw >> 3;
      const bool neg = cw & 1;
      auto add = [&](std::string& acc, const std::string& t) {
        if (acc.empty()) acc = neg ? "-" + t : t;
        else acc = "(" + acc + (neg ? " - " : " + ") + t + ")";
      };
      if (src == 2) { add(re, "1.0"); continue; }
      need_elem(idx);
      const std::string E = std::to_string(idx) + ph();
      if (src == 1) { add(re, "jr" + E); add(im, "ji" + E); }
      else if (in.el_type[idx] == 0) add(re, "yr" + E);
      else add(im, "yi" + E);
    }
    if (has_re) s += "    const double er" + EN + " = " + re + ";\n";
    if (has_im) s += "    const double ei" + EN + " = " + im + ";\n";
    return o;
  };
  // raw values needed by op t (forward: stamped operands; back: stamped operands and element currents)
  auto raw_slots_of_entry = [&](int en, std::vector<int>& out) {
    for (int c = in.ent_ptr[en]; c < in.ent_ptr[en + 1]; ++c) {
      const int cw = in.contrib[c], src = (cw >> 1) & 3, idx = cw >> 3;
      if (src == 2) continue;
      const int vi = in.el_vidx[idx];
      if (in.el_type[idx] == 3) { out.push_back(vi + 1); out.push_back(vi + 2); } else out.push_back(vi);
    }
  };

  auto opnd = [&](int o) -> Opnd {
    Opnd r;
    if (o == kNoOperand) { r.re = r.im = "0.0"; r.re0 = r.im0 = true; r.re_lit = true; return r; }
    if (o < 0) return in.eager ? entry_opnd_eager(~o) : entry_opnd(~o);
    const std::string nm = "v" + std::to_string(o);
    r.re = nm + ".x"; r.im = nm + ".y";
    return r;
  };

  s += "#define BLOCK " + std::to_string(opt.block) + "\n";
  s += "extern \"C\" __global__ void __launch_bounds__(BLOCK, " + std::to_string(opt.min_blocks) + ") spicey_sparse_jit(JitArgs a) {\n";
  s += "  if (a.p_count <= 0) return;\n";
  // long sweeps: the CTAs start in four phases, so that the SMs do not all reach their store-heavy
  // back-substitution at the same moment (-2 % on cfg 2)
  if (opt.stagger_ns > 0)
    s += "  { const unsigned phs = blockIdx.x & 3u; if (phs && a.p_count > 16ll * gridDim.x * BLOCK) __nanosleep(phs * " +
         std::to_string(opt.stagger_ns) + "u); }\n";
  // Several CTAs per SM: the back-substitution issues every result store of a point and is bound by the SM's store
  // port (32 B/clk measured, tools/micro/write_bw.cu) while the elimination stores nothing, so with all warps of an SM in
  // the same phase the port idles for the first 40 % of an iteration and throttles the rest.  The second wave of the
  // grid — CTA i + grid/min_blocks shares an SM with CTA i on B200 (measured: delaying the odd CTAs instead changes
  // nothing) — starts half an iteration late, so that one CTA eliminates while the other stores.
  if (opt.antiphase_ns > 0 && opt.min_blocks >= 2) {
    const std::string mb = std::to_string(opt.min_blocks) + "u";
    s += "  if (a.p_count > 4ll * gridDim.x * BLOCK) { const unsigned wv = blockIdx.x / max(1u, gridDim.x / " + mb + "); if (wv) __nanosleep(wv * " +
         std::to_string(opt.antiphase_ns) + "u); }\n";
  }
  s += "  const unsigned ld = a.series_ld ? (unsigned)a.series_ld : 1u;\n";
  s += "  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 16u;\n";
  s += "  const long long stride = (long long)gridDim.x * BLOCK, plast = a.p_count - 1;\n";
  if (ng > 0) s += "  const size_t gT = (size_t)stride;\n  double2* const gw = a.work + (size_t)blockIdx.x * BLOCK + threadIdx.x;\n";
  const bool hinted = ng > 0 && opt.l2_policy;
  if (hinted)
    s += "  unsigned long long pol_keep, pol_stream;\n"
         "  asm(\"createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\" : \"=l\"(pol_keep));\n"
         "  asm(\"createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\" : \"=l\"(pol_stream));\n";
  if (!in.eager) s += "  double fnext = a.freqs[min((long long)blockIdx.x * BLOCK + threadIdx.x, plast)];\n";
  // block-uniform trip count: lanes past the end solve the last point again and store nothing
  s += "  for (long long base = (long long)blockIdx.x * BLOCK; base < a.p_count; base += stride) {\n";
  s += "    const long long p = base + threadIdx.x;\n    const bool valid = p < a.p_count;\n";
  if (in.eager) {
    // point index = instance * n_freq + frequency index (the layout of spicey_ac_solve for sweeps)
    s += "    const long long gp = a.p_begin + min(p, plast), inst = gp / a.n_freq;\n";
    s += "    const double w = 6.283185307179586 * a.freqs[gp - inst * a.n_freq];\n";
    s += "    const double* __restrict__ vv = a.var_values + inst;   // swept slot v of this instance: vv[v * n_inst]\n";
    s += "    const size_t vs = (size_t)a.n_inst;\n";
  } else {
    s += "    const double w = 6.283185307179586 * fnext;\n";
    s += "    fnext = a.freqs[min(p + stride, plast)];\n";
    if (need_iw) s += "    const double iw = 1.0 / w;\n";
  }
  // Lanes past the end of the sweep solve the LAST point again (same frequency, same instructions, so the same bits)
  // and store it to the last point's rows: a benign duplicate write instead of a predicate, which keeps every result
  // store unconditional and the back-substitution one basic block (no BSSY / BRA / BSYNC around each row's stores).
  s += "    const long long pc = min(p, plast);\n";
  // (two opaque copies of the column's base, one per phase: the compiler otherwise keeps every slot's 64-bit address in a
  //  register from the store in the elimination to the load in the back-substitution — hundreds of them, all spilled)
  if (ng > 0) s += "    double2 *gwp = gw, *gwq = gw; asm volatile(\"\" : \"+l\"(gwp), \"+l\"(gwq));\n";
  s += "    char* const xb = (char*)(a.series_ld ? a.x + pc : a.x + pc * a.n);\n";
  if (opt.with_ielem) s += "    char* const ib = (char*)(a.series_ld ? a.ielem + pc : a.ielem + pc * a.n_ac_elem);\n";
  s += "    bool ok = true, bad = false;\n    double mp, m, inv;\n    double2 r, fm;\n";
  if (!in.eager)
    for (double L : sp.ind_L)  // inductor guards of simulateAC.ts:47-51 are value dependent: dense kernel decides
      s += "    { const double d = w * " + lit(L) + "; bad = bad || fabs(d) < EPS || d * d < EPS; }\n";
  if (named_classes && !in.eager)
    for (int c = 0; c < st.n_classes; ++c) {
      const int en = cls_rep[c];
      if (sp.ent_beta[en] != 0.0 || sp.ent_gamma[en] != 0.0 || sp.ent_jim[en] != 0.0)
        s += "    const double q" + std::to_string(c) + " = " + im_expr(sp.ent_beta[en], sp.ent_gamma[en], sp.ent_jim[en]) + ";\n";
    }

  // a - f*p, component expressions; f is the double2 named `fn`
  auto submul = [&](const Opnd& a, const std::string& fn, const Opnd& p, std::string& re, std::string& im) {
    // re = a.re - f.x p.re + f.y p.im ; im = a.im - f.x p.im - f.y p.re
    std::string inner_re = a.re, inner_im = a.im;
    if (!p.im0) inner_re = "fma(" + fn + ".y, " + p.im + ", " + a.re + ")";
    if (!p.re0) inner_im = "fma(-" + fn + ".y, " + p.re + ", " + a.im + ")";
    re = p.re0 ? inner_re : "fma(-" + fn + ".x, " + p.re + ", " + inner_re + ")";
    im = p.im0 ? inner_im : "fma(-" + fn + ".x, " + p.im + ", " + inner_im + ")";
  };
  // a * b, component expressions
  auto cmul = [&](const Opnd& a, const Opnd& b, std::string& re, std::string& im) {
    auto prod = [&](const std::string& x, bool x0, const std::string& y, bool y0) { return (x0 || y0) ? std::string() : x + " * " + y; };
    // re = a.re b.re - a.im b.im ; im = a.re b.im + a.im b.re
    const std::string t1 = prod(a.im, a.im0, b.im, b.im0), t2 = prod(a.im, a.im0, b.re, b.re0);
    if (a.re0 || b.re0) re = t1.empty() ? "0.0" : "-(" + t1 + ")";
    else re = t1.empty() ? a.re + " * " + b.re : "fma(" + a.re + ", " + b.re + ", -(" + t1 + "))";
    if (a.re0 || b.im0) im = t2.empty() ? "0.0" : t2;
    else im = t2.empty() ? a.re + " * " + b.im : "fma(" + a.re + ", " + b.im + ", " + t2 + ")";
  };
  auto nrm = [&](const Opnd& a) -> std::string {
    if (a.re0 && a.im0) return "0.0";
    if (a.im0) return a.re + " * " + a.re;
    if (a.re0) return a.im + " * " + a.im;
    return "fma(" + a.re + ", " + a.re + ", " + a.im + " * " + a.im + ")";
  };
  auto off = [&](int k) { return "(size_t)ld * " + std::to_string(16ll * k) + "u"; };

  // ---- results: predicated stores, emitted as soon as the values exist ----
  struct Out { bool cur; int k; std::string re, im; };
  std::vector<Out> pending;
  auto flush_outputs = [&]() {
    for (const Out& o : pending)
      if (hinted) s += std::string("    RST(") + (o.cur ? "ib" : "xb") + " + " + off(o.k) + ", " + o.re + ", " + o.im + ");\n";
      else s += std::string("    *(double2*)(") + (o.cur ? "ib" : "xb") + " + " + off(o.k) + ") = D2(" + o.re + ", " + o.im + ");\n";
    pending.clear();
  };

  auto emit_currents = [&](int t) {
    if (cur_at[t].empty() || !opt.with_ielem) return;
    for (int e : cur_at[t]) {
      std::string re, im;
      if (e >= in.v_first) {
        const Opnd x = opnd(sp.x_virtual[in.nn + e - in.v_first]);
        re = x.re; im = x.im;
      } else {
        const int va = xv_of(in.n1[e]), vb = xv_of(in.n2[e]);
        Opnd d;  // v1 - v2
        if (va >= 0 && vb >= 0) {
          const Opnd A = opnd(va), Bq = opnd(vb);
          d.re = "(" + A.re + " - " + Bq.re + ")"; d.im = "(" + A.im + " - " + Bq.im + ")";
        } else if (va >= 0) d = opnd(va);
        else if (vb >= 0) { const Opnd Bq = opnd(vb); d.re = "(0.0 - " + Bq.re + ")"; d.im = "(0.0 - " + Bq.im + ")"; }
        else { d.re = d.im = "0.0"; d.re0 = d.im0 = true; }
        Opnd y;  // element admittance ya + j(w*yb - yg/w)
        if (in.eager) {
          need_elem(e);
          const std::string E = std::to_string(e) + ph();
          if (in.el_type[e] == 0) { y.re = "yr" + E; y.im = "0.0"; y.im0 = true; }
          else { y.re = "0.0"; y.re0 = true; y.im = "yi" + E; }
        } else {
          y.re = lit(sp.el_a[e]); y.re0 = sp.el_a[e] == 0.0;
          y.im0 = sp.el_b[e] == 0.0 && sp.el_g[e] == 0.0;
          y.im = y.im0 ? "0.0" : "(" + im_expr(sp.el_b[e], sp.el_g[e], 0.0) + ")";
        }
        cmul(y, d, re, im);
      }
      pending.push_back(Out{true, e, re, im});
    }
  };

  // per-instance stamping: value slots op t needs, so that their loads can be issued a few steps ahead
  std::vector<std::vector<int>> slots_at(in.eager ? n_ir : 0);
  if (in.eager)
    for (int t = 0; t < n_ir; ++t) {
      const IrOp& op = ir[t];
      auto use = [&](int o) { if (o < 0 && o != kNoOperand) raw_slots_of_entry(~o, slots_at[t]); };
      for (int o : op.reads) use(o);
      for (const Update& u : op.upd) { use(u.dst_old); use(u.src); }
      if (t >= B && opt.with_ielem)
        for (int e : cur_at[t]) if (e < in.v_first) slots_at[t].push_back(in.el_vidx[e]);
    }
  auto prefetch_from = [&](int t) {   // called at a PIVOT / BSUB: loads for this and the next prefetch_steps steps
    if (!in.eager) return;
    int steps = 0;
    for (int q = t; q < n_ir && (t < B ? q < B : true); ++q) {
      if (q > t && (ir[q].kind == SOP_PIVOT || ir[q].kind == SOP_BSUB) && ++steps > opt.prefetch_steps) break;
      for (int slot : slots_at[q]) raw(slot);
    }
  };

  // global-column values: loaded opt.gmem_ahead back-substitution rows before the row that reads them
  std::vector<char> g_loaded(sp.n_virtual, 0);
  auto gload_ahead = [&](int t) {
    if (ng == 0) return;
    int rows = 0;
    for (int q = t; q < n_ir; ++q) {
      if (q > t && ir[q].kind == SOP_BSUB && ++rows > opt.gmem_ahead) break;
      for (int o : ir[q].reads)
        if (o >= 0 && gslot_of[o] >= 0 && !g_loaded[o]) {
          g_loaded[o] = 1;
          if (opt.l2_policy) s += "    double2 g" + std::to_string(o) + "; GLD(g" + std::to_string(o) + ", gwq + " + std::to_string(gslot_of[o]) + " * gT);\n";
          else s += "    const double2 g" + std::to_string(o) + " = __ldcg(gwq + " + std::to_string(gslot_of[o]) + " * gT);\n";
        }
    }
  };

  int n_piv = 0, n_bs = 0;
  for (int t = 0; t < n_ir; ++t) {
    const IrOp& op = ir[t];
    if (t == B) phase = 'b';
    if (op.kind == SOP_BSUB) gload_ahead(t);
    if (op.kind == SOP_PIVOT || op.kind == SOP_BSUB) prefetch_from(t);
    if (t == B) {
      // Every pivot has been verified.  A system whose pivot order differs from the pilot's, or that trips a
      // guard of the reference (singular, Complex.div, inductor), goes to the dense kernel, which also
      // reports the exact status; whatever the rows below write for it is overwritten there.
      s += "    { const bool fb = bad || !ok;\n";
      s += "      if (valid) { a.status[p] = fb ? -1 : 0; if (fb) a.fb_list[atomicAdd(a.fb_count, 1)] = p; } }\n";
    }
    if (op.kind == SOP_PIVOT) {
      if (opt.sync_every > 0 && n_piv % opt.sync_every == 0) s += "    __syncthreads();\n";
      ++n_piv;
      const Opnd ap = opnd(op.reads[op.pidx]);
      s += "    mp = " + nrm(ap) + ";\n";
      for (int c = 0; c < (int)op.reads.size(); ++c) {
        if (c == op.pidx) continue;
        s += "    m = " + nrm(opnd(op.reads[c])) + "; ok = ok && " + (c < op.pidx ? "(m < mp)" : "!(m > mp)") + ";\n";
      }
      s += "    ok = ok && (mp >= EPS);   // singular / Complex.div guard (or NaN): the dense kernel reports which\n";
      s += "    inv = rcp_nr(mp);\n";
      const std::string v = "v" + std::to_string(op.def);
      s += "    const double2 " + v + " = D2(" + (ap.re0 ? std::string("0.0") : ap.re + " * inv") + ", " +
           (ap.im0 ? std::string("0.0") : "-" + ap.im + " * inv") + ");\n";
      s += "    r = " + v + ";\n";
      if (slot_of[op.def] >= 0) s += "    SMST(" + soff(slot_of[op.def]) + ", " + v + ");\n";
      s += gst(op.def, v);
    } else if (op.kind == SOP_ELIM) {
      std::string re, im;
      Opnd rr; rr.re = "r.x"; rr.im = "r.y";
      cmul(opnd(op.reads[0]), rr, re, im);
      // solveComplex.ts:46 skips the row when |f| < EPS: a zero multiplier leaves every updated value as it was
      // (strongly attenuating circuits do produce such multipliers at the far end of a sweep)
      s += "    fm = D2(" + re + ", " + im + "); if (fma(fm.x, fm.x, fm.y * fm.y) < THR) fm = D2(0.0, 0.0);\n";
      for (const Update& u : op.upd) {
        const Opnd a = opnd(u.dst_old), pq = opnd(u.src);
        const std::string v = "v" + std::to_string(u.dst_new);
        submul(a, "fm", pq, re, im);
        s += "    const double2 " + v + " = D2(" + re + ", " + im + ");\n";
        if (slot_of[u.dst_new] >= 0) s += "    SMST(" + soff(slot_of[u.dst_new]) + ", " + v + ");\n";
        s += gst(u.dst_new, v);
      }
    } else {
      if (opt.sync_every > 0 && n_bs % opt.sync_every == 0) s += "    __syncthreads();\n";
      ++n_bs;
      // x_i = (b_i - sum u_ij x_j) * (1/u_ii); the most recently produced x_j is applied last
      std::vector<std::pair<int, int>> terms;  // (def time of x_j, index into reads)
      for (size_t q = 2; q + 1 < op.reads.size(); q += 2) terms.push_back(std::make_pair(deft[op.reads[q + 1]], (int)q));
      std::sort(terms.begin(), terms.end());
      // shared-memory residents read by this op are loaded under a per-use name
      std::map<int, std::string> local;
      for (int o : op.reads)
        if (o >= 0 && slot_of[o] >= 0 && !local.count(o)) {
          const std::string nm = "s" + std::to_string(o) + "_" + std::to_string(t);
          s += "    double2 " + nm + "; SMLD(" + nm + ", " + soff(slot_of[o]) + ");\n";
          local[o] = nm;
        }
      auto bop = [&](int o) -> Opnd {
        if (o >= 0 && local.count(o)) { Opnd r2; r2.re = local[o] + ".x"; r2.im = local[o] + ".y"; return r2; }
        if (o >= 0 && gslot_of[o] >= 0) { Opnd r2; r2.re = "g" + std::to_string(o) + ".x"; r2.im = "g" + std::to_string(o) + ".y"; return r2; }
        return opnd(o);
      };
      Opnd acc = bop(op.reads[0]);
      int k = 0;
      for (const auto& tm : terms) {
        const Opnd u = bop(op.reads[tm.second]), xj = bop(op.reads[tm.second + 1]);
        // acc - u * x_j, u in the role of the structurally known factor
        std::string re, im;
        const std::string xn = "xj" + std::to_string(t) + "_" + std::to_string(k);
        s += "    const double2 " + xn + " = D2(" + xj.re + ", " + xj.im + ");\n";
        submul(acc, xn, u, re, im);
        const std::string an = "ac" + std::to_string(t) + "_" + std::to_string(k);
        s += "    const double2 " + an + " = D2(" + re + ", " + im + ");\n";
        acc = Opnd(); acc.re = an + ".x"; acc.im = an + ".y";
        ++k;
      }
      std::string re, im;
      cmul(acc, bop(op.reads[1]), re, im);
      const std::string v = "v" + std::to_string(op.def);
      s += "    const double2 " + v + " = D2(" + re + ", " + im + ");\n";
      pending.push_back(Out{false, op.var, v + ".x", v + ".y"});
    }
    if (t >= B) { emit_currents(t); flush_outputs(); }
  }
  s += "  }\n";
  s += "}\n";
  {
    std::string head = sparse_jit_prelude();
    head += "__constant__ double KC[" + std::to_string(std::max<size_t>(1, ktab.size())) + "] = {";
    for (size_t i = 0; i < ktab.size(); ++i) head += (i ? ", " : "") + hexlit(ktab[i]);
    if (ktab.empty()) head += "0.0";
    head += "};\n";
    s = head + s;
  }
  if (stats_out) *stats_out = st;
  return s;
}

}  // namespace spicey
