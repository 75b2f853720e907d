// C ABI of the B200-native batched MNA engine (include/spicey_native.h).
//
// Host side of the hot path: validates the flat element table, builds the per-topology
// plan (gather-form stamping lists + structural row masks), uploads it, picks the kernel
// tier, shards the batch axis over the handle's devices in contiguous ranges and, for the
// host-buffer entry points, pipelines kernel chunks against device->host copies on a
// second stream.  No CPU compute path exists here: without a CUDA device every solve
// entry point fails with SPICEY_ERR_NO_DEVICE.
#include "../../include/spicey_native.h"

#include <cuda.h>   // CUtensorMap types only: cuTensorMapEncodeTiled is reached through cudaGetDriverEntryPoint (no libcuda link)

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ac_kernels.cuh"
#include "ac_sparse.cuh"
#include "ac_warp.cuh"
#include "band_kernel_embed.h"
#include "tile_kernel_embed.h"
#include "warp_lu_kernel_embed.h"
#include "band_plan.h"
#include "tile_plan.h"
#include "host_plan.h"
#include "jit_runtime.h"
#include "sparse_codegen.h"
#include "tran_codegen.h"
#include "tran_kernels.cuh"
#include "tran_small.cuh"

using namespace spicey;
using namespace spicey::host;

namespace {

// ---------------------------------------------------------------------------------
// Device context
struct Buffer {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return SPICEY_SUCCESS;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CUDA_TRY(cudaMalloc(&p, want));
    cap = want;
    return SPICEY_SUCCESS;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct DeviceCtx {
  int dev = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t compute = nullptr, copy = nullptr;
  Buffer plan, scratch, in0, in1, in2, out_x[2], out_i[2], out_s[2], aux0, aux1, tr_iters[2], tr_state[2];
  // sparse AC path: cached program (keyed by the element table), workspace, fallback list
  Buffer sp_blob, sp_work, sp_fb;
  Buffer sp_cnt;              // two counters: [0] fallback solves of the current call, [1] since the cached program was built
  Buffer tl_fb;               // second-level fallback list (points of a fallback list that trip an inductor guard in the tile tier)
  long long sp_launched = 0;  // points launched through the cached program since it was built
  int sp_adapt = 0;           // 0 undecided, 1 the pilot's pivot order holds (program tiers), 2 it does not (dense tiers)
  SparseProgram sp;
  SparseArgs sp_args;
  uint64_t sp_key = 0;
  long long sp_T = 0;
  bool sp_unified = false;
  bool sp_eager = false;
  // straight-line (NVRTC-compiled) variant of the program, when it was built
  // compiled straight-line variants of the sparse program: [0] without, [1] with element currents
  struct JitVariant {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    uint64_t key = 0;   // plan key the module was compiled for (0 = none)
    bool failed = false;
    size_t smem_bytes = 0;
    int block = 0, min_blocks = 1;
    int gmem_slots = 0;   // double2 per thread of the global column (JitArgs.work)
  } sp_jit[2];
  Buffer sp_jit_work;
  // Launch shape of the compiled kernel, measured on cfg2 (tools/jit_sweep.py): one CTA of 6 warps per SM with
  // 255 registers per thread and 75 shared-memory slots per thread for the factor values (0.76 ms per 1e6
  // points); 5 warps x 90 slots: 0.81 ms, 4 x 113: 1.03 ms, 8 x 55 (spills): 1.0 ms.  A __syncthreads every
  // 4 pivots keeps the warps of a CTA on the same instruction-cache lines (-8 %).  Tried and removed: a bulk-copy
  // (TMA) epilogue (one cp.async.bulk issue costs its warp ~90 cycles, tools/micro/bulk_store.cu: 2.3 ms) and
  // reserving a stored value's registers with an empty asm while the store drains (no gain).  Overridable for experiments: SPICEY_JIT_CFG=block,minb,slots[,sync,prefetch].
  // Since then: the back-substitution issues all 3,080 B of a point's results and is bound by the SM's store port
  // (32 B/clk, tools/micro/write_bw.cu) while the elimination stores nothing, so the shape is now TWO CTAs of 3 warps
  // per SM started half an iteration apart (CTA i and i + grid/2 share an SM): one eliminates while the other stores.
  // 1 x 192 in phase: 0.67 ms; 2 x 96 in phase: 0.67; 2 x 96 half an iteration apart: 0.61 - 0.64 ms.
  int sp_jit_block = 96, sp_jit_minb = 2, sp_jit_slots = 75, sp_jit_sync = 4;
  int sp_jit_stagger = 0;           // ns between four start phases of the CTAs (helped the 1 x 192 shape by 2 %)
  double sp_jit_antiphase_ns_per_op = 10.4;   // start delay of the second wave: half an iteration, ~20.8 ns per micro-op
  // per-instance (eager) stamping keeps element values in flight: fewer threads, more shared memory each.  cfg2mc:
  // 2 CTAs x 64 threads x 112 slots 1.18 ms, 1 x 128 x 113 1.28 ms, 2 x 96 x 75 1.57, 1 x 160 x 90 1.50 (a start delay
  // between the two CTAs changes nothing here: the kernel is bound by its arithmetic, not by the store port);
  // prefetch 8 pivots ahead (4: +5 %, 2: +12 %).  Since the global column exists (sparse_codegen.h) the values that do not
  // fit beside the element values in flight go there instead of asking for more shared memory per thread: 6 warps per SM
  // (2 x 96 x 75 slots + 40 register values + 14 in the column) 0.98 ms against 1.17 ms for the 4 warps of 2 x 64 x 112
  // (profiles/r3s_cfg2mc_shapes.txt; 3 x 64 x 75: the same, 4 x 64 x 56: 1.00, 6 x 32 x 75: 1.14)
  int sp_jit_minb_eager = 2, sp_jit_stagger_eager = 0;
  int sp_jit_block_eager = 96, sp_jit_slots_eager = 75, sp_jit_prefetch = 8;
  double sp_jit_compile_ms = 0;
  uint64_t sp_jit_fit_key = 0;   // sparse program the fit check below was made for
  bool sp_jit_fits = false;
  bool sp_jit_column = false;   // the compiled kernel of this topology needs its global column
  std::string sp_jit_note;
  // warp-cooperative form of the sparse program (warp_program.h): large programs, one warp per system
  WarpProgram wp;
  uint64_t wp_key = 0;        // sparse-program key the warp program was lowered from (0 = none)
  bool wp_valid = false;
  bool wp_chainlike = false;   // fewer than 32 updates per pivot step on average: the warp tier is only used when forced
  Buffer wp_blob, wp_work;
  WarpArgs wp_args;
  // banded + bordered tier (band_plan.h / band_kernel.cuh): plan of the last topology, its tables, the workspace
  BandPlan bp;
  uint64_t bp_key = 0;        // sparse-program key the band plan was built for (0 = none)
  bool bp_valid = false;
  Buffer bp_blob, bp_work;
  double bp_order_dev = 0;   // band_order_deviation of the last renumbered plan
  struct BandDev { const void *tab, *flags, *newvar, *el_rec; } bp_dev = {nullptr, nullptr, nullptr, nullptr};
  struct BandJit {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    uint64_t key = 0;
    bool failed = false;
    int warps = 0, minb = 0;
    int umode = 0;            // BAND_UMODE of band_kernel.cuh: how U leaves for the workspace (2 = TMA tensor store)
    size_t smem_bytes = 0;
  } band_jit[2];              // [0] without, [1] with element currents
  bool band_tma_refused = false;   // cuTensorMapEncodeTiled unavailable or refused the workspace: BAND_UMODE 2 -> 0
  // dense register-tile tier (tile_kernel.cuh): per-entry gather lists of the last topology and the compiled kernels
  Buffer tl_blob;
  uint64_t tl_key = 0;        // plan key the entry lists were uploaded for (0 = none)
  struct TileDev {
    const void *ent_rc = nullptr, *ent_ptr = nullptr, *contrib = nullptr, *ctab = nullptr, *el_rec = nullptr, *ind_L = nullptr;
    const void* wtab = nullptr;   // the constants in the one-warp-per-system layout [column][lane = row] (Nvar <= 32)
    int n_ent = 0, n_ind = 0;
    bool has_const = false, rc_only = false;
  } tl_dev;
  struct TileJit {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    uint64_t key = 0;
    bool failed = false;
    int n = 0, tr = 0, tc = 0, warps = 0, minb = 0;
  } tile_jit[16];             // by variant: bit 0 element currents, bit 1 constant tables, bit 2 (alpha, beta)-only tables,
                              // bit 3 the one-warp-per-system kernel (Nvar <= 32)
  std::string tl_note;
  double sp_pilot_f = 0;      // frequency of the pilot point the cached programs were built from
  double sp_f_lo = 0, sp_f_hi = 0;   // first / last frequency of the call that built them (0: unknown)
  uint64_t plan_up_key = 0;   // plan currently resident in `plan` (its device pointers are in plan_dp)
  DevPlan plan_dp;
  JitVariant tr_jit[2];   // compiled transient kernel of the last topology: [0] without, [1] with element currents
  std::vector<int4> sp_code_scaled;  // program with slot operands scaled by the pool strides
  bool sp_valid = false;
  std::vector<cudaEvent_t> events;
  cudaEvent_t get_event(size_t i) {
    while (events.size() <= i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      events.push_back(e);
    }
    return events[i];
  }
};

template <typename T> size_t push_blob(std::vector<unsigned char>& blob, const std::vector<T>& v) {
  size_t off = (blob.size() + 15) & ~(size_t)15;
  blob.resize(off + sizeof(T) * v.size());
  if (!v.empty()) memcpy(blob.data() + off, v.data(), sizeof(T) * v.size());
  return off;
}

// Uploads the plan into ctx.plan on `stream` and fills the device-pointer view.
int upload_plan(DeviceCtx& ctx, const HostPlan& hp, cudaStream_t stream, DevPlan& dp,
                std::vector<unsigned char>& blob) {
  uint64_t key = plan_key(hp);
  key = fnv1a(key, hp.var_of_slot.data(), sizeof(int) * hp.var_of_slot.size());
  if (!key) key = 1;
  if (ctx.plan_up_key == key) {  // same netlist as the previous call on this device: nothing to copy
    dp = ctx.plan_dp;
    return SPICEY_SUCCESS;
  }
  ctx.plan_up_key = 0;
  blob.clear();
  size_t o_ends = push_blob(blob, hp.ends), o_meta = push_blob(blob, hp.meta);
  size_t o_sidx = push_blob(blob, hp.state_idx), o_val = push_blob(blob, hp.values);
  size_t o_vos = push_blob(blob, hp.var_of_slot);
  size_t o[2][5];
  const HostGather* gs[2] = {&hp.ac, &hp.tran};
  for (int k = 0; k < 2; ++k) {
    o[k][0] = push_blob(blob, gs[k]->row_ptr);
    o[k][1] = push_blob(blob, gs[k]->ent_col);
    o[k][2] = push_blob(blob, gs[k]->ent_ptr);
    o[k][3] = push_blob(blob, gs[k]->contrib);
    o[k][4] = push_blob(blob, gs[k]->rowmask);
  }
  int rc = ctx.plan.ensure(blob.size());
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(ctx.plan.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, stream));
  unsigned char* b = (unsigned char*)ctx.plan.p;
  memset(&dp, 0, sizeof(dp));
  dp.nn = hp.nn; dp.nV = hp.nV; dp.nvar = hp.nvar; dp.n_elem = hp.n_elem; dp.n_values = hp.n_values;
  dp.n_ac_elem = hp.n_ac_elem; dp.n_state = hp.n_state; dp.MW = hp.MW;
  for (int k = 0; k < 8; ++k) dp.off[k] = hp.off[k];
  dp.ends = (const int4*)(b + o_ends);
  dp.meta = (const int2*)(b + o_meta);
  dp.state_idx = (const int*)(b + o_sidx);
  dp.values = (const double*)(b + o_val);
  dp.var_of_slot = (const int*)(b + o_vos);
  GatherPlan* gp[2] = {&dp.ac, &dp.tran};
  for (int k = 0; k < 2; ++k) {
    gp[k]->row_ptr = (const int*)(b + o[k][0]);
    gp[k]->ent_col = (const int*)(b + o[k][1]);
    gp[k]->ent_ptr = (const int*)(b + o[k][2]);
    gp[k]->contrib = (const int*)(b + o[k][3]);
    gp[k]->rowmask = (const unsigned*)(b + o[k][4]);
  }
  // the cached copy may be used from another stream by a later call: make sure it has landed (cache misses only)
  CUDA_TRY(cudaStreamSynchronize(stream));
  ctx.plan_dp = dp;
  ctx.plan_up_key = key;
  return SPICEY_SUCCESS;
}

int round32(int n) { return std::max(32, (n + 31) / 32 * 32); }

// The fallback counters of a device ([0] current call, [1] since the cached sparse program was built): allocated once,
// [0] cleared at the start of every call.
int reset_call_counters(DeviceCtx& ctx, cudaStream_t stream) {
  if (!ctx.sp_cnt.p) {
    int rc = ctx.sp_cnt.ensure(16);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(ctx.sp_cnt.p, 0, 16, stream));
  } else {
    CUDA_TRY(cudaMemsetAsync(ctx.sp_cnt.p, 0, 8, stream));
  }
  return SPICEY_SUCCESS;
}

}  // namespace

struct spicey_handle {
  std::vector<DeviceCtx> devs;
  spicey_stats stats;
  std::vector<unsigned char> blob;  // plan staging (kept alive until the stream has consumed it)
  // host plan of the last element table: repeated calls on one netlist (a bench loop, chunked sweeps, a
  // Monte-Carlo driver) skip the rebuild (~0.1 ms of std::map work that otherwise sits in front of every launch)
  uint64_t hp_key = 0;
  HostPlan hp;
};

namespace {

// Restores the caller's current device when a multi-device call returns (a handle inside a torch process must not
// leave later allocations on another GPU).
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// On an early (error) return of a host-buffer call, waits for the copies already queued on the handle's devices:
// they write into the caller's buffers, which the caller is free to release as soon as the call has returned.
struct DrainOnError {
  spicey_handle* h;
  bool ok = false;
  explicit DrainOnError(spicey_handle* hh) : h(hh) {}
  ~DrainOnError() {
    if (ok) return;
    for (auto& c : h->devs) {
      if (cudaSetDevice(c.dev) != cudaSuccess) continue;
      cudaStreamSynchronize(c.copy);
      cudaStreamSynchronize(c.compute);
    }
    cudaGetLastError();
  }
};

// build_plan() through the handle's one-entry cache.
int cached_plan(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep) {
  const uint64_t key = table_key(table, sweep);
  if (key && key == h->hp_key) return SPICEY_SUCCESS;
  h->hp_key = 0;
  h->hp = HostPlan();
  int rc = build_plan(table, sweep, h->hp);
  if (rc) return rc;
  h->hp_key = key;
  return SPICEY_SUCCESS;
}

double now_ms();

struct ConstProgOwner {
  std::mutex mu;
  uint64_t key = 0;
  long long T = 0;
  cudaEvent_t ev = nullptr;
};
ConstProgOwner& const_owner(int dev) {
  static ConstProgOwner owners[64];
  return owners[dev & 63];
}


// ---------------------------------------------------------------------------------
// NVRTC (loaded lazily with dlopen so that the library itself has no link-time dependency on it):
// compiles the straight-line kernel of sparse_codegen.h for sm_100a.
struct JitArgs {   // must match sparse_jit_prelude()
  const double* freqs; long long p_count;
  double2* x; double2* ielem; int* status; long long series_ld;
  long long* fb_list; int* fb_count; int n; int n_ac_elem;
  const double* var_values; long long n_inst; long long n_freq; long long p_begin;   // per-instance stamping
  double2* work;   // global column of the factor values beyond registers + shared memory (CodegenStats.gmem_slots)
};

// Host half of the sparse path: per-entry constants, pilot matrix and the program itself (no device needed).
// Leaves sp.ok=false when the sparse path does not apply (R<=0, singular pilot).
// Per-entry constants of a plain frequency sweep: entry = alpha + Re J + j (w beta - gamma / w + Im J), the
// contributions of simulateAC.ts:36-57 summed in the reference's stamping order; per element Y = ya + j (w yb - yg / w).
// false: an R <= 0 (simulateAC.ts:37) — the dense kernels report it per point.
bool entry_constants(const HostPlan& hp, SparseProgram& sp) {
  const HostGather& G = hp.ac;
  const int n_ent = (int)G.ent_col.size();
  sp.ent_alpha.assign(n_ent, 0.0); sp.ent_beta.assign(n_ent, 0.0); sp.ent_gamma.assign(n_ent, 0.0);
  sp.ent_jre.assign(n_ent, 0.0); sp.ent_jim.assign(n_ent, 0.0);
  for (int e = 0; e < hp.n_ac_elem; ++e)
    if (hp.meta[e].x == ELEM_R && !(hp.values[hp.meta[e].y] > 0)) return false;
  for (int en = 0; en < n_ent; ++en) {
    for (int c = G.ent_ptr[en]; c < G.ent_ptr[en + 1]; ++c) {
      const int w = G.contrib[c], src = (w >> 1) & 3, idx = w >> 3;
      const double sgn = (w & 1) ? -1.0 : 1.0;
      if (src == SRC_ONE) { sp.ent_alpha[en] += sgn; continue; }
      const int ty = hp.meta[idx].x;
      const double* v = &hp.values[hp.meta[idx].y];
      if (src == SRC_J) {  // source phasor, Complex.fromPolar (Complex.ts:16-19)
        const double ph = (v[2] * kPi) / 180;
        sp.ent_jre[en] += sgn * (v[1] * cos(ph));
        sp.ent_jim[en] += sgn * (v[1] * sin(ph));
      } else if (ty == ELEM_R) sp.ent_alpha[en] += sgn * (1 / v[0]);
      else if (ty == ELEM_C) sp.ent_beta[en] += sgn * v[0];
      else if (ty == ELEM_L) sp.ent_gamma[en] += sgn * (1 / v[0]);
    }
  }
  sp.el_a.assign(hp.n_ac_elem, 0.0); sp.el_b.assign(hp.n_ac_elem, 0.0); sp.el_g.assign(hp.n_ac_elem, 0.0);
  sp.ind_L.clear();
  for (int e = 0; e < hp.n_ac_elem; ++e) {
    const int ty = hp.meta[e].x;
    const double v = hp.values[hp.meta[e].y];
    if (ty == ELEM_R) sp.el_a[e] = 1 / v;
    else if (ty == ELEM_C) sp.el_b[e] = v;
    else if (ty == ELEM_L) { sp.el_g[e] = 1 / v; sp.ind_L.push_back(v); }
  }
  return true;
}

void build_sparse_host(const HostPlan& hp, double pilot_f, bool eager, SparseProgram& sp) {
  sp = SparseProgram();
  const HostGather& G = hp.ac;
  const int n_ent = (int)G.ent_col.size();
  if (!entry_constants(hp, sp)) return;  // R<=0: dense kernel reports it
  PilotInput pin;
  pin.n = hp.nvar;
  pin.row_ptr = &G.row_ptr;
  pin.ent_col = &G.ent_col;
  const double w = (2 * kPi) * pilot_f;
  pin.ent_val.resize(n_ent);
  for (int en = 0; en < n_ent; ++en)
    pin.ent_val[en] = std::complex<double>(sp.ent_alpha[en] + sp.ent_jre[en],
                                           w * sp.ent_beta[en] - sp.ent_gamma[en] / w + sp.ent_jim[en]);
  // Entries with identical constants share one value: when few distinct values exist they are
  // materialised once per system in the fast pool instead of being recomputed where used.
  std::vector<int> entry_class(n_ent, 0);
  int n_class = 0;
  {
    std::map<std::vector<double>, int> seen;
    for (int en = 0; en < n_ent; ++en) {
      std::vector<double> key = {sp.ent_alpha[en] + sp.ent_jre[en], sp.ent_jim[en], sp.ent_beta[en], sp.ent_gamma[en]};
      auto it = seen.find(key);
      if (it == seen.end()) it = seen.insert(std::make_pair(key, n_class++)).first;
      entry_class[en] = it->second;
    }
  }
  const int kFastSlots = 12;  // 12 x 16 B x 128 threads = 24 KiB per CTA, 8 CTAs per SM
  // sweep mode: entry values differ per instance, so no constants are materialised (n_class = 0)
  build_sparse_program(pin, sp, kFastSlots, &entry_class, eager ? 0 : n_class);
}

// Builds (or reuses) the sparse program of this topology on ctx.  pilot_f: a representative frequency.
// Returns SPICEY_SUCCESS with ctx.sp_valid=false when the sparse path does not apply.
int prepare_sparse(DeviceCtx& ctx, const HostPlan& hp, double pilot_f, bool eager, cudaStream_t stream) {
  const uint64_t key = plan_key(hp) ^ (eager ? 0x9e3779b97f4a7c15ull : 0ull);
  if (ctx.sp_key == key) return SPICEY_SUCCESS;  // cached (valid or known not to apply)
  ctx.sp_key = key;
  ctx.sp_valid = false;
  ctx.sp_launched = 0;
  ctx.sp_adapt = 0;
  if (ctx.sp_cnt.p) CUDA_TRY(cudaMemsetAsync((char*)ctx.sp_cnt.p + 8, 0, 8, stream));
  SparseProgram& sp = ctx.sp;
  build_sparse_host(hp, pilot_f, eager, sp);
  const int n_ent = (int)hp.ac.ent_col.size();
  ctx.sp_eager = eager;
  ctx.sp_pilot_f = pilot_f;
  if (!sp.ok) return SPICEY_SUCCESS;
  // Workspace stride: the resident grid the workspace is sized for (offsets are baked into the program).
  {
    const int block = 128;
    const size_t per_thread = sizeof(double2) * (size_t)std::max(1, sp.n_slots + sp.n_fast + (eager ? sp.n_stamp + hp.n_ac_elem : 0));
    long long T = (long long)ctx.sm_count * 10 * block;
    const long long t_cap = std::max<long long>(block, (long long)(((size_t)12 << 30) / per_thread) / block * block);
    T = std::min(T, t_cap);                       // workspace <= 12 GiB
    while ((unsigned long long)(sp.n_slots + sp.n_fast + (eager ? sp.n_stamp + hp.n_ac_elem : 0)) * (unsigned long long)T > 0x7ffffff0ull && T > block) T -= block;  // 31-bit offsets
    ctx.sp_T = T;
  }
  const long long T = ctx.sp_T;
  // A/B-measured alternative (fast pool as the tail of the global workspace, kept hot by L1): within 5 % on
  // cfg2 and 15 % slower on cfg4 than the shared-memory pool, so it stays off.
  const bool unified = ctx.sp_unified = false;
  std::vector<int4>& code = ctx.sp_code_scaled;
  code.assign(sp.code.size(), make_int4(0, 0, 0, 0));
  for (size_t i = 0; i < sp.code.size(); ++i) {
    const MicroWord& m = sp.code[i];
    const int ka = (m.hdr >> 4) & 3, kb = (m.hdr >> 6) & 3, kc = (m.hdr >> 10) & 3;
    int4 q = make_int4(m.hdr, m.a, m.b, m.c);
    // pointer operands: bit 31 = fast pool, low bits = offset in double2 units
    auto enc = [&](int kind, int v) -> int {
      if (kind == 1) return (int)(unsigned)((long long)v * T);
      if (kind == 3) return unified ? (int)(unsigned)((long long)(sp.n_slots + v) * T) : (int)(0x80000000u | (unsigned)(v * 128));
      return v;
    };
    q.y = enc(ka, m.a); q.z = enc(kb, m.b); q.w = enc(kc, m.c);
    code[i] = q;
  }
  std::vector<unsigned> x_off(hp.nvar);
  for (int i = 0; i < hp.nvar; ++i) x_off[i] = (unsigned)((long long)sp.x_slot[i] * T);
  std::vector<uint2> el_x(std::max(1, hp.n_ac_elem));
  for (int e = 0; e < hp.n_ac_elem; ++e) {
    const int n1 = hp.ends[e].x, n2 = hp.ends[e].y;
    el_x[e] = make_uint2(n1 ? x_off[n1 - 1] : 0xffffffffu, n2 ? x_off[n2 - 1] : 0xffffffffu);
  }
  std::vector<double2> c0(n_ent), c1(n_ent);
  for (int en = 0; en < n_ent; ++en) {
    c0[en] = make_double2(sp.ent_alpha[en] + sp.ent_jre[en], sp.ent_jim[en]);
    c1[en] = make_double2(sp.ent_beta[en], sp.ent_gamma[en]);
  }
  std::vector<unsigned char> blob;
  size_t o_code = push_blob(blob, code), o_xs = push_blob(blob, x_off), o_ex = push_blob(blob, el_x);
  size_t o_c0 = push_blob(blob, c0), o_c1 = push_blob(blob, c1), o_ea = push_blob(blob, sp.el_a);
  size_t o_eb = push_blob(blob, sp.el_b), o_eg = push_blob(blob, sp.el_g), o_l = push_blob(blob, sp.ind_L);
  size_t o_ce = push_blob(blob, sp.const_entry);
  int rc = ctx.sp_blob.ensure(blob.size() + 16);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(ctx.sp_blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  unsigned char* b = (unsigned char*)ctx.sp_blob.p;
  SparseArgs& a = ctx.sp_args;
  memset(&a, 0, sizeof(a));
  a.code = (const int4*)(b + o_code);
  a.x_off = (const unsigned*)(b + o_xs);
  a.el_x = (const uint2*)(b + o_ex);
  a.n = hp.nvar; a.n_stamp = sp.n_stamp; a.n_slots = sp.n_slots; a.n_fast = sp.n_fast; a.n_const = sp.n_const;
  a.const_entry = (const int*)(b + o_ce);
  a.ent_c0 = (const double2*)(b + o_c0); a.ent_c1 = (const double2*)(b + o_c1);
  a.el_a = (const double*)(b + o_ea); a.el_b = (const double*)(b + o_eb); a.el_g = (const double*)(b + o_eg);
  a.ind_L = (const double*)(b + o_l); a.n_ind = (int)sp.ind_L.size();
  a.n_ac_elem = hp.n_ac_elem; a.nn = hp.nn; a.v_first = hp.off[ELEM_V];
  ctx.sp_valid = true;
  return SPICEY_SUCCESS;
}

// Sparse launch + dense fallback over the diverged list.  args.p_begin must be the start of a
// single-instance frequency range (point index == frequency index).
int launch_ac_dense(DeviceCtx& ctx, const HostPlan& hp, const DevPlan& dp, const AcArgs& args, uint32_t flags,
                    cudaStream_t stream, int* tier_out, int64_t* launches);

constexpr long long kTileMinPoints = 4096;    // below this the ~2 s compile of a new (Nvar, shape) dense kernel does not pay off (unless forced)
constexpr long long kSparseMinPoints = 2048;  // batches from which the sparse program path pays for its host-side analysis
constexpr size_t kJitMaxOps = 3000;          // larger programs stay on the interpreter (compile time: cfg2's 831 micro-ops take 4 s)
constexpr int kJitSpareValues = 64;          // cross-phase values the registers can hold beside the shared-memory slots
constexpr int kJitRegValuesWithColumn = 40;  // ... when the global column is in use (its loads in flight need registers too)
constexpr int kJitRegValuesWide = 8;         // ... in the 16-warps-per-SM shape of large column programs (128 registers per thread)
constexpr size_t kJitWideOps = 1600;         // program size from which a column program takes that shape
constexpr int kJitMaxGlobalValues = 2048;    // ... and beside the kernel's [slot][thread] column of global memory (32 KB per thread)
constexpr long long kJitMinPoints = 200000;  // below this the ~4 s compile does not pay off (unless forced)

void jit_source(const SparseProgram& sp, const HostPlan& hp, const CodegenOptions& opt, bool eager, std::string& src,
                CodegenStats& st) {
  std::vector<int> n1(hp.n_ac_elem), n2(hp.n_ac_elem), ty(hp.n_ac_elem), vi(hp.n_ac_elem);
  for (int e = 0; e < hp.n_ac_elem; ++e) { n1[e] = hp.ends[e].x; n2[e] = hp.ends[e].y; ty[e] = hp.meta[e].x; vi[e] = hp.meta[e].y; }
  CodegenInput ci;
  ci.sp = &sp; ci.nn = hp.nn; ci.n_ac_elem = hp.n_ac_elem; ci.v_first = hp.off[ELEM_V];
  ci.n1 = n1.data(); ci.n2 = n2.data();
  ci.eager = eager; ci.ent_ptr = hp.ac.ent_ptr.data(); ci.contrib = hp.ac.contrib.data();
  ci.el_type = ty.data(); ci.el_vidx = vi.data(); ci.var_of_slot = hp.var_of_slot.data(); ci.values = hp.values.data();
  src = generate_sparse_kernel_source(ci, opt, &st);
}

// source -> loaded kernel, through the disk cache.  A cached cubin that does not load (foreign architecture,
// damaged file) is evicted and the source compiled again once before the variant is given up.
bool load_jit_kernel(const std::string& src, const char* name, cudaLibrary_t* lib, cudaKernel_t* kernel, std::string& note) {
  for (int attempt = 0; attempt < 2; ++attempt) {
    std::vector<char> cubin;
    bool from_cache = false;
    if (!jit_compile_cached(src, cubin, note, &from_cache)) return false;
    if (cudaLibraryLoadData(lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess) {
      if (cudaLibraryGetKernel(kernel, *lib, name) == cudaSuccess) return true;
      cudaLibraryUnload(*lib);
    }
    *lib = nullptr; *kernel = nullptr;
    note = std::string("loading the compiled kernel failed: ") + cudaGetErrorString(cudaGetLastError());
    if (!from_cache) return false;
    jit_cache_evict(src);
  }
  return false;
}

// Compiles (once per topology, handle and variant) the straight-line kernel of the cached sparse program.
// Returns the usable variant or nullptr.
// Returns the usable variant or nullptr.
DeviceCtx::JitVariant* ensure_jit(DeviceCtx& ctx, const HostPlan& hp, bool with_ielem) {
  DeviceCtx::JitVariant& jv = ctx.sp_jit[with_ielem ? 1 : 0];
  // the per-instance (eager) kernel also depends on WHICH value slots are swept
  uint64_t key = ctx.sp_eager ? fnv1a(ctx.sp_key, hp.var_of_slot.data(), sizeof(int) * hp.var_of_slot.size()) : ctx.sp_key;
  if (!key) key = 1;
  if (jv.key == key) return jv.failed ? nullptr : &jv;
  jv.key = key;
  jv.failed = true;
  if (jv.lib) { cudaLibraryUnload(jv.lib); jv.lib = nullptr; jv.kernel = nullptr; }
  const double t0 = now_ms();
  if (const char* e = getenv("SPICEY_JIT_CFG")) {
    int b = 0, m = 0, sl = 0, sy = ctx.sp_jit_sync, pf = ctx.sp_jit_prefetch, sg = ctx.sp_jit_stagger;
    if (sscanf(e, "%d,%d,%d,%d,%d,%d", &b, &m, &sl, &sy, &pf, &sg) >= 3 && b >= 32 && b <= 1024 && b % 32 == 0 && m >= 1 && sl >= 0) {
      ctx.sp_jit_stagger = ctx.sp_jit_stagger_eager = sg;
      ctx.sp_jit_block = ctx.sp_jit_block_eager = b; ctx.sp_jit_minb = ctx.sp_jit_minb_eager = m; ctx.sp_jit_slots = ctx.sp_jit_slots_eager = sl;
      ctx.sp_jit_sync = sy; ctx.sp_jit_prefetch = pf;
    }
  }
  CodegenOptions opt;
  opt.block = ctx.sp_eager ? ctx.sp_jit_block_eager : ctx.sp_jit_block; opt.with_ielem = with_ielem;
  opt.min_blocks = ctx.sp_eager ? ctx.sp_jit_minb_eager : ctx.sp_jit_minb;
  int slots_cfg = ctx.sp_eager ? ctx.sp_jit_slots_eager : ctx.sp_jit_slots;
  const int crossing = count_cross_phase_values(ctx.sp);
  // (per-instance stamping keeps element values in flight: 40 register values at most, the rest in the column)
  const bool with_column = crossing > slots_cfg + (ctx.sp_eager ? kJitRegValuesWithColumn : kJitSpareValues);
  // Large programs that use the global column anyway run from L2 (their code is several times the instruction cache)
  // and are bound by latency, not by HBM: 16 warps per SM with few values on chip beat 6 warps with many (measured,
  // profiles/r3k_ladder_probe.txt: 150 / 200-node ladders 230 -> 339 / 163 -> 250 M solves/s; the 100-node ladder,
  // which is bound by HBM, loses 25 % that way and keeps the cfg-2 shape)
  const bool wide = with_column && !ctx.sp_eager && ctx.sp.code.size() >= kJitWideOps && !getenv("SPICEY_JIT_CFG");
  if (wide) { opt.block = 128; opt.min_blocks = 4; slots_cfg = 27; }
  opt.prefetch_steps = ctx.sp_jit_prefetch; opt.stagger_ns = ctx.sp_eager ? ctx.sp_jit_stagger_eager : ctx.sp_jit_stagger;
  opt.sync_every = ctx.sp_jit_sync;
  if (!ctx.sp_eager && opt.min_blocks >= 2)
    opt.antiphase_ns = (int)std::min(200000.0, ctx.sp_jit_antiphase_ns_per_op * (double)ctx.sp.code.size() * 2.0 / opt.min_blocks);
  if (const char* e = getenv("SPICEY_JIT_ANTIPHASE")) opt.antiphase_ns = atoi(e);   // experiments
  opt.smem_slots = std::min<int>(slots_cfg, (int)((size_t)(227 * 1024 / opt.min_blocks - 1024) / ((size_t)opt.block * 16)));
  // What exceeds registers + shared memory lives in the kernel's global column.  Programs that fit without it (cfg 2: 129
  // values = 75 slots + 54 registers) are generated exactly as before; the others keep fewer values in registers, because
  // the loads of the column run a few rows ahead of their use and need registers of their own.
  opt.reg_values = !with_column ? kJitSpareValues : wide ? kJitRegValuesWide : kJitRegValuesWithColumn;
  if (const char* e = getenv("SPICEY_JIT_REGVALUES")) opt.reg_values = std::max(0, atoi(e));   // experiments
  if (const char* e = getenv("SPICEY_JIT_GAHEAD")) opt.gmem_ahead = std::max(1, atoi(e));
  if (const char* e = getenv("SPICEY_JIT_L2POLICY")) opt.l2_policy = atoi(e) != 0;
  std::string src;
  CodegenStats st;
  jit_source(ctx.sp, hp, opt, ctx.sp_eager, src, st);
  if (!load_jit_kernel(src, "spicey_sparse_jit", &jv.lib, &jv.kernel, ctx.sp_jit_note)) return nullptr;
  if (cudaFuncSetAttribute((const void*)jv.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st.smem_bytes) != cudaSuccess) {
    ctx.sp_jit_note = std::string("the compiled kernel does not accept its shared-memory size: ") + cudaGetErrorString(cudaGetLastError());
    jv.kernel = nullptr;
    return nullptr;
  }
  jv.smem_bytes = st.smem_bytes;
  jv.gmem_slots = st.gmem_slots;
  jv.block = opt.block;
  jv.min_blocks = opt.min_blocks;
  jv.failed = false;
  ctx.sp_jit_compile_ms = now_ms() - t0;
  ctx.sp_jit_note = "ok";
  return &jv;
}

constexpr int kWarpTierWarps = 8;            // warps (systems) per CTA of ac_warp_kernel
constexpr int kWarpPoolSpare = 64;            // free slots from which the bank-residue rule of the pool is relaxed
constexpr int kWarpTierMinSlots = 512;       // thread-per-system workspace (slots) from which the warp tier takes over

// Lowers the cached sparse program to its warp-cooperative form and uploads it (once per topology and handle).
int prepare_warp(DeviceCtx& ctx, const HostPlan& hp, cudaStream_t stream) {
  if (ctx.wp_key == ctx.sp_key) return SPICEY_SUCCESS;
  ctx.wp_key = ctx.sp_key;
  ctx.wp_valid = false;
  const int per_warp_cap = (int)((ctx.smem_optin - 1024 - 32 * 1024) / kWarpTierWarps / sizeof(double2));
  // a value may take a slot of the wrong bank residue once 64 slots lie free: cfg 4's pool 482 -> 362 slots,
  // 24 -> 32 warps per SM, 3.59 -> 3.83 M solves/s (32 / 16 free slots: 3.80 / 3.79)
  int spare = kWarpPoolSpare;
  if (const char* e = getenv("SPICEY_WARP_SPARE")) spare = std::max(1, atoi(e));
  build_warp_program(ctx.sp, per_warp_cap - 64, ctx.wp, spare);
  WarpProgram& wp = ctx.wp;
  // lanes = columns of a pivot row: chain-like circuits (a ladder updates 2-3 entries per row) would leave the
  // warp idle; they stay with one thread per system
  if (!wp.ok || (size_t)(std::max(wp.n_pool, wp.n) + wp.max_elim) > (size_t)per_warp_cap || wp.max_rec16 > 1024) { wp.ok = false; return SPICEY_SUCCESS; }
  ctx.wp_chainlike = wp.n_upd_total < 32ll * wp.n;
  std::vector<unsigned char> blob;
  size_t o_st = push_blob(blob, wp.stream), o_ft = push_blob(blob, wp.fwd_tab), o_bt = push_blob(blob, wp.back_tab);
  size_t o_rh = push_blob(blob, wp.rhs_init);
  int rc = ctx.wp_blob.ensure(blob.size() + 16);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(ctx.wp_blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  unsigned char* b = (unsigned char*)ctx.wp_blob.p;
  WarpArgs& a = ctx.wp_args;
  memset(&a, 0, sizeof(a));
  a.stream = (const int4*)(b + o_st); a.fwd_tab = (const int2*)(b + o_ft); a.back_tab = (const int2*)(b + o_bt);
  a.rhs_init = (const int*)(b + o_rh);
  a.n_groups = wp.n_groups; a.max_rec16 = wp.max_rec16; a.g_first0 = wp.g_first0; a.g_count0 = wp.g_count0;
  const SparseArgs& sa = ctx.sp_args;   // per-entry / per-element constants already uploaded for the interpreter
  a.ent_c0 = sa.ent_c0; a.ent_c1 = sa.ent_c1; a.el_a = sa.el_a; a.el_b = sa.el_b; a.el_g = sa.el_g;
  a.ind_L = sa.ind_L; a.n_ind = sa.n_ind;
  a.n = hp.nvar; a.nn = hp.nn; a.n_ac_elem = hp.n_ac_elem; a.v_first = hp.off[ELEM_V];
  a.n_pool = wp.n_pool; a.n_gslots = wp.n_gslots; a.max_elim = std::max(1, wp.max_elim);
  ctx.wp_valid = true;
  return SPICEY_SUCCESS;
}

int launch_ac_warp(DeviceCtx& ctx, const HostPlan& hp, const DevPlan& dp, const AcArgs& args, uint32_t flags,
                   cudaStream_t stream, long long* fb_list, int* fb_count, int64_t* launches) {
  (void)hp; (void)flags;
  WarpArgs a = ctx.wp_args;
  const size_t per_warp = sizeof(double2) * (size_t)(std::max(a.n_pool, a.n) + a.max_elim);
  const size_t smem = per_warp * kWarpTierWarps + 2 * sizeof(int4) * (size_t)a.max_rec16;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (ctx.smem_optin + 1024) / (smem + 1024)));
  const long long want = (args.p_count + kWarpTierWarps - 1) / kWarpTierWarps;
  const unsigned grid = (unsigned)std::min<long long>(want, (long long)ctx.sm_count * per_sm);
  int rc = ctx.wp_work.ensure(sizeof(double2) * (size_t)a.n_gslots * grid * kWarpTierWarps);
  if (rc) return rc;
  a.el_ends = dp.ends;
  a.freqs = args.freqs + args.p_begin; a.p_count = args.p_count;
  a.G = (double2*)ctx.wp_work.p;
  a.x = args.x; a.ielem = args.ielem; a.status = args.status; a.series_ld = args.series_ld;
  a.fb_list = fb_list; a.fb_count = fb_count;
  CUDA_TRY(cudaFuncSetAttribute(ac_warp_kernel<kWarpTierWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ac_warp_kernel<kWarpTierWarps><<<grid, kWarpTierWarps * 32, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  if (launches) ++*launches;
  return SPICEY_SUCCESS;
}


// ---------------------------------------------------------------------------------
// Banded + bordered tier (SPICEY_TIER_BAND): band_plan.h decides whether the topology qualifies and lays out the
// tables, band_kernel.cuh (embedded as text, compiled by NVRTC once per band shape and cached on disk) solves.
struct BandArgs {   // must match band_kernel.cuh
  const double* freqs; long long p_count;
  double2* x; double2* ielem; int* status; long long series_ld;
  long long* fb_list; int* fb_count;
  double2* G; long long g_stride;
  const double2* tab; const uint2* flags; const int* newvar; const double4* el_rec;
  const double* ind_L;
  int n, nb, n_out, n_ac_elem, n_ind;
  int o_init, o_initb, o_brd0, o_bb0, o_step;
  unsigned long long* prof;
};

// Half-bandwidth from which the banded tier beats the interpreted thread-per-system program (measured, tools/tier_sweep.py,
// 200,000 points: mesh4 (W 4) 567 against 440 M solves/s, mesh6 138 / 97, mesh8 87 / 25, mesh16 11.9 / 1.45; but mesh3 (W 3)
// 735 / 1,124 and the 400-node ladder (W 1) 14 / 38: a step of the band kernel costs ~150 instructions whatever the width)
constexpr int kBandMinBandwidth = 4;
constexpr double kBandOrderTol = 5e-10;   // per-entry agreement of a renumbered plan with the netlist order (parity bar: 1e-9)
constexpr int kBandUModeDefault = 2;   // band_kernel.cuh BAND_UMODE: the pivot rows leave through the TMA unit

void band_input(const HostPlan& hp, const SparseProgram& sp, double pilot_f, BandInput& in) {
  in.n = hp.nvar; in.nn = hp.nn; in.nV = hp.nV;
  in.row_ptr = &hp.ac.row_ptr; in.ent_col = &hp.ac.ent_col;
  in.ent_alpha = &sp.ent_alpha; in.ent_beta = &sp.ent_beta; in.ent_gamma = &sp.ent_gamma;
  in.ent_jre = &sp.ent_jre; in.ent_jim = &sp.ent_jim;
  in.pilot_w = (2 * kPi) * pilot_f;
}

void band_force_shape(int& L, int& RPL) {   // experiments: SPICEY_BAND_SHAPE=L,RPL
  L = RPL = 0;
  if (const char* e = getenv("SPICEY_BAND_SHAPE")) {
    int a = 0, b = 0;
    if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 1 && a <= 32 && (a & (a - 1)) == 0 && b >= 1 && b <= 4 && (b & (b - 1)) == 0) { L = a; RPL = b; }
  }
}

// Builds (once per topology and handle) the band plan of the cached sparse program's circuit and uploads its tables.
int prepare_band(DeviceCtx& ctx, const HostPlan& hp, cudaStream_t stream, bool force) {
  const uint64_t want = ctx.sp_key ^ (force ? 0x5bd1e995ull : 0ull);
  if (ctx.bp_key == want) return SPICEY_SUCCESS;
  ctx.bp_key = want;
  ctx.bp_valid = false;
  BandInput in;
  band_input(hp, ctx.sp, ctx.sp_pilot_f, in);
  int fl = 0, fr = 0;
  band_force_shape(fl, fr);
  build_band_plan(in, ctx.bp, fl, fr);
  BandPlan& bp = ctx.bp;
  if (!bp.ok || (bp.bandwidth < kBandMinBandwidth && !force)) return SPICEY_SUCCESS;
  // A renumbered plan eliminates in another order than the reference: before it becomes the default for a topology,
  // the reference's algorithm is run on the host in both orders at the pilot frequency and at both ends of the sweep
  // (band_plan.h: band_order_deviation); a plan whose solution differs per entry by more than kBandOrderTol — strongly
  // attenuating networks whose small node voltages keep their relative accuracy only in the netlist order — is declined
  // (SPICEY_FLAG_BAND still forces it).  cfg 4's mesh: a few 1e-11.
  if (bp.renumbered && !force) {
    double worst = 0.0;
    for (double f : {ctx.sp_pilot_f, ctx.sp_f_lo, ctx.sp_f_hi}) {
      if (!(f > 0.0)) continue;
      const double d = band_order_deviation(in, bp, (2 * kPi) * f);
      worst = std::max(worst, (d < 0.0 || d != d) ? 1.0 : d);   // a guard tripped / NaN: no evidence, no renumbering
    }
    ctx.bp_order_dev = worst;
    if (worst > kBandOrderTol) return SPICEY_SUCCESS;
  }
  // element records of the unpack phase: current = Y (x[i1] - x[i2]) in elimination-order indices, index n = the
  // zero slot (ground; bp.n counts the padding rows of the band); a V element's current is its branch unknown:
  // (branch, zero slot, Y = 1)
  std::vector<double4> el_idx(std::max(1, hp.n_ac_elem));
  for (int e = 0; e < hp.n_ac_elem; ++e) {
    const int n1 = hp.ends[e].x, n2 = hp.ends[e].y;
    long long ij;
    double4 r = make_double4(0.0, 0.0, 0.0, 0.0);
    if (e >= hp.off[ELEM_V]) {
      ij = (long long)(bp.nb + (e - hp.off[ELEM_V])) | ((long long)bp.n << 32);
      r.y = 1.0;
    } else {
      ij = (long long)(n1 ? bp.newvar[n1 - 1] : bp.n) | ((long long)(n2 ? bp.newvar[n2 - 1] : bp.n) << 32);
      r.y = ctx.sp.el_a[e]; r.z = ctx.sp.el_b[e]; r.w = ctx.sp.el_g[e];
    }
    memcpy(&r.x, &ij, sizeof ij);
    el_idx[e] = r;
  }
  std::vector<unsigned char> blob;
  size_t o_tab = 0;
  if (bp.rc_only) {   // (alpha, beta) per entry: half the table bytes, one multiply per stamped value
    std::vector<double2> packed(bp.tab.size());
    for (size_t i = 0; i < bp.tab.size(); ++i) packed[i] = make_double2(bp.tab[i].alpha_jre, bp.tab[i].beta);
    o_tab = push_blob(blob, packed);
  } else {
    o_tab = push_blob(blob, bp.tab);
  }
  while (blob.size() % 128) blob.push_back(0);   // step records start on 128-byte lines (the kernel prefetches by line)
  const size_t o_fl = push_blob(blob, bp.flags), o_nv = push_blob(blob, bp.newvar), o_ei = push_blob(blob, el_idx);
  int rc = ctx.bp_blob.ensure(blob.size() + 16);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(ctx.bp_blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  unsigned char* b = (unsigned char*)ctx.bp_blob.p;
  ctx.bp_dev.tab = b + o_tab; ctx.bp_dev.flags = b + o_fl; ctx.bp_dev.newvar = b + o_nv; ctx.bp_dev.el_rec = b + o_ei;
  ctx.bp_valid = true;
  return SPICEY_SUCCESS;
}

int band_sync_default() {
  if (const char* e = getenv("SPICEY_BAND_SYNC")) return std::max(0, std::min(16, atoi(e)));   // experiments: barriers per W steps
  return 1;
}

// How the pivot rows (U) reach the workspace (BAND_UMODE of band_kernel.cuh).  The column-major workspace makes the plain
// store (0) touch 32 lines per instruction; the tensor store (2) takes the pivot record from shared memory through the
// TMA unit instead; 1 pairs two steps into one 32-byte store per lane.
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeTiledFn tensor_map_encoder() {
  static TensorMapEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return (TensorMapEncodeTiledFn)p;
  }();
  return fn;
}

int band_umode(const DeviceCtx& ctx, const BandPlan& bp) {
  int m = kBandUModeDefault;
  if (const char* e = getenv("SPICEY_BAND_UMODE")) m = atoi(e) == 2 ? 2 : 0;   // experiments
  if (bp.W < 4) m = 0;
  if (m == 2 && (ctx.band_tma_refused || !tensor_map_encoder())) m = 0;
  return m;
}

std::string band_source(const BandPlan& bp, bool with_ielem, int warps, int minb, int umode) {
  char head[512];
  snprintf(head, sizeof head,
           "#define BAND_L %d\n#define BAND_RPL %d\n#define BAND_NB %d\n#define BAND_ABMASK %uu\n#define BAND_IELEM %d\n"
           "#define BAND_WARPS %d\n#define BAND_MINB %d\n#define BAND_RC %d\n#define BAND_SYNC %d\n#define BAND_UMODE %d\n#define BAND_PROF %d\n",
           bp.L, bp.RPL, bp.NB, bp.abmask, with_ielem ? 1 : 0, warps, minb, bp.rc_only ? 1 : 0, band_sync_default(), umode,
           getenv("SPICEY_BAND_PROF") ? 1 : 0);
  return std::string(head) + kBandKernelSource;
}

// Shared memory of one CTA: per warp the ring of staged step records, per system the solution vector and the pivot
// records (umode 2: their band part per warp, [4][W][systems of the warp], in front: the tensor store's source).
size_t band_smem_bytes(const BandPlan& bp, int warps, int umode) {
  const size_t ring = (size_t)4 * bp.step_stride * (bp.rc_only ? 1 : 2);   // BRING step records per warp
  const size_t recs = umode == 2 ? 2 * (size_t)(bp.NB + 2) : (umode ? 4 : 2) * (size_t)(bp.W + bp.NB + 2);
  const size_t sys = ((size_t)bp.n + 1 + recs) | 1;   // odd: see band_kernel.cuh
  const size_t tma = umode == 2 ? (size_t)4 * bp.W * (32 / bp.L) : 0;
  return sizeof(double2) * (size_t)warps * (tma + ring + (32 / bp.L) * sys);
}

// Launch shape: as many warps per SM as shared memory and the register file (255 per thread) allow.
bool band_launch_shape(const DeviceCtx& ctx, const BandPlan& bp, int umode, int& warps, int& minb) {
  if (const char* e = getenv("SPICEY_BAND_CFG")) {   // experiments: warps per CTA, CTAs per SM
    int a = 0, b = 0;
    if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 1 && a <= 16 && b >= 1 && b <= 8 && band_smem_bytes(bp, a, umode) * b + 1024 * b <= ctx.smem_optin + 1024) {
      warps = a; minb = b;
      return true;
    }
  }
  const int shapes[][2] = {{4, 2}, {2, 2}, {2, 1}, {1, 1}};
  for (const auto& sh : shapes)
    if ((band_smem_bytes(bp, sh[0], umode) + 1024) * sh[1] <= ctx.smem_optin) { warps = sh[0]; minb = sh[1]; return true; }
  return false;
}

DeviceCtx::BandJit* ensure_band_jit(DeviceCtx& ctx, bool with_ielem) {
  DeviceCtx::BandJit& jv = ctx.band_jit[with_ielem ? 1 : 0];
  const BandPlan& bp = ctx.bp;
  int warps = 0, minb = 0;
  const int umode = band_umode(ctx, bp);
  if (!band_launch_shape(ctx, bp, umode, warps, minb)) return nullptr;
  const int shape[10] = {bp.L, bp.RPL, bp.NB, (int)bp.abmask, warps, minb, with_ielem ? 1 : 0, bp.rc_only ? 1 : 0, band_sync_default(), umode};
  uint64_t key = fnv1a(1469598103934665603ull, shape, sizeof shape);
  if (!key) key = 1;
  if (jv.key == key) return jv.failed ? nullptr : &jv;
  jv.key = key;
  jv.failed = true;
  if (jv.lib) { cudaLibraryUnload(jv.lib); jv.lib = nullptr; jv.kernel = nullptr; }
  const std::string src = band_source(bp, with_ielem, warps, minb, umode);
  if (!load_jit_kernel(src, "spicey_band_jit", &jv.lib, &jv.kernel, ctx.sp_jit_note)) return nullptr;
  jv.warps = warps; jv.minb = minb; jv.umode = umode;
  jv.failed = false;
  return &jv;
}

int launch_ac_band(DeviceCtx& ctx, const HostPlan& hp, const AcArgs& args, DeviceCtx::BandJit* jv, cudaStream_t stream,
                   long long* fb_list, int* fb_count, int64_t* launches) {
  const BandPlan& bp = ctx.bp;
  const int gpb = jv->warps * (32 / bp.L);   // systems per CTA
  const size_t smem = band_smem_bytes(bp, jv->warps, jv->umode);
  unsigned grid = (unsigned)std::min<long long>((args.p_count + gpb - 1) / gpb, (long long)ctx.sm_count * jv->minb);
  if (const char* e = getenv("SPICEY_BAND_GRID")) grid = std::min<unsigned>(grid, (unsigned)std::max(1, atoi(e)));   // experiments
  int rc = ctx.bp_work.ensure(sizeof(double2) * (size_t)bp.g_stride * grid * gpb);
  if (rc) return rc;
  CUDA_TRY(cudaFuncSetAttribute((const void*)jv->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const SparseArgs& sa = ctx.sp_args;   // per-element constants already uploaded for the interpreter
  BandArgs a;
  memset(&a, 0, sizeof a);
  a.freqs = args.freqs + args.p_begin; a.p_count = args.p_count;
  a.x = args.x; a.ielem = args.ielem; a.status = args.status; a.series_ld = args.series_ld;
  a.fb_list = fb_list; a.fb_count = fb_count;
  a.G = (double2*)ctx.bp_work.p; a.g_stride = bp.g_stride;
  a.tab = (const double2*)ctx.bp_dev.tab; a.flags = (const uint2*)ctx.bp_dev.flags;
  a.newvar = (const int*)ctx.bp_dev.newvar; a.el_rec = (const double4*)ctx.bp_dev.el_rec;
  a.ind_L = sa.ind_L;
  a.n = bp.n; a.nb = bp.nb; a.n_out = hp.nvar; a.n_ac_elem = hp.n_ac_elem; a.n_ind = sa.n_ind;
  a.o_init = bp.o_init; a.o_initb = bp.o_initb; a.o_brd0 = bp.o_brd0; a.o_bb0 = bp.o_bb0; a.o_step = bp.o_step;
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  if (jv->umode == 2) {
    // The U part of every system's workspace as one tensor of doubles: (2 W: row slot x re / im | systems | nb + W columns);
    // a box is one row slot of the W columns k + 1 .. k + W for the systems of one warp, [column][system] in shared memory.
    const cuuint64_t dims[3] = {(cuuint64_t)2 * bp.W, (cuuint64_t)grid * gpb, (cuuint64_t)(bp.nb + bp.W)};
    const cuuint64_t strides[2] = {(cuuint64_t)bp.g_stride * sizeof(double2), (cuuint64_t)bp.W * sizeof(double2)};
    const cuuint32_t box[3] = {2, (cuuint32_t)(32 / bp.L), (cuuint32_t)bp.W};
    const cuuint32_t estr[3] = {1, 1, 1};
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    const CUresult er = enc ? enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ctx.bp_work.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                            : CUDA_ERROR_NOT_SUPPORTED;
    if (er != CUDA_SUCCESS) {   // this driver will not describe the workspace: compile and run the plain-store kernel instead
      ctx.band_tma_refused = true;
      ctx.sp_jit_note = "cuTensorMapEncodeTiled refused the band workspace (CUresult " + std::to_string((int)er) + "): plain U stores";
      DeviceCtx::BandJit* alt = ensure_band_jit(ctx, args.ielem != nullptr);
      if (!alt || alt->umode == 2) return fail(SPICEY_ERR_CUDA, "the band kernel could not be rebuilt without the tensor store");
      return launch_ac_band(ctx, hp, args, alt, stream, fb_list, fb_count, launches);
    }
  }
  void* kargs[] = {&a, &tm};   // the plain-store kernels take the first parameter only
  if (getenv("SPICEY_BAND_PROF")) {   // experiments: cycles per phase, summed over the warps, printed per launch (synchronous)
    unsigned long long* d = nullptr;
    unsigned long long h[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMalloc(&d, sizeof h));
    CUDA_TRY(cudaMemsetAsync(d, 0, sizeof h, stream));
    a.prof = d;
    CUDA_TRY(cudaLaunchKernel((const void*)jv->kernel, dim3(grid), dim3(jv->warps * 32), kargs, smem, stream));
    CUDA_TRY(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    cudaFree(d);
    const double tot = (double)(h[0] + h[1] + h[2] + h[3]) + 1e-9;
    fprintf(stderr, "[band prof] warps %u  cycles/warp: stamp %.0f (%.1f%%)  elim %.0f (%.1f%%)  backsub %.0f (%.1f%%)  results %.0f (%.1f%%)\n",
            grid * jv->warps, h[0] / (double)(grid * jv->warps), 100 * h[0] / tot, h[1] / (double)(grid * jv->warps), 100 * h[1] / tot,
            h[2] / (double)(grid * jv->warps), 100 * h[2] / tot, h[3] / (double)(grid * jv->warps), 100 * h[3] / tot);
    if (launches) ++*launches;
    return SPICEY_SUCCESS;
  }
  CUDA_TRY(cudaLaunchKernel((const void*)jv->kernel, dim3(grid), dim3(jv->warps * 32), kargs, smem, stream));
  if (launches) ++*launches;
  return SPICEY_SUCCESS;
}

int launch_ac_sparse(DeviceCtx& ctx, const HostPlan& hp, const DevPlan& dp, const AcArgs& args, uint32_t flags,
                     cudaStream_t stream, int* tier_out, int64_t* launches) {
  const int block = 128;
  const long long T = ctx.sp_T;  // workspace stride (fixed when the program was uploaded)
  const long long nthreads = std::min<long long>((args.p_count + block - 1) / block * block, T);
  size_t wbytes = sizeof(double2) * (size_t)std::max(1, ctx.sp.n_slots + ctx.sp.n_fast + (ctx.sp_eager ? ctx.sp.n_stamp + hp.n_ac_elem : 0)) * T;
  int rc = ctx.sp_work.ensure(wbytes);
  if (rc) return rc;
  if ((rc = ctx.sp_fb.ensure(sizeof(long long) * args.p_count + 64))) return rc;
  if (!ctx.sp_cnt.p && (rc = reset_call_counters(ctx, stream))) return rc;
  int* fb_count = (int*)ctx.sp_fb.p;
  long long* fb_list = (long long*)((char*)ctx.sp_fb.p + 64);
  CUDA_TRY(cudaMemsetAsync(fb_count, 0, sizeof(int), stream));
  ctx.sp_launched += args.p_count;
  // The factorisation of one system fits one thread's registers + shared memory (cfg2: 129 values), or what exceeds
  // them goes to the compiled kernel's global column (sparse_codegen.h: mid-size ladders, trees, sparse networks;
  // bounded, so that the column of the resident grid stays a few hundred MB).  kJitMaxOps bounds the compile time.
  if (ctx.sp_jit_fit_key != ctx.sp_key) {
    ctx.sp_jit_fit_key = ctx.sp_key;
    const int on_chip = (ctx.sp_eager ? ctx.sp_jit_slots_eager : ctx.sp_jit_slots) + kJitSpareValues;
    const int crossing = count_cross_phase_values(ctx.sp);
    ctx.sp_jit_fits = crossing <= on_chip + kJitMaxGlobalValues;
    ctx.sp_jit_column = crossing > on_chip;
  }
  // Circuits with inductors change their pivot order along a sweep more often than not (see launch_ac): their first
  // >= 4,096 points run interpreted, and a kernel is compiled for the topology only once the pilot order has been seen to hold.
  const bool order_known = ctx.sp.ind_L.empty() || ctx.sp_adapt == 1 || (flags & (SPICEY_FLAG_BAND | SPICEY_FLAG_WARP)) ||
                           args.p_count < kTileMinPoints;
  bool want_jit = !(flags & (SPICEY_FLAG_NO_JIT | SPICEY_FLAG_BAND)) && ctx.sp.code.size() <= kJitMaxOps && ctx.sp_jit_fits &&
                  order_known && (args.p_count >= kJitMinPoints || (flags & SPICEY_FLAG_JIT));
  // A program that needs the global column and is banded enough for the register-blocked banded tier (meshes from
  // half-bandwidth 4: measured faster there than one thread per system) is left to that tier, as before the column existed.
  if (want_jit && ctx.sp_jit_column && !ctx.sp_eager && !(flags & SPICEY_FLAG_NO_BAND) && args.series_ld < (1ll << 31)) {
    rc = prepare_band(ctx, hp, stream, false);
    if (rc) return rc;
    if (ctx.bp_valid) want_jit = false;
  }
  DeviceCtx::JitVariant* jv = (want_jit && args.series_ld < (1ll << 32)) ? ensure_jit(ctx, hp, args.ielem != nullptr) : nullptr;
  if (jv) {
    JitArgs j;
    j.freqs = ctx.sp_eager ? args.freqs : args.freqs + args.p_begin; j.p_count = args.p_count;
    j.x = args.x; j.ielem = args.ielem; j.status = args.status; j.series_ld = args.series_ld;
    j.fb_list = fb_list; j.fb_count = fb_count; j.n = hp.nvar; j.n_ac_elem = hp.n_ac_elem;
    j.var_values = dp.var_values; j.n_inst = dp.n_inst; j.n_freq = args.n_freq; j.p_begin = args.p_begin;
    const int jblock = jv->block;
    const long long resident = (long long)ctx.sm_count * jv->min_blocks;
    unsigned jgrid = (unsigned)std::min<long long>((args.p_count + jblock - 1) / jblock, resident);
    if (const char* e = getenv("SPICEY_JIT_GRID")) jgrid = std::min<unsigned>(jgrid, (unsigned)std::max(1, atoi(e)));   // experiments: fewer SMs
    if (jv->gmem_slots > 0) {
      if ((rc = ctx.sp_jit_work.ensure(sizeof(double2) * (size_t)jv->gmem_slots * jgrid * jblock))) return rc;
      j.work = (double2*)ctx.sp_jit_work.p;
    } else {
      j.work = nullptr;
    }
    void* kargs[] = {&j};
    CUDA_TRY(cudaLaunchKernel((const void*)jv->kernel, dim3(jgrid), dim3(jblock), kargs, jv->smem_bytes, stream));
    if (launches) ++*launches;
    AcArgs d = args;
    d.plist = fb_list;
    d.pcount = fb_count;
    d.fb_total = (unsigned long long*)ctx.sp_cnt.p;
    rc = launch_ac_dense(ctx, hp, dp, d, flags, stream, nullptr, launches);
    if (rc) return rc;
    if (tier_out) *tier_out = SPICEY_TIER_SPARSE_JIT;
    return SPICEY_SUCCESS;
  }
  // Banded + bordered circuits the thread-per-system compiled kernel does not take (meshes, long ladders): a few
  // lanes per system, register-blocked (band_plan.h / band_kernel.cuh).  Compiled once per band shape and machine.
  if (!ctx.sp_eager && !(flags & (SPICEY_FLAG_NO_BAND | SPICEY_FLAG_NO_JIT)) && args.series_ld < (1ll << 31) && order_known &&
      ((flags & SPICEY_FLAG_BAND) || args.p_count >= kJitMinPoints || (flags & SPICEY_FLAG_JIT))) {
    rc = prepare_band(ctx, hp, stream, (flags & SPICEY_FLAG_BAND) != 0);
    if (rc) return rc;
    if (ctx.bp_valid) {
      if (DeviceCtx::BandJit* bj = ensure_band_jit(ctx, args.ielem != nullptr)) {
        rc = launch_ac_band(ctx, hp, args, bj, stream, fb_list, fb_count, launches);
        if (rc) return rc;
        AcArgs d = args;
        d.plist = fb_list;
        d.pcount = fb_count;
        d.fb_total = (unsigned long long*)ctx.sp_cnt.p;
        rc = launch_ac_dense(ctx, hp, dp, d, flags, stream, nullptr, launches);
        if (rc) return rc;
        if (tier_out) *tier_out = SPICEY_TIER_BAND;
        return SPICEY_SUCCESS;
      }
    }
  }
  // Large programs of a plain frequency sweep: one warp per system (warp_program.h / ac_warp.cuh).
  if (!ctx.sp_eager && !(flags & SPICEY_FLAG_NO_WARP) && (ctx.sp.n_slots >= kWarpTierMinSlots || (flags & SPICEY_FLAG_WARP))) {
    rc = prepare_warp(ctx, hp, stream);
    if (rc) return rc;
    if (ctx.wp_valid && (!ctx.wp_chainlike || (flags & SPICEY_FLAG_WARP))) {
      rc = launch_ac_warp(ctx, hp, dp, args, flags, stream, fb_list, fb_count, launches);
      if (rc) return rc;
      AcArgs d = args;
      d.plist = fb_list;
      d.pcount = fb_count;
      d.fb_total = (unsigned long long*)ctx.sp_cnt.p;
      rc = launch_ac_dense(ctx, hp, dp, d, flags, stream, nullptr, launches);
      if (rc) return rc;
      if (tier_out) *tier_out = SPICEY_TIER_SPARSE_WARP;
      return SPICEY_SUCCESS;
    }
  }
  SparseArgs a = ctx.sp_args;
  a.freqs = ctx.sp_eager ? args.freqs : args.freqs + args.p_begin;
  a.n_freq = args.n_freq;
  a.p_begin = args.p_begin;
  a.p_count = args.p_count;
  a.plan = dp;
  a.W = (double2*)ctx.sp_work.p;
  a.T = T;
  a.x = args.x; a.ielem = args.ielem; a.status = args.status; a.series_ld = args.series_ld;
  a.fb_list = fb_list; a.fb_count = fb_count;
  const size_t fast_bytes = sizeof(double2) * (size_t)std::max(1, ctx.sp.n_fast) * block;
  const bool pristine_ops = ctx.sp.n_const == 0;
  const bool uni = ctx.sp_unified;
  const unsigned grid = (unsigned)(nthreads / block);
#define SPARSE_LAUNCH(CP)                                                                                  \
  do {                                                                                                     \
    if (ctx.sp_eager) ac_sparse_kernel<CP, true, false, true><<<grid, block, fast_bytes, stream>>>(a);     \
    else if (uni) {                                                                                        \
      if (pristine_ops) ac_sparse_kernel<CP, true, true, false><<<grid, block, 0, stream>>>(a);            \
      else ac_sparse_kernel<CP, false, true, false><<<grid, block, 0, stream>>>(a);                        \
    } else {                                                                                               \
      if (pristine_ops) ac_sparse_kernel<CP, true, false, false><<<grid, block, fast_bytes, stream>>>(a);  \
      else ac_sparse_kernel<CP, false, false, false><<<grid, block, fast_bytes, stream>>>(a);              \
    }                                                                                                      \
  } while (0)
  if ((int)ctx.sp.code.size() <= kConstProgWords) {
    // Constant-memory program: one resident program per device at a time.  The upload is ordered after
    // the last kernel that used the previous contents.
    ConstProgOwner& own = const_owner(ctx.dev);
    std::lock_guard<std::mutex> lock(own.mu);
    if (own.key != ctx.sp_key || own.T != T) {
      if (own.ev) CUDA_TRY(cudaStreamWaitEvent(stream, own.ev, 0));
      CUDA_TRY(cudaMemcpyToSymbolAsync(c_sparse_prog, ctx.sp_code_scaled.data(), sizeof(int4) * ctx.sp_code_scaled.size(),
                                       0, cudaMemcpyHostToDevice, stream));
      own.key = ctx.sp_key;
      own.T = T;
    }
    SPARSE_LAUNCH(true);
    if (!own.ev) CUDA_TRY(cudaEventCreateWithFlags(&own.ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(own.ev, stream));
  } else {
    SPARSE_LAUNCH(false);
  }
#undef SPARSE_LAUNCH
  CUDA_TRY(cudaGetLastError());
  if (launches) ++*launches;
  AcArgs d = args;
  d.plist = fb_list;
  d.pcount = fb_count;
  d.fb_total = (unsigned long long*)ctx.sp_cnt.p;
  rc = launch_ac_dense(ctx, hp, dp, d, flags, stream, nullptr, launches);
  if (rc) return rc;
  if (tier_out) *tier_out = SPICEY_TIER_SPARSE;
  return SPICEY_SUCCESS;
}

// ---------------------------------------------------------------------------------
// Dense register-tile tier (SPICEY_TIER_TILE): tile_plan.h picks the thread grid for Nvar, tile_kernel.cuh (embedded as
// text, compiled by NVRTC once per (Nvar, shape) and cached on disk) stamps, factors with partial pivoting and unpacks.
struct TileArgs {   // must match tile_kernel.cuh
  const double* freqs; long long n_freq, p_begin, p_count;
  double2* x; double2* ielem; int* status; long long series_ld;
  const int4* ends; const int2* meta; const double* values; const int* var_of_slot; const double* var_values; long long n_inst;
  const int* ent_rc; const int* ent_ptr; const int* contrib;
  const double2* ctab; const double4* el_rec; const double* ind_L;
  long long* fb_list; int* fb_count;
  const long long* plist; const int* pcount; unsigned long long* fb_total;
  int n_ind, n_ent, nn, nV, n_elem, n_ac_elem, off_v, off_v_end, off_i;
};


bool tile_shape_for(const DeviceCtx& ctx, const HostPlan& hp, TileShape& sh) {
  const int n_src = hp.nV + hp.nI;
  if (hp.nvar > 0xffff) return false;
  if (const char* e = getenv("SPICEY_TILE_SHAPE")) {   // experiments: TR,TC[,CTAs per SM]
    int a = 0, b = 0, c = 0;
    if (sscanf(e, "%d,%d,%d", &a, &b, &c) >= 2) {
      sh = tile_shape_eval(hp.nvar, a, b, hp.n_elem, n_src, ctx.smem_optin, c);
      if (sh.ok && c > 0) sh.minb = c;
      if (sh.ok) return true;
    }
  }
  sh = choose_tile_shape(hp.nvar, hp.n_elem, n_src, ctx.smem_optin);
  return sh.ok;
}

// variant: bit 0 element currents, bit 1 constant tables (plain frequency sweep), bit 2 (alpha, beta)-only tables
std::string tile_source(const TileShape& sh, int variant) {
  char head[320];
  snprintf(head, sizeof head, "#define TL_N %d\n#define TL_TR %d\n#define TL_TC %d\n#define TL_WARPS %d\n#define TL_MINB %d\n#define TL_IELEM %d\n"
           "#define TL_CONST %d\n#define TL_RC %d\n",
           sh.n, sh.tr, sh.tc, sh.warps, sh.minb, variant & 1, (variant >> 1) & 1, (variant >> 2) & 1);
  return std::string(head) + kTileKernelSource;
}

// Tables of the topology for one tile shape: the per-entry gather lists (per-instance stamping: the plan's lists are
// per row, one thread per ENTRY balances a dense matrix) and, for plain frequency sweeps, every thread's tile entries
// as constants in the order the kernel reads them, the element records of the unpack phase and the inductances.
int prepare_tile(DeviceCtx& ctx, const HostPlan& hp, const TileShape& sh, cudaStream_t stream) {
  const int shape[3] = {sh.tr, sh.tc, sh.warps};
  uint64_t key = fnv1a(plan_key(hp), shape, sizeof shape);
  key = fnv1a(key, hp.var_of_slot.data(), sizeof(int) * hp.var_of_slot.size());
  if (!key) key = 1;
  if (ctx.tl_key == key) return SPICEY_SUCCESS;
  ctx.tl_key = 0;
  const int n_ent = (int)hp.ac.ent_col.size();
  std::vector<int> ent_rc(std::max(1, n_ent));
  for (int r = 0; r < hp.nvar; ++r)
    for (int en = hp.ac.row_ptr[r]; en < hp.ac.row_ptr[r + 1]; ++en) ent_rc[en] = r | (hp.ac.ent_col[en] << 16);
  std::vector<unsigned char> blob;
  const size_t o_rc = push_blob(blob, ent_rc), o_ep = push_blob(blob, hp.ac.ent_ptr), o_co = push_blob(blob, hp.ac.contrib);
  // constants: only when no value slot varies per instance and every R > 0
  bool swept = false;
  for (int v : hp.var_of_slot) swept = swept || v >= 0;
  SparseProgram sc;
  ctx.tl_dev.has_const = !swept && entry_constants(hp, sc);
  size_t o_ct = 0, o_er = 0, o_il = 0, o_wt = 0;
  if (ctx.tl_dev.has_const) {
    bool rc_only = true;
    for (int en = 0; en < n_ent; ++en) rc_only = rc_only && sc.ent_gamma[en] == 0.0 && sc.ent_jim[en] == 0.0;
    ctx.tl_dev.rc_only = rc_only;
    const int n = hp.nvar, threads = sh.warps * 32, tpw = 32 / sh.tr;
    std::vector<int> ent_of((size_t)n * (n + 1), -1);
    for (int en = 0; en < n_ent; ++en) ent_of[(size_t)(ent_rc[en] & 0xffff) * (n + 1) + (ent_rc[en] >> 16)] = en;
    std::vector<double2> tab((size_t)sh.mr * sh.mc * threads * (rc_only ? 1 : 2), make_double2(0.0, 0.0));
    for (int tid = 0; tid < threads; ++tid) {
      const int lane = tid & 31, warp = tid >> 5;
      const int tr = lane % sh.tr, tc = std::min(warp * tpw + lane / sh.tr, sh.tc - 1);   // as the kernel maps (and mirrors) its lanes
      for (int m = 0; m < sh.mr; ++m)
        for (int c = 0; c < sh.mc; ++c) {
          const int i = m * sh.tr + tr, j = c * sh.tc + tc;
          const int en = (i < n && j <= n) ? ent_of[(size_t)i * (n + 1) + j] : -1;
          if (en < 0) continue;
          const size_t o = (size_t)(m * sh.mc + c) * threads + tid;
          if (rc_only) tab[o] = make_double2(sc.ent_alpha[en] + sc.ent_jre[en], sc.ent_beta[en]);
          else {
            tab[2 * o] = make_double2(sc.ent_alpha[en] + sc.ent_jre[en], sc.ent_jim[en]);
            tab[2 * o + 1] = make_double2(sc.ent_beta[en], sc.ent_gamma[en]);
          }
        }
    }
    std::vector<double2> wtab;
    if (n <= 32) {   // warp_lu_kernel.cuh: lane = row, [column][lane]
      wtab.assign((size_t)(n + 1) * 32 * (rc_only ? 1 : 2), make_double2(0.0, 0.0));
      for (int en = 0; en < n_ent; ++en) {
        const int i = ent_rc[en] & 0xffff, j = ent_rc[en] >> 16;
        const size_t o = (size_t)j * 32 + i;
        if (rc_only) wtab[o] = make_double2(sc.ent_alpha[en] + sc.ent_jre[en], sc.ent_beta[en]);
        else {
          wtab[2 * o] = make_double2(sc.ent_alpha[en] + sc.ent_jre[en], sc.ent_jim[en]);
          wtab[2 * o + 1] = make_double2(sc.ent_beta[en], sc.ent_gamma[en]);
        }
      }
    } else {
      wtab.push_back(make_double2(0.0, 0.0));
    }
    std::vector<double4> el_rec(std::max(1, hp.n_ac_elem));
    for (int e = 0; e < hp.n_ac_elem; ++e) {
      const int n1 = hp.ends[e].x, n2 = hp.ends[e].y;
      long long ij;
      double4 r = make_double4(0.0, 0.0, 0.0, 0.0);
      if (e >= hp.off[ELEM_V]) {   // a V element's current is its branch unknown
        ij = (long long)(hp.nn + (e - hp.off[ELEM_V])) | ((long long)n << 32);
        r.y = 1.0;
      } else {
        ij = (long long)(n1 ? n1 - 1 : n) | ((long long)(n2 ? n2 - 1 : n) << 32);
        r.y = sc.el_a[e]; r.z = sc.el_b[e]; r.w = sc.el_g[e];
      }
      memcpy(&r.x, &ij, sizeof ij);
      el_rec[e] = r;
    }
    ctx.tl_dev.n_ind = (int)sc.ind_L.size();
    if (sc.ind_L.empty()) sc.ind_L.push_back(0.0);
    o_ct = push_blob(blob, tab); o_er = push_blob(blob, el_rec); o_il = push_blob(blob, sc.ind_L);
    o_wt = push_blob(blob, wtab);
  }
  int rc = ctx.tl_blob.ensure(blob.size() + 16);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(ctx.tl_blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  unsigned char* b = (unsigned char*)ctx.tl_blob.p;
  ctx.tl_dev.ent_rc = b + o_rc; ctx.tl_dev.ent_ptr = b + o_ep; ctx.tl_dev.contrib = b + o_co; ctx.tl_dev.n_ent = n_ent;
  ctx.tl_dev.ctab = b + o_ct; ctx.tl_dev.el_rec = b + o_er; ctx.tl_dev.ind_L = b + o_il; ctx.tl_dev.wtab = b + o_wt;
  ctx.tl_key = key;
  return SPICEY_SUCCESS;
}

// One warp per system (warp_lu_kernel.cuh, Nvar <= 32, plain frequency sweeps): 4 warps per CTA, as many CTAs per SM as
// the row in registers (4 per entry + ~36) allows.
constexpr int kWarpLuWarps = 4;
int warp_lu_minb(int n) {
  if (const char* e = getenv("SPICEY_WARP_LU_MINB")) return std::max(1, std::min(16, atoi(e)));   // experiments
  const int regs = (4 * (n + 1) + 36 + 7) / 8 * 8;
  return std::max(1, std::min(8, 65536 / (32 * kWarpLuWarps * regs)));
}
int warp_lu_sync() {
  if (const char* e = getenv("SPICEY_WARP_LU_SYNC")) return std::max(0, std::min(32, atoi(e)));   // experiments
  return 4;
}
// per warp: two step records, x; per-instance stamping adds the element admittances, source phasors and the stamped image
size_t warp_lu_smem_bytes(int n, bool cst, int n_elem, int n_src) {
  const size_t per_warp = 2 * (size_t)(n + 3) + n + 1 + (cst ? 0 : (size_t)n_elem + std::max(1, n_src) + (size_t)(n + 1) * 32);
  return sizeof(double2) * (size_t)kWarpLuWarps * per_warp;
}
std::string warp_lu_source(int n, int variant, int minb) {
  char head[320];
  snprintf(head, sizeof head, "#define WL_N %d\n#define WL_WARPS %d\n#define WL_MINB %d\n#define WL_IELEM %d\n#define WL_CONST %d\n#define WL_RC %d\n#define WL_SYNC %d\n",
           n, kWarpLuWarps, minb, variant & 1, (variant >> 1) & 1, (variant >> 2) & 1, warp_lu_sync());
  return std::string(head) + kWarpLuKernelSource;
}

DeviceCtx::TileJit* ensure_warp_lu_jit(DeviceCtx& ctx, int n, int variant, int minb) {
  DeviceCtx::TileJit& jv = ctx.tile_jit[variant & 15];
  const int shape[5] = {n, kWarpLuWarps, minb, variant, warp_lu_sync()};
  uint64_t key = fnv1a(1469598103934665603ull, shape, sizeof shape);
  if (!key) key = 1;
  if (jv.key == key) return jv.failed ? nullptr : &jv;
  jv.key = key;
  jv.failed = true;
  if (jv.lib) { cudaLibraryUnload(jv.lib); jv.lib = nullptr; jv.kernel = nullptr; }
  if (!load_jit_kernel(warp_lu_source(n, variant, minb), "spicey_warp_lu_jit", &jv.lib, &jv.kernel, ctx.tl_note)) return nullptr;
  jv.n = n; jv.tr = 32; jv.tc = 1; jv.warps = kWarpLuWarps; jv.minb = minb;
  jv.failed = false;
  ctx.tl_note = "ok";
  return &jv;
}

DeviceCtx::TileJit* ensure_tile_jit(DeviceCtx& ctx, const TileShape& sh, int variant) {
  DeviceCtx::TileJit& jv = ctx.tile_jit[variant & 7];
  const int shape[6] = {sh.n, sh.tr, sh.tc, sh.warps, sh.minb, variant};
  uint64_t key = fnv1a(1469598103934665603ull, shape, sizeof shape);
  if (!key) key = 1;
  if (jv.key == key) return jv.failed ? nullptr : &jv;
  jv.key = key;
  jv.failed = true;
  if (jv.lib) { cudaLibraryUnload(jv.lib); jv.lib = nullptr; jv.kernel = nullptr; }
  const std::string src = tile_source(sh, variant);
  if (!load_jit_kernel(src, "spicey_tile_jit", &jv.lib, &jv.kernel, ctx.tl_note)) return nullptr;
  jv.n = sh.n; jv.tr = sh.tr; jv.tc = sh.tc; jv.warps = sh.warps; jv.minb = sh.minb;
  jv.failed = false;
  ctx.tl_note = "ok";
  return &jv;
}

// Stamps, factors and unpacks args.p_count points with the register-tile kernel; *used = false when the tier does not
// apply (no shape fits, the compile failed): the caller then takes the one-thread-per-row kernel.
int launch_ac_tile(DeviceCtx& ctx, const HostPlan& hp, const DevPlan& dp, const AcArgs& args, uint32_t flags,
                   cudaStream_t stream, int64_t* launches, bool* used) {
  *used = false;
  TileShape sh;
  if (!tile_shape_for(ctx, hp, sh)) return SPICEY_SUCCESS;
  int rc = prepare_tile(ctx, hp, sh, stream);
  if (rc) return rc;
  const bool cst = ctx.tl_dev.has_const && dp.n_inst == 1 && !(flags & SPICEY_FLAG_TILE_GENERIC);
  // Nvar <= 32: one warp per system, the row of a lane in registers (unless a thread grid is forced)
  const size_t wl_smem = warp_lu_smem_bytes(hp.nvar, cst, hp.n_elem, hp.nV + hp.nI);
  bool warp_lu = hp.nvar <= 32 && !getenv("SPICEY_TILE_SHAPE") && wl_smem + 1024 <= ctx.smem_optin;
  if (const char* e = getenv("SPICEY_WARP_LU")) warp_lu = warp_lu && atoi(e) != 0;   // experiments
  const int wl_minb = warp_lu ? std::max(1, std::min<int>(warp_lu_minb(hp.nvar), (int)(ctx.smem_optin / (wl_smem + 1024)))) : 0;
  const int variant = (args.ielem ? 1 : 0) | (cst ? 2 : 0) | (cst && ctx.tl_dev.rc_only ? 4 : 0) | (warp_lu ? 8 : 0);
  DeviceCtx::TileJit* jv = warp_lu ? ensure_warp_lu_jit(ctx, hp.nvar, variant, wl_minb) : ensure_tile_jit(ctx, sh, variant);
  if (!jv) return SPICEY_SUCCESS;
  const size_t smem = warp_lu ? wl_smem
                              : tile_smem_bytes(hp.nvar, cst ? 0 : hp.n_elem, cst ? 0 : hp.nV + hp.nI, jv->tr, jv->tc, cst);
  CUDA_TRY(cudaFuncSetAttribute((const void*)jv->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned grid = (unsigned)std::min<long long>(warp_lu ? (args.p_count + kWarpLuWarps - 1) / kWarpLuWarps : args.p_count,
                                                (long long)ctx.sm_count * jv->minb);
  if (const char* e = getenv("SPICEY_TILE_GRID")) grid = std::min<unsigned>(grid, (unsigned)std::max(1, atoi(e)));   // experiments
  TileArgs a;
  memset(&a, 0, sizeof a);
  a.freqs = args.freqs; a.n_freq = args.n_freq; a.p_begin = args.p_begin; a.p_count = args.p_count;
  a.x = args.x; a.ielem = args.ielem; a.status = args.status; a.series_ld = args.series_ld;
  a.ends = dp.ends; a.meta = dp.meta; a.values = dp.values; a.var_of_slot = dp.var_of_slot; a.var_values = dp.var_values;
  a.n_inst = dp.n_inst;
  a.ent_rc = (const int*)ctx.tl_dev.ent_rc; a.ent_ptr = (const int*)ctx.tl_dev.ent_ptr; a.contrib = (const int*)ctx.tl_dev.contrib;
  a.n_ent = ctx.tl_dev.n_ent; a.nn = hp.nn; a.nV = hp.nV; a.n_elem = hp.n_elem; a.n_ac_elem = hp.n_ac_elem;
  a.off_v = hp.off[ELEM_V]; a.off_v_end = hp.off[ELEM_V + 1]; a.off_i = hp.off[ELEM_I];
  a.plist = args.plist; a.pcount = args.pcount; a.fb_total = args.plist ? args.fb_total : nullptr;
  const bool guards = cst && ctx.tl_dev.n_ind > 0;   // points that trip an inductor guard go to the one-thread-per-row kernel
  if (cst) {
    a.ctab = (const double2*)(warp_lu ? ctx.tl_dev.wtab : ctx.tl_dev.ctab); a.el_rec = (const double4*)ctx.tl_dev.el_rec; a.ind_L = (const double*)ctx.tl_dev.ind_L;
    a.n_ind = ctx.tl_dev.n_ind;
    if (guards) {
      // (the input may itself be the fallback list of a program tier, which lives in sp_fb: a list of its own then)
      Buffer& fb = args.plist ? ctx.tl_fb : ctx.sp_fb;
      if ((rc = fb.ensure(sizeof(long long) * args.p_count + 64))) return rc;
      a.fb_count = (int*)fb.p;
      a.fb_list = (long long*)((char*)fb.p + 64);
      CUDA_TRY(cudaMemsetAsync(a.fb_count, 0, sizeof(int), stream));
    }
  }
  void* kargs[] = {&a};
  CUDA_TRY(cudaLaunchKernel((const void*)jv->kernel, dim3(grid), dim3(jv->warps * 32), kargs, smem, stream));
  if (launches) ++*launches;
  *used = true;
  if (guards) {
    AcArgs d = args;
    d.plist = a.fb_list;
    d.pcount = a.fb_count;
    d.fb_total = args.plist ? nullptr : (unsigned long long*)ctx.sp_cnt.p;   // (points of a fallback list are counted once)
    return launch_ac_dense(ctx, hp, dp, d, flags | SPICEY_FLAG_NO_TILE, stream, nullptr, launches);
  }
  return SPICEY_SUCCESS;
}

// ---------------------------------------------------------------------------------
// AC launch on one device (device pointers), asynchronous on `stream`.
int launch_ac_dense(DeviceCtx& ctx, const HostPlan& hp, const DevPlan& dp, const AcArgs& args, uint32_t flags,
              cudaStream_t stream, int* tier_out, int64_t* launches) {
  if (args.p_count <= 0) return SPICEY_SUCCESS;
  const bool strict = flags & SPICEY_FLAG_STRICT;
  // Large batches: the matrix in registers, 2-D tiles (tile_kernel.cuh).  Strict mode, the global-scratch tier, the
  // fallback lists of the sparse tiers and small batches stay with one thread per row in shared memory.
  // Matrices that are mostly structural zeros stay with the row kernel as well: it walks the set bits of its row masks
  // (cfg 2's ladder forced dense: 5.9 M solves/s against 4.2 M with every tile entry updated).
  const bool dense_matrix = (long long)hp.ac.ent_col.size() * 4 >= (long long)hp.nvar * (hp.nvar + 1);
  // The fallback list of a program tier (systems whose pivot order differs from the pilot order) takes the register
  // kernels too when the launch is large and the matrix is not nearly empty: on circuits with inductors most points of a
  // many-decade sweep can end up there (measured, random RLC networks: 98 - 99.9 % of the points).
  const bool fb_tile = args.plist && args.p_count >= kTileMinPoints &&
                       (hp.nvar <= 32 || (long long)hp.ac.ent_col.size() * 12 >= (long long)hp.nvar * (hp.nvar + 1));
  if (!strict && !(flags & (SPICEY_FLAG_FORCE_GMEM | SPICEY_FLAG_NO_JIT | SPICEY_FLAG_NO_TILE)) && (!args.plist || fb_tile) &&
      ((flags & SPICEY_FLAG_TILE) || fb_tile || (dense_matrix && (args.p_count >= kTileMinPoints || (flags & SPICEY_FLAG_JIT)))) &&
      args.series_ld < (1ll << 40)) {
    bool used = false;
    int rc = launch_ac_tile(ctx, hp, dp, args, flags, stream, launches, &used);
    if (rc) return rc;
    if (used) {
      if (tier_out) *tier_out = SPICEY_TIER_TILE;
      return SPICEY_SUCCESS;
    }
  }
  const int NT = round32(hp.nvar);
  const int nwarps = NT / 32;
  AcSmem sm(hp.nvar, hp.n_elem, hp.nV + hp.nI, hp.MW, nwarps, false);
  bool gmem = (flags & SPICEY_FLAG_FORCE_GMEM) || sm.total > ctx.smem_optin;
  AcSmem L(hp.nvar, hp.n_elem, hp.nV + hp.nI, hp.MW, nwarps, gmem);
  if (L.total > ctx.smem_optin) return fail(SPICEY_ERR_UNSUPPORTED, "element table too large for shared memory");
  void (*kern)(DevPlan, AcArgs) =
      gmem ? (strict ? ac_cta_kernel<true, true> : ac_cta_kernel<false, true>)
           : (strict ? ac_cta_kernel<true, false> : ac_cta_kernel<false, false>);
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, L.total));
  if (occ < 1) return fail(SPICEY_ERR_UNSUPPORTED, "AC kernel does not fit on an SM");
  long long grid = std::min<long long>(args.p_count, (long long)occ * ctx.sm_count);
  AcArgs a = args;
  if (gmem) {
    size_t per = AcSmem::scratch_bytes(hp.nvar, hp.MW);
    int rc = ctx.scratch.ensure(per * grid);
    if (rc) return rc;
    a.scratch = (double2*)ctx.scratch.p;
  }
  kern<<<(unsigned)grid, NT, L.total, stream>>>(dp, a);
  CUDA_TRY(cudaGetLastError());
  if (tier_out) *tier_out = gmem ? SPICEY_TIER_CTA_GMEM : SPICEY_TIER_CTA_SMEM;
  if (launches) ++*launches;
  return SPICEY_SUCCESS;
}


// Dispatcher: sparse program path for a large single-instance sweep, dense pivoting kernel otherwise.
int launch_ac(DeviceCtx& ctx, const HostPlan& hp, const DevPlan& dp, const AcArgs& args, uint32_t flags,
              cudaStream_t stream, int* tier_out, int64_t* launches, double pilot_f, bool pilot_known, double f_lo = 0.0,
              double f_hi = 0.0) {
  const bool want_sparse = !(flags & (SPICEY_FLAG_STRICT | SPICEY_FLAG_FORCE_GMEM | SPICEY_FLAG_DENSE)) &&
                           (args.p_count >= kSparseMinPoints || (flags & SPICEY_FLAG_SPARSE));
  if (want_sparse) {
    const bool eager = dp.n_inst > 1 || dp.n_var > 0;  // component sweep / Monte-Carlo: per-instance stamping
    // the per-instance stamping of the sparse tiers indexes the R, C, L, V elements only: sweeps of circuits with
    // current sources stay with the dense kernel
    if (eager && hp.nI > 0) return launch_ac_dense(ctx, hp, dp, args, flags, stream, tier_out, launches);
    const uint64_t key = plan_key(hp) ^ (eager ? 0x9e3779b97f4a7c15ull : 0ull);
    if (ctx.sp_key != key) {
      if (!pilot_known) {  // device-resident frequencies: fetch a representative value and both ends
        const long long first = eager ? 0 : args.p_begin, count = eager ? args.n_freq : args.p_count;
        CUDA_TRY(cudaMemcpyAsync(&pilot_f, args.freqs + first + count / 2, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(&f_lo, args.freqs + first, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(&f_hi, args.freqs + first + count - 1, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
      }
      int rc = prepare_sparse(ctx, hp, pilot_f, eager, stream);
      if (rc) return rc;
      ctx.sp_f_lo = f_lo; ctx.sp_f_hi = f_hi;
    }
    // Once per cached program: how did the pilot order hold up so far?  When a quarter of the points launched through
    // it came back on the fallback list, the program tiers only add their own pass in front of the dense solve: later
    // launches of this topology go to the dense tiers directly.  (One stream synchronisation per topology.)
    if (ctx.sp_valid && ctx.sp_adapt == 0 && ctx.sp_launched >= kTileMinPoints && ctx.sp_cnt.p &&
        !(flags & (SPICEY_FLAG_BAND | SPICEY_FLAG_WARP))) {
      unsigned long long life = 0;
      CUDA_TRY(cudaMemcpyAsync(&life, (const char*)ctx.sp_cnt.p + 8, sizeof life, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      ctx.sp_adapt = ((long long)life * 4 >= ctx.sp_launched) ? 2 : 1;
    }
    if (ctx.sp_valid && ctx.sp_adapt == 2 && !(flags & (SPICEY_FLAG_BAND | SPICEY_FLAG_WARP | SPICEY_FLAG_NO_TILE | SPICEY_FLAG_NO_JIT))) {
      TileShape sh;
      if (tile_shape_for(ctx, hp, sh)) return launch_ac_dense(ctx, hp, dp, args, flags | SPICEY_FLAG_TILE, stream, tier_out, launches);
    }
    if (ctx.sp_valid) {
      // A program that executes most of the dense elimination's work has nothing to gain from the program tiers (their
      // per-operation bookkeeping, workspaces in memory): dense circuits go to the register-tile kernel.
      const long long nv = hp.nvar;
      const long long dense_cfma = nv * (nv - 1) * (2 * nv + 5) / 6;   // sum_k (n - 1 - k)(n + 1 - k): the update of :47-52 without skips
      const bool dense_like = ctx.sp.n_fma * 2 >= dense_cfma && nv >= 8;
      if ((flags & SPICEY_FLAG_TILE) || (dense_like && args.p_count >= kTileMinPoints && !(flags & (SPICEY_FLAG_BAND | SPICEY_FLAG_WARP | SPICEY_FLAG_NO_TILE | SPICEY_FLAG_NO_JIT)))) {
        TileShape sh;
        if (tile_shape_for(ctx, hp, sh)) return launch_ac_dense(ctx, hp, dp, args, flags | SPICEY_FLAG_TILE, stream, tier_out, launches);
      }
      return launch_ac_sparse(ctx, hp, dp, args, flags, stream, tier_out, launches);
    }
  }
  return launch_ac_dense(ctx, hp, dp, args, flags, stream, tier_out, launches);
}

struct TranJitArgs {   // must match tran_jit_prelude()
  const double* var_values; long long n_inst;
  double dt; long long steps;
  const double* vsrc; unsigned vmask;
  const double* state0; long long inst0; long long n_local;
  double* v; double* ielem; double* state_out; int* iters; int* status;
};
constexpr int kTranJitBlock = 64;                 // 65,536 instances -> 1024 CTAs: 6.9 per SM, 1 % tail
constexpr long long kTranJitMinSteps = 2000000;   // instance-steps below which the ~1 s compile does not pay off

// Host copy of the per-source waveform descriptors {kind, first parameter slot, PWL pair count, 0}.
typedef std::vector<int4> WaveList;

// Validates a caller's spicey_waves (or the legacy has-row mask) against the plan.
int build_waves(const HostPlan& hp, const spicey_waves* wv, const int32_t* vsrc_mask, bool have_rows, WaveList& out) {
  out.assign(std::max(1, hp.nV), make_int4(WAVE_DC, 0, 0, 0));
  if (!wv) {
    for (int k = 0; k < hp.nV; ++k)
      if (vsrc_mask && vsrc_mask[k]) {
        if (!have_rows) return fail(SPICEY_ERR_INVALID, "vsrc_mask set but vsrc is NULL");
        out[k].x = WAVE_TABLE;
      }
    return SPICEY_SUCCESS;
  }
  if (wv->n_vsrc != hp.nV) return fail(SPICEY_ERR_INVALID, "waves->n_vsrc differs from the table's V count");
  if (hp.nV > 0 && !wv->kind) return fail(SPICEY_ERR_INVALID, "waves->kind is NULL");
  for (int k = 0; k < hp.nV; ++k) {
    const int kind = wv->kind[k];
    if (kind == SPICEY_WAVE_DC) continue;
    if (kind == SPICEY_WAVE_TABLE) {
      if (!have_rows) return fail(SPICEY_ERR_INVALID, "a source of kind SPICEY_WAVE_TABLE needs vsrc");
      out[k].x = WAVE_TABLE;
      continue;
    }
    if (kind != SPICEY_WAVE_PULSE && kind != SPICEY_WAVE_PWL) return fail(SPICEY_ERR_INVALID, "unknown waveform kind");
    if (!wv->value_idx) return fail(SPICEY_ERR_INVALID, "waves->value_idx is NULL");
    const int vi = wv->value_idx[k];
    const int np = kind == SPICEY_WAVE_PWL ? (wv->n_pairs ? wv->n_pairs[k] : -1) : 0;
    if (np < 0) return fail(SPICEY_ERR_INVALID, "PWL source without a pair count");
    const long long need = kind == SPICEY_WAVE_PULSE ? 8 : 2ll * np;
    if (vi < 0 || vi + need > hp.n_values) return fail(SPICEY_ERR_INVALID, "waveform parameter slots out of range");
    out[k] = make_int4(kind == SPICEY_WAVE_PULSE ? WAVE_PULSE : WAVE_PWL, vi, np, 0);
  }
  return SPICEY_SUCCESS;
}

void wave_masks(const HostPlan& hp, const WaveList& w, unsigned& table_bits, unsigned& dev_bits) {
  table_bits = dev_bits = 0;
  for (int k = 0; k < hp.nV && k < 32; ++k) {
    if (w[k].x == WAVE_TABLE) table_bits |= 1u << k;
    else if (w[k].x >= WAVE_PULSE) dev_bits |= 1u << k;
  }
}

void tran_jit_source(const HostPlan& hp, const WaveList& waves, bool with_ielem, std::string& src) {
  std::vector<int> n1(hp.n_elem), n2(hp.n_elem), c1(hp.n_elem), c2(hp.n_elem), vi(hp.n_elem);
  for (int e = 0; e < hp.n_elem; ++e) {
    n1[e] = hp.ends[e].x; n2[e] = hp.ends[e].y; c1[e] = hp.ends[e].z; c2[e] = hp.ends[e].w; vi[e] = hp.meta[e].y;
  }
  TranCodegenInput in;
  in.nn = hp.nn; in.nV = hp.nV; in.nvar = hp.nvar; in.n_elem = hp.n_elem; in.n_state = hp.n_state;
  in.off = hp.off; in.n1 = n1.data(); in.n2 = n2.data(); in.nc1 = c1.data(); in.nc2 = c2.data();
  in.value_idx = vi.data(); in.state_idx = hp.state_idx.data(); in.values = hp.values.data();
  in.var_of_slot = hp.var_of_slot.data(); in.with_ielem = with_ielem; in.block = kTranJitBlock;
  // dc / pre-sampled sources are chosen at run time by a.vmask: only device-evaluated waveforms shape the code
  std::vector<int> wk(std::max(1, hp.nV), -1), wi(std::max(1, hp.nV), 0), wn(std::max(1, hp.nV), 0);
  for (int k = 0; k < hp.nV; ++k)
    if (waves[k].x >= WAVE_PULSE) { wk[k] = waves[k].x; wi[k] = waves[k].y; wn[k] = waves[k].z; }
  in.wave_kind = wk.data(); in.wave_vidx = wi.data(); in.wave_npairs = wn.data();
  src = generate_tran_kernel_source(in);
}

DeviceCtx::JitVariant* ensure_tran_jit(DeviceCtx& ctx, const HostPlan& hp, const WaveList& waves, bool with_ielem) {
  DeviceCtx::JitVariant& jv = ctx.tr_jit[with_ielem ? 1 : 0];
  uint64_t key = plan_key(hp);
  key = fnv1a(key, hp.var_of_slot.data(), sizeof(int) * hp.var_of_slot.size());
  key = fnv1a(key, hp.state_idx.data(), sizeof(int) * hp.state_idx.size());
  for (int k = 0; k < hp.nV; ++k)
    if (waves[k].x >= WAVE_PULSE) { key = fnv1a(key, &k, sizeof k); key = fnv1a(key, &waves[k], sizeof(int4)); }
  if (!key) key = 1;
  if (jv.key == key) return jv.failed ? nullptr : &jv;
  jv.key = key;
  jv.failed = true;
  if (jv.lib) { cudaLibraryUnload(jv.lib); jv.lib = nullptr; jv.kernel = nullptr; }
  std::string src;
  tran_jit_source(hp, waves, with_ielem, src);
  if (!load_jit_kernel(src, "spicey_tran_jit", &jv.lib, &jv.kernel, ctx.sp_jit_note)) return nullptr;
  jv.failed = false;
  return &jv;
}

int launch_tran(DeviceCtx& ctx, const HostPlan& hp, const DevPlan& dp, const TranArgs& args, const WaveList& waves,
                uint32_t flags, cudaStream_t stream, int* tier_out, int64_t* launches) {
  if (args.n_local <= 0) return SPICEY_SUCCESS;
  const bool strict = flags & SPICEY_FLAG_STRICT;
  bool jit_waves_ok = true;
  for (int k = 0; k < hp.nV; ++k) jit_waves_ok &= !(waves[k].x == WAVE_PWL && waves[k].z > kTranJitMaxPwlPairs);
  // Compiled per-topology kernel (tran_codegen.h): small systems, batches large enough to pay for the compile.
  if (!strict && !(flags & (SPICEY_FLAG_FORCE_CTA | SPICEY_FLAG_FORCE_GMEM | SPICEY_FLAG_GENERIC_THREAD | SPICEY_FLAG_NO_JIT)) &&
      hp.nvar <= 8 && hp.n_elem <= 48 && hp.nV <= 32 && hp.nI == 0 && args.n_local < (1ll << 29) && jit_waves_ok &&
      ((flags & SPICEY_FLAG_JIT) || args.n_local * (args.steps + 1) >= kTranJitMinSteps)) {
    if (DeviceCtx::JitVariant* jv = ensure_tran_jit(ctx, hp, waves, args.ielem != nullptr)) {
      TranJitArgs j;
      j.var_values = dp.var_values; j.n_inst = dp.n_inst; j.dt = args.dt; j.steps = args.steps;
      j.vsrc = args.vsrc; j.vmask = args.vmask_bits; j.state0 = args.state0; j.inst0 = args.inst0; j.n_local = args.n_local;
      j.v = args.v; j.ielem = args.ielem; j.state_out = args.state_out; j.iters = args.iters; j.status = args.status;
      void* kargs[] = {&j};
      const unsigned grid = (unsigned)((args.n_local + kTranJitBlock - 1) / kTranJitBlock);
      CUDA_TRY(cudaLaunchKernel((const void*)jv->kernel, dim3(grid), dim3(kTranJitBlock), kargs, 0, stream));
      if (tier_out) *tier_out = SPICEY_TIER_TRAN_JIT;
      if (launches) ++*launches;
      return SPICEY_SUCCESS;
    }
  }
  const int n_ent = (int)hp.tran.ent_col.size(), n_con = (int)hp.tran.contrib.size();
  // Register-resident small-system kernel (Nvar <= 6).
  if (!(flags & (SPICEY_FLAG_FORCE_CTA | SPICEY_FLAG_FORCE_GMEM | SPICEY_FLAG_GENERIC_THREAD)) && hp.nvar <= 6 && hp.nI == 0) {
    const bool dyn = hp.off[ELEM_D + 1] > hp.off[ELEM_S];
    typedef void (*KernT)(DevPlan, TranArgs);
    KernT kern = nullptr;
    constexpr int NT = 64;  // small CTAs: more CTAs per SM, better balance for ~1e5-instance batches
    int nv = hp.nvar <= 2 ? 2 : hp.nvar <= 3 ? 3 : hp.nvar <= 4 ? 4 : 6;
    switch (nv) {
      case 2: kern = strict ? (KernT)tran_small_kernel<2, NT, true> : (KernT)tran_small_kernel<2, NT, false>; break;
      case 3: kern = strict ? (KernT)tran_small_kernel<3, NT, true> : (KernT)tran_small_kernel<3, NT, false>; break;
      case 4: kern = strict ? (KernT)tran_small_kernel<4, NT, true> : (KernT)tran_small_kernel<4, NT, false>; break;
      default: kern = strict ? (KernT)tran_small_kernel<6, NT, true> : (KernT)tran_small_kernel<6, NT, false>; break;
    }
    TranSmallSmem L(nv, hp.n_elem, hp.n_state, dyn, NT);
    if (L.total <= ctx.smem_optin) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
      long long grid = (args.n_local + NT - 1) / NT;
      kern<<<(unsigned)grid, NT, L.total, stream>>>(dp, args);
      CUDA_TRY(cudaGetLastError());
      if (tier_out) *tier_out = SPICEY_TIER_THREAD;
      if (launches) ++*launches;
      return SPICEY_SUCCESS;
    }
  }
  // Thread tier when the per-thread footprint leaves room for >= 32 threads per CTA.
  if (!(flags & (SPICEY_FLAG_FORCE_CTA | SPICEY_FLAG_FORCE_GMEM)) && hp.nvar <= 16) {
    for (int nt = 128; nt >= 32; nt >>= 1) {
      TranThreadSmem L(hp.nvar, hp.n_elem, hp.n_state, n_ent, n_con, nt);
      if (L.total > std::min<size_t>(ctx.smem_optin, nt == 32 ? ctx.smem_optin : 110 * 1024)) continue;
      void (*kern)(DevPlan, TranArgs, int, int) = strict ? tran_thread_kernel<true> : tran_thread_kernel<false>;
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
      long long grid = (args.n_local + nt - 1) / nt;
      kern<<<(unsigned)grid, nt, L.total, stream>>>(dp, args, n_ent, n_con);
      CUDA_TRY(cudaGetLastError());
      if (tier_out) *tier_out = SPICEY_TIER_THREAD;
      if (launches) ++*launches;
      return SPICEY_SUCCESS;
    }
  }
  const int NT = round32(hp.nvar);
  const int nwarps = NT / 32;
  TranCtaSmem sm(hp.nvar, hp.n_elem, hp.n_state, hp.MW, nwarps, false);
  bool gmem = (flags & SPICEY_FLAG_FORCE_GMEM) || sm.total > ctx.smem_optin;
  TranCtaSmem L(hp.nvar, hp.n_elem, hp.n_state, hp.MW, nwarps, gmem);
  if (L.total > ctx.smem_optin) return fail(SPICEY_ERR_UNSUPPORTED, "element table too large for shared memory");
  void (*kern)(DevPlan, TranArgs, double*) =
      gmem ? (strict ? tran_cta_kernel<true, true> : tran_cta_kernel<false, true>)
           : (strict ? tran_cta_kernel<true, false> : tran_cta_kernel<false, false>);
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, L.total));
  if (occ < 1) return fail(SPICEY_ERR_UNSUPPORTED, "TRAN kernel does not fit on an SM");
  long long grid = std::min<long long>(args.n_local, (long long)occ * ctx.sm_count);
  double* scratch = nullptr;
  if (gmem) {
    int rc = ctx.scratch.ensure(TranCtaSmem::scratch_bytes(hp.nvar, hp.MW) * grid);
    if (rc) return rc;
    scratch = (double*)ctx.scratch.p;
  }
  kern<<<(unsigned)grid, NT, L.total, stream>>>(dp, args, scratch);
  CUDA_TRY(cudaGetLastError());
  if (tier_out) *tier_out = gmem ? SPICEY_TIER_CTA_GMEM : SPICEY_TIER_CTA_SMEM;
  if (launches) ++*launches;
  return SPICEY_SUCCESS;
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

__global__ void dfma_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 123.456) out[0] = a0;
}

}  // namespace

// ---------------------------------------------------------------------------------
extern "C" {

int32_t spicey_native_abi_version(void) { return SPICEY_NATIVE_ABI_VERSION; }

int32_t spicey_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* spicey_last_error(void) { return g_err.c_str(); }

int32_t spicey_create(const int32_t* devices, int32_t n_devices, spicey_handle** out) {
  if (!out) return fail(SPICEY_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int avail = spicey_device_count();
  if (avail <= 0) return fail(SPICEY_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
  DeviceGuard device_guard;
  std::vector<int> ids;
  if (!devices || n_devices <= 0) ids.push_back(0);
  else ids.assign(devices, devices + n_devices);
  spicey_handle* h = new spicey_handle();
  memset(&h->stats, 0, sizeof(h->stats));
  for (int id : ids) {
    if (id < 0 || id >= avail) {
      delete h;
      return fail(SPICEY_ERR_INVALID, "device id out of range");
    }
    DeviceCtx c;
    c.dev = id;
    cudaDeviceProp prop;
    if (cudaSetDevice(id) != cudaSuccess || cudaGetDeviceProperties(&prop, id) != cudaSuccess) {
      delete h;
      return fail(SPICEY_ERR_CUDA, "cannot query device");
    }
    c.sm_count = prop.multiProcessorCount;
    c.smem_optin = prop.sharedMemPerBlockOptin;
    cudaStreamCreateWithFlags(&c.compute, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking);
    h->devs.push_back(c);
  }
  h->stats.n_devices = (int)h->devs.size();
  *out = h;
  return SPICEY_SUCCESS;
}

void spicey_destroy(spicey_handle* h) {
  if (!h) return;
  DeviceGuard device_guard;
  for (auto& c : h->devs) {
    cudaSetDevice(c.dev);
    cudaDeviceSynchronize();
    Buffer* bufs[] = {&c.plan, &c.scratch, &c.in0, &c.in1, &c.in2, &c.out_x[0], &c.out_x[1], &c.out_i[0],
                      &c.out_i[1], &c.out_s[0], &c.out_s[1], &c.aux0, &c.aux1, &c.sp_blob, &c.sp_work, &c.sp_fb,
                      &c.wp_blob, &c.wp_work, &c.sp_jit_work, &c.bp_blob, &c.bp_work, &c.tl_blob, &c.tr_iters[0], &c.tr_iters[1], &c.tr_state[0], &c.tr_state[1], &c.sp_cnt, &c.tl_fb};
    for (Buffer* b : bufs) b->release();
    for (auto& jv : c.sp_jit) if (jv.lib) cudaLibraryUnload(jv.lib);
    for (auto& jv : c.tr_jit) if (jv.lib) cudaLibraryUnload(jv.lib);
    for (auto& jv : c.band_jit) if (jv.lib) cudaLibraryUnload(jv.lib);
    for (auto& jv : c.tile_jit) if (jv.lib) cudaLibraryUnload(jv.lib);
    for (auto e : c.events) cudaEventDestroy(e);
    cudaStreamDestroy(c.compute);
    cudaStreamDestroy(c.copy);
  }
  delete h;
}

int32_t spicey_get_stats(const spicey_handle* h, spicey_stats* out) {
  if (!h || !out) return fail(SPICEY_ERR_INVALID, "NULL argument");
  *out = h->stats;
  DeviceGuard device_guard;
  if (out->fallback_solves < 0) {  // read the device-side counters (synchronises the devices)
    long long total = 0;
    for (const auto& c : h->devs) {
      if (!c.sp_cnt.p) continue;
      unsigned long long v = 0;
      cudaSetDevice(c.dev);
      if (cudaMemcpy(&v, c.sp_cnt.p, sizeof(v), cudaMemcpyDeviceToHost) == cudaSuccess) total += (long long)v;
    }
    out->fallback_solves = total;
  }
  return SPICEY_SUCCESS;
}

void* spicey_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void spicey_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int32_t spicey_ac_solve_device(spicey_handle* h, int32_t dev_index, const spicey_elem_table* table,
                               const spicey_sweep* sweep, const double* d_freqs, int64_t n_freq, double* d_x,
                               double* d_ielem, int32_t* d_status, int64_t series_ld, uint32_t flags, void* stream) {
  if (!h) return fail(SPICEY_ERR_INVALID, "handle is NULL");
  if (dev_index < 0 || dev_index >= (int)h->devs.size()) return fail(SPICEY_ERR_INVALID, "dev_index out of range");
  if (!d_freqs || n_freq < 1 || !d_x || !d_status) return fail(SPICEY_ERR_INVALID, "NULL buffer or empty sweep");
  if (series_ld != 0 && series_ld < (sweep ? sweep->n_inst : 1) * n_freq) return fail(SPICEY_ERR_INVALID, "series_ld smaller than the number of points");
  DeviceCtx& ctx = h->devs[dev_index];
  int rc = cached_plan(h, table, sweep);
  if (rc) return rc;
  DeviceGuard device_guard;
  const HostPlan& hp = h->hp;
  CUDA_TRY(cudaSetDevice(ctx.dev));
  cudaStream_t st = (cudaStream_t)stream;
  DevPlan dp;
  rc = upload_plan(ctx, hp, st, dp, h->blob);
  if (rc) return rc;
  dp.n_inst = sweep ? sweep->n_inst : 1;
  dp.n_var = sweep ? sweep->n_var : 0;
  dp.var_values = sweep ? sweep->var_values : nullptr;
  AcArgs a;
  a.freqs = d_freqs; a.n_freq = n_freq; a.p_begin = 0; a.p_count = dp.n_inst * n_freq;
  a.x = (double2*)d_x; a.ielem = (double2*)d_ielem; a.status = d_status; a.scratch = nullptr; a.plist = nullptr; a.pcount = nullptr; a.fb_total = nullptr;
  a.series_ld = series_ld ? series_ld : ((flags & SPICEY_FLAG_SERIES_MAJOR) ? a.p_count : 0);
  int tier = 0;
  int64_t launches = 0;
  if ((rc = ctx.sp_fb.ensure(sizeof(long long) * a.p_count + 64))) return rc;
  CUDA_TRY(cudaMemsetAsync(ctx.sp_fb.p, 0, 64, st));
  if ((rc = reset_call_counters(ctx, st))) return rc;
  rc = launch_ac(ctx, hp, dp, a, flags, st, &tier, &launches, 0.0, false);
  if (rc) return rc;
  h->stats.kernel_launches = launches;
  h->stats.tier = tier;
  h->stats.fallback_solves = -1;  // resolved lazily by spicey_get_stats
  h->stats.program_cfma = tier == SPICEY_TIER_BAND ? ctx.bp.n_cfma : (tier == SPICEY_TIER_SPARSE || tier == SPICEY_TIER_SPARSE_JIT || tier == SPICEY_TIER_SPARSE_WARP) ? ctx.sp.n_fma : 0;
  h->stats.solves = a.p_count;
  h->stats.h2d_bytes = (int64_t)h->blob.size();
  h->stats.d2h_bytes = 0;
  return SPICEY_SUCCESS;
}

int32_t spicey_ac_solve(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep,
                        const double* freqs, int64_t n_freq, double* x, double* ielem, int32_t* status,
                        uint32_t flags) {
  if (!h) return fail(SPICEY_ERR_INVALID, "handle is NULL");
  if (!freqs || n_freq < 1 || !x || !status) return fail(SPICEY_ERR_INVALID, "NULL buffer or empty sweep");
  const double t0 = now_ms();
  int rc = cached_plan(h, table, sweep);
  if (rc) return rc;
  DeviceGuard device_guard;   // destroyed last: the caller's device is current again when the call returns
  DrainOnError drain(h);
  const HostPlan& hp = h->hp;
  const long long n_inst = sweep ? sweep->n_inst : 1;
  const int n_var = sweep ? sweep->n_var : 0;
  const long long P = n_inst * n_freq;
  const int D = (int)h->devs.size();
  // the compile-or-interpret decision looks at the whole call, not at one pipeline chunk
  if (P >= kJitMinPoints && !(flags & SPICEY_FLAG_NO_JIT)) flags |= SPICEY_FLAG_JIT;
  // ... and so does the sparse-or-dense decision: eight devices' shards of a 10,000-point sweep are 1,250 points each
  if (P >= kSparseMinPoints) flags |= SPICEY_FLAG_SPARSE;
  const size_t xrow = sizeof(double2) * hp.nvar, irow = sizeof(double2) * hp.n_ac_elem;
  const bool series = (flags & SPICEY_FLAG_SERIES_MAJOR) != 0;
  // Chunk size: ~96 MiB of results per chunk so that copies overlap the next chunk's kernel.
  // ... but never fewer points than fill the GPU (one thread per point in the sparse tier), within 4 GiB per buffer.
  const size_t prow = xrow + (ielem ? irow : 0) + 4;
  size_t chunk_bytes = 96ull << 20;
  if (const char* e = getenv("SPICEY_AC_CHUNK_BYTES")) chunk_bytes = (size_t)std::max(1ll << 20, atoll(e));   // experiments
  long long chunk = std::max<long long>(1024, (long long)(chunk_bytes / prow));
  chunk = std::max<long long>(chunk, std::min<long long>((long long)h->devs[0].sm_count * 1024, (long long)((4ull << 30) / prow)));
  int64_t launches = 0, h2d = 0, d2h = 0;
  int tier = 0;
  struct Shard { long long lo, hi; size_t ev0; int nchunks; };
  std::vector<Shard> shards(D);
  for (int d = 0; d < D; ++d) {
    DeviceCtx& ctx = h->devs[d];
    Shard& s = shards[d];
    s.lo = P * d / D; s.hi = P * (d + 1) / D; s.nchunks = 0; s.ev0 = 0;
    if (s.hi <= s.lo) continue;
    CUDA_TRY(cudaSetDevice(ctx.dev));
    DevPlan dp;
    rc = upload_plan(ctx, hp, ctx.compute, dp, h->blob);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx.compute));  // h->blob is reused by the next device
    h2d += (int64_t)h->blob.size();
    dp.n_inst = n_inst; dp.n_var = n_var;
    rc = ctx.in0.ensure(sizeof(double) * n_freq);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(ctx.in0.p, freqs, sizeof(double) * n_freq, cudaMemcpyHostToDevice, ctx.compute));
    h2d += sizeof(double) * n_freq;
    if (n_var > 0) {
      size_t vb = sizeof(double) * (size_t)n_var * n_inst;
      rc = ctx.in1.ensure(vb);
      if (rc) return rc;
      CUDA_TRY(cudaMemcpyAsync(ctx.in1.p, sweep->var_values, vb, cudaMemcpyHostToDevice, ctx.compute));
      h2d += vb;
      dp.var_values = (const double*)ctx.in1.p;
    }
    const long long cnt = s.hi - s.lo;
    const long long csz = std::min(chunk, cnt);
    const long long cld = series ? spicey_series_ld(csz) : csz;  // device rows start on 512-byte boundaries
    if ((rc = ctx.sp_fb.ensure(sizeof(long long) * csz + 64))) return rc;
    CUDA_TRY(cudaMemsetAsync(ctx.sp_fb.p, 0, 64, ctx.compute));
    if ((rc = reset_call_counters(ctx, ctx.compute))) return rc;
    for (int b = 0; b < 2; ++b) {
      if ((rc = ctx.out_x[b].ensure(xrow * cld))) return rc;
      if (ielem && (rc = ctx.out_i[b].ensure(irow * cld))) return rc;
      if ((rc = ctx.out_s[b].ensure(sizeof(int) * csz))) return rc;
    }
    int ci = 0;
    for (long long lo = s.lo; lo < s.hi; lo += csz, ++ci) {
      const long long n = std::min(csz, s.hi - lo);
      const int b = ci & 1;
      // events per chunk: [4*ci] kernel start, [4*ci+1] kernel end, [4*ci+2] copy done
      cudaEvent_t ks = ctx.get_event(4 * ci), ke = ctx.get_event(4 * ci + 1), cd = ctx.get_event(4 * ci + 2);
      if (ci >= 2) CUDA_TRY(cudaStreamWaitEvent(ctx.compute, ctx.get_event(4 * (ci - 2) + 2), 0));
      AcArgs a;
      a.freqs = (const double*)ctx.in0.p; a.n_freq = n_freq; a.p_begin = lo; a.p_count = n;
      a.x = (double2*)ctx.out_x[b].p; a.ielem = ielem ? (double2*)ctx.out_i[b].p : nullptr;
      a.status = (int*)ctx.out_s[b].p; a.scratch = nullptr; a.plist = nullptr; a.pcount = nullptr; a.fb_total = nullptr;
      a.series_ld = series ? cld : 0;
      CUDA_TRY(cudaEventRecord(ks, ctx.compute));
      rc = launch_ac(ctx, hp, dp, a, flags, ctx.compute, &tier, &launches, freqs[n_freq / 2], true, freqs[0], freqs[n_freq - 1]);
      if (rc) return rc;
      CUDA_TRY(cudaEventRecord(ke, ctx.compute));
      CUDA_TRY(cudaStreamWaitEvent(ctx.copy, ke, 0));
      if (series) {  // device chunk [rows][n] -> host [rows][P] at column lo
        const size_t w = sizeof(double2) * n, dpitch = sizeof(double2) * cld;
        CUDA_TRY(cudaMemcpy2DAsync((char*)x + sizeof(double2) * lo, sizeof(double2) * P, a.x, dpitch, w, hp.nvar,
                                   cudaMemcpyDeviceToHost, ctx.copy));
        if (ielem && hp.n_ac_elem > 0)
          CUDA_TRY(cudaMemcpy2DAsync((char*)ielem + sizeof(double2) * lo, sizeof(double2) * P, a.ielem, dpitch, w,
                                     hp.n_ac_elem, cudaMemcpyDeviceToHost, ctx.copy));
      } else {
        CUDA_TRY(cudaMemcpyAsync((char*)x + xrow * lo, a.x, xrow * n, cudaMemcpyDeviceToHost, ctx.copy));
        if (ielem)
          CUDA_TRY(cudaMemcpyAsync((char*)ielem + irow * lo, a.ielem, irow * n, cudaMemcpyDeviceToHost, ctx.copy));
      }
      CUDA_TRY(cudaMemcpyAsync(status + lo, a.status, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx.copy));
      CUDA_TRY(cudaEventRecord(cd, ctx.copy));
      d2h += (int64_t)((xrow + (ielem ? irow : 0) + 4) * n);
    }
    s.nchunks = ci;
  }
  double kmax = 0;
  for (int d = 0; d < D; ++d) {
    DeviceCtx& ctx = h->devs[d];
    if (shards[d].hi <= shards[d].lo) continue;
    CUDA_TRY(cudaSetDevice(ctx.dev));
    CUDA_TRY(cudaStreamSynchronize(ctx.copy));
    CUDA_TRY(cudaStreamSynchronize(ctx.compute));
    double k = 0;
    for (int ci = 0; ci < shards[d].nchunks; ++ci) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx.get_event(4 * ci), ctx.get_event(4 * ci + 1));
      k += ms;
    }
    kmax = std::max(kmax, k);
  }
  h->stats.kernel_ms = kmax;
  h->stats.total_ms = now_ms() - t0;
  h->stats.kernel_launches = launches;
  h->stats.h2d_bytes = h2d;
  h->stats.d2h_bytes = d2h;
  h->stats.solves = P;
  h->stats.tier = tier;
  h->stats.fallback_solves = -1;
  h->stats.program_cfma = tier == SPICEY_TIER_BAND ? h->devs[0].bp.n_cfma : (tier == SPICEY_TIER_SPARSE || tier == SPICEY_TIER_SPARSE_JIT || tier == SPICEY_TIER_SPARSE_WARP) ? h->devs[0].sp.n_fma : 0;
  drain.ok = true;
  return SPICEY_SUCCESS;
}

}  // extern "C"

// Uploads the waveform descriptors of one call (a few int4) into ctx.aux0 and fills the TranArgs fields.
static int stage_waves(DeviceCtx& ctx, const HostPlan& hp, const WaveList& waves, cudaStream_t st, TranArgs& a) {
  int rc = ctx.aux0.ensure(sizeof(int4) * waves.size());
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(ctx.aux0.p, waves.data(), sizeof(int4) * waves.size(), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));  // the list is a call-lifetime staging buffer
  a.waves = (const int4*)ctx.aux0.p;
  wave_masks(hp, waves, a.vmask_bits, a.wmask_bits);
  return SPICEY_SUCCESS;
}

static int32_t tran_solve_device_impl(spicey_handle* h, int32_t dev_index, const spicey_elem_table* table,
                                 const spicey_sweep* sweep, double dt, int64_t steps, const double* d_vsrc,
                                 const int32_t* vsrc_mask, const spicey_waves* wv, const double* d_state0, double* d_v, double* d_ielem,
                                 double* d_state_out, int32_t* d_iters, int32_t* d_status, uint32_t flags,
                                 void* stream) {
  if (!h) return fail(SPICEY_ERR_INVALID, "handle is NULL");
  if (dev_index < 0 || dev_index >= (int)h->devs.size()) return fail(SPICEY_ERR_INVALID, "dev_index out of range");
  if (steps < 1 || !d_v || !d_status) return fail(SPICEY_ERR_INVALID, "NULL buffer or steps < 1");
  DeviceCtx& ctx = h->devs[dev_index];
  int rc = cached_plan(h, table, sweep);
  if (rc) return rc;
  DeviceGuard device_guard;
  const HostPlan& hp = h->hp;
  CUDA_TRY(cudaSetDevice(ctx.dev));
  cudaStream_t st = (cudaStream_t)stream;
  WaveList waves;
  if ((rc = build_waves(hp, wv, vsrc_mask, d_vsrc != nullptr, waves))) return rc;
  DevPlan dp;
  rc = upload_plan(ctx, hp, st, dp, h->blob);
  if (rc) return rc;
  dp.n_inst = sweep ? sweep->n_inst : 1;
  dp.n_var = sweep ? sweep->n_var : 0;
  dp.var_values = sweep ? sweep->var_values : nullptr;
  TranArgs a;
  a.dt = dt; a.steps = steps; a.vsrc = d_vsrc; a.state0 = d_state0;
  if ((rc = stage_waves(ctx, hp, waves, st, a))) return rc;
  a.inst0 = 0; a.n_local = dp.n_inst; a.v = d_v; a.ielem = d_ielem; a.state_out = d_state_out;
  a.iters = d_iters; a.status = d_status;
  int tier = 0;
  int64_t launches = 0;
  rc = launch_tran(ctx, hp, dp, a, waves, flags, st, &tier, &launches);
  if (rc) return rc;
  h->stats.kernel_launches = launches;
  h->stats.tier = tier;
  h->stats.solves = dp.n_inst * (steps + 1);
  return SPICEY_SUCCESS;
}

static int32_t tran_solve_impl(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep, double dt,
                          int64_t steps, const double* vsrc, const int32_t* vsrc_mask, const spicey_waves* wv, const double* state0,
                          double* v, double* ielem, double* state_out, int32_t* iters, int32_t* status,
                          uint32_t flags, const int32_t* node_sel = nullptr, int32_t n_sel = -1) {
  if (!h) return fail(SPICEY_ERR_INVALID, "handle is NULL");
  if (steps < 1 || !status || (!v && n_sel != 0)) return fail(SPICEY_ERR_INVALID, "NULL buffer or steps < 1");
  if (n_sel > 0 && !node_sel) return fail(SPICEY_ERR_INVALID, "node_sel is NULL");
  const double t0 = now_ms();
  int rc = cached_plan(h, table, sweep);
  if (rc) return rc;
  DeviceGuard device_guard;
  DrainOnError drain(h);
  const HostPlan& hp = h->hp;
  const long long n_inst = sweep ? sweep->n_inst : 1;
  const int n_var = sweep ? sweep->n_var : 0;
  const long long S1 = steps + 1;
  const int D = (int)h->devs.size();
  for (int k = 0; k < n_sel; ++k)
    if (node_sel[k] < 1 || node_sel[k] > hp.nn) return fail(SPICEY_ERR_INVALID, "node_sel entry is not a node id 1..n_nodes");
  WaveList waves;
  if ((rc = build_waves(hp, wv, vsrc_mask, vsrc != nullptr, waves))) return rc;
  bool any_wave = false;   // any pre-sampled row to upload
  for (int k = 0; k < hp.nV; ++k) any_wave |= waves[k].x == WAVE_TABLE;
  int64_t launches = 0, h2d = 0, d2h = 0;
  int tier = 0;
  std::vector<std::pair<long long, long long>> shards(D);
  std::vector<int> nchunks(D, 0);
  for (int d = 0; d < D; ++d) {
    DeviceCtx& ctx = h->devs[d];
    const long long lo = n_inst * d / D, hi = n_inst * (d + 1) / D, nl = hi - lo;
    shards[d] = std::make_pair(lo, hi);
    if (nl <= 0) continue;
    CUDA_TRY(cudaSetDevice(ctx.dev));
    DevPlan dp;
    rc = upload_plan(ctx, hp, ctx.compute, dp, h->blob);
    if (rc) return rc;
    h2d += (int64_t)h->blob.size();
    dp.n_inst = n_inst; dp.n_var = n_var;
    TranArgs a;
    a.dt = dt; a.steps = steps; a.vsrc = nullptr; a.state0 = nullptr;
    if ((rc = stage_waves(ctx, hp, waves, ctx.compute, a))) return rc;
    if (any_wave) {
      size_t b = sizeof(double) * (size_t)hp.nV * S1;
      if ((rc = ctx.in0.ensure(b))) return rc;
      CUDA_TRY(cudaMemcpyAsync(ctx.in0.p, vsrc, b, cudaMemcpyHostToDevice, ctx.compute));
      h2d += b;
      a.vsrc = (const double*)ctx.in0.p;
    }
    if (n_var > 0) {
      size_t b = sizeof(double) * (size_t)n_var * n_inst;
      if ((rc = ctx.in1.ensure(b))) return rc;
      CUDA_TRY(cudaMemcpyAsync(ctx.in1.p, sweep->var_values, b, cudaMemcpyHostToDevice, ctx.compute));
      h2d += b;
      dp.var_values = (const double*)ctx.in1.p;
    }
    if (state0 && hp.n_state > 0) {
      size_t b = sizeof(double) * (size_t)hp.n_state * n_inst;
      if ((rc = ctx.in2.ensure(b))) return rc;
      CUDA_TRY(cudaMemcpyAsync(ctx.in2.p, state0, b, cudaMemcpyHostToDevice, ctx.compute));
      h2d += b;
      a.state0 = (const double*)ctx.in2.p;
    }
    // Pipeline over instance chunks (every instance is independent: simulateTRAN.ts:146-238 runs one circuit): the
    // persistent kernel of chunk i + 1 runs while chunk i's waveforms travel as 2-D copies, and the device holds two
    // chunks of results (<= ~1 GiB each), not the whole batch (cfg 5: 14.4 GB).
    const size_t per_inst = sizeof(double) * (size_t)S1 * (hp.nn + (ielem ? hp.n_elem : 0)) + (iters ? sizeof(int) * (size_t)S1 : 0);
    size_t chunk_bytes = (size_t)1 << 30;
    if (const char* e = getenv("SPICEY_TRAN_CHUNK_BYTES")) chunk_bytes = (size_t)std::max(1ll, atoll(e));   // tests: force several chunks
    long long csz = std::max<long long>(1, (long long)(chunk_bytes / std::max<size_t>(1, per_inst)));
    if (!getenv("SPICEY_TRAN_CHUNK_BYTES")) csz = std::max<long long>(csz, std::min<long long>(4096, nl));   // ... but not so few instances that a launch is all latency
    if (csz >= nl) csz = nl;
    else csz = (csz + 31) / 32 * 32;
    for (int b = 0; b < 2; ++b) {
      if (b == 1 && csz >= nl) break;
      if ((rc = ctx.out_x[b].ensure(sizeof(double) * S1 * hp.nn * csz))) return rc;
      if (ielem && (rc = ctx.out_i[b].ensure(sizeof(double) * S1 * hp.n_elem * csz))) return rc;
      if ((rc = ctx.out_s[b].ensure(sizeof(int) * csz))) return rc;
      if (state_out && (rc = ctx.tr_state[b].ensure(sizeof(double) * std::max(1, hp.n_state) * csz))) return rc;
      if (iters && (rc = ctx.tr_iters[b].ensure(sizeof(int) * S1 * csz))) return rc;
    }
    int ci = 0;
    for (long long c0 = lo; c0 < hi; c0 += csz, ++ci) {
      const long long n = std::min(csz, hi - c0);
      const int b = ci & 1;
      // events per chunk: [4*ci] kernel start, [4*ci+1] kernel end, [4*ci+2] copies done
      cudaEvent_t ks = ctx.get_event(4 * ci), ke = ctx.get_event(4 * ci + 1), cd = ctx.get_event(4 * ci + 2);
      if (ci >= 2) CUDA_TRY(cudaStreamWaitEvent(ctx.compute, ctx.get_event(4 * (ci - 2) + 2), 0));
      a.inst0 = c0; a.n_local = n;
      a.v = (double*)ctx.out_x[b].p;
      a.ielem = ielem ? (double*)ctx.out_i[b].p : nullptr;
      a.state_out = state_out ? (double*)ctx.tr_state[b].p : nullptr;
      a.iters = iters ? (int*)ctx.tr_iters[b].p : nullptr;
      a.status = (int*)ctx.out_s[b].p;
      CUDA_TRY(cudaEventRecord(ks, ctx.compute));
      rc = launch_tran(ctx, hp, dp, a, waves, flags, ctx.compute, &tier, &launches);
      if (rc) return rc;
      CUDA_TRY(cudaEventRecord(ke, ctx.compute));
      CUDA_TRY(cudaStreamWaitEvent(ctx.copy, ke, 0));
      // [rows][n] device slabs -> [rows][n_inst] host arrays at column offset c0
      const size_t dp_ = sizeof(double) * n_inst, sp_ = sizeof(double) * n;
      if (n_sel < 0) {
        CUDA_TRY(cudaMemcpy2DAsync((char*)v + sizeof(double) * c0, dp_, a.v, sp_, sp_, S1 * hp.nn,
                                   cudaMemcpyDeviceToHost, ctx.copy));
        d2h += (int64_t)(sp_ * S1 * hp.nn);
      } else {
        // probes only (simulateTRAN.ts:240-249 keeps the probed node voltages): one strided copy per selected node,
        // device rows [step][node] of n instances -> host [step][k] of n_inst
        for (int k = 0; k < n_sel; ++k)
          CUDA_TRY(cudaMemcpy2DAsync((char*)v + sizeof(double) * ((size_t)k * n_inst + c0), dp_ * n_sel,
                                     a.v + (size_t)(node_sel[k] - 1) * n, sp_ * hp.nn, sp_, S1, cudaMemcpyDeviceToHost, ctx.copy));
        d2h += (int64_t)(sp_ * S1 * n_sel);
      }
      if (ielem && hp.n_elem > 0) {
        CUDA_TRY(cudaMemcpy2DAsync((char*)ielem + sizeof(double) * c0, dp_, a.ielem, sp_, sp_, S1 * hp.n_elem,
                                   cudaMemcpyDeviceToHost, ctx.copy));
        d2h += (int64_t)(sp_ * S1 * hp.n_elem);
      }
      if (state_out && hp.n_state > 0)
        CUDA_TRY(cudaMemcpy2DAsync((char*)state_out + sizeof(double) * c0, dp_, a.state_out, sp_, sp_, hp.n_state,
                                   cudaMemcpyDeviceToHost, ctx.copy));
      if (iters)
        CUDA_TRY(cudaMemcpy2DAsync((char*)iters + sizeof(int) * c0, sizeof(int) * n_inst, a.iters, sizeof(int) * n,
                                   sizeof(int) * n, S1, cudaMemcpyDeviceToHost, ctx.copy));
      CUDA_TRY(cudaMemcpyAsync(status + c0, a.status, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx.copy));
      CUDA_TRY(cudaEventRecord(cd, ctx.copy));
    }
    nchunks[d] = ci;
  }
  double kmax = 0;
  for (int d = 0; d < D; ++d) {
    DeviceCtx& ctx = h->devs[d];
    if (shards[d].second <= shards[d].first) continue;
    CUDA_TRY(cudaSetDevice(ctx.dev));
    CUDA_TRY(cudaStreamSynchronize(ctx.copy));
    CUDA_TRY(cudaStreamSynchronize(ctx.compute));
    double k = 0;
    for (int ci = 0; ci < nchunks[d]; ++ci) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx.get_event(4 * ci), ctx.get_event(4 * ci + 1));
      k += ms;
    }
    kmax = std::max(kmax, k);
  }
  h->stats.kernel_ms = kmax;
  h->stats.total_ms = now_ms() - t0;
  h->stats.kernel_launches = launches;
  h->stats.h2d_bytes = h2d;
  h->stats.d2h_bytes = d2h;
  h->stats.solves = n_inst * S1;
  h->stats.tier = tier;
  drain.ok = true;
  return SPICEY_SUCCESS;
}

extern "C" {

int32_t spicey_tran_solve_device(spicey_handle* h, int32_t dev_index, const spicey_elem_table* table,
                                 const spicey_sweep* sweep, double dt, int64_t steps, const double* d_vsrc,
                                 const int32_t* vsrc_mask, const double* d_state0, double* d_v, double* d_ielem,
                                 double* d_state_out, int32_t* d_iters, int32_t* d_status, uint32_t flags,
                                 void* stream) {
  return tran_solve_device_impl(h, dev_index, table, sweep, dt, steps, d_vsrc, vsrc_mask, nullptr, d_state0, d_v, d_ielem,
                                d_state_out, d_iters, d_status, flags, stream);
}

int32_t spicey_tran_solve(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep, double dt,
                          int64_t steps, const double* vsrc, const int32_t* vsrc_mask, const double* state0,
                          double* v, double* ielem, double* state_out, int32_t* iters, int32_t* status,
                          uint32_t flags) {
  return tran_solve_impl(h, table, sweep, dt, steps, vsrc, vsrc_mask, nullptr, state0, v, ielem, state_out, iters, status, flags);
}

int32_t spicey_tran_solve_probes(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep,
                                 double dt, int64_t steps, const spicey_waves* waves, const double* vsrc, const int32_t* vsrc_mask,
                                 const double* state0, const int32_t* node_sel, int32_t n_sel, double* v, double* ielem,
                                 double* state_out, int32_t* iters, int32_t* status, uint32_t flags) {
  if (n_sel < 0) return fail(SPICEY_ERR_INVALID, "n_sel < 0");
  return tran_solve_impl(h, table, sweep, dt, steps, vsrc, waves ? nullptr : vsrc_mask, waves, state0, v, ielem, state_out, iters,
                         status, flags, node_sel, n_sel);
}

int32_t spicey_tran_solve_waves(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep, double dt,
                                int64_t steps, const spicey_waves* waves, const double* vsrc, const double* state0,
                                double* v, double* ielem, double* state_out, int32_t* iters, int32_t* status,
                                uint32_t flags) {
  if (!waves) return fail(SPICEY_ERR_INVALID, "waves is NULL");
  return tran_solve_impl(h, table, sweep, dt, steps, vsrc, nullptr, waves, state0, v, ielem, state_out, iters, status, flags);
}

int32_t spicey_tran_solve_waves_device(spicey_handle* h, int32_t dev_index, const spicey_elem_table* table,
                                       const spicey_sweep* sweep, double dt, int64_t steps, const spicey_waves* waves,
                                       const double* d_vsrc, const double* d_state0, double* d_v, double* d_ielem,
                                       double* d_state_out, int32_t* d_iters, int32_t* d_status, uint32_t flags,
                                       void* stream) {
  if (!waves) return fail(SPICEY_ERR_INVALID, "waves is NULL");
  return tran_solve_device_impl(h, dev_index, table, sweep, dt, steps, d_vsrc, nullptr, waves, d_state0, d_v, d_ielem,
                                d_state_out, d_iters, d_status, flags, stream);
}

int64_t spicey_debug_sparse_source(const spicey_elem_table* table, const spicey_sweep* sweep, double pilot_f, int32_t block,
                                   int32_t min_blocks, int32_t smem_slots, int32_t with_ielem, char* buf, int64_t cap,
                                   int32_t* stats_out) {
  HostPlan hp;
  if (build_plan(table, sweep, hp) != SPICEY_SUCCESS) return -1;
  const bool eager = sweep && (sweep->n_inst > 1 || sweep->n_var > 0);
  SparseProgram sp;
  build_sparse_host(hp, pilot_f, eager, sp);
  if (!sp.ok) { fail(SPICEY_ERR_UNSUPPORTED, "the sparse path does not apply to this circuit"); return -1; }
  CodegenOptions opt;
  opt.block = block; opt.min_blocks = min_blocks; opt.smem_slots = smem_slots; opt.with_ielem = (with_ielem & 1) != 0;
  opt.sync_every = (with_ielem >> 16) & 0xff;
  if ((with_ielem >> 1) & 1) opt.reg_values = (with_ielem >> 24) & 0x7f;   // bit 1: global column beyond that many register values
  if (const char* e = getenv("SPICEY_JIT_GAHEAD")) opt.gmem_ahead = std::max(1, atoi(e));
  if (const char* e = getenv("SPICEY_JIT_L2POLICY")) opt.l2_policy = atoi(e) != 0;
  if (const char* e = getenv("SPICEY_JIT_ANTIPHASE")) opt.antiphase_ns = atoi(e);
  std::string src;
  CodegenStats st;
  jit_source(sp, hp, opt, eager, src, st);
  if (stats_out) {
    stats_out[0] = st.n_saved; stats_out[1] = st.smem_slots; stats_out[2] = st.n_classes; stats_out[3] = (int32_t)sp.code.size();
    stats_out[4] = (int32_t)sp.n_fma; stats_out[5] = (int32_t)sp.n_div; stats_out[6] = sp.n_virtual;
    stats_out[7] = (with_ielem >> 1) & 1 ? st.gmem_slots : sp.n_slots;
  }
  if (buf && cap > 0) {
    const size_t n = std::min<size_t>(src.size(), (size_t)cap - 1);
    memcpy(buf, src.data(), n);
    buf[n] = 0;
  }
  return (int64_t)src.size() + 1;
}

int64_t spicey_debug_tran_source(const spicey_elem_table* table, const spicey_sweep* sweep, int32_t with_ielem,
                                 char* buf, int64_t cap) {
  return spicey_debug_tran_source_waves(table, sweep, nullptr, with_ielem, buf, cap);
}

int64_t spicey_debug_tran_source_waves(const spicey_elem_table* table, const spicey_sweep* sweep, const spicey_waves* waves,
                                       int32_t with_ielem, char* buf, int64_t cap) {
  HostPlan hp;
  if (build_plan(table, sweep, hp) != SPICEY_SUCCESS) return -1;
  if (hp.nvar > 8 || hp.n_elem > 48 || hp.nV > 32 || hp.nI > 0) { fail(SPICEY_ERR_UNSUPPORTED, "the compiled transient kernel covers Nvar <= 8 without current sources"); return -1; }
  std::string src;
  WaveList wl;
  if (build_waves(hp, waves, nullptr, true, wl) != SPICEY_SUCCESS) return -1;
  for (int k = 0; k < hp.nV; ++k)
    if (wl[k].x == WAVE_PWL && wl[k].z > kTranJitMaxPwlPairs) { fail(SPICEY_ERR_UNSUPPORTED, "PWL source with more pairs than the compiled kernel unrolls"); return -1; }
  tran_jit_source(hp, wl, with_ielem != 0, src);
  if (buf && cap > 0) {
    const size_t n = std::min<size_t>(src.size(), (size_t)cap - 1);
    memcpy(buf, src.data(), n);
    buf[n] = 0;
  }
  return (int64_t)src.size() + 1;
}

int32_t spicey_debug_warp_stats(const spicey_elem_table* table, double pilot_f, int32_t* out) {
  HostPlan hp;
  int rc = build_plan(table, nullptr, hp);
  if (rc) return rc;
  SparseProgram sp;
  build_sparse_host(hp, pilot_f, false, sp);
  if (!sp.ok) return fail(SPICEY_ERR_UNSUPPORTED, "the sparse path does not apply to this circuit");
  WarpProgram wp;
  int spare = kWarpPoolSpare;
  if (const char* e = getenv("SPICEY_WARP_SPARE")) spare = std::max(1, atoi(e));
  build_warp_program(sp, 1 << 24, wp, spare);
  if (!wp.ok) return fail(SPICEY_ERR_UNSUPPORTED, "the warp program builder refused this circuit");
  long long chunks = 0;
  for (const WarpStep& st : wp.steps) chunks += (long long)st.n_elim * ((st.n_cols + 31) / 32);
  out[0] = wp.n; out[1] = wp.n_pool; out[2] = wp.n_gslots; out[3] = wp.max_elim;
  out[4] = (int32_t)wp.n_upd_total; out[5] = (int32_t)chunks; out[6] = (int32_t)wp.colent.size(); out[7] = sp.n_slots;
  return SPICEY_SUCCESS;
}

double spicey_debug_band_order_deviation(const spicey_elem_table* table, double pilot_f, double f) {
  HostPlan hp;
  if (build_plan(table, nullptr, hp) != SPICEY_SUCCESS) return -1.0;
  SparseProgram sp;
  build_sparse_host(hp, pilot_f, false, sp);
  if (!sp.ok) return -1.0;
  BandInput in;
  band_input(hp, sp, pilot_f, in);
  BandPlan bp;
  build_band_plan(in, bp);
  if (!bp.ok) return -1.0;
  return bp.renumbered ? band_order_deviation(in, bp, (2 * kPi) * f) : 0.0;
}

int32_t spicey_debug_band_stats(const spicey_elem_table* table, double pilot_f, int32_t* out) {
  HostPlan hp;
  int rc = build_plan(table, nullptr, hp);
  if (rc) return rc;
  SparseProgram sp;
  build_sparse_host(hp, pilot_f, false, sp);
  if (!sp.ok) return fail(SPICEY_ERR_UNSUPPORTED, "the sparse path does not apply to this circuit");
  BandInput in;
  band_input(hp, sp, pilot_f, in);
  BandPlan bp;
  int fl = 0, fr = 0;
  band_force_shape(fl, fr);
  build_band_plan(in, bp, fl, fr);
  if (!bp.ok) return fail(SPICEY_ERR_UNSUPPORTED, "the circuit is not banded + bordered within the kernel's window");
  out[0] = bp.W; out[1] = bp.L; out[2] = bp.RPL; out[3] = bp.bandwidth; out[4] = bp.renumbered ? 1 : 0; out[5] = bp.NB;
  out[6] = (int32_t)bp.abmask; out[7] = (int32_t)std::min<long long>(bp.g_stride, 0x7fffffff);
  return SPICEY_SUCCESS;
}

int64_t spicey_debug_band_source(int32_t L, int32_t RPL, int32_t NB, uint32_t abmask, int32_t with_ielem, int32_t warps,
                                 int32_t minb, char* buf, int64_t cap) {
  BandPlan bp;
  bp.L = L; bp.RPL = RPL; bp.NB = NB; bp.abmask = abmask & 0xffffu; bp.rc_only = (abmask >> 16) & 1u;
  const std::string src = band_source(bp, with_ielem != 0, warps, minb, (int)((abmask >> 17) & 3u));
  if (buf && cap > 0) {
    const size_t n = std::min<size_t>(src.size(), (size_t)cap - 1);
    memcpy(buf, src.data(), n);
    buf[n] = 0;
  }
  return (int64_t)src.size() + 1;
}

int64_t spicey_debug_tile_source(int32_t nvar, int32_t n_elem, int32_t n_src, int32_t tr, int32_t tc, int32_t with_ielem,
                                 int32_t* shape_out, char* buf, int64_t cap) {
  const size_t smem_optin = 227 * 1024;   // an sm_100 SM's opt-in shared memory per CTA
  const TileShape sh = (tr > 0 && tc > 0) ? tile_shape_eval(nvar, tr, tc, n_elem, n_src, smem_optin) : choose_tile_shape(nvar, n_elem, n_src, smem_optin);
  if (!sh.ok) return -1;
  if (shape_out) {
    const bool cst = (with_ielem >> 1) & 1;
    const int32_t v[8] = {sh.tr, sh.tc, sh.mr, sh.mc, sh.warps, sh.minb, sh.regs,
                          (int32_t)tile_smem_bytes(nvar, cst ? 0 : n_elem, cst ? 0 : n_src, sh.tr, sh.tc, cst)};
    memcpy(shape_out, v, sizeof v);
  }
  const std::string src = tile_source(sh, with_ielem);
  if (buf && cap > 0) {
    const size_t n = std::min<size_t>(src.size(), (size_t)cap - 1);
    memcpy(buf, src.data(), n);
    buf[n] = 0;
  }
  return (int64_t)src.size() + 1;
}

int64_t spicey_debug_warp_lu_source(int32_t nvar, int32_t variant, int32_t* shape_out, char* buf, int64_t cap) {
  if (nvar < 1 || nvar > 32) return -1;
  const bool cst = (variant >> 1) & 1;
  if (shape_out) { shape_out[0] = kWarpLuWarps; shape_out[1] = warp_lu_minb(nvar); shape_out[2] = (int32_t)warp_lu_smem_bytes(nvar, cst, 64, 2); }
  const std::string src = warp_lu_source(nvar, variant, warp_lu_minb(nvar));
  if (buf && cap > 0) {
    const size_t n = std::min<size_t>(src.size(), (size_t)cap - 1);
    memcpy(buf, src.data(), n);
    buf[n] = 0;
  }
  return (int64_t)src.size() + 1;
}

int64_t spicey_series_ld(int64_t n_points) { return (n_points + 31) / 32 * 32; }

int32_t spicey_measure_fp64_peak(spicey_handle* h, int32_t dev_index, double* gflops_out) {
  if (!h || !gflops_out) return fail(SPICEY_ERR_INVALID, "NULL argument");
  if (dev_index < 0 || dev_index >= (int)h->devs.size()) return fail(SPICEY_ERR_INVALID, "dev_index out of range");
  DeviceCtx& ctx = h->devs[dev_index];
  DeviceGuard device_guard;
  CUDA_TRY(cudaSetDevice(ctx.dev));
  int rc = ctx.aux1.ensure(64);
  if (rc) return rc;
  const int iters = 1 << 16, threads = 256, blocks = ctx.sm_count * 8;
  cudaEvent_t e0 = ctx.get_event(0), e1 = ctx.get_event(1);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, ctx.compute));
    dfma_peak_kernel<<<blocks, threads, 0, ctx.compute>>>((double*)ctx.aux1.p, iters);
    CUDA_TRY(cudaEventRecord(e1, ctx.compute));
    CUDA_TRY(cudaStreamSynchronize(ctx.compute));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * (double)iters * threads * blocks;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e9);
  }
  *gflops_out = best;
  return SPICEY_SUCCESS;
}

}  // extern "C"
