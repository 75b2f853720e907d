// Host side of the dense register-tile tier (tile_kernel.cuh): the thread grid / tile shape for a given Nvar and the
// shared-memory footprint of one CTA.  No device code; included by spicey_native.cu and tests/cpp/tile_shape_check.cpp.
//
// The augmented (Nvar) x (Nvar + 1) matrix of lib/math/solveComplex.ts:5-13 is distributed 2-D cyclically over TR x TC
// threads, MR x MC entries each, in registers.  A thread column (TR lanes) must sit inside one warp (its pivot search is
// a redux.sync), so a warp carries 32 / TR thread columns; lanes left over mirror a neighbour.  What the choice trades:
// a large tile means few shared-memory loads per complex FMA (MR + MC per MR * MC) but many registers, i.e. few
// systems resident per SM to hide the per-step latency (pivot search -> barrier -> row exchange -> barrier); a fine
// grid means exact skipping of finished rows and columns is lost at the granularity of the tile.
#pragma once
#include <algorithm>
#include <cstddef>

namespace spicey {
namespace host {

struct TileShape {
  int n = 0, tr = 0, tc = 0, mr = 0, mc = 0, warps = 0, minb = 0;
  int regs = 0;              // registers per thread the shape is expected to need
  long long tile_cfma = 0;   // complex FMAs one THREAD executes per system (live part of the tile, summed over the steps)
  long long tile_lds = 0;    // 16-byte shared-memory loads one thread issues for them (multipliers + pivot-row entries)
  double rate = 0;           // modelled systems per SM and cycle
  bool ok = false;
};

// Shared memory of one CTA; must match the carve-up at the top of spicey_tile_jit.
// const_tables: the variant that reads its entries from per-topology constants keeps no element admittances.
inline size_t tile_smem_bytes(int n, int n_elem, int n_src, int tr, int tc, bool const_tables = false) {
  const int nc = n + 1, ld = n | 1;
  const int mr = (n + tr - 1) / tr, mc = (nc + tc - 1) / tc;
  const size_t cplx = (size_t)nc * ld + (const_tables ? 0 : n_elem + std::max(1, n_src)) + 2 * (size_t)(mr * tr + 1) + 2 * (size_t)(mc * tc) + n + (n + 1);
  return 16 * cplx + 32 /* two winner records */ + 16 /* status */;
}

inline long long tile_thread_cfma(int n, int tr, int tc) {
  const int mr = (n + tr - 1) / tr, mc = (n + 1 + tc - 1) / tc;
  long long s = 0;
  for (int k = 0; k < n; ++k) s += (long long)(mr - k / tr) * (mc - k / tc);
  return s;
}

inline long long tile_thread_lds(int n, int tr, int tc) {
  const int mr = (n + tr - 1) / tr, mc = (n + 1 + tc - 1) / tc;
  long long s = 0;
  for (int k = 0; k < n; ++k) s += (mr - k / tr) + (mc - k / tc);
  return s;
}

// Fills the derived fields of a shape; ok = it fits an SM.
inline TileShape tile_shape_eval(int n, int tr, int tc, int n_elem, int n_src, size_t smem_optin, int minb_force = 0) {
  TileShape s;
  s.n = n; s.tr = tr; s.tc = tc;
  if (tr < 1 || tr > 32 || tc < 1) return s;
  s.mr = (n + tr - 1) / tr; s.mc = (n + 1 + tc - 1) / tc;
  const int tpw = 32 / tr;
  s.warps = (tc + tpw - 1) / tpw;
  if (s.warps > 32 || s.mr * s.mc > 36) return s;
  // measured with ptxas on the kernel: 4 registers per tile entry + ~80, of which ~35 can be squeezed without spilling
  s.regs = std::min(255, (4 * s.mr * s.mc + 48 + 7) / 8 * 8);
  const int warps_alloc = (s.warps + 1) / 2 * 2;   // ptxas sizes the register budget of a launch bound for an even warp count
  const size_t smem = tile_smem_bytes(n, n_elem, n_src, tr, tc) + 1024;
  if (smem > smem_optin) return s;
  int res = std::min<long long>(65536 / ((long long)warps_alloc * 32 * s.regs), (long long)(smem_optin / smem));
  res = std::min(res, std::min(16, 64 / s.warps));
  if (minb_force > 0) res = std::min(res, minb_force);
  if (res < 1) return s;
  s.minb = res;
  s.tile_cfma = tile_thread_cfma(n, tr, tc);
  s.tile_lds = tile_thread_lds(n, tr, tc);
  // model: a DFMA warp instruction takes two issue cycles of its scheduler (four schedulers per SM); a 16-byte
  // shared-memory load of a warp takes four cycles of the SM's one load/store data path
  const double lsu = 4.0 * s.tile_lds * s.warps;                       // load/store-path cycles per system
  const double own = 8.0 * s.tile_cfma * ((s.warps + 3) / 4) + lsu;    // cycles a system's own update needs
  const double pipe = 8.0 * s.tile_cfma * s.warps / 4.0;               // FP64-pipe cycles per scheduler and system
  const double latency = 450.0 * n + own + 4000.0 + 60.0 * n;          // per-step chain + stamping + back-substitution
  s.rate = std::min(res / latency, std::min(1.0 / pipe, 0.7 / lsu));
  s.ok = true;
  return s;
}

inline TileShape choose_tile_shape(int n, int n_elem, int n_src, size_t smem_optin) {
  TileShape best;
  for (int tr = 1; tr <= 32; ++tr)
    for (int tc = 1; tc <= n + 1; ++tc) {
      const TileShape s = tile_shape_eval(n, tr, tc, n_elem, n_src, smem_optin);
      if (!s.ok) continue;
      if (!best.ok || s.rate > best.rate * 1.0001 || (s.rate > best.rate * 0.9999 && s.warps < best.warps)) best = s;
    }
  return best;
}

}  // namespace host
}  // namespace spicey
