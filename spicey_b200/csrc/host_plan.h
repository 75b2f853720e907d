// Host side of the element-table boundary: validation of the flat table, the per-topology plan
// (gather-form stamping lists in the reference's stamping order + structural row masks) and the keys
// the per-handle caches are indexed by.  No device code; included by spicey_native.cu only.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/spicey_native.h"
#include "common.cuh"

namespace spicey {
namespace host {

inline uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

inline thread_local std::string g_err;

inline int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return fail(SPICEY_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// ---------------------------------------------------------------------------------
// Host plan
struct HostGather {
  std::vector<int> row_ptr, ent_col, ent_ptr, contrib;
  std::vector<unsigned> rowmask;
};

struct HostPlan {
  int nn = 0, nV = 0, nvar = 0, n_elem = 0, n_values = 0, n_ac_elem = 0, n_state = 0, MW = 0;
  int off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int nI = 0;   // independent current sources (stampCurrentReal.ts / stampCurrentComplex.ts)
  std::vector<int4> ends;
  std::vector<int2> meta;
  std::vector<int> state_idx;
  std::vector<double> values;
  std::vector<int> var_of_slot;
  HostGather ac, tran;
};

const int kValueSlots[7] = {1, 1, 1, 3, 4, 2, 3};

struct GatherBuilder {
  int nvar;
  std::map<std::pair<int, int>, std::vector<int>> ent;  // (row, col) -> ordered contributions
  explicit GatherBuilder(int n) : nvar(n) {}
  void add(int row, int col, int idx, int src, bool neg) {
    ent[std::make_pair(row, col)].push_back((idx << 3) | (src << 1) | (neg ? 1 : 0));
  }
  // stampAdmittance{Real,Complex}.ts: +Y (i1,i1), +Y (i2,i2), -Y (i1,i2), -Y (i2,i1); ground skipped.
  void admittance(int e, int n1, int n2) {
    int i1 = n1 - 1, i2 = n2 - 1;
    if (i1 >= 0) add(i1, i1, e, SRC_Y, false);
    if (i2 >= 0) add(i2, i2, e, SRC_Y, false);
    if (i1 >= 0 && i2 >= 0) { add(i1, i2, e, SRC_Y, true); add(i2, i1, e, SRC_Y, true); }
  }
  // stampCurrentReal.ts: b[n+] -= I, b[n-] += I.
  void current(int e, int np, int nm) {
    int ip = np - 1, im = nm - 1;
    if (ip >= 0) add(ip, nvar, e, SRC_J, true);
    if (im >= 0) add(im, nvar, e, SRC_J, false);
  }
  // stampVoltageSource{Real,Complex}.ts: +-1 in column/row j, b[j] += V.
  void vsource(int e, int n1, int n2, int j) {
    int i1 = n1 - 1, i2 = n2 - 1;
    if (i1 >= 0) add(i1, j, 0, SRC_ONE, false);
    if (i2 >= 0) add(i2, j, 0, SRC_ONE, true);
    if (i1 >= 0) add(j, i1, 0, SRC_ONE, false);
    if (i2 >= 0) add(j, i2, 0, SRC_ONE, true);
    add(j, nvar, e, SRC_J, false);
  }
  void finish(HostGather& g, int MW) const {
    g.row_ptr.assign(nvar + 1, 0);
    g.rowmask.assign((size_t)nvar * MW, 0u);
    g.ent_ptr.push_back(0);
    int row = 0;
    for (const auto& kv : ent) {
      int r = kv.first.first, c = kv.first.second;
      while (row < r) g.row_ptr[++row] = (int)g.ent_col.size();
      g.ent_col.push_back(c);
      for (int w : kv.second) g.contrib.push_back(w);
      g.ent_ptr.push_back((int)g.contrib.size());
      g.rowmask[(size_t)r * MW + (c >> 5)] |= 1u << (c & 31);
    }
    while (row < nvar) g.row_ptr[++row] = (int)g.ent_col.size();
  }
};

inline int build_plan(const spicey_elem_table* tb, const spicey_sweep* sw, HostPlan& hp) {
  if (!tb) return fail(SPICEY_ERR_INVALID, "element table is NULL");
  if (tb->n_nodes < 0 || tb->n_elem < 0 || tb->n_values < 0)
    return fail(SPICEY_ERR_INVALID, "negative size in element table");
  if (tb->n_elem > 0 && (!tb->type || !tb->n1 || !tb->n2 || !tb->value_idx || !tb->values))
    return fail(SPICEY_ERR_INVALID, "element table array is NULL");
  hp.nn = tb->n_nodes;
  hp.n_elem = tb->n_elem;
  hp.n_values = tb->n_values;
  hp.ends.resize(hp.n_elem);
  hp.meta.resize(hp.n_elem);
  hp.state_idx.assign(hp.n_elem, -1);
  int prev = 0, ns = 0;
  int count[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int e = 0; e < hp.n_elem; ++e) {
    int ty = tb->type[e];
    if (ty < 0 || ty > 6) return fail(SPICEY_ERR_INVALID, "unknown element type");
    if (ty < prev) return fail(SPICEY_ERR_INVALID, "elements must be grouped in the order R,C,L,V,S,D,I");
    prev = ty;
    count[ty]++;
    int n1 = tb->n1[e], n2 = tb->n2[e];
    int c1 = (ty == ELEM_S && tb->nc1) ? tb->nc1[e] : 0, c2 = (ty == ELEM_S && tb->nc2) ? tb->nc2[e] : 0;
    if (n1 < 0 || n1 > hp.nn || n2 < 0 || n2 > hp.nn || c1 < 0 || c1 > hp.nn || c2 < 0 || c2 > hp.nn)
      return fail(SPICEY_ERR_INVALID, "node id out of range");
    int vi = tb->value_idx[e];
    if (vi < 0 || vi + kValueSlots[ty] > hp.n_values) return fail(SPICEY_ERR_INVALID, "value_idx out of range");
    hp.ends[e] = make_int4(n1, n2, c1, c2);
    hp.meta[e] = make_int2(ty, vi);
    if (ty == ELEM_C || ty == ELEM_L || ty == ELEM_S || ty == ELEM_D) hp.state_idx[e] = ns++;
  }
  hp.n_state = ns;
  hp.off[0] = 0;
  for (int k = 0; k < 7; ++k) hp.off[k + 1] = hp.off[k] + count[k];
  hp.nV = count[ELEM_V];
  hp.nI = count[ELEM_I];
  hp.nvar = hp.nn + hp.nV;
  hp.n_ac_elem = hp.off[ELEM_V + 1];
  hp.MW = (hp.nvar + 1 + 31) / 32;
  hp.values.assign(tb->values, tb->values + hp.n_values);
  hp.var_of_slot.assign(hp.n_values, -1);
  if (sw) {
    if (sw->n_inst < 1 || sw->n_var < 0) return fail(SPICEY_ERR_INVALID, "bad sweep sizes");
    if (sw->n_var > 0 && (!sw->var_slot || !sw->var_values)) return fail(SPICEY_ERR_INVALID, "sweep array is NULL");
    for (int v = 0; v < sw->n_var; ++v) {
      int s = sw->var_slot[v];
      if (s < 0 || s >= hp.n_values) return fail(SPICEY_ERR_INVALID, "sweep slot out of range");
      hp.var_of_slot[s] = v;
    }
  }
  if (hp.nvar < 1) return fail(SPICEY_ERR_INVALID, "circuit has no unknowns");
  if (hp.nvar > 1024) return fail(SPICEY_ERR_UNSUPPORTED, "Nvar > 1024 exceeds the largest kernel tier");

  GatherBuilder ac(hp.nvar), tr(hp.nvar);
  // AC stamping order R, C, L, V (simulateAC.ts:36-57); S and D are not stamped.
  for (int e = 0; e < hp.n_ac_elem; ++e) {
    int ty = hp.meta[e].x;
    if (ty == ELEM_V) ac.vsource(e, hp.ends[e].x, hp.ends[e].y, hp.nn + (e - hp.off[ELEM_V]));
    else ac.admittance(e, hp.ends[e].x, hp.ends[e].y);
  }
  // independent current sources: b[n+] -= I, b[n-] += I (stampCurrentComplex.ts:4-15 / stampCurrentReal.ts:3-14).  The
  // reference ships the stamps but its parser skips `I` lines (parseNetlist.ts:444-446), so it defines no position for
  // them in the stamping order; they contribute to the right-hand side only, after every other element.
  for (int e = hp.off[ELEM_I]; e < hp.off[ELEM_I + 1]; ++e) ac.current(e, hp.ends[e].x, hp.ends[e].y);
  // TRAN stamping order R, C, L, S, V, D (simulateTRAN.ts:35-101).
  const int order[6] = {ELEM_R, ELEM_C, ELEM_L, ELEM_S, ELEM_V, ELEM_D};
  for (int oi = 0; oi < 6; ++oi) {
    int ty = order[oi];
    for (int e = hp.off[ty]; e < hp.off[ty + 1]; ++e) {
      int n1 = hp.ends[e].x, n2 = hp.ends[e].y;
      if (ty == ELEM_V) { tr.vsource(e, n1, n2, hp.nn + (e - hp.off[ELEM_V])); continue; }
      tr.admittance(e, n1, n2);
      if (ty == ELEM_C || ty == ELEM_L || ty == ELEM_D) tr.current(e, n1, n2);
    }
  }
  for (int e = hp.off[ELEM_I]; e < hp.off[ELEM_I + 1]; ++e) tr.current(e, hp.ends[e].x, hp.ends[e].y);
  ac.finish(hp.ac, hp.MW);
  tr.finish(hp.tran, hp.MW);
  return SPICEY_SUCCESS;
}


// Key of the raw inputs build_plan() reads (everything but the per-instance values themselves).
inline uint64_t table_key(const spicey_elem_table* tb, const spicey_sweep* sw) {
  if (!tb || tb->n_elem < 0 || tb->n_values < 0 || (tb->n_elem > 0 && (!tb->type || !tb->n1 || !tb->n2 || !tb->value_idx || !tb->values)))
    return 0;  // let build_plan produce the error
  uint64_t h = 1469598103934665603ull;
  const int hdr[3] = {tb->n_nodes, tb->n_elem, tb->n_values};
  h = fnv1a(h, hdr, sizeof hdr);
  h = fnv1a(h, tb->type, sizeof(int32_t) * tb->n_elem);
  h = fnv1a(h, tb->n1, sizeof(int32_t) * tb->n_elem);
  h = fnv1a(h, tb->n2, sizeof(int32_t) * tb->n_elem);
  if (tb->nc1) h = fnv1a(h, tb->nc1, sizeof(int32_t) * tb->n_elem);
  if (tb->nc2) h = fnv1a(h, tb->nc2, sizeof(int32_t) * tb->n_elem);
  h = fnv1a(h, tb->value_idx, sizeof(int32_t) * tb->n_elem);
  h = fnv1a(h, tb->values, sizeof(double) * tb->n_values);
  if (sw) {
    if (sw->n_inst < 1 || sw->n_var < 0 || (sw->n_var > 0 && (!sw->var_slot || !sw->var_values))) return 0;
    const int nv = sw->n_var;
    h = fnv1a(h, &nv, sizeof nv);
    if (nv > 0) h = fnv1a(h, sw->var_slot, sizeof(int32_t) * nv);
  } else {
    const int nv = -1;
    h = fnv1a(h, &nv, sizeof nv);
  }
  return h ? h : 1;
}


inline uint64_t plan_key(const HostPlan& hp) {
  uint64_t h = 1469598103934665603ull;
  h = fnv1a(h, &hp.nn, sizeof(int));
  h = fnv1a(h, hp.ends.data(), sizeof(int4) * hp.ends.size());
  h = fnv1a(h, hp.meta.data(), sizeof(int2) * hp.meta.size());
  h = fnv1a(h, hp.values.data(), sizeof(double) * hp.values.size());
  return h ? h : 1;
}


}  // namespace host
}  // namespace spicey
