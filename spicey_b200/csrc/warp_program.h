// Warp-cooperative form of the sparse LU program (host side).
//
// sparse_program.h runs one THREAD per system; that stops scaling when the factorisation of one system no
// longer fits a thread's share of the chip (cfg 4, the 16x16 mesh: 4,400 U entries = 70 KB per system, so
// the thread-per-system workspace lives in L2/HBM and every update moves 48 bytes through it).  Here one
// WARP owns one system: the single-assignment program is re-scheduled into levels of mutually independent
// operations that the 32 lanes execute side by side,
//
//   step k:  A  verify the pilot's pivot against every structural candidate, r = 1/pivot
//            B  one multiplier per row to eliminate           (lanes = rows)
//            C  a_ij <- a_ij - f_i * a_kj                     (lanes = (row, column) pairs)
//   back-substitution, column oriented: x_j = acc_j / u_jj, then acc_i -= u_ij * x_j for the rows i of
//   column j (lanes = rows; no reduction across lanes)
//
// and the values are placed by what reads them: everything the elimination itself reads lives in a
// per-warp pool in shared memory (linear-scan allocation in program order, LIFO free list, in-place
// updates); what the back-substitution reads (U, 1/u_kk, the eliminated right-hand side) is written once
// to a per-warp global workspace whose slots are numbered in the order the back-substitution consumes them
// (contiguous reads).  A value with readers in both phases is stored to both.
//
// Safety of the slot reuse under parallel execution: slots are released at a value's last reader in
// program order and handed out to later operations only, so a writer never precedes a reader of the old
// value in program order; the kernel executes a level in chunks of 32 consecutive operations and loads all
// operands of a chunk before it stores any result (warp barrier in between), and levels are separated by
// warp barriers.
#pragma once
#include <algorithm>
#include <climits>
#include <cstring>
#include <map>
#include <vector>

#include "sparse_program.h"

namespace spicey {

struct WarpStep {
  int cand_begin, n_cand, pidx, rcp_g;   // candidates [cand_begin, +n_cand) in scan order, pilot's choice, global slot of 1/pivot
  int elim_begin, n_elim;                // rows to eliminate (a_ik operands)
  int col_begin, n_cols;                 // columns of the pivot row right of the pivot (+ rhs): src slot per column
  int upd_begin;                         // n_elim x n_cols updates, row-major
  int stamp_begin, n_stamp;              // stamped entries this step reads, materialised into temporary pool slots
};
struct WarpStamp { int entry, slot; };
struct WarpUpd { int old_enc, dst; };   // pool slots; dst -1: no forward reader
struct WarpCol { int rcp_g, begin, count, pad; };        // column j of the back-substitution
struct WarpColEnt { int row, u_enc; };

constexpr int kWarpZero = INT_MIN;   // back-phase operand: structural zero

struct WarpProgram {
  bool ok = false;
  int n = 0;
  int n_pool = 0;     // shared-memory pool slots per system (slot 0 holds zero)
  int n_gslots = 0;   // global workspace slots per system
  int max_elim = 0;   // largest number of rows eliminated in one step (size of the multiplier array)
  long long n_upd_total = 0;
  std::vector<WarpStep> steps;
  std::vector<WarpStamp> stamp;
  std::vector<int> cand;         // forward operands: pool slots (stamped entries are materialised per step, see stamp)
  std::vector<int> elim;         // a_ik per eliminated row
  std::vector<int> src;          // pivot-row entry per column of a step
  std::vector<WarpUpd> upd;
  std::vector<int> upd_g;        // per update: global copy slot or -1
  std::vector<int> rhs_init;     // [n] back operands: >= 0 global slot, < 0 stamped entry ~idx, kWarpZero
  std::vector<WarpCol> cols;     // [n]
  std::vector<WarpColEnt> colent;
  // Packed form the kernel stages through shared memory, one record per pivot step / per group of
  // kWarpColGroup back-substitution columns (all sizes in 16-byte units; see pack_warp_program):
  std::vector<int> stream;        // records, each a multiple of 4 ints
  std::vector<int> fwd_tab;       // [n][2]  (offset, length) of step s
  std::vector<int> back_tab;      // [n_groups][2]
  int n_groups = 0;
  int max_rec16 = 0;              // largest record
  int g_first0 = 0, g_count0 = 0; // global slots the first back-substitution group reads
};

constexpr int kWarpColGroup = 16;

// step record:  {n_cand, pidx, rcp_g, n_elim, n_cols, n_stamp, 0, 0}
//               stamp[n_stamp][12] = {slot, 0, 0, 0, (alpha + Re J), Im J, beta, gamma as doubles}
//               cand[n_cand] elim[n_elim] src[n_cols] pad4
//               op[n_elim][n_cols] = (16 * old | has_global_copy) | (16 * dst) << 16, 0xfff0 = no destination
//               (byte offsets into the pool: no scaling on the device; the low 4 bits are free for the flag)   pad4
//               opg[n_elim][n_cols]  global copy slot (read only when flagged)                          pad4
// back record:  {n_cols, first global slot the NEXT group reads, number of slots it reads, 0}
//               cols[n_cols][4] = {rcp_g, ent_begin (ints from record start), count, j} ents[...][2] pad4
inline void pack_warp_program(WarpProgram& wp, const SparseProgram& sp) {
  wp.stream.clear(); wp.fwd_tab.clear(); wp.back_tab.clear();
  wp.max_rec16 = 0;
  auto pad4 = [&]() { while (wp.stream.size() & 3) wp.stream.push_back(0); };
  for (const WarpStep& st : wp.steps) {
    const size_t o = wp.stream.size();
    const int hdr[8] = {st.n_cand, st.pidx, st.rcp_g, st.n_elim, st.n_cols, st.n_stamp, 0, 0};
    wp.stream.insert(wp.stream.end(), hdr, hdr + 8);
    for (int q = 0; q < st.n_stamp; ++q) {
      const WarpStamp& m = wp.stamp[st.stamp_begin + q];
      double c[4] = {0.0, 0.0, 0.0, 0.0};
      if ((size_t)m.entry < sp.ent_alpha.size()) {
        c[0] = sp.ent_alpha[m.entry] + sp.ent_jre[m.entry]; c[1] = sp.ent_jim[m.entry];
        c[2] = sp.ent_beta[m.entry]; c[3] = sp.ent_gamma[m.entry];
      }
      int w[12] = {m.slot, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      memcpy(w + 4, c, sizeof c);
      wp.stream.insert(wp.stream.end(), w, w + 12);
    }
    for (int c = 0; c < st.n_cand; ++c) wp.stream.push_back(wp.cand[st.cand_begin + c]);
    for (int e = 0; e < st.n_elim; ++e) wp.stream.push_back(wp.elim[st.elim_begin + e]);
    for (int c = 0; c < st.n_cols; ++c) wp.stream.push_back(wp.src[st.col_begin + c]);
    pad4();
    const int n_upd = st.n_elim * st.n_cols;
    for (int u = 0; u < n_upd; ++u) {
      const WarpUpd& w = wp.upd[st.upd_begin + u];
      const int g = wp.upd_g[st.upd_begin + u];
      wp.stream.push_back((int)(((unsigned)w.old_enc * 16u) | (g >= 0 ? 1u : 0u) | ((w.dst < 0 ? 0xfff0u : (unsigned)w.dst * 16u) << 16)));
    }
    pad4();
    for (int u = 0; u < n_upd; ++u) wp.stream.push_back(wp.upd_g[st.upd_begin + u]);
    pad4();
    const int len16 = (int)((wp.stream.size() - o) / 4);
    wp.fwd_tab.push_back((int)(o / 4)); wp.fwd_tab.push_back(len16);
    wp.max_rec16 = std::max(wp.max_rec16, len16);
  }
  wp.n_groups = 0;
  // global slots a group reads: a contiguous range (slots are numbered in consumption order)
  auto g_range = [&](int hi, int& first, int& count) {
    const int lo = std::max(0, hi - kWarpColGroup + 1);
    int mn = INT_MAX, mx = -1;
    for (int j = hi; j >= lo && j >= 0; --j) {
      const WarpCol& c = wp.cols[j];
      if (c.rcp_g >= 0) { mn = std::min(mn, c.rcp_g); mx = std::max(mx, c.rcp_g); }
      for (int q = 0; q < c.count; ++q) {
        const int e = wp.colent[c.begin + q].u_enc;
        if (e >= 0) { mn = std::min(mn, e); mx = std::max(mx, e); }
      }
    }
    first = mx >= 0 ? mn : 0;
    count = mx >= 0 ? mx - mn + 1 : 0;
  };
  g_range(wp.n - 1, wp.g_first0, wp.g_count0);
  for (int hi = wp.n - 1; hi >= 0; hi -= kWarpColGroup) {
    const int lo = std::max(0, hi - kWarpColGroup + 1), nc = hi - lo + 1;
    const size_t o = wp.stream.size();
    int nf = 0, ncnt = 0;
    if (lo > 0) g_range(lo - 1, nf, ncnt);
    const int hdr[4] = {nc, nf, ncnt, 0};
    wp.stream.insert(wp.stream.end(), hdr, hdr + 4);
    int ent = 4 + 4 * nc;
    for (int j = hi; j >= lo; --j) {
      const WarpCol& c = wp.cols[j];
      wp.stream.push_back(c.rcp_g); wp.stream.push_back(ent); wp.stream.push_back(c.count); wp.stream.push_back(j);
      ent += 2 * c.count;
    }
    for (int j = hi; j >= lo; --j) {
      const WarpCol& c = wp.cols[j];
      for (int q = 0; q < c.count; ++q) { wp.stream.push_back(wp.colent[c.begin + q].row); wp.stream.push_back(wp.colent[c.begin + q].u_enc); }
    }
    pad4();
    const int len16 = (int)((wp.stream.size() - o) / 4);
    wp.back_tab.push_back((int)(o / 4)); wp.back_tab.push_back(len16);
    wp.max_rec16 = std::max(wp.max_rec16, len16);
    ++wp.n_groups;
  }
}

// Lowers sp.ir.  pool_cap: largest pool the kernel can give one warp; the program is rejected (ok=false)
// when the elimination's working set does not fit, instead of spilling forward operands to global memory.
inline void build_warp_program(const SparseProgram& sp, int pool_cap, WarpProgram& wp, int pool_spare = 1 << 30) {
  using namespace sparse_detail;
  wp = WarpProgram();
  const int n = sp.n, nv = sp.n_virtual;
  wp.n = n;
  const std::vector<IrOp>& ir = sp.ir;
  const int n_ir = (int)ir.size();
  int B = n_ir;
  for (int t = 0; t < n_ir; ++t) if (ir[t].kind == SOP_BSUB) { B = t; break; }
  // ---- steps: a PIVOT and the ELIM rows that follow it; every row updates the same columns (the pivot row's
  //      structural non-zeros right of the pivot, plus the right-hand side) ----
  struct StepIr { const IrOp* piv; std::vector<const IrOp*> rows; };
  std::vector<StepIr> sir;
  for (int t = 0; t < B; ++t) {
    if (ir[t].kind == SOP_PIVOT) { sir.push_back(StepIr()); sir.back().piv = &ir[t]; }
    else if (!sir.empty()) sir.back().rows.push_back(&ir[t]);
    else return;
  }
  if ((int)sir.size() != n) return;
  for (const StepIr& si : sir)
    for (const IrOp* r : si.rows) {
      if (r->upd.size() != si.rows[0]->upd.size()) return;
      for (size_t c = 0; c < r->upd.size(); ++c)
        if (r->upd[c].col != si.rows[0]->upd[c].col || r->upd[c].src != si.rows[0]->upd[c].src) return;
    }
  // Execution order of the kernel (what the slot reuse must be safe for): per step the candidates, the rows'
  // multipliers, then the updates in passes of 32 columns, each pass row by row.
  // ---- uses: last forward reader in that order, and whether the back-substitution reads the value ----
  std::vector<int> fwd_last(nv, -1);
  std::vector<char> back_use(nv, 0);
  {
    int tt = 0;
    for (const StepIr& si : sir) {
      for (int o : si.piv->reads) if (o >= 0) fwd_last[o] = tt;
      ++tt;
      for (const IrOp* r : si.rows) { if (r->reads[0] >= 0) fwd_last[r->reads[0]] = tt; ++tt; }
      const int nc = si.rows.empty() ? 0 : (int)si.rows[0]->upd.size();
      for (int c0 = 0; c0 < nc; c0 += 32) {
        for (int c = c0; c < std::min(nc, c0 + 32); ++c) {   // the pass loads its pivot-row entries once, up front
          const int o = si.rows[0]->upd[c].src;
          if (o >= 0) fwd_last[o] = tt;
        }
        ++tt;
        for (const IrOp* r : si.rows)
          for (int c = c0; c < std::min(nc, c0 + 32); ++c) {
            if (r->upd[c].dst_old >= 0) fwd_last[r->upd[c].dst_old] = tt;
            ++tt;
          }
      }
    }
    for (int t = B; t < n_ir; ++t)
      for (int o : ir[t].reads) if (o >= 0) back_use[o] = 1;
  }
  // ---- global slots, numbered in back-substitution order: rhs of every row, then per column (descending)
  //      1/u_jj and the column's U entries ----
  std::vector<int> var_of_x(nv, -1);      // virtual id of x_j -> j
  std::vector<const IrOp*> bs_of(n, nullptr);
  for (int t = B; t < n_ir; ++t) { var_of_x[ir[t].def] = ir[t].var; bs_of[ir[t].var] = &ir[t]; }
  for (int i = 0; i < n; ++i) if (!bs_of[i]) return;
  std::vector<int> gslot(nv, -1);
  int ng = 0;
  auto g_of = [&](int o) -> int {
    if (o == kNoOperand) return kWarpZero;
    if (o < 0) return o;  // stamped entry ~idx, recomputed where it is read
    if (gslot[o] < 0) gslot[o] = ng++;
    return gslot[o];
  };
  wp.rhs_init.resize(n);
  for (int i = 0; i < n; ++i) wp.rhs_init[i] = g_of(bs_of[i]->reads[0]);
  std::vector<std::vector<WarpColEnt>> col(n);
  for (int i = 0; i < n; ++i) {
    const IrOp& bs = *bs_of[i];
    for (size_t q = 2; q + 1 < bs.reads.size(); q += 2) {
      const int j = var_of_x[bs.reads[q + 1]];
      if (j < 0) return;
      col[j].push_back(WarpColEnt{i, bs.reads[q]});   // operand resolved below, in consumption order
    }
  }
  wp.cols.resize(n);
  for (int j = n - 1; j >= 0; --j) {
    WarpCol c;
    c.rcp_g = g_of(bs_of[j]->reads[1]);
    c.begin = (int)wp.colent.size();
    c.count = (int)col[j].size();
    c.pad = 0;
    for (WarpColEnt e : col[j]) { e.u_enc = g_of(e.u_enc); wp.colent.push_back(e); }
    wp.cols[j] = c;
  }
  wp.n_gslots = std::max(1, ng);
  // ---- forward phase: pool allocation in program order ----
  // Slots are handed out by bank group: a 16-byte pool access of a warp is served a quarter-warp at a time,
  // conflict-free when the eight slots differ mod 8.  The lanes of a chunk work on consecutive columns of a
  // row, so a value of column j gets a slot with slot % 8 == j % 8 (in-place updates keep it).
  std::vector<int> pool(nv, -1), free_list[8];
  int high = 1;  // slot 0 = zero
  const int kSpare = pool_spare;
  auto alloc = [&](int col) -> int {
    const int b = col & 7;
    if (!free_list[b].empty()) { int s = free_list[b].back(); free_list[b].pop_back(); return s; }
    // no free slot of the right residue: growing the pool costs occupancy, a wrong residue costs a bank
    // conflict — take a foreign slot when plenty are lying around
    int total = 0, best = -1;
    for (int q = 0; q < 8; ++q) { total += (int)free_list[q].size(); if (best < 0 || free_list[q].size() > free_list[best].size()) best = q; }
    if (total >= kSpare && !free_list[best].empty()) { int s = free_list[best].back(); free_list[best].pop_back(); return s; }
    while ((high & 7) != b) { free_list[high & 7].push_back(high); ++high; }
    return high++;
  };
  // A stamped entry read by a step is written to a temporary pool slot at the start of that step (so that
  // the update loop reads nothing but pool slots) and the slot is released when the step ends.
  // Temporary slots form their own free list: they are written at the START of a step, so they must not be
  // slots that the same step releases (and still reads) further down.
  std::map<int, int> stamped;
  std::vector<int> temp_free;
  auto alloc_temp = [&]() -> int {
    if (!temp_free.empty()) { int t = temp_free.back(); temp_free.pop_back(); return t; }
    return high++;
  };
  (void)0;
  auto enc = [&](int o) -> int {
    if (o == kNoOperand) return 0;
    if (o < 0) {
      const int en = ~o;
      auto it = stamped.find(en);
      if (it == stamped.end()) {
        it = stamped.insert(std::make_pair(en, alloc_temp())).first;
        wp.stamp.push_back(WarpStamp{en, it->second});
      }
      return it->second;
    }
    return pool[o];   // defined earlier with a forward reader, hence in the pool
  };
  int tt = 0;
  auto release = [&](int o) {
    if (o >= 0 && fwd_last[o] == tt && pool[o] >= 0) { free_list[pool[o] & 7].push_back(pool[o]); fwd_last[o] = -2; }
  };
  for (const StepIr& si : sir) {
    WarpStep cur = WarpStep();
    cur.stamp_begin = (int)wp.stamp.size();
    cur.cand_begin = (int)wp.cand.size();
    cur.n_cand = (int)si.piv->reads.size();
    cur.pidx = si.piv->pidx;
    for (int o : si.piv->reads) { const int e = enc(o); if (e < 0) return; wp.cand.push_back(e); }
    for (int o : si.piv->reads) release(o);
    cur.rcp_g = (si.piv->def >= 0 && gslot[si.piv->def] >= 0) ? gslot[si.piv->def] : -1;
    ++tt;
    cur.elim_begin = (int)wp.elim.size();
    cur.n_elim = (int)si.rows.size();
    for (const IrOp* r : si.rows) {
      const int e = enc(r->reads[0]);
      if (e < 0) return;
      wp.elim.push_back(e);
      release(r->reads[0]);
      ++tt;
    }
    const int nc = si.rows.empty() ? 0 : (int)si.rows[0]->upd.size();
    cur.col_begin = (int)wp.src.size();
    cur.n_cols = nc;
    cur.upd_begin = (int)wp.upd.size();
    wp.upd.resize(wp.upd.size() + (size_t)cur.n_elim * nc);
    wp.upd_g.resize(wp.upd.size(), -1);
    wp.src.resize(wp.src.size() + nc);
    for (int c0 = 0; c0 < nc; c0 += 32) {
      for (int c = c0; c < std::min(nc, c0 + 32); ++c) {
        const int e = enc(si.rows[0]->upd[c].src);
        if (e < 0) return;
        wp.src[cur.col_begin + c] = e;
      }
      for (int c = c0; c < std::min(nc, c0 + 32); ++c) release(si.rows[0]->upd[c].src);
      ++tt;
      for (int e = 0; e < cur.n_elim; ++e)
        for (int c = c0; c < std::min(nc, c0 + 32); ++c) {
          const Update& u = si.rows[e]->upd[c];
          WarpUpd w;
          w.old_enc = enc(u.dst_old);
          if (w.old_enc < 0) return;
          release(u.dst_old);   // in-place update when the old version dies here
          w.dst = -1;
          if (fwd_last[u.dst_new] >= 0) { pool[u.dst_new] = alloc(u.col); w.dst = pool[u.dst_new]; }
          wp.upd[cur.upd_begin + (size_t)e * nc + c] = w;
          wp.upd_g[cur.upd_begin + (size_t)e * nc + c] = gslot[u.dst_new];
          ++tt;
        }
    }
    cur.n_stamp = (int)wp.stamp.size() - cur.stamp_begin;
    for (const auto& kv : stamped) temp_free.push_back(kv.second);
    stamped.clear();
    wp.max_elim = std::max(wp.max_elim, cur.n_elim);
    wp.steps.push_back(cur);
  }
  wp.n_pool = high;
  wp.n_upd_total = (long long)wp.upd.size();
  if (high > pool_cap || high >= 0xfff) return;   // 16-bit byte offsets, 0xfff0 reserved
  pack_warp_program(wp, sp);
  wp.ok = true;
}

}  // namespace spicey
