// Sparse AC kernel: one THREAD per frequency point executes the straight-line LU program
// built by sparse_program.h (static pivot sequence from a pilot point, verified per point).
//
// Replaces, for a pure frequency sweep (one instance, no swept values), the same reference
// code as ac_kernels.cuh: simulateAC.ts:24-60 (stamp), solveComplex.ts:4-73 (solve),
// simulateAC.ts:85-126 (unpack) — with the dense kernel as the exact fallback for any point
// whose pivot choice (solveComplex.ts:18-28 rule) differs from the pilot's.
//
// B200 mapping
//  * Lanes of a warp are 32 consecutive frequency points.  Each thread's working set
//    (matrix slots incl. fill, then x) lives in a [slot][thread] workspace, so every
//    slot access of a warp is one contiguous 512-byte request; the workspace is sized for
//    the resident grid (SMs x 1024 threads), stays L1/L2-resident in flight and is reused
//    across the grid-stride loop — the matrix itself never round-trips to HBM per point
//    beyond that workspace.
//  * Stamping needs no element loop and no memory: with no swept values every matrix entry is
//    alpha + j*(w*beta - gamma/w) with per-entry constants summed on the host in the
//    reference's stamping order (R: 1/R, C: C, L: 1/L).  An operand that still holds its stamped
//    value is recomputed from those constants where it is used (lazy stamping); only values
//    the elimination produced ever touch the workspace, and the host's liveness analysis lets
//    short-lived ones share slots.
//  * The program words are uniform across the warp (same address for all lanes: one
//    broadcast load through the read-only path); control flow is divergence-free except
//    for the rare early exit of a failed / diverged point.
#pragma once
#include "common.cuh"
#include "sparse_program.h"

namespace spicey {

struct SparseArgs {
  const int* code;
  const int* x_slot;                               // [n] slot of x_i after the program
  int n, n_stamp, n_slots;
  const double2 *ent_c0, *ent_c1;                  // [n_stamp] (alpha + Re J, Im J), (beta, gamma)
  const double *el_a, *el_b, *el_g;                // [n_ac_elem]
  const double* ind_L;
  int n_ind;
  const int4* ends;
  int n_ac_elem, nn, v_first;
  const double* freqs;
  long long p_count;
  double2* W;       // [n_slots][T]
  long long T;      // threads of the grid
  double2* x;       // [p_count][n]
  double2* ielem;   // [p_count][n_ac_elem] or null
  int* status;      // [p_count]
  long long* fb_list;  // points handed to the dense kernel
  int* fb_count;
};

__global__ void __launch_bounds__(128) ac_sparse_kernel(SparseArgs a) {
  typedef Num<cplx> N;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long T = a.T;
  double2* __restrict__ W = a.W + tid;
  const int* __restrict__ code = a.code;
  const double thr = kEps * kEps;  // squared-magnitude metric

  for (long long p = tid; p < a.p_count; p += T) {
    const double f = a.freqs[p];
    const double w = (2 * kPi) * f;
    const double iw = 1.0 / w;
    // Operand fetch: >= 0 workspace slot, < 0 pristine stamped entry ~idx (lazy stamping), INT_MIN zero.
    auto fetch = [&](int o) -> cplx {
      if (o >= 0) return W[(long long)o * T];
      if (o == kNoOperand) return make_double2(0.0, 0.0);
      const double2 c0 = __ldg(a.ent_c0 + ~o), c1 = __ldg(a.ent_c1 + ~o);
      return make_double2(c0.x, fma(w, c1.x, -c1.y * iw) + c0.y);
    };
    bool diverged = false;
    int status = ST_OK;
    // inductor guards of simulateAC.ts:47-51 / Complex.ts:41 are value dependent: leave them to the dense kernel
    for (int l = 0; l < a.n_ind; ++l) {
      double d = w * a.ind_L[l];
      if (fabs(d) < kEps || d * d < kEps) diverged = true;
    }
    int pc = 0;
    cplx r = N::zero();
    while (!diverged && status == ST_OK) {
      const int op = __ldg(code + pc);
      if (op == SOP_PIVOT) {
        const int nc = __ldg(code + pc + 1), pidx = __ldg(code + pc + 2);
        const cplx ap = fetch(__ldg(code + pc + 3 + pidx));
        const double mp = N::metric<false>(ap);
        bool ok = (mp == mp);
        for (int c = 0; c < nc; ++c) {
          if (c == pidx) continue;
          const double m = N::metric<false>(fetch(__ldg(code + pc + 3 + c)));
          ok = ok && (c < pidx ? (m < mp) : !(m > mp));   // first maximum wins (solveComplex.ts:20-28)
        }
        if (!ok) { diverged = true; break; }
        if (mp < thr) { status = ST_SINGULAR; break; }      // :29
        if (mp < kEps) { status = ST_CDIV; break; }         // Complex.ts:41-42
        r = N::recip(ap);
        W[(long long)__ldg(code + pc + 3 + nc) * T] = r;    // 1/u_kk for the back-substitution
        pc += 4 + nc;
      } else if (op == SOP_ELIM) {
        cplx fm = N::mul(fetch(__ldg(code + pc + 1)), r);
        if (N::metric<false>(fm) < thr) fm = N::zero();     // :46 (a zero multiplier leaves the row unchanged)
        const int nu = __ldg(code + pc + 2);
        pc += 3;
        for (int u = 0; u < nu; ++u, pc += 3) {
          const cplx dst = fetch(__ldg(code + pc));
          const cplx src = fetch(__ldg(code + pc + 1));
          W[(long long)__ldg(code + pc + 2) * T] = N::submul<false>(dst, fm, src);   // :47-52
        }
      } else if (op == SOP_BSUB) {
        const int nt = __ldg(code + pc + 4);
        cplx acc = fetch(__ldg(code + pc + 2));
        const cplx rc = fetch(__ldg(code + pc + 3));
        pc += 5;
        for (int q = 0; q < nt; ++q, pc += 2)
          acc = N::submul<false>(acc, fetch(__ldg(code + pc)), fetch(__ldg(code + pc + 1)));
        W[(long long)__ldg(code + pc) * T] = N::mul(acc, rc);   // :56-71
        pc += 1;
      } else {
        break;
      }
    }
    // ---- unpack (simulateAC.ts:85-126) ----
    if (diverged) {
      a.status[p] = -1;
      int slot = atomicAdd(a.fb_count, 1);
      a.fb_list[slot] = p;
      continue;
    }
    cplx* xo = a.x + p * a.n;
    if (status != ST_OK) {
      for (int i = 0; i < a.n; ++i) xo[i] = N::nan();
      if (a.ielem)
        for (int e = 0; e < a.n_ac_elem; ++e) a.ielem[p * a.n_ac_elem + e] = N::nan();
      a.status[p] = status;
      continue;
    }
    for (int i = 0; i < a.n; ++i) xo[i] = W[(long long)__ldg(a.x_slot + i) * T];
    if (a.ielem) {
      cplx* io = a.ielem + p * a.n_ac_elem;
      for (int e = 0; e < a.n_ac_elem; ++e) {
        const int4 en = __ldg(a.ends + e);
        cplx cur;
        if (e >= a.v_first) {
          cur = W[(long long)__ldg(a.x_slot + a.nn + e - a.v_first) * T];
        } else {
          cplx v1 = en.x == 0 ? N::zero() : W[(long long)__ldg(a.x_slot + en.x - 1) * T];
          cplx v2 = en.y == 0 ? N::zero() : W[(long long)__ldg(a.x_slot + en.y - 1) * T];
          cplx y = make_double2(__ldg(a.el_a + e), fma(w, __ldg(a.el_b + e), -__ldg(a.el_g + e) * iw));
          cur = N::mul(y, csub(v1, v2));
        }
        io[e] = cur;
      }
    }
    a.status[p] = ST_OK;
  }
}

}  // namespace spicey
