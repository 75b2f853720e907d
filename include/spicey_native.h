/*
 * spicey_native.h — C ABI of the B200-native batched MNA solve engine.
 *
 * This is the drop-in boundary for the reference's hot path (SURVEY.md §8 b).  The
 * reference (tscircuit/spicey, TypeScript) has no FFI of its own; the seam is its
 * exported functions, so each entry point below names the reference code it replaces:
 *
 *   spicey_ac_solve    replaces the body of the per-frequency loop of simulateAC
 *                      (lib/analysis/simulateAC.ts:80-127): buildLinearSystemForAC
 *                      (:24-60, lib/stamping/stamp*Complex.ts), solveComplex
 *                      (lib/math/solveComplex.ts:4-73, lib/math/Complex.ts) and the
 *                      node-voltage / element-current unpack (:85-126).
 *   spicey_tran_solve  replaces the time-step loop of simulateTRAN
 *                      (lib/analysis/simulateTRAN.ts:146-238): stampAllElementsAtTime
 *                      (:25-102, lib/stamping/stamp*Real.ts), solveReal
 *                      (lib/math/solveReal.ts:3-73), updateSwitchStatesFromSolution
 *                      (:108-128), the recording (:164-219) and state update (:221-237).
 *
 * The host side (netlist parsing, frequency list, step count, waveform sampling, result
 * objects) stays in the caller's language; lib/native binds this header with bun:ffi
 * (INTEGRATION.md), spicey_b200/native.py binds it with ctypes.
 *
 * Conventions
 *  - Plain pointers and sizes only.  The caller owns every buffer; the library never
 *    keeps a host pointer after a call returns and never returns memory to be freed
 *    (except spicey_host_alloc / spicey_host_free, an optional pinned-buffer helper).
 *  - All floating point is IEEE binary64.  Complex numbers are interleaved (re, im).
 *  - Call-level failures return a non-zero code; text via spicey_last_error().
 *    Per-instance numerical failures (the reference's synchronous throws) are reported
 *    in status[] with SPICEY_ST_*; the wrapper re-throws the reference's message for
 *    the first failing index.  One bad instance never poisons a batch.
 *  - A handle is safe for one caller at a time.  There is NO CPU fallback: every
 *    solve entry point fails with SPICEY_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef SPICEY_NATIVE_H
#define SPICEY_NATIVE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPICEY_NATIVE_ABI_VERSION 4

/* Element kinds of the flat element table (ParsedCircuit, lib/parsing/parseNetlist.ts:12-105). */
enum {
  SPICEY_ELEM_R = 0, /* values: R                         (ParsedResistor  :12)   */
  SPICEY_ELEM_C = 1, /* values: C            state: vPrev (ParsedCapacitor :13-19) */
  SPICEY_ELEM_L = 2, /* values: L            state: iPrev (ParsedInductor  :20-26) */
  SPICEY_ELEM_V = 3, /* values: dc, acMag, acPhaseDeg     (ParsedVoltageSource :34-43) */
  SPICEY_ELEM_S = 4, /* values: Ron, Roff, Von, Voff  state: isOn (ParsedSwitch :62-71) */
  SPICEY_ELEM_D = 5, /* values: Is, N        state: vdPrev (ParsedDiode :53-60)   */
  SPICEY_ELEM_I = 6  /* values: dc, acMag, acPhaseDeg: independent current source from n1 (n+) to n2 (n-) through the
                        source, b[n+] -= I, b[n-] += I (lib/stamping/stampCurrentReal.ts:3-14, stampCurrentComplex.ts:4-15;
                        the reference's parser skips `I` lines, parseNetlist.ts:444-446, its stamps exist).  Constant in
                        transient analysis; appears in the transient element currents (its own value), not in the AC ones */
};

/* Per-instance status, mapped by the wrapper to the reference's error messages. */
enum {
  SPICEY_ST_OK = 0,
  SPICEY_ST_SINGULAR = 1,   /* "Singular matrix (complex|real)"  solveComplex.ts:29 / solveReal.ts:28 */
  SPICEY_ST_CDIV = 2,       /* "Complex divide by ~0"            Complex.ts:42 */
  SPICEY_ST_R_NONPOS = 3    /* "R <name> must be > 0"            simulateAC.ts:37 (AC only) */
};

/* Call-level return codes. */
enum {
  SPICEY_SUCCESS = 0,
  SPICEY_ERR_INVALID = 1,    /* bad argument / malformed table */
  SPICEY_ERR_NO_DEVICE = 2,  /* no usable CUDA device: there is no CPU path */
  SPICEY_ERR_CUDA = 3,       /* CUDA runtime error (text in spicey_last_error) */
  SPICEY_ERR_UNSUPPORTED = 4 /* system larger than the largest kernel tier */
};

/*
 * Flat element table (struct of arrays).  Elements are grouped by kind in the order
 * R, C, L, V, S, D, I and keep netlist order inside a kind — the order in which the
 * reference pushes element currents (simulateAC.ts:94-126, simulateTRAN.ts:173-219).
 * Node ids are the reference's: 0 is ground, the matrix row of node id is id-1
 * (NodeIndex.ts:28-31); the k-th V element owns branch row n_nodes + k
 * (parseNetlist.ts:455-460).
 */
typedef struct spicey_elem_table {
  int32_t n_nodes;          /* non-ground nodes (nn) */
  int32_t n_elem;
  int32_t n_values;         /* length of values[] */
  int32_t reserved;
  const int32_t* type;      /* [n_elem] SPICEY_ELEM_* (grouped, see above) */
  const int32_t* n1;        /* [n_elem] first node  (n1 / nPlus) */
  const int32_t* n2;        /* [n_elem] second node (n2 / nMinus) */
  const int32_t* nc1;       /* [n_elem] switch control + (ncPos), else 0 */
  const int32_t* nc2;       /* [n_elem] switch control - (ncNeg), else 0 */
  const int32_t* value_idx; /* [n_elem] first slot of the element in values[] */
  const double* values;     /* [n_values] nominal values, slots per kind as listed above */
} spicey_elem_table;

/*
 * Sweep / Monte-Carlo batch: n_inst instances of the same topology whose value slots
 * var_slot[v] take per-instance values var_values[v*n_inst + inst].  NULL or n_inst==1
 * with n_var==0 means one nominal instance.
 */
typedef struct spicey_sweep {
  int64_t n_inst;
  int32_t n_var;
  int32_t reserved;
  const int32_t* var_slot;   /* [n_var] indices into values[] */
  const double* var_values;  /* [n_var][n_inst] */
} spicey_sweep;

/* Counters of the most recent solve call on a handle (all devices of the handle). */
typedef struct spicey_stats {
  double kernel_ms;          /* CUDA-event time of the solve kernels, max over devices */
  double total_ms;           /* host wall time of the call */
  int64_t kernel_launches;   /* kernels of this library launched by the call */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int64_t solves;            /* AC points, or TRAN matrix solves (sum of re-solve iterations) */
  int32_t tier;              /* kernel tier used (SPICEY_TIER_*) */
  int32_t n_devices;
  int64_t fallback_solves;   /* sparse tier: systems re-solved by the dense kernel (pivot differed) */
  int64_t program_cfma;      /* sparse tier: complex FMAs executed per system */
} spicey_stats;

enum {
  SPICEY_TIER_THREAD = 1,    /* one thread per system (Nvar <= 16) */
  SPICEY_TIER_CTA_SMEM = 2,  /* one CTA per system, matrix resident in shared memory */
  SPICEY_TIER_CTA_GMEM = 3,  /* one CTA per system, matrix in an L2-resident global scratch */
  SPICEY_TIER_SPARSE = 4,    /* one thread per system, static-pivot sparse LU program verified per
                                system (interpreted), dense pivoting kernel as fallback */
  SPICEY_TIER_SPARSE_JIT = 5, /* the same program written out as a straight-line sm_100a kernel and compiled
                                with NVRTC once per topology (large single-instance sweeps, small programs) */
  SPICEY_TIER_TRAN_JIT = 6   /* transient: the persistent time loop written out for the netlist at hand (every
                                element a few named registers) and compiled with NVRTC once per topology;
                                Nvar <= 8, batches of >= 2e6 instance-steps or SPICEY_FLAG_JIT */
  ,SPICEY_TIER_SPARSE_WARP = 7 /* the sparse program level-scheduled for one WARP per system: elimination working
                                set in shared memory, U streamed to a per-warp workspace (programs whose
                                per-system factorisation is too large for one thread's share of the chip) */
  ,SPICEY_TIER_BAND = 8       /* banded + bordered systems (meshes, long ladders) after a bandwidth-reducing
                                renumbering of the nodes: a few lanes per system, the sliding window of rows in
                                registers, partial pivoting verified per system (band_kernel.cuh) */
  ,SPICEY_TIER_TILE = 9       /* dense LU with partial pivoting (lib/math/solveComplex.ts:15-53), the augmented
                                matrix of a system resident in REGISTERS as 2-D cyclic tiles of one CTA (one warp
                                for small Nvar), compiled per (Nvar, tile shape): dense circuits and
                                SPICEY_FLAG_DENSE batches of >= 4096 points (tile_kernel.cuh); plain sweeps of
                                Nvar <= 32: one WARP per system, a lane's row in registers (warp_lu_kernel.cuh) */
};

typedef struct spicey_handle spicey_handle;

int32_t spicey_native_abi_version(void);
/* Number of usable CUDA devices (0 when none). */
int32_t spicey_device_count(void);
/* Text of the last call-level error on this thread ("" when none). */
const char* spicey_last_error(void);

/* devices == NULL or n_devices <= 0 selects device 0. Work is sharded across the
 * handle's devices in contiguous index ranges (SURVEY.md §8 e); no collective.
 * All devices of a handle are driven from the calling thread: pass page-locked output
 * buffers (spicey_host_alloc, or memory the caller registered) to a multi-device handle —
 * a device->host copy into pageable memory blocks the thread, and the next device then
 * starts late.  Every entry point leaves the caller's current CUDA device unchanged. */
int32_t spicey_create(const int32_t* devices, int32_t n_devices, spicey_handle** out);
void spicey_destroy(spicey_handle* h);
int32_t spicey_get_stats(const spicey_handle* h, spicey_stats* out);

/* Optional page-locked host buffers for the caller's inputs/outputs (faster copies). */
void* spicey_host_alloc(int64_t bytes);
void spicey_host_free(void* p);

/*
 * AC small-signal batch.  Point p = inst*n_freq + k solves the complex MNA system of
 * instance inst at freqs[k].  Nvar = n_nodes + nV.  S and D elements are ignored, as in
 * the reference (simulateAC.ts:36-57).
 *   x      [n_inst*n_freq][Nvar][2]    solution: node voltages then V branch currents
 *   ielem  [n_inst*n_freq][nAc][2]     currents of the R, C, L, V elements in table order
 *                                      (nAc = their count); may be NULL
 *          with SPICEY_FLAG_SERIES_MAJOR the two are transposed: x[Nvar][P][2], ielem[nAc][P][2]
 *   status [n_inst*n_freq]             SPICEY_ST_*; rows of a failed point are NaN
 * Host pointers.  flags: SPICEY_FLAG_*.
 */
int32_t spicey_ac_solve(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep,
                        const double* freqs, int64_t n_freq, double* x, double* ielem,
                        int32_t* status, uint32_t flags);

/* Same, on ONE device with device pointers (freqs, sweep->var_values, x, ielem, status
 * live on device devices[dev_index]); the table arrays stay on the host.  Launches on
 * `stream` (a cudaStream_t, NULL = default stream) and does not synchronise.
 * series_ld: 0 = point-major x[P][Nvar] (or, with SPICEY_FLAG_SERIES_MAJOR, series-major with
 * leading dimension P); otherwise series-major x[Nvar][series_ld], ielem[nAc][series_ld] with
 * series_ld >= P points per row.  Use spicey_series_ld(P): rows that start on a 512-byte boundary
 * make every warp's store whole 128-byte lines (measured on B200: 6.0 instead of 3.4 TB/s of
 * store bandwidth for this pattern). */
int32_t spicey_ac_solve_device(spicey_handle* h, int32_t dev_index, const spicey_elem_table* table,
                               const spicey_sweep* sweep, const double* d_freqs, int64_t n_freq,
                               double* d_x, double* d_ielem, int32_t* d_status, int64_t series_ld,
                               uint32_t flags, void* stream);

/* Recommended leading dimension (in points) of a series-major device result of n_points points:
 * n_points rounded up to a multiple of 32 (32 points x 16 B = 512 B). */
int64_t spicey_series_ld(int64_t n_points);

/*
 * Transient batch: fixed-step backward Euler from state0, steps+1 recorded samples
 * (t_k = k*dt, k = 0..steps; dt and steps come from computeEffectiveTimeStep on the
 * host, simulateTRAN.ts:14-19 — hazard H2).  The re-solve policy is the reference's:
 * another iteration only while a switch toggled, at most 20 (simulateTRAN.ts:151-162).
 *   vsrc      [nV][steps+1] pre-sampled waveform(t_k) per V element (NULL if none)
 *   vsrc_mask [nV] non-zero: use vsrc row; zero: use the element's dc value
 *   state0    [nState][n_inst] initial vPrev/iPrev/isOn/vdPrev of the C, L, S, D
 *             elements in table order; NULL = all zero / off
 *   v         [steps+1][n_nodes][n_inst]   node voltages
 *   ielem     [steps+1][n_elem][n_inst]    element currents, table order; may be NULL
 *   state_out [nState][n_inst]             final state (the reference mutates ckt); may be NULL
 *   iters     [steps+1][n_inst]            solves done per step; may be NULL
 *   status    [n_inst]
 */
int32_t spicey_tran_solve(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep,
                          double dt, int64_t steps, const double* vsrc, const int32_t* vsrc_mask,
                          const double* state0, double* v, double* ielem, double* state_out,
                          int32_t* iters, int32_t* status, uint32_t flags);

int32_t spicey_tran_solve_device(spicey_handle* h, int32_t dev_index, const spicey_elem_table* table,
                                 const spicey_sweep* sweep, double dt, int64_t steps,
                                 const double* d_vsrc, const int32_t* vsrc_mask /* host */,
                                 const double* d_state0, double* d_v, double* d_ielem,
                                 double* d_state_out, int32_t* d_iters, int32_t* d_status,
                                 uint32_t flags, void* stream);

/*
 * Source waveforms evaluated ON THE DEVICE: the reference hides a V element's PULSE / PWL
 * in a closure (parseNetlist.ts:366-383: spec.waveform = t => pulseValue(p, t)), evaluated once per step at
 * t = step*dt (simulateTRAN.ts:66-69, :147).  Here the parameters are value slots of the element table like
 * any R or C, so a spicey_sweep can vary them per instance (amplitude, delay, duty ... sweeps), which a
 * pre-sampled row shared by all instances cannot express.  Evaluation is pulseValue.ts:4-22 / pwlValue.ts:3-16
 * with every operation separately rounded: bit-identical to the pre-sampled row of the same parameters.
 *   kind[k]       SPICEY_WAVE_* of the k-th V element
 *   value_idx[k]  first slot in table->values of the parameters:
 *                   PULSE  v1, v2, td, tr, tf, ton, period, ncycles   (8 slots, PulseSpec, lib/types/simulation.ts:1-10;
 *                          ncycles = +Infinity when the netlist gives 7 arguments, parsePulseArgs.ts)
 *                   PWL    t0, v0, t1, v1, ...                        (2*n_pairs[k] slots)
 *   n_pairs[k]    PWL pair count (ignored for the other kinds)
 */
enum { SPICEY_WAVE_DC = 0, SPICEY_WAVE_TABLE = 1, SPICEY_WAVE_PULSE = 2, SPICEY_WAVE_PWL = 3 };
typedef struct spicey_waves {
  int32_t n_vsrc;            /* must equal the number of V elements of the table */
  int32_t reserved;
  const int32_t* kind;       /* [n_vsrc] */
  const int32_t* value_idx;  /* [n_vsrc] */
  const int32_t* n_pairs;    /* [n_vsrc] (may be NULL when no source is PWL) */
} spicey_waves;

/* spicey_tran_solve / spicey_tran_solve_device with per-source waveform descriptors instead of the
 * has-row mask; vsrc holds the rows of the sources of kind SPICEY_WAVE_TABLE (NULL if none). */
int32_t spicey_tran_solve_waves(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep,
                                double dt, int64_t steps, const spicey_waves* waves, const double* vsrc,
                                const double* state0, double* v, double* ielem, double* state_out,
                                int32_t* iters, int32_t* status, uint32_t flags);

/* spicey_tran_solve / spicey_tran_solve_waves (waves NULL: vsrc_mask as in spicey_tran_solve; else the descriptors)
 * returning only the node voltages a caller will keep: simulateTRAN filters result.nodeVoltages by the `.PRINT TRAN`
 * probes (lib/analysis/simulateTRAN.ts:240-249), so the others need not cross the bus.  node_sel[n_sel]: node ids
 * (1 .. n_nodes) in the order wanted; v: [steps+1][n_sel][n_inst] (v may be NULL when n_sel == 0).  ielem, state_out,
 * iters, status as in spicey_tran_solve (element currents are not filtered by the reference; pass ielem = NULL to leave
 * them on the device too). */
int32_t spicey_tran_solve_probes(spicey_handle* h, const spicey_elem_table* table, const spicey_sweep* sweep,
                                 double dt, int64_t steps, const spicey_waves* waves, const double* vsrc, const int32_t* vsrc_mask,
                                 const double* state0, const int32_t* node_sel, int32_t n_sel, double* v, double* ielem,
                                 double* state_out, int32_t* iters, int32_t* status, uint32_t flags);

int32_t spicey_tran_solve_waves_device(spicey_handle* h, int32_t dev_index, const spicey_elem_table* table,
                                       const spicey_sweep* sweep, double dt, int64_t steps,
                                       const spicey_waves* waves /* host */, const double* d_vsrc,
                                       const double* d_state0, double* d_v, double* d_ielem,
                                       double* d_state_out, int32_t* d_iters, int32_t* d_status,
                                       uint32_t flags, void* stream);

enum {
  SPICEY_FLAG_STRICT = 1u,      /* reference-order, unfused arithmetic (slow; parity testing) */
  SPICEY_FLAG_FORCE_GMEM = 2u,  /* testing: force the global-scratch tier */
  SPICEY_FLAG_FORCE_CTA = 4u,   /* testing: force a CTA tier even for tiny systems */
  SPICEY_FLAG_DENSE = 8u,       /* never use the sparse program path */
  SPICEY_FLAG_SPARSE = 16u,     /* use the sparse program path even for small batches */
  SPICEY_FLAG_GENERIC_THREAD = 32u, /* testing: transient thread tier without the register-resident kernel */
  SPICEY_FLAG_SERIES_MAJOR = 64u,  /* AC: x is [Nvar][P] and ielem [nAc][P] (one contiguous series per node /
                                      element, coalesced stores on the device) instead of [P][Nvar] / [P][nAc] */
  SPICEY_FLAG_JIT = 128u,          /* compile the per-topology kernel (AC tier 5, TRAN tier 6) even for small batches */
  SPICEY_FLAG_WARP = 512u,         /* AC: use the warp-per-system sparse tier even for small programs (testing) */
  SPICEY_FLAG_NO_WARP = 1024u,     /* AC: never use the warp-per-system sparse tier */
  SPICEY_FLAG_NO_JIT = 256u,       /* never compile: interpreted sparse program (AC), generic kernels (TRAN) */
  SPICEY_FLAG_BAND = 2048u,        /* AC: use the banded + bordered tier even for small batches / small programs, and even when
                                      its renumbered order fails the per-entry check against the netlist order (testing) */
  SPICEY_FLAG_NO_BAND = 4096u,     /* AC: never use the banded + bordered tier */
  SPICEY_FLAG_TILE = 8192u,        /* AC: use the dense register-tile tier even for small batches / sparse circuits (testing) */
  SPICEY_FLAG_NO_TILE = 16384u,    /* AC: never use the dense register-tile tier */
  SPICEY_FLAG_TILE_GENERIC = 32768u /* AC, testing: the register-tile tier stamps from the element table (as it does for
                                      per-instance values) even when the per-topology constants of a plain sweep apply */
};

/* Tooling (no device needed): writes the CUDA source of the compiled straight-line sparse kernel
 * (tier 5) for this element table (sweep: which value slots vary per instance, or NULL) into buf (NUL-terminated, truncated to cap) and returns the
 * size needed, or -1 when the sparse path does not apply.  stats_out[8], optional: values crossing
 * into the back-substitution, of those in shared memory, distinct stamped values, micro-ops,
 * complex FMAs, reciprocals, virtual values, interpreter workspace slots.  with_ielem: bit 0 element
 * currents, bits 16-23 __syncthreads period. */
int64_t spicey_debug_sparse_source(const spicey_elem_table* table, const spicey_sweep* sweep, double pilot_f, int32_t block,
                                   int32_t min_blocks, int32_t smem_slots, int32_t with_ielem, char* buf, int64_t cap,
                                   int32_t* stats_out);

/* Tooling (no device needed): CUDA source of the compiled transient kernel (tier 6) for this element
 * table and sweep (which value slots vary per instance); returns the size needed or -1. */
int64_t spicey_debug_tran_source(const spicey_elem_table* table, const spicey_sweep* sweep, int32_t with_ielem,
                                 char* buf, int64_t cap);

/* The same with device-evaluated waveforms (waves may be NULL). */
int64_t spicey_debug_tran_source_waves(const spicey_elem_table* table, const spicey_sweep* sweep,
                                       const spicey_waves* waves, int32_t with_ielem, char* buf, int64_t cap);

/* Tooling (no device needed): sizes of the warp-per-system program (tier 7) of this element table:
 * out[8] = Nvar, shared-memory pool slots, global workspace slots, rows per step (max), updates, update
 * chunks of 32, back-substitution entries, thread-per-system workspace slots of the same circuit. */
int32_t spicey_debug_warp_stats(const spicey_elem_table* table, double pilot_f, int32_t* out);

/* Tooling (no device needed): the banded + bordered plan (tier 8) of this element table: out[8] = window W, lanes per
 * system, rows per lane, measured half-bandwidth, 1 when the nodes were renumbered, border rows, active border-column
 * mask, workspace values per system.  SPICEY_ERR_UNSUPPORTED when the circuit does not qualify. */
int32_t spicey_debug_band_stats(const spicey_elem_table* table, double pilot_f, int32_t* out);

/* Tooling (no device needed): how far the solution in the banded plan's elimination order is from the solution in the
 * netlist order (the reference's), per entry, both by the reference's algorithm on the host at frequency f:
 * max |x_plan - x_netlist| / max(|x_netlist|, 1e-12 max|x_netlist|); 0 when the plan keeps the netlist order, negative
 * when there is no plan.  A renumbered plan above 5e-10 at the pilot or either end of the sweep is not used by default. */
double spicey_debug_band_order_deviation(const spicey_elem_table* table, double pilot_f, double f);

/* Tooling (no device needed): the CUDA source NVRTC compiles for one band shape; returns the size needed.
 * abmask: bits 0-15 active border columns of band rows, bit 16: (alpha, beta)-only tables (RC circuits),
 * bits 17-18: how the pivot rows reach the workspace (0 plain stores, 1 paired 32-byte stores, 2 TMA tensor store). */
int64_t spicey_debug_band_source(int32_t L, int32_t RPL, int32_t NB, uint32_t abmask, int32_t with_ielem, int32_t warps,
                                 int32_t minb, char* buf, int64_t cap);

/* Tooling (no device needed): the CUDA source NVRTC compiles for the dense register-tile tier (tier 9) of an
 * Nvar-unknown circuit with n_elem elements and n_src sources; tr, tc > 0 force the thread grid, else the library's
 * choice.  with_ielem: bit 0 element currents, bit 1 constant tables (plain frequency sweep), bit 2 (alpha, beta)-only tables.  shape_out[8], optional: TR, TC, MR, MC, warps per CTA, CTAs per SM, expected registers per thread, shared
 * memory per CTA.  Returns the size needed, or -1 when no tile shape fits an SM (Nvar too large). */
int64_t spicey_debug_tile_source(int32_t nvar, int32_t n_elem, int32_t n_src, int32_t tr, int32_t tc, int32_t with_ielem,
                                 int32_t* shape_out, char* buf, int64_t cap);

/* Tooling (no device needed): the CUDA source of the one-warp-per-system dense LU (tier 9, Nvar <= 32, plain sweeps).
 * variant: bit 0 element currents, bit 2 (alpha, beta)-only tables.  shape_out[3], optional: warps per CTA, CTAs per
 * SM, shared memory per CTA.  Returns the size needed, or -1 when Nvar > 32. */
int64_t spicey_debug_warp_lu_source(int32_t nvar, int32_t variant, int32_t* shape_out, char* buf, int64_t cap);

/* Measures this GPU's FP64 FMA peak with a register-only DFMA loop (GFLOP/s), the
 * denominator the FP64-bound roofline is reported against. */
int32_t spicey_measure_fp64_peak(spicey_handle* h, int32_t dev_index, double* gflops_out);

#ifdef __cplusplus
}
#endif
#endif /* SPICEY_NATIVE_H */
