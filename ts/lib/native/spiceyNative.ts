// lib/native/spiceyNative.ts — bun:ffi binding of include/spicey_native.h.
//
// NOT EXECUTED IN THIS REPO: the build image and the GPU box have no bun/node
// (SURVEY.md §8c), so this file is the reference-side half of the drop-in, kept
// thin and mechanical; the same ABI is exercised from Python (spicey_b200/native.py)
// by tests/.  Drop it into the reference as lib/native/spiceyNative.ts.
import { dlopen, FFIType, ptr, CString, type Pointer } from "bun:ffi"

const { i32, i64, u32, f64, ptr: p, cstring } = FFIType

const LIB_PATH =
  process.env.SPICEY_NATIVE_LIB ?? `${import.meta.dir}/libspicey_native.so`

const { symbols: C } = dlopen(LIB_PATH, {
  spicey_native_abi_version: { args: [], returns: i32 },
  spicey_device_count: { args: [], returns: i32 },
  spicey_last_error: { args: [], returns: cstring },
  spicey_create: { args: [p, i32, p], returns: i32 },
  spicey_destroy: { args: [p], returns: FFIType.void },
  spicey_get_stats: { args: [p, p], returns: i32 },
  spicey_ac_solve: { args: [p, p, p, p, i64, p, p, p, u32], returns: i32 },
  spicey_tran_solve: {
    args: [p, p, p, f64, i64, p, p, p, p, p, p, p, p, u32],
    returns: i32,
  },
  // (h, table, sweep, dt, steps, waves, vsrc, vsrc_mask, state0, node_sel, n_sel, v, ielem, state_out, iters, status, flags)
  spicey_tran_solve_probes: {
    args: [p, p, p, f64, i64, p, p, p, p, p, i32, p, p, p, p, p, u32],
    returns: i32,
  },
  // (h, table, sweep, dt, steps, waves, vsrc, state0, v, ielem, state_out, iters, status, flags)
  spicey_tran_solve_waves: {
    args: [p, p, p, f64, i64, p, p, p, p, p, p, p, p, u32],
    returns: i32,
  },
})

/** SPICEY_ELEM_*; I (independent current source) has no ParsedCircuit counterpart yet: the parser skips `I` lines. */
export const ELEM = { R: 0, C: 1, L: 2, V: 3, S: 4, D: 5, I: 6 } as const
/** SPICEY_FLAG_SERIES_MAJOR: x is [Nvar][P], ielem [nAc][P] — one contiguous slab per series. */
export const FLAG_SERIES_MAJOR = 64
export const STATUS = { OK: 0, SINGULAR: 1, CDIV: 2, R_NONPOS: 3 } as const
export const WAVE = { DC: 0, TABLE: 1, PULSE: 2, PWL: 3 } as const

/** Sweep / Monte-Carlo batch: value slots varSlot[v] take varValues[v * nInst + inst]. */
export type Sweep = { nInst: number; varSlot: Int32Array; varValues: Float64Array }
/** Per-source waveform descriptors (spicey_waves): parameters are value slots of the table. */
export type Waves = { kind: Int32Array; valueIdx: Int32Array; nPairs: Int32Array }

/** Flat element table: typed arrays in the layout of `spicey_elem_table`. */
export type ElemTable = {
  nNodes: number
  type: Int32Array
  n1: Int32Array
  n2: Int32Array
  nc1: Int32Array
  nc2: Int32Array
  valueIdx: Int32Array
  values: Float64Array
  names: string[]
  nVsrc: number
  nAcElem: number
  nState: number
}

function check(rc: number) {
  if (rc !== 0)
    throw new Error(`spicey_native error ${rc}: ${C.spicey_last_error()}`)
}

/** struct spicey_elem_table: 4 x int32 then 7 pointers (72 bytes). */
function tableStruct(t: ElemTable) {
  const buf = new ArrayBuffer(16 + 7 * 8)
  const dv = new DataView(buf)
  dv.setInt32(0, t.nNodes, true)
  dv.setInt32(4, t.type.length, true)
  dv.setInt32(8, t.values.length, true)
  const ptrs = [t.type, t.n1, t.n2, t.nc1, t.nc2, t.valueIdx, t.values]
  ptrs.forEach((a, i) =>
    dv.setBigUint64(16 + 8 * i, BigInt(a.length ? ptr(a) : 0), true),
  )
  return new Uint8Array(buf)
}

/** struct spicey_sweep: int64 n_inst, int32 n_var, int32 reserved, 2 pointers (32 bytes). */
function sweepStruct(s: Sweep) {
  const dv = new DataView(new ArrayBuffer(32))
  dv.setBigInt64(0, BigInt(s.nInst), true)
  dv.setInt32(8, s.varSlot.length, true)
  dv.setBigUint64(16, BigInt(s.varSlot.length ? ptr(s.varSlot) : 0), true)
  dv.setBigUint64(24, BigInt(s.varValues.length ? ptr(s.varValues) : 0), true)
  return new Uint8Array(dv.buffer)
}

/** struct spicey_waves: int32 n_vsrc, int32 reserved, 3 pointers (32 bytes). */
function wavesStruct(w: Waves) {
  const dv = new DataView(new ArrayBuffer(32))
  dv.setInt32(0, w.kind.length, true)
  ;[w.kind, w.valueIdx, w.nPairs].forEach((a, i) =>
    dv.setBigUint64(8 + 8 * i, BigInt(a.length ? ptr(a) : 0), true),
  )
  return new Uint8Array(dv.buffer)
}

let handle: Pointer | null = null
function getHandle(): Pointer {
  if (handle) return handle
  if (C.spicey_native_abi_version() !== 4)
    throw new Error("spicey_native ABI version mismatch")
  const out = new BigUint64Array(1)
  check(C.spicey_create(null, 0, ptr(out)))
  handle = Number(out[0]) as unknown as Pointer
  return handle
}

/** wantCurrents = false: ielem is NULL for the library — the solution vector alone crosses PCIe (16 Nvar bytes per
 *  point instead of 16 (Nvar + nAc)); simulateAC then computes element currents on access. */
export function acSolve(t: ElemTable, freqs: Float64Array, wantCurrents = true) {
  const nvar = t.nNodes + t.nVsrc
  const P = freqs.length
  const x = new Float64Array(P * nvar * 2)
  const ielem = new Float64Array(wantCurrents ? P * t.nAcElem * 2 : 0)
  const status = new Int32Array(P)
  const ts = tableStruct(t)
  check(
    C.spicey_ac_solve(
      getHandle(), ptr(ts), null, ptr(freqs), BigInt(P), ptr(x),
      wantCurrents && t.nAcElem ? ptr(ielem) : null, ptr(status), FLAG_SERIES_MAJOR,
    ),
  )
  return { x, ielem, status, nvar, nPoints: P }
}

export function tranSolve(
  t: ElemTable,
  dt: number,
  steps: number,
  vsrc: Float64Array,
  vsrcMask: Int32Array,
  state0: Float64Array,
  /** node ids (1-based) whose voltages are wanted — the `.PRINT TRAN` probes: only those rows cross the bus, v is [S1][nodeSel.length] */
  nodeSel?: Int32Array,
) {
  const S1 = steps + 1
  const nOut = nodeSel ? nodeSel.length : t.nNodes
  const v = new Float64Array(Math.max(1, S1 * nOut))
  const ielem = new Float64Array(S1 * t.type.length)
  const stateOut = new Float64Array(Math.max(1, t.nState))
  const status = new Int32Array(1)
  const ts = tableStruct(t)
  if (nodeSel) {
    check(
      C.spicey_tran_solve_probes(
        getHandle(), ptr(ts), null, dt, BigInt(steps), null,
        vsrc.length ? ptr(vsrc) : null, ptr(vsrcMask),
        state0.length ? ptr(state0) : null, nodeSel.length ? ptr(nodeSel) : null, nodeSel.length,
        nodeSel.length ? ptr(v) : null, t.type.length ? ptr(ielem) : null, ptr(stateOut), null, ptr(status), 0,
      ),
    )
    return { v, ielem, stateOut, status, nOut }
  }
  check(
    C.spicey_tran_solve(
      getHandle(), ptr(ts), null, dt, BigInt(steps),
      vsrc.length ? ptr(vsrc) : null, ptr(vsrcMask),
      state0.length ? ptr(state0) : null, ptr(v),
      t.type.length ? ptr(ielem) : null, ptr(stateOut), null, ptr(status), 0,
    ),
  )
  return { v, ielem, stateOut, status, nOut }
}

/** AC batch: point p = inst * F + k; series-major slabs x[Nvar][P], ielem[nAc][P]. */
export function acSolveBatch(t: ElemTable, freqs: Float64Array, sweep: Sweep) {
  const nvar = t.nNodes + t.nVsrc
  const P = freqs.length * sweep.nInst
  const x = new Float64Array(P * nvar * 2)
  const ielem = new Float64Array(P * t.nAcElem * 2)
  const status = new Int32Array(P)
  const ts = tableStruct(t), ss = sweepStruct(sweep)
  check(
    C.spicey_ac_solve(
      getHandle(), ptr(ts), ptr(ss), ptr(freqs), BigInt(freqs.length), ptr(x),
      t.nAcElem ? ptr(ielem) : null, ptr(status), FLAG_SERIES_MAJOR,
    ),
  )
  return { x, ielem, status, nvar, nPoints: P }
}

/** TRAN batch with the sources evaluated on the device: v[steps+1][nn][nInst], ielem[steps+1][nElem][nInst]. */
export function tranSolveBatch(
  t: ElemTable, dt: number, steps: number, sweep: Sweep, waves: Waves, state0: Float64Array,
) {
  const S1 = steps + 1, n = sweep.nInst
  const v = new Float64Array(S1 * t.nNodes * n)
  const ielem = new Float64Array(S1 * t.type.length * n)
  const status = new Int32Array(n)
  const ts = tableStruct(t), ss = sweepStruct(sweep), ws = wavesStruct(waves)
  check(
    C.spicey_tran_solve_waves(
      getHandle(), ptr(ts), ptr(ss), dt, BigInt(steps), ptr(ws), null,
      state0.length ? ptr(state0) : null, ptr(v), t.type.length ? ptr(ielem) : null,
      null, null, ptr(status), 0,
    ),
  )
  return { v, ielem, status }
}
