// Drop-in replacement of lib/analysis/simulateTRAN.ts: same signature, result shape and
// side effects (final vPrev/iPrev/vdPrev/isOn written back into ckt), step loop
// (:146-238 of the reference) replaced by pack -> FFI -> unpack.
import { EPS } from "../constants/EPS"
import type { ParsedCircuit } from "../parsing/parseNetlist"
import { tranSolve, STATUS } from "../native/spiceyNative"
import { packCircuit } from "./packCircuit"

function computeEffectiveTimeStep(dtRequested: number, tstop: number) {
  const dtEff = dtRequested > EPS ? dtRequested : Math.max(tstop / 1000, EPS)
  const steps = Math.max(1, Math.ceil(tstop / Math.max(dtEff, EPS)))
  const dt = steps > 0 ? tstop / steps : tstop
  return { dt, steps }
}

function simulateTRAN(ckt: ParsedCircuit) {
  if (!ckt.analyses.tran) return null
  const { dt, steps } = computeEffectiveTimeStep(ckt.analyses.tran.dt, ckt.analyses.tran.tstop)
  const table = packCircuit(ckt)
  const S1 = steps + 1
  const times = Array.from({ length: S1 }, (_, k) => k * dt)
  // waveforms are closures in the reference: pre-sample them on the host (SURVEY H-G)
  const vsrc = new Float64Array(ckt.V.length * S1)
  const mask = new Int32Array(Math.max(1, ckt.V.length))
  ckt.V.forEach((vs, i) => {
    if (!vs.waveform) return
    mask[i] = 1
    for (let k = 0; k < S1; k++) vsrc[i * S1 + k] = vs.waveform(times[k]!)
  })
  const S = ckt.S.filter((s) => s.model), D = ckt.D.filter((d) => d.model)
  const state0 = Float64Array.from([
    ...ckt.C.map((c) => c.vPrev), ...ckt.L.map((l) => l.iPrev),
    ...S.map((s) => (s.isOn ? 1 : 0)), ...D.map((d) => d.vdPrev),
  ])
  // `.PRINT TRAN` probes (lib/analysis/simulateTRAN.ts:240-249 keeps those node voltages only): the others stay on the GPU
  const names = ckt.nodes.rev.slice(1)
  let sel: number[] | undefined
  if (ckt.probes.tran.length > 0) {
    const upper = ckt.probes.tran.map((p) => p.toUpperCase())
    sel = names.flatMap((n, i) => (upper.includes(n.toUpperCase()) ? [i + 1] : []))
  }
  const { v, ielem, stateOut, status, nOut } = tranSolve(
    table, dt, steps, vsrc, mask, state0, sel ? Int32Array.from(sel) : undefined,
  )
  if (status[0] !== STATUS.OK) throw new Error("Singular matrix (real)")
  let si = 0
  for (const c of ckt.C) c.vPrev = stateOut[si++]!
  for (const l of ckt.L) l.iPrev = stateOut[si++]!
  for (const s of S) s.isOn = stateOut[si++]! !== 0
  for (const d of D) d.vdPrev = stateOut[si++]!
  const ne = table.type.length
  const nodeVoltages: Record<string, number[]> = {}
  const outIds = sel ?? names.map((_, i) => i + 1)
  outIds.forEach((id, col) => {
    nodeVoltages[names[id - 1]!] = Array.from({ length: S1 }, (_, k) => v[k * nOut + col]!)
  })
  const elementCurrents: Record<string, number[]> = {}
  for (let k = 0; k < S1; k++)
    table.names.forEach((name, e) => (elementCurrents[name] ||= []).push(ielem[k * ne + e]!))
  return { times, nodeVoltages, elementCurrents }
}

export { simulateTRAN }
