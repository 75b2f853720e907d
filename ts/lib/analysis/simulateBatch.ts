// lib/analysis/simulateBatch.ts — sweep / Monte-Carlo entry points (SURVEY.md 8 f2, f3).
//
// The reference has no batch API: a caller loops simulate() and re-parses every time because
// simulateTRAN mutates the circuit (simulateTRAN.ts:221-237).  Here one parsed circuit plus a map
// "element[.param]" -> Float64Array(nInst) becomes ONE library call; results stay typed-array slabs
// (instance-contiguous), never Complex / number[] objects per point.
//
// Source-parameter sweeps ("v1.pulse.v2", "vctrl.pwl.t1") need the parsed PulseSpec / PWL pairs, which
// parseNetlist.ts:366-383 hides in closures.  Two added lines there keep them on the source record:
//     spec.pulse = p      // next to  spec.waveform = (t) => pulseValue(p, t)
//     spec.pwl = pairs    // next to  spec.waveform = (t) => pwlValue(pairs, t)
// The kernels then evaluate pulseValue / pwlValue themselves, bit-identically (include/spicey_native.h).
//
// NOT EXECUTED IN THIS REPO (no JS runtime, SURVEY.md 8 c); spicey_b200/analysis.py is the tested mirror.
import type { ParsedCircuit } from "../parsing/parseNetlist"
import { acSolveBatch, tranSolveBatch, ELEM, WAVE, type ElemTable, type Sweep, type Waves } from "../native/spiceyNative"
import { packCircuit } from "./packCircuit"

const PARAMS: Record<number, Record<string, number>> = {
  [ELEM.V]: { dc: 0, acmag: 1, acphase: 2 },
  [ELEM.S]: { ron: 0, roff: 1, von: 2, voff: 3 },
  [ELEM.D]: { is: 0, n: 1 },
}
const PULSE_PARAMS = ["v1", "v2", "td", "tr", "tf", "ton", "period", "ncycles"] as const

/** Appends the PULSE / PWL parameters of the V elements as value slots; returns descriptors + slot names. */
function packWaves(ckt: ParsedCircuit, table: ElemTable) {
  const values = Array.from(table.values)
  const kind: number[] = [], valueIdx: number[] = [], nPairs: number[] = []
  const slots: Record<string, number> = {}
  for (const v of ckt.V as any[]) {
    const name = v.name.toLowerCase()
    if (v.pulse) {
      kind.push(WAVE.PULSE); valueIdx.push(values.length); nPairs.push(0)
      PULSE_PARAMS.forEach((nm, q) => (slots[`${name}.pulse.${nm}`] = values.length + q))
      values.push(...PULSE_PARAMS.map((nm) => v.pulse[nm] as number))
    } else if (v.pwl) {
      kind.push(WAVE.PWL); valueIdx.push(values.length); nPairs.push(v.pwl.length)
      v.pwl.forEach((pr: { t: number; v: number }, q: number) => {
        slots[`${name}.pwl.t${q}`] = values.length
        slots[`${name}.pwl.v${q}`] = values.length + 1
        values.push(pr.t, pr.v)
      })
    } else {
      kind.push(WAVE.DC); valueIdx.push(0); nPairs.push(0)
    }
  }
  table.values = Float64Array.from(values)
  const waves: Waves = { kind: Int32Array.from(kind), valueIdx: Int32Array.from(valueIdx), nPairs: Int32Array.from(nPairs) }
  return { waves, slots }
}

function makeSweep(table: ElemTable, nInst: number, overrides: Record<string, Float64Array>, waveSlots: Record<string, number> = {}): Sweep {
  const lower = new Map<string, number>()
  table.names.forEach((n, i) => { if (!lower.has(n.toLowerCase())) lower.set(n.toLowerCase(), i) })
  const keys = Object.keys(overrides)
  const varSlot = new Int32Array(keys.length)
  const varValues = new Float64Array(keys.length * nInst)
  keys.forEach((key, v) => {
    const k = key.toLowerCase()
    let slot = waveSlots[k]
    if (slot === undefined) {
      const [name, param] = lower.has(k) ? [k, ""] : (k.split(".", 2) as [string, string])
      const e = lower.get(name)
      if (e === undefined) throw new Error(`no element named ${key}`)
      const off = param ? PARAMS[table.type[e]!]?.[param] : 0
      if (off === undefined) throw new Error(`element ${name} has no parameter ${param}`)
      slot = table.valueIdx[e]! + off
    }
    if (overrides[key]!.length !== nInst) throw new Error(`override ${key}: expected ${nInst} values`)
    varSlot[v] = slot
    varValues.set(overrides[key]!, v * nInst)
  })
  return { nInst, varSlot, varValues }
}

/** x[Nvar][nInst*F], ielem[nAc][nInst*F] (complex interleaved), status[nInst*F]. */
export function simulateACBatch(ckt: ParsedCircuit, freqs: Float64Array, nInst: number, overrides: Record<string, Float64Array>) {
  const table = packCircuit(ckt)
  return { ...acSolveBatch(table, freqs, makeSweep(table, nInst, overrides)), table }
}

/** v[steps+1][nn][nInst], ielem[steps+1][nElem][nInst], status[nInst]; the circuit is not mutated. */
export function simulateTRANBatch(ckt: ParsedCircuit, dt: number, steps: number, nInst: number, overrides: Record<string, Float64Array>) {
  const table = packCircuit(ckt)
  const { waves, slots } = packWaves(ckt, table)
  const S = ckt.S.filter((s) => s.model), D = ckt.D.filter((d) => d.model)
  const row = [...ckt.C.map((c) => c.vPrev), ...ckt.L.map((l) => l.iPrev), ...S.map((s) => (s.isOn ? 1 : 0)), ...D.map((d) => d.vdPrev)]
  const state0 = new Float64Array(row.length * nInst)
  row.forEach((x, s) => state0.fill(x, s * nInst, (s + 1) * nInst))
  return { ...tranSolveBatch(table, dt, steps, makeSweep(table, nInst, overrides, slots), waves, state0), table }
}
