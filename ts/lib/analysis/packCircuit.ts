// lib/analysis/packCircuit.ts — ParsedCircuit -> flat element table (north star:
// "lib/analysis packs each parsed netlist into a flat element table").
// Order R, C, L, V, S, D = the order simulateAC/simulateTRAN push element currents.
import type { ParsedCircuit } from "../parsing/parseNetlist"
import { ELEM, type ElemTable } from "../native/spiceyNative"

export function packCircuit(ckt: ParsedCircuit): ElemTable {
  const type: number[] = [], n1: number[] = [], n2: number[] = []
  const c1: number[] = [], c2: number[] = [], vidx: number[] = []
  const values: number[] = [], names: string[] = []
  const add = (t: number, a: number, b: number, vals: number[], name: string, ca = 0, cb = 0) => {
    type.push(t); n1.push(a); n2.push(b); c1.push(ca); c2.push(cb)
    vidx.push(values.length); values.push(...vals); names.push(name)
  }
  for (const r of ckt.R) add(ELEM.R, r.n1, r.n2, [r.R], r.name)
  for (const c of ckt.C) add(ELEM.C, c.n1, c.n2, [c.C], c.name)
  for (const l of ckt.L) add(ELEM.L, l.n1, l.n2, [l.L], l.name)
  for (const v of ckt.V) add(ELEM.V, v.n1, v.n2, [v.dc || 0, v.acMag || 0, v.acPhaseDeg || 0], v.name)
  for (const s of ckt.S)
    if (s.model) add(ELEM.S, s.n1, s.n2, [s.model.Ron, s.model.Roff, s.model.Von, s.model.Voff], s.name, s.ncPos, s.ncNeg)
  for (const d of ckt.D)
    if (d.model) add(ELEM.D, d.nPlus, d.nMinus, [d.model.Is, d.model.N], d.name)
  const nState = type.filter((t) => t === ELEM.C || t === ELEM.L || t === ELEM.S || t === ELEM.D).length
  return {
    nNodes: ckt.nodes.count() - 1,
    type: Int32Array.from(type), n1: Int32Array.from(n1), n2: Int32Array.from(n2),
    nc1: Int32Array.from(c1), nc2: Int32Array.from(c2), valueIdx: Int32Array.from(vidx),
    values: Float64Array.from(values), names,
    nVsrc: ckt.V.length, nAcElem: ckt.R.length + ckt.C.length + ckt.L.length + ckt.V.length, nState,
  }
}
