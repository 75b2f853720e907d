// Drop-in replacement of lib/analysis/simulateAC.ts: same signature and result shape,
// loop body (:80-127 of the reference) replaced by pack -> FFI -> unpack.
import { Complex } from "../math/Complex"
import type { ParsedCircuit } from "../parsing/parseNetlist"
import { logspace } from "../utils/logspace"
import { acSolve, STATUS } from "../native/spiceyNative"
import { packCircuit } from "./packCircuit"

function buildFrequencyArray(mode: "dec" | "lin", N: number, f1: number, f2: number) {
  if (mode === "dec") return logspace(f1, f2, N)
  const npts = Math.max(2, N)
  const step = (f2 - f1) / (npts - 1)
  return Array.from({ length: npts }, (_, i) => f1 + i * step)
}

/** Complex[] view over one contiguous series of a series-major slab (SURVEY.md §8 f1):
 *  row `row` of a [rows][count] complex array, materialising Complex objects only on access. */
function complexSeries(slab: Float64Array, row: number, count: number): Complex[] {
  const view = slab.subarray(2 * row * count, 2 * (row + 1) * count)
  return new Proxy([] as Complex[], {
    get(_t, key) {
      if (key === "length") return count
      const k = typeof key === "string" ? Number(key) : NaN
      if (Number.isInteger(k) && k >= 0 && k < count) return new Complex(view[2 * k], view[2 * k + 1])
      return (Array.prototype as any)[key]
    },
  })
}

/** Element-current series computed on access from the node-voltage slab with the reference's own expression
 *  (simulateAC.ts:94-126 of the reference: Y.mul(v1.sub(v2))): nothing is transferred from the device for it. */
function lazyCurrentSeries(
  freqs: number[], count: number,
  admittance: (f: number) => Complex, v1: Complex[] | null, v2: Complex[] | null,
): Complex[] {
  const zero = Complex.from(0, 0)
  return new Proxy([] as Complex[], {
    get(_t, key) {
      if (key === "length") return count
      const k = typeof key === "string" ? Number(key) : NaN
      if (Number.isInteger(k) && k >= 0 && k < count)
        return admittance(freqs[k]!).mul((v1 ? v1[k]! : zero).sub(v2 ? v2[k]! : zero))
      return (Array.prototype as any)[key]
    },
  })
}

function simulateAC(ckt: ParsedCircuit, opts: { lazyCurrents?: boolean } = {}) {
  if (!ckt.analyses.ac) return null
  const { mode, N, f1, f2 } = ckt.analyses.ac
  const freqs = buildFrequencyArray(mode, N, f1, f2)
  const table = packCircuit(ckt)
  const { x, ielem, status } = acSolve(table, Float64Array.from(freqs), !opts.lazyCurrents)
  for (let k = 0; k < status.length; k++) {
    const st = status[k]
    if (st === STATUS.OK) continue
    if (st === STATUS.R_NONPOS) {
      const bad = ckt.R.find((r) => r.R <= 0)
      throw new Error(`R ${bad?.name} must be > 0`)
    }
    throw new Error(st === STATUS.SINGULAR ? "Singular matrix (complex)" : "Complex divide by ~0")
  }
  const nodeVoltages: Record<string, Complex[]> = {}
  ckt.nodes.rev.forEach((name, id) => {
    if (id !== 0) nodeVoltages[name] = complexSeries(x, id - 1, freqs.length)
  })
  const elementCurrents: Record<string, Complex[]> = {}
  if (opts.lazyCurrents) {
    const twoPi = 2 * Math.PI, F = freqs.length
    const volt = (id: number) => (id === 0 ? null : complexSeries(x, id - 1, F))
    for (const r of ckt.R) elementCurrents[r.name] ||= lazyCurrentSeries(freqs, F, () => Complex.from(1 / r.R, 0), volt(r.n1), volt(r.n2))
    for (const c of ckt.C) elementCurrents[c.name] ||= lazyCurrentSeries(freqs, F, (f) => Complex.from(0, twoPi * f * c.C), volt(c.n1), volt(c.n2))
    for (const l of ckt.L)
      elementCurrents[l.name] ||= lazyCurrentSeries(freqs, F, (f) => {
        const denom = Complex.from(0, twoPi * f * l.L)
        return denom.abs() < 1e-15 ? Complex.from(0, 0) : Complex.from(1, 0).div(denom)
      }, volt(l.n1), volt(l.n2))
    for (const vs of ckt.V) elementCurrents[vs.name] ||= complexSeries(x, vs.index, F)
    return { freqs, nodeVoltages, elementCurrents }
  }
  table.names.slice(0, table.nAcElem).forEach((name, e) => {
    elementCurrents[name] ||= complexSeries(ielem, e, freqs.length)
  })
  return { freqs, nodeVoltages, elementCurrents }
}

export { simulateAC }
