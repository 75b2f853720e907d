"""The bun:ffi binding (ts/lib/native/spiceyNative.ts) cannot run here (no JS runtime), but its three hand-packed
structs can be checked: this test builds the same byte blobs — offsets read out of the TypeScript source itself — and
passes them to the library as raw pointers, next to the ctypes.Structure path the rest of the suite uses.  If the
header, the TypeScript offsets or the ctypes structures drift apart, the two calls disagree (or the offsets asserted
below no longer match the source).  No GPU needed: the tooling exports parse the structs on the host."""
import ctypes as C
import os
import re
import struct

import numpy as np

from spicey_b200 import native, workloads as w
from spicey_b200.packing import make_sweep, pack_circuit
from spicey_b200.parsing import parse_netlist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TS = open(os.path.join(ROOT, "ts", "lib", "native", "spiceyNative.ts")).read()


def _fn(name):
    m = re.search(r"function %s\(.*?\n}\n" % name, TS, re.S)
    assert m, name
    return m.group(0)


def ts_table_blob(t):
    """tableStruct(): ArrayBuffer(16 + 7 * 8); int32 nNodes @0, nElem @4, nValues @8; pointers @16 + 8 i in the order
    type, n1, n2, nc1, nc2, valueIdx, values."""
    src = _fn("tableStruct")
    assert "new ArrayBuffer(16 + 7 * 8)" in src
    assert "dv.setInt32(0, t.nNodes, true)" in src and "dv.setInt32(4, t.type.length, true)" in src
    assert "dv.setInt32(8, t.values.length, true)" in src
    assert "[t.type, t.n1, t.n2, t.nc1, t.nc2, t.valueIdx, t.values]" in src and "16 + 8 * i" in src
    buf = bytearray(16 + 7 * 8)
    struct.pack_into("<iii", buf, 0, t.n_nodes, t.n_elem, int(t.values.shape[0]))
    for i, a in enumerate((t.type, t.n1, t.n2, t.nc1, t.nc2, t.value_idx, t.values)):
        struct.pack_into("<Q", buf, 16 + 8 * i, a.ctypes.data if a.size else 0)
    return (C.c_char * len(buf)).from_buffer(buf)


def ts_sweep_blob(s):
    """sweepStruct(): 32 bytes; int64 nInst @0, int32 nVar @8, pointers varSlot @16, varValues @24."""
    src = _fn("sweepStruct")
    assert "new ArrayBuffer(32)" in src and "dv.setBigInt64(0, BigInt(s.nInst), true)" in src
    assert "dv.setInt32(8, s.varSlot.length, true)" in src and "dv.setBigUint64(16," in src and "dv.setBigUint64(24," in src
    buf = bytearray(32)
    struct.pack_into("<qi", buf, 0, s.n_inst, int(s.var_slot.shape[0]))
    struct.pack_into("<QQ", buf, 16, s.var_slot.ctypes.data, s.var_values.ctypes.data)
    return (C.c_char * len(buf)).from_buffer(buf)


def ts_waves_blob(wv):
    """wavesStruct(): 32 bytes; int32 nVsrc @0, pointers kind, valueIdx, nPairs @8 + 8 i."""
    src = _fn("wavesStruct")
    assert "new ArrayBuffer(32)" in src and "dv.setInt32(0, w.kind.length, true)" in src
    assert "[w.kind, w.valueIdx, w.nPairs]" in src and "8 + 8 * i" in src
    buf = bytearray(32)
    struct.pack_into("<i", buf, 0, int(wv.kind.shape[0]))
    for i, a in enumerate((wv.kind, wv.value_idx, wv.n_pairs)):
        struct.pack_into("<Q", buf, 8 + 8 * i, a.ctypes.data)
    return (C.c_char * len(buf)).from_buffer(buf)


def test_struct_sizes_match_the_header():
    assert C.sizeof(native.ElemTableStruct) == 72 and C.sizeof(native.SweepStruct) == 32 and C.sizeof(native.WavesStruct) == 32
    assert C.sizeof(native.StatsStruct) == 72
    assert native.ElemTableStruct.type.offset == 16 and native.ElemTableStruct.values.offset == 64
    assert native.SweepStruct.var_slot.offset == 16 and native.SweepStruct.var_values.offset == 24
    assert native.WavesStruct.kind.offset == 8 and native.WavesStruct.n_pairs.offset == 24
    assert "spicey_native_abi_version() !== 4" in TS and "#define SPICEY_NATIVE_ABI_VERSION 4" in open(
        os.path.join(ROOT, "include", "spicey_native.h")).read()


def _raw(fn, argtypes):
    """The same export bound with raw void pointers, as bun:ffi passes them."""
    f = getattr(C.CDLL(native.load_library()._name), fn)
    f.argtypes, f.restype = argtypes, C.c_int64
    return f


def test_typescript_blobs_drive_the_library_like_the_ctypes_structures():
    lib = native.load_library()
    vp = C.c_void_p
    # element table: the band / warp planners read every array of the table
    tb = pack_circuit(parse_netlist(w.rc_mesh(8)))
    want = native.band_plan_stats(tb, 300.0)
    out = (C.c_int32 * 8)()
    f = _raw("spicey_debug_band_stats", [vp, C.c_double, vp])
    blob = ts_table_blob(tb)
    assert f(C.addressof(blob), 300.0, C.addressof(out)) == 0
    assert list(out) == list(want.values())
    # element table + sweep: the generated per-instance kernel source names the swept slots
    ck = parse_netlist(w.rc_ladder(6))
    tb = pack_circuit(ck)
    n = 5
    sw = make_sweep(tb, n, {"r2": np.linspace(900, 1100, n), "c3": np.linspace(0.9e-9, 1.1e-9, n)})
    ref_src, _ = native.sparse_kernel_source(tb, 1000.0, sweep=sw)
    g = _raw("spicey_debug_sparse_source", [vp, vp, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int64, vp])
    tblob, sblob = ts_table_blob(tb), ts_sweep_blob(sw)
    args = (C.addressof(tblob), C.addressof(sblob), 1000.0, 192, 1, 75, 1 | (4 << 16))
    need = g(*args, None, 0, None)
    assert need > 0
    buf = C.create_string_buffer(need)
    g(*args, C.addressof(buf), need, None)
    assert buf.value.decode() == ref_src
    # element table + waves: the compiled transient kernel evaluates the PULSE from the slots the descriptor names
    ckp = parse_netlist("* p\nV1 1 0 PULSE(0 5 1u 1n 1n 5u 10u)\nR1 1 2 1k\nC1 2 0 1u\n.tran 0.1u 20u\n.end\n")
    tbw = pack_circuit(ckp, device_waves=True)
    ref_t = native.tran_kernel_source(tbw, waves=tbw.waves)
    h = _raw("spicey_debug_tran_source_waves", [vp, vp, vp, C.c_int32, vp, C.c_int64])
    tblob, wblob = ts_table_blob(tbw), ts_waves_blob(tbw.waves)
    need = h(C.addressof(tblob), None, C.addressof(wblob), 1, None, 0)
    assert need > 0
    buf = C.create_string_buffer(need)
    h(C.addressof(tblob), None, C.addressof(wblob), 1, C.addressof(buf), need)
    assert buf.value.decode() == ref_t and "pulse" in ref_t.lower()
    del lib
