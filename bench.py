#!/usr/bin/env python3
"""Benchmark of the batched MNA-solve hot path (BASELINE.json metric: solves/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5|cfg2mc|dense64] [--scaling weak|strong]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # the reference algorithm's CPU port on all host cores

A "step" is one pass of the hot path over one batch.  The job is ONE batch sharded over the N ranks in contiguous
ranges of its batch axis (SURVEY.md 8 e: frequency index for AC, instance index for TRAN; rank r takes
shard_range(total, r, N); element table replicated; no collective on the data path):
  cfg2 (default, the configuration the metric is quoted on), weak scaling: the 64-node RC ladder swept over
       N x 1,000,000 + 1 log-spaced points from 1 Hz to 100 kHz (.ac dec 200000*N 1 100k); N = 1 is BASELINE's
       1,000,001-point sweep.  --scaling strong keeps the 1,000,001 points and splits them.
  cfg4 strong scaling: the 8,000,001-point sweep of the 16x16 mesh split over the ranks (solved in chunks of the
       device result buffers); cfg3 / cfg5 strong scaling: the 65,536 / 100,000 instances split over the ranks.
  value      whole-job solves/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same through the host-buffer C-ABI call (pinned host buffers, H2D + D2H inside)
  roofline   dominant kernel vs the FP64 (AC) or HBM-write (TRAN) roofline, SURVEY.md 8(d) figures
  cpu_baseline  oracle/oracle.c (C restatement of the reference algorithm) on a bounded sample
  secondary  (default line only) short legs of cfg3, cfg4 and cfg5 at their BASELINE sizes: value, roofline, e2e
After the timed region every rank checks a sample of each of its chunks against the oracle (1e-9 AC, 1e-6 TRAN).
Timing is barrier + synchronize on both sides, CUDA events, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

from spicey_b200 import workloads as W  # noqa: E402
from spicey_b200.parsing import compute_effective_time_step, parse_netlist  # noqa: E402


def fma_count(n):
    return (n - 1) * n * (2 * n - 1) // 6 + n * (n - 1)


def f_cplx(n):  # complex FMA = 8 flop, complex divide = 11
    return 8 * fma_count(n) + 11 * (n * (n + 1) // 2)


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    NVML is polled from a thread every millisecond (a cfg2 step is 0.7 ms, far below nvidia-smi's 100 ms
    loop); nvidia-smi is the fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.rows, self.proc = index, uuid, [], None
        self.nvml, self.handle, self.samples, self.masks, self.run = None, None, [], 0, False
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        try:
                            h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                            break
                        except Exception:
                            h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.handle = pynvml, h
        except Exception:
            self.nvml = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        while self.run:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.masks |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.masks |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.nvml is not None:
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def pause(self):
        """End of a sampled region (NVML path): keeps what was collected."""
        if self.nvml is not None and self.run:
            self.run = False
            self.thread.join(timeout=1)

    def stop(self):
        if self.nvml is not None:
            self.pause()
            try:
                mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            except Exception:
                mx = None
            reasons = [n for n, bit in self.BITS if self.masks & bit]
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": mx,
                    "reasons": reasons, "samples": len(self.samples), "source": "nvml, 1 ms polling"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


def bind_to_gpu_numa(local):
    """Multi-rank runs: pin this rank to the CPUs NVML names as local to its GPU, so that the pinned host buffers
    of the e2e leg are allocated on the NUMA node the GPU's PCIe link hangs off (eight ranks left floating put
    most buffers on one node and share its memory controller and the inter-socket link).  Returns a note."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = None
        for cand in ("GPU-" + uuid, uuid):
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                break
            except Exception:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode())
                    break
                except Exception:
                    h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = cpus & allowed
        if not pick or pick == allowed:
            return "no NUMA binding (affinity mask %d cpus, allowed %d)" % (len(cpus), len(allowed))
        os.sched_setaffinity(0, pick)
        return "rank bound to the %d CPUs local to its GPU" % len(pick)
    except Exception as e:  # no NVML / no permission: run unbound
        return "no NUMA binding (%s)" % type(e).__name__


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ---- workloads --------------------------------------------------------------------------

AC_NETLISTS = {"cfg1": lambda n: W.README_RC, "cfg2": lambda n: W.rc_ladder(64, ppd=200000 * n),
               "cfg2mc": lambda n: W.rc_ladder(64), "cfg4": lambda n: W.rc_mesh(16),
               # not a BASELINE config: cfg2's size with a matrix that has no structural zero (complete RC graph) — the
               # workload of the dense pivoting LU of north_star's piece (2); 2,080 element currents per solve
               "dense64": lambda n: W.rc_dense(64)}
DEFAULT_SCALING = {"cfg1": "strong", "cfg2": "weak", "cfg2mc": "strong", "cfg3": "strong", "cfg4": "strong", "cfg5": "strong",
                   "dense64": "strong"}
DENSE64_POINTS = 200000   # of the sweep's 1,000,001: 34 KB of results per solve


def load_workload(name, world=1, scaling=None, points=None, instances=None, device_waves=False):
    """The WHOLE job (all ranks): netlist, batch axis, labels.  `units` counts the metric's unit over the whole job."""
    import spicey_b200 as sp
    from spicey_b200.packing import make_sweep, pack_circuit, sample_sources, initial_state
    scaling = scaling or DEFAULT_SCALING[name]
    mult = world if scaling == "weak" else 1
    wl = {"name": name, "scaling": scaling}
    if name == "cfg2mc":  # not a BASELINE config: cfg2's ladder with every R and C swept +-5 % (AC Monte-Carlo axis)
        ck = parse_netlist(W.rc_ladder(64))
        n = (instances or 4096) * mult
        freqs = np.logspace(0, 5, 245)
        u = W.splitmix_uniform_pm1(n, 126)
        ov = {}
        for k in range(1, 64):
            ov["r%d" % k] = 1000.0 * (1 + 0.05 * u[:, k - 1])
            ov["c%d" % k] = 1e-9 * (1 + 0.05 * u[:, 62 + k])
        table = pack_circuit(ck)
        wl.update(kind="ac", axis="instance", ckt=ck, table=table, freqs=freqs, n_inst=n, overrides=ov,
                  units=int(n * freqs.shape[0]), batch=n, unit_per_batch=int(freqs.shape[0]),
                  label="cfg2mc: 64-node RC ladder, %d Monte-Carlo instances x %d frequencies (all R, C +-5 %%)" % (
                      n, freqs.shape[0]),
                  flops_per_unit=f_cplx(table.nvar), bytes_per_unit=8 + 16 * table.nvar + 16 * table.n_ac_elem)
        return wl
    if name in AC_NETLISTS:
        ck = parse_netlist(AC_NETLISTS[name](mult))
        freqs = np.array(sp.analysis.ac_frequencies(ck), dtype=np.float64)
        full = int(freqs.shape[0])
        if name == "dense64" and not points:
            points = DENSE64_POINTS
        if points:
            freqs = np.ascontiguousarray(freqs[:: max(1, freqs.shape[0] // points)][:points])
        table = pack_circuit(ck)
        P = int(freqs.shape[0])
        what = {"cfg2": "64-node RC ladder .ac dec %d 1 100k" % (200000 * mult), "cfg4": "16x16 RC mesh .ac dec 1600000 1 100k",
                "cfg1": "README RC low-pass", "dense64": "complete RC graph on 64 nodes (dense MNA matrix) .ac dec 200000 10k 1g"}[name]
        wl.update(kind="ac", axis="frequency", ckt=ck, table=table, freqs=freqs, n_inst=1, overrides=None,
                  units=P, batch=P, unit_per_batch=1,
                  label="%s: %s (%s points%s, Nvar=%d, c128 LU)" % (
                      name, what, "{:,}".format(P), "" if P == full else " of the {:,}".format(full), table.nvar),
                  flops_per_unit=f_cplx(table.nvar), bytes_per_unit=8 + 16 * table.nvar + 16 * table.n_ac_elem)
        return wl
    text, ovf, n_full = {"cfg3": (W.RLC_TANK, W.rlc_tank_overrides, 65536),
                         "cfg5": (W.RECTIFIER, W.rectifier_overrides, 100000)}[name]
    n = (instances or n_full) * mult
    ck = parse_netlist(text)
    ov = {k: v[:n] for k, v in ovf(max(n, n_full)).items()}
    dt, steps = compute_effective_time_step(ck.analyses.tran.dt, ck.analyses.tran.tstop)
    table = pack_circuit(ck, device_waves=device_waves)   # device_waves: PULSE evaluated by the kernel
    vsrc, mask = sample_sources(ck, dt, steps)
    wl.update(kind="tran", axis="instance", ckt=ck, table=table, dt=dt, steps=steps, vsrc=vsrc, mask=mask, overrides=ov,
              n_inst=n, units=n * (steps + 1), batch=n, unit_per_batch=steps + 1,
              state0=initial_state(ck, table, n), device_waves=device_waves,
              label="%s: %s, %s instances x %d recorded steps (Nvar=%d, f64)" % (
                  name, "RLC tank .tran 1u 1m Monte-Carlo" if name == "cfg3" else "diode half-wave rectifier .tran 1u 3m sweep",
                  "{:,}".format(n), steps + 1, table.nvar),
              flops_per_unit=55, bytes_per_unit=8 * (table.n_nodes + table.n_elem))
    del make_sweep
    return wl


def config_of(wl, world):
    """The part of the JSON line both arms print identically."""
    return {"workload": wl["label"], "scaling": wl["scaling"],
            "parallelism": "one batch sharded over %d rank%s in contiguous %s ranges, no collective on the data path" % (
                world, "" if world == 1 else "s", wl["axis"])}


# ---- reference arm: the reference algorithm's CPU port on all host cores ------------------

def cpu_rate(wl, target_s=12.0, threads=None):
    """solves/s of oracle/oracle.c on a bounded sample of the workload (about target_s of CPU work)."""
    from oracle import c_oracle as co
    threads = threads or os.cpu_count() or 1
    if wl["kind"] == "ac" and wl.get("overrides"):
        freqs, n_inst = wl["freqs"], wl["n_inst"]
        m = min(n_inst, 4 * threads)
        t0 = time.perf_counter()
        co.ac_solve(wl["ckt"], freqs, n_inst=m, overrides={k: v[:m] for k, v in wl["overrides"].items()}, nthreads=threads)
        rate = m * len(freqs) / max(1e-9, time.perf_counter() - t0)
        m2 = int(min(n_inst, max(m, rate * target_s / len(freqs))))
        t0 = time.perf_counter()
        co.ac_solve(wl["ckt"], freqs, n_inst=m2, overrides={k: v[:m2] for k, v in wl["overrides"].items()}, nthreads=threads)
        dt = time.perf_counter() - t0
        return m2 * len(freqs) / dt, threads, "first %d of %d instances x %d frequencies (%.1f s)" % (m2, n_inst, len(freqs), dt)
    if wl["kind"] == "ac":
        freqs = wl["freqs"]
        probe = freqs[:: max(1, len(freqs) // (64 * threads))][: 64 * threads]
        t0 = time.perf_counter()
        co.ac_solve(wl["ckt"], probe, nthreads=threads)
        rate = len(probe) / max(1e-9, time.perf_counter() - t0)
        n = int(min(len(freqs), max(len(probe), rate * target_s)))
        stride = max(1, len(freqs) // n)
        sample = freqs[::stride][:n]
        t0 = time.perf_counter()
        co.ac_solve(wl["ckt"], sample, nthreads=threads)
        dt = time.perf_counter() - t0
        return len(sample) / dt, threads, "every %d-th of %d frequencies (%d points, %.1f s)" % (
            stride, len(freqs), len(sample), dt)
    n = min(wl["n_inst"], 8 * threads)
    ov = {k: v[:n] for k, v in wl["overrides"].items()}
    t0 = time.perf_counter()
    co.tran_solve(wl["ckt"], wl["dt"], wl["steps"], n_inst=n, overrides=ov, nthreads=threads)
    rate = n / max(1e-9, time.perf_counter() - t0)
    n2 = int(min(wl["n_inst"], max(n, rate * target_s)))
    ov = {k: v[:n2] for k, v in wl["overrides"].items()}
    t0 = time.perf_counter()
    co.tran_solve(wl["ckt"], wl["dt"], wl["steps"], n_inst=n2, overrides=ov, nthreads=threads)
    dt = time.perf_counter() - t0
    return n2 * (wl["steps"] + 1) / dt, threads, "first %d of %d instances (%d recorded steps, %.1f s)" % (
        n2, wl["n_inst"], n2 * (wl["steps"] + 1), dt)


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    wl = load_workload(args.workload, max(1, args.gpus), args.scaling, args.points, args.instances)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_rate(wl, target_s=0.5)
    rates, sample = [], ""
    per_step = max(2.0, min(15.0, 120.0 / max(1, args.steps)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, threads, sample = cpu_rate(wl, target_s=per_step)
        rates.append(r)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "batched MNA solves/sec", "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (time.perf_counter() - t0) / max(1, args.steps), "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "c128" if wl["kind"] == "ac" else "f64",
        "data": "synthetic", "config": config_of(wl, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference algorithm (oracle/oracle.c); the reference itself "
                                 "is TypeScript and no JS runtime exists in this image"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---- native arm -------------------------------------------------------------------------

DEVICE_RESULT_BYTES = 12 << 30     # device result buffers of one AC chunk (x + ielem): larger sweeps are solved in chunks
HOST_RESULT_BYTES = 8 << 30        # pinned host buffers of the end-to-end leg: larger shards time a slab of this size


class Leg:
    """One rank's share [lo, hi) of one workload's batch axis: buffers, the resident step, the end-to-end step, checks."""

    def __init__(self, eng, wl, rank, world, dev, stream, dense=False):
        import torch
        from spicey_b200 import native
        from spicey_b200.packing import make_sweep
        from spicey_b200.sharding import shard_range
        self.eng, self.wl, self.dev, self.stream, self.native, self.torch = eng, wl, dev, stream, native, torch
        self.table = table = wl["table"]
        self.lo, self.hi = shard_range(wl["batch"], rank, world)
        self.nb = nb = self.hi - self.lo                      # my slice of the batch axis
        self.units = nb * wl["unit_per_batch"]
        self.pins, self.launches_per_step, self.tier, self.fallback = [], 0, 0, 0
        ov = wl.get("overrides")
        self.ov = None if ov is None else {k: np.ascontiguousarray(v[self.lo:self.hi]) for k, v in ov.items()}
        if wl["kind"] == "ac":
            self.flags = native.FLAG_SERIES_MAJOR | (native.FLAG_DENSE if dense else 0)
            if wl["axis"] == "frequency":
                self.freqs = np.ascontiguousarray(wl["freqs"][self.lo:self.hi])
                self.sweep, self.P = None, nb
            else:
                self.freqs = wl["freqs"]
                self.sweep, self.P = make_sweep(table, nb, self.ov), nb * int(wl["freqs"].shape[0])
            per_point = 16 * (table.nvar + table.n_ac_elem)
            self.chunk = self.P if self.sweep is not None else max(1, min(self.P, DEVICE_RESULT_BYTES // per_point))
            if self.chunk < self.P and self.P >= native_jit_min_points():
                self.flags |= native.FLAG_JIT    # the compile-or-interpret decision looks at the whole sweep, not at one chunk
            self.ld = eng.series_ld(self.chunk)
            self.d_freqs = torch.from_numpy(self.freqs).to(dev)
            self.d_var = torch.from_numpy(self.sweep.var_values).to(dev) if self.sweep is not None else None
            self.d_x = torch.empty((table.nvar, self.ld), dtype=torch.complex128, device=dev)
            self.d_i = torch.empty((table.n_ac_elem, self.ld), dtype=torch.complex128, device=dev)
            self.d_s = torch.empty(self.chunk, dtype=torch.int32, device=dev)
            self.result_bytes = self.P * per_point
        else:
            self.sweep = make_sweep(table, nb, self.ov)
            S1 = wl["steps"] + 1
            self.d_var = torch.from_numpy(self.sweep.var_values).to(dev)
            self.d_vsrc = torch.from_numpy(wl["vsrc"]).to(dev)
            self.state0 = np.ascontiguousarray(wl["state0"][:, self.lo:self.hi])
            self.d_st0 = torch.from_numpy(self.state0).to(dev)
            self.d_v = torch.empty((S1, table.n_nodes, nb), dtype=torch.float64, device=dev)
            self.d_i = torch.empty((S1, table.n_elem, nb), dtype=torch.float64, device=dev)
            self.d_s = torch.empty(nb, dtype=torch.int32, device=dev)
            self.result_bytes = self.units * wl["bytes_per_unit"]

    # -- device-resident --------------------------------------------------------------
    def chunks(self):
        return [(c, min(self.chunk, self.P - c)) for c in range(0, self.P, self.chunk)] if self.wl["kind"] == "ac" else [(0, self.nb)]

    def solve_chunk(self, c0, n):
        if self.sweep is not None:   # instance axis: one call
            self.eng.ac_solve_device(self.table, self.d_freqs.data_ptr(), int(self.freqs.shape[0]), self.d_x.data_ptr(),
                                     self.d_i.data_ptr(), self.d_s.data_ptr(), sweep=self.sweep,
                                     d_var_values=self.d_var.data_ptr(), flags=self.flags, stream=self.stream.cuda_stream,
                                     series_ld=self.ld)
        else:
            self.eng.ac_solve_device(self.table, self.d_freqs.data_ptr() + 8 * c0, n, self.d_x.data_ptr(), self.d_i.data_ptr(),
                                     self.d_s.data_ptr(), flags=self.flags, stream=self.stream.cuda_stream, series_ld=self.ld)

    def step_resident(self):
        wl = self.wl
        if wl["kind"] == "ac":
            for c0, n in self.chunks():
                self.solve_chunk(c0, n)
        else:
            self.eng.tran_solve_device(self.table, wl["dt"], wl["steps"], self.d_vsrc.data_ptr(), wl["mask"], self.d_st0.data_ptr(),
                                       self.d_v.data_ptr(), self.d_i.data_ptr(), None, None, self.d_s.data_ptr(), sweep=self.sweep,
                                       d_var_values=self.d_var.data_ptr(), stream=self.stream.cuda_stream, waves=self.table.waves)

    def note_stats(self):
        st = self.eng.stats()   # reads the fallback counter back: synchronises, hence outside any timed loop
        self.tier, self.fallback, self.cfma = st["tier"], int(st["fallback_solves"]), int(st["program_cfma"])
        self.launches_per_step = int(st["kernel_launches"]) * len(self.chunks())

    def verify(self, per_chunk=4):
        """Untimed pass over every chunk of my range: status 0 everywhere, a sample of each chunk against the oracle."""
        from oracle import c_oracle as co
        torch, wl = self.torch, self.wl
        worst = 0.0
        if wl["kind"] == "ac":
            nac, nv = self.table.n_ac_elem, self.table.nvar
            for c0, n in self.chunks():
                self.solve_chunk(c0, n)
                torch.cuda.synchronize()
                npts = n if self.sweep is None else self.P
                assert int(self.d_s[:npts].max().item()) == 0, "non-zero status in the bench run (%s)" % wl["name"]
                pick = np.unique(np.linspace(0, npts - 1, per_chunk).astype(np.int64))
                sel = torch.from_numpy(pick).to(self.dev)
                x = self.d_x[:, sel].T.cpu().numpy()
                ie = self.d_i[:, sel].T.cpu().numpy()
                if self.sweep is None:
                    xr, ir, st = co.ac_solve(wl["ckt"], self.freqs[c0 + pick], nthreads=4)
                else:
                    F = int(self.freqs.shape[0])
                    rows = []
                    for p_ in pick:
                        inst, k = int(p_) // F, int(p_) % F
                        rows.append(co.ac_solve(wl["ckt"], self.freqs[k:k + 1], n_inst=1,
                                                overrides={key: v[inst:inst + 1] for key, v in self.ov.items()}))
                    xr = np.concatenate([r_[0] for r_ in rows]); ir = np.concatenate([r_[1] for r_ in rows])
                    st = np.concatenate([r_[2] for r_ in rows])
                assert int(st.max()) == 0
                xr, ir = xr.reshape(len(pick), nv), ir.reshape(len(pick), nac)
                ex = float(np.max(np.abs(x - xr) / np.maximum(np.abs(xr), 1e-300)))
                if wl["name"] == "dense64":   # a resistor between two nodes of nearly equal voltage: relative to the point's largest current
                    ei = float(np.max(np.abs(ie - ir) / np.max(np.abs(ir), axis=1, keepdims=True)))
                else:
                    ei = float(np.max(np.abs(ie - ir) / np.maximum(np.abs(ir), 1e-300))) if nac else 0.0
                worst = max(worst, ex, ei)
            assert worst <= 1e-9, "bench results differ from the oracle: %.3e (%s)" % (worst, wl["name"])
        else:
            self.step_resident()
            torch.cuda.synchronize()
            assert int(self.d_s.max().item()) == 0, "non-zero status in the bench run (%s)" % wl["name"]
            pick = np.unique(np.linspace(0, self.nb - 1, per_chunk).astype(np.int64))
            sel = torch.from_numpy(pick).to(self.dev)
            v = self.d_v[:, :, sel].cpu().numpy()
            ie = self.d_i[:, :, sel].cpu().numpy()
            for j, inst in enumerate(pick):
                vr, ir, _, st, _ = co.tran_solve(wl["ckt"], wl["dt"], wl["steps"], n_inst=1,
                                                 overrides={key: val[inst:inst + 1] for key, val in self.ov.items()})
                assert int(np.max(st)) == 0
                vr, ir = vr[0], ir[0]   # [steps+1][nn], [steps+1][n_elem]; errors relative to each series' maximum
                worst = max(worst, float(np.max(np.abs(v[:, :, j] - vr) / np.maximum(np.max(np.abs(vr), axis=0, keepdims=True), 1e-30))),
                            float(np.max(np.abs(ie[:, :, j] - ir) / np.maximum(np.max(np.abs(ir), axis=0, keepdims=True), 1e-30))))
            assert worst <= 1e-6, "bench results differ from the oracle: %.3e (%s)" % (worst, wl["name"])
        return worst

    # -- end to end through the host-buffer C ABI ---------------------------------------
    def prepare_e2e(self):
        native, eng, wl, table = self.native, self.eng, self.wl, self.table
        if wl["kind"] == "ac":
            per_point = 16 * (table.nvar + table.n_ac_elem)
            if self.sweep is None:
                self.e2e_points = max(1, min(self.P, HOST_RESULT_BYTES // per_point))
                self.e2e_freqs = np.ascontiguousarray(self.freqs[: self.e2e_points])
            else:
                self.e2e_points, self.e2e_freqs = self.P, self.freqs
            F = int(self.e2e_freqs.shape[0])
            self.h_freqs, p0 = native.pinned_empty(eng.lib, (F,), np.float64)
            self.h_freqs[:] = self.e2e_freqs
            self.h_x, p1 = native.pinned_empty(eng.lib, (table.nvar, self.e2e_points), np.complex128)
            self.h_i, p2 = native.pinned_empty(eng.lib, (table.n_ac_elem, self.e2e_points), np.complex128)
            self.h_s, p3 = native.pinned_empty(eng.lib, (self.e2e_points,), np.int32)
            self.pins += [p0, p1, p2, p3]
            self.e2e_units = self.e2e_points
            self.e2e_flags = self.flags | (native.FLAG_JIT if self.P >= native_jit_min_points() else 0)
        else:
            S1 = wl["steps"] + 1
            per_inst = 8 * S1 * (table.n_nodes + table.n_elem)
            self.e2e_inst = max(1, min(self.nb, HOST_RESULT_BYTES // per_inst))
            from spicey_b200.packing import make_sweep
            self.e2e_sweep = self.sweep if self.e2e_inst == self.nb else make_sweep(
                table, self.e2e_inst, {k: v[: self.e2e_inst] for k, v in self.ov.items()})
            self.h_v, p1 = native.pinned_empty(eng.lib, (S1, table.n_nodes, self.e2e_inst), np.float64)
            self.h_i, p2 = native.pinned_empty(eng.lib, (S1, table.n_elem, self.e2e_inst), np.float64)
            self.pins += [p1, p2]
            self.e2e_units = self.e2e_inst * S1

    def step_e2e(self, want_currents=True):
        wl = self.wl
        if wl["kind"] == "ac":
            self.eng.ac_solve(self.table, self.h_freqs, sweep=self.sweep, out=(self.h_x, self.h_i if want_currents else None, self.h_s),
                              flags=self.e2e_flags, want_currents=want_currents)
            return int(self.h_s.max())
        r = self.eng.tran_solve(self.table, wl["dt"], wl["steps"], vsrc=wl["vsrc"], vsrc_mask=wl["mask"], sweep=self.e2e_sweep,
                                state0=np.ascontiguousarray(self.state0[:, : self.e2e_inst]), out=(self.h_v, self.h_i if want_currents else None),
                                waves=self.table.waves, want_currents=want_currents)
        return int(r["status"].max())

    def release(self):
        for p in self.pins:
            self.eng.lib.spicey_host_free(p)
        self.pins = []
        for name in ("d_x", "d_i", "d_v", "d_s", "d_freqs", "d_var", "d_vsrc", "d_st0"):
            if hasattr(self, name):
                delattr(self, name)
        self.torch.cuda.empty_cache()


def native_jit_min_points():
    return 200000   # spicey_native.cu: kJitMinPoints (the host entry point sets SPICEY_FLAG_JIT from the whole call's size)


def roofline_of(leg, wl, world, kern_ms, peak_hbm, peak_src, fp64_peak):
    """Roofline of the dominant kernel of one leg, from the per-rank kernel time (max over ranks) and per-rank units."""
    native = leg.native
    units = leg.units
    if wl["kind"] == "tran":
        ach = units * wl["bytes_per_unit"] / (kern_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak_hbm, "unit": "GB/s", "frac": ach / peak_hbm,
                "traffic": None, "peak_source": peak_src, "bytes_per_solve": wl["bytes_per_unit"]}
    else:
        ach_f = units * wl["flops_per_unit"] / (kern_ms * 1e-3) / 1e12
        dense = {"bound": "fp64", "achieved": ach_f, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_f / fp64_peak,
                 "flops_per_solve_dense": wl["flops_per_unit"],
                 "peak_source": "own DFMA-loop microbenchmark in this run (MEASURED_PEAKS.json has no FP64 figure)"}
        ach_b = units * wl["bytes_per_unit"] / (kern_ms * 1e-3) / 1e9
        hbm = {"bound": "hbm", "achieved": ach_b, "peak": peak_hbm, "unit": "GB/s", "frac": ach_b / peak_hbm,
               "peak_source": peak_src, "bytes_per_solve": wl["bytes_per_unit"]}
        sparse_tiers = (native.TIER_SPARSE, native.TIER_SPARSE_JIT, native.TIER_SPARSE_WARP, native.TIER_BAND)
        if leg.tier in sparse_tiers:
            ex = units * leg.cfma * 8 / (kern_ms * 1e-3) / 1e12
            executed = {"bound": "fp64", "complex_fma_per_solve": leg.cfma, "achieved": ex, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": ex / fp64_peak}
            # which roofline binds: the structure-exploiting tiers execute 1e3 - 8e4 complex FMAs per solve instead of the
            # dense 1e5 - 6e6; for the ladder (cfg2) the FP64 pipe cannot bind and HBM does, for the mesh (cfg4: 78,592
            # complex FMAs against 15,896 result bytes per solve) the FP64 pipe does
            if executed["frac"] >= hbm["frac"]:
                roof = dict(executed, traffic=None, hbm=hbm, dense_fp64_equivalent=dense,
                            note="banded / sparse LU verified per point: bound by the FP64 pipe on the flops it executes; "
                                 "dense_fp64_equivalent is SURVEY 8(d)'s dense flop figure per solve (structurally zero work is never executed)")
            else:
                roof = dict(hbm, traffic=None, executed_fp64=executed, dense_fp64_equivalent=dense,
                            note="sparse static-pivot LU program verified per point: HBM-bound on SURVEY 8(d)'s algorithmic bytes; "
                                 "dense_fp64_equivalent exceeds 1 because structurally zero work is never executed")
        else:
            roof = dict(dense, traffic=None, hbm=hbm)
    tr = os.path.join(ROOT, "profiles", "traffic_%s.json" % wl["name"])
    if os.path.exists(tr):   # DRAM bytes of one ncu --set full capture: only quoted for the kernel tier it was taken on
        try:
            cap = json.load(open(tr))
            if int(cap.get("tier", -1)) == int(leg.tier):
                roof["traffic"] = cap.get("dram_bytes_per_unit", 0) * units
                roof["traffic_capture"] = {k: cap.get(k) for k in ("capture", "kernel", "units_in_capture", "dram_bytes_per_unit")}
        except Exception:
            pass
    return roof


def run_native(args):
    import torch
    import torch.distributed as dist
    from spicey_b200 import native

    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU port)")
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner does, whatever
    # NCCL_DEBUG_FILE says) is sent to stderr at the file-descriptor level, the line itself goes to the saved fd
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa(local) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = native.Engine([local])
    stream = torch.cuda.current_stream()
    peaks, peak_src = read_peaks()
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    try:
        dev_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local, dev_uuid)
    fp64_peak = eng.fp64_peak_gflops() / 1e3  # TFLOP/s, own DFMA-loop microbenchmark

    def time_leg(wl, steps, warmup, e2e_steps, sample_clocks, sustain_s):
        """value / roofline / e2e of one workload, every rank on its shard.  Returns the leg's dict on every rank."""
        leg = Leg(eng, wl, rank, world, dev, stream, dense=args.dense)
        for _ in range(max(1, warmup)):
            leg.step_resident()
        barrier()
        leg.note_stats()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sample_clocks and rank == 0:
            sampler.start()
        barrier()
        e0.record(stream)
        for k in range(steps):
            ev[k][0].record(stream)
            leg.step_resident()
            ev[k][1].record(stream)
        e1.record(stream)
        barrier()
        total_ms = e0.elapsed_time(e1)
        kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        # sustained region: the same step repeated until sustain_s seconds have passed (clock samples need more than
        # the few milliseconds K steps take); reported beside the K-step value, never instead of it
        sus = None
        if sustain_s > 0:
            reps = int(max(1, min(2000, np.ceil(sustain_s / max(1e-6, kern_ms * 1e-3)))))
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            s0.record(stream)
            for _ in range(reps):
                leg.step_resident()
            s1.record(stream)
            barrier()
            sus = (reps, s0.elapsed_time(s1))
        if sample_clocks and rank == 0:
            sampler.pause()
        n_samples = len(sampler.samples)
        worst = leg.verify()
        total_ms, kern_ms, worst = allmax([total_ms, kern_ms, worst])
        out = {"workload": wl["label"], "tier": leg.tier, "units_per_step": wl["units"],
               "value": wl["units"] * steps / (total_ms * 1e-3), "unit": "solves/s", "steps": steps, "ms_per_step": total_ms / steps,
               "kernel_ms_per_step": kern_ms, "chunks_per_step": len(leg.chunks()),
               "fallback_solves_last_step": leg.fallback, "rank0_range": [leg.lo, leg.hi],
               "checked_against_oracle": {"max_rel_err": worst, "tolerance": 1e-9 if wl["kind"] == "ac" else 1e-6,
                                          "sample": "4 units of every chunk of every rank's range, plus status == 0 everywhere"},
               "roofline": roofline_of(leg, wl, world, kern_ms, peak_hbm, peak_src, fp64_peak),
               "gpu_launches": leg.launches_per_step * steps}
        if sus is not None:
            (sus_ms,) = allmax([sus[1]])
            out["sustained"] = {"steps": sus[0], "seconds": sus_ms * 1e-3, "value": wl["units"] * sus[0] / (sus_ms * 1e-3)}
        # ---- end-to-end through the host-buffer C ABI ----
        if e2e_steps > 0:
            leg.prepare_e2e()
            assert leg.step_e2e() == 0
            barrier()
            if sample_clocks and rank == 0 and sampler.nvml is not None:
                sampler.start()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                assert leg.step_e2e() == 0
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            if sample_clocks and rank == 0:
                sampler.pause()
            es = eng.stats()
            # the same without element currents: the lazy result objects compute Y (v1 - v2) on access, exactly
            # as the reference does (simulateAC.ts:94-126), from the node voltages alone
            assert leg.step_e2e(want_currents=False) == 0
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                assert leg.step_e2e(want_currents=False) == 0
            torch.cuda.synchronize()
            lazy_s = time.perf_counter() - t0
            ls = eng.stats()
            if world > 1:
                dist.barrier()
            e2e_s, lazy_s = allmax([e2e_s, lazy_s])
            tot_units, = allmax([float(leg.e2e_units)])   # shards differ by at most one unit
            tot_units = tot_units * world
            out["e2e"] = {"value": tot_units * e2e_steps / e2e_s, "unit": "solves/s",
                          "h2d_bytes_per_step": int(es["h2d_bytes"]), "d2h_bytes_per_step": int(es["d2h_bytes"]),
                          "steps": e2e_steps, "kernel_ms_per_step": es["kernel_ms"],
                          "pcie_gbs_per_gpu": (int(es["h2d_bytes"]) + int(es["d2h_bytes"])) * e2e_steps / e2e_s / 1e9,
                          "units_per_step": int(tot_units),
                          "sample": "the whole job" if int(tot_units) >= wl["units"] - world else
                                    "the first %d units of every rank's shard (pinned host result buffers are capped at %d GiB per rank)" % (
                                        leg.e2e_units, HOST_RESULT_BYTES >> 30),
                          "node_voltages_only": {"value": tot_units * e2e_steps / lazy_s, "d2h_bytes_per_step": int(ls["d2h_bytes"]),
                                                 "note": "element currents left to the lazy result objects (computed on access from the node voltages, as simulateAC.ts:94-126 does)"},
                          "note": "host buffers pinned; bound by the PCIe link when pcie_gbs_per_gpu is near the link rate"}
        out["n_clock_samples"] = n_samples
        leg.release()
        return out

    wl = load_workload(args.workload, world, args.scaling, args.points, args.instances, getattr(args, "device_waves", False))
    main = time_leg(wl, args.steps, args.warmup, max(1, min(args.steps, 3)), True, 0.5)
    secondary = {}
    if args.workload == "cfg2" and not args.no_secondary and not args.points and not args.dense:
        for name in ("cfg3", "cfg5", "cfg4", "dense64"):
            w2 = load_workload(name, world, None)
            k2 = 2 if name in ("cfg4", "dense64") else 5
            try:
                secondary[name] = time_leg(w2, k2, 1, 2 if name not in ("cfg4", "dense64") else 1, False, 0.0)
                secondary[name]["scaling"] = w2["scaling"]
            except torch.cuda.OutOfMemoryError as exc:   # a secondary leg must not take the headline line down with it
                secondary[name] = {"workload": w2["label"], "error": "out of device memory: %s" % str(exc)[:200]}
                torch.cuda.empty_cache()
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        if clocks is not None:
            clocks["samples_in_timed_region"] = main.pop("n_clock_samples")
        for v in secondary.values():
            v.pop("n_clock_samples", None)
        # the CPU port is timed beside the GPU at N = 1 only (the contract's cpu_baseline); larger N reuse that line
        cpu_v, cores, sample = cpu_rate(wl, target_s=12.0) if world == 1 else (None, None, "measured at N=1 only")
        cfg = config_of(wl, world)
        line = {
            "metric": "batched MNA solves/sec", "value": main["value"], "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "c128" if wl["kind"] == "ac" else "f64", "data": "synthetic",
            "config": cfg,
            "details": {"units_per_step_whole_job": wl["units"], "rank0_range": main["rank0_range"], "tier": main["tier"],
                        "chunks_per_step": main["chunks_per_step"],
                        "result_layout": "series-major x[Nvar][ld], ielem[nAc][ld], ld = chunk rounded up to 32 points" if wl["kind"] == "ac" else "v[step][node][inst]",
                        "fallback_solves_last_step": main["fallback_solves_last_step"],
                        "l2": "no flush: every step writes its results (GBs per rank), far more than the 126 MB L2",
                        "checked_against_oracle": main["checked_against_oracle"]},
            "kernel_ms_per_step": main["kernel_ms_per_step"],
            "sustained": main.get("sustained"),
            "roofline": main["roofline"],
            "cpu_baseline": {"value": cpu_v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": main["e2e"],
            "gpu_launches": int(main["gpu_launches"]),
            "clocks": clocks,
        }
        if wl["kind"] == "tran":
            line["details"]["sources"] = "PULSE evaluated on the device from its parameters" if wl.get("device_waves") \
                else "pre-sampled row [nV][steps+1], one load per step"
        line["e2e"]["host_numa"] = numa_note
        if secondary:
            line["secondary"] = secondary
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg2mc", "cfg3", "cfg4", "cfg5", "dense64"])
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: the batch grows with the number of ranks (default for cfg2); strong: BASELINE's batch is split (default for the rest)")
    ap.add_argument("--no-secondary", action="store_true", help="default line only: skip the cfg3 / cfg4 / cfg5 legs")
    ap.add_argument("--dense", action="store_true", help="AC: force the dense pivoted-LU kernel (no sparse program)")
    ap.add_argument("--points", type=int, default=None, help="AC: subsample to this many frequency points")
    ap.add_argument("--instances", type=int, default=None, help="TRAN: number of instances")
    ap.add_argument("--device-waves", action="store_true",
                    help="TRAN: evaluate the PULSE source on the device from its parameters instead of a pre-sampled row")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
