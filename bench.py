#!/usr/bin/env python
"""bench.py -- WWZ tau x f cells/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload of the timed region (config.workload): BASELINE.json configs[1] -- synthetic unevenly sampled AR(1) series
n=2,000 (seed 0), ntau=1,000, nfreq=500 (log grid).  One STEP = wwz (c=1/(8 pi^2), all six per-cell outputs) +
wwz_psd (c=1e-3, transform fused with the wwa2psd tau-reduction) = 2 x 500,000 cells.

N > 1 (torchrun, one rank per GPU): weak scaling by tau-block -- the grid has 1,000*N tau rows, rank r computes rows
[1000 r, 1000 (r+1)); there is no data-path exchange inside the transform, one grouped NCCL all-gather assembles the
six output planes on every GPU (side stream, under the second transform) and wwz_psd all-reduces the [2, nf] partial
sums of wwa2psd.

`value`   : whole-job cells/s with inputs resident in HBM, device time (CUDA events, max over ranks).
`e2e`     : same metric through the public API on host NumPy arrays: H2D of the inputs and D2H of every output inside
            the timed region, every step (N > 1: every rank's tau block over its own PCIe link into one page-locked
            shared-memory buffer read by rank 0).
`roofline`: the kernel of the timed region (wwz_tile_kernel<8, true, 168>, the tau-recurrence warp-tile kernel):
            executed FP64-pipe instructions / launch time against the FP64 issue rate, live from a device counter;
            the dense direct kernel (50 algorithmic flop per (cell, point)) is a separate labelled field.
`signif_test`, `config5`, `ensemble`: the other configurations BASELINE.json names, at every N: the 10,000-surrogate
            significance test of config 3 (surrogates sharded over the ranks), config 5 sharded by tau-block with the
            NCCL all-gather, config 4 through the member-batch entry point.
`parity`  : results of the runs above checked inside this job: against the CPU oracle on sampled rows and, for N > 1,
            sharded against unsharded.
`cpu_baseline`: the oracle port (oracle/wwz_oracle.c, the reference's two-pass algorithm, OpenMP on all host cores)
            on a bounded sample of the same workload (rank 0, N = 1).
--impl reference: the CPU arm, rank 0 only: the reference's own kirchner_numba (utils/wavelet.py:872-1049) through
            tests/golden/ref_shim.py when a reference tree is present (/root/reference, baseline/_ref,
            $WWZ_REFERENCE_ROOT), else the oracle port.
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
_REAL_STDOUT = sys.stdout

C_WWZ = 1.0 / (8.0 * np.pi ** 2)
C_PSD = 1e-3
FLOP_PER_CELL_POINT = 50.0
REC_FP64_PER_WARP_POINT = 88     # FP64-pipe instructions of the recurrence kernel's loop body per (warp, point): 8 cells
WORKLOAD = "C2: uneven AR(1) n=2000 (seed 0), ntau=1000/GPU, nfreq=500, wwz(c=1/(8pi^2)) + wwz_psd(c=1e-3)"


def workload(n_gpus):
    from pyleoclim_util_b200 import synth, wavelet as WV
    n, ntau, nf, seed = synth.CONFIGS["C2"]
    t, y = synth.uneven_ar1(n, seed)
    tau = np.linspace(t.min(), t.max(), ntau * n_gpus)
    freq = synth.freq_log(t, nf)
    ys = WV.standardize(y)[0]
    omega = WV.make_omega(t, freq)
    return t, ys, tau, freq, omega


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU legs
def oracle_step(t, ys, tau, omega, nthreads=0):
    """One CPU step on the oracle port: both transforms + psd (faithful two-pass algorithm)."""
    from oracle import wwz_oracle as O
    Ne, a0, a1, a2 = O.kirchner_cells(t, ys, tau, omega, C_WWZ, 3, mode=0, nthreads=nthreads)
    Ne2, b0, b1, b2 = O.kirchner_cells(t, ys, tau, omega, C_PSD, 3, mode=0, nthreads=nthreads)
    amp = np.sqrt(b1 ** 2 + b2 ** 2)
    psd = O.wwa2psd(amp, t, Ne2, Neff_threshold=3)
    return Ne, psd


def cpu_baseline(t, ys, tau, omega, target_seconds=12.0):
    from oracle import wwz_oracle as O
    cores = O.num_threads()
    rows = max(8, min(tau.size, 16 * cores))
    sub = tau[:: max(1, tau.size // rows)][:rows]
    t0 = time.perf_counter()
    oracle_step(t, ys, sub, omega)
    dt = time.perf_counter() - t0
    reps = int(max(1, min(20, target_seconds / max(dt, 1e-3) - 1)))
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle_step(t, ys, sub, omega)
    dt2 = (time.perf_counter() - t0) / reps
    cells = 2 * sub.size * omega.size
    return {"value": cells / dt2, "unit": "cells/s", "cores": cores, "kind": "port",
            "sample": "%d of %d tau rows x %d freq, both transforms (wwz + wwz_psd), %d repetitions, "
                      "oracle/wwz_oracle.c two-pass Kirchner with OpenMP" % (sub.size, tau.size, omega.size, reps)}


def _reference_modules():
    """The reference's own utils modules through the import shim, or None (no tree, or numba missing)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import ref_shim
        if not ref_shim.available():
            return None, "no reference tree"
        import numba  # noqa: F401  (kirchner_numba needs it)
        return ref_shim.load(), ref_shim.REF_ROOT
    except Exception as exc:            # pragma: no cover - depends on the box
        return None, "%s: %s" % (type(exc).__name__, exc)


def reference_numba(t, ys, tau, freq, omega, more_rows=192):
    """The reference's own kernel, timed beside the port: pyleoclim kirchner_numba
    (utils/wavelet.py:872-1049) through the import shim, all numba threads.  numba re-compiles the kernel closure on
    EVERY call (utils/wavelet.py:985), 8-20 s per call depending on the host, so K timed steps of it cannot fit the time
    bound of the arm: it is measured on two bounded samples of tau rows instead -- one wwz + wwz_psd pair on 8 rows and
    one on 8 + more_rows -- which separates the per-call JIT (intercept) from the compute per row (slope), and the
    full-workload figures are the linear extrapolation, labelled as such."""
    mods, where = _reference_modules()
    if mods is None:
        return {"available": False, "why": where}
    import numba
    W = mods["wavelet"]

    def pair(taus):
        t0 = time.perf_counter()
        W.kirchner_numba(ys, t, freq, taus, c=C_WWZ, Neff_threshold=3, nproc=1, detrend=False, sg_kwargs=None,
                         gaussianize=False, standardize=False)
        p = W.kirchner_numba(ys, t, freq, taus, c=C_PSD, Neff_threshold=3, nproc=1, detrend=False, sg_kwargs=None,
                             gaussianize=False, standardize=False)
        W.wwa2psd(p[0], t, p[2], freq=freq, Neff_threshold=3, anti_alias=False, avgs=2)
        return time.perf_counter() - t0

    small = np.ascontiguousarray(tau[:: max(1, tau.size // 8)][:8])
    r2 = min(tau.size, 8 + more_rows)
    large = np.ascontiguousarray(tau[:: max(1, tau.size // r2)][:r2])
    pair(small[:2])                                     # first import / LLVM start-up
    ta, tb = pair(small), pair(large)
    slope = max((tb - ta) / max(large.size - small.size, 1), 1e-12)        # seconds per tau row, both transforms
    jit = max(ta - small.size * slope, 0.0)
    full = jit + tau.size * slope
    cells = 2.0 * tau.size * omega.size
    return {"available": True, "kind": "reference", "from": where, "cores": int(numba.get_num_threads()),
            "sample": "one wwz + wwz_psd pair on %d tau rows (%.2f s) and one on %d (%.2f s), %d freq, n=%d" % (
                small.size, ta, large.size, tb, omega.size, t.size),
            "jit_s_per_step": jit, "compute_s_per_tau_row": slope,
            "full_workload_s_per_step_extrapolated": full,
            "cells_per_s_incl_jit_extrapolated": cells / full,
            "cells_per_s_steady_extrapolated": cells / (tau.size * slope)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t, ys, tau, freq, omega = workload(1)
    from oracle import wwz_oracle as O
    cores = O.num_threads()
    rows = max(8, min(tau.size, 8 * cores))
    sub = tau[:: max(1, tau.size // rows)][:rows]
    for _ in range(args.warmup):
        oracle_step(t, ys, sub[:8], omega)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(t, ys, sub, omega)
    dt = time.perf_counter() - t0
    value = 2 * sub.size * omega.size * args.steps / dt
    sample = "%d of %d tau rows x %d freq per step, both transforms" % (sub.size, tau.size, omega.size)
    numba_leg = None
    if not os.environ.get("WWZ_BENCH_SKIP_NUMBA"):
        try:
            numba_leg = reference_numba(t, ys, tau, freq, omega)
        except Exception as exc:            # pragma: no cover - depends on the box
            numba_leg = {"available": False, "why": "%s: %s" % (type(exc).__name__, exc)}
    line = {"impl": "reference", "metric": "WWZ tau x f cells/s", "value": value, "unit": "cells/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_numba": numba_leg,
            "note": "value = CPU restatement of pyleoclim kirchner_numba (utils/wavelet.py:985-1045) in C with OpenMP on all host "
                    "cores, K bounded steps (the conservative baseline: several times faster per core than the reference's numba "
                    "kernel); reference_numba = the reference's own kernel timed beside it on two bounded samples -- its per-call "
                    "JIT (8-20 s) rules out K timed steps of it inside the arm's time bound"}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


# ------------------------------------------------------------------------------------------------ helpers
def _relmax(a, b):
    """max |a - b| / |b| over the cells where b is finite (NumPy arrays)."""
    m = ~np.isnan(b)
    if not m.any():
        return 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a[m] - b[m]) / np.abs(b[m])
    r[a[m] == b[m]] = 0.0
    return float(np.nanmax(r)) if r.size else 0.0


def _phase_absmax(a, b):
    m = ~np.isnan(b)
    if not m.any():
        return 0.0
    d = np.abs(a[m] - b[m])
    return float(np.max(np.minimum(d, np.abs(2 * np.pi - d))))


def _oracle_rows(t, ys, tau_rows, omega, c):
    """(Neffs, a0, a1, a2, wwa, phase) of a few tau rows from the CPU oracle (two-pass Kirchner)."""
    from oracle import wwz_oracle as O
    Ne, a0, a1, a2 = O.kirchner_cells(t, ys, np.ascontiguousarray(tau_rows), omega, c, 3, mode=0)
    return Ne, a0, a1, a2, np.sqrt(a1 ** 2 + a2 ** 2), np.arctan2(a2, a1)


def _rows_vs_oracle(dev6_rows, t, ys, tau_rows, omega, c):
    """Parity of device rows [6, r, nf] (library order Neffs, a0, a1, a2, wwa, phase) against the oracle."""
    Ne, a0, a1, a2, wwa, ph = _oracle_rows(t, ys, tau_rows, omega, c)
    g = dev6_rows
    return {"cells": int(wwa.size),
            "mask_identical": bool(np.array_equal(np.isnan(g[4]), np.isnan(wwa)) and np.array_equal(np.isnan(g[0]), np.isnan(Ne))),
            "amplitude_max_rel": _relmax(g[4], wwa), "phase_max_abs": _phase_absmax(g[5], ph),
            "neff_max_rel": _relmax(g[0], Ne)}


def run_ours(args):
    import torch
    from pyleoclim_util_b200 import _lib, wavelet as WV

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    _lib.set_device(dev.index)
    L = _lib.lib()
    from pyleoclim_util_b200 import distributed as _D
    # (N > 1 only: at N = 1 the CPU baseline of this process uses all host cores)
    numa_cpus = _D.bind_to_gpu_numa_node(dev.index) if world > 1 else None   # before the first page-locked allocation

    t, ys, tau_all, freq, omega = workload(world)
    n, nf = t.size, freq.size
    nt_local = tau_all.size // world
    tau = np.ascontiguousarray(tau_all[rank * nt_local:(rank + 1) * nt_local])
    nt_total = tau_all.size
    cells_step_local = 2 * nt_local * nf

    fp64_tf, _ = _lib.measure_fp64_peak()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        tm = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item())

    def gather_planes(full, loc):
        """All six output planes of a tau-sharded grid in ONE grouped NCCL launch: plane i of `full` is
        [world, rows, nf] = the row blocks of all ranks side by side, i.e. its final place (no re-layout copy)."""
        try:
            with dist._coalescing_manager(device=dev):
                for i in range(loc.shape[0]):
                    dist.all_gather_into_tensor(full[i].view(-1), loc[i].view(-1))
        except Exception:               # (an older torch without the coalescing manager: six launches)
            for i in range(loc.shape[0]):
                dist.all_gather_into_tensor(full[i].view(-1), loc[i].view(-1))

    # ---------------------------------------------------------------- device-resident leg
    d_t, d_y, d_tau, d_om = (torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (t, ys, tau, omega))
    loc6 = torch.empty((6, nt_local, nf), dtype=torch.float64, device=dev)   # Neffs,a0,a1,a2,wwa,phase of wwz
    loc2 = torch.empty((2, nt_local, nf), dtype=torch.float64, device=dev)   # Neffs, wwa of the psd pass
    psd = torch.empty(nf, dtype=torch.float64, device=dev)
    if world > 1:
        full6 = torch.empty((6, nt_total, nf), dtype=torch.float64, device=dev)
        parts = torch.empty((2, nf), dtype=torch.float64, device=dev)
        comm = torch.cuda.Stream(device=dev)
        ev_a, ev_c = torch.cuda.Event(), torch.cuda.Event()
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)             # > 126 MB L2
    tspan = float(t.max() - t.min())

    def step_device():
        o = [loc6[i].data_ptr() for i in range(6)]
        _lib.check(L.wwzb200_kirchner_dev(d_t.data_ptr(), d_y.data_ptr(), n, d_tau.data_ptr(), nt_local,
                                          d_om.data_ptr(), nf, C_WWZ, 3, 3, 0, o[0], o[1], o[2], o[3], o[4], o[5],
                                          None, nf, stream))
        if world > 1:
            # tau-row blocks are the shards: one grouped NCCL all-gather of the six wwz outputs over NVLink on a side
            # stream, under the second transform
            ev_a.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ev_a)
                gather_planes(full6, loc6)
                ev_c.record()
        _lib.check(L.wwzb200_kirchner_dev(d_t.data_ptr(), d_y.data_ptr(), n, d_tau.data_ptr(), nt_local,
                                          d_om.data_ptr(), nf, C_PSD, 3, 3, 0, loc2[0].data_ptr(), None, None, None,
                                          loc2[1].data_ptr(), None, psd.data_ptr() if world == 1 else None, nf,
                                          stream))
        if world > 1:
            # wwz_psd returns psd[nf] only: each rank reduces its own rows to the partial column sums of wwa2psd,
            # the [2, nf] parts are all-reduced and divided (utils/wavelet.py:1275-1278)
            _lib.check(L.wwzb200_wwa2psd_parts_dev(loc2[1].data_ptr(), loc2[0].data_ptr(), nt_local, nf, tspan, n, 3,
                                                   parts.data_ptr(), stream))
            dist.all_reduce(parts)
            torch.div(parts[0], parts[1], out=psd)
            torch.cuda.current_stream().wait_event(ev_c)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    sampler = ClockSampler(dev.index)
    sampler.start()
    time.sleep(0.15)
    _lib.profile(True)
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                  # L2 flush between timed iterations (outside the events)
        ev[i][0].record()
        step_device()
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = _lib.launch_count() - launches0
    cell_ms, cell_launches, cell_points = _lib.profile_read()
    walked = _lib.profile_walked()     # (warp, point) pairs the recurrence kernel evaluated in the timed region
    _lib.profile(False)
    dev_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    value = world * cells_step_local * args.steps / (dev_ms * 1e-3)

    parity = {"bar": "identical three-state mask; amplitude / psd <= 1e-9 relative, phase <= 1e-9 rad, Neff <= 1e-12"}
    # parity of the timed path, checked in this job: sampled rows of this run's wwz output against the CPU oracle
    if rank == 0:
        pick = np.linspace(0, nt_local - 1, 12).astype(int)
        parity["c2_rows_vs_oracle"] = _rows_vs_oracle(loc6[:, pick].cpu().numpy(), t, ys, tau[pick], omega, C_WWZ)
    if world > 1:
        # the assembled (all-gathered) grid against the same grid computed unsharded on this GPU
        ref6 = torch.empty((6, nt_total, nf), dtype=torch.float64, device=dev)
        d_tau_all = torch.from_numpy(np.ascontiguousarray(tau_all)).to(dev)
        _lib.check(L.wwzb200_kirchner_dev(d_t.data_ptr(), d_y.data_ptr(), n, d_tau_all.data_ptr(), nt_total, d_om.data_ptr(),
                                          nf, C_WWZ, 3, 3, 0, *[ref6[i].data_ptr() for i in range(6)], None, nf, stream))
        torch.cuda.synchronize()
        same_mask = bool(torch.equal(torch.isnan(full6), torch.isnan(ref6)))
        m = ~torch.isnan(ref6[4])
        amp_err = float(((full6[4][m] - ref6[4][m]).abs() / ref6[4][m].abs()).max()) if bool(m.any()) else 0.0
        ne_ok = ~torch.isnan(ref6[0])
        ne_err = float(((full6[0][ne_ok] - ref6[0][ne_ok]).abs() / ref6[0][ne_ok].abs()).max())
        worst = max_over_ranks(amp_err)
        all_same = max_over_ranks(0.0 if same_mask else 1.0) == 0.0
        parity["c2_sharded_vs_unsharded"] = {"cells": int(nt_total * nf), "mask_identical": all_same,
                                             "amplitude_max_rel": worst, "neff_max_rel": max_over_ranks(ne_err),
                                             "what": "the grid assembled by the NCCL all-gather on every rank against the "
                                                     "whole grid transformed on that rank's GPU alone (max over ranks)"}
        del ref6, d_tau_all

    # roofline leg: the same two transforms with WWZB200_DENSE (every (cell, point) pair evaluated, nothing
    # skipped), so that algorithmic flops / kernel time is a statement about the FP64 pipe
    def step_dense():
        o = [loc6[i].data_ptr() for i in range(6)]
        for cc in (C_WWZ, C_PSD):
            _lib.check(L.wwzb200_kirchner_dev(d_t.data_ptr(), d_y.data_ptr(), n, d_tau.data_ptr(), nt_local,
                                              d_om.data_ptr(), nf, cc, 3, 3, _lib.DENSE, o[0], o[1], o[2], o[3], o[4],
                                              o[5], None, nf, stream))
    step_dense()
    torch.cuda.synchronize()
    _lib.profile(True)
    for _ in range(max(2, min(args.steps, 10))):
        flush.zero_()
        step_dense()
    torch.cuda.synchronize()
    dense_ms, dense_launches, dense_points = _lib.profile_read()
    _lib.profile(False)

    # ---------------------------------------------------------------- end-to-end leg (host buffers)
    h_in = [_lib.pinned_empty(v.shape) for v in (t, ys, tau, omega)]
    for dst, src in zip(h_in, (t, ys, tau, omega)):
        dst[...] = src
    h_out = [_lib.pinned_empty((nt_local, nf)) for _ in range(6)]
    h_psd = _lib.pinned_empty((nf,))
    h_ne2 = _lib.pinned_empty((nt_local, nf))

    def step_e2e():
        _lib.check(L.wwzb200_kirchner(_lib.ptr(h_in[0]), _lib.ptr(h_in[1]), n, _lib.ptr(h_in[2]), nt_local,
                                      _lib.ptr(h_in[3]), nf, C_WWZ, 3, 0, *[_lib.ptr(o) for o in h_out]))
        _lib.check(L.wwzb200_kirchner_psd(_lib.ptr(h_in[0]), _lib.ptr(h_in[1]), n, _lib.ptr(h_in[2]), nt_local,
                                          _lib.ptr(h_in[3]), nf, C_PSD, 3, 3, 0, _lib.ptr(h_psd), _lib.ptr(h_ne2),
                                          None, None, None, None, None))

    for _ in range(3):
        step_e2e()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    cabi_s = max_over_ranks(time.perf_counter() - e0)
    h2d = 2 * 8 * (2 * n + nt_local + nf)
    d2h = 8 * (6 * nt_local * nf + nf)

    # the public Python API: what a user (or Pyleoclim through register()) calls.  Host NumPy arrays in, NumPy result
    # arrays out (page-locked, pooled); prepare_wwz / standardize / make_omega / make_coi on the host, H2D of the
    # inputs and D2H of every output inside the timed region, every step.
    def step_api():
        if world == 1:
            r = WV.wwz(ys, t, tau=tau, freq=freq)
            p = WV.wwz_psd(ys, t, tau=tau, freq=freq)
        else:       # the sharded public API: prepare on the full grid, tau-block per rank, results to rank 0
            from pyleoclim_util_b200 import distributed as D
            r = D.wwz_tau_sharded(ys, t, tau=tau_all, freq=freq, host="root")
            p = D.wwz_psd_tau_sharded(ys, t, tau=tau_all, freq=freq)
        return (r.amplitude[0, 0] if r.amplitude is not None else 0.0) + p.psd[0]

    for _ in range(3):
        step_api()
    barrier()
    a0 = time.perf_counter()
    for _ in range(args.steps):
        step_api()
    barrier()
    api_s = max_over_ranks(time.perf_counter() - a0)
    e2e_value = world * cells_step_local * args.steps / api_s
    if world > 1:
        from pyleoclim_util_b200 import distributed as D
        r = D.wwz_tau_sharded(ys, t, tau=tau_all, freq=freq, host="root")
        if rank == 0:       # what rank 0 received through the shared page-locked buffer, against the device grid
            g = full6.cpu().numpy()     # (the API call z-scores the already standardised input once more: equal to rounding)
            parity["c2_e2e_root_vs_device"] = {
                "mask_identical": bool(np.array_equal(np.isnan(r.amplitude), np.isnan(g[4])) and np.array_equal(np.isnan(r.Neffs), np.isnan(g[0]))),
                "amplitude_max_rel": _relmax(r.amplitude, g[4]), "neff_max_rel": _relmax(r.Neffs, g[0]),
                "phase_max_abs": _phase_absmax(r.phase, g[5]),
                "what": "rank 0's host result of wwz_tau_sharded(host='root') (every rank's tau block DMA'd over its own PCIe "
                        "link into one shared page-locked buffer) against the all-gathered device grid"}
        del r

    # ---------------------------------------------------------------- config 3: the 10,000-surrogate significance test
    # n=5000, default 50 x 501 grid; surrogates generated, z-scored, transformed and reduced to the 95% quantile on the
    # device, host to host.  N > 1: the surrogates are sharded over the ranks (strong scaling).
    extras = {}
    from pyleoclim_util_b200 import batch as B, synth, distributed as D
    t3, y3, tau3, freq3 = synth.config_grid("C3")
    S3 = 10000

    def signif_run(number):
        if world == 1:
            return B.wwz_signif(y3, t3, tau=tau3, freq=freq3, number=number, qs=[0.95], seed=1)
        return D.wwz_signif_sharded(y3, t3, tau=tau3, freq=freq3, number=number, qs=[0.95], seed=1)

    signif_run(256)
    signif_run(S3)                     # workspace (series tiles, stored weights) allocated at full size
    s_times = []
    res3 = None
    for rep in range(3):
        barrier()
        if rep == 2:
            _lib.profile(True)
        s0 = time.perf_counter()
        res3 = signif_run(S3)
        barrier()
        s_times.append(max_over_ranks(time.perf_counter() - s0))
    sh_ms, sh_n, sh_cp = _lib.profile_read()
    _lib.profile(False)
    s1 = min(s_times)
    dense3 = None
    if world == 1:
        # the same test with every (cell, point, surrogate) triple evaluated (WWZB200_DENSE: no point windows):
        # algorithmic flops / kernel time is then a statement about the FP64 (tensor) pipe
        _lib.profile(True)
        B.wwz_signif(y3, t3, tau=tau3, freq=freq3, number=S3, qs=[0.95], seed=1, flags=_lib.DENSE)
        torch.cuda.synchronize()
        d_ms, d_n, d_cp = _lib.profile_read()
        _lib.profile(False)
        dense3 = {"kernel_ms": d_ms, "contraction_tflops": 6.0 * d_cp / (d_ms * 1e-3) / 1e12,
                  "frac_of_fp64_peak": 6.0 * d_cp / (d_ms * 1e-3) / 1e12 / fp64_tf}
    signif = {
        "workload": "C3: n=5000, 50 tau x %d f, %d AR(1) surrogates, generated, z-scored, transformed and reduced to the "
                    "95%% quantile on the device, host to host%s"
                    % (freq3.size, S3, "" if world == 1 else "; surrogates sharded over the ranks, all-to-all of the "
                       "cell-major amplitudes over NVLink, quantiles from the receive buffer, all-gather of the tiles"),
        "surrogates_per_s": S3 / s1, "seconds": s1, "seconds_all_runs": s_times, "n_gpus": world, "scaling": "strong",
        "kernel": "wwz_tile_weights_kernel + wwz_shared_mma_kernel<64, stored weights> (mma.sync m8n8k4 f64, SASS "
                  "DMMA.8x8x4): weights once per (cell, point), contracted against 64 surrogates per tile",
        "shared_axis_kernel_ms_rank0": sh_ms,
        "contraction_tflops_rank0": 6.0 * sh_cp / (sh_ms * 1e-3) / 1e12 if sh_ms else None,
        "contraction_tflops_note": "algorithmic rate over all (cell, point, surrogate) triples, including the points "
                                   "the tile windows skip (weights below 2^-100 of a cell's largest weight)",
        "dense": dense3, "algorithmic": "6 flop per (cell, point, surrogate) (SURVEY.md 8d)"}
    # parity: eight full-grid surrogates of the same Philox streams through the shared-axis DMMA path against the
    # oracle (rank 0), and for N > 1 the sharded quantiles against the single-GPU test run on rank 0 alone
    if rank == 0:
        from oracle import wwz_oracle as O
        ys3c, ts3c, fr3c, ta3c = WV.prepare_wwz(y3, t3, tau=tau3, freq=freq3)
        tau_est = float(B.tau_estimation(ys3c, ts3c))
        Y8 = B.ar1_surrogates(ts3c, tau_est, sigma=float(np.std(ys3c)), mu=float(np.mean(ys3c)), number=8, seed=1)
        got = B.wwz_shared_axis(Y8, ts3c, ta3c, fr3c)
        worst, same = 0.0, True
        for s in range(8):
            wo, _, neo, _ = O.kirchner(Y8[s], ts3c, fr3c, ta3c, standardize=True)
            worst = max(worst, _relmax(got.amplitude[s], wo))
            same = same and bool(np.array_equal(np.isnan(got.amplitude[s]), np.isnan(wo)))
        parity["c3_surrogates_vs_oracle"] = {"surrogates": 8, "cells_each": int(ta3c.size * fr3c.size),
                                             "mask_identical": same, "amplitude_max_rel": worst,
                                             "what": "the first 8 surrogates of the test (same generator, same seed) on the "
                                                     "full 50 x %d grid through the shared-axis DMMA path against the oracle" % fr3c.size}
        if world > 1:
            one = B.wwz_signif(y3, t3, tau=tau3, freq=freq3, number=S3, qs=[0.95], seed=1)
            parity["c3_sharded_vs_single_gpu"] = {
                "mask_identical": bool(np.array_equal(np.isnan(res3.amplitude_qs), np.isnan(one.amplitude_qs))),
                "quantile_max_rel": _relmax(res3.amplitude_qs, one.amplitude_qs),
                "neff_max_rel": _relmax(res3.Neffs, one.Neffs),
                "what": "95% levels of the surrogate-sharded test against wwzb200_ar1_signif on rank 0's GPU alone"}
    barrier()

    # ---------------------------------------------------------------- config 4: 1,000 age-ensemble members
    # (n = 1500 each, own time axis, default 50 x ~151 grid), host to host through batch.wwz_ragged -> one
    # wwzb200_kirchner_members call (every kernel once per group of members, groups' D2H under the next group's kernels).
    # N > 1: members are independent -- each rank takes 1000 members of its own (weak scaling, no exchange).
    axes, y4 = synth.age_ensemble(1500, 1000)
    members = []
    for tm in axes:
        ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y4, tm)
        members.append((ys_c, ts_c, ta_c, fr_c))
    res4 = B.wwz_ragged(members)       # warm-up at full size: the page-locked result buffers come from a pool
    del res4
    m_times = []
    for rep in range(3):
        barrier()
        m0 = time.perf_counter()
        res4 = B.wwz_ragged(members)
        m_times.append(max_over_ranks(time.perf_counter() - m0))
        cells4 = sum(r.amplitude.size for r in res4)
        if rep < 2:
            del res4
    m1 = min(m_times)
    ensemble = {"workload": "C4: 1000 age-ensemble members per GPU, n=1500, default grid per member (50 tau x ~151 f), "
                            "standardised, all six outputs, host to host through batch.wwz_ragged (wwzb200_kirchner_members)",
                "members_per_s": world * len(members) / m1, "cells_per_s": world * cells4 / m1, "seconds": m1,
                "seconds_all_runs": m_times, "n_gpus": world, "scaling": "weak"}
    if rank == 0:
        from oracle import wwz_oracle as O
        worst, same = 0.0, True
        for m in (0, 333, 999):
            ys_c, ts_c, ta_c, fr_c = members[m]
            wo, _, neo, _ = O.kirchner(ys_c, ts_c, fr_c, ta_c, standardize=True)
            worst = max(worst, _relmax(res4[m].amplitude, wo))
            same = same and bool(np.array_equal(np.isnan(res4[m].amplitude), np.isnan(wo)))
        parity["c4_members_vs_oracle"] = {"members": 3, "mask_identical": same, "amplitude_max_rel": worst}
    del res4

    # ---------------------------------------------------------------- config 5: n=100,000, 4096 tau x 2047 f
    # tau-blocks sharded over the ranks (strong scaling), device resident, one NCCL all-gather of the six planes
    t5, y5, tau5, freq5 = synth.config_grid("C5")
    ys5 = WV.standardize(y5)[0]
    om5 = WV.make_omega(t5, freq5)
    nt5 = tau5.size
    blk = (nt5 + world - 1) // world
    a5, b5 = min(rank * blk, nt5), min((rank + 1) * blk, nt5)
    d5 = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (t5, ys5, tau5[a5:b5], om5)]
    o5 = torch.zeros((6, blk, freq5.size), dtype=torch.float64, device=dev)
    f5 = torch.empty((6, blk * world, freq5.size), dtype=torch.float64, device=dev) if world > 1 else o5

    def run5(flags, gather=True):
        if b5 > a5:
            _lib.check(L.wwzb200_kirchner_dev(d5[0].data_ptr(), d5[1].data_ptr(), t5.size, d5[2].data_ptr(), b5 - a5,
                                              d5[3].data_ptr(), freq5.size, C_WWZ, 3, 3, flags,
                                              *[o5[i].data_ptr() for i in range(6)], None, freq5.size, stream))
        if world > 1 and gather:
            gather_planes(f5, o5)

    c5 = {}
    for name, fl in (("default", 0), ("dense", _lib.DENSE)):
        run5(fl)
        barrier()
        _lib.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run5(fl)
        e1.record()
        barrier()
        k_ms, k_n, k_cp = _lib.profile_read()
        _lib.profile(False)
        tot = max_over_ranks(e0.elapsed_time(e1))
        c5[name] = {"ms": tot, "cells_per_s": nt5 * freq5.size / (tot * 1e-3), "cell_kernel_ms_rank0": k_ms,
                    "cell_kernel_launches_rank0": k_n,
                    "cell_kernel_algorithmic_tflops_rank0": FLOP_PER_CELL_POINT * k_cp / (k_ms * 1e-3) / 1e12 if k_ms else None}
    c5["dense"]["frac_of_fp64_peak_rank0"] = c5["dense"]["cell_kernel_algorithmic_tflops_rank0"] / fp64_tf
    c5["dense"]["frac_of_fp64_peak_whole_job"] = (FLOP_PER_CELL_POINT * t5.size * nt5 * freq5.size
                                                  / (c5["dense"]["ms"] * 1e-3) / 1e12 / (fp64_tf * world))
    c5["n_gpus"] = world
    c5["scaling"] = "strong"
    c5["workload"] = ("C5: n=100000, 4096 tau x %d f, one transform, device resident; %s"
                      % (freq5.size, "single GPU" if world == 1 else
                         "tau-blocks of %d rows per rank, one grouped NCCL all-gather of the six output planes into the "
                         "assembled grid on every GPU, inside the timed region" % blk))
    run5(0)                     # leave the default path's result in o5 / f5 for the parity check
    barrier()
    if rank == 0:
        # stratified sample of the assembled grid against the oracle: first / last / middle tau rows, every frequency
        # batch (low frequencies walk the whole series, high ones a few points), 16 rows x 128 columns
        rows = np.unique(np.concatenate([[0, 1, nt5 - 2, nt5 - 1], np.linspace(2, nt5 - 3, 12).astype(int)]))
        cols = np.unique(np.concatenate([np.arange(0, freq5.size, 17), [freq5.size - 2, freq5.size - 1]]))
        dev_rows = f5[:, torch.from_numpy(rows).to(dev)][:, :, torch.from_numpy(cols).to(dev)].cpu().numpy()
        parity["c5_sample_vs_oracle"] = _rows_vs_oracle(dev_rows, t5, ys5, tau5[rows], om5[cols], C_WWZ)
        parity["c5_sample_vs_oracle"]["what"] = ("%d tau rows (both ends and the interior) x %d frequencies (every 17th and the "
                                                 "last two) of the %s grid against the oracle" %
                                                 (rows.size, cols.size, "all-gathered" if world > 1 else "device"))
    del d5, o5, f5
    barrier()

    clocks = sampler.stop()
    sm_count = _lib.sm_count()
    sm_hz = (clocks.get("sm_mhz") or 0) * 1e6

    # ---------------------------------------------------------------- roofline + CPU baseline
    dense_tf = FLOP_PER_CELL_POINT * dense_points / (dense_ms * 1e-3) / 1e12 if dense_ms > 0 else None
    windowed_tf = FLOP_PER_CELL_POINT * cell_points / (cell_ms * 1e-3) / 1e12 if cell_ms > 0 else None
    # executed FP64-pipe instructions of the timed kernel / its time / the pipe's issue rate (one warp instruction per
    # 2 clocks per SM sub-partition = 2 per clock per SM), at the SM clock sampled during the run
    util = (walked * REC_FP64_PER_WARP_POINT / (cell_ms * 1e-3) / (sm_count * 2 * sm_hz)) if cell_ms > 0 and sm_hz else None
    executed_tf = (walked * REC_FP64_PER_WARP_POINT * 32 * 2 / (cell_ms * 1e-3) / 1e12) if cell_ms > 0 else None
    roofline = {
        "bound": "fp64", "kernel": "wwz_tile_kernel<8, true, 168> (tau-recurrence warp-tile kernel: the kernel of the timed region)",
        "achieved": executed_tf, "peak": fp64_tf, "unit": "TFLOP/s", "frac": util,
        "frac_definition": "executed FP64-pipe warp instructions (88 per (warp, point) of 8 tau-adjacent cells x 32 lanes, "
                           "counted in the SASS loop body; (warp, point) pairs from a device counter) per launch / CUDA-event "
                           "launch time / (2 warp instructions per clock per SM at the SM clock sampled in the run); `achieved` "
                           "counts every executed FP64 instruction as one FMA (2 flop) per lane, `peak` is the measured DFMA rate",
        "avg_launch_ms": cell_ms / max(cell_launches, 1), "launches_timed": cell_launches,
        "kernel_share_of_step": cell_ms / (sum(a.elapsed_time(b) for a, b in ev)) if cell_launches else None,
        "warp_points_evaluated": walked,
        "fraction_of_dense_warp_points": walked / (cell_points / 8 / 32) if cell_points else None,
        "algorithmic_tflops_incl_skipped_points": windowed_tf,
        "algorithmic_note": "50 flop x every (cell, point) pair of the grid / time: above the FP64 peak because the kernel shares "
                            "one exp2 pair between 8 tau-adjacent cells and never touches points whose weight is below 2^-100 "
                            "of a cell's largest; both are parity-tested cell for cell",
        "traffic": 46.0e6, "traffic_note": "dram__bytes_read + write per launch from ncu --set full (profiles/r02_rec_kernel_ncu.txt): "
                   "33.8 MB read (the basis rows once) + 12.5 MB written (the partial-sum slabs) against 24 MB of algorithmic output; "
                   "FP64-pipe bound, not HBM bound",
        "dense_direct_kernel": {
            "kernel": "wwz_cells_kernel<2, 8> with WWZB200_DENSE: every (cell, point) pair evaluated with its own exp2",
            "algorithmic_tflops": dense_tf, "frac_of_fp64_peak": (dense_tf / fp64_tf) if dense_tf else None,
            "avg_launch_ms": dense_ms / max(dense_launches, 1), "launches": dense_launches,
            "executed_fp64_instr_per_cell_point": 19,
            "hardware_utilisation": (19.0 * 2 / FLOP_PER_CELL_POINT) * (dense_tf / fp64_tf) if dense_tf else None,
            "note": "algorithmic 50 flop per (cell, point) (SURVEY.md 8d: exp counted as 26 flop) / CUDA-event time of launches "
                    "after the timed region; the kernel executes 19 FP64 instructions per (cell, point), so the FP64 pipe itself "
                    "is busy `hardware_utilisation` of the time (ncu: 0.71, stall math_pipe_throttle: DFMAs with three distinct "
                    "register operands issue at 2/3 rate)"},
        "algorithmic": "50 flop per (cell, point) x n=%d x %d cells per launch (SURVEY.md 8d)" % (n, nt_local * nf),
        "peak_source": "of measured: register-resident DFMA-chain microbenchmark on this device "
                       "(wwzb200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; nominal 37.2"}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cpu = cpu_baseline(t, ys, tau, omega) if world == 1 else None
    if cpu is not None and not os.environ.get("WWZ_BENCH_SKIP_NUMBA"):
        try:
            cpu["reference_numba"] = reference_numba(t, ys, tau, freq, omega)
        except Exception as exc:            # pragma: no cover - depends on the box
            cpu["reference_numba"] = {"available": False, "why": "%s: %s" % (type(exc).__name__, exc)}
    line = {"metric": "WWZ tau x f cells/s", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": n, "ntau_per_gpu": nt_local, "nfreq": nf,
                       "cells_per_step_per_gpu": cells_step_local, "partition": "tau-block per GPU; one grouped NCCL all-gather of "
                       "the six wwz output planes into the assembled grid on every GPU (side stream, under the second "
                       "transform); wwz_psd: all-reduce of the [2, nf] partial sums of wwa2psd"
                       if world > 1 else "single GPU", "l2": "256 MiB buffer written between timed iterations",
                       "window": "per tile of cells only the points whose weight reaches 2^-100 of a cell's largest weight are walked",
                       "kernels": "tau-recurrence warp-tile kernel on rows with kappa*dtau <= 8 (all rows here), direct kernel otherwise"},
            "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "bytes_note": "per rank: the four input vectors of both calls up, the rank's six output planes and psd back",
                    "ms_per_step": 1e3 * api_s / args.steps,
                    "path": ("pyleoclim_util_b200.wavelet.wwz + .wwz_psd (the reference's wwz / wwz_psd signatures) on host "
                             "NumPy arrays; results land in pooled page-locked NumPy arrays; the wwz call streams finished blocks of "
                             "tau rows to the host while the kernel is still computing (fused epilogue + progress flags)") if world == 1 else
                            ("pyleoclim_util_b200.distributed.wwz_tau_sharded(host='root') + .wwz_psd_tau_sharded on host "
                             "NumPy arrays, every rank: prepare, tau block on its GPU through the host entry point, finished "
                             "blocks of tau rows streamed over the rank's own PCIe link into one page-locked shared-memory "
                             "buffer that rank 0 returns; psd all-reduced"),
                    "c_abi_ms_per_step": 1e3 * cabi_s / args.steps,
                    "c_abi_path": "wwzb200_kirchner + wwzb200_kirchner_psd on pre-allocated pinned host buffers"},
            "host_binding": ("rank pinned to the %d CPUs next to its GPU (sysfs local_cpulist)" % len(numa_cpus)) if numa_cpus
                            else "no CPU binding",
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "signif_test": signif, "config5": c5, "ensemble": ensemble, "parity": parity,
            "wall_ms_per_step_incl_l2_flush": 1e3 * wall / args.steps, "extras": extras,
            "hbm_gbs_measured": peaks.get("hbm_gbs")}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _protect_stdout():
    """The contract is ONE JSON line on stdout: route everything else that writes to fd 1 (NCCL's version banner,
    library chatter) to stderr and hand back a file object on the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    global _REAL_STDOUT
    _REAL_STDOUT = _protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
