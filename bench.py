#!/usr/bin/env python
"""bench.py -- WWZ tau x f cells/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] -- synthetic unevenly sampled AR(1) series
n=2,000 (seed 0), ntau=1,000, nfreq=500 (log grid).  One STEP = wwz (c=1/(8 pi^2), all six per-cell
outputs) + wwz_psd (c=1e-3, transform fused with the wwa2psd tau-reduction) = 2 x 500,000 cells.

N > 1 (torchrun, one rank per GPU): weak scaling by tau-block -- the grid has 1,000*N tau rows, rank r
computes rows [1000 r, 1000 (r+1)); there is no data-path exchange inside the transform, one NCCL
all-gather assembles the sharded outputs and wwa2psd runs on the gathered scalogram.

`value`   : whole-job cells/s with inputs resident in HBM, device time (CUDA events, max over ranks).
`e2e`     : same metric through the C-ABI host-pointer entry points (wwzb200_kirchner +
            wwzb200_kirchner_psd) with host buffers: H2D of the inputs and D2H of every output inside
            the timed region, every step.
`roofline`: the dominant kernel (wwz_cells_kernel) -- algorithmic 50 flop per (cell, point)
             / its CUDA-event time, against the FP64 FMA peak measured on this device.
`cpu_baseline`: the oracle port (oracle/wwz_oracle.c, the reference's two-pass algorithm, OpenMP on
            all host cores) on a bounded sample of the same workload.
--impl reference: the CPU arm (same oracle port; the reference's own kernel is Python/numba +
            Fortran and cannot travel to the GPU box), rank 0 only.
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
    cells = 2 * sub.size * omega.size * args.steps
    value = cells / dt
    sample = "%d of %d tau rows x %d freq per step, both transforms" % (sub.size, tau.size, omega.size)
    line = {"impl": "reference", "metric": "WWZ tau x f cells/s", "value": value, "unit": "cells/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference arm = CPU restatement of pyleoclim kirchner_numba (utils/wavelet.py:985-1045) in C with "
                    "OpenMP on all host cores; the reference's numba path adds a ~7-8 s JIT per call (BASELINE.md)"}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


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

    t, ys, tau_all, freq, omega = workload(world)
    n, nf = t.size, freq.size
    nt_local = tau_all.size // world
    tau = np.ascontiguousarray(tau_all[rank * nt_local:(rank + 1) * nt_local])
    nt_total = tau_all.size
    cells_step_local = 2 * nt_local * nf

    fp64_tf, _ = _lib.measure_fp64_peak()

    # ---------------------------------------------------------------- device-resident leg
    d_t, d_y, d_tau, d_om = (torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (t, ys, tau, omega))
    loc6 = torch.empty((6, nt_local, nf), dtype=torch.float64, device=dev)   # Neffs,a0,a1,a2,wwa,phase of wwz
    loc2 = torch.empty((2, nt_local, nf), dtype=torch.float64, device=dev)   # Neffs, wwa of the psd pass
    psd = torch.empty(nf, dtype=torch.float64, device=dev)
    if world > 1:
        # the assembled outputs: plane i of full6 is [world, nt_local, nf] = the tau-row blocks of all ranks side by
        # side, so each output plane is gathered straight into its final place (no re-layout copy)
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
            # tau-row blocks are the shards: NCCL all-gathers of the six wwz outputs over NVLink on a side stream,
            # overlapped with the second transform
            ev_a.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ev_a)
                for i in range(6):
                    dist.all_gather_into_tensor(full6[i].view(-1), loc6[i].view(-1))
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

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    if dist is not None:
        tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dev_ms = float(tmax.item())
    value = world * cells_step_local * args.steps / (dev_ms * 1e-3)

    # ---------------------------------------------------------------- end-to-end leg (host buffers, C ABI)
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
    e2e_s = time.perf_counter() - e0
    if dist is not None:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    cabi_s = e2e_s
    # through the public API, rank 0: the inputs of both calls go up, the assembled result of both calls comes back
    # (N > 1: the grid is assembled in the HBM of every GPU by the all-gather; rank 0 copies it to the host)
    h2d = 2 * 8 * (2 * n + nt_local + nf)
    d2h = 8 * (6 * (nt_local if world == 1 else nt_total) * nf + nf)

    # the public Python API: what a user (or Pyleoclim through register()) calls.  Host NumPy arrays in, fresh
    # NumPy result arrays out (page-locked, pooled); prepare_wwz / standardize / make_omega / make_coi on the host,
    # H2D of the inputs and D2H of every output inside the timed region, every step.
    def step_api():
        if world == 1:
            r = WV.wwz(ys, t, tau=tau, freq=freq)
            p = WV.wwz_psd(ys, t, tau=tau, freq=freq)
        else:       # the sharded public API: prepare on the full grid, tau-block per rank, NCCL all-gather / all-reduce
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
    api_s = time.perf_counter() - a0
    if dist is not None:
        tmax = torch.tensor([api_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        api_s = float(tmax.item())
    e2e_value = world * cells_step_local * args.steps / api_s

    # ---------------------------------------------------------------- the other half of the metric: surrogate tests/s
    # config 3: n=5000, default 50 x 501 grid; surrogates generated, z-scored, transformed and reduced to the 95%
    # quantile on the device, host to host.  N > 1: the grid is sharded by frequency block over the ranks.
    extras = {}
    from pyleoclim_util_b200 import batch as B, synth, distributed as D
    t3, y3, tau3, freq3 = synth.config_grid("C3")
    S3 = 10000          # BASELINE.json configs[2]: a 10,000-surrogate test, strong scaling over the GPUs

    def signif_run(number):
        if world == 1:
            return B.wwz_signif(y3, t3, tau=tau3, freq=freq3, number=number, qs=[0.95], seed=1)
        return D.wwz_signif_sharded(y3, t3, tau=tau3, freq=freq3, number=number, qs=[0.95], seed=1)

    signif_run(256)
    signif_run(S3)                     # workspace (series tiles, stored weights) allocated at full size
    barrier()
    _lib.profile(True)
    s0 = time.perf_counter()
    signif_run(S3)
    barrier()
    s1 = time.perf_counter() - s0
    sh_ms, sh_n, sh_cp = _lib.profile_read()
    _lib.profile(False)
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
    if dist is not None:
        tmax = torch.tensor([s1], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        s1 = float(tmax.item())
    extras["signif_test"] = {
        "workload": "C3: n=5000, 50 tau x %d f, %d AR(1) surrogates, generated, z-scored, transformed and "
                    "reduced to the 95%% quantile on the device (wwzb200_ar1_signif), host to host%s"
                    % (freq3.size, S3, "" if world == 1 else "; frequency blocks sharded over the ranks"),
        "surrogates_per_s": S3 / s1, "seconds": s1, "n_gpus": world, "scaling": "strong",
        "kernel": "wwz_tile_weights_kernel + wwz_shared_mma_kernel<64, stored weights> (mma.sync m8n8k4 f64, SASS "
                  "DMMA.8x8x4): weights once per (cell, point), contracted against 64 surrogates per tile",
        "shared_axis_kernel_ms_rank0": sh_ms,
        "contraction_tflops_rank0": 6.0 * sh_cp / (sh_ms * 1e-3) / 1e12 if sh_ms else None,
        "contraction_tflops_note": "algorithmic rate over all (cell, point, surrogate) triples, including the points "
                                   "the tile windows skip (weights below 2^-100 of a cell's largest weight)",
        "dense": dense3,
        "algorithmic": "6 flop per (cell, point, surrogate) (SURVEY.md 8d)"}

    # config 4: 1,000 age-ensemble members (n = 1500 each, own time axis, default 50 x 151 grid), one ragged batch,
    # host to host (wwzb200_kirchner_ragged: packed inputs up once, members over 8 streams, packed outputs back once)
    if world == 1:
        axes, y4 = synth.age_ensemble(1500, 1000)
        members = []
        for tm in axes:
            ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y4, tm)
            members.append((ys_c, ts_c, ta_c, fr_c))
        res4 = B.wwz_ragged(members)       # warm-up at full size: the page-locked result buffers come from a pool
        del res4
        torch.cuda.synchronize()
        m0 = time.perf_counter()
        res4 = B.wwz_ragged(members)
        m1 = time.perf_counter() - m0
        cells4 = sum(r.amplitude.size for r in res4)
        extras["ensemble"] = {"workload": "C4: 1000 age-ensemble members, n=1500, default grid per member (50 tau x ~151 f), "
                                          "standardised, all six outputs, host to host through batch.wwz_ragged",
                              "members_per_s": len(members) / m1, "cells_per_s": cells4 / m1, "seconds": m1}
        del res4

    # config 5 (n=100,000, 4096 tau x 2047 f = 8.4e11 cell-points), device resident, rank 0 at N = 1 only:
    # once dense (the FP64-pipe statement BASELINE.json asks for) and once with the default path
    if world == 1:
        t5, y5, tau5, freq5 = synth.config_grid("C5")
        d5 = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in
              (t5, WV.standardize(y5)[0], tau5, WV.make_omega(t5, freq5))]
        o5 = torch.empty((6, tau5.size, freq5.size), dtype=torch.float64, device=dev)

        def run5(flags):
            _lib.check(L.wwzb200_kirchner_dev(d5[0].data_ptr(), d5[1].data_ptr(), t5.size, d5[2].data_ptr(), tau5.size,
                                              d5[3].data_ptr(), freq5.size, C_WWZ, 3, 3, flags,
                                              *[o5[i].data_ptr() for i in range(6)], None, freq5.size, stream))

        c5 = {}
        for name, fl in (("default", 0), ("dense", _lib.DENSE)):
            run5(fl)
            torch.cuda.synchronize()
            _lib.profile(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run5(fl)
            e1.record()
            torch.cuda.synchronize()
            k_ms, k_n, k_cp = _lib.profile_read()
            _lib.profile(False)
            tot = e0.elapsed_time(e1)
            c5[name] = {"ms": tot, "cells_per_s": tau5.size * freq5.size / (tot * 1e-3), "cell_kernel_ms": k_ms,
                        "cell_kernel_launches": k_n,
                        "cell_kernel_algorithmic_tflops": FLOP_PER_CELL_POINT * k_cp / (k_ms * 1e-3) / 1e12}
        c5["dense"]["frac_of_fp64_peak"] = c5["dense"]["cell_kernel_algorithmic_tflops"] / fp64_tf
        c5["workload"] = "C5: n=100000, 4096 tau x %d f, one transform, device resident (basis rows in %d frequency batches)" \
                         % (freq5.size, c5["dense"]["cell_kernel_launches"])
        extras["config5"] = c5
        del d5, o5

    clocks = sampler.stop()
    sm_count = _lib.sm_count()
    sm_hz = (clocks.get("sm_mhz") or 0) * 1e6

    # ---------------------------------------------------------------- roofline + CPU baseline
    achieved_tf = FLOP_PER_CELL_POINT * dense_points / (dense_ms * 1e-3) / 1e12 if dense_ms > 0 else None
    windowed_tf = FLOP_PER_CELL_POINT * cell_points / (cell_ms * 1e-3) / 1e12 if cell_ms > 0 else None
    roofline = {"bound": "fp64", "kernel": "wwz_cells_kernel", "achieved": achieved_tf, "peak": fp64_tf,
                "unit": "TFLOP/s", "frac": (achieved_tf / fp64_tf) if achieved_tf else None,
                "traffic": 63.4e6, "traffic_note": "dram__bytes_read+write per launch from ncu --set full "
                "(profiles/r01_cells_kernel_ncu.txt): 33.0 MB read (32 MB of basis rows) + 30.4 MB written (partial-sum "
                "slabs) against 24 MB of algorithmic output (48 B per cell); the kernel is FP64-pipe bound, not HBM bound",
                "mode": "WWZB200_DENSE launches (nothing skipped), %d launches after the timed region" % dense_launches,
                "avg_launch_ms": dense_ms / max(dense_launches, 1),
                "timed_region": {"launches": cell_launches, "avg_launch_ms": cell_ms / max(cell_launches, 1),
                                 "algorithmic_tflops_incl_skipped_points": windowed_tf,
                                 "kernel_share_of_step": cell_ms / dev_ms if dev_ms else None,
                                 "kernel": "wwz_rec_kernel (tau-recurrence, point windows) + wwz_cells_kernel side by side",
                                 "warp_points_evaluated": walked,
                                 "fraction_of_dense_warp_points": walked / (cell_points / 8 / 32) if cell_points else None,
                                 "executed_fp64_instr_per_warp_point": 88,
                                 "fp64_pipe_utilisation": (walked * 88 / (cell_ms * 1e-3) / (sm_count * 2 * sm_hz))
                                 if cell_ms > 0 and sm_hz else None,
                                 "fp64_pipe_note": "executed FP64 instructions (88 per (warp, point) of 8 tau-adjacent cells, "
                                 "counted in the SASS loop body) / kernel time / (2 warp instructions per clock per SM at the "
                                 "sampled SM clock): the live counterpart of ncu's sm__inst_executed_pipe_fp64"},
                "executed_fp64_instr_per_cell_point": 19,
                "note": "frac is for the direct kernel (one exp2 per (cell, point)); the timed region runs the tau-recurrence "
                        "kernel, which shares one exp2 pair between 8 tau-adjacent cells (about 11 FP64 instructions per "
                        "(cell, point)), so its algorithmic rate can exceed the FP64 peak",
                "algorithmic": "50 flop per (cell, point) x n=%d x %d cells per launch (SURVEY.md 8d)" % (n, nt_local * nf),
                "peak_source": "of measured: register-resident DFMA-chain microbenchmark on this device "
                               "(wwzb200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; nominal 37.2"}
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cpu = cpu_baseline(t, ys, tau_all[:nt_local] if world > 1 else tau, omega) if world == 1 else None
    line = {"metric": "WWZ tau x f cells/s", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": n, "ntau_per_gpu": nt_local, "nfreq": nf,
                       "cells_per_step_per_gpu": cells_step_local, "partition": "tau-block per GPU; NCCL all-gather of every wwz output "
                       "plane into the assembled grid on every GPU (side stream, overlapped with the second transform); "
                       "wwz_psd: all-reduce of the [2, nf] partial sums of wwa2psd"
                       if world > 1 else "single GPU", "l2": "256 MiB buffer written between timed iterations",
                       "window": "per tile of cells only the points whose weight reaches 2^-100 of a cell's largest weight are walked",
                       "kernels": "tau-recurrence kernel on rows with kappa*dtau <= 8 (all rows here), direct kernel otherwise"},
            "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * api_s / args.steps,
                    "path": ("pyleoclim_util_b200.wavelet.wwz + .wwz_psd (the reference's wwz / wwz_psd signatures) on host "
                             "NumPy arrays; results land in pooled page-locked NumPy arrays") if world == 1 else
                            ("pyleoclim_util_b200.distributed.wwz_tau_sharded(host='root') + .wwz_psd_tau_sharded on host "
                             "NumPy arrays, every rank: prepare, tau block on its GPU, NCCL all-gather / all-reduce; the "
                             "assembled [6, ntau, nf] grid is copied to page-locked host memory by rank 0, psd by every rank"),
                    "c_abi_ms_per_step": 1e3 * cabi_s / args.steps,
                    "c_abi_path": "wwzb200_kirchner + wwzb200_kirchner_psd on pre-allocated pinned host buffers"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
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
