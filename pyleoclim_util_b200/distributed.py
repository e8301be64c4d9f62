"""One process per GPU: tau-block sharding of the (tau, f) grid with a single all-gather of the
output shards.  Cells are independent, so there is no data-path exchange inside
the transform; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.

The surrogate significance test is sharded the same way -- every rank draws the same counter-based
surrogate streams (same seed) and evaluates all S surrogates on its own tau rows, so the per-cell
quantiles need no exchange either; only the [nq, rows, nf] quantile tiles are gathered.
"""
import ctypes

import numpy as np

from . import wavelet as _wv


def row_blocks(n, world):
    """Contiguous balanced partition of range(n) into `world` blocks: [(start, stop), ...]."""
    base, extra = divmod(int(n), int(world))
    out, start = [], 0
    for r in range(world):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


TILE_ROWS = 32      # frequency rows of one CTA tile of the shared-axis kernel (kMmaRows, csrc/wwz_mma.cu)


def balanced_freq_groups(ts_cut, tau, freq, c, world, cut_bits=100.0):
    """Frequency rows for each rank of a frequency-sharded significance test: [index arrays].

    The shared-axis kernel walks, per tile of TILE_ROWS consecutive frequencies, only the points inside the
    Gaussian support of the tile (weights above 2^-cut_bits), so a low-frequency row costs up to the whole
    series and a high-frequency row a few points: equal-count contiguous blocks would leave the first rank with
    most of the work.  Rows are kept in groups of TILE_ROWS consecutive frequencies (compact windows), the groups
    are dealt to the ranks by modelled cost (longest-processing-time first), and each rank's groups are put back
    in ascending order (a short last group stays last, so tiles never straddle two groups)."""
    freq = np.asarray(freq, dtype=float)
    nf = freq.size
    span = float(np.max(ts_cut) - np.min(ts_cut)) or 1.0
    kappa4 = 4.0 * 2.0 * np.pi * np.abs(freq) * np.sqrt(c * np.log2(np.e))
    with np.errstate(divide="ignore", invalid="ignore"):
        dcut = 4.0 * np.sqrt(cut_bits) / kappa4
    dtau = 3.0 * float(np.median(np.abs(np.diff(tau)))) if np.size(tau) > 1 else 0.0
    frac = np.clip(np.nan_to_num((2.0 * dcut + dtau) / span, nan=0.0, posinf=1.0), 0.0, 1.0)
    groups = [np.arange(g, min(g + TILE_ROWS, nf)) for g in range(0, nf, TILE_ROWS)]
    # a tile costs its widest window for all of its row slots, plus a fixed part per tile
    cost = np.array([TILE_ROWS * (frac[g].max() + 0.02) for g in groups])
    load = np.zeros(world)
    mine = [[] for _ in range(world)]
    for gi in np.argsort(-cost, kind="stable"):
        r = int(np.argmin(load))
        mine[r].append(int(gi))
        load[r] += cost[gi]
    return [np.concatenate([groups[g] for g in sorted(m)]) if m else np.zeros(0, dtype=int) for m in mine]


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun, one rank per GPU)")
    return dist


def all_gather_rows(local, counts, group=None):
    """Gather row blocks of different heights: local is [counts[rank], ...]; returns the
    concatenation [sum(counts), ...] on every rank (one all-gather on blocks padded to the tallest)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    local = np.ascontiguousarray(local, dtype=np.float64)
    assert local.shape[0] == counts[rank], "row count of this rank does not match the partition"
    tail = local.shape[1:]
    hmax = max(max(counts), 1)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    pad = torch.zeros((hmax,) + tail, dtype=torch.float64, device=device)
    if local.shape[0]:
        pad[: local.shape[0]] = torch.from_numpy(local).to(device)
    out = torch.empty((world, hmax) + tail, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    out = out.cpu().numpy()
    return np.concatenate([out[r, : counts[r]] for r in range(world)], axis=0)


def bind_to_gpu_numa_node(device_index=None):
    """Pin this process to the CPUs next to its GPU (the PCI device's `local_cpulist` in sysfs), so that page-locked
    result buffers are allocated on -- and the GPU's DMA writes go to -- the memory of the GPU's own socket.  torchrun
    does not bind its workers; with eight ranks moving results to the host at once the difference is the aggregate
    host-side bandwidth.  Returns the CPU set, or None when the topology cannot be read (or WWZB200_NO_NUMA_BIND)."""
    import os
    if os.environ.get("WWZB200_NO_NUMA_BIND"):
        return None
    try:
        import torch
        i = torch.cuda.current_device() if device_index is None else int(device_index)
        pr = torch.cuda.get_device_properties(i)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def _shm_class():
    from multiprocessing import shared_memory

    class _Shm(shared_memory.SharedMemory):
        """A mapping that lives as long as the process: result arrays and the page-lock keep pointers into it, so
        it is never closed by the garbage collector (the kernel unmaps it at exit; the creator unlinks the name)."""

        def __del__(self):
            pass

    return _Shm


class _SharedHostSegments:
    """Page-locked POSIX shared-memory segments mapped by every rank of the box (host='root' results).

    Rank 0 -- the consumer -- owns the free list: a result array handed to the caller is a view of a segment, and the
    segment returns to the free list when the array (and every view of it) has been collected, like the page-locked
    pool of `_lib`.  Every rank maps each segment once and page-locks its own mapping (`wwzb200_host_register`), so
    that its GPU copies its block of rows over its own PCIe link straight into the consumer's buffer.

    The ranks agree on the segment of a call, and rank 0 learns that every block has landed, through a small control
    segment (a mailbox of int64 words in shared memory): no collective and no device synchronisation beyond each
    rank's own stream.  Only the creation of a new segment (first call, or a larger result) is a collective."""

    CTRL_WORDS = 512            # [0] sequence number of the published pick, [1] picked segment, [8 + r] rank r's done flag

    def __init__(self):
        self.segs = []          # [dict(shm, nbytes, addr, free)]
        self.same_host = {}     # group -> bool
        self.ctrl = None        # int64 view of the control segment
        self.seq = 0

    def on_one_host(self, group):
        key = id(group)
        if key not in self.same_host:
            import socket
            dist = _dist()
            names = [None] * dist.get_world_size(group)
            dist.all_gather_object(names, socket.gethostname(), group=group)
            self.same_host[key] = len(set(names)) == 1 and dist.get_world_size(group) <= self.CTRL_WORDS - 8
        return self.same_host[key]

    def _create(self, size, group, register=True):
        """Collective: a new segment of `size` bytes mapped (and page-locked) by every rank."""
        import atexit
        import ctypes
        from multiprocessing import resource_tracker
        from . import _lib
        dist = _dist()
        rank = dist.get_rank(group)
        Shm = _shm_class()
        name = [None]
        shm = None
        if rank == 0:
            shm = Shm(create=True, size=size)
            name[0] = shm.name
            atexit.register(self._unlink, shm)
        dist.broadcast_object_list(name, src=0, group=group)
        if rank != 0:
            shm = Shm(name=name[0])
            try:                     # the creator unlinks; an attaching process must not (Python < 3.13 tracks both)
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        addr = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
        if register:
            # first touch: every rank faults in its share of the pages from the CPUs next to its GPU, so the segment is
            # spread over the sockets of the box instead of sitting on the node of whoever page-locks it first
            world = dist.get_world_size(group)
            lo, hi = rank * size // world, (rank + 1) * size // world
            ctypes.memset(addr + lo, 0, hi - lo)
            dist.barrier(group=group)
            _lib.check(_lib.lib().wwzb200_host_register(addr, size))
        dist.barrier(group=group)    # every rank has mapped the segment before anybody uses it
        return dict(shm=shm, nbytes=size, addr=addr, free=True)

    def acquire(self, nbytes, group):
        """Every rank returns its mapping (a dict) of the same segment, of at least nbytes.  Called by all ranks in
        the same order."""
        dist = _dist()
        rank = dist.get_rank(group)
        if self.ctrl is None:
            bind_to_gpu_numa_node()
            c = self._create(self.CTRL_WORDS * 8, group, register=False)
            self.ctrl_seg = c
            self.ctrl = np.ndarray((self.CTRL_WORDS,), dtype=np.int64, buffer=c["shm"].buf)
            if rank == 0:
                self.ctrl[:] = 0
            dist.barrier(group=group)
        self.seq += 1
        ctrl = self.ctrl
        if rank == 0:
            idx = -1
            for i, sg in enumerate(self.segs):
                if sg["free"] and sg["nbytes"] >= nbytes:
                    idx = i
                    break
            ctrl[1] = idx
            ctrl[0] = self.seq                       # publish
        else:
            while ctrl[0] != self.seq:
                pass
            idx = int(ctrl[1])
        if idx < 0:
            size = max(int(nbytes), 1 << 20)
            size = (size + (1 << 21) - 1) & ~((1 << 21) - 1)
            self.segs.append(self._create(size, group))
            idx = len(self.segs) - 1
        sg = self.segs[idx]
        sg["free"] = False
        return sg

    def landed(self, group):
        """This rank's copies into the segment of the current call are complete; rank 0 waits for every rank."""
        dist = _dist()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.ctrl[8 + rank] = self.seq
        if rank == 0:
            flags = self.ctrl[8:8 + world]
            while not (flags == self.seq).all():
                pass

    @staticmethod
    def _unlink(shm):
        try:
            shm.unlink()
        except Exception:
            pass

    def view(self, sg, shape):
        """float64 array over the segment; the segment is free again once the array has been collected."""
        import weakref
        count = int(np.prod(shape, dtype=np.int64))
        arr = np.frombuffer(sg["shm"].buf, dtype=np.float64, count=count).reshape(shape)
        holder = arr.base if arr.base is not None else arr
        weakref.finalize(holder, sg.__setitem__, "free", True)
        return arr


_shared_segments = _SharedHostSegments()


def _device_blocks(pd_ys, ts_cut, tau, omega, c, Neff_threshold, flags, group, want, to_host=True, gather=True,
                   root_only=False):
    """NCCL fast path: this rank's tau block is transformed into a device buffer (C-ABI *_dev entry point on
    torch's current stream), the row blocks are all-gathered on the device and only the assembled
    [len(want), nt, nf] result crosses PCIe, into page-locked host memory.  `want` indexes the library's output
    order (0 Neffs, 1 a0, 2 a1, 3 a2, 4 wwa, 5 phase)."""
    import torch
    from . import _lib
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    nt, nf = tau.size, omega.size
    blocks = row_blocks(nt, world)
    counts = [b - a for a, b in blocks]
    a, b = blocks[rank]
    hmax = max(max(counts), 1)
    if gather and root_only and tuple(want) == (4, 5, 0, 1, 2, 3) and _shared_segments.on_one_host(group):
        # host='root', all six outputs: the consumer is rank 0 and there is no all-gather.  Every rank runs the HOST entry
        # point on its tau block with its six output pointers inside one page-locked shared-memory buffer (library
        # order, one pitch): finished tau blocks cross the rank's own PCIe link while the kernel is still computing
        # (streamed output, wwz_capi.cu) and rank 0 reads the assembled planes in place.
        L = _lib.lib()
        sg = _shared_segments.acquire(6 * nt * nf * 8, group)
        if b > a:
            f64 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
            h_ts, h_ys, h_tau, h_om = f64(ts_cut), f64(pd_ys), f64(tau[a:b]), f64(omega)
            outs = [ctypes.cast(sg["addr"] + (i * nt + a) * nf * 8, _lib._dp) for i in range(6)]
            _lib.check(L.wwzb200_kirchner(_lib.ptr(h_ts), _lib.ptr(h_ys), h_ts.size, _lib.ptr(h_tau), b - a, _lib.ptr(h_om),
                                          nf, float(c), int(Neff_threshold), int(flags), *outs))
        _shared_segments.landed(group)                  # rank 0 waits until every block has landed (shared-memory flags)
        if rank == 0:
            v = _shared_segments.view(sg, (6, nt, nf))
            return [v[i] for i in want], None
        sg["free"] = True
        return None, None
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev, non_blocking=True)
    d_ts, d_ys, d_om = up(ts_cut), up(pd_ys), up(omega)
    d_tau = up(tau[a:b]) if b > a else torch.zeros(1, dtype=torch.float64, device=dev)
    loc = torch.zeros((6, hmax, nf), dtype=torch.float64, device=dev)
    if b > a:
        ptrs = [loc[i].data_ptr() for i in range(6)]
        _lib.check(_lib.lib().wwzb200_kirchner_dev(d_ts.data_ptr(), d_ys.data_ptr(), ts_cut.size, d_tau.data_ptr(), b - a,
                                                   d_om.data_ptr(), nf, float(c), int(Neff_threshold), 0, int(flags),
                                                   *ptrs, None, nf, torch.cuda.current_stream().cuda_stream))
    k = len(want)
    if not gather:
        return None, loc[:, : b - a]
    if root_only and _shared_segments.on_one_host(group):
        # host='root': the consumer is rank 0.  No all-gather: every rank copies its own row block of every wanted
        # plane over its own PCIe link into one page-locked shared-memory buffer that rank 0 then reads in place.
        L = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        sg = _shared_segments.acquire(k * nt * nf * 8, group)
        if b > a:
            for i, w_i in enumerate(want):
                _lib.check(L.wwzb200_copy_to_host_async(sg["addr"] + (i * nt + a) * nf * 8, loc[w_i].data_ptr(),
                                                        (b - a) * nf * 8, st))
        _lib.check(L.wwzb200_stream_synchronize(st))
        _shared_segments.landed(group)                  # rank 0 waits until every block has landed (shared-memory flags)
        if rank == 0:
            return _shared_segments.view(sg, (k, nt, nf)), None
        sg["free"] = True
        return None, None
    if len(set(counts)) == 1:
        # equal row blocks: plane i of the result is [world, hmax, nf], so every wanted output plane is gathered
        # straight into its final place (no selection copy, no re-layout)
        full = torch.empty((k, nt, nf), dtype=torch.float64, device=dev)
        for i, w_i in enumerate(want):
            dist.all_gather_into_tensor(full[i].view(-1), loc[w_i].view(-1), group=group)
    else:
        sel = loc[list(want)].contiguous()
        g = torch.empty((world, k, hmax, nf), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(g.view(-1), sel.view(-1), group=group)
        full = torch.cat([g[r, :, : counts[r]] for r in range(world)], dim=1)
    host = None
    if to_host:
        host = _lib.output_empty((k, nt, nf))
        full = full.contiguous()
        # one D2H of the assembled result, straight into the page-locked result buffer
        _lib.check(_lib.lib().wwzb200_copy_to_host(host.ctypes.data, full.data_ptr(), host.nbytes,
                                                   torch.cuda.current_stream().cuda_stream))
    return host, full


def wwz_tau_sharded(ys, ts, tau=None, freq=None, freq_method="log", freq_kwargs=None, c=1 / (8 * np.pi ** 2),
                    Neff_threshold=3, Neff_coi=3, standardize=True, group=None, kernel=None, flags=0, host="all"):
    """`wavelet.wwz` with the tau rows sharded over the ranks of `group`.  Every rank returns the
    full Results (amplitude, phase, coi, freq, time, Neffs, coeff, scale), like the reference's wwz
    (utils/wavelet.py:1294-1494).  `kernel` defaults to `wavelet.kirchner_cuda`.

    host='all' (default): the all-gather assembles the grid in the HBM of every GPU and every rank copies it to its
    host memory.  host='root' (NCCL path, ranks on one box): rank 0 is the consumer; there is no all-gather -- every
    rank copies its own tau block over its own PCIe link into one page-locked shared-memory buffer (8 links in
    parallel instead of one carrying the whole grid) which rank 0 returns in place; the other ranks get Results whose
    per-cell arrays are None."""
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    kernel = _wv.kirchner_cuda if kernel is None else kernel
    ys_cut, ts_cut, freq, tau = _wv.prepare_wwz(ys, ts, freq=freq, freq_method=freq_method, freq_kwargs=freq_kwargs,
                                                tau=tau)
    pd_ys = _wv.preprocess(ys_cut, ts_cut, standardize=standardize)
    if kernel is _wv.kirchner_cuda and dist.get_backend(group) == "nccl":
        _wv.assertPositiveInt(Neff_threshold)
        if host not in ("all", "root"):
            raise ValueError("host must be 'all' or 'root'")
        to_host = host == "all" or rank == 0
        host_arr, _ = _device_blocks(pd_ys, ts_cut, tau, _wv.make_omega(ts_cut, freq), c, Neff_threshold, flags, group,
                                     want=(4, 5, 0, 1, 2, 3), to_host=to_host, root_only=host == "root")
        wwa, phase, Neffs, a0, a1, a2 = host_arr if host_arr is not None else (None,) * 6
        return _wv.WwzResults(amplitude=wwa, phase=phase, coi=_wv.make_coi(tau, Neff_threshold=Neff_coi), freq=freq,
                              time=tau, Neffs=Neffs, coeff=(a0, a1, a2), scale=1 / freq)
    blocks = row_blocks(tau.size, world)
    counts = [b - a for a, b in blocks]
    a, b = blocks[rank]
    kw = dict(c=c, Neff_threshold=Neff_threshold, standardize=False)
    if kernel is _wv.kirchner_cuda:
        kw["flags"] = flags
    wwa, phase, Neffs, (a0, a1, a2) = kernel(pd_ys, ts_cut, freq, tau[a:b], **kw)
    stacked = np.stack([wwa, phase, Neffs, a0, a1, a2], axis=1) if b > a else np.zeros((0, 6, freq.size))
    full = all_gather_rows(stacked, counts, group)
    wwa, phase, Neffs, a0, a1, a2 = (np.ascontiguousarray(full[:, i]) for i in range(6))
    return _wv.WwzResults(amplitude=wwa, phase=phase, coi=_wv.make_coi(tau, Neff_threshold=Neff_coi), freq=freq,
                          time=tau, Neffs=Neffs, coeff=(a0, a1, a2), scale=1 / freq)


def wwz_psd_tau_sharded(ys, ts, tau=None, freq=None, freq_method="log", freq_kwargs=None, c=1e-3, standardize=True,
                        Neff_threshold=3, group=None, kernel=None, flags=0):
    """`wavelet.wwz_psd` (utils/spectral.py:728-900) with the tau rows sharded.  The tau-reduction of
    wwa2psd (utils/wavelet.py:1275-1278) is a ratio of two column sums, so the ranks all-reduce their
    partial numerators and denominators ([2, nf]) before the divide."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    kernel = _wv.kirchner_cuda if kernel is None else kernel
    ys_cut, ts_cut, freq, tau = _wv.prepare_wwz(ys, ts, freq=freq, freq_method=freq_method, freq_kwargs=freq_kwargs,
                                                tau=tau)
    pd_ys = _wv.preprocess(ys_cut, ts_cut, standardize=standardize)
    if kernel is _wv.kirchner_cuda and dist.get_backend(group) == "nccl":
        # each rank reduces its own tau rows to the partial column sums of wwa2psd on the device
        # (wwzb200_wwa2psd_parts_dev), the [2, nf] parts are all-reduced over NVLink, then divided
        from . import _lib
        _, loc = _device_blocks(pd_ys, ts_cut, tau, _wv.make_omega(ts_cut, freq), c, 3, flags, group, want=(4, 0),
                                to_host=False, gather=False)
        parts = torch.zeros((2, freq.size), dtype=torch.float64, device=loc.device)
        if loc.shape[1]:
            wwa_loc, ne_loc = loc[4].contiguous(), loc[0].contiguous()
            _lib.check(_lib.lib().wwzb200_wwa2psd_parts_dev(wwa_loc.data_ptr(), ne_loc.data_ptr(), loc.shape[1], freq.size,
                                                            float(np.max(ts_cut) - np.min(ts_cut)), int(np.size(ts_cut)),
                                                            int(Neff_threshold), parts.data_ptr(),
                                                            torch.cuda.current_stream().cuda_stream))
        dist.all_reduce(parts, group=group)
        return _wv.PsdResults(psd=(parts[0] / parts[1]).cpu().numpy(), freq=freq)
    a, b = row_blocks(tau.size, world)[rank]
    kw = dict(c=c, Neff_threshold=3, standardize=False)
    if kernel is _wv.kirchner_cuda:
        kw["flags"] = flags
    part = np.zeros((2, freq.size))
    if b > a:
        wwa, _, Neffs, _ = kernel(pd_ys, ts_cut, freq, tau[a:b], **kw)
        power = wwa ** 2 * 0.5 * (np.max(ts_cut) - np.min(ts_cut)) / np.size(ts_cut) * Neffs   # utils/wavelet.py:1270
        Neff_diff = Neffs - Neff_threshold
        Neff_diff[Neff_diff < 0] = 0
        part[0] = np.nansum(power * Neff_diff, axis=0)
        part[1] = np.nansum(Neff_diff, axis=0)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    tot = torch.from_numpy(part).to(device)
    dist.all_reduce(tot, group=group)
    tot = tot.cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        psd = tot[0] / tot[1]
    return _wv.PsdResults(psd=psd, freq=freq)


def wwz_signif_sharded(ys, ts, tau=None, freq=None, number=200, qs=(0.95,), seed=0, c=1 / (8 * np.pi ** 2),
                       Neff_threshold=3, freq_method="log", freq_kwargs=None, group=None, signif=None, axis="auto"):
    """Scalogram significance levels sharded over the ranks.  axis='surrogates' (the default under NCCL): every rank
    transforms its block of the surrogates on the whole grid, an all-to-all regroups the amplitudes by cell block and
    the quantiles are taken there (`_signif_by_surrogates`).  Grid sharding: each rank runs the fused
    device pipeline (`batch.signif_prepared`) for all `number` surrogates on its block of tau rows
    (axis='tau') or on its share of the frequency columns (axis='freq': groups of 32 consecutive frequencies dealt
    to the ranks by modelled cost, `balanced_freq_groups`; 'auto' takes the longer axis -- the reference's default
    grid is 50 tau x n/10 frequencies).  The same seed gives every rank the same
    surrogates, so the per-cell quantiles are exchange-free; only the quantile tiles are gathered.
    Returns `batch.SignifResults`."""
    from . import batch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    signif = batch.signif_prepared if signif is None else signif
    ys_cut, ts_cut, freq, tau = _wv.prepare_wwz(ys, ts, freq=freq, freq_method=freq_method, freq_kwargs=freq_kwargs,
                                                tau=tau)
    probs = np.atleast_1d(np.asarray(qs, dtype=float))
    nq = probs.size
    if axis == "auto":
        # NCCL: shard the surrogates (balanced by construction; the all-to-all of the amplitudes rides NVLink);
        # otherwise (CPU tests, custom `signif`) shard the longer grid axis
        if signif is batch.signif_prepared and dist.get_backend(group) == "nccl" and number >= world:
            axis = "surrogates"
        else:
            axis = "freq" if freq.size >= tau.size else "tau"
    if axis == "surrogates":
        q, Neffs = _signif_by_surrogates(ys_cut, ts_cut, tau, freq, number, probs, seed, c, Neff_threshold, group)
        return batch.SignifResults(qs=probs, amplitude_qs=q, psd_qs=None, Neffs=Neffs, freq=freq, time=tau)
    if axis == "tau":
        blocks = row_blocks(tau.size, world)
        a, b = blocks[rank]
        tau_loc, freq_loc = tau[a:b], freq
    elif axis == "freq":
        # cost-balanced groups of consecutive frequencies per rank (not contiguous blocks: see balanced_freq_groups)
        rows = balanced_freq_groups(ts_cut, tau, freq, c, world)
        blocks = [(0, r.size) for r in rows]
        a, b = blocks[rank]
        tau_loc, freq_loc = tau, freq[rows[rank]]
    else:
        raise ValueError("axis must be 'surrogates', 'tau', 'freq' or 'auto'")
    counts = [y - x for x, y in blocks]
    if b > a:
        q_wwa, _, Neffs = signif(ys_cut, ts_cut, tau_loc, freq_loc, number=number, probs=probs, seed=seed, c=c,
                                 Neff_threshold=Neff_threshold)
        local = np.concatenate([q_wwa, Neffs[None]], axis=0)                 # [nq+1, nt_loc, nf_loc]
        local = np.moveaxis(local, 1 if axis == "tau" else 2, 0)             # sharded axis first
    else:
        local = np.zeros((0, nq + 1, freq.size if axis == "tau" else tau.size))
    full = all_gather_rows(np.ascontiguousarray(local), counts, group)
    if axis == "freq":                                                       # undo the dealing of the frequency rows
        ordered = np.empty_like(full)
        ordered[np.concatenate(rows)] = full
        full = ordered
    full = np.moveaxis(full, 0, 1 if axis == "tau" else 2)                   # back to [nq+1, nt, nf]
    return batch.SignifResults(qs=probs, amplitude_qs=np.ascontiguousarray(full[:nq]), psd_qs=None,
                               Neffs=np.ascontiguousarray(full[nq]), freq=freq, time=tau)


class _PeerBuffers:
    """One device buffer per rank that every other rank of the box has mapped (CUDA IPC through the C ABI's
    wwzb200_ipc_*): the destination of the fused exchange of the surrogate-sharded significance test.  Grown
    collectively; reused across calls (a call ends with a collective that follows the last read of the buffer on
    every rank's stream, so the next call's peers cannot overwrite data still in use)."""

    def __init__(self):
        self.entries = {}       # id(group) -> dict(nbytes, own, peers (ctypes array of void*))

    def get(self, nbytes, group):
        import ctypes
        import torch
        from . import _lib
        dist = _dist()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        key = id(group)
        ent = self.entries.get(key)
        if ent is not None and ent["nbytes"] >= nbytes:
            return ent
        L = _lib.lib()
        dev = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.synchronize()
        if ent is not None:                          # grow: unmap the peers, then (after everybody did) free
            for r in range(world):
                if r != rank:
                    _lib.check(L.wwzb200_ipc_close(ent["peers"][r]))
            dist.barrier(group=group)
            _lib.check(L.wwzb200_ipc_free(ent["own"]))
        size = (int(nbytes) + (1 << 21) - 1) & ~((1 << 21) - 1)
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(L.wwzb200_ipc_alloc(size, ctypes.byref(own), handle))
        mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=dev)
        allh = torch.empty((world, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh.view(-1), mine, group=group)
        allh = allh.cpu().numpy()
        peers = (ctypes.c_void_p * world)()
        for r in range(world):
            if r == rank:
                peers[r] = own.value
            else:
                p = ctypes.c_void_p()
                _lib.check(L.wwzb200_ipc_open(allh[r].tobytes(), ctypes.byref(p)))
                peers[r] = p.value
        dist.barrier(group=group)
        ent = dict(nbytes=size, own=own.value, peers=peers)
        self.entries[key] = ent
        return ent


_peer_buffers = _PeerBuffers()


def surrogate_blocks(number, world):
    """Partition of the surrogates over the ranks: blocks of ceil(number / world), the last ranks take what is left
    (so that every block but the last has the same length: the all-to-all receive buffer is then a row of equal
    segments, which `wwzb200_quantiles_segments_dev` reads in place)."""
    per = -(-int(number) // int(world))
    return per, [(min(r * per, number), min((r + 1) * per, number)) for r in range(world)]


def _signif_by_surrogates(ys_cut, ts_cut, tau, freq, number, probs, seed, c, Neff_threshold, group, method="ar1sim"):
    """Surrogates sharded over the ranks (NCCL): rank r generates rows [s0_r, s1_r) of the surrogate batch (the Philox
    stream of a series is keyed by its global index: same surrogates as the single-GPU test), transforms them as one
    shared-axis batch on the whole (tau, f) grid -- perfectly balanced, whatever the point windows -- and keeps the
    amplitudes of a cell with the rank that owns the cell's block: the contraction kernel stores them straight into
    that rank's HBM over NVLink (fused exchange, `wwzb200_shared_axis_cells_scatter_dev`; on several hosts, or with
    WWZB200_NO_PEER_SCATTER, a staged all-to-all instead), the per-cell quantiles are taken in place, and the
    [nq, cells/world] tiles are all-gathered.  The part of the transform that does not depend on the surrogates (setup, basis rows, stored weights,
    Neffs) is enqueued first and runs under the host-side AR(1) fit.  Returns (q [nq, nt, nf], Neffs [nt, nf]) as
    NumPy arrays, equal on every rank and to the single-GPU result."""
    import ctypes
    import torch
    from . import _lib, batch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream
    L = _lib.lib()
    ys_cut = np.asarray(ys_cut, dtype=float)
    flags = _lib.STANDARDIZE
    if method == "uar1":
        flags |= _lib.UAR1
    elif method != "ar1sim":
        raise ValueError("surrogate method must be 'ar1sim' or 'uar1'")
    omega = _wv.make_omega(ts_cut, freq)
    n, nt, nf, nq = ts_cut.size, tau.size, freq.size, probs.size
    ncell = nt * nf
    per, sblocks = surrogate_blocks(number, world)
    s0, s1 = sblocks[rank]
    S_loc = s1 - s0
    pad = ctypes.c_int64()
    _lib.check(L.wwzb200_series_pad(max(per, 1), ctypes.byref(pad)))
    S_pad = int(pad.value)                                   # same row stride on every rank (the all-to-all is regular)
    # one upload for the four small inputs
    packed = np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in (ts_cut, tau, omega, probs)])
    d_in = torch.from_numpy(packed).to(dev, non_blocking=True)
    d_ts, d_tau, d_om, d_probs = d_in[:n], d_in[n:n + nt], d_in[n + nt:n + nt + nf], d_in[n + nt + nf:]
    cells_per = (ncell + world - 1) // world
    d_ne = torch.empty((nt, nf), dtype=torch.float64, device=dev)
    prepared = ctypes.c_int(0)
    if S_loc > 0:
        # y-independent pass first: it runs on the device while the host fits the AR(1) model below
        _lib.check(L.wwzb200_shared_axis_prepare_dev(d_ts.data_ptr(), n, S_loc, d_tau.data_ptr(), nt, d_om.data_ptr(), nf,
                                                     float(c), int(Neff_threshold), flags, d_ne.data_ptr(),
                                                     ctypes.byref(prepared), st))
    mu = float(np.mean(ys_cut))
    if method == "ar1sim":
        sig, tau_est = float(np.std(ys_cut)), float(batch.tau_estimation(ys_cut, ts_cut))
    else:
        tau_est, sig = (float(v) for v in batch.uar1_fit(ys_cut, ts_cut))
    import os
    fused = world <= 16 and not os.environ.get("WWZB200_NO_PEER_SCATTER") and _shared_segments.on_one_host(group)
    d_q = torch.empty((nq, cells_per), dtype=torch.float64, device=dev)
    if fused:
        # FUSED EXCHANGE: every rank owns a block of cells_per cells and a buffer [cells_per][world][S_pad] that all
        # ranks have mapped; the contraction kernel stores each amplitude straight into the owner's buffer over NVLink
        # (peer memory), column slot `rank` -- the transfer rides under the contraction, tile by tile.  No staging
        # buffer and no all-to-all.
        ent = _peer_buffers.get(cells_per * world * S_pad * 8, group)
        if S_loc > 0:
            d_Y = torch.empty((S_loc, n), dtype=torch.float64, device=dev)
            _lib.check(L.wwzb200_surrogates_block_dev(d_ts.data_ptr(), n, tau_est, sig, mu, S_loc, s0,
                                                      int(seed) & 0xFFFFFFFFFFFFFFFF, flags & _lib.UAR1, d_Y.data_ptr(), st))
            fl = flags | (_lib.PREPARED if prepared.value else 0)
            _lib.check(L.wwzb200_shared_axis_cells_scatter_dev(
                d_ts.data_ptr(), d_Y.data_ptr(), n, S_loc, d_tau.data_ptr(), nt, d_om.data_ptr(), nf, float(c),
                int(Neff_threshold), fl, d_ne.data_ptr(), ent["peers"], world, cells_per, world * S_pad, rank * S_pad, st))
        # "every peer has finished writing into my buffer": a collective in stream order behind the kernels
        token = torch.zeros(1, dtype=torch.float64, device=dev)
        dist.all_reduce(token, group=group)
        _lib.check(L.wwzb200_quantiles_segments_dev(ent["own"], number, per, S_pad, world * S_pad, cells_per,
                                                    d_probs.data_ptr(), nq, d_q.data_ptr(), st))
    else:
        # cell-major amplitudes, cells padded to a multiple of world (padding rows are never read back)
        d_amp = torch.empty((cells_per * world, S_pad), dtype=torch.float64, device=dev)
        if S_loc > 0:
            own = ctypes.c_int64()
            _lib.check(L.wwzb200_series_pad(S_loc, ctypes.byref(own)))
            d_Y = torch.empty((S_loc, n), dtype=torch.float64, device=dev)
            _lib.check(L.wwzb200_surrogates_block_dev(d_ts.data_ptr(), n, tau_est, sig, mu, S_loc, s0,
                                                      int(seed) & 0xFFFFFFFFFFFFFFFF, flags & _lib.UAR1, d_Y.data_ptr(), st))
            if int(own.value) == S_pad:
                tgt = d_amp
            else:                                           # a short last block pads to fewer columns: its own buffer
                tgt = torch.empty((ncell, int(own.value)), dtype=torch.float64, device=dev)
            fl = flags | (_lib.PREPARED if prepared.value else 0)
            _lib.check(L.wwzb200_shared_axis_cells_dev(d_ts.data_ptr(), d_Y.data_ptr(), n, S_loc, d_tau.data_ptr(), nt,
                                                       d_om.data_ptr(), nf, float(c), int(Neff_threshold), fl,
                                                       d_ne.data_ptr(), tgt.data_ptr(), st))
            if tgt is not d_amp:
                d_amp[:ncell, :S_loc] = tgt[:, :S_loc]
        # every rank gets all surrogates of its block of cells: recv[r] = rank r's surrogates of my cells
        recv = torch.empty((world, cells_per, S_pad), dtype=torch.float64, device=dev)
        dist.all_to_all_single(recv.view(-1), d_amp.view(-1), group=group)
        _lib.check(L.wwzb200_quantiles_segments_dev(recv.data_ptr(), number, per, cells_per * S_pad, S_pad, cells_per,
                                                    d_probs.data_ptr(), nq, d_q.data_ptr(), st))
    qall = torch.empty((world, nq, cells_per), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(qall.view(-1), d_q.view(-1), group=group)
    q = qall.permute(1, 0, 2).reshape(nq, world * cells_per)[:, :ncell].reshape(nq, nt, nf)
    if number < world:                                      # some ranks hold no surrogate: Neffs from rank 0
        if S_loc == 0:
            d_ne.zero_()
        dist.broadcast(d_ne, src=0, group=group)
    out = torch.cat([q.reshape(-1), d_ne.reshape(-1)]).cpu().numpy()      # one device-to-host copy
    return out[:nq * ncell].reshape(nq, nt, nf), out[nq * ncell:].reshape(nt, nf)


def wwz_signif_tau_sharded(ys, ts, **kw):
    """`wwz_signif_sharded` with axis='tau'."""
    return wwz_signif_sharded(ys, ts, axis="tau", **kw)
