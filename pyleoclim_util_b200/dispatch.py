"""Registration of ``method='Kirchner_cuda'`` inside an installed Pyleoclim.

Kernel level.  The reference resolves the kernel through ``pyleoclim.utils.wavelet.get_wwz_func``
(utils/wavelet.py:2052-2092), looked up as a module global at call time (:1480), so wrapping that
one function makes ``wwz``, ``wwz_psd`` (utils/spectral.py:889), ``wwz_coherence`` (:1660-1665),
``Series.wavelet`` (core/series.py:3559) and ``Series.spectral`` (core/series.py:3386) accept
``settings={'method': 'Kirchner_cuda'}``; the method name is stored in ``wave_args`` /
``spec_args`` and re-passed by ``signif_test`` (core/scalograms.py:517, core/psds.py:280).

Class level.  ``MultipleSeries.wavelet/.spectral`` (core/multipleseries.py:1291-1534; inherited by
``SurrogateSeries`` and ``EnsembleSeries``) fan out over a ``spawn`` process pool, whose workers see
neither this patch nor a shared CUDA context.  For the CUDA method they are replaced by a
record / replay scheme: every member's own ``Series.wavelet`` / ``Series.spectral`` runs in-process
with a recording kernel (so grids, arguments and the ``Scalogram`` / ``PSD`` objects are built by the
reference's own code), then all recorded requests are issued as ONE batched launch (shared time axis
or ragged) and the results are written into the objects.  ``MultipleScalogram.quantiles`` /
``MultiplePSD.quantiles`` run on the device, and ``Scalogram.signif_test`` / ``PSD.signif_test`` take
the fully fused device pipeline (surrogates -> transform -> quantiles) when nothing but the quantiles
is asked for.
"""
import contextlib
import functools
import warnings

import numpy as np

from . import _lib
from . import wavelet as _wavelet

_state = {}
_recorder = None          # list collecting kernel requests while a batch is being recorded

CUDA_METHOD = "Kirchner_cuda"
FOSTER_METHOD = "Foster_cuda"          # the reference's method='Foster' (wwz_basic) on the same kernels
CUDA_METHODS = (CUDA_METHOD, FOSTER_METHOD)


def _fresh_seed():
    """An unseeded significance test draws fresh surrogates on every call, like the reference's use of NumPy's
    global stream (utils/tsmodel.py:66): one 64-bit seed from the OS entropy pool per call."""
    return int(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF)


# ----------------------------------------------------------------------------------- kernel level
def _make_kernel(mod, foster=False):
    pre = getattr(mod, "preprocess", None)
    mko = getattr(mod, "make_omega", None)
    flags = _lib.FOSTER if foster else 0

    def kirchner_cuda(ys, ts, freq, tau, c=1 / (8 * np.pi ** 2), Neff_threshold=3, nproc=1, detrend=False,
                      sg_kwargs=None, gaussianize=False, standardize=False):
        if _recorder is not None:
            _wavelet.assertPositiveInt(Neff_threshold)
            nt, nf = np.size(tau), np.size(freq)
            _recorder.append(dict(ys=np.asarray(ys, dtype=float), ts=np.asarray(ts, dtype=float),
                                  freq=np.asarray(freq, dtype=float), tau=np.asarray(tau, dtype=float), c=c,
                                  Neff_threshold=Neff_threshold, detrend=detrend, sg_kwargs=sg_kwargs,
                                  gaussianize=gaussianize, standardize=standardize, flags=flags))
            hole = np.full((nt, nf), np.nan)
            return hole, hole, hole, (hole, hole, hole)
        return _wavelet.kirchner_cuda(ys, ts, freq, tau, c=c, Neff_threshold=Neff_threshold, nproc=nproc,
                                      detrend=detrend, sg_kwargs=sg_kwargs, gaussianize=gaussianize,
                                      standardize=standardize, flags=flags, _preprocess=pre, _make_omega=mko)

    kirchner_cuda.__doc__ = (_wavelet.foster_cuda if foster else _wavelet.kirchner_cuda).__doc__
    if foster:
        kirchner_cuda.__name__ = kirchner_cuda.__qualname__ = "foster_cuda"
    return kirchner_cuda


@contextlib.contextmanager
def _recording():
    global _recorder
    previous, _recorder = _recorder, []
    try:
        yield _recorder
    finally:
        _recorder = previous


def _plain(req):
    return (req["detrend"] in (False, None, "none")) and not req["gaussianize"]


def _replay(requests, mod, psd_threshold=None, batch=None):
    """Run the recorded kernel requests as one batched launch.

    Returns a list of (wwa, Neffs, psd-or-None) in request order.  All requests of a
    MultipleSeries call share c / threshold / preprocessing flags (they come from one settings
    dict); series with one common time axis and grid go through the shared-axis entry point,
    everything else through the ragged one."""
    if batch is None:
        from . import batch
    if not requests:
        return []
    r0 = requests[0]
    uniform = all(r["c"] == r0["c"] and r["Neff_threshold"] == r0["Neff_threshold"]
                  and r["standardize"] == r0["standardize"] and _plain(r) == _plain(r0)
                  and r.get("flags", 0) == r0.get("flags", 0) for r in requests)
    if not uniform:
        raise ValueError("members of one batched WWZ call must share c, Neff_threshold and preprocessing flags")
    pre = getattr(mod, "preprocess", _wavelet.preprocess)
    kflags = int(r0.get("flags", 0))
    # Foster's solve is not linear in the shared basis sums: its batches take the ragged (per-member) entry point
    same_axis = len(requests) > 1 and _plain(r0) and not (kflags & _lib.FOSTER) and all(
        r["ts"].shape == r0["ts"].shape and np.array_equal(r["ts"], r0["ts"]) and np.array_equal(r["tau"], r0["tau"])
        and np.array_equal(r["freq"], r0["freq"]) for r in requests[1:])
    out = []
    if same_axis:
        Y = np.stack([r["ys"] for r in requests])
        res = batch.wwz_shared_axis(Y, r0["ts"], r0["tau"], r0["freq"], c=r0["c"], Neff_threshold=r0["Neff_threshold"],
                                    standardize=bool(r0["standardize"]), outputs=("amplitude",),
                                    psd_Neff_threshold=psd_threshold)
        for s in range(len(requests)):
            out.append((res.amplitude[s], res.Neffs.copy(), None if res.psd is None else res.psd[s]))
        return out
    members = []
    for r in requests:
        pd_ys = pre(r["ys"], r["ts"], detrend=r["detrend"], sg_kwargs=r["sg_kwargs"], gaussianize=r["gaussianize"],
                    standardize=r["standardize"])
        members.append((pd_ys, r["ts"], r["tau"], r["freq"]))
    res = batch.wwz_ragged(members, c=r0["c"], Neff_threshold=r0["Neff_threshold"], standardize=False,
                           psd_Neff_threshold=psd_threshold, **({"flags": kflags} if kflags else {}))
    if psd_threshold is not None:
        res, psds = res
    else:
        psds = [None] * len(members)
    for r, p in zip(res, psds):
        out.append((r.amplitude, r.Neffs, p))
    return out


def _wants_cuda(method, settings):
    return method == "wwz" and isinstance(settings, dict) and settings.get("method") in CUDA_METHODS


# ----------------------------------------------------------------------------------- class level
def _patch_core(pyleo, wavemod):
    core = pyleo.core
    MS = core.multipleseries.MultipleSeries
    MScal = core.scalograms.MultipleScalogram
    Scal = core.scalograms.Scalogram
    MPSD = core.psds.MultiplePSD
    PSD = core.psds.PSD
    Coh = core.coherences.Coherence
    orig = dict(ms_wavelet=MS.wavelet, ms_spectral=MS.spectral, mscal_q=MScal.quantiles, mpsd_q=MPSD.quantiles,
                scal_signif=Scal.signif_test, psd_signif=PSD.signif_test, coh_signif=Coh.signif_test)

    @functools.wraps(orig["ms_wavelet"])
    def ms_wavelet(self, method="cwt", settings={}, freq=None, freq_kwargs=None, verbose=False, mute_pbar=False):
        if not _wants_cuda(method, settings):
            return orig["ms_wavelet"](self, method=method, settings=settings, freq=freq, freq_kwargs=freq_kwargs,
                                      verbose=verbose, mute_pbar=mute_pbar)
        settings = settings.copy()
        with _recording() as rec:
            scal_list = [s.wavelet(method=method, settings=settings, freq=freq, freq_kwargs=freq_kwargs,
                                   verbose=verbose) for s in self.series_list]
        if len(rec) != len(scal_list):
            raise RuntimeError("expected one WWZ kernel request per series")
        for scal, (wwa, neffs, _) in zip(scal_list, _replay(rec, wavemod)):
            scal.amplitude = wwa
            scal.wwz_Neffs = neffs
        return MScal(scalogram_list=scal_list)

    @functools.wraps(orig["ms_spectral"])
    def ms_spectral(self, method="lomb_scargle", freq=None, settings=None, mute_pbar=False, freq_kwargs=None,
                    label=None, verbose=False, scalogram_list=None):
        if not _wants_cuda(method, settings):
            return orig["ms_spectral"](self, method=method, freq=freq, settings=settings, mute_pbar=mute_pbar,
                                       freq_kwargs=freq_kwargs, label=label, verbose=verbose,
                                       scalogram_list=scalogram_list)
        settings = settings.copy()
        n_reuse = len(scalogram_list.scalogram_list) if scalogram_list else 0
        if settings.get("anti_alias", False) or n_reuse:
            # stored scalograms are reused member by member (core/multipleseries.py:44-58); the alias filter
            # is host post-processing: run in-process, member by member (never through the spawn pool)
            psd_list = []
            for idx, s in enumerate(self.series_list):
                kw = dict(method=method, settings=settings, freq=freq, freq_kwargs=freq_kwargs, label=label,
                          verbose=verbose)
                if n_reuse and idx < n_reuse:
                    kw["scalogram"] = scalogram_list.scalogram_list[idx]
                psd_list.append(s.spectral(**kw))
            return MPSD(psd_list=psd_list)
        with _recording() as rec, warnings.catch_warnings():
            warnings.simplefilter("ignore")       # wwa2psd of the placeholder scalogram divides 0 by 0
            psd_list = [s.spectral(method=method, settings=settings, freq=freq, freq_kwargs=freq_kwargs, label=label,
                                   verbose=verbose) for s in self.series_list]
        if len(rec) != len(psd_list):
            raise RuntimeError("expected one WWZ kernel request per series")
        thr = int(settings.get("Neff_threshold", 3))
        for psd, (_, _, p) in zip(psd_list, _replay(rec, wavemod, psd_threshold=thr)):
            psd.amplitude = p
        return MPSD(psd_list=psd_list)

    def _all_cuda(objs, attr):
        return all(isinstance(getattr(o, attr, None), dict) and getattr(o, attr).get("method") in CUDA_METHODS
                   for o in objs)

    @functools.wraps(orig["mscal_q"])
    def mscal_quantiles(self, qs=[0.05, 0.5, 0.95]):
        if not self.scalogram_list or not _all_cuda(self.scalogram_list, "wave_args"):
            return orig["mscal_q"](self, qs=qs)
        from . import batch
        first = self.scalogram_list[0]
        freq, scale, time, coi = (np.copy(first.frequency), np.copy(first.scale), np.copy(first.time),
                                  np.copy(first.coi))
        for scal in self.scalogram_list:                              # core/scalograms.py:627-631
            if not np.array_equal(scal.frequency, freq):
                raise ValueError("Frequency axis not consistent across the scalogram list!")
            if not np.array_equal(scal.time, time):
                raise ValueError("Time axis not consistent across the scalogram list!")
        amps = np.stack([scal.amplitude for scal in self.scalogram_list])
        amp_qs = batch.quantiles(amps, qs)                            # :635-641 on the device
        scal_list = [Scal(frequency=freq, time=time, amplitude=amp, scale=scale, coi=coi, label=f"{qs[i]*100:g}%")
                     for i, amp in enumerate(amp_qs)]                 # :643-647
        return MScal(scalogram_list=scal_list)

    @functools.wraps(orig["mpsd_q"])
    def mpsd_quantiles(self, qs=[0.05, 0.5, 0.95], lw=[0.5, 1.5, 0.5]):
        if not self.psd_list or not _all_cuda(self.psd_list, "spec_args"):
            return orig["mpsd_q"](self, qs=qs, lw=lw)
        from . import batch
        period_unit = None
        if self.psd_list[0].timeseries is not None:
            period_unit = self.psd_list[0].timeseries.time_unit
        freq = np.copy(self.psd_list[0].frequency)
        for psd in self.psd_list:                                     # core/psds.py:905-909
            if not np.array_equal(psd.frequency, freq):
                raise ValueError("Frequency axis not consistent across the PSD list!")
        amps = np.stack([psd.amplitude for psd in self.psd_list])
        amp_qs = batch.quantiles(amps, qs)                            # :913 on the device
        psd_list = [PSD(frequency=freq, amplitude=amp, label=f"{qs[i]*100:g}%",
                        plot_kwargs={"color": "gray", "linewidth": lw[i]}, period_unit=period_unit)
                    for i, amp in enumerate(amp_qs)]                  # :915-919
        return MPSD(psd_list=psd_list)

    def _fusable(ts, args, method="ar1sim"):
        """The fused device pipeline covers the default preprocessing; 'ar1sim' on an uneven axis only."""
        if method not in ("ar1sim", "uar1"):
            return False
        if ts is None or not isinstance(args, dict) or args.get("method") != CUDA_METHOD:
            return False
        if args.get("detrend", False) not in (False, None, "none") or args.get("gaussianize", False):
            return False
        if args.get("len_bd", 0) or args.get("anti_alias", False) or not args.get("standardize", True):
            return False
        # the even branch of ar1_sim is a statsmodels fit; uar1_fit / uar1_sim take any axis (utils/tsmodel.py:646-716)
        return method == "uar1" or not ts.is_evenly_spaced()

    def _surrogate_grid(ts, args):
        """tau / freq exactly as the surrogates' Series.wavelet / Series.spectral would build them
        (core/series.py:3545-3551: tau is rebuilt from ntau on the surrogate's own axis)."""
        freq = np.asarray(args["freq"], dtype=float)
        ntau = args.get("ntau", np.min([np.size(ts.time), 50]))
        tau = np.linspace(np.min(ts.time), np.max(ts.time), ntau)
        return wavemod.prepare_wwz(ts.value, ts.time, freq=freq, tau=tau)

    @functools.wraps(orig["scal_signif"])
    def scal_signif_test(self, method="ar1sim", number=None, seed=None, qs=[0.95], settings=None, export_scal=False):
        fused = (self.wave_method == "wwz" and not export_scal
                 and not getattr(self, "signif_scals", None) and _fusable(self.timeseries, self.wave_args, method))
        if not fused:
            return orig["scal_signif"](self, method=method, number=number, seed=seed, qs=qs, settings=settings,
                                       export_scal=export_scal)
        from . import batch
        if number is None:
            number = 200                                               # core/scalograms.py:490
        elif number == 0:
            return self
        if type(qs) is not list:
            raise TypeError("qs should be a list")                     # :519-520
        ts, args = self.timeseries, self.wave_args
        ys_cut, ts_cut, freq, tau = _surrogate_grid(ts, args)
        sg = batch.wwz_signif(ys_cut, ts_cut, tau=tau, freq=freq, number=number, qs=qs,
                              seed=_fresh_seed() if seed is None else seed, c=args.get("c", 1 / (8 * np.pi ** 2)),
                              Neff_threshold=args.get("Neff_threshold", 3), method=method)
        new = self.copy()
        coi = wavemod.make_coi(tau, Neff_threshold=args.get("Neff_coi", 3))
        new.signif_qs = MScal(scalogram_list=[
            Scal(frequency=freq, time=tau, amplitude=amp, scale=1 / freq, coi=coi, label=f"{qs[i]*100:g}%")
            for i, amp in enumerate(sg.amplitude_qs)])
        new.signif_method = method
        new.qs = qs
        return new

    @functools.wraps(orig["psd_signif"])
    def psd_signif_test(self, method="ar1sim", number=None, seed=None, qs=[0.95], settings=None, scalogram=None):
        fused = (self.spec_method == "wwz" and scalogram is None
                 and _fusable(self.timeseries, self.spec_args, method) and "tau" not in self.spec_args)
        if not fused:
            return orig["psd_signif"](self, method=method, number=number, seed=seed, qs=qs, settings=settings,
                                      scalogram=scalogram)
        from . import batch
        if number is None:
            number = 200
        elif number == 0:
            return self
        ts, args = self.timeseries, self.spec_args
        freq = np.asarray(args["freq"], dtype=float)
        ys_cut, ts_cut, freq, tau = wavemod.prepare_wwz(ts.value, ts.time, freq=freq)   # utils/spectral.py:882
        sg = batch.wwz_signif(ys_cut, ts_cut, tau=tau, freq=freq, number=number, qs=qs,
                              seed=_fresh_seed() if seed is None else seed, c=args.get("c", 1e-3), Neff_threshold=3, psd=True,
                              psd_Neff_threshold=args.get("Neff_threshold", 3), method=method)
        new = self.copy()
        lw = [0.5, 1.5, 0.5] if len(qs) == 3 else [1.0] * len(qs)
        new.signif_qs = MPSD(psd_list=[
            PSD(frequency=freq, amplitude=amp, label=f"{qs[i]*100:g}%", plot_kwargs={"color": "gray", "linewidth": lw[i]},
                period_unit=ts.time_unit) for i, amp in enumerate(sg.psd_qs)])
        new.signif_method = method
        return new

    @functools.wraps(orig["coh_signif"])
    def coh_signif_test(self, number=200, method="ar1sim", seed=None, qs=[0.95], settings=None, mute_pbar=False):
        """core/coherences.py:748-963 with the surrogate pairs as two shared-axis batches and the coherency smoothing
        of all pairs on the device (batch.coherence_signif); every other case takes the reference's own path."""
        args = getattr(self, "wave_args", None)
        ok = (self.wave_method == "wwz" and isinstance(args, dict) and _fusable(self.timeseries1, args, method)
              and _fusable(self.timeseries2, args, method) and "tau" in args and "freq" in args and not settings)
        if not ok:
            return orig["coh_signif"](self, number=number, method=method, seed=seed, qs=qs, settings=settings,
                                      mute_pbar=mute_pbar)
        if number == 0:
            return self                                                # :872-873
        from . import batch
        ts1, ts2 = self.timeseries1, self.timeseries2
        sg = batch.coherence_signif(ts1.value, ts1.time, ts2.value, ts2.time, tau=np.asarray(args["tau"], dtype=float),
                                    freq=np.asarray(args["freq"], dtype=float), number=number, qs=qs,
                                    seed=_fresh_seed() if seed is None else seed,
                                    smooth_factor=args.get("smooth_factor", 0.25), c=args.get("c", 1 / (8 * np.pi ** 2)),
                                    Neff_threshold=args.get("Neff_threshold", 3), method=method)
        new = self.copy()

        def levels(q):
            return MScal(scalogram_list=[
                Scal(frequency=self.frequency, time=self.time, amplitude=q[i], coi=self.coi, scale=self.scale,
                     freq_method=self.freq_method, freq_kwargs=self.freq_kwargs, label=f"{qs[i]*100:g}%")
                for i in range(len(qs))])

        new.signif_qs = [levels(sg.wtc_qs), levels(sg.xwt_qs)]          # :944-951: WTC quantiles, then XWT quantiles
        new.signif_method = method
        new.qs = qs
        return new

    MS.wavelet, MS.spectral = ms_wavelet, ms_spectral
    MScal.quantiles, MPSD.quantiles = mscal_quantiles, mpsd_quantiles
    Scal.signif_test, PSD.signif_test = scal_signif_test, psd_signif_test
    Coh.signif_test = coh_signif_test

    def undo():
        MS.wavelet, MS.spectral = orig["ms_wavelet"], orig["ms_spectral"]
        MScal.quantiles, MPSD.quantiles = orig["mscal_q"], orig["mpsd_q"]
        Scal.signif_test, PSD.signif_test = orig["scal_signif"], orig["psd_signif"]
        Coh.signif_test = orig["coh_signif"]

    return undo


def register(pyleoclim_wavelet=None, core=True, surrogates=False):
    """Add ``'Kirchner_cuda'`` (and ``'Foster_cuda'``, the reference's ``method='Foster'`` on the same kernels) to
    Pyleoclim.

    * wraps ``pyleoclim.utils.wavelet.get_wwz_func`` (kernel level); the registered kernel calls the
      reference's own ``preprocess`` and ``make_omega`` exactly as every reference kernel variant
      does (utils/wavelet.py:976-978);
    * ``core=True``: when the full package is importable, batches ``MultipleSeries.wavelet`` /
      ``.spectral``, moves the ``quantiles`` of CUDA results to the device and fuses
      ``signif_test`` (class level, see the module docstring);
    * ``surrogates=True``: additionally routes ``pyleoclim.utils.tsmodel.uar1_sim`` (shared time axis) and
      ``pyleoclim.utils.tsmodel.ar1_sim`` for unevenly spaced
      axes to the device generator (the reference draws from NumPy's unseeded global stream, so only
      the distribution, not the stream, is preserved).

    ``pyleoclim_wavelet`` may be the module object to patch (tests); otherwise Pyleoclim is imported.
    Returns the patched wavelet module.  Raises ImportError if Pyleoclim is not importable."""
    pyleo = None
    if pyleoclim_wavelet is None:
        import pyleoclim as pyleo
        import pyleoclim.utils.wavelet as pyleoclim_wavelet  # noqa: N813
    mod = pyleoclim_wavelet
    if getattr(mod.get_wwz_func, "_wwz_b200", False):
        return mod
    original = mod.get_wwz_func
    kernel = _make_kernel(mod)
    foster_kernel = _make_kernel(mod, foster=True)

    @functools.wraps(original)
    def get_wwz_func(nproc, method):
        if method in CUDA_METHODS:
            if hasattr(mod, "assertPositiveInt"):
                mod.assertPositiveInt(nproc)
            return kernel if method == CUDA_METHOD else foster_kernel
        return original(nproc, method)

    get_wwz_func._wwz_b200 = True
    undo = []
    mod.get_wwz_func = get_wwz_func
    mod.kirchner_cuda = kernel
    mod.foster_cuda = foster_kernel

    def undo_kernel():
        mod.get_wwz_func = original
        for name in ("kirchner_cuda", "foster_cuda"):
            if hasattr(mod, name):
                delattr(mod, name)

    undo.append(undo_kernel)
    if core and pyleo is not None and hasattr(pyleo, "core"):
        undo.append(_patch_core(pyleo, mod))
    if surrogates and pyleo is not None:
        tsmodel = pyleo.utils.tsmodel
        orig_sim = tsmodel.ar1_sim

        @functools.wraps(orig_sim)
        def ar1_sim(y, p, t=None):
            if t is None or tsmodel.is_evenly_spaced(t):
                return orig_sim(y, p, t)
            from . import batch
            return batch.ar1_sim(y, p, t, seed=int(np.random.randint(0, 2 ** 31 - 1)))

        tsmodel.ar1_sim = ar1_sim
        undo.append(lambda: setattr(tsmodel, "ar1_sim", orig_sim))
        orig_usim = tsmodel.uar1_sim

        @functools.wraps(orig_usim)
        def uar1_sim(t, tau, sigma_2=1):
            t_arr = np.asarray(t, dtype=float)
            if t_arr.ndim == 2 and not (t_arr == t_arr[:, :1]).all():
                return orig_usim(t, tau, sigma_2)                 # per-column axes: not one shared-axis batch
            from . import batch
            return batch.uar1_sim(t_arr, tau, sigma_2, seed=int(np.random.randint(0, 2 ** 31 - 1)))

        tsmodel.uar1_sim = uar1_sim
        undo.append(lambda: setattr(tsmodel, "uar1_sim", orig_usim))
    _state[id(mod)] = undo
    return mod


def unregister():
    """Undo every ``register()``."""
    for undo in list(_state.values()):
        for f in reversed(undo):
            f()
    _state.clear()
