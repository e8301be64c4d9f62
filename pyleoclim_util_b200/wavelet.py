"""Host-side mirror of the reference's WWZ interface, backed by the CUDA library.

Names, argument meaning, defaults, return layout and error behaviour follow
``pyleoclim/utils/wavelet.py`` and ``pyleoclim/utils/spectral.py`` (file:line in each
docstring) so that this module is a drop-in for that path.  All (tau, f)-cell arithmetic runs in
``lib/libwwz_b200.so``; the functions here only do what the reference does on the host around its
kernels (cleaning, grid construction, z-scoring, cone of influence).
"""
import collections
import warnings

import numpy as np

from . import _lib

__all__ = [
    "kirchner_cuda", "wwz", "wwz_psd", "wwa2psd", "make_omega", "make_coi", "prepare_wwz",
    "freq_vector_log", "make_freq_vector", "clean_ts", "standardize", "preprocess", "get_wwz_func",
]

WwzResults = collections.namedtuple(
    "Results", ["amplitude", "phase", "coi", "freq", "time", "Neffs", "coeff", "scale"])
PsdResults = collections.namedtuple("Results", ["psd", "freq"])


# ------------------------------------------------------------------------------ host helpers
def assertPositiveInt(*args):
    """utils/wavelet.py: assertPositiveInt (used at :971, :2072)."""
    for arg in args:
        assert isinstance(arg, (int, np.integer)) and arg >= 1


def clean_ts(ys, ts, verbose=False):
    """What utils/tsbase.py:316-352 does to a series before WWZ: entries with a NaN value or stamp are
    dropped, the rest is put in ascending time order, and samples sharing a stamp are replaced by their mean."""
    ys = np.asarray(ys, dtype=float)
    ts = np.asarray(ts, dtype=float)
    assert ys.size == ts.size, "The size of time axis and data value should be equal!"
    if ts.ndim == 1 and ts.size and not np.isnan(ys).any() and (ts[1:] > ts[:-1]).all():
        return ys, ts                      # already clean (strictly ascending stamps, no NaN): nothing to do
    keep = ~(np.isnan(ys) | np.isnan(ts))
    ys, ts = ys[keep], ts[keep]
    order = np.argsort(ts)
    ys, ts = ys[order], ts[order]
    stamps, first, count = np.unique(ts, return_index=True, return_counts=True)
    if stamps.size != ts.size:
        ys = np.array([ys[i:i + c].mean() for i, c in zip(first, count)])
        ts = stamps
    return ys, ts


def standardize(x, scale=1, axis=0, ddof=0, eps=1e-3):
    """z-score with the conventions of utils/tsutils.py:1096-1156: NaN-aware mean and standard deviation
    (``ddof`` = 0), ``scale`` divides the deviation, and a series whose deviation is below ``eps`` is returned
    unshifted and unscaled with a warning.  Returns ``(z, mean, std)``."""
    x = np.asanyarray(x)
    assert x.ndim <= 2, "The time series x should be a vector or 2-D array!"
    if x.ndim == 1 and x.dtype == np.float64 and axis == 0:
        mu, sig = x.mean(), x.std(ddof=ddof)          # same reductions as the NaN-aware ones on NaN-free data
        if np.isnan(sig):
            mu, sig = np.nanmean(x, axis=axis), np.nanstd(x, axis=axis, ddof=ddof)
    else:
        mu = np.nanmean(x, axis=axis)
        sig = np.nanstd(x, axis=axis, ddof=ddof)
    flat = np.abs(sig) < eps
    shift = np.where(flat, 0.0, mu)
    spread = np.where(flat, 1.0, sig / scale)
    if np.any(flat):
        warnings.warn("Constant or nearly constant time series not rescaled.", stacklevel=2)
    if axis and np.ndim(mu) < x.ndim:
        shift, spread = np.expand_dims(shift, axis=axis), np.expand_dims(spread, axis=axis)
    return (x - shift) / spread, mu, sig


def preprocess(ys, ts, detrend=False, sg_kwargs=None, gaussianize=False, standardize=True):
    """The default branch of utils/tsutils.py:1757-1825 (z-score only).  Detrending and gaussianization are
    optional Pyleoclim preprocessing that is off by default on the WWZ path (utils/wavelet.py:1296-1297); when this
    package is registered inside Pyleoclim, ``register()`` routes through Pyleoclim's own ``preprocess`` so those
    options keep working."""
    if detrend not in ("none", False, None) or gaussianize:
        raise NotImplementedError(
            "detrend/gaussianize are Pyleoclim preprocessing options; use pyleoclim_util_b200.register() "
            "inside Pyleoclim to have its preprocess() applied before the CUDA kernel")
    return globals()["standardize"](ys)[0] if standardize else ys


def make_omega(ts, freq):
    """Angular frequencies for the kernel (utils/wavelet.py:1210-1234): 2*pi*f, NaN for every f above the Nyquist
    frequency of the median sampling interval."""
    freq = np.array(freq, dtype=float, copy=True)
    nyquist = 0.5 / np.median(np.diff(ts))
    return 2 * np.pi * np.where(freq > nyquist, np.nan, freq)


def make_coi(tau, Neff_threshold=3):
    """Cone of influence (utils/wavelet.py:1168-1208): a tent over the tau axis, distance (in steps of the median
    tau spacing) to the nearer end, with 1e-5 at the two ends, times the e-folding constant of the Morlet-like
    wavelet for ``Neff_threshold``."""
    assert isinstance(Neff_threshold, (int, np.integer)) and Neff_threshold >= 1
    nt = np.size(tau)
    efold = 4 * np.pi / (Neff_threshold + np.sqrt(2 + Neff_threshold ** 2)) / np.sqrt(2)
    step = np.median(np.diff(tau))
    idx = np.arange(nt)
    tent = np.minimum(idx, nt - 1 - idx).astype(float)
    tent[tent == 0] = 0.00001
    return efold * step * tent


def freq_vector_log(ts, fmin=None, fmax=None, nf=None):
    """Log-spaced frequency axis (utils/wavelet.py:1935-1984): nf = n//10 + 1 points from the Rayleigh
    frequency 1/span to the Nyquist frequency of the median sampling interval, unless given."""
    n = np.size(ts)
    lo = 1 / np.ptp(ts) if fmin is None else fmin
    hi = 0.5 / np.median(np.diff(ts)) if fmax is None else fmax
    return np.logspace(np.log2(lo), np.log2(hi), n // 10 + 1 if nf is None else nf, base=2)


def freq_vector_lomb_scargle(ts, dt=None, nf=None, ofac=4, hifac=0.95):
    """REDFIT-style axis (utils/wavelet.py:1689-1742): from the Nyquist frequency over ``n*ofac`` up to ``hifac``
    times the Nyquist frequency of the median (or given) sampling interval, in steps of the lowest frequency."""
    assert ofac >= 1 and hifac <= 1, "`ofac` should be >= 1, and `hifac` should be <= 1"
    step = np.median(np.diff(ts)) if dt is None else dt
    lowest = (1 / (2 * step)) / (np.size(ts) * ofac)
    highest = hifac / (2 * step)
    count = int((highest - lowest) / lowest + 1) if nf is None else nf
    return np.linspace(lowest, highest, count)


def freq_vector_welch(ts):
    """Welch / periodogram axis (utils/wavelet.py:1744-1787): ``k * fs / n`` for the non-negative FFT bins."""
    n = np.size(ts)
    fs = 1 / np.median(np.diff(ts))
    bins = n // 2 + 1 if n % 2 == 0 else (n + 1) // 2
    return np.arange(bins) * fs / n


def freq_vector_nfft(ts):
    """NFFT axis (utils/wavelet.py:1789-1823): ``n//2 + 1`` points from 0 to the Nyquist frequency."""
    fs = 1 / np.median(np.diff(ts))
    return np.linspace(0, fs / 2, np.size(ts) // 2 + 1)


def freq_vector_scale(ts, dj=0.25, s0=None, j1=None, mother="MORLET", param=None):
    """Torrence-Compo scale axis turned into frequencies (utils/wavelet.py:1825-1933): scales ``s0 * 2^(j dj)``,
    ``j = 0..j1``, divided into the Fourier factor of the mother wavelet, ascending."""
    kind = mother.upper()
    if kind not in ("MORLET", "DOG", "PAUL"):
        raise ValueError('The mother wavelet should be either "MORLET","PAUL", or "DOG"')
    step = np.diff(ts).mean()
    smallest = 2 * step if s0 is None else s0
    last = np.fix((np.log(len(ts) * step / smallest) / np.log(2)) / dj) if j1 is None else j1
    if kind == "MORLET":
        k0 = 6. if param is None else param
        factor = 4 * np.pi / (k0 + np.sqrt(2 + k0 ** 2))
    elif kind == "PAUL":
        m = 4. if param is None else param
        factor = 4 * np.pi / (2 * m + 1)
    else:
        m = 2. if param is None else param
        factor = 2 * np.pi * np.sqrt(2. / (2 * m + 1))
    scales = smallest * 2. ** (np.arange(0, last + 1) * dj)
    return np.sort(1. / (factor * scales))


def make_freq_vector(ts, method="log", rayleigh_bound=True, **kwargs):
    """utils/wavelet.py:1986-2049: the frequency axis of ``wwz`` / ``wwz_psd`` by one of the reference's five
    generators, cut at the Rayleigh frequency ``1/span`` unless ``rayleigh_bound`` is False."""
    makers = {"lomb_scargle": freq_vector_lomb_scargle, "welch": freq_vector_welch, "nfft": freq_vector_nfft,
              "scale": freq_vector_scale, "log": freq_vector_log}
    if method not in makers:
        raise ValueError("This method is not supported")
    fr = makers[method](ts, **kwargs) if method in ("lomb_scargle", "scale", "log") else makers[method](ts)
    return fr[fr >= 1 / np.ptp(ts)] if rayleigh_bound else fr


def _repair_tau(tau, ts):
    """tau as wwz will use it (utils/wavelet.py:2147-2169): default 50 points over the series, regenerated over
    the series span when it holds NaNs or overhangs both ends, and pulled onto time-axis points when its ends
    are not samples of the series."""
    t_lo, t_hi = np.min(ts), np.max(ts)
    if tau is None:
        return np.linspace(t_lo, t_hi, min(np.size(ts), 50))
    tau = np.asarray(tau, dtype=float)
    if np.isnan(tau).any():
        warnings.warn("The input tau contains some NaNs; it is regenerated over the time span of the series "
                      "(NaNs deleted) with the same number of points.")
        return np.linspace(t_lo, t_hi, tau.size)
    if np.min(tau) < t_lo and np.max(tau) > t_hi:
        warnings.warn("tau should lie within the time span of the series (leading NaNs of the series are deleted "
                      "first); a new tau with the same number of points is generated over the series span.")
        return np.linspace(t_lo, t_hi, tau.size)
    if np.min(tau) not in ts or np.max(tau) not in ts:
        warnings.warn("The ends of tau are not points of the time axis; tau is adjusted so that they are.")
        return np.linspace(np.min(ts[ts > np.min(tau)]), np.max(ts[ts < np.max(tau)]), tau.size)
    return tau


def _ghost_pad(ys, ts, tau, len_bd, bc_mode, reflect_type):
    """``len_bd`` ghost points on both sides of the series and the matching extension of tau
    (utils/wavelet.py:2172-2189)."""
    dt, dtau = np.median(np.diff(ts)), np.median(np.diff(tau))
    n_tau = int(len_bd * dt // dtau)
    pad_kw = {"reflect_type": reflect_type} if bc_mode in ("reflect", "symmetric") else {}
    ys = np.pad(ys, (len_bd, len_bd), bc_mode, **pad_kw)
    ts = np.concatenate((np.linspace(ts[0] - dt * len_bd, ts[0] - dt, len_bd), ts,
                         np.linspace(ts[-1] + dt, ts[-1] + dt * len_bd, len_bd)))
    warnings.warn("The tau will be regenerated to fit the boundary condition.")
    tau = np.concatenate((np.linspace(tau[0] - dtau * n_tau, tau[0] - dtau, n_tau), tau,
                          np.linspace(tau[-1] + dtau, tau[-1] + dtau * n_tau, n_tau)))
    return ys, ts, tau


def prepare_wwz(ys, ts, freq=None, freq_method="log", freq_kwargs=None, tau=None, len_bd=0,
                bc_mode="reflect", reflect_type="odd", **kwargs):
    """utils/wavelet.py:2094-2202: clean the series, build or repair tau, optional ghost padding, cut the series to
    the tau span, build freq when it is not given, drop f == 0.  Returns ``ys_cut, ts_cut, freq, tau``."""
    ys, ts = clean_ts(ys, ts)
    tau = _repair_tau(tau, ts)
    if len_bd > 0:
        ys, ts, tau = _ghost_pad(ys, ts, tau, len_bd, bc_mode, reflect_type)
    inside = (np.min(tau) <= ts) & (ts <= np.max(tau))
    ts_cut, ys_cut = ts[inside], ys[inside]
    if freq is None:
        freq = make_freq_vector(ts_cut, method=freq_method, **({} if freq_kwargs is None else dict(freq_kwargs)))
    freq = np.asarray(freq, dtype=float)
    return ys_cut, ts_cut, freq[freq != 0], np.asarray(tau, dtype=float)


# ------------------------------------------------------------------------------ the kernel
def _transform(pd_ys, ts, tau, omega, c, Neff_threshold, flags=0, psd_threshold=None, pinned=True, psd_only=False):
    """One call into the C ABI: (1) wwzb200_kirchner or (2) wwzb200_kirchner_psd.
    Returns (wwa, phase, Neffs, (a0, a1, a2), psd); with psd_only the per-cell arrays stay on the device
    and come back as None."""
    L = _lib.lib()
    ts, pd_ys, tau, omega = _lib.f64(ts), _lib.f64(pd_ys), _lib.f64(tau), _lib.f64(omega)
    if ts.ndim != 1 or pd_ys.shape != ts.shape:
        raise ValueError("ys and ts must be vectors of equal length")
    nt, nf = tau.size, omega.size
    alloc = _lib.output_empty if pinned else np.empty
    if psd_only and psd_threshold is not None:
        Neffs = a0 = a1 = a2 = wwa = phase = None
    else:
        Neffs, a0, a1, a2, wwa, phase = alloc((6, nt, nf))      # one block, the library's output order: one DMA
    if psd_threshold is None:
        _lib.check(L.wwzb200_kirchner(_lib.ptr(ts), _lib.ptr(pd_ys), ts.size, _lib.ptr(tau), nt, _lib.ptr(omega),
                                      nf, float(c), int(Neff_threshold), int(flags), _lib.ptr(Neffs), _lib.ptr(a0),
                                      _lib.ptr(a1), _lib.ptr(a2), _lib.ptr(wwa), _lib.ptr(phase)))
        return wwa, phase, Neffs, (a0, a1, a2), None
    psd = np.empty(nf)
    _lib.check(L.wwzb200_kirchner_psd(_lib.ptr(ts), _lib.ptr(pd_ys), ts.size, _lib.ptr(tau), nt, _lib.ptr(omega),
                                      nf, float(c), int(Neff_threshold), int(psd_threshold), int(flags),
                                      _lib.ptr(psd), _lib.ptr(Neffs), _lib.ptr(a0), _lib.ptr(a1), _lib.ptr(a2),
                                      _lib.ptr(wwa), _lib.ptr(phase)))
    return wwa, phase, Neffs, (a0, a1, a2), psd


def kirchner_cuda(ys, ts, freq, tau, c=1 / (8 * np.pi ** 2), Neff_threshold=3, nproc=1, detrend=False,
                  sg_kwargs=None, gaussianize=False, standardize=False, flags=0,
                  _preprocess=None, _make_omega=None):
    """The CUDA member of the reference's kernel family (same signature and return as
    ``kirchner_numba``, utils/wavelet.py:872-1049; replaces the f2py kernel, :1051-1166).

    Returns ``wwa, phase, Neffs, (a0, a1, a2)`` -- fresh float64 arrays [ntau, nfreq], NaN in
    masked cells.  ``nproc`` is accepted and ignored (as in kirchner_numba)."""
    assertPositiveInt(Neff_threshold)                                            # :971
    pre = preprocess if _preprocess is None else _preprocess
    mko = make_omega if _make_omega is None else _make_omega
    pd_ys = pre(ys, ts, detrend=detrend, sg_kwargs=sg_kwargs, gaussianize=gaussianize,
                standardize=standardize)                                         # :976
    omega = mko(ts, freq)                                                        # :978
    wwa, phase, Neffs, coeff, _ = _transform(pd_ys, ts, tau, omega, c, Neff_threshold, flags)
    return wwa, phase, Neffs, coeff


def foster_cuda(ys, ts, freq, tau, c=1 / (8 * np.pi ** 2), Neff_threshold=3, nproc=1, detrend=False,
                sg_kwargs=None, gaussianize=False, standardize=False, flags=0, _preprocess=None, _make_omega=None):
    """Foster's original projection (the reference's ``method='Foster'``, ``wwz_basic``, utils/wavelet.py:228-381:
    3x3 normal equations per cell solved with a pseudo-inverse, :357-374) on the same CUDA kernels.
    Same signature and return as ``wwz_basic``; the coefficients are ``(ywave_1, ywave_2, ywave_3)``."""
    return kirchner_cuda(ys, ts, freq, tau, c=c, Neff_threshold=Neff_threshold, nproc=nproc, detrend=detrend,
                         sg_kwargs=sg_kwargs, gaussianize=gaussianize, standardize=standardize,
                         flags=int(flags) | _lib.FOSTER, _preprocess=_preprocess, _make_omega=_make_omega)


def get_wwz_func(nproc, method):
    """utils/wavelet.py:2052-2092 restricted to the methods this package provides."""
    assertPositiveInt(nproc)
    if method == "Kirchner_cuda":
        return kirchner_cuda
    if method == "Foster_cuda":
        return foster_cuda
    raise ValueError('Wrong specific method name for WWZ. pyleoclim_util_b200 provides {"Kirchner_cuda", "Foster_cuda"}; '
                     'use pyleoclim_util_b200.register() to add them to Pyleoclim\'s own dispatch')


def wwz(ys, ts, tau=None, ntau=None, freq=None, freq_method="log", freq_kwargs={}, c=1 / (8 * np.pi ** 2),
        Neff_threshold=3, Neff_coi=3, nproc=8, detrend=False, sg_kwargs=None, method="Kirchner_cuda",
        gaussianize=False, standardize=True, len_bd=0, bc_mode="reflect", reflect_type="odd", flags=0):
    """utils/wavelet.py:1294-1494 (body :1471-1494): weighted wavelet Z-transform of an unevenly
    sampled series.  ``Results(amplitude, phase, coi, freq, time, Neffs, coeff, scale)``."""
    if standardize == True:  # noqa: E712  (mirrors :1471)
        warnings.warn("Standardizing the timeseries")
    ys_cut, ts_cut, freq, tau = prepare_wwz(ys, ts, freq=freq, freq_method=freq_method, freq_kwargs=freq_kwargs,
                                            tau=tau, len_bd=len_bd, bc_mode=bc_mode, reflect_type=reflect_type)
    wwz_func = get_wwz_func(nproc, method)
    wwa, phase, Neffs, coeff = wwz_func(ys_cut, ts_cut, freq, tau, Neff_threshold=Neff_threshold, c=c, nproc=nproc,
                                        detrend=detrend, sg_kwargs=sg_kwargs, gaussianize=gaussianize,
                                        standardize=standardize, flags=flags)
    coi = make_coi(tau, Neff_threshold=Neff_coi)
    return WwzResults(amplitude=wwa, phase=phase, coi=coi, freq=freq, time=tau, Neffs=Neffs, coeff=coeff,
                      scale=1 / freq)


def wwa2psd(wwa, ts, Neffs, freq=None, Neff_threshold=3, anti_alias=False, avgs=2):
    """utils/wavelet.py:1236-1292: PSD from an existing scalogram (tau-reduction on the device).
    The optional anti-alias filter (:1281-1290, off by default) is Pyleoclim post-processing."""
    if anti_alias:
        raise NotImplementedError("anti_alias is post-processing outside the WWZ hot path")
    wwa, Neffs = _lib.f64(wwa), _lib.f64(Neffs)
    nt, nf = wwa.shape
    psd = np.empty(nf)
    _lib.check(_lib.lib().wwzb200_wwa2psd(_lib.ptr(wwa), _lib.ptr(Neffs), nt, nf,
                                          float(np.max(ts) - np.min(ts)), int(np.size(ts)), int(Neff_threshold),
                                          _lib.ptr(psd)))
    return psd


def wwz_psd(ys, ts, freq=None, freq_method="log", freq_kwargs=None, tau=None, c=1e-3, nproc=8, detrend=False,
            sg_kwargs=None, gaussianize=False, standardize=True, Neff_threshold=3, anti_alias=False, avgs=2,
            method="Kirchner_cuda", wwa=None, wwz_Neffs=None, wwz_freq=None, flags=0):
    """utils/spectral.py:728-900.  The transform and the wwa2psd reduction run as one fused call
    (C-ABI entry 2).  As in the reference (:889-891) ``Neff_threshold`` is used by wwa2psd only;
    the transform keeps wwz's default threshold 3."""
    if standardize == True:  # noqa: E712  (mirrors :879)
        warnings.warn("Standardizing the timeseries")
    if anti_alias:
        raise NotImplementedError("anti_alias is post-processing outside the WWZ hot path")
    ys_cut, ts_cut, freq, tau = prepare_wwz(ys, ts, freq=freq, freq_method=freq_method, freq_kwargs=freq_kwargs,
                                            tau=tau)
    if wwa is None or wwz_Neffs is None or wwz_freq is None:
        # what wwz(ys_cut, ts_cut, freq=freq, tau=tau, c=c, ...) does (:889), fused with wwa2psd (:896)
        if standardize == True:  # noqa: E712  (the nested wwz call warns again, utils/wavelet.py:1471)
            warnings.warn("Standardizing the timeseries")
        ys2, ts2, freq2, tau2 = prepare_wwz(ys_cut, ts_cut, freq=freq, tau=tau)
        if get_wwz_func(nproc, method) is foster_cuda:
            flags = int(flags) | _lib.FOSTER
        pd_ys = preprocess(ys2, ts2, detrend=detrend, sg_kwargs=sg_kwargs, gaussianize=gaussianize,
                           standardize=standardize)
        omega = make_omega(ts2, freq2)
        # wwa2psd takes its span and count from ts_cut (:896)
        if ts2.size == ts_cut.size:
            psd = _transform(pd_ys, ts2, tau2, omega, c, 3, flags, psd_threshold=Neff_threshold, psd_only=True)[4]
        else:
            amp, _, ne, _, _ = _transform(pd_ys, ts2, tau2, omega, c, 3, flags)
            psd = wwa2psd(amp, ts_cut, ne, freq=freq2, Neff_threshold=Neff_threshold)
    else:
        psd = wwa2psd(wwa, ts_cut, wwz_Neffs, freq=wwz_freq, Neff_threshold=Neff_threshold)
    return PsdResults(psd=psd, freq=freq)


# ------------------------------------------------------------------------------------------------
# wavelet coherence: utils/wavelet.py:1497-1687 (wwz_coherence), :2201-2364 (xwt, wtc)
# ------------------------------------------------------------------------------------------------
CoherenceResults = collections.namedtuple(
    "Results", ["xw_coherence", "xw_amplitude", "xw_phase", "xwt", "freq", "time", "coi", "scale"])


def coherence_grid(y1, t1, y2, t2, tau=None, freq=None, freq_method="log", freq_kwargs=None, verbose=False):
    """The common (tau, freq) grid and the two truncated series of ``wwz_coherence`` (utils/wavelet.py:1621-1658):
    default tau over the overlap of the two axes, default freq from the first axis, both series prepared on it, and
    the two prepared grids reconciled when they differ.  Returns (y1_cut, t1_cut, y2_cut, t2_cut, tau, freq)."""
    if tau is None:                                                      # :1621-1629
        lb, ub = max(np.min(t1), np.min(t2)), min(np.max(t1), np.max(t2))
        inside = t1[(t1 >= lb) & (t1 <= ub)]
        tau = np.linspace(lb, ub, np.size(inside) // 10)
        print(f'Setting tau={tau[:3]}...{tau[-3:]}, ntau={np.size(tau)}')
    if freq is None:                                                     # :1631-1634
        freq = make_freq_vector(t1, method=freq_method, **({} if freq_kwargs is None else freq_kwargs.copy()))
        print(f'Setting freq={freq[:3]}...{freq[-3:]}, nf={np.size(freq)}')
    y1_cut, t1_cut, freq1, tau1 = prepare_wwz(y1, t1, freq=freq, tau=tau)   # :1636-1637
    y2_cut, t2_cut, freq2, tau2 = prepare_wwz(y2, t2, freq=freq, tau=tau)
    if np.size(tau1) != np.size(tau2) or np.any(tau1 != tau2):            # :1639-1646
        if verbose:
            print('inconsistent `tau`, recalculating...')
        tau = np.linspace(min(np.min(tau1), np.min(tau2)), max(np.max(tau1), np.max(tau2)),
                          max(np.size(tau1), np.size(tau2)))
    else:
        tau = tau1
    if np.size(freq1) != np.size(freq2) or np.any(freq1 != freq2):        # :1648-1655
        if verbose:
            print('inconsistent `freq`, recalculating...')
        freq = np.linspace(min(np.min(freq1), np.min(freq2)), max(np.max(freq1), np.max(freq2)),
                           max(np.size(freq1), np.size(freq2)))
    else:
        freq = freq1
    if freq[0] == 0:
        freq = freq[1:]                                                  # :1657-1658
    return y1_cut, t1_cut, y2_cut, t2_cut, np.ascontiguousarray(tau, dtype=float), np.ascontiguousarray(freq, dtype=float)


def coherence_inputs(y1, t1, y2, t2, tau=None, freq=None, freq_method="log", freq_kwargs=None, verbose=False):
    """What the two transforms of ``wwz_coherence`` actually see.  The reference hands the truncated series and the
    common grid to ``wwz`` (utils/wavelet.py:1660-1663), whose own ``prepare_wwz`` (:1474) runs once more on each
    series: it may repair tau for that series (:2151-2169) and truncates the series to the tau span (:2192-2193), so
    the two coefficient sets can sit on slightly different tau values while wtc, coi and the returned `time` use the
    common tau.  Returns ((ys1, ts1, tau1), (ys2, ts2, tau2), tau, freq)."""
    y1c, t1c, y2c, t2c, tau, freq = coherence_grid(y1, t1, y2, t2, tau=tau, freq=freq, freq_method=freq_method,
                                                   freq_kwargs=freq_kwargs, verbose=verbose)
    y1p, t1p, f1, tau1 = prepare_wwz(y1c, t1c, freq=freq, tau=tau)
    y2p, t2p, f2, tau2 = prepare_wwz(y2c, t2c, freq=freq, tau=tau)
    if np.size(tau1) != np.size(tau) or np.size(tau2) != np.size(tau) or np.size(f1) != np.size(freq):
        raise ValueError("the two series do not share one (tau, freq) grid")   # the reference fails in wtc's broadcast
    return (y1p, t1p, np.ascontiguousarray(tau1, dtype=float)), (y2p, t2p, np.ascontiguousarray(tau2, dtype=float)), tau, freq


def wtc_boxcar_length(freq):
    """int(round(0.6 / dj * 2)) with dj = log2(scale[0] / scale[-1]) / N, scale = 1 / freq (utils/wavelet.py:2331-2349)."""
    scales = 1 / np.asarray(freq, dtype=float)
    dj = np.log2(scales[0] / scales[-1]) / np.size(scales)
    return int(np.round(0.6 / dj * 2))


def wtc_cuda(a1_1, a2_1, a1_2, a2_2, freq, tau, smooth_factor=0.25, want_phase=True, want_amplitude=True):
    """Wavelet transform coherency (utils/wavelet.py:2240-2364) and cross-wavelet amplitude (:2230) of coefficient
    pairs on the device (C-ABI ``wwzb200_wtc``).  Coefficient arrays [ntau, nf] or [S, ntau, nf] (coeff = a1 - i a2)."""
    arrs = [_lib.f64(x) for x in (a1_1, a2_1, a1_2, a2_2)]
    shape = arrs[0].shape
    if any(a.shape != shape for a in arrs) or len(shape) not in (2, 3):
        raise ValueError("coefficient arrays must share the shape [ntau, nf] or [S, ntau, nf]")
    S = 1 if len(shape) == 2 else shape[0]
    nt, nf = shape[-2], shape[-1]
    freq = _lib.f64(freq)
    L = wtc_boxcar_length(freq)
    if L < 1:
        raise IndexError("the scale-smoothing window of wtc is empty for this frequency axis")     # rect(0) in the reference
    coh = np.empty(shape)
    ph = np.empty(shape) if want_phase else None
    amp = np.empty(shape) if want_amplitude else None
    _lib.check(_lib.lib().wwzb200_wtc(*[_lib.ptr(a) for a in arrs], S, nt, nf, _lib.ptr(freq),
                                      float(np.median(np.diff(tau))), float(smooth_factor), L, _lib.ptr(coh), _lib.ptr(ph),
                                      _lib.ptr(amp)))
    return coh, ph, amp


def wwz_coherence(y1, t1, y2, t2, smooth_factor=0.25, tau=None, freq=None, freq_method="log", freq_kwargs=None,
                  c=1 / (8 * np.pi ** 2), Neff_threshold=3, nproc=8, detrend=False, sg_kwargs=None, verbose=False,
                  method="Kirchner_cuda", gaussianize=False, standardize=True):
    """``pyleoclim.utils.wavelet.wwz_coherence`` (utils/wavelet.py:1497-1687) with both transforms and the coherency
    smoothing on the device.  Same signature and Results (xw_coherence, xw_amplitude, xw_phase, xwt, freq, time, coi,
    scale)."""
    if standardize:
        warnings.warn("Standardizing the timeseries")                    # :1618-1619
    (y1p, t1p, tau1), (y2p, t2p, tau2), tau, freq = coherence_inputs(y1, t1, y2, t2, tau=tau, freq=freq,
                                                                     freq_method=freq_method, freq_kwargs=freq_kwargs,
                                                                     verbose=verbose)
    kern = get_wwz_func(nproc, method)
    kw = dict(c=c, Neff_threshold=Neff_threshold, nproc=nproc, detrend=detrend, sg_kwargs=sg_kwargs,
              gaussianize=gaussianize, standardize=standardize)
    _, _, _, co1 = kern(y1p, t1p, freq, tau1, **kw)                      # :1660-1663 (each on its own prepared grid)
    _, _, _, co2 = kern(y2p, t2p, freq, tau2, **kw)
    coh, ph, amp = wtc_cuda(co1[1], co1[2], co2[1], co2[2], freq, tau, smooth_factor=smooth_factor)
    prod = (co1[1] - co1[2] * 1j) * np.conj(co2[1] - co2[2] * 1j)        # :1664-1665, :2229
    return CoherenceResults(xw_coherence=coh, xw_amplitude=amp, xw_phase=ph, xwt=prod, freq=freq, time=tau,
                            coi=make_coi(tau, Neff_threshold=Neff_threshold), scale=1 / freq)
