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
    """utils/tsbase.py:316-352: drop NaNs, sort ascending, average duplicated time stamps."""
    ys = np.asarray(ys, dtype=float)
    ts = np.asarray(ts, dtype=float)
    assert ys.size == ts.size, "The size of time axis and data value should be equal!"
    good = ~np.isnan(ys)
    ys, ts = ys[good], ts[good]
    good = ~np.isnan(ts)
    ys, ts = ys[good], ts[good]
    order = np.argsort(ts)
    ys, ts = ys[order], ts[order]
    if ts.size != np.unique(ts).size:
        uts, start, counts = np.unique(ts, return_index=True, return_counts=True)
        ys = np.array([np.mean(ys[s:s + c]) for s, c in zip(start, counts)])
        ts = uts
    return ys, ts


def standardize(x, scale=1, axis=0, ddof=0, eps=1e-3):
    """utils/tsutils.py:1096-1156: z-score; (nearly) constant series are left unscaled."""
    x = np.asanyarray(x)
    assert x.ndim <= 2, "The time series x should be a vector or 2-D array!"
    mu = np.nanmean(x, axis=axis)
    sig = np.nanstd(x, axis=axis, ddof=ddof)
    mu2 = np.asarray(np.copy(mu))
    sig2 = np.asarray(np.copy(sig) / scale)
    if np.any(np.abs(sig) < eps):
        warnings.warn("Constant or nearly constant time series not rescaled.", stacklevel=2)
        where_const = np.abs(sig) < eps
        mu2[where_const] = 0
        sig2[where_const] = 1
    if axis and mu.ndim < x.ndim:
        z = (x - np.expand_dims(mu2, axis=axis)) / np.expand_dims(sig2, axis=axis)
    else:
        z = (x - mu2) / sig2
    return z, mu, sig


def preprocess(ys, ts, detrend=False, sg_kwargs=None, gaussianize=False, standardize=True):
    """utils/tsutils.py:1757-1825.  Detrending and gaussianization are optional preprocessing that
    is off by default on the WWZ path (utils/wavelet.py:1296-1297) and stays with Pyleoclim: when
    the package is registered inside Pyleoclim, ``register()`` routes through the reference's own
    ``preprocess``; standalone only the default (z-score) is provided."""
    if not (detrend == "none" or detrend is False or detrend is None) or gaussianize:
        raise NotImplementedError(
            "detrend/gaussianize are Pyleoclim preprocessing options; use pyleoclim_util_b200.register() "
            "inside Pyleoclim to have its preprocess() applied before the CUDA kernel")
    if standardize:
        return globals()["standardize"](ys)[0]
    return ys


def make_omega(ts, freq):
    """utils/wavelet.py:1210-1234: omega = 2*pi*freq, NaN above f_Nyquist = 0.5/median(diff(ts))."""
    f_Nyquist = 0.5 / np.median(np.diff(ts))
    freq_with_nan = np.array(freq, dtype=float, copy=True)
    freq_with_nan[freq_with_nan > f_Nyquist] = np.nan
    return 2 * np.pi * freq_with_nan


def make_coi(tau, Neff_threshold=3):
    """utils/wavelet.py:1168-1208: cone of influence."""
    assert isinstance(Neff_threshold, (int, np.integer)) and Neff_threshold >= 1
    nt = np.size(tau)
    fourier_factor = 4 * np.pi / (Neff_threshold + np.sqrt(2 + Neff_threshold ** 2))
    coi_const = fourier_factor / np.sqrt(2)
    dt = np.median(np.diff(tau))
    nt_half = (nt + 1) // 2 - 1
    A = np.append(0.00001, np.arange(nt_half) + 1)
    B = A[::-1]
    C = np.append(A, B) if nt % 2 == 0 else np.append(A, B[1:])
    return coi_const * dt * C


def freq_vector_log(ts, fmin=None, fmax=None, nf=None):
    """utils/wavelet.py:1935-1984."""
    nt = np.size(ts)
    dt = np.median(np.diff(ts))
    if nf is None:
        nf = nt // 10 + 1
    if fmin is None:
        fmin = 1 / np.ptp(ts)
    if fmax is None:
        fmax = (1 / dt) / 2
    return np.logspace(np.log2(fmin), np.log2(fmax), nf, base=2)


def make_freq_vector(ts, method="log", rayleigh_bound=True, **kwargs):
    """utils/wavelet.py:1986-2049.  Only the 'log' generator (the default of wwz / wwz_psd) is
    mirrored; the other generators belong to estimators outside this path."""
    if method == "log":
        fr = freq_vector_log(ts, **kwargs)
    elif method in ("lomb_scargle", "welch", "nfft", "scale"):
        raise NotImplementedError("freq_method=%r is not on the WWZ hot path; pass freq explicitly" % method)
    else:
        raise ValueError("This method is not supported")
    if rayleigh_bound:
        return fr[fr >= 1 / np.ptp(ts)]
    return fr


def prepare_wwz(ys, ts, freq=None, freq_method="log", freq_kwargs=None, tau=None, len_bd=0,
                bc_mode="reflect", reflect_type="odd", **kwargs):
    """utils/wavelet.py:2094-2202: clean the series, build / repair tau, optional ghost padding,
    truncate the series to the tau span, build freq, drop f == 0."""
    ys, ts = clean_ts(ys, ts)
    if tau is None:
        ntau = np.min([np.size(ts), 50])
        tau = np.linspace(np.min(ts), np.max(ts), ntau)
    elif np.isnan(tau).any():
        warnings.warn("The input tau contains some NaNs." +
                      "It will be regenerated using the boundarys of the time axis of the time series with NaNs deleted," +
                      "with the length of the size of the input tau.")
        tau = np.linspace(np.min(ts), np.max(ts), np.size(tau))
    elif np.min(tau) < np.min(ts) and np.max(tau) > np.max(ts):
        warnings.warn("tau should be within the time span of the time series." +
                      "Note that sometimes if the leading points of the time series are NaNs," +
                      "they will be deleted and cause np.min(tau) < np.min(ts)." +
                      "A new tau with the same size of the input tau will be generated.")
        tau = np.linspace(np.min(ts), np.max(ts), np.size(tau))
    elif np.min(tau) not in ts or np.max(tau) not in ts:
        warnings.warn("The boundaries of tau are not exactly on two of the time axis points," +
                      "and it will be adjusted to be so.")
        tau_lb = np.min(ts[ts > np.min(tau)])
        tau_ub = np.max(ts[ts < np.max(tau)])
        tau = np.linspace(tau_lb, tau_ub, np.size(tau))

    if len_bd > 0:                                               # :2172-2189
        dt = np.median(np.diff(ts))
        dtau = np.median(np.diff(tau))
        len_bd_tau = int(len_bd * dt // dtau)
        if bc_mode in ["reflect", "symmetric"]:
            ys = np.pad(ys, (len_bd, len_bd), bc_mode, reflect_type=reflect_type)
        else:
            ys = np.pad(ys, (len_bd, len_bd), bc_mode)
        ts_left_bd = np.linspace(ts[0] - dt * len_bd, ts[0] - dt, len_bd)
        ts_right_bd = np.linspace(ts[-1] + dt, ts[-1] + dt * len_bd, len_bd)
        ts = np.concatenate((ts_left_bd, ts, ts_right_bd))
        warnings.warn("The tau will be regenerated to fit the boundary condition.")
        tau_left_bd = np.linspace(tau[0] - dtau * len_bd_tau, tau[0] - dtau, len_bd_tau)
        tau_right_bd = np.linspace(tau[-1] + dtau, tau[-1] + dtau * len_bd_tau, len_bd_tau)
        tau = np.concatenate((tau_left_bd, tau, tau_right_bd))

    sel = (np.min(tau) <= ts) & (ts <= np.max(tau))
    ts_cut, ys_cut = ts[sel], ys[sel]
    if freq is None:
        freq_kwargs = {} if freq_kwargs is None else freq_kwargs.copy()
        freq = make_freq_vector(ts_cut, method=freq_method, **freq_kwargs)
    freq = np.asarray(freq, dtype=float)
    freq = freq[freq != 0]
    return ys_cut, ts_cut, freq, np.asarray(tau, dtype=float)


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
        Neffs, a0, a1, a2, wwa, phase = (alloc((nt, nf)) for _ in range(6))
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


def get_wwz_func(nproc, method):
    """utils/wavelet.py:2052-2092 restricted to the method this package provides."""
    assertPositiveInt(nproc)
    if method == "Kirchner_cuda":
        return kirchner_cuda
    raise ValueError('Wrong specific method name for WWZ. pyleoclim_util_b200 provides {"Kirchner_cuda"}; '
                     'use pyleoclim_util_b200.register() to add it to Pyleoclim\'s own dispatch')


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
        get_wwz_func(nproc, method)
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
