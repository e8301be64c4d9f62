"""CPU oracle for the WWZ hot path -- TEST INFRASTRUCTURE ONLY.

Python face of ``wwz_oracle.c`` plus NumPy restatements of the thin host-side
functions around the kernel.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module;
the product package ``pyleoclim_util_b200`` never does.

Every function cites the reference lines it follows (paths relative to
``/root/reference/pyleoclim/``).  The oracle is pinned against the reference's golden
scalogram and against outputs of the reference run in the build container
(``tests/test_oracle_golden.py``; fixtures made by ``tests/golden/make_golden.py``).
"""
import collections
import ctypes
import os
import subprocess
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwwz_oracle.so")
_lib_handle = None

_dp = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64


def build(force=False):
    """Compile wwz_oracle.c with the Makefile next to it (gcc -O2 -fopenmp, no fast-math)."""
    src = os.path.join(_HERE, "wwz_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libwwz_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def _lib():
    global _lib_handle
    if _lib_handle is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.oracle_kirchner.restype = ctypes.c_int
        lib.oracle_kirchner.argtypes = [_dp, _dp, _i64, _dp, _i64, _dp, _i64, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, _dp]
        lib.oracle_amp_phase.restype = None
        lib.oracle_amp_phase.argtypes = [_dp, _dp, _i64, _dp, _dp]
        lib.oracle_make_omega.restype = ctypes.c_int
        lib.oracle_make_omega.argtypes = [_dp, _i64, _dp, _i64, _dp]
        lib.oracle_make_coi.restype = ctypes.c_int
        lib.oracle_make_coi.argtypes = [_dp, _i64, ctypes.c_int, _dp]
        lib.oracle_wwa2psd.restype = None
        lib.oracle_wwa2psd.argtypes = [_dp, _dp, _i64, _i64, ctypes.c_double, _i64, ctypes.c_double, _dp]
        lib.oracle_standardize.restype = ctypes.c_int
        lib.oracle_standardize.argtypes = [_dp, _i64, _dp, _dp, _dp]
        lib.oracle_ar1_model.restype = None
        lib.oracle_ar1_model.argtypes = [_dp, _i64, ctypes.c_double, ctypes.c_double, _dp, _dp]
        lib.oracle_ar1_objective.restype = ctypes.c_double
        lib.oracle_ar1_objective.argtypes = [_dp, _dp, _i64, ctypes.c_double]
        lib.oracle_num_threads.restype = ctypes.c_int
        _lib_handle = lib
    return _lib_handle


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def num_threads():
    return int(_lib().oracle_num_threads())


# --------------------------------------------------------------------------- kernel level
def make_omega(ts, freq):
    """utils/wavelet.py:1210-1234."""
    ts, freq = _c(ts), _c(freq)
    omega = np.empty_like(freq)
    rc = _lib().oracle_make_omega(_p(ts), ts.size, _p(freq), freq.size, _p(omega))
    if rc:
        raise ValueError("make_omega needs at least two time points")
    return omega


def make_coi(tau, Neff_threshold=3):
    """utils/wavelet.py:1168-1208."""
    assert isinstance(Neff_threshold, int) and Neff_threshold >= 1
    tau = _c(tau)
    coi = np.empty_like(tau)
    rc = _lib().oracle_make_coi(_p(tau), tau.size, Neff_threshold, _p(coi))
    if rc:
        raise ValueError("make_coi needs at least two tau points")
    return coi


def standardize(y):
    """utils/tsutils.py:1096-1156 for a NaN-free vector (scale=1, ddof=0, eps=1e-3)."""
    y = _c(y)
    z = np.empty_like(y)
    mu = ctypes.c_double()
    sig = ctypes.c_double()
    const = _lib().oracle_standardize(_p(y), y.size, _p(z), ctypes.byref(mu), ctypes.byref(sig))
    if const:
        warnings.warn("Constant or nearly constant time series not rescaled.", stacklevel=2)
    return z, mu.value, sig.value


def kirchner_cells(ts, pd_ys, tau, omega, c, Neff_threshold, mode=0, nthreads=0):
    """The cell loop of utils/wavelet.py:985-1042 on already-preprocessed ys and a given omega.

    Returns Neffs, a0, a1, a2 each [ntau, nf].  mode bit0: 0 = faithful two-pass, 1 = one-pass check; bit1: 1 = kirchner_basic NaN-mask semantics."""
    ts, pd_ys, tau, omega = _c(ts), _c(pd_ys), _c(tau), _c(omega)
    nt, nf = tau.size, omega.size
    out = [np.empty((nt, nf)) for _ in range(4)]
    rc = _lib().oracle_kirchner(_p(ts), _p(pd_ys), ts.size, _p(tau), nt, _p(omega), nf, float(c),
                                float(Neff_threshold), int(mode), int(nthreads), *[_p(o) for o in out])
    if rc:
        raise RuntimeError("oracle_kirchner failed: %d" % rc)
    return tuple(out)


def kirchner(ys, ts, freq, tau, c=1 / (8 * np.pi ** 2), Neff_threshold=3, nproc=1, detrend=False,
             sg_kwargs=None, gaussianize=False, standardize=False, mode=0, nthreads=0):
    """Same signature and return as the reference kernels (utils/wavelet.py:872-1049):
    ``wwa, phase, Neffs, (a0, a1, a2)``.  detrend / gaussianize are off on this path."""
    assert isinstance(Neff_threshold, int) and Neff_threshold >= 1   # assertPositiveInt, :971
    if detrend not in (False, None, "none") or gaussianize:
        raise NotImplementedError("oracle covers the default preprocessing only")
    pd_ys = globals()["standardize"](ys)[0] if standardize else _c(ys)   # :976, tsutils.py:1816-1819
    omega = make_omega(ts, freq)                                          # :978
    Neffs, a0, a1, a2 = kirchner_cells(ts, pd_ys, tau, omega, c, Neff_threshold, mode, nthreads)
    wwa = np.empty_like(a1)
    phase = np.empty_like(a1)
    _lib().oracle_amp_phase(_p(a1), _p(a2), a1.size, _p(wwa), _p(phase))   # :1044-1045
    return wwa, phase, Neffs, (a0, a1, a2)


def foster(ys, ts, freq, tau, c=1 / (8 * np.pi ** 2), Neff_threshold=3, nproc=1, detrend=False, sg_kwargs=None,
           gaussianize=False, standardize=False):
    """wwz_basic (the reference's method='Foster'), utils/wavelet.py:328-381, restated with NumPy: the cell loop
    :341-374 vectorised over tau for one frequency at a time, the 3x3 pseudo-inverses (:366) as one stacked
    np.linalg.pinv call (third-party, same function and default rcond as the reference)."""
    assert nproc == 1 and Neff_threshold >= 1
    ts, tau, freq = _c(ts), _c(tau), _c(freq)
    pd_ys = globals()["standardize"](ys)[0] if standardize else _c(ys)   # :334
    omega = make_omega(ts, freq)                                          # :336
    nt, nf = tau.size, freq.size
    Neffs = np.full((nt, nf), np.nan)
    y1 = np.full((nt, nf), np.nan)
    y2 = np.full((nt, nf), np.nan)
    y3 = np.full((nt, nf), np.nan)
    for k in range(nf):
        dz = omega[k] * (ts[None, :] - tau[:, None])                      # :343
        w = np.exp(-c * dz ** 2)                                          # :344
        sum_w = np.sum(w, axis=1)                                         # :346
        with np.errstate(invalid="ignore", divide="ignore"):
            Neffs[:, k] = sum_w ** 2 / np.sum(w ** 2, axis=1)             # :347
        ok = Neffs[:, k] > Neff_threshold                                 # :349 (NaN Neff: left NaN here)
        if not ok.any():
            continue
        w, dz, sw = w[ok], dz[ok], sum_w[ok, None]
        phi2, phi3 = np.cos(dz), np.sin(dz)                               # :354-355
        S = np.empty((w.shape[0], 3, 3))
        S[:, 0, 0] = 1
        S[:, 1, 1] = np.sum(w * phi2 * phi2, axis=1) / sw[:, 0]
        S[:, 2, 2] = np.sum(w * phi3 * phi3, axis=1) / sw[:, 0]
        S[:, 1, 0] = S[:, 0, 1] = np.sum(w * phi2, axis=1) / sw[:, 0]
        S[:, 2, 0] = S[:, 0, 2] = np.sum(w * phi3, axis=1) / sw[:, 0]
        S[:, 2, 1] = S[:, 1, 2] = np.sum(w * phi2 * phi3, axis=1) / sw[:, 0]
        S_inv = np.linalg.pinv(S)                                         # :366
        p = np.stack([np.sum(w * pd_ys, axis=1), np.sum(w * phi2 * pd_ys, axis=1),
                      np.sum(w * phi3 * pd_ys, axis=1)], axis=1) / sw     # :368-370
        sol = np.einsum("cij,cj->ci", S_inv, p)                           # :372-374
        y1[ok, k], y2[ok, k], y3[ok, k] = sol[:, 0], sol[:, 1], sol[:, 2]
    wwa = np.sqrt(y2 ** 2 + y3 ** 2)                                      # :376
    phase = np.arctan2(y3, y2)                                            # :377
    return wwa, phase, Neffs, (y1, y2, y3)


def wwa2psd(wwa, ts, Neffs, freq=None, Neff_threshold=3, anti_alias=False, avgs=2):
    """utils/wavelet.py:1236-1292 with anti_alias off (the default on this path)."""
    if anti_alias:
        raise NotImplementedError("anti_alias is off on the hot path")
    wwa, Neffs, ts = _c(wwa), _c(Neffs), _c(ts)
    nt, nf = wwa.shape
    psd = np.empty(nf)
    _lib().oracle_wwa2psd(_p(wwa), _p(Neffs), nt, nf, float(np.max(ts) - np.min(ts)), ts.size,
                          float(Neff_threshold), _p(psd))
    return psd


# --------------------------------------------------------------------------- host-side front end
def clean_ts(ys, ts):
    """utils/tsbase.py:316-352: dropna -> sort -> average duplicated timestamps."""
    ys = np.asarray(ys, dtype=float)
    ts = np.asarray(ts, dtype=float)
    keep = ~(np.isnan(ys) | np.isnan(ts))
    ys, ts = ys[keep], ts[keep]
    order = np.argsort(ts, kind="stable")
    ys, ts = ys[order], ts[order]
    if ts.size and np.any(np.diff(ts) == 0):
        uts, inv = np.unique(ts, return_inverse=True)
        sums = np.zeros(uts.size)
        cnts = np.zeros(uts.size)
        np.add.at(sums, inv, ys)
        np.add.at(cnts, inv, 1)
        ys, ts = sums / cnts, uts
    return ys, ts


def freq_vector_log(ts, fmin=None, fmax=None, nf=None):
    """utils/wavelet.py:1935-1984."""
    nt = np.size(ts)
    dt = np.median(np.diff(ts))
    fs = 1 / dt
    if nf is None:
        nf = nt // 10 + 1
    if fmin is None:
        fmin = 1 / np.ptp(ts)
    if fmax is None:
        fmax = fs / 2
    return np.logspace(np.log2(fmin), np.log2(fmax), nf, base=2)


def make_freq_vector(ts, method="log", rayleigh_bound=True, **kwargs):
    """utils/wavelet.py:1986-2049, 'log' method only (the default on this path)."""
    if method != "log":
        raise NotImplementedError("oracle covers freq_method='log' only")
    fr = freq_vector_log(ts, **kwargs)
    if rayleigh_bound:
        return fr[fr >= 1 / np.ptp(ts)]
    return fr


def prepare_wwz(ys, ts, freq=None, freq_method="log", freq_kwargs=None, tau=None):
    """utils/wavelet.py:2094-2202 with len_bd=0 (no ghost padding, the default)."""
    ys, ts = clean_ts(ys, ts)
    if tau is None:
        tau = np.linspace(np.min(ts), np.max(ts), np.min([np.size(ts), 50]))          # :2147-2149
    elif np.isnan(tau).any():
        tau = np.linspace(np.min(ts), np.max(ts), np.size(tau))                       # :2151-2155
    elif np.min(tau) < np.min(ts) and np.max(tau) > np.max(ts):
        tau = np.linspace(np.min(ts), np.max(ts), np.size(tau))                       # :2157-2162
    elif np.min(tau) not in ts or np.max(tau) not in ts:
        tau_lb = np.min(ts[ts > np.min(tau)])
        tau_ub = np.max(ts[ts < np.max(tau)])
        tau = np.linspace(tau_lb, tau_ub, np.size(tau))                               # :2164-2169
    sel = (np.min(tau) <= ts) & (ts <= np.max(tau))                                   # :2192-2193
    ts_cut, ys_cut = ts[sel], ys[sel]
    if freq is None:
        freq_kwargs = {} if freq_kwargs is None else dict(freq_kwargs)
        freq = make_freq_vector(ts_cut, method=freq_method, **freq_kwargs)            # :2195-2197
    freq = np.asarray(freq, dtype=float)
    freq = freq[freq != 0]                                                            # :2200
    return ys_cut, ts_cut, freq, np.asarray(tau, dtype=float)


WwzResults = collections.namedtuple(
    "Results", ["amplitude", "phase", "coi", "freq", "time", "Neffs", "coeff", "scale"])
PsdResults = collections.namedtuple("Results", ["psd", "freq"])


def wwz(ys, ts, tau=None, freq=None, freq_method="log", freq_kwargs=None, c=1 / (8 * np.pi ** 2),
        Neff_threshold=3, Neff_coi=3, standardize=True, mode=0, nthreads=0):
    """utils/wavelet.py:1294-1494 (body :1471-1494)."""
    ys_cut, ts_cut, freq, tau = prepare_wwz(ys, ts, freq=freq, freq_method=freq_method,
                                            freq_kwargs=freq_kwargs, tau=tau)
    wwa, phase, Neffs, coeff = kirchner(ys_cut, ts_cut, freq, tau, c=c, Neff_threshold=Neff_threshold,
                                        standardize=standardize, mode=mode, nthreads=nthreads)
    coi = make_coi(tau, Neff_threshold=Neff_coi)
    return WwzResults(amplitude=wwa, phase=phase, coi=coi, freq=freq, time=tau, Neffs=Neffs,
                      coeff=coeff, scale=1 / freq)


def wwz_psd(ys, ts, freq=None, freq_method="log", freq_kwargs=None, tau=None, c=1e-3,
            standardize=True, Neff_threshold=3, mode=0, nthreads=0):
    """utils/spectral.py:728-900.  As in the reference (:889-891) Neff_threshold is *not*
    forwarded to wwz (which uses its default 3) and only enters wwa2psd (:896)."""
    ys_cut, ts_cut, freq, tau = prepare_wwz(ys, ts, freq=freq, freq_method=freq_method,
                                            freq_kwargs=freq_kwargs, tau=tau)
    res = wwz(ys_cut, ts_cut, freq=freq, tau=tau, c=c, standardize=standardize, mode=mode,
              nthreads=nthreads)
    psd = wwa2psd(res.amplitude, ts_cut, res.Neffs, freq=res.freq, Neff_threshold=Neff_threshold)
    return PsdResults(psd=psd, freq=freq)


# --------------------------------------------------------------------------- surrogates
def ar1_model(t, tau, output_sigma=1.0, z=None, rng=None):
    """utils/tsmodel.py:35-69 with the standard-normal draws made explicit.

    ``z[i]`` is the draw consumed by step i (z[0] unused).  When ``z`` is None the draws
    come from ``rng`` (a ``np.random.RandomState``) or from NumPy's legacy global generator,
    which is what the reference consumes one value at a time (:66)."""
    t = _c(t)
    n = t.size
    if z is None:
        src = np.random if rng is None else rng
        z = np.concatenate([[0.0], src.standard_normal(n - 1)]) if n > 1 else np.zeros(n)
    z = _c(z)
    y = np.empty(n)
    _lib().oracle_ar1_model(_p(t), n, float(tau), float(output_sigma), _p(z), _p(y))
    return y


def tau_estimation(y, t):
    """utils/tsmodel.py:243-275: bounded scalar minimisation (scipy, third-party, same call
    as the reference) of the C objective oracle_ar1_objective (:268-271)."""
    from scipy import optimize
    y, t = _c(y), _c(t)
    lib = _lib()

    def ar1_fun(a):
        return lib.oracle_ar1_objective(_p(y), _p(t), y.size, float(a))

    a_est = optimize.minimize_scalar(ar1_fun, bounds=[0, 1], method="bounded").x
    return -1 / np.log(a_est)


def ar1_sim(y, p, t, rng=None):
    """utils/tsmodel.py:110-170, uneven-axis branch (:159-165): ysim[n, p]."""
    y, t = _c(y), _c(t)
    sig = np.std(y)
    tau_est = tau_estimation(y, t)
    ysim = np.empty((y.size, p))
    for i in range(p):
        ysim[:, i] = ar1_model(t, tau_est, output_sigma=sig, rng=rng)
    return ysim[:, 0] if p == 1 else ysim


def n_ll_unevenly_spaced_ar1(theta, y, t):
    """utils/tsmodel.py:612-644, term by term (Python ``sum`` over the arrays like the reference)."""
    n = len(y)
    log_tau, log_sigma_2 = theta[0], theta[1]
    tau = np.exp(log_tau)
    sigma_2 = np.exp(log_sigma_2)
    delta = np.diff(t)
    phi = np.exp((-delta / tau))
    term_1 = y[1:n] - (phi * y[0:(n - 1)])
    term_2 = pow(term_1, 2)
    term_3 = term_2 / (1 - pow(phi, 2))
    term_4 = 1 / (sigma_2 * n) * sum(term_3)
    term_5 = 1 / n * sum(np.log(1 - pow(phi, 2)))
    return np.log(2 * np.pi) + np.log(sigma_2) + term_5 + term_4


def uar1_fit(y, t):
    """utils/tsmodel.py:646-671 (scipy Nelder-Mead, third-party, same call and options as the reference)."""
    from scipy.optimize import minimize
    y, t = _c(y), _c(t)
    x0 = [np.log(tau_estimation(y, t)), np.log(np.var(y))]
    res = minimize(n_ll_unevenly_spaced_ar1, x0=x0, args=(y, t), method="nelder-mead",
                   options={"xatol": 1e-10, "disp": False, "maxiter": 1000})
    return np.exp(res.x)


def uar1_sim(t, tau, sigma_2=1, z=None):
    """utils/tsmodel.py:673-716 with the innovations made explicit: ``z`` has the shape of the reference's
    ``np.random.normal(size=(n, p))`` matrix (``[n]`` for a 1-D ``t``); None draws it from NumPy's legacy
    global generator exactly like the reference (:702)."""
    t = np.asarray(t, dtype=float)
    squeeze = t.ndim == 1
    if squeeze:
        t = t[:, np.newaxis]
    n, p = t.shape
    z = np.random.normal(loc=0, scale=1, size=(n, p)) if z is None else np.asarray(z, dtype=float).reshape(n, p)
    y = np.copy(z)
    for j in range(p):
        for i in range(1, n):
            delta_i = t[i, j] - t[i - 1, j]
            phi = np.exp(-delta_i / tau)
            sigma_i = np.sqrt(sigma_2 * (1 - phi ** 2))
            y[i, j] = phi * y[i - 1, j] + sigma_i * z[i, j]
    return np.squeeze(y) if squeeze or p == 1 else y


def mquantiles(a, prob, axis=0, alphap=0.4, betap=0.4):
    """scipy.stats.mstats.mquantiles (scipy 1.18.1, third-party; called at
    core/scalograms.py:641 and core/psds.py:913) restated for NaN-free data:
    aleph = n*p + m, m = alphap + p*(1-alphap-betap), k = floor(clip(aleph, 1, n-1)),
    gamma = clip(aleph-k, 0, 1), q = (1-gamma)*x[k-1] + gamma*x[k] on the sorted sample."""
    x = np.sort(np.asarray(a, dtype=float), axis=axis)
    x = np.moveaxis(x, axis, 0)
    n = x.shape[0]
    p = np.atleast_1d(np.asarray(prob, dtype=float))
    m = alphap + p * (1.0 - alphap - betap)
    aleph = n * p + m
    k = np.floor(aleph.clip(1, n - 1)).astype(int)
    gamma = (aleph - k).clip(0, 1)
    out = np.stack([(1.0 - g) * x[kk - 1] + g * x[kk] for kk, g in zip(k, gamma)], axis=0)
    return out


# --------------------------------------------------------------------------- GPU surrogate stream
def philox_normals(seed, series, n):
    """The standard-normal draws the CUDA surrogate generator consumes for one series, restated in
    NumPy: Philox4x32-10 (Salmon et al., SC'11; counter = (pair lo, pair hi, series lo, series hi),
    key = (seed lo, seed hi)), two 53-bit uniforms per counter, Box-Muller.  Step i (1 <= i < n) of the
    recurrence uses z0 of pair i//2 when i is even and z1 of pair i//2 when i is odd.
    Returns z[n]; z[0] (z0 of pair 0) is the start value of ``uar1_sim`` and is ignored by ``ar1_model``."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    npair = n // 2 + 1
    pair = np.arange(npair, dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    c0 = pair & mask
    c1 = pair >> np.uint64(32)
    c2 = np.full(npair, series & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(npair, (series >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF

    def u01(a, b):
        return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64) + 0.5) \
            / 9007199254740992.0

    u1, u2 = u01(c0, c1), u01(c2, c3)
    rad = np.sqrt(-2.0 * np.log(u1))
    z0, z1 = rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)
    i = np.arange(n)
    return np.where(i % 2 == 0, z0[i // 2], z1[i // 2])


# ------------------------------------------------------------------------------------------------
# wavelet coherence
# ------------------------------------------------------------------------------------------------
def xwt(coeff1, coeff2):
    """utils/wavelet.py:2201-2238: cross wavelet transform, its amplitude and phase."""
    prod = coeff1 * np.conj(coeff2)                                                       # :2229
    return prod, np.sqrt(prod.real ** 2 + prod.imag ** 2), np.arctan2(prod.imag, prod.real)   # :2230-2231


def boxcar_length(scales):
    """Length of the scale-smoothing window of wtc: dj = log2(scales[0] / scales[-1]) / N (utils/wavelet.py:2346-2349),
    window int(round(0.6 / dj * 2)) (:2331-2332)."""
    dj = np.log2(scales[0] / scales[-1]) / np.size(scales)
    return int(np.round(0.6 / dj * 2))


def _smooth(coeff, snorm, length, smooth_factor):
    """The `smoothing` closure of wtc (utils/wavelet.py:2288-2338): Gaussian in time through an FFT of the
    zero-padded rows, then a normalised boxcar over the scales."""
    from scipy import fft, signal
    rows = coeff.transpose()                                                              # :2313 [scale, time]
    n = rows.shape[1]
    npad = int(2 ** np.ceil(np.log2(n)))                                                  # :2310
    k2 = (2 * np.pi * fft.fftfreq(npad)) ** 2                                             # :2317-2318
    gauss = np.exp(-smooth_factor * (snorm[:, np.newaxis] ** 2) * k2)                     # :2323
    sm = fft.ifft(gauss * fft.fft(rows, axis=1, n=npad), axis=1, n=npad)[:, :n]           # :2324-2327
    if np.isreal(rows).all():                                                             # :2328-2329
        sm = sm.real
    win = np.zeros(length)                                                                # rect(), :2265-2286
    win[0] = win[-1] = 0.5
    win[1:-1] = 1
    win /= win.sum()
    return signal.convolve2d(sm, win[:, np.newaxis], "same").transpose()                  # :2334-2336


def wtc(coeff1, coeff2, scales, tau, smooth_factor=0.25):
    """utils/wavelet.py:2240-2364: wavelet transform coherency and its phase."""
    cross = coeff1 * np.conj(coeff2)                                                      # :2340
    p1, p2 = np.abs(coeff1) ** 2, np.abs(coeff2) ** 2                                     # :2341-2342
    snorm = scales / np.median(np.diff(tau))                                              # :2344-2345
    length = boxcar_length(scales)
    s12 = _smooth(cross / scales, snorm, length, smooth_factor)                           # :2350
    s1 = _smooth(p1 / scales, snorm, length, smooth_factor)                               # :2351
    s2 = _smooth(p2 / scales, snorm, length, smooth_factor)                               # :2352
    coherence = np.abs(s12) ** 2 / (s1 * s2)                                              # :2355
    return coherence, np.angle(s12 / (np.sqrt(s1) * np.sqrt(s2)))                         # :2356-2357


def wwz_coherence(y1, t1, y2, t2, smooth_factor=0.25, tau=None, freq=None, c=1 / (8 * np.pi ** 2), Neff_threshold=3,
                  standardize=True):
    """utils/wavelet.py:1660-1687 for the truncated series and the common grid that :1621-1658 arrive at (that part is
    the caller's business here): two transforms through wwz -- whose own prepare_wwz may repair tau per series --,
    coeff = a1 - i a2 (:1664-1665), then wtc, xwt and coi on the COMMON tau."""
    r1 = wwz(y1, t1, tau=tau, freq=freq, c=c, Neff_threshold=Neff_threshold, standardize=standardize)
    r2 = wwz(y2, t2, tau=tau, freq=freq, c=c, Neff_threshold=Neff_threshold, standardize=standardize)
    c1 = r1.coeff[1] - r1.coeff[2] * 1j
    c2 = r2.coeff[1] - r2.coeff[2] * 1j
    freq = np.asarray(freq, dtype=float)
    tau = np.asarray(tau, dtype=float)
    coh, ph = wtc(c1, c2, 1 / freq, tau, smooth_factor=smooth_factor)                     # :1669-1670
    prod, amp, _ = xwt(c1, c2)
    Res = collections.namedtuple("CoherenceResults", ["xw_coherence", "xw_amplitude", "xw_phase", "xwt", "freq", "time",
                                                      "coi", "scale"])
    return Res(xw_coherence=coh, xw_amplitude=amp, xw_phase=ph, xwt=prod, freq=freq, time=tau,
               coi=make_coi(tau, Neff_threshold=Neff_threshold), scale=1 / freq)
