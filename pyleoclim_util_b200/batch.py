"""Batched entry points: AR(1) surrogates, shared-time-axis transforms, ragged ensembles, per-cell
quantiles and the fused significance test.  Thin NumPy wrappers over the C ABI (include/wwz_b200.h);
reference semantics are cited per function (paths relative to pyleoclim/).
"""
import collections

import numpy as np

from . import _lib
from . import wavelet as _wv

__all__ = ["tau_estimation", "ar1_surrogates", "ar1_sim", "n_ll_unevenly_spaced_ar1", "uar1_fit", "uar1_surrogates",
           "uar1_sim", "wwz_shared_axis", "wwz_ragged", "wwz_members", "quantiles",
           "wwz_signif", "signif_prepared", "coherence_signif", "SharedAxisResults", "SignifResults",
           "CoherenceSignifResults"]

SharedAxisResults = collections.namedtuple(
    "SharedAxisResults", ["amplitude", "Neffs", "phase", "coeff", "psd", "freq", "time", "coi", "scale"])
SignifResults = collections.namedtuple("SignifResults", ["qs", "amplitude_qs", "psd_qs", "Neffs", "freq", "time"])


def tau_estimation(y, t):
    """utils/tsmodel.py:243-275: persistence of an unevenly sampled AR(1) process (TAUEST),
    a bounded scalar minimisation on the host (scipy, the same third-party call as the reference)."""
    from scipy import optimize
    y = np.asarray(y, dtype=float)
    dt = np.diff(np.asarray(t, dtype=float))

    def ar1_fun(a):
        return np.sum((y[1:] - y[:-1] * a ** dt) ** 2)

    a_est = optimize.minimize_scalar(ar1_fun, bounds=[0, 1], method="bounded").x
    return -1 / np.log(a_est)


def ar1_surrogates(t, tau_est, sigma=1.0, mu=0.0, number=1, seed=0):
    """S realisations of ``ar1_model`` (utils/tsmodel.py:59-69) on the time axis ``t``, generated on
    the device (Philox4x32-10 + Box-Muller).  Returns ``Y[number, n]``; ``mu`` is added like
    ``SurrogateSeries.from_series`` does (core/surrogateseries.py:139)."""
    t = _lib.f64(t)
    Y = np.empty((int(number), t.size))
    _lib.check(_lib.lib().wwzb200_ar1_surrogates(_lib.ptr(t), t.size, float(tau_est), float(sigma), float(mu),
                                                 int(number), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(Y)))
    return Y


def ar1_sim(y, p, t=None, seed=0):
    """utils/tsmodel.py:110-170, uneven-axis branch: ``ysim[n, p]`` (``[n]`` when p == 1).
    The evenly spaced branch of the reference is a statsmodels ARIMA fit and is not on this path."""
    y = np.asarray(y, dtype=float)
    if t is None:
        raise ValueError("the CUDA surrogate generator needs the (uneven) time axis t")
    sig = np.std(y)
    tau_est = tau_estimation(y, t)
    ysim = ar1_surrogates(t, tau_est, sigma=sig, number=p, seed=seed).T
    return ysim[:, 0] if p == 1 else np.ascontiguousarray(ysim)


def n_ll_unevenly_spaced_ar1(theta, y, t):
    """utils/tsmodel.py:612-644: negative log-likelihood of an AR(1) process on an uneven axis,
    ``theta = (log tau, log sigma_2)``."""
    y = np.asarray(y, dtype=float)
    n = len(y)
    tau, sigma_2 = np.exp(theta[0]), np.exp(theta[1])
    phi = np.exp(-np.diff(np.asarray(t, dtype=float)) / tau)
    term_3 = (y[1:n] - phi * y[0:n - 1]) ** 2 / (1 - phi ** 2)
    term_4 = 1 / (sigma_2 * n) * sum(term_3)
    term_5 = 1 / n * sum(np.log(1 - phi ** 2))
    return np.log(2 * np.pi) + np.log(sigma_2) + term_5 + term_4


def uar1_fit(y, t):
    """utils/tsmodel.py:646-671: maximum-likelihood ``(tau, sigma_2)`` of an unevenly sampled AR(1) process
    (Nelder-Mead on the host from the TAUEST / variance starting point; scipy, the same third-party call as
    the reference)."""
    from scipy.optimize import minimize
    y = np.asarray(y, dtype=float)
    t = np.asarray(t, dtype=float)
    x0 = [np.log(tau_estimation(y=y, t=t)), np.log(np.var(y))]
    res = minimize(n_ll_unevenly_spaced_ar1, x0=x0, args=(y, t), method="nelder-mead",
                   options={"xatol": 1e-10, "disp": False, "maxiter": 1000})
    return np.exp(res.x)


def uar1_surrogates(t, tau, sigma_2=1.0, mu=0.0, number=1, seed=0):
    """S realisations of ``uar1_sim`` (utils/tsmodel.py:673-716) on the time axis ``t``, generated on the device:
    ``y[0] = z[0]``, ``y[i] = phi_i y[i-1] + sqrt(sigma_2 (1 - phi_i^2)) z[i]``.  Returns ``Y[number, n]``;
    ``mu`` is added like ``SurrogateSeries.from_series`` does (core/surrogateseries.py:156)."""
    t = _lib.f64(t)
    Y = np.empty((int(number), t.size))
    _lib.check(_lib.lib().wwzb200_uar1_surrogates(_lib.ptr(t), t.size, float(tau), float(sigma_2), float(mu),
                                                  int(number), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(Y)))
    return Y


def uar1_sim(t, tau, sigma_2=1, seed=0):
    """utils/tsmodel.py:673-716 for a time axis shared by all realisations: ``t`` is ``[n]`` (one series, returns
    ``[n]``) or ``[n, p]`` with identical columns (what SurrogateSeries.from_series passes, core/surrogateseries.py:154;
    returns ``[n, p]``).  Different axes per column are independent series: call once per axis."""
    t = np.asarray(t, dtype=float)
    if t.ndim == 1:
        return uar1_surrogates(t, tau, sigma_2, number=1, seed=seed)[0]
    if not (t == t[:, :1]).all():
        raise ValueError("the batched generator needs one time axis shared by all columns")
    return np.ascontiguousarray(uar1_surrogates(t[:, 0], tau, sigma_2, number=t.shape[1], seed=seed).T)


def wwz_shared_axis(Y, ts, tau, freq, c=1 / (8 * np.pi ** 2), Neff_threshold=3, standardize=True,
                    outputs=("amplitude",), psd_Neff_threshold=None, flags=0):
    """Transform ``S`` series that share the time axis ``ts`` in one batched launch
    (what ``MultipleSeries.wavelet`` does member by member through a process pool,
    core/multipleseries.py:1520-1532, for the surrogates of core/scalograms.py:515-517).

    Y : [S, n].  ``outputs`` selects among 'amplitude', 'phase', 'coeff'; ``psd_Neff_threshold`` adds
    the per-series ``wwa2psd`` (psd [S, nf]).  ts must already be clean (sorted, no NaN) and the grid
    final: call ``prepare_wwz`` first, as ``wwz`` does."""
    _wv.assertPositiveInt(Neff_threshold)
    ts, tau, freq = _lib.f64(ts), _lib.f64(tau), _lib.f64(freq)
    Y = _lib.f64(Y)
    if Y.ndim != 2 or Y.shape[1] != ts.size:
        raise ValueError("Y must be [S, len(ts)]")
    S, nt, nf = Y.shape[0], tau.size, freq.size
    omega = _wv.make_omega(ts, freq)
    Neffs = np.empty((nt, nf))
    want = set(outputs)
    amp = _lib.output_empty((S, nt, nf)) if "amplitude" in want else None
    phase = _lib.output_empty((S, nt, nf)) if "phase" in want else None
    a0 = a1 = a2 = None
    if "coeff" in want:
        a0, a1, a2 = (_lib.output_empty((S, nt, nf)) for _ in range(3))
    psd = np.empty((S, nf)) if psd_Neff_threshold is not None else None
    fl = int(flags) | (_lib.STANDARDIZE if standardize else 0)
    _lib.check(_lib.lib().wwzb200_kirchner_shared_axis(
        _lib.ptr(ts), _lib.ptr(Y), ts.size, S, _lib.ptr(tau), nt, _lib.ptr(omega), nf, float(c),
        int(Neff_threshold), int(psd_Neff_threshold or 0), fl, _lib.ptr(Neffs), _lib.ptr(amp), _lib.ptr(phase),
        _lib.ptr(a0), _lib.ptr(a1), _lib.ptr(a2), _lib.ptr(psd)))
    coi = _wv.make_coi(tau) if nt > 1 else None
    return SharedAxisResults(amplitude=amp, Neffs=Neffs, phase=phase, coeff=(a0, a1, a2) if a0 is not None else None,
                             psd=psd, freq=freq, time=tau, coi=coi, scale=1 / freq)


def _cois(tau):
    """make_coi (utils/wavelet.py:1168-1208) of every row of tau [M, nt] at once."""
    nt0 = tau.shape[1]
    efold = 4 * np.pi / (3 + np.sqrt(2 + 3 ** 2)) / np.sqrt(2)
    step = np.median(np.diff(tau, axis=1), axis=1)
    idx = np.arange(nt0)
    tent = np.minimum(idx, nt0 - 1 - idx).astype(float)
    tent[tent == 0] = 0.00001
    return efold * step[:, None] * tent[None, :]


def wwz_members(T, Y, tau, freq, c=1 / (8 * np.pi ** 2), Neff_threshold=3, standardize=True, psd_Neff_threshold=None,
                flags=0):
    """M members of EQUAL shape -- an age ensemble (core/ensembleseries.py:47-175), which the reference transforms
    member by member in a process pool (core/multipleseries.py:1520-1532) -- through ONE C call
    (``wwzb200_kirchner_members``): z-score, Nyquist clipping of the frequencies (make_omega), setup, basis rows,
    cell kernels, epilogue and wwa2psd each run once for all members of a group on the device, and finished groups
    cross PCIe under the next group's kernels.

    T [M, n] time axes (prepared: sorted, no NaN); Y [M, n] values, or [n] when all members share them; tau [M, nt];
    freq [M, nf].  Returns a list of ``wavelet.wwz``-style Results whose arrays are views of one page-locked block
    (plus the list of psd vectors when ``psd_Neff_threshold`` is given)."""
    _wv.assertPositiveInt(Neff_threshold)
    T, tau, freq, Y = _lib.f64(T), _lib.f64(tau), _lib.f64(freq), _lib.f64(Y)
    if T.ndim != 2 or tau.ndim != 2 or freq.ndim != 2 or not (T.shape[0] == tau.shape[0] == freq.shape[0]):
        raise ValueError("T, tau and freq must be [M, n], [M, nt], [M, nf]")
    M, n = T.shape
    nt, nf = tau.shape[1], freq.shape[1]
    if Y.ndim == 1:
        if Y.size != n:
            raise ValueError("shared values must have the length of a time axis")
        ys_stride = 0
    elif Y.shape == (M, n):
        ys_stride = n
    else:
        raise ValueError("Y must be [n] or [M, n]")
    fl = int(flags) | (_lib.STANDARDIZE if standardize else 0)
    if n - 1 <= 16384 and n >= 2:
        fl |= _lib.FREQ_HZ                      # make_omega on the device
        om = freq
    else:
        nyq = 0.5 / np.median(np.diff(T, axis=1), axis=1)
        om = _lib.f64(2 * np.pi * np.where(freq > nyq[:, None], np.nan, freq))
    out = _lib.output_empty((6, M, nt, nf))
    psd = np.empty((M, nf)) if psd_Neff_threshold is not None else None
    stats = np.empty((M, 2)) if standardize else None
    _lib.check(_lib.lib().wwzb200_kirchner_members(
        _lib.ptr(T), _lib.ptr(Y), ys_stride, M, n, _lib.ptr(tau), nt, _lib.ptr(om), nf, float(c), int(Neff_threshold),
        int(psd_Neff_threshold or 0), fl, _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(out[3]),
        _lib.ptr(out[4]), _lib.ptr(out[5]), _lib.ptr(psd), _lib.ptr(stats)))
    if stats is not None and ((stats[:, 0] == 0) & (stats[:, 1] == 1)).any():
        import warnings
        warnings.warn("Constant or nearly constant time series not rescaled.", stacklevel=2)
    coi = _cois(tau) if nt > 1 else [None] * M
    scale = 1 / freq
    Ne, a0, a1, a2, wwa, ph = out
    res = [_wv.WwzResults(amplitude=wwa[m], phase=ph[m], coi=coi[m], freq=freq[m], time=tau[m], Neffs=Ne[m],
                          coeff=(a0[m], a1[m], a2[m]), scale=scale[m]) for m in range(M)]
    if psd is not None:
        return res, list(psd)
    return res


def wwz_ragged(members, c=1 / (8 * np.pi ** 2), Neff_threshold=3, standardize=True, psd_Neff_threshold=None,
               flags=0):
    """Transform M series with their own time axes and grids (age ensembles:
    core/ensembleseries.py:169-173 -> core/multipleseries.py:1520-1532) in one call.

    members : iterable of ``(ys, ts, tau, freq)`` (already prepared with ``prepare_wwz``).
    Returns a list of ``wavelet.wwz``-style Results (plus ``.psd`` lists when requested)."""
    _wv.assertPositiveInt(Neff_threshold)
    members = list(members)
    ts_l = [_lib.f64(m[1]) for m in members]
    tau_l = [_lib.f64(m[2]) for m in members]
    fr_l = [_lib.f64(m[3]) for m in members]
    nfs = [f.size for f in fr_l]
    if len(members) > 1 and not (int(flags) & (_lib.FOSTER | _lib.FAITHFUL)) and ts_l[0].size > 1 \
            and len({t.size for t in ts_l}) == 1 and len({a.size for a in tau_l}) == 1 and tau_l[0].size \
            and min(nfs) > 0 and max(nfs) - min(nfs) <= max(2, max(nfs) // 8) \
            and all(np.ndim(m[0]) == 1 and np.size(m[0]) == ts_l[0].size for m in members):
        # equally shaped members (an age ensemble on its default grids: same n, same ntau, frequency axes that differ by
        # a point or two): one C call, every kernel launched once per group.  Shorter frequency axes are padded with NaN
        # (a NaN frequency gives a NaN column, like a frequency above Nyquist) and the padding is sliced off again.
        nf = max(nfs)
        M_, n_, nt_ = len(members), ts_l[0].size, tau_l[0].size
        # packed inputs in page-locked (pooled) buffers: one concatenate per matrix (np.stack on a thousand small arrays
        # costs more than the transform)
        T, TAU = _lib.output_empty((M_, n_)), np.empty((M_, nt_))
        np.concatenate(ts_l, out=T.reshape(-1))
        np.concatenate(tau_l, out=TAU.reshape(-1))
        if min(nfs) == nf:
            F = np.empty((M_, nf))
            np.concatenate(fr_l, out=F.reshape(-1))
        else:
            F = np.full((M_, nf), np.nan)
            for m in range(M_):
                F[m, : nfs[m]] = fr_l[m]
        first = members[0][0]
        if all(m[0] is first for m in members):
            Yv = _lib.f64(first)                        # one value vector shared by all members (ys_stride = 0)
        else:
            Yv = _lib.output_empty((M_, n_))
            np.concatenate([_lib.f64(m[0]) for m in members], out=Yv.reshape(-1))
        out = wwz_members(T, Yv, TAU, F, c=c, Neff_threshold=Neff_threshold, standardize=standardize,
                          psd_Neff_threshold=psd_Neff_threshold, flags=flags)
        res, psds = out if psd_Neff_threshold is not None else (out, None)
        if min(nfs) != nf:
            for m, k in enumerate(nfs):
                if k != nf:
                    r = res[m]
                    res[m] = _wv.WwzResults(amplitude=r.amplitude[:, :k], phase=r.phase[:, :k], coi=r.coi, freq=fr_l[m],
                                            time=r.time, Neffs=r.Neffs[:, :k], coeff=tuple(x[:, :k] for x in r.coeff),
                                            scale=1 / fr_l[m])
                    if psds is not None:
                        psds[m] = psds[m][:k]
        return (res, psds) if psds is not None else res
    same_n = len(members) > 1 and len({t.size for t in ts_l}) == 1 and ts_l[0].size > 1 \
        and all(np.ndim(m[0]) == 1 and np.size(m[0]) == ts_l[0].size for m in members)
    if same_n:
        # equal-length members (an age ensemble): z-scores and Nyquist frequencies of all members in one pass each,
        # row by row the same arithmetic as wavelet.standardize / wavelet.make_omega
        T = np.stack(ts_l)
        Ymat = np.stack([_lib.f64(m[0]) for m in members])
        if standardize:
            mu, sig = Ymat.mean(axis=1), Ymat.std(axis=1)
            if np.isnan(sig).any():
                mu, sig = np.nanmean(Ymat, axis=1), np.nanstd(Ymat, axis=1)
            flat = np.abs(sig) < 1e-3
            if flat.any():
                import warnings
                warnings.warn("Constant or nearly constant time series not rescaled.", stacklevel=2)
            Ymat = (Ymat - np.where(flat, 0.0, mu)[:, None]) / np.where(flat, 1.0, sig)[:, None]
        ys_l = list(Ymat)
        nyq = 0.5 / np.median(np.diff(T, axis=1), axis=1)
        om_l = [2 * np.pi * np.where(f > q, np.nan, f) for f, q in zip(fr_l, nyq)]
    else:
        ys_l, om_l = [], []
        for (ys, _, _, _), ts, fr in zip(members, ts_l, fr_l):
            ys_l.append(_lib.f64(_wv.preprocess(_lib.f64(ys), ts, standardize=standardize)))
            om_l.append(_wv.make_omega(ts, fr))
    M = len(ts_l)
    off = lambda xs: np.concatenate([[0], np.cumsum([x.size for x in xs])]).astype(np.int64)
    ts_off, tau_off, om_off = off(ts_l), off(tau_l), off(om_l)
    out_off = np.concatenate([[0], np.cumsum([a.size * b.size for a, b in zip(tau_l, om_l)])]).astype(np.int64)
    cat = lambda xs: np.concatenate(xs) if xs else np.empty(0)
    ts_p, ys_p, tau_p, om_p = cat(ts_l), cat(ys_l), cat(tau_l), cat(om_l)
    outs = [_lib.output_empty((int(out_off[-1]),)) for _ in range(6)]
    psd = np.empty(int(om_off[-1])) if psd_Neff_threshold is not None else None
    _lib.check(_lib.lib().wwzb200_kirchner_ragged(
        _lib.ptr(ts_p), _lib.ptr(ys_p), _lib.iptr(ts_off), _lib.ptr(tau_p), _lib.iptr(tau_off), _lib.ptr(om_p),
        _lib.iptr(om_off), _lib.iptr(out_off), M, float(c), int(Neff_threshold), int(psd_Neff_threshold or 0),
        int(flags), *[_lib.ptr(o) for o in outs], _lib.ptr(psd)))
    res = []
    coi_l = None
    if M > 1 and len({a.size for a in tau_l}) == 1 and tau_l[0].size > 1:
        # equal-length tau axes: the cones of influence of all members at once (utils/wavelet.py:1168-1208)
        nt0 = tau_l[0].size
        efold = 4 * np.pi / (3 + np.sqrt(2 + 3 ** 2)) / np.sqrt(2)
        step = np.median(np.diff(np.stack(tau_l), axis=1), axis=1)
        idx = np.arange(nt0)
        tent = np.minimum(idx, nt0 - 1 - idx).astype(float)
        tent[tent == 0] = 0.00001
        coi_l = efold * step[:, None] * tent[None, :]
    for m in range(M):
        nt, nf = tau_l[m].size, om_l[m].size
        sl = slice(int(out_off[m]), int(out_off[m + 1]))
        Neffs, a0, a1, a2, wwa, phase = (o[sl].reshape(nt, nf) for o in outs)
        coi = coi_l[m] if coi_l is not None else (_wv.make_coi(tau_l[m]) if nt > 1 else None)
        r = _wv.WwzResults(amplitude=wwa, phase=phase, coi=coi, freq=fr_l[m],
                           time=tau_l[m], Neffs=Neffs, coeff=(a0, a1, a2), scale=1 / fr_l[m])
        res.append(r)
    if psd is not None:
        return res, [psd[int(om_off[m]):int(om_off[m + 1])] for m in range(M)]
    return res


def quantiles(X, qs):
    """Per-cell quantiles over a batch with ``scipy.stats.mstats.mquantiles(alphap=betap=0.4)``
    semantics (MultipleScalogram.quantiles core/scalograms.py:635-641, MultiplePSD.quantiles
    core/psds.py:913).  X : [S, ...]; returns [len(qs), ...]."""
    X = _lib.f64(X)
    S = X.shape[0]
    cell_shape = X.shape[1:]
    ncell = int(np.prod(cell_shape, dtype=np.int64))
    probs = _lib.f64(np.atleast_1d(qs))
    q = np.empty((probs.size, ncell))
    _lib.check(_lib.lib().wwzb200_quantiles(_lib.ptr(X.reshape(S, ncell)), S, ncell, _lib.ptr(probs), probs.size,
                                            _lib.ptr(q)))
    return q.reshape((probs.size,) + cell_shape)


def wwz_signif(ys, ts, tau=None, freq=None, number=200, qs=(0.95,), seed=0, c=1 / (8 * np.pi ** 2),
               Neff_threshold=3, psd=False, psd_c=None, psd_Neff_threshold=3, freq_method="log", freq_kwargs=None,
               flags=0, method="ar1sim"):
    """AR(1)-surrogate significance levels of a WWZ scalogram (and optionally of the WWZ spectrum) with
    everything between the time axis and the quantiles on the device (C-ABI ``wwzb200_ar1_signif``).

    Follows Scalogram.signif_test (core/scalograms.py:515-522): surrogates of the target from
    ``ar1_sim`` (+ target mean, core/surrogateseries.py:137-139), each transformed like the target
    (standardised, same tau/freq), per-cell ``mquantiles`` over the surrogates.  ``number`` surrogates,
    probabilities ``qs``.  With ``psd=True`` the quantiles of the surrogate spectra are returned too
    (PSD.signif_test core/psds.py:246-282), computed with decay constant ``psd_c`` (default ``c``).
    ``method='uar1'`` draws the surrogates from the maximum-likelihood AR(1) fit instead
    (core/surrogateseries.py:147-156: uar1_fit + uar1_sim)."""
    ys_cut, ts_cut, freq, tau = _wv.prepare_wwz(ys, ts, freq=freq, freq_method=freq_method, freq_kwargs=freq_kwargs,
                                                tau=tau)
    probs = _lib.f64(np.atleast_1d(qs))
    cc = float(c if not psd or psd_c is None else psd_c)
    q_wwa, q_psd, Neffs = signif_prepared(ys_cut, ts_cut, tau, freq, number=number, probs=probs, seed=seed, c=cc,
                                          Neff_threshold=Neff_threshold, psd=psd,
                                          psd_Neff_threshold=psd_Neff_threshold, flags=flags, method=method)
    return SignifResults(qs=probs, amplitude_qs=q_wwa, psd_qs=q_psd, Neffs=Neffs, freq=freq, time=tau)


def signif_prepared(ys_cut, ts_cut, tau, freq, number, probs, seed=0, c=1 / (8 * np.pi ** 2), Neff_threshold=3,
                    psd=False, psd_Neff_threshold=3, flags=0, method="ar1sim"):
    """The fused significance pipeline on an already prepared series / grid (no prepare_wwz: `tau`
    may be a row block of the full grid, as in distributed.wwz_signif_tau_sharded).
    Returns (q_wwa [nq, nt, nf], q_psd [nq, nf] or None, Neffs [nt, nf])."""
    ts_cut, tau, freq = _lib.f64(ts_cut), _lib.f64(tau), _lib.f64(freq)
    ys_cut = np.asarray(ys_cut, dtype=float)
    mu = float(np.mean(ys_cut))                     # core/surrogateseries.py:137-139 (removed again by the z-score)
    if method == "ar1sim":
        sig = float(np.std(ys_cut))                     # utils/tsmodel.py:147
        tau_est = float(tau_estimation(ys_cut, ts_cut))  # utils/tsmodel.py:161
    elif method == "uar1":
        tau_est, sig = (float(v) for v in uar1_fit(ys_cut, ts_cut))   # core/surrogateseries.py:149; sig = sigma_2
        flags = int(flags) | _lib.UAR1
    else:
        raise ValueError("surrogate method must be 'ar1sim' or 'uar1'")
    omega = _wv.make_omega(ts_cut, freq)
    probs = _lib.f64(np.atleast_1d(probs))
    nt, nf = tau.size, freq.size
    Neffs = np.empty((nt, nf))
    q_wwa = np.empty((probs.size, nt, nf))
    q_psd = np.empty((probs.size, nf)) if psd else None
    _lib.check(_lib.lib().wwzb200_ar1_signif(
        _lib.ptr(ts_cut), ts_cut.size, tau_est, sig, mu, int(number), int(seed) & 0xFFFFFFFFFFFFFFFF,
        _lib.ptr(tau), nt, _lib.ptr(omega), nf, float(c), int(Neff_threshold), int(psd_Neff_threshold), int(flags),
        _lib.ptr(probs), probs.size, _lib.ptr(Neffs), _lib.ptr(q_wwa), _lib.ptr(q_psd)))
    return q_wwa, q_psd, Neffs


CoherenceSignifResults = collections.namedtuple("CoherenceSignifResults", ["qs", "wtc_qs", "xwt_qs", "freq", "time", "coi"])


def coherence_signif(y1, t1, y2, t2, tau=None, freq=None, number=200, qs=(0.95,), seed=0, smooth_factor=0.25,
                     c=1 / (8 * np.pi ** 2), Neff_threshold=3, method="ar1sim", freq_method="log", freq_kwargs=None):
    """Significance levels of a WWZ wavelet coherence from AR(1) surrogate PAIRS, everything between the two time axes
    and the quantiles on the device.

    Follows Coherence.signif_test (core/coherences.py:748-963): `number` surrogates of each series on its own time
    axis (core/surrogateseries.py:90-181), for every pair the wavelet coherence of core/series.py:3586-3800 ->
    utils/wavelet.py:1497-1687 on the grid of the original analysis, then per-cell ``mquantiles`` of the coherence
    (wtc) and of the cross-wavelet amplitude (xwt) over the pairs.  Where the reference runs one process-pool task per
    pair, here the surrogates of a series are ONE shared-axis batch (coefficients a1, a2 kept on the device), the
    coherency smoothing of all pairs is one launch sequence (``wwzb200_wtc_dev``) and only the [nq, ntau, nf]
    quantiles return.  The two surrogate sets draw from independent Philox streams (seed, seed + 1).
    Device memory: about 10 x number x ntau x nf doubles."""
    import ctypes
    import torch
    (y1c, t1c, tau1), (y2c, t2c, tau2), tau, freq = _wv.coherence_inputs(
        np.asarray(y1, dtype=float), np.asarray(t1, dtype=float), np.asarray(y2, dtype=float), np.asarray(t2, dtype=float),
        tau=tau, freq=freq, freq_method=freq_method, freq_kwargs=freq_kwargs)
    _wv.assertPositiveInt(Neff_threshold)
    L_box = _wv.wtc_boxcar_length(freq)
    if L_box < 1:
        raise IndexError("the scale-smoothing window of wtc is empty for this frequency axis")
    S, nt, nf = int(number), tau.size, freq.size
    ncell = nt * nf
    probs = _lib.f64(np.atleast_1d(qs))
    nq = probs.size
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    d_freq, d_probs = up(freq), up(probs)
    coeff = []
    for k, (yc, tc, tau_k) in enumerate(((y1c, t1c, tau1), (y2c, t2c, tau2))):
        d_tau = up(tau_k)               # each series on its own prepared tau (see wavelet.coherence_inputs)
        mu = float(np.mean(yc))                                          # core/surrogateseries.py:137-139
        flags = _lib.STANDARDIZE
        if method == "ar1sim":
            sig, tau_est = float(np.std(yc)), float(tau_estimation(yc, tc))
        elif method == "uar1":
            tau_est, sig = (float(v) for v in uar1_fit(yc, tc))
            flags |= _lib.UAR1
        else:
            raise ValueError("surrogate method must be 'ar1sim' or 'uar1'")
        d_ts, d_om = up(tc), up(_wv.make_omega(tc, freq))
        d_Y = torch.empty((S, tc.size), dtype=torch.float64, device=dev)
        _lib.check(L.wwzb200_surrogates_block_dev(d_ts.data_ptr(), tc.size, tau_est, sig, mu, S, 0,
                                                  (int(seed) + k) & 0xFFFFFFFFFFFFFFFF, flags & _lib.UAR1, d_Y.data_ptr(), st))
        d_a1 = torch.empty((S, nt, nf), dtype=torch.float64, device=dev)
        d_a2 = torch.empty((S, nt, nf), dtype=torch.float64, device=dev)
        d_ne = torch.empty((nt, nf), dtype=torch.float64, device=dev)
        _lib.check(L.wwzb200_kirchner_shared_axis_dev(d_ts.data_ptr(), d_Y.data_ptr(), tc.size, S, d_tau.data_ptr(), nt,
                                                      d_om.data_ptr(), nf, float(c), int(Neff_threshold), 0,
                                                      flags & ~_lib.UAR1, d_ne.data_ptr(), None, None, None, d_a1.data_ptr(),
                                                      d_a2.data_ptr(), None, st))
        coeff.append((d_a1, d_a2))
        del d_Y
    d_coh = torch.empty((S, nt, nf), dtype=torch.float64, device=dev)
    d_xwa = torch.empty((S, nt, nf), dtype=torch.float64, device=dev)
    _lib.check(L.wwzb200_wtc_dev(coeff[0][0].data_ptr(), coeff[0][1].data_ptr(), coeff[1][0].data_ptr(), coeff[1][1].data_ptr(),
                                 S, nt, nf, d_freq.data_ptr(), float(np.median(np.diff(tau))), float(smooth_factor), L_box,
                                 d_coh.data_ptr(), None, d_xwa.data_ptr(), st))
    del coeff
    d_q = torch.empty((2, nq, ncell), dtype=torch.float64, device=dev)
    _lib.check(L.wwzb200_quantiles_dev(d_coh.data_ptr(), S, ncell, d_probs.data_ptr(), nq, d_q[0].data_ptr(), st))
    _lib.check(L.wwzb200_quantiles_dev(d_xwa.data_ptr(), S, ncell, d_probs.data_ptr(), nq, d_q[1].data_ptr(), st))
    q = d_q.cpu().numpy().reshape(2, nq, nt, nf)
    return CoherenceSignifResults(qs=probs, wtc_qs=q[0], xwt_qs=q[1], freq=freq, time=tau,
                                  coi=_wv.make_coi(tau, Neff_threshold=Neff_threshold))
