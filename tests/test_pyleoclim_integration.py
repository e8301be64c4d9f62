"""Class-level integration with the *real* Pyleoclim (Series / MultipleSeries / Scalogram / PSD),
imported from the read-only reference tree through tests/golden/ref_shim.py.  CPU only: the batched
C-ABI calls are replaced by oracle-backed stand-ins so that the host logic of
pyleoclim_util_b200.dispatch (record / replay, object construction, fused signif_test) is checked
against the reference's own serial path.  Skipped on the GPU box (no reference tree there)."""
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_shim  # noqa: E402

if not ref_shim.available():
    pytest.skip("reference tree not mounted (expected on the GPU box)", allow_module_level=True)

from conftest import max_rel, assert_same_nan_mask  # noqa: E402
from oracle import wwz_oracle as O  # noqa: E402
from pyleoclim_util_b200 import batch as B, dispatch as D, wavelet as WV, synth  # noqa: E402


@pytest.fixture(scope="module")
def pyleo():
    return ref_shim.load_full()


@pytest.fixture()
def cuda_standins(monkeypatch, pyleo):
    """Oracle-backed replacements for the C-ABI calls (this container has no GPU)."""
    calls = {"shared": 0, "ragged": 0, "quantiles": 0, "signif": 0, "single": 0}

    def shared(Y, ts, tau, freq, c=1 / (8 * np.pi ** 2), Neff_threshold=3, standardize=True, outputs=("amplitude",),
               psd_Neff_threshold=None, flags=0):
        calls["shared"] += 1
        amps, psds, ne = [], [], None
        for y in Y:
            wwa, ph, ne, co = O.kirchner(y, ts, freq, tau, c=c, Neff_threshold=Neff_threshold, standardize=standardize)
            amps.append(wwa)
            if psd_Neff_threshold is not None:
                psds.append(O.wwa2psd(wwa, ts, ne, Neff_threshold=psd_Neff_threshold))
        return B.SharedAxisResults(amplitude=np.stack(amps), Neffs=ne, phase=None, coeff=None,
                                   psd=np.stack(psds) if psds else None, freq=freq, time=tau, coi=None, scale=1 / freq)

    def ragged(members, c=1 / (8 * np.pi ** 2), Neff_threshold=3, standardize=True, psd_Neff_threshold=None, flags=0):
        calls["ragged"] += 1
        res, psds = [], []
        for ys, ts, tau, freq in members:
            wwa, ph, ne, co = O.kirchner(ys, ts, freq, tau, c=c, Neff_threshold=Neff_threshold, standardize=standardize)
            res.append(WV.WwzResults(amplitude=wwa, phase=ph, coi=None, freq=freq, time=tau, Neffs=ne, coeff=co,
                                     scale=1 / freq))
            if psd_Neff_threshold is not None:
                psds.append(O.wwa2psd(wwa, ts, ne, Neff_threshold=psd_Neff_threshold))
        return (res, psds) if psd_Neff_threshold is not None else res

    def quantiles(X, qs):
        calls["quantiles"] += 1
        X = np.asarray(X)
        return O.mquantiles(X.reshape(X.shape[0], -1), qs, axis=0).reshape((len(qs),) + X.shape[1:])

    def signif(ys, ts, tau=None, freq=None, number=200, qs=(0.95,), seed=0, c=1 / (8 * np.pi ** 2), Neff_threshold=3,
               psd=False, psd_c=None, psd_Neff_threshold=3, **kw):
        calls["signif"] += 1
        calls["signif_method"] = kw.get("method", "ar1sim")
        rng = np.random.RandomState(seed % (2 ** 32))
        tau_est = O.tau_estimation(ys, ts)
        Y = np.stack([O.ar1_model(ts, tau_est, output_sigma=np.std(ys), rng=rng) + np.mean(ys) for _ in range(number)])
        r = shared(Y, ts, tau, freq, c=c, Neff_threshold=Neff_threshold, psd_Neff_threshold=psd_Neff_threshold if psd else None)
        return B.SignifResults(qs=np.asarray(qs), amplitude_qs=quantiles(r.amplitude, list(qs)),
                               psd_qs=quantiles(r.psd, list(qs)) if psd else None, Neffs=r.Neffs, freq=freq, time=tau)

    def single(pd_ys, ts, tau, omega, c, Neff_threshold, flags=0, psd_threshold=None, pinned=False, psd_only=False):
        calls["single"] += 1
        if flags & 64:                                            # WWZB200_FOSTER
            calls["foster"] = calls.get("foster", 0) + 1
            wwa, ph, ne, co = O.foster(pd_ys, ts, omega / (2 * np.pi), tau, c=c, Neff_threshold=Neff_threshold)
            psd = None if psd_threshold is None else O.wwa2psd(wwa, ts, ne, Neff_threshold=psd_threshold)
            return wwa, ph, ne, co, psd
        ne, a0, a1, a2 = O.kirchner_cells(ts, pd_ys, tau, omega, c, Neff_threshold)
        wwa, ph = np.sqrt(a1 ** 2 + a2 ** 2), np.arctan2(a2, a1)
        psd = None if psd_threshold is None else O.wwa2psd(wwa, ts, ne, Neff_threshold=psd_threshold)
        return wwa, ph, ne, (a0, a1, a2), psd

    monkeypatch.setattr(B, "wwz_shared_axis", shared)
    monkeypatch.setattr(B, "wwz_ragged", ragged)
    monkeypatch.setattr(B, "quantiles", quantiles)
    monkeypatch.setattr(B, "wwz_signif", signif)
    monkeypatch.setattr(WV, "_transform", single)
    D.register()
    yield calls
    D.unregister()


def _series(pyleo, n, seed, label):
    t, y = synth.uneven_ar1(n, seed)
    return pyleo.Series(time=t, value=3 * y + 1, label=label, time_unit="yr", verbose=False)


FREQ = np.array([0.005, 0.01, 0.03, 0.08, 0.2])
CUDA = {"method": "Kirchner_cuda", "ntau": 7}
REF = {"method": "Kirchner", "nproc": 1, "ntau": 7}


def test_series_wavelet_and_spectral_accept_the_method(pyleo, cuda_standins):
    s = _series(pyleo, 180, 1, "a")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = s.wavelet(method="wwz", settings=CUDA, freq=FREQ)
        ref = s.wavelet(method="wwz", settings=REF, freq=FREQ)
        pg = s.spectral(method="wwz", settings={"method": "Kirchner_cuda"}, freq=FREQ)
        pr = s.spectral(method="wwz", settings={"method": "Kirchner", "nproc": 1}, freq=FREQ)
    assert sorted(got.wave_args) == ["freq", "method", "ntau", "tau"] and got.wave_args["method"] == "Kirchner_cuda"
    assert max_rel(got.amplitude, ref.amplitude) < 1e-9 and max_rel(got.wwz_Neffs, ref.wwz_Neffs) < 1e-12
    assert np.array_equal(got.coi, ref.coi) and np.array_equal(got.time, ref.time)
    assert sorted(pg.spec_args) == ["freq", "method"]
    assert max_rel(pg.amplitude, pr.amplitude) < 1e-9
    assert cuda_standins["single"] == 2


def test_multipleseries_wavelet_ragged_batch(pyleo, cuda_standins):
    ms = pyleo.MultipleSeries([_series(pyleo, 150 + 10 * i, i, "s%d" % i) for i in range(3)])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = ms.wavelet(method="wwz", settings=CUDA, freq=FREQ)
        refs = [s.wavelet(method="wwz", settings=REF, freq=FREQ) for s in ms.series_list]
    assert cuda_standins["ragged"] == 1 and cuda_standins["shared"] == 0 and cuda_standins["single"] == 0
    assert type(got).__name__ == "MultipleScalogram" and len(got.scalogram_list) == 3
    for g, r, s in zip(got.scalogram_list, refs, ms.series_list):
        assert_same_nan_mask(g.amplitude, r.amplitude)
        assert max_rel(g.amplitude, r.amplitude) < 1e-9 and max_rel(g.wwz_Neffs, r.wwz_Neffs) < 1e-12
        assert g.timeseries is s and g.label == r.label and g.wave_method == "wwz" and g.freq_method == r.freq_method
        assert np.array_equal(g.frequency, r.frequency) and np.array_equal(g.time, r.time)


def test_multipleseries_shared_axis_batch_and_spectral(pyleo, cuda_standins):
    base = _series(pyleo, 160, 5, "target")
    members = []
    for i in range(4):
        m = base.copy()
        m.value = np.random.default_rng(i).standard_normal(base.value.size)
        members.append(m)
    ms = pyleo.MultipleSeries(members)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = ms.wavelet(method="wwz", settings=CUDA, freq=FREQ)
        refs = [s.wavelet(method="wwz", settings=REF, freq=FREQ) for s in members]
        pg = ms.spectral(method="wwz", settings={"method": "Kirchner_cuda", "Neff_threshold": 4}, freq=FREQ)
        prs = [s.spectral(method="wwz", settings={"method": "Kirchner", "nproc": 1, "Neff_threshold": 4}, freq=FREQ)
               for s in members]
    assert cuda_standins["shared"] == 2 and cuda_standins["ragged"] == 0
    for g, r in zip(got.scalogram_list, refs):
        assert max_rel(g.amplitude, r.amplitude) < 1e-9
    assert type(pg).__name__ == "MultiplePSD"
    for g, r in zip(pg.psd_list, prs):
        assert max_rel(g.amplitude, r.amplitude) < 1e-9 and g.spec_args["method"] == "Kirchner_cuda"
    # quantiles of CUDA results go to the (stand-in) device path and equal scipy's on the reference results
    q = got.quantiles(qs=[0.5, 0.9])
    qr = pyleo.MultipleScalogram(scalogram_list=refs).quantiles(qs=[0.5, 0.9])
    assert cuda_standins["quantiles"] == 1
    for a, b in zip(q.scalogram_list, qr.scalogram_list):
        assert max_rel(a.amplitude, b.amplitude) < 1e-9 and a.label == b.label
    qp = pg.quantiles(qs=[0.5, 0.9], lw=[1, 2])
    qpr = pyleo.MultiplePSD(psd_list=prs).quantiles(qs=[0.5, 0.9], lw=[1, 2])
    for a, b in zip(qp.psd_list, qpr.psd_list):
        assert max_rel(a.amplitude, b.amplitude) < 1e-9 and a.label == b.label and a.plot_kwargs == b.plot_kwargs


def test_other_methods_still_use_the_reference_path(pyleo, cuda_standins):
    s = _series(pyleo, 120, 2, "x")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = s.wavelet(method="wwz", settings=REF, freq=FREQ)
    assert cuda_standins["single"] == 0 and r.wave_args["method"] == "Kirchner"
    with pytest.raises(ValueError):
        s.wavelet(method="wwz", settings={"method": "no_such_method"}, freq=FREQ)


def test_fused_and_unfused_signif_test(pyleo, cuda_standins):
    s = _series(pyleo, 140, 9, "sig")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scal = s.wavelet(method="wwz", settings=CUDA, freq=FREQ)
        sig = scal.signif_test(number=6, qs=[0.9, 0.95], seed=3)
    assert cuda_standins["signif"] == 1
    assert type(sig.signif_qs).__name__ == "MultipleScalogram" and len(sig.signif_qs.scalogram_list) == 2
    assert [q.label for q in sig.signif_qs.scalogram_list] == ["90%", "95%"]
    for q in sig.signif_qs.scalogram_list:
        assert q.amplitude.shape == scal.amplitude.shape and np.array_equal(q.time, scal.time)
        assert np.array_equal(q.frequency, scal.frequency) and np.array_equal(q.coi, scal.coi)
    assert sig.signif_method == "ar1sim" and sig.qs == [0.9, 0.95] and sig is not scal
    assert np.all(sig.signif_qs.scalogram_list[1].amplitude >= sig.signif_qs.scalogram_list[0].amplitude)
    # export_scal=True needs the surrogate scalograms: reference flow, batched through MultipleSeries.wavelet
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sig2 = scal.signif_test(number=3, qs=[0.95], export_scal=True)
    assert cuda_standins["shared"] >= 1 and len(sig2.signif_scals.scalogram_list) == 3
    assert sig2.signif_qs.scalogram_list[0].amplitude.shape == scal.amplitude.shape
    # PSD significance, fused
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        psd = s.spectral(method="wwz", settings={"method": "Kirchner_cuda"}, freq=FREQ)
        ps = psd.signif_test(number=5, qs=[0.95])
    assert cuda_standins["signif"] == 2 and type(ps.signif_qs).__name__ == "MultiplePSD"
    assert ps.signif_qs.psd_list[0].amplitude.shape == psd.amplitude.shape and ps.signif_qs.psd_list[0].label == "95%"


def test_wwz_coherence_uses_the_registered_kernel(pyleo, cuda_standins):
    """wwz_coherence (utils/wavelet.py:1497-1687) makes two wwz() calls on a common grid
    (:1660-1665); with the dispatch patched both go through the CUDA kernel, the smoothing stays Pyleoclim's."""
    W = pyleo.utils.wavelet
    t1, y1 = synth.uneven_ar1(150, 1)
    t2, y2 = synth.uneven_ar1(150, 2)
    tau = np.linspace(max(t1.min(), t2.min()), min(t1.max(), t2.max()), 12)
    kw = dict(tau=tau, freq=FREQ, nproc=1, standardize=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = W.wwz_coherence(y1, t1, y2, t2, method="Kirchner_cuda", **kw)
        ref = W.wwz_coherence(y1, t1, y2, t2, method="Kirchner", **kw)
    assert cuda_standins["single"] == 2
    assert max_rel(got.xw_coherence, ref.xw_coherence) < 1e-8 and max_rel(got.xw_amplitude, ref.xw_amplitude) < 1e-8


def test_surrogates_option_routes_ar1_sim_to_the_device(pyleo, monkeypatch):
    calls = []

    def fake_ar1_sim(y, p, t, seed=0):
        calls.append((p, seed))
        return np.zeros((np.size(y), p)) if p > 1 else np.zeros(np.size(y))

    monkeypatch.setattr(B, "ar1_sim", fake_ar1_sim)
    D.register(surrogates=True)
    try:
        t, y = synth.uneven_ar1(80, 4)
        out = pyleo.utils.tsmodel.ar1_sim(y, 3, t)
        assert out.shape == (80, 3) and calls and calls[0][0] == 3
    finally:
        D.unregister()
    assert pyleo.utils.tsmodel.ar1_sim.__module__.endswith("tsmodel")


def test_uar1_signif_is_fused_and_uar1_sim_is_routed(pyleo, cuda_standins, monkeypatch):
    """method='uar1' (core/surrogateseries.py:147-156) takes the fused device pipeline, and
    register(surrogates=True) routes tsmodel.uar1_sim for a shared time axis to the device generator."""
    s = _series(pyleo, 120, 5, "u")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scal = s.wavelet(method="wwz", settings=CUDA, freq=FREQ)
        sig = scal.signif_test(method="uar1", number=4, qs=[0.95], seed=1)
    assert cuda_standins["signif"] == 1 and cuda_standins["signif_method"] == "uar1"
    assert sig.signif_method == "uar1" and sig.signif_qs.scalogram_list[0].amplitude.shape == scal.amplitude.shape
    D.unregister()
    calls = []

    def fake_uar1_sim(t, tau, sigma_2=1, seed=0):
        calls.append((np.shape(t), tau, sigma_2))
        return np.zeros(np.shape(t))

    monkeypatch.setattr(B, "uar1_sim", fake_uar1_sim)
    D.register(surrogates=True)
    try:
        t, _ = synth.uneven_ar1(60, 2)
        out = pyleo.utils.tsmodel.uar1_sim(np.tile(t, (4, 1)).T, tau=3.0, sigma_2=1.5)
        assert out.shape == (60, 4) and calls == [((60, 4), 3.0, 1.5)]
        T = np.stack([t, t * 1.01], axis=1)                      # per-column axes stay with the reference
        np.random.seed(0)
        assert pyleo.utils.tsmodel.uar1_sim(T, tau=3.0).shape == (60, 2) and len(calls) == 1
    finally:
        D.unregister()
    assert pyleo.utils.tsmodel.uar1_sim.__module__.endswith("tsmodel")


def test_foster_cuda_is_registered_next_to_foster(pyleo, cuda_standins):
    """method='Foster_cuda' goes through Pyleoclim's own wwz() front end (prepare_wwz,
    preprocess, make_omega, make_coi are the reference's) and matches method='Foster' (wwz_basic)."""
    W = pyleo.utils.wavelet
    t, y = synth.uneven_ar1(130, 6)
    tau = np.linspace(t.min(), t.max(), 9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = W.wwz(y, t, tau=tau, freq=FREQ, method="Foster_cuda", nproc=1)
        ref = W.wwz(y, t, tau=tau, freq=FREQ, method="Foster", nproc=1)
    assert cuda_standins.get("foster") == 1
    assert np.array_equal(np.isnan(got.amplitude), np.isnan(ref.amplitude))
    assert max_rel(got.amplitude, ref.amplitude) < 1e-8 and np.array_equal(got.coi, ref.coi)
    assert W.get_wwz_func(1, "Foster_cuda").__name__ == "foster_cuda"
    s = _series(pyleo, 120, 8, "f")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scal = s.wavelet(method="wwz", settings={"method": "Foster_cuda", "ntau": 6}, freq=FREQ)
    assert scal.wave_args["method"] == "Foster_cuda" and cuda_standins["foster"] == 2
