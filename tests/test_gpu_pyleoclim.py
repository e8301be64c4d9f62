"""The class-level path on hardware: the REAL Pyleoclim classes (imported through tests/golden/ref_shim.py from the
reference tree, or from the offline install under baseline/_ref that travels to the GPU box) driving the REAL
libwwz_b200.so.  The test bodies are those of tests/test_pyleoclim_integration.py -- there the C-ABI calls are replaced
by oracle-backed stand-ins because the build container has no GPU; here the `cuda_standins` fixture only counts the
calls and lets them through to the device.  Reference behaviour: core/series.py:3415-3584 (Series.wavelet),
core/multipleseries.py:1436-1534, core/scalograms.py:401-562 (signif_test), core/ensembleseries.py:47-175."""
import os
import sys
import time
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.gpu

if not ref_shim.available():
    pytest.skip("no reference tree (/root/reference, baseline/_ref or $WWZ_REFERENCE_ROOT)", allow_module_level=True)

import test_pyleoclim_integration as T  # noqa: E402
from conftest import max_rel, assert_same_nan_mask, has_cuda  # noqa: E402
from pyleoclim_util_b200 import batch as B, dispatch as D, wavelet as WV, synth  # noqa: E402

pyleo = T.pyleo


@pytest.fixture()
def cuda_standins(monkeypatch, pyleo):
    """Counting pass-throughs to the real library (same keys as the CPU stand-ins)."""
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    calls = {"shared": 0, "ragged": 0, "quantiles": 0, "signif": 0, "single": 0}
    real = dict(shared=B.wwz_shared_axis, ragged=B.wwz_ragged, quantiles=B.quantiles, signif=B.wwz_signif,
                single=WV._transform)

    def counted(key):
        def f(*a, **k):
            calls[key] += 1
            if key == "signif":
                calls["signif_method"] = k.get("method", "ar1sim")
            if key == "single" and (k.get("flags", a[6] if len(a) > 6 else 0) & 64):
                calls["foster"] = calls.get("foster", 0) + 1
            return real[key](*a, **k)
        return f

    monkeypatch.setattr(B, "wwz_shared_axis", counted("shared"))
    monkeypatch.setattr(B, "wwz_ragged", counted("ragged"))
    monkeypatch.setattr(B, "quantiles", counted("quantiles"))
    monkeypatch.setattr(B, "wwz_signif", counted("signif"))
    monkeypatch.setattr(WV, "_transform", counted("single"))
    D.register()
    yield calls
    D.unregister()


test_series_wavelet_and_spectral_accept_the_method = T.test_series_wavelet_and_spectral_accept_the_method
test_multipleseries_wavelet_ragged_batch = T.test_multipleseries_wavelet_ragged_batch
test_multipleseries_shared_axis_batch_and_spectral = T.test_multipleseries_shared_axis_batch_and_spectral
test_other_methods_still_use_the_reference_path = T.test_other_methods_still_use_the_reference_path
test_fused_and_unfused_signif_test = T.test_fused_and_unfused_signif_test
test_wwz_coherence_uses_the_registered_kernel = T.test_wwz_coherence_uses_the_registered_kernel
test_uar1_signif_is_fused_and_uar1_sim_is_routed = T.test_uar1_signif_is_fused_and_uar1_sim_is_routed
test_foster_cuda_is_registered_next_to_foster = T.test_foster_cuda_is_registered_next_to_foster


def test_kirchner_cuda_against_the_references_own_numba_kernel(pyleo, cuda_standins):
    """The parity target itself, run here: utils/wavelet.py:872-1049 (kirchner_numba, JIT-compiled by numba on the
    host) against method='Kirchner_cuda' through the reference's own wwz() front end, default grids."""
    W = pyleo.utils.wavelet
    t, y = synth.uneven_ar1(400, 21)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = W.wwz(y, t, method="Kirchner_cuda", nproc=1)
        ref = W.wwz(y, t, method="Kirchner_numba", nproc=1)
    assert_same_nan_mask(got.amplitude, ref.amplitude)
    assert_same_nan_mask(got.Neffs, ref.Neffs)
    assert max_rel(got.amplitude, ref.amplitude) < 1e-9 and max_rel(got.Neffs, ref.Neffs) < 1e-12
    d = np.abs(got.phase - ref.phase)
    m = ~np.isnan(d)
    assert np.max(np.minimum(d[m], 2 * np.pi - d[m])) < 1e-9


def test_class_level_timings(pyleo, cuda_standins, capsys):
    """BASELINE.json configs 3 and 4 as a Pyleoclim user writes them, on the GPU with the real classes:
    Series.wavelet('wwz', settings={'method': 'Kirchner_cuda'}).signif_test(number=10000) on n = 5,000, and
    EnsembleSeries.wavelet on age-ensemble members (record / replay through MultipleSeries.wavelet)."""
    t3, y3, tau3, freq3 = synth.config_grid("C3")
    s = pyleo.Series(time=t3, value=y3, label="c3", time_unit="yr", verbose=False)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scal = s.wavelet(method="wwz", settings={"method": "Kirchner_cuda"})
        for _ in range(2):
            t0 = time.perf_counter()
            sig = scal.signif_test(number=10000, qs=[0.95], seed=1)
            out["Series.wavelet().signif_test(number=10000), n=5000 [s]"] = time.perf_counter() - t0
    assert sig.signif_qs.scalogram_list[0].amplitude.shape == scal.amplitude.shape
    assert np.isfinite(sig.signif_qs.scalogram_list[0].amplitude).mean() > 0.9
    M = 200
    axes, y4 = synth.age_ensemble(1500, M)
    members = [pyleo.Series(time=tm, value=y4, label="m%d" % i, time_unit="yr", verbose=False) for i, tm in enumerate(axes)]
    ens = pyleo.EnsembleSeries(members)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(2):
            t0 = time.perf_counter()
            ms = ens.wavelet(method="wwz", settings={"method": "Kirchner_cuda"})
            out["EnsembleSeries.wavelet, %d members of n=1500 [s]" % M] = time.perf_counter() - t0
    assert len(ms.scalogram_list) == M and cuda_standins["ragged"] >= 1
    one = WV.wwz(y4, axes[7])
    assert max_rel(ms.scalogram_list[7].amplitude, one.amplitude) < 1e-11
    with capsys.disabled():
        for k, v in out.items():
            print("\n[class-level timing] %s: %.4f" % (k, v))
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out", "class_level_timings.txt"), "w") as f:
        for k, v in out.items():
            f.write("%s: %.4f\n" % (k, v))


def test_coherence_and_its_signif_test_on_the_device(pyleo, cuda_standins, monkeypatch):
    """SURVEY.md 8f rank 2 through the real classes: Series.wavelet_coherence(method='wwz') with the CUDA kernel against
    the reference's own Kirchner path, and Coherence.signif_test (core/coherences.py:748-963) taking the batched device
    pipeline (two shared-axis batches + wwzb200_wtc_dev + device quantiles) instead of one pool task per pair."""
    s1, s2 = T._series(pyleo, 220, 3, "a"), T._series(pyleo, 200, 4, "b")
    calls = []
    real = B.coherence_signif
    monkeypatch.setattr(B, "coherence_signif", lambda *a, **k: (calls.append(k.get("number")), real(*a, **k))[1])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        coh = s1.wavelet_coherence(s2, method="wwz", settings={"method": "Kirchner_cuda"})
        ref = s1.wavelet_coherence(s2, method="wwz", settings={"method": "Kirchner", "nproc": 1})
        sig = coh.signif_test(number=12, qs=[0.9, 0.95], seed=5)
    assert_same_nan_mask(coh.wtc, ref.wtc)
    assert max_rel(coh.wtc, ref.wtc) < 1e-8 and max_rel(coh.xwt, ref.xwt) < 1e-8
    assert calls == [12] and sig is not coh and sig.signif_method == "ar1sim" and sig.qs == [0.9, 0.95]
    assert len(sig.signif_qs) == 2 and all(type(m).__name__ == "MultipleScalogram" for m in sig.signif_qs)
    for ms in sig.signif_qs:
        assert [q.label for q in ms.scalogram_list] == ["90%", "95%"]
        for q in ms.scalogram_list:
            assert q.amplitude.shape == coh.wtc.shape and np.array_equal(q.time, coh.time)
            assert np.array_equal(q.frequency, coh.frequency)
    lo, hi = (q.amplitude for q in sig.signif_qs[0].scalogram_list)
    m = ~np.isnan(lo)
    assert m.any() and (hi[m] >= lo[m]).all() and (hi[m] <= 1 + 1e-12).all()
    # anything the batched path does not cover goes to the reference's own implementation
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert ref.signif_test(number=0) is ref
