"""GPU tests of the wavelet-coherence path: wtc / xwt on the device against outputs of the
reference's own wwz_coherence (tests/golden/coherence.npz, utils/wavelet.py:1497-1687, :2201-2364), and the batched
surrogate-pair significance test against the unfused pipeline built from the oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import golden, max_rel, max_phase_abs, assert_same_nan_mask, has_cuda

pytestmark = pytest.mark.gpu

from pyleoclim_util_b200 import wavelet as WV, batch as B   # noqa: E402


@pytest.fixture(autouse=True)
def _need_gpu():
    if not has_cuda():
        pytest.skip("needs a CUDA device")


def oracle():
    from oracle import wwz_oracle as O
    return O


@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_wtc_on_the_references_coefficients(case):
    """The coherency smoothing alone: the reference's coefficients in, its xw_coherence / xw_phase / xw_amplitude out
    (case A: 50 taus -> 64-point circular convolution, boxcar of 7; B: even boxcar; C: NaN cells of a hiatus)."""
    g = golden("coherence")
    G = lambda k: g[case + "_" + k]
    coh, ph, amp = WV.wtc_cuda(G("a1_1"), G("a2_1"), G("a1_2"), G("a2_2"), G("freq"), G("tau"))
    assert_same_nan_mask(coh, G("xw_coherence"), "coherence " + case)
    assert_same_nan_mask(amp, G("xw_amplitude"), "xw amplitude " + case)
    assert max_rel(coh, G("xw_coherence")) < 1e-9
    assert max_rel(amp, G("xw_amplitude")) < 1e-12
    assert max_phase_abs(ph, G("xw_phase")) < 1e-9
    # a batch of pairs gives the same rows
    stack = lambda k: np.stack([G(k), 0.5 * G(k), G(k)])
    coh3, _, amp3 = WV.wtc_cuda(stack("a1_1"), stack("a2_1"), stack("a1_2"), stack("a2_2"), G("freq"), G("tau"), want_phase=False)
    assert np.array_equal(coh3[0], coh, equal_nan=True) and np.array_equal(coh3[2], coh, equal_nan=True)
    assert max_rel(coh3[1], coh) < 1e-12            # coherence is scale invariant
    assert max_rel(amp3[1], 0.25 * amp) < 1e-14      # both coefficient sets were halved


@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_wwz_coherence_against_the_reference(case):
    """wavelet.wwz_coherence end to end (grid reconciliation, two transforms, wtc, xwt) against the reference's Results."""
    g = golden("coherence")
    G = lambda k: g[case + "_" + k]
    with pytest.warns(UserWarning):
        r = WV.wwz_coherence(G("y1"), G("t1"), G("y2"), G("t2"), tau=G("tau_in"), freq=G("freq"))
    assert np.array_equal(r.time, G("tau")) and np.array_equal(r.freq, G("freq")) and np.array_equal(r.coi, G("coi"))
    assert_same_nan_mask(r.xw_coherence, G("xw_coherence"))
    assert max_rel(r.xw_coherence, G("xw_coherence")) < 1e-9
    assert max_rel(r.xw_amplitude, G("xw_amplitude")) < 1e-9
    assert max_phase_abs(r.xw_phase, G("xw_phase")) < 1e-9
    m = ~np.isnan(G("xwt"))
    assert np.max(np.abs(r.xwt[m] - G("xwt")[m])) <= 1e-9 * np.nanmax(np.abs(G("xwt")))


@pytest.mark.parametrize("method", ["ar1sim", "uar1"])
def test_coherence_signif_equals_the_unfused_pipeline(method):
    """The batched surrogate-pair test (Coherence.signif_test, core/coherences.py:748-963) against the same pipeline
    assembled pair by pair from the oracle: same surrogates (device generator, seeds s and s + 1), oracle wwz_coherence
    of every pair, scipy mquantiles."""
    from scipy.stats.mstats import mquantiles
    O = oracle()
    g = golden("coherence")
    y1, t1, y2, t2, tau, freq = (g["B_" + k] for k in ("y1", "t1", "y2", "t2", "tau_in", "freq"))
    S, qs, seed = 24, [0.5, 0.9], 7
    sg = B.coherence_signif(y1, t1, y2, t2, tau=tau, freq=freq, number=S, qs=qs, seed=seed, method=method)
    (y1c, t1c, _), (y2c, t2c, _), tau_c, freq_c = WV.coherence_inputs(y1, t1, y2, t2, tau=tau, freq=freq)
    assert np.array_equal(sg.time, tau_c) and np.array_equal(sg.freq, freq_c)
    sets = []
    for k, (yc, tc) in enumerate(((y1c, t1c), (y2c, t2c))):
        if method == "ar1sim":
            sets.append(B.ar1_surrogates(tc, B.tau_estimation(yc, tc), sigma=np.std(yc), mu=np.mean(yc), number=S, seed=seed + k))
        else:
            tau_u, s2 = B.uar1_fit(yc, tc)
            sets.append(B.uar1_surrogates(tc, tau_u, s2, mu=np.mean(yc), number=S, seed=seed + k))
    coh, amp = [], []
    for s in range(S):
        r = O.wwz_coherence(sets[0][s], t1c, sets[1][s], t2c, tau=tau_c, freq=freq_c)
        coh.append(r.xw_coherence)
        amp.append(r.xw_amplitude)
    for got, ref in ((sg.wtc_qs, np.stack(coh)), (sg.xwt_qs, np.stack(amp))):
        q = np.asarray(mquantiles(ref.reshape(S, -1), qs, axis=0).filled(np.nan)).reshape(len(qs), tau_c.size, freq_c.size)
        assert_same_nan_mask(got, q)
        assert max_rel(got, q) < 1e-8
    m = ~np.isnan(sg.wtc_qs[0])
    assert (sg.wtc_qs[1][m] >= sg.wtc_qs[0][m]).all() and (sg.wtc_qs[:, m] <= 1 + 1e-12).all()
