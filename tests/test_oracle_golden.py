"""The oracle (oracle/wwz_oracle.c + .py) against the reference's own golden scalogram and
against outputs of the reference's kirchner_numba / wwz_psd / ar1_model run in the build container
(fixtures: tests/golden/*.npz, generator: tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden, assert_scalogram_parity, as_dict, max_rel, assert_same_nan_mask
from oracle import wwz_oracle as O
from pyleoclim_util_b200 import synth


def _res(r):
    return as_dict(r.amplitude, r.phase, r.Neffs, r.coeff)


def test_benthic_known_answer_file():
    """example_data/scal_signif_benthic.json: the one artefact the reference ships."""
    g = golden("benthic")
    r = O.wwz(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"])
    assert_same_nan_mask(r.amplitude, g["file_amplitude"])
    assert max_rel(r.amplitude, g["file_amplitude"]) < 1e-11
    assert max_rel(r.Neffs, g["file_Neffs"]) < 1e-12
    assert np.max(np.abs(r.coi - g["file_coi"])) == 0.0
    assert_scalogram_parity(_res(r), g, amp_rtol=1e-10, phase_atol=1e-10, what="benthic vs numba")


def test_soi_defaults_and_psd():
    g = golden("soi")
    r = O.wwz(g["ys"], g["ts"])
    assert np.array_equal(r.freq, g["freq"]) and np.array_equal(r.time, g["tau"])
    assert np.allclose(r.coi, g["coi"], rtol=1e-15)
    assert_scalogram_parity(_res(r), g, amp_rtol=1e-10, phase_atol=1e-10, what="soi")
    p = O.wwz_psd(g["ys"], g["ts"])
    assert np.array_equal(p.freq, g["psd_freq"])
    assert max_rel(p.psd, g["psd"]) < 1e-11


@pytest.mark.parametrize("tag,kw", [("", {}), ("c1e3_", {"c": 1e-3}), ("thr5_", {"Neff_threshold": 5}),
                                     ("raw_", {"standardize": False})])
def test_synth_small_variants(tag, kw):
    g = golden("synth_small")
    r = O.wwz(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"], **kw)
    ref = {k: g[tag + k] for k in ("amplitude", "phase", "Neffs", "a0", "a1", "a2")}
    assert_scalogram_parity(_res(r), ref, amp_rtol=1e-11, phase_atol=1e-11, what="synth_small " + tag)


def test_synth_small_psd():
    g = golden("synth_small")
    p = O.wwz_psd(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"])
    assert max_rel(p.psd, g["psd"]) < 1e-12
    p = O.wwz_psd(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"], Neff_threshold=6)
    assert max_rel(p.psd, g["psd_thr6"]) < 1e-12
    p = O.wwz_psd(g["ys"], g["ts"], freq=g["freq_in"])
    assert max_rel(p.psd, g["psd_default_tau"]) < 1e-12


def test_edge_three_state_mask():
    """tau inside a long gap, f above Nyquist, Neff <= threshold: all three mask states, including the
    subnormal-weight cell where fastmath numba masks a NaN Neff and kirchner_basic does not."""
    g = golden("edge")
    wwa, ph, ne, co = O.kirchner(g["ys"], g["ts"], g["freq_in"], g["tau_in"], standardize=True)
    assert_scalogram_parity(as_dict(wwa, ph, ne, co), g, amp_rtol=1e-12, what="edge (numba semantics)")
    assert np.isnan(ne).sum() == 13 and np.isnan(wwa).sum() == 19
    wwa_b, _, ne_b, _ = O.kirchner(g["ys"], g["ts"], g["freq_in"], g["tau_in"], standardize=True, mode=2)
    assert_same_nan_mask(wwa_b, g["basic_amplitude"])
    assert max_rel(wwa_b, g["basic_amplitude"]) < 1e-12
    assert np.isnan(wwa_b).sum() == 18      # kirchner_basic keeps the subnormal cell


def test_config2_full_grid():
    """Config 2 (n=2000, 1000 tau x 500 f): full grid on the CPU, stored rows + mask + psd."""
    g = golden("c2_sub")
    t, y, tau, freq = synth.config_grid("C2")
    r = O.wwz(y, t, tau=tau, freq=freq)
    rows = g["rows"]
    got = {k: v[rows] for k, v in _res(r).items()}
    assert_scalogram_parity(got, g, amp_rtol=1e-10, phase_atol=1e-10, what="C2 rows")
    assert np.array_equal(np.packbits(np.isnan(r.amplitude)), g["nan_mask_packed"])
    assert int(np.isnan(r.amplitude).sum()) == int(g["full_nan_count"][0])
    assert abs(np.nansum(r.amplitude) / g["full_amp_nansum"][0] - 1) < 1e-12
    p = O.wwz_psd(y, t, tau=tau, freq=freq)
    assert max_rel(p.psd, g["psd"]) < 1e-12


def test_config5_sample():
    g = golden("c5_sample")
    t, y, tau, freq = synth.config_grid("C5")
    ys = (y - y.mean()) / y.std()
    wwa, ph, ne, co = O.kirchner(ys, t, freq[g["kf"]], tau[g["jt"]])
    assert_scalogram_parity(as_dict(wwa, ph, ne, co), g, amp_rtol=1e-10, phase_atol=1e-10, what="C5 sample")


def test_onepass_algebra_matches_faithful():
    """The reduced formulation the CUDA kernels use (basis rotation cancelled) against the
    reference-faithful two-pass evaluation, on the CPU."""
    g = golden("synth_small")
    a = O.kirchner(g["ys"], g["ts"], g["freq_in"], g["tau_in"], standardize=True, mode=0)
    b = O.kirchner(g["ys"], g["ts"], g["freq_in"], g["tau_in"], standardize=True, mode=1)
    assert_scalogram_parity(as_dict(*b), as_dict(*a), amp_rtol=1e-11, phase_atol=1e-11, what="one-pass")


def test_surrogate_generator():
    g = golden("surrogates")
    t, y = g["ts"], g["ys"]
    assert abs(O.tau_estimation(y, t) / g["tau_est"][0] - 1) < 1e-6
    np.random.seed(1234)
    m = O.ar1_model(t, 4.0, output_sigma=1.7)
    assert np.max(np.abs(m - g["model_seed1234_tau4_sig1p7"])) < 1e-13
    np.random.seed(99)
    sim = O.ar1_sim(y, 3, t)
    assert np.max(np.abs(sim - g["sim_seed99_p3"])) < 1e-9


def test_foster_restatement_vs_reference_outputs():
    """method='Foster' (wwz_basic, utils/wavelet.py:228-381)."""
    g = golden("foster")
    wwa, ph, ne, co = O.foster(g["ys"], g["ts"], g["freq_in"], g["tau_in"], standardize=True)
    assert np.array_equal(np.isnan(wwa), np.isnan(g["k_wwa"]))
    m = ~np.isnan(g["k_wwa"])
    assert np.max(np.abs(wwa[m] / g["k_wwa"][m] - 1)) < 1e-12 and np.max(np.abs(ne / g["k_Neffs"] - 1)) < 1e-13
    for c, name in zip(co, ("k_y1", "k_y2", "k_y3")):
        assert np.nanmax(np.abs(c - g[name])) < 1e-13
    assert np.array_equal(g["amplitude"], g["k_wwa"], equal_nan=True)       # wwz(method='Foster') is this kernel
    wwa2 = O.foster(g["ys2"], g["ts2"], g["freq2"], g["tau2"], standardize=True)[0]
    assert np.array_equal(np.isnan(wwa2), np.isnan(g["s_wwa"]))
    m2 = ~np.isnan(g["s_wwa"])
    assert np.max(np.abs(wwa2[m2] / g["s_wwa"][m2] - 1)) < 1e-11
    psd = O.wwa2psd(g["amplitude"], g["ts"], g["Neffs"], Neff_threshold=3)   # c = 1e-3 in wwz_psd: its own transform
    assert psd.shape == g["psd"].shape


def test_uar1_fit_and_sim():
    """uar1 surrogates: likelihood, Nelder-Mead fit and the recurrence against outputs of the
    reference's own uar1_fit / uar1_sim (utils/tsmodel.py:612-716), draw for draw."""
    g = golden("uar1")
    t, y = g["ts"], g["ys"]
    assert abs(O.n_ll_unevenly_spaced_ar1([np.log(4.0), np.log(1.3)], y, t) / g["nll_tau4_s1p3"][0] - 1) < 1e-13
    theta = O.uar1_fit(y, t)
    assert np.max(np.abs(theta / g["theta_hat"] - 1)) < 1e-6
    sim1 = O.uar1_sim(t, tau=g["theta_hat"][0], sigma_2=g["theta_hat"][1], z=g["z_seed4321"])
    assert sim1.shape == g["sim_seed4321"].shape and np.max(np.abs(sim1 - g["sim_seed4321"])) < 1e-13
    T = np.tile(t, (3, 1)).T
    sim3 = O.uar1_sim(T, tau=3.5, sigma_2=2.2, z=g["z_seed77"])
    assert np.max(np.abs(sim3 - g["sim_seed77_tau3p5_s2p2"])) < 1e-13
    np.random.seed(77)
    assert np.array_equal(O.uar1_sim(T, tau=3.5, sigma_2=2.2), sim3)      # same draws from the global generator


def test_prepare_wwz_cleaning():
    g = golden("cleaning")
    ys_cut, ts_cut, freq, tau = O.prepare_wwz(g["ys"], g["ts"], tau=g["tau_in"])
    assert np.array_equal(ts_cut, g["ts_cut"]) and np.allclose(ys_cut, g["ys_cut"], rtol=1e-15, atol=0)
    assert np.array_equal(tau, g["tau_out"]) and np.array_equal(freq, g["freq_out"])
    _, ts2, freq2, tau2 = O.prepare_wwz(g["ys"], g["ts"])
    assert np.array_equal(ts2, g["ts_cut_default"]) and np.array_equal(freq2, g["freq_default"])
    assert np.array_equal(tau2, g["tau_default"])
    r = O.wwz(g["ys"], g["ts"], tau=g["tau_in"])
    assert_scalogram_parity(dict(amplitude=r.amplitude, phase=r.phase, Neffs=r.Neffs), g, amp_rtol=1e-11,
                            phase_atol=1e-11, what="cleaning")


def test_mquantiles_matches_scipy():
    from scipy.stats.mstats import mquantiles
    rng = np.random.default_rng(3)
    for n in (2, 3, 7, 200):
        a = rng.standard_normal((n, 11))
        qs = [0.05, 0.5, 0.9, 0.95, 0.99]
        assert np.allclose(O.mquantiles(a, qs, axis=0), np.asarray(mquantiles(a, qs, axis=0)), rtol=0, atol=1e-15)


def test_oracle_against_live_reference_if_present():
    """When the reference tree is mounted (build container only) compare live, not only fixtures."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not mounted (expected on the GPU box)")
    W = ref_shim.load()["wavelet"]
    t, y = synth.uneven_ar1(150, 123)
    tau = np.linspace(t.min(), t.max(), 9)
    freq = synth.freq_log(t, 12)
    ref = W.kirchner_basic(y, t, freq, tau, standardize=True)
    got = O.kirchner(y, t, freq, tau, standardize=True, mode=2)
    assert_scalogram_parity(as_dict(*got), as_dict(*ref), amp_rtol=1e-12, phase_atol=1e-12, what="live kirchner_basic")


@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_coherence_restatement_vs_reference_outputs(case):
    """the oracle's wtc / xwt / wwz_coherence (utils/wavelet.py:1497-1687, :2201-2364) against
    outputs of the reference's own wwz_coherence -- default 50 taus with a grid the reference re-derives (A), an even
    scale-smoothing window (B), Neff-masked cells whose NaN spreads through the smoothing (C)."""
    g = golden("coherence")
    G = lambda k: g[case + "_" + k]
    c1, c2 = G("a1_1") - 1j * G("a2_1"), G("a1_2") - 1j * G("a2_2")
    coh, ph = O.wtc(c1, c2, 1 / G("freq"), G("tau"))
    assert np.array_equal(coh, G("xw_coherence"), equal_nan=True) and np.array_equal(ph, G("xw_phase"), equal_nan=True)
    prod, amp, _ = O.xwt(c1, c2)
    assert np.array_equal(amp, G("xw_amplitude"), equal_nan=True)
    # the reference transforms the series as truncated by the FIRST prepare_wwz (on the tau it was given, :1636-1637)
    y1c, t1c, _, _ = O.prepare_wwz(G("y1"), G("t1"), freq=G("freq"), tau=G("tau_in"))
    y2c, t2c, _, _ = O.prepare_wwz(G("y2"), G("t2"), freq=G("freq"), tau=G("tau_in"))
    r = O.wwz_coherence(y1c, t1c, y2c, t2c, tau=G("tau"), freq=G("freq"))
    assert np.array_equal(np.isnan(r.xw_coherence), np.isnan(G("xw_coherence")))
    assert max_rel(r.xw_coherence, G("xw_coherence")) < 1e-9 and max_rel(r.xw_amplitude, G("xw_amplitude")) < 1e-9
    # the product's host-side grid logic reproduces the grid the reference settled on
    from pyleoclim_util_b200 import wavelet as WV
    _, _, _, _, tau, freq = WV.coherence_grid(G("y1"), G("t1"), G("y2"), G("t2"), tau=G("tau_in"), freq=G("freq"))
    assert np.array_equal(tau, G("tau")) and np.array_equal(freq, G("freq"))
    assert WV.wtc_boxcar_length(freq) == O.boxcar_length(1 / freq)

