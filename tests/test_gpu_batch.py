"""Batched paths on the GPU: AR(1) surrogates, shared-time-axis transform, ragged ensembles,
per-cell quantiles, fused significance test -- against the oracle and against the single-series path."""
import os

import numpy as np
import pytest

from conftest import assert_scalogram_parity, as_dict, max_rel, assert_same_nan_mask, has_cuda

pytestmark = pytest.mark.gpu

from pyleoclim_util_b200 import wavelet as WV, batch as B, _lib, synth   # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert has_cuda(), "these tests need a CUDA device and the built libwwz_b200.so (no CPU fallback)"


def oracle():
    from oracle import wwz_oracle as O
    return O


# ------------------------------------------------------------------ surrogates
def test_ar1_surrogates_match_the_oracle_recurrence():
    """Same recurrence as ar1_model (utils/tsmodel.py:59-69) driven by the documented Philox stream."""
    O = oracle()
    t, _ = synth.uneven_ar1(333, 5)
    Y = B.ar1_surrogates(t, tau_est=3.7, sigma=1.9, mu=0.25, number=37, seed=12345)
    assert Y.shape == (37, 333)
    for s in (0, 1, 17, 36):
        z = O.philox_normals(12345, s, t.size)
        ref = O.ar1_model(t, 3.7, output_sigma=1.9, z=z) + 0.25
        assert np.max(np.abs(Y[s] - ref)) < 1e-11
    assert np.all(Y[:, 0] == 0.25)                       # y[0] = 0 (+ mu)
    Y2 = B.ar1_surrogates(t, tau_est=3.7, sigma=1.9, mu=0.25, number=37, seed=12345)
    assert np.array_equal(Y, Y2)                         # counter-based: reproducible
    Y3 = B.ar1_surrogates(t, tau_est=3.7, sigma=1.9, mu=0.25, number=37, seed=12346)
    assert not np.allclose(Y, Y3)


def test_ar1_surrogate_statistics():
    """The reference's own acceptance criterion (tests/test_core_SurrogateSeries.py:90-104): the AR(1)
    parameters are recovered from the surrogates; here with a tighter tolerance and many series."""
    t, y = synth.uneven_ar1(2000, 3)
    tau_true, sigma = 5.0, 2.0
    Y = B.ar1_surrogates(t, tau_est=tau_true, sigma=sigma, number=256, seed=7)
    assert abs(Y[:, 500:].std() / sigma - 1) < 0.02
    assert abs(Y.mean()) < 0.05
    taus = [B.tau_estimation(Y[s], t) for s in range(0, 256, 8)]
    assert abs(np.mean(taus) / tau_true - 1) < 0.1
    sim = B.ar1_sim(y, 5, t, seed=1)
    assert sim.shape == (2000, 5) and abs(sim.std() / y.std() - 1) < 0.1
    assert B.ar1_sim(y, 1, t).shape == (2000,)


def test_uar1_surrogates_match_the_oracle_recurrence():
    """uar1_sim (utils/tsmodel.py:673-716): y[0] = z[0], innovations sqrt(sigma_2 (1 - phi^2)) z, same Philox stream."""
    O = oracle()
    t, y = synth.uneven_ar1(301, 9)
    Y = B.uar1_surrogates(t, tau=2.9, sigma_2=1.7, mu=-0.5, number=40, seed=99)
    assert Y.shape == (40, 301)
    for s in (0, 3, 39):
        z = O.philox_normals(99, s, t.size)
        ref = O.uar1_sim(t, tau=2.9, sigma_2=1.7, z=z) - 0.5
        assert np.max(np.abs(Y[s] - ref)) < 1e-11
        assert abs(Y[s, 0] - (z[0] - 0.5)) < 1e-14              # y[0] = z[0] (+ mu)
    sim = B.uar1_sim(np.tile(t, (6, 1)).T, tau=2.9, sigma_2=1.7, seed=99)
    assert sim.shape == (301, 6) and np.max(np.abs(sim - (Y[:6].T + 0.5))) < 1e-14
    assert B.uar1_sim(t, tau=2.9, sigma_2=1.7, seed=99).shape == (301,)
    # the fit recovers the parameters of its own surrogates (reference criterion: tests/test_core_SurrogateSeries.py:90-104)
    t2, _ = synth.uneven_ar1(2000, 4)
    Y2 = B.uar1_surrogates(t2, tau=5.0, sigma_2=2.0, number=8, seed=1)
    th = np.array([B.uar1_fit(Y2[s], t2) for s in range(8)])
    assert abs(th[:, 0].mean() / 5.0 - 1) < 0.15 and abs(th[:, 1].mean() / 2.0 - 1) < 0.1
    gfit = np.load(os.path.join(os.path.dirname(__file__), "golden", "uar1.npz"))
    assert np.max(np.abs(B.uar1_fit(gfit["ys"], gfit["ts"]) / gfit["theta_hat"] - 1)) < 1e-6


def test_wwz_signif_with_uar1_surrogates():
    from scipy.stats.mstats import mquantiles
    t, y = synth.uneven_ar1(350, 13)
    number, qs, seed = 40, [0.9, 0.95], 5
    sg = B.wwz_signif(y, t, number=number, qs=qs, seed=seed, method="uar1")
    ys_cut, ts_cut, freq, tau = WV.prepare_wwz(y, t)
    tau_hat, s2 = B.uar1_fit(ys_cut, ts_cut)
    Y = B.uar1_surrogates(ts_cut, tau_hat, s2, mu=np.mean(ys_cut), number=number, seed=seed)
    amps = np.stack([WV.kirchner_cuda(Y[s], ts_cut, freq, tau, standardize=True)[0] for s in range(number)])
    ref = np.asarray(mquantiles(amps.reshape(number, -1), qs, axis=0).filled(np.nan)).reshape(2, tau.size, freq.size)
    assert_same_nan_mask(sg.amplitude_qs, ref)
    assert max_rel(sg.amplitude_qs, ref) < 1e-10
    with pytest.raises(ValueError):
        B.wwz_signif(y, t, number=4, method="phaseran")


# ------------------------------------------------------------------ shared time axis
def _shared_case(n=500, nt=21, nf=17, S=37, seed=2):
    t, y = synth.uneven_ar1(n, seed)
    tau = np.linspace(t.min(), t.max(), nt)
    freq = synth.freq_log(t, nf)
    Y = B.ar1_surrogates(t, 4.0, sigma=1.3, mu=0.7, number=S, seed=seed)
    return t, Y, tau, freq


@pytest.mark.parametrize("standardize", [True, False])
def test_shared_axis_matches_single_series_path(standardize):
    t, Y, tau, freq = _shared_case()
    r = B.wwz_shared_axis(Y, t, tau, freq, standardize=standardize, outputs=("amplitude", "phase", "coeff"),
                          psd_Neff_threshold=3, c=1e-3)
    for s in range(Y.shape[0]):
        one = WV.kirchner_cuda(Y[s], t, freq, tau, c=1e-3, standardize=standardize)
        got = dict(amplitude=r.amplitude[s], phase=r.phase[s], Neffs=r.Neffs, a0=r.coeff[0][s], a1=r.coeff[1][s],
                   a2=r.coeff[2][s])
        assert_scalogram_parity(got, as_dict(*one), amp_rtol=1e-11, phase_atol=1e-11, what="series %d" % s)
        p = WV.wwz_psd(Y[s], t, freq=freq, tau=tau, c=1e-3, standardize=standardize)
        assert max_rel(r.psd[s], p.psd) < 1e-11


def test_shared_axis_vs_oracle_default_grid_shape():
    """nt = 50 (the reference's default) with a frequency above Nyquist and Neff-masked cells."""
    O = oracle()
    t, Y, tau, freq = _shared_case(n=400, nt=50, nf=12, S=5)
    freq = np.append(freq, 3.0)                               # > Nyquist: NaN column
    r = B.wwz_shared_axis(Y, t, tau, freq, outputs=("amplitude", "phase", "coeff"))
    for s in range(Y.shape[0]):
        o = O.kirchner(Y[s], t, freq, tau, standardize=True)
        got = dict(amplitude=r.amplitude[s], phase=r.phase[s], Neffs=r.Neffs, a0=r.coeff[0][s], a1=r.coeff[1][s],
                   a2=r.coeff[2][s])
        assert_scalogram_parity(got, as_dict(*o), what="oracle series %d" % s)
    assert np.isnan(r.amplitude[:, :, -1]).all()


def test_shared_axis_gap_series_deep_underflow():
    O = oracle()
    rng = np.random.default_rng(8)
    t = np.concatenate([np.sort(rng.uniform(0, 50, 100)), np.sort(rng.uniform(400, 450, 100))])
    Y = rng.standard_normal((3, t.size))
    tau = np.linspace(t.min(), t.max(), 30)
    freq = np.logspace(-2.5, 0.05, 20)
    r = B.wwz_shared_axis(Y, t, tau, freq, outputs=("amplitude", "phase"))
    for s in range(3):
        o = O.kirchner(Y[s], t, freq, tau, standardize=True)
        assert_scalogram_parity(dict(amplitude=r.amplitude[s], phase=r.phase[s], Neffs=r.Neffs),
                                dict(amplitude=o[0], phase=o[1], Neffs=o[2]), what="gap series %d" % s)


@pytest.mark.parametrize("S", [1, 15, 16, 17, 100])
def test_shared_axis_series_count_padding(S):
    t, Y, tau, freq = _shared_case(n=200, nt=9, nf=7, S=S)
    r = B.wwz_shared_axis(Y, t, tau, freq)
    assert r.amplitude.shape == (S, 9, 7)
    one = WV.kirchner_cuda(Y[S - 1], t, freq, tau, standardize=True)
    assert max_rel(r.amplitude[S - 1], one[0]) < 1e-11


# ------------------------------------------------------------------ ragged ensembles
def test_ragged_members_match_single_calls():
    axes, y = synth.age_ensemble(300, 7)
    members = []
    for m, t in enumerate(axes):
        ys_cut, ts_cut, freq, tau = WV.prepare_wwz(y, t, tau=np.linspace(t.min(), t.max(), 10 + m))
        members.append((ys_cut, ts_cut, tau, freq[: 20 + m]))
    res, psds = B.wwz_ragged(members, psd_Neff_threshold=3)
    assert len(res) == 7
    for (ys, ts, tau, freq), r, p in zip(members, res, psds):
        one = WV.kirchner_cuda(ys, ts, freq, tau, standardize=True)
        assert r.amplitude.shape == (tau.size, freq.size)
        assert np.array_equal(r.amplitude.view(np.uint64), one[0].view(np.uint64))
        assert np.array_equal(r.Neffs.view(np.uint64), one[2].view(np.uint64))
        ps = WV.wwa2psd(one[0], ts, one[2], Neff_threshold=3)
        assert max_rel(p, ps) < 1e-13


def test_ragged_equal_length_members_take_the_vectorised_host_path():
    """Age-ensemble shape (config 4): same values and length, own time axis per member.  The host side z-scores all
    members, takes their Nyquist frequencies and cones of influence in one pass each; results equal single calls."""
    axes, y = synth.age_ensemble(300, 12)
    members = []
    for tm in axes:
        ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y, tm)
        members.append((ys_c, ts_c, ta_c, fr_c))
    res, psds = B.wwz_ragged(members, psd_Neff_threshold=3)
    assert len(res) == 12
    for (ys_c, ts_c, ta_c, fr_c), r, p in zip(members, res, psds):
        one = WV.kirchner_cuda(ys_c, ts_c, fr_c, ta_c, standardize=True)
        assert_same_nan_mask(r.amplitude, one[0])
        assert max_rel(r.amplitude, one[0]) < 1e-12 and max_rel(r.Neffs, one[2]) < 1e-13
        assert np.array_equal(r.coi, WV.make_coi(ta_c)) and np.array_equal(r.time, ta_c)
        assert max_rel(p, WV.wwa2psd(one[0], ts_c, one[2], Neff_threshold=3)) < 1e-12
    flat = [(np.full(50, 2.0), np.arange(50.0), np.linspace(0, 49, 5), np.array([0.05, 0.1]))] * 2
    with pytest.warns(UserWarning):
        B.wwz_ragged(flat)                                   # constant series: left unscaled, with the warning


def test_config4_size_ensemble_against_the_oracle():
    """BASELINE.json configs[3] at its size: 1,000 age-ensemble members of n = 1,500 on their default 50 x 151 grids
    through the member-batch entry point (core/ensembleseries.py:169-173 -> core/multipleseries.py:1520-1532).
    32 sampled members against the ORACLE (two-pass Kirchner, host z-score, host make_omega) at the parity bar, and the
    three-state Neff mask and Neff values of EVERY member against single calls of the single-series path."""
    O = oracle()
    axes, y = synth.age_ensemble(1500, 1000)
    members = []
    for tm in axes:
        ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y, tm)
        members.append((ys_c, ts_c, ta_c, fr_c))
    res, psds = B.wwz_ragged(members, psd_Neff_threshold=3)
    assert len(res) == 1000 and res[0].amplitude.shape == (50, members[0][3].size)
    worst = 0.0
    for m in range(0, 1000, 32)[:32]:
        ys_c, ts_c, ta_c, fr_c = members[m]
        o = O.kirchner(ys_c, ts_c, fr_c, ta_c, standardize=True)
        r = res[m]
        got = dict(amplitude=r.amplitude, phase=r.phase, Neffs=r.Neffs, a0=r.coeff[0], a1=r.coeff[1], a2=r.coeff[2])
        assert_scalogram_parity(got, as_dict(*o), what="member %d vs oracle" % m)
        worst = max(worst, max_rel(r.amplitude, o[0]))
        po = O.wwa2psd(o[0], ts_c, o[2], Neff_threshold=3)
        assert max_rel(psds[m], po) < 1e-9
        assert np.array_equal(r.coi, WV.make_coi(ta_c))
    print("config-4 ensemble: worst amplitude error of 32 sampled members vs the oracle %.2e" % worst)
    for m, (ys_c, ts_c, ta_c, fr_c) in enumerate(members):
        one = WV.kirchner_cuda(ys_c, ts_c, fr_c, ta_c, standardize=True)
        assert_same_nan_mask(res[m].Neffs, one[2], "member %d Neffs" % m)
        assert_same_nan_mask(res[m].amplitude, one[0], "member %d amplitude" % m)
        assert max_rel(res[m].Neffs, one[2]) < 1e-12
        assert max_rel(res[m].amplitude, one[0]) < 1e-11


def test_member_batch_shared_values_strides_and_flags():
    """wwz_members: one value vector shared by all members (ys_stride = 0) equals per-member copies bit for bit; the
    dense flag, explicit omega (no device make_omega) and a frequency above every member's Nyquist agree with single
    calls; groups (forced by a small basis budget) change nothing."""
    axes, y = synth.age_ensemble(400, 9)
    T = np.stack(axes)
    tau = np.stack([np.linspace(t.min(), t.max(), 23) for t in axes])
    freq = np.stack([np.append(synth.freq_log(t, 30), 3.0) for t in axes])          # last column above Nyquist
    a = B.wwz_members(T, y, tau, freq)
    b = B.wwz_members(T, np.tile(y, (9, 1)), tau, freq)
    for ra, rb in zip(a, b):
        assert np.array_equal(ra.amplitude.view(np.uint64), rb.amplitude.view(np.uint64))
    old = os.environ.get("WWZB200_BASIS_BYTES")
    os.environ["WWZB200_BASIS_BYTES"] = str(3 << 20)          # a few members per group
    try:
        g = B.wwz_members(T, y, tau, freq)
    finally:
        if old is None:
            os.environ.pop("WWZB200_BASIS_BYTES")
        else:
            os.environ["WWZB200_BASIS_BYTES"] = old
    d = B.wwz_members(T, y, tau, freq, flags=_lib.DENSE)
    for m in range(9):
        one = WV.kirchner_cuda(y, T[m], freq[m], tau[m], standardize=True)
        for r in (a[m], g[m], d[m]):
            assert_scalogram_parity(dict(amplitude=r.amplitude, phase=r.phase, Neffs=r.Neffs, a0=r.coeff[0],
                                         a1=r.coeff[1], a2=r.coeff[2]), as_dict(*one), amp_rtol=1e-11, phase_atol=1e-11)
        assert np.isnan(a[m].amplitude[:, -1]).all() and np.isnan(a[m].Neffs[:, -1]).all()


# ------------------------------------------------------------------ quantiles
@pytest.mark.parametrize("S", [1, 2, 3, 10, 200, 1000, 20000])
def test_quantiles_match_scipy_mquantiles(S):
    from scipy.stats.mstats import mquantiles
    rng = np.random.default_rng(S)
    ncell = 13 if S > 5000 else 57
    X = rng.standard_normal((S, ncell))
    X[:, 3] = np.nan                                    # a masked cell: NaN for every series
    qs = [0.05, 0.5, 0.9, 0.95, 0.99]
    got = B.quantiles(X, qs)
    ref = np.asarray(mquantiles(X, qs, axis=0).filled(np.nan))
    assert got.shape == ref.shape == (5, ncell)
    assert_same_nan_mask(got, ref)
    m = ~np.isnan(ref)
    assert np.max(np.abs(got[m] - ref[m])) < 1e-14
    got3 = B.quantiles(X.reshape(S, 1, ncell), [0.9])
    assert got3.shape == (1, 1, ncell)


@pytest.mark.parametrize("S,nq", [(64, 1), (257, 3), (5000, 2), (30000, 1), (300, 17)])
def test_quantiles_with_ties_and_scattered_nans(S, nq):
    """Radix-select path (few probabilities) and sort path (nq > 16) against scipy on data with repeated values,
    negative values, zeros of both signs, infinities and NaNs scattered through the sample (NaN sorts last)."""
    from scipy.stats.mstats import mquantiles
    rng = np.random.default_rng(S + nq)
    ncell = 9
    X = np.round(rng.standard_normal((S, ncell)), 1)             # many ties
    X[rng.random((S, ncell)) < 0.02] = np.nan
    X[rng.random((S, ncell)) < 0.01] = np.inf
    X[rng.random((S, ncell)) < 0.01] = -0.0
    X[:, 1] = 4.25                                               # constant cell
    X[: S // 2, 2] = np.nan                                      # half NaN: upper quantiles land on NaN
    qs = list(np.linspace(0.03, 0.97, nq))
    got = B.quantiles(X, qs)
    with np.errstate(invalid="ignore"):
        ref = np.asarray(mquantiles(X, qs, axis=0).filled(np.nan))
    assert_same_nan_mask(got, ref)
    m = np.isfinite(ref)
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    assert np.max(np.abs(got[m] - ref[m])) < 1e-14


# ------------------------------------------------------------------ fused significance test
def test_wwz_signif_equals_the_unfused_pipeline():
    from scipy.stats.mstats import mquantiles
    t, y = synth.uneven_ar1(400, 11)
    y = 2.0 * y + 3.0
    number, qs, seed = 48, [0.9, 0.95], 77
    sg = B.wwz_signif(y, t, number=number, qs=qs, seed=seed, psd=True, c=1e-3)
    ys_cut, ts_cut, freq, tau = WV.prepare_wwz(y, t)
    assert np.array_equal(sg.freq, freq) and np.array_equal(sg.time, tau)
    tau_est = B.tau_estimation(ys_cut, ts_cut)
    Y = B.ar1_surrogates(ts_cut, tau_est, sigma=np.std(ys_cut), mu=np.mean(ys_cut), number=number, seed=seed)
    amps = np.stack([WV.kirchner_cuda(Y[s], ts_cut, freq, tau, c=1e-3, standardize=True)[0] for s in range(number)])
    ref = np.asarray(mquantiles(amps.reshape(number, -1), qs, axis=0).filled(np.nan)).reshape(2, tau.size, freq.size)
    assert_same_nan_mask(sg.amplitude_qs, ref)
    assert max_rel(sg.amplitude_qs, ref) < 1e-10
    psds = np.stack([WV.wwz_psd(Y[s], ts_cut, freq=freq, tau=tau, c=1e-3).psd for s in range(number)])
    pref = np.asarray(mquantiles(psds, qs, axis=0))
    assert max_rel(sg.psd_qs, pref) < 1e-10


def test_surrogate_blocks_reassemble_the_fused_test():
    """The building blocks of the surrogate-sharded test (distributed._signif_by_surrogates) on one GPU: two blocks of
    surrogates generated with their global offsets equal the rows of the whole batch bit for bit; their cell-major
    amplitudes side by side, reduced by wwzb200_quantiles_cells_dev, give the levels of the fused single-GPU test."""
    import ctypes
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    L = _lib.lib()
    t, y = synth.uneven_ar1(420, 3)
    number, seed, qs = 150, 9, np.array([0.5, 0.95])
    sg = B.wwz_signif(y, t, number=number, qs=qs, seed=seed)
    ys_cut, ts_cut, freq, tau = WV.prepare_wwz(y, t)
    tau_est, sig, mu = B.tau_estimation(ys_cut, ts_cut), np.std(ys_cut), np.mean(ys_cut)
    whole = B.ar1_surrogates(ts_cut, tau_est, sigma=sig, mu=mu, number=number, seed=seed)
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    d_ts, d_tau, d_om, d_p = up(ts_cut), up(tau), up(WV.make_omega(ts_cut, freq)), up(qs)
    n, nt, nf = ts_cut.size, tau.size, freq.size
    ncell = nt * nf
    parts, ne = [], None
    for s0, s1 in ((0, 90), (90, 150)):
        S = s1 - s0
        d_Y = torch.empty((S, n), dtype=torch.float64, device=dev)
        _lib.check(L.wwzb200_surrogates_block_dev(d_ts.data_ptr(), n, float(tau_est), float(sig), float(mu), S, s0, seed, 0,
                                                  d_Y.data_ptr(), st))
        assert np.array_equal(d_Y.cpu().numpy(), whole[s0:s1])
        pad = ctypes.c_int64()
        _lib.check(L.wwzb200_series_pad(S, ctypes.byref(pad)))
        assert pad.value % 16 == 0 and pad.value >= S
        d_amp = torch.empty((ncell, pad.value), dtype=torch.float64, device=dev)
        d_ne = torch.empty((nt, nf), dtype=torch.float64, device=dev)
        _lib.check(L.wwzb200_shared_axis_cells_dev(d_ts.data_ptr(), d_Y.data_ptr(), n, S, d_tau.data_ptr(), nt,
                                                   d_om.data_ptr(), nf, 1 / (8 * np.pi ** 2), 3, _lib.STANDARDIZE,
                                                   d_ne.data_ptr(), d_amp.data_ptr(), st))
        parts.append(d_amp[:, :S])
        ne = d_ne
    X = torch.cat(parts, dim=1).contiguous()
    d_q = torch.empty((2, ncell), dtype=torch.float64, device=dev)
    _lib.check(L.wwzb200_quantiles_cells_dev(X.data_ptr(), number, number, ncell, d_p.data_ptr(), 2, d_q.data_ptr(), st))
    q = d_q.cpu().numpy().reshape(2, nt, nf)
    assert_same_nan_mask(q, sg.amplitude_qs)
    assert max_rel(q, sg.amplitude_qs) < 1e-12 and max_rel(ne.cpu().numpy(), sg.Neffs) < 1e-13


def test_config3_full_grid_surrogates_against_the_oracle():
    """Surrogates of config 3 (n = 5,000, default 50 x 501 grid) through the path the 10,000-surrogate test takes --
    64-series tiles, stored weight fragments, DMMA contraction -- checked DIRECTLY against the oracle on the full grid:
    320 surrogates are transformed in one batch (5 series blocks: the stored-weights mode), 8 of them, from every
    block and both ends of a block, are compared cell for cell (25,050 cells each)."""
    O = oracle()
    t, y, tau, freq = synth.config_grid("C3")
    ys_cut, ts_cut, freq_c, tau_c = WV.prepare_wwz(y, t, tau=tau, freq=freq)
    Y = B.ar1_surrogates(ts_cut, B.tau_estimation(ys_cut, ts_cut), sigma=np.std(ys_cut), mu=np.mean(ys_cut),
                         number=320, seed=5)
    r = B.wwz_shared_axis(Y, ts_cut, tau_c, freq_c, outputs=("amplitude", "phase"))
    worst = 0.0
    for s in (0, 63, 64, 130, 191, 200, 256, 319):
        o = O.kirchner(Y[s], ts_cut, freq_c, tau_c, standardize=True)
        assert_same_nan_mask(r.amplitude[s], o[0], "surrogate %d" % s)
        assert max_rel(r.Neffs, o[2]) < 1e-12
        e = max_rel(r.amplitude[s], o[0])
        worst = max(worst, e)
        assert e < 1e-9, "surrogate %d amplitude" % s
        d = np.abs(r.phase[s] - o[1])
        m = ~np.isnan(d)
        assert np.max(np.minimum(d[m], 2 * np.pi - d[m])) < 1e-9
    print("config 3, 8 full-grid surrogates (stored-weights DMMA path) vs the oracle: worst amplitude error %.2e" % worst)


def test_config3_size_significance_test():
    """Config 3 at full size: 10,000 AR(1) surrogates of an n=5,000 series on the default 50 x 501 grid, fused on
    the device.  Size-independent properties: quantiles ordered, mask equal to the target's, reproducible for a
    seed, and the fused result equals the unfused pipeline (same surrogates, scipy mquantiles) on a tau block."""
    from scipy.stats.mstats import mquantiles
    t, y, tau, freq = synth.config_grid("C3")
    qs = [0.5, 0.9, 0.95]
    sg = B.wwz_signif(y, t, tau=tau, freq=freq, number=10000, qs=qs, seed=5)
    assert sg.amplitude_qs.shape == (3, 50, freq.size)
    target = WV.wwz(y, t, tau=tau, freq=freq)
    for q in sg.amplitude_qs:
        assert_same_nan_mask(q, target.amplitude)
    m = ~np.isnan(target.amplitude)
    assert (sg.amplitude_qs[1][m] >= sg.amplitude_qs[0][m]).all() and (sg.amplitude_qs[2][m] >= sg.amplitude_qs[1][m]).all()
    assert_same_nan_mask(sg.Neffs, target.Neffs) and None
    assert max_rel(sg.Neffs, target.Neffs) < 1e-12
    sg2 = B.wwz_signif(y, t, tau=tau, freq=freq, number=10000, qs=qs, seed=5)
    assert np.array_equal(sg.amplitude_qs, sg2.amplitude_qs, equal_nan=True)
    # unfused pipeline on 4 tau rows: same surrogates, single-batch transform, scipy quantiles
    ys_cut, ts_cut, freq_c, tau_c = WV.prepare_wwz(y, t, tau=tau, freq=freq)
    Y = B.ar1_surrogates(ts_cut, B.tau_estimation(ys_cut, ts_cut), sigma=np.std(ys_cut), mu=np.mean(ys_cut),
                         number=10000, seed=5)
    rows = slice(20, 24)
    r = B.wwz_shared_axis(Y, ts_cut, tau_c[rows], freq_c)
    ref = np.asarray(mquantiles(r.amplitude.reshape(10000, -1), qs, axis=0).filled(np.nan)).reshape(3, 4, freq_c.size)
    assert max_rel(sg.amplitude_qs[:, rows], ref) < 1e-10
    # about 5% of the target's cells exceed the 95% level of red noise of the same persistence
    frac = np.mean(target.amplitude[m] > sg.amplitude_qs[2][m])
    assert 0.0 < frac < 0.5


def test_calls_on_different_streams_and_threads_do_not_race_on_the_scratch():
    """the *_dev entry points share one scratch set per device.  Calls enqueued on different streams
    are ordered on the device (each call waits for the end marker of the previous one), and the host entry points hold
    the library lock through their final synchronisation, so concurrent Python threads get the results of serial calls."""
    import threading
    import torch
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    cases = []
    for seed, (n, nt, nf) in enumerate([(900, 64, 40), (700, 48, 33), (1100, 80, 25)]):
        t, y = synth.uneven_ar1(n, 50 + seed)
        tau, freq = np.linspace(t.min(), t.max(), nt), synth.freq_log(t, nf)
        ys, om = WV.standardize(y)[0], WV.make_omega(t, freq)
        serial = WV.kirchner_cuda(ys, t, freq, tau)
        cases.append((t, ys, tau, om, serial))
    streams = [torch.cuda.Stream(device=dev) for _ in cases]
    outs, keep = [], []
    for (t, ys, tau, om, _), st in zip(cases, streams):       # back to back, no synchronisation in between
        d = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (t, ys, tau, om)]
        o = torch.empty((6, tau.size, om.size), dtype=torch.float64, device=dev)
        keep.append(d)
        outs.append(o)
    torch.cuda.synchronize()
    for (t, ys, tau, om, _), st, d, o in zip(cases, streams, keep, outs):
        _lib.check(L.wwzb200_kirchner_dev(d[0].data_ptr(), d[1].data_ptr(), t.size, d[2].data_ptr(), tau.size, d[3].data_ptr(),
                                          om.size, 1 / (8 * np.pi ** 2), 3, 3, 0, *[o[i].data_ptr() for i in range(6)], None,
                                          om.size, st.cuda_stream))
    torch.cuda.synchronize()
    for (_, _, _, _, serial), o in zip(cases, outs):
        g = o.cpu().numpy()
        assert np.array_equal(g[4], serial[0], equal_nan=True) and np.array_equal(g[0], serial[2], equal_nan=True)
    # host entry points from concurrent threads
    rng = np.random.default_rng(3)
    X = [rng.standard_normal((300 + 40 * i, 500)) for i in range(4)]
    want = [B.quantiles(x, [0.5, 0.95]) for x in X]
    tt, _ = synth.uneven_ar1(400, 9)
    want_s = [B.ar1_surrogates(tt, 4.0, number=50 + i, seed=i) for i in range(4)]
    got, got_s = [None] * 4, [None] * 4

    def work(i):
        for _ in range(5):
            got[i] = B.quantiles(X[i], [0.5, 0.95])
            got_s[i] = B.ar1_surrogates(tt, 4.0, number=50 + i, seed=i)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [x.start() for x in th]
    [x.join() for x in th]
    for i in range(4):
        assert np.array_equal(got[i], want[i], equal_nan=True) and np.array_equal(got_s[i], want_s[i])



def test_member_batch_on_device_pointers():
    """wwzb200_kirchner_members_dev (members resident in HBM, caller's stream) equals the host entry point."""
    import torch
    axes, y = synth.age_ensemble(350, 6)
    T = np.stack(axes)
    tau = np.stack([np.linspace(t.min(), t.max(), 19) for t in axes])
    freq = np.stack([synth.freq_log(t, 21) for t in axes])
    host = B.wwz_members(T, y, tau, freq, psd_Neff_threshold=3)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (T, y, tau, freq)]
    out = torch.empty((6, 6, 19, 21), dtype=torch.float64, device=dev)
    psd = torch.empty((6, 21), dtype=torch.float64, device=dev)
    stats = torch.empty((6, 2), dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().wwzb200_kirchner_members_dev(
        d[0].data_ptr(), d[1].data_ptr(), 0, 6, 350, d[2].data_ptr(), 19, d[3].data_ptr(), 21, 1 / (8 * np.pi ** 2), 3, 3,
        _lib.STANDARDIZE | _lib.FREQ_HZ, *[out[i].data_ptr() for i in range(6)], psd.data_ptr(), stats.data_ptr(), st.cuda_stream))
    st.synchronize()
    g, p = out.cpu().numpy(), psd.cpu().numpy()
    for m in range(6):
        assert np.array_equal(g[4, m], host[0][m].amplitude, equal_nan=True)
        assert np.array_equal(g[0, m], host[0][m].Neffs, equal_nan=True)
        assert np.array_equal(p[m], host[1][m], equal_nan=True)
    s = stats.cpu().numpy()
    assert np.allclose(s[:, 0], np.mean(y)) and np.allclose(s[:, 1], np.std(y))
