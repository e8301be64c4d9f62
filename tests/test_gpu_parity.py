"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's goldens.

Bar (BASELINE.json north_star): identical Neff mask; amplitude, phase, PSD within 1e-9 of
kirchner_numba (relative for amplitude/PSD, absolute mod 2 pi for phase)."""
import os

import numpy as np
import pytest

from conftest import (golden, assert_scalogram_parity, as_dict, max_rel, assert_same_nan_mask, has_cuda,
                      max_phase_abs)

pytestmark = pytest.mark.gpu

from pyleoclim_util_b200 import wavelet as WV   # noqa: E402
from pyleoclim_util_b200 import _lib, synth     # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert has_cuda(), "these tests need a CUDA device and the built libwwz_b200.so (no CPU fallback)"


def _res(r):
    return as_dict(r.amplitude, r.phase, r.Neffs, r.coeff)


def oracle():
    from oracle import wwz_oracle as O
    return O


# ------------------------------------------------------------------ goldens from the reference
def test_benthic_known_answer_file():
    g = golden("benthic")
    r = WV.wwz(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"])
    assert_same_nan_mask(r.amplitude, g["file_amplitude"])
    assert max_rel(r.amplitude, g["file_amplitude"]) < 1e-9
    assert max_rel(r.Neffs, g["file_Neffs"]) < 1e-12
    assert np.array_equal(r.coi, g["file_coi"])
    assert_scalogram_parity(_res(r), g, what="benthic vs numba")


def test_soi_defaults_and_psd():
    g = golden("soi")
    r = WV.wwz(g["ys"], g["ts"])
    assert np.array_equal(r.freq, g["freq"]) and np.array_equal(r.time, g["tau"])
    assert np.array_equal(r.scale, 1 / g["freq"])
    assert_scalogram_parity(_res(r), g, what="soi")
    p = WV.wwz_psd(g["ys"], g["ts"])
    assert np.array_equal(p.freq, g["psd_freq"])
    assert max_rel(p.psd, g["psd"]) < 1e-9


@pytest.mark.parametrize("tag,kw", [("", {}), ("c1e3_", {"c": 1e-3}), ("thr5_", {"Neff_threshold": 5}),
                                     ("raw_", {"standardize": False})])
def test_synth_small_variants(tag, kw):
    g = golden("synth_small")
    r = WV.wwz(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"], **kw)
    ref = {k: g[tag + k] for k in ("amplitude", "phase", "Neffs", "a0", "a1", "a2")}
    assert_scalogram_parity(_res(r), ref, what="synth_small " + tag)


def test_synth_small_psd():
    g = golden("synth_small")
    p = WV.wwz_psd(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"])
    assert max_rel(p.psd, g["psd"]) < 1e-9
    p = WV.wwz_psd(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"], Neff_threshold=6)
    assert max_rel(p.psd, g["psd_thr6"]) < 1e-9
    p = WV.wwz_psd(g["ys"], g["ts"], freq=g["freq_in"])
    assert max_rel(p.psd, g["psd_default_tau"]) < 1e-9
    # scalogram reuse path (utils/spectral.py:888): wwa2psd alone
    r = WV.wwz(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"], c=1e-3)
    p2 = WV.wwz_psd(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"], wwa=r.amplitude, wwz_Neffs=r.Neffs,
                    wwz_freq=r.freq)
    assert max_rel(p2.psd, g["psd"]) < 1e-9


def test_edge_three_state_mask():
    g = golden("edge")
    out = WV.kirchner_cuda(g["ys"], g["ts"], g["freq_in"], g["tau_in"], standardize=True)
    assert_scalogram_parity(as_dict(*out), g, what="edge")
    assert np.isnan(out[2]).sum() == 13 and np.isnan(out[0]).sum() == 19


def test_wwz_on_a_scale_based_frequency_axis():
    """freq_method='scale' (Torrence-Compo scales, utils/wavelet.py:1825-1933) through the whole front end."""
    g = golden("freq_vectors")
    r = WV.wwz(g["ys"], g["ts"], freq_method="scale", ntau=9)      # ntau is accepted and unused, like in the reference
    assert np.array_equal(r.freq, g["wwz_scale_freq"]) and np.array_equal(r.time, g["wwz_scale_tau"])
    assert_same_nan_mask(r.amplitude, g["wwz_scale_amplitude"])
    assert max_rel(r.amplitude, g["wwz_scale_amplitude"]) < 1e-9


def test_cleaning_inputs():
    g = golden("cleaning")
    r = WV.wwz(g["ys"], g["ts"], tau=g["tau_in"])
    assert_scalogram_parity(dict(amplitude=r.amplitude, phase=r.phase, Neffs=r.Neffs), g, what="cleaning")


def test_config2_full_grid_vs_golden_and_oracle():
    g = golden("c2_sub")
    t, y, tau, freq = synth.config_grid("C2")
    r = WV.wwz(y, t, tau=tau, freq=freq)
    rows = g["rows"]
    got = {k: v[rows] for k, v in _res(r).items()}
    assert_scalogram_parity(got, g, what="C2 golden rows")
    assert np.array_equal(np.packbits(np.isnan(r.amplitude)), g["nan_mask_packed"])
    p = WV.wwz_psd(y, t, tau=tau, freq=freq)
    assert max_rel(p.psd, g["psd"]) < 1e-9
    # every cell against the oracle (CPU, all host threads)
    ro = oracle().wwz(y, t, tau=tau, freq=freq)
    assert_scalogram_parity(_res(r), _res(ro), what="C2 full grid vs oracle")


def test_config5_sample_vs_golden():
    g = golden("c5_sample")
    t, y, tau, freq = synth.config_grid("C5")
    ys = (y - y.mean()) / y.std()
    out = WV.kirchner_cuda(ys, t, freq[g["kf"]], tau[g["jt"]])
    assert_scalogram_parity(as_dict(*out), g, what="C5 sample")


# ------------------------------------------------------------------ kernel variants agree bit for bit
def _grid_case(n=700, nt=37, nf=23, seed=4):
    t, y = synth.uneven_ar1(n, seed)
    y = (y - y.mean()) / y.std()
    tau = np.linspace(t.min(), t.max(), nt)
    freq = synth.freq_log(t, nf)
    return t, y, tau, freq


def _bits(out):
    return [np.ascontiguousarray(a).view(np.uint64) for a in (out[0], out[1], out[2], *out[3])]


def test_windowed_is_bit_identical_to_dense():
    """WWZB200_EXACT_WINDOW skips only points whose weight is exactly +0.0 in the kernel: same bits as dense."""
    t, y, tau, freq = _grid_case(n=3000, nt=64, nf=48)
    a = WV.kirchner_cuda(y, t, freq, tau, flags=_lib.NO_RECURRENCE | _lib.EXACT_WINDOW)
    b = WV.kirchner_cuda(y, t, freq, tau, flags=_lib.DENSE)
    for x, z in zip(_bits(a), _bits(b)):
        assert np.array_equal(x, z)


def test_relative_window_inside_a_hiatus():
    """The point window is measured relative to each cell's own largest weight (it widens by the distance from tau
    to the nearest sample): with tau inside a long hiatus, where every weight of a cell is tiny, the default path
    still agrees with the dense run to rounding -- single-series kernels and the shared-axis batch."""
    from pyleoclim_util_b200 import batch as B
    rng = np.random.default_rng(21)
    t = np.concatenate([np.sort(rng.uniform(0, 300, 700)), np.sort(rng.uniform(520, 900, 800))])
    y = rng.standard_normal(t.size)
    tau = np.linspace(t.min(), t.max(), 96)                 # a quarter of the rows lie inside the 220-wide gap
    freq = np.logspace(-2.6, -0.4, 36)
    d = WV.kirchner_cuda(y, t, freq, tau, standardize=True, flags=_lib.DENSE)
    assert np.sum(~np.isnan(d[0][(tau > 330) & (tau < 490)])) > 200      # valid cells inside the hiatus exist
    for fl in (_lib.NO_RECURRENCE, 0):
        a = WV.kirchner_cuda(y, t, freq, tau, standardize=True, flags=fl)
        assert_scalogram_parity(as_dict(*a), as_dict(*d), amp_rtol=1e-11, phase_atol=1e-11, neff_rtol=1e-12,
                                what="hiatus, flags=%d" % fl)
    Y = np.stack([y, rng.standard_normal(t.size), y[::-1].copy()])
    sa = B.wwz_shared_axis(Y, t, tau, freq, outputs=("amplitude", "phase"))
    sd = B.wwz_shared_axis(Y, t, tau, freq, outputs=("amplitude", "phase"), flags=_lib.DENSE)
    assert_same_nan_mask(sa.amplitude, sd.amplitude)
    assert max_rel(sa.amplitude, sd.amplitude) < 1e-11 and max_rel(sa.Neffs, sd.Neffs) < 1e-12
    assert max_rel(sa.amplitude[0], d[0]) < 1e-10


def test_default_window_is_invisible():
    """The default window drops weights below 2^-100 of the largest weight of a cell (hence of its sum(w)): against
    the dense run the results agree to rounding, with the same mask, for both kernels and both decay constants."""
    for c in (1 / (8 * np.pi ** 2), 1e-3):
        t, y, tau, freq = _grid_case(n=3000, nt=64, nf=48)
        d = WV.kirchner_cuda(y, t, freq, tau, c=c, flags=_lib.DENSE)
        for fl in (_lib.NO_RECURRENCE, 0):
            a = WV.kirchner_cuda(y, t, freq, tau, c=c, flags=fl)
            assert_scalogram_parity(as_dict(*a), as_dict(*d), amp_rtol=1e-11 if fl == 0 else 1e-14,
                                    phase_atol=1e-11 if fl == 0 else 1e-14, neff_rtol=1e-12 if fl == 0 else 1e-15,
                                    what="default window vs dense (flags=%d, c=%g)" % (fl, c))


def test_tau_recurrence_kernel_agrees_with_direct_kernel():
    """Equally spaced tau: rows with kappa*dtau small run the tau-recurrence kernel (one exp2 pair per
    8 cells); it must agree with the direct kernel far inside the parity bar, cell for cell, and keep the
    mask.  Covers a partial last group (nt % 8 != 0), both decay constants and a gap series."""
    for (n, nt, nf, c) in [(3000, 101, 48, 1 / (8 * np.pi ** 2)), (1500, 64, 40, 1e-3), (800, 333, 25, 0.0126)]:
        t, y, tau, freq = _grid_case(n=n, nt=nt, nf=nf)
        a = WV.kirchner_cuda(y, t, freq, tau, c=c)
        b = WV.kirchner_cuda(y, t, freq, tau, c=c, flags=_lib.NO_RECURRENCE)
        assert_scalogram_parity(as_dict(*a), as_dict(*b), amp_rtol=1e-11, phase_atol=1e-11, neff_rtol=1e-12,
                                what="recurrence vs direct %s" % ((n, nt, nf),))
    rng = np.random.default_rng(8)
    t = np.concatenate([np.sort(rng.uniform(0, 50, 150)), np.sort(rng.uniform(400, 450, 150))])
    y = rng.standard_normal(t.size)
    tau = np.linspace(t.min(), t.max(), 200)
    freq = np.logspace(-2.5, 0.05, 40)
    a = WV.kirchner_cuda(y, t, freq, tau, standardize=True)
    o = oracle().kirchner(y, t, freq, tau, standardize=True)
    assert_scalogram_parity(as_dict(*a), as_dict(*o), what="recurrence on a gap series")


@pytest.mark.parametrize("m", ["1", "2", "4"])
def test_cells_per_lane_variants_bit_identical(m, monkeypatch):
    t, y, tau, freq = _grid_case()
    monkeypatch.setenv("WWZB200_CELLS_PER_LANE", "2")
    a = WV.kirchner_cuda(y, t, freq, tau)
    monkeypatch.setenv("WWZB200_CELLS_PER_LANE", m)
    b = WV.kirchner_cuda(y, t, freq, tau)
    for x, z in zip(_bits(a), _bits(b)):
        assert np.array_equal(x, z)


def test_frequency_batching_bit_identical(monkeypatch):
    t, y, tau, freq = _grid_case(n=5000, nt=20, nf=40)
    a = WV.kirchner_cuda(y, t, freq, tau)
    monkeypatch.setenv("WWZB200_BASIS_BYTES", str(2 << 20))     # forces several frequency batches
    b = WV.kirchner_cuda(y, t, freq, tau)
    for x, z in zip(_bits(a), _bits(b)):
        assert np.array_equal(x, z)


def test_faithful_mode_matches_oracle_tightly():
    t, y, tau, freq = _grid_case()
    a = WV.kirchner_cuda(y, t, freq, tau, flags=_lib.FAITHFUL)
    o = oracle().kirchner(y, t, freq, tau)
    assert_scalogram_parity(as_dict(*a), as_dict(*o), amp_rtol=1e-11, phase_atol=1e-11, what="faithful")


# ------------------------------------------------------------------ ragged shapes and edge inputs
@pytest.mark.parametrize("n,nt,nf", [(2, 1, 1), (7, 3, 5), (33, 1, 17), (100, 51, 1), (257, 5, 300), (1000, 50, 101),
                                      (513, 2, 2), (4097, 13, 9)])
def test_ragged_shapes_vs_oracle(n, nt, nf):
    t, y = synth.uneven_ar1(n, 100 + n)
    tau = np.linspace(t.min(), t.max(), nt) if nt > 1 else np.array([t.mean()])
    freq = synth.freq_log(t, nf) if nf > 1 else np.array([0.5 / np.ptp(t) * 4])
    a = WV.kirchner_cuda(y, t, freq, tau, standardize=True, Neff_threshold=1)
    o = oracle().kirchner(y, t, freq, tau, standardize=True, Neff_threshold=1)
    assert_scalogram_parity(as_dict(*a), as_dict(*o), what="shape %s" % ((n, nt, nf),))


def test_empty_grids():
    t, y = synth.uneven_ar1(20, 1)
    a = WV.kirchner_cuda(y, t, np.array([]), np.array([1.0, 2.0]))
    assert a[0].shape == (2, 0)
    a = WV.kirchner_cuda(y, t, np.array([0.1]), np.array([]))
    assert a[0].shape == (0, 1)


def test_unsorted_time_axis():
    """The raw kernel accepts an unsorted axis: windowing falls back to dense on the device."""
    t, y = synth.uneven_ar1(400, 9)
    perm = np.random.default_rng(0).permutation(t.size)
    tau = np.linspace(t.min(), t.max(), 11)
    freq = np.array([0.01, 0.1, 0.3])
    a = WV.kirchner_cuda(y[perm], t[perm], freq, tau)
    o = oracle().kirchner(y[perm], t[perm], freq, tau)
    assert_scalogram_parity(as_dict(*a), as_dict(*o), what="unsorted")
    b = WV.kirchner_cuda(y, t, freq, tau)
    assert max_rel(a[0], b[0]) < 1e-12


def test_frequency_above_nyquist_gives_nan_column():
    """utils/wavelet.py:1229-1232: f > f_Nyquist -> omega = NaN -> NaN Neff and coefficients."""
    t, y = synth.uneven_ar1(400, 9)
    tau = np.linspace(t.min(), t.max(), 11)
    freq = np.array([0.01, 5.0, 0.1, 0.3, 7.0])
    a = WV.kirchner_cuda(y, t, freq, tau)
    o = oracle().kirchner(y, t, freq, tau)
    for col in (1, 4):
        assert np.isnan(a[2][:, col]).all() and np.isnan(a[0][:, col]).all() and np.isnan(a[3][0][:, col]).all()
    assert_scalogram_parity(as_dict(*a), as_dict(*o), what="nyquist")


def test_gap_series_flags_deep_underflow_cells():
    """Long hiatus at high frequency: cells whose weights are all (sub)normal-tiny go through the
    faithful fix-up kernel; all three mask states must match the oracle."""
    rng = np.random.default_rng(8)
    t = np.concatenate([np.sort(rng.uniform(0, 50, 120)), np.sort(rng.uniform(400, 450, 120))])
    y = rng.standard_normal(t.size)
    tau = np.linspace(t.min(), t.max(), 60)
    freq = np.logspace(-2.5, 0.05, 40)
    a = WV.kirchner_cuda(y, t, freq, tau, standardize=True)
    o = oracle().kirchner(y, t, freq, tau, standardize=True)
    assert np.isnan(o[2]).any() and np.isnan(o[0]).sum() > np.isnan(o[2]).sum()
    assert_scalogram_parity(as_dict(*a), as_dict(*o), what="gap")


# ------------------------------------------------------------------ size-independent properties
def test_linearity_in_y_and_neff_independent_of_y_at_config2_size():
    t, y, tau, freq = synth.config_grid("C2")
    rng = np.random.default_rng(2)
    y2 = rng.standard_normal(y.size)
    a = WV.kirchner_cuda(y, t, freq, tau)
    b = WV.kirchner_cuda(y2, t, freq, tau)
    c = WV.kirchner_cuda(y + 2.0 * y2, t, freq, tau)
    assert np.array_equal(a[2].view(np.uint64), b[2].view(np.uint64))      # Neffs do not depend on y
    scale = np.nanmax(c[0])
    for i in range(3):
        lin = a[3][i] + 2.0 * b[3][i]
        m = ~np.isnan(lin)
        assert np.array_equal(np.isnan(lin), np.isnan(c[3][i]))
        assert np.max(np.abs(lin[m] - c[3][i][m])) < 1e-11 * scale


def test_time_origin_shift_invariance_of_amplitude():
    t, y, tau, freq = _grid_case(n=1500, nt=40, nf=30)
    a = WV.kirchner_cuda(y, t, freq, tau)
    b = WV.kirchner_cuda(y, t + 1000.0, freq, tau + 1000.0)
    assert_same_nan_mask(a[0], b[0])
    assert max_rel(b[0], a[0]) < 1e-9


def test_wwa2psd_parts_of_tau_blocks_add_up_to_the_whole():
    """wwzb200_wwa2psd_parts_dev (the tau-sharded wwz_psd): the column sums of the blocks of a tau partition, added and
    divided, are the wwa2psd of the whole scalogram (utils/wavelet.py:1270-1278) -- NaN cells and masked cells included."""
    torch = pytest.importorskip("torch")
    t, y, tau, freq = _grid_case(n=1200, nt=77, nf=41)
    r = WV.wwz(y, t, tau=tau, freq=freq, c=1e-3)
    ref = WV.wwa2psd(r.amplitude, t, r.Neffs, freq=freq, Neff_threshold=3)
    orc = oracle().wwa2psd(r.amplitude, t, r.Neffs, Neff_threshold=3)
    assert max_rel(ref, orc) < 1e-13
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    L = _lib.lib()
    total = torch.zeros((2, freq.size), dtype=torch.float64, device=dev)
    for a, b in ((0, 30), (30, 31), (31, 77)):
        wwa = torch.from_numpy(np.ascontiguousarray(r.amplitude[a:b])).to(dev)
        ne = torch.from_numpy(np.ascontiguousarray(r.Neffs[a:b])).to(dev)
        parts = torch.empty((2, freq.size), dtype=torch.float64, device=dev)
        _lib.check(L.wwzb200_wwa2psd_parts_dev(wwa.data_ptr(), ne.data_ptr(), b - a, freq.size, float(t.max() - t.min()),
                                               t.size, 3, parts.data_ptr(), st))
        total += parts
    got = (total[0] / total[1]).cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and max_rel(got, ref) < 1e-13


def test_profile_counts_the_points_the_recurrence_kernel_walks():
    """wwzb200_profile_walked: with point windows the recurrence kernel evaluates fewer (warp, point) pairs than the
    dense count n * ntau/8 * nf / 32, never more, and none when the recurrence is switched off."""
    t, y, tau, freq = synth.config_grid("C2")
    ys = WV.standardize(y)[0]
    om = WV.make_omega(t, freq)
    dense = t.size * (tau.size / 8) * freq.size / 32
    _lib.profile(True)
    WV._transform(ys, t, tau, om, 1 / (8 * np.pi ** 2), 3)
    walked = _lib.profile_walked()
    _lib.profile(False)
    assert 0.3 * dense < walked <= 1.05 * dense
    _lib.profile(True)
    WV._transform(ys, t, tau, om, 1 / (8 * np.pi ** 2), 3, flags=_lib.NO_RECURRENCE)
    assert _lib.profile_walked() == 0
    _lib.profile(False)


# ------------------------------------------------------------------ device-pointer entry point
def test_dev_entry_point_matches_host_entry_point():
    torch = pytest.importorskip("torch")
    t, y, tau, freq = _grid_case(n=900, nt=30, nf=25)
    om = WV.make_omega(t, freq)
    host = WV._transform(y, t, tau, om, 1 / (8 * np.pi ** 2), 3, psd_threshold=3)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in dict(t=t, y=y, tau=tau, om=om).items()}
    outs = [torch.empty((tau.size, freq.size), dtype=torch.float64, device=dev) for _ in range(6)]
    psd = torch.empty(freq.size, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = _lib.lib().wwzb200_kirchner_dev(d["t"].data_ptr(), d["y"].data_ptr(), t.size, d["tau"].data_ptr(), tau.size,
                                         d["om"].data_ptr(), freq.size, 1 / (8 * np.pi ** 2), 3, 3, 0,
                                         *[o.data_ptr() for o in outs], psd.data_ptr(), 0, st)
    _lib.check(rc)
    torch.cuda.synchronize()
    neffs, a0, a1, a2, wwa, phase = [o.cpu().numpy() for o in outs]
    assert np.array_equal(wwa.view(np.uint64), host[0].view(np.uint64))
    assert np.array_equal(neffs.view(np.uint64), host[2].view(np.uint64))
    assert np.array_equal(phase.view(np.uint64), host[1].view(np.uint64))
    assert max_rel(psd.cpu().numpy(), host[4]) < 1e-14


def test_pinned_outputs_and_launch_counter():
    t, y, tau, freq = _grid_case()
    om = WV.make_omega(t, freq)
    before = _lib.launch_count()
    a = WV._transform(y, t, tau, om, 0.0126, 3, pinned=True)
    assert _lib.launch_count() - before >= 5
    b = WV._transform(y, t, tau, om, 0.0126, 3, pinned=False)
    assert np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))
    del a
    c = WV._transform(y, t, tau, om, 0.0126, 3, pinned=True)      # recycled buffers, same values
    assert np.array_equal(c[0].view(np.uint64), b[0].view(np.uint64))


# ------------------------------------------------------------------ BASELINE.json's full sizes
def test_config5_full_grid_properties_and_golden_sample():
    """Config 5 at full size (n=100,000, 4096 tau x 2047 f = 8.4e6 cells, 8.4e11 cell-points): the cells the
    reference was run on (tests/golden/c5_sample.npz) are picked out of the full-grid result, the three-state
    mask is consistent, Neff does not depend on y, and the default path (windowing + tau-recurrence, frequency
    batches) agrees with the direct kernel on a tau block."""
    g = golden("c5_sample")
    t, y, tau, freq = synth.config_grid("C5")
    ys = (y - y.mean()) / y.std()
    wwa, phase, Neffs, (a0, a1, a2) = WV.kirchner_cuda(ys, t, freq, tau)
    assert wwa.shape == (4096, freq.size)
    sub = np.ix_(g["jt"], g["kf"])
    got = dict(amplitude=wwa[sub], phase=phase[sub], Neffs=Neffs[sub], a0=a0[sub], a1=a1[sub], a2=a2[sub])
    assert_scalogram_parity(got, g, what="C5 full grid, golden cells")
    # mask consistency over all 8.4e6 cells
    masked = ~(Neffs > 3)
    assert np.array_equal(np.isnan(wwa), masked) and np.array_equal(np.isnan(a0), masked)
    assert np.isfinite(wwa[~masked]).all() and (wwa[~masked] >= 0).all()
    # amplitude / phase are consistent with the coefficients everywhere
    m = ~masked
    assert np.max(np.abs(wwa[m] - np.hypot(a1[m], a2[m])) / wwa[m]) < 1e-14
    # a tau block through the direct kernel (one exp2 per (cell, point), dense) against the default path
    blk = slice(1024, 1024 + 64)
    d = WV.kirchner_cuda(ys, t, freq, tau[blk], flags=_lib.DENSE)
    assert_same_nan_mask(d[0], wwa[blk])
    assert max_rel(wwa[blk], d[0]) < 1e-10 and max_rel(Neffs[blk], d[2]) < 1e-12
    # Neff is a property of the time axis alone
    y2 = np.random.default_rng(0).standard_normal(t.size)
    e = WV.kirchner_cuda(y2, t, freq, tau[blk])
    f = WV.kirchner_cuda(ys, t, freq, tau[blk])
    assert np.array_equal(e[2].view(np.uint64), f[2].view(np.uint64))
    assert max_rel(f[2], Neffs[blk]) < 1e-12       # (the full grid uses a different point segmentation)


def test_arbitrary_tau_axes_use_the_direct_kernel_and_match_the_oracle():
    """tau need not be np.linspace: unequal spacing, descending order and nearly-but-not equally spaced axes
    (the tau-recurrence kernel must not take those) all match the oracle."""
    t, y = synth.uneven_ar1(900, 12)
    y = (y - y.mean()) / y.std()
    freq = synth.freq_log(t, 19)
    rng = np.random.default_rng(4)
    lin = np.linspace(t.min(), t.max(), 48)
    cases = {
        "random sorted": np.sort(rng.uniform(t.min(), t.max(), 48)),
        "descending": lin[::-1].copy(),
        "jittered by 1e-9": lin + 1e-9 * rng.standard_normal(48) * np.arange(48),
        "quadratic": t.min() + (t.max() - t.min()) * np.linspace(0, 1, 48) ** 2,
    }
    for name, tau in cases.items():
        a = WV.kirchner_cuda(y, t, freq, tau)
        o = oracle().kirchner(y, t, freq, tau)
        assert_scalogram_parity(as_dict(*a), as_dict(*o), what=name)
        b = WV.kirchner_cuda(y, t, freq, tau, flags=_lib.NO_RECURRENCE)
        for x, z in zip(_bits(a), _bits(b)):          # the recurrence kernel was not used
            assert np.array_equal(x, z), name


def test_config5_stratified_cells_against_the_oracle():
    """Config 5 at full size against the ORACLE on more than 20,000 stratified cells (the golden sample holds 960): 68 tau rows
    -- both ends of the axis, where the point windows are cut by the series' end, and a seeded interior sample -- times
    ~300 frequencies -- the first and last ones and every 7th in between, so every frequency batch of the basis scratch
    and both the long-window (low-frequency) and few-point (high-frequency) tiles are hit.  The worst margins are
    printed; the lowest-amplitude cells of the sample are checked at the same relative bar."""
    from oracle import wwz_oracle as O
    t, y, tau, freq = synth.config_grid("C5")
    ys = (y - y.mean()) / y.std()
    wwa, phase, Neffs, (a0, a1, a2) = WV.kirchner_cuda(ys, t, freq, tau)
    rng = np.random.default_rng(12)
    rows = np.unique(np.concatenate([np.arange(4), tau.size - 1 - np.arange(4), rng.choice(np.arange(4, tau.size - 4), 60, replace=False)]))
    cols = np.unique(np.concatenate([np.arange(8), freq.size - 1 - np.arange(8), np.arange(8, freq.size - 8, 7)]))[:320]
    assert rows.size == 68 and rows.size * cols.size >= 20000
    om = WV.make_omega(t, freq)
    Ne, b0, b1, b2 = O.kirchner_cells(t, ys, np.ascontiguousarray(tau[rows]), np.ascontiguousarray(om[cols]), 1 / (8 * np.pi ** 2), 3, mode=0)
    ref = dict(amplitude=np.sqrt(b1 ** 2 + b2 ** 2), phase=np.arctan2(b2, b1), Neffs=Ne, a0=b0, a1=b1, a2=b2)
    sub = np.ix_(rows, cols)
    got = dict(amplitude=wwa[sub], phase=phase[sub], Neffs=Neffs[sub], a0=a0[sub], a1=a1[sub], a2=a2[sub])
    assert_scalogram_parity(got, ref, what="C5 stratified cells vs oracle")
    amp_err, ph_err = max_rel(got["amplitude"], ref["amplitude"]), max_phase_abs(got["phase"], ref["phase"])
    m = ~np.isnan(ref["amplitude"])
    low = ref["amplitude"] <= np.nanpercentile(ref["amplitude"], 2)
    low_err = max_rel(np.where(low, got["amplitude"], np.nan), np.where(low, ref["amplitude"], np.nan))
    print("config 5, %d stratified cells vs the oracle: amplitude %.2e rel (lowest-amplitude 2%%: %.2e), phase %.2e rad, "
          "%d masked" % (ref["amplitude"].size, amp_err, low_err, ph_err, int((~m).sum())))
    assert low_err < 1e-9



# ------------------------------------------------------------------ streamed output of the host entry points
def _stream_case(monkeypatch, t, y, tau, freq, c=1 / (8 * np.pi ** 2), pinned=True, separate=False):
    """The same host call with the output copied in one piece after the last kernel and streamed block by block
    (fused epilogue of the recurrence kernel + progress flags, wwz_capi.cu); same point segments in both runs."""
    om = WV.make_omega(t, freq)
    monkeypatch.setenv("WWZB200_POINT_SEGMENTS", "8")
    monkeypatch.setenv("WWZB200_NO_STREAMED_OUTPUT", "1")
    one = WV._transform(y, t, tau, om, c, 3, pinned=pinned)
    monkeypatch.delenv("WWZB200_NO_STREAMED_OUTPUT")
    if not separate:
        return one, WV._transform(y, t, tau, om, c, 3, pinned=pinned)
    # six separately allocated outputs (no common pitch): the blocks leave as six copies each
    L = _lib.lib()
    outs = [np.full((tau.size, freq.size), -7.0) for _ in range(6)]
    _lib.check(L.wwzb200_kirchner(_lib.ptr(_lib.f64(t)), _lib.ptr(_lib.f64(y)), t.size, _lib.ptr(_lib.f64(tau)), tau.size,
                                  _lib.ptr(om), freq.size, float(c), 3, 0, *[_lib.ptr(o) for o in outs]))
    neffs, a0, a1, a2, wwa, phase = outs
    return one, (wwa, phase, neffs, (a0, a1, a2), None)


def test_streamed_output_is_the_single_copy_result_at_config2_size(monkeypatch):
    t, y, tau, freq = synth.config_grid("C2")
    ys = WV.standardize(y)[0]
    for c in (1 / (8 * np.pi ** 2), 1e-3):
        one, streamed = _stream_case(monkeypatch, t, ys, tau, freq, c=c)
        for a, b in zip(_bits(one), _bits(streamed)):
            assert np.array_equal(a, b)
    # and against the oracle on rows of both ends and the interior of the tau axis (every block boundary region)
    O = oracle()
    rows = np.r_[0:3, 62:66, 498:502, 997:1000]
    ref = O.kirchner(ys, t, freq, tau[rows], c=1 / (8 * np.pi ** 2), mode=0)
    monkeypatch.delenv("WWZB200_POINT_SEGMENTS")
    got = WV._transform(ys, t, tau, WV.make_omega(t, freq), 1 / (8 * np.pi ** 2), 3)
    assert_scalogram_parity(as_dict(got[0][rows], got[1][rows], got[2][rows], tuple(x[rows] for x in got[3])), as_dict(*ref),
                            what="streamed output, config 2")


def test_streamed_output_with_separately_allocated_outputs(monkeypatch):
    t, y, tau, freq = synth.config_grid("C2")
    ys = WV.standardize(y)[0]
    one, streamed = _stream_case(monkeypatch, t, ys, tau[:640], freq, pinned=False, separate=True)
    for a, b in zip(_bits(one), _bits(streamed)):
        assert np.array_equal(a, b)


def test_streamed_output_falls_back_when_other_kernels_own_cells(monkeypatch):
    """Blocks that left early are only final if the recurrence kernel owned every row and no cell went to the faithful
    kernel; otherwise everything is copied again at the end.  (a) a coarse tau axis: the highest rows exceed
    kappa*dtau = 8 and belong to the direct kernel; (b) a hiatus: deep-underflow cells recomputed by the faithful kernel."""
    t, y, tau, freq = synth.config_grid("C2")
    ys = WV.standardize(y)[0]
    tau_coarse = np.linspace(t.min(), t.max(), 1200)[::4]              # 300 taus, dtau ~ 6.7
    freq_big = synth.freq_log(t, 2000)                                  # 300 x 2000 cells = 28.8 MB of output
    one, streamed = _stream_case(monkeypatch, t, ys, tau_coarse, freq_big)
    for a, b in zip(_bits(one), _bits(streamed)):
        assert np.array_equal(a, b)
    om = WV.make_omega(t, freq_big)
    kap_d = 4 * om * np.sqrt(np.log2(np.e) / (8 * np.pi ** 2)) * (tau_coarse[1] - tau_coarse[0])
    assert (kap_d > 8).any() and (kap_d <= 8).any()                     # both kernels had rows
    tg = np.concatenate([t[:800], t[800:] + 900.0])
    tau_g = np.linspace(tg.min(), tg.max(), 640)
    one, streamed = _stream_case(monkeypatch, tg, ys, tau_g, freq)
    for a, b in zip(_bits(one), _bits(streamed)):
        assert np.array_equal(a, b)
    assert np.isnan(one[0]).any()                                       # the hiatus masks cells


# ------------------------------------------------------------------ tabulated growth of rho (recurrence flavour, long series)
def test_rho_growth_table_agrees_with_the_exact_pair(monkeypatch):
    """Long series advance rho = 2^(a D/8) from point to point with a tabulated factor (one exp2 chain per point instead of
    two), re-evaluated exactly at every chunk start; chunks in which a lane's exponent could leave the range of a double
    (a hiatus) and unsorted axes take the exact pair.  Forced on here at config 2 size and on a series with a hiatus."""
    t, y, tau, freq = synth.config_grid("C2")
    ys = WV.standardize(y)[0]
    tg = np.concatenate([t[:800], t[800:] + 5000.0])                     # a hiatus 2.5 times the span of the series
    cases = [(t, tau), (tg, np.linspace(tg.min(), tg.max(), 640))]
    for tt, ta in cases:
        monkeypatch.setenv("WWZB200_RHO_GROWTH", "0")
        exact = WV.kirchner_cuda(ys, tt, freq, ta)
        monkeypatch.setenv("WWZB200_RHO_GROWTH", "1")
        table = WV.kirchner_cuda(ys, tt, freq, ta)
        assert_same_nan_mask(table[0], exact[0])
        assert_same_nan_mask(table[2], exact[2])
        assert max_rel(table[0], exact[0]) < 1e-11
        assert max_rel(table[2], exact[2]) < 1e-12
        assert max_phase_abs(table[1], exact[1]) < 1e-11
    O = oracle()
    rows = np.r_[0:2, 499:501, 998:1000]
    ref = O.kirchner(ys, t, freq, tau[rows], mode=0)
    got = WV.kirchner_cuda(ys, t, freq, tau)
    assert_scalogram_parity(as_dict(got[0][rows], got[1][rows], got[2][rows], tuple(x[rows] for x in got[3])), as_dict(*ref),
                            what="rho growth table, config 2")
    # an unsorted axis never uses the table (the growth factors assume ascending time)
    perm = np.random.default_rng(0).permutation(t.size)
    a = WV.kirchner_cuda(ys[perm], t[perm], freq, tau[:256])
    monkeypatch.setenv("WWZB200_RHO_GROWTH", "0")
    b = WV.kirchner_cuda(ys[perm], t[perm], freq, tau[:256])
    for x, z in zip(_bits(a), _bits(b)):
        assert np.array_equal(x, z)
