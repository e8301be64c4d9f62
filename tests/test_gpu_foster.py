"""Foster's projection (the reference's method='Foster', wwz_basic, utils/wavelet.py:228-381 -- SURVEY.md 8f rank 4) on
the GPU against outputs of the reference itself (tests/golden/foster.npz) and against the oracle restatement.

Bar: identical Neff mask; amplitude / coefficients within 1e-9 relative, phase within 1e-9 (mod 2 pi)."""
import numpy as np
import pytest

from conftest import golden, max_rel, assert_same_nan_mask, has_cuda, max_phase_abs

pytestmark = pytest.mark.gpu

from pyleoclim_util_b200 import wavelet as WV, batch as B, _lib, synth   # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert has_cuda(), "these tests need a CUDA device and the built libwwz_b200.so (no CPU fallback)"


def oracle():
    from oracle import wwz_oracle as O
    return O


def _check(got, ref_wwa, ref_phase, ref_neffs, ref_coeff, tol=1e-9):
    wwa, phase, neffs, coeff = got
    assert_same_nan_mask(wwa, ref_wwa)
    assert_same_nan_mask(neffs, ref_neffs)
    assert max_rel(neffs, ref_neffs) < 1e-12
    assert max_rel(wwa, ref_wwa) < tol
    assert max_phase_abs(phase, ref_phase) < tol
    scale = np.nanmax(np.abs(ref_wwa))
    for c, r in zip(coeff, ref_coeff):
        assert_same_nan_mask(c, r)
        assert np.nanmax(np.abs(c - r)) < tol * scale


@pytest.mark.parametrize("flags", [0, _lib.DENSE, _lib.NO_RECURRENCE, _lib.FAITHFUL])
def test_foster_kernel_vs_reference_outputs(flags):
    g = golden("foster")
    got = WV.foster_cuda(g["ys"], g["ts"], g["freq_in"], g["tau_in"], standardize=True, flags=flags)
    _check(got, g["k_wwa"], g["k_phase"], g["k_Neffs"], (g["k_y1"], g["k_y2"], g["k_y3"]))


def test_foster_sparse_series_ill_conditioned_cells():
    g = golden("foster")
    got = WV.foster_cuda(g["ys2"], g["ts2"], g["freq2"], g["tau2"], standardize=True)
    _check(got, g["s_wwa"], g["s_phase"], g["s_Neffs"], (g["s_y1"], g["s_y2"], g["s_y3"]), tol=1e-8)


def test_foster_front_ends():
    g = golden("foster")
    r = WV.wwz(g["ys"], g["ts"], tau=g["tau_in"], freq=g["freq_in"], method="Foster_cuda")
    assert_same_nan_mask(r.amplitude, g["amplitude"])
    assert max_rel(r.amplitude, g["amplitude"]) < 1e-9 and max_phase_abs(r.phase, g["phase"]) < 1e-9
    assert np.array_equal(r.coi, g["coi"])
    p = WV.wwz_psd(g["ys"], g["ts"], freq=g["freq_in"], tau=g["tau_in"], method="Foster_cuda")
    assert np.array_equal(p.freq, g["psd_freq"]) and max_rel(p.psd, g["psd"]) < 1e-9
    with pytest.raises(ValueError):
        WV.get_wwz_func(1, "Foster")


def test_foster_larger_grid_vs_oracle_and_differs_from_kirchner():
    """n = 1500, 60 x 80 cells: all three kernel flavours against the NumPy restatement; and the result is not
    Kirchner's (the two differ by up to tens of percent)."""
    O = oracle()
    t, y = synth.uneven_ar1(1500, 8)
    tau = np.linspace(t.min(), t.max(), 60)
    freq = synth.freq_log(t, 80)
    ref = O.foster(y, t, freq, tau, standardize=True)
    for flags in (0, _lib.DENSE):
        got = WV.foster_cuda(y, t, freq, tau, standardize=True, flags=flags)
        _check(got, ref[0], ref[1], ref[2], ref[3])
    kir = WV.kirchner_cuda(y, t, freq, tau, standardize=True)
    m = ~np.isnan(ref[0])
    assert np.array_equal(np.isnan(kir[0]), np.isnan(got[0]))
    assert np.max(np.abs(kir[0][m] - got[0][m]) / got[0][m]) > 1e-3


def test_foster_ragged_members_and_shared_axis_refusal():
    O = oracle()
    members = []
    for seed, n in ((1, 200), (2, 260), (3, 90)):
        t, y = synth.uneven_ar1(n, seed)
        members.append((y, t, np.linspace(t.min(), t.max(), 9 + seed), synth.freq_log(t, 11)))
    res = B.wwz_ragged(members, flags=_lib.FOSTER)
    for (y, t, tau, freq), r in zip(members, res):
        ref = O.foster(y, t, freq, tau, standardize=True)
        _check((r.amplitude, r.phase, r.Neffs, r.coeff), ref[0], ref[1], ref[2], ref[3])
    t, y, tau, freq = members[0]
    with pytest.raises(ValueError):
        B.wwz_shared_axis(np.stack([y, y]), t, tau, freq, flags=_lib.FOSTER)
