import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


# ---- parity helpers: tolerances are the ones BASELINE.json's north_star states -------------
AMP_RTOL = 1e-9      # amplitude, PSD, coefficients: relative
PHASE_ATOL = 1e-9    # phase: absolute, modulo 2*pi
NEFF_RTOL = 1e-12    # Neff: relative (it only depends on the weights)


def assert_same_nan_mask(a, b, what=""):
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN mask differs " + what


def max_rel(a, b):
    m = ~np.isnan(b)
    if not m.any():
        return 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a[m] - b[m]) / np.abs(b[m])
    r[(a[m] == b[m])] = 0.0
    return float(np.max(r))


def max_phase_abs(a, b):
    m = ~np.isnan(b)
    if not m.any():
        return 0.0
    d = np.abs(a[m] - b[m])
    d = np.minimum(d, np.abs(2 * np.pi - d))
    return float(np.max(d))


def assert_scalogram_parity(got, ref, amp_rtol=AMP_RTOL, phase_atol=PHASE_ATOL, neff_rtol=NEFF_RTOL, what=""):
    """got/ref: dicts with amplitude, phase, Neffs and optionally a0, a1, a2.

    Bar: identical three-state mask (NaN Neff / masked coefficients / valid), amplitude within
    amp_rtol relative, phase within phase_atol rad (mod 2 pi), Neff within neff_rtol relative."""
    assert_same_nan_mask(got["Neffs"], ref["Neffs"], what + " (Neffs)")
    assert_same_nan_mask(got["amplitude"], ref["amplitude"], what + " (amplitude)")
    assert max_rel(got["Neffs"], ref["Neffs"]) <= neff_rtol, what + " Neffs"
    assert max_rel(got["amplitude"], ref["amplitude"]) <= amp_rtol, what + " amplitude"
    assert max_phase_abs(got["phase"], ref["phase"]) <= phase_atol, what + " phase"
    scale = np.nanmax(np.abs(ref["amplitude"])) if np.isfinite(ref["amplitude"]).any() else 1.0
    for k in ("a0", "a1", "a2"):
        if k in got and k in ref:
            assert_same_nan_mask(got[k], ref[k], what + " (%s)" % k)
            m = ~np.isnan(ref[k])
            if m.any():
                # coefficients pass through zero; compare against the amplitude scale of the grid
                assert np.max(np.abs(got[k][m] - ref[k][m])) <= 10 * amp_rtol * max(scale, 1e-300), what + " " + k


def as_dict(wwa, phase, Neffs, coeff):
    return dict(amplitude=wwa, phase=phase, Neffs=Neffs, a0=coeff[0], a1=coeff[1], a2=coeff[2])


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def has_cuda():
    try:
        from pyleoclim_util_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False
