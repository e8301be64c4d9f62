"""The C-ABI shared library loads on a CPU-only host and exports every symbol that
include/wwz_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "wwz_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wwzb200_\w+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ["wwzb200_kirchner", "wwzb200_kirchner_psd", "wwzb200_kirchner_shared_axis",
                 "wwzb200_kirchner_ragged", "wwzb200_ar1_surrogates", "wwzb200_quantiles",
                 "wwzb200_set_device", "wwzb200_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from pyleoclim_util_b200 import build, _lib
    build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(handle, n)]
    assert not missing, "declared in include/wwz_b200.h but not exported: %s" % missing
    assert _lib.lib().wwzb200_version() >= 100


def test_binding_covers_every_declared_symbol():
    from pyleoclim_util_b200 import _lib
    bound = set(_lib._SIGS) | {"wwzb200_last_error"}
    assert set(declared_symbols()) <= bound


def test_no_signature_mentions_torch_or_python():
    src = open(os.path.join(ROOT, "include", "wwz_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in code and "PyObject" not in code and "at::" not in code


def test_bad_arguments_are_rejected_without_a_gpu():
    """Argument checks happen before any CUDA call, so they can be exercised on the CPU."""
    import numpy as np
    from pyleoclim_util_b200 import _lib
    L = _lib.lib()
    a = np.zeros(4)
    out = np.zeros(4)
    rc = L.wwzb200_kirchner(_lib.ptr(a), _lib.ptr(a), 4, _lib.ptr(a), 2, _lib.ptr(a), 2, 0.0126, 0, 0,
                            _lib.ptr(out), None, None, None, None, None)
    assert rc == -1 and b"Neff_threshold" in L.wwzb200_last_error()
    rc = L.wwzb200_kirchner(_lib.ptr(a), _lib.ptr(a), 4, _lib.ptr(a), 2, _lib.ptr(a), 2, -1.0, 3, 0,
                            _lib.ptr(out), None, None, None, None, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "pyleoclim_util_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "wwz_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
