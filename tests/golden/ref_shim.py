"""Import shim that loads the *reference's own* WWZ modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by ``make_golden.py`` (to generate the committed
fixtures) and by the optional ``-m "not gpu"`` cross-checks that are skipped when
the reference tree is absent (it never exists on the GPU box).

The reference package cannot be imported normally in this image (matplotlib,
cartopy, statsmodels, pathos, nitime ... are absent), so the package ``__init__``
files are bypassed: ``pyleoclim`` and ``pyleoclim.utils`` are registered as bare
namespace modules whose ``__path__`` points into the reference tree, and the few
absent third-party roots that ``utils/wavelet.py`` / ``utils/spectral.py`` /
``utils/tsmodel.py`` import at module top are replaced by permissive stubs.  None
of the stubbed modules is on the code path exercised here (Kirchner WWZ, wwa2psd,
uneven-axis ar1_sim).
"""
import importlib
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("WWZ_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "pyleoclim")


def available():
    return os.path.isdir(REF_PKG)


class _Stub(types.ModuleType):
    """Stand-in for an absent third-party module: any attribute is another stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed module %s was called" % self.__name__)


def _stub(dotted):
    parts = dotted.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _Stub(n)
            m.__path__ = []
            m.__spec__ = importlib.machinery.ModuleSpec(n, None, is_package=True)
            sys.modules[n] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[: i - 1])], parts[i - 1], m)


_loaded = {}


def load():
    """Return a dict with the reference modules ``wavelet``, ``spectral``, ``tsmodel``,
    ``tsutils``, ``tsbase`` (the real files under /root/reference/pyleoclim/utils)."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_PKG)
    for n in [
        "pathos.multiprocessing",
        "statsmodels.api",
        "statsmodels.tsa.stattools",
        "statsmodels.tsa.arima_process",
        "statsmodels.tsa.arima.model",
        "statsmodels.multivariate.pca",
        "nitime.algorithms",
    ]:
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
    for name, path in [("pyleoclim", REF_PKG), ("pyleoclim.utils", REF_PKG + "/utils")]:
        if name in sys.modules and not isinstance(sys.modules[name], types.ModuleType):
            continue
        m = types.ModuleType(name)
        m.__path__ = [path]
        m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
        m.__spec__.submodule_search_locations = [path]
        sys.modules[name] = m
    for short in ["wavelet", "spectral", "tsmodel", "tsutils", "tsbase"]:
        _loaded[short] = importlib.import_module("pyleoclim.utils." + short)
    return _loaded
