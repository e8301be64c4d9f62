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

def _find_root():
    """$WWZ_REFERENCE_ROOT, else the read-only reference tree of the build container, else the (git-ignored) offline
    install `pip install --no-deps --target baseline/_ref /root/reference`, which travels to the GPU box."""
    env = os.environ.get("WWZ_REFERENCE_ROOT")
    if env:
        return env
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for cand in ("/root/reference", os.path.join(repo, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "pyleoclim")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()
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
    if _full:      # the whole package is already imported (load_full): hand out its modules
        for short in ["wavelet", "spectral", "tsmodel", "tsutils", "tsbase"]:
            _loaded[short] = importlib.import_module("pyleoclim.utils." + short)
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


# ------------------------------------------------------------------------------------------------
# Class-level import (Series, MultipleSeries, Scalogram, PSD ...): a meta-path finder that serves
# permissive stubs for the absent third-party roots, so that `import pyleoclim` itself succeeds.
# ------------------------------------------------------------------------------------------------
_STUB_ROOTS = {"matplotlib", "mpl_toolkits", "seaborn", "cartopy", "shapely", "pyproj", "statsmodels", "pathos",
               "nitime", "kneed", "wget", "bs4", "pyhht", "lipd", "pylipd"}


class _FullStub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name not in ("__mro_entries__",):
            raise AttributeError(name)
        m = _FullStub(self.__name__ + "." + name)
        object.__setattr__(self, name, m)
        return m

    def __call__(self, *a, **k):
        return _FullStub(self.__name__ + "()")

    def __getitem__(self, k):
        return _FullStub(self.__name__ + "[]")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubLoader:
    def create_module(self, spec):
        m = _FullStub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _StubFinder:
    def find_spec(self, name, path=None, target=None):
        root = name.split(".")[0]
        if root in _STUB_ROOTS:
            try:
                # prefer a real installation when there is one
                for f in sys.meta_path:
                    if f is self or not hasattr(f, "find_spec"):
                        continue
                    spec = f.find_spec(name, path, target)
                    if spec is not None:
                        return spec
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(name, _StubLoader(), is_package=True)
        return None


_full = {}


def load_full():
    """`import pyleoclim` (the real package under /root/reference) with stubs for absent deps."""
    if _full:
        return _full["pyleo"]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_PKG)
    for k in [k for k in sys.modules if k == "pyleoclim" or k.startswith("pyleoclim.")]:
        del sys.modules[k]          # drop the namespace shims of load()
    _loaded.clear()
    sys.meta_path.insert(0, _StubFinder())
    import importlib.metadata as md
    _orig_version = md.version

    def version(name):
        if name.lower() == "pyleoclim":
            return "1.3.1b"
        return _orig_version(name)

    md.version = version
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import pyleoclim as pyleo
    _full["pyleo"] = pyleo
    return pyleo
