"""Registration of ``method='Kirchner_cuda'`` inside an installed Pyleoclim.

The reference resolves the kernel through ``pyleoclim.utils.wavelet.get_wwz_func``
(utils/wavelet.py:2052-2092), looked up as a module global at call time (:1480), so wrapping that
one function makes ``wwz``, ``wwz_psd`` (utils/spectral.py:889), ``wwz_coherence`` (:1660-1665),
``Series.wavelet`` (core/series.py:3559) and ``Series.spectral`` (core/series.py:3386) accept
``settings={'method': 'Kirchner_cuda'}``; the method name is stored in ``wave_args`` /
``spec_args`` and re-passed by ``signif_test`` (core/scalograms.py:517, core/psds.py:280).
"""
import functools

from . import wavelet as _wavelet

_state = {}


def register(pyleoclim_wavelet=None):
    """Patch ``pyleoclim.utils.wavelet.get_wwz_func`` to serve ``'Kirchner_cuda'``.

    The registered kernel calls the reference's own ``preprocess`` and ``make_omega`` exactly as
    every reference kernel variant does (utils/wavelet.py:976-978).  Returns the patched module.
    Raises ImportError if Pyleoclim is not importable."""
    if pyleoclim_wavelet is None:
        import pyleoclim.utils.wavelet as pyleoclim_wavelet  # noqa: N813
    mod = pyleoclim_wavelet
    if getattr(mod.get_wwz_func, "_wwz_b200", False):
        return mod
    original = mod.get_wwz_func
    kernel = functools.partial(_wavelet.kirchner_cuda, _preprocess=getattr(mod, "preprocess", None),
                               _make_omega=getattr(mod, "make_omega", None))
    functools.update_wrapper(kernel, _wavelet.kirchner_cuda)

    @functools.wraps(original)
    def get_wwz_func(nproc, method):
        if method == "Kirchner_cuda":
            mod.assertPositiveInt(nproc) if hasattr(mod, "assertPositiveInt") else None
            return kernel
        return original(nproc, method)

    get_wwz_func._wwz_b200 = True
    _state[id(mod)] = (mod, original)
    mod.get_wwz_func = get_wwz_func
    mod.kirchner_cuda = kernel
    return mod


def unregister():
    """Undo every ``register()``."""
    for mod, original in list(_state.values()):
        mod.get_wwz_func = original
        if hasattr(mod, "kirchner_cuda"):
            delattr(mod, "kirchner_cuda")
    _state.clear()
