"""pyleoclim_util_b200 -- B200-native weighted wavelet Z-transform (Kirchner WWZ) for Pyleoclim.

One hot path of LinkedEarth/Pyleoclim_util rebuilt as hand-written sm_100a CUDA behind the
reference's own ``method=`` dispatch:

* ``wavelet.kirchner_cuda`` / ``wavelet.wwz`` / ``wavelet.wwz_psd``  -- single series
* ``batch``      -- shared-time-axis surrogate batches, ragged age ensembles, AR(1) surrogates,
                    per-cell quantiles
* ``register()`` -- adds ``method='Kirchner_cuda'`` to ``pyleoclim.utils.wavelet`` when Pyleoclim
                    is importable
* ``_lib``       -- ctypes binding of the C ABI in ``include/wwz_b200.h``

There is no CPU fallback: every compute call goes to ``lib/libwwz_b200.so`` and raises if the
library or a CUDA device is missing.
"""
from . import _lib  # noqa: F401
from .wavelet import (kirchner_cuda, wwz, wwz_psd, wwa2psd, make_omega, make_coi, prepare_wwz,  # noqa: F401
                      make_freq_vector, freq_vector_log)
from .dispatch import register, unregister  # noqa: F401

__version__ = "0.1.0"
