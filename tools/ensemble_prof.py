"""Developer timing: config 4 (1,000 age-ensemble members, n = 1,500, default 50 x 151 grids) through
batch.wwz_ragged -> wwzb200_kirchner_members, host to host, with the split C call / Python."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyleoclim_util_b200 import _lib, synth, wavelet as WV, batch as B
M = int(os.environ.get("M", "1000"))
axes, y = synth.age_ensemble(1500, M)
members = []
for tm in axes:
    ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y, tm)
    members.append((ys_c, ts_c, ta_c, fr_c))
for rep in range(4):
    t0 = time.perf_counter()
    res = B.wwz_ragged(members)
    t1 = time.perf_counter() - t0
    print("wwz_ragged(%d members): %.2f ms  -> %.0f members/s" % (M, t1 * 1e3, M / t1), flush=True)
    del res
T = np.stack([m[1] for m in members]); tau = np.stack([m[2] for m in members])
nfm = max(m[3].size for m in members)
fr = np.full((M, nfm), np.nan)
for i, m in enumerate(members):
    fr[i, : m[3].size] = m[3]
for rep in range(3):
    t0 = time.perf_counter()
    res = B.wwz_members(T, y, tau, fr)
    t1 = time.perf_counter() - t0
    print("wwz_members (arrays, shared values): %.2f ms -> %.0f members/s" % (t1 * 1e3, M / t1), flush=True)
    del res
# the C call alone on preallocated pinned buffers
out = _lib.pinned_empty((6, M, tau.shape[1], fr.shape[1]))
Tp = _lib.pinned_empty(T.shape); Tp[...] = T
taup = _lib.pinned_empty(tau.shape); taup[...] = tau
frp = _lib.pinned_empty(fr.shape); frp[...] = fr
yp = _lib.f64(y)
for fl in (_lib.STANDARDIZE | _lib.FREQ_HZ, _lib.STANDARDIZE | _lib.FREQ_HZ | _lib.NO_RECURRENCE, _lib.STANDARDIZE | _lib.FREQ_HZ | _lib.DENSE):
    for rep in range(3):
        _lib.profile(True)
        t0 = time.perf_counter()
        _lib.check(_lib.lib().wwzb200_kirchner_members(_lib.ptr(Tp), _lib.ptr(yp), 0, M, T.shape[1], _lib.ptr(taup), tau.shape[1],
                   _lib.ptr(frp), fr.shape[1], 1 / (8 * np.pi ** 2), 3, 0, fl, *[_lib.ptr(out[i]) for i in range(6)], None, None))
        t1 = time.perf_counter() - t0
        kms, kn, kcp = _lib.profile_read()
        _lib.profile(False)
    print("C call flags=%d: %.2f ms (cell kernels %.2f ms in %d launches) -> %.0f members/s" % (fl, t1 * 1e3, kms, kn, M / t1), flush=True)
