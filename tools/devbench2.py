"""Developer timing helper for the batched paths (config 3 / config 4 shapes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, warnings
warnings.filterwarnings("ignore")
from pyleoclim_util_b200 import _lib, synth, wavelet as WV, batch as B

def t_(f, reps=3):
    f(); ts=[]
    for _ in range(reps):
        t0=time.perf_counter(); r=f(); ts.append(time.perf_counter()-t0)
    return min(ts), r

t, y, tau, freq = synth.config_grid("C3")
print("C3 grid", t.size, tau.size, freq.size)
for S in (256, 2048):
    _lib.profile(True)
    dt, sg = t_(lambda: B.wwz_signif(y, t, tau=tau, freq=freq, number=S, qs=[0.95], seed=1), reps=2)
    ms, nl, cp = _lib.profile_read(); _lib.profile(False)
    print("signif S=%d: %.3f s wall -> %.1f surrogates/s ; shared kernel %.2f ms/launch, %.3e cell-pt-series/s, contraction %.2f TF/s (6 flop)" % (
        S, dt, S/dt, ms/max(nl,1), cp/(ms*1e-3), 6*cp/(ms*1e-3)/1e12))
axes, yy = synth.age_ensemble(1500, 64)
members=[]
for ta in axes:
    ys_cut, ts_cut, fr, ta_ = WV.prepare_wwz(yy, ta)
    members.append((ys_cut, ts_cut, ta_, fr))
dt, r = t_(lambda: B.wwz_ragged(members), reps=2)
print("ragged 64 members: %.3f s -> %.1f members/s" % (dt, 64/dt))
