"""Developer tool (one GPU): time every rank's share of the frequency-sharded significance test one after the other,
to check the cost model of distributed.balanced_freq_groups for a given world size."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
from pyleoclim_util_b200 import _lib, synth, batch as B, wavelet as WV, distributed as D
t, y, tau, freq = synth.config_grid("C3")
ys_cut, ts_cut, fr, ta = WV.prepare_wwz(y, t, tau=tau, freq=freq)
probs = np.array([0.95])
B.signif_prepared(ys_cut, ts_cut, ta, fr, number=10000, probs=probs, seed=1)
for world in (2, 4, 8):
    rows = D.balanced_freq_groups(ts_cut, ta, fr, 1 / (8 * np.pi ** 2), world)
    out = []
    for r in rows:
        B.signif_prepared(ys_cut, ts_cut, ta, fr[r], number=10000, probs=probs, seed=1)
        _lib.profile(True)
        t0 = time.perf_counter()
        B.signif_prepared(ys_cut, ts_cut, ta, fr[r], number=10000, probs=probs, seed=1)
        dt = time.perf_counter() - t0
        ms, _, _ = _lib.profile_read()
        _lib.profile(False)
        out.append((r.size, dt * 1e3, ms))
    print("world %d: rows %s" % (world, [o[0] for o in out]))
    print("   total ms  %s   max %.1f" % (" ".join("%.1f" % o[1] for o in out), max(o[1] for o in out)))
    print("   kernel ms %s   max %.1f  mean %.1f" % (" ".join("%.1f" % o[2] for o in out), max(o[2] for o in out),
                                                     np.mean([o[2] for o in out])))
