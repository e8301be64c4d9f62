import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
from pyleoclim_util_b200 import _lib, synth, batch as B, wavelet as WV
t, y, tau, freq = synth.config_grid("C3")
ys_cut, ts_cut, fr, ta = WV.prepare_wwz(y, t, tau=tau, freq=freq)
for nf in (1, 8, 63, 501):
    f = fr[:nf]
    B.signif_prepared(ys_cut, ts_cut, ta, f, number=10000, probs=[0.95], seed=1)
    t0 = time.perf_counter()
    for _ in range(3): B.signif_prepared(ys_cut, ts_cut, ta, f, number=10000, probs=[0.95], seed=1)
    print("nf=%d: %.1f ms per call" % (nf, (time.perf_counter() - t0) / 3 * 1e3))
t0 = time.perf_counter(); 
for _ in range(5): B.tau_estimation(ys_cut, ts_cut)
print("tau_estimation %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
t0 = time.perf_counter()
for _ in range(5): WV.prepare_wwz(y, t, tau=tau, freq=freq)
print("prepare_wwz %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
