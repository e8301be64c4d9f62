import os, sys, warnings, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
from pyleoclim_util_b200 import synth, wavelet as WV
t, y, tau, freq = synth.config_grid("C2")
for _ in range(3): WV.wwz(y, t, tau=tau, freq=freq)
t0 = time.perf_counter()
for _ in range(10): WV.wwz(y, t, tau=tau, freq=freq)
print("wwz ms", (time.perf_counter() - t0) * 100)
t0 = time.perf_counter()
for _ in range(10): WV.wwz_psd(y, t, tau=tau, freq=freq)
print("wwz_psd ms", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): WV.wwz(y, t, tau=tau, freq=freq)
pr.disable(); pstats.Stats(pr).sort_stats("cumtime").print_stats(14)
