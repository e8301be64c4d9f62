"""Developer tool: where the time of the config-4 ragged batch goes (Python wrapper vs the C call)."""
import os, sys, time, warnings, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
from pyleoclim_util_b200 import _lib, synth, batch as B, wavelet as WV
axes, y4 = synth.age_ensemble(1500, 1000)
members = []
for tm in axes:
    ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y4, tm)
    members.append((ys_c, ts_c, ta_c, fr_c))
B.wwz_ragged(members[:64])
for rep in range(2):
    t0 = time.perf_counter(); r = B.wwz_ragged(members); dt = time.perf_counter() - t0
    print("wwz_ragged 1000 members: %.1f ms" % (dt * 1e3))
L = _lib.lib()
orig = L.wwzb200_kirchner_ragged
tc = [0.0]
def timed(*a):
    t0 = time.perf_counter(); rc = orig(*a); tc[0] += time.perf_counter() - t0; return rc
class Shim:
    def __getattr__(self, k): return timed if k == "wwzb200_kirchner_ragged" else getattr(L, k)
_lib_lib = _lib.lib
_lib.lib = lambda: Shim()
t0 = time.perf_counter(); r = B.wwz_ragged(members); dt = time.perf_counter() - t0
print("total %.1f ms, C call %.1f ms" % (dt * 1e3, tc[0] * 1e3))
_lib.lib = _lib_lib
pr = cProfile.Profile(); pr.enable(); B.wwz_ragged(members); pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(12)
