"""Developer tool: time the pieces of the config-3 significance test (shared-axis kernel via the library profile)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
from pyleoclim_util_b200 import _lib, synth, batch as B
t, y, tau, freq = synth.config_grid("C3")
S = int(os.environ.get("S", "10000"))
flags = int(os.environ.get("WWZ_FLAGS", "0"))
reps = int(os.environ.get("REPS", "2"))
B.wwz_signif(y, t, tau=tau, freq=freq, number=256, qs=[0.95], seed=1, flags=flags)
for rep in range(reps):
    _lib.profile(True)
    t0 = time.perf_counter()
    r = B.wwz_signif(y, t, tau=tau, freq=freq, number=S, qs=[0.95], seed=1, flags=flags)
    dt = time.perf_counter() - t0
    ms, nl, cp = _lib.profile_read()
    _lib.profile(False)
    print("S=%d flags=%d P=%s: %.1f ms total, shared kernel %.1f ms (%d launches), %.2f TFLOP/s algorithmic (6 flop), q95 mean %.9f"
          % (S, flags, os.environ.get("WWZB200_SHARED_POINTS", "-"), dt * 1e3, ms, nl, 6 * cp / (ms * 1e-3) / 1e12, np.nanmean(r.amplitude_qs)), flush=True)
