"""Developer timing: host-to-host wwz / wwz_psd calls on config 2 with and without streamed output; results compared."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
from pyleoclim_util_b200 import synth, wavelet as WV
t, y, tau, freq = synth.config_grid("C2")
def timeit(f, reps=20):
    for _ in range(3): f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts), 1e3 * float(np.median(ts))
ref = None
for label, env in (("separate epilogue, one copy", {"WWZB200_NO_STREAMED_OUTPUT": "1"}),
                   ("fused epilogue, one copy", {"WWZB200_NO_STREAMED_OUTPUT": "1", "WWZB200_FUSED_EPILOGUE": "1"}),
                   ("fused epilogue, streamed", {}),
                   ("fused, streamed, 8 blocks", {"WWZB200_PROGRESS_BLOCKS": "8"}),
                   ("fused, streamed, 32 blocks", {"WWZB200_PROGRESS_BLOCKS": "32"})):
    for k in ("WWZB200_FUSED_EPILOGUE", "WWZB200_NO_STREAMED_OUTPUT", "WWZB200_PROGRESS_BLOCKS"): os.environ.pop(k, None)
    os.environ.update(env)
    r = WV.wwz(y, t, tau=tau, freq=freq)
    arr = np.stack([r.amplitude, r.phase, r.Neffs, *r.coeff])
    if ref is None: ref = arr.copy()
    same = np.array_equal(arr, ref, equal_nan=True)
    a = timeit(lambda: WV.wwz(y, t, tau=tau, freq=freq)); b = timeit(lambda: WV.wwz_psd(y, t, tau=tau, freq=freq))
    print("%-32s wwz %.3f ms (med %.3f)  wwz_psd %.3f ms (med %.3f)  identical to first: %s" % (label, a[0], a[1], b[0], b[1], same), flush=True)
