"""Developer check: fused epilogue against the separate epilogue, cell by cell."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyleoclim_util_b200 import _lib, synth, wavelet as WV
t, y, tau, freq = synth.config_grid("C2")
os.environ["WWZB200_FUSED_EPILOGUE"] = "0"
a = WV.kirchner_cuda(y, t, freq, tau)
os.environ["WWZB200_FUSED_EPILOGUE"] = "1"
b = WV.kirchner_cuda(y, t, freq, tau)
names = ["wwa", "phase", "Neffs", "a0", "a1", "a2"]
arrs_a = [a[0], a[1], a[2], *a[3]]; arrs_b = [b[0], b[1], b[2], *b[3]]
for nm, x, z in zip(names, arrs_a, arrs_b):
    d = ~((x == z) | (np.isnan(x) & np.isnan(z)))
    print(nm, "cells differing:", d.sum(), "of", d.size)
    if d.any():
        jj, kk = np.nonzero(d)
        print("   tau idx range", jj.min(), jj.max(), "freq idx range", kk.min(), kk.max(), "max rel", np.nanmax(np.abs(x[d] - z[d]) / np.abs(x[d])))
        print("   first few:", list(zip(jj[:8], kk[:8])), x[d][:4], z[d][:4])
