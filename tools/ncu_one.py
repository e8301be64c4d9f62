"""Smallest program that launches the hot kernel a few times on config 2 (for ncu --set full)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyleoclim_util_b200 import synth, wavelet as WV, _lib
t, y, tau, freq = synth.config_grid("C2")
ys = WV.standardize(y)[0]
om = WV.make_omega(t, freq)
flags = int(os.environ.get("WWZ_FLAGS", "1"))
for _ in range(4):
    out = WV._transform(ys, t, tau, om, 1 / (8 * np.pi ** 2), 3, flags=flags)
print("ok", np.nansum(out[0]))
