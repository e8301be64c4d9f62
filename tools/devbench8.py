"""Developer timing: the default path after the round-2 kernel work (C2 both decay constants, full C5, small grids)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyleoclim_util_b200 import _lib, synth, wavelet as WV
from tools.devbench6 import run
t, y, tau, freq = synth.config_grid("C2")
ref = run("C2 dense", t, y, tau, freq, flags=_lib.DENSE)
run("C2", t, y, tau, freq, ref=ref)
run("C2 c=1e-3", t, y, tau, freq, c=1e-3)
t5, y5, tau5, freq5 = synth.config_grid("C5")
for segs in os.environ.get("SEGS5", "0,1,4,16").split(","):
    if segs == "0":
        os.environ.pop("WWZB200_POINT_SEGMENTS", None)
    else:
        os.environ["WWZB200_POINT_SEGMENTS"] = segs
    run("C5 full segs=%s" % segs, t5, y5, tau5, freq5, reps=2)
os.environ.pop("WWZB200_POINT_SEGMENTS", None)
axes, y4 = synth.age_ensemble(1500, 2)
ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y4, axes[0])
refs = run("C4 member dense", ts_c, ys_c, ta_c, fr_c, flags=_lib.DENSE)
run("C4 member", ts_c, ys_c, ta_c, fr_c, ref=refs)
