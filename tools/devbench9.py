"""Developer experiment (round 2, session 3): point-segment sweep of the persistent recurrence kernel on config 2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyleoclim_util_b200 import _lib, synth
from tools.devbench6 import run
t, y, tau, freq = synth.config_grid("C2")
ref = run("C2 dense", t, y, tau, freq, flags=_lib.DENSE)
ref3 = run("C2 dense c=1e-3", t, y, tau, freq, c=1e-3, flags=_lib.DENSE)
for segs in os.environ.get("SEGS", "0,1,2,3,4,6,8,12,16").split(","):
    if segs == "0":
        os.environ.pop("WWZB200_POINT_SEGMENTS", None)
    else:
        os.environ["WWZB200_POINT_SEGMENTS"] = segs
    run("C2 segs=%s" % segs, t, y, tau, freq, ref=ref)
    run("C2 c=1e-3 segs=%s" % segs, t, y, tau, freq, c=1e-3, ref=ref3)
