"""Developer experiment: point-segment count (work-item granularity of the persistent warp-tile kernel) on C2 and a C5 slab."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyleoclim_util_b200 import _lib, synth, wavelet as WV
from tools.devbench6 import run
t, y, tau, freq = synth.config_grid("C2")
shapes = os.environ.get("SHAPES", "4x168,4x255").split(",")
for shape in shapes:
  for var in os.environ.get("VARS", "0,1").split(","):
    os.environ["WWZB200_REC_VARIANT"] = var
    os.environ["WWZB200_REC_SHAPE"] = shape
    print("variant", var)
    for segs in os.environ.get("SEGS2", "2,3,4,6,8").split(","):
        os.environ["WWZB200_POINT_SEGMENTS"] = segs
        run("C2 %s segs=%s" % (shape, segs), t, y, tau, freq)
        run("C2 c=1e-3 %s segs=%s" % (shape, segs), t, y, tau, freq, c=1e-3)
t5, y5, tau5, freq5 = synth.config_grid("C5")
sl = slice(0, 4096, 16)
for shape in shapes:
  for var in os.environ.get("VARS", "0,1").split(","):
    os.environ["WWZB200_REC_VARIANT"] = var
    os.environ["WWZB200_REC_SHAPE"] = shape
    print("variant", var)
    for segs in os.environ.get("SEGS5", "1,4,16,64").split(","):
        os.environ["WWZB200_POINT_SEGMENTS"] = segs
        run("C5 slab 256tau %s segs=%s" % (shape, segs), t5, y5, tau5[sl], freq5, reps=3)
