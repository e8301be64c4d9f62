"""Developer timing helper: tile-shape sweep of the warp-tile recurrence kernel on C2 and a C5 slab."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyleoclim_util_b200 import _lib, synth
from devbench import run
print(torch.cuda.get_device_name(0), "SMs", _lib.sm_count())
t, y, tau, freq = synth.config_grid("C2")
t5, y5, tau5, freq5 = synth.config_grid("C5")
for tg in ("3", "2", "4", "5", "1"):
    os.environ["WWZB200_REC_TG_LOG2"] = tg
    print("TG_LOG2 =", tg)
    run("C2", t, y, tau, freq, flags=0)
    run("C2 c=1e-3", t, y, tau, freq, c=1e-3, flags=0)
    run("C5 slab 256tau", t5, y5, tau5[:256], freq5, flags=0, reps=3)
os.environ.pop("WWZB200_REC_TG_LOG2")
for segs in ("1", "2", "3", "8"):
    os.environ["WWZB200_POINT_SEGMENTS"] = segs
    print("segments =", segs)
    run("C2", t, y, tau, freq, flags=0)
    run("C2 c=1e-3", t, y, tau, freq, c=1e-3, flags=0)
os.environ.pop("WWZB200_POINT_SEGMENTS")
t3, y3, tau3, freq3 = synth.config_grid("C3")
run("C3 single", t3, y3, tau3, freq3, flags=0)
run("C3 single dense", t3, y3, tau3, freq3, flags=1)
