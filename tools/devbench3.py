import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.devbench as d
from pyleoclim_util_b200 import synth
import numpy as np
t, y, tau, freq = synth.config_grid("C2")
d.run("C2 default", t, y, tau, freq, flags=0)
d.run("C2 no-recurrence", t, y, tau, freq, flags=8)
d.run("C2 c=1e-3 default", t, y, tau, freq, c=1e-3, flags=0)
t, y, tau, freq = synth.config_grid("C5")
d.run("C5 tau[0:256] default", t, y, tau[:256], freq, flags=0, reps=3)
d.run("C5 tau[0:256] no-rec", t, y, tau[:256], freq, flags=8, reps=3)
d.run("C5 tau[0:256] dense", t, y, tau[:256], freq, flags=1, reps=3)
