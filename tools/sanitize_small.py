"""Smallest program that runs every kernel once (for compute-sanitizer memcheck)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
from pyleoclim_util_b200 import synth, wavelet as WV, batch as B, _lib
t, y = synth.uneven_ar1(333, 3)
tau = np.linspace(t.min(), t.max(), 37); freq = synth.freq_log(t, 21)
for fl in (0, _lib.DENSE, _lib.NO_RECURRENCE, _lib.FAITHFUL):
    r = WV.wwz(y, t, tau=tau, freq=freq, flags=fl)
p = WV.wwz_psd(y, t, tau=tau, freq=freq)
r = WV.wwz(y, t)                                   # default grid (nt = 50)
g = np.concatenate([np.sort(np.random.default_rng(1).uniform(0, 20, 60)), np.sort(np.random.default_rng(2).uniform(300, 320, 60))])
r = WV.kirchner_cuda(np.random.default_rng(3).standard_normal(120), g, np.logspace(-2, 0, 15), np.linspace(0, 320, 40))
Y = B.ar1_surrogates(t, 3.0, sigma=1.0, number=37, seed=1)
s = B.wwz_shared_axis(Y, t, tau, freq, outputs=("amplitude", "phase", "coeff"), psd_Neff_threshold=3)
q = B.quantiles(s.amplitude, [0.5, 0.95])
sg = B.wwz_signif(y, t, number=40, qs=[0.9], psd=True)
axes, yy = synth.age_ensemble(200, 5)
res = B.wwz_ragged([WV.prepare_wwz(yy, a)[:2] + (np.linspace(a.min(), a.max(), 20), synth.freq_log(a, 9)) for a in axes], psd_Neff_threshold=3)
print("ok", np.nansum(r[0]), np.nansum(q), np.nansum(sg.amplitude_qs))
