"""torchrun --nproc-per-node N tools/nccl_check.py : the sharded API under NCCL against the unsharded run."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np, torch, torch.distributed as dist
from pyleoclim_util_b200 import _lib, synth, wavelet as WV, distributed as D, batch as B
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); _lib.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t, y = synth.uneven_ar1(800, 4)
tau = np.linspace(t.min(), t.max(), 37); freq = synth.freq_log(t, 29)
r = D.wwz_tau_sharded(y, t, tau=tau, freq=freq)
p = D.wwz_psd_tau_sharded(y, t, tau=tau, freq=freq)
s = D.wwz_signif_tau_sharded(y, t, tau=tau, freq=freq, number=64, qs=[0.9, 0.95], seed=3)
sf = D.wwz_signif_sharded(y, t, tau=tau, freq=freq, number=64, qs=[0.9, 0.95], seed=3, axis="freq")
ss = D.wwz_signif_sharded(y, t, tau=tau, freq=freq, number=333, qs=[0.5, 0.95], seed=4)          # by surrogates (NCCL default)
ss1 = B.wwz_signif(y, t, tau=tau, freq=freq, number=333, qs=[0.5, 0.95], seed=4)
rr = D.wwz_tau_sharded(y, t, tau=tau, freq=freq, host="root")       # rank 0: blocks DMA'd into one shared pinned buffer
rr2 = D.wwz_tau_sharded(y, t, tau=tau, freq=freq, host="root")      # a second result while the first is alive
r1 = WV.wwz(y, t, tau=tau, freq=freq); p1 = WV.wwz_psd(y, t, tau=tau, freq=freq)
s1 = B.wwz_signif(y, t, tau=tau, freq=freq, number=64, qs=[0.9, 0.95], seed=3)
# (a tau block and the full grid use different point segmentations: equal to rounding, not bit for bit)
def relmax(a, b):
    m = ~np.isnan(b)
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300))) if m.any() else 0.0
checks = {
    "wwz amplitude": np.allclose(r.amplitude, r1.amplitude, rtol=1e-12, atol=0, equal_nan=True),
    "wwz Neffs": np.allclose(r.Neffs, r1.Neffs, rtol=1e-13, atol=0, equal_nan=True),
    "wwz phase": np.allclose(r.phase, r1.phase, rtol=0, atol=1e-12, equal_nan=True),
    "wwz a0": np.allclose(r.coeff[0], r1.coeff[0], rtol=0, atol=1e-13, equal_nan=True),
    "psd": np.allclose(p.psd, p1.psd, rtol=1e-12),
    "wwz host=root": (rr.amplitude is None) if dist.get_rank() else
                     (np.array_equal(rr.amplitude, r.amplitude, equal_nan=True) and np.array_equal(rr.Neffs, r.Neffs, equal_nan=True)
                      and np.array_equal(rr2.phase, r.phase, equal_nan=True) and rr2.amplitude is not rr.amplitude
                      and not np.shares_memory(rr2.amplitude, rr.amplitude)),
    "signif tau-sharded": np.allclose(s.amplitude_qs, s1.amplitude_qs, rtol=1e-11, equal_nan=True),
    "signif freq-sharded": np.allclose(sf.amplitude_qs, s1.amplitude_qs, rtol=1e-11, equal_nan=True),
    # same surrogates (Philox streams keyed by the global series index), same per-series arithmetic
    "signif surrogate-sharded mask": np.array_equal(np.isnan(ss.amplitude_qs), np.isnan(ss1.amplitude_qs)),
    "signif surrogate-sharded": relmax(ss.amplitude_qs, ss1.amplitude_qs) < 1e-12,
    "signif surrogate-sharded Neffs": np.allclose(ss.Neffs, ss1.Neffs, rtol=1e-13, equal_nan=True),
}
ok = all(checks.values())
if not ok:
    print("rank", dist.get_rank(), {k: v for k, v in checks.items() if not v},
          "surrogate-sharded relmax", relmax(ss.amplitude_qs, ss1.amplitude_qs), flush=True)
print("rank", dist.get_rank(), "of", dist.get_world_size(), "sharded == unsharded:", ok, flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
