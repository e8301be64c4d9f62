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
r1 = WV.wwz(y, t, tau=tau, freq=freq); p1 = WV.wwz_psd(y, t, tau=tau, freq=freq)
s1 = B.wwz_signif(y, t, tau=tau, freq=freq, number=64, qs=[0.9, 0.95], seed=3)
ok = (np.array_equal(r.amplitude, r1.amplitude, equal_nan=True) and np.array_equal(r.Neffs, r1.Neffs, equal_nan=True)
      and np.allclose(p.psd, p1.psd, rtol=1e-12) and np.allclose(s.amplitude_qs, s1.amplitude_qs, rtol=1e-12, equal_nan=True))
print("rank", dist.get_rank(), "of", dist.get_world_size(), "sharded == unsharded:", ok, flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
