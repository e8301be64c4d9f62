"""torchrun helper: strong scaling of the 10,000-surrogate significance test (config 3) over the ranks."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np, torch, torch.distributed as dist
from pyleoclim_util_b200 import _lib, synth, batch as B, distributed as D
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); _lib.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t, y, tau, freq = synth.config_grid("C3")
def run(S):
    if world == 1: return B.wwz_signif(y, t, tau=tau, freq=freq, number=S, qs=[0.95], seed=1)
    return D.wwz_signif_sharded(y, t, tau=tau, freq=freq, number=S, qs=[0.95], seed=1)
run(512)
for rep in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = run(10000); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
    if local == 0: print("world %d: 10000 surrogates in %.3f s -> %.0f surrogates/s" % (world, dt, 10000 / dt), flush=True)
if world > 1: dist.destroy_process_group()
