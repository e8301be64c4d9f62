"""Developer timing helper (not the contract bench): kernel time of the transform on device-resident
inputs for a few grid shapes, plus the FP64 peak microbenchmark."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyleoclim_util_b200 import _lib, synth, wavelet as WV

def run(name, t, y, tau, freq, c=1/(8*np.pi**2), flags=0, reps=5):
    dev = torch.device("cuda:0")
    om = WV.make_omega(t, freq)
    ys = (y - y.mean())/y.std()
    d = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (t, ys, tau, om)]
    outs = [torch.empty((tau.size, freq.size), dtype=torch.float64, device=dev) for _ in range(6)]
    st = torch.cuda.current_stream().cuda_stream
    L = _lib.lib()
    def call():
        _lib.check(L.wwzb200_kirchner_dev(d[0].data_ptr(), d[1].data_ptr(), t.size, d[2].data_ptr(), tau.size,
                   d[3].data_ptr(), freq.size, c, 3, 3, flags, *[o.data_ptr() for o in outs], None, 0, st))
    for _ in range(2): call()
    torch.cuda.synchronize()
    ts_ = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts_.append(e0.elapsed_time(e1))
    ms = min(ts_)
    cp = t.size * tau.size * freq.size
    print("%-28s n=%d nt=%d nf=%d flags=%d  %.3f ms (med %.3f)  %.3e cell-pts/s  alg %.2f TF/s  %.3e cells/s" % (
        name, t.size, tau.size, freq.size, flags, ms, float(np.median(ts_)), cp/ms*1e3, 50*cp/ms*1e3/1e12, tau.size*freq.size/ms*1e3), flush=True)
    return ms

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "SMs", _lib.sm_count())
    tf, ms = _lib.measure_fp64_peak(); print("fp64 peak %.2f TFLOP/s (%.3f ms)" % (tf, ms))
    t, y, tau, freq = synth.config_grid("C2")
    for fl in (1, 0):
        run("C2", t, y, tau, freq, flags=fl)
        run("C2 c=1e-3", t, y, tau, freq, c=1e-3, flags=fl)
    for m in ("1", "2", "4"):
        os.environ["WWZB200_CELLS_PER_LANE"] = m
        run("C2 dense M=" + m, t, y, tau, freq, flags=1)
    os.environ.pop("WWZB200_CELLS_PER_LANE")
    t, y, tau, freq = synth.config_grid("C5")
    # a 1/16 tau-slab of C5 (256 tau x 2048 f), dense and windowed
    sl = slice(0, 4096, 16)
    for fl in (1, 0):
        run("C5 slab 256tau", t, y, tau[sl], freq, flags=fl, reps=3)
    t3, y3, tau3, freq3 = synth.config_grid("C3")
    run("C3 single", t3, y3, tau3, freq3, flags=1)
    tf, ms = _lib.measure_fp64_peak(); print("fp64 peak %.2f TFLOP/s (%.3f ms)" % (tf, ms))
