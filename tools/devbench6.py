"""Developer experiment (round 2): CTA shape / register budget sweep of the warp-tile recurrence kernel on config 2 and a
config-5 slab, and tile-shape sweep on the small-tau grids (SOI-like config 1, a config-4 member), each checked
against the direct dense kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyleoclim_util_b200 import _lib, synth, wavelet as WV


def run(name, t, y, tau, freq, c=1 / (8 * np.pi ** 2), flags=0, reps=7, ref=None):
    dev = torch.device("cuda:0")
    om = WV.make_omega(t, freq)
    ys = (y - y.mean()) / y.std()
    d = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (t, ys, tau, om)]
    outs = torch.empty((6, tau.size, freq.size), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L = _lib.lib()

    def call():
        _lib.check(L.wwzb200_kirchner_dev(d[0].data_ptr(), d[1].data_ptr(), t.size, d[2].data_ptr(), tau.size,
                                          d[3].data_ptr(), freq.size, c, 3, 3, flags, *[outs[i].data_ptr() for i in range(6)],
                                          None, 0, st))
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    _lib.profile(True)
    ts_ = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts_.append(e0.elapsed_time(e1))
    kms, kn, kcp = _lib.profile_read()
    _lib.profile(False)
    ms = min(ts_)
    res = outs.cpu().numpy()
    err = ""
    if ref is not None:
        m = ~np.isnan(ref[4])
        same = np.array_equal(np.isnan(res[4]), np.isnan(ref[4]))
        e = np.max(np.abs(res[4][m] - ref[4][m]) / np.abs(ref[4][m])) if m.any() else 0.0
        err = " mask_same=%s amp_err=%.2e" % (same, e)
    print("%-34s nt=%d nf=%d fl=%d  %.3f ms (med %.3f) cellk %.3f ms/launch %.3e cells/s%s" % (
        name, tau.size, freq.size, flags, ms, float(np.median(ts_)), kms / max(kn, 1), tau.size * freq.size / ms * 1e3, err), flush=True)
    return res


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "SMs", _lib.sm_count())
    tf, ms = _lib.measure_fp64_peak(); print("fp64 peak %.2f TFLOP/s" % tf)
    t, y, tau, freq = synth.config_grid("C2")
    ref = run("C2 dense direct", t, y, tau, freq, flags=_lib.DENSE)
    ref3 = run("C2 dense direct c=1e-3", t, y, tau, freq, c=1e-3, flags=_lib.DENSE)
    for shape in ("4x168", "4x255", "8x255"):
        os.environ["WWZB200_REC_SHAPE"] = shape
        run("C2 rec " + shape, t, y, tau, freq, ref=ref)
        run("C2 rec c=1e-3 " + shape, t, y, tau, freq, c=1e-3, ref=ref3)
    t5, y5, tau5, freq5 = synth.config_grid("C5")
    sl = slice(0, 4096, 16)
    for shape in ("4x168", "4x255", "8x255"):
        os.environ["WWZB200_REC_SHAPE"] = shape
        run("C5 slab 256tau " + shape, t5, y5, tau5[sl], freq5, reps=3)
    os.environ.pop("WWZB200_REC_SHAPE")

    if os.environ.get('SMALL'):
      pass
    if not os.environ.get("SMALL"): sys.exit(0)
    # ---- small tau grids
    import json
    axes, y4 = synth.age_ensemble(1500, 2)
    ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y4, axes[0])
    t1, y1 = synth.uneven_ar1(1910, 11)
    ys1, ts1, fr1, ta1 = WV.prepare_wwz(y1, t1)
    t3, y3, tau3, freq3 = synth.config_grid("C3")
    for nm, (tt, yy, ta, fr) in {"C4 member": (ts_c, ys_c, ta_c, fr_c), "C1-like": (ts1, ys1, ta1, fr1),
                                 "C3 single": (t3, y3, tau3, freq3)}.items():
        refs = run(nm + " dense direct", tt, yy, ta, fr, flags=_lib.DENSE)
        os.environ["WWZB200_NO_TILE_DIRECT"] = "1"
        os.environ["WWZB200_REC_TG_LOG2"] = "3"
        run(nm + " r1 shapes (cells+rec 8x4)", tt, yy, ta, fr, ref=refs)
        os.environ.pop("WWZB200_NO_TILE_DIRECT")
        for tg in ("0", "1", "2"):
            for dtg in ("0", "1", "2"):
                os.environ["WWZB200_REC_TG_LOG2"] = tg
                os.environ["WWZB200_TILE_TG_LOG2"] = dtg
                run(nm + " rec tg=%s direct tg=%s" % (tg, dtg), tt, yy, ta, fr, ref=refs)
        os.environ.pop("WWZB200_REC_TG_LOG2"); os.environ.pop("WWZB200_TILE_TG_LOG2")
        run(nm + " NO_RECURRENCE (tile direct only)", tt, yy, ta, fr, flags=_lib.NO_RECURRENCE, ref=refs)
