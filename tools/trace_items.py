"""Developer experiment: per-item timeline of the persistent recurrence kernel (library built with -DWWZ_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
path = os.environ.setdefault("WWZB200_TRACE", "/tmp/wwz_trace.bin")
if os.path.exists(path): os.remove(path)
from pyleoclim_util_b200 import _lib, synth
from tools.devbench6 import run
t, y, tau, freq = synth.config_grid("C2")
cases = []
for segs in os.environ.get("SEGS", "2,8").split(","):
    os.environ["WWZB200_POINT_SEGMENTS"] = segs
    for c in (1 / (8 * np.pi ** 2), 1e-3):
        run("C2 c=%.3g segs=%s" % (c, segs), t, y, tau, freq, c=c, reps=1)   # 3 launches each
        cases += [(c, segs)] * 3
run("flush", t, y, tau, freq, reps=1)
raw = np.fromfile(path, dtype=np.uint64)
pos = 0; k = 0
while pos < raw.size and k < len(cases):
    n = int(raw[pos]); rec = raw[pos + 1: pos + 1 + 8 * n].reshape(n, 8).astype(np.int64); pos += 1 + 8 * n
    c, segs = cases[k]; k += 1
    if k % 3: continue   # keep the third launch of each case
    t0 = rec[:, 0].min()
    st, rd, en = rec[:, 0] - t0, rec[:, 1] - t0, rec[:, 2] - t0
    pts = rec[:, 3]; wid = rec[:, 6]
    mhz = np.zeros(n)
    big = pts >= 900
    if big.any():
        inloop_ns = rec[:, 4][big] / 1.965; tot_ns = (en - rd)[big]
        v = rec[:, 5][big]; nch = pts[big] / 32.0
        print('   per chunk (cycles, median): syncwarp %.0f fence %.0f issue %.0f wait %.0f' % (np.median((v >> 48) / nch), np.median(((v >> 32) & 0xffff) / nch), np.median(((v >> 16) & 0xffff) / nch), np.median((v & 0xffff) / nch)))
        print('   items >= 900 pts: point loops take %.1f%% of the item (median); cycles per point inside the loops p10 %.0f p50 %.0f p90 %.0f' % (100 * np.median(inloop_ns / tot_ns), np.percentile(rec[:, 4][big] / pts[big], 10), np.median(rec[:, 4][big] / pts[big]), np.percentile(rec[:, 4][big] / pts[big], 90)))

    dur = en.max()
    print("c=%.3g segs=%s items=%d kernel span %.1f us; total pts %d; setup mean %.2f us (p50 %.2f p99 %.2f); ns/pt busy items p50 %.1f" % (
        c, segs, n, dur / 1e3, pts.sum(), (rd - st).mean() / 1e3, np.median(rd - st) / 1e3, np.percentile(rd - st, 99) / 1e3,
        np.median(((en - rd)[pts > 0]) / pts[pts > 0])))
    first = rec[:, 7] >> 32; epi = rec[:, 7] & 0xffffffff
    print('   first-chunk wait p50 %.2f p90 %.2f us; epilogue p50 %.2f p90 %.2f us (empty items: total p50 %.2f us, %d of them)' % (np.median(first[pts > 0]) / 1e3, np.percentile(first[pts > 0], 90) / 1e3, np.median(epi) / 1e3, np.percentile(epi, 90) / 1e3, np.median((en - st)[pts == 0]) / 1e3 if (pts == 0).any() else 0, (pts == 0).sum()))
    # per-warp finish time and busy points
    uw = np.unique(wid)
    fin = np.array([en[wid == w].max() for w in uw]); wp = np.array([pts[wid == w].sum() for w in uw]); cnt = np.array([(wid == w).sum() for w in uw])
    print("   warps %d: finish p0 %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f us; pts/warp min %d p50 %d max %d; items/warp min %d max %d" % (
        uw.size, fin.min() / 1e3, np.percentile(fin, 10) / 1e3, np.median(fin) / 1e3, np.percentile(fin, 90) / 1e3, fin.max() / 1e3,
        wp.min(), np.median(wp), wp.max(), cnt.min(), cnt.max()))
    # machine occupancy over time: number of items in flight in 20 slices
    edges = np.linspace(0, dur, 21)
    occ = [(np.minimum(en, edges[i + 1]) - np.maximum(st, edges[i])).clip(0).sum() / (edges[i + 1] - edges[i]) for i in range(20)]
    print("   items in flight per 5%% slice: " + " ".join("%d" % o for o in occ))
    ptsrate = [(((np.minimum(en, edges[i + 1]) - np.maximum(rd, edges[i])).clip(0) / np.maximum(en - rd, 1)) * pts).sum() / (edges[i + 1] - edges[i]) * 176 / 1.965 / 592 for i in range(20)]
    print("   FP64 pipe share per slice (176 cyc/pt): " + " ".join("%.2f" % o for o in ptsrate))
    np.save("gpurun_out/trace_c%.3g_s%s.npy" % (c, segs), rec)
