"""Developer tool: per-item timeline of the persistent recurrence kernel on config 2.

Needs the library built with the trace compiled in:
    WWZ_NVCC_EXTRA=-DWWZ_TRACE python -m pyleoclim_util_b200.build --force
    python tools/trace_items.py            # SEGS=2,8 selects the point-segment counts traced
Every work item (warp tile x point segment) records when it was taken from the queue, when its window was known and its
first bulk copies were issued, when its sums were written, the points it walked and the SM cycles spent inside the
point loops (wwz_rec.cu, #ifdef WWZ_TRACE).  Rebuild without the flag afterwards: the trace build is slower."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
path = os.environ.setdefault("WWZB200_TRACE", "/tmp/wwz_trace.bin")
if os.path.exists(path):
    os.remove(path)
from pyleoclim_util_b200 import _lib, synth
from tools.devbench6 import run
FP64_PER_WARP_POINT = 88            # FP64-pipe instructions per (warp, point) in the SASS loop body
CYC = 2 * FP64_PER_WARP_POINT       # pipe cycles per (warp, point): one FP64 warp instruction per 2 clocks per SMSP
t, y, tau, freq = synth.config_grid("C2")
sms = _lib.sm_count()
cases = []
for segs in os.environ.get("SEGS", "2").split(","):
    os.environ["WWZB200_POINT_SEGMENTS"] = segs
    for c in (1 / (8 * np.pi ** 2), 1e-3):
        run("C2 c=%.3g segs=%s" % (c, segs), t, y, tau, freq, c=c, reps=1)   # 3 launches each (2 warm-up + 1)
        cases += [(c, segs)] * 3
run("flush", t, y, tau, freq, reps=1)
raw = np.fromfile(path, dtype=np.uint64)
pos = 0
k = 0
while pos < raw.size and k < len(cases):
    n = int(raw[pos])
    rec = raw[pos + 1: pos + 1 + 8 * n].reshape(n, 8).astype(np.int64)
    pos += 1 + 8 * n
    c, segs = cases[k]
    k += 1
    if k % 3:
        continue   # keep the third launch of each case
    t0 = rec[:, 0].min()
    st, rd, en = rec[:, 0] - t0, rec[:, 1] - t0, rec[:, 2] - t0
    pts, clk, wid = rec[:, 3], rec[:, 4], rec[:, 6]
    first, epi = rec[:, 7] >> 32, rec[:, 7] & 0xffffffff
    busy = pts > 0
    dur = en.max()
    print("c=%.3g, %s point segments: %d items (%d empty), launch %.1f us, %d warp-points" % (c, segs, n, (~busy).sum(), dur / 1e3, pts.sum()))
    print("   per item (median): window + first copies %.2f us, wait for the first chunk %.2f us, epilogue %.2f us; empty item %.2f us"
          % (np.median(rd - st) / 1e3, np.median(first[busy]) / 1e3, np.median(epi) / 1e3,
             np.median((en - st)[~busy]) / 1e3 if (~busy).any() else 0.0))
    big = pts >= 900
    if big.any():
        print("   items of >= 900 points: %.1f%% of their time inside the point loops; SM cycles per point inside the loops: p10 %.0f, median %.0f, p90 %.0f (3 warps per SMSP share one FP64 pipe: %d when it is saturated)"
              % (100 * np.median(clk[big] / 1.965 / (en - rd)[big]), np.percentile(clk[big] / pts[big], 10), np.median(clk[big] / pts[big]),
                 np.percentile(clk[big] / pts[big], 90), 3 * CYC))
    # the first wave: every warp holds its first item; CTAs are resident in launch order (3 per SM)
    fw = busy & (st < 5000)
    cta = wid // 4
    line = []
    for slot in range(3):
        m = fw & (cta >= slot * sms) & (cta < (slot + 1) * sms)
        if m.any():
            line.append("CTAs %d-%d: %.0f" % (slot * sms, (slot + 1) * sms - 1, clk[m].sum() / pts[m].sum()))
    print("   first wave, SM cycles per point by residency order of the CTA: " + "; ".join(line))
    edges = np.linspace(0, dur, 21)
    occ = [(np.minimum(en, edges[i + 1]) - np.maximum(st, edges[i])).clip(0).sum() / (edges[i + 1] - edges[i]) for i in range(20)]
    share = [(((np.minimum(en, edges[i + 1]) - np.maximum(rd, edges[i])).clip(0) / np.maximum(en - rd, 1)) * pts).sum()
             / (edges[i + 1] - edges[i]) * CYC / 1.965 / (4 * sms) for i in range(20)]
    print("   items in flight per 5% slice of the launch:   " + " ".join("%4d" % o for o in occ))
    print("   FP64 pipe share per slice (%d cycles/point): " % CYC + " ".join("%.2f" % o for o in share))
    sm = rec[:, 5]
    tot = np.array([pts[sm == s].sum() for s in range(sms)])
    fin = np.array([en[sm == s].max() if (sm == s).any() else 0 for s in range(sms)])
    print("   per SM: points walked min %d, median %d, max %d; last item done at min %.0f, median %.0f, max %.0f us; points per us min %.0f, median %.0f, max %.0f"
          % (tot.min(), np.median(tot), tot.max(), fin.min() / 1e3, np.median(fin) / 1e3, fin.max() / 1e3,
             (tot / np.maximum(fin, 1)).min() * 1e3, np.median(tot / np.maximum(fin, 1)) * 1e3, (tot / np.maximum(fin, 1)).max() * 1e3))
