import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, warnings
warnings.filterwarnings("ignore")
from pyleoclim_util_b200 import _lib, synth, wavelet as WV, batch as B
M = int(os.environ.get("MEMBERS", "256"))
axes, yy = synth.age_ensemble(1500, M)
t0 = time.perf_counter()
members = []
for ta in axes:
    ys_cut, ts_cut, fr, ta_ = WV.prepare_wwz(yy, ta)
    members.append((ys_cut, ts_cut, ta_, fr))
print("prepare_wwz per member: %.1f us" % ((time.perf_counter() - t0) / M * 1e6))
for rep in range(3):
    t0 = time.perf_counter(); r = B.wwz_ragged(members); dt = time.perf_counter() - t0
print("B.wwz_ragged (python wrapper + C): %.3f s -> %.0f members/s" % (dt, M / dt))
# C call alone
ys_l = [WV.standardize(m[0])[0] for m in members]; ts_l = [m[1] for m in members]; tau_l = [m[2] for m in members]
om_l = [WV.make_omega(m[1], m[3]) for m in members]
off = lambda xs: np.concatenate([[0], np.cumsum([x.size for x in xs])]).astype(np.int64)
ts_off, tau_off, om_off = off(ts_l), off(tau_l), off(om_l)
out_off = np.concatenate([[0], np.cumsum([a.size * b.size for a, b in zip(tau_l, om_l)])]).astype(np.int64)
ts_p, ys_p, tau_p, om_p = map(np.concatenate, (ts_l, ys_l, tau_l, om_l))
outs = [np.empty(int(out_off[-1])) for _ in range(6)]
L = _lib.lib()
for flags in (0, 8, 1):
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(L.wwzb200_kirchner_ragged(_lib.ptr(ts_p), _lib.ptr(ys_p), _lib.iptr(ts_off), _lib.ptr(tau_p), _lib.iptr(tau_off),
                   _lib.ptr(om_p), _lib.iptr(om_off), _lib.iptr(out_off), M, 1/(8*np.pi**2), 3, 0, flags, *[_lib.ptr(o) for o in outs], None))
        dt = time.perf_counter() - t0
    cp = sum(t.size * a.size * b.size for t, a, b in zip(ts_l, tau_l, om_l))
    print("C call flags=%d: %.3f s -> %.0f members/s, %.3e cell-pts/s, launches so far %d" % (flags, dt, M / dt, cp / dt, _lib.launch_count()))
