"""Developer check: hashes of the outputs of a fixed set of transforms (to compare two builds of the library bit for bit).
   python tools/regress_hash.py out.json [ref.json]"""
import sys, os, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyleoclim_util_b200 import _lib, synth, wavelet as WV, batch as B

def h(*arrs):
    m = hashlib.sha256()
    for a in arrs:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()[:16]

out = {}
t, y, tau, freq = synth.config_grid("C2")
for c in (1 / (8 * np.pi ** 2), 1e-3):
    for fl in (0, _lib.NO_RECURRENCE, _lib.EXACT_WINDOW):
        r = WV.kirchner_cuda(y, t, freq, tau, c=c, flags=fl)
        out["C2 c=%.3g fl=%d" % (c, fl)] = h(r[0], r[1], r[2], *r[3])
r = WV.kirchner_cuda(y, t, freq, tau[:333], flags=0)
out["C2 333 taus"] = h(r[0], r[1], r[2], *r[3])
r = WV.kirchner_cuda(y[:700], t[:700], freq[::3], tau[:97], flags=0)
out["C2 small ragged"] = h(r[0], r[1], r[2], *r[3])
t5, y5, tau5, freq5 = synth.config_grid("C5")
r = WV.kirchner_cuda(y5, t5, freq5[::8], tau5[::16], flags=0)
out["C5 slab"] = h(r[0], r[1], r[2], *r[3])
os.environ["WWZB200_BASIS_BYTES"] = str(64 << 20)
r = WV.kirchner_cuda(y5, t5, freq5[::8], tau5[::16], flags=0)
out["C5 slab, frequency batches"] = h(r[0], r[1], r[2], *r[3])
os.environ.pop("WWZB200_BASIS_BYTES")
# gap series (hiatus), unsorted, NaN omega
tg = np.concatenate([t[:800], t[800:] + 900.0]); 
r = WV.kirchner_cuda(y, tg, freq, np.linspace(tg.min(), tg.max(), 400), flags=0)
out["gap series"] = h(r[0], r[1], r[2], *r[3])
t3, y3, tau3, freq3 = synth.config_grid("C3")
r = WV.kirchner_cuda(y3, t3, freq3, tau3, flags=0)
out["C3 single (50 taus)"] = h(r[0], r[1], r[2], *r[3])
r = WV.kirchner_cuda(y3, t3, freq3, tau3, flags=_lib.FOSTER)
out["C3 single foster"] = h(r[0], r[1], r[2], *r[3])
r = WV.kirchner_cuda(y, t, freq, tau, flags=_lib.FOSTER)
out["C2 foster"] = h(r[0], r[1], r[2], *r[3])
axes, y4 = synth.age_ensemble(1500, 24)
mem = []
for a in axes:
    ys_c, ts_c, fr_c, ta_c = WV.prepare_wwz(y4, a)
    mem.append((ys_c, ts_c, ta_c, fr_c))
res = B.wwz_ragged(mem)
out["C4 24 members"] = h(*[x.amplitude for x in res], *[x.Neffs for x in res])
res = B.wwz_ragged(mem[:5] + [(m[0][:1200], m[1][:1200], m[2], m[3]) for m in mem[5:9]])
out["ragged 9 members"] = h(*[x.amplitude for x in res], *[x.Neffs for x in res])
psd = WV.wwz_psd(y, t, freq=freq, tau=tau)
out["C2 psd"] = h(psd.psd)
json.dump(out, open(sys.argv[1], "w"), indent=1)
if len(sys.argv) > 2:
    ref = json.load(open(sys.argv[2]))
    for k in out:
        print("%-32s %s" % (k, "same" if ref.get(k) == out[k] else "DIFFERENT (%s vs %s)" % (out[k], ref.get(k))))
else:
    for k, v in out.items():
        print("%-32s %s" % (k, v))
