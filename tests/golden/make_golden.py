"""Generate the golden fixtures in tests/golden/ by running the *reference itself*.

Run once in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It imports the reference's own modules through ref_shim.py (no reference source is
copied), runs ``method='Kirchner_numba'`` (the parity target) and writes small ``.npz``
files.  The fixtures travel to the GPU box; the reference tree does not.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from pyleoclim_util_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore")
mods = ref_shim.load()
W, SP, TM = mods["wavelet"], mods["spectral"], mods["tsmodel"]


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-22s %8.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def run_wwz(ys, ts, method="Kirchner_numba", **kw):
    r = W.wwz(ys, ts, method=method, nproc=1, **kw)
    a0, a1, a2 = r.coeff
    return dict(amplitude=r.amplitude, phase=r.phase, Neffs=r.Neffs, a0=a0, a1=a1, a2=a2,
                coi=r.coi, freq=r.freq, tau=r.time)


def benthic():
    """The one known-answer artefact the reference ships: a serialized WWZ Scalogram."""
    d = json.load(open(os.path.join(ref_shim.REF_ROOT, "example_data", "scal_signif_benthic.json")))
    ts = np.array(d["timeseries"]["time"], float)
    ys = np.array(d["timeseries"]["value"], float)
    tau = np.array(d["wave_args"]["tau"], float)
    freq = np.array(d["wave_args"]["freq"], float)
    ref = run_wwz(ys, ts, tau=tau, freq=freq)
    save("benthic", ts=ts, ys=ys, tau_in=tau, freq_in=freq,
         file_amplitude=np.array(d["amplitude"], float), file_Neffs=np.array(d["wwz_Neffs"], float),
         file_coi=np.array(d["coi"], float), **ref)


def soi():
    """Config 1: SOI, default tau / freq grid, wwz and wwz_psd."""
    d = json.load(open(os.path.join(ref_shim.REF_ROOT, "doc_build", "soi.json")))
    ts = np.array(d["time"], float)
    ys = np.array(d["value"], float)
    ref = run_wwz(ys, ts)
    psd = SP.wwz_psd(ys, ts, nproc=1)
    save("soi", ts=ts, ys=ys, psd=psd.psd, psd_freq=psd.freq, **ref)


def synth_small():
    """Small uneven AR(1) case with both decay constants, the numpy kernel beside the numba one,
    a non-default Neff threshold and standardize on/off."""
    t, y = synth.uneven_ar1(300, 10)
    y = 3.0 * y + 7.0                      # non-trivial mean/scale so standardize matters
    tau = np.linspace(t.min(), t.max(), 40)
    freq = synth.freq_log(t, 30)
    ref = run_wwz(y, t, tau=tau, freq=freq)
    basic = run_wwz(y, t, tau=tau, freq=freq, method="Kirchner")
    raw = run_wwz(y, t, tau=tau, freq=freq, standardize=False)
    thr5 = run_wwz(y, t, tau=tau, freq=freq, Neff_threshold=5)
    c3 = run_wwz(y, t, tau=tau, freq=freq, c=1e-3)
    psd = SP.wwz_psd(y, t, freq=freq, tau=tau, nproc=1)
    psd_thr = SP.wwz_psd(y, t, freq=freq, tau=tau, nproc=1, Neff_threshold=6)
    psd_default_tau = SP.wwz_psd(y, t, freq=freq, nproc=1)
    out = dict(ts=t, ys=y, tau_in=tau, freq_in=freq, psd=psd.psd, psd_thr6=psd_thr.psd,
               psd_default_tau=psd_default_tau.psd)
    for tag, r in [("", ref), ("basic_", basic), ("raw_", raw), ("thr5_", thr5), ("c1e3_", c3)]:
        for k in ("amplitude", "phase", "Neffs", "a0", "a1", "a2"):
            out[tag + k] = r[k]
    out["coi"] = ref["coi"]
    save("synth_small", **out)


def edge():
    """Three-state Neff mask: two clusters with a long gap, a frequency above Nyquist."""
    t = np.concatenate([np.arange(10.0), 1000.0 + np.arange(10.0)])
    rng = np.random.default_rng(5)
    y = rng.standard_normal(t.size)
    tau = np.array([0.0, 5.0, 30.0, 60.0, 200.0, 500.0, 1005.0, 1009.0])
    freq = np.array([1e-3, 1e-2, 0.03, 0.05, 0.1, 0.2, 0.5, 0.6])
    ref = W.kirchner_numba(y, t, freq, tau, standardize=True)
    basic = W.kirchner_basic(y, t, freq, tau, standardize=True)
    save("edge", ts=t, ys=y, tau_in=tau, freq_in=freq,
         amplitude=ref[0], phase=ref[1], Neffs=ref[2], a0=ref[3][0], a1=ref[3][1], a2=ref[3][2],
         basic_amplitude=basic[0], basic_Neffs=basic[2])


def c2_subsample():
    """Config 2 (n=2000, 1000 tau x 500 f): full reference run, every 16th tau row stored,
    plus full-grid checksums and the full wwz_psd vector."""
    t, y, tau, freq = synth.config_grid("C2")
    ref = run_wwz(y, t, tau=tau, freq=freq)
    psd = SP.wwz_psd(y, t, freq=freq, tau=tau, nproc=1)
    rows = np.arange(0, tau.size, 16)
    out = dict(rows=rows, psd=psd.psd,
               full_nan_count=np.array([np.isnan(ref["amplitude"]).sum()]),
               full_amp_nansum=np.array([np.nansum(ref["amplitude"])]),
               full_neff_nansum=np.array([np.nansum(ref["Neffs"])]),
               nan_mask_packed=np.packbits(np.isnan(ref["amplitude"])))
    for k in ("amplitude", "phase", "Neffs", "a0", "a1", "a2"):
        out[k] = ref[k][rows]
    save("c2_sub", **out)


def c5_sample():
    """Config 5 inputs (n=100,000): a 24 tau x 40 f sample of the 4096 x 2048 grid."""
    t, y, tau, freq = synth.config_grid("C5")
    jt = np.unique(np.round(np.linspace(0, tau.size - 1, 24)).astype(int))
    kf = np.unique(np.round(np.linspace(0, freq.size - 1, 40)).astype(int))
    # prepare_wwz would re-truncate the series to the span of the sub-sampled tau; the sample
    # keeps both grid ends so the series is unchanged.
    ys_std = (y - y.mean()) / y.std()
    ref = W.kirchner_numba(ys_std, t, freq[kf], tau[jt], standardize=False)
    save("c5_sample", jt=jt, kf=kf, amplitude=ref[0], phase=ref[1], Neffs=ref[2],
         a0=ref[3][0], a1=ref[3][1], a2=ref[3][2])


def surrogates():
    """ar1_model / tau_estimation / ar1_sim on an uneven axis with NumPy's legacy global RNG."""
    t, y = synth.uneven_ar1(400, 21)
    y = 2.5 * y + 1.0
    tau_est = TM.tau_estimation(y, t)
    np.random.seed(1234)
    model = TM.ar1_model(t, 4.0, output_sigma=1.7)
    np.random.seed(99)
    sim = TM.ar1_sim(y, 3, t)
    save("surrogates", ts=t, ys=y, tau_est=np.array([tau_est]), model_seed1234_tau4_sig1p7=model,
         sim_seed99_p3=sim)


def foster():
    """method='Foster' (wwz_basic, utils/wavelet.py:228-381): kernel-level call and wwz / wwz_psd front ends."""
    t, y = synth.uneven_ar1(420, 41)
    y = 1.3 * y - 0.4
    tau = np.linspace(t.min(), t.max(), 23)
    freq = synth.freq_log(t, 19)
    wwa, phase, Neffs, coeff = W.wwz_basic(y, t, freq, tau, c=1 / (8 * np.pi ** 2), Neff_threshold=3, nproc=1,
                                           standardize=True)
    ref = run_wwz(y, t, method="Foster", tau=tau, freq=freq)
    psd = SP.wwz_psd(y, t, freq=freq, tau=tau, method="Foster", nproc=1)
    # a sparse series: cells with few effective points (ill-conditioned normal matrices)
    t2, y2 = synth.uneven_ar1(60, 42)
    tau2 = np.linspace(t2.min(), t2.max(), 11)
    freq2 = synth.freq_log(t2, 13)
    k2 = W.wwz_basic(y2, t2, freq2, tau2, c=1 / (8 * np.pi ** 2), Neff_threshold=3, nproc=1, standardize=True)
    save("foster", ts=t, ys=y, tau_in=tau, freq_in=freq, k_wwa=wwa, k_phase=phase, k_Neffs=Neffs, k_y1=coeff[0],
         k_y2=coeff[1], k_y3=coeff[2], amplitude=ref["amplitude"], phase=ref["phase"], Neffs=ref["Neffs"],
         coi=ref["coi"], psd=psd.psd, psd_freq=psd.freq,
         ts2=t2, ys2=y2, tau2=tau2, freq2=freq2, s_wwa=k2[0], s_phase=k2[1], s_Neffs=k2[2], s_y1=k2[3][0],
         s_y2=k2[3][1], s_y3=k2[3][2])


def freq_vectors():
    """make_freq_vector with each of the reference's generators (utils/wavelet.py:1689-2049), and a wwz run on one."""
    t, y = synth.uneven_ar1(333, 51)
    out = {"ts": t, "ys": y}
    out["lomb_scargle"] = W.make_freq_vector(t, method="lomb_scargle")
    out["lomb_scargle_ofac2_hifac05"] = W.make_freq_vector(t, method="lomb_scargle", ofac=2, hifac=0.5)
    out["welch"] = W.make_freq_vector(t, method="welch")
    out["nfft"] = W.make_freq_vector(t, method="nfft")
    out["scale"] = W.make_freq_vector(t, method="scale")
    out["scale_paul_dj05"] = W.make_freq_vector(t, method="scale", mother="PAUL", dj=0.5)
    out["scale_dog"] = W.make_freq_vector(t, method="scale", mother="DOG")
    out["log"] = W.make_freq_vector(t, method="log")
    out["log_nf20_unbounded"] = W.make_freq_vector(t, method="log", nf=20, rayleigh_bound=False)
    ref = run_wwz(y, t, freq_method="scale", ntau=9)
    save("freq_vectors", wwz_scale_amplitude=ref["amplitude"], wwz_scale_freq=ref["freq"], wwz_scale_tau=ref["tau"], **out)


def uar1():
    """uar1_fit / uar1_sim / the likelihood (utils/tsmodel.py:612-716) with NumPy's legacy global RNG; the
    innovation matrices are stored so that a restatement can be checked draw for draw."""
    t, y = synth.uneven_ar1(300, 31)
    y = 1.8 * y + 0.5
    theta = TM.uar1_fit(y, t)
    nll = TM.n_ll_unevenly_spaced_ar1([np.log(4.0), np.log(1.3)], y, t)
    np.random.seed(4321)
    z1 = np.random.normal(loc=0, scale=1, size=(t.size, 1))
    np.random.seed(4321)
    sim1 = TM.uar1_sim(t, tau=theta[0], sigma_2=theta[1])
    T = np.tile(t, (3, 1)).T
    np.random.seed(77)
    z3 = np.random.normal(loc=0, scale=1, size=T.shape)
    np.random.seed(77)
    sim3 = TM.uar1_sim(T, tau=3.5, sigma_2=2.2)
    save("uar1", ts=t, ys=y, theta_hat=np.asarray(theta), nll_tau4_s1p3=np.array([nll]), z_seed4321=z1[:, 0],
         sim_seed4321=sim1, z_seed77=z3, sim_seed77_tau3p5_s2p2=sim3)


def cleaning():
    """prepare_wwz on NaNs, unsorted stamps, duplicates and a tau that is off the time axis."""
    rng = np.random.default_rng(77)
    t = np.sort(rng.uniform(0, 200, 260))
    y = rng.standard_normal(260)
    t[50] = t[49]                     # duplicate stamp
    y[10] = np.nan
    t[200] = np.nan
    perm = rng.permutation(260)
    t, y = t[perm], y[perm]
    tau = np.linspace(5.3, 190.2, 17)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ys_cut, ts_cut, freq, tau_out = W.prepare_wwz(y, t, tau=tau)
        ys_cut2, ts_cut2, freq2, tau_out2 = W.prepare_wwz(y, t)
        ref = run_wwz(y, t, tau=tau)
    save("cleaning", ts=t, ys=y, tau_in=tau, ys_cut=ys_cut, ts_cut=ts_cut, freq_out=freq,
         tau_out=tau_out, ts_cut_default=ts_cut2, freq_default=freq2, tau_default=tau_out2,
         amplitude=ref["amplitude"], Neffs=ref["Neffs"], phase=ref["phase"])


def coherence():
    """wwz_coherence (utils/wavelet.py:1497-1687) of two uneven series on a common grid, and wtc / xwt (:2201-2364) of
    their coefficients: case A with the default 50 taus (npad = 64, even boxcar), case B with 21 taus and a frequency
    axis whose boxcar length is odd, case C whose second series has a hiatus (Neff-masked cells: NaN propagation
    through the FFT smoothing and the boxcar)."""
    out = {}
    cases = {"A": (260, 240, None, 41), "B": (150, 170, 21, 23), "C": (200, 220, 30, 17)}
    for name, (n1, n2, ntau, nf) in cases.items():
        t1, y1 = synth.uneven_ar1(n1, 31)
        t2, y2 = synth.uneven_ar1(n2, 32)
        if name == "C":
            keep = (t2 < 0.45 * t2.max()) | (t2 > 0.7 * t2.max())
            t2, y2 = t2[keep], y2[keep]
        lo, hi = max(t1.min(), t2.min()), min(t1.max(), t2.max())
        tau = np.linspace(lo, hi, 50 if ntau is None else ntau)
        freq = synth.freq_log(t1, nf)
        r = W.wwz_coherence(y1, t1, y2, t2, tau=tau, freq=freq, nproc=1, method="Kirchner_numba")
        # the grid the reference settled on (:1640-1657 reconcile the two prepared grids) and the coefficients the two
        # transforms return on the truncated series (:1638-1639, :1660-1665)
        y1c, t1c, _, _ = W.prepare_wwz(y1, t1, freq=freq, tau=tau)
        y2c, t2c, _, _ = W.prepare_wwz(y2, t2, freq=freq, tau=tau)
        r1 = W.wwz(y1c, t1c, tau=r.time, freq=r.freq, nproc=1, method="Kirchner_numba")
        r2 = W.wwz(y2c, t2c, tau=r.time, freq=r.freq, nproc=1, method="Kirchner_numba")
        for k, v in dict(t1=t1, y1=y1, t2=t2, y2=y2, tau_in=tau, tau=r.time, freq=r.freq, xw_coherence=r.xw_coherence, xw_phase=r.xw_phase,
                         xw_amplitude=r.xw_amplitude, xwt=r.xwt, coi=r.coi, a1_1=r1.coeff[1], a2_1=r1.coeff[2],
                         a1_2=r2.coeff[1], a2_2=r2.coeff[2]).items():
            out[name + "_" + k] = v
    save("coherence", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["benthic", "soi", "synth_small", "edge", "c2_subsample", "c5_sample",
                             "surrogates", "uar1", "foster", "freq_vectors", "cleaning", "coherence"]
    for name in which:
        globals()[name]()
