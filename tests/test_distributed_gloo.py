"""The N > 1 path on the CPU: world_size-2 gloo process groups exercising the tau-block partition,
the ragged all-gather, the psd all-reduce and the sharded significance test.  The per-rank compute is
the oracle (no GPU here); what is under test is the host-side sharding logic."""
import os
import socket

import numpy as np
import pytest

from pyleoclim_util_b200 import distributed as DD


def test_row_blocks_partition():
    for n in (0, 1, 2, 7, 50, 4096):
        for w in (1, 2, 3, 8):
            b = DD.row_blocks(n, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nt, out_dir):
    import warnings
    warnings.filterwarnings("ignore")
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import wwz_oracle as O
        from pyleoclim_util_b200 import synth, distributed as D

        def kernel(ys, ts, freq, tau, c, Neff_threshold, standardize):
            return O.kirchner(ys, ts, freq, tau, c=c, Neff_threshold=Neff_threshold, standardize=standardize, nthreads=1)

        def signif(ys_cut, ts_cut, tau, freq, number, probs, seed, c, Neff_threshold):
            rng = np.random.RandomState(seed)                     # same seed on every rank -> same surrogates
            tau_est = O.tau_estimation(ys_cut, ts_cut)
            amps = []
            for _ in range(number):
                y = O.ar1_model(ts_cut, tau_est, output_sigma=np.std(ys_cut), rng=rng)
                amps.append(O.kirchner(y, ts_cut, freq, tau, c=c, Neff_threshold=Neff_threshold, standardize=True,
                                       nthreads=1)[0])
            amps = np.stack(amps)
            ne = O.kirchner(ys_cut, ts_cut, freq, tau, c=c, Neff_threshold=Neff_threshold, nthreads=1)[2]
            return O.mquantiles(amps.reshape(number, -1), probs, axis=0).reshape(len(probs), tau.size, freq.size), None, ne

        t, y = synth.uneven_ar1(120, 3)
        tau = np.linspace(t.min(), t.max(), nt)
        freq = synth.freq_log(t, 9)
        r = D.wwz_tau_sharded(y, t, tau=tau, freq=freq, kernel=kernel)
        p = D.wwz_psd_tau_sharded(y, t, tau=tau, freq=freq, kernel=kernel)
        s = D.wwz_signif_tau_sharded(y, t, tau=tau, freq=freq, number=4, qs=[0.5, 0.9], seed=5, signif=signif)
        sf = D.wwz_signif_sharded(y, t, tau=tau, freq=freq, number=4, qs=[0.5, 0.9], seed=5, signif=signif, axis="freq")
        assert np.array_equal(s.amplitude_qs, sf.amplitude_qs, equal_nan=True)      # either axis, same quantiles
        assert np.array_equal(s.Neffs, sf.Neffs, equal_nan=True)
        g = D.all_gather_rows(np.full((rank + 1, 2), float(rank)), [1, 2])
        g0 = D.all_gather_rows(np.full((2 * rank, 3), 7.0), [0, 2])            # rank 0 holds an empty block
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), amp=r.amplitude, phase=r.phase, Neffs=r.Neffs, a0=r.coeff[0],
                 coi=r.coi, psd=p.psd, sq=s.amplitude_qs, sne=s.Neffs, g=g, g0=g0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nt", [7, 2])
def test_tau_sharded_world2_matches_unsharded(tmp_path, nt):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nt, str(tmp_path)), nprocs=world, join=True)
    from oracle import wwz_oracle as O
    from pyleoclim_util_b200 import synth
    t, y = synth.uneven_ar1(120, 3)
    tau = np.linspace(t.min(), t.max(), nt)
    freq = synth.freq_log(t, 9)
    ref = O.wwz(y, t, tau=tau, freq=freq, nthreads=1)
    pref = O.wwz_psd(y, t, tau=tau, freq=freq, nthreads=1)
    outs = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for o in outs:
        # same arithmetic, row-sharded (the z-score is taken with NumPy here, with the C loop in the oracle)
        assert np.allclose(o["amp"], ref.amplitude, rtol=1e-12, atol=0, equal_nan=True)
        assert np.allclose(o["phase"], ref.phase, rtol=0, atol=1e-12, equal_nan=True)
        assert np.array_equal(o["Neffs"], ref.Neffs, equal_nan=True)
        assert np.allclose(o["a0"], ref.coeff[0], rtol=0, atol=1e-13, equal_nan=True)
        assert np.array_equal(o["coi"], ref.coi)
        assert np.allclose(o["psd"], pref.psd, rtol=1e-12, atol=0)
        assert o["sq"].shape == (2, nt, freq.size) and np.array_equal(o["sne"], ref.Neffs, equal_nan=True)
        assert np.array_equal(o["g"], np.array([[0, 0], [1, 1], [1, 1]], dtype=float))
        assert np.array_equal(o["g0"], np.full((2, 3), 7.0))
    assert np.array_equal(outs[0]["sq"], outs[1]["sq"], equal_nan=True)
