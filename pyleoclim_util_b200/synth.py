"""Synthetic unevenly sampled AR(1) inputs.

Shared by the golden-vector generator, the parity tests and bench.py so that every leg
sees bit-identical inputs for a given seed.
"""
import numpy as np


def uneven_ar1(n, seed, tau=5.0):
    """t = cumsum(Exp(1)) and the REDFIT AR(1) recurrence of the reference's ``ar1_model``
    (pyleoclim/utils/tsmodel.py:63-67) with persistence ``tau`` and unit output sigma."""
    rng = np.random.default_rng(seed)
    dt = rng.exponential(1.0, n)
    t = np.cumsum(dt)
    z = rng.standard_normal(n)
    y = np.zeros(n)
    for i in range(1, n):
        rho = np.exp(-dt[i] / tau)
        y[i] = rho * y[i - 1] + np.sqrt(1.0 - rho * rho) * z[i]
    return t, y


def freq_log(t, nf):
    """The reference's ``freq_vector_log`` (pyleoclim/utils/wavelet.py:1972-1982) for an explicit nf."""
    dt = np.median(np.diff(t))
    return np.logspace(np.log2(1.0 / np.ptp(t)), np.log2(0.5 / dt), nf, base=2)


def age_ensemble(n, members, seed=2, jitter=0.05):
    """Config 4: one value vector on ``members`` perturbed time axes."""
    rng = np.random.default_rng(seed)
    dt = rng.exponential(1.0, n)
    z = rng.standard_normal(n)
    y = np.zeros(n)
    for i in range(1, n):
        rho = np.exp(-dt[i] / 5.0)
        y[i] = rho * y[i - 1] + np.sqrt(1.0 - rho * rho) * z[i]
    axes = []
    for m in range(members):
        eps = np.random.default_rng(1000 + m).standard_normal(n)
        f = np.clip(1.0 + jitter * eps, 0.05, None)
        axes.append(np.cumsum(dt * f))
    return axes, y


CONFIGS = {
    # name: (n, ntau, nf, seed)
    "C2": (2000, 1000, 500, 0),
    "C3": (5000, 50, None, 1),
    "C5": (100000, 4096, 2048, 3),
}


def config_grid(name):
    """Inputs of a named BASELINE.json config: t, y, tau, freq."""
    n, ntau, nf, seed = CONFIGS[name]
    t, y = uneven_ar1(n, seed)
    tau = np.linspace(t.min(), t.max(), ntau)
    if nf is None:
        nf = n // 10 + 1
    freq = freq_log(t, nf)
    freq = freq[freq >= 1.0 / np.ptp(t)]
    return t, y, tau, freq
