"""Rate tables of the synthetic `data-beans-sim topic` counts (data-beans-sim/src/core.rs:253-345): pure numpy, no device
and no library handle, so that the CPU baseline leg of bench.py can load this file on its own (by path) without mapping
the product library.  See sim.py for the generator itself."""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = np.asarray(z, np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


class SimTables:
    def __init__(self, D, ntopic, nbatch, lam, p0, npiece, seed):
        self.D, self.ntopic, self.nbatch, self.seed = D, ntopic, nbatch, seed
        self.lam, self.p0, self.npiece = lam, p0, npiece
        self._dev = None

    def cell_labels(self, col_lo, col_hi):
        """topic and batch of cells [col_lo, col_hi): pure functions of (seed, cell)"""
        j = np.arange(col_lo, col_hi, dtype=np.uint64)
        with np.errstate(over="ignore"):
            ht = _mix64(np.uint64(self.seed) ^ (j * np.uint64(0xA0761D6478BD642F) + np.uint64(1)))
            hb = _mix64(np.uint64(self.seed) ^ (j * np.uint64(0xE7037ED1A0B428DB) + np.uint64(2)))
        topic = ((ht >> np.uint64(33)) % np.uint64(self.ntopic)).astype(np.uint8)
        batch = ((hb >> np.uint64(33)) % np.uint64(self.nbatch)).astype(np.uint8)
        return topic, batch

    def device_tables(self, device):
        import torch
        if self._dev is None:
            self._dev = tuple(torch.from_numpy(a).to(f"cuda:{device}") for a in (self.lam, self.p0, self.npiece))
        return self._dev


def make_tables(D, ntopic=8, nbatch=1, depth=1000, beta_scale=1.0, pve_topic=1.0, pve_batch=1.0, seed=42) -> SimTables:
    rng = np.random.default_rng(seed)
    pvb = min(max(pve_batch, 0.0), 1.0)
    pvt = min(max(pve_topic, 0.0), 1.0)
    # core.rs:89-115 log delta = sqrt(pve) z (z-scored per batch) + sqrt(1-pve) w (z-scored)
    z = rng.standard_normal((D, nbatch))
    z = (z - z.mean(0)) / np.maximum(z.std(0), 1e-12) * np.sqrt(pvb)
    w = rng.standard_normal(D)
    w = (w - w.mean()) / max(w.std(), 1e-12) * np.sqrt(1.0 - pvb)
    delta = np.exp(z + w[:, None]) if nbatch > 1 else np.ones((D, 1))  # core.rs:316 delta only when B > 1
    # core.rs:52-76 log beta = s (sqrt(pve) u + sqrt(1-pve) v) - s^2/2
    v = rng.standard_normal(D)
    u = rng.standard_normal((D, ntopic))
    beta = np.exp(beta_scale * (np.sqrt(pvt) * u + np.sqrt(1.0 - pvt) * v[:, None]) - 0.5 * beta_scale ** 2)
    # core.rs:19-38 theta = pve * onehot + (1 - pve)/K  =>  sum_k beta theta = pve beta[:,k*] + (1-pve) mean_k beta
    mix = pvt * beta + ((1.0 - pvt) * beta.mean(1, keepdims=True) if ntopic > 1 else 0.0)
    rate = np.maximum((depth / D) * delta[None, :, :].transpose(0, 2, 1) * mix.T[:, None, :], 1e-8)  # (topic, batch, D)
    npiece = np.clip(np.ceil(rate / 8.0), 1, 255).astype(np.uint8)
    lam = (rate / npiece).astype(np.float32)
    p0 = np.exp(-lam.astype(np.float64)).astype(np.float32)
    flat = lambda a: np.ascontiguousarray(a.reshape(-1))
    return SimTables(D, ntopic, nbatch, flat(lam), flat(p0), flat(npiece), seed)
