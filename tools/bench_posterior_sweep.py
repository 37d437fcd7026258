"""BASELINE configs[4]: Poisson-Gamma posterior update (lg_optimize_single, K6) swept over 2^8 .. 2^14 pseudobulk groups
at D = 30000 genes for a 10M-cell collapse.  observed_sum_ds is a Poisson fill with the per-entry mean the collapse of
10M cells x 30k genes at ~4.7% density would give (N * rate_g / S, log-normal gene rates as in data-beans-sim),
size_s = N / S, (a0, b0) = (1, 1).  Reports ms, GB/s of algorithmic bytes 4*D*S*(1 + planes) and the fraction of the
measured HBM peak, for the three calibration targets.  Also a sparse fill (1M cells, the configs[1] regime, where the
small-argument recurrences of digamma / trigamma are exercised) at S = 2^10 and 2^14."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200")]
import numpy as np, torch
import legume_b200 as lg
from legume_b200._lib import lib

D = 30000
peak = 6545.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ctx = lg.Context(0); ctx.use_torch_stream()
g = torch.Generator(device="cuda").manual_seed(0)
rate = 0.05 * torch.exp(torch.randn(D, device="cuda", generator=g) - 0.5)  # per-gene Poisson rate per cell
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2


def one(ncells, S, target):
    lam = (rate * (ncells / S))[None, :].expand(S, D).contiguous()
    sums = torch.poisson(lam, generator=g)
    size = torch.full((S,), ncells / S, device="cuda")
    planes = {0: 4, 1: 1, 2: 2}[target]
    outs = [torch.empty((S, D), device="cuda") for _ in range(planes)]
    mean = outs[0]
    sd = outs[1] if target == 0 else None
    lm = outs[2] if target == 0 else (outs[1] if target == 2 else None)
    ls = outs[3] if target == 0 else None
    p = lambda t: t.data_ptr() if t is not None else None

    def run():
        ctx.check(lib.lg_optimize_single(ctx.h, sums.data_ptr(), size.data_ptr(), D, S, 1.0, 1.0, target, p(mean), p(sd), p(lm), p(ls)))
    for _ in range(3):
        run()
    ts = []
    for _ in range(7):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    nbytes = 4.0 * D * S * (1 + planes)
    # spot check against the closed form of the mean
    want = (1.0 + sums[:2]) / (1.0 + size[:2, None])
    if target == 1:
        want[sums[:2] == 0] = 0.0  # MeanOnly drops the prior baseline where nothing was observed
    assert torch.allclose(mean[:2], want, rtol=1e-6)
    return {"cells": ncells, "S": S, "target": ["All", "MeanOnly", "MeanAndLogMean"][target], "elements": D * S,
            "zero_frac": float((sums == 0).float().mean()), "ms": ms, "GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}


rows = []
for kk in range(8, 15):
    rows.append(one(10_000_000, 1 << kk, 0))
for t in (1, 2):
    rows.append(one(10_000_000, 1 << 14, t))
for kk in (10, 14):
    rows.append(one(1_000_000, 1 << kk, 0))
print(json.dumps({"D": D, "hbm_peak_GBps": peak, "l2": "256 MiB flush between timed launches", "rows": rows}))
