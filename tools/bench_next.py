"""The nnz streams either side of the hot path at the configs[1] shape (30000 genes x 1M cells): K10 per-gene running
statistics (lg_row_stats; HBM-bound, algorithmic bytes 8*nnz + 8*(N+1) per pass) and K11 Nystrom re-projection
(lg_nystrom_project; CUDA-core gather-FMA), with the oracle port timed on a sample of the same cells beside them."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200")]
import numpy as np, torch
import legume_b200 as lg
from legume_b200 import sim
from legume_b200._lib import lib

_nums = [a for a in sys.argv[1:] if a.isdigit()]
N = int(_nums[0]) if _nums else 1_000_000
D, K, S = 30000, 50, 1024
peak = 6545.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ctx = lg.Context(0); ctx.use_torch_stream()
tabs = sim.make_tables(D, ntopic=8, nbatch=1, depth=1500, seed=42)
blk, _, _ = sim.sim_block(ctx, tabs, 0, N)
nnz = blk.nnz
g = torch.Generator(device="cuda").manual_seed(0)
basis = torch.randn((K, D), device="cuda", generator=g)
delta = torch.exp(0.3 * torch.randn((S, D), device="cuda", generator=g)).contiguous()
pb = torch.randint(0, S, (N,), device="cuda", generator=g, dtype=torch.int32)
pb_sorted = torch.sort(pb).values.contiguous()  # cells ordered by pseudobulk: the delta column stays cache-resident


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


st = torch.empty((3, D), dtype=torch.float64, device="cuda")
ms_rs = timed(lambda: ctx.check(lib.lg_row_stats(ctx.h, blk.h, st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr())))
out = torch.empty((N, K), device="cuda")
ms_ny = timed(lambda: ctx.check(lib.lg_nystrom_project(ctx.h, blk.h, basis.data_ptr(), K, None, None, 0, 1e4, out.data_ptr())), 3)
ms_nyd = timed(lambda: ctx.check(lib.lg_nystrom_project(ctx.h, blk.h, basis.data_ptr(), K, delta.data_ptr(), pb.data_ptr(), S, 1e4,
                                                        out.data_ptr())), 3)
ms_nys = timed(lambda: ctx.check(lib.lg_nystrom_project(ctx.h, blk.h, basis.data_ptr(), K, delta.data_ptr(), pb_sorted.data_ptr(), S,
                                                        1e4, out.data_ptr())), 3)
bytes_rs = 8.0 * nnz + 8.0 * (N + 1)
bytes_ny = bytes_rs + 4.0 * K * N
res = {"cells": N, "genes": D, "nnz": int(nnz), "hbm_peak_GBps": peak, "l2": "inputs larger than L2 (nnz stream %.1f GB)" % (8e-9 * nnz),
       "row_stats": {"ms": ms_rs, "GBps": bytes_rs / ms_rs / 1e6, "frac_of_hbm_peak": bytes_rs / ms_rs / 1e6 / peak, "cells_per_s": N / ms_rs * 1e3},
       "nystrom_no_delta": {"ms": ms_ny, "GBps": bytes_ny / ms_ny / 1e6, "frac_of_hbm_peak": bytes_ny / ms_ny / 1e6 / peak, "cells_per_s": N / ms_ny * 1e3},
       "nystrom_delta_1024_random_pb": {"ms": ms_nyd, "cells_per_s": N / ms_nyd * 1e3},
       "nystrom_delta_1024_sorted_pb": {"ms": ms_nys, "cells_per_s": N / ms_nys * 1e3}}
# CPU beside it: the oracle port (single thread, as the reference's visitor body) on the first cells
if "--no-cpu" not in sys.argv:
    import oracle as orc
    n_s = min(N, 4000)
    sub, _, _ = sim.sim_block(ctx, tabs, 0, n_s)
    ip, ix, v = sub.download()
    t0 = time.perf_counter(); orc.row_stats(ip, ix, v, D); t1 = time.perf_counter()
    orc.nystrom_project(ip, ix, v, D, basis.cpu().numpy(), None, None, 1e4); t2 = time.perf_counter()
    res["cpu_port"] = {"sample_cells": n_s, "cores": 1, "row_stats_cells_per_s": n_s / (t1 - t0), "nystrom_cells_per_s": n_s / (t2 - t1)}
print(json.dumps(res))
