"""K1 A/B: the fused projection kernel against the two-kernel form (LG_K1_FUSED=0) on the same block — bit comparison of
the raw projections and device times (LG_K1_TRACE=1 prints them on stderr).  Diagnostic, not a test.

    python tools/k1_check.py [D] [N] [depth] [reps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "legume-rs_b200"))

import numpy as np
import torch

import legume_b200 as lg
from legume_b200 import sim
from legume_b200._lib import lib
from legume_b200.pipeline import _ptr


def main():
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    K = 50
    ctx = lg.Context(0)
    tabs = sim.make_tables(D, ntopic=8, nbatch=1, depth=depth, seed=42)
    blk, _, _ = sim.sim_block(ctx, tabs, 0, N)
    basis = torch.from_numpy(np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)).cuda()
    outs = {}
    for mode in ("0", "1"):
        os.environ["LG_K1_FUSED"] = mode
        proj = torch.full((N, K), float("nan"), dtype=torch.float32, device="cuda")
        ms = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream())
            t0 = time.perf_counter()
            ctx.check(lib.lg_project_raw(ctx.h, blk.h, _ptr(basis), K, _ptr(proj)))
            torch.cuda.synchronize()
            ms.append((time.perf_counter() - t0) * 1e3)
        outs[mode] = proj.cpu().numpy()
        print(f"LG_K1_FUSED={mode}: wall ms {['%.3f' % x for x in ms]} nnz {blk.nnz}", flush=True)
    a, b = outs["0"], outs["1"]
    print("nan in fused:", int(np.isnan(b).sum()), " bit-identical:", bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))),
          " max abs diff:", float(np.nanmax(np.abs(a - b))), " rows differing:", int((a != b).any(axis=1).sum()))


if __name__ == "__main__":
    main()
