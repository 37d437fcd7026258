"""K1 A/B: lg_project_raw on the same block under two settings of one environment switch — bit comparison of the raw
projections and the per-kernel device times (LG_K1_TRACE=1 prints them on stderr).  Diagnostic, not a test.

    python tools/k1_ab.py ENVVAR A B [D] [N] [depth] [reps]      e.g.  LG_K1_SCAN 0 1 30000 262144
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
    var, va, vb = sys.argv[1], sys.argv[2], sys.argv[3]
    D = int(sys.argv[4]) if len(sys.argv) > 4 else 30000
    N = int(sys.argv[5]) if len(sys.argv) > 5 else 262144
    depth = int(sys.argv[6]) if len(sys.argv) > 6 else 1000
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
    K = 50
    os.environ["LG_K1_TRACE"] = "1"
    ctx = lg.Context(0)
    tabs = sim.make_tables(D, ntopic=8, nbatch=1, depth=depth, seed=42)
    blk, _, _ = sim.sim_block(ctx, tabs, 0, N)
    basis = torch.from_numpy(np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)).cuda()
    outs = {}
    for val in (va, vb):
        os.environ[var] = val
        proj = torch.full((N, K), float("nan"), dtype=torch.float32, device="cuda")
        ms = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.check(lib.lg_project_raw(ctx.h, blk.h, _ptr(basis), K, _ptr(proj)))
            torch.cuda.synchronize()
            ms.append((time.perf_counter() - t0) * 1e3)
        outs[val] = proj.cpu().numpy()
        print(f"{var}={val}: wall ms {['%.3f' % x for x in ms]} nnz {blk.nnz}", file=sys.stderr, flush=True)
    a, b = outs[va], outs[vb]
    print(f"nan in {var}={vb}:", int(np.isnan(b).sum()), " bit-identical:", bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))),
          " max abs diff:", float(np.nanmax(np.abs(a - b))), " rows differing:", int((a != b).any(axis=1).sum()), file=sys.stderr)


if __name__ == "__main__":
    main()
