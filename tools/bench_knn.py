"""kNN stage benchmark (BASELINE config 3 shape): queries x one batch's reference cells, d = 50, k = 10."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200")]
import numpy as np, torch
import legume_b200 as lg
from legume_b200._lib import lib

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 250_000
d, k = 50, 10
ctx = lg.Context(0); ctx.use_torch_stream()
g = torch.Generator(device="cuda").manual_seed(0)
ref = torch.randn((nr, d), device="cuda", generator=g)
ref = (ref - ref.mean(1, keepdim=True)) / ref.std(1, keepdim=True, unbiased=False)
qry = torch.randn((nq, d), device="cuda", generator=g)
qry = (qry - qry.mean(1, keepdim=True)) / qry.std(1, keepdim=True, unbiased=False)
idx = torch.empty((nq, k), dtype=torch.int32, device="cuda"); dist = torch.empty((nq, k), device="cuda")
def run():
    ctx.check(lib.lg_knn_topk(ctx.h, ref.data_ptr(), nr, qry.data_ptr(), nq, d, k, None, idx.data_ptr(), dist.data_ptr()))
for _ in range(2): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
a.record()
for _ in range(reps): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
flop = 2.0 * d * nq * nr
print(json.dumps({"nr": nr, "nq": nq, "d": d, "k": k, "ms": ms, "algorithmic_TFLOPs": flop / ms / 1e9,
                  "queries_per_s": nq / ms * 1e3}))
