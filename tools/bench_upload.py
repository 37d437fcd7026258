"""Where the end-to-end time goes: host CSC (u64/u64/f32, pinned) -> lg_csc_upload -> hot path -> results to host."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200")]
import numpy as np, torch
import legume_b200 as lg
from legume_b200 import sim
from legume_b200._lib import lib
from legume_b200.pipeline import HotPath
_nums = [a for a in sys.argv[1:] if a.isdigit()]
N = int(_nums[0]) if _nums else 1_000_000
D, K, kk = 30000, 50, 10
ctx = lg.Context(0); hp = HotPath(ctx)
tabs = sim.make_tables(D, ntopic=8, nbatch=1, depth=1500, seed=42)
blk, _, _ = sim.sim_block(ctx, tabs, 0, N)
ip, ix, v = blk.download(); blk.free()
pin = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).pin_memory()
h_ip, h_ix, h_v = pin(ip), pin(ix), pin(v); del ip, ix, v
basis = torch.from_numpy(np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)).cuda()
batch = torch.zeros(N, dtype=torch.int32, device="cuda")
def now():
    torch.cuda.synchronize(); return time.perf_counter()
res = {}
PAGEABLE = "--pageable" in sys.argv  # the reference's Vec<u64> / Vec<f32> slices are ordinary (pageable) host memory
if PAGEABLE:
    h_ip, h_ix, h_v = (torch.from_numpy(t.numpy().copy()) for t in (h_ip, h_ix, h_v))
modes = [("device_narrow", {"LG_UPLOAD_THREADS": "0"}), ("default", {}), ("no_gaps", {"LG_UPLOAD_NO_GAPS": "1"}),
         ("nowide", {"LG_UPLOAD_NO_WIDE": "1"}), ("host4", {"LG_UPLOAD_THREADS": "4"}),
         ("host8", {"LG_UPLOAD_THREADS": "8"}), ("host16", {"LG_UPLOAD_THREADS": "16"}), ("host32", {"LG_UPLOAD_THREADS": "32"}),
         ("host16_nowide", {"LG_UPLOAD_THREADS": "16", "LG_UPLOAD_NO_WIDE": "1"}),
         ("host16_nopack", {"LG_UPLOAD_THREADS": "16", "LG_UPLOAD_NO_PACK": "1"}),
         ("wide2", {"LG_UPLOAD_WIDE_DEPTH": "2"}), ("wide3", {"LG_UPLOAD_WIDE_DEPTH": "3"}), ("wide4", {"LG_UPLOAD_WIDE_DEPTH": "4"}),
         ("wide6", {"LG_UPLOAD_WIDE_DEPTH": "6"}), ("wide8", {"LG_UPLOAD_WIDE_DEPTH": "8"})]
if "--wide-only" in sys.argv:
    modes = [m for m in modes if m[0].startswith("wide")]
if "--quick" in sys.argv:
    modes = [m for m in modes if m[0] in ("default", "no_gaps", "nowide", "wide3", "wide4")]
os.environ["LG_UPLOAD_TRACE"] = "1"
for name, env in modes:
    for k in ("LG_UPLOAD_THREADS", "LG_UPLOAD_NO_WIDE", "LG_UPLOAD_NO_PACK", "LG_UPLOAD_NO_GAPS", "LG_UPLOAD_WIDE_DEPTH"):
        os.environ.pop(k, None)
    os.environ.update(env)
    rows = []
    for it in range(3):
        t0 = now()
        h = C.c_void_p()
        ctx.check(lib.lg_csc_upload(ctx.h, h_ip.data_ptr(), h_ix.data_ptr(), h_v.data_ptr(), D, 0, N, None, C.byref(h)))
        t1 = now()
        b = lg.CscBlock(ctx, h)
        o = hp.run(b, basis, batch, 1, kk)
        t2 = now()
        b.free()
        t3 = now()
        rows.append({"upload_ms": 1e3 * (t1 - t0), "run_ms": 1e3 * (t2 - t1), "free_ms": 1e3 * (t3 - t2)})
    res[name] = rows
    print(name, [round(r["upload_ms"], 1) for r in rows], file=sys.stderr, flush=True)
bytes_up = h_ip.numel() * 8 + h_ix.numel() * 8 + h_v.numel() * 4
print(json.dumps({"cells": N, "pageable_host_arrays": PAGEABLE, "host_bytes": bytes_up, "host_cores": os.cpu_count(), "modes": res}))
