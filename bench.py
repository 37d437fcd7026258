#!/usr/bin/env python
"""bench.py — cells/sec of the legume-rs hot path (projection + codes + groups + collapse +
posterior) on N B200 GPUs of one node, with the roofline of the dominant kernel and the CPU
restatement of the reference timed beside it.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic `data-beans-sim topic`-shaped
counts.  N = 1 runs BASELINE.json configs[1] (1M cells x 30k genes, single batch, K = 50, 2^10 bins);
N > 1 shards cells (1.25M per GPU, configs[3] at N = 8) with NCCL all-reduce of the gene x group
sums (weak scaling).  Inputs are resident in HBM for `value`; `e2e` goes through the C ABI with host
buffers in the reference's own form (u64 indptr / u64 indices / f32 data) and reads results back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE JSON line.  NCCL prints its version banner on fd 1, so the real stdout is kept aside and
# fd 1 points at stderr while the run lasts; emit() writes the result line to the real stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "legume-rs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "cells/sec projection+collapse+kNN at 1/2/4/8 B200; nnz-stream GB/s vs HBM peak"
UNIT = "cells/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=0, help="cells per GPU (default: 1M at N=1, 1.25M at N>1)")
    ap.add_argument("--genes", type=int, default=30000)
    ap.add_argument("--depth", type=int, default=1500)
    ap.add_argument("--proj-dim", type=int, default=50)
    ap.add_argument("--sort-dim", type=int, default=10)
    ap.add_argument("--cpu-cells", type=int, default=0, help="cells in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-knn", action="store_true")
    return ap.parse_args()


def workload_name(args, world, cells_per_gpu):
    tot = cells_per_gpu * world
    return (f"data-beans-sim topic synthetic, {args.genes} genes x {tot} cells (~5% nnz), single batch, "
            f"proj d={args.proj_dim}, 2^{args.sort_dim} pseudobulk bins"
            + (f", cell-sharded over {world} GPUs" if world > 1 else ", 1 B200"))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """dense 16-bit tensor throughput (the kNN filter runs kind::f16): burst figure, the kernel is timed alone"""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json, cuBLAS bf16 burst)"
    except Exception:
        return 2250.0, "fallback (nominal dense bf16, B200_PROFILING.md)"


def k1_traffic(cells, genes, nnz):
    """dram__bytes_read.sum + dram__bytes_write.sum of the two K1 kernels from the committed ncu --set full capture
    (profiles/k1_traffic.json), if it was taken on this workload"""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            t = json.load(f)
        if int(t["cells"]) == int(cells) and int(t["genes"]) == int(genes) and abs(int(t["nnz"]) - int(nnz)) <= 0.001 * nnz:
            return float(t["dram_bytes_prep"]) + float(t["dram_bytes_umma"])
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle (a port of the reference's arithmetic) on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_pass(orc, ip, ix, v, D, basis, K, kk, threads):
    n = len(ip) - 1
    t0 = time.perf_counter()
    proj = orc.project(ip, ix, v, basis, np.zeros(n, np.uint32), 1, nthreads=threads)
    codes = orc.binary_codes(proj, kk)
    grp, ng = orc.assign_groups(codes)
    s, size = orc.collapse_basic(ip, ix, v, D, grp, ng)
    orc.optimize_single(s, size, 1.0, 1.0, orc.TARGET_ALL)
    return time.perf_counter() - t0


def cpu_sample(args, tabs, ncells, ctx=None):
    """the first `ncells` cells of the workload as host arrays: from the GPU generator when a ctx is
    at hand, else from the oracle's CPU twin (bit-identical by construction)"""
    import oracle as orc
    if ctx is not None:
        from legume_b200 import sim
        blk, _, _ = sim.sim_block(ctx, tabs, 0, ncells)
        out = blk.download()
        blk.free()
        return out
    topic, batch = tabs.cell_labels(0, ncells)
    return orc.sim_poisson_csc(tabs.seed, tabs.D, 0, ncells, topic, batch, tabs.ntopic, tabs.nbatch, tabs.lam, tabs.p0,
                               tabs.npiece)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Rust crates cannot be
    built in this image, so this times the oracle port (kind = "port") on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc
    from legume_b200 import sim
    world = args.gpus
    cells = args.cells or (1_000_000 if world == 1 else 1_250_000)
    cores = os.cpu_count() or 1
    ncpu = args.cpu_cells or 20_000
    tabs = sim.make_tables(args.genes, ntopic=8, nbatch=1, depth=args.depth, seed=42)
    ctx = None
    try:
        import torch
        if torch.cuda.is_available():
            import legume_b200 as lg
            ctx = lg.Context(0)
    except Exception:
        ctx = None
    ip, ix, v = cpu_sample(args, tabs, ncpu, ctx)
    basis = np.random.default_rng(0).standard_normal((args.genes, args.proj_dim)).astype(np.float32)
    for _ in range(max(args.warmup, 1)):
        cpu_pass(orc, ip, ix, v, args.genes, basis, args.proj_dim, args.sort_dim, cores)
    ts = [cpu_pass(orc, ip, ix, v, args.genes, basis, args.proj_dim, args.sort_dim, cores) for _ in range(args.steps)]
    t = float(np.mean(ts))
    val = ncpu / t
    sample = f"first {ncpu} cells ({len(v)} nnz) of the workload per step, OpenMP projection + serial collapse/codes"
    emit(({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, world, cells), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import legume_b200 as lg
    from legume_b200 import sim
    from legume_b200._lib import lib
    from legume_b200.pipeline import HotPath, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    cells_per_gpu = args.cells or (1_000_000 if world == 1 else 1_250_000)
    ntotal = cells_per_gpu * world
    D, K, kk = args.genes, args.proj_dim, args.sort_dim

    ctx = lg.Context(local)
    hp = HotPath(ctx)
    tabs = sim.make_tables(D, ntopic=8, nbatch=1, depth=args.depth, seed=42)
    lo, hi = shard_range(ntotal, rank, world)
    blk, _, _ = sim.sim_block(ctx, tabs, lo, hi)
    n_local = hi - lo
    basis_h = np.random.default_rng(0).standard_normal((D, K)).astype(np.float32)
    basis = torch.from_numpy(basis_h).to(dev)
    batch = torch.zeros(n_local, dtype=torch.int32, device=dev)
    nnz_local = blk.nnz

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return hp.run(blk, basis, batch, 1, kk, lg.TARGET_ALL)

    for _ in range(args.warmup):
        out = step()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    sync_all()
    launches = ctx.launch_count - l0
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = ntotal / (ms_per_step * 1e-3)
    ngroups = out["num_groups"]

    # ---- per-stage device times + roofline of the dominant kernel (live CUDA events, same stream) ----
    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    proj_raw = torch.empty((n_local, K), dtype=torch.float32, device=dev)
    reps = max(args.steps, 3)
    t_k1 = timed(lambda: ctx.check(lib.lg_project_raw(ctx.h, blk.h, basis.data_ptr(), K, proj_raw.data_ptr())), reps)
    stages = {"project_raw_ms": t_k1}
    if world == 1:
        proj = out["proj"]
        group = out["group"]
        stages["project_total_ms"] = timed(lambda: hp.project(blk, basis, batch, 1), reps)
        stages["binary_codes_ms"] = timed(lambda: hp.binary_codes(proj, kk), reps)
        stages["assign_groups_ms"] = timed(lambda: hp.assign_groups(out["codes"], kk), reps)
        stages["collapse_ms"] = timed(lambda: hp.collapse_basic(blk, group, ngroups), reps)
        stages["posterior_ms"] = timed(lambda: hp.optimize_single(out["sum_ds"], out["size_s"]), reps)
    peak, peak_src = measured_peak()
    k1_bytes = 8.0 * nnz_local + 8.0 * (n_local + 1) + 4.0 * K * n_local
    achieved = k1_bytes / (t_k1 * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "K1 projection = k_project_prep + k_project_umma (one nnz stream)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": k1_traffic(n_local, D, nnz_local),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes, "launch_ms": t_k1}
    if "collapse_ms" in stages:
        cb = 8.0 * nnz_local + 8.0 * (n_local + 1) + 4.0 * n_local + 4.0 * D * ngroups
        stages["collapse_GBps"] = cb / (stages["collapse_ms"] * 1e-3) / 1e9

    # ---- K7 at the shape of BASELINE configs[2]: one batch's 125k cells as references, 250k queries, d = 50, k = 10 ----
    roofline_knn = None
    if rank == 0 and not args.no_knn:
        nr_k, nq_k, k_k = 125_000, 250_000, 10
        g = torch.Generator(device=dev).manual_seed(0)
        std = lambda x: (x - x.mean(1, keepdim=True)) / x.std(1, keepdim=True, unbiased=False)
        kref, kqry = std(torch.randn((nr_k, K), device=dev, generator=g)), std(torch.randn((nq_k, K), device=dev, generator=g))
        kidx = torch.empty((nq_k, k_k), dtype=torch.int32, device=dev)
        kdist = torch.empty((nq_k, k_k), dtype=torch.float32, device=dev)
        t_knn = timed(lambda: ctx.check(lib.lg_knn_topk(ctx.h, kref.data_ptr(), nr_k, kqry.data_ptr(), nq_k, K, k_k, None, kidx.data_ptr(),
                                                        kdist.data_ptr())), reps)
        tpeak, tsrc = measured_tensor_peak()
        flop = 2.0 * K * nq_k * nr_k
        ach = flop / (t_knn * 1e-3) / 1e12
        stages["knn_250k_x_125k_ms"] = t_knn
        roofline_knn = {"bound": "tensor", "kernel": "K7 exact kNN = k_knn_umma (split-f16 tcgen05 filter) + k_knn_refine", "achieved": ach,
                        "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": None, "peak_source": tsrc,
                        "algorithmic_flops_per_launch": flop, "launch_ms": t_knn, "queries_per_s": nq_k / (t_knn * 1e-3),
                        # tensor-pipe occupancy: what the kernel actually issues on tcgen05, against the same measured peak
                        "issued": ach * 3.84, "issued_frac": ach * 3.84 / tpeak,
                        "note": "achieved / frac count ALGORITHMIC flops 2*d*Nq*Nr; the filter issues 3 f16 passes over a K axis "
                                "padded 50 -> 64, i.e. 3.84x these flops on the tensor pipe (issued / issued_frac; cf. "
                                "sm__pipe_tensor_cycles_active in profiles/)"}
        del kref, kqry, kidx, kdist

    # ---- e2e: host buffers in the reference's form through the C ABI, results read back ----
    e2e = None
    if not args.no_e2e:
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64 << 30
        # host copies (download + pinned) of this rank's block; every rank of the node draws on the same host memory
        need = nnz_local * 12 * 2.5 + n_local * 8
        budget = 0.4 * avail / max(world, 1)
        e2e_cells = n_local
        if need > budget:
            e2e_cells = max(1024, int(n_local * budget / need) // 1024 * 1024)
        if world > 1:  # same sample size on every rank
            t_cells = torch.tensor([e2e_cells], dtype=torch.int64, device=dev)
            dist.all_reduce(t_cells, op=dist.ReduceOp.MIN)
            e2e_cells = int(t_cells.item())
        sub, _, _ = (blk, None, None) if e2e_cells == n_local else sim.sim_block(ctx, tabs, lo, lo + e2e_cells)
        ip, ix, v = sub.download()
        if sub is not blk:
            sub.free()
        pin = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).pin_memory()
        h_ip, h_ix, h_v = pin(ip), pin(ix), pin(v)
        del ip, ix, v
        h_basis = torch.from_numpy(basis_h).pin_memory()
        h_proj = torch.empty((e2e_cells, K), dtype=torch.float32).pin_memory()
        h_group = torch.empty(e2e_cells, dtype=torch.int32).pin_memory()
        res_host = {}
        h2d = h_ip.numel() * 8 + h_ix.numel() * 8 + h_v.numel() * 4 + h_basis.numel() * 4
        d2h = [0]

        def e2e_step():
            h = lg.C.c_void_p()
            ctx.check(lib.lg_csc_upload(ctx.h, h_ip.data_ptr(), h_ix.data_ptr(), h_v.data_ptr(), D, 0, e2e_cells, None,
                                        lg.C.byref(h)))
            b = lg.CscBlock(ctx, h)
            o = hp.run(b, basis.copy_(h_basis, non_blocking=True), batch[:e2e_cells], 1, kk, lg.TARGET_ALL)
            h_proj.copy_(o["proj"], non_blocking=True)
            h_group.copy_(o["group"], non_blocking=True)
            nb = h_proj.numel() * 4 + h_group.numel() * 4
            for key in ("mean", "log_mean"):
                t = o["posterior"][key]
                if key not in res_host or res_host[key].shape != t.shape:
                    res_host[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                res_host[key].copy_(t, non_blocking=True)
                nb += t.numel() * 4
            torch.cuda.synchronize()
            d2h[0] = nb
            b.free()

        e2e_step()
        sync_all()
        esteps = min(args.steps, 3)
        wire0 = ctx.h2d_bytes
        t0 = time.perf_counter()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(esteps):
            e2e_step()
        b2.record()
        sync_all()
        wall = (time.perf_counter() - t0) / esteps
        ems = torch.tensor([max(a.elapsed_time(b2) / esteps, wall * 1e3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": e2e_cells * world / (float(ems.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h[0]), "cells_per_gpu": int(e2e_cells), "ms_per_step": float(ems.item()),
               "steps": esteps,
               # h2d_bytes_per_step is the size of the pinned host arrays handed to the C ABI (u64 indptr / u64 indices /
               # f32 values + basis); lg_csc_upload narrows the indices and packs count values on the host cores before
               # they travel, so fewer bytes cross the link:
               "h2d_wire_bytes_per_step": int((ctx.h2d_bytes - wire0) // esteps + h_basis.numel() * 4)}
        del h_ip, h_ix, h_v

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle as orc
        cores = os.cpu_count() or 1
        ncpu = args.cpu_cells or min(20_000, n_local)
        ip, ix, v = cpu_sample(args, tabs, ncpu, ctx)
        cpu_pass(orc, ip, ix, v, D, basis_h, K, kk, cores)
        reps_cpu, tsum = 0, 0.0
        while tsum < 10.0 and reps_cpu < 20:
            tsum += cpu_pass(orc, ip, ix, v, D, basis_h, K, kk, cores)
            reps_cpu += 1
        cpu = {"value": ncpu * reps_cpu / tsum, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {ncpu} cells ({len(v)} nnz) x {reps_cpu} passes; OpenMP projection on {cores} threads, "
                         "serial codes/collapse/posterior as in the reference's locked visitors"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world, cells_per_gpu), "cells_per_gpu": cells_per_gpu,
                       "genes": D, "nnz_per_gpu": int(nnz_local), "proj_dim": K, "sort_dim": kk, "groups": int(ngroups),
                       "l2": "inputs larger than L2 (nnz stream %.1f GB per GPU)" % (8e-9 * nnz_local),
                       "parallelism": f"cells sharded x{world}"},
            "roofline": roofline, "roofline_knn": roofline_knn, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "stages": stages,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
