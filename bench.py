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
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs[2] / configs[4] / sharded-kNN extras")
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
# CPU leg: the oracle port of the reference (oracle/oracle_bench.cpp: the reference's arithmetic in the
# reference's execution structure) on a bounded sample of the workload, on all host cores.
#
# The path has two kinds of cost: stages proportional to the number of cells (projection over 100-cell blocks,
# codes, group-wise collapse) and ONE Poisson-Gamma fit of the D x S sums whatever N is.  Timing a 20 000-cell
# sample and dividing by 20 000 would charge that fixed fit 50x too often (round 1's bias), so each is timed on
# its own and the throughput AT THE GPU ARM'S CONFIG is   N / (N * t_cell + t_post).
# ------------------------------------------------------------------------------------------------
def load_sim_tables():
    """legume_b200/sim_tables.py by PATH: pure numpy, does not import the product package nor map its library"""
    import importlib.util
    sp = importlib.util.spec_from_file_location("lg_sim_tables", os.path.join(ROOT, "legume-rs_b200", "legume_b200", "sim_tables.py"))
    mod = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(mod)
    return mod


def cpu_sample(tabs, ncells):
    """the first `ncells` cells of the workload as host arrays, from the oracle's CPU twin of the generator
    (bit-identical to the GPU generator: tests/test_gpu_parity.py::test_sim_matches_cpu_twin)"""
    import oracle as orc
    topic, batch = tabs.cell_labels(0, ncells)
    return orc.sim_poisson_csc(tabs.seed, tabs.D, 0, ncells, topic, batch, tabs.ntopic, tabs.nbatch, tabs.lam, tabs.p0, tabs.npiece)


class CpuModel:
    """per-stage CPU times on the sample + the fixed posterior; .rate(N) extrapolates to N cells"""

    def __init__(self, orc, ip, ix, v, D, basis, K, kk, S_full, threads):
        self.orc, self.a, self.D, self.basis, self.K, self.kk, self.threads = orc, (ip, ix, v), D, basis, K, kk, threads
        self.n = len(ip) - 1
        self.S_full = S_full
        # the D x S_full sufficient statistics of the full workload stand-in: Poisson fill (values do not change the cost)
        rng = np.random.default_rng(1)
        self.sum_full = rng.poisson(0.05 * 1000, size=(S_full, D)).astype(np.float32)
        self.size_full = np.full(S_full, 1000.0, np.float32)
        self.t = {}

    def per_cell_pass(self):
        orc, (ip, ix, v) = self.orc, self.a
        t0 = time.perf_counter()
        proj = orc.bench_project_blocks(ip, ix, v, self.basis, 0, self.threads)
        proj = orc.project_finish(proj, np.zeros(self.n, np.uint32), 1)
        t1 = time.perf_counter()
        codes = orc.binary_codes(proj, self.kk)
        grp, ng = orc.assign_groups(codes)
        t2 = time.perf_counter()
        orc.bench_collapse_groups(ip, ix, v, self.D, grp, ng, True, self.threads)
        t3 = time.perf_counter()
        orc.bench_collapse_groups(ip, ix, v, self.D, grp, ng, False, self.threads)
        t4 = time.perf_counter()
        return {"project_s": t1 - t0, "codes_groups_s": t2 - t1, "collapse_locked_s": t3 - t2, "collapse_lockfree_s": t4 - t3}

    def posterior_pass(self):
        t0 = time.perf_counter()
        self.orc.bench_optimize_single_mt(self.sum_full, self.size_full, 1.0, 1.0, 0, self.threads)
        return time.perf_counter() - t0

    def step(self):
        """one bounded step: the per-cell stages over the sample + the fixed fit once; returns (stage dict, seconds)"""
        st = self.per_cell_pass()
        st["posterior_DxS_s"] = self.posterior_pass()
        return st

    @staticmethod
    def rate(st, n_sample, n_cells, lockfree=False):
        coll = st["collapse_lockfree_s"] if lockfree else st["collapse_locked_s"]
        t_cell = (st["project_s"] + st["codes_groups_s"] + coll) / n_sample
        return n_cells / (n_cells * t_cell + st["posterior_DxS_s"])


def cpu_baseline(args, n_cells, cores, steps, warmup, budget_s=25.0):
    """returns (value at n_cells, cpu_baseline dict, mean step seconds)"""
    import oracle as orc
    tabs = load_sim_tables().make_tables(args.genes, ntopic=8, nbatch=1, depth=args.depth, seed=42)
    ncpu = args.cpu_cells or 20_000
    ip, ix, v = cpu_sample(tabs, ncpu)
    basis = np.random.default_rng(0).standard_normal((args.genes, args.proj_dim)).astype(np.float32)
    model = CpuModel(orc, ip, ix, v, args.genes, basis, args.proj_dim, args.sort_dim, 1 << args.sort_dim, cores)
    for _ in range(max(warmup, 1)):
        model.step()
    acc, nst, t_begin = {}, 0, time.perf_counter()
    while nst < steps and (nst == 0 or time.perf_counter() - t_begin < budget_s):
        st = model.step()
        for k, x in st.items():
            acc[k] = acc.get(k, 0.0) + x
        nst += 1
    st = {k: x / nst for k, x in acc.items()}
    val = CpuModel.rate(st, ncpu, n_cells)
    sample = (f"first {ncpu} cells ({len(v)} nnz) of the workload x {nst} steps on {cores} threads: projection over "
              f"{orc.lib().orc_default_block_size(C_u64(args.genes))}-cell blocks with per-block CSC repack, serial codes (nalgebra is "
              f"single-threaded), group-wise collapse under the reference's global lock, plus ONE {args.genes} x {1 << args.sort_dim} "
              f"posterior fit over gene blocks; extrapolated to {n_cells} cells as N / (N * t_cell + t_post)")
    info = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "same_config": "extrapolated: per-cell stages timed on the sample, the fixed D x S fit timed once at full size",
            "stage_seconds": {k: round(x, 4) for k, x in st.items()},
            "value_lockfree_collapse": CpuModel.rate(st, ncpu, n_cells, lockfree=True),
            "value_on_sample_only": ncpu / (st["project_s"] + st["codes_groups_s"] + st["collapse_locked_s"] + st["posterior_DxS_s"])}
    step_s = st["project_s"] + st["codes_groups_s"] + st["collapse_locked_s"] + st["collapse_lockfree_s"] + st["posterior_DxS_s"]
    return val, info, step_s


def C_u64(x):
    import ctypes
    return ctypes.c_uint64(x)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Rust crates cannot be built in this image,
    so this times the oracle port (kind = "port") on all host cores.  Nothing of the product is imported or mapped."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    cells = args.cells or (1_000_000 if world == 1 else 1_250_000)
    cores = os.cpu_count() or 1
    val, info, step_s = cpu_baseline(args, cells * world, cores, args.steps, args.warmup, budget_s=150.0)
    emit(({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, world, cells), "sample": info["sample"]},
        "cpu_baseline": info,
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

    # the timed path is ONE C-ABI call per step (lg_hotpath_run_sharded: what a Rust host calls, exchanges on the
    # library's own NCCL communicator); LG_BENCH_STAGED=1 times the staged calls of legume_b200.pipeline instead
    staged = os.environ.get("LG_BENCH_STAGED") == "1"
    run_path = hp.run if staged else hp.run_native

    def step():
        return run_path(blk, basis, batch, 1, kk, lg.TARGET_ALL)

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
    # every rank runs the stages (those with an exchange are collective calls); rank 0's device times are reported
    proj = out["proj"]
    group = out["group"]
    stages["project_total_ms"] = timed(lambda: hp.project(blk, basis, batch, 1), reps)
    stages["binary_codes_ms"] = timed(lambda: hp.binary_codes(proj, kk), reps)
    stages["assign_groups_ms"] = timed(lambda: hp.assign_groups(out["codes"], kk), reps)
    stages["collapse_ms"] = timed(lambda: hp.collapse_basic(blk, group, ngroups), reps)
    stages["posterior_ms"] = timed(lambda: hp.optimize_single(out["sum_ds"], out["size_s"]), reps)
    if world > 1:
        # the exchange steps on their own: the D x S all-reduce after K5 and the small ordered gathers of K1 / K3
        scratch = torch.zeros_like(out["sum_ds"])
        stages["allreduce_sum_ds_ms"] = timed(lambda: hp.ex.sum_(scratch), reps)
        stages["allreduce_sum_ds_bytes"] = int(scratch.numel() * 4)
        nblk = (n_local + lg.BLOCK_CELLS - 1) // lg.BLOCK_CELLS
        part = torch.zeros((max(nblk, 1), K + 1), dtype=torch.float64, device=dev)
        stages["gather_block_partials_ms"] = timed(lambda: hp._sum_partials(part, nblk, K + 1), reps)
        local_only = lambda: ctx.check(lib.lg_collapse_basic(ctx.h, blk.h, group.data_ptr(), None, ngroups, scratch.data_ptr(),
                                                             out["size_s"].clone().data_ptr()))
        stages["collapse_kernel_only_ms"] = timed(local_only, reps)
        stages["collectives"] = "NCCL all-reduce(sum) of the D x S sums; all-gather of f64 block partials (K1 batch sums, K3 Gram, K3 means)"
    # the six stages INSIDE the one-call path (lg_ctx_time_stages: events around each stage of lg_hotpath_run_sharded, outside
    # the timed region): there K5 sums K1's 1-bit pattern instead of streaming the CSC arrays again, which the stand-alone
    # lg_collapse_basic above cannot do
    if not staged:
        import ctypes as _C
        lib.lg_ctx_time_stages(ctx.h, 1)
        rows = []
        pc0 = lib.lg_ctx_pattern_collapse_count(ctx.h)
        for _ in range(3):
            step()
            buf = (_C.c_float * 6)()
            if lib.lg_hotpath_last_stage_ms(ctx.h, buf) == 0:
                rows.append([float(x) for x in buf])
        lib.lg_ctx_time_stages(ctx.h, 0)
        if rows:
            med = np.median(np.asarray(rows), axis=0)
            names = ("project", "codes", "groups", "collapse", "allreduce", "posterior")
            stages["in_path_ms"] = {k: float(v) for k, v in zip(names, med)}
            stages["in_path_collapse_from_pattern"] = bool(lib.lg_ctx_pattern_collapse_count(ctx.h) - pc0 == 3)
    peak, peak_src = measured_peak()
    k1_bytes = 8.0 * nnz_local + 8.0 * (n_local + 1) + 4.0 * K * n_local
    achieved = k1_bytes / (t_k1 * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "K1 projection = k_project_prep + k_project_umma (one nnz stream)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": k1_traffic(n_local, D, nnz_local),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes, "launch_ms": t_k1}
    if "collapse_ms" in stages:
        cb = 8.0 * nnz_local + 8.0 * (n_local + 1) + 4.0 * n_local + 4.0 * D * ngroups
        stages["collapse_GBps"] = cb / (stages["collapse_ms"] * 1e-3) / 1e9
        if "in_path_ms" in stages:  # ALGORITHMIC bytes of the stage over its time in the path (the pattern form moves fewer)
            ip = stages["in_path_ms"]
            stages["in_path_collapse_GBps"] = cb / (ip["collapse"] * 1e-3) / 1e9
            stages["in_path_project_plus_collapse_frac"] = (k1_bytes + cb) / ((ip["project"] + ip["collapse"]) * 1e-3) / 1e9 / peak

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
                        # every one of the Nq x Nr f32 accumulators has to leave TMEM to be ranked: 64 B / clk / SM (tcgen05.ld),
                        # 148 SMs at the sampled SM clock — the bound of an exact brute-force search at this d
                        "tmem_read_bytes": 4.0 * nq_k * nr_k,
                        "tmem_read_peak_TBps": 148 * 64 * 1.965e9 / 1e12,
                        "tmem_read_achieved_TBps": 4.0 * nq_k * nr_k / (t_knn * 1e-3) / 1e12,
                        "tmem_read_frac": (4.0 * nq_k * nr_k / (t_knn * 1e-3)) / (148 * 64 * 1.965e9),
                        "tensor_frac_ceiling": (flop / (4.0 * nq_k * nr_k / (148 * 64 * 1.965e9))) / 1e12 / tpeak,
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
            e2e_cells = int(t_cells.item()) // 1024 * 1024  # shards are cut at multiples of the 1024-cell reduction block
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
            o = run_path(b, basis.copy_(h_basis, non_blocking=True), batch[:e2e_cells], 1, kk, lg.TARGET_ALL)
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
        _, cpu, _ = cpu_baseline(args, ntotal, os.cpu_count() or 1, steps=8, warmup=1, budget_s=20.0)

    # ---- end-to-end parity FROM COUNTS at configs[0] (outside every timed region; the oracle is the checker) ----
    parity = None
    if rank == 0 and world == 1 and not args.no_parity:
        from tools.e2e_parity import run_e2e_parity
        parity = run_e2e_parity(ctx, counts="gpu")
        parity["note"] = ("configs[0], GPU and CPU oracle each run counts + basis -> projection -> codes -> groups -> sums -> "
                          "posterior on their own. exact = lg_project_exact (the reference's operation order), fast = the "
                          "tensor-core path timed above (1e-5 projection contract).  K3 is held to the mirror oracle; against "
                          "the independent f32 SVD oracle the per-bit partitions agree on > 99.9 % of the cells (tests).")

    # ---- the other BASELINE configs, measured by the same command ----
    extras = None
    if not args.no_extras:
        from tools import bench_extras
        extras = {}
        try:
            extras["knn_sharded"] = bench_extras.knn_sharded(ctx, hp, 125_000, K, 10)   # every rank takes part
            if rank == 0 and world == 1:
                peak_hbm, _ = measured_peak()
                extras["c5"] = bench_extras.c5_posterior(ctx, D, peak=peak_hbm)
                extras["c3"] = bench_extras.c3_adjust(ctx, hp, 1_000_000, 8, D, K, kk, 10)
        except Exception as e:  # an extra must never cost the headline line
            extras["error"] = repr(e)[:300]

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world, cells_per_gpu), "cells_per_gpu": cells_per_gpu,
                       "genes": D, "nnz_per_gpu": int(nnz_local), "proj_dim": K, "sort_dim": kk, "groups": int(ngroups),
                       "l2": "inputs larger than L2 (nnz stream %.1f GB per GPU)" % (8e-9 * nnz_local),
                       "parallelism": f"cells sharded x{world}",
                       "api": "staged C-ABI calls (legume_b200.pipeline)" if staged else "lg_hotpath_run_sharded (one C-ABI call per step)"},
            "roofline": roofline, "roofline_knn": roofline_knn, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "stages": stages, "parity": parity, "extras": extras,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
