"""Device-resident, cell-sharded driver of the hot path.

One process per GPU; every rank owns a contiguous range of cells whose start is a multiple of
BLOCK_CELLS.  All heavy data stays in torch CUDA tensors; the C ABI is called on torch's current
stream.  Collectives (torch.distributed / NCCL) appear only where the reference reduces over cells:

    after K1   all-gather of per-block (batch, dim) partial sums; all-reduce(min,max) of 2 floats
    in K3      broadcast of the first r cells' K-vectors; all-gather of Gram / column-sum partials
    in K4      all-reduce(max) of the 2^kk code-presence flags
    after K5   all-reduce(sum) of the gene x group sums and group sizes

Order-sensitive reductions are all-gathered block partials summed in global block order, so results
are bit-identical for any GPU count.  With world size 1 the same code runs without collectives.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np
import torch

from . import BLOCK_CELLS, Context, CscBlock, LegumeError, _ptr
from ._lib import TARGET_ALL, TARGET_MEAN_ONLY, lib
from .exchange import Exchange


def shard_range(ncols: int, rank: int, world: int):
    """contiguous, BLOCK_CELLS-aligned cell range of `rank`"""
    nblk = (ncols + BLOCK_CELLS - 1) // BLOCK_CELLS
    per = (nblk + world - 1) // world
    lo = min(rank * per * BLOCK_CELLS, ncols)
    hi = min((rank + 1) * per * BLOCK_CELLS, ncols)
    return lo, hi


class HotPath:
    def __init__(self, ctx: Context, group=None):
        self.ctx = ctx
        self.dev = torch.device(f"cuda:{ctx.device}")
        self.ex = Exchange(group)
        self.world, self.rank = self.ex.world, self.ex.rank
        ctx.use_torch_stream()

    # ---- helpers ---------------------------------------------------------------------------------
    def _nblk(self, n):
        return (n + BLOCK_CELLS - 1) // BLOCK_CELLS

    def _sum_partials(self, partials: torch.Tensor, nblk_local: int, M: int):
        """partials (nblk_local, M) f64 -> (M,) f64 summed over ALL ranks' blocks in global block order"""
        out = torch.empty(M, dtype=torch.float64, device=self.dev)
        allp = self.ex.gather_block_partials(partials, nblk_local).contiguous()
        self.ctx.check(lib.lg_block_partials_finalize(self.ctx.h, _ptr(allp), allp.shape[0], M, _ptr(out)))
        return out

    def _total(self, n_local: int) -> int:
        return self.ex.total(n_local, self.dev)

    # ---- stage 1 -----------------------------------------------------------------------------------
    def project_exact(self, block: CscBlock, basis_kd: torch.Tensor, batch: torch.Tensor | None, nbatch: int):
        """the exact-order mode (lg_project_exact) over cell shards: the batch means are serial f32 folds over the cells
        in ascending order, so the running folds are handed on shard after shard in rank order — the result is the
        single-GPU (and the reference's) fold for any GPU count"""
        ctx, ex, n = self.ctx, self.ex, block.ncols
        K = basis_kd.shape[1]
        proj = torch.empty((n, K), dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_project_raw_exact(ctx.h, block.h, _ptr(basis_kd), K, _ptr(proj)))
        fsum = fcnt = None
        if batch is not None and nbatch >= 1:
            fsum = torch.zeros((nbatch, K), dtype=torch.float32, device=self.dev)
            fcnt = torch.zeros(nbatch, dtype=torch.int64, device=self.dev)
            for r in range(self.world):
                if r == self.rank:
                    ctx.check(lib.lg_proj_batch_fold(ctx.h, _ptr(proj), K, n, _ptr(batch), nbatch, _ptr(fsum), _ptr(fcnt)))
                ex.broadcast_(fsum, r)
                ex.broadcast_(fcnt, r)
        mm = torch.empty(2, dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_proj_centre_scale_exact(ctx.h, _ptr(proj), K, n, _ptr(batch) if fsum is not None else None,
                                                 nbatch if fsum is not None else 0, _ptr(fsum), _ptr(fcnt), _ptr(mm)))
        mn, mx = ex.minmax(mm)
        if mx > 4.0 or mn < -4.0:
            ctx.check(lib.lg_proj_clamp_rescale(ctx.h, _ptr(proj), K, n))
        return proj

    def project(self, block: CscBlock, basis_kd: torch.Tensor, batch: torch.Tensor | None, nbatch: int, exact: bool = False):
        """random_projection.rs:341-415 on this rank's cells; basis_kd (D, K) f32, batch int32 (n_local,)"""
        if exact:
            return self.project_exact(block, basis_kd, batch, nbatch)
        ctx, n = self.ctx, block.ncols
        K = basis_kd.shape[1]
        proj = torch.empty((n, K), dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_project_raw(ctx.h, block.h, _ptr(basis_kd), K, _ptr(proj)))
        sums = None
        if batch is not None and nbatch >= 1:
            M = nbatch * (K + 1)
            nblk = self._nblk(n)
            part = torch.empty((max(nblk, 1), M), dtype=torch.float64, device=self.dev)
            ctx.check(lib.lg_proj_batch_partials(ctx.h, _ptr(proj), K, n, _ptr(batch), nbatch, _ptr(part)))
            sums = self._sum_partials(part, nblk, M)
        mm = torch.empty(2, dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_proj_centre_scale(ctx.h, _ptr(proj), K, n, _ptr(batch) if sums is not None else None,
                                           nbatch if sums is not None else 0, _ptr(sums), _ptr(mm)))
        mn, mx = self.ex.minmax(mm)
        if mx > 4.0 or mn < -4.0:  # global decision, random_projection.rs:401
            ctx.check(lib.lg_proj_clamp_rescale(ctx.h, _ptr(proj), K, n))
        return proj

    # ---- stage 2 -----------------------------------------------------------------------------------
    def binary_codes(self, proj: torch.Tensor, kk: int):
        """random_projection.rs:535-564 over all ranks' cells; returns int64 codes (n_local,)"""
        ctx = self.ctx
        n, K = proj.shape
        ntot = self._total(n)
        rank_all = min(K, ntot)
        r = kk + 5 if rank_all > kk else rank_all
        r = min(r, ntot)
        first = torch.zeros((r, K), dtype=torch.float32, device=self.dev)
        if self.rank == 0:
            if n < r:
                raise LegumeError(1, "rank 0 must hold at least kk+5 cells")
            first.copy_(proj[:r])
        self.ex.broadcast_(first, src=0)
        # K3 stays on the stream: the K x r QR and the kk x kk eigen-decomposition are one-warp kernels
        q = torch.empty((kk, K), dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_codes_basis(ctx.h, _ptr(first), K, r, kk, _ptr(q)))
        nblk = self._nblk(n)
        M = kk * (kk + 1) // 2
        b = torch.empty((n, kk), dtype=torch.float32, device=self.dev)
        part = torch.empty((max(nblk, 1), M), dtype=torch.float64, device=self.dev)
        ctx.check(lib.lg_codes_gram(ctx.h, _ptr(proj), K, n, _ptr(q), kk, _ptr(b), _ptr(part)))
        gram = self._sum_partials(part, nblk, M)
        u = torch.empty((kk, kk), dtype=torch.float32, device=self.dev)
        sig = torch.empty(kk, dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_codes_factor(ctx.h, _ptr(gram), _ptr(q), K, kk, _ptr(u), _ptr(sig)))
        v = torch.empty((n, kk), dtype=torch.float32, device=self.dev)
        part2 = torch.empty((max(nblk, 1), kk), dtype=torch.float64, device=self.dev)
        ctx.check(lib.lg_codes_vproj(ctx.h, _ptr(b), kk, n, _ptr(u), _ptr(sig), _ptr(v), _ptr(part2)))
        sums = self._sum_partials(part2, nblk, kk)
        mean = torch.empty(kk, dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_codes_means(ctx.h, _ptr(sums), kk, ntot, _ptr(mean)))
        codes = torch.empty(n, dtype=torch.int64, device=self.dev)
        ctx.check(lib.lg_codes_pack(ctx.h, _ptr(v), kk, n, _ptr(mean), _ptr(codes)))
        return codes

    # ---- stage 3 -----------------------------------------------------------------------------------
    def assign_groups(self, codes: torch.Tensor, kk: int, padded: bool = False):
        """groups.rs:13-37 over all ranks' codes; returns (int32 group ids (n_local,), num_groups)"""
        ctx, n = self.ctx, codes.shape[0]
        present = torch.empty(1 << kk, dtype=torch.int32, device=self.dev)
        ctx.check(lib.lg_code_presence(ctx.h, _ptr(codes), n, kk, _ptr(present)))
        self.ex.max_(present)
        present_h = present.cpu().numpy().astype(np.uint32)
        lut_h = np.empty(1 << kk, np.uint32)
        ng = C.c_uint32()
        ctx.check(lib.lg_group_lut(ctx.h, _ptr(present_h), kk, int(padded), _ptr(lut_h), C.byref(ng)))
        lut = torch.from_numpy(lut_h.astype(np.int64).astype(np.int32, casting="unsafe")).to(self.dev)
        group = torch.empty(n, dtype=torch.int32, device=self.dev)
        ctx.check(lib.lg_codes_to_groups(ctx.h, _ptr(codes), n, kk, _ptr(lut), _ptr(group)))
        return group, int(ng.value)

    # ---- stage 4 -----------------------------------------------------------------------------------
    def collapse_basic(self, block: CscBlock, group: torch.Tensor, S: int, mult: torch.Tensor | None = None):
        """stats.rs:110-134 + all-reduce of the sums; returns (sum_ds (S, D), size_s (S,))"""
        ctx = self.ctx
        sum_ds = torch.empty((S, block.nrows), dtype=torch.float32, device=self.dev)
        size_s = torch.empty(S, dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_collapse_basic(ctx.h, block.h, _ptr(group), _ptr(mult), S, _ptr(sum_ds), _ptr(size_s)))
        self.ex.sum_(sum_ds)
        self.ex.sum_(size_s)
        return sum_ds, size_s

    def collapse_batch(self, block: CscBlock, group, batch, S: int, B: int, mult=None):
        """stats.rs:136-164 + all-reduce; returns (sum_db (B, D), n_bs (S, B))"""
        ctx = self.ctx
        sum_db = torch.empty((B, block.nrows), dtype=torch.float32, device=self.dev)
        n_bs = torch.empty((S, B), dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_collapse_batch(ctx.h, block.h, _ptr(group), _ptr(batch), _ptr(mult), S, B, _ptr(sum_db),
                                        _ptr(n_bs)))
        self.ex.sum_(sum_db)
        self.ex.sum_(n_bs)
        return sum_db, n_bs

    # ---- either side of the path ------------------------------------------------
    def row_stats(self, block: CscBlock):
        """streaming_sparse_running_stats over every shard (sparse_streaming.rs:23-60): (npos, s1, s2) as f64 (D,)
        tensors + the global column count.  Count data gives exact whole numbers, so the all-reduce(sum) returns the
        same totals for any GPU count (merge = add, sparse_stat.rs:183-196)."""
        ctx = self.ctx
        out = torch.empty((3, block.nrows), dtype=torch.float64, device=self.dev)
        ctx.check(lib.lg_row_stats(ctx.h, block.h, _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        self.ex.sum_(out)
        return out[0], out[1], out[2], self._total(block.ncols)

    def nystrom_project(self, block: CscBlock, basis_dk: torch.Tensor, delta_dp=None, pb_of_cell=None, column_sum_norm=1e4):
        """nystrom_proj_visitor (senna/src/svd/fit.rs:433-466): cells are independent, every rank projects its own
        shard against the replicated basis / delta; no exchange"""
        ctx = self.ctx
        K = int(basis_dk.shape[0])
        P = 0 if delta_dp is None else int(delta_dp.shape[0])
        out = torch.empty((block.ncols, K), dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_nystrom_project(ctx.h, block.h, _ptr(basis_dk), K, _ptr(delta_dp),
                                         _ptr(pb_of_cell) if delta_dp is not None else None, P, float(column_sum_norm), _ptr(out)))
        return out

    # ---- stage 5 -----------------------------------------------------------------------------------
    def optimize_single(self, sum_ds: torch.Tensor, size_s: torch.Tensor, a0=1.0, b0=1.0, target=TARGET_ALL):
        """stats.rs:351-368 (replicated on every rank after the all-reduce)"""
        ctx = self.ctx
        S, D = sum_ds.shape
        new = lambda: torch.empty((S, D), dtype=torch.float32, device=self.dev)
        mean = new()
        sd = new() if target == TARGET_ALL else None
        ls = new() if target == TARGET_ALL else None
        lm = new() if target != TARGET_MEAN_ONLY else None
        ctx.check(lib.lg_optimize_single(ctx.h, _ptr(sum_ds), _ptr(size_s), D, S, a0, b0, target, _ptr(mean), _ptr(sd),
                                         _ptr(lm), _ptr(ls)))
        return dict(mean=mean, sd=sd, log_mean=lm, log_sd=ls)

    # ---- stage 6 -----------------------------------------------------------------------------------
    def knn_topk_sharded(self, ref_local: torch.Tensor, qry_local: torch.Tensor, k: int, exclude_global=None):
        """exact kNN with the REFERENCE cells sharded over ranks: queries are all-gathered, every
        rank searches its own reference shard (squared distances, local indices), the k-lists travel back to the
        rank that owns the query and are merged by (squared distance, lower global index).  Returns
        (idx int32 (nq_local, k) global reference indices, dist f32 (nq_local, k)), identical to one search over
        all reference cells.  exclude_global: int32 (nq_local,) global index to drop per query, or None."""
        ctx = self.ctx
        d = qry_local.shape[1]
        kq = k + (1 if exclude_global is not None else 0)
        allq, qcnt = self.ex.all_gather_rows(qry_local.contiguous())
        nr_cnt = self.ex.counts(ref_local.shape[0], self.dev)
        NQ = allq.shape[0]
        idx = torch.empty((NQ, kq), dtype=torch.int32, device=self.dev)
        sq = torch.empty((NQ, kq), dtype=torch.float32, device=self.dev)
        ctx.check(lib.lg_knn_topk_sq(ctx.h, _ptr(ref_local.contiguous()), ref_local.shape[0], _ptr(allq.contiguous()), NQ, d, kq,
                                     None, _ptr(idx), _ptr(sq)))
        sidx = self.ex.exchange_query_lists(idx, qcnt).contiguous()
        ssq = self.ex.exchange_query_lists(sq, qcnt).contiguous()
        off = torch.tensor(np.concatenate([[0], np.cumsum(nr_cnt)[:-1]]).astype(np.int64), device=self.dev)
        nq = qry_local.shape[0]
        out_idx = torch.empty((nq, k), dtype=torch.int32, device=self.dev)
        out_dist = torch.empty((nq, k), dtype=torch.float32, device=self.dev)
        # the merge reads lists of length kq and writes the k best after the exclusion
        if kq != k:
            full_i = torch.empty((nq, kq), dtype=torch.int32, device=self.dev)
            full_d = torch.empty((nq, kq), dtype=torch.float32, device=self.dev)
            ctx.check(lib.lg_knn_merge_topk(ctx.h, _ptr(sidx), _ptr(ssq), self.world, nq, kq, _ptr(off), _ptr(exclude_global),
                                            _ptr(full_i), _ptr(full_d)))
            out_idx.copy_(full_i[:, :k])
            out_dist.copy_(full_d[:, :k])
        else:
            ctx.check(lib.lg_knn_merge_topk(ctx.h, _ptr(sidx), _ptr(ssq), self.world, nq, k, _ptr(off), None, _ptr(out_idx),
                                            _ptr(out_dist)))
        return out_idx, out_dist

    # ---- stage 7, pb-sample arm ---------------------------------------------------------------------
    def pb_matched_stat(self, block: CscBlock, proj: torch.Tensor, group: torch.Tensor, S: int, batch: torch.Tensor, B: int,
                        knn: int, mult: torch.Tensor | None = None):
        """collapse_columns_multilevel_vec's batch arm (collapse_data/mod.rs:940-990) over cell shards:
        build_pb_samples + per_batch_sc_neighbors + collect_matched_stat_coarse.  Exchanges: all-reduce(max) of the
        (group, batch) presence flags; the centroid folds handed on in rank order (they are serial f32 folds over
        cells, so the result is the one a single GPU computes); all-reduce(sum) of the per-pb-sample gene sums;
        all-reduce(min) of the (squared distance, global cell) keys.  Returns a dict with the layout, the matches and
        the imputed / residual sums (S, D) — identical on every rank and for every GPU count."""
        ctx, ex, dev = self.ctx, self.ex, self.dev
        n, K = proj.shape
        cap = S * B
        present = torch.empty(cap, dtype=torch.int32, device=dev)
        ctx.check(lib.lg_pair_presence(ctx.h, _ptr(group), _ptr(batch), n, S, B, _ptr(present)))
        ex.max_(present)
        present_h = present.cpu().numpy().astype(np.uint32)
        id_h, pg_h, pb_h = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty(cap, np.uint32)
        npb_c = C.c_uint32()
        ctx.check(lib.lg_pb_ids(ctx.h, _ptr(present_h), S, B, _ptr(id_h), _ptr(pg_h), _ptr(pb_h), C.byref(npb_c)))
        npb = int(npb_c.value)
        if npb == 0:
            raise LegumeError(1, "no pb-samples built")
        as_i32 = lambda a: torch.from_numpy(a.view(np.int32).copy()).to(dev)
        ids, pb_group, pb_batch = as_i32(id_h), as_i32(pg_h[:npb]), as_i32(pb_h[:npb])
        c2p = torch.empty(n, dtype=torch.int32, device=dev)
        ctx.check(lib.lg_cells_to_pb(ctx.h, _ptr(group), _ptr(batch), n, S, B, _ptr(ids), _ptr(c2p)))
        # centroids: one serial fold per pb-sample, continued shard after shard
        csum = torch.zeros((npb, K), dtype=torch.float32, device=dev)
        ccnt = torch.zeros(npb, dtype=torch.float32, device=dev)
        for r in range(self.world):
            if r == self.rank:
                ctx.check(lib.lg_pb_centroid_fold(ctx.h, _ptr(proj), K, n, _ptr(c2p), npb, _ptr(mult), _ptr(csum), _ptr(ccnt)))
            ex.broadcast_(csum, r)
            ex.broadcast_(ccnt, r)
        cen = torch.empty((npb, K), dtype=torch.float32, device=dev)
        ctx.check(lib.lg_pb_centroid_finish(ctx.h, _ptr(csum), _ptr(ccnt), npb, K, _ptr(cen)))
        # per-pb-sample gene sums: the collapse with the pb-sample as the label (integer counts: exact in any order)
        gene_sums, _ = self.collapse_basic(block, c2p, npb, mult)
        # matches: min over every shard's cells, then the per-batch top-k
        T = B * knn
        mp = torch.empty((npb, T), dtype=torch.int32, device=dev)
        md = torch.empty((npb, T), dtype=torch.float32, device=dev)
        counts = ex.counts(n, dev)
        offset = int(sum(counts[:self.rank]))
        qc = max(256, min(npb, (1 << 27) // npb))
        keys = torch.empty((qc, npb), dtype=torch.int64, device=dev)
        for q0 in range(0, npb, qc):
            nq = min(qc, npb - q0)
            ctx.check(lib.lg_pb_min_keys(ctx.h, _ptr(proj), K, n, _ptr(c2p), offset, _ptr(cen), _ptr(pb_batch), npb, q0, nq, _ptr(keys)))
            if self.world > 1:  # keys are < 2^63 (non-negative f32 bits << 32 | index): signed order == unsigned order;
                # the all-ones "nothing here" key is -1 as int64, so shift it out of the way of the minimum
                view = keys[:nq]
                view[view < 0] = torch.iinfo(torch.int64).max
                ex.dist.all_reduce(view, op=ex.dist.ReduceOp.MIN, group=ex.pg)
                view[view == torch.iinfo(torch.int64).max] = -1
            ctx.check(lib.lg_pb_topk_keys(ctx.h, _ptr(keys), npb, q0, nq, B, _ptr(pb_batch), knn, _ptr(mp), _ptr(md)))
        imp = torch.empty((S, block.nrows), dtype=torch.float32, device=dev)
        res = torch.empty((S, block.nrows), dtype=torch.float32, device=dev)
        ctx.check(lib.lg_collect_matched_stat_coarse(ctx.h, _ptr(gene_sums), block.nrows, npb, _ptr(ccnt), _ptr(pb_group), S, _ptr(mp),
                                                     _ptr(md), T, _ptr(imp), _ptr(res)))
        return dict(num_pb=npb, cell_to_pb=c2p, pb_group=pb_group, pb_batch=pb_batch, pb_count=ccnt, centroids=cen,
                    gene_sums=gene_sums, matched_pb=mp, matched_dist=md, imputed_sum_ds=imp, residual_sum_ds=res)

    # ---- the same path inside the library (lg_comm.cu): what a Rust host calls ----------------------------------------
    def init_native_comm(self):
        """give the library's context its own NCCL communicator over the ranks of this process group: rank 0 draws the
        id (lg_comm_unique_id), torch.distributed carries the 128 bytes, every rank calls lg_comm_init"""
        if getattr(self, "_native_comm", False):
            return
        ident = torch.zeros(128, dtype=torch.uint8, device=self.dev)
        if self.world > 1:
            if self.rank == 0:
                buf = (C.c_ubyte * 128)()
                self.ctx.check(lib.lg_comm_unique_id(self.ctx.h, buf))
                ident.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
            self.ex.broadcast_(ident, 0)
        raw = bytes(ident.cpu().numpy().tobytes())
        self.ctx.check(lib.lg_comm_init(self.ctx.h, raw, self.rank, self.world))
        self._native_comm = True

    def run_native(self, block: CscBlock, basis_kd: torch.Tensor, batch, nbatch: int, kk: int, target=TARGET_ALL):
        """lg_hotpath_run_sharded: projection -> codes -> groups -> collapse (+ all-reduce) -> posterior in ONE library
        call, the exchanges on the library's own NCCL communicator; same results, bit for bit, as run()"""
        self.init_native_comm()
        ctx, n, D = self.ctx, block.ncols, block.nrows
        K = basis_kd.shape[1]
        cap = 1 << kk
        f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.dev)
        proj = f32(n, K)
        codes = torch.empty(n, dtype=torch.int64, device=self.dev)
        group = torch.empty(n, dtype=torch.int32, device=self.dev)
        sum_ds, size_s, mean = f32(cap, D), f32(cap), f32(cap, D)
        sd = f32(cap, D) if target == TARGET_ALL else None
        ls = f32(cap, D) if target == TARGET_ALL else None
        lm = f32(cap, D) if target != TARGET_MEAN_ONLY else None
        ng = C.c_uint32()
        ctx.check(lib.lg_hotpath_run_sharded(ctx.h, block.h, _ptr(basis_kd), K, _ptr(batch) if batch is not None and nbatch >= 1 else None,
                                             nbatch if batch is not None else 0, kk, target, _ptr(proj), _ptr(codes), _ptr(group),
                                             C.byref(ng), _ptr(sum_ds), _ptr(size_s), _ptr(mean), _ptr(sd), _ptr(lm), _ptr(ls)))
        S = int(ng.value)
        cut = lambda t: None if t is None else t[:S]
        return dict(proj=proj, codes=codes, group=group, num_groups=S, sum_ds=sum_ds[:S], size_s=size_s[:S],
                    posterior=dict(mean=cut(mean), sd=cut(sd), log_mean=cut(lm), log_sd=cut(ls)))

    # ---- whole path --------------------------------------------------------------------------------
    def run(self, block: CscBlock, basis_kd: torch.Tensor, batch, nbatch: int, kk: int, target=TARGET_ALL, exact: bool = False):
        """projection -> codes -> groups -> collapse -> posterior (single-batch arm of the path); exact=True takes the
        exact-order projection (bit-identical groups and sums from counts, several times slower)"""
        trace = os.environ.get("LG_TRACE")
        t = [time.perf_counter()]

        def mark():
            if trace:
                torch.cuda.synchronize()
                t.append(time.perf_counter())

        proj = self.project(block, basis_kd, batch, nbatch, exact)
        mark()
        codes = self.binary_codes(proj, kk)
        mark()
        group, ng = self.assign_groups(codes, kk)
        mark()
        sum_ds, size_s = self.collapse_basic(block, group, ng)
        mark()
        post = self.optimize_single(sum_ds, size_s, 1.0, 1.0, target)
        mark()
        if trace and self.rank == 0:
            names = ["project", "codes", "groups", "collapse", "posterior"]
            print("LG_TRACE " + " ".join(f"{n}={1e3 * (b - a):.2f}ms" for n, a, b in zip(names, t, t[1:])), flush=True)
        return dict(proj=proj, codes=codes, group=group, num_groups=ng, sum_ds=sum_ds, size_s=size_s, posterior=post)
