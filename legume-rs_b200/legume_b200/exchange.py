"""The exchange steps of the cell-sharded hot path, as device-agnostic torch.distributed
calls: NCCL over NVLink on the GPUs, gloo on CPU tensors in the host-logic tests.

Every order-sensitive reduction is an all-gather of per-block partials that are then summed in GLOBAL block
order, so a result never depends on how many ranks the cells were cut into.  Only payloads that are exact in
any order (integer-valued count sums, presence flags, min / max) use all-reduce.
"""
from __future__ import annotations

import torch


class Exchange:
    def __init__(self, group=None):
        d = torch.distributed
        self.on = d.is_available() and d.is_initialized()
        self.dist, self.pg = (d if self.on else None), group
        self.world = d.get_world_size(group) if self.on else 1
        self.rank = d.get_rank(group) if self.on else 0

    # ---- scalars -------------------------------------------------------------------------------------
    def total(self, n_local: int, device) -> int:
        if self.world == 1:
            return int(n_local)
        t = torch.tensor([n_local], dtype=torch.int64, device=device)
        self.dist.all_reduce(t, group=self.pg)
        return int(t.item())

    def counts(self, n_local: int, device) -> list[int]:
        """n_local of every rank, in rank order"""
        if self.world == 1:
            return [int(n_local)]
        mine = torch.tensor([n_local], dtype=torch.int64, device=device)
        out = torch.empty(self.world, dtype=torch.int64, device=device)
        self.dist.all_gather_into_tensor(out, mine, group=self.pg)
        return [int(x) for x in out.tolist()]

    def minmax(self, mm: torch.Tensor):
        """mm = [min, max] of this rank -> global (min, max) as floats (random_projection.rs:401 is a global test)"""
        if self.world > 1:
            lo, hi = mm[0:1].clone(), mm[1:2].clone()
            self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN, group=self.pg)
            self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX, group=self.pg)
            return float(lo.item()), float(hi.item())
        a, b = mm.tolist()
        return float(a), float(b)

    # ---- tensors -------------------------------------------------------------------------------------
    def sum_(self, t: torch.Tensor):
        """all-reduce(sum) in place: only for payloads that are exact in any order (count sums below 2^24)"""
        if self.world > 1:
            self.dist.all_reduce(t, group=self.pg)
        return t

    def max_(self, t: torch.Tensor):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.pg)
        return t

    def broadcast_(self, t: torch.Tensor, src: int = 0):
        if self.world > 1:
            self.dist.broadcast(t, src=src, group=self.pg)
        return t

    def gather_block_partials(self, partials: torch.Tensor, nblk_local: int):
        """partials (>= nblk_local, M) -> (world * mx, M): every rank's blocks in rank order, each rank padded to
        the largest block count with +0.0 rows (exact no-ops in the ordered sum that follows)"""
        M = partials.shape[1]
        if self.world == 1:
            return partials[:nblk_local]
        mx = max(self.counts(nblk_local, partials.device))
        padded = torch.zeros((mx, M), dtype=partials.dtype, device=partials.device)
        padded[:nblk_local] = partials[:nblk_local]
        out = torch.empty((self.world * mx, M), dtype=partials.dtype, device=partials.device)
        self.dist.all_gather_into_tensor(out, padded, group=self.pg)
        return out

    def all_gather_rows(self, rows: torch.Tensor):
        """rows (n_local, ...) -> (concatenation over ranks in rank order, counts)"""
        cnt = self.counts(rows.shape[0], rows.device)
        if self.world == 1:
            return rows, cnt
        mx = max(cnt)
        padded = torch.zeros((mx,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        padded[:rows.shape[0]] = rows
        out = torch.empty((self.world * mx,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        self.dist.all_gather_into_tensor(out, padded, group=self.pg)
        return torch.cat([out[r * mx:r * mx + cnt[r]] for r in range(self.world)]), cnt

    def exchange_query_lists(self, lists: torch.Tensor, qcounts: list[int]):
        """kNN shard merge, exchange half: `lists` (sum(qcounts), k) holds THIS rank's answers (from its reference
        shard) for every rank's queries; returns (world, qcounts[rank], k): every shard's answers for this rank's
        own queries, in shard (= rank) order"""
        if self.world == 1:
            return lists[None]
        k = lists.shape[1]
        mine = qcounts[self.rank]
        out = torch.empty((self.world * mine, k), dtype=lists.dtype, device=lists.device)
        try:
            self.dist.all_to_all_single(out, lists.contiguous(), output_split_sizes=[mine] * self.world,
                                        input_split_sizes=list(qcounts), group=self.pg)
        except (RuntimeError, NotImplementedError):
            # backends without all-to-all: gather everything, keep this rank's slice of every shard
            mx = sum(qcounts)
            allv = torch.empty((self.world * mx, k), dtype=lists.dtype, device=lists.device)
            self.dist.all_gather_into_tensor(allv, lists.contiguous(), group=self.pg)
            q0 = sum(qcounts[:self.rank])
            out = torch.cat([allv[s * mx + q0:s * mx + q0 + mine] for s in range(self.world)])
        return out.view(self.world, mine, k)
