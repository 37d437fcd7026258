"""legume_b200 — host-side mirror of the legume-rs hot-path API over liblegume_b200.so.

The reference's host code is Rust (no toolchain in this image), so this module mirrors the
reference's operator interface for the path — same names, argument meaning and error behaviour —
in Python over the C ABI, so that the parity tests read like the reference's own tests:

    RandProjOps            data-beans-alg/src/random_projection.rs:43-162
    binary_sort_columns    data-beans-alg/src/random_projection.rs:535
    SparseIoVec.assign_groups / register_batch_membership / register_column_multiplicity
                           data-beans/src/sparse_io_vector/{groups.rs:13-37, batch.rs:259-336}
    CollapsingOps          data-beans-alg/src/collapse_data/mod.rs:315-361
    CollapsedStat/optimize data-beans-alg/src/collapse_data/stats.rs:378-582
    GammaMatrix            matrix-param/src/dmatrix_gamma.rs, traits.rs
    ColumnDict             matrix-util/src/knn/mod.rs:62-299

Matrix convention: a nalgebra `DMatrix` of shape R x C (column-major) is represented by an array of
shape (C, R) in C order, so `proj[j]` is cell j's K-vector and `sum_ds[s]` is group s's gene vector.
Arrays may be numpy (host) or torch CUDA tensors (device, used in place, no copies).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import (BLOCK_CELLS, EXPORTED, LIB_PATH, TARGET_ALL, TARGET_MEAN_AND_LOG_MEAN, TARGET_MEAN_ONLY,
                   LG_OK, LegumeError, lib)

DEFAULT_PROJECTION_SEED = 0x50524F4A_50524F4A  # random_projection.rs:41
DEFAULT_KNN = 10                                # collapse_data/mod.rs:27
NONE_U32 = 0xFFFFFFFF
DEFAULT_OPT_ITER = 100                          # collapse_data/mod.rs:28

__all__ = ["Context", "CscBlock", "SparseIoVec", "binary_sort_columns", "GammaMatrix", "CollapsedStat",
           "CollapsedOut", "optimize", "ColumnDict", "LegumeError", "CalibrateTarget", "compute_level_sort_dims",
           "pad_numeric_labels", "merge_stat", "MultilevelParams", "PbSampleLayout", "build_pb_sample_layout",
           "per_batch_sc_neighbors", "collect_matched_stat_coarse", "compute_fine_to_coarse_mapping",
           "sort_batch_proximity", "knn_match_batches", "SparseRunningStatistics", "nystrom_project", "SparseIoStack", "mix_seed",
           "SparseMtxData", "RefineParams", "refine_assignments", "build_reproject_offsets", "project_to_refinement",
           "child_offset_within_parent", "compute_sibling_sets", "intersect_with_siblings_fallback", "build_candidate_sets",
           "compact_labels"]
DEFAULT_NUM_LEVELS = 2                          # collapse_data/stats.rs:688


class CalibrateTarget:
    All = TARGET_ALL
    MeanOnly = TARGET_MEAN_ONLY
    MeanAndLogMean = TARGET_MEAN_AND_LOG_MEAN


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if _is_torch(x):
        assert x.is_contiguous(), "device tensors must be contiguous"
        return x.data_ptr()
    assert x.flags["C_CONTIGUOUS"], "host arrays must be C-contiguous"
    return x.ctypes.data


def _as(x, dtype):
    """host arrays are converted/copied to the ABI dtype; device tensors must already match"""
    if x is None:
        return None
    if _is_torch(x):
        import torch
        want = {np.float32: torch.float32, np.uint32: torch.int32, np.uint64: torch.int64, np.float64: torch.float64,
                np.uint8: torch.uint8}[dtype]
        ok = {x.dtype} & {want, getattr(torch, "uint32", want), getattr(torch, "uint64", want)}
        assert ok or x.element_size() == np.dtype(dtype).itemsize, f"device tensor dtype {x.dtype} != {dtype}"
        return x.contiguous()
    return np.ascontiguousarray(x, dtype)


class Context:
    """One per device (lg_ctx).  Fails loudly without a CUDA device: there is no CPU path."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib.lg_ctx_create(device, C.byref(h))
        if rc != 0:
            raise LegumeError(rc, f"lg_ctx_create(device={device}) failed: no usable CUDA device (no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise LegumeError(rc, lib.lg_last_error(self.h).decode())

    def sync(self):
        self.check(lib.lg_ctx_sync(self.h))

    def set_stream(self, stream_ptr):
        self.check(lib.lg_ctx_set_stream(self.h, stream_ptr))

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launch_count(self) -> int:
        return int(lib.lg_ctx_launch_count(self.h))

    @property
    def h2d_bytes(self) -> int:
        """bytes lg_csc_upload has put on the host->device link"""
        return int(lib.lg_ctx_h2d_bytes(self.h))

    def empty(self, shape, dtype, device: bool):
        """allocate an output next to the inputs (torch CUDA tensor or numpy array)"""
        if device:
            import torch
            tdt = {np.float32: torch.float32, np.uint32: torch.int32, np.uint64: torch.int64,
                   np.float64: torch.float64}[dtype]
            return torch.empty(shape, dtype=tdt, device=f"cuda:{self.device}")
        return np.empty(shape, dtype)

    def close(self):
        if self.h:
            lib.lg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CscBlock:
    """Device-resident CSC block (lg_csc): the data feed of the path (SparseIo::csc_column_arrays)."""

    def __init__(self, ctx: Context, handle, keepalive=None):
        self.ctx, self.h, self._keep = ctx, handle, keepalive
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.lg_csc_shape(self.h, C.byref(a), C.byref(b), C.byref(c))
        self.nrows, self.ncols, self.nnz = a.value, b.value, c.value

    @classmethod
    def upload(cls, ctx, indptr, indices, data, nrows, col_lo=0, col_hi=None, row_remap=None):
        indptr = np.ascontiguousarray(indptr, np.uint64)
        indices = np.ascontiguousarray(indices, np.uint64)
        data = np.ascontiguousarray(data, np.float32)
        if col_hi is None:
            col_hi = len(indptr) - 1
        h = C.c_void_p()
        if row_remap is None:
            ctx.check(lib.lg_csc_upload(ctx.h, _ptr(indptr), _ptr(indices), _ptr(data), nrows, col_lo, col_hi, None, C.byref(h)))
        else:
            # remapped rows are made canonical on the way (sorted, duplicates summed, absent rows dropped: read.rs:246-281)
            rr = np.ascontiguousarray(row_remap, np.uint32)
            ctx.check(lib.lg_csc_upload_remap(ctx.h, _ptr(indptr), _ptr(indices), _ptr(data), nrows, col_lo, col_hi, _ptr(rr),
                                              len(rr), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def concat(cls, ctx, blocks):
        """columns of several blocks side by side (lg_csc_concat)"""
        arr = (C.c_void_p * len(blocks))(*[b.h for b in blocks])
        h = C.c_void_p()
        ctx.check(lib.lg_csc_concat(ctx.h, arr, len(blocks), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def wrap_device(cls, ctx, d_indptr, d_indices, d_values, nrows):
        """torch CUDA tensors: int64 indptr (ncols+1), int32 indices, float32 values; kept alive by the block"""
        h = C.c_void_p()
        ncols, nnz = d_indptr.numel() - 1, d_values.numel()
        ctx.check(lib.lg_csc_wrap_device(ctx.h, _ptr(d_indptr), _ptr(d_indices), _ptr(d_values), nrows, ncols, nnz,
                                         C.byref(h)))
        return cls(ctx, h, keepalive=(d_indptr, d_indices, d_values))

    def download(self):
        indptr = np.empty(self.ncols + 1, np.uint64)
        indices = np.empty(self.nnz, np.uint64)
        data = np.empty(self.nnz, np.float32)
        self.ctx.check(lib.lg_csc_download(self.ctx.h, self.h, _ptr(indptr), _ptr(indices), _ptr(data)))
        return indptr, indices, data

    def keep_pattern(self, on: bool = True):
        """lg_csc_keep_pattern: every projection of this block leaves its 1-bit sparsity pattern + the list of counts != 1 in
        buffers owned by the block, and the collapses that follow (unit multiplicities) sum those instead of streaming the
        arrays again — one projection, then one collapse per level of the multilevel scheme.  Same sums, bit for bit."""
        self.ctx.check(lib.lg_csc_keep_pattern(self.ctx.h, self.h, 1 if on else 0))
        return self

    def free(self):
        if self.h:
            lib.lg_csc_free(self.ctx.h if self.ctx.h else None, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------
# label helpers (host side, exactly the reference's string-rank rules)
# --------------------------------------------------------------------------------------------------
class SparseMtxData:
    """The reference's zarr backend as a reader (data-beans/src/sparse_backend/zarr.rs: `SparseMtxData::open`, the
    `SparseIo` trait methods the path uses).  The store is the Zarr V3 directory the reference writes; chunks are inflated
    on the host cores inside the library (lg_zarr_*, csrc/lg_ingest.cu) and blocks reach the device through
    lg_csc_upload.  Read-only: writing stores is the reference's `data-beans` CLI, outside the path."""

    def __init__(self, handle, file_name):
        self.h, self.file_name = handle, file_name
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.lg_zarr_shape(self.h, C.byref(a), C.byref(b), C.byref(c))
        self._shape = (a.value, b.value, c.value)
        self._preloaded = None

    @classmethod
    def open(cls, zarr_file):
        """zarr.rs:640-690 (`open`): shape from the root attributes nrow / ncol / nnz (:515-523)"""
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib.lg_zarr_open(os.fsencode(zarr_file), C.byref(h), err, len(err))
        if rc != LG_OK:
            raise LegumeError(rc, err.value.decode(errors="replace"))
        return cls(h, str(zarr_file))

    def _check(self, rc):
        if rc != LG_OK:
            raise LegumeError(rc, lib.lg_zarr_last_error(self.h).decode(errors="replace"))

    def num_rows(self):
        return self._shape[0]

    def num_columns(self):
        return self._shape[1]

    def num_non_zeros(self):
        return self._shape[2]

    def read_columns_host(self, col_lo=0, col_hi=None):
        """(indptr rebased to 0, indices, data) of columns [col_lo, col_hi) as the reference's u64 / u64 / f32 arrays"""
        col_hi = self.num_columns() if col_hi is None else col_hi
        first, last = C.c_uint64(), C.c_uint64()
        self._check(lib.lg_zarr_column_extent(self.h, col_lo, col_hi, C.byref(first), C.byref(last)))
        n = last.value - first.value
        ip, ix, v = np.empty(col_hi - col_lo + 1, np.uint64), np.empty(n, np.uint64), np.empty(n, np.float32)
        self._check(lib.lg_zarr_read_columns_host(self.h, col_lo, col_hi, _ptr(ip), _ptr(ix), _ptr(v)))
        return ip, ix, v

    def preload_columns(self):
        """zarr.rs:573-587"""
        self._preloaded = self.read_columns_host()

    def clean_preloaded_columns(self):
        self._preloaded = None

    def csc_column_arrays(self):
        """zarr.rs:982-994: None until preload_columns"""
        return self._preloaded

    def read_columns_csc(self, ctx: Context, col_lo=0, col_hi=None):
        """the column range as a device-resident block (SparseIoVec::read_columns_csc, read.rs:172-285)"""
        col_hi = self.num_columns() if col_hi is None else col_hi
        h = C.c_void_p()
        ctx.check(lib.lg_zarr_read_columns(ctx.h, self.h, col_lo, col_hi, C.byref(h)))
        return CscBlock(ctx, h)

    def close(self):
        if self.h:
            lib.lg_zarr_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _rank_labels(labels):
    """index = rank of label.to_string() in byte-wise order (batch.rs:274-275, groups.rs:20-24)"""
    strs = [str(int(x)) if isinstance(x, (np.integer, int)) else str(x) for x in labels]
    keys = sorted(set(strs), key=lambda s: s.encode())
    lut = {k: i for i, k in enumerate(keys)}
    return np.fromiter((lut[s] for s in strs), np.uint32, len(strs)), keys


def pad_numeric_labels(cell_to_group, k):
    """collapse_data/refine.rs:21-35"""
    width, n = 1, max(k, 1) - 1
    while n >= 10:
        width += 1
        n //= 10
    return [f"{int(g):0{width}d}" for g in cell_to_group]


PARTITION_SHUFFLE_SEED = 0x5041525453485546  # "PARTSHUF", matrix-util/src/utils.rs:12


class _Xoshiro256pp:
    """xoshiro256++ seeded through SplitMix64 (the construction behind rand's SmallRng::seed_from_u64 on 64-bit targets)"""
    M = 0xFFFFFFFFFFFFFFFF

    def __init__(self, seed):
        s, st = seed & self.M, []
        for _ in range(4):
            s = (s + 0x9E3779B97F4A7C15) & self.M
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.M
            st.append(z ^ (z >> 31))
        self.s = st

    def next(self):
        s, M = self.s, self.M
        rot = lambda x, k: ((x << k) | (x >> (64 - k))) & M
        out = (rot((s[0] + s[3]) & M, 23) + s[0]) & M
        t = (s[1] << 17) & M
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = rot(s[3], 45)
        return out

    def below(self, n):
        """uniform integer in [0, n): widening multiply with rejection (unbiased)"""
        lim = (1 << 64) - ((1 << 64) % n)
        while True:
            x = self.next()
            if x < lim:
                return x % n


def downsample_groups(col_to_group, num_groups, ntarget):
    """partition_by_membership with nelem_per_group (matrix-util/src/utils.rs:36-66): every group with more than `ntarget`
    columns keeps `ntarget` of them — a Fisher-Yates shuffle of its members (ascending) by a generator seeded with
    mix_seed(PARTITION_SHUFFLE_SEED, smallest member), truncated; the others get the label 0xFFFFFFFF"""
    g = np.asarray(col_to_group, np.uint32).copy()
    if ntarget < 0:
        raise LegumeError(1, "ncolumns_per_group must not be negative")
    order = np.argsort(g, kind="stable")
    bounds = np.searchsorted(g[order], np.arange(num_groups + 1))
    for k in range(num_groups):
        cells = order[bounds[k]:bounds[k + 1]]
        if len(cells) <= ntarget:
            continue
        rng = _Xoshiro256pp(mix_seed(PARTITION_SHUFFLE_SEED, int(cells[0])))
        perm = cells.copy()
        for i in range(len(perm) - 1, 0, -1):
            j = rng.below(i + 1)
            perm[i], perm[j] = perm[j], perm[i]
        g[perm[ntarget:]] = NONE_U32
    return g


def compute_level_sort_dims(finest_sort_dim, num_levels):
    """collapse_data/refine.rs:718-734 (f32 arithmetic, round half away from zero)"""
    if num_levels <= 1:
        return [finest_sort_dim]
    coarsest = min(7, finest_sort_dim)
    dims = []
    for level in range(num_levels):
        t = np.float32(level) / np.float32(num_levels - 1)
        dim = np.float32(finest_sort_dim) - t * np.float32(finest_sort_dim - coarsest)
        d = int(np.floor(abs(float(dim)) + 0.5) * (1 if dim >= 0 else -1))
        if not dims or dims[-1] != d:
            dims.append(d)
    return dims


# --------------------------------------------------------------------------------------------------
# stage 2/3 free functions
# --------------------------------------------------------------------------------------------------
def binary_sort_columns(ctx: Context, proj_kn, kk: int):
    """random_projection.rs:535-564: proj_kn (N, K) -> codes uint64[N] (< 2^kk)"""
    proj_kn = _as(proj_kn, np.float32)
    n, K = proj_kn.shape
    dev = _is_torch(proj_kn)
    codes = ctx.empty((n,), np.uint64, dev)
    ctx.check(lib.lg_binary_codes(ctx.h, _ptr(proj_kn), K, n, kk, _ptr(codes)))
    return codes


def assign_groups_from_codes(ctx: Context, codes, kk: int, padded: bool = False):
    """groups.rs:13-37 on integer codes: returns (group_of_cell uint32[N], num_groups)"""
    codes = _as(codes, np.uint64)
    dev = _is_torch(codes)
    n = codes.shape[0]
    group = ctx.empty((n,), np.uint32, dev)
    ng = C.c_uint32()
    ctx.check(lib.lg_assign_groups(ctx.h, _ptr(codes), n, kk, int(padded), _ptr(group), C.byref(ng)))
    return group, ng.value


def merge_stat(ctx: Context, fine_ds, fine_to_coarse, ncoarse):
    """stats.rs:790-833 for one D x S plane"""
    fine_ds = _as(fine_ds, np.float32)
    nfine, D = fine_ds.shape
    f2c = _as(fine_to_coarse, np.uint32)
    out = ctx.empty((ncoarse, D), np.float32, _is_torch(fine_ds))
    ctx.check(lib.lg_merge_stat(ctx.h, _ptr(fine_ds), D, nfine, _ptr(f2c), ncoarse, _ptr(out)))
    return out


# --------------------------------------------------------------------------------------------------
# stage 7: cross-batch neighbourhood adjustment (free functions; SparseIoVec methods wrap them)
# --------------------------------------------------------------------------------------------------
def sort_batch_proximity(ctx: Context, proj_kn, col_to_batch, nbatch):
    """batch.rs:182-234: (order (B, B) uint32 — row b = every batch by centroid distance, b first; centroids (B, K))"""
    proj_kn = _as(proj_kn, np.float32)
    n, K = proj_kn.shape
    batch = _as(col_to_batch, np.uint32)
    order = np.empty((nbatch, nbatch), np.uint32)
    cen = np.empty((nbatch, K), np.float32)
    ctx.check(lib.lg_batch_proximity(ctx.h, _ptr(proj_kn), K, n, _ptr(batch), nbatch, _ptr(order), _ptr(cen)))
    return order, cen


def knn_match_batches(ctx: Context, proj_kn, col_to_batch, nbatch, knn, target_order=None):
    """matched.rs:173-260 (kNN part): (matched_idx (N, nt*knn) global cell indices, matched_dist)"""
    proj_kn = _as(proj_kn, np.float32)
    n, K = proj_kn.shape
    batch = _as(col_to_batch, np.uint32)
    to = None if target_order is None else np.ascontiguousarray(target_order, np.uint32)
    nt = nbatch if to is None else to.shape[1]
    dev = _is_torch(proj_kn)
    idx = ctx.empty((n, nt * knn), np.uint32, dev)
    dist = ctx.empty((n, nt * knn), np.float32, dev)
    ctx.check(lib.lg_knn_match_batches(ctx.h, _ptr(proj_kn), K, n, _ptr(batch), nbatch, knn, _ptr(to), nt, _ptr(idx),
                                       _ptr(dist)))
    return idx, dist


class PbSampleLayout(dict):
    """pb_samples.rs:33-60: centroids (npb, K), cell_counts, pb_sample_to_batch, pb_sample_to_group, cell_to_pbsamp"""

    def __getattr__(self, k):
        return self[k]


def build_pb_sample_layout(ctx: Context, col_to_group, num_groups, col_to_batch, nbatch, proj_kn, col_weight=None,
                           anchor_batches=(), bulk_batches=()):
    """pb_samples.rs:94-219 (ordinary (batch, group) blocks; anchored / bulk singletons are outside the hot path)"""
    if len(anchor_batches) or len(bulk_batches):
        raise LegumeError(1, "anchor / bulk batches are outside the hot path (pb_samples.rs:75-88)")
    proj_kn = _as(proj_kn, np.float32)
    n, K = proj_kn.shape
    grp, bat = _as(col_to_group, np.uint32), _as(col_to_batch, np.uint32)
    w = None if col_weight is None else _as(col_weight, np.float32)
    cap = num_groups * nbatch
    c2p = np.empty(n, np.uint32)
    pg, pb = np.empty(cap, np.uint32), np.empty(cap, np.uint32)
    cnt = np.empty(cap, np.float32)
    cen = np.empty((cap, K), np.float32)
    npb = C.c_uint32()
    ctx.check(lib.lg_pb_layout(ctx.h, _ptr(proj_kn), K, n, _ptr(grp), num_groups, _ptr(bat), nbatch, _ptr(w), _ptr(c2p),
                               _ptr(pg), _ptr(pb), _ptr(cnt), _ptr(cen), C.byref(npb)))
    k = npb.value
    return PbSampleLayout(centroids=cen[:k].copy(), cell_counts=cnt[:k].copy(), pb_sample_to_batch=pb[:k].copy(),
                          pb_sample_to_group=pg[:k].copy(), cell_to_pbsamp=c2p, num_pb=k)


def per_batch_sc_neighbors(ctx: Context, layout: PbSampleLayout, proj_kn, col_to_batch, nbatch, knn, anchor_batches=None):
    """pb_samples.rs:442-459 with pooled matching: (matched_pb (npb, B*knn) uint32, matched_dist)"""
    if anchor_batches is not None:
        raise LegumeError(1, "anchored matching is outside the hot path (pb_samples.rs:407-424)")
    proj_kn = _as(proj_kn, np.float32)
    n, K = proj_kn.shape
    npb = layout.num_pb
    mp = np.empty((npb, nbatch * knn), np.uint32)
    md = np.empty((npb, nbatch * knn), np.float32)
    ctx.check(lib.lg_pb_match(ctx.h, _ptr(proj_kn), K, n, _ptr(_as(col_to_batch, np.uint32)), nbatch,
                              _ptr(_as(layout.cell_to_pbsamp, np.uint32)), _ptr(_as(layout.centroids, np.float32)),
                              _ptr(_as(layout.pb_sample_to_batch, np.uint32)), npb, knn, _ptr(mp), _ptr(md)))
    return mp, md


def collect_matched_stat_coarse(ctx: Context, layout: PbSampleLayout, gene_sums, pbsamp_to_group, matched, stat):
    """stats.rs:698-784: fills stat.imputed_sum_ds / stat.residual_sum_ds; gene_sums (npb, D) dense"""
    mp, md = matched
    gs = _as(gene_sums, np.float32)
    npb, D = gs.shape
    S = stat.num_samples()
    dev = _is_torch(gs)
    imp, res = ctx.empty((S, D), np.float32, dev), ctx.empty((S, D), np.float32, dev)
    ctx.check(lib.lg_collect_matched_stat_coarse(ctx.h, _ptr(gs), D, npb, _ptr(_as(layout.cell_counts, np.float32)),
                                                 _ptr(_as(pbsamp_to_group, np.uint32)), S, _ptr(_as(mp, np.uint32)),
                                                 _ptr(_as(md, np.float32)), mp.shape[1], _ptr(imp), _ptr(res)))
    stat.imputed_sum_ds, stat.residual_sum_ds = imp, res


def compute_fine_to_coarse_mapping(ctx: Context, fine_codes, col_to_group, num_fine, coarse_dim):
    """refine.rs:741-769: (fine_to_coarse uint32[num_fine], num_coarse)"""
    codes, grp = _as(fine_codes, np.uint64), _as(col_to_group, np.uint32)
    f2c = np.empty(num_fine, np.uint32)
    k = C.c_uint32()
    ctx.check(lib.lg_fine_to_coarse(ctx.h, _ptr(codes), _ptr(grp), codes.shape[0], num_fine, coarse_dim, _ptr(f2c),
                                    C.byref(k)))
    return f2c, k.value


class RefineParams:
    """dc_poisson.rs:71-117 `RefineParams::default()`: the BBKNN + Poisson refinement of the pb-sample partition.  With one batch
    the refinement is the identity (refine.rs:126-147); with two or more it runs the Jacobi sweeps of dc_poisson.rs:778-915
    through lg_dcp_refine_level.  profile_source = "Projected" scores on the pb-samples' summed projection columns instead of
    their gene sums (Profiles::from_projection, dc_poisson.rs:164-195; feature weighting is then skipped, refine_multilevel.rs:
    228-236).  The sequential Gauss-Seidel sweeps (parallel = false) are outside the hot path and refused."""

    def __init__(self, num_gibbs=20, num_greedy=10, feature_weighting="FisherInfoNb", seed=42, gibbs_stagnation=0.005,
                 profile_source="Raw", parallel=True):
        if feature_weighting not in ("None", "FisherInfoNb"):
            raise LegumeError(1, "RefineParams.feature_weighting is None or FisherInfoNb (dc_poisson.rs:54-66)")
        if profile_source not in ("Raw", "Projected"):
            raise LegumeError(1, "RefineParams.profile_source is Raw or Projected (dc_poisson.rs:39-49)")
        if not parallel:
            raise LegumeError(1, "RefineParams.parallel = false (Gauss-Seidel sweeps, dc_poisson.rs:688-731) is outside the hot path")
        self.num_gibbs, self.num_greedy = int(num_gibbs), int(num_greedy)
        self.feature_weighting, self.seed, self.gibbs_stagnation = feature_weighting, int(seed), float(gibbs_stagnation)
        self.profile_source, self.parallel = profile_source, parallel


class MultilevelParams:
    """collapse_data/mod.rs:64-130.  `MultilevelParams::new` sets refine = Some(default) and observe_panels = true; pass
    refine=None for the legacy un-refined descent (mod.rs:943-1046)."""

    def __init__(self, proj_dim, knn_pb_samples=DEFAULT_KNN, num_levels=DEFAULT_NUM_LEVELS, sort_dim=None,
                 num_opt_iter=DEFAULT_OPT_ITER, refine="default", output_calibration=TARGET_ALL, observe_panels=True,
                 anchor_batches=None, bulk_batches=None):
        self.knn_pb_samples = knn_pb_samples
        self.num_levels = num_levels
        self.sort_dim = min(proj_dim, 12) if sort_dim is None else sort_dim  # mod.rs:119
        self.num_opt_iter = num_opt_iter
        self.refine = RefineParams() if isinstance(refine, str) and refine == "default" else refine
        self.output_calibration = output_calibration
        self.observe_panels = observe_panels
        self.anchor_batches, self.bulk_batches = anchor_batches, bulk_batches


# ---- the refinement arm's bookkeeping (dc_poisson.rs:493-509, refine.rs:43-88, collapse_data/mod.rs:823-841) ----
def compact_labels(labels):
    """labels -> 0..k in order of first appearance: (compact uint32, k)"""
    labels = np.asarray(labels)
    if labels.size == 0:
        return np.zeros(0, np.uint32), 0
    uniq, first, inv = np.unique(labels, return_index=True, return_inverse=True)
    rank = np.empty(len(uniq), np.uint32)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq), dtype=np.uint32)
    return rank[inv.reshape(-1)], len(uniq)


def initial_per_level_from_hash(fine_codes, first_cell_of_pb, level_dims):
    """refine.rs:68-88"""
    codes = np.asarray(fine_codes, np.uint64)[first_cell_of_pb]
    out = []
    for d in level_dims:
        mask = np.uint64(0xFFFFFFFFFFFFFFFF) if d >= 64 else np.uint64((1 << d) - 1)
        out.append(compact_labels(codes & mask)[0])
    return out


def fine_to_coarse_from_refined(pbsamp_to_fine, pbsamp_to_coarse, num_fine):
    """refine.rs:43-62: the coarse label of the first pb-sample of every fine group"""
    p2f = np.asarray(pbsamp_to_fine)
    _, first = np.unique(p2f, return_index=True)
    m = np.full(num_fine, NONE_U32, np.uint32)
    m[p2f[first]] = np.asarray(pbsamp_to_coarse, np.uint32)[first]
    return m


def build_reproject_offsets(fine_codes, first_cell_of_pb, level_dims):
    """refine.rs:95-125: per level the finest hash bits above the parent level's sort dim; empty for the coarsest level"""
    raw = np.asarray(fine_codes, np.uint64)[first_cell_of_pb]
    out = []
    for level in range(len(level_dims)):
        if level + 1 < len(level_dims):
            parent_dim = level_dims[level + 1]
            nbits = max(level_dims[level] - parent_dim, 0)
            mask = np.uint64(0xFFFFFFFFFFFFFFFF) if nbits >= 64 else np.uint64((1 << nbits) - 1)
            out.append(((raw >> np.uint64(parent_dim)) & mask).astype(np.uint32))
        else:
            out.append(np.zeros(0, np.uint32))
    return out


def project_to_refinement(child, parent):
    """refine_multilevel.rs:315-320: dense labels of the (child, parent) pairs in order of first appearance"""
    c, p = np.asarray(child, np.uint64), np.asarray(parent, np.uint64)
    return compact_labels(c * (np.uint64(int(p.max()) + 1) if p.size else np.uint64(1)) + p)


def child_offset_within_parent(child, parent):
    """refine_multilevel.rs:333-345: the index of an entity's child label inside its parent, in first-seen order"""
    per, out = {}, np.zeros(len(child), np.uint32)
    for i, (c, q) in enumerate(zip(np.asarray(child).tolist(), np.asarray(parent).tolist())):
        local = per.setdefault(q, {})
        out[i] = local.setdefault(c, len(local))
    return out


def compute_sibling_sets(refined, level, num_groups_at_level):
    """dc_poisson.rs:518-550: per entity the sorted groups of `level` that share its parent at `level + 1` (all groups at the
    coarsest level); entities with the same parent share one list object"""
    n = len(refined[level])
    if level + 1 >= len(refined):
        allg = list(range(num_groups_at_level))
        return [allg] * n
    kids = {}
    for c, q in zip(np.asarray(refined[level]).tolist(), np.asarray(refined[level + 1]).tolist()):
        kids.setdefault(q, set()).add(c)
    kids = {q: sorted(v) for q, v in kids.items()}
    return [kids[q] for q in np.asarray(refined[level + 1]).tolist()]


def intersect_with_siblings_fallback(siblings, neighbor_groups, current):
    """dc_poisson.rs:599-633"""
    if len(siblings) <= 1:
        return list(siblings)
    inter = [g for g in siblings if g in neighbor_groups]
    if not inter:
        return list(siblings)
    if current not in inter:
        inter = sorted(inter + [current])
    return inter


def build_candidate_sets(siblings, bbknn, pbsamp_to_group):
    """refine_multilevel.rs:85-112: siblings ∩ (groups of the BBKNN neighbours), sibling fall-back, own group always in"""
    lab = np.asarray(pbsamp_to_group).tolist()
    return [intersect_with_siblings_fallback(sib, {lab[j] for j in bbknn[e]}, lab[e]) for e, sib in enumerate(siblings)]


def _candidate_csr(refined, level, k, bbknn_matrix):
    """compute_sibling_sets + build_candidate_sets for every entity at once, as (cand_ptr, cand) of the CSR the ABI takes.
    bbknn_matrix: (npb, T) matched entities, NONE_U32 where there is none.  Same sets as the list forms above (tested); used
    when the npb x k membership matrices are small enough, which they are for every partition the path produces."""
    lab = np.asarray(refined[level], np.int64)
    n = len(lab)
    if level + 1 < len(refined):
        par = np.asarray(refined[level + 1], np.int64)
        parent_of_group = np.zeros(k, np.int64)
        parent_of_group[lab] = par  # a strict hierarchy: every entity of a group names the same parent
        sib = parent_of_group[None, :] == par[:, None]
    else:
        sib = np.ones((n, k), bool)
    bb = np.asarray(bbknn_matrix)
    ok = bb != NONE_U32
    nb_groups = np.zeros((n, k), bool)
    rows = np.broadcast_to(np.arange(n)[:, None], bb.shape)[ok]
    nb_groups[rows, lab[bb[ok].astype(np.int64)]] = True
    inter = sib & nb_groups
    some = inter.any(axis=1)
    inter[np.arange(n)[some], lab[some]] = True          # staying put is always legal (dc_poisson.rs:624-632)
    single = sib.sum(axis=1) <= 1
    use_sib = single | ~some                               # a lone sibling, or an empty intersection: the siblings (:604-617)
    cand = np.where(use_sib[:, None], sib, inter)
    ptr = np.zeros(n + 1, np.uint32)
    ptr[1:] = np.cumsum(cand.sum(axis=1))
    return ptr, np.nonzero(cand)[1].astype(np.uint32)


def refine_assignments(ctx: Context, gene_sums, bbknn, initial_per_level, reproject_offsets, params: "RefineParams"):
    """refine_multilevel.rs:170-298: top-down BBKNN + DC-Poisson refinement of the pb-sample -> group maps (finest first).
    gene_sums: (npb, D) dense pb-sample gene sums; bbknn: per pb-sample the matched foreign pb-samples.  Returns
    (pbsamp_to_group per level, num_groups per level, accepted moves)."""
    L = len(initial_per_level)
    if L == 0:
        raise LegumeError(1, "no levels")
    gs = _as(gene_sums, np.float32)
    npb, D = gs.shape
    for i, lvl in enumerate(initial_per_level):
        if len(lvl) != npb:
            raise LegumeError(1, f"level {i} has {len(lvl)} entries, expected {npb}")
    refined, ks = [], []
    for lvl in initial_per_level:
        c, k = compact_labels(lvl)
        refined.append(c)
        ks.append(k)
    if params.num_gibbs == 0 and params.num_greedy == 0:  # :215-222
        return refined, ks, 0
    # profiles once: stored entries of the gene sums, NB Fisher-information weights, size factors (:224-243)
    dev = _is_torch(gs)
    prof = gs.clone() if dev else np.array(gs, np.float32, copy=True)
    w = None
    if params.feature_weighting == "FisherInfoNb" and params.profile_source == "Raw":
        w = ctx.empty((D,), np.float32, dev)
        ctx.check(lib.lg_dcp_fisher_weights(ctx.h, _ptr(prof), D, npb, _ptr(w)))
    sf = ctx.empty((npb,), np.float32, dev)
    ctx.check(lib.lg_dcp_profiles(ctx.h, _ptr(prof), D, npb, _ptr(w), _ptr(sf)))
    rng = _Xoshiro256pp(params.seed)
    total = 0
    # bbknn: per entity a list of matched entities, or the (npb, T) matrix of per_batch_sc_neighbors with NONE_U32 padding
    bb_matrix = bbknn if isinstance(bbknn, np.ndarray) and bbknn.ndim == 2 else None
    if bb_matrix is not None:
        bbknn = None
    for level in range(L - 1, -1, -1):
        if level + 1 < L:  # re-anchor this level in its REFINED parent by the child hash relative to the parent (:255-280)
            off = reproject_offsets[level] if reproject_offsets is not None and level < len(reproject_offsets) else ()
            if len(off) == 0:
                off = child_offset_within_parent(initial_per_level[level], initial_per_level[level + 1])
            refined[level], ks[level] = project_to_refinement(off, refined[level + 1])
        k = ks[level]
        if bb_matrix is not None and npb * k <= (1 << 28):
            cptr, cflat = _candidate_csr(refined, level, k, bb_matrix)
        else:
            if bbknn is None:
                bbknn = [row[row != NONE_U32].tolist() for row in bb_matrix]
            cand = build_candidate_sets(compute_sibling_sets(refined, level, k), bbknn, refined[level])
            cptr = np.zeros(npb + 1, np.uint32)
            cptr[1:] = np.cumsum([len(c) for c in cand])
            cflat = np.fromiter((g for c in cand for g in c), np.uint32, int(cptr[-1]))
        base_seed = rng.next() | 1  # dc_poisson.rs:824
        labels = np.ascontiguousarray(refined[level], np.uint32)
        moves = C.c_uint64(0)
        ctx.check(lib.lg_dcp_refine_level(ctx.h, _ptr(prof), _ptr(sf), D, npb, _ptr(cptr), _ptr(cflat), k, params.num_gibbs,
                                          params.num_greedy, base_seed, params.gibbs_stagnation, _ptr(labels), C.byref(moves)))
        total += moves.value
        refined[level], ks[level] = compact_labels(labels)  # greedy sweeps can empty a group (:292-295)
    return refined, ks, total


def modal_groups(cell_to_pbsamp, num_pb, lvl):
    """collapse_data/mod.rs:823-841 for every pb-sample at once: the most frequent inherited label among its cells (the
    reference leaves the choice among equally frequent labels to its hash map; here the smallest label wins), 0 for an
    empty pb-sample"""
    c2p = np.asarray(cell_to_pbsamp).astype(np.int64)
    lvl = np.asarray(lvl).astype(np.int64)
    live = c2p != NONE_U32
    L = int(lvl.max()) + 1 if lvl.size else 1
    key, cnt = np.unique(c2p[live] * L + lvl[live], return_counts=True)
    pb, lab = key // L, key % L
    order = np.lexsort((lab, -cnt, pb))  # per pb-sample: highest count first, then the smallest label
    pb, lab = pb[order], lab[order]
    head = np.ones(len(pb), bool)
    head[1:] = pb[1:] != pb[:-1]
    out = np.zeros(num_pb, np.uint32)
    out[pb[head]] = lab[head]
    return out


# --------------------------------------------------------------------------------------------------
# GammaMatrix (matrix-param/src/dmatrix_gamma.rs) — TwoStatParam + Inference
# --------------------------------------------------------------------------------------------------
class GammaMatrix:
    def __init__(self, ctx: Context, dims, a0: float, b0: float):
        self.ctx = ctx
        self.num_rows, self.num_columns = dims
        self.a0, self.b0 = float(a0), float(b0)
        self.a_stat = np.full((self.num_columns, self.num_rows), a0, np.float32)
        self.b_stat = np.full((self.num_columns, self.num_rows), b0, np.float32)
        self.estimated_mean = np.zeros((self.num_columns, self.num_rows), np.float32)  # eager, zero start (:49-52)
        self.estimated_sd = self.estimated_log_mean = self.estimated_log_sd = None

    @classmethod
    def new(cls, ctx, dims, a0, b0):
        return cls(ctx, dims, a0, b0)

    def update_stat(self, update_a, update_b):
        self.reset_stat()
        self.add_stat(update_a, update_b)

    def add_stat(self, add_a, add_b):
        self.a_stat = self.a_stat + np.asarray(add_a, np.float32)
        self.b_stat = self.b_stat + np.asarray(add_b, np.float32)

    def reset_stat(self):
        self.a_stat = np.full_like(self.a_stat, self.a0)
        self.b_stat = np.full_like(self.b_stat, self.b0)

    def calibrate(self):
        self.calibrate_with(CalibrateTarget.All)

    def calibrate_with(self, target):
        n = self.a_stat.size
        mean = np.empty_like(self.a_stat)
        sd = np.empty_like(self.a_stat) if target == TARGET_ALL else None
        lm = np.empty_like(self.a_stat) if target != TARGET_MEAN_ONLY else None
        ls = np.empty_like(self.a_stat) if target == TARGET_ALL else None
        # a_stat/b_stat already carry the hyper-parameters (fill + add, dmatrix_gamma.rs:64-75)
        a = np.ascontiguousarray(self.a_stat)
        b = np.ascontiguousarray(self.b_stat)
        self.ctx.check(lib.lg_gamma_calibrate(self.ctx.h, _ptr(a), _ptr(b), n, 0.0, 0.0, target,
                                              _ptr(mean), _ptr(sd), _ptr(lm), _ptr(ls)))
        self.estimated_mean = mean
        if sd is not None:
            self.estimated_sd = sd
        if lm is not None:
            self.estimated_log_mean = lm
        if ls is not None:
            self.estimated_log_sd = ls

    @classmethod
    def vconcat(cls, blocks, stack_stats: bool):
        """dmatrix_gamma.rs:301-326: row blocks (same columns, same hyper-parameters) stacked into one parameter; the
        planes the first block has calibrated are stacked, the sufficient statistics only when asked for.  Planes are
        stored (columns, rows) here, so the rows are the trailing axis."""
        if not blocks:
            raise LegumeError(1, "vconcat of empty block list")
        first = blocks[0]
        if any(b.num_columns != first.num_columns for b in blocks):
            raise LegumeError(1, "vconcat: blocks differ in their column count")
        out = cls(first.ctx, (sum(b.num_rows for b in blocks), first.num_columns), first.a0, first.b0)
        stack = lambda name: np.concatenate([np.asarray(getattr(b, name)) for b in blocks], axis=1)
        out.a_stat = stack("a_stat") if stack_stats else None
        out.b_stat = stack("b_stat") if stack_stats else None
        for name in ("estimated_mean", "estimated_sd", "estimated_log_mean", "estimated_log_sd"):
            setattr(out, name, stack(name) if getattr(first, name) is not None else None)
        return out

    def posterior_mean(self):
        return self.estimated_mean

    def posterior_sd(self):
        return self.estimated_sd

    def posterior_log_mean(self):
        return self.estimated_log_mean

    def posterior_log_sd(self):
        return self.estimated_log_sd

    def nrows(self):
        return self.num_rows

    def ncols(self):
        return self.num_columns


# --------------------------------------------------------------------------------------------------
# CollapsedStat / optimize (collapse_data/stats.rs)
# --------------------------------------------------------------------------------------------------
class CollapsedStat:
    """stats.rs:546-582: sufficient statistics, all (S, D) / (B, D) / (S, B) in this module's layout"""

    def __init__(self, ngene, nsample, nbatch):
        z = lambda *s: np.zeros(s, np.float32)
        self.observed_sum_ds, self.imputed_sum_ds, self.residual_sum_ds = z(nsample, ngene), z(nsample, ngene), z(nsample, ngene)
        self.size_s = z(nsample)
        self.observed_sum_db = z(nbatch, ngene)
        self.n_bs = z(nsample, nbatch)
        # panel observability (stats.rs:556-567): (S, D) effective sizes and the (B, D) mask of delta, None = fully observed
        self.size_ds = None
        self.obs_mask_db = None

    def num_genes(self):
        return self.observed_sum_ds.shape[1]

    def num_samples(self):
        return self.observed_sum_ds.shape[0]

    def num_batches(self):
        return self.observed_sum_db.shape[0]

    def select_rows(self, r0, nrows):
        out = CollapsedStat(nrows, self.num_samples(), self.num_batches())
        sl = slice(r0, r0 + nrows)
        out.observed_sum_ds = np.ascontiguousarray(self.observed_sum_ds[:, sl])
        out.imputed_sum_ds = np.ascontiguousarray(self.imputed_sum_ds[:, sl])
        out.residual_sum_ds = np.ascontiguousarray(self.residual_sum_ds[:, sl])
        out.observed_sum_db = np.ascontiguousarray(self.observed_sum_db[:, sl])
        out.size_s, out.n_bs = self.size_s.copy(), self.n_bs.copy()
        if self.size_ds is not None:
            out.size_ds = np.ascontiguousarray(np.asarray(self.size_ds)[:, sl])
        if self.obs_mask_db is not None:
            out.obs_mask_db = np.ascontiguousarray(np.asarray(self.obs_mask_db)[:, sl])
        return out


class CollapsedOut(dict):
    """stats.rs:516-522: mu_observed, mu_adjusted, mu_residual, gamma, delta (dicts of posterior planes)"""

    def __getattr__(self, k):
        return self[k]


def optimize(ctx: Context, stat: CollapsedStat, hyper=(1.0, 1.0), num_iter=DEFAULT_OPT_ITER, out_target=TARGET_ALL):
    """stats.rs:378-512.  Gene blocking is numerically inert (:370-377), so the fit runs in one launch."""
    a0, b0 = hyper
    for name in ("observed_sum_ds", "imputed_sum_ds", "residual_sum_ds", "size_s", "observed_sum_db", "n_bs"):
        setattr(stat, name, _as(getattr(stat, name), np.float32))
    S, D = stat.observed_sum_ds.shape
    B = stat.num_batches()
    dev = _is_torch(stat.observed_sum_ds)
    new = lambda shape=(S, D): ctx.empty(shape, np.float32, dev)
    if B <= 1:
        mean, sd, lm, ls = new(), None, None, None
        if out_target == TARGET_ALL:
            sd, ls = new(), new()
        if out_target != TARGET_MEAN_ONLY:
            lm = new()
        size_ds = None if stat.size_ds is None else _as(stat.size_ds, np.float32)
        ctx.check(lib.lg_optimize_single_obs(ctx.h, _ptr(stat.observed_sum_ds), _ptr(stat.size_s), _ptr(size_ds), D, S, a0, b0,
                                             out_target, _ptr(mean), _ptr(sd), _ptr(lm), _ptr(ls)))
        return CollapsedOut(mu_observed=dict(mean=mean, sd=sd, log_mean=lm, log_sd=ls), mu_adjusted=None,
                            mu_residual=None, gamma=None, delta=None)
    mu_obs, mu_adj, mu_res, gam, delta = new(), new(), new(), new(), new((B, D))
    lm = new() if out_target != TARGET_MEAN_ONLY else None
    size_ds = None if stat.size_ds is None else _as(stat.size_ds, np.float32)
    mask = None if stat.obs_mask_db is None else _as(stat.obs_mask_db, np.float32)
    ctx.check(lib.lg_optimize_batched_obs(ctx.h, _ptr(stat.observed_sum_ds), _ptr(stat.imputed_sum_ds),
                                          _ptr(stat.residual_sum_ds), _ptr(stat.size_s), _ptr(size_ds),
                                          _ptr(stat.observed_sum_db), _ptr(stat.n_bs), _ptr(mask), D, S, B, a0, b0, num_iter,
                                          out_target, _ptr(mu_obs), _ptr(mu_adj), _ptr(mu_res), _ptr(gam), _ptr(delta),
                                          _ptr(lm)))
    return CollapsedOut(mu_observed=dict(mean=mu_obs), mu_adjusted=dict(mean=mu_adj, log_mean=lm),
                        mu_residual=dict(mean=mu_res), gamma=dict(mean=gam), delta=dict(mean=delta))


# --------------------------------------------------------------------------------------------------
# SparseIoVec: the data handle the reference's traits are implemented on
# --------------------------------------------------------------------------------------------------
class SparseRunningStatistics:
    """matrix-util/src/sparse_stat.rs:33-198, 404-431 (T = f32): per-row (npos, s1, s2) over the columns seen so far.
    The sufficient statistics are held as f64 — exact whole numbers for count data, so blocks and shards merge to
    the same totals in any order — and narrowed to the reference's f32 by the accessors."""

    def __init__(self, nrows):
        self._nrows, self._ncols = int(nrows), 0
        self._npos, self._s1, self._s2 = (np.zeros(self._nrows, np.float64) for _ in range(3))

    def nrows(self):
        return self._nrows

    def ncols_processed(self):
        return self._ncols

    def add_block(self, ctx: "Context", block: "CscBlock"):
        """add_csc over a device-resident block (sparse_stat.rs:97-108)"""
        if block.nrows != self._nrows:
            raise LegumeError(1, "SparseRunningStatistics: row count mismatch")
        out = [np.empty(self._nrows, np.float64) for _ in range(3)]
        ctx.check(lib.lg_row_stats(ctx.h, block.h, _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        self._npos += out[0]
        self._s1 += out[1]
        self._s2 += out[2]
        self._ncols += block.ncols

    def merge(self, other: "SparseRunningStatistics"):
        """sparse_stat.rs:183-196"""
        self._npos += other._npos
        self._s1 += other._s1
        self._s2 += other._s2
        self._ncols += other._ncols

    def _denom(self):
        return np.float32(self._ncols) if self._ncols > 0 else np.float32(1e-8)  # safe_denom, :16-23

    def count_positives(self):
        return self._npos.astype(np.float32)

    def sum(self):
        return self._s1.astype(np.float32)

    def mean(self):
        return self.sum() / self._denom()

    def variance(self):
        mu = self.mean()
        return self._s2.astype(np.float32) / self._denom() - mu * mu

    def std(self):
        with np.errstate(invalid="ignore"):
            return np.sqrt(self.variance())

    def to_vecs(self):
        return self.count_positives(), self.sum(), self.mean(), self.std()


def nystrom_project(ctx: "Context", block: "CscBlock", basis_dk, delta_dp=None, pb_of_cell=None, column_sum_norm=1e4):
    """nystrom_proj_visitor over a block (senna/src/svd/fit.rs:433-466).  basis_dk: (K, D) array = the reference's
    D x K column-major DMatrix; delta_dp: (P, D) = D x P or None; returns (N, K) = K x N column-major, next to the inputs."""
    basis_dk = _as(basis_dk, np.float32)
    K = int(basis_dk.shape[0])
    if int(basis_dk.shape[1]) != block.nrows:
        raise LegumeError(1, "nystrom_project: basis rows mismatch the number of genes")
    P = 0
    if delta_dp is not None:
        delta_dp = _as(delta_dp, np.float32)
        P = int(delta_dp.shape[0])
        if int(delta_dp.shape[1]) != block.nrows:
            raise LegumeError(1, "nystrom_project: delta rows mismatch the number of genes")
        if pb_of_cell is None or len(pb_of_cell) != block.ncols:
            raise LegumeError(1, "nystrom_project: delta needs the pseudobulk of every cell")
        pb_of_cell = _as(pb_of_cell, np.uint32)
    out = ctx.empty((block.ncols, K), np.float32, _is_torch(basis_dk))
    ctx.check(lib.lg_nystrom_project(ctx.h, block.h, _ptr(basis_dk), K, _ptr(delta_dp),
                                     _ptr(pb_of_cell) if delta_dp is not None else None, P, float(column_sum_norm), _ptr(out)))
    return out


class SparseIoVec:
    """One preloaded backend's columns on the device, with the derived caches of
    data-beans/src/sparse_io_vector/mod.rs:70-85 (groups, batches, multiplicity)."""

    def __init__(self, ctx: Context, block: CscBlock):
        self.ctx, self.block = ctx, block
        # the reference's flow is one projection and then one collapse per level (and per label kind): the block keeps the
        # projection's 1-bit pattern for them (lg_csc_keep_pattern; same sums) unless LG_KEEP_PATTERN=0 or memory is short —
        # the buffers take about half of what the block itself does
        if block.ncols:
            ctx.check(lib.lg_csc_keep_pattern(ctx.h, block.h, 2))
        self.col_to_group = None      # uint32[N]
        self.group_keys = None
        self.col_to_batch = None      # uint32[N]
        self.batch_names = None
        self.multiplicity = None      # float32[N] or None
        self.batch_proj = None        # (N, K) features the per-batch kNN dictionaries were built on
        self.between_batch_proximity = None  # uint32 (B, B) or None (batch.rs:176-178: only when B > 2)
        self.row_coverage = None      # bool (nbackends, D): which rows every backend measures; None = all of them
        self.col_source = None        # uint32[N]: the backend every column came from

    @classmethod
    def from_csc(cls, ctx, indptr, indices, data, nrows):
        return cls(ctx, CscBlock.upload(ctx, indptr, indices, data, nrows))

    @classmethod
    def from_zarr(cls, ctx, zarr_file, col_lo=0, col_hi=None):
        """open_sparse_matrix + SparseIoVec::push of one zarr backend (sparse_io_vector/mod.rs `push`)"""
        be = SparseMtxData.open(zarr_file)
        try:
            return cls(ctx, be.read_columns_csc(ctx, col_lo, col_hi))
        finally:
            be.close()

    @classmethod
    def from_backends(cls, ctx, backends, nrows):
        """Several backends' columns side by side, as SparseIoVec::push + read_columns_csc see them
        (data-beans/src/sparse_io_vector/read.rs:172-285): every backend is (indptr, indices, data, row_remap or None),
        row_remap[local row] = row in the union (`g2c[l2g[row]]`, :202-219; rows a backend lacks simply never occur).
        Every backend goes through lg_csc_upload_remap, which makes its columns canonical the way the reference does
        (a remap that is not monotone leaves a column unsorted, a many-to-one remap leaves duplicate rows: sorted and
        summed, read.rs:246-281); the blocks are then joined on the device (lg_csc_concat)."""
        blocks = [CscBlock.upload(ctx, ip, ix, v, nrows, row_remap=remap) for ip, ix, v, remap in backends]
        # row_coverage_by_backend / column_source (sparse_io_vector/mod.rs:378-403): a backend without a remap measures
        # every row, one with a remap the rows its map reaches
        cov = np.ones((len(backends), nrows), bool)
        for d, (_, _, _, remap) in enumerate(backends):
            if remap is not None:
                r = np.asarray(remap).astype(np.int64)
                cov[d] = False
                cov[d, r[(r >= 0) & (r < nrows)]] = True
        src = np.concatenate([np.full(b.ncols, d, np.uint32) for d, b in enumerate(blocks)]) if blocks else np.zeros(0, np.uint32)
        if len(blocks) == 1:
            out = cls(ctx, blocks[0])
        else:
            blk = CscBlock.concat(ctx, blocks)
            ctx.sync()
            for b in blocks:
                b.free()
            out = cls(ctx, blk)
        out.row_coverage, out.col_source = (cov if not cov.all() else None), src
        return out

    def row_coverage_by_backend(self):
        return self.row_coverage

    def attach_observability(self, stat: "CollapsedStat"):
        """collapse_data/mod.rs:221-301: no-op when every backend measures every row"""
        cov = self.row_coverage_by_backend()
        if cov is None:
            return
        S, B, D = stat.num_samples(), stat.num_batches(), self.num_rows()
        grp = _as(self.get_group_membership(), np.uint32)
        bat = None if self.col_to_batch is None or B == 0 else _as(self.col_to_batch, np.uint32)
        size_ds = np.empty((S, D), np.float32)
        mask = np.empty((B, D), np.float32) if bat is not None else None
        has_zero = C.c_int(0)
        self.ctx.check(lib.lg_attach_observability(self.ctx.h, _ptr(np.ascontiguousarray(cov, np.uint8)), cov.shape[0],
                                                   _ptr(self.col_source), _ptr(grp), _ptr(bat), _ptr(self.multiplicity),
                                                   self.num_columns(), D, S, B, _ptr(size_ds), _ptr(mask),
                                                   C.byref(has_zero) if mask is not None else None))
        stat.size_ds = size_ds
        stat.obs_mask_db = mask if (mask is not None and has_zero.value) else None

    def num_rows(self):
        return self.block.nrows

    def num_columns(self):
        return self.block.ncols

    # ---- batch.rs:259-336 ----
    def register_batch_membership(self, labels):
        if len(labels) != self.num_columns():
            raise LegumeError(1, "batch membership length mismatches the number of columns")
        self.col_to_batch, self.batch_names = _rank_labels(labels)

    def num_batches(self):
        return 0 if self.batch_names is None else len(self.batch_names)

    def register_column_multiplicity(self, weights):
        w = np.ascontiguousarray(weights, np.float32)
        if len(w) != self.num_columns():
            raise LegumeError(1, "column multiplicity length mismatches the number of columns")
        if not np.all(w > 0):
            raise LegumeError(1, "column multiplicity must be strictly positive")
        self.multiplicity = w

    # ---- groups.rs:13-37 ----
    def assign_groups(self, column_to_group, ncolumns_per_group=None):
        """groups.rs:13-37.  ncolumns_per_group: groups with more columns are down-sampled to that many
        (partition_by_membership, matrix-util/src/utils.rs:36-66): one generator per group, seeded with
        mix_seed(PARTITION_SHUFFLE_SEED, smallest member) so the subset does not depend on thread order; the columns
        left out belong to no group (label 0xFFFFFFFF: every statistic skips them).  The subset itself is NOT the
        reference's: that is rand 0.9's SmallRng + SliceRandom::shuffle, a third-party stream that is not restated
        (parity unpinned) — the sizes, the seeding rule and the run-to-run stability are."""
        if len(column_to_group) != self.num_columns():
            raise LegumeError(1, "group membership length mismatches the number of columns")
        self.col_to_group, self.group_keys = _rank_labels(column_to_group)
        if ncolumns_per_group is not None:
            self.col_to_group = downsample_groups(self.col_to_group, len(self.group_keys), int(ncolumns_per_group))

    def num_groups(self):
        return 0 if self.group_keys is None else len(self.group_keys)

    def get_group_membership(self):
        if self.col_to_group is None:
            raise LegumeError(1, "groups were not assigned")
        return self.col_to_group

    # ---- RandProjOps (random_projection.rs:341-527) ----
    def _basis(self, target_dim, seed, basis):
        if basis is None:
            # NOTE: the reference draws the basis from rand 0.10 StdRng + rand_distr ziggurat
            # (rand_util.rs:54-76); that stream is not reproducible here, so parity runs pass `basis`.
            basis = np.random.default_rng(seed & 0xFFFFFFFF).standard_normal((self.num_rows(), target_dim)).astype(np.float32)
        basis = _as(basis, np.float32)
        if tuple(basis.shape) != (self.num_rows(), target_dim):
            raise LegumeError(1, f"basis must be K x D given as shape (D, K) = ({self.num_rows()}, {target_dim})")
        return basis

    def project_columns(self, target_dim, block_size=None, basis=None, exact=False):
        return self.project_columns_with_batch_correction(target_dim, block_size, None, basis=basis, exact=exact)

    def project_columns_with_batch_correction(self, target_dim, block_size=None, batch_membership=None, basis=None,
                                              seed=DEFAULT_PROJECTION_SEED, exact=False):
        """returns (basis_kd as (D, K), proj as (N, K)); block_size is accepted and ignored (the GPU streams
        whole column ranges).  exact=True runs the exact-order kernels (lg_project_exact): bit-identical to the CPU
        path on count data, several times slower; the default is the tensor-core path with a 1e-5 contract."""
        basis = self._basis(target_dim, seed, basis)
        n = self.num_columns()
        batch, nb = None, 0
        if batch_membership is not None:
            if len(batch_membership) == n:
                batch, names = _rank_labels(batch_membership)
                nb = len(names)
            # else: the reference warns and skips the centring (:389-395)
        dev = _is_torch(basis)
        proj = self.ctx.empty((n, target_dim), np.float32, dev)
        if dev and batch is not None:
            import torch
            batch = torch.from_numpy(batch.astype(np.int32)).to(basis.device)
        fn = lib.lg_project_exact if exact else lib.lg_project
        self.ctx.check(fn(self.ctx.h, self.block.h, _ptr(basis), target_dim, _ptr(batch), nb, _ptr(proj)))
        return basis, proj

    def project_columns_weighted(self, target_dim, block_size, batch_membership, row_weights, basis=None,
                                 seed=DEFAULT_PROJECTION_SEED, exact=False):
        """random_projection.rs:417-495: rows with w <= 0 are zeroed, |w - 1| > 1e-6 scaled"""
        row_weights = np.asarray(row_weights, np.float32)
        if len(row_weights) != self.num_rows():
            raise LegumeError(1, "row_weights length mismatch")
        basis = np.array(self._basis(target_dim, seed, basis), np.float32, copy=True)
        for r, w in enumerate(row_weights):
            if w <= 0.0:
                basis[r, :] = 0.0
            elif abs(w - 1.0) > 1e-6:
                basis[r, :] *= w
        return self.project_columns_with_batch_correction(target_dim, block_size, batch_membership, basis=basis, exact=exact)

    def partition_columns_to_groups(self, proj_kn, num_features=None, ncols_per_group=None):
        """random_projection.rs:506-527: returns max code + 1 and assigns groups"""
        n, K = proj_kn.shape
        if n != self.num_columns():
            raise LegumeError(1, "number of columns mismatch")
        kk = min(K, num_features if num_features is not None else K, n)
        codes = binary_sort_columns(self.ctx, proj_kn, kk)
        codes_h = codes.cpu().numpy().astype(np.uint64) if _is_torch(codes) else codes
        self.binary_codes = codes_h
        group, ng = assign_groups_from_codes(self.ctx, codes_h, kk)
        self.col_to_group = group if ncols_per_group is None else downsample_groups(np.asarray(group), ng, int(ncols_per_group))
        self.group_keys = sorted({str(int(c)) for c in np.unique(codes_h)}, key=lambda s: s.encode())
        return int(codes_h.max()) + 1

    # ---- the nnz streams either side of the path ----
    def streaming_sparse_running_stats(self, block_size=None, progress_label=""):
        """data-beans-alg/src/sparse_streaming.rs:23-60: per-gene (npos, sum, sum_sq) over every column"""
        stats = SparseRunningStatistics(self.num_rows())
        stats.add_block(self.ctx, self.block)
        return stats

    def nystrom_project(self, basis_dk, delta_dp=None, column_sum_norm=1e4):
        """do_nystrom_proj's visitor pass (senna/src/svd/fit.rs:421-466); the pseudobulk of a cell is its group"""
        pb = self.get_group_membership() if delta_dp is not None else None
        return nystrom_project(self.ctx, self.block, basis_dk, delta_dp, pb, column_sum_norm)

    # ---- CollapsingOps (collapse_data/mod.rs:315-483) ----
    def collect_basic_stat(self, stat: CollapsedStat):
        S = stat.num_samples()
        stat.observed_sum_ds, stat.size_s = _as(stat.observed_sum_ds, np.float32), _as(stat.size_s, np.float32)
        self.ctx.check(lib.lg_collapse_basic(self.ctx.h, self.block.h, _ptr(self.get_group_membership()),
                                             _ptr(self.multiplicity), S, _ptr(stat.observed_sum_ds), _ptr(stat.size_s)))

    def collect_batch_stat(self, stat: CollapsedStat):
        if self.col_to_batch is None:
            raise LegumeError(1, "batches were not registered")
        S, B = stat.num_samples(), stat.num_batches()
        self.ctx.check(lib.lg_collapse_batch(self.ctx.h, self.block.h, _ptr(self.get_group_membership()),
                                             _ptr(self.col_to_batch), _ptr(self.multiplicity), S, B,
                                             _ptr(stat.observed_sum_db), _ptr(stat.n_bs)))

    # ---- batch.rs:46-234 ----
    def build_hnsw_per_batch(self, proj_kn, batch_membership):
        """collapse_data/mod.rs:364-383 -> register_batches_dmatrix (batch.rs:46-180).  The per-batch
        dictionaries are the exact backend at every size: they are not materialised; the
        features are kept and searched by lg_knn_match_batches / lg_pb_match."""
        if len(batch_membership) != self.num_columns():
            raise LegumeError(1, "batch membership length mismatches the number of columns")
        self.register_batch_membership(batch_membership)
        self.batch_proj = _as(proj_kn, np.float32)
        self.between_batch_proximity = None
        if self.num_batches() > 2:
            self.between_batch_proximity, _ = sort_batch_proximity(self.ctx, self.batch_proj, self.col_to_batch,
                                                                   self.num_batches())

    register_batches_dmatrix = build_hnsw_per_batch

    def neighbouring_matches(self, knn_batches, knn_columns, skip_same_batch=True, skip_batches=None,
                             target_batches=None):
        """the kNN part of read_neighbouring_columns_csc / read_matched_columns_csc (matched.rs:173-260, 97-158):
        (matched_columns (N, T) global indices, distances (N, T)); slot i*knn + r = r-th nearest cell of the i-th
        neighbouring batch.  knn_batches only sizes a Vec in the reference (matched.rs:201) and is ignored."""
        if self.batch_proj is None:
            raise LegumeError(1, "no knn lookup")
        if not skip_same_batch:
            raise LegumeError(1, "skip_same_batch = false is not used on the hot path")
        B = self.num_batches()
        if target_batches is not None:
            order = np.tile(np.asarray(target_batches, np.uint32)[None, :], (B, 1))
        elif self.between_batch_proximity is not None:
            order = np.array(self.between_batch_proximity, np.uint32, copy=True)
        else:
            order = np.tile(np.arange(B, dtype=np.uint32)[None, :], (B, 1))
        if skip_batches is not None:
            order[np.isin(order, np.asarray(skip_batches, np.uint32))] = 0xFFFFFFFF
        return knn_match_batches(self.ctx, self.batch_proj, self.col_to_batch, B, knn_columns, order)

    def collect_matched_stat(self, knn_batches, knn_cells, reference_indices, stat: CollapsedStat):
        """collapse_data/mod.rs:485-500 -> collect_matched_stat_visitor (stats.rs:26-108)"""
        midx, mdist = self.neighbouring_matches(knn_batches, knn_cells, True, None, reference_indices)
        S, D = stat.num_samples(), self.num_rows()
        dev = _is_torch(midx)
        imp, res = self.ctx.empty((S, D), np.float32, dev), self.ctx.empty((S, D), np.float32, dev)
        grp = self.get_group_membership()
        if dev:
            import torch
            grp = torch.from_numpy(np.asarray(grp).astype(np.int32)).to(midx.device)
        self.ctx.check(lib.lg_collect_matched_stat(self.ctx.h, self.block.h, _ptr(grp), S, _ptr(midx), _ptr(mdist),
                                                   midx.shape[1], _ptr(imp), _ptr(res)))
        stat.imputed_sum_ds, stat.residual_sum_ds = imp, res

    def collapse_columns(self, knn_batches=None, knn_cells=None, reference_batch_names=None, num_opt_iter=None,
                         out_target=TARGET_ALL):
        """collapse_data/mod.rs:384-475: basic stats; with B > 1 also batch stats and the per-cell matched stats
        (cross-batch kNN neighbourhood adjustment), then the Poisson-Gamma fit"""
        if self.col_to_group is None:
            raise LegumeError(1, "The columns were not assigned before. Call `assign_columns_to_groups`")
        nb = self.num_batches()
        stat = CollapsedStat(self.num_rows(), self.num_groups(), nb)
        self.collect_basic_stat(stat)
        if nb > 1:
            ref = None
            if reference_batch_names is not None:
                lut = {k: i for i, k in enumerate(self.batch_names)}
                ref = [lut[str(b)] for b in reference_batch_names if str(b) in lut]
                if not ref:
                    raise LegumeError(1, "no reference batch names matched!")
            self.collect_batch_stat(stat)
            self.collect_matched_stat(2 if knn_batches is None else knn_batches,
                                      DEFAULT_KNN if knn_cells is None else knn_cells, ref, stat)
        return optimize(self.ctx, stat, (1.0, 1.0), num_opt_iter or DEFAULT_OPT_ITER, out_target), stat

    # ---- MultilevelCollapsingOps (collapse_data/mod.rs:503-1050) ----
    def _multilevel_prologue(self, proj_kn, batch_membership, params):
        """what all three entries share (mod.rs:542-560, 640-662, 876-910): batches, per-batch dictionaries, level dims,
        finest codes, finest hash groups"""
        if params.anchor_batches or params.bulk_batches:
            raise LegumeError(1, "anchor / bulk batches are outside the hot path (pb_samples.rs:75-88)")
        ctx = self.ctx
        proj_kn = _as(proj_kn, np.float32)
        n, K = proj_kn.shape
        self.register_batch_membership(batch_membership)
        nb = self.num_batches()
        if nb >= 2:
            self.build_hnsw_per_batch(proj_kn, batch_membership)
        level_dims = compute_level_sort_dims(params.sort_dim, params.num_levels)
        kk = min(K, level_dims[0], n)
        fine_codes = binary_sort_columns(ctx, proj_kn, kk)
        codes_h = fine_codes.cpu().numpy().astype(np.uint64) if _is_torch(fine_codes) else fine_codes
        group, ng = assign_groups_from_codes(ctx, codes_h, kk)
        self.col_to_group, self.binary_codes = group, codes_h
        self.group_keys = sorted({str(int(c)) for c in np.unique(codes_h)}, key=lambda s: s.encode())
        return proj_kn, nb, level_dims, codes_h, np.asarray(group), ng

    def _build_pb_samples(self, proj_kn, group, ng, nb):
        """pb_samples.rs:274-307: layout + per-pb-sample gene sums"""
        ctx = self.ctx
        bat = self.col_to_batch if self.col_to_batch is not None else np.zeros(self.num_columns(), np.uint32)
        layout = build_pb_sample_layout(ctx, group, ng, bat, max(nb, 1), proj_kn, self.multiplicity)
        gs = CollapsedStat(self.num_rows(), layout.num_pb, 1)
        ctx.check(lib.lg_collapse_basic(ctx.h, self.block.h, _ptr(layout.cell_to_pbsamp), _ptr(self.multiplicity),
                                        layout.num_pb, _ptr(gs.observed_sum_ds), _ptr(gs.size_s)))
        return layout, gs.observed_sum_ds

    def _merge_level(self, prev: CollapsedStat, f2c, nc, nb):
        """merge_stat (stats.rs:790-833): sums, sizes and effective sizes add per coarse group in ascending fine index"""
        ctx = self.ctx
        coarse = CollapsedStat(self.num_rows(), nc, nb)
        for name in ("observed_sum_ds", "imputed_sum_ds", "residual_sum_ds"):
            setattr(coarse, name, merge_stat(ctx, getattr(prev, name), f2c, nc))
        prev_n = len(f2c)
        size, nbs = np.zeros(nc, np.float32), np.zeros((nc, max(nb, 1)), np.float32)
        psize, pnbs = np.asarray(prev.size_s, np.float32), np.asarray(prev.n_bs, np.float32).reshape(prev_n, -1)
        for f, c in enumerate(f2c):  # ascending fine index, f32 adds (stats.rs:813-816)
            size[c] += psize[f]
            nbs[c] += pnbs[f]
        coarse.size_s, coarse.n_bs = size, nbs
        coarse.observed_sum_db = prev.observed_sum_db
        if prev.size_ds is not None:
            coarse.size_ds = merge_stat(ctx, prev.size_ds, f2c, nc)
        coarse.obs_mask_db = prev.obs_mask_db
        return coarse

    def _collect_refined_levels(self, proj_kn, nb, layout, gene_sums, p2g_levels, k_levels, params, target):
        """the post-refinement tail shared by refine_and_collect_single_layer (refine.rs:380-500) and
        collapse_columns_multilevel_with_partition (mod.rs:715-815): finest groups from the pb-sample partition, one data
        pass for the finest statistics, merge_stat descent, the per-level cell -> pb map"""
        ctx = self.ctx
        c2p = np.asarray(layout.cell_to_pbsamp).astype(np.int64)
        k0 = k_levels[0]
        fine = np.asarray(p2g_levels[0], np.uint32)[c2p]
        # pad_numeric_labels + assign_groups (refine.rs:21-35, groups.rs:13-37): the rank of a zero-padded number is the number
        self.col_to_group, self.group_keys = fine, pad_numeric_labels(range(k0), k0)
        fine_stat = CollapsedStat(self.num_rows(), k0, nb)
        self.collect_basic_stat(fine_stat)
        if nb >= 2:
            self.collect_batch_stat(fine_stat)
            matched = per_batch_sc_neighbors(ctx, layout, proj_kn, self.col_to_batch, nb, params.knn_pb_samples)
            collect_matched_stat_coarse(ctx, layout, gene_sums, p2g_levels[0], matched, fine_stat)
        if params.observe_panels:
            self.attach_observability(fine_stat)
        outs = [optimize(ctx, fine_stat, (1.0, 1.0), params.num_opt_iter, target)]
        stats = [fine_stat]
        prev = fine_stat
        for level in range(1, len(k_levels)):
            f2c = fine_to_coarse_from_refined(p2g_levels[level - 1], p2g_levels[level], k_levels[level - 1])
            coarse = self._merge_level(prev, f2c, k_levels[level], nb)
            outs.append(optimize(ctx, coarse, (1.0, 1.0), max(params.num_opt_iter // 2, 10), target))
            stats.append(coarse)
            prev = coarse
        cell_to_pb = [np.asarray(p2g, np.uint32)[c2p] for p2g in p2g_levels]
        return dict(levels=outs, cell_to_pb_per_level=cell_to_pb, stats=stats)

    def _refined_partition(self, proj_kn, nb, level_dims, codes_h, layout, gene_sums, params):
        """every level's pb-sample -> group and group count: refine_or_identity(num_batches >= 2, ..) (refine.rs:126-147,
        313-345) — with one batch the compacted hash partition, with more the BBKNN candidates + DC-Poisson sweeps"""
        c2p = np.asarray(layout.cell_to_pbsamp).astype(np.int64)
        first = np.full(layout.num_pb, len(c2p), np.int64)
        np.minimum.at(first, c2p, np.arange(len(c2p)))  # a pb-sample's first cell (pb_samples.rs:472-481)
        p2g = initial_per_level_from_hash(codes_h, first, level_dims)
        if nb >= 2:
            matched, _ = per_batch_sc_neighbors(self.ctx, layout, proj_kn, self.col_to_batch, nb, params.knn_pb_samples)
            bbknn = np.asarray(matched.cpu() if _is_torch(matched) else matched)  # build_bbknn_neighbors (refine_multilevel.rs:60-83)
            offsets = build_reproject_offsets(codes_h, first, level_dims)
            profiles = gene_sums
            if params.refine.profile_source == "Projected":  # Profiles::from_projection: a pb-sample's projection columns summed
                profiles = self._projection_sums(proj_kn, layout)  # in ascending cell order (serial f32 folds, dc_poisson.rs:171-178)
            p2g, k, moves = refine_assignments(self.ctx, profiles, bbknn, p2g, offsets, params.refine)
            self.refine_moves = moves
        else:
            p2g, k = zip(*(compact_labels(l) for l in p2g))
        return list(p2g), list(k)

    def _projection_sums(self, proj_kn, layout):
        """(npb, K): sum of every pb-sample's projection columns, cells ascending (lg_pb_centroid_fold on zeroed running sums)"""
        import torch
        n, K = proj_kn.shape
        dev = torch.device("cuda", self.ctx.device)
        dproj = proj_kn.to(dev) if _is_torch(proj_kn) else torch.from_numpy(np.ascontiguousarray(proj_kn, np.float32)).to(dev)
        c2p = torch.from_numpy(np.asarray(layout.cell_to_pbsamp).astype(np.int32)).to(dev)
        csum = torch.zeros((layout.num_pb, K), dtype=torch.float32, device=dev)
        ccnt = torch.zeros(layout.num_pb, dtype=torch.float32, device=dev)
        self.ctx.check(lib.lg_pb_centroid_fold(self.ctx.h, _ptr(dproj), K, n, _ptr(c2p), layout.num_pb, None, _ptr(csum), _ptr(ccnt)))
        self.ctx.sync()
        return csum.cpu().numpy()

    def _refine_and_collect(self, proj_kn, nb, level_dims, codes_h, group, ng, params):
        """refine_and_collect_single_layer (refine.rs:264-500)"""
        layout, gene_sums = self._build_pb_samples(proj_kn, group, ng, nb)
        p2g, k = self._refined_partition(proj_kn, nb, level_dims, codes_h, layout, gene_sums, params)
        return self._collect_refined_levels(proj_kn, nb, layout, gene_sums, p2g, k, params, params.output_calibration)

    def _legacy_levels(self, proj_kn, nb, level_dims, codes_h, group, ng, params):
        """the un-refined descent (mod.rs:943-1046): finest statistics from the hash groups, coarser levels by masked codes"""
        ctx = self.ctx
        fine_stat = CollapsedStat(self.num_rows(), ng, nb)
        self.collect_basic_stat(fine_stat)
        if nb >= 2:
            self.collect_batch_stat(fine_stat)
            layout, gene_sums = self._build_pb_samples(proj_kn, group, ng, nb)
            matched = per_batch_sc_neighbors(ctx, layout, proj_kn, self.col_to_batch, nb, params.knn_pb_samples)
            collect_matched_stat_coarse(ctx, layout, gene_sums, layout.pb_sample_to_group, matched, fine_stat)
        if params.observe_panels:
            self.attach_observability(fine_stat)
        outs = [optimize(ctx, fine_stat, (1.0, 1.0), params.num_opt_iter, TARGET_ALL)]
        stats = [fine_stat]
        prev, prev_group, prev_n = fine_stat, group, ng
        for dim in level_dims[1:]:
            f2c, nc = compute_fine_to_coarse_mapping(ctx, codes_h, prev_group, prev_n, dim)
            coarse = self._merge_level(prev, f2c, nc, nb)
            outs.append(optimize(ctx, coarse, (1.0, 1.0), max(params.num_opt_iter // 2, 10), TARGET_ALL))
            stats.append(coarse)
            prev, prev_group, prev_n = coarse, f2c[prev_group], nc
        return outs, stats

    def collapse_columns_multilevel_with_hierarchy(self, proj_kn, batch_membership, params: MultilevelParams):
        """collapse_data/mod.rs:534-607: the levels plus the per-level cell -> pb map; needs params.refine"""
        if params.refine is None:
            raise LegumeError(1, "collapse_columns_multilevel_with_hierarchy requires MultilevelParams.refine = Some(..); "
                                 "the legacy non-refinement path doesn't surface per-level cell->pb mappings")
        proj_kn, nb, level_dims, codes_h, group, ng = self._multilevel_prologue(proj_kn, batch_membership, params)
        return self._refine_and_collect(proj_kn, nb, level_dims, codes_h, group, ng, params)

    def collapse_columns_multilevel_with_partition(self, proj_kn, batch_membership, params: MultilevelParams,
                                                   cell_to_pb_per_level):
        """collapse_data/mod.rs:617-821: skip the refinement and take every level's pb-sample -> group from an inherited
        cell -> pb map (finest first) by majority vote inside each pb-sample"""
        proj_kn, nb, level_dims, codes_h, group, ng = self._multilevel_prologue(proj_kn, batch_membership, params)
        if len(cell_to_pb_per_level) != len(level_dims):
            raise LegumeError(1, f"inherited cell_to_pb has {len(cell_to_pb_per_level)} levels but --num-levels is "
                                 f"{len(level_dims)}; pass --num-levels to match the source run")
        layout, gene_sums = self._build_pb_samples(proj_kn, group, ng, nb)
        p2g, k = [], []
        for i, lvl in enumerate(cell_to_pb_per_level):
            if len(lvl) != self.num_columns():
                raise LegumeError(1, f"inherited cell_to_pb level {i} has {len(lvl)} cells, data has {self.num_columns()}")
            compact, kl = compact_labels(modal_groups(layout.cell_to_pbsamp, layout.num_pb, lvl))
            p2g.append(compact)
            k.append(kl)
        return self._collect_refined_levels(proj_kn, nb, layout, gene_sums, p2g, k, params, TARGET_ALL)

    def collapse_columns_multilevel_vec(self, proj_kn, batch_membership, params: MultilevelParams):
        """collapse_data/mod.rs:867-1046; returns (levels finest-first: list of CollapsedOut, list of CollapsedStat)"""
        proj_kn, nb, level_dims, codes_h, group, ng = self._multilevel_prologue(proj_kn, batch_membership, params)
        ctx = self.ctx
        if params.refine is not None:  # mod.rs:914-941
            out = self._refine_and_collect(proj_kn, nb, level_dims, codes_h, group, ng, params)
            return out["levels"], out["stats"]
        return self._legacy_levels(proj_kn, nb, level_dims, codes_h, group, ng, params)


# --------------------------------------------------------------------------------------------------
# ColumnDict (matrix-util/src/knn/mod.rs) with the exact backend
# --------------------------------------------------------------------------------------------------
def mix_seed(base: int, salt: int) -> int:
    """SplitMix64 avalanche of (base, salt): matrix-util/src/rand_util.rs:30-35"""
    m = 0xFFFFFFFFFFFFFFFF
    z = (base ^ (salt * 0x9E3779B97F4A7C15)) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


class SparseIoStack:
    """Several modalities over the same cells (data-beans/src/sparse_io_stack.rs): RandProjOps runs every modality with
    its own sub-seed mix_seed(seed, m) and stacks the results vertically (random_projection.rs:200-330)."""

    def __init__(self, stack):
        self.stack = list(stack)
        if not self.stack:
            raise LegumeError(1, "SparseIoStack: empty stack")

    def num_columns(self):
        return max(x.num_columns() for x in self.stack)  # sparse_io_stack.rs:48-54

    def num_rows(self):
        return sum(x.num_rows() for x in self.stack)

    def _stacked(self, one, target_dim, batch_membership, bases, seed):
        n = self.num_columns()
        if batch_membership is not None:
            batch_membership = batch_membership[:n]  # `.get(0..ncols)` (:213)
        if bases is not None and len(bases) != len(self.stack):
            raise LegumeError(1, "SparseIoStack: one basis per modality")
        outs = [one(m, x, batch_membership, None if bases is None else bases[m], mix_seed(seed, m))
                for m, x in enumerate(self.stack)]
        if any(p.shape[0] != outs[0][1].shape[0] for _, p in outs):
            raise LegumeError(1, "SparseIoStack: modalities disagree on the number of columns")  # concatenate_vertical fails
        basis = np.concatenate([np.asarray(b) for b, _ in outs], axis=0)              # (sum D_m, K): vertical concat of D x K
        proj = np.concatenate([np.asarray(p) for _, p in outs], axis=1)               # (N, M * K): vertical concat of K x N
        return basis, proj

    def project_columns_with_batch_correction(self, target_dim, block_size=None, batch_membership=None, bases=None,
                                              seed=DEFAULT_PROJECTION_SEED):
        return self._stacked(lambda m, x, bm, b, sd: x.project_columns_with_batch_correction(target_dim, block_size, bm, basis=b, seed=sd),
                             target_dim, batch_membership, bases, seed)

    def project_columns(self, target_dim, block_size=None, bases=None):
        return self.project_columns_with_batch_correction(target_dim, block_size, None, bases=bases)

    def partition_columns_to_groups(self, proj_kn, num_features=None, ncols_per_group=None):
        """random_projection.rs:311-340: one set of binary codes from the stacked projection, assigned to every modality"""
        n, K = proj_kn.shape
        if n != self.num_columns():
            raise LegumeError(1, "number of columns mismatch")
        first = self.stack[0]
        out = first.partition_columns_to_groups(proj_kn, num_features, ncols_per_group)
        for x in self.stack[1:]:
            x.binary_codes, x.col_to_group, x.group_keys = first.binary_codes, first.col_to_group, first.group_keys
        return out

    def collapse_columns_multilevel_vec(self, proj_kn, batch_membership, params: "MultilevelParams"):
        """MultilevelCollapsingOps for SparseIoStack (collapse_data/mod.rs:1050-1260, refine.rs:503-716): ONE partition for every
        modality — the first layer owns the grouping decision (its gene sums drive the refinement) — and per level one
        CollapsedOut per layer.  Returns (levels x layers of CollapsedOut, levels x layers of CollapsedStat).  Panels are not
        observed on the stack path (mod.rs:1128) and anchor / bulk batches are refused there (:1116-1120)."""
        import copy
        first = self.stack[0]
        proj_kn, nb, level_dims, codes_h, group, ng = first._multilevel_prologue(proj_kn, batch_membership, params)
        for x in self.stack[1:]:
            x.register_batch_membership(batch_membership)
            if nb >= 2:
                x.build_hnsw_per_batch(proj_kn, batch_membership)
            x.col_to_group, x.binary_codes, x.group_keys = first.col_to_group, first.binary_codes, first.group_keys
        p = copy.copy(params)
        p.observe_panels = False
        per_layer = []
        if params.refine is not None:
            layout, gs0 = first._build_pb_samples(proj_kn, group, ng, nb)
            p2g, k = first._refined_partition(proj_kn, nb, level_dims, codes_h, layout, gs0, p)
            for d, x in enumerate(self.stack):
                gs = gs0 if d == 0 else x._build_pb_samples(proj_kn, group, ng, nb)[1]
                out = x._collect_refined_levels(proj_kn, nb, layout, gs, p2g, k, p, p.output_calibration)
                per_layer.append((out["levels"], out["stats"]))
        else:
            for x in self.stack:
                per_layer.append(x._legacy_levels(proj_kn, nb, level_dims, codes_h, group, ng, p))
        nlev = len(per_layer[0][0])
        return ([[per_layer[d][0][l] for d in range(len(self.stack))] for l in range(nlev)],
                [[per_layer[d][1][l] for d in range(len(self.stack))] for l in range(nlev)])

    def collapse_columns_multilevel(self, proj_kn, batch_membership, params: "MultilevelParams"):
        """the finest level only (mod.rs:1053-1068)"""
        outs, _ = self.collapse_columns_multilevel_vec(proj_kn, batch_membership, params)
        if not outs:
            raise LegumeError(1, "no levels processed")
        return outs[0]

    def project_columns_weighted(self, target_dim, block_size, batch_membership, row_weights, bases=None,
                                 seed=DEFAULT_PROJECTION_SEED):
        """row_weights covers the stacked rows; every modality takes its own slice (random_projection.rs:253-330)"""
        w = np.ascontiguousarray(row_weights, np.float32)
        if len(w) != self.num_rows():
            raise LegumeError(1, "SparseIoStack: row weights must cover the stacked rows")
        offs = np.cumsum([0] + [x.num_rows() for x in self.stack])
        return self._stacked(lambda m, x, bm, b, sd: x.project_columns_weighted(target_dim, block_size, bm, w[offs[m]:offs[m + 1]], basis=b, seed=sd),
                             target_dim, batch_membership, bases, seed)


class ColumnDict:
    def __init__(self, ctx: Context, data, names):
        """data: (n, d) array, one point per row here = one column of the reference's DMatrix"""
        self.ctx = ctx
        self.data = _as(data, np.float32)
        self._names = list(names)
        if len(self._names) != self.data.shape[0]:
            raise LegumeError(1, "Data and names must have the same length")
        self.name2index = {k: i for i, k in enumerate(self._names)}

    @classmethod
    def from_dmatrix(cls, ctx, data, names):
        return cls(ctx, data, names)

    def names(self):
        return self._names

    def dim(self):
        return None if self.data.shape[0] == 0 else self.data.shape[1]

    def num_points(self):
        return self.data.shape[0]

    def _index_of(self, name):
        if name not in self.name2index:
            raise LegumeError(1, f"name {name} not found")
        return self.name2index[name]

    def search_indices(self, queries, knn, exclude=None):
        """batched core: queries (nq, d) -> (idx (nq, k'), dist (nq, k')) with k' = min(knn, available)"""
        queries = _as(queries, np.float32)
        nq = queries.shape[0]
        nr = self.num_points()
        dev = _is_torch(queries) or _is_torch(self.data)
        if knn == 0 or nr == 0 or nq == 0:
            return np.zeros((nq, 0), np.uint32), np.zeros((nq, 0), np.float32)
        ex = None if exclude is None else _as(exclude, np.uint32)
        idx = self.ctx.empty((nq, knn), np.uint32, dev)
        dist = self.ctx.empty((nq, knn), np.float32, dev)
        self.ctx.check(lib.lg_knn_topk(self.ctx.h, _ptr(self.data), nr, _ptr(queries), nq, self.data.shape[1], knn,
                                       _ptr(ex), _ptr(idx), _ptr(dist)))
        return idx, dist

    def _trim(self, idx, dist):
        keep = idx != np.uint32(0xFFFFFFFF)
        return [int(i) for i in idx[keep]], [float(x) for x in dist[keep]]

    def search_by_query_data(self, query, knn):
        query = np.asarray(query, np.float32)
        if (self.dim() or 0) != query.shape[0]:
            raise LegumeError(1, "query's dim does not match")
        idx, dist = self.search_indices(query[None, :], knn)
        ii, dd = self._trim(np.asarray(idx)[0], np.asarray(dist)[0])
        return [self._names[i] for i in ii], dd

    def search_by_query_name(self, query_name, knn, exclude_same):
        q = self._index_of(query_name)
        ex = np.array([q], np.uint32) if exclude_same else None
        idx, dist = self.search_indices(np.asarray(self.data)[q:q + 1], knn, ex)
        ii, dd = self._trim(np.asarray(idx)[0], np.asarray(dist)[0])
        return [self._names[i] for i in ii], dd

    def search_others(self, query_name, knn):
        return self.search_by_query_name(query_name, knn, True)

    def match_by_query_name_against(self, query_name, knn, against: "ColumnDict"):
        q = self._index_of(query_name)
        idx, dist = against.search_indices(np.asarray(self.data)[q:q + 1], knn)
        ii, dd = against._trim(np.asarray(idx)[0], np.asarray(dist)[0])
        return [against._names[i] for i in ii], dd
