"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(legume-rs_b200/) never imports this module.

All dense matrices are column-major f32 like nalgebra's DMatrix; here they are
numpy arrays of shape (ncols, nrows) in C order, i.e. `proj[j]` is cell j's
K-vector.  Helper names follow the reference functions they restate.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblegume_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with its committed Makefile (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_adjust.cpp", "oracle_next.cpp", "oracle_svd.cpp", "oracle_bench.cpp", "oracle_refine.cpp", "oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None

_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.orc_digamma.restype = C.c_float
        L.orc_digamma.argtypes = [C.c_float]
        L.orc_trigamma.restype = C.c_float
        L.orc_trigamma.argtypes = [C.c_float]
        L.orc_l2_sq.restype = C.c_float
        L.orc_l2_sq.argtypes = [_f32p, _f32p, C.c_int]
        L.orc_assign_groups.restype = C.c_uint32
        L.orc_assign_groups_padded.restype = C.c_uint32
        L.orc_level_sort_dims.restype = C.c_int
        L.orc_binary_codes.restype = C.c_int
        L.orc_binary_codes_svd.restype = C.c_int
        L.orc_sim_poisson_csc.restype = C.c_uint64
        L.orc_pb_layout.restype = C.c_uint32
        L.orc_fine_to_coarse.restype = C.c_uint32
        for f in ("orc_smallrng_u64", "orc_sibling_sets", "orc_candidate_sets", "orc_dcp_refine_level", "orc_refine_assignments"):
            getattr(L, f).restype = C.c_uint64
        L.orc_smallrng_range_f64.restype = C.c_double
        L.orc_compact_labels.restype = C.c_uint32
        L.orc_project_to_refinement.restype = C.c_uint32
    return _lib


def _csc(indptr, indices, data):
    return (np.ascontiguousarray(indptr, np.uint64), np.ascontiguousarray(indices, np.uint64),
            np.ascontiguousarray(data, np.float32))


# ---- stage 1 -------------------------------------------------------------------------------
def project_raw(indptr, indices, data, basis_kd, nthreads=1):
    """project_columns_visitor (random_projection.rs:169-199). basis_kd: (D, K) array = K×D col-major."""
    indptr, indices, data = _csc(indptr, indices, data)
    basis_kd = np.ascontiguousarray(basis_kd, np.float32)
    n = len(indptr) - 1
    K = basis_kd.shape[1]
    out = np.empty((n, K), np.float32)
    lib().orc_project_raw(_ptr(indptr, C.c_uint64), _ptr(indices, C.c_uint64), _ptr(data, C.c_float),
                          C.c_uint64(n), _ptr(basis_kd, C.c_float), C.c_int(K), _ptr(out, C.c_float),
                          C.c_int(nthreads))
    return out


def project_finish(proj, batch=None, nbatch=0):
    """batch centring + standardise + clamp (random_projection.rs:378-407); returns a new array."""
    out = np.ascontiguousarray(proj, np.float32).copy()
    n, K = out.shape
    b = None if batch is None else np.ascontiguousarray(batch, np.uint32)
    lib().orc_project_finish(_ptr(out, C.c_float), C.c_int(K), C.c_uint64(n), _ptr(b, C.c_uint32),
                             C.c_uint32(nbatch))
    return out


def project(indptr, indices, data, basis_kd, batch=None, nbatch=0, nthreads=1):
    return project_finish(project_raw(indptr, indices, data, basis_kd, nthreads), batch, nbatch)


# ---- stage 2 -------------------------------------------------------------------------------
def binary_codes(proj, kk, details=False):
    """binary_sort_columns (random_projection.rs:535-564)."""
    proj = np.ascontiguousarray(proj, np.float32)
    n, K = proj.shape
    codes = np.zeros(n, np.uint64)
    q = np.zeros((kk, K), np.float32)
    u = np.zeros((kk, kk), np.float32)
    sig = np.zeros(kk, np.float32)
    mean = np.zeros(kk, np.float32)
    rc = lib().orc_binary_codes(_ptr(proj, C.c_float), C.c_int(K), C.c_uint64(n), C.c_int(kk),
                                _ptr(codes, C.c_uint64), _ptr(q, C.c_float), _ptr(u, C.c_float),
                                _ptr(sig, C.c_float), _ptr(mean, C.c_float))
    if rc != 0:
        raise ValueError("orc_binary_codes: bad arguments")
    return (codes, q, u, sig, mean) if details else codes


def binary_codes_svd(proj, kk, details=False):
    """the independent restatement (oracle_svd.cpp): f32 bidiagonalisation + implicit QR SVD, the reference's route.
    Bits are defined up to per-bit complement."""
    proj = np.ascontiguousarray(proj, np.float32)
    n, K = proj.shape
    codes = np.zeros(n, np.uint64)
    v = np.zeros((kk, n), np.float32)
    sig = np.zeros(kk, np.float32)
    rc = lib().orc_binary_codes_svd(_ptr(proj, C.c_float), C.c_int(K), C.c_uint64(n), C.c_int(kk), _ptr(codes, C.c_uint64),
                                    _ptr(v, C.c_float), _ptr(sig, C.c_float))
    if rc != 0:
        raise ValueError(f"orc_binary_codes_svd failed ({rc})")
    return (codes, v, sig) if details else codes


def partition_agreement(codes_a, codes_b, kk):
    """per bit: the fraction of cells two code sets put on the same side, up to complement of the bit"""
    a, b = np.asarray(codes_a, np.uint64), np.asarray(codes_b, np.uint64)
    out = []
    for k in range(kk):
        same = float(np.mean(((a >> np.uint64(k)) & np.uint64(1)) == ((b >> np.uint64(k)) & np.uint64(1))))
        out.append(max(same, 1.0 - same))
    return out


def householder_q(a):
    """a: (r, K) array = K×r col-major -> thin Q same shape."""
    a = np.ascontiguousarray(a, np.float32)
    r, K = a.shape
    q = np.zeros_like(a)
    lib().orc_householder_q(_ptr(a, C.c_float), C.c_int(K), C.c_int(r), _ptr(q, C.c_float))
    return q


def jacobi_eig(g):
    g = np.ascontiguousarray(g, np.float64)
    n = g.shape[0]
    ev = np.zeros(n)
    vec = np.zeros((n, n))
    lib().orc_jacobi_eig(_ptr(g, C.c_double), C.c_int(n), _ptr(ev, C.c_double), _ptr(vec, C.c_double))
    return ev, vec  # vec[k] = k-th eigenvector


# ---- stage 3 -------------------------------------------------------------------------------
def assign_groups(codes):
    codes = np.ascontiguousarray(codes, np.uint64)
    out = np.zeros(len(codes), np.uint32)
    ng = lib().orc_assign_groups(_ptr(codes, C.c_uint64), C.c_uint64(len(codes)), _ptr(out, C.c_uint32))
    return out, int(ng)


def assign_groups_padded(labels, k):
    labels = np.ascontiguousarray(labels, np.uint64)
    out = np.zeros(len(labels), np.uint32)
    ng = lib().orc_assign_groups_padded(_ptr(labels, C.c_uint64), C.c_uint64(len(labels)), C.c_uint64(k),
                                        _ptr(out, C.c_uint32))
    return out, int(ng)


def level_sort_dims(sort_dim, num_levels):
    out = (C.c_int * max(num_levels, 1))()
    n = lib().orc_level_sort_dims(C.c_int(sort_dim), C.c_int(num_levels), out)
    return [out[i] for i in range(n)]


# ---- stage 4 -------------------------------------------------------------------------------
def collapse_basic(indptr, indices, data, nrows, group_of_cell, S, mult=None):
    indptr, indices, data = _csc(indptr, indices, data)
    g = np.ascontiguousarray(group_of_cell, np.uint32)
    m = None if mult is None else np.ascontiguousarray(mult, np.float32)
    n = len(indptr) - 1
    sum_ds = np.zeros((S, nrows), np.float32)
    size_s = np.zeros(S, np.float32)
    lib().orc_collapse_basic(_ptr(indptr, C.c_uint64), _ptr(indices, C.c_uint64), _ptr(data, C.c_float),
                             C.c_uint64(nrows), C.c_uint64(n), _ptr(g, C.c_uint32), _ptr(m, C.c_float),
                             C.c_uint32(S), _ptr(sum_ds, C.c_float), _ptr(size_s, C.c_float))
    return sum_ds, size_s


def collapse_batch(indptr, indices, data, nrows, group_of_cell, batch_of_cell, S, B, mult=None):
    indptr, indices, data = _csc(indptr, indices, data)
    g = np.ascontiguousarray(group_of_cell, np.uint32)
    b = np.ascontiguousarray(batch_of_cell, np.uint32)
    m = None if mult is None else np.ascontiguousarray(mult, np.float32)
    n = len(indptr) - 1
    sum_db = np.zeros((B, nrows), np.float32)
    n_bs = np.zeros((S, B), np.float32)
    lib().orc_collapse_batch(_ptr(indptr, C.c_uint64), _ptr(indices, C.c_uint64), _ptr(data, C.c_float),
                             C.c_uint64(nrows), C.c_uint64(n), _ptr(g, C.c_uint32), _ptr(b, C.c_uint32),
                             _ptr(m, C.c_float), C.c_uint32(S), C.c_uint32(B), _ptr(sum_db, C.c_float),
                             _ptr(n_bs, C.c_float))
    return sum_db, n_bs


def merge_stat(fine_ds, fine_to_coarse, ncoarse):
    fine_ds = np.ascontiguousarray(fine_ds, np.float32)
    nfine, D = fine_ds.shape
    f2c = np.ascontiguousarray(fine_to_coarse, np.uint32)
    out = np.zeros((ncoarse, D), np.float32)
    lib().orc_merge_stat(_ptr(fine_ds, C.c_float), C.c_uint64(D), C.c_uint32(nfine), _ptr(f2c, C.c_uint32),
                         C.c_uint32(ncoarse), _ptr(out, C.c_float))
    return out


# ---- stage 5 -------------------------------------------------------------------------------
TARGET_ALL, TARGET_MEAN_ONLY, TARGET_MEAN_AND_LOG_MEAN = 0, 1, 2


def digamma(x):
    return float(lib().orc_digamma(C.c_float(x)))


def trigamma(x):
    return float(lib().orc_trigamma(C.c_float(x)))


def gamma_calibrate(num, den, a0=1.0, b0=1.0, target=TARGET_ALL):
    num = np.ascontiguousarray(num, np.float32)
    den = np.ascontiguousarray(den, np.float32)
    outs = [np.zeros_like(num) for _ in range(4)]
    lib().orc_gamma_calibrate(_ptr(num, C.c_float), _ptr(den, C.c_float), C.c_uint64(num.size), C.c_float(a0),
                              C.c_float(b0), C.c_int(target), *[_ptr(o, C.c_float) for o in outs])
    return dict(mean=outs[0], sd=outs[1], log_mean=outs[2], log_sd=outs[3])


def optimize_single(sum_ds, size_s, a0=1.0, b0=1.0, target=TARGET_ALL):
    sum_ds = np.ascontiguousarray(sum_ds, np.float32)
    size_s = np.ascontiguousarray(size_s, np.float32)
    S, D = sum_ds.shape
    outs = [np.zeros_like(sum_ds) for _ in range(4)]
    lib().orc_optimize_single(_ptr(sum_ds, C.c_float), _ptr(size_s, C.c_float), C.c_uint64(D), C.c_uint32(S),
                              C.c_float(a0), C.c_float(b0), C.c_int(target), *[_ptr(o, C.c_float) for o in outs])
    return dict(mean=outs[0], sd=outs[1], log_mean=outs[2], log_sd=outs[3])


def optimize_batched(obs, imp, res, size_s, obs_db, n_bs, a0=1.0, b0=1.0, num_iter=30, target=TARGET_ALL):
    obs = np.ascontiguousarray(obs, np.float32)
    imp = np.ascontiguousarray(imp, np.float32)
    res = np.ascontiguousarray(res, np.float32)
    size_s = np.ascontiguousarray(size_s, np.float32)
    obs_db = np.ascontiguousarray(obs_db, np.float32)
    n_bs = np.ascontiguousarray(n_bs, np.float32)
    S, D = obs.shape
    B = obs_db.shape[0]
    mu_obs, mu_adj, mu_res, gam, lm = [np.zeros_like(obs) for _ in range(5)]
    delta = np.zeros_like(obs_db)
    lib().orc_optimize_batched(_ptr(obs, C.c_float), _ptr(imp, C.c_float), _ptr(res, C.c_float),
                               _ptr(size_s, C.c_float), _ptr(obs_db, C.c_float), _ptr(n_bs, C.c_float),
                               C.c_uint64(D), C.c_uint32(S), C.c_uint32(B), C.c_float(a0), C.c_float(b0),
                               C.c_int(num_iter), C.c_int(target), _ptr(mu_obs, C.c_float),
                               _ptr(mu_adj, C.c_float), _ptr(mu_res, C.c_float), _ptr(gam, C.c_float),
                               _ptr(delta, C.c_float), _ptr(lm, C.c_float))
    return dict(mu_observed=mu_obs, mu_adjusted=mu_adj, mu_residual=mu_res, gamma=gam, delta=delta,
                mu_adjusted_log_mean=lm)


def optimize_single_obs(sum_ds, size_s, size_ds, a0=1.0, b0=1.0, target=TARGET_ALL):
    """optimize_block, B <= 1 arm, with per-(gene, sample) effective sizes (stats.rs:176-186, 351-368)"""
    sum_ds = np.ascontiguousarray(sum_ds, np.float32)
    size_s = np.ascontiguousarray(size_s, np.float32)
    size_ds = None if size_ds is None else np.ascontiguousarray(size_ds, np.float32)
    S, D = sum_ds.shape
    outs = [np.zeros_like(sum_ds) for _ in range(4)]
    lib().orc_optimize_single_obs(_ptr(sum_ds, C.c_float), _ptr(size_s, C.c_float),
                                  None if size_ds is None else _ptr(size_ds, C.c_float), C.c_uint64(D), C.c_uint32(S),
                                  C.c_float(a0), C.c_float(b0), C.c_int(target), *[_ptr(o, C.c_float) for o in outs])
    return dict(mean=outs[0], sd=outs[1], log_mean=outs[2], log_sd=outs[3])


def optimize_batched_obs(obs, imp, res, size_s, size_ds, obs_db, n_bs, mask_db, a0=1.0, b0=1.0, num_iter=30, target=TARGET_ALL):
    """optimize_block, B > 1 arm, with size_ds (S, D) and obs_mask_db (B, D), either may be None (stats.rs:176-204, 299-322)"""
    obs, imp, res = (np.ascontiguousarray(x, np.float32) for x in (obs, imp, res))
    size_s, obs_db, n_bs = (np.ascontiguousarray(x, np.float32) for x in (size_s, obs_db, n_bs))
    size_ds = None if size_ds is None else np.ascontiguousarray(size_ds, np.float32)
    mask_db = None if mask_db is None else np.ascontiguousarray(mask_db, np.float32)
    S, D = obs.shape
    B = obs_db.shape[0]
    mu_obs, mu_adj, mu_res, gam, lm = [np.zeros_like(obs) for _ in range(5)]
    delta = np.zeros_like(obs_db)
    fp = lambda x: None if x is None else _ptr(x, C.c_float)
    lib().orc_optimize_batched_obs(fp(obs), fp(imp), fp(res), fp(size_s), fp(size_ds), fp(obs_db), fp(n_bs), fp(mask_db),
                                   C.c_uint64(D), C.c_uint32(S), C.c_uint32(B), C.c_float(a0), C.c_float(b0),
                                   C.c_int(num_iter), C.c_int(target), fp(mu_obs), fp(mu_adj), fp(mu_res), fp(gam),
                                   fp(delta), fp(lm))
    return dict(mu_observed=mu_obs, mu_adjusted=mu_adj, mu_residual=mu_res, gamma=gam, delta=delta,
                mu_adjusted_log_mean=lm)


def attach_observability(coverage, source, group, batch, mult, S, B):
    """collapse_data/mod.rs:221-301 restated loop for loop.  coverage: (nsrc, D) bool; source / group / batch: per column.
    Returns (size_ds (S, D) f32, mask_db (B, D) f32 or None when no entry is zero)."""
    coverage = np.asarray(coverage, bool)
    nsrc, D = coverage.shape
    count_bs = np.zeros((nsrc, S), np.float32)
    used = np.zeros((nsrc, B), bool)
    for c in range(len(source)):
        s = int(group[c])
        if s < S:
            count_bs[source[c], s] = np.float32(count_bs[source[c], s] + np.float32(1.0 if mult is None else mult[c]))
        if batch is not None and int(batch[c]) < B:
            used[source[c], int(batch[c])] = True
    size_ds = np.zeros((S, D), np.float32)
    for src in range(nsrc):
        for g in np.nonzero(coverage[src])[0]:
            size_ds[:, g] = size_ds[:, g] + count_bs[src]
    mask = np.zeros((B, D), np.float32)
    for src in range(nsrc):
        for b in np.nonzero(used[src])[0]:
            mask[b, coverage[src]] = 1.0
    return size_ds, (mask if (mask == 0.0).any() else None)


# ---- refine.rs / dc_poisson.rs helpers of the refinement arm without refinement (B = 1, inherited partitions) ----------
def compact_labels(labels):
    """dc_poisson.rs:493-509: labels -> 0..k in order of first appearance"""
    lut, out = {}, []
    for g in labels:
        out.append(lut.setdefault(int(g), len(lut)))
    return np.array(out, np.uint32), len(lut)


def pb_sample_to_cells(cell_to_pb, num_pb):
    """pb_samples.rs:472-481"""
    out = [[] for _ in range(num_pb)]
    for c, p in enumerate(cell_to_pb):
        if int(p) != 0xFFFFFFFF:
            out[int(p)].append(c)
    return out


def initial_per_level_from_hash(fine_codes, pb_cells, level_dims):
    """refine.rs:68-88: every level's pb-sample -> group from the finest code of the pb-sample's first cell, masked"""
    out = []
    for d in level_dims:
        mask = 0xFFFFFFFFFFFFFFFF if d >= 64 else (1 << d) - 1
        out.append(compact_labels([int(fine_codes[cells[0]]) & mask for cells in pb_cells])[0])
    return out


def fine_to_coarse_from_refined(p2f, p2c, num_fine):
    """refine.rs:43-62: the coarse label of the first pb-sample of every fine group"""
    m = np.full(num_fine, 0xFFFFFFFF, np.uint32)
    for p in range(len(p2f)):
        if m[p2f[p]] == 0xFFFFFFFF:
            m[p2f[p]] = p2c[p]
    return m


def modal_group(cells, lvl):
    """collapse_data/mod.rs:823-841; among equally frequent labels the reference takes whichever its hash map yields last
    (unspecified) - the oracle fixes the smallest label"""
    if not cells:
        return 0
    cnt = {}
    for c in cells:
        cnt[int(lvl[c])] = cnt.get(int(lvl[c]), 0) + 1
    best = max(cnt.values())
    return min(g for g, n in cnt.items() if n == best)


# ---- stage 6 -------------------------------------------------------------------------------
def l2_sq(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().orc_l2_sq(a, b, C.c_int(len(a))))


def knn_topk(ref, qry, k, exclude=None, nthreads=1):
    """ref: (nr, d), qry: (nq, d). Returns (idx (nq,k) uint32, dist (nq,k) f32), nearest first."""
    ref = np.ascontiguousarray(ref, np.float32)
    qry = np.ascontiguousarray(qry, np.float32)
    nr, d = ref.shape
    nq = qry.shape[0]
    ex = None if exclude is None else np.ascontiguousarray(exclude, np.uint32)
    idx = np.zeros((nq, k), np.uint32)
    dist = np.zeros((nq, k), np.float32)
    lib().orc_knn_topk(_ptr(ref, C.c_float), C.c_uint64(nr), _ptr(qry, C.c_float), C.c_uint64(nq), C.c_int(d),
                       C.c_int(k), _ptr(ex, C.c_uint32), _ptr(idx, C.c_uint32), _ptr(dist, C.c_float),
                       C.c_int(nthreads))
    return idx, dist


# ---- stage 7: cross-batch neighbourhood adjustment ---------------------------------------------
def batch_proximity(proj, batch, B):
    """sort_batch_proximity (batch.rs:182-234): (order (B, B) uint32, centroids (B, K))"""
    proj = np.ascontiguousarray(proj, np.float32)
    n, K = proj.shape
    b = np.ascontiguousarray(batch, np.uint32)
    order = np.zeros((B, B), np.uint32)
    cen = np.zeros((B, K), np.float32)
    lib().orc_batch_proximity(_ptr(proj, C.c_float), C.c_int(K), C.c_uint64(n), _ptr(b, C.c_uint32), C.c_uint32(B),
                              _ptr(order, C.c_uint32), _ptr(cen, C.c_float))
    return order, cen


def knn_match_batches(proj, batch, B, knn, target_order=None, nthreads=0):
    """neighbouring_columns_triplets (matched.rs:173-260): (idx (N, nt*knn) uint32 global, dist)"""
    proj = np.ascontiguousarray(proj, np.float32)
    n, K = proj.shape
    b = np.ascontiguousarray(batch, np.uint32)
    to = None if target_order is None else np.ascontiguousarray(target_order, np.uint32)
    nt = B if to is None else to.shape[1]
    idx = np.zeros((n, nt * knn), np.uint32)
    dist = np.zeros((n, nt * knn), np.float32)
    lib().orc_knn_match_batches(_ptr(proj, C.c_float), C.c_int(K), C.c_uint64(n), _ptr(b, C.c_uint32), C.c_uint32(B),
                                C.c_int(knn), _ptr(to, C.c_uint32), C.c_uint32(nt), _ptr(idx, C.c_uint32),
                                _ptr(dist, C.c_float), C.c_int(nthreads))
    return idx, dist


def collect_matched_stat(indptr, indices, data, nrows, group_of_cell, S, matched_idx, matched_dist):
    """collect_matched_stat_visitor (stats.rs:26-108): (imputed_sum_ds (S, D), residual_sum_ds (S, D))"""
    indptr, indices, data = _csc(indptr, indices, data)
    g = np.ascontiguousarray(group_of_cell, np.uint32)
    mi = np.ascontiguousarray(matched_idx, np.uint32)
    md = np.ascontiguousarray(matched_dist, np.float32)
    n = len(indptr) - 1
    imp = np.zeros((S, nrows), np.float32)
    res = np.zeros((S, nrows), np.float32)
    lib().orc_collect_matched_stat(_ptr(indptr, C.c_uint64), _ptr(indices, C.c_uint64), _ptr(data, C.c_float),
                                   C.c_uint64(nrows), C.c_uint64(n), _ptr(g, C.c_uint32), C.c_uint32(S),
                                   _ptr(mi, C.c_uint32), _ptr(md, C.c_float), C.c_uint32(mi.shape[1]),
                                   _ptr(imp, C.c_float), _ptr(res, C.c_float))
    return imp, res


def pb_layout(proj, group_of_cell, S, batch_of_cell, B, mult=None):
    """build_pb_sample_layout (pb_samples.rs:94-219): dict(cell_to_pb, pb_group, pb_batch, pb_count, centroids)"""
    proj = np.ascontiguousarray(proj, np.float32)
    n, K = proj.shape
    g = np.ascontiguousarray(group_of_cell, np.uint32)
    b = np.ascontiguousarray(batch_of_cell, np.uint32)
    m = None if mult is None else np.ascontiguousarray(mult, np.float32)
    c2p = np.zeros(n, np.uint32)
    pg, pbt = np.zeros(S * B, np.uint32), np.zeros(S * B, np.uint32)
    cnt = np.zeros(S * B, np.float32)
    cen = np.zeros((S * B, K), np.float32)
    npb = int(lib().orc_pb_layout(_ptr(proj, C.c_float), C.c_int(K), C.c_uint64(n), _ptr(g, C.c_uint32), C.c_uint32(S),
                                  _ptr(b, C.c_uint32), C.c_uint32(B), _ptr(m, C.c_float), _ptr(c2p, C.c_uint32),
                                  _ptr(pg, C.c_uint32), _ptr(pbt, C.c_uint32), _ptr(cnt, C.c_float), _ptr(cen, C.c_float)))
    return dict(cell_to_pb=c2p, pb_group=pg[:npb].copy(), pb_batch=pbt[:npb].copy(), pb_count=cnt[:npb].copy(),
                centroids=cen[:npb].copy(), num_pb=npb)


def pb_match(proj, batch_of_cell, B, layout, knn, nthreads=0):
    """per_batch_sc_neighbors (pb_samples.rs:442-459): (matched_pb (npb, B*knn) uint32, dist)"""
    proj = np.ascontiguousarray(proj, np.float32)
    n, K = proj.shape
    b = np.ascontiguousarray(batch_of_cell, np.uint32)
    npb = layout["num_pb"]
    mp = np.zeros((npb, B * knn), np.uint32)
    md = np.zeros((npb, B * knn), np.float32)
    cen = np.ascontiguousarray(layout["centroids"], np.float32)
    lib().orc_pb_match(_ptr(proj, C.c_float), C.c_int(K), C.c_uint64(n), _ptr(b, C.c_uint32), C.c_uint32(B),
                       _ptr(layout["cell_to_pb"], C.c_uint32), _ptr(cen, C.c_float), _ptr(layout["pb_batch"], C.c_uint32),
                       C.c_uint32(npb), C.c_int(knn), _ptr(mp, C.c_uint32), _ptr(md, C.c_float), C.c_int(nthreads))
    return mp, md


def collect_matched_stat_coarse(gene_sums, pb_count, pb_to_group, S, matched_pb, matched_dist):
    """collect_matched_stat_coarse (stats.rs:698-784); gene_sums (npb, D) dense"""
    gs = np.ascontiguousarray(gene_sums, np.float32)
    npb, D = gs.shape
    cnt = np.ascontiguousarray(pb_count, np.float32)
    p2g = np.ascontiguousarray(pb_to_group, np.uint32)
    mp = np.ascontiguousarray(matched_pb, np.uint32)
    md = np.ascontiguousarray(matched_dist, np.float32)
    imp = np.zeros((S, D), np.float32)
    res = np.zeros((S, D), np.float32)
    lib().orc_collect_matched_stat_coarse(_ptr(gs, C.c_float), C.c_uint64(D), C.c_uint32(npb), _ptr(cnt, C.c_float),
                                          _ptr(p2g, C.c_uint32), C.c_uint32(S), _ptr(mp, C.c_uint32), _ptr(md, C.c_float),
                                          C.c_uint32(mp.shape[1]), _ptr(imp, C.c_float), _ptr(res, C.c_float))
    return imp, res


def fine_to_coarse(group_code, coarse_dim):
    """compute_fine_to_coarse_mapping (refine.rs:741-769): (fine_to_coarse uint32[nfine], num_coarse)"""
    gc = np.ascontiguousarray(group_code, np.uint64)
    out = np.zeros(len(gc), np.uint32)
    k = int(lib().orc_fine_to_coarse(_ptr(gc, C.c_uint64), C.c_uint32(len(gc)), C.c_int(coarse_dim), _ptr(out, C.c_uint32)))
    return out, k


# ---- synthetic counts ------------------------------------------------------------------------
def sim_poisson_csc(seed, D, col_lo, col_hi, topic_of_cell, batch_of_cell, ntopic, nbatch, lam, p0, npiece):
    """CPU twin of lg_sim_poisson_csc.  topic/batch arrays cover [col_lo, col_hi)."""
    n = col_hi - col_lo
    t = np.ascontiguousarray(topic_of_cell, np.uint8)
    b = np.ascontiguousarray(batch_of_cell, np.uint8)
    lam = np.ascontiguousarray(lam, np.float32)
    p0 = np.ascontiguousarray(p0, np.float32)
    npiece = np.ascontiguousarray(npiece, np.uint8)
    indptr = np.zeros(n + 1, np.uint64)
    args = [C.c_uint64(seed), C.c_uint64(D), C.c_uint64(col_lo), C.c_uint64(col_hi), _ptr(t, C.c_uint8),
            _ptr(b, C.c_uint8), C.c_uint32(ntopic), C.c_uint32(nbatch), _ptr(lam, C.c_float), _ptr(p0, C.c_float),
            _ptr(npiece, C.c_uint8), _ptr(indptr, C.c_uint64)]
    nnz = int(lib().orc_sim_poisson_csc(*args, None, None))
    indices = np.zeros(nnz, np.uint64)
    data = np.zeros(nnz, np.float32)
    lib().orc_sim_poisson_csc(*args, _ptr(indices, C.c_uint64), _ptr(data, C.c_float))
    return indptr, indices, data


# ---- the steps either side of the path -------------------------------------
def row_stats(indptr, indices, data, nrows):
    """SparseRunningStatistics::add_csc (matrix-util/src/sparse_stat.rs:64-108): (npos, s1, s2), f32, column order"""
    indptr, indices, data = _csc(indptr, indices, data)
    n = len(indptr) - 1
    npos, s1, s2 = (np.zeros(nrows, np.float32) for _ in range(3))
    lib().orc_row_stats(_ptr(indptr, C.c_uint64), _ptr(indices, C.c_uint64), _ptr(data, C.c_float), C.c_uint64(nrows),
                        C.c_uint64(n), _ptr(npos, C.c_float), _ptr(s1, C.c_float), _ptr(s2, C.c_float))
    return npos, s1, s2


def row_stats_moments(s1, s2, ncols_processed):
    """mean, variance, std (sparse_stat.rs:412-431)"""
    s1 = np.ascontiguousarray(s1, np.float32)
    s2 = np.ascontiguousarray(s2, np.float32)
    mean, var, sd = (np.zeros(len(s1), np.float32) for _ in range(3))
    lib().orc_row_stats_moments(_ptr(s1, C.c_float), _ptr(s2, C.c_float), C.c_uint64(len(s1)), C.c_uint64(ncols_processed),
                                _ptr(mean, C.c_float), _ptr(var, C.c_float), _ptr(sd, C.c_float))
    return mean, var, sd


def nystrom_project(indptr, indices, data, nrows, basis_dk, delta_dp=None, pb_of_cell=None, column_sum_norm=1e4):
    """nystrom_proj_visitor (senna/src/svd/fit.rs:433-466).  basis_dk: (K, D) array = D x K column-major;
    delta_dp: (P, D) array = D x P column-major or None.  Returns (N, K) = K x N column-major."""
    indptr, indices, data = _csc(indptr, indices, data)
    basis = np.ascontiguousarray(basis_dk, np.float32)
    K = basis.shape[0]
    n = len(indptr) - 1
    delta = None if delta_dp is None else np.ascontiguousarray(delta_dp, np.float32)
    pb = None if delta is None else np.ascontiguousarray(pb_of_cell, np.uint32)
    P = 0 if delta is None else delta.shape[0]
    out = np.zeros((n, K), np.float32)
    lib().orc_nystrom_project(_ptr(indptr, C.c_uint64), _ptr(indices, C.c_uint64), _ptr(data, C.c_float), C.c_uint64(nrows),
                              C.c_uint64(n), _ptr(basis, C.c_float), C.c_int(K), _ptr(delta, C.c_float), _ptr(pb, C.c_uint32),
                              C.c_uint32(P), C.c_float(column_sum_norm), _ptr(out, C.c_float))
    return out


# ---- the CPU baseline leg (oracle_bench.cpp) -------------------------------------------------
def bench_project_blocks(indptr, indices, data, basis_kd, block=0, nthreads=0):
    """visit_columns_by_block + project_columns_visitor, 100-cell blocks, per-block repack, Mutex copy-out"""
    ip, ix, v = _csc(indptr, indices, data)
    basis_kd = np.ascontiguousarray(basis_kd, np.float32)
    D, K = basis_kd.shape
    n = len(ip) - 1
    out = np.zeros((n, K), np.float32)
    lib().orc_bench_project_blocks(_ptr(ip, C.c_uint64), _ptr(ix, C.c_uint64), _ptr(v, C.c_float), C.c_uint64(D), C.c_uint64(n),
                                   _ptr(basis_kd, C.c_float), C.c_int(K), C.c_uint64(block), C.c_int(nthreads), _ptr(out, C.c_float))
    return out


def bench_collapse_groups(indptr, indices, data, nrows, group_of_cell, S, locked=True, nthreads=0):
    ip, ix, v = _csc(indptr, indices, data)
    grp = np.ascontiguousarray(group_of_cell, np.uint32)
    s = np.zeros((S, nrows), np.float32)
    size = np.zeros(S, np.float32)
    lib().orc_bench_collapse_groups(_ptr(ip, C.c_uint64), _ptr(ix, C.c_uint64), _ptr(v, C.c_float), C.c_uint64(nrows),
                                    C.c_uint64(len(ip) - 1), _ptr(grp, C.c_uint32), C.c_uint32(S), C.c_int(int(locked)),
                                    C.c_int(nthreads), _ptr(s, C.c_float), _ptr(size, C.c_float))
    return s, size


def bench_optimize_single_mt(sum_ds, size_s, a0=1.0, b0=1.0, target=TARGET_ALL, nthreads=0):
    sum_ds = np.ascontiguousarray(sum_ds, np.float32)
    size_s = np.ascontiguousarray(size_s, np.float32)
    S, D = sum_ds.shape
    outs = {k: np.empty_like(sum_ds) for k in ("mean", "sd", "log_mean", "log_sd")}
    lib().orc_bench_optimize_single_mt(_ptr(sum_ds, C.c_float), _ptr(size_s, C.c_float), C.c_uint64(D), C.c_uint32(S),
                                       C.c_float(a0), C.c_float(b0), C.c_int(target), C.c_int(nthreads),
                                       _ptr(outs["mean"], C.c_float), _ptr(outs["sd"], C.c_float),
                                       _ptr(outs["log_mean"], C.c_float), _ptr(outs["log_sd"], C.c_float))
    return outs


# ---- BBKNN + DC-Poisson refinement (oracle_refine.cpp) -----------------------
def _csr(sets):
    ptr = np.zeros(len(sets) + 1, np.uint32)
    ptr[1:] = np.cumsum([len(s) for s in sets])
    flat = np.fromiter((x for s in sets for x in s), np.uint32, int(ptr[-1]))
    return ptr, np.ascontiguousarray(flat)


def _sets(ptr, flat):
    return [flat[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]


def smallrng_u64(seed, skip=0):
    return int(lib().orc_smallrng_u64(C.c_uint64(seed), C.c_int(skip)))


def smallrng_range_f64(seed, lo, hi, skip=0):
    return float(lib().orc_smallrng_range_f64(C.c_uint64(seed), C.c_int(skip), C.c_double(lo), C.c_double(hi)))


def project_to_refinement(child, parent):
    """refine_multilevel.rs:315-320"""
    c, p = np.ascontiguousarray(child, np.uint32), np.ascontiguousarray(parent, np.uint32)
    out = np.empty(len(c), np.uint32)
    k = lib().orc_project_to_refinement(_ptr(c, C.c_uint32), _ptr(p, C.c_uint32), C.c_uint64(len(c)), _ptr(out, C.c_uint32))
    return out, int(k)


def child_offset_within_parent(child, parent):
    """refine_multilevel.rs:333-345"""
    c, p = np.ascontiguousarray(child, np.uint32), np.ascontiguousarray(parent, np.uint32)
    out = np.empty(len(c), np.uint32)
    lib().orc_child_offset_within_parent(_ptr(c, C.c_uint32), _ptr(p, C.c_uint32), C.c_uint64(len(c)), _ptr(out, C.c_uint32))
    return out


def sibling_sets(level, parent, k):
    """dc_poisson.rs:518-550; parent=None at the coarsest level"""
    lv = np.ascontiguousarray(level, np.uint32)
    pa = None if parent is None else np.ascontiguousarray(parent, np.uint32)
    E = len(lv)
    ptr = np.empty(E + 1, np.uint32)
    n = lib().orc_sibling_sets(_ptr(lv, C.c_uint32), _ptr(pa, C.c_uint32), C.c_uint64(E), C.c_uint32(k), _ptr(ptr, C.c_uint32), None, C.c_uint64(0))
    flat = np.empty(max(int(n), 1), np.uint32)
    lib().orc_sibling_sets(_ptr(lv, C.c_uint32), _ptr(pa, C.c_uint32), C.c_uint64(E), C.c_uint32(k), _ptr(ptr, C.c_uint32), _ptr(flat, C.c_uint32),
                           C.c_uint64(len(flat)))
    return _sets(ptr, flat)


def candidate_sets(siblings, bbknn, labels):
    """refine_multilevel.rs:85-112"""
    sp, sf = _csr(siblings)
    bp, bf = _csr(bbknn)
    lb = np.ascontiguousarray(labels, np.uint32)
    E = len(lb)
    ptr = np.empty(E + 1, np.uint32)
    args = (_ptr(sp, C.c_uint32), _ptr(sf, C.c_uint32), _ptr(bp, C.c_uint32), _ptr(bf, C.c_uint32), _ptr(lb, C.c_uint32), C.c_uint64(E),
            _ptr(ptr, C.c_uint32))
    n = lib().orc_candidate_sets(*args, None, C.c_uint64(0))
    flat = np.empty(max(int(n), 1), np.uint32)
    lib().orc_candidate_sets(*args, _ptr(flat, C.c_uint32), C.c_uint64(len(flat)))
    return _sets(ptr, flat)


def dcp_fisher_weights(profiles):
    """Profiles::nb_fisher_weights (dc_poisson.rs:230-295) of a dense entity x feature matrix"""
    P = np.ascontiguousarray(profiles, np.float32)
    w = np.empty(P.shape[1], np.float32)
    lib().orc_dcp_fisher_weights(_ptr(P, C.c_float), C.c_uint32(P.shape[0]), C.c_uint64(P.shape[1]), _ptr(w, C.c_float))
    return w


def dcp_profiles(gene_sums, weights=None):
    """(weighted profile values as a dense matrix, size factors): from_gene_sums + weight_by_vec"""
    P = np.ascontiguousarray(gene_sums, np.float32).copy()
    w = None if weights is None else np.ascontiguousarray(weights, np.float32)
    sf = np.empty(P.shape[0], np.float32)
    lib().orc_dcp_profiles(_ptr(P, C.c_float), C.c_uint32(P.shape[0]), C.c_uint64(P.shape[1]), _ptr(w, C.c_float), _ptr(sf, C.c_float))
    return P, sf


def dcp_stats(profiles, k, labels, moves=()):
    """DcPoissonStats::from_profiles followed by delta_move for every (entity, to) of `moves`"""
    P = np.ascontiguousarray(profiles, np.float32)
    E, M = P.shape
    lb = np.ascontiguousarray(labels, np.uint32)
    me = np.ascontiguousarray([m[0] for m in moves], np.uint32)
    mt = np.ascontiguousarray([m[1] for m in moves], np.uint32)
    gs, lg = np.empty((k, M), np.float64), np.empty((k, M), np.float32)
    ss, lso, mem = np.empty(k, np.float64), np.empty(k, np.float32), np.empty(E, np.uint32)
    lib().orc_dcp_stats(_ptr(P, C.c_float), C.c_uint32(E), C.c_uint64(M), C.c_uint32(k), _ptr(lb, C.c_uint32), _ptr(me, C.c_uint32),
                        _ptr(mt, C.c_uint32), C.c_uint64(len(me)), _ptr(gs, C.c_double), _ptr(lg, C.c_float), _ptr(ss, C.c_double),
                        _ptr(lso, C.c_float), _ptr(mem, C.c_uint32))
    return dict(gene_sum=gs, log_gene=lg, size_sum=ss, log_size_offset=lso, membership=mem)


def dcp_scores(profiles, k, labels, e):
    P = np.ascontiguousarray(profiles, np.float32)
    lb = np.ascontiguousarray(labels, np.uint32)
    out = np.empty(k, np.float64)
    lib().orc_dcp_scores(_ptr(P, C.c_float), C.c_uint32(P.shape[0]), C.c_uint64(P.shape[1]), C.c_uint32(k), _ptr(lb, C.c_uint32), C.c_uint32(e),
                         _ptr(out, C.c_double))
    return out


def dcp_refine_level(profiles, candidates, k, labels, num_gibbs, num_greedy, jacobi_base_seed, stagnation=0.005):
    """refine_with_candidates_guarded (dc_poisson.rs:778-915), Jacobi sweeps: (labels, moves)"""
    P = np.ascontiguousarray(profiles, np.float32)
    cp, cf = _csr(candidates)
    lb = np.ascontiguousarray(labels, np.uint32).copy()
    moves = lib().orc_dcp_refine_level(_ptr(P, C.c_float), C.c_uint32(P.shape[0]), C.c_uint64(P.shape[1]), _ptr(cp, C.c_uint32),
                                       _ptr(cf, C.c_uint32), C.c_uint32(k), C.c_int(num_gibbs), C.c_int(num_greedy),
                                       C.c_uint64(jacobi_base_seed), C.c_double(stagnation), _ptr(lb, C.c_uint32))
    return lb, int(moves)


def refine_assignments(gene_sums, bbknn, initial_per_level, reproject_offsets=None, num_gibbs=20, num_greedy=10, fisher=True, seed=42,
                       stagnation=0.005):
    """refine_multilevel.rs:170-298: (pbsamp_to_group per level finest first, num_groups per level, moves)"""
    P = np.ascontiguousarray(gene_sums, np.float32)
    E, M = P.shape
    bp, bf = _csr(bbknn)
    init = np.ascontiguousarray(np.stack([np.asarray(l, np.uint32) for l in initial_per_level]), np.uint32)
    L = init.shape[0]
    off = None
    if reproject_offsets is not None:
        off = np.zeros((L, E), np.uint32)
        for l, o in enumerate(reproject_offsets):
            if len(o):
                off[l] = np.asarray(o, np.uint32)
    out, ks = np.empty((L, E), np.uint32), np.empty(L, np.uint32)
    moves = lib().orc_refine_assignments(_ptr(P, C.c_float), C.c_uint32(E), C.c_uint64(M), _ptr(bp, C.c_uint32), _ptr(bf, C.c_uint32), C.c_int(L),
                                         _ptr(init, C.c_uint32), _ptr(off, C.c_uint32), C.c_int(num_gibbs), C.c_int(num_greedy),
                                         C.c_int(1 if fisher else 0), C.c_uint64(seed), C.c_double(stagnation), _ptr(out, C.c_uint32),
                                         _ptr(ks, C.c_uint32))
    return [out[l].copy() for l in range(L)], [int(x) for x in ks], int(moves)
