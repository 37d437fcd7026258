"""Generates tests/golden/hotpath_small.npz: one small seeded input set and the oracle's outputs for every stage
of the hot path.  The reference is Rust and cannot run here, so these vectors pin the ORACLE
(a later edit to oracle/ that changes any result fails tests/test_golden.py) and give the CUDA path a fixed target
that does not depend on the oracle being rebuilt on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle as orc  # noqa: E402
from util import nystrom_f64, random_csc  # noqa: E402


def build():
    D, N, K, kk, B, knn = 120, 400, 12, 6, 3, 4
    rng = np.random.default_rng(20240517)
    ip, ix, v = random_csc(rng, D, N, density=0.12, empty_every=57)
    basis = rng.standard_normal((D, K)).astype(np.float32)
    batch = rng.integers(0, B, N).astype(np.uint32)
    g = dict(D=D, N=N, K=K, kk=kk, B=B, knn=knn, indptr=ip, indices=ix, data=v, basis=basis, batch=batch)
    g["proj"] = orc.project(ip, ix, v, basis, batch, B)
    g["codes"] = orc.binary_codes(g["proj"], kk)
    grp, S = orc.assign_groups(g["codes"])
    g["group"], g["S"] = grp, S
    g["sum_ds"], g["size_s"] = orc.collapse_basic(ip, ix, v, D, grp, S)
    g["sum_db"], g["n_bs"] = orc.collapse_batch(ip, ix, v, D, grp, batch, S, B)
    post = orc.optimize_single(g["sum_ds"], g["size_s"], 1.0, 1.0, 0)
    for key in ("mean", "sd", "log_mean", "log_sd"):
        g["post_" + key] = post[key]
    g["knn_idx"], g["knn_dist"] = orc.knn_topk(g["proj"][:300], g["proj"][300:], 5)
    g["prox_order"], g["prox_centroids"] = orc.batch_proximity(g["proj"], batch, B)
    g["matched_idx"], g["matched_dist"] = orc.knn_match_batches(g["proj"], batch, B, knn, g["prox_order"])
    g["imputed_cell"], g["residual_cell"] = orc.collect_matched_stat(ip, ix, v, D, grp, S, g["matched_idx"], g["matched_dist"])
    lay = orc.pb_layout(g["proj"], grp, S, batch, B)
    g["npb"] = lay["num_pb"]
    for key in ("cell_to_pb", "pb_group", "pb_batch", "pb_count", "centroids"):
        g["pb_" + key if not key.startswith("pb_") else key] = lay[key]
    g["pb_gene_sums"], _ = orc.collapse_basic(ip, ix, v, D, lay["cell_to_pb"], lay["num_pb"])
    g["pb_matched"], g["pb_matched_dist"] = orc.pb_match(g["proj"], batch, B, lay, knn)
    g["imputed_pb"], g["residual_pb"] = orc.collect_matched_stat_coarse(g["pb_gene_sums"], lay["pb_count"], lay["pb_group"], S,
                                                                          g["pb_matched"], g["pb_matched_dist"])
    bat = orc.optimize_batched(g["sum_ds"], g["imputed_pb"], g["residual_pb"], g["size_s"], g["sum_db"], g["n_bs"], 1.0, 1.0, 12, 0)
    g["mu_adjusted"], g["delta"] = bat["mu_adjusted"], bat["delta"]
    gcode = np.zeros(S, np.uint64)
    gcode[grp] = g["codes"]
    g["f2c_dim4"], g["ncoarse_dim4"] = orc.fine_to_coarse(gcode, 4)
    return g


def build_next():
    """tests/golden/next_small.npz: the steps either side of the path on a small seeded input — per-gene
    running statistics (oracle = reference f32 folds) and the Nystrom re-projection (the oracle's f32 result AND the float64
    restatement the CUDA path is held to)"""
    D, N, K, P = 150, 500, 10, 6
    rng = np.random.default_rng(20241018)
    ip, ix, v = random_csc(rng, D, N, density=0.15, empty_every=61)
    basis_dk = (rng.standard_normal((K, D)) * (10.0 ** rng.uniform(-2, 1, K))[:, None]).astype(np.float32)
    delta = np.exp(0.4 * rng.standard_normal((P, D))).astype(np.float32)
    delta[:, ::11] = 0.0
    pb = rng.integers(0, P, N).astype(np.uint32)
    g = dict(D=D, N=N, K=K, P=P, indptr=ip, indices=ix, data=v, basis_dk=basis_dk, delta_dp=delta, pb=pb)
    g["npos"], g["s1"], g["s2"] = orc.row_stats(ip, ix, v, D)
    g["mean"], g["variance"], g["sd"] = orc.row_stats_moments(g["s1"], g["s2"], N)
    g["nystrom_oracle"] = orc.nystrom_project(ip, ix, v, D, basis_dk, None, None, 1e4)
    g["nystrom_oracle_delta"] = orc.nystrom_project(ip, ix, v, D, basis_dk, delta, pb, 1e4)
    g["nystrom_exact"] = nystrom_f64(ip, ix, v, D, basis_dk, None, None, 1e4)
    g["nystrom_exact_delta"] = nystrom_f64(ip, ix, v, D, basis_dk, delta, pb, 1e4)
    return g


if __name__ == "__main__":
    gn = build_next()
    outn = os.path.join(HERE, "next_small.npz")
    np.savez_compressed(outn, **gn)
    print(outn, os.path.getsize(outn), "bytes;", len(gn), "arrays")
    g = build()
    out = os.path.join(HERE, "hotpath_small.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes;", len(g), "arrays; groups", g["S"], "pb-samples", g["npb"])
