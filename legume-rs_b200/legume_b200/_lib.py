"""ctypes binding of liblegume_b200.so (include/legume_b200.h).

There is no fallback: if the CUDA library has not been built, importing fails loudly, and
creating a Context without a CUDA device raises.  Nothing here touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG), "liblegume_b200.so")

LG_OK = 0
TARGET_ALL, TARGET_MEAN_ONLY, TARGET_MEAN_AND_LOG_MEAN = 0, 1, 2
BLOCK_CELLS = 1024


class LegumeError(RuntimeError):
    """Raised for any non-zero status from the C ABI (the reference returns anyhow::Error)."""

    def __init__(self, code, msg):
        super().__init__(f"legume_b200 error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C legume-rs_b200/csrc` (or __graft_entry__.build()). "
        "legume_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _u64, _u32, _i, _f = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
_sig = {
    "lg_ctx_create": [_i, C.POINTER(_vp)],
    "lg_ctx_destroy": [_vp],
    "lg_ctx_set_stream": [_vp, _vp],
    "lg_ctx_sync": [_vp],
    "lg_csc_upload": [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp, C.POINTER(_vp)],
    "lg_csc_upload_remap": [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp, _u64, C.POINTER(_vp)],
    "lg_csc_concat": [_vp, _vp, _u32, C.POINTER(_vp)],
    "lg_csc_wrap_device": [_vp, _vp, _vp, _vp, _u64, _u64, _u64, C.POINTER(_vp)],
    "lg_csc_free": [_vp, _vp],
    "lg_csc_keep_pattern": [_vp, _vp, _i],
    "lg_csc_shape": [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)],
    "lg_csc_device_arrays": [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)],
    "lg_csc_download": [_vp, _vp, _vp, _vp, _vp],
    "lg_project": [_vp, _vp, _vp, _i, _vp, _u32, _vp],
    "lg_project_raw": [_vp, _vp, _vp, _i, _vp],
    "lg_proj_batch_partials": [_vp, _vp, _i, _u64, _vp, _u32, _vp],
    "lg_block_partials_finalize": [_vp, _vp, _u64, _u32, _vp],
    "lg_proj_centre_scale": [_vp, _vp, _i, _u64, _vp, _u32, _vp, _vp],
    "lg_proj_clamp_rescale": [_vp, _vp, _i, _u64],
    "lg_project_exact": [_vp, _vp, _vp, _i, _vp, _u32, _vp],
    "lg_project_raw_exact": [_vp, _vp, _vp, _i, _vp],
    "lg_proj_batch_fold": [_vp, _vp, _i, _u64, _vp, _u32, _vp, _vp],
    "lg_proj_centre_scale_exact": [_vp, _vp, _i, _u64, _vp, _u32, _vp, _vp, _vp],
    "lg_binary_codes": [_vp, _vp, _i, _u64, _i, _vp],
    "lg_codes_basis": [_vp, _vp, _i, _i, _i, _vp],
    "lg_codes_gram": [_vp, _vp, _i, _u64, _vp, _i, _vp, _vp],
    "lg_codes_factor": [_vp, _vp, _vp, _i, _i, _vp, _vp],
    "lg_codes_vproj": [_vp, _vp, _i, _u64, _vp, _vp, _vp, _vp],
    "lg_codes_means": [_vp, _vp, _i, _u64, _vp],
    "lg_codes_pack": [_vp, _vp, _i, _u64, _vp, _vp],
    "lg_assign_groups": [_vp, _vp, _u64, _i, _i, _vp, C.POINTER(_u32)],
    "lg_code_presence": [_vp, _vp, _u64, _i, _vp],
    "lg_group_lut": [_vp, _vp, _i, _i, _vp, C.POINTER(_u32)],
    "lg_codes_to_groups": [_vp, _vp, _u64, _i, _vp, _vp],
    "lg_collapse_basic": [_vp, _vp, _vp, _vp, _u32, _vp, _vp],
    "lg_collapse_batch": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _vp, _vp],
    "lg_merge_stat": [_vp, _vp, _u64, _u32, _vp, _u32, _vp],
    "lg_gamma_calibrate": [_vp, _vp, _vp, _u64, _f, _f, _i, _vp, _vp, _vp, _vp],
    "lg_optimize_single": [_vp, _vp, _vp, _u64, _u32, _f, _f, _i, _vp, _vp, _vp, _vp],
    "lg_optimize_single_obs": [_vp, _vp, _vp, _vp, _u64, _u32, _f, _f, _i, _vp, _vp, _vp, _vp],
    "lg_optimize_batched_obs": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _f, _f, _i, _i, _vp, _vp, _vp,
                                _vp, _vp, _vp],
    "lg_attach_observability": [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _u64, _u64, _u32, _u32, _vp, _vp, _vp],
    "lg_proj_clamp_rescale_if": [_vp, _vp, _i, _u64, _vp],
    "lg_comm_unique_id": [_vp, _vp],
    "lg_comm_init": [_vp, _vp, _i, _i],
    "lg_comm_info": [_vp, _vp, _vp],
    "lg_comm_destroy": [_vp],
    "lg_allreduce_stats": [_vp, _vp, _vp, _vp, _vp, _u64, _u32, _u32],
    "lg_hotpath_run_sharded": [_vp, _vp, _vp, _i, _vp, _u32, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "lg_optimize_batched": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _f, _f, _i, _i, _vp, _vp, _vp, _vp,
                            _vp, _vp],
    "lg_knn_topk": [_vp, _vp, _u64, _vp, _u64, _i, _i, _vp, _vp, _vp],
    "lg_knn_topk_sq": [_vp, _vp, _u64, _vp, _u64, _i, _i, _vp, _vp, _vp],
    "lg_knn_merge_topk": [_vp, _vp, _vp, _u32, _u64, _i, _vp, _vp, _vp, _vp],
    "lg_batch_proximity": [_vp, _vp, _i, _u64, _vp, _u32, _vp, _vp],
    "lg_knn_match_batches": [_vp, _vp, _i, _u64, _vp, _u32, _i, _vp, _u32, _vp, _vp],
    "lg_collect_matched_stat": [_vp, _vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp],
    "lg_pb_layout": [_vp, _vp, _i, _u64, _vp, _u32, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_u32)],
    "lg_pb_match": [_vp, _vp, _i, _u64, _vp, _u32, _vp, _vp, _vp, _u32, _i, _vp, _vp],
    "lg_collect_matched_stat_coarse": [_vp, _vp, _u64, _u32, _vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp],
    "lg_pair_presence": [_vp, _vp, _vp, _u64, _u32, _u32, _vp],
    "lg_pb_ids": [_vp, _vp, _u32, _u32, _vp, _vp, _vp, C.POINTER(_u32)],
    "lg_cells_to_pb": [_vp, _vp, _vp, _u64, _u32, _u32, _vp, _vp],
    "lg_pb_centroid_fold": [_vp, _vp, _i, _u64, _vp, _u32, _vp, _vp, _vp],
    "lg_pb_centroid_finish": [_vp, _vp, _vp, _u32, _i, _vp],
    "lg_pb_min_keys": [_vp, _vp, _i, _u64, _vp, _u64, _vp, _vp, _u32, _u32, _u32, _vp],
    "lg_pb_topk_keys": [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _i, _vp, _vp],
    "lg_fine_to_coarse": [_vp, _vp, _vp, _u64, _u32, _i, _vp, C.POINTER(_u32)],
    "lg_row_stats": [_vp, _vp, _vp, _vp, _vp],
    "lg_nystrom_project": [_vp, _vp, _vp, _i, _vp, _vp, _u32, _f, _vp],
    "lg_dcp_fisher_weights": [_vp, _vp, _u64, _u32, _vp],
    "lg_dcp_profiles": [_vp, _vp, _u64, _u32, _vp, _vp],
    "lg_dcp_refine_level": [_vp, _vp, _vp, _u64, _u32, _vp, _vp, _u32, _i, _i, _u64, C.c_double, _vp, C.POINTER(_u64)],
    "lg_zarr_open": [C.c_char_p, C.POINTER(_vp), C.c_char_p, C.c_size_t],
    "lg_zarr_shape": [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)],
    "lg_zarr_column_extent": [_vp, _u64, _u64, C.POINTER(_u64), C.POINTER(_u64)],
    "lg_zarr_read_columns_host": [_vp, _u64, _u64, _vp, _vp, _vp],
    "lg_zarr_read_columns": [_vp, _vp, _u64, _u64, C.POINTER(_vp)],
    "lg_sim_poisson_csc": [_vp, _u64, _u64, _u64, _u64, _vp, _vp, _u32, _u32, _vp, _vp, _vp, C.POINTER(_vp)],
}
for _name, _args in _sig.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = C.c_int
lib.lg_last_error.argtypes = [_vp]
lib.lg_last_error.restype = C.c_char_p
lib.lg_ctx_launch_count.argtypes = [_vp]
lib.lg_ctx_launch_count.restype = C.c_uint64
lib.lg_ctx_h2d_bytes.argtypes = [_vp]
lib.lg_ctx_h2d_bytes.restype = C.c_uint64
lib.lg_ctx_fallback_count.argtypes = [_vp]
lib.lg_ctx_fallback_count.restype = C.c_uint64
lib.lg_ctx_pattern_collapse_count.argtypes = [_vp]
lib.lg_ctx_pattern_collapse_count.restype = C.c_uint64
lib.lg_ctx_time_stages.argtypes = [_vp, C.c_int]
lib.lg_ctx_time_stages.restype = None
lib.lg_hotpath_last_stage_ms.argtypes = [_vp, _vp]
lib.lg_hotpath_last_stage_ms.restype = C.c_int
lib.lg_ctx_last_fallback.argtypes = [_vp]
lib.lg_ctx_last_fallback.restype = C.c_char_p
lib.lg_zarr_close.argtypes = [_vp]
lib.lg_zarr_close.restype = None
lib.lg_zarr_last_error.argtypes = [_vp]
lib.lg_zarr_last_error.restype = C.c_char_p
lib.lg_version.argtypes = []
lib.lg_version.restype = C.c_char_p

EXPORTED = sorted(list(_sig) + ["lg_last_error", "lg_ctx_launch_count", "lg_ctx_h2d_bytes", "lg_ctx_fallback_count", "lg_ctx_last_fallback",
                          "lg_ctx_pattern_collapse_count", "lg_ctx_time_stages", "lg_hotpath_last_stage_ms", "lg_version", "lg_zarr_close", "lg_zarr_last_error"])
