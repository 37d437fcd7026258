"""End-to-end parity FROM COUNTS at BASELINE configs[0] (20k genes x 50k cells, ~5 % nnz, d = 50, 2^10 bins):
GPU and oracle each run counts + basis -> projection -> codes -> groups -> sums -> posterior independently.

  exact-order mode (lg_project_exact): the projection is bit-identical to the CPU path, hence zero flipped code bits,
      zero group mismatches, zero differing sums; posterior within 1e-5.
  throughput mode (lg_project, tensor cores): projection within 1e-5; a cell whose standardised V sits within that
      error of a column mean can flip a bit — the count is REPORTED (and bounded), not assumed to be zero.
"""
import numpy as np
import pytest

import legume_b200 as lg
from tools.e2e_parity import run_e2e_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return lg.Context(0)


@pytest.fixture(scope="module")
def report(ctx):
    rep = run_e2e_parity(ctx)  # configs[0]
    print("\ne2e parity (configs[0]):", rep)
    return rep


def test_configs0_exact_mode_is_bit_exact_from_counts(report):
    r = report["exact"]
    assert r["proj_bit_identical"], r
    assert r["codes_bits_flipped"] == 0 and r["groups_mismatch"] == 0, r
    assert r["num_groups"][0] == r["num_groups"][1]
    assert r["sums_mismatch"] == 0 and r["sizes_mismatch"] == 0, r
    assert r["posterior_mean_max_err"] <= 1e-5 and r["posterior_log_mean_max_err"] <= 1e-5, r


def test_configs0_fast_mode_projection_within_contract_and_mismatch_is_counted(report):
    r = report["fast"]
    assert r["proj_max_err"] <= 1e-5, r
    # one flipped bit moves one cell between two groups; the rate is a measurement, bounded here
    assert r["cells_with_flipped_code"] <= r["cells"] // 1000, r


def test_exact_mode_small_shapes_with_batches(ctx):
    """ragged / empty columns and three batches: exact projection == oracle bytes, groups equal"""
    import oracle as orc
    from util import random_csc
    rng = np.random.default_rng(7)
    D, N, K, kk = 700, 3000, 50, 8
    ip, ix, v = random_csc(rng, D, N, density=0.04, max_count=9, empty_every=97)
    basis = rng.standard_normal((D, K)).astype(np.float32)
    for nb in (1, 3):
        batch = rng.integers(0, nb, N).astype(np.uint32)
        want = orc.project(ip, ix, v, basis, batch, nb, nthreads=4)
        x = lg.SparseIoVec.from_csc(ctx, ip, ix, v, D)
        _, got = x.project_columns_with_batch_correction(K, None, [int(b) for b in batch], basis=basis, exact=True)
        assert got.tobytes() == want.tobytes(), (nb, np.max(np.abs(got - want)))
        codes = lg.binary_sort_columns(ctx, got, kk)
        assert np.array_equal(codes, orc.binary_codes(want, kk))


def test_batch_label_out_of_range_is_rejected(ctx):
    from util import random_csc
    rng = np.random.default_rng(3)
    ip, ix, v = random_csc(rng, 100, 64, density=0.1)
    blk = lg.CscBlock.upload(ctx, ip, ix, v, 100)
    basis = rng.standard_normal((100, 8)).astype(np.float32)
    proj = np.empty((64, 8), np.float32)
    bad = np.full(64, 2, np.uint32)
    from legume_b200._lib import lib
    rc = lib.lg_project(ctx.h, blk.h, basis.ctypes.data, 8, bad.ctypes.data, 2, proj.ctypes.data)
    assert rc == 1 and b"out of range" in lib.lg_last_error(ctx.h)
