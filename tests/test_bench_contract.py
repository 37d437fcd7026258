"""bench.py's output contract, checked on CPU through the reference arm (the oracle port on a tiny sample): exactly ONE
line on stdout, JSON, with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_exactly_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-cells", "300"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cells/s" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]
    # the reference arm times per-cell stages and the fixed D x S fit separately and extrapolates to the GPU arm's config
    assert "same_config" in d["cpu_baseline"] and "posterior_DxS_s" in d["cpu_baseline"]["stage_seconds"]


def test_reference_arm_never_maps_the_product_library():
    """the CPU arm may execute oracle/ only: importing bench's reference leg must not load liblegume_b200.so"""
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--cpu-cells','200'];"
            "import runpy; runpy.run_path(os.path.join(%r, 'bench.py'), run_name='__main__');"
            "maps=open('/proc/self/maps').read(); assert 'liblegume_b200' not in maps, 'product library mapped';"
            "assert 'legume_b200' not in sys.modules; assert 'liblegume_oracle' in maps") % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]


def test_cpu_baseline_structure_matches_the_serial_oracle():
    """oracle_bench.cpp (block-wise projection with repack, group-wise collapse locked / lock-free, gene-blocked fit)
    returns exactly what the serial oracle does"""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import numpy as np

    import oracle as orc
    from util import random_csc
    rng = np.random.default_rng(12)
    D, N, K = 900, 1234, 50
    ip, ix, v = random_csc(rng, D, N, 0.05, empty_every=41)
    basis = rng.standard_normal((D, K)).astype(np.float32)
    assert orc.bench_project_blocks(ip, ix, v, basis, 0, 3).tobytes() == orc.project_raw(ip, ix, v, basis, 2).tobytes()
    assert orc.bench_project_blocks(ip, ix, v, basis, 37, 2).tobytes() == orc.project_raw(ip, ix, v, basis, 1).tobytes()
    grp = rng.integers(0, 17, N).astype(np.uint32)
    s, size = orc.collapse_basic(ip, ix, v, D, grp, 17)
    for locked in (True, False):
        s2, size2 = orc.bench_collapse_groups(ip, ix, v, D, grp, 17, locked, 3)
        assert np.array_equal(s, s2) and np.array_equal(size, size2)
    a, b = orc.optimize_single(s, size), orc.bench_optimize_single_mt(s, size, nthreads=3)
    assert all(a[k].tobytes() == b[k].tobytes() for k in a)
