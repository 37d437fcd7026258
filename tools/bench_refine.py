"""configs[2] (1M cells, 8 batches) through the pb-sample arm with the DC-Poisson refinement of the partition timed beside
the other stages (tools/bench_extras.c3_adjust).  One JSON line.  python tools/bench_refine.py [cells=1000000] [batches=8]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legume-rs_b200")]
import legume_b200 as lg
from legume_b200.pipeline import HotPath
from tools import bench_extras
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = lg.Context(0)
print(json.dumps(bench_extras.c3_adjust(ctx, HotPath(ctx), N, B, per_cell=False, do_refine=True)))
