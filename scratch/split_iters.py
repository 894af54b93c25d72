"""Where a benchmarked scan's time goes: the replay of bench.py (resident leg) with iter_count_thres = 1, 2, 4, 10.
T(1) = map assembly + search grids + the first LM iteration; (T(10) - T(2)) / 8 = one steady-state iteration.
LLB200_LIB selects a library build; NSTEPS the timed scans (default 40); NOFLUSH=1 keeps L2 warm."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from lego_loam_sr_b200 import abi
if os.environ.get("LLB200_LIB"):
    abi.LIB_PATH = os.environ["LLB200_LIB"]
import bench
steps = int(os.environ.get("NSTEPS", "40"))
flush = not os.environ.get("NOFLUSH")
if os.environ.get("LEGACY_WORLD"):  # round 1's smaller workload (~62 k-point map, ~6.7 k queries)
    bench._WORLD, bench._STEP = bench.R01_WORLD, bench.R01_STEP_M
data = bench.make_data(2, bench.WINDOW + 5 + steps)
out = {}
for it in [int(v) for v in os.environ.get("ITERS", "1,2,4,10").split(",")]:
    r = bench.run_gpu(data, 0, bench.WINDOW, 5, steps, mode="resident", flush=flush, params_kw=dict(iter_count_thres=it))
    out[it] = statistics.median(r["step_ms"]) * 1e3
    print("iter_count_thres=%2d  p50 %.1f us/scan  (mean %.1f)" % (it, out[it], 1e3 * sum(r["step_ms"]) / len(r["step_ms"])), flush=True)
if 10 in out and 2 in out:
    print("steady-state iteration ~ %.1f us; map + grids + first iteration ~ %.1f us" % ((out[10] - out[2]) / 8, out.get(1, float('nan'))))
