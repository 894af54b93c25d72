import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, bench
data = bench.make_data(2, 50 + 5 + 100)
for resident in (True, False):
    r = bench.run_gpu(data, 0, 50, 5, 100, resident=resident, flush=True)
    ms = np.array(r["step_ms"]); wall = np.array(r["wall_ms"])
    print("resident", resident, "mean %.3f p50 %.3f max %.3f" % (ms.mean(), np.median(ms), ms.max()))
    idx = np.argsort(-ms)[:8]
    print("  worst steps", [(int(i), round(float(ms[i]), 3), round(float(wall[i]), 3)) for i in idx])
    print("  first 12", np.round(ms[:12], 3))
