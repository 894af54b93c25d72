"""configs[2] / configs[3] records of bench.py alone (no CPU leg): e2e ms per scan and the per-kernel-class table."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
for kind in sys.argv[1:] or ["c2", "c3"]:
    r = bench.run_single_config(kind, 0, peaks, with_cpu=False)
    print(kind, json.dumps({k: r[k] for k in r if k != "kernels"}))
    print("   ", json.dumps({k: v["us_per_launch"] for k, v in r["kernels"].items()}))
