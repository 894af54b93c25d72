"""Where the e2e leg's extra time goes: resident vs host clouds (pcl layout, 32-byte records) vs host clouds already packed
(16-byte records: the pack becomes a memcpy). p50 of the CUDA-event time, the wall clock and the C++ mirror's own frame time."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
steps = int(os.environ.get("NSTEPS", "100"))
data = bench.make_data(2, bench.WINDOW + 5 + steps)
for mode in ("resident", "host"):
    r = bench.run_gpu(data, 0, bench.WINDOW, 5, steps, mode=mode, flush=True)
    print("%-9s device %.1f us  wall %.1f us  mirror frame %.1f us" % (mode, 1e3 * statistics.median(r["step_ms"]), 1e3 * statistics.median(r["wall_ms"]),
                                                                      1e3 * statistics.median(r["frame_ms"])), flush=True)
# host clouds already packed
orig = bench.to_pcl_layout
bench.to_pcl_layout = lambda a: np.ascontiguousarray(a, np.float32)
r = bench.run_gpu(data, 0, bench.WINDOW, 5, steps, mode="host", flush=True)
print("%-9s device %.1f us  wall %.1f us  mirror frame %.1f us" % ("host16", 1e3 * statistics.median(r["step_ms"]), 1e3 * statistics.median(r["wall_ms"]),
                                                                  1e3 * statistics.median(r["frame_ms"])), flush=True)
