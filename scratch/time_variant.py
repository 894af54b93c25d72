"""Per-kernel-class timing of one config-1-shaped registration with a chosen library build (LLB200_LIB)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from lego_loam_sr_b200 import synth, abi
if os.environ.get("LLB200_LIB"):
    abi.LIB_PATH = os.environ["LLB200_LIB"]
import bench
if os.environ.get("LLB200_R01_WORLD"):
    d = synth.make_single(seed=1, n_map_scans=50, step=bench.R01_STEP_M, world_kw=bench.R01_WORLD)
else:  # the benchmark's workload (~118 k-point map, ~10 k queries)
    d = synth.make_single(seed=1, n_map_scans=50, step=bench.BENCH_STEP_M, world_kw=bench.BENCH_WORLD)
if os.environ.get("LLB200_SMALL_PRIOR"):  # a prior as close as the replay's odometry gives (the default is 0.5 deg / 5 cm off)
    d["prior"] = synth.perturb(d["gt"], 0.05, 0.01, np.random.default_rng(1)).astype(np.float32)
prm = abi.default_params(apply_update=1, eigen_mode=1, iter_count_thres=10, stop_thres=0.0, lm_solver=1)
with abi.Handle(prm) as h:
    for i, f in enumerate(d["keyframes"]):
        h.add_keyframe(i, f["corner_scan"], f["surf_last_ds"], f["outlier_last_ds"], d["key_poses"][i])
    for r in range(13):
        if r == 3:
            h.profile_enable(True); h.profile_read()
        h.build_local_map(list(range(len(d["keyframes"]))))
        pose, res = h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
    prof = h.profile_read()
    print(os.environ.get("LLB200_LIB", "default"), "sel", res.sel_num, " ".join("%s=%.1fus" % (k, v[0] * 1e3 / max(v[1], 1)) for k, v in prof.items()))
