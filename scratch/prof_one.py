"""One config-1-shaped registration (50-keyframe map, one VLP-16 scan, 10 LM iterations), repeated a few times."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from lego_loam_sr_b200 import synth, abi
WORLD = dict(n_blocks=60, n_poles=600, half_x=80, half_z=120)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
d = synth.make_single(seed=1, n_map_scans=50, step=1.0, world_kw=WORLD)
prm = abi.default_params(apply_update=1, eigen_mode=1, iter_count_thres=10, stop_thres=0.0)
with abi.Handle(prm) as h:
    for i, f in enumerate(d["keyframes"]):
        h.add_keyframe(i, f["corner_scan"], f["surf_last_ds"], f["outlier_last_ds"], d["key_poses"][i])
    for r in range(reps):
        h.build_local_map(list(range(len(d["keyframes"]))))
        pose, res = h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
    print("pose err", pose - d["gt"], "sel", res.sel_num, "maps", res.corner_map_ds, res.surf_map_ds, "launches", h.launch_count)
