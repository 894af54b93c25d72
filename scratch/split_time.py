import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from lego_loam_sr_b200 import synth, abi
WORLD = dict(n_blocks=60, n_poles=600, half_x=80, half_z=120)
d = synth.make_single(seed=1, n_map_scans=50, step=1.0, world_kw=WORLD)
prm = abi.default_params(apply_update=1, eigen_mode=1, iter_count_thres=10, stop_thres=0.0, lm_solver=1)
with abi.Handle(prm) as h:
    h._lib.ll_reserve(h._h, 1 << 20, 1 << 21, 1 << 17)
    for i, f in enumerate(d["keyframes"]):
        h.add_keyframe(i, f["corner_scan"], f["surf_last_ds"], f["outlier_last_ds"], d["key_poses"][i])
    ids = list(range(len(d["keyframes"])))
    ext = torch.cuda.ExternalStream(h.stream)
    cq = torch.from_numpy(d["query"]["corner_scan_ds"]).cuda(); sq = torch.from_numpy(d["query"]["surf_total_ds"]).cuda()
    tb, ts, tt = [], [], []
    for r in range(60):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(ext)
        h.build_local_map(ids)
        e1.record(ext)
        pose, res = h.scan2map_dev(cq.data_ptr(), len(cq), sq.data_ptr(), len(sq), d["prior"])
        e2.record(ext); e2.synchronize()
        if r >= 10:
            tb.append(e0.elapsed_time(e1)); ts.append(e1.elapsed_time(e2)); tt.append(e0.elapsed_time(e2))
    print("build %.1f us, scan2map %.1f us, total %.1f us (median of 50, L2 warm)" % (np.median(tb) * 1e3, np.median(ts) * 1e3, np.median(tt) * 1e3))
