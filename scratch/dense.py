"""BASELINE configs[3] shape: ~2 M surrounding-map points WITHOUT map downsampling, 20 k corner + 80 k surf queries."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from lego_loam_sr_b200 import abi
rng = np.random.default_rng(4)
def world(n):
    p = np.empty((n, 4), np.float32)
    p[:, 0] = rng.uniform(-70, 70, n); p[:, 2] = rng.uniform(-70, 70, n)
    p[:, 1] = np.where(rng.random(n) < 0.7, -1.8 + rng.normal(0, 0.02, n), rng.uniform(-1.8, 6, n))
    p[:, 3] = rng.uniform(0, 16, n)
    return p
for ds in (0, 1):
    surf = world(2_000_000 if ds == 0 else 2_500_000); corner = world(200_000)
    prm = abi.default_params(apply_update=1, eigen_mode=1, iter_count_thres=10, stop_thres=0.0, map_downsample=ds)
    def vox_order(q, leaf):  # the scan-side VoxelGrid emits its points in ascending voxel index (z-major)
        k = np.floor(q[:, :3] / leaf).astype(np.int64)
        return q[np.lexsort((k[:, 0], k[:, 1], k[:, 2]))]
    cq = vox_order(corner[rng.choice(len(corner), 20_000, replace=False)], 0.2).copy(); sq = vox_order(surf[rng.choice(len(surf), 80_000, replace=False)], 0.4).copy()
    pose = np.array([0.002, -0.003, 0.002, 0.03, -0.02, 0.03], np.float32)
    with abi.Handle(prm) as h:
        for r in range(4):
            if r == 1: h.profile_enable(True); h.profile_read()
            h.set_local_map_raw(corner, surf)
            p, res = h.scan2map(cq, sq, pose)
        prof = h.profile_read()
        print("map_downsample", ds, "maps", res.corner_map_ds, res.surf_map_ds, "sel", res.sel_num, "iters", res.iter_num)
        print("  ", " ".join("%s=%.1fus" % (k, v[0] * 1e3 / max(v[1], 1)) for k, v in prof.items()))
