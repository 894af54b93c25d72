import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, orc
from lego_loam_sr_b200 import synth, abi
from conftest import SMALL_WORLD, upstream_params
np.set_printoptions(precision=5, suppress=True, linewidth=200)
d = synth.make_single(seed=1, n_map_scans=30, step=0.7, world_kw=SMALL_WORLD)
kf = d["keyframes"]
prm0 = orc.default_params()
cds, sds, _, _ = orc.build_local_map(d["key_poses"], [f["corner_scan"] for f in kf], [f["surf_last_ds"] for f in kf], [f["outlier_last_ds"] for f in kf], prm0)
for eigen_mode in (1, 0):
    oprm = upstream_params(orc, eigen_mode=eigen_mode)
    rc = orc.corner_residuals(cds, d["query"]["corner_scan_ds"], d["prior"], oprm)
    rs = orc.surf_residuals(sds, d["query"]["surf_total_ds"], d["prior"], oprm)
    ori = np.concatenate([d["query"]["corner_scan_ds"][rc["mask"] == 1], d["query"]["surf_total_ds"][rs["mask"] == 1]])
    coeff = np.concatenate([rc["coeff"][rc["mask"] == 1], rs["coeff"][rs["mask"] == 1]])
    for th in (100.0, 1e9):
        ost = orc.new_state(); oprm.eig_thresh = th
        o = orc.lm_step(ori, coeff, 0, d["prior"], oprm, ost)
        with abi.Handle(upstream_params(abi, eigen_mode=eigen_mode, eig_thresh=th)) as h:
            g = h.lm_step(ori, coeff, 0, d["prior"]); deg, P = h.get_degeneracy()
        print("mode", eigen_mode, "thresh", th, "ret", g["ret"], o["ret"], "deg", deg, ost.is_degenerate, "minlam", ost.min_lambda)
        print(" x gpu", g["x"]); print(" x orc", o["x"])
        print(" P gpu diag", np.diag(P), " P orc diag", np.diag(np.array(ost.matP).reshape(6,6)))
        print(" eig(AtA)", np.linalg.eigvalsh(o["AtA"].astype(np.float64)))
