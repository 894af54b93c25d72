"""Writes tests/golden/oracle_golden.npz: oracle outputs on a small seeded problem.

Run in the authoring container (where /root/reference exists) so that the k-NN answers come from the
REFERENCE'S OWN vendored nanoflann (oracle/_ref, compiled in place from
/root/reference/LeGO-LOAM/include/lego_loam/nanoflann.hpp); the script asserts they equal the canonical
(brute-force, ties-by-index) answers before writing. Everything else comes from the restatement in
oracle/oracle.cpp ("parity unpinned" for the PCL / Eigen pieces: the reference has no fixtures for them).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import orc  # noqa: E402
from lego_loam_sr_b200 import synth  # noqa: E402


def main():
    orc.build()
    world = dict(n_blocks=10, n_poles=60, half_x=30, half_z=40)
    d = synth.make_single(seed=7, n_map_scans=8, step=0.7, world_kw=world)
    prm = orc.default_params(apply_update=1, eigen_mode=1, iter_count_thres=10, stop_thres=0.0)
    kf = d["keyframes"]
    cds, sds, _, _ = orc.build_local_map(d["key_poses"], [f["corner_scan"] for f in kf], [f["surf_last_ds"] for f in kf],
                                         [f["outlier_last_ds"] for f in kf], prm)
    cq, sq = d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"][::3]
    knn_q = orc.transform_cloud(sq, d["gt"])
    idx, d2 = orc.knn5(sds, knn_q, orc.KNN_BRUTE)
    if orc.have_ref():
        inn, dn = orc.knn5(sds, knn_q, orc.KNN_NANOFLANN)
        assert np.array_equal(dn.view(np.uint32), d2.view(np.uint32)), "nanoflann distances differ from brute force"
        tie_free = (np.diff(d2, axis=1) > 0).all(1)
        assert np.array_equal(inn[tie_free], idx[tie_free]), "nanoflann neighbours differ from brute force"
        print("k-NN golden vectors confirmed against the reference's nanoflann (%d queries, %d tie-free)" %
              (len(knn_q), tie_free.sum()))
        orc.set_knn_backend(orc.KNN_NANOFLANN)
    else:
        print("WARNING: oracle/_ref missing; k-NN golden vectors come from brute force only")
    voxel_in = orc.transform_cloud(kf[0]["surf_last_ds"], d["key_poses"][0])
    voxel_in = np.concatenate([voxel_in, orc.transform_cloud(kf[1]["surf_last_ds"], d["key_poses"][1])])
    vout, vmem, _ = orc.voxel_downsample(voxel_in, 0.4)
    rs = orc.surf_residuals(sds, sq, d["prior"], prm)
    rc = orc.corner_residuals(cds, cq, d["prior"], prm)
    res, pose, st = orc.scan2map(cds, sds, cq, sq, d["prior"], prm)
    orc.set_knn_backend(orc.KNN_GRID)
    out = os.path.join(HERE, "oracle_golden.npz")
    np.savez_compressed(out, corner_map=cds, surf_map=sds, corner_q=cq, surf_q=sq, prior=d["prior"], gt=d["gt"],
                        knn_queries=knn_q, knn_idx=idx, knn_d2=d2, voxel_in=voxel_in, voxel_out=vout, voxel_member=vmem,
                        surf_mask=rs["mask"], surf_coeff=rs["coeff"], corner_mask=rc["mask"], corner_coeff=rc["coeff"],
                        pose10=pose, sel_num=res.sel_num)
    print("wrote", out, os.path.getsize(out), "bytes; maps", len(cds), len(sds), "queries", len(cq), len(sq),
          "final err", pose - d["gt"])


if __name__ == "__main__":
    main()
