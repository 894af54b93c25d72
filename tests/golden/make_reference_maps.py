"""Writes tests/golden/reference_maps.npz: the REFERENCE'S OWN recorded maps as real-distribution inputs.

/root/reference/Result/0318_test/{cornerMap,surfaceMap}.pcd are maps the reference node itself saved from a real
VLP-16 run (ASCII PCD, FIELDS x y z intensity). They exist only in the authoring container, so this script (run there)
commits what the tests need: the two clouds (the corner map thinned 1:2 to keep the fixture below 1 MB), a query cloud
taken from the surface map, and the 5-NN answers of the reference's own vendored nanoflann (oracle/_ref, compiled in
place from /root/reference/LeGO-LOAM/include/lego_loam/nanoflann.hpp) for those queries against the VoxelGrid-ed
corner map — asserted equal to brute force (ties by index) before they are written.

    python tests/golden/make_reference_maps.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import orc  # noqa: E402

SRC = "/root/reference/Result/0318_test"


def read_pcd(path):
    with open(path) as f:
        lines = f.read().split("\n")
    start = next(i for i, ln in enumerate(lines) if ln.startswith("DATA")) + 1
    assert lines[2].strip() == "FIELDS x y z intensity", lines[2]
    a = np.array([ln.split() for ln in lines[start:] if ln.strip()], dtype=np.float32)
    assert a.shape[1] == 4
    return np.ascontiguousarray(a)


def main():
    orc.build()
    assert orc.have_ref(), "run where /root/reference exists (oracle/_ref must be built)"
    corner = read_pcd(os.path.join(SRC, "cornerMap.pcd"))[::2].copy()
    surf = read_pcd(os.path.join(SRC, "surfaceMap.pcd"))
    cds = orc.voxel_downsample(corner, 0.2)[0]
    rng = np.random.default_rng(5)
    # queries: surface-map points and corner-map points, nudged so that they are not map points themselves
    q = np.concatenate([surf[:, :], corner[rng.choice(len(corner), 3000, replace=False)]]).astype(np.float32)
    q[:, :3] += rng.normal(0, 0.05, (len(q), 3)).astype(np.float32)
    idx, d2 = orc.knn5(cds, q, orc.KNN_BRUTE)
    inn, dn = orc.knn5(cds, q, orc.KNN_NANOFLANN)
    assert np.array_equal(dn.view(np.uint32), d2.view(np.uint32)), "nanoflann distances differ from brute force"
    tie_free = (np.diff(d2, axis=1) > 0).all(1)
    assert np.array_equal(inn[tie_free], idx[tie_free]), "nanoflann neighbours differ from brute force"
    out = os.path.join(HERE, "reference_maps.npz")
    np.savez_compressed(out, corner_map=corner, surf_map=surf, queries=q, knn_idx=inn, knn_d2=dn, tie_free=tie_free)
    print("wrote %s (%d bytes): corner %d (ds %d), surf %d, %d queries (%d tie-free, %d with 5 neighbours inside 1 m)" %
          (out, os.path.getsize(out), len(corner), len(cds), len(surf), len(q), tie_free.sum(), (d2[:, 4] < 1.0).sum()))


if __name__ == "__main__":
    main()
