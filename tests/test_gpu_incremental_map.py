"""The incremental local map (csrc/map_incremental.cuh) against the oracle's full rebuild, bit for bit (`pytest -m gpu`).

A sliding set of keyframes is assembled on the device by merging the added keyframes into a resident, voxel-ordered
window instead of re-sorting the whole surrounding cloud; laserCloud{Corner,Surf}FromMapDS must stay IDENTICAL to what
extractSurroundingKeyFrames + pcl::VoxelGrid produce from scratch (mapOptmization.cpp:1218-1231; oracle.cpp) for every
list the two branches of extractSurroundingKeyFrames can generate: deque pop / push (:1127-1144), erase from the middle +
several push_backs (:1171-1216), an unchanged list, a corrected pose (correctPoses :1757-1785), and the cases the window
cannot represent, which must fall back to the full path without a trace.
"""
import numpy as np
import pytest

from conftest import SMALL_WORLD, rot_err, trans_err, upstream_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    import __graft_entry__ as ge
    ge.build()
    from lego_loam_sr_b200 import abi as _abi
    return _abi


@pytest.fixture(scope="module")
def frames(synth):
    seq = synth.make_sequence(seed=7, n_scans=64, step=0.9, world_kw=SMALL_WORLD)
    return seq["frames"], seq["gt"].astype(np.float32)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def oracle_maps(orc, frames, poses, ids, prm):
    return orc.build_local_map(np.array([poses[i] for i in ids], np.float32), [frames[i]["corner_scan"] for i in ids],
                               [frames[i]["surf_last_ds"] for i in ids], [frames[i]["outlier_last_ds"] for i in ids], prm)[:2]


def same(a, b):
    """Bit for bit, except that a NaN only has to be a NaN: the pass-through case keeps non-finite points, and a NaN born
    in the keyframe transform (inf - inf) is 0xffc00000 on x86 and 0x7fffffff on the GPU — the payload carries nothing."""
    if a.shape != b.shape:
        return False
    ba, bb = bits(a), bits(b)
    return bool(((ba == bb) | (np.isnan(a) & np.isnan(b))).all())


def check(h, orc, frames, poses, ids, prm, what):
    cds, sds = oracle_maps(orc, frames, poses, ids, prm)
    gc, gs = h.get_map_ds(0), h.get_map_ds(1)
    if not same(gc, cds) and gc.shape == cds.shape:
        bad = np.nonzero((bits(gc) != bits(cds)).any(axis=1))[0]
        print("corner rows that differ:", bad[:10], gc[bad[:5]], cds[bad[:5]])
    assert same(gc, cds), "corner map differs from the oracle: " + what
    assert same(gs, sds), "surf map differs from the oracle: " + what
    return cds, sds


@pytest.mark.parametrize("max_blocks", [0, 20, 3])
def test_sliding_window_bit_exact(abi, orc, frames, max_blocks):
    """Loop-closure-on shape: the last W keyframes. Odd frames read the map back before any scan (the merge runs as its
    own launch), even frames run a scan first (the merge runs at the head of the scan kernel)."""
    fr, gt = frames
    W = 12
    oprm = upstream_params(orc)
    poses = {}
    with abi.Handle(upstream_params(abi, lm_solver=1, lm_max_blocks=max_blocks)) as h:
        for k in range(1, 40):
            poses[k - 1] = gt[k - 1]
            h.add_keyframe(k - 1, fr[k - 1]["corner_scan"], fr[k - 1]["surf_last_ds"], fr[k - 1]["outlier_last_ds"], gt[k - 1])
            ids = list(range(max(0, k - W), k))
            h.build_local_map(ids)
            if k % 2 == 0 and k > 3:
                pose, res = h.scan2map(fr[k]["corner_scan_ds"], fr[k]["surf_total_ds"], gt[k])
                cds, sds = check(h, orc, fr, poses, ids, oprm, "frame %d after the scan" % k)
                ores, opose, _ = orc.scan2map(cds, sds, fr[k]["corner_scan_ds"], fr[k]["surf_total_ds"], gt[k], oprm)
                assert res.status == ores.status == 0 and res.iter_num == ores.iter_num
                assert (res.corner_map_ds, res.surf_map_ds) == (len(cds), len(sds))
                assert trans_err(pose, opose) <= 1e-4 and rot_err(pose, opose) <= 1e-5, k
            else:
                check(h, orc, fr, poses, ids, oprm, "frame %d before any scan" % k)


def test_default_branch_patterns_and_fallbacks(abi, orc, frames):
    """erase from the middle + several push_backs, the same list twice, a list that shrinks, a jump to unrelated
    keyframes (full path), and back to sliding."""
    fr, gt = frames
    oprm = upstream_params(orc)
    poses = {i: gt[i] for i in range(40)}
    with abi.Handle(upstream_params(abi, lm_solver=1)) as h:
        for i in range(40):
            h.add_keyframe(i, fr[i]["corner_scan"], fr[i]["surf_last_ds"], fr[i]["outlier_last_ds"], gt[i])
        lists = [[0], [0, 1], [0, 1, 2], [0, 1, 2, 3, 4], [0, 2, 3, 4, 5, 6], [0, 2, 3, 4, 5, 6], [2, 4, 6, 7], [2, 4, 6, 7, 8, 9, 10],
                 [4, 6, 7, 8, 9, 10, 11], [6, 7, 8, 9, 10, 11, 3],          # 3 comes BACK, at the end of the list
                 [30, 31, 32], [30, 31, 32, 33], [31, 32, 33, 34], [31, 33, 34, 35, 36, 37, 38, 39], [39], [39, 38]]
        for n, ids in enumerate(lists):
            h.build_local_map(ids)
            if n % 3 == 1:
                h.scan2map(fr[ids[-1] + 1 if ids[-1] + 1 < 40 else 0]["corner_scan_ds"], fr[0]["surf_total_ds"], gt[ids[-1]])
            check(h, orc, fr, poses, ids, oprm, "list %d %s" % (n, ids))


def test_pose_correction_and_replaced_keyframe(abi, orc, frames):
    """correctPoses (:1757-1785) moves key poses: the window must not keep points transformed by the old pose; a
    keyframe id that is added again with other clouds must not leave its old points behind either."""
    fr, gt = frames
    oprm = upstream_params(orc)
    poses = {i: gt[i].copy() for i in range(12)}
    clouds = {i: fr[i] for i in range(12)}
    with abi.Handle(upstream_params(abi, lm_solver=1)) as h:
        for k in range(1, 9):
            h.add_keyframe(k - 1, fr[k - 1]["corner_scan"], fr[k - 1]["surf_last_ds"], fr[k - 1]["outlier_last_ds"], gt[k - 1])
            h.build_local_map(list(range(k)))
            h.scan2map(fr[k]["corner_scan_ds"], fr[k]["surf_total_ds"], gt[k])
        poses[3] = poses[3] + np.array([0.001, -0.002, 0.0015, 0.07, -0.03, 0.05], np.float32)
        h.set_keyframe_pose(3, poses[3])
        ids = list(range(8))
        h.build_local_map(ids)
        check(h, orc, clouds, poses, ids, oprm, "after a pose correction")
        for k in (9, 10):  # sliding again after the correction
            h.add_keyframe(k - 1, fr[k - 1]["corner_scan"], fr[k - 1]["surf_last_ds"], fr[k - 1]["outlier_last_ds"], gt[k - 1])
            ids = list(range(k - 8, k))
            h.build_local_map(ids)
            h.scan2map(fr[k]["corner_scan_ds"], fr[k]["surf_total_ds"], gt[k])
            check(h, orc, clouds, poses, ids, oprm, "sliding after the correction, frame %d" % k)
        clouds[5] = fr[11]  # keyframe 5 replaced by other clouds
        h.add_keyframe(5, fr[11]["corner_scan"], fr[11]["surf_last_ds"], fr[11]["outlier_last_ds"], poses[5])
        h.build_local_map(ids)
        check(h, orc, clouds, poses, ids, oprm, "after a keyframe was replaced")


def test_window_edge_cases(abi, orc, frames):
    """Keyframes with an empty cloud, non-finite points, and a keyframe so far away that the merged box would overflow
    PCL's int32 voxel index (the pass-through case: not representable in the window, must come out of the full path)."""
    fr, gt = frames
    oprm = upstream_params(orc)
    rng = np.random.default_rng(3)
    clouds, poses = {}, {}
    for i in range(10):
        f = {k: fr[i][k].copy() for k in ("corner_scan", "surf_last_ds", "outlier_last_ds")}
        if i == 2:
            f["corner_scan"] = f["corner_scan"][:0]
        if i == 4:
            f["outlier_last_ds"] = f["outlier_last_ds"][:0]
        if i == 5:
            f["surf_last_ds"][rng.choice(len(f["surf_last_ds"]), 40, replace=False), 1] = np.nan
            f["corner_scan"][rng.choice(len(f["corner_scan"]), 9, replace=False), 0] = np.inf
        clouds[i], poses[i] = f, gt[i].copy()
    with abi.Handle(upstream_params(abi, lm_solver=1)) as h:
        for k in range(1, 9):
            c = clouds[k - 1]
            h.add_keyframe(k - 1, c["corner_scan"], c["surf_last_ds"], c["outlier_last_ds"], poses[k - 1])
            ids = list(range(max(0, k - 5), k))
            h.build_local_map(ids)
            if k % 2:
                h.scan2map(fr[k]["corner_scan_ds"], fr[k]["surf_total_ds"], gt[k])
            check(h, orc, clouds, poses, ids, oprm, "edge cases, frame %d" % k)
        # a keyframe 300 km away: (dx * dy * dz) of the 0.2 m corner grid exceeds int32 -> PCL passes the cloud through
        far = {k: clouds[8][k].copy() for k in clouds[8]}
        poses[8] = poses[8].copy()
        poses[8][3] += 3.0e5
        clouds[8] = far
        h.add_keyframe(8, far["corner_scan"], far["surf_last_ds"], far["outlier_last_ds"], poses[8])
        ids = [4, 5, 6, 7, 8]
        h.build_local_map(ids)
        pose, res = h.scan2map(fr[8]["corner_scan_ds"], fr[8]["surf_total_ds"], gt[8])
        assert res.flags & 3, "the voxel-index overflow must be reported as by the full path"
        check(h, orc, clouds, poses, ids, oprm, "voxel index overflow")
        ids = [5, 6, 7, 8, 9]
        h.add_keyframe(9, clouds[9]["corner_scan"], clouds[9]["surf_last_ds"], clouds[9]["outlier_last_ds"], poses[9])
        h.build_local_map(ids)
        check(h, orc, clouds, poses, ids, oprm, "after the overflow")
