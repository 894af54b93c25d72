"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on identical
seeded inputs (B200 box; `pytest -m gpu`).

Bars (BASELINE.json north_star):
  * voxel-downsampled map membership, output order AND coordinates: bit-exact;
  * 5-NN accept mask and neighbour index sets (ties by index) and squared distances: bit-exact;
  * per-point residual records: |diff| <= 2e-5 (same masks except points within float rounding of a threshold);
  * final 6-DoF pose after 10 LM iterations: <= 1e-4 m translation, <= 1e-5 rad rotation.
"""
import os
import sys

import numpy as np
import pytest

from conftest import rot_err, trans_err, upstream_params

pytestmark = pytest.mark.gpu

TOL_T = 1e-4   # metres, BASELINE.json
TOL_R = 1e-5   # radians, BASELINE.json
TOL_COEFF = 2e-5


@pytest.fixture(scope="module")
def abi():
    import __graft_entry__ as ge
    ge.build()
    from lego_loam_sr_b200 import abi as _abi
    return _abi


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def load_map(h, d):
    kf = d["keyframes"]
    for i, f in enumerate(kf):
        h.add_keyframe(i, f["corner_scan"], f["surf_last_ds"], f["outlier_last_ds"], d["key_poses"][i])
    h.build_local_map(list(range(len(kf))))


# ------------------------------------------------------------------------------------------ a3 VoxelGrid
@pytest.mark.parametrize("leaf", [0.2, 0.4, 1.0])
def test_voxel_downsample_bit_exact(abi, orc, single_problem, leaf):
    d = single_problem
    pts = np.concatenate([orc.transform_cloud(f["surf_last_ds"], d["key_poses"][i]) for i, f in enumerate(d["keyframes"][:12])])
    with abi.Handle() as h:
        out, member, flags = h.voxel_downsample(pts, leaf)
    o_out, o_member, o_flag = orc.voxel_downsample(pts, leaf)
    assert flags == o_flag == 0
    assert len(out) == len(o_out)
    assert np.array_equal(member, o_member)            # membership + output order
    assert np.array_equal(bits(out), bits(o_out))      # centroids, accumulated in the canonical order


def test_voxel_downsample_single_launch_path(abi, orc, single_problem):
    """Without the membership output ll_voxel_downsample runs rows a2-a3 as ONE cooperative launch (voxel_inline_kernel,
    what downsampleCurrentScan's six filters use): same bits as the oracle and as the stand-alone kernels."""
    d = single_problem
    kf = d["keyframes"]
    big = np.concatenate([f["surf_last_ds"] for f in kf]).astype(np.float32)
    rng = np.random.default_rng(3)
    with_nan = big[:20000].copy()
    with_nan[rng.choice(len(with_nan), 40, replace=False), 2] = np.nan
    all_nan = np.full((5, 4), np.nan, np.float32)
    one_voxel = np.tile(big[:1], (900, 1)) + np.float32(1e-3) * rng.random((900, 4)).astype(np.float32)
    cases = [(big, 0.4), (big, 0.2), (big[:1], 0.4), (big[:37], 1.0), (with_nan, 0.4), (big[:3000], 1e-4),
             # <= 1024 points: the one-block shared-memory kernel (voxel_small_kernel)
             (big[:1024], 0.4), (big[:1025], 0.4), (big[5000:5700], 0.2), (with_nan[:1000], 0.4), (big[:900], 1e-4),
             (all_nan, 0.4), (one_voxel, 2.0), (big[:65], 0.4), (big[:64], 0.4)]
    with abi.Handle(abi.default_params()) as h:
        for cloud, leaf in cases:
            out, _, flags = h.voxel_downsample(cloud, leaf, want_member=False)
            launches0 = h.launch_count
            out2, member, flags2 = h.voxel_downsample(cloud, leaf, want_member=True)
            o_out, _, o_flag = orc.voxel_downsample(cloud, leaf)
            assert flags == flags2 == o_flag
            assert np.array_equal(bits(out), bits(o_out)) and np.array_equal(bits(out2), bits(o_out)), (len(cloud), leaf)
            assert h.launch_count - launches0 >= 7  # (the membership variant is the seven-launch pipeline)


def test_downsample_scan_parity(abi, orc, synth, single_problem, single_maps):
    """f-2 as one call (ll_downsample_scan): the six clouds of downsampleCurrentScan (:1234-1267) bit-equal to the
    oracle's VoxelGrid chain, also while a local map is pending, also with empty clouds; the device copies it leaves
    feed ll_scan2map_dev to the same bits as the host clouds."""
    from conftest import SMALL_WORLD
    world = synth.World(seed=1, **SMALL_WORLD)  # single_problem's world: a raw scan at its query pose
    raw = synth.scan(world, single_problem["gt"], seed=77)
    z = np.zeros((0, 4), np.float32)
    cs, sl, ol = (np.ascontiguousarray(raw[k][:, :4], np.float32) for k in ("corner_scan", "surf_last", "outlier_last"))
    full = [cs[::2].copy(), sl, cs, sl[1::2].copy(), ol]  # corner_last, surf_last, corner_scan, surf_scan, outlier_last
    prm = abi.default_params()

    def oracle_chain(clouds):
        leaf = (prm.corner_leaf, prm.surf_leaf, prm.corner_leaf, prm.surf_leaf, prm.surf_leaf)
        ds = [orc.voxel_downsample(c, l)[0] if len(c) else z for c, l in zip(clouds, leaf)]
        total = np.concatenate([ds[1], ds[4]])
        ds.append(orc.voxel_downsample(total, prm.surf_leaf)[0] if len(total) else z)
        return ds

    d = single_problem
    kf = d["keyframes"]
    with abi.Handle(upstream_params(abi)) as h:
        for i, f in enumerate(kf):
            h.add_keyframe(i, f["corner_scan"], f["surf_last_ds"], f["outlier_last_ds"], d["key_poses"][i])
        for case, clouds in enumerate((full, [full[0], z, full[2], full[3], z], [z, full[1], z, z, z], [z] * 5,
                                       [full[0][:1], full[1][:3], full[2], full[3], full[4][:2]])):
            if case == 1:
                h.build_local_map(list(range(len(kf))))  # from here on a map is pending while the filters run
            outs, dev = h.downsample_scan(clouds)
            want = oracle_chain(clouds)
            for k in range(6):
                assert dev.n_ds[k] == len(want[k]) and dev.flags[k] == 0, (case, k)
                assert np.array_equal(bits(outs[k]), bits(want[k])), (case, k)
            _, dev2 = h.downsample_scan(clouds, want_host=False)
            assert list(dev2.n_ds) == list(dev.n_ds)
        # the device copies are what the scan consumes: same result bits as the host clouds
        h.build_local_map(list(range(len(kf))))
        outs, dev = h.downsample_scan(full)
        pose_d, res_d = h.scan2map_dev(dev.d_ds[2], dev.n_ds[2], dev.d_ds[5], dev.n_ds[5], d["prior"])
        h.build_local_map(list(range(len(kf))))
        pose_h, res_h = h.scan2map(outs[2], outs[5], d["prior"])
        assert res_d.status == 0 and res_d.sel_num == res_h.sel_num
        assert np.array_equal(bits(pose_d), bits(pose_h))


def test_voxel_downsample_edge_cases(abi, orc):
    rng = np.random.default_rng(0)
    with abi.Handle() as h:
        out, member, flags = h.voxel_downsample(np.zeros((0, 4), np.float32), 0.2)
        assert len(out) == 0
        one = np.array([[1.5, 2.5, -3.5, 7.0]], np.float32)
        out, member, flags = h.voxel_downsample(one, 0.2)
        assert np.array_equal(out, one) and member.tolist() == [0]
        same = np.tile(np.array([[1.0, 1.0, 1.0, 2.0]], np.float32), (700, 1))
        out, member, flags = h.voxel_downsample(same, 0.2)
        o_out, o_member, _ = orc.voxel_downsample(same, 0.2)
        assert np.array_equal(bits(out), bits(o_out)) and np.array_equal(member, o_member)
        # ragged sizes around tile boundaries (1024 / 2048) and a PointXYZI-strided (32-byte) input
        for n in (1023, 1024, 1025, 2047, 2049, 5000):
            pts = np.concatenate([rng.uniform(-8, 8, (n, 3)), rng.uniform(0, 16, (n, 1))], 1).astype(np.float32)
            out, member, flags = h.voxel_downsample(pts, 0.4)
            o_out, o_member, _ = orc.voxel_downsample(pts, 0.4)
            assert np.array_equal(member, o_member) and np.array_equal(bits(out), bits(o_out)), n
        wide = np.zeros((len(pts), 8), np.float32)
        wide[:, :3] = pts[:, :3]
        wide[:, 3] = 1.0
        wide[:, 4] = pts[:, 3]
        out32, member32, _ = h.voxel_downsample(wide, 0.4)
        assert np.array_equal(bits(out32), bits(o_out))
        # negative coordinates, large offsets
        pts = np.concatenate([rng.uniform(-500, -480, (3000, 3)), rng.uniform(0, 16, (3000, 1))], 1).astype(np.float32)
        out, member, flags = h.voxel_downsample(pts, 0.2)
        o_out, o_member, _ = orc.voxel_downsample(pts, 0.2)
        assert np.array_equal(member, o_member) and np.array_equal(bits(out), bits(o_out))
        # int32 voxel index overflow: PCL passes the input through; so do the oracle and the device
        far = np.array([[0, 0, 0, 0], [3000, 3000, 3000, 1], [10, 10, 10, 2]], np.float32)
        out, member, flags = h.voxel_downsample(far, 0.01)
        o_out, o_member, o_flag = orc.voxel_downsample(far, 0.01)
        assert flags == 1 and o_flag == 1
        assert np.array_equal(bits(out), bits(o_out)) and np.array_equal(member, o_member)


# ------------------------------------------------------------------------------------------ a1/a2 local map
def test_local_map_assembly_bit_exact(abi, orc, single_problem, single_maps):
    with abi.Handle() as h:
        load_map(h, single_problem)
        cds, sds = h.get_map_ds(0), h.get_map_ds(1)
    assert len(cds) == len(single_maps["corner_ds"]) and len(sds) == len(single_maps["surf_ds"])
    assert np.array_equal(bits(cds), bits(single_maps["corner_ds"]))
    assert np.array_equal(bits(sds), bits(single_maps["surf_ds"]))


def test_local_map_subset_order_and_pose_update(abi, orc, single_problem):
    """Cache order = ids order (:1218-1222); set_keyframe_pose is what correctPoses does (:1757-1785)."""
    d = single_problem
    kf = d["keyframes"]
    ids = [7, 2, 11, 5]
    poses = d["key_poses"].copy()
    poses[2] = poses[2] + np.array([0.01, -0.02, 0.005, 0.3, -0.1, 0.2], np.float32)
    prm = orc.default_params()
    o_c, o_s, _, _ = orc.build_local_map(poses[ids], [kf[i]["corner_scan"] for i in ids], [kf[i]["surf_last_ds"] for i in ids],
                                         [kf[i]["outlier_last_ds"] for i in ids], prm)
    with abi.Handle() as h:
        for i, f in enumerate(kf):
            h.add_keyframe(i, f["corner_scan"], f["surf_last_ds"], f["outlier_last_ds"], d["key_poses"][i])
        h.set_keyframe_pose(2, poses[2])
        h.build_local_map(ids)
        assert np.array_equal(bits(h.get_map_ds(0)), bits(o_c))
        assert np.array_equal(bits(h.get_map_ds(1)), bits(o_s))
        with pytest.raises(abi.LegoLoamError) as e:
            h.build_local_map([0, 999])
        assert e.value.code == -4
    # map_downsample = 0 (BASELINE config 4): the raw concatenation is the map
    prm0 = orc.default_params(map_downsample=0)
    r_c, r_s, _, _ = orc.build_local_map(poses[ids], [kf[i]["corner_scan"] for i in ids], [kf[i]["surf_last_ds"] for i in ids],
                                         [kf[i]["outlier_last_ds"] for i in ids], prm0)
    with abi.Handle(abi.default_params(map_downsample=0)) as h:
        for i in ids:
            h.add_keyframe(i, kf[i]["corner_scan"], kf[i]["surf_last_ds"], kf[i]["outlier_last_ds"], poses[i])
        h.build_local_map(ids)
        assert np.array_equal(bits(h.get_map_ds(0)), bits(r_c))
        assert np.array_equal(bits(h.get_map_ds(1)), bits(r_s))


# ------------------------------------------------------------------------------------------ a4/a5 k-NN
def test_knn5_bit_exact(abi, orc, single_problem, single_maps):
    d = single_problem
    with abi.Handle() as h:
        h.set_local_map_raw(single_maps["corner_ds"], single_maps["surf_ds"])  # goes through VoxelGrid again
        # a voxel-filtered map is a fixed point of the filter at the same leaf only up to rounding, so query the
        # map the device actually holds
        maps = [h.get_map_ds(0), h.get_map_ds(1)]
        for which, cloud in ((0, "corner_scan_ds"), (1, "surf_total_ds")):
            q = orc.transform_cloud(d["query"][cloud], d["gt"])
            q = np.concatenate([q, q[:20] + np.float32(40.0)])  # some queries far outside the map
            idx, d2, accept = h.knn5(which, q)
            o_idx, o_d2 = orc.knn5(maps[which], q, orc.KNN_BRUTE)
            o_accept = (o_idx[:, 4] >= 0) & (o_d2[:, 4] < 1.0)
            assert np.array_equal(accept.astype(bool), o_accept)
            assert o_accept.sum() > 0.5 * len(d["query"][cloud])
            assert np.array_equal(idx[o_accept], o_idx[o_accept])
            assert np.array_equal(bits(d2[o_accept]), bits(o_d2[o_accept]))
            if orc.have_ref():  # and against the reference's own nanoflann
                n_idx, n_d2 = orc.knn5(maps[which], q, orc.KNN_NANOFLANN)
                tie_free = (np.diff(o_d2, axis=1) > 0).all(1) & o_accept
                assert np.array_equal(idx[tie_free], n_idx[tie_free])


def test_knn5_exact_ties_and_small_maps(abi, orc):
    """Equal distances: the five lowest indices win, in index order. Fewer than five candidates: rejected."""
    m = (np.array([[1, 0, 0, 0], [-1, 0, 0, 0], [0, 1, 0, 0], [0, -1, 0, 0], [0, 0, 1, 0], [0, 0, -1, 0], [1, 0, 0, 0],
                   [0, 1, 0, 0]], np.float32) * 0.5)
    surf = np.concatenate([np.random.default_rng(1).uniform(5, 9, (300, 3)), np.zeros((300, 1))], 1).astype(np.float32)
    with abi.Handle(abi.default_params(map_downsample=0)) as h:
        h.set_local_map_raw(m, surf)
        q = np.array([[0, 0, 0, 0], [0.3, 0.2, -0.1, 0], [30, 30, 30, 0]], np.float32)
        idx, d2, accept = h.knn5(0, q)
        o_idx, o_d2 = orc.knn5(m, q, orc.KNN_BRUTE)
        assert idx[0].tolist() == [0, 1, 2, 3, 4] and accept.tolist() == [1, 1, 0]
        assert np.array_equal(idx[:2], o_idx[:2]) and np.array_equal(bits(d2[:2]), bits(o_d2[:2]))
        h.set_local_map_raw(m[:4], surf)
        idx, d2, accept = h.knn5(0, q)
        assert accept.tolist() == [0, 0, 0]


def test_cell_table_stamps_survive_many_builds(abi, orc):
    """The cell tables are stamped, not cleared (csrc/map_build.cuh: cell_entry_x): 8-bit stamps, table reset every 255
    builds. More than 255 grid builds on ONE handle, over maps that move and shrink (so stale entries of earlier builds lie
    in cells the current map does not touch, and inside its box), through the stand-alone grid kernels (ll_knn5) and, every
    few builds, through the grid built inside the scan kernel: neighbours stay bit-exact across the wrap."""
    rng = np.random.default_rng(11)
    base = np.concatenate([rng.uniform(-6, 6, (4000, 3)), np.zeros((4000, 1))], 1).astype(np.float32)
    base[:, 1] *= 0.2
    prm = abi.default_params(map_downsample=0, apply_update=1, eigen_mode=1, iter_count_thres=2, stop_thres=0.0)
    with abi.Handle(prm) as h:
        for k in range(300):
            n = 4000 - 13 * (k % 50)
            shift = np.array([0.37 * (k % 7), 0.0, -0.41 * (k % 5), 0.0], np.float32)
            corner = (base[:n] + shift).astype(np.float32)
            surf = (base[::-1][:n] * np.float32(1.3) + shift).astype(np.float32)
            h.set_local_map_raw(corner, surf)
            if k % 4 == 3:   # the grid of this build is made inside the single-kernel scan
                h.scan2map(corner[::40], surf[::10], np.zeros(6, np.float32))
            if k < 250 and k % 25:
                continue     # (every build counts; only some are checked — all of them around the wrap)
            q = (corner[rng.choice(n, 200, replace=False)] + np.float32(0.05)).astype(np.float32)
            for which, m in ((0, corner), (1, surf)):
                idx, d2, acc = h.knn5(which, q)
                oidx, od2 = orc.knn5(m, q, orc.KNN_GRID)
                oacc = (oidx[:, 4] >= 0) & (od2[:, 4] < 1.0)
                assert np.array_equal(acc.astype(bool), oacc), (k, which)
                assert np.array_equal(idx[oacc], oidx[oacc]) and np.array_equal(bits(d2[oacc]), bits(od2[oacc])), (k, which)


# ------------------------------------------------------------------------------------------ a6/a7/a8 residuals
@pytest.mark.parametrize("eigen_mode", [1, 0])
def test_residuals_parity(abi, orc, single_problem, single_maps, eigen_mode):
    d = single_problem
    prm = upstream_params(abi, eigen_mode=eigen_mode)
    oprm = upstream_params(orc, eigen_mode=eigen_mode)
    with abi.Handle(prm) as h:
        load_map(h, d)
        for name, fn, ofn, mp, cloud in (("corner", h.corner_residuals, orc.corner_residuals, single_maps["corner_ds"], "corner_scan_ds"),
                                         ("surf", h.surf_residuals, orc.surf_residuals, single_maps["surf_ds"], "surf_total_ds")):
            q = d["query"][cloud]
            g = fn(q, d["prior"])
            o = ofn(mp, q, d["prior"], oprm)
            assert np.array_equal(g["accept"], o["accept"]), name
            acc = o["accept"].astype(bool)
            assert np.array_equal(g["knn_idx"][acc], o["knn_idx"][acc]), name
            assert np.array_equal(bits(g["knn_d2"][acc]), bits(o["knn_d2"][acc])), name
            mism = int((g["mask"] != o["mask"]).sum())
            both = (g["mask"] == 1) & (o["mask"] == 1)
            assert both.sum() > 0.3 * len(q), (name, both.sum(), len(q))
            assert mism <= max(2, len(q) // 1000), (name, mism)  # only points sitting on a threshold may flip
            err = np.abs(g["coeff"][both] - o["coeff"][both]).max()
            assert err <= TOL_COEFF, (name, err)
            nbit = int((bits(g["coeff"][both]) != bits(o["coeff"][both])).any(1).sum())
            print("%s eigen_mode=%d: %d kept, %d mask mismatches, max |dcoeff| %.2e, %d records not bit-identical"
                  % (name, eigen_mode, both.sum(), mism, err, nbit))


# ------------------------------------------------------------------------------------------ a9 LM step
@pytest.mark.parametrize("eigen_mode", [1, 0])
def test_lm_step_parity(abi, orc, single_problem, single_maps, eigen_mode):
    d = single_problem
    oprm = upstream_params(orc, eigen_mode=eigen_mode)
    rc = orc.corner_residuals(single_maps["corner_ds"], d["query"]["corner_scan_ds"], d["prior"], oprm)
    rs = orc.surf_residuals(single_maps["surf_ds"], d["query"]["surf_total_ds"], d["prior"], oprm)
    ori = np.concatenate([d["query"]["corner_scan_ds"][rc["mask"] == 1], d["query"]["surf_total_ds"][rs["mask"] == 1]])
    coeff = np.concatenate([rc["coeff"][rc["mask"] == 1], rs["coeff"][rs["mask"] == 1]])
    for eig_thresh in (100.0, 1e9):  # 1e9: every direction is "degenerate" -> exercises the projector
        ost = orc.new_state()
        oprm.eig_thresh = eig_thresh
        o = orc.lm_step(ori, coeff, 0, d["prior"], oprm, ost)
        prm = upstream_params(abi, eigen_mode=eigen_mode, eig_thresh=eig_thresh)
        with abi.Handle(prm) as h:
            g = h.lm_step(ori, coeff, 0, d["prior"])
            deg, P = h.get_degeneracy()
        assert g["ret"] == o["ret"]
        scale = np.abs(o["AtA"]).max()
        assert np.abs(g["AtA"] - o["AtA"]).max() <= 2e-5 * scale
        assert np.abs(g["Atb"] - o["Atb"]).max() <= 2e-5 * np.abs(o["Atb"]).max() + 1e-6
        assert deg == ost.is_degenerate
        if eigen_mode == 1 or eig_thresh == 100.0:
            assert np.abs(P - np.array(ost.matP).reshape(6, 6)).max() <= 1e-3
        assert np.abs(g["x"] - o["x"]).max() <= 5e-5, (g["x"], o["x"])
        assert np.abs(g["pose"] - o["pose"]).max() <= 5e-5
        assert abs(g["deltaR"] - o["deltaR"]) <= 1e-3 * max(1.0, o["deltaR"])
    # fewer than 50 residuals: skipped, pose untouched (:1453)
    with abi.Handle(upstream_params(abi)) as h:
        g = h.lm_step(ori[:20], coeff[:20], 0, d["prior"])
    assert g["ret"] == -1 and np.array_equal(g["pose"], d["prior"])


# ------------------------------------------------------------------------------------------ f-1 FeatureAssociation solves
@pytest.mark.parametrize("eigen_mode", [0, 1])
def test_solve_normal_eq_parity(abi, orc, eigen_mode):
    """calculateTransformationSurf / Corner / calculateTransformation (featureAssociation.cpp:1846-2289): the same
    reduce + small solve + degeneracy kernel with a column mask ([rx,rz,ty], [ry,tx,tz], all six), eigen threshold 10."""
    rng = np.random.default_rng(9)
    n = 1500
    rows = (rng.normal(size=(n, 6)) * np.array([3.0, 2.0, 3.0, 0.6, 0.3, 0.6])).astype(np.float32)
    x_true = np.array([0.004, -0.003, 0.002, 0.02, -0.01, 0.03])
    b = (rows.astype(np.float64) @ x_true + rng.normal(0, 1e-3, n)).astype(np.float32)
    with abi.Handle(abi.default_params(eigen_mode=eigen_mode)) as h:
        for mask in (0b010101, 0b101010, 0b111111):
            for thresh in (10.0, 1e9, 3000.0):
                gx, gP, gdeg = h.solve_normal_eq(rows, b, mask, 0, thresh)
                ox, oP, odeg = orc.solve_normal_eq(rows, b, mask, 0, thresh, eigen_mode)
                assert gdeg == odeg, (mask, thresh)
                assert np.abs(gx - ox).max() <= 2e-5 * max(1.0, np.abs(ox).max()), (mask, thresh, gx, ox)
                assert np.abs(gP - oP).max() <= 2e-3, (mask, thresh)
                # iteration > 0 re-uses the caller's projector (function-local matP in the reference)
                gx1, gP1, gdeg1 = h.solve_normal_eq(rows, b, mask, 3, thresh, matP=gP, is_degenerate=gdeg)
                ox1, oP1, odeg1 = orc.solve_normal_eq(rows, b, mask, 3, thresh, eigen_mode, matP=oP, is_degenerate=odeg)
                assert gdeg1 == odeg1 and np.abs(gx1 - ox1).max() <= 2e-5 * max(1.0, np.abs(ox1).max())
        # not degenerate: the solution is the least-squares one
        gx, gP, gdeg = h.solve_normal_eq(rows, b, 0b111111, 0, 10.0)
        assert gdeg == 0 and np.abs(gx - x_true).max() < 2e-4


def test_voxel_downsample_skips_non_finite_points(abi, orc):
    """pcl::VoxelGrid drops non-finite points (they are neither binned nor counted)."""
    rng = np.random.default_rng(12)
    pts = np.concatenate([rng.uniform(-6, 6, (3000, 3)), rng.uniform(0, 16, (3000, 1))], 1).astype(np.float32)
    bad = rng.choice(3000, 40, replace=False)
    pts[bad[:20], 0] = np.nan
    pts[bad[20:30], 1] = np.inf
    pts[bad[30:], 2] = -np.inf
    with abi.Handle() as h:
        out, member, flags = h.voxel_downsample(pts, 0.4)
    o_out, o_member, _ = orc.voxel_downsample(pts, 0.4)
    assert np.array_equal(member, o_member) and (member[bad] == -1).all()
    assert np.array_equal(bits(out), bits(o_out))


# ------------------------------------------------------------------------------------------ a10 full loop
@pytest.mark.parametrize("lm_solver,persistent_lm", [(0, 1), (1, 1), (0, 0), (1, 0)])
def test_scan2map_pose_within_tolerance(abi, orc, single_problem, single_maps, lm_solver, persistent_lm):
    """lm_solver: Eigen-sequence QR / LDL^T in double; persistent_lm: one cooperative launch / one launch per iteration."""
    d = single_problem
    with abi.Handle(upstream_params(abi, lm_solver=lm_solver, persistent_lm=persistent_lm)) as h:
        load_map(h, d)
        pose, res = h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
        pose2, res2 = h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
        # with persistent_lm = 1 the local map was assembled INSIDE the scan kernel (map_inline.cuh): same bits
        cds_gpu, sds_gpu = h.get_map_ds(0), h.get_map_ds(1)
    assert np.array_equal(bits(cds_gpu), bits(single_maps["corner_ds"]))
    assert np.array_equal(bits(sds_gpu), bits(single_maps["surf_ds"]))
    ores, opose, _ = orc.scan2map(single_maps["corner_ds"], single_maps["surf_ds"], d["query"]["corner_scan_ds"],
                                  d["query"]["surf_total_ds"], d["prior"], upstream_params(orc))
    assert res.status == 0 and res.iter_num == ores.iter_num == 10
    assert res.corner_map_ds == len(single_maps["corner_ds"]) and res.surf_map_ds == len(single_maps["surf_ds"])
    assert abs(res.sel_num - ores.sel_num) <= 3
    assert trans_err(pose, opose) <= TOL_T, (pose, opose)
    assert rot_err(pose, opose) <= TOL_R, (pose, opose)
    assert abs(res.min_lambda - ores.min_lambda) <= 1e-3 * ores.min_lambda
    assert abs(res.cf_mean - ores.cf_mean) <= 1e-5
    assert res.is_degenerate == ores.is_degenerate
    assert trans_err(pose, d["gt"]) < 0.02
    assert np.array_equal(bits(pose), bits(pose2)) and res2.sel_num == res.sel_num  # deterministic reduction order
    print("scan2map: |dt| %.2e m, |dr| %.2e rad vs oracle; sel %d vs %d" % (trans_err(pose, opose), rot_err(pose, opose), res.sel_num, ores.sel_num))


def _scan_both_ways(abi, corner, surf, cq, sq, prior, **kw):
    """The same scan through the single-kernel path (persistent_lm = 1: map assembly, search grids and the LM loop in one
    cooperative launch, map_inline.cuh) and through the stand-alone kernels (persistent_lm = 0)."""
    out = []
    for persistent in (1, 0):
        with abi.Handle(upstream_params(abi, persistent_lm=persistent, **kw)) as h:
            h.set_local_map_raw(corner, surf)
            pose, res = h.scan2map(cq, sq, prior)
            pose2, res2 = h.scan2map(cq, sq, prior)  # the map is built: the second scan only runs the LM loop
            out.append((pose, res, pose2, res2, h.get_map_ds(0), h.get_map_ds(1)))
    return out


def test_single_kernel_scan_equals_stand_alone_kernels(abi, orc, single_problem, single_maps):
    """Bit-for-bit: maps, poses and statistics of the two execution paths, on ordinary and on degenerate inputs."""
    d = single_problem
    rng = np.random.default_rng(11)
    kf = d["keyframes"]
    raw_c = np.concatenate([f["corner_scan"] for f in kf[:6]]).astype(np.float32)
    raw_s = np.concatenate([np.concatenate([f["surf_last_ds"], f["outlier_last_ds"]]) for f in kf[:6]]).astype(np.float32)
    cq, sq, prior = d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], np.zeros(6, np.float32)
    nan_s = raw_s.copy()
    nan_s[rng.choice(len(nan_s), 50, replace=False), 1] = np.nan           # non-finite points are dropped by VoxelGrid
    nan_s[rng.choice(len(nan_s), 7, replace=False), 0] = np.inf
    one_cell_c = (raw_c[:300] * np.float32(1e-3) + np.float32(0.5)).astype(np.float32)   # everything inside one 1 m cell
    one_cell_c[:, 3] = raw_c[:300, 3]
    cases = {
        "ordinary": (raw_c, raw_s, {}),
        "non-finite points": (raw_c, nan_s, {}),
        "tiny corner map (guard :1573)": (raw_c[:8], raw_s, {}),
        "empty corner map": (raw_c[:0], raw_s, {}),
        "single-cell corner map": (one_cell_c, raw_s, {}),
        "no map downsampling": (raw_c, raw_s, dict(map_downsample=0)),
        "voxel index overflow (PCL pass-through)": (raw_c, raw_s, dict(corner_leaf=1e-4, surf_leaf=1e-4)),
        "as shipped": (raw_c, raw_s, dict(apply_update=0, eigen_mode=0, iter_count_thres=200)),
    }
    for name, (c, s_, kw) in cases.items():
        (p1, r1, p1b, r1b, c1, s1), (p0, r0, p0b, r0b, c0, s0) = _scan_both_ways(abi, c, s_, cq, sq, prior, **kw)
        assert np.array_equal(bits(c1), bits(c0)) and np.array_equal(bits(s1), bits(s0)), name
        for (pa, ra), (pb, rb) in (((p1, r1), (p0, r0)), ((p1b, r1b), (p0b, r0b)), ((p1, r1), (p1b, r1b))):
            assert np.array_equal(bits(pa), bits(pb)), (name, pa, pb)
            assert (ra.status, ra.iter_num, ra.sel_num, ra.flags, ra.corner_map_ds, ra.surf_map_ds, ra.is_degenerate) == \
                   (rb.status, rb.iter_num, rb.sel_num, rb.flags, rb.corner_map_ds, rb.surf_map_ds, rb.is_degenerate), name
            assert bits(np.float32(ra.min_lambda)) == bits(np.float32(rb.min_lambda)) and \
                   bits(np.float32(ra.cf_mean)) == bits(np.float32(rb.cf_mean)), name
    # mixed path: asking for the map before the scan builds it with the stand-alone kernels; the scan kernel then only
    # builds the search grids itself
    with abi.Handle(upstream_params(abi)) as h:
        h.set_local_map_raw(raw_c, raw_s)
        c_first = h.get_map_ds(0)
        p_mixed, r_mixed = h.scan2map(cq, sq, prior)
    (p1, r1, _, _, c1, _), _ = _scan_both_ways(abi, raw_c, raw_s, cq, sq, prior)
    assert np.array_equal(bits(c_first), bits(c1)) and np.array_equal(bits(p_mixed), bits(p1)) and r_mixed.sel_num == r1.sel_num
    # and against the oracle for the ordinary case
    prm = upstream_params(orc)
    ocd = orc.voxel_downsample(raw_c, prm.corner_leaf)[0]
    osd = orc.voxel_downsample(raw_s, prm.surf_leaf)[0]
    (p1, r1, _, _, c1, s1), _ = _scan_both_ways(abi, raw_c, raw_s, cq, sq, prior)
    assert np.array_equal(bits(c1), bits(ocd)) and np.array_equal(bits(s1), bits(osd))


def test_call_sequences_around_a_pending_map(abi, orc, single_problem):
    """ll_build_local_map only PREPARES the map when persistent_lm = 1; whatever is called between it and ll_scan2map must
    neither disturb it nor see a half-built map."""
    d = single_problem
    kf = d["keyframes"]
    raw_c = np.concatenate([f["corner_scan"] for f in kf[:6]]).astype(np.float32)
    raw_s = np.concatenate([np.concatenate([f["surf_last_ds"], f["outlier_last_ds"]]) for f in kf[:6]]).astype(np.float32)
    cq, sq, prior = d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], np.zeros(6, np.float32)
    with abi.Handle(upstream_params(abi)) as h:
        h.set_local_map_raw(raw_c, raw_s)
        ref_pose, ref_res = h.scan2map(cq, sq, prior)
        ref_c = h.get_map_ds(0)
    o_small = orc.voxel_downsample(raw_s[:5000], 0.4)[0]
    with abi.Handle(upstream_params(abi)) as h:
        h.set_local_map_raw(raw_s[:100], raw_c[:100])      # a map that is never used ...
        h.set_local_map_raw(raw_c, raw_s)                  # ... is replaced by the next one
        out, _, _ = h.voxel_downsample(raw_s[:5000], 0.4)  # downsampleCurrentScan in between: own buffers
        assert np.array_equal(bits(out), bits(o_small))
        near = raw_s[::max(1, len(raw_s) // 64)][:64].copy()  # queries next to map points
        near[:, :3] += np.float32(0.01)
        idx, d2, acc = h.knn5(1, near)                      # a stage-level call builds map + grids the stand-alone way
        assert acc.sum() > 32
        pose, res = h.scan2map(cq, sq, prior)
        assert np.array_equal(bits(pose), bits(ref_pose)) and res.sel_num == ref_res.sel_num
        h.set_local_map_raw(raw_c, raw_s)
        out, _, _ = h.voxel_downsample(raw_s[:5000], 0.4)
        pose, res = h.scan2map(cq, sq, prior)              # pending map + aux problem in between: single-kernel path
        assert np.array_equal(bits(out), bits(o_small))
        assert np.array_equal(bits(pose), bits(ref_pose)) and np.array_equal(bits(h.get_map_ds(0)), bits(ref_c))
        pose0, res0 = h.scan2map(cq[:0], sq[:0], prior)     # no queries: nothing to optimise, pose unchanged
        assert res0.sel_num == 0 and np.array_equal(bits(pose0), bits(prior))


@pytest.mark.parametrize("max_blocks", [74, 20, 3])
def test_scan2map_on_part_of_the_gpu(abi, orc, single_problem, single_maps, max_blocks):
    """lm_max_blocks: the single-kernel scan on a fraction of the SMs (batched replay keeps several in flight). 74 blocks
    still hold the map (in-kernel assembly), 20 and 3 do not (stand-alone kernels, persistent LM loop only)."""
    d = single_problem
    with abi.Handle(upstream_params(abi, lm_max_blocks=max_blocks)) as h:
        load_map(h, d)
        pose, res = h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
        cds_gpu, sds_gpu = h.get_map_ds(0), h.get_map_ds(1)
    ores, opose, _ = orc.scan2map(single_maps["corner_ds"], single_maps["surf_ds"], d["query"]["corner_scan_ds"],
                                  d["query"]["surf_total_ds"], d["prior"], upstream_params(orc))
    assert np.array_equal(bits(cds_gpu), bits(single_maps["corner_ds"])) and np.array_equal(bits(sds_gpu), bits(single_maps["surf_ds"]))
    assert res.status == 0 and res.iter_num == 10 and abs(res.sel_num - ores.sel_num) <= 3
    assert trans_err(pose, opose) <= TOL_T and rot_err(pose, opose) <= TOL_R


def test_scan2map_early_exit_and_thresholds(abi, orc, single_problem, single_maps):
    d = single_problem
    for kw in (dict(stop_thres=0.05, iter_count_thres=30), dict(stop_thres=0.05, iter_count_thres=200),
               dict(stop_thres=0.0, iter_count_thres=3), dict(stop_thres=0.0, iter_count_thres=25)):
        ores, opose, _ = orc.scan2map(single_maps["corner_ds"], single_maps["surf_ds"], d["query"]["corner_scan_ds"],
                                      d["query"]["surf_total_ds"], d["prior"], upstream_params(orc, **kw))
        for dev_kw in (dict(lm_solver=0, persistent_lm=1), dict(lm_solver=1, persistent_lm=1), dict(lm_solver=1, persistent_lm=0)):
            with abi.Handle(upstream_params(abi, **kw, **dev_kw)) as h:
                load_map(h, d)
                pose, res = h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
            assert (res.iter_num, res.converged) == (ores.iter_num, ores.converged), (kw, dev_kw)
            assert trans_err(pose, opose) <= TOL_T and rot_err(pose, opose) <= TOL_R, (kw, dev_kw)


def test_scan2map_as_shipped(abi, orc, single_problem, single_maps):
    """apply_update = 0, eigen_mode = 0, 200 iterations: the reference as it ships (pose never moves)."""
    d = single_problem
    with abi.Handle(abi.default_params()) as h:
        load_map(h, d)
        pose, res = h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
    ores, opose, _ = orc.scan2map(single_maps["corner_ds"], single_maps["surf_ds"], d["query"]["corner_scan_ds"],
                                  d["query"]["surf_total_ds"], d["prior"], orc.default_params(iter_count_thres=200))
    assert np.array_equal(pose, d["prior"]) and np.array_equal(opose, d["prior"])
    assert (res.iter_num, res.converged, res.is_degenerate) == (ores.iter_num, ores.converged, ores.is_degenerate) == (200, 0, 0)
    assert abs(res.sel_num - ores.sel_num) <= 3
    assert abs(res.cf_mean - ores.cf_mean) <= 1e-5
    assert abs(res.min_lambda - ores.min_lambda) <= 2e-3 * abs(ores.min_lambda)
    assert np.abs(np.array(res.last_x) - np.array(ores.last_x)).max() <= 1e-4


def test_scan2map_guards(abi, orc, single_problem, single_maps):
    d = single_problem
    with abi.Handle(upstream_params(abi, map_downsample=0)) as h:
        h.set_local_map_raw(single_maps["corner_ds"][:10], single_maps["surf_ds"])
        pose, res = h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
        assert res.status == 1 and res.iter_num == 0 and np.array_equal(pose, d["prior"])  # :1573
        h.set_local_map_raw(single_maps["corner_ds"], single_maps["surf_ds"])
        pose, res = h.scan2map(d["query"]["corner_scan_ds"][:5], d["query"]["surf_total_ds"][:20], d["prior"])
        assert res.status == 0 and res.iter_num == 10 and res.sel_num < 50 and np.array_equal(pose, d["prior"])  # :1453
        pose, res = h.scan2map(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), d["prior"])
        assert res.status == 0 and res.sel_num == 0 and np.array_equal(pose, d["prior"])
    with abi.Handle(upstream_params(abi)) as h:
        with pytest.raises(abi.LegoLoamError) as e:
            h.scan2map(d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"], d["prior"])
        assert e.value.code == -7  # no local map yet


# ------------------------------------------------------------------------------------------ replay (config 2 shape)
def test_replay_matches_oracle(abi, orc, sequence):
    """The C++ MapOptimization mirror (host/) fed frame by frame vs the oracle's replay of the same frames."""
    from lego_loam_sr_b200 import host
    kw = dict(iter_count_thres=10, stop_thres=0.0)
    mo = host.MapOptimization(upstream_params(abi, **kw), loop_closure_enabled=True, search_num=50)
    view = mo.abi_handle()
    rp = orc.Replay(upstream_params(orc, **kw), window=50)
    oprm = upstream_params(orc, **kw)
    worst_t = worst_r = drift_t = drift_r = 0.0
    for k, f in enumerate(sequence["frames"]):
        pose, res = mo.process_frame(sequence["odom"][k], f, pre_downsampled=True)
        ores, nc, ns = rp.step(sequence["odom"][k], f["corner_scan"], f["corner_scan_ds"], f["surf_last_ds"],
                               f["outlier_last_ds"], f["surf_total_ds"])
        assert res.status == ores.status, k
        if res.status != 0:
            continue
        # (1) strict, per frame: the oracle registers the SAME scan against the map and prior the device used
        gc, gs = view.get_map_ds(0), view.get_map_ds(1)
        assert (len(gc), len(gs)) == (res.corner_map_ds, res.surf_map_ds)
        fres, fpose, _ = orc.scan2map(gc, gs, f["corner_scan_ds"], f["surf_total_ds"], mo.last_prior, oprm)
        worst_t = max(worst_t, trans_err(pose, fpose))
        worst_r = max(worst_r, rot_err(pose, fpose))
        assert fres.iter_num == res.iter_num and abs(fres.sel_num - res.sel_num) <= 3, k
        # (2) loose, end to end: two independent replays; rounding-level differences compound through the map
        drift_t = max(drift_t, trans_err(pose, ores.pose))
        drift_r = max(drift_r, rot_err(pose, ores.pose))
        assert abs(res.corner_map_ds - nc) <= max(3, nc // 200) and abs(res.surf_map_ds - ns) <= max(3, ns // 200), k
    assert mo.num_keyframes == len(sequence["frames"])
    assert worst_t <= TOL_T and worst_r <= TOL_R, (worst_t, worst_r)
    assert drift_t <= 5e-3 and drift_r <= 5e-4, (drift_t, drift_r)
    print("replay: per-frame worst |dt| %.2e m |dr| %.2e rad; end-to-end drift |dt| %.2e m |dr| %.2e rad over %d frames"
          % (worst_t, worst_r, drift_t, drift_r, len(sequence["frames"])))
    mo.close()


def test_replay_scan_side_downsample_on_device(abi, orc, synth, sequence):
    """The raw-frame path — downsampleCurrentScan (:1234-1267) on the device as part of the frame, loop-closure-off mirror —
    against the SAME replay fed with the oracle's VoxelGrid chain of the same raw clouds (the reference's six filters,
    :1236-1265): the two pose streams, the keyframe-id lists and the downsampled maps must be identical bit for bit (the
    device filters are bit-exact, so everything downstream sees the same numbers), and every frame's pose must match the
    oracle registering that scan on the device's map and prior."""
    from conftest import SMALL_WORLD
    from lego_loam_sr_b200 import host
    world = synth.World(seed=2, **SMALL_WORLD)
    n = 12
    gt = synth.trajectory(world, n, step=0.7, seed=2)
    kw = dict(iter_count_thres=10, stop_thres=0.0)
    prm, oprm = upstream_params(abi, lm_solver=1, **kw), upstream_params(orc, **kw)
    mo = host.MapOptimization(prm, loop_closure_enabled=False, search_num=50, search_radius=50.0)
    ref = host.MapOptimization(prm, loop_closure_enabled=False, search_num=50, search_radius=50.0)
    view = mo.abi_handle()
    worst_t = worst_r = 0.0
    for k in range(n):
        raw = synth.scan(world, gt[k], seed=2 * 100003 + k)
        pose, res = mo.process_frame(gt[k].astype(np.float32), raw, pre_downsampled=False)
        # the reference's chain on the CPU: corner clouds 0.2, the others 0.4, SurfTotal = VoxelGrid(SurfLastDS + OutlierLastDS)
        cs = np.ascontiguousarray(raw["corner_scan"][:, :4], np.float32)
        sl = np.ascontiguousarray(raw["surf_last"][:, :4], np.float32)
        ol = np.ascontiguousarray(raw["outlier_last"][:, :4], np.float32)
        csd = orc.voxel_downsample(cs, oprm.corner_leaf)[0]
        sld = orc.voxel_downsample(sl, oprm.surf_leaf)[0]
        old = orc.voxel_downsample(ol, oprm.surf_leaf)[0]
        std = orc.voxel_downsample(np.concatenate([sld, old]), oprm.surf_leaf)[0]
        frame = dict(corner_scan=cs, corner_scan_ds=csd, surf_last_ds=sld, outlier_last_ds=old, surf_total_ds=std)
        pose2, res2 = ref.process_frame(gt[k].astype(np.float32), frame, pre_downsampled=True)
        assert res.status == res2.status == (1 if k == 0 else 0), k
        assert np.array_equal(pose.view(np.uint32), pose2.view(np.uint32)), (k, pose, pose2)
        assert np.array_equal(mo.map_ids, ref.map_ids), k
        if k:
            assert (res.iter_num, res.sel_num, res.corner_map_ds, res.surf_map_ds) == (res2.iter_num, res2.sel_num, res2.corner_map_ds, res2.surf_map_ds)
            gc, gs = view.get_map_ds(0), view.get_map_ds(1)
            fres, fpose, _ = orc.scan2map(gc, gs, csd, std, mo.last_prior, oprm)
            worst_t, worst_r = max(worst_t, trans_err(pose, fpose)), max(worst_r, rot_err(pose, fpose))
            assert trans_err(pose, gt[k]) < 0.05
    assert worst_t <= 1e-4 and worst_r <= 1e-5, (worst_t, worst_r)
    mo.close()
    ref.close()


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties(abi, orc):
    """BASELINE config 3/4 sizes (2 M-point map, 100 k queries): too big for the brute-force oracle, so check
    size-independent properties — membership partition, sortedness, centroid containment, idempotence of the
    voxel stage, and k-NN against the oracle's grid index on a sample."""
    rng = np.random.default_rng(4)
    n = 2_000_000
    # a few planes + clutter, ~120 m extent
    pts = np.empty((n, 4), np.float32)
    pts[:, 0] = rng.uniform(-60, 60, n)
    pts[:, 2] = rng.uniform(-60, 60, n)
    pts[:, 1] = np.where(rng.random(n) < 0.7, -1.8 + rng.normal(0, 0.02, n), rng.uniform(-1.8, 6, n))
    pts[:, 3] = rng.uniform(0, 16, n)
    with abi.Handle(abi.default_params(map_downsample=0)) as h:
        out, member, flags = h.voxel_downsample(pts, 0.4)
        assert flags == 0 and member.min() == 0 and member.max() == len(out) - 1
        cnt = np.bincount(member, minlength=len(out))
        assert cnt.min() >= 1 and cnt.sum() == n
        inv = np.float32(1.0) / np.float32(0.4)
        vox = np.floor(pts[:, :3] * inv).astype(np.int64)
        vo = np.floor(out[:, :3] * inv).astype(np.int64)
        # each centroid lies in (or on the rounding edge of) the voxel of its members
        assert np.abs(vo[member] - vox).max() <= 1
        mn = vox.min(0)
        dd = vox.max(0) - mn + 1
        key = (vox[:, 0] - mn[0]) + dd[0] * ((vox[:, 1] - mn[1]) + dd[1] * (vox[:, 2] - mn[2]))
        first = np.full(len(out), -1, np.int64)
        first[member[::-1]] = key[::-1]
        assert (np.diff(first) > 0).all()                    # output ascends in voxel index
        assert len(np.unique(key)) == len(out)
        sums = np.zeros((len(out), 4))
        np.add.at(sums, member, pts.astype(np.float64))
        assert np.abs(sums / cnt[:, None] - out).max() < 5e-4  # centroid = mean
        out2, member2, _ = h.voxel_downsample(out, 0.4)
        assert abs(len(out2) - len(out)) <= len(out) // 1000   # (near-)idempotent
        # k-NN at scale on the un-downsampled 2 M map vs the oracle grid (exact, canonical)
        corner = pts[:200_000]
        h.set_local_map_raw(corner, pts)
        q = pts[rng.choice(n, 100_000, replace=False)].copy()
        q[:, :3] += rng.normal(0, 0.05, (len(q), 3)).astype(np.float32)
        idx, d2, accept = h.knn5(1, q)
        assert accept.mean() > 0.9
        sample = rng.choice(len(q), 3000, replace=False)
        o_idx, o_d2 = orc.knn5(pts, q[sample], orc.KNN_GRID)
        o_acc = o_d2[:, 4] < 1.0
        assert np.array_equal(accept[sample].astype(bool), o_acc)
        assert np.array_equal(idx[sample][o_acc], o_idx[o_acc])
        assert np.array_equal(bits(d2[sample][o_acc]), bits(o_d2[o_acc]))
        assert (np.diff(d2[accept == 1], axis=1) >= 0).all()


def test_reference_recorded_maps(abi, orc):
    """The reference's own recorded maps (tests/golden/reference_maps.npz, see make_reference_maps.py): VoxelGrid bit-exact
    vs the oracle, 5-NN bit-exact vs the answers of the reference's vendored nanoflann stored in the fixture."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_maps.npz"))
    with abi.Handle(abi.default_params()) as h:
        for cloud, leaf in ((g["corner_map"], 0.2), (g["surf_map"], 0.4), (g["corner_map"], 0.4)):
            out, _, flags = h.voxel_downsample(cloud, leaf)
            o_out = orc.voxel_downsample(cloud, leaf)[0]
            assert flags == 0 and np.array_equal(bits(out), bits(o_out)), leaf
    cds = orc.voxel_downsample(g["corner_map"], 0.2)[0]
    with abi.Handle(abi.default_params(map_downsample=0)) as h:
        h.set_local_map_raw(cds, g["surf_map"])
        idx, d2, accept = h.knn5(0, g["queries"])
    acc = g["knn_d2"][:, 4] < 1.0
    assert np.array_equal(accept.astype(bool), acc)
    assert np.array_equal(bits(d2[acc]), bits(g["knn_d2"][acc]))
    assert np.array_equal(idx[acc & g["tie_free"]], g["knn_idx"][acc & g["tie_free"]])


# ------------------------------------------------------------------------------------------ f-3 FeatureAssociation search
def test_fa_correspondences_parity(abi, orc, synth):
    """findCorrespondingCornerFeatures / findCorrespondingSurfFeatures on the device vs the oracle restatement: indices,
    keep flags and coefficient records bit for bit, over a search iteration, a reuse iteration and weighted iterations."""
    from test_oracle import _fa_problem
    for which in (0, 1):
        q, last, tc = _fa_problem(synth, which)
        with abi.Handle(abi.default_params()) as h:
            h.fa_set_last(which, last)
            ind_d = ind_o = None
            for it in (0, 3, 5, 7, 10):
                tci = (tc * np.float32(1.0 + 0.01 * it)).astype(np.float32)  # transformCur moves between iterations
                gd = h.fa_find_correspondences(which, q, tci, it, 25.0, ind=ind_d)
                go = orc.fa_correspondences(which, q, last, tci, it, 25.0, ind=ind_o)
                for name, a, b in zip(("ind1", "ind2", "ind3", "coeff", "keep"), gd, go):
                    if which == 0 and name == "ind3":
                        continue
                    assert np.array_equal(bits(a) if a.dtype == np.float32 else a, bits(b) if b.dtype == np.float32 else b), \
                        (which, it, name)
                ind_d, ind_o = gd[:3], go[:3]
                if it < 5:  # (from iteration 5 on the weight 1 - 1.8 |d| drops most pairs of this unconverged guess)
                    assert gd[4].sum() > len(q) // 4
        # the walks are ORDERED (first strict minimum, stop at the first far scan line): exact ties between duplicated
        # points and a cloud that is not sorted by scan line must come out like the point-by-point walk
        rng = np.random.default_rng(11)
        dup = last.copy()
        dup[1::3] = dup[0:len(dup) - 1:3][:len(dup[1::3])]           # every third point twice: equal distances
        mixed = dup.copy()
        blocks = np.array_split(np.arange(len(mixed)), 40)
        mixed = mixed[np.concatenate([blocks[k] for k in rng.permutation(40)])]  # scan lines interleaved in 40 runs
        for cloud in (dup, mixed):
            with abi.Handle(abi.default_params()) as h:
                h.fa_set_last(which, cloud)
                gd = h.fa_find_correspondences(which, q, tc, 0)
                go = orc.fa_correspondences(which, q, cloud, tc, 0)
                for name, a, b in zip(("ind1", "ind2", "ind3", "coeff", "keep"), gd, go):
                    if which == 0 and name == "ind3":
                        continue
                    if name == "coeff":  # (a plane through a duplicated point is 0/0: NaN on both sides, payloads differ)
                        nan = np.isnan(a)
                        assert np.array_equal(nan, np.isnan(b)), (which, name)
                        a, b = np.where(nan, np.float32(0), a), np.where(nan, np.float32(0), b)
                    assert np.array_equal(bits(a) if a.dtype == np.float32 else a, bits(b) if b.dtype == np.float32 else b), \
                        (which, name)
                assert (gd[1] >= 0).sum() > len(q) // 8
        # degenerate inputs: an empty / tiny previous cloud gives no correspondences and no error
        with abi.Handle(abi.default_params()) as h:
            h.fa_set_last(which, last[:0])
            i1, i2, i3, coeff, keep = h.fa_find_correspondences(which, q, tc, 0)
            assert (i1 == -1).all() and (keep == 0).all()
            h.fa_set_last(which, last[:3])
            g1 = h.fa_find_correspondences(which, q, tc, 0)
            o1 = orc.fa_correspondences(which, q, last[:3], tc, 0)
            assert np.array_equal(g1[0], o1[0]) and np.array_equal(g1[1], o1[1]) and np.array_equal(g1[4], o1[4])


# ------------------------------------------------------------------------------------------ batched replay (config 5 shape)
def test_batched_replay_is_independent_of_concurrency(abi, synth):
    """Several sequences in flight on one GPU (one handle + stream + host thread each) give bit-identical poses to
    replaying them one after the other: handles share nothing."""
    from lego_loam_sr_b200 import batch, host
    small = dict(n_blocks=8, n_poles=50, half_x=25, half_z=35)
    seqs = {s: synth.make_sequence(seed=100 + s, n_scans=6, step=0.7, world_kw=small) for s in range(4)}

    def make_node():
        return host.MapOptimization(upstream_params(abi, iter_count_thres=6), loop_closure_enabled=True, search_num=50)

    def feed(node, data, k):
        f = data["frames"][k]
        return node.process_frame_host(data["odom"][k], [f["corner_scan"], f["corner_scan_ds"], f["surf_last_ds"],
                                                         f["outlier_last_ds"], f["surf_total_ds"]])

    seq_rec = batch.replay_sequences(seqs, make_node, feed, n_threads=1)
    par_rec = batch.replay_sequences(seqs, make_node, feed, n_threads=4)
    a = batch.gather_records(seq_rec, 4, 6, 1, 0)
    b = batch.gather_records(par_rec, 4, 6, 1, 0)
    assert np.array_equal(a, b)
    assert (a[:, 1:, batch.RECORD_FIELDS.index("status")] == 0).all()
    assert (a[:, 1:, batch.RECORD_FIELDS.index("iter_num")] == 6).all()


def test_scan2map_with_overlapped_query_upload():
    """LLB200_QUERY_OVERLAP=1 (off by default: profiles/r02_ab_records.txt) launches the scan kernel before the query clouds
    are uploaded and lets it wait for their arrival stamp; the switch is read once per process, so the scan-level parity
    tests are run again in a child process with it on."""
    import subprocess
    env = dict(os.environ, LLB200_QUERY_OVERLAP="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__), "-k",
                        "scan2map_pose_within_tolerance or replay_matches_oracle or single_kernel_scan_equals"],
                       env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
