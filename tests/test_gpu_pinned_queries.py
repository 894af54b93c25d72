"""Query clouds in page-locked host memory (`pytest -m gpu`).

ll_scan2map(_async) copies pcl::PointXYZI clouds that lie in page-locked memory by DMA straight from the caller's
buffers (32-byte records, the kernel steps over them: LMArgs::q_step) instead of packing them on the CPU first
(include/legoloam_b200.h). The result must not depend on where the clouds lie or how they are laid out: pageable or
page-locked, 16- or 32-byte records, persistent or per-iteration kernels — the same pose bit for bit, and the oracle's
pose (scan2MapOptimization, mapOptmization.cpp:1572-1610) within the usual tolerance.
"""
import numpy as np
import pytest

from conftest import rot_err, trans_err, upstream_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    import __graft_entry__ as ge
    ge.build()
    from lego_loam_sr_b200 import abi as _abi
    return _abi


def pcl_layout(a):
    out = np.zeros((len(a), 8), np.float32)
    out[:, :3] = a[:, :3]
    out[:, 3] = 1.0
    out[:, 4] = a[:, 3]
    return out


def pinned(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).pin_memory().numpy()


def test_pinned_clouds_give_the_same_scan(abi, orc, single_problem, single_maps):
    d = single_problem
    cq, sq = d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"]
    kf = d["keyframes"]
    variants = {
        "pageable16": (np.ascontiguousarray(cq, np.float32), np.ascontiguousarray(sq, np.float32)),
        "pageable32": (pcl_layout(cq), pcl_layout(sq)),
        "pinned16": (pinned(cq), pinned(sq)),
        "pinned32": (pinned(pcl_layout(cq)), pinned(pcl_layout(sq))),
        "pinned32_no_corner": (np.zeros((0, 8), np.float32), pinned(pcl_layout(sq))),
        "pageable32_no_corner": (np.zeros((0, 8), np.float32), pcl_layout(sq)),
    }
    out = {}
    for persistent in (1, 0):
        with abi.Handle(upstream_params(abi, persistent_lm=persistent)) as h:
            for i, f in enumerate(kf):
                h.add_keyframe(i, f["corner_scan"], f["surf_last_ds"], f["outlier_last_ds"], d["key_poses"][i])
            for name, (c, s) in variants.items():
                h.build_local_map(list(range(len(kf))))
                pose, res = h.scan2map(c, s, d["prior"])
                assert res.status == 0 and res.iter_num == 10, name
                out[(persistent, name)] = (pose.copy(), res.sel_num)
            # twice in a row on the same (already built) map, alternating layouts: the device staging is reused
            for name in ("pinned32", "pageable16", "pinned32"):
                pose, res = h.scan2map(*variants[name], d["prior"])
                assert np.array_equal(pose.view(np.uint32), out[(persistent, name)][0].view(np.uint32)), name
    for persistent in (1, 0):
        ref_pose, ref_sel = out[(persistent, "pageable16")]
        for name in ("pageable32", "pinned16", "pinned32"):
            pose, sel = out[(persistent, name)]
            assert sel == ref_sel, (persistent, name)
            assert np.array_equal(pose.view(np.uint32), ref_pose.view(np.uint32)), (persistent, name)
        a, b = out[(persistent, "pinned32_no_corner")], out[(persistent, "pageable32_no_corner")]
        assert a[1] == b[1] and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    # and the oracle on the same maps
    oprm = upstream_params(orc)
    ores, opose, _ = orc.scan2map(single_maps["corner_ds"], single_maps["surf_ds"], cq, sq, d["prior"], oprm)
    gpose, gsel = out[(1, "pinned32")]
    assert trans_err(gpose, opose) <= 1e-4 and rot_err(gpose, opose) <= 1e-5
    assert abs(int(gsel) - int(ores.sel_num)) <= 3
