"""f-1 + f-3 as ONE device-resident call: ll_fa_update_transformation against the oracle's restatement of
FeatureAssociation::updateTransformation (featureAssociation.cpp:2505-2535 with calculateTransformationSurf :1846-2010
and calculateTransformationCorner :2013-2143 inside). `pytest -m gpu`.

Tolerances: the correspondence records are bit-identical to the oracle's (test_gpu_parity.py::test_fa_correspondences_parity)
and the Jacobian rows are the same float expressions, so the two sides differ only in how the 3x3 normal equations are
accumulated (the oracle adds float products one by one, the device adds them in double) and, in upstream eigen mode, in
the eigen solver behind the projector. transformCur after both loops: <= 1e-5 (rad and m); iteration counts and the
per-iteration row counts: equal.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def abi():
    import __graft_entry__ as ge
    ge.build()
    from lego_loam_sr_b200 import abi as _abi
    return _abi


def _sweeps(synth, orc, seed=2, step=0.7):
    """Two consecutive synthetic sweeps as FeatureAssociation sees them: `last` = the previous sweep's feature clouds in
    extraction order (by scan line; intensity = ring + column / 10000), queries = the next sweep's sharp / flat points."""
    world = synth.World(seed=seed)
    gt = synth.trajectory(world, 2, step=step, seed=seed)
    raw = [synth.scan(world, gt[k], seed=1 + k) for k in range(2)]

    def ds(c, leaf):
        return orc.voxel_downsample(np.ascontiguousarray(c[:, :4], np.float32), leaf)[0]

    cur_c, cur_s = ds(raw[0]["corner_scan"], 0.2), ds(raw[0]["surf_last"], 0.4)[::4].copy()
    prev_c, prev_s = ds(raw[1]["corner_scan"], 0.2), ds(raw[1]["surf_last"], 0.4)
    last_c = prev_c[np.argsort(prev_c[:, 3].astype(np.int32), kind="stable")]
    last_s = prev_s[np.argsort(prev_s[:, 3].astype(np.int32), kind="stable")]
    return cur_c, cur_s, last_c, last_s


def _compare(g, o, what):
    assert g["skipped"] == o["skipped"], what
    assert (g["iters_surf"], g["iters_corner"]) == (o["iters_surf"], o["iters_corner"]), (what, g["iters_surf"], g["iters_corner"], o["iters_surf"], o["iters_corner"])
    for name in ("trace_surf", "trace_corner"):
        assert g[name].shape == o[name].shape, (what, name)
        assert np.array_equal(g[name][:, 3], o[name][:, 3]), (what, name, "row counts per iteration differ")
        assert np.abs(g[name][:, :3] - o[name][:, :3]).max(initial=0.0) <= TOL, (what, name)
    d = np.abs(np.asarray(g["transform_cur"]) - np.asarray(o["transform_cur"])).max()
    assert d <= TOL, (what, d)
    assert g["is_degenerate"] == o["is_degenerate"], what
    return d


@pytest.mark.parametrize("eigen_mode", [0, 1])
def test_update_transformation_matches_oracle(abi, orc, synth, eigen_mode):
    cur_c, cur_s, last_c, last_s = _sweeps(synth, orc)
    with abi.Handle(abi.default_params(eigen_mode=eigen_mode)) as h:
        h.fa_set_last(0, last_c)
        h.fa_set_last(1, last_s)
        for t0 in (np.zeros(6, np.float32), np.array([0.004, -0.006, 0.003, 0.05, -0.02, 0.30], np.float32)):
            h.fa_reset_state()
            g = h.fa_update_transformation(cur_c, cur_s, t0, want_trace=True)
            o = orc.fa_update_transformation(cur_c, cur_s, last_c, last_s, t0, eigen_mode=eigen_mode)
            d = _compare(g, o, "eigen_mode %d" % eigen_mode)
            assert g["iters_surf"] + g["iters_corner"] > 5           # the loops really ran
            assert g["n_sel_surf"] > 100 and g["n_sel_corner"] > 50
            print("eigen_mode %d: %d + %d iterations, |d transformCur| %.2e vs oracle" % (eigen_mode, g["iters_surf"], g["iters_corner"], d))
        # isDegenerate / matP persist from call to call (class member / reused stack slot): a second call WITHOUT a reset
        st = o["state"]
        g2 = h.fa_update_transformation(cur_c[::2], cur_s[::2], np.zeros(6, np.float32), want_trace=True)
        o2 = orc.fa_update_transformation(cur_c[::2], cur_s[::2], last_c, last_s, np.zeros(6, np.float32), eigen_mode=eigen_mode, state=st)
        _compare(g2, o2, "second call, state carried over")
        if orc.have_ref():  # the same loop with the reference's own nanoflann as kdtree*Last
            o3 = orc.fa_update_transformation(cur_c, cur_s, last_c, last_s, np.zeros(6, np.float32), eigen_mode=eigen_mode, use_nanoflann=True)
            h.fa_reset_state()
            g3 = h.fa_update_transformation(cur_c, cur_s, np.zeros(6, np.float32), want_trace=True)
            _compare(g3, o3, "oracle with nanoflann")


def test_update_transformation_guards_and_small_inputs(abi, orc, synth):
    cur_c, cur_s, last_c, last_s = _sweeps(synth, orc)
    t0 = np.array([0.001, 0.002, -0.001, 0.01, 0.0, 0.1], np.float32)
    with abi.Handle(abi.default_params(eigen_mode=1)) as h:
        # :2506 — fewer than 10 corner or 100 surf points in the previous sweep: nothing runs
        h.fa_set_last(0, last_c[:9])
        h.fa_set_last(1, last_s)
        g = h.fa_update_transformation(cur_c, cur_s, t0, want_trace=True)
        o = orc.fa_update_transformation(cur_c, cur_s, last_c[:9], last_s, t0, eigen_mode=1)
        assert g["skipped"] == o["skipped"] == 1 and np.array_equal(g["transform_cur"], t0)
        h.fa_set_last(0, last_c)
        h.fa_set_last(1, last_s[:99])
        assert h.fa_update_transformation(cur_c, cur_s, t0)["skipped"] == 1
        # fewer than 10 pairs (:2514 / :2527): both loops spin to 100 without ever solving
        h.fa_set_last(1, last_s)
        h.fa_reset_state()
        g = h.fa_update_transformation(cur_c[:6], cur_s[:7], t0, want_trace=True)
        o = orc.fa_update_transformation(cur_c[:6], cur_s[:7], last_c, last_s, t0, eigen_mode=1)
        _compare(g, o, "fewer than 10 pairs")
        assert g["iters_surf"] == g["iters_corner"] == 100 and np.array_equal(g["transform_cur"], t0)
        # empty feature clouds
        h.fa_reset_state()
        g = h.fa_update_transformation(cur_c[:0], cur_s[:0], t0)
        assert g["iters_surf"] == g["iters_corner"] == 100 and np.array_equal(g["transform_cur"], t0)


@pytest.mark.parametrize("max_blocks", [1, 7])
def test_update_transformation_on_part_of_the_gpu(abi, orc, synth, max_blocks):
    cur_c, cur_s, last_c, last_s = _sweeps(synth, orc, seed=5)
    with abi.Handle(abi.default_params(eigen_mode=1, lm_max_blocks=max_blocks)) as h:
        h.fa_set_last(0, last_c)
        h.fa_set_last(1, last_s)
        g = h.fa_update_transformation(cur_c, cur_s, np.zeros(6, np.float32), want_trace=True)
        o = orc.fa_update_transformation(cur_c, cur_s, last_c, last_s, np.zeros(6, np.float32), eigen_mode=1)
        _compare(g, o, "%d blocks" % max_blocks)


def test_update_transformation_with_staged_cloud():
    """The cp.async.bulk staging of the previous sweep's cloud (LLB200_FA_STAGE=1, off by default: profiles/r02_ab_records.txt)
    is read once per process, so the parity test above is run again in a child process with the switch on."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, LLB200_FA_STAGE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__),
                        "-k", "matches_oracle or on_part_of_the_gpu"], env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
