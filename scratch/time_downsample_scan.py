"""downsampleCurrentScan on one VLP-16 frame, host clouds in and out: six ll_voxel_downsample calls against one
ll_downsample_scan; and the raw-frame mirror (llh_process_frame, pre_ds=0) per frame."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
abi = importlib.import_module("lego-loam-sr_b200.abi")
synth = importlib.import_module("lego-loam-sr_b200.synth")
host = importlib.import_module("lego-loam-sr_b200.host")
world = synth.World(seed=2)
gt = synth.trajectory(world, 60, step=0.7, seed=2)
raw = synth.scan(world, gt[0], seed=1)
cs, sl, ol = (np.ascontiguousarray(raw[k][:, :4], np.float32) for k in ("corner_scan", "surf_last", "outlier_last"))
clouds = [cs, sl, cs, sl, ol]
print("points:", [len(c) for c in clouds])
prm = abi.default_params()
leaf = (prm.corner_leaf, prm.surf_leaf, prm.corner_leaf, prm.surf_leaf, prm.surf_leaf)
with abi.Handle(prm) as h:
    def six():
        ds = [h.voxel_downsample(c, l, want_member=False)[0] for c, l in zip(clouds, leaf)]
        ds.append(h.voxel_downsample(np.concatenate([ds[1], ds[4]]), prm.surf_leaf, want_member=False)[0])
        return ds
    def one():
        return h.downsample_scan(clouds)[0]
    def one_dev():
        return h.downsample_scan(clouds, want_host=False)
    for name, fn in (("six ll_voxel_downsample calls", six), ("one ll_downsample_scan (host results)", one),
                     ("one ll_downsample_scan (device results only)", one_dev)):
        for _ in range(20):
            fn()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        print(f"{name}: p50 {1e6*np.median(ts):.1f} us", flush=True)
mo = host.MapOptimization(abi.default_params(apply_update=1, eigen_mode=1, iter_count_thres=10, stop_thres=0.0, lm_solver=1),
                          loop_closure_enabled=True, search_num=50, search_radius=50.0)
ts = []
for k in range(60):
    r = synth.scan(world, gt[k], seed=100 + k)
    t0 = time.perf_counter(); mo.process_frame(gt[k].astype(np.float32), r, pre_downsampled=False); ts.append(time.perf_counter() - t0)
print(f"raw-frame mirror: p50 {1e6*np.median(ts[10:]):.1f} us per frame (downsampleCurrentScan + map build + scan + keyframe)")
mo.close()
