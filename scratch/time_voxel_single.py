"""Time ll_voxel_downsample (no membership output) on downsampleCurrentScan-sized clouds: the single cooperative
launch against the stand-alone pipeline (LLB200_NO_INLINE_VOXEL=1). Host buffers in, host buffers out."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
abi = importlib.import_module("lego-loam-sr_b200.abi")
rng = np.random.default_rng(0)
for n, leaf in ((1200, 0.2), (5000, 0.4), (20000, 0.4), (120000, 0.4)):
    cloud = np.concatenate([rng.uniform(-40, 40, (n, 2)), rng.uniform(-2, 4, (n, 1)), np.zeros((n, 1))], 1).astype(np.float32)
    with abi.Handle(abi.default_params()) as h:
        for _ in range(20):
            h.voxel_downsample(cloud, leaf, want_member=False)
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); h.voxel_downsample(cloud, leaf, want_member=False); ts.append(time.perf_counter() - t0)
        print(f"n={n} leaf={leaf} inline={'LLB200_NO_INLINE_VOXEL' not in os.environ} p50={1e6*np.median(ts):.1f} us", flush=True)
