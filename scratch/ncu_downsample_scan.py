"""A few ll_downsample_scan calls on one VLP-16 frame (for the ncu launch list of the f-2 path)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
abi = importlib.import_module("lego-loam-sr_b200.abi")
synth = importlib.import_module("lego-loam-sr_b200.synth")
world = synth.World(seed=2)
gt = synth.trajectory(world, 2, step=0.7, seed=2)
raw = synth.scan(world, gt[0], seed=1)
cs, sl, ol = (np.ascontiguousarray(raw[k][:, :4], np.float32) for k in ("corner_scan", "surf_last", "outlier_last"))
with abi.Handle(abi.default_params()) as h:
    for _ in range(5):
        outs, dev = h.downsample_scan([cs, sl, cs, sl, ol])
    print([len(o) for o in outs], "launches:", h.launch_count)
