"""A few f-3 search iterations (surf, then corner) for ncu captures of fa_corr_kernel."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
abi = importlib.import_module("lego-loam-sr_b200.abi")
synth = importlib.import_module("lego-loam-sr_b200.synth")
d = synth.make_sequence(seed=3, n_scans=3, step=0.7)
tc = np.array([0.004, -0.006, 0.003, 0.05, -0.02, 0.30], np.float32)
with abi.Handle(abi.default_params()) as h:
    for which, key, sub in ((1, "surf_last_ds", 4), (0, "corner_scan", 3)):
        last = np.ascontiguousarray(d["frames"][1][key], np.float32)
        last = last[np.argsort(last[:, 3].astype(np.int32), kind="stable")]
        q = np.ascontiguousarray(d["frames"][2][key], np.float32)[::sub].copy()
        h.fa_set_last(which, last)
        for _ in range(3):
            out = h.fa_find_correspondences(which, q, tc, 0)
        print(which, len(q), len(last), int(out[4].sum()))
