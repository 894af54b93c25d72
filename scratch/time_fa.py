"""f-1 / f-3 timing, host call to host result, with the CPU restatement (nanoflann 1-NN where oracle/_ref is built)
timed beside it on the same inputs. scratch/ measurement helper: it imports oracle/ only as the timed CPU baseline."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
abi = importlib.import_module("lego-loam-sr_b200.abi")
synth = importlib.import_module("lego-loam-sr_b200.synth")
if os.environ.get("LLB200_LIB"):  # A/B against another build of the library
    abi.LIB_PATH = os.environ["LLB200_LIB"]
import orc

def p50(fn, n=100, warm=10):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))

d = synth.make_sequence(seed=3, n_scans=3, step=0.7)
tc = np.array([0.004, -0.006, 0.003, 0.05, -0.02, 0.30], np.float32)
with abi.Handle(abi.default_params()) as h:
    for which, key, sub in ((0, "corner_scan", 3), (1, "surf_last_ds", 4)):
        last = np.ascontiguousarray(d["frames"][1][key], np.float32)
        last = last[np.argsort(last[:, 3].astype(np.int32), kind="stable")]
        q = np.ascontiguousarray(d["frames"][2][key], np.float32)[::sub].copy()
        h.fa_set_last(which, last)
        t_set = p50(lambda: h.fa_set_last(which, last))
        t_gpu = p50(lambda: h.fa_find_correspondences(which, q, tc, 0))
        ind = h.fa_find_correspondences(which, q, tc, 0)[:3]
        t_gpu_reuse = p50(lambda: h.fa_find_correspondences(which, q, tc, 1, ind=ind))
        nf = orc.have_ref()
        t_cpu = p50(lambda: orc.fa_correspondences(which, q, last, tc, 0, use_nanoflann=nf), n=20, warm=2)
        print(f"f-3 {'corner' if which == 0 else 'surf'}: {len(q)} feature points vs {len(last)} last-sweep points: "
              f"ll_fa_set_last {t_set:.0f} us; ll_fa_find_correspondences search iteration {t_gpu:.0f} us, reuse iteration "
              f"{t_gpu_reuse:.0f} us; CPU ({'nanoflann' if nf else 'brute force'} 1-NN, tree build included) {t_cpu:.0f} us", flush=True)
    rng = np.random.default_rng(0)
    for n in (400, 4000):
        rows = rng.normal(0, 1, (n, 6)).astype(np.float32)
        b = rng.normal(0, 0.05, n).astype(np.float32)
        for mask, name in ((0b010101, "calculateTransformationSurf (3 dof)"), (0b111111, "calculateTransformation (6 dof)")):
            t_gpu = p50(lambda: h.solve_normal_eq(rows, b, mask, 0, 10.0))
            t_cpu = p50(lambda: orc.solve_normal_eq(rows, b, mask, 0, 10.0, 1), n=50, warm=5)
            print(f"f-1 {name}, {n} rows: ll_solve_normal_eq {t_gpu:.0f} us; CPU restatement {t_cpu:.0f} us", flush=True)
