#!/usr/bin/env python
"""Benchmark of the scan-to-map registration hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload replay|single|dense]

Workload at N = 1 (BASELINE.json configs[1]): replay of a synthetic VLP-16 trajectory (10 Hz, ground plane +
walls + poles) through MapOptimization's per-frame hot path with a sliding local map of the last 50
keyframes: extractSurroundingKeyFrames (transform + concatenate + VoxelGrid 0.2/0.4 + search-grid build),
scan2MapOptimization (10 LM iterations of 5-NN + line/plane fit + normal equations + 6x6 solve) and the
keyframe store. One STEP = one scan registered. 50 scans are registered first to fill the window
(untimed set-up), then W warm-up scans, then K timed scans.

  value : scans/s with every input cloud already resident in HBM (device pointers through the C ABI).
  e2e   : the same through the reference-facing C++ mirror with HOST clouds: pack + H2D of the scan's five
          clouds and D2H of the result inside the timed region.
  N > 1 : one process per GPU (torchrun), each rank replays its own sequence (weak scaling); poses are
          gathered with NCCL after the loop; value = all ranks' scans / max-over-ranks device time.
  --impl reference : the CPU path (oracle restatement + the reference's vendored nanoflann) on the host cores.

Timing: CUDA events on the library's own stream around every step; L2 is flushed (256 MiB memset) before each
timed step; SM clocks and throttle reasons are sampled with nvidia-smi during the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "scan-to-map registrations/sec (VLP-16 scan vs sliding 50-keyframe local map)"
UNIT = "scans/s"
BENCH_WORLD = dict(n_blocks=60, n_poles=600, half_x=80, half_z=120)
BENCH_STEP_M = 1.0   # metres travelled per scan (10 Hz at 36 km/h) so that 50 keyframes span ~50 m
WINDOW = 50          # surroundingKeyframeSearchNum, loam_config.yaml:27
ITERS = 10           # LM iterations per scan (exactly 10, stop_thres = 0)
CACHE_DIR = os.environ.get("LLB200_CACHE", "/tmp/llb200_cache")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="replay", choices=["replay", "batch"],
                    help="replay = configs[1] (default, the headline); batch = configs[4]: independent sequences "
                         "partitioned across ranks, several streams per GPU")
    ap.add_argument("--sequences", type=int, default=64)
    ap.add_argument("--scans-per-seq", type=int, default=40)
    ap.add_argument("--threads", type=int, default=8, help="concurrent sequences (host threads / streams) per GPU")
    ap.add_argument("--batch-blocks", type=int, default=-1,
                    help="batch workload: > 0 = every sequence runs the persistent scan kernel on at most this many SMs "
                         "(lm_max_blocks); 0 = one launch per LM iteration (persistent_lm = 0); -1 = a quarter of the SMs "
                         "(measured best on B200: 37 blocks x 8 sequences in flight)")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-at-scale", action="store_true", help="skip the configs[2]/[3]-shaped per-kernel measurements")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="wall budget of the cpu_baseline sample")
    ap.add_argument("--seed", type=int, default=2)
    return ap.parse_args()


def headline_params(mod, **device_kw):
    p = mod.default_params(apply_update=1, eigen_mode=1, iter_count_thres=ITERS, stop_thres=0.0)
    if hasattr(p, "lm_solver"):  # execution knobs of the CUDA library (the oracle has no such choice)
        p.lm_solver = 1          # 6x6 normal equations by LDL^T in double
        for k, v in device_kw.items():
            setattr(p, k, v)
    return p


def make_data(seed, n_scans):
    from lego_loam_sr_b200 import synth
    return synth.make_sequence(seed=seed, n_scans=n_scans, step=BENCH_STEP_M, world_kw=BENCH_WORLD, cache_dir=CACHE_DIR)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- GPU arm
def frame_bytes(f):
    return 16 * sum(len(f[k]) for k in ("corner_scan", "corner_scan_ds", "surf_last_ds", "outlier_last_ds", "surf_total_ds"))


def run_gpu(data, device, prefill, warmup, steps, resident, flush, profile=False):
    """Replay prefill + warmup + steps frames; time the last `steps`. Returns a dict of measurements."""
    import torch
    from lego_loam_sr_b200 import abi, host
    torch.cuda.set_device(device)
    mo = host.MapOptimization(headline_params(abi), device=device, loop_closure_enabled=True, search_num=WINDOW)
    ext = torch.cuda.ExternalStream(mo.stream, device=device)
    view = mo.abi_handle()
    frames, odom = data["frames"], data["odom"]
    total = prefill + warmup + steps
    dev_frames = None
    if resident:
        keys = ("corner_scan", "corner_scan_ds", "surf_last_ds", "outlier_last_ds", "surf_total_ds")
        dev_frames = []
        for f in frames[:total]:
            ts = [torch.from_numpy(np.ascontiguousarray(f[k], np.float32)).to("cuda:%d" % device) for k in keys]
            dev_frames.append((ts, [(t.data_ptr() if t.numel() else 0, t.shape[0]) for t in ts]))
        torch.cuda.synchronize(device)
    host_frames = None
    if not resident:
        # the reference's clouds are pcl::PointCloud<PointXYZI>: 32-byte records read in place by the library
        keys = ("corner_scan", "corner_scan_ds", "surf_last_ds", "outlier_last_ds", "surf_total_ds")
        host_frames = []
        for f in frames[:total]:
            arrs = []
            for k in keys:
                a = np.zeros((len(f[k]), 8), np.float32)
                a[:, :3] = f[k][:, :3]
                a[:, 3] = 1.0
                a[:, 4] = f[k][:, 3]
                arrs.append(a)
            host_frames.append(arrs)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % device) if flush else None
    step_ms, wall_ms, results, poses = [], [], [], []
    launches0 = 0
    for k in range(total):
        timed = k >= prefill + warmup
        if k == prefill + warmup:
            launches0 = mo.launch_count
            if profile:
                view.profile_enable(True)
                view.profile_read()
        if timed:
            if flush_buf is not None:
                flush_buf.zero_()
                torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(ext)
        if resident:
            pose, res = mo.process_frame_dev(odom[k], dev_frames[k][1])
        else:
            pose, res = mo.process_frame_host(odom[k], host_frames[k])
        if timed:
            e1.record(ext)
            e1.synchronize()
            wall_ms.append((time.perf_counter() - t0) * 1e3)
            step_ms.append(e0.elapsed_time(e1))
            results.append((res.status, res.iter_num, res.sel_num, res.corner_map_ds, res.surf_map_ds))
            poses.append(pose.copy())
    out = dict(step_ms=step_ms, wall_ms=wall_ms, results=results, poses=np.array(poses, np.float32),
               launches=mo.launch_count - launches0)
    if profile:
        out["profile"] = view.profile_read()
        view.profile_enable(False)
    # raw (pre-VoxelGrid) sizes of the timed steps' local maps, for the algorithmic-bytes model
    raw_c, raw_s = [], []
    for k in range(prefill + warmup, total):
        win = frames[max(0, k - WINDOW):k]
        raw_c.append(sum(len(f["corner_scan"]) for f in win))
        raw_s.append(sum(len(f["surf_last_ds"]) + len(f["outlier_last_ds"]) for f in win))
    out["raw_corner"], out["raw_surf"] = raw_c, raw_s
    out["queries"] = [len(frames[k]["corner_scan_ds"]) + len(frames[k]["surf_total_ds"]) for k in range(prefill + warmup, total)]
    out["h2d_bytes"] = [frame_bytes(frames[k]) for k in range(prefill + warmup, total)]
    mo.close()
    return out


def run_batch(args, rank, world, device):
    """BASELINE configs[4]: `--sequences` independent VLP-16 sequences x `--scans-per-seq` scans, seq -> rank =
    seq mod world, `--threads` sequences in flight per GPU (one handle + stream + host thread each), NCCL only
    to gather the result records. Host clouds (e2e path)."""
    import torch
    import torch.distributed as dist
    from lego_loam_sr_b200 import abi, batch, host, synth
    ids = batch.shard_sequences(args.sequences, world, rank)
    seqs = {}
    for sid in ids:
        d = synth.make_sequence(seed=100 + sid, n_scans=args.scans_per_seq, step=BENCH_STEP_M, world_kw=BENCH_WORLD,
                                cache_dir=CACHE_DIR)
        keys = ("corner_scan", "corner_scan_ds", "surf_last_ds", "outlier_last_ds", "surf_total_ds")
        d["host"] = []
        for f in d["frames"]:
            arrs = []
            for k in keys:
                a = np.zeros((len(f[k]), 8), np.float32)
                a[:, :3] = f[k][:, :3]
                a[:, 3] = 1.0
                a[:, 4] = f[k][:, 3]
                arrs.append(a)
            d["host"].append(arrs)
        seqs[sid] = d

    def make_node():
        # several handles share the GPU: one launch per LM iteration so that their kernels interleave
        nb = args.batch_blocks if args.batch_blocks >= 0 else max(1, torch.cuda.get_device_properties(device).multi_processor_count // 4)
        kw = dict(persistent_lm=0) if nb == 0 else dict(persistent_lm=1, lm_max_blocks=nb)
        return host.MapOptimization(headline_params(abi, **kw), device=device, loop_closure_enabled=True, search_num=WINDOW)

    def feed(node, data, k):
        return node.process_frame_host(data["odom"][k], data["host"][k])

    warm = {ids[0]: seqs[ids[0]]} if ids else {}
    batch.replay_sequences(warm, make_node, feed, n_threads=1)  # warm-up: allocator, module load
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    nodes = {sid: make_node() for sid in ids}  # handle creation (cudaMalloc) stays outside the timed region
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    recs = batch.replay_sequences(seqs, make_node, feed, n_threads=args.threads, nodes=nodes)
    torch.cuda.synchronize(device)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda:%d" % device)
    for n in nodes.values():
        n.close()
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    allrec = batch.gather_records(recs, args.sequences, args.scans_per_seq, world, rank, dist=dist if world > 1 else None,
                                  device="cuda:%d" % device)
    return float(dt[0]), allrec


def algorithmic_bytes(cls, n_raw, n_ds, q):
    """Algorithmic bytes ONE launch of a kernel class moves (packed 16-byte points, 4-byte keys and
    values). n_raw / n_ds count corner + surf points together (both maps go through the same launch)."""
    return {
        "transform_minmax": 32 * n_raw,             # read sensor-frame cloud, write map-frame cloud
        "voxel_keys": 24 * n_raw,                   # read 16, write key + value
        "voxel_sort_pass": 16 * n_raw,              # read + write (key, value)
        "voxel_reduce": 24 * n_raw + 16 * n_ds,     # read pairs, gather points, write centroids
        "grid_keys": 24 * n_ds,
        "grid_sort_pass": 16 * n_ds,
        "grid_finalize": 24 * n_ds + 16 * n_ds,
        "lm_begin": 0,
        "lm_iter": 96 * q * ITERS,                  # SURVEY.md §8d B_iter (query + five neighbours, 16 B each) x the
                                                    # iterations one persistent launch runs
    }[cls]


def roofline_from_profile(prof, run, peaks):
    n = len(run["results"])
    n_raw = sum(run["raw_corner"]) / n + sum(run["raw_surf"]) / n
    n_ds = sum(r[3] + r[4] for r in run["results"]) / n
    q = sum(run["queries"]) / n
    total_ms = sum(ms for ms, _ in prof.values()) or 1e-9
    table = {}
    # single-kernel scan: the map assembly ran inside the one launch of the step, so that launch moves B_scan =
    # B_voxel + B_grid + iterations * B_iter, not just the LM part
    whole_scan = not prof.get("voxel_keys", (0, 0))[1]
    for cls, (ms, cnt) in prof.items():
        if not cnt:
            continue
        per_launch_us = ms * 1e3 / cnt
        b = algorithmic_bytes(cls, n_raw, n_ds, q)
        if cls == "lm_iter" and whole_scan:
            b += (16 * n_raw + 16 * n_ds) + 32 * n_ds
        table[cls] = dict(launches_per_step=cnt / n, us_per_launch=round(per_launch_us, 3), share=round(ms / total_ms, 4),
                          gbs=round(b / (per_launch_us * 1e-6) / 1e9, 2) if per_launch_us > 0 else None)
    top = max(table, key=lambda c: table[c]["share"])
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = table[top]["gbs"]
    traffic = None  # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture
    for name in ("r01c_ncu_scan_kernel_full.json" if top == "lm_iter" else "", "r01_ncu_%s_full.json" % top):
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", name)))
            traffic = int(cap["traffic_bytes_per_launch"])
            break
        except Exception:
            pass
    roof = {"kernel": "lm_persistent_kernel (the whole scan: transform, VoxelGrid, search grids, %d LM iterations)" % ITERS if top == "lm_iter" else top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 5), "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)",
            "algorithmic_bytes_per_launch": int(round(achieved * 1e9 * table[top]["us_per_launch"] * 1e-6)),
            "us_per_launch": table[top]["us_per_launch"], "share_of_step": table[top]["share"],
            "note": "one scan is latency-bound (10 dependent iterations, ~7 k queries): see DESIGN.md §5 and profiles/"}
    return roof, table, dict(n_raw=round(n_raw), n_ds=round(n_ds), queries=round(q))


def run_at_scale(device, peaks):
    """BASELINE configs[2]/[3]-shaped single problems (parity-test cases, not the headline): per-kernel-class device
    time and achieved algorithmic GB/s at sizes where bandwidth, not launch latency, is what bounds the kernels.
    c4: ~2.2 M surrounding-map points WITHOUT map downsampling, 20 k corner + 80 k surf queries.
    c3: 2.5 M raw surf + 0.2 M raw corner points WITH VoxelGrid, 8 k + 32 k queries (HDL-64E-like volume)."""
    from lego_loam_sr_b200 import abi
    rng = np.random.default_rng(4)

    def world(n):
        p = np.empty((n, 4), np.float32)
        p[:, 0] = rng.uniform(-70, 70, n)
        p[:, 2] = rng.uniform(-70, 70, n)
        p[:, 1] = np.where(rng.random(n) < 0.7, -1.8 + rng.normal(0, 0.02, n), rng.uniform(-1.8, 6, n))
        p[:, 3] = rng.uniform(0, 16, n)
        return p

    out = {}
    peak = peaks.get("hbm_gbs", 6650.0)
    for name, ds, n_surf, n_corner, nqc, nqs in (("c4_dense_no_downsample", 0, 2_000_000, 200_000, 20_000, 80_000),
                                                 ("c3_like_with_voxelgrid", 1, 2_500_000, 200_000, 8_000, 32_000)):
        surf, corner = world(n_surf), world(n_corner)
        prm = headline_params(abi, map_downsample=ds)
        def vox_order(q, leaf):  # downsampleCurrentScan's VoxelGrid emits points in ascending voxel index (z-major)
            k = np.floor(q[:, :3] / leaf).astype(np.int64)
            return np.ascontiguousarray(q[np.lexsort((k[:, 0], k[:, 1], k[:, 2]))])

        cq = vox_order(corner[rng.choice(len(corner), nqc, replace=False)], 0.2)
        sq = vox_order(surf[rng.choice(len(surf), nqs, replace=False)], 0.4)
        pose = np.array([0.002, -0.003, 0.002, 0.03, -0.02, 0.03], np.float32)
        with abi.Handle(prm, device=device) as h:
            for r in range(4):
                if r == 1:
                    h.profile_enable(True)
                    h.profile_read()
                h.set_local_map_raw(corner, surf)
                _, res = h.scan2map(cq, sq, pose)
            prof = h.profile_read()
        n_raw, n_ds, q = n_surf + n_corner, res.corner_map_ds + res.surf_map_ds, nqc + nqs
        table = {}
        for cls, (ms, cnt) in prof.items():
            if not cnt:
                continue
            us = ms * 1e3 / cnt
            b = algorithmic_bytes(cls, n_raw, n_ds, q)
            table[cls] = {"us_per_launch": round(us, 1), "gbs": round(b / (us * 1e-6) / 1e9, 1), "frac_of_hbm_peak": round(b / (us * 1e-6) / 1e9 / peak, 4)}
        out[name] = {"n_raw": n_raw, "n_ds": int(n_ds), "queries": q, "iterations": int(res.iter_num), "residuals": int(res.sel_num),
                     "kernels": table}
    return out


# ----------------------------------------------------------------------------------------------- CPU arm
def run_next_rows(with_cpu):
    """SURVEY.md §8 "next" rows (not the headline): p50 microseconds per call, host buffers in, host results out, on one
    synthetic VLP-16 frame; with_cpu adds the CPU restatement of the same call (the bench's CPU leg) beside each."""
    from lego_loam_sr_b200 import abi, synth

    def p50(fn, n=60, warm=8):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return round(1e6 * statistics.median(ts), 1)

    world = synth.World(seed=2)
    gt = synth.trajectory(world, 2, step=0.7, seed=2)
    raw = [synth.scan(world, gt[k], seed=1 + k) for k in range(2)]
    cs, sl, ol = (np.ascontiguousarray(raw[0][k][:, :4], np.float32) for k in ("corner_scan", "surf_last", "outlier_last"))
    clouds = [cs, sl, cs, sl, ol]
    prm = abi.default_params()
    leaf = (prm.corner_leaf, prm.surf_leaf, prm.corner_leaf, prm.surf_leaf, prm.surf_leaf)
    tc = np.array([0.004, -0.006, 0.003, 0.05, -0.02, 0.30], np.float32)
    out = {"unit": "us per call (p50), host buffers both ways"}
    with abi.Handle(prm) as h:
        ds = h.downsample_scan(clouds)[0]
        out["f2_downsample_scan"] = {"points_in": [len(c) for c in clouds], "points_out": [len(d) for d in ds],
                                     "gpu": p50(lambda: h.downsample_scan(clouds)),
                                     "gpu_device_results": p50(lambda: h.downsample_scan(clouds, want_host=False))}
        fa = {}
        for which, name, cur, prev in ((0, "corner", ds[2], h.voxel_downsample(np.ascontiguousarray(raw[1]["corner_scan"][:, :4], np.float32), prm.corner_leaf, want_member=False)[0]),
                                       (1, "surf", ds[1][::4].copy(), h.voxel_downsample(np.ascontiguousarray(raw[1]["surf_last"][:, :4], np.float32), prm.surf_leaf, want_member=False)[0])):
            last = prev[np.argsort(prev[:, 3].astype(np.int32), kind="stable")]
            h.fa_set_last(which, last)
            ind = h.fa_find_correspondences(which, cur, tc, 0)[:3]
            fa[name] = {"feature_points": len(cur), "last_points": len(last),
                        "gpu_search_iteration": p50(lambda: h.fa_find_correspondences(which, cur, tc, 0)),
                        "gpu_reuse_iteration": p50(lambda: h.fa_find_correspondences(which, cur, tc, 1, ind=ind)),
                        "_inputs": (cur, last)}
        out["f3_fa_correspondences"] = fa
        rng = np.random.default_rng(0)
        rows, b = rng.normal(0, 1, (1500, 6)).astype(np.float32), rng.normal(0, 0.05, 1500).astype(np.float32)
        out["f1_solve_normal_eq"] = {"rows": 1500, "gpu_3dof": p50(lambda: h.solve_normal_eq(rows, b, 0b010101, 0, 10.0)),
                                     "gpu_6dof": p50(lambda: h.solve_normal_eq(rows, b, 0b111111, 0, 10.0))}
    if with_cpu:
        import orc
        orc.build()

        def chain():
            d = [orc.voxel_downsample(c, l)[0] for c, l in zip(clouds, leaf)]
            orc.voxel_downsample(np.concatenate([d[1], d[4]]), prm.surf_leaf)
        out["f2_downsample_scan"]["cpu"] = p50(chain, n=10, warm=2)
        nf = orc.have_ref()
        for which, name in ((0, "corner"), (1, "surf")):
            cur, last = fa[name]["_inputs"]
            fa[name]["cpu_search_iteration"] = p50(lambda: orc.fa_correspondences(which, cur, last, tc, 0, use_nanoflann=nf), n=10, warm=2)
        out["f1_solve_normal_eq"]["cpu_3dof"] = p50(lambda: orc.solve_normal_eq(rows, b, 0b010101, 0, 10.0, 1), n=20, warm=2)
        out["f1_solve_normal_eq"]["cpu_6dof"] = p50(lambda: orc.solve_normal_eq(rows, b, 0b111111, 0, 10.0, 1), n=20, warm=2)
    for v in fa.values():
        v.pop("_inputs", None)
    return out


def run_cpu(data, prefill, warmup, steps, budget_s):
    """The reference's CPU path (oracle restatement; k-NN = the reference's vendored nanoflann when oracle/_ref was
    built), single thread like the reference's _run_thread (mapOptmization.cpp:175). Times `steps` scans after
    prefill + warmup untimed ones, or as many as fit in budget_s."""
    import orc
    orc.build()
    knn = "nanoflann (reference, oracle/_ref)" if orc.have_ref() else "oracle uniform grid"
    orc.set_knn_backend(orc.KNN_NANOFLANN if orc.have_ref() else orc.KNN_GRID)
    rp = orc.Replay(headline_params(orc), window=WINDOW)
    frames, odom = data["frames"], data["odom"]
    times = []
    t_start = time.perf_counter()
    for k in range(min(prefill + warmup + steps, len(frames))):
        f = frames[k]
        t0 = time.perf_counter()
        rp.step(odom[k], f["corner_scan"], f["corner_scan_ds"], f["surf_last_ds"], f["outlier_last_ds"], f["surf_total_ds"])
        dt = time.perf_counter() - t0
        if k >= prefill + warmup:
            times.append(dt)
            if time.perf_counter() - t_start > budget_s and len(times) >= 5:
                break
    orc.set_knn_backend(orc.KNN_GRID)
    rp.close()
    return times, knn


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    prefill = WINDOW
    n_scans = prefill + args.warmup + args.steps
    config = {"workload": "configs[1]: synthetic VLP-16 trajectory replay, sliding local map of the last %d keyframes, "
                          "%d LM iterations per scan" % (WINDOW, ITERS),
              "sensor": "VLP-16 16x1800", "metres_per_scan": BENCH_STEP_M, "window_keyframes": WINDOW,
              "lm_iterations": ITERS, "apply_update": 1, "eigen_mode": "upstream", "lm_solver": "ldlt-f64", "persistent_lm": 1, "prefill_scans": prefill,
              "sequences_per_gpu": 1, "l2": "flushed before every timed step (256 MiB memset)" if not args.no_l2_flush else "not flushed"}

    if args.impl == "reference":
        if rank != 0:
            return
        data = make_data(args.seed, n_scans)
        times, knn = run_cpu(data, prefill, args.warmup, args.steps, budget_s=150.0)
        v = len(times) / sum(times)
        sample = "%d consecutive scans of the same replay after %d untimed scans filled the window; k-NN = %s" % (len(times), prefill + args.warmup, knn)
        line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / v, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "p50_ms_per_scan": round(statistics.median(times) * 1e3, 3),
                "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                                 "host_cores_available": host_cores(),
                                 "why_one_core": "one registration stream is sequential; the reference runs it on one thread (mapOptmization.cpp:175)"},
                "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # Keep NCCL's own output (the "NCCL version ..." banner is printed at every debug level from VERSION up, to stdout)
    # off stdout: the driver reads ONE JSON line there.
    # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION is raised to WARN.)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    import __graft_entry__ as ge
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    device = local_rank
    if args.workload == "batch":
        secs, allrec = run_batch(args, rank, world, device)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            total = args.sequences * args.scans_per_seq
            ok = int((allrec[:, 1:, 10] == 0).all())
            print(json.dumps({"metric": METRIC, "value": round(total / secs, 2), "unit": UNIT, "n_gpus": world,
                              "steps": args.scans_per_seq, "warmup": 0, "ms_per_step": round(1e3 * secs / total, 4),
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic",
                              "config": {"workload": "configs[4]: batched replay of %d independent VLP-16 sequences x %d scans, "
                                                     "seq -> rank = seq mod world, %d sequences in flight per GPU (persistent scan kernels on a quarter "
                                                     "of the SMs each unless --batch-blocks says otherwise), NCCL only gathers the result records" % (args.sequences, args.scans_per_seq, args.threads),
                                         "timing": "wall clock between device synchronisations, max over ranks (concurrent streams)"},
                              "checks": {"all_scans_optimised": ok, "records": list(allrec.shape)}}))
        return
    data = make_data(args.seed + rank, n_scans)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    flush = not args.no_l2_flush

    def sync_all():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(device)
    # ---- e2e leg: host clouds through the C++ mirror (SM clocks / throttle reasons are sampled over both timed legs)
    sync_all()
    sampler.start()
    e2e = run_gpu(data, device, prefill, args.warmup, args.steps, resident=False, flush=flush)
    # ---- value leg: inputs resident in HBM
    sync_all()
    res = run_gpu(data, device, prefill, args.warmup, args.steps, resident=True, flush=flush)
    sync_all()
    clocks = sampler.stop()
    # gather per-rank device times and poses (NCCL carries results only)
    t_res = torch.tensor([sum(res["step_ms"]), sum(e2e["step_ms"])], dtype=torch.float64, device="cuda:%d" % device)
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        poses = torch.from_numpy(res["poses"]).to("cuda:%d" % device)
        gathered = [torch.empty_like(poses) for _ in range(world)]
        dist.all_gather(gathered, poses)
    total_ms, total_e2e_ms = float(t_res[0]), float(t_res[1])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- per-kernel timing pass (rank 0): same steps with CUDA events around every launch
    prof_run = run_gpu(data, device, prefill, args.warmup, args.steps, resident=True, flush=flush, profile=True)
    roof, table, sizes = roofline_from_profile(prof_run["profile"], prof_run, peaks)
    value = world * args.steps / (total_ms * 1e-3)
    e2e_value = world * args.steps / (total_e2e_ms * 1e-3)
    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "p50_ms_per_scan": round(statistics.median(res["step_ms"]), 4),
            "p50_ms_per_scan_e2e": round(statistics.median(e2e["step_ms"]), 4),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(np.mean(e2e["h2d_bytes"])),
                    "d2h_bytes_per_step": 484, "wall_ms_per_step": round(statistics.mean(e2e["wall_ms"]), 4)},
            "gpu_launches": int(res["launches"]), "gpu_launches_per_step": round(res["launches"] / args.steps, 2),
            "clocks": clocks, "roofline": roof, "kernels": table, "sizes": sizes,
            "checks": {"all_scans_optimised": all(r[0] == 0 for r in res["results"]),
                       "iterations_per_scan": sorted(set(r[1] for r in res["results"])),
                       "resident_equals_e2e_pose": bool(np.array_equal(res["poses"], e2e["poses"]))}}
    if world == 1 and not args.no_at_scale:
        line["at_scale"] = run_at_scale(device, peaks)
    if world == 1 and not args.no_at_scale:
        try:
            line["next_rows"] = run_next_rows(with_cpu=not args.no_cpu_baseline)
        except Exception as e:  # (never lose the headline line to a side measurement)
            line["next_rows"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if world == 1 and not args.no_cpu_baseline:
        times, knn = run_cpu(data, prefill, 0, args.steps, budget_s=args.cpu_seconds)
        v = len(times) / sum(times)
        line["cpu_baseline"] = {"value": round(v, 3), "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "%d consecutive scans of the same replay after %d untimed scans filled the window; "
                                          "restated PCL/Eigen pieces, k-NN = %s" % (len(times), prefill, knn),
                                "p50_ms_per_scan": round(statistics.median(times) * 1e3, 3), "host_cores_available": host_cores()}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
