#!/usr/bin/env python
"""Benchmark of the scan-to-map registration hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload replay|batch]

Workload at N = 1 (BASELINE.json configs[1], on the metric's stated sizes: a VLP-16 scan of ~2 k corner + ~8 k surf
queries against a ~100 k-point local map): replay of a synthetic VLP-16 trajectory (10 Hz, ground plane + walls +
poles) through MapOptimization's per-frame hot path with a sliding local map of the last 50 keyframes:
extractSurroundingKeyFrames (transform + concatenate + VoxelGrid 0.2/0.4 + search-grid build), scan2MapOptimization
(10 LM iterations of 5-NN + line/plane fit + normal equations + 6x6 solve) and the keyframe store. One STEP = one scan
registered. 50 scans are registered first to fill the window (untimed set-up), then W warm-up scans, then K timed scans.

  value : scans/s with every input cloud already resident in HBM (device pointers through the C ABI); CUDA events.
  e2e   : the same through the reference-facing C++ mirror with HOST pcl::PointXYZI clouds in page-locked memory (the
          contract's "inputs in pinned host memory"; BENCH_PAGEABLE=1 for pageable ones): H2D of the scan's five clouds
          (the two query clouds as the 32-byte records they are, the three keyframe clouds packed) and D2H of the result
          inside the timed region; WALL clock per step (the call a user makes).
  checks.parity : the timed frames replayed once more with the oracle registering the same scan against the device's
          own map and prior; the line is marked failed when a pose differs by more than 1e-4 m / 1e-5 rad, or when the
          checked replay's poses are not bit-identical to the timed legs'.
  N > 1 : one process per GPU (torchrun), each rank replays its own sequence (weak scaling); poses are gathered with
          NCCL after the loop; value = all ranks' scans / max-over-ranks time. A `batch` record (configs[4]: 64
          sequences partitioned over the ranks, windows pre-filled untimed, C++ host threads) is appended.
  --impl reference : the CPU path (oracle restatement + the reference's vendored nanoflann) on the host cores, one
          process per requested GPU count (N independent replays on N cores, summed).

Extra records on the N = 1 line (none of them the headline): `raw_frame` (downsampleCurrentScan on the device too),
`as_shipped` (the fork's shipped behaviour: update commented out, 200 iterations), `configs2` / `configs3` (HDL-64E vs a
~500 k map; ~2 M-point map without downsampling and ~100 k queries), `next_rows`, `cpu_baseline`.

Timing: L2 is flushed (256 MiB memset) before each timed step; SM clocks and throttle reasons are sampled with
nvidia-smi during the timed region.
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

METRIC = "scan-to-map registrations/sec (VLP-16 scan vs sliding 50-keyframe local map of ~100 k points)"
UNIT = "scans/s"
# World and speed chosen so that the workload sits on BASELINE.json's stated sizes: ~1.8 k corner + ~7.7 k surf queries
# per scan, 50 keyframes -> ~0.59 M raw / ~118 k downsampled map points (round 1 ran on ~62 k / ~6.7 k: R01_WORLD).
BENCH_WORLD = dict(n_blocks=70, n_poles=1800, half_x=100, half_z=150, wall_height=20.0)
BENCH_STEP_M = 1.2   # metres travelled per scan (10 Hz at 43 km/h) so that 50 keyframes span ~60 m
R01_WORLD = dict(n_blocks=60, n_poles=600, half_x=80, half_z=120)
R01_STEP_M = 1.0
HDL64_WORLD = dict(n_blocks=150, n_poles=4000, half_x=160, half_z=240, wall_height=20.0)
WINDOW = 50          # surroundingKeyframeSearchNum, loam_config.yaml:27
ITERS = 10           # LM iterations per scan (exactly 10, stop_thres = 0)
TOL_T, TOL_R = 1e-4, 1e-5   # north_star: pose within 1e-4 m / 1e-5 rad of the reference after 10 LM iterations
CACHE_DIR = os.environ.get("LLB200_CACHE", "/tmp/llb200_cache")
CLOUD_KEYS = ("corner_scan", "corner_scan_ds", "surf_last_ds", "outlier_last_ds", "surf_total_ds")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="replay", choices=["replay", "batch"],
                    help="replay = configs[1] (default, the headline; a batch record is appended); batch = configs[4] only")
    ap.add_argument("--sequences", type=int, default=64)
    ap.add_argument("--scans-per-seq", type=int, default=24, help="timed scans per sequence (after the untimed window fill)")
    ap.add_argument("--threads", type=int, default=8, help="batch: host threads (= sequences in flight) per GPU")
    ap.add_argument("--batch-blocks", type=int, default=-1,
                    help="batch: > 0 = every sequence runs the persistent scan kernel on at most this many SMs "
                         "(lm_max_blocks); 0 = one launch per LM iteration (persistent_lm = 0); -1 = SMs / threads")
    ap.add_argument("--batch-dump", default="", help="batch: write the gathered records to this .npy (rank 0)")
    ap.add_argument("--batch-group", type=int, default=0,
                    help="batch: sequences each host thread keeps in flight (0 = 1, or more when the rank is short of cores)")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-at-scale", action="store_true", help="skip configs[2]/[3], raw-frame, as-shipped and next-row records")
    ap.add_argument("--no-batch", action="store_true", help="skip the appended batch record")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed frames")
    ap.add_argument("--legacy-world", action="store_true", help="round 1's smaller workload (62 k-point map, 6.7 k queries)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="wall budget of the cpu_baseline sample")
    ap.add_argument("--seed", type=int, default=2)
    return ap.parse_args()


def headline_params(mod, **device_kw):
    p = mod.default_params(apply_update=1, eigen_mode=1, iter_count_thres=ITERS, stop_thres=0.0)
    if hasattr(p, "lm_solver"):  # execution knobs of the CUDA library (the oracle has no such choice)
        p.lm_solver = 1          # 6x6 normal equations by LDL^T in double
    for k, v in device_kw.items():
        if hasattr(p, k):
            setattr(p, k, v)
    return p


_WORLD, _STEP = BENCH_WORLD, BENCH_STEP_M


def make_data(seed, n_scans, keep_raw=False):
    from lego_loam_sr_b200 import synth
    return synth.make_sequence(seed=seed, n_scans=n_scans, step=_STEP, world_kw=_WORLD, cache_dir=CACHE_DIR, keep_raw=keep_raw)


def make_hdl64_problem(kind):
    """configs[2] ("c2"): one HDL-64E scan (64 x 1800, downsampled queries) against the local map of 150 keyframes 4 m
    apart (~0.5 M points after VoxelGrid, ~3.5 M raw). configs[3] ("c3"): ~2 M raw surrounding-map points (85 keyframes,
    used WITHOUT map downsampling) and the RAW corner / surf clouds of one HDL-64E scan as ~100 k queries."""
    from lego_loam_sr_b200 import synth
    n = 150 if kind == "c2" else 85
    d = synth.make_single(seed=3, n_map_scans=n, step=4.0, sensor="hdl64", world_kw=HDL64_WORLD, cache_dir=CACHE_DIR)
    if kind == "c3":
        world = synth.World(seed=3, **HDL64_WORLD)
        raw = synth.scan(world, d["gt"].astype(np.float64), sensor="hdl64", seed=3 * 100003 + n)
        d["raw_query"] = (np.ascontiguousarray(raw["corner_scan"], np.float32),
                          np.ascontiguousarray(np.concatenate([raw["surf_last"], raw["outlier_last"]]), np.float32))
    return d


def to_pcl_layout(a):
    """[n, 4] packed -> [n, 8]: pcl::PointXYZI's 32-byte record (x, y, z, 1, intensity, pad x 3)."""
    out = np.zeros((len(a), 8), np.float32)
    out[:, :3] = a[:, :3]
    out[:, 3] = 1.0
    out[:, 4] = a[:, 3]
    return out


def to_pinned_pcl_layout(a):
    """to_pcl_layout in PAGE-LOCKED host memory (the e2e contract's "inputs in pinned host memory"): the library copies
    such clouds by DMA straight from the caller's buffer (include/legoloam_b200.h, ll_scan2map_async).
    BENCH_PAGEABLE=1 keeps them pageable (the library then packs them into its own staging memory) for A/B runs."""
    out = to_pcl_layout(a)
    if len(out) == 0 or os.environ.get("BENCH_PAGEABLE"):
        return out
    import torch
    return torch.from_numpy(out).pin_memory().numpy()  # (the array keeps the pinned tensor alive)


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
    """Host-to-device bytes of one e2e frame: the three keyframe clouds go up packed (16 bytes per point); the two query
    clouds as the 32-byte pcl::PointXYZI records they are (page-locked: copied in place), packed when BENCH_PAGEABLE is set."""
    q = 16 if os.environ.get("BENCH_PAGEABLE") else 32
    return sum((q if k in ("corner_scan_ds", "surf_total_ds") else 16) * len(f[k]) for k in CLOUD_KEYS)


def run_gpu(data, device, prefill, warmup, steps, mode, flush, profile=False, params_kw=None, param_override=None):
    """Replay prefill + warmup + steps frames; time the last `steps`. mode: "resident" (device pointers), "host"
    (pcl::PointXYZI host clouds read in place: the e2e leg), "raw" (raw host clouds: downsampleCurrentScan runs on the
    device as part of the frame; timed inside the C++ mirror). Returns a dict of measurements."""
    import torch
    from lego_loam_sr_b200 import abi, host
    torch.cuda.set_device(device)
    prm = param_override if param_override is not None else headline_params(abi, **(params_kw or {}))
    mo = host.MapOptimization(prm, device=device, loop_closure_enabled=True, search_num=WINDOW)
    ext = torch.cuda.ExternalStream(mo.stream, device=device)
    view = mo.abi_handle()
    frames, odom = data["frames"], data["odom"]
    total = prefill + warmup + steps
    dev_frames = host_frames = None
    if mode == "resident":
        dev_frames = []
        for f in frames[:total]:
            ts = [torch.from_numpy(np.ascontiguousarray(f[k], np.float32)).to("cuda:%d" % device) for k in CLOUD_KEYS]
            dev_frames.append((ts, mo.prepare_dev([(t.data_ptr() if t.numel() else 0, t.shape[0]) for t in ts])))
        torch.cuda.synchronize(device)
    elif mode == "host":
        # the reference's clouds are pcl::PointCloud<PointXYZI>: 32-byte records read in place by the library; the
        # (pointer, count) views are made once, outside the timed region — a C++ caller has them at hand
        host_frames = [mo.prepare_host([to_pinned_pcl_layout(f[k]) for k in CLOUD_KEYS]) for f in frames[:total]]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % device) if flush else None
    step_ms, wall_ms, frame_ms, results, poses = [], [], [], [], []
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
            e0.record(ext)
            t0 = time.perf_counter()  # (the wall clock brackets the API call alone: the call returns with the result on the host)
        if mode == "resident":
            pose, res = mo.process_prepared(odom[k], dev_frames[k][1])
        elif mode == "host":
            pose, res = mo.process_prepared(odom[k], host_frames[k])
        else:
            pose, res = mo.process_frame(odom[k], frames[k], pre_downsampled=False)
        if timed:
            t1 = time.perf_counter()
            e1.record(ext)
            e1.synchronize()
            wall_ms.append((t1 - t0) * 1e3)
            step_ms.append(e0.elapsed_time(e1))
            frame_ms.append(mo.last_frame_ms)
            results.append((res.status, res.iter_num, res.sel_num, res.corner_map_ds, res.surf_map_ds))
            poses.append(pose.copy())
    out = dict(step_ms=step_ms, wall_ms=wall_ms, frame_ms=frame_ms, results=results, poses=np.array(poses, np.float32),
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


def replay_parity(data, device, prefill, steps, every=1, params_kw=None):
    """The parity check behind `checks.parity` and tests/test_gpu_configs.py::test_bench_replay_parity: replay
    prefill + steps frames through the e2e path (host clouds, C++ mirror); for every `every`-th of the last `steps`
    frames (and the last one) fetch the device's own local map and the prior the scan started from, let the ORACLE
    register the same scan on them, and compare the poses (strict form: identical inputs on both sides)."""
    import orc
    from lego_loam_sr_b200 import abi, host
    orc.build()
    orc.set_knn_backend(orc.KNN_GRID)
    oprm = headline_params(orc)
    mo = host.MapOptimization(headline_params(abi, **(params_kw or {})), device=device, loop_closure_enabled=True, search_num=WINDOW)
    view = mo.abi_handle()
    frames, odom = data["frames"], data["odom"]
    poses, checked, worst_t, worst_r, maps, qs, sel_diff = [], 0, 0.0, 0.0, [], [], 0
    for k in range(prefill + steps):
        f = frames[k]
        pose, res = mo.process_frame_host(odom[k], [to_pcl_layout(f[c]) for c in CLOUD_KEYS])
        if k < prefill:
            continue
        poses.append(pose.copy())
        i = k - prefill
        if i % every and i != steps - 1:
            continue
        gc, gs = view.get_map_ds(0), view.get_map_ds(1)
        ores, opose, _ = orc.scan2map(gc, gs, f["corner_scan_ds"], f["surf_total_ds"], mo.last_prior, oprm)
        d = np.abs(np.asarray(pose, np.float64) - np.asarray(opose, np.float64))
        d[:3] = np.minimum(d[:3], np.abs(d[:3] - 2 * np.pi))
        worst_r, worst_t = max(worst_r, float(d[:3].max())), max(worst_t, float(d[3:].max()))
        sel_diff = max(sel_diff, abs(int(res.sel_num) - int(ores.sel_num)))
        ok_state = res.status == ores.status == 0 and res.iter_num == ores.iter_num
        if not ok_state:
            worst_t = worst_r = float("inf")
        maps.append(len(gc) + len(gs))
        qs.append(len(f["corner_scan_ds"]) + len(f["surf_total_ds"]))
        checked += 1
    mo.close()
    return dict(frames_checked=checked, of_timed_frames=steps, max_dt_m=worst_t, max_dr_rad=worst_r, tol_dt_m=TOL_T, tol_dr_rad=TOL_R,
                max_residual_count_diff=sel_diff, ok=bool(worst_t <= TOL_T and worst_r <= TOL_R),
                map_points_mean=float(np.mean(maps)) if maps else 0.0, queries_mean=float(np.mean(qs)) if qs else 0.0,
                oracle="oracle.cpp scan2map (uniform-grid 5-NN backend) on the device's own downsampled maps and prior",
                poses=np.array(poses, np.float32))


def algorithmic_bytes(cls, n_raw, n_ds, q, iters=ITERS):
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
        "lm_iter": 96 * q * iters,                  # SURVEY.md §8d B_iter (query + five neighbours, 16 B each) x the
                                                    # iterations one persistent launch runs
    }[cls]


def load_traffic(top):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture of THIS command (profiles/r02_ncu_scan_kernel_full.json; ncu replays every kernel with cold caches, which is
    also the state of the L2-flushed timed steps). Returns (bytes or None, label)."""
    name = "r02_ncu_scan_kernel_full.json" if top == "lm_iter" else "r02_ncu_%s_full.json" % top
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", name)))
        return int(cap["traffic_bytes_per_launch"]), "profiles/%s: %s" % (name, cap.get("traffic_note", "ncu --set full, cold caches per replay pass"))
    except Exception:
        return None, "no committed ncu capture for this kernel"


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
    traffic, traffic_src = load_traffic(top)
    roof = {"kernel": "lm_persistent_kernel (the whole scan: transform, VoxelGrid, search grids, %d LM iterations)" % ITERS if top == "lm_iter" else top,
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 5), "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)",
            "algorithmic_bytes_per_launch": int(round(achieved * 1e9 * table[top]["us_per_launch"] * 1e-6)),
            "us_per_launch": table[top]["us_per_launch"], "share_of_step": table[top]["share"],
            "note": "one scan is latency-bound (10 dependent iterations, ~10 k queries): see DESIGN.md §5 and profiles/"}
    return roof, table, dict(n_raw=round(n_raw), n_ds=round(n_ds), queries=round(q))


def _pose_diff(a, b):
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    d[:3] = np.minimum(d[:3], np.abs(d[:3] - 2 * np.pi))
    return float(d[3:].max()), float(d[:3].max())


def run_single_config(kind, device, peaks, with_cpu):
    """BASELINE configs[2] / configs[3] as real workloads: one HDL-64E scan against its local map, per scan =
    ll_build_local_map (the reference rebuilds map and trees every frame) + ll_scan2map with HOST query clouds.
    Reports e2e ms per scan (wall, p50 of 5), the per-kernel-class device times with achieved algorithmic GB/s, pose /
    map parity against the oracle, and the CPU path (oracle + the reference's nanoflann) timed once beside it."""
    from lego_loam_sr_b200 import abi
    d = make_hdl64_problem(kind)
    ds = 0 if kind == "c3" else 1
    cq, sq = d["raw_query"] if kind == "c3" else (d["query"]["corner_scan_ds"], d["query"]["surf_total_ds"])
    kf = d["keyframes"]
    ids = list(range(len(kf)))
    peak = peaks.get("hbm_gbs", 6650.0)
    with abi.Handle(headline_params(abi, map_downsample=ds), device=device) as h:
        for i, f in enumerate(kf):
            h.add_keyframe(i, f["corner_scan"], f["surf_last_ds"], f["outlier_last_ds"], d["key_poses"][i])
        wall = []
        for r in range(7):
            if r == 2:
                h.profile_enable(True)
                h.profile_read()
            t0 = time.perf_counter()
            h.build_local_map(ids)
            pose, res = h.scan2map(cq, sq, d["prior"])
            if r >= 2:
                wall.append((time.perf_counter() - t0) * 1e3)
        prof = h.profile_read()
        gc, gs = h.get_map_ds(0), h.get_map_ds(1)
    n_raw = sum(len(f["corner_scan"]) + len(f["surf_last_ds"]) + len(f["outlier_last_ds"]) for f in kf)
    n_ds, q = len(gc) + len(gs), len(cq) + len(sq)
    table = {}
    for cls, (ms, cnt) in prof.items():
        if not cnt:
            continue
        us = ms * 1e3 / cnt
        b = algorithmic_bytes(cls, n_raw, n_ds, q, iters=res.iter_num)
        table[cls] = {"us_per_launch": round(us, 1), "launches_per_scan": round(cnt / len(wall), 2), "gbs": round(b / (us * 1e-6) / 1e9, 1),
                      "frac_of_hbm_peak": round(b / (us * 1e-6) / 1e9 / peak, 4)}
    out = {"workload": ("configs[2]: HDL-64E 64x1800 scan vs local map of %d keyframes" if kind == "c2" else
                        "configs[3]: raw HDL-64E scan (no query downsampling) vs %d keyframes' raw map, map_downsample = 0") % len(kf),
           "n_raw": n_raw, "n_ds": int(n_ds), "queries": q, "iterations": int(res.iter_num), "residuals": int(res.sel_num),
           "e2e_ms_per_scan_p50": round(statistics.median(wall), 3), "scans_per_s": round(1e3 / statistics.median(wall), 2),
           "h2d_bytes_per_scan": 16 * q, "kernels": table}
    if with_cpu:
        import orc
        orc.build()
        oprm = headline_params(orc, map_downsample=ds)
        use_ref = orc.have_ref()
        orc.set_knn_backend(orc.KNN_NANOFLANN if use_ref else orc.KNN_GRID)
        t0 = time.perf_counter()
        cds, sds, _, _ = orc.build_local_map(d["key_poses"], [f["corner_scan"] for f in kf], [f["surf_last_ds"] for f in kf],
                                             [f["outlier_last_ds"] for f in kf], oprm)
        ores, opose, _ = orc.scan2map(cds, sds, cq, sq, d["prior"], oprm)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        orc.set_knn_backend(orc.KNN_GRID)
        dt, dr = _pose_diff(pose, opose)
        out["cpu_ms_per_scan"] = round(cpu_ms, 1)
        out["cpu_knn"] = "nanoflann (reference, oracle/_ref)" if use_ref else "oracle uniform grid"
        out["parity"] = {"maps_bit_exact": bool(np.array_equal(gc.view(np.uint32), cds.view(np.uint32)) and
                                                np.array_equal(gs.view(np.uint32), sds.view(np.uint32))),
                         "dt_m": dt, "dr_rad": dr, "ok": bool(dt <= TOL_T and dr <= TOL_R),
                         "residuals_gpu_vs_cpu": [int(res.sel_num), int(ores.sel_num)]}
    return out


def run_as_shipped(data, device, with_cpu):
    """Fidelity record: the fork AS SHIPPED — pose update commented out
    (mapOptmization.cpp:1539-1545), row-0 / row-mask eigen port bugs, iterCountThres = 200. The pose never moves, so the
    reference repeats the same iteration 200 times; the device computes it once and reports iter_num = 200."""
    from lego_loam_sr_b200 import abi
    kw = dict(apply_update=0, eigen_mode=0, iter_count_thres=200, stop_thres=0.05)
    n_t = 10
    r = run_gpu(data, device, WINDOW, 3, n_t, mode="host", flush=True, param_override=abi.default_params(**kw))
    out = {"params": "apply_update=0, eigen_mode=0 (as shipped), iter_count_thres=200, stop_thres=0.05",
           "gpu_e2e_ms_per_scan": round(statistics.mean(r["wall_ms"]), 4), "gpu_scans_per_s": round(1e3 / statistics.mean(r["wall_ms"]), 1),
           "iterations_reported": sorted(set(x[1] for x in r["results"])), "timed_scans": n_t}
    if with_cpu:
        import orc
        orc.build()
        use_ref = orc.have_ref()
        orc.set_knn_backend(orc.KNN_NANOFLANN if use_ref else orc.KNN_GRID)
        # untimed window fill with ONE iteration per scan: as shipped the pose never moves, so the state after the fill is
        # the same as with 200; the timed scans then run the shipped 200 iterations
        rp = orc.Replay(orc.default_params(**dict(kw, iter_count_thres=1)), window=WINDOW)
        frames, odom = data["frames"], data["odom"]
        times, oposes = [], []
        for k in range(WINDOW + 3 + 2):
            f = frames[k]
            if k == WINDOW + 3:
                rp.set_params(orc.default_params(**kw))
            t0 = time.perf_counter()
            ores, _, _ = rp.step(odom[k], *[f[c] for c in CLOUD_KEYS])
            if k >= WINDOW + 3:
                times.append(time.perf_counter() - t0)
                oposes.append(np.array(ores.pose[:], np.float32))
        orc.set_knn_backend(orc.KNN_GRID)
        rp.close()
        out["cpu_ms_per_scan"] = round(1e3 * statistics.mean(times), 1)
        out["cpu_timed_scans"] = len(times)
        out["cpu_iterations"] = int(ores.iter_num)
        out["max_abs_pose_diff_vs_cpu"] = float(np.abs(np.array(oposes, np.float64) - r["poses"][:len(oposes)].astype(np.float64)).max())
    return out


def run_raw_frame(seed, device, with_cpu):
    """The whole frame body of MapOptimization::run with RAW feature clouds: downsampleCurrentScan (six VoxelGrids,
    mapOptmization.cpp:1234-1267) runs on the device too (ll_downsample_scan), then a1-a11 and the keyframe store.
    Timed inside the C++ mirror with steady_clock around the frame body, the way the reference fills mapt.txt."""
    n_t = 30
    data = make_data(seed, WINDOW + 5 + n_t, keep_raw=True)
    r = run_gpu(data, device, WINDOW, 5, n_t, mode="raw", flush=False)
    out = {"frame_ms_p50": round(statistics.median(r["frame_ms"]), 4), "frame_ms_mean": round(statistics.mean(r["frame_ms"]), 4),
           "frames_per_s": round(1e3 / statistics.mean(r["frame_ms"]), 1), "timed_frames": n_t,
           "includes": "ll_downsample_scan + extractSurroundingKeyFrames + scan2MapOptimization + keyframe store, host clouds in, host results out",
           "raw_points_per_frame": int(np.mean([len(f["corner_scan"]) * 2 + len(f["surf_last"]) * 2 + len(f["outlier_last"]) for f in data["frames"][WINDOW:]])),
           "all_scans_optimised": all(x[0] == 0 for x in r["results"])}
    if with_cpu:
        import orc
        orc.build()
        use_ref = orc.have_ref()
        orc.set_knn_backend(orc.KNN_NANOFLANN if use_ref else orc.KNN_GRID)
        prm = headline_params(orc)
        rp = orc.Replay(prm, window=WINDOW)
        times = []
        for k in range(WINDOW + 5 + 5):
            f = data["frames"][k]
            t0 = time.perf_counter()
            c_ds = orc.voxel_downsample(f["corner_scan"], prm.corner_leaf)[0]     # CornerLastDS and CornerScanDS (:1236-1247)
            orc.voxel_downsample(f["corner_scan"], prm.corner_leaf)
            s_ds = orc.voxel_downsample(f["surf_last"], prm.surf_leaf)[0]         # SurfLastDS and SurfScanDS
            orc.voxel_downsample(f["surf_last"], prm.surf_leaf)
            o_ds = orc.voxel_downsample(f["outlier_last"], prm.surf_leaf)[0]
            st_ds = orc.voxel_downsample(np.concatenate([s_ds, o_ds]), prm.surf_leaf)[0]
            rp.step(data["odom"][k], f["corner_scan"], c_ds, s_ds, o_ds, st_ds)
            if k >= WINDOW + 5:
                times.append(time.perf_counter() - t0)
        orc.set_knn_backend(orc.KNN_GRID)
        rp.close()
        out["cpu_frame_ms_mean"] = round(1e3 * statistics.mean(times), 1)
    return out


# ----------------------------------------------------------------------------------------------- batched replay
_RANK_CPUS = None


def rank_cpus(local_rank, world):
    """Disjoint slice of the host cores for this rank (its threads stay next to each other and off the other ranks').
    Computed ONCE per process, from the affinity the process started with: main() then restricts the process to its
    slice, and a second slicing of that slice (the batched replay asks again) left 2 of 8 cores at N = 4."""
    global _RANK_CPUS
    if _RANK_CPUS is not None:
        return _RANK_CPUS
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        return []
    if world > 1 and len(cores) >= 2 * world:
        per = len(cores) // world
        cores = cores[local_rank * per:(local_rank + 1) * per]
    _RANK_CPUS = cores
    return cores


def batch_schedule(want_in_flight, n_sequences, n_cpus, forced_group=0):
    """How a rank drives its sequences: -> (host threads, sequences each thread keeps in flight, sequences in flight).
    One host thread per core at most (each spins on its scan's completion). Short of cores, a thread keeps several
    sequences in flight (host/batch_replay.cpp: begin the frame of each, then end them), aiming at 1.5 x the wanted number:
    measured on one GPU restricted to 4 cores (taskset), 8 x 1: 7 669 scans/s; 4 x 1: 6 258; 4 x 2: 6 995; 4 x 3: 7 309."""
    want = max(1, min(want_in_flight, n_sequences))
    threads = max(1, min(want, n_cpus if n_cpus > 0 else want))
    group = 1 if threads >= want else (3 * want + 2 * threads - 1) // (2 * threads)
    if forced_group > 0:
        group = forced_group
    return threads, group, max(1, min(n_sequences, threads * group))


def run_batch(args, rank, world, device, dist):
    """BASELINE configs[4]: `--sequences` independent VLP-16 sequences, seq -> rank = seq mod world, one handle + stream per
    sequence, `--threads` C++ host threads per GPU (host/batch_replay.cpp). Every sequence first registers WINDOW scans
    UNTIMED (so that the timed scans run against full 50-keyframe windows, like the headline), then `--scans-per-seq`
    timed scans. Host clouds (e2e path). NCCL only gathers the result records."""
    import torch
    from lego_loam_sr_b200 import abi, batch, host, synth
    ids = batch.shard_sequences(args.sequences, world, rank)
    n_frames = WINDOW + args.scans_per_seq
    n_base = 4
    bases = {}
    seqs = []
    for sid in ids:
        b = sid % n_base
        if b not in bases:  # a few distinct trajectories; every sequence is one of them moved to its own start pose
            d = make_data(100 + b, n_frames)
            bases[b] = (d, [[to_pcl_layout(f[k]) for k in CLOUD_KEYS] for f in d["frames"][:n_frames]])
        d, clouds = bases[b]
        G = synth.pose_to_mat([0.0, 0.1 * (sid // n_base), 0.0, 3.0 * (sid // n_base), 0.0, -2.0 * (sid // n_base)])
        odom = np.array([synth.mat_to_pose(G @ synth.pose_to_mat(p)) for p in d["odom"][:n_frames]], np.float32)
        seqs.append(dict(odom=odom, clouds=clouds))
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    cpus = rank_cpus(int(os.environ.get("LOCAL_RANK", "0")), world)
    # one host thread per sequence in flight, each spinning on its own core while its scan runs: never more threads than
    # this rank has cores (a 2-GPU box with 12 cores gave each rank 6: eight threads on them ran the batch at 65 % per GPU)
    threads, group, in_flight = batch_schedule(args.threads, len(ids), len(cpus), args.batch_group)
    nb = args.batch_blocks if args.batch_blocks >= 0 else max(1, sms // in_flight)
    kw = dict(persistent_lm=0) if nb == 0 else dict(persistent_lm=1, lm_max_blocks=nb)
    nodes = [host.MapOptimization(headline_params(abi, **kw), device=device, loop_closure_enabled=True, search_num=WINDOW,
                                  reserve=(1 << 18, 1 << 20, 1 << 15)) for _ in ids]
    host.batch_replay(nodes, seqs, 0, WINDOW, threads, cpus, group)   # untimed: fill the windows
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    rec, secs = host.batch_replay(nodes, seqs, WINDOW, args.scans_per_seq, threads, cpus, group)
    torch.cuda.synchronize(device)
    for n in nodes:
        n.close()
    dt = torch.tensor([secs], dtype=torch.float64, device="cuda:%d" % device)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    local = {sid: rec[i].astype(np.float64) for i, sid in enumerate(ids)}
    # gather_records expects its own 13-column layout; carry the 16-float records as they are
    allrec = np.full((args.sequences, args.scans_per_seq, 16), np.nan, np.float64)
    if world > 1:
        max_owned = (args.sequences + world - 1) // world
        block = np.full((max_owned, args.scans_per_seq, 16), np.nan, np.float64)
        for i, sid in enumerate(ids):
            block[i] = local[sid]
        t = torch.from_numpy(block).to("cuda:%d" % device)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        for r in range(world):
            arr = parts[r].cpu().numpy()
            for i, sid in enumerate(batch.shard_sequences(args.sequences, world, r)):
                allrec[sid] = arr[i]
    else:
        for sid, v in local.items():
            allrec[sid] = v
    total = args.sequences * args.scans_per_seq
    secs_max = float(dt[0])
    record = {"workload": "configs[4]: %d independent VLP-16 sequences, seq -> rank = seq mod world, windows of %d keyframes pre-filled "
                          "untimed, then %d timed scans per sequence; %d sequences in flight per GPU (%d C++ host threads x %d), %s"
                          % (args.sequences, WINDOW, args.scans_per_seq, in_flight, threads, group,
                             "persistent scan kernels on <= %d SMs each" % nb if nb else "one launch per LM iteration"),
              "value": round(total / secs_max, 2), "unit": UNIT, "per_gpu": round(total / secs_max / world, 2), "n_gpus": world,
              "seconds": round(secs_max, 4), "timing": "steady_clock in the C++ driver from the threads' common start to the last result on the host, max over ranks",
              "all_scans_optimised": bool((allrec[:, :, 6] == 0).all()), "iterations": sorted(set(int(v) for v in np.unique(allrec[:, :, 7]))),
              "mean_frame_ms": round(float(np.nanmean(allrec[:, :, 14])), 4), "cpus_per_rank": len(cpus)}
    return record, allrec


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
        if hasattr(h, "fa_update_transformation"):
            out["f1_f3_fa_update_transformation"] = run_fa_loop_row(h, fa, np.zeros(6, np.float32), p50, with_cpu)
            # the same call from a poor initial guess: both loops run tens of iterations, the surf loop's search iterations
            # (1.4 k feature points against 5.5 k) included
            out["f1_f3_fa_update_transformation_far"] = run_fa_loop_row(h, fa, np.array([0.01, -0.012, 0.008, 0.15, -0.05, 0.5], np.float32), p50, with_cpu)
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


def run_fa_loop_row(h, fa, t0, p50, with_cpu):
    """f-1 + f-3 as ONE device-resident call: FeatureAssociation::updateTransformation (featureAssociation.cpp:2505-2535)."""
    cs, cl = fa["corner"]["_inputs"]
    ss, sl = fa["surf"]["_inputs"]
    out = {"gpu": p50(lambda: h.fa_update_transformation(cs, ss, t0), n=30, warm=4)}
    h.fa_reset_state()
    res = h.fa_update_transformation(cs, ss, t0)
    out["iterations"] = [int(res["iters_surf"]), int(res["iters_corner"])]
    if with_cpu:
        import orc
        nf = orc.have_ref()
        out["cpu"] = p50(lambda: orc.fa_update_transformation(cs, ss, cl, sl, t0, use_nanoflann=nf), n=5, warm=1)
        o = orc.fa_update_transformation(cs, ss, cl, sl, t0, use_nanoflann=False)
        out["max_abs_diff_transform_cur"] = float(np.abs(np.asarray(res["transform_cur"]) - np.asarray(o["transform_cur"])).max())
    return out


def run_cpu(data, prefill, warmup, steps, budget_s):
    """The reference's CPU path (oracle restatement; k-NN = the reference's vendored nanoflann when oracle/_ref was
    built), single thread like the reference's _run_thread (mapOptmization.cpp:175). Times `steps` scans after
    prefill + warmup untimed ones, or as many as fit in budget_s. Returns (times, k-NN label, poses of the timed scans)."""
    import orc
    orc.build()
    knn = "nanoflann (reference, oracle/_ref)" if orc.have_ref() else "oracle uniform grid"
    orc.set_knn_backend(orc.KNN_NANOFLANN if orc.have_ref() else orc.KNN_GRID)
    rp = orc.Replay(headline_params(orc), window=WINDOW)
    frames, odom = data["frames"], data["odom"]
    times, poses = [], []
    t_start = time.perf_counter()
    for k in range(min(prefill + warmup + steps, len(frames))):
        f = frames[k]
        t0 = time.perf_counter()
        res, _, _ = rp.step(odom[k], f["corner_scan"], f["corner_scan_ds"], f["surf_last_ds"], f["outlier_last_ds"], f["surf_total_ds"])
        dt = time.perf_counter() - t0
        if k >= prefill + warmup:
            times.append(dt)
            poses.append(np.array(res.pose[:], np.float32))
            if time.perf_counter() - t_start > budget_s and len(times) >= 5:
                break
    orc.set_knn_backend(orc.KNN_GRID)
    rp.close()
    return times, knn, np.array(poses, np.float32)


def _cpu_worker(seed, n_scans, prefill, warmup, steps, budget_s, core, q):
    try:
        if core is not None:
            os.sched_setaffinity(0, {core})
    except Exception:
        pass
    data = make_data(seed, n_scans)
    times, knn, _ = run_cpu(data, prefill, warmup, steps, budget_s)
    q.put((seed, times, knn))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args, config, prefill, n_scans):
    """--impl reference: N independent oracle replays (one per requested GPU of the other arm: seeds seed .. seed + N - 1,
    the sequences that arm's ranks replay), one process each on its own core; value = all scans / the slowest process."""
    import multiprocessing as mp
    n = max(1, args.gpus)
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        cores = []
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    budget = 150.0
    procs = [ctx.Process(target=_cpu_worker, args=(args.seed + r, n_scans, prefill, args.warmup, args.steps, budget,
                                                   cores[r % len(cores)] if cores else None, q)) for r in range(n)]
    for p in procs:
        p.start()
    got = [q.get() for _ in procs]
    for p in procs:
        p.join()
    all_times = [t for _, t, _ in got]
    knn = got[0][2]
    v = sum(len(t) for t in all_times) / max(sum(t) for t in all_times)
    flat = [x for t in all_times for x in t]
    sample = ("%d process(es) x %d consecutive scans of the same replay(s) after %d untimed scans filled the window; k-NN = %s"
              % (n, min(len(t) for t in all_times), prefill + args.warmup, knn))
    return {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / v, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "p50_ms_per_scan": round(statistics.median(flat) * 1e3, 3),
            "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": n, "kind": "port", "sample": sample,
                             "host_cores_available": host_cores(),
                             "why_one_core_per_stream": "one registration stream is sequential; the reference runs it on one thread (mapOptmization.cpp:175); N streams = N processes"},
            "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    global _WORLD, _STEP
    args = parse_args()
    if args.legacy_world:
        _WORLD, _STEP = R01_WORLD, R01_STEP_M
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    prefill = WINDOW
    n_scans = prefill + args.warmup + args.steps
    config = {"workload": "configs[1]: synthetic VLP-16 trajectory replay, sliding local map of the last %d keyframes, "
                          "%d LM iterations per scan; sized to the metric's statement (VLP-16 scan of ~2 k corner + ~8 k surf "
                          "queries vs a ~100 k-point local map)%s" % (WINDOW, ITERS, " [LEGACY round-1 world]" if args.legacy_world else ""),
              "sensor": "VLP-16 16x1800", "metres_per_scan": _STEP, "window_keyframes": WINDOW,
              "lm_iterations": ITERS, "apply_update": 1, "eigen_mode": "upstream", "lm_solver": "ldlt-f64", "persistent_lm": 1, "prefill_scans": prefill,
              "sequences_per_gpu": 1, "l2": "flushed before every timed step (256 MiB memset)" if not args.no_l2_flush else "not flushed"}

    if args.impl == "reference":
        if rank != 0:
            return
        print(json.dumps(reference_arm(args, config, prefill, n_scans)))
        return

    # Keep NCCL's own output (the "NCCL version ..." banner is printed at every debug level from VERSION up, to stdout)
    # off stdout: the driver reads ONE JSON line there.
    # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION is raised to WARN.)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:  # this rank's threads on their own slice of the host cores
        cpus = rank_cpus(local_rank, world)
        try:
            if cpus:
                os.sched_setaffinity(0, set(cpus))
        except Exception:
            pass
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    import __graft_entry__ as ge
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ge.build_product()
    if world > 1:
        dist.barrier()
    device = local_rank
    if args.workload == "batch":
        record, allrec = run_batch(args, rank, world, device, dist)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            if args.batch_dump:
                np.save(args.batch_dump, allrec)
            total = args.sequences * args.scans_per_seq
            print(json.dumps({"metric": METRIC, "value": record["value"], "unit": UNIT, "n_gpus": world,
                              "steps": args.scans_per_seq, "warmup": 0, "ms_per_step": round(1e3 * record["seconds"] / total, 4),
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": record["workload"], "timing": record["timing"]},
                              "batch": record}))
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
    e2e = run_gpu(data, device, prefill, args.warmup, args.steps, mode="host", flush=flush)
    # ---- value leg: inputs resident in HBM
    sync_all()
    res = run_gpu(data, device, prefill, args.warmup, args.steps, mode="resident", flush=flush)
    sync_all()
    clocks = sampler.stop()
    # per-rank times (device time of the resident leg, WALL time of the e2e leg), max over ranks; NCCL carries results only
    t_res = torch.tensor([sum(res["step_ms"]), sum(e2e["wall_ms"]), sum(e2e["step_ms"])], dtype=torch.float64, device="cuda:%d" % device)
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        poses = torch.from_numpy(res["poses"]).to("cuda:%d" % device)
        gathered = [torch.empty_like(poses) for _ in range(world)]
        dist.all_gather(gathered, poses)
    total_ms, total_e2e_wall_ms, total_e2e_dev_ms = float(t_res[0]), float(t_res[1]), float(t_res[2])
    batch_record = None
    if not args.no_batch and not args.legacy_world:
        try:
            batch_record, _ = run_batch(args, rank, world, device, dist)
        except Exception as e:
            batch_record = {"error": "%s: %s" % (type(e).__name__, e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- per-kernel timing pass (rank 0): same steps with CUDA events around every launch
    prof_run = run_gpu(data, device, prefill, args.warmup, args.steps, mode="resident", flush=flush, profile=True)
    roof, table, sizes = roofline_from_profile(prof_run["profile"], prof_run, peaks)
    value = world * args.steps / (total_ms * 1e-3)
    e2e_value = world * args.steps / (total_e2e_wall_ms * 1e-3)
    checks = {"all_scans_optimised": all(r[0] == 0 for r in res["results"]),
              "iterations_per_scan": sorted(set(r[1] for r in res["results"])),
              "resident_equals_e2e_pose": bool(np.array_equal(res["poses"], e2e["poses"]))}
    if not args.no_parity:
        par = replay_parity(data, device, prefill + args.warmup, args.steps, every=max(1, args.steps // 20))
        chk_poses = par.pop("poses")
        par["checked_replay_equals_timed_poses_bitwise"] = bool(np.array_equal(chk_poses.view(np.uint32), e2e["poses"].view(np.uint32)))
        par["ok"] = bool(par["ok"] and par["checked_replay_equals_timed_poses_bitwise"] and checks["resident_equals_e2e_pose"])
        checks["parity"] = par
    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "p50_ms_per_scan": round(statistics.median(res["step_ms"]), 4),
            "p50_ms_per_scan_e2e": round(statistics.median(e2e["wall_ms"]), 4),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(np.mean(e2e["h2d_bytes"])),
                    "d2h_bytes_per_step": 484, "timing": "wall clock (perf_counter) around each call of the C++ mirror, host clouds (pcl::PointXYZI records in page-locked memory) in, host pose out",
                    "wall_ms_per_step": round(total_e2e_wall_ms / args.steps, 4), "device_ms_per_step": round(total_e2e_dev_ms / args.steps, 4),
                    "note": "query clouds are pre-downsampled on the host; the frame with downsampleCurrentScan on the device is `raw_frame`"},
            "gpu_launches": int(res["launches"]), "gpu_launches_per_step": round(res["launches"] / args.steps, 2),
            "clocks": clocks, "roofline": roof, "kernels": table, "sizes": sizes, "checks": checks}
    if "parity" in checks and not checks["parity"]["ok"]:
        line["failed"] = "parity: a timed scan's pose differs from the oracle's beyond tolerance (see checks.parity)"
    if batch_record is not None:
        line["batch"] = batch_record
    with_cpu = not args.no_cpu_baseline
    if world == 1 and not args.no_at_scale:
        for name, fn in (("raw_frame", lambda: run_raw_frame(args.seed, device, with_cpu)),
                         ("as_shipped", lambda: run_as_shipped(data, device, with_cpu)),
                         ("configs2", lambda: run_single_config("c2", device, peaks, with_cpu)),
                         ("configs3", lambda: run_single_config("c3", device, peaks, with_cpu)),
                         ("next_rows", lambda: run_next_rows(with_cpu=with_cpu))):
            try:
                line[name] = fn()
            except Exception as e:  # (never lose the headline line to a side measurement)
                line[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    if world == 1 and with_cpu:
        times, knn, cpu_poses = run_cpu(data, prefill, args.warmup, args.steps, budget_s=args.cpu_seconds)
        v = len(times) / sum(times)
        dt, dr = _pose_diff(e2e["poses"][len(times) - 1], cpu_poses[-1]) if len(times) else (None, None)
        line["cpu_baseline"] = {"value": round(v, 3), "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "%d consecutive scans of the same replay after %d untimed scans filled the window; "
                                          "restated PCL/Eigen pieces, k-NN = %s" % (len(times), prefill + args.warmup, knn),
                                "p50_ms_per_scan": round(statistics.median(times) * 1e3, 3), "host_cores_available": host_cores(),
                                "independent_replay_drift_vs_gpu": {"after_scans": prefill + args.warmup + len(times), "dt_m": dt, "dr_rad": dr,
                                                                    "note": "two independent replays (maps built from each side's own poses); the strict per-frame check is checks.parity"}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
