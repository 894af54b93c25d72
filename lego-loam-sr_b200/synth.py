"""Seeded synthetic VLP-16 / HDL-64E data for the scan-to-map path.

The reference ships no input scans (its rosbag is empty), so tests and bench.py feed
both the CUDA path and the CPU oracle from this generator: a ground plane, axis-aligned wall
segments and vertical poles are ray-cast from a moving sensor, and the hits are split into the three
clouds MapOptimization receives from FeatureAssociation (utility.h:75-83): corner (poles and wall
ends), surf (ground and walls) and outlier (sparse noisy rest). Frames: the LOAM camera frame
(x left, y up, z forward); a pose is transformTobeMapped's layout [rx, ry, rz, tx, ty, tz] with
R = R_y(ry) R_x(rx) R_z(rz) (mapOptmization.cpp:606-620). Intensity = ring + column / 10000 as in
imageProjection.cpp:342.

This module is data generation only: no registration maths lives here.
"""
import os

import numpy as np

SENSORS = {
    # rings, lowest elevation (deg), highest elevation (deg), azimuth steps (loam_config.yaml:5-8, :141-144)
    "vlp16": (16, -15.0, 15.0, 1800),
    "hdl64": (64, -24.8, 2.0, 1800),
}


# ----------------------------------------------------------------------------- pose utilities
def pose_to_mat(p):
    rx, ry, rz, tx, ty, tz = [float(v) for v in p]
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    R = np.array([
        [cy * cz + sy * sx * sz, -cy * sz + sy * sx * cz, sy * cx],
        [cx * sz, cx * cz, -sx],
        [-sy * cz + cy * sx * sz, sy * sz + cy * sx * cz, cy * cx],
    ])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = [tx, ty, tz]
    return T


def mat_to_pose(T):
    R = T[:3, :3]
    rx = np.arcsin(np.clip(-R[1, 2], -1.0, 1.0))
    ry = np.arctan2(R[0, 2], R[2, 2])
    rz = np.arctan2(R[1, 0], R[1, 1])
    return np.array([rx, ry, rz, T[0, 3], T[1, 3], T[2, 3]], np.float64)


def perturb(pose, d_rot_deg, d_trans, rng=None):
    """Add (d_rot_deg, d_trans) to every rotation / translation component (signs random if rng)."""
    p = np.array(pose, np.float64)
    sg = np.ones(6) if rng is None else rng.choice([-1.0, 1.0], 6)
    p[:3] += sg[:3] * np.deg2rad(d_rot_deg)
    p[3:] += sg[3:] * d_trans
    return p


# ----------------------------------------------------------------------------- world
class World:
    """Ground plane y = -sensor_height, wall segments (vertical rectangles) and poles (cylinders)."""

    def __init__(self, seed=0, half_x=60.0, half_z=90.0, n_blocks=14, n_poles=60, sensor_height=1.8,
                 wall_height=7.0):
        rng = np.random.default_rng(seed)
        self.ground_y = -sensor_height
        self.wall_height = wall_height
        segs = []
        # outer boundary
        hx, hz = half_x, half_z
        segs += [(-hx, -hz, hx, -hz), (hx, -hz, hx, hz), (hx, hz, -hx, hz), (-hx, hz, -hx, -hz)]
        # rectangular "building" blocks: four wall segments each, kept clear of the driving corridor
        for _ in range(n_blocks):
            for _try in range(50):
                w, d = rng.uniform(6, 18), rng.uniform(6, 18)
                cx, cz = rng.uniform(-hx + 10, hx - 10), rng.uniform(-hz + 10, hz - 10)
                if self._clear_of_path(cx, cz, max(w, d) * 0.75 + 4.0, hx, hz):
                    segs += [(cx - w / 2, cz - d / 2, cx + w / 2, cz - d / 2),
                             (cx + w / 2, cz - d / 2, cx + w / 2, cz + d / 2),
                             (cx + w / 2, cz + d / 2, cx - w / 2, cz + d / 2),
                             (cx - w / 2, cz + d / 2, cx - w / 2, cz - d / 2)]
                    break
        self.segs = np.array(segs, np.float64)
        poles = []
        for _ in range(n_poles):
            for _try in range(50):
                px, pz = rng.uniform(-hx + 3, hx - 3), rng.uniform(-hz + 3, hz - 3)
                if self._clear_of_path(px, pz, 2.5, hx, hz):
                    poles.append((px, pz, rng.uniform(0.08, 0.2), rng.uniform(3.0, 6.0)))
                    break
        self.poles = np.array(poles, np.float64)
        self.half_x, self.half_z = hx, hz

    @staticmethod
    def path_xy(s, hx, hz):
        """Figure-eight driving line, parameter s in [0, 1)."""
        a = 2 * np.pi * s
        return 0.55 * hx * np.sin(2 * a), 0.7 * hz * np.sin(a)

    def _clear_of_path(self, x, z, margin, hx, hz):
        s = np.linspace(0, 1, 400, endpoint=False)
        px, pz = self.path_xy(s, hx, hz)
        return np.min(np.hypot(px - x, pz - z)) > margin


def trajectory(world, n, step=0.5, start=0.0, wobble=0.01, seed=0):
    """n ground-truth poses along the figure-eight at `step` metres per scan; heading follows the path."""
    rng = np.random.default_rng(seed + 7919)
    s_dense = np.linspace(0, 1, 20000, endpoint=False)
    x, z = World.path_xy(s_dense, world.half_x, world.half_z)
    seg = np.hypot(np.diff(x, append=x[0]), np.diff(z, append=z[0]))
    arc = np.concatenate([[0.0], np.cumsum(seg)[:-1]])
    total = arc[-1] + seg[-1]
    poses = []
    for k in range(n):
        a = (start * total + k * step) % total
        i = int(np.searchsorted(arc, a, side="right") - 1)
        j = (i + 1) % len(x)
        f = (a - arc[i]) / max(seg[i], 1e-9)
        px, pz = x[i] + f * (x[j] - x[i]), z[i] + f * (z[j] - z[i])
        ry = np.arctan2(x[j] - x[i], z[j] - z[i])  # z forward, x left
        rx = wobble * np.sin(0.13 * k) + rng.normal(0, wobble * 0.2)
        rz = wobble * np.cos(0.17 * k) + rng.normal(0, wobble * 0.2)
        poses.append([rx, ry, rz, px, 0.02 * np.sin(0.05 * k), pz])
    return np.array(poses, np.float64)


# ----------------------------------------------------------------------------- ray casting
def scan(world, pose, sensor="vlp16", seed=0, range_sigma=0.02, max_range=100.0, outlier_frac=0.06):
    """Ray-cast one revolution. Returns raw clouds in the SENSOR frame as float32 [n, 4]:
    corner_scan, surf_last, outlier_last."""
    rings, lo, hi, cols = SENSORS[sensor]
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(lo, hi, rings))
    azim = np.linspace(-np.pi, np.pi, cols, endpoint=False)
    E, A = np.meshgrid(elev, azim, indexing="ij")
    ring_id = np.repeat(np.arange(rings), cols).astype(np.float32)
    col_id = np.tile(np.arange(cols), rings).astype(np.float32)
    d_s = np.stack([np.cos(E) * np.sin(A), np.sin(E), np.cos(E) * np.cos(A)], -1).reshape(-1, 3)
    T = pose_to_mat(pose)
    d_w = d_s @ T[:3, :3].T
    o = T[:3, 3]
    n = len(d_w)
    best = np.full(n, np.inf)
    kind = np.zeros(n, np.int8)  # 0 none, 1 ground, 2 wall, 3 wall end (edge), 4 pole

    # ground
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (world.ground_y - o[1]) / d_w[:, 1]
    ok = (d_w[:, 1] < -1e-9) & (t > 0.5)
    upd = ok & (t < best)
    best[upd] = t[upd]
    kind[upd] = 1

    # Rays sorted by world azimuth so that each wall / pole only touches the rays that can see it.
    dx, dz = d_w[:, 0], d_w[:, 2]
    az = np.arctan2(dx, dz)
    order = np.argsort(az, kind="stable")
    az_sorted = az[order]

    def rays_between(a_lo, a_hi):
        """Indices of rays with world azimuth in [a_lo, a_hi] (angles may wrap)."""
        a_lo = (a_lo + np.pi) % (2 * np.pi) - np.pi
        a_hi = (a_hi + np.pi) % (2 * np.pi) - np.pi
        i0, i1 = np.searchsorted(az_sorted, a_lo, "left"), np.searchsorted(az_sorted, a_hi, "right")
        if a_lo <= a_hi:
            return order[i0:i1]
        return np.concatenate([order[i0:], order[:i1]])

    margin = 0.01
    # walls: ray (o + t d) in the xz plane against segment a + u (b - a)
    for (ax, az0, bx, bz) in world.segs:
        da, db = np.hypot(ax - o[0], az0 - o[2]), np.hypot(bx - o[0], bz - o[2])
        if min(da, db) > max_range + np.hypot(bx - ax, bz - az0):
            continue
        a0, a1 = np.arctan2(ax - o[0], az0 - o[2]), np.arctan2(bx - o[0], bz - o[2])
        dlt = (a1 - a0 + np.pi) % (2 * np.pi) - np.pi
        lo_a, hi_a = (a0, a0 + dlt) if dlt >= 0 else (a0 + dlt, a0)
        idx = rays_between(lo_a - margin, hi_a + margin)
        if len(idx) == 0:
            continue
        ex, ez = bx - ax, bz - az0
        rdx, rdz, rdy = dx[idx], dz[idx], d_w[idx, 1]
        den = rdx * ez - rdz * ex
        with np.errstate(divide="ignore", invalid="ignore"):
            tt = ((ax - o[0]) * ez - (az0 - o[2]) * ex) / den
            u = ((ax - o[0]) * rdz - (az0 - o[2]) * rdx) / den
        y = o[1] + tt * rdy
        ok = (np.abs(den) > 1e-12) & (tt > 0.5) & (u >= 0) & (u <= 1) & (y >= world.ground_y) & \
             (y <= world.ground_y + world.wall_height) & (tt < best[idx])
        sel = idx[ok]
        best[sel] = tt[ok]
        seg_len = np.hypot(ex, ez)
        edge = (u[ok] * seg_len < 0.12) | ((1 - u[ok]) * seg_len < 0.12)
        kind[sel] = np.where(edge, 3, 2)

    # poles: 2-D ray / circle
    for (px, pz, pr, ph) in world.poles:
        dist = np.hypot(px - o[0], pz - o[2])
        if dist > max_range or dist <= pr + 0.5:
            continue
        a0 = np.arctan2(px - o[0], pz - o[2])
        half = np.arcsin(min(1.0, pr / dist)) + margin
        idx = rays_between(a0 - half, a0 + half)
        if len(idx) == 0:
            continue
        rdx, rdz, rdy = dx[idx], dz[idx], d_w[idx, 1]
        dn2 = rdx * rdx + rdz * rdz
        fx, fz = o[0] - px, o[2] - pz
        b = fx * rdx + fz * rdz
        c = fx * fx + fz * fz - pr * pr
        disc = b * b - dn2 * c
        with np.errstate(divide="ignore", invalid="ignore"):
            tt = (-b - np.sqrt(np.maximum(disc, 0))) / dn2
        y = o[1] + tt * rdy
        ok = (disc > 0) & (tt > 0.5) & (y >= world.ground_y) & (y <= world.ground_y + ph) & (tt < best[idx])
        sel = idx[ok]
        best[sel] = tt[ok]
        kind[sel] = 4

    hit = np.isfinite(best) & (best < max_range)
    r = np.where(hit, best, 0.0) + rng.normal(0, range_sigma, n)
    pts = d_s * r[:, None]  # sensor frame
    inten = ring_id + col_id / 10000.0
    cloud = np.concatenate([pts, inten[:, None]], 1).astype(np.float32)

    is_out = hit & (kind != 1) & (kind != 4) & (rng.random(n) < outlier_frac)
    corner = hit & ((kind == 4) | (kind == 3)) & ~is_out
    surf = hit & ((kind == 1) | (kind == 2)) & ~is_out
    out = cloud[is_out].copy()
    out[:, :3] += rng.normal(0, 0.05, (len(out), 3)).astype(np.float32)
    return dict(corner_scan=cloud[corner], surf_last=cloud[surf], outlier_last=out)


# ----------------------------------------------------------------------------- input preparation
def voxel_np(pts, leaf):
    """Plain numpy voxel-grid mean filter used only to PREPARE inputs (the query clouds the reference
    gets from downsampleCurrentScan, mapOptmization.cpp:1234-1267). Parity of the VoxelGrid stage
    itself is tested CUDA-vs-oracle elsewhere; this is not the checker."""
    pts = np.ascontiguousarray(pts, np.float32)
    if len(pts) == 0:
        return pts.reshape(0, 4)
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(pts[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(0)
    dims = ijk.max(0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.flatnonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))
    sums = np.add.reduceat(pts[order].astype(np.float64), starts, axis=0)
    cnt = np.diff(np.concatenate([starts, [len(ks)]]))[:, None]
    return (sums / cnt).astype(np.float32)


def frame_inputs(raw, corner_leaf=0.2, surf_leaf=0.4, keep_raw=False):
    """The five clouds one MapOptimization frame consumes after downsampleCurrentScan (keep_raw: plus the raw surf and
    outlier clouds, for the legs that run downsampleCurrentScan on the device)."""
    corner_scan_ds = voxel_np(raw["corner_scan"], corner_leaf)
    surf_last_ds = voxel_np(raw["surf_last"], surf_leaf)
    outlier_last_ds = voxel_np(raw["outlier_last"], surf_leaf)
    surf_total_ds = voxel_np(np.concatenate([surf_last_ds, outlier_last_ds]), surf_leaf)
    out = dict(corner_scan=np.ascontiguousarray(raw["corner_scan"], np.float32),
               corner_scan_ds=corner_scan_ds, surf_last_ds=surf_last_ds,
               outlier_last_ds=outlier_last_ds, surf_total_ds=surf_total_ds)
    if keep_raw:
        out["surf_last"] = np.ascontiguousarray(raw["surf_last"], np.float32)
        out["outlier_last"] = np.ascontiguousarray(raw["outlier_last"], np.float32)
    return out


def noisy_odometry(gt_poses, seed=0, rot_sigma_deg=0.05, trans_sigma=0.01):
    """Odometry chain (transformSum per scan): ground-truth increments with a small random
    perturbation each, composed — so it drifts like laser odometry does."""
    rng = np.random.default_rng(seed + 104729)
    odom = [np.array(gt_poses[0], np.float64)]
    T_prev_gt = pose_to_mat(gt_poses[0])
    T_odom = pose_to_mat(gt_poses[0])
    for k in range(1, len(gt_poses)):
        T_gt = pose_to_mat(gt_poses[k])
        inc = np.linalg.inv(T_prev_gt) @ T_gt
        noise = np.concatenate([np.deg2rad(rng.normal(0, rot_sigma_deg, 3)), rng.normal(0, trans_sigma, 3)])
        T_odom = T_odom @ inc @ pose_to_mat(noise)
        odom.append(mat_to_pose(T_odom))
        T_prev_gt = T_gt
    return np.array(odom, np.float32)


def make_sequence(seed=2, n_scans=100, step=0.5, sensor="vlp16", world_kw=None, cache_dir=None, keep_raw=False):
    """A trajectory replay problem (BASELINE config 2 shape): per scan the five input clouds, the
    odometry pose and the ground-truth pose. The first n scans are the same for every n_scans >= n."""
    key = "seq_s%d_n%d_st%.3f_%s_%s%s.npz" % (seed, n_scans, step, sensor,
                                              "-".join("%s%s" % kv for kv in sorted((world_kw or {}).items())),
                                              "_raw" if keep_raw else "")
    if cache_dir:
        path = os.path.join(cache_dir, key)
        if os.path.exists(path):
            z = np.load(path, allow_pickle=True)
            return dict(frames=list(z["frames"]), odom=z["odom"], gt=z["gt"])
    world = World(seed=seed, **(world_kw or {}))
    gt = trajectory(world, n_scans, step=step, seed=seed)
    frames = [frame_inputs(scan(world, gt[k], sensor=sensor, seed=seed * 100003 + k), keep_raw=keep_raw) for k in range(n_scans)]
    odom = noisy_odometry(gt, seed=seed)
    out = dict(frames=frames, odom=odom, gt=gt.astype(np.float32))
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        np.savez(path, frames=np.array(frames, dtype=object), odom=odom, gt=out["gt"])
    return out


def make_single(seed=1, n_map_scans=50, step=0.5, sensor="vlp16", rot_deg=0.5, trans=0.05, world_kw=None, cache_dir=None):
    """BASELINE config 1/3 shape: one query scan against a local map assembled from n_map_scans prior
    scans at their ground-truth poses; the prior is the ground truth perturbed by (rot_deg, trans)."""
    world = World(seed=seed, **(world_kw or {}))
    gt = trajectory(world, n_map_scans + 1, step=step, seed=seed)
    path = None
    if cache_dir:
        path = os.path.join(cache_dir, "single_s%d_n%d_st%.3f_%s_%s.npz" % (
            seed, n_map_scans, step, sensor, "-".join("%s%s" % kv for kv in sorted((world_kw or {}).items()))))
    if path and os.path.exists(path):
        frames = list(np.load(path, allow_pickle=True)["frames"])
    else:
        frames = [frame_inputs(scan(world, gt[k], sensor=sensor, seed=seed * 100003 + k)) for k in range(n_map_scans + 1)]
        if path:
            os.makedirs(cache_dir, exist_ok=True)
            np.savez(path, frames=np.array(frames, dtype=object))
    q = frames[-1]
    rng = np.random.default_rng(seed)
    prior = perturb(gt[-1], rot_deg, trans, rng).astype(np.float32)
    return dict(keyframes=frames[:-1], key_poses=gt[:-1].astype(np.float32), query=q,
                gt=gt[-1].astype(np.float32), prior=prior)
