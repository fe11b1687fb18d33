"""Synthetic, seeded inputs for the ROS-free harness: box-corridor depth frames
(D435i-/L515-like), 128-beam LiDAR scans of a hall with boxes, trajectories and query streams.
Pure numpy; RandomState (MT19937) streams are frozen across numpy versions, so the frames are
reproducible bit for bit.  Used identically for the CUDA path and the CPU oracle."""
from __future__ import annotations

import numpy as np


def quat_to_rot(q):
    w, x, y, z = [float(v) for v in q]
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def pose_from_xyz_yaw(x, y, z, yaw):
    """pose[7] = (tx,ty,tz,qw,qx,qy,qz) for a yaw-only body orientation"""
    return np.array([x, y, z, np.cos(yaw / 2), 0.0, 0.0, np.sin(yaw / 2)], dtype=np.float64)


def corridor_trajectory_pose(k: int, y_offset: float = 0.0, step: float = 0.05):
    """BASELINE config 2: x=5+0.05k, y=0.3 sin(0.02k), z=1.2, yaw=0.2 sin(0.015k)"""
    return pose_from_xyz_yaw(5.0 + step * k, y_offset + 0.3 * np.sin(0.02 * k), 1.2, 0.2 * np.sin(0.015 * k))


def _ray_box_exit(o, d, lo, hi):
    """distance parameter at which rays starting inside the box [lo,hi] leave it"""
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(d > 0, (hi - o) / d, np.where(d < 0, (lo - o) / d, np.inf))
    return t.min(axis=-1)


def corridor_depth_frame(cfg, T_wb, rows=480, cols=640, frame_idx=0, seed_drop=1, seed_noise=2,
                         length=60.0, half_width=1.2, height=2.4, y_offset=0.0):
    """Depth image (uint16 mm) of an axis-aligned box corridor seen by the configured camera.

    depth = exact ray/box intersection along the optical axis, x1000, rounded; pixels deeper than
    10 m are zeroed in 5 % of the rows (mt19937 stream seed_drop); additive integer noise
    U{-2..2} mm (mt19937 stream seed_noise) on the non-zero pixels."""
    T_wb = np.asarray(T_wb, dtype=np.float64)
    R_wb = quat_to_rot(T_wb[3:7])
    T_bs = np.array(list(cfg.T_bs), dtype=np.float64)
    R_bs = quat_to_rot(T_bs[3:7])
    o = T_wb[:3] + R_wb @ T_bs[:3]
    u = np.arange(cols, dtype=np.float64)
    v = np.arange(rows, dtype=np.float64)
    uu, vv = np.meshgrid(u, v)
    d_s = np.stack([(uu - cfg.cam_cx) / cfg.cam_fx, (vv - cfg.cam_cy) / cfg.cam_fy, np.ones_like(uu)], axis=-1)
    d_w = d_s @ (R_wb @ R_bs).T
    lo = np.array([0.0, y_offset - half_width, 0.0])
    hi = np.array([length, y_offset + half_width, height])
    t = _ray_box_exit(o, d_w, lo, hi)
    mm = np.rint(t * 1000.0)
    mm = np.where((mm > 65535) | ~np.isfinite(mm), 0, mm).astype(np.int64)
    rs = np.random.RandomState(seed_drop * 1000003 + frame_idx)
    drop_rows = rs.rand(rows) < 0.05
    mm[drop_rows[:, None] & (mm > 10000)] = 0
    rn = np.random.RandomState(seed_noise * 1000003 + frame_idx)
    noise = rn.randint(-2, 3, size=mm.shape)
    mm = np.where(mm > 0, np.clip(mm + noise, 1, 65535), 0)
    return mm.astype(np.uint16)


def hall_boxes(seed=4, n_boxes=64, hall=120.0, height=12.0):
    rs = np.random.RandomState(seed)
    c = np.stack([rs.uniform(-hall / 2 + 5, hall / 2 - 5, n_boxes), rs.uniform(-hall / 2 + 5, hall / 2 - 5, n_boxes)], 1)
    sz = rs.uniform(1.0, 6.0, (n_boxes, 2))
    hz = rs.uniform(1.0, height - 1.0, n_boxes)
    lo = np.concatenate([c - sz / 2, np.zeros((n_boxes, 1))], 1)
    hi = np.concatenate([c + sz / 2, hz[:, None]], 1)
    return lo, hi


def lidar_scan(T_wb, frame_idx=0, beams=128, azimuths=2048, fov_deg=22.5, max_range=50.0, seed_noise=3,
               seed_boxes=4, hall=120.0, height=12.0):
    """128-beam LiDAR scan (sensor frame == body frame) of a hall with random boxes.  Returns the
    returns within max_range as (n,3) float64 sensor-frame points, beam-major order."""
    T_wb = np.asarray(T_wb, dtype=np.float64)
    R = quat_to_rot(T_wb[3:7])
    o = T_wb[:3]
    el = np.deg2rad(np.linspace(-fov_deg, fov_deg, beams))
    az = np.linspace(0.0, 2 * np.pi, azimuths, endpoint=False)
    ee, aa = np.meshgrid(el, az, indexing="ij")
    d_s = np.stack([np.cos(ee) * np.cos(aa), np.cos(ee) * np.sin(aa), np.sin(ee)], -1).reshape(-1, 3)
    d_w = d_s @ R.T
    lo_h = np.array([-hall / 2, -hall / 2, 0.0])
    hi_h = np.array([hall / 2, hall / 2, height])
    t = _ray_box_exit(o, d_w, lo_h, hi_h)
    blo, bhi = hall_boxes(seed_boxes, hall=hall, height=height)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d_w
        for b in range(blo.shape[0]):
            t1 = (blo[b] - o) * inv
            t2 = (bhi[b] - o) * inv
            tn = np.nanmax(np.minimum(t1, t2), axis=1)
            tf = np.nanmin(np.maximum(t1, t2), axis=1)
            hit = (tn <= tf) & (tn > 0.05)
            t = np.where(hit & (tn < t), tn, t)
    rn = np.random.RandomState(seed_noise * 1000003 + frame_idx)
    t = t + rn.uniform(-0.01, 0.01, t.shape)
    keep = (t > 0.3) & (t <= max_range)
    return np.ascontiguousarray(d_s[keep] * t[keep, None])


def lidar_loop_pose(k: int, n_frames=100, radius=40.0 / (2 * np.pi)):
    """100 scans along a 40 m loop (circle) at 1.5 m height, heading tangent"""
    a = 2 * np.pi * k / n_frames
    return pose_from_xyz_yaw(radius * np.cos(a), radius * np.sin(a), 1.5, a + np.pi / 2)


def query_positions(n, aabb_min, aabb_max, seed=5, inflate=5.0, frac_inside=0.8):
    """CFG-D query stream: 80 % uniform in the AABB of allocated subboxes, 20 % in the AABB inflated by 5 m"""
    rs = np.random.RandomState(seed)
    aabb_min = np.asarray(aabb_min, dtype=np.float64)
    aabb_max = np.asarray(aabb_max, dtype=np.float64)
    n_in = int(n * frac_inside)
    a = rs.uniform(0, 1, (n_in, 3)) * (aabb_max - aabb_min) + aabb_min
    b = rs.uniform(0, 1, (n - n_in, 3)) * (aabb_max - aabb_min + 2 * inflate) + (aabb_min - inflate)
    out = np.concatenate([a, b], 0)
    rs.shuffle(out)
    return np.ascontiguousarray(out)
