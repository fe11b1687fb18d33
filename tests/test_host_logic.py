"""CPU: harness-side logic — synthetic scene determinism, config presets, golden fixtures of the oracle."""
import hashlib
import json
from pathlib import Path

import numpy as np

from mlmapping_b200 import config_cfg_a, config_cfg_b, config_cfg_c, scenes
from oracle_binding import Oracle

GOLDEN = Path(__file__).resolve().parent / "golden"


def _digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_scene_generators_are_deterministic():
    cfg = config_cfg_a()
    pose = scenes.corridor_trajectory_pose(17)
    a = scenes.corridor_depth_frame(cfg, pose, frame_idx=17)
    b = scenes.corridor_depth_frame(cfg, pose, frame_idx=17)
    assert a.dtype == np.uint16 and a.shape == (480, 640) and np.array_equal(a, b)
    assert (a == 0).mean() < 0.05 and a.max() > 20000
    pts = scenes.lidar_scan(scenes.lidar_loop_pose(3), frame_idx=3, beams=16, azimuths=256)
    assert pts.shape[1] == 3 and 1000 < pts.shape[0] <= 16 * 256
    assert np.linalg.norm(pts, axis=1).max() <= 50.02
    q = scenes.query_positions(1000, [0, -2, 0], [10, 2, 3])
    assert q.shape == (1000, 3)


def test_config_presets_match_survey():
    a, b, c = config_cfg_a(), config_cfg_b(), config_cfg_c()
    assert (a.am_n_rho, a.am_n_z_below + a.am_n_z_over + 1, int(360 / a.am_d_phi_deg)) == (65, 41, 360)
    assert 65 * 41 * 360 == 959400
    assert (b.am_n_rho, b.am_n_z_below + b.am_n_z_over + 1) == (180, 81)
    assert (c.am_n_rho, c.am_n_z_below + c.am_n_z_over + 1) == (250, 201)
    for cfg in (a, b, c):
        assert cfg.subbox_n == 10 and cfg.use_exploration_frontiers == 0


def test_oracle_matches_golden_fixtures():
    """oracle dumps on the SURVEY §8d seeds, committed under tests/golden/ by make_golden.py"""
    meta = json.loads((GOLDEN / "golden.json").read_text())
    from golden.make_golden import run_case
    for name, want in meta["cases"].items():
        got = run_case(name)
        assert got == want, name


def test_pose_forwarding_matches_the_callback_and_the_oracle():
    """mlm_compensate_pose = depth_odom_input_callback's linear pose forwarding (src/mlmap.cpp:470-498) with Sophus'
    SO3::log / SO3::exp (so3.cpp:127-199): known answers, then the product (host code of the C ABI) against the oracle"""
    import ctypes as C
    from mlmapping_b200 import compensate_pose
    from oracle_binding import load_oracle
    lib = load_oracle()

    def oracle(pos, q, v, w, go, gi, lat):
        arr = [(C.c_double * len(a))(*a) for a in (pos, q, v, w)]
        out = (C.c_double * 7)()
        lib.orc_compensate_pose(arr[0], arr[1], arr[2], arr[3], go, gi, lat, out)
        return np.array(out[:])

    # zero gaps: the odometry pose itself (log followed by exp returns the unit quaternion up to rounding)
    q0 = np.array([np.cos(0.3), 0.0, 0.0, np.sin(0.3)])
    T = compensate_pose([1, 2, 3], q0, [0.5, 0, 0], [0, 0, 1.0], 0.0, 0.0, 0.0)
    assert np.allclose(T[:3], [1, 2, 3]) and np.allclose(T[3:], q0, atol=1e-15)
    # pure yaw rate 1 rad/s for 0.1 s after a 20 ms latency: yaw grows by 0.08 rad; position moves 0.03 * v
    T = compensate_pose([1, 2, 3], q0, [0.5, -1.0, 0.25], [0, 0, 1.0], 0.05, 0.1, 0.02)
    assert np.allclose(T[:3], [1 + 0.03 * 0.5, 2 - 0.03, 3 + 0.03 * 0.25], atol=1e-15)
    yaw = 2 * np.arctan2(T[6], T[3])
    assert abs(yaw - (0.6 + 0.08)) < 1e-12 and abs(T[4]) < 1e-15 and abs(T[5]) < 1e-15
    # body-frame rate: rot_dot = R * omega (a roll rate seen from a yawed body turns about the rotated axis)
    T = compensate_pose([0, 0, 0], [np.cos(np.pi / 4), 0, 0, np.sin(np.pi / 4)], [0, 0, 0], [1.0, 0, 0], 0.0, 1e-3, 0.0)
    lg = 2 * np.arccos(T[3]) * T[4:7] / np.linalg.norm(T[4:7])
    assert np.allclose(lg, [0.0, 1e-3, np.pi / 2], atol=1e-12)
    # random cases incl. the small-angle branches: bit-identical to the oracle restatement
    rs = np.random.RandomState(11)
    for i in range(200):
        q = rs.normal(size=4)
        if i % 10 == 0:
            q = np.array([1.0, 0, 0, 0]) + rs.normal(size=4) * 1e-12  # n < SMALL_EPS branch of log
        pos, v, w = rs.normal(size=3) * 5, rs.normal(size=3), rs.normal(size=3) * (0.0 if i % 25 == 0 else 1.0)
        go, gi, lat = rs.uniform(-0.05, 0.05), rs.uniform(-0.05, 0.05), rs.uniform(0, 0.03)
        a, b = compensate_pose(pos, q, v, w, go, gi, lat), oracle(list(pos), list(q), list(v), list(w), go, gi, lat)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (i, a, b)
        assert abs(np.linalg.norm(a[3:]) - 1) < 1e-15


def test_multiply_high_division_and_bucket_shortcut():
    """host-built constants of the device's division-free paths (mlmap_capi.cu make_div_magic / pow64_mod):
    x / d == umulhi(x, mul) >> shift for every x < 2^31, and libstdc++'s bucket of a negative int hash
    ((size_t)(int64)h % B) from 32-bit arithmetic with c64 = 2^64 mod B"""
    rs = np.random.RandomState(7)

    def magic(d):
        sh = 0
        while (2 << sh) <= d - 1:
            sh += 1
        return ((1 << (32 + sh)) // d) + 1, sh

    for d in [2, 3, 7, 10, 16, 64, 154, 155, 524, 154 * 154, 524 * 524, 1000003, 2 ** 20 + 1, 2 ** 30 - 1]:
        mul, sh = magic(d)
        assert mul < 2 ** 32
        xs = np.concatenate([rs.randint(0, 2 ** 31, 200000, dtype=np.int64), np.arange(0, 4096, dtype=np.int64),
                             (np.arange(1, 3000, dtype=np.int64) * d).clip(0, 2 ** 31 - 1),
                             (np.arange(1, 3000, dtype=np.int64) * d - 1).clip(0, 2 ** 31 - 1),
                             2 ** 31 - 1 - np.arange(0, 4096, dtype=np.int64)])
        q = ((xs.astype(object) * mul) >> 32) >> sh
        assert np.array_equal(np.array(q, dtype=np.int64), xs // d), d
    for B in [13, 29, 59, 172933, 351061, 712697, 2938679]:
        c64 = (1 << 64) % B
        hs = np.concatenate([rs.randint(-2 ** 31, 2 ** 31, 20000, dtype=np.int64), [-2 ** 31, -1, 0, 1, 2 ** 31 - 1, -B, -B - 1, -B + 1]])
        for h in hs.tolist():
            want = (h % (1 << 64)) % B
            if h >= 0:
                got = h % B
            else:
                r = (-h) % B
                got = c64 - r if c64 >= r else c64 + (B - r)
            assert got == want, (h, B)
