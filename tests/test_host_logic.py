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
