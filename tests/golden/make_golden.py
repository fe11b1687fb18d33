"""Generates tests/golden/golden.json: digests + counters of the CPU oracle on the seeded inputs of
SURVEY §8d.  The reference itself cannot run here (no Eigen/PCL/ROS), so these fixtures pin the
ORACLE (regression guard for the restatement), not the reference: parity stays 'unpinned'.
Run:  python tests/golden/make_golden.py"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

from mlmapping_b200 import config_cfg_a, config_cfg_c, scenes  # noqa: E402
from oracle_binding import Oracle  # noqa: E402


def _d(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _summ(o, st, extra=None):
    keys, p = o.last_frame_hits()
    miss = o.last_frame_misses()
    m = o.export_map()
    out = {"n_points": st.n_points, "n_inside": st.n_inside, "n_cast": st.n_cast, "n_hit": st.n_hit_cells,
           "n_miss": st.n_miss_cells, "n_touched": st.n_touched_voxels, "buckets": st.hit_bucket_count,
           "ram_expand_cnt": st.ram_expand_cnt, "obs_cnt": st.obs_cnt,
           "hits_in_order": _d(keys), "hit_p": _d(p), "miss": _d(miss), "subboxes": _d(m["glb"]),
           "occupancy": _d(m["occupancy"]), "log_odds": _d(m["log_odds"]),
           "n_occupied": int((m["occupancy"] == b"o").sum()), "n_free": int((m["occupancy"] == b"f").sum())}
    if extra:
        out.update(extra)
    return out


def run_case(name):
    if name == "cfg_a_config1_single_frame":
        cfg = config_cfg_a()
        pose = scenes.pose_from_xyz_yaw(5.0, 0.0, 1.2, 0.0)
        img = scenes.corridor_depth_frame(cfg, pose)
        o = Oracle(cfg)
        st = o.integrate_depth(img, pose)
        return _summ(o, st, {"image": _d(img)})
    if name == "cfg_a_config2_first_8_frames_stride_25":
        cfg = config_cfg_a()
        o = Oracle(cfg)
        for k in range(8):
            pose = scenes.corridor_trajectory_pose(25 * k)
            st = o.integrate_depth(scenes.corridor_depth_frame(cfg, pose, frame_idx=25 * k), pose)
        m = o.export_map()
        lo = m["glb"].min(0) * 1.0
        hi = (m["glb"].max(0) + 1) * 1.0
        pos = scenes.query_positions(20000, lo, hi, seed=5, inflate=2.0)
        return _summ(o, st, {"occ_q": _d(o.getOccupancy(pos)), "odd_q": _d(o.getOdd(pos)),
                             "grad_q": _d(o.getOddGrad(pos[:5000]))})
    if name == "cfg_c_small_lidar_3_scans":
        cfg = config_cfg_c()
        cfg.am_n_rho, cfg.am_n_z_below, cfg.am_n_z_over = 120, 30, 30
        o = Oracle(cfg)
        for k in range(3):
            pose = scenes.lidar_loop_pose(3 * k)
            st = o.integrate_points(scenes.lidar_scan(pose, frame_idx=k, beams=32, azimuths=512), pose)
        return _summ(o, st)
    raise KeyError(name)


CASES = ["cfg_a_config1_single_frame", "cfg_a_config2_first_8_frames_stride_25", "cfg_c_small_lidar_3_scans"]

if __name__ == "__main__":
    out = {"generator": "tests/golden/make_golden.py", "note": "pins the oracle restatement, not the reference",
           "cases": {c: run_case(c) for c in CASES}}
    (HERE / "golden.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))
