"""planner-query timing on the CFG-A map of the bench (200 corridor frames): 10 M queries 4:4:2, L2 flushed, CUDA events"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mlmapping_b200 import MLMap, config_cfg_a, scenes
cfg = config_cfg_a()
m = MLMap(cfg)
for k in range(200):
    pose = scenes.corridor_trajectory_pose(k)
    m.integrate_depth(scenes.corridor_depth_frame(cfg, pose, frame_idx=k), pose)
ex = m.export_map()
d = cfg.subbox_d_xyz * cfg.subbox_n
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pos = scenes.query_positions(nq, ex["glb"].min(0) * d, (ex["glb"].max(0) + 1) * d, seed=5)
n_odd = n_occ = int(0.4 * nq)
n_grad = nq - n_odd - n_occ
d_pos = m.to_device(pos)
o1, o2, o3 = m.device_alloc(4 * n_odd), m.device_alloc(4 * n_occ), m.device_alloc(24 * n_grad)
acc = {}
for rep in range(13):
    t = {}
    m.flush_l2(); m.timer_start(); m.getOdd_device(d_pos, n_odd, o1); t["getOdd"] = m.timer_stop_ms()
    m.flush_l2(); m.timer_start(); m.getOccupancy_device(d_pos + 24 * n_odd, n_occ, o2); t["getOccupancy"] = m.timer_stop_ms()
    m.flush_l2(); m.timer_start(); m.getOddGrad_device(d_pos + 24 * (n_odd + n_occ), n_grad, o3, 5); t["getOddGrad"] = m.timer_stop_ms()
    if rep >= 3:
        for k_, v_ in t.items():
            acc[k_] = acc.get(k_, 0.0) + v_ / 10
print({k_: round(v_, 4) for k_, v_ in acc.items()}, "total ms", round(sum(acc.values()), 4), "subboxes", ex["glb"].shape[0])
