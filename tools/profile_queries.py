"""driver for ncu captures of the query kernels: builds a 20-frame map, then runs 4M getOdd/getOccupancy + 2M getOddGrad"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from mlmapping_b200 import MLMap, config_cfg_a, scenes
cfg = config_cfg_a()
m = MLMap(cfg)
for k in range(20):
    pose = scenes.corridor_trajectory_pose(k)
    m.integrate_depth(scenes.corridor_depth_frame(cfg, pose, frame_idx=k), pose)
g = m.export_map()["glb"]
pos = scenes.query_positions(10_000_000, g.min(0) * 1.0, (g.max(0) + 1) * 1.0, seed=5)
d_pos = m.to_device(pos)
o1, o2, o3 = m.device_alloc(16_000_000), m.device_alloc(16_000_000), m.device_alloc(48_000_000)
for _ in range(2):
    m.getOdd_device(d_pos, 4_000_000, o1)
    m.getOccupancy_device(d_pos + 96_000_000, 4_000_000, o2)
    m.getOddGrad_device(d_pos + 192_000_000, 2_000_000, o3, 5)
    m.sync()
print("ok")
