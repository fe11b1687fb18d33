"""CFG-C scan timing on one GPU (device-resident points, warm L2), for A/B runs"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mlmapping_b200 import MLMap, config_cfg_c, scenes
cfg = config_cfg_c()
m = MLMap(cfg)
data = []
for k in range(8):
    pose = scenes.lidar_loop_pose(k)
    p = scenes.lidar_scan(pose, frame_idx=k)
    data.append((m.to_device(p), p.shape[0], pose))
ms = 0.0
for k, (dp, n, pose) in enumerate(data):
    m.timer_start()
    m.integrate_points_device(dp, n, pose)
    t = m.timer_stop_ms()
    if k >= 3:
        ms += t
print(f"lidar scan {1e3 * ms / 5:.1f} us")
