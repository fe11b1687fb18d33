"""where does the frame time go besides the kernels?  device-resident frames, with/without L2 flush"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from mlmapping_b200 import MLMap, config_cfg_a, scenes
cfg = config_cfg_a()
m = MLMap(cfg)
frames, poses = [], []
for k in range(60):
    pose = scenes.corridor_trajectory_pose(k)
    frames.append(m.to_device(scenes.corridor_depth_frame(cfg, pose, frame_idx=k)))
    poses.append(pose)
for k in range(10):
    m.integrate_depth_device(frames[k], 480, 640, poses[k])
for flush in (True, False):
    ev, wall = 0.0, 0.0
    for k in range(10, 60):
        if flush:
            m.flush_l2()
        t0 = time.perf_counter()
        m.timer_start()
        m.integrate_depth_device(frames[k], 480, 640, poses[k])
        ev += m.timer_stop_ms()
        wall += time.perf_counter() - t0
    print(f"flush={flush}: events {1e3*ev/50:.1f} us/frame, host wall {1e6*wall/50:.1f} us/frame")
