"""host-buffer (pinned) end-to-end latency of mlm_integrate_depth_u16, with and without L2 flush"""
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
    f = scenes.corridor_depth_frame(cfg, pose, frame_idx=k)
    pf = m.pinned_array(f.shape, np.uint16)
    pf[...] = f
    frames.append(pf)
    poses.append(pose)
for k in range(10):
    m.integrate_depth(frames[k], poses[k])
for flush in (True, False):
    wall = 0.0
    for k in range(10, 60):
        if flush:
            m.flush_l2()
            m.sync()
        t0 = time.perf_counter()
        m.integrate_depth(frames[k], poses[k])
        wall += time.perf_counter() - t0
    print(f"flush={flush}: e2e host wall {1e6*wall/50:.1f} us/frame")
