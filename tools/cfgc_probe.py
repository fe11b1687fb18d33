import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from mlmapping_b200 import MLMap, config_cfg_c, scenes
cfg = config_cfg_c()
m = MLMap(cfg)
for k in range(8):
    pose = scenes.lidar_loop_pose(k)
    pts = scenes.lidar_scan(pose, frame_idx=k)
    d = m.to_device(pts)
    m.timer_start(); t0=time.perf_counter()
    st = m.integrate_points_device(d, pts.shape[0], pose)
    ms = m.timer_stop_ms(); w=time.perf_counter()-t0
    print(k, f"dev {ms*1e3:.0f} us wall {w*1e6:.0f} us", st.n_hit_cells, st.hit_bucket_count, st.ordering_slow_path, st.n_new_submaps)
