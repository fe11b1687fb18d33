"""frame timing of the other BASELINE configs (CFG-B L515-like, CFG-C LiDAR) with per-kernel breakdown"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from mlmapping_b200 import MLMap, config_cfg_b, config_cfg_c, scenes

def run(name, cfg, gen, n=12):
    m = MLMap(cfg)
    m.set_profiling(True)
    acc, ev, rays = {}, 0.0, 0
    for k in range(n):
        data, pose = gen(k)
        m.flush_l2()
        m.timer_start()
        st = data(m, pose)
        t = m.timer_stop_ms()
        if k >= 3:
            ev += t; rays += st.n_points
            for kk, v in m.last_frame_kernel_ms().items(): acc[kk] = acc.get(kk, 0) + v
    nn = n - 3
    print(name, f"{1e3*ev/nn:.1f} us/frame (host input, events)", f"{rays/(ev*1e-3)/1e9:.2f} Grays/s", {k: round(1e3*v/nn, 1) for k, v in acc.items()}, st.as_dict())

cfgb = config_cfg_b(); cfgb.pool_submaps = 65536
def gen_b(k):
    pose = scenes.corridor_trajectory_pose(k * 2, step=0.1)
    img = scenes.corridor_depth_frame(cfgb, pose, rows=768, cols=1024, frame_idx=k, length=200.0)
    return (lambda m, p: m.integrate_depth(img, p)), pose
run("CFG-B 1024x768 @0.05m", cfgb, gen_b)
cfgc = config_cfg_c()
def gen_c(k):
    pose = scenes.lidar_loop_pose(k)
    pts = scenes.lidar_scan(pose, frame_idx=k)
    return (lambda m, p: m.integrate_points(pts, p)), pose
run("CFG-C 128x2048 LiDAR @0.2m", cfgc, gen_c, n=8)
