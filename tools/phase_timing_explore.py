"""k_frame_explore timeline for CFG-A exploration frames (needs a -DMLM_PHASE_TIMING build, MLM_LIB_PATH)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from mlmapping_b200 import MLMap, config_cfg_a, scenes
cfg = config_cfg_a()
cfg.use_exploration_frontiers = 1
m = MLMap(cfg)
acc = []
for k in range(24):
    pose = scenes.corridor_trajectory_pose(k)
    d = m.to_device(scenes.corridor_depth_frame(cfg, pose, frame_idx=k))
    m.flush_l2()
    st = m.integrate_depth_device(d, 480, 640, pose)
    if k < 12 or st.ordering_slow_path:
        continue
    c = m.debug_phase_cycles()
    fa = c[(c[:, 0] > 10 ** 17) & (c[:, 12] > 10 ** 17)]   # the per-CTA rows hold globaltimer stamps (ns since the epoch)
    t0 = fa[:, 0].min()
    acc.append(((fa[:, [0, 1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 8]] - t0) / 1e3).max(0))
print(st.as_dict())
rel = np.array(acc).mean(0)
names = ["start", "proj done", "bar1 out", "cols done", "bar2 out", "fuse hits done", "tkey done", "explore_a done", "explore_b done",
         "fuse misses done", "release done", "finish ticket"]
prev = 0.0
for n, v in zip(names, rel):
    print(f"{n:18s} last CTA at {v:7.1f} us  (+{v - prev:5.1f})")
    prev = v
