"""k_frame timeline for a CFG-C LiDAR scan (needs a -DMLM_PHASE_TIMING build)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from mlmapping_b200 import MLMap, config_cfg_c, scenes
cfg = config_cfg_c()
m = MLMap(cfg)
for k in range(5):
    pose = scenes.lidar_loop_pose(k)
    pts = scenes.lidar_scan(pose, frame_idx=k)
    st = m.integrate_points(pts, pose)
print(st.as_dict())
c = m.debug_phase_cycles()
nc = c.shape[0] - 256
fr, c = c[nc:], c[:nc]
act = c[:, 10] > 0
ca = c[act]
names = ["gather", "bound", "contrib", "radix", "fold||walks", "miss-stage"]
cc = np.concatenate([ca[:, 15:16], ca[:, :4], ca[:, 6:8]], axis=1)
d = np.diff(cc, axis=1)
print("items", act.sum(), "cycles per item: mean", d.sum(1).mean().astype(int), "max", d.sum(1).max(), "records mean", ca[:, 10].mean(), "contribs mean", ca[:, 11].mean())
print("phase", names)
print("mean  ", d.mean(0).astype(int))
print("max   ", d.max(0).astype(int))
fa = fr[fr[:, 0] > 0]
t0 = fa[:, 0].min()
rel = (fa[:, :9] - t0) / 1e3
for i, n in enumerate(["start", "proj done", "bar1 out", "cols done", "bar2 out", "bar3 out", "end", "col prologue", "fuse ticket"]):
    print(f"k_frame {n:12s}: min {rel[:, i].min():6.1f}  mean {rel[:, i].mean():6.1f}  max {rel[:, i].max():6.1f} us")
