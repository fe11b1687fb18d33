import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from mlmapping_b200 import MLMap, config_cfg_a, scenes
cfg = config_cfg_a()
m = MLMap(cfg)
for k in range(6):
    pose = scenes.corridor_trajectory_pose(100 + k)
    img = scenes.corridor_depth_frame(cfg, pose, frame_idx=100 + k)
    st = m.integrate_depth(img, pose)
c = m.debug_phase_cycles().reshape(-1)[:3 * 592].reshape(592, 3)
t0 = c[:, 0].min()
start, mid, end = c[:, 0] - t0, c[:, 1] - t0, c[:, 2] - t0
print("touched voxels", st.n_touched_voxels)
print("CTA start ns: min %d max %d" % (start.min(), start.max()))
print("thread0 done ns: median %d max %d" % (np.median(mid), mid.max()))
print("CTA end ns: median %d p90 %d max %d" % (np.median(end), np.percentile(end, 90), end.max()))
dur = end - start
order = np.argsort(-dur)[:8]
print("longest CTAs:", [(int(i), int(dur[i])) for i in order])
print("dur of CTAs 0..9:", dur[:10].tolist(), " CTAs 200..205:", dur[200:206].tolist())
