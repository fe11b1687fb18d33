"""per-column phase breakdown of k_column (needs a -DMLM_PHASE_TIMING build: sh csrc/build.sh -DMLM_PHASE_TIMING)"""
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
    m.integrate_depth(img, pose)
c = m.debug_phase_cycles()
act = c[:, 10] > 0
names = ["gather", "bound", "contrib", "radix", "fold||walks", "miss-stage"]
cc = np.concatenate([c[:, 15:16], c[:, :4], c[:, 6:8]], axis=1)
d = np.diff(cc, axis=1)[act]
tot = d.sum(1)
order = np.argsort(-tot)
print("columns", act.sum(), "cycles: total max", tot.max(), "mean", tot.mean())
print("phase", names)
print("mean  ", d.mean(0).astype(int))
print("max   ", d.max(0).astype(int))
for i in order[:6]:
    print("col", np.nonzero(act)[0][i], "n_c", c[act][i, 10], "n_k", c[act][i, 11], "tot", tot[i], d[i])
