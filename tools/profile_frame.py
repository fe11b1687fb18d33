"""small driver for ncu captures: integrates a few frames of the CFG-A trajectory (device-resident inputs)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mlmapping_b200 import MLMap, config_cfg_a, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = config_cfg_a()
m = MLMap(cfg)
for k in range(n):
    pose = scenes.corridor_trajectory_pose(100 + k)
    img = scenes.corridor_depth_frame(cfg, pose, frame_idx=100 + k)
    st = m.integrate_depth(img, pose)
print(st.as_dict())
