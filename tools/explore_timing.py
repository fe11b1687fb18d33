"""exploration-mode frame time (CFG-A, device-resident frames, L2 flushed) next to the normal frame"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mlmapping_b200 import MLMap, config_cfg_a, scenes
ROWS, COLS = 480, 640
for explore in (0, 1):
    cfg = config_cfg_a()
    cfg.use_exploration_frontiers = explore
    m = MLMap(cfg)
    fr = []
    for k in range(30):
        pose = scenes.corridor_trajectory_pose(k)
        fr.append((m.to_device(scenes.corridor_depth_frame(cfg, pose, frame_idx=k)), pose))
    ms, slow = 0.0, 0
    for k, (d, pose) in enumerate(fr):
        m.flush_l2()
        m.timer_start()
        st = m.integrate_depth_device(d, ROWS, COLS, pose)
        t = m.timer_stop_ms()
        if k >= 10:
            ms += t
            slow += st.ordering_slow_path
    print("exploration" if explore else "normal", f"{1e3 * ms / 20:.1f} us/frame", "rehash frames", slow)
    m.close()
