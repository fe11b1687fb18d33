"""CFG-C scan timing of the sharded path on ONE GPU (world = 1: same kernels, no peers) next to the fused single-GPU
frame, device-resident points, CUDA events on the library's stream; per-kernel times via MLM_SHARD_PROFILE"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mlmapping_b200 import MLMap, config_cfg_c, scenes
from mlmapping_b200.sharded import sharded_group_in_process
cfg = config_cfg_c()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
data = []
for k in range(10):
    pose = scenes.lidar_loop_pose(k)
    data.append((scenes.lidar_scan(pose, frame_idx=k), pose))
m = MLMap(cfg)
dev = [(m.to_device(p), p.shape[0], pose) for p, pose in data]
ms = 0.0
for k, (dp, n, pose) in enumerate(dev):
    m.flush_l2()
    m.timer_start()
    m.integrate_points_device(dp, n, pose)
    t = m.timer_stop_ms()
    if k >= 3:
        ms += t
print(f"fused single-GPU scan: {1e3 * ms / 7:.1f} us")
m.close()
ranks = sharded_group_in_process(cfg, world)
dev = [[(r.map.to_device(p), p.shape[0], pose) for p, pose in data] for r in ranks]
ms = 0.0
for k in range(10):
    for r in ranks:
        r.map.flush_l2()
    ranks[0].map.timer_start()
    for i, r in enumerate(ranks):
        dp, n, pose = dev[i][k]
        r.submit_device(dp, n, pose)
    for r in ranks:
        st = r.finish()
    t = ranks[0].map.timer_stop_ms()
    if k >= 3:
        ms += t
    print(k, f"{1e3 * t:.1f} us", ranks[0].last)
print(f"sharded path, world {world} on one GPU: {1e3 * ms / 7:.1f} us per scan")
