"""debug: world=1 sharded flow vs oracle, compared after every scan"""
import os, sys
from pathlib import Path
import numpy as np
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from mlmapping_b200 import config_cfg_c, scenes
from mlmapping_b200.sharded import ShardedMLMap
from oracle_binding import Oracle
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
cfg = config_cfg_c()
sh = ShardedMLMap(cfg, rank=0, world=1, device=0)
orc = Oracle(cfg)
for k in range(3):
    pose = scenes.lidar_loop_pose(k * 3)
    pts = scenes.lidar_scan(pose, frame_idx=k, beams=128, azimuths=2048)
    st = sh.integrate_points(pts, pose)
    so = orc.integrate_points(pts, pose)
    g, o = sh.export_map(), orc.export_map()
    order = np.lexsort((g["glb"][:, 2], g["glb"][:, 1], g["glb"][:, 0]))
    same_glb = g["glb"].shape == o["glb"].shape and np.array_equal(g["glb"][order], o["glb"])
    print("scan", k, "fast" , sh.last["fast_order"], "hits", sh.last["n_hit_total"], so.n_hit_cells, "miss", sh.last["n_miss_local"], so.n_miss_cells,
          "subboxes", g["glb"].shape[0], o["glb"].shape[0], "same set", same_glb, "touched", st.n_touched_voxels, so.n_touched_voxels)
    if same_glb:
        for name in ("occupancy", "log_odds"):
            u, v = g[name][order], o[name]
            if name == "log_odds":
                d = u.view(np.uint32) != v.view(np.uint32)
            else:
                d = u.view(np.uint8) != v.view(np.uint8)
            print("  ", name, "mismatching cells", int(d.sum()), "in subboxes", int(d.reshape(d.shape[0], -1).any(1).sum()))
            if d.any() and name == "log_odds":
                idx = np.argwhere(d.reshape(d.shape[0], -1))[:5]
                for s_, c_ in idx:
                    print("     subbox", o["glb"][s_], "cell", c_, "gpu", u.reshape(u.shape[0], -1)[s_, c_], "oracle", v.reshape(v.shape[0], -1)[s_, c_])
dist.destroy_process_group()
