"""Run under torchrun on N GPUs: the sharded LiDAR map (mlmapping_b200.sharded) over N ranks must equal the CPU
oracle's single map: every subbox is owned by exactly one rank and all owned subboxes match bit for bit."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from mlmapping_b200 import config_cfg_c, scenes  # noqa: E402
from mlmapping_b200.sharded import ShardedMLMap  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    full = "--full" in sys.argv
    cfg = config_cfg_c()
    beams, az, scans = (128, 2048, 3) if full else (32, 512, 4)
    if "--scans" in sys.argv:
        scans = int(sys.argv[sys.argv.index("--scans") + 1])
    if not full:
        cfg.am_n_rho, cfg.am_n_z_below, cfg.am_n_z_over = 120, 30, 30
        cfg.max_points = 32 * 512
        cfg.pool_submaps = 8192
    sh = ShardedMLMap(cfg, rank=rank, world=world, device=local)
    orc = None
    if rank == 0:
        from oracle_binding import Oracle
        orc = Oracle(cfg)
    times = []
    for k in range(scans):
        pose = scenes.lidar_loop_pose(k * 3)
        pts = scenes.lidar_scan(pose, frame_idx=k, beams=beams, azimuths=az)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if "--slices" in sys.argv:   # every rank copies only its part of the scan from the host
            lo, hi = (rank * pts.shape[0]) // world, ((rank + 1) * pts.shape[0]) // world
            sh.submit_slice(pts[lo:hi], lo, pts.shape[0], pose)
            st = sh.finish()
        else:
            st = sh.integrate_points(pts, pose)
        times.append(time.perf_counter() - t0)
        if orc is not None:
            st_o = orc.integrate_points(pts, pose)
            assert st.n_hit_cells == st_o.n_hit_cells, (sh.last, st_o.n_hit_cells)
    mine = sh.export_map()
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v for k, v in mine.items()})
    ok = True
    if rank == 0:
        o = orc.export_map()
        glb = np.concatenate([g["glb"] for g in gathered])
        order = np.lexsort((glb[:, 2], glb[:, 1], glb[:, 0]))
        glb = glb[order]
        assert np.array_equal(glb, o["glb"]), ("union of owned subboxes differs from the oracle", glb.shape, o["glb"].shape)
        for name in ("occupancy", "inflate", "log_odds", "collapsed"):
            u = np.concatenate([g[name] for g in gathered])[order]
            same = np.array_equal(u.view(np.uint8) if u.dtype.kind == "S" else u, o[name].view(np.uint8) if o[name].dtype.kind == "S" else o[name])
            if name == "log_odds":
                same = np.array_equal(u.view(np.uint32), o[name].view(np.uint32))
            assert same, name
        print(json.dumps({"sharded_check": "ok", "world": world, "subboxes": int(glb.shape[0]),
                          "owned_per_rank": [int(g["glb"].shape[0]) for g in gathered],
                          "ms_per_scan": [round(1e3 * t, 3) for t in times],
                          "median_ms_after_warmup": round(1e3 * float(np.median(times[2:])), 3) if len(times) > 3 else None,
                          "last": sh.last}))
    dist.barrier()
    sh.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
