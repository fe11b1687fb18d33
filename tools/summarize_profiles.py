"""turn the ncu captures under gpurun_out/ into the tracked summaries under profiles/ (round-tagged)"""
import csv, json, subprocess, sys
from collections import defaultdict
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

def raw(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    return [{k: (f"{v} {u}".strip() if u and k != "Kernel Name" else v) for k, v, u in zip(h, r, units)} for r in rows[2:]]

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        # atomic throughput (north star): L2 sectors touched by atomics / reductions, shared-memory atomics
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "smsp__inst_executed_op_shared_atom.sum",
        "smsp__inst_executed_op_global_atom.sum", "smsp__inst_executed_op_global_red.sum",
        # warp divergence (north star): active threads per executed warp instruction (32 = none)
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct"]
UNITS = {}
summary = {}
for name in ["prof_frame", "prof_queries", "prof_lidar", "prof_shard", "prof_explore"]:
    rep = ROOT / "gpurun_out" / f"{name}_{tag}.ncu-rep"
    if not rep.exists():
        continue
    for d in raw(rep):
        k = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
        summary.setdefault(k, []).append({m: d.get(m) for m in KEYS if d.get(m) not in (None, "")})
(out / f"ncu_full_summary_{tag}.json").write_text(json.dumps(summary, indent=1))

# launch list of `bench.py` under ncu: per-kernel durations and share of a frame
ll = ROOT / "gpurun_out" / f"launches_{tag}.csv"
if ll.exists():
    rows = list(csv.reader(open(ll)))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[i]
    acc = defaultdict(list)
    for r in rows[i + 1:]:
        if len(r) > 5:
            acc[r[h.index("Kernel Name")].split("(")[0].replace("void ", "")].append(float(r[h.index("Metric Value")].replace(",", "")) / 1000)
    (out / f"launches_{tag}.csv").write_text(open(ll).read())
    med = {k: sorted(v)[len(v) // 2] for k, v in acc.items()}
    phases = {k: v for k, v in med.items() if k.startswith(("k_project", "k_column", "k_fuse"))}
    tot = sum(phases.values())
    lines = [f"# ncu launch list of `python bench.py --steps 6 --warmup 3 --no-cpu --no-lidar --no-agents` ({tag}); cold-cache, serialised: compare SHARES",
             "# a timed step launches ONE map kernel, k_frame (share 1.0 of the step); k_project / k_column / k_fuse are the same three",
             "# phases run as stand-alone kernels by bench.py's profiling pass: their shares split the step",
             "kernel,launches,median_us,share_of_step"]
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        share = "1.000" if k.startswith("k_frame") else (f"{phases[k] / tot:.3f}" if k in phases else "")
        lines.append(f"{k},{len(v)},{med[k]:.2f},{share}")
    (out / f"launch_summary_{tag}.csv").write_text("\n".join(lines) + "\n")
    print("\n".join(lines))
print(json.dumps({k: v[0] for k, v in summary.items()}, indent=1)[:6000])
