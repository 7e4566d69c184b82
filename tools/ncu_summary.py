#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into the per-kernel lines DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("gpc__cycles_elapsed.avg.per_second", "gpc clock"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long scoreboard %"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stall short scoreboard %"),
    ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "stall mio throttle %"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall barrier %"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall math pipe throttle %"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "stall wait %"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "stall lg throttle %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}  (ncu --set full --clock-control none; per launch, cold cache, serialised)")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        print(f"\n== {r[col['Kernel Name']][:110]}")
        seen = set()
        for key, label in KEYS:
            if key in col and label not in seen and r[col[key]] != "":
                seen.add(label)
                print(f"   {label:38s} {r[col[key]]:>16s} {units[col[key]]}")


if __name__ == "__main__":
    main()
