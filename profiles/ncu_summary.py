#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md cites.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rNN_kernel.txt]"""
import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200.build import kernel_source_sha  # noqa: E402  (stamps the capture with the kernel sources' hash)

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__inst_issued.avg.per_cycle_active",
    "smsp__average_warp_latency_issue_stalled_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


JSON_KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
             "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
             "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
             "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def to_json(rep, frames_per_launch, source):
    """--json: counters of the longest launch of each search kernel -> profiles/ncu_counters.json
    (bench.py's roofline.traffic / roofline.ncu)."""
    import json
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    best = {}
    for r in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
        stage = "inter" if "k_inter_search" in r[name_col] else ("intra" if ("k_intra_rows" in r[name_col] or "k_intra_wavefront" in r[name_col]) else None)
        if stage is None:
            continue
        t = float(d["gpu__time_duration.sum"][1].replace(",", ""))
        if stage in best and best[stage][0] >= t:
            continue
        o = {"kernel": r[name_col].split("(")[0], "frames_per_launch": frames_per_launch, "source": source,
             "source_sha16": kernel_source_sha(stage)}
        for k in JSON_KEYS:
            if k in d:
                u, v = d[k]
                x = float(v.replace(",", ""))
                if k.startswith("dram__bytes"):
                    x *= UNIT_SCALE.get(u, 1.0)
                    u = "byte"
                o[k] = x
                o[k + ".unit"] = u
        o["dram_bytes_per_frame"] = (o.get("dram__bytes_read.sum", 0.0) + o.get("dram__bytes_write.sum", 0.0)) / frames_per_launch
        best[stage] = (t, o)
    print(json.dumps({k: v[1] for k, v in best.items()}, indent=1))


def main():
    if len(sys.argv) > 2 and sys.argv[2] == "--json":
        to_json(sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 2, sys.argv[4] if len(sys.argv) > 4 else sys.argv[1])
        return
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[name_col]}  (id {r[0]})")
        d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
        for k in KEYS:
            if k in d:
                print(f"  {k:90s} {d[k][1]:>18s} {d[k][0]}")


if __name__ == "__main__":
    main()
