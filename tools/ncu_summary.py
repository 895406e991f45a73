#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small JSON: one entry per captured launch with the
metrics the roofline discussion needs.  Usage: python tools/ncu_summary.py rep.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    res = []
    for r in data:
        e = {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for k in KEEP:
            if k in col and r[col[k]] != "":
                e[k] = [r[col[k]], units[col[k]]]
        res.append(e)
    json.dump(res, open(out, "w"), indent=1)
    for e in res:
        def g(k):
            return e.get(k, ["-"])[0]
        print(f"{e['kernel'][:60]:60s} t={g('gpu__time_duration.sum')}{e.get('gpu__time_duration.sum',['',''])[1]} "
              f"dramR={g('dram__bytes_read.sum')} dramW={g('dram__bytes_write.sum')} dram%={g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} "
              f"issue%={g('smsp__issue_active.avg.pct_of_peak_sustained_active')} regs={g('launch__registers_per_thread')} "
              f"warps%={g('sm__warps_active.avg.pct_of_peak_sustained_active')}")


if __name__ == "__main__":
    main()
