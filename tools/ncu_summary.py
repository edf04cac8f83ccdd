"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one block of selected metrics per kernel launch.

usage: python tools/ncu_summary.py report.ncu-rep "header comment" > profiles/<name>.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__inst_issued.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
    # instructions per cycle and pipe, in % of one per cycle and sub-partition: issue-port accounting (an FFMA2 is one
    # instruction here but holds the FMA pipe, and the port, for two cycles: compare with sm__pipe_fma_cycles_active)
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    for c in sys.argv[2:]:
        print("# " + c)
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        print("Kernel Name,%s," % d.get("Kernel Name", "?"))
        for m in METRICS:
            if m in d:
                print("%s,%s,%s" % (m, d[m], u.get(m, "")))
        print()


if __name__ == "__main__":
    main()
