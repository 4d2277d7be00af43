"""Summarise an .ncu-rep (read here, without a GPU): python tools/ncu_summary.py file.ncu-rep [substring ...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64",
        "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled", "local_load", "local_store", "smsp__inst_executed_op_local",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fp64_pred_on.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "issue_stalled"] + extra
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if any(w in h for w in want):
            print(f"{h:110s} {units[i]:14s} {r[i]}")
