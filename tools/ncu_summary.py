#!/usr/bin/env python3
"""Condense an `ncu --set full` report into the metric,value,unit table kept under profiles/.
usage: ncu_summary.py report.ncu-rep [kernel-substring] > profiles/<name>_summary.csv   (last matching launch)"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
cand = [r for r in rows[2:] if pat in r[ki]]
r = cand[-1]
w = csv.writer(sys.stdout)
w.writerow(["metric", "value", "unit"])
w.writerow(["Kernel Name", r[ki], ""])
col = {h: i for i, h in enumerate(hdr)}
for k in KEYS:
    if k in col:
        w.writerow([k, r[col[k]], units[col[k]]])
for h in hdr:
    if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
        try:
            if float(r[col[h]]) >= 0.3:
                w.writerow([h, r[col[h]], units[col[h]]])
        except ValueError:
            pass
