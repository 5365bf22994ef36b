#!/usr/bin/env python
"""Prints the handful of ncu metrics this project reads from a .ncu-rep (raw page): python tools/ncu_summary.py rep [rep ...]"""
import csv, subprocess, sys, io
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_requests_srcunit_tex.sum", "lts__t_sectors_srcunit_tex.sum",
        "l1tex__m_xbar2l1tex_read_sectors.sum", "l1tex__m_l1tex2xbar_write_sectors.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", rep, "|", r[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                print(f"   {w:90s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
