#!/usr/bin/env python
"""Print the handful of ncu metrics that matter for these kernels from a .ncu-rep (details + raw pages)."""
import csv
import subprocess
import sys

KEEP = ['Duration', 'Registers Per Thread', 'Theoretical Occupancy', 'Achieved Occupancy', 'Executed Ipc Active',
        'Issue Slots Busy', 'Memory Throughput', 'DRAM Throughput', 'L1/TEX Hit Rate', 'L2 Hit Rate', 'Mem Busy',
        'Compute (SM) Throughput', 'Avg. Active Threads Per Warp', 'Avg. Not Predicated Off Threads Per Warp',
        'No Eligible', 'Eligible Warps Per Scheduler', 'Active Warps Per Scheduler',
        'Warp Cycles Per Issued Instruction', 'Grid Size', 'Block Size', 'Block Limit Registers',
        'Block Limit Shared Mem', 'Executed Instructions']
RAW = ['l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
       'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
       'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
       'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
       'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct',
       'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
       'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    rows = page(rep, 'details')
    ix = {h: i for i, h in enumerate(rows[0])}
    for r in rows[1:]:
        n = r[ix['Metric Name']]
        if n in KEEP:
            print(f"{n:45s} {r[ix['Metric Value']]:>16s} {r[ix['Metric Unit']]}")
    rows = page(rep, 'raw')
    hdr, units, vals = rows[0], rows[1], rows[2] if len(rows) > 2 else rows[1]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    for k in RAW:
        if k in d:
            print(f"{k:95s} {d[k]} {u.get(k, '')}")


if __name__ == '__main__':
    main()
