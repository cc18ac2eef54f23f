#!/usr/bin/env python
"""Turns an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the compact per-kernel
summary committed under profiles/.   usage: python profiles/summarize.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
    ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'L1 global-load sectors'),
    ('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'L1 global-load requests'),
    ('lts__t_sectors_srcunit_tex_op_read.sum', 'L2 read sectors (from SMs)'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads/instr'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64 pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
    ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'stall no_instruction'),
    ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'stall branch_resolving'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg_throttle'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    print(f'# {path}  (ncu --set full --clock-control none; per-launch values, cold-cache, serialised)')
    for r in rows[2:]:
        print(f'\n## {r[ki]}')
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f'  {label:28s} {r[i]:>22s} {units[i]}')


if __name__ == '__main__':
    main(sys.argv[1])
