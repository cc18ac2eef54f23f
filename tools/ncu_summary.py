#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: per captured launch the metrics the roofline needs
(duration, DRAM bytes and % of peak, occupancy, issue rate, registers, stall mix), and -- with
--traffic-json -- profiles/r02_traffic.json for bench.py (DRAM bytes per launch, tied to the kernel
sources by digest and to the workload).

    python tools/ncu_summary.py gpurun_out/X.ncu-rep --queries 8388608 --workload 'C4@None/near/linear/kdtree' \
        [--traffic-json profiles/r02_traffic.json] > profiles/r02x_ncu.txt"""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__cycles_active.avg']


def to_bytes(v, unit):
    f = float(v.replace(',', ''))
    return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--queries', type=int, required=True)
    ap.add_argument('--workload', default='')
    ap.add_argument('--traffic-json')
    args = ap.parse_args()
    raw = subprocess.run(['ncu', '-i', args.rep, '--page', 'raw', '--csv'], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split('\n')))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    src = subprocess.run(['ncu', '-i', args.rep, '--page', 'source', '--csv'], check=True, capture_output=True, text=True).stdout
    srows = list(csv.reader(src.split('\n')))
    stalls, shdr = [], None
    for r in srows:
        if r and r[0] == 'Kernel Name':
            stalls.append(collections.Counter())
        elif r and r[0] == 'Address':
            shdr = r
        elif r and r[0].startswith('0x') and stalls:
            for i, h in enumerate(shdr):
                if h.startswith('stall_') and 'Not Issued' not in h:
                    try:
                        stalls[-1][h] += int(r[i])
                    except ValueError:
                        pass
    traffic = None
    for k, r in enumerate(x for x in rows[2:] if len(x) > 10):
        name = r[col['Kernel Name']]
        print(f'== launch {k}: {name}   ({args.queries} queries, {args.workload})')
        for key in KEYS:
            if key in col:
                print(f'   {key:75s} {r[col[key]]:>16s} {units[col[key]]}')
        rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
        wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
        inst = float(r[col['smsp__inst_executed.sum']].replace(',', ''))
        print(f'   -> DRAM bytes per query: {(rd + wr) / args.queries:.1f} (read {rd / args.queries:.1f}, write {wr / args.queries:.1f}); '
              f'warp instructions per 32-query tile: {inst / (args.queries / 32):.0f}')
        if k < len(stalls):
            tot = sum(stalls[k].values()) or 1
            print('   stall samples: ' + ', '.join(f'{h[6:]} {100 * v / tot:.0f}%' for h, v in stalls[k].most_common(7)))
        traffic = {'kernel': name, 'queries_per_launch': args.queries, 'dram_bytes_per_launch': rd + wr,
                   'dram_read': rd, 'dram_write': wr, 'workload': args.workload}
    if args.traffic_json and traffic:
        import bench
        traffic['source_digest'] = bench.kernel_source_digest()
        traffic['kernel_mangled'] = bench.mangled_fused(traffic['kernel'])
        traffic['sass_digest'] = bench.kernel_sass_digest(traffic['kernel_mangled'])
        traffic['report'] = os.path.basename(args.rep)
        json.dump(traffic, open(args.traffic_json, 'w'), indent=1)


if __name__ == '__main__':
    main()
