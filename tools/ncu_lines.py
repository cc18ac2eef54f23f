#!/usr/bin/env python
"""Per-source-line executed-instruction counts of one kernel from an ncu report.

ncu's CSV source page lists SASS with per-instruction counters but no line numbers; `nvdisasm -g` lists
the same SASS with line markers.  This joins the two by instruction offset.

    python tools/ncu_lines.py gpurun_out/r02b_prof.ncu-rep '_Z7k_fusedILi5ELi2ELi1ELi5EEv9FusedArgs' \
        --tu ndp_tu_fused.cu -DNDP_FUSED_N=5 --queries 8388608 [--top 60]

The translation unit is recompiled here with build.py's flags (the report must come from the same sources)."""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('symbol')
    ap.add_argument('--tu', required=True)
    ap.add_argument('--queries', type=int, required=True, help='queries per launch (per-tile = per 32 queries)')
    ap.add_argument('--top', type=int, default=50)
    ap.add_argument('--launch', type=int, default=0, help='which captured launch of the kernel')
    ap.add_argument('--stalls', action='store_true', help='also rank source lines by warp-stall samples (with the reasons)')
    args, defines = ap.parse_known_args()
    from ndpolator_b200 import build
    tmp = tempfile.mkdtemp()
    cubin = os.path.join(tmp, 'k.cubin')
    subprocess.run([build._nvcc()] + [f for f in build.NVCC_FLAGS if f not in ('-Xcompiler', '-fPIC')] + defines +
                   ['-cubin', '-o', cubin, os.path.join(build.CSRC, args.tu)], check=True, capture_output=True)
    dis = subprocess.run(['nvdisasm', '-g', cubin], check=True, capture_output=True, text=True).stdout.split('\n')
    start = [i for i, l in enumerate(dis) if l.startswith(f'.text.{args.symbol}:')][0]
    insts, cur = [], None
    for l in dis[start + 1:]:
        if l.startswith('//-----'):
            break
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            insts.append((int(m.group(1), 16), cur, m.group(2)))
    sass = subprocess.run(['ncu', '-i', args.rep, '--page', 'source', '--csv'], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(sass.split('\n')))
    # reports hold one block per captured launch: "Kernel Name" row, header row, instruction rows
    blocks, hdr = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            blocks.append({'name': r[1], 'rows': []})
        elif r and r[0] == 'Address':
            hdr = r
        elif r and r[0].startswith('0x') and blocks:
            blocks[-1]['rows'].append(r)
    blk = blocks[args.launch]
    ia, ie, it = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    base = int(blk['rows'][0][ia], 16)
    cnt = {int(r[ia], 16) - base: (int(r[ie]), int(r[it])) for r in blk['rows']}
    if len(cnt) != len(insts):
        print(f'warning: report has {len(cnt)} instructions, this build {len(insts)}: sources differ?', file=sys.stderr)
    agg, aggt, byfile = collections.Counter(), collections.Counter(), collections.Counter()
    for off, cur, _ in insts:
        n, t = cnt.get(off, (0, 0))
        agg[cur] += n
        aggt[cur] += t
        byfile[cur[0] if cur else None] += n
    tiles = args.queries / 32
    total = sum(agg.values())
    print(f'{blk["name"]}: {total / tiles:.1f} warp instructions per 32-query tile '
          f'({sum(aggt.values()) / max(total, 1):.1f} active threads per instruction)')
    print('by file:', {k: round(v / tiles, 1) for k, v in byfile.most_common()})
    for k, v in agg.most_common(args.top):
        src = ''
        if k:
            path = os.path.join(build.CSRC, k[0])
            if os.path.exists(path):
                src = open(path).read().split('\n')[k[1] - 1].strip()[:100]
        print(f'{v / tiles:8.1f} {aggt[k] / max(v, 1):5.1f}  {k[0] if k else "?"}:{k[1] if k else 0}  {src}')
    if args.stalls:
        cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        per_line, reasons = collections.Counter(), collections.defaultdict(collections.Counter)
        by_off = {int(r[ia], 16) - base: r for r in blk['rows']}
        for off, cur, _ in insts:
            r = by_off.get(off)
            if r is None:
                continue
            for i, name in cols:
                try:
                    v = int(r[i])
                except ValueError:
                    v = 0
                per_line[cur] += v
                reasons[cur][name] += v
        tot = sum(per_line.values()) or 1
        print(f'\nwarp-stall samples by source line (total {tot}):')
        for k, v in per_line.most_common(args.top):
            src = ''
            if k:
                path = os.path.join(build.CSRC, k[0])
                if os.path.exists(path):
                    src = open(path).read().split('\n')[k[1] - 1].strip()[:90]
            why = ', '.join(f'{n} {100 * c / max(v, 1):.0f}%' for n, c in reasons[k].most_common(3))
            print(f'{100 * v / tot:6.2f}%  {k[0] if k else "?"}:{k[1] if k else 0}  [{why}]  {src}')


if __name__ == '__main__':
    main()
