#!/usr/bin/env python
"""Device-resident hypercube-caching path against the plain call (one GPU, CUDA events):
import_query_pts once, then ndpolate_cached from the kept products -- e.g. for several tables on the same axes.

    python tools/bench_cached.py [--workload C4] [--scale N] [--n 67108864]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C4')
ap.add_argument('--scale', type=int, default=None)
ap.add_argument('--n', type=int, default=1 << 26)
ap.add_argument('--steps', type=int, default=5)
args = ap.parse_args()

import torch  # noqa: E402
from ndpolator_b200 import synth  # noqa: E402
from ndpolator_b200.cndpolator import Table  # noqa: E402

w = synth.workload(args.workload, args.scale)
dev = torch.device('cuda', 0)
grid = w.grid(dev)
torch.cuda.empty_cache()
T = Table(w.axes, grid, len(w.basic_axes), device=0)
del grid
q = torch.cat([w.queries(lo, min(1 << 23, args.n - lo), dev) for lo in range(0, args.n, 1 << 23)])
method = {'none': 0, 'nearest': 1, 'linear': 2}[w.method]


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps, out


ms_find, (idx, flg, nrm) = timed(lambda: T.find(q))
ms_plain, plain = timed(lambda: T.ndpolate(q, method, 0))
ms_cached, cached = timed(lambda: T.ndpolate_cached(nrm, idx, flg, method, 0))
same = bool(torch.equal(plain[0].view(torch.int64), cached[0].view(torch.int64)) and
            torch.equal(plain[1].view(torch.int64), cached[1].view(torch.int64)))
print(json.dumps({'workload': w.name, 'queries': args.n, 'method': w.method, 'find_ms': ms_find, 'ndpolate_ms': ms_plain,
                  'ndpolate_cached_ms': ms_cached, 'cached_pts_per_s': args.n / ms_cached * 1e3,
                  'plain_pts_per_s': args.n / ms_plain * 1e3, 'cached_fused': T.stat('fused'), 'identical_bits': same}))
