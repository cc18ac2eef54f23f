"""Debug aid: C4 full size, method NONE, fused kernel run repeatedly against the strided three-pass path and the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ndpolator_b200 import synth
from ndpolator_b200.cndpolator import Table
from oracle.oracle import Oracle

w = synth.workload('C4', None)
grid_d = w.grid('cuda')
T = Table(w.axes, grid_d, 5)
n = 1 << 22
q = w.queries(0, n, 'cuda')
runs = []
for rep in range(6):
    gi, gd = T.ndpolate(q, 0, 0)
    runs.append(gi.clone())
T.set_option('gather', 1)
si, sd = T.ndpolate(q, 0, 0)
for rep, gi in enumerate(runs):
    bad = torch.nonzero(gi.view(torch.int64)[:, 0] != si.view(torch.int64)[:, 0])[:, 0]
    print('rep', rep, 'rows differing from the strided path:', bad.numel(), bad[:10].tolist())
    if bad.numel():
        rows = bad[:5].cpu().numpy()
        qq = q[bad[:5]].cpu().numpy()
        O = Oracle('port', w.axes, w.grid(), 5)
        oi, od = O.ndpolate(qq, 0, 0)
        for k, r in enumerate(rows):
            print('  row', r, 'tile', r // 32, 'lane', r % 32, 'q', qq[k], 'fused', gi[r].item(), 'strided', si[r].item(), 'oracle', oi[k, 0])
        break
for a in range(1, len(runs)):
    print('fused run', a, 'equals run 0:', bool(torch.equal(runs[a].view(torch.int64), runs[0].view(torch.int64))))
