"""Small end-to-end exercise of every kernel family for compute-sanitizer (development tool).
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from cases import adversarial_queries, build_cases  # noqa: E402
from ndpolator_b200.cndpolator import Table  # noqa: E402

QUICK = os.environ.get('NDPB_SANITIZE_QUICK') == '1'      # one pass per case (memcheck is ~50x slower than native)
for name, axes, nbasic, grid in build_cases():
    if name not in ('holed3d', '2basic_1assoc_v2', '5d_corner_void', '7d_runtime_n', 'c3_like_v16', '2basic_2assoc_v1'):
        continue
    q = adversarial_queries(axes, 600 if QUICK else 3000, np.random.default_rng(1))
    for gather, staged, fused in (((2, 0, 0), (1, 0, 1)) if QUICK else ((1, 0, 0), (2, 0, 0), (2, 0, 1), (2, 1, 1), (2, 2, 1))):
        T = Table(axes, grid, nbasic)
        T.set_option('gather', gather)
        T.set_option('staged', staged)
        T.set_option('fused', fused)
        T.set_option('chunk', 1024)
        for m in (0, 1, 2):
            for s in (0, 1):
                T.ndpolate(q, m, s)
        T.find_nearest(q, 1, 0)
        T.find_nearest(q, 2, 1)
        T.distance(q)
        idx, flg, nrm = T.find(q)
        T.hypercubes_raw(nrm, idx, flg)
        for m in (0, 1, 2):
            T.ndpolate_cached(nrm, idx, flg, m, 0)           # k_fused<..., CACHED> / k_fused_vec / k_slow on cached products
        if len(axes) <= 6 and grid.shape[-1] == 1 and gather != 1:
            for order in range(1, len(axes) + 1):            # every blocked layout k_fused gathers from
                T.set_option('block_order', order)
                T.ndpolate(q, 2, 0)
        import torch
        qd = torch.as_tensor(q, device='cuda')
        T.set_option('async', 1)                             # stream-ordered calls
        outs = [T.ndpolate(qd, m, 0) for m in (0, 1, 2)]
        T.sync()
        T.set_option('async', 0)
        del outs
        T.masks()
        T.tree_topology(1)
        T.close()
    print(name, 'ok')
