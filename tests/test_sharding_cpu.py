"""CPU, world_size 2 over gloo: query sharding + final gather reproduce the unsharded
result (the local engine here is the CPU oracle; on the GPU box it is the CUDA path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    from ndpolator_b200.sharding import shard_range
    for n in (0, 1, 7, 1000, 10 ** 9 + 3):
        for ws in (1, 2, 3, 8):
            cuts = [shard_range(n, r, ws) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(ws - 1))
            sizes = [h - l for l, h in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, ws, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from cases import build_cases
    from ndpolator_b200 import synth
    from ndpolator_b200.sharding import ndpolate_sharded
    from oracle.oracle import Oracle
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    w = synth.workload('C4', 8, n_queries=5001)
    O = Oracle('port', w.axes, w.grid(), 5)
    res = ndpolate_sharded(lambda q: O.ndpolate(q, 2, 0), w.n_queries, lambda lo, hi: w.queries(lo, hi - lo))
    if rank == 0:
        full = O.ndpolate(w.queries(0, w.n_queries), 2, 0)
        ok = (np.array_equal(res[0].numpy().view(np.uint64), full[0].view(np.uint64)) and
              np.array_equal(res[1].numpy().view(np.uint64), full[1].view(np.uint64)))
        open(out_path, 'w').write('ok' if ok else 'mismatch')
    else:
        assert res == (None, None)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_equals_unsharded(tmp_path):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'result.txt')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == 'ok'
