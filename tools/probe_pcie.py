#!/usr/bin/env python
"""Host<->device copy ceiling of the box for the end-to-end leg of bench.py: every rank (one per GPU, torchrun)
streams pinned host memory to its GPU and back with plain cudaMemcpyAsync -- H2D alone, D2H alone, and both at
the ndpolate ratio (40 bytes in : 16 bytes out per C4 query) -- for about a second each, with and without binding
the rank to its GPU's NUMA node.  Rank 0 prints one JSON line per mode with the aggregate GB/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/probe_pcie.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench


def main():
    rank, ws, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if ws > 1:
        dist.init_process_group('nccl', device_id=dev)
    for bind in (False, True):
        note = bench.bind_to_gpu_numa_node(local) if bind else 'unbound'
        nin, nout = 40 << 24, 16 << 24                       # one "step" of 2^24 C4 queries: 671 MB in, 268 MB out
        hin = torch.empty(nin, dtype=torch.uint8).pin_memory()
        hout = torch.empty(nout, dtype=torch.uint8).pin_memory()
        hin.fill_(1)
        din = torch.empty(nin, dtype=torch.uint8, device=dev)
        dout = torch.zeros(nout, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        for mode in ('h2d', 'd2h', 'both'):
            torch.cuda.synchronize()
            if ws > 1:
                dist.barrier()
            t0 = time.perf_counter()
            steps = 0
            while time.perf_counter() - t0 < 1.0:
                if mode in ('h2d', 'both'):
                    with torch.cuda.stream(s1):
                        din.copy_(hin, non_blocking=True)
                if mode in ('d2h', 'both'):
                    with torch.cuda.stream(s2):
                        hout.copy_(dout, non_blocking=True)
                torch.cuda.synchronize()
                steps += 1
            dt = time.perf_counter() - t0
            nbytes = steps * ((nin if mode != 'd2h' else 0) + (nout if mode != 'h2d' else 0))
            t = torch.tensor([nbytes / dt / 1e9], dtype=torch.float64, device=dev)
            if ws > 1:
                lo = t.clone()
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            else:
                lo = t
            if rank == 0:
                print(json.dumps({'probe': 'pinned cudaMemcpyAsync', 'n_gpus': ws, 'mode': mode, 'numa': note if bind else 'unbound',
                                  'aggregate_GBps': float(t.item()), 'slowest_rank_GBps': float(lo.item()),
                                  'c4_queries_per_s_at_this_rate': float(t.item()) * 1e9 / (56 if mode == 'both' else (40 if mode == 'h2d' else 16))}), flush=True)
        del hin, hout
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
