#!/usr/bin/env python
"""bench.py -- query points/sec of the batched ndpolate hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path on the box's host cores

Workload (BASELINE.json metric / config): C4 = 5-D 40^5 sparse grid (30 % structured
void, 819 MB fp64, exceeds L2), vdim 1, 'near' query mix (56 % in a fully defined
hypercube, 44 % off-grid within ~3 cells), extrapolation_method='linear',
search_algorithm='kdtree'.  A step = one ndpolate pass over one batch of `--batch`
synthetic queries per GPU (device-resident, > L2, two alternating batches); N GPUs run
N query shards against replicated tables with no collective on the data path
(weak scaling).  One JSON line on stdout (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'query_pts_per_sec'
UNIT = 'pts/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='C4')
    ap.add_argument('--scale', type=int, default=None, help='ticks per axis (default: the size BASELINE.json names)')
    ap.add_argument('--mix', default='near', choices=['near', 'deep'])
    ap.add_argument('--method', default=None, choices=[None, 'none', 'nearest', 'linear'])
    ap.add_argument('--search', default='kdtree', choices=['kdtree', 'linear'])
    ap.add_argument('--batch', type=int, default=1 << 27, help='queries per GPU per step')
    ap.add_argument('--e2e-batch', type=int, default=1 << 24, help='queries per GPU per end-to-end (host buffer) step')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--gather', type=int, default=0, help='0 auto, 1 strided, 2 cell-major')
    ap.add_argument('--staged', type=int, default=0, help='0 auto (cp.async ring), 1 off (same results)')
    ap.add_argument('--opt', action='append', default=[], help='table option key=value (tuning sweeps), repeatable')
    ap.add_argument('--cpu-scale', type=int, default=20, help='ticks per axis of the grid the CPU reference is timed on')
    ap.add_argument('--cpu-queries', type=int, default=20000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


METHODS = {'none': 0, 'nearest': 1, 'linear': 2}
SEARCHES = {'kdtree': 0, 'linear': 1}


# ---------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def kernel_source_digest():
    """sha256 over the CUDA sources: an ncu capture is only quoted for the code it was taken on."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'ndpolator_b200', 'csrc')
    for name in sorted(os.listdir(d)):
        h.update(name.encode())
        h.update(open(os.path.join(d, name), 'rb').read())
    return h.hexdigest()[:16]


def measured_traffic(batch, workload_key):
    """DRAM bytes of one launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum from
    an `ncu --set full` capture, tools/traffic_from_ncu.py -> profiles/r02_traffic.json), scaled to this
    batch.  None unless the capture was taken on exactly these kernel sources and this workload."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
        if t.get('source_digest') != kernel_source_digest() or t.get('workload') != workload_key:
            return None
        return t['dram_bytes_per_launch'] * batch / t['queries_per_launch']
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref when it was compiled here, else the oracle port)
# ---------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup):
    """Times the reference C implementation of this path on the host cores, on a bounded
    sample: the same query mix on a `cpu_scale`^5 version of the grid (its insertion-built
    kd-tree for the full 40^5 grid needs ~15 min and 6.6 GB per process),
    tree warm (like the GPU table), all host threads sharing one table."""
    import numpy as np
    from ndpolator_b200 import synth
    from oracle import oracle as orc
    kind = 'reference' if orc.have('reference') else 'port'
    w = synth.workload(args.workload, args.cpu_scale, mix=args.mix, n_queries=args.cpu_queries)
    method = METHODS[args.method or w.method]
    search = SEARCHES[args.search]
    O = orc.Oracle(kind, w.axes, w.grid(), len(w.basic_axes))
    if method and search == 0:
        O.warm(method)
    cores = os.cpu_count() or 1
    n = args.cpu_queries
    batches = [np.ascontiguousarray(w.queries(s * n, n)) for s in range(2)]
    for s in range(warmup):
        O.ndpolate(batches[s % 2], method, search, threads=cores)
    t0 = time.perf_counter()
    for s in range(steps):
        O.ndpolate(batches[s % 2], method, search, threads=cores)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    O.ndpolate(batches[0][: max(1, n // 4)], method, search, threads=1)
    single = max(1, n // 4) / (time.perf_counter() - t1)
    return {
        'value': steps * n / dt, 'unit': UNIT, 'cores': cores, 'kind': kind,
        'single_thread_value': single,
        'sample': (f'{w.name} ({args.cpu_scale}^5 grid, {w.mix} mix, method={args.method or w.method}, '
                   f'search={args.search}), {n} queries/step x {steps} steps, kd-tree warm, {cores} host threads '
                   f'sharing one table; full 40^5 grid not used on CPU: reference kd-tree build ~15 min/6.6 GB'),
        'ms_per_step': dt / steps * 1e3,
    }


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cb = cpu_reference_run(args, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'single_thread_value')},
        'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args):
    from ndpolator_b200 import synth
    w = synth.workload(args.workload, args.scale, mix=args.mix)
    return {
        'workload': (f'{w.name}: {w.naxes}-D grid {"x".join(str(len(a)) for a in w.axes)}, vdim {w.vdim}, '
                     f'{w.mix} query mix, extrapolation_method={args.method or w.method}, search_algorithm={args.search}'),
        'queries_per_gpu_per_step': args.batch,
        'bytes_per_query_algorithmic': w.bytes_per_query,
        'l2': 'inputs larger than L2 (two alternating device-resident batches)',
        'parallelism': f'query shards x{args.gpus}, table replicated, no data-path collective',
    }


# ---------------------------------------------------------------------------
# the CUDA path
# ---------------------------------------------------------------------------
def gen_queries(w, first, count, device, piece=1 << 23):
    import torch
    out = torch.empty((count, w.naxes), dtype=torch.float64, device=device)
    for lo in range(0, count, piece):
        c = min(piece, count - lo)
        out[lo:lo + c] = w.queries(first + lo, c, device)
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ndpolator_b200 import synth
    from ndpolator_b200.cndpolator import Table
    from ndpolator_b200.sharding import world

    rank, ws, local = world()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if ws > 1:
        dist.init_process_group('nccl', device_id=dev)

    w = synth.workload(args.workload, args.scale, mix=args.mix)
    method = METHODS[args.method or w.method]
    search = SEARCHES[args.search]
    B, K, W = args.batch, args.steps, args.warmup

    t_build = time.perf_counter()
    grid = w.grid(dev)
    T = Table(w.axes, grid, len(w.basic_axes), device=local)
    del grid
    if args.gather:
        T.set_option('gather', args.gather)
    if args.staged:
        T.set_option('staged', args.staged)
    for kv in args.opt:
        k, v = kv.split('=')
        T.set_option(k, int(v))
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build

    # per-rank, per-buffer distinct rows of the global synthetic batch
    qs = [gen_queries(w, (b * ws + rank) * B, B, dev) for b in range(2)]
    interps = torch.empty((B, w.vdim), dtype=torch.float64, device=dev)
    dists = torch.empty((B, 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(dev)                # kernels and timing events share this stream
    T.set_option('stream', stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        T.ndpolate(qs[s % 2], method, search, out=(interps, dists))

    keys = ('launches', 'offgrid', 'hard', 'ties', 'slow', 'fused', 'main_ns', 'search_ns', 'finish_ns')
    acc = dict.fromkeys(keys, 0)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(K):
        T.ndpolate(qs[s % 2], method, search, out=(interps, dists))
        for k in keys:
            acc[k] += T.stat(k)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = ws * B * K / (ms_total / 1e3)
    checksum = float(torch.nan_to_num(interps).sum().item())

    # end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    T.set_option('stream', 0)
    nb_e2e = min(args.e2e_batch, B)
    qh = torch.empty((nb_e2e, w.naxes), dtype=torch.float64).pin_memory()
    qh.copy_(qs[0][:nb_e2e])
    ih = torch.empty((nb_e2e, w.vdim), dtype=torch.float64).pin_memory()
    dh = torch.empty((nb_e2e, 1), dtype=torch.float64).pin_memory()
    qn, inn, dn = qh.numpy(), ih.numpy(), dh.numpy()
    T.ndpolate(qn, method, search, out=(inn, dn))
    barrier()
    t0 = time.perf_counter()
    for s in range(args.e2e_steps):
        T.ndpolate(qn, method, search, out=(inn, dn))
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    h2d, d2h = T.stat('h2d_bytes'), T.stat('d2h_bytes')
    if ws > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = ws * nb_e2e * args.e2e_steps / float(e2e_s.item())
    # the pipelined host path (chunks, two slots, three streams) against a device-resident run of the same rows
    ref_i, ref_d = T.ndpolate(qs[0][:nb_e2e], method, search)
    same = bool(torch.equal(ih.view(torch.int64), ref_i.cpu().view(torch.int64)) and
                torch.equal(dh.view(torch.int64), ref_d.cpu().view(torch.int64)))
    if not same:
        raise SystemExit('bench.py: end-to-end (host buffer) results differ from the device-resident run')

    if ws > 1:    # the only collective besides timing: a final gather of per-rank checksums to rank 0
        sums = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(ws)] if rank == 0 else None
        dist.gather(torch.tensor([checksum], dtype=torch.float64, device=dev), sums, dst=0)
        if rank == 0:
            checksum = float(sum(t.item() for t in sums))

    if rank == 0:
        peak, peak_src = measured_peak()
        n_fdhc = B * K - acc['offgrid']
        main_s = acc['main_ns'] / 1e9 / K
        cell_bytes = (1 << w.naxes) * w.vdim * 8
        out_bytes = w.vdim * 8 + 8
        fused = acc['fused'] == K
        if fused:
            # k_fused, per launch: every query row read (naxes*8), ONE gather per query -- its own cell, or for
            # an extrapolated query the cell below the found hypercube ('linear') / the found vertex ('nearest')
            # -- and every interps row and dist written; queries handed to k_slow gather nothing here
            n_slow = acc['slow'] / K
            off = acc['offgrid'] / K - n_slow
            gather = cell_bytes if method == 2 else (w.vdim * 8 if method == 1 else 0)
            main_bytes = B * w.naxes * 8 + (n_fdhc / K) * cell_bytes + off * gather + (B - n_slow) * out_bytes
            kernel_name = f'k_fused<{w.naxes},{method},{(0 if search == 0 else 2) + (1 if method == 2 else 0)}>'
        else:
            # k_gather_staged / k_main, per launch: every query row read + for the queries it completes (fully
            # defined hypercube) the 2^n-corner gather, the interps row and the dist
            main_bytes = B * w.naxes * 8 + (n_fdhc / K) * (cell_bytes + out_bytes)
            kernel_name = 'k_gather_staged<N,0> / k_main (fused import + cell gather + reduce, first of three passes)'
        achieved = main_bytes / main_s / 1e9 if main_s > 0 else None
        wl_key = f'{args.workload}@{args.scale}/{args.mix}/{args.method or w.method}/{args.search}'
        traffic = measured_traffic(B, wl_key) if fused else None
        step_gbs = B * w.bytes_per_query / (ms_total / 1e3 / K) / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': ws, 'steps': K, 'warmup': W,
            'ms_per_step': ms_total / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'queries_per_gpu_per_step': nb_e2e, 'steps': args.e2e_steps, 'host_memory': 'pinned',
                    'matches_device_run': same},
            'gpu_launches': acc['launches'],
            'clocks': clocks,
            'roofline': {
                'bound': 'hbm', 'kernel': kernel_name,
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': (achieved / peak) if achieved else None,
                'traffic': traffic,
                'traffic_source': ('profiles/r02_traffic.json (ncu --set full on these sources, scaled to this batch)'
                                   if traffic else 'no ncu capture of exactly these kernel sources / this workload'),
                'peak_source': peak_src,
                'algorithmic_bytes_per_launch': main_bytes, 'avg_launch_ms': main_s * 1e3,
                'whole_step': {'achieved': step_gbs, 'frac': step_gbs / peak,
                               'note': 'queries/step x SURVEY 8d bytes/query over the whole step (all kernels)'},
                'phase_ms_per_step': {'k_main': acc['main_ns'] / 1e6 / K, 'k_search': acc['search_ns'] / 1e6 / K,
                                      'k_finish': acc['finish_ns'] / 1e6 / K},
            },
            'mix_measured': {'in_fully_defined_hypercube': n_fdhc / (B * K), 'offgrid': acc['offgrid'] / (B * K),
                             'needed_lattice_descent': acc['hard'] / (B * K), 'kd_ties': acc['ties'],
                             'left_to_exact_generic_kernel': acc['slow'] / (B * K)},
            'path': 'fused (k_fused + k_slow)' if fused else 'three-pass (gather, search, finish)',
            'table': {'cellmajor': bool(T.stat('cellmajor')), 'build_s': build_s},
            'checksum': checksum,
        }
        if ws == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_run(args, steps=2, warmup=1)
                line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'single_thread_value')}
            except Exception as e:       # the baseline is reported beside the number, never instead of it
                line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                                        'sample': f'failed: {e!r}'}
        print(json.dumps(line))
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    # Only the JSON line may reach stdout: libraries (NCCL's version banner, for one) write to
    # fd 1, so everything else is sent to stderr and the line is written to the saved descriptor.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, 'w')
    global print
    import builtins

    def print(*a, **k):          # noqa: A001 -- the JSON line goes to the real stdout
        k.setdefault('file', real_stdout)
        builtins.print(*a, **k)
        real_stdout.flush()

    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
