#!/usr/bin/env python
"""bench.py -- query points/sec of the batched ndpolate hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path on the box's host cores

Workload (BASELINE.json metric / config): C4 = 5-D 40^5 sparse grid (30 % structured
void, 819 MB fp64, exceeds L2), vdim 1, 'near' query mix (56 % in a fully defined
hypercube, 44 % off-grid within ~3 cells), extrapolation_method='linear',
search_algorithm='kdtree'.  A step = one ndpolate pass over the configuration's 10^9
synthetic queries (device-resident, 40 GB >> L2), query-sharded over the N GPUs of the
run against replicated tables with no collective on the data path (strong scaling: the
total is fixed, each GPU takes 10^9 / N).  One JSON line on stdout (rank 0)."""
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
    ap.add_argument('--variant', default=None, choices=[None, 'iid', 'allbasic'],
                    help="SURVEY 8d variants: 'iid' = C4 with i.i.d. 30 %% voids (mask depends on all axes, ~100 %% off-grid), "
                         "'allbasic' = C3 with mu as a fourth basic axis")
    ap.add_argument('--spacing', type=float, default=None,
                    help='tick spacing of the synthetic axes (default: 1, the integers); e.g. 0.3 exercises the general '
                         'index search with IEEE divisions instead of the exact power-of-two-spacing multiply')
    ap.add_argument('--method', default=None, choices=[None, 'none', 'nearest', 'linear'])
    ap.add_argument('--search', default='kdtree', choices=['kdtree', 'linear'])
    ap.add_argument('--total', type=int, default=None,
                    help='queries per step over ALL GPUs (default: the number BASELINE.json names for the workload, 10^9 for C4)')
    ap.add_argument('--batch', type=int, default=None, help='queries per GPU per step (overrides --total: weak scaling)')
    ap.add_argument('--sweep-totals', default=None,
                    help='comma-separated totals (e.g. 1e3,1e5,1e7,1e9): one JSON line per total on the same table; steps of '
                         '2^24 queries per GPU and more are timed over at most 10 steps')
    ap.add_argument('--e2e-batch', type=int, default=1 << 26, help='queries per GPU per end-to-end (host buffer) step')
    ap.add_argument('--e2e-steps', type=int, default=20)
    ap.add_argument('--gather', type=int, default=0, help='0 auto, 1 strided, 2 cell-major')
    ap.add_argument('--staged', type=int, default=0, help='0 auto (cp.async ring), 1 off (same results)')
    ap.add_argument('--opt', action='append', default=[], help='table option key=value (tuning sweeps), repeatable')
    ap.add_argument('--cpu-scale', type=int, default=None,
                    help='ticks per axis of the grid the CPU reference is timed on (default: 24 for C4 -- BASELINE.md 4.5 -- else full size)')
    ap.add_argument('--cpu-queries', type=int, default=20000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--api-mode', default='auto', choices=['auto', 'sync', 'async'],
                    help="device-resident timed steps: 'sync' = every call returns with its results complete (the reference's "
                         "semantics), 'async' = stream-ordered calls (table option async=1) and ONE ndpb_table_sync after the K "
                         "steps, inside the timed region; auto = async for steps of at most 2^24 queries per GPU, where the "
                         "per-call host round trip is a visible share of the step")
    ap.add_argument('--presort', action='store_true',
                    help='experiment (SURVEY 8 f3): order each device-resident batch by the cell its query falls into before '
                         'timing -- the ceiling of what a query-binning front end could give the gather; the time of the '
                         'sort itself (torch.sort of the cell keys + the row gather) is reported as presort_ms')
    ap.add_argument('--inprocess', action='store_true',
                    help='ONE process, --gpus N devices through ndpb_group (one ndpolate call split over N GPUs): end-to-end '
                         'from pinned host arrays only; launch with plain python, not torchrun')
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


def mangled_fused(name):
    """'void k_fused<5, 2, 1, 4, 1, 3, 0>(FusedArgs)' (ncu's kernel name) -> its mangled symbol."""
    import re
    m = re.search(r'k_fused<([0-9, ]+)>', name or '')
    if not m:
        return None
    return '_Z7k_fusedI' + ''.join(f'Li{int(x)}E' for x in m.group(1).split(',')) + 'Ev9FusedArgs'


def kernel_sass_digest(mangled, lib=None):
    """sha256 over the SASS instruction text of one kernel of the built library (cuobjdump -sass -fun): a traffic
    capture stays valid exactly as long as the machine code of the kernel it was taken on is unchanged."""
    import hashlib
    import re
    if not mangled:
        return None
    lib = lib or os.environ.get('NDPB_LIB') or os.path.join(ROOT, 'ndpolator_b200', 'libndpb200.so')
    try:
        out = subprocess.run(['cuobjdump', '-sass', '-fun', mangled, lib], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    ins = [re.sub(r'/\*[0-9a-f]+\*/', '', l).strip() for l in out.splitlines() if re.match(r'^\s+/\*[0-9a-f]{4,}\*/', l)]
    if not ins:
        return None
    return hashlib.sha256('\n'.join(ins).encode()).hexdigest()[:16]


def measured_traffic(batch, workload_key):
    """DRAM bytes of one launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum from an
    `ncu --set full` capture, tools/ncu_summary.py --traffic-json -> profiles/r02_traffic.json), scaled to this batch.
    None unless the capture was taken on exactly this kernel's machine code (SASS digest of the built library; the
    digest of the CUDA sources when the capture carries no SASS digest) and this workload."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
        if t.get('workload') != workload_key:
            return None
        if t.get('sass_digest'):
            if kernel_sass_digest(t.get('kernel_mangled')) != t['sass_digest']:
                return None
        elif t.get('source_digest') != kernel_source_digest():
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
def _pool_worker(args):
    """Process-pool worker (fork): the parent's warm table is shared copy-on-write, as BASELINE.md 4.5 allows."""
    lo, hi = args
    O, q, method, search = _POOL_STATE
    i, d = O.ndpolate(q[lo:hi], method, search)
    return float(i[~(i != i)].sum())


_POOL_STATE = None


def cpu_scale_of(args):
    if args.cpu_scale is not None:
        return args.cpu_scale
    return 24 if (args.workload == 'C4' and args.scale is None) else args.scale


def cpu_reference_run(args, steps, warmup):
    """Times the reference C implementation of this path (oracle/_ref = the unmodified reference sources; the
    oracle port only where they could not be compiled) on the host cores, on a bounded sample of the workload:
    the same query mix, for C4 on a 24^5 version of the grid (the reference's insertion-built kd-tree for the
    full 40^5 grid takes ~12 min and ~8 GB per process to build; BASELINE.md 4.5 names this substitution).
    Reported: cold first call (masks + kd-tree build inside), warm single thread as shipped, all host threads
    sharing one warm table, and a fork()ed process pool sharing it copy-on-write."""
    global _POOL_STATE
    import multiprocessing as mp
    import numpy as np
    from ndpolator_b200 import synth
    from oracle import oracle as orc
    kind = 'reference' if orc.have('reference') else 'port'
    scale = cpu_scale_of(args)
    w = synth.workload(args.workload, scale, mix=args.mix, n_queries=args.cpu_queries, spacing=args.spacing, variant=args.variant)
    method = METHODS[args.method or w.method]
    search = SEARCHES[args.search]
    cores = os.cpu_count() or 1
    n = args.cpu_queries
    batches = [np.ascontiguousarray(w.queries(s * n, n)) for s in range(2)]
    t0 = time.perf_counter()
    O = orc.Oracle(kind, w.axes, w.grid(), len(w.basic_axes))
    register_s = time.perf_counter() - t0
    n_cold = max(1, n // 8)
    t0 = time.perf_counter()
    O.ndpolate(batches[1][:n_cold], method, search, threads=1)          # lazy kd-tree build happens here
    cold_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.ndpolate(batches[0][: max(1, n // 4)], method, search, threads=1)
    single = max(1, n // 4) / (time.perf_counter() - t0)
    for s in range(warmup):
        O.ndpolate(batches[s % 2], method, search, threads=cores)
    t0 = time.perf_counter()
    for s in range(steps):
        O.ndpolate(batches[s % 2], method, search, threads=cores)
    dt = time.perf_counter() - t0
    threads_value = steps * n / dt
    pool_value = None
    try:
        _POOL_STATE = (O, batches[0], method, search)
        cuts = np.linspace(0, n, cores + 1).astype(int)
        with mp.get_context('fork').Pool(cores) as pool:
            pool.map(_pool_worker, [(int(cuts[k]), int(cuts[k + 1])) for k in range(cores)])       # warm the workers
            t0 = time.perf_counter()
            reps = max(1, min(steps, 4))
            for _ in range(reps):
                pool.map(_pool_worker, [(int(cuts[k]), int(cuts[k + 1])) for k in range(cores)])
            pool_value = reps * n / (time.perf_counter() - t0)
    except Exception:
        pool_value = None
    finally:
        _POOL_STATE = None
    grid_txt = 'x'.join(str(len(a)) for a in w.axes)
    full = synth.workload(args.workload, args.scale, mix=args.mix, spacing=args.spacing, variant=args.variant)
    subst = ('' if len(full.axes[0]) == len(w.axes[0]) else
             f' -- SUBSTITUTED for the {"x".join(str(len(a)) for a in full.axes)} grid, whose reference kd-tree takes ~12 min '
             f'and ~8 GB per process to build (BASELINE.md 4.5)')
    return {
        'value': max(threads_value, pool_value or 0.0), 'unit': UNIT, 'cores': cores, 'kind': kind,
        'single_thread_value': single, 'threads_value': threads_value, 'process_pool_value': pool_value,
        'cold_first_call_s': cold_s, 'cold_first_call_queries': n_cold, 'register_s': register_s,
        'grid': grid_txt,
        'sample': (f'{w.name}: grid {grid_txt}{subst}; {w.mix} mix, method={args.method or w.method}, search={args.search}; '
                   f'{n} queries/step x {steps} steps on {cores} host threads sharing one warm table (value = best of '
                   f'threads / fork()ed process pool of {cores}); single thread as shipped and the cold first call '
                   f'(masks + kd-tree build) reported beside it'),
        'ms_per_step': n / max(threads_value, pool_value or 0.0) * 1e3,
    }


CPU_KEYS = ('value', 'unit', 'cores', 'kind', 'sample', 'grid', 'single_thread_value', 'threads_value', 'process_pool_value',
            'cold_first_call_s', 'cold_first_call_queries', 'register_s')


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cb = cpu_reference_run(args, args.steps, args.warmup)
    cfg = workload_config(args)
    cfg['workload'] += f' [CPU arm timed on grid {cb["grid"]}: see cpu_baseline.sample]'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': cfg,
        'cpu_baseline': {k: cb[k] for k in CPU_KEYS},
        'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def default_total(w):
    return {'C1': 10**6, 'C2': 10**7, 'C3': 10**8, 'C4': 10**9, 'C5': 10**8}[w.name.split('@')[0].split('/')[0]]


def batch_of(args, w, ws):
    """(queries per GPU per step, scaling): --batch fixes the per-GPU work (weak); otherwise the total is
    fixed -- BASELINE.json's number for the configuration -- and sharded (strong)."""
    if args.batch:
        return args.batch, 'weak'
    total = args.total or default_total(w)
    return max(1, total // ws), 'strong'


def workload_config(args):
    from ndpolator_b200 import synth
    w = synth.workload(args.workload, args.scale, mix=args.mix, spacing=args.spacing, variant=args.variant)
    B, scaling = batch_of(args, w, max(1, args.gpus))
    nbuf = 1 if B * w.naxes * 8 >= (1 << 30) else 2
    return {
        'workload': (f'{w.name}: {w.naxes}-D grid {"x".join(str(len(a)) for a in w.axes)}, vdim {w.vdim}, '
                     f'{w.mix} query mix, extrapolation_method={args.method or w.method}, search_algorithm={args.search}'),
        'queries_per_step_all_gpus': B * max(1, args.gpus),
        'queries_per_gpu_per_step': B,
        'bytes_per_query_algorithmic': w.bytes_per_query,
        'l2': (f'inputs larger than L2: one device-resident batch of {B * w.naxes * 8 / 1e9:.1f} GB per GPU re-read every step'
               if nbuf == 1 else 'inputs larger than L2 where the batch allows: two alternating device-resident batches'),
        'parallelism': f'query shards x{args.gpus}, table replicated, no data-path collective',
    }


# ---------------------------------------------------------------------------
# the CUDA path
# ---------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Runs this process (and first-touches its pinned host buffers) on the CPUs of the NUMA node the GPU hangs
    off, so that eight ranks do not pull their H2D / D2H traffic through one socket.  Best effort."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), 'pci_bus_id') else None
        if bdf is None:
            out = subprocess.run(['nvidia-smi', '-i', str(local), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                                 capture_output=True, text=True).stdout.strip()
            bdf = out[-12:] if len(out) >= 12 else None
        if not bdf:
            return ''
        bdf = bdf.lower()
        if len(bdf) == 16:                      # 00000000:17:00.0 -> 0000:17:00.0
            bdf = bdf[4:]
        node = int(open(f'/sys/bus/pci/devices/{bdf}/numa_node').read())
        if node < 0:
            return ''
        cpus = []
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return ''
        os.sched_setaffinity(0, cpus)
        return f'process bound to NUMA node {node} of GPU {local}'
    except Exception:
        return ''


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

    w = synth.workload(args.workload, args.scale, mix=args.mix, spacing=args.spacing, variant=args.variant)
    method = METHODS[args.method or w.method]
    search = SEARCHES[args.search]
    B, scaling = batch_of(args, w, ws)
    K, W = args.steps, args.warmup
    nbuf = 1 if B * w.naxes * 8 >= (1 << 30) else 2

    bind_note = bind_to_gpu_numa_node(local)          # host buffers of the end-to-end leg land next to this GPU
    t_build = time.perf_counter()
    grid = w.grid(dev)
    torch.cuda.empty_cache()                       # the generator's temporaries: the table sizes its cell copy by the free memory
    T = Table(w.axes, grid, len(w.basic_axes), device=local)
    del grid
    torch.cuda.empty_cache()
    if args.gather:
        T.set_option('gather', args.gather)
    if args.staged:
        T.set_option('staged', args.staged)
    for kv in args.opt:
        k, v = kv.split('=')
        T.set_option(k, int(v))
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build

    sweep = [int(float(x)) for x in args.sweep_totals.split(',')] if args.sweep_totals else [None]
    for sweep_total in sweep:
        if sweep_total is not None:                       # one JSON line per total, same table (--sweep-totals)
            args.total, args.batch = sweep_total, None
            B, scaling = batch_of(args, w, ws)
            nbuf = 1 if B * w.naxes * 8 >= (1 << 30) else 2
            K = args.steps if B < (1 << 24) else max(3, min(args.steps, 10))
            T.set_option('stream', 0)
            qs = interps = dists = None
            torch.cuda.empty_cache()
        # per-rank, per-buffer distinct rows of the global synthetic batch
        qs = [gen_queries(w, (b * ws + rank) * B, B, dev) for b in range(nbuf)]
        presort_ms = None
        if args.presort:
            lo = torch.as_tensor([a[0] for a in w.axes], dtype=torch.float64, device=dev)
            step = torch.as_tensor([(a[-1] - a[0]) / (len(a) - 1) for a in w.axes], dtype=torch.float64, device=dev)
            ncell = torch.as_tensor([len(a) - 1 for a in w.axes], dtype=torch.int64, device=dev)
            for b in range(nbuf):
                torch.cuda.synchronize()
                t_sort = time.perf_counter()
                key = torch.zeros(B, dtype=torch.int64, device=dev)
                for a in range(w.naxes):                   # cell number of the (clamped) query, axis 0 most significant
                    c = ((qs[b][:, a] - lo[a]) / step[a]).floor_().clamp_(0, float(ncell[a].item() - 1)).to(torch.int64)
                    key.mul_(int(ncell[a].item())).add_(c)
                    del c
                order = torch.sort(key).indices
                del key
                qs[b] = qs[b][order]
                del order
                torch.cuda.synchronize()
                presort_ms = (time.perf_counter() - t_sort) * 1e3
            torch.cuda.empty_cache()
        interps = torch.empty((B, w.vdim), dtype=torch.float64, device=dev)
        dists = torch.empty((B, 1), dtype=torch.float64, device=dev)
        stream = torch.cuda.Stream(dev)                # kernels and timing events share this stream
        T.set_option('stream', stream.cuda_stream)

        def barrier():
            torch.cuda.synchronize()
            if ws > 1:
                dist.barrier()
            torch.cuda.synchronize()

        use_async = args.api_mode == 'async' or (args.api_mode == 'auto' and B <= (1 << 24))
        torch.cuda.synchronize()
        t_cold = time.perf_counter()
        T.ndpolate(qs[0][:min(B, 1 << 20)], method, search)      # cold first call: lazy nearest-site map / topology builds
        torch.cuda.synchronize()
        cold_s = time.perf_counter() - t_cold
        if use_async:
            T.set_option('async', 1)
            T.ndpolate(qs[0][:min(B, 1 << 20)], method, search)  # builds the kd topology the stream-ordered mode needs
            T.sync()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()                                        # before the warm-up, so that short runs are sampled too
        with torch.cuda.stream(stream):
            for s in range(W):
                T.ndpolate(qs[s % nbuf], method, search, out=(interps, dists))
            T.sync()

        keys = ('launches', 'offgrid', 'hard', 'ties', 'slow', 'fused', 'main_ns', 'search_ns', 'finish_ns')
        acc = dict.fromkeys(keys, 0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for s in range(K):
                T.ndpolate(qs[s % nbuf], method, search, out=(interps, dists))
                if not use_async:
                    for k in keys:
                        acc[k] += T.stat(k)
            if use_async:
                T.sync()                                           # completes the K stream-ordered calls, inside the timed region
            e1.record(stream)
        barrier()
        if use_async:
            # per-step counters and phase times of the same steps, replayed call by call after the timed region
            with torch.cuda.stream(stream):
                per_buf = []
                for b in range(nbuf):
                    T.ndpolate(qs[b], method, search, out=(interps, dists))
                    T.sync()
                    per_buf.append({k: T.stat(k) for k in keys})
            for s in range(K):
                for k in keys:
                    acc[k] += per_buf[s % nbuf][k]
            T.set_option('async', 0)
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())
        if rank == 0 and ms_total < 400.0:
            # a short timed region: keep the GPU busy with the same work until the sampler has seen it under load
            t_keep = time.perf_counter()
            with torch.cuda.stream(stream):
                while time.perf_counter() - t_keep < 0.5:
                    T.ndpolate(qs[0], method, search, out=(interps, dists))
            torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None:
            clocks['note'] = ('sampled every 20 ms from before the warm-up to after the timed steps' +
                              ('; the timed region was shorter than the sampling period, so the same step was repeated '
                               'for 0.5 s after it to record the clocks under this load' if ms_total < 400.0 else ''))
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
        e2e_steps = 0
        while e2e_steps < args.e2e_steps or (time.perf_counter() - t0 < 1.0 and e2e_steps < 50 * args.e2e_steps):
            T.ndpolate(qn, method, search, out=(inn, dn))       # returns when the results are in the host arrays
            e2e_steps += 1
        torch.cuda.synchronize()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if ws > 1:                                               # every rank runs the same number of steps
            steps_t = torch.tensor([e2e_steps], dtype=torch.int64, device=dev)
            dist.all_reduce(steps_t, op=dist.ReduceOp.MIN)
            e2e_steps_min = int(steps_t.item())
        else:
            e2e_steps_min = e2e_steps
        h2d, d2h = T.stat('h2d_bytes'), T.stat('d2h_bytes')
        if ws > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        # per-rank rate (its own steps over its own time), summed over ranks would double-count nothing but hide the
        # slowest: report the conservative whole-job figure = ranks x per-rank queries x common steps / slowest time
        e2e_value = ws * nb_e2e * e2e_steps_min / float(e2e_s.item())
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
            table_bytes = w.vdim * 8
            for ax in w.axes:
                table_bytes *= len(ax)
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
                'ms_per_step': ms_total / K, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None,
                'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args),
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'queries_per_gpu_per_step': nb_e2e, 'steps': e2e_steps_min, 'seconds': float(e2e_s.item()),
                        'host_memory': 'pinned' + (f' ({bind_note})' if bind_note else ''),
                        'matches_device_run': same},
                'gpu_launches': acc['launches'],
                'clocks': clocks,
                'roofline': {
                    'bound': 'hbm', 'kernel': kernel_name,
                    'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': (achieved / peak) if achieved else None,
                    'traffic': traffic,
                'note': ('table larger than L2: every gather is HBM traffic' if table_bytes > 126 * 2**20 else
                         'table fits in the 126 MB L2: gathers are served from L2, so the fraction of the HBM roofline can exceed 1'),
                'traffic_source': ('profiles/r02_traffic.json (ncu --set full on this kernel: same SASS digest; scaled to this batch)'
                                       if traffic else 'no ncu capture of exactly this kernel / this workload'),
                    'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': main_bytes, 'avg_launch_ms': main_s * 1e3,
                    'whole_step': {'achieved': step_gbs, 'frac': step_gbs / peak,
                                   'note': 'queries/step x SURVEY 8d bytes/query (query + full 2^n cell + outputs for EVERY query) over '
                                           'the whole step, all kernels; frac_algorithmic counts what this method really has to move '
                                           "(a 'nearest'-extrapolated query fetches one vertex, not a cell)",
                                   'frac_algorithmic': main_bytes / (ms_total / 1e3 / K) / 1e9 / peak},
                    'phase_ms_per_step': {'k_main': acc['main_ns'] / 1e6 / K, 'k_search': acc['search_ns'] / 1e6 / K,
                                          'k_finish': acc['finish_ns'] / 1e6 / K},
                },
                'mix_measured': {'in_fully_defined_hypercube': n_fdhc / (B * K), 'offgrid': acc['offgrid'] / (B * K),
                                 'needed_lattice_descent': acc['hard'] / (B * K), 'kd_ties': acc['ties'],
                                 'left_to_exact_generic_kernel': acc['slow'] / (B * K)},
                'path': 'fused (k_fused + k_slow)' if fused else 'three-pass (gather, search, finish)',
                'api_mode': ('async: K stream-ordered ndpolate calls on device tensors + one ndpb_table_sync, all inside the timed '
                             'region; counters / phase times from a call-by-call replay after it' if use_async else
                             'sync: every ndpolate call returns with its results complete'),
                'table': {'cellmajor': bool(T.stat('cellmajor')), 'build_s': build_s, 'cold_first_call_s': cold_s,
                          'note': 'build_s: upload + masks + pyramids + cell-major copy; cold_first_call_s: first ndpolate of 2^20 '
                                  'queries incl. the lazy nearest-site map (the reference builds its kd-tree there)'},
                'checksum': checksum,
            }
            if presort_ms is not None:
                line['presort'] = {'presort_ms_per_batch': presort_ms,
                                   'note': 'queries ordered by cell before the timed region (torch: cell keys, sort, row gather); '
                                           'value / ms_per_step do NOT include it'}
            if ws == 1 and not args.no_cpu_baseline and sweep_total is None:
                try:
                    cb = cpu_reference_run(args, steps=2, warmup=1)
                    line['cpu_baseline'] = {k: cb[k] for k in CPU_KEYS}
                except Exception as e:       # the baseline is reported beside the number, never instead of it
                    line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                                            'sample': f'failed: {e!r}'}
            print(json.dumps(line))
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


def interleave_host_memory():
    """set_mempolicy(MPOL_INTERLEAVE, all nodes) for this thread: the pinned batch of an in-process multi-GPU
    call is then spread over the NUMA nodes instead of sitting next to one socket.  Best effort (x86-64 syscall)."""
    try:
        import ctypes
        nodes = [int(d[4:]) for d in os.listdir('/sys/devices/system/node') if d.startswith('node') and d[4:].isdigit()]
        if len(nodes) < 2:
            return f'{len(nodes)} NUMA node'
        mask = ctypes.c_ulong(sum(1 << n for n in nodes))
        libc = ctypes.CDLL(None, use_errno=True)
        rc = libc.syscall(238, 3, ctypes.byref(mask), ctypes.c_ulong(max(nodes) + 2))      # SYS_set_mempolicy, MPOL_INTERLEAVE
        return f'interleaved over {len(nodes)} NUMA nodes' if rc == 0 else f'set_mempolicy failed ({ctypes.get_errno()})'
    except Exception as e:
        return f'not interleaved ({e!r})'


def run_inprocess(args):
    """One process, one ndpolate call per step over N GPUs (ndpb_group / cndpolator.Group): host arrays in, host
    arrays out, every GPU copying its slice back itself -- the API-level multi-GPU number."""
    import numpy as np
    import torch
    from ndpolator_b200 import synth
    from ndpolator_b200.cndpolator import Group, Table
    n_dev = args.gpus
    w = synth.workload(args.workload, args.scale, mix=args.mix, spacing=args.spacing, variant=args.variant)
    method = METHODS[args.method or w.method]
    search = SEARCHES[args.search]
    policy = interleave_host_memory()
    per_gpu = args.e2e_batch
    total = per_gpu * n_dev
    grid = w.grid()                                            # host grid: every member uploads its own replica
    t0 = time.perf_counter()
    G = Group(w.axes, grid, len(w.basic_axes), devices=list(range(n_dev)))
    build_s = time.perf_counter() - t0
    qh = torch.empty((total, w.naxes), dtype=torch.float64).pin_memory()
    for lo in range(0, total, 1 << 23):
        c = min(1 << 23, total - lo)
        qh[lo:lo + c] = w.queries(lo, c, 'cuda:0').cpu()
    ih = torch.empty((total, w.vdim), dtype=torch.float64).pin_memory()
    dh = torch.empty((total, 1), dtype=torch.float64).pin_memory()
    qn, inn, dn = qh.numpy(), ih.numpy(), dh.numpy()
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(max(1, args.warmup)):
        G.ndpolate(qn, method, search, out=(inn, dn))
    t0 = time.perf_counter()
    steps = 0
    while steps < args.steps or time.perf_counter() - t0 < 1.0:
        G.ndpolate(qn, method, search, out=(inn, dn))
        steps += 1
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    # same rows through ONE table: identical bits
    T = G.member(0)
    sub = min(total, 1 << 22)
    ri, rd = T.ndpolate(qn[:sub], method, search)
    same = bool(np.array_equal(ri.view(np.uint64), inn[:sub].view(np.uint64)) and np.array_equal(rd.view(np.uint64), dn[:sub].view(np.uint64)))
    if not same:
        raise SystemExit('bench.py --inprocess: group results differ from a single table')
    value = total * steps / dt
    cfg = workload_config(args)
    cfg['queries_per_step_all_gpus'] = total
    cfg['queries_per_gpu_per_step'] = per_gpu
    cfg['parallelism'] = f'ONE process, one ndpolate call per step split over {n_dev} GPUs (ndpb_group): contiguous query ranges, table replicated, every GPU copies its slice of the results back; no collective'
    print(json.dumps({
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_dev, 'steps': steps, 'warmup': max(1, args.warmup),
        'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': cfg, 'mode': 'inprocess-group (value IS the end-to-end number: host arrays in and out)',
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': G.stat('h2d_bytes'), 'd2h_bytes_per_step': G.stat('d2h_bytes'),
                'queries_per_gpu_per_step': per_gpu, 'steps': steps, 'seconds': dt, 'host_memory': f'pinned, {policy}',
                'matches_single_table': same},
        'gpu_launches': G.stat('launches') * steps, 'clocks': clocks,
        'table': {'build_s': build_s, 'note': 'N replicas built concurrently, one host thread per device'},
    }))


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
    elif args.inprocess:
        run_inprocess(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
