#!/usr/bin/env python
"""Fuzz campaign (CPU only): the host build of the device logic (tests/hostsim: import, plan, nearest-site map,
exact lattice search, blocked layout) against the oracle port -- and the live reference where it is built -- on random
small tables: random axis counts / lengths / spacings, structured + i.i.d. + slab voids, associated axes, vector
values, adversarial and far-out queries, every method and search.  Development tool; prints the first mismatch.

    python tools/fuzz_hostsim.py [--minutes 20] [--seed 1]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from cases import adversarial_queries  # noqa: E402
from hostsim.driver import HostSim  # noqa: E402
from oracle.oracle import Oracle, have  # noqa: E402


MAXDIM, MAXLEN = 5, 9


def random_table(rng):
    n = int(rng.integers(1, MAXDIM + 1))
    nbasic = int(rng.integers(1, n + 1))
    lens = [int(rng.integers(2, MAXLEN + 1 if n <= 3 else max(4, MAXLEN - 3 - (n > 4)))) for _ in range(n)]
    axes = []
    for L in lens:
        kind = rng.integers(0, 4)
        if kind == 0:
            ax = np.arange(L, dtype=np.float64)                                   # unit spacing (power of two)
        elif kind == 1:
            ax = rng.uniform(-5, 5) + rng.choice([0.25, 0.5, 2.0, 0.3, 1e-3, 17.0]) * np.arange(L)
        elif kind == 2:
            ax = np.sort(rng.uniform(-3, 3, L)); ax += np.arange(L) * 1e-3        # irregular, strictly increasing
        else:
            ax = np.geomspace(rng.uniform(0.5, 5), rng.uniform(50, 500), L)
        axes.append(np.ascontiguousarray(ax))
    vdim = int(rng.choice([1, 1, 1, 2, 5]))
    grid = rng.random(tuple(lens) + (vdim,))
    bshape = tuple(lens[:nbasic])
    style = rng.integers(0, 5)
    void = np.zeros(bshape, bool)
    if style == 1:
        void = rng.random(bshape) < rng.choice([0.05, 0.2, 0.5, 0.8])
    elif style == 2 and nbasic >= 2:                                               # structured corner on two axes
        i = np.indices(bshape)
        a, b = rng.choice(nbasic, 2, replace=False)
        void = (i[a] / max(1, lens[a] - 1) + i[b] / max(1, lens[b] - 1)) > rng.uniform(0.8, 1.6)
    elif style == 3:                                                               # slab on one axis
        i = np.indices(bshape)
        a = int(rng.integers(0, nbasic))
        void = i[a] >= rng.integers(1, lens[a] + 1)
    elif style == 4:
        void = rng.random(bshape) < 0.97                                           # almost empty
    grid[void] = np.nan
    return axes, nbasic, grid


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--minutes', type=float, default=20)
    ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--maxdim', type=int, default=5)
    ap.add_argument('--maxlen', type=int, default=9)
    args = ap.parse_args()
    global MAXDIM, MAXLEN
    MAXDIM, MAXLEN = args.maxdim, args.maxlen
    rng = np.random.default_rng(args.seed)
    use_ref = have('reference')
    t_end = time.time() + args.minutes * 60
    ntab = ncmp = 0
    while time.time() < t_end:
        axes, nbasic, grid = random_table(rng)
        q = adversarial_queries(axes, 600, rng)
        far = q.copy()
        k = rng.integers(0, len(axes), len(far))
        span = np.array([a[-1] - a[0] for a in axes])
        far[np.arange(len(far)), k] += rng.choice([-1.0, 1.0], len(far)) * span[k] * rng.uniform(0.5, 40, len(far))
        # a family a hair away from lattice points and cell centres (bisectors of the two metrics), on every axis
        bis = q[:200].copy()
        for a_i, ax in enumerate(axes):
            pick = rng.random(200) < 0.5
            j = rng.integers(0, len(ax) - 1, 200)
            frac = rng.choice([0.0, 0.5, 1.0], 200) + rng.choice([-1.0, 1.0], 200) * rng.choice([0.0, 1e-16, 1e-12, 1e-9, 3e-7], 200)
            bis[:, a_i] = np.where(pick, ax[j] + frac * (ax[j + 1] - ax[j]), bis[:, a_i])
        q = np.concatenate([q, far[:200], bis])
        P = Oracle('port', axes, grid, nbasic)
        H = HostSim(axes, grid, nbasic)
        vm, hm = P.masks()
        R = Oracle('reference', axes, grid, nbasic) if use_ref and vm.any() and hm.any() else None   # the reference crashes on an empty kd-tree
        ntab += 1
        for a, b, what in zip(P.masks(), H.masks(), ('vmask', 'hcmask')):
            if not np.array_equal(a, b):
                print('MISMATCH', what, 'seed', args.seed, 'table', ntab); return 1
        if ntab % 4 == 0 and vm.any() and hm.any():        # the taps: nearest coordinates, rebuilt kd topology
            fo = P.find(q)
            for m in (1, 2):
                for s in (0, 1):
                    a, b = P.find_nearest(*fo, m, s), H.find_nearest(q, m, s)
                    if not (np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))):
                        print(f'MISMATCH find_nearest method {m} search {s}: seed {args.seed} table {ntab}')
                        np.savez(os.path.join(ROOT, 'gpurun_out', f'fuzz_fail_{args.seed}_{ntab}.npz'), q=q, grid=grid, nbasic=nbasic,
                                 **{f'ax{i}': a for i, a in enumerate(axes)})
                        return 1
                pp, dp = P.tree_topology(m)
                ph, dh = H.topology(m)
                if not (np.array_equal(pp, ph.astype(np.int64)) and np.array_equal(dp, dh)):
                    print(f'MISMATCH topology method {m}: seed {args.seed} table {ntab}')
                    return 1
        for m in (0, 1, 2):
            for s in (0, 1):
                want = P.ndpolate(q, m, s)
                engines = [('3pass', H.ndpolate(q, m, s, 0)), ('fused', H.ndpolate_fused(q, m, s))]
                if R is not None:
                    engines.append(('reference', R.ndpolate(q, m, s)))
                for name, got in engines:
                    if got is None:
                        continue
                    ncmp += 1
                    for x, y, what in ((want[0], got[0], 'interps'), (want[1], got[1], 'dists')):
                        if m == 0 and what == 'dists':
                            continue
                        if not np.array_equal(np.asarray(x).view(np.uint64), np.asarray(y).reshape(np.asarray(x).shape).view(np.uint64)):
                            bad = np.nonzero(np.asarray(x).view(np.uint64) != np.asarray(y).reshape(np.asarray(x).shape).view(np.uint64))[0]
                            print(f'MISMATCH {name} {what} method {m} search {s}: seed {args.seed} table {ntab} '
                                  f'(axes {[len(a) for a in axes]}, nbasic {nbasic}, vdim {grid.shape[-1]}), {len(bad)} rows, first {bad[0]}: q = {q[bad[0]]!r}')
                            np.savez(os.path.join(ROOT, 'gpurun_out', f'fuzz_fail_{args.seed}_{ntab}.npz'), q=q, grid=grid, nbasic=nbasic,
                                     **{f'ax{i}': a for i, a in enumerate(axes)})
                            return 1
        del H, P, R
    print(f'fuzz ok: seed {args.seed}, {ntab} tables, {ncmp} engine comparisons, {args.minutes} min')
    return 0


if __name__ == '__main__':
    sys.exit(main())
