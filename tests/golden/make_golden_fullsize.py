"""Full-size golden vectors: the UNMODIFIED reference C core (oracle/_ref/libndp_ref.so) run
on BASELINE.json's configurations AT THEIR FULL GRID SIZE -- C4 40^5 (its insertion-built
kd-trees take ~15 min and ~8 GB each), C5 16^6, C3 80x20x10x50 vdim 16 -- on seeded query
sub-ranges of the very batches bench.py times (ndpolator_b200/synth.py), kd search,
extrapolation none / nearest / linear.

    make -C oracle && python tests/golden/make_golden_fullsize.py [C4 C5 C3]

Only the query spec (workload, mix, first row, count) and the reference's OUTPUTS are stored
(tests/golden/full_<tag>.npz, a few MB each); the grid and the queries are regenerated from
the seeded stream at test time.  Nothing reads /root/reference at test time."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from ndpolator_b200 import synth  # noqa: E402
from oracle.oracle import Oracle, have  # noqa: E402

# (tag, workload, mix, first row, count, methods stored in full)
SPECS = {
    'C4': [('c4_near', 'C4', 'near', 7_000_000, 100_000), ('c4_deep', 'C4', 'deep', 3_000_000, 50_000)],
    'C5': [('c5', 'C5', 'near', 500_000, 100_000)],
    'C3': [('c3', 'C3', 'near', 2_000_000, 20_000)],
}


def main(which):
    if not have('reference'):
        raise SystemExit('oracle/_ref/libndp_ref.so missing: run `make -C oracle` where /root/reference is mounted')
    threads = os.cpu_count() or 1
    for key in which:
        specs = SPECS[key]
        w0 = synth.workload(specs[0][1], None)
        t0 = time.time()
        grid = w0.grid()
        nb = len(w0.basic_axes)
        R = Oracle('reference', w0.axes, grid, nb)
        print(f'{key}: table {grid.shape} registered in {time.time() - t0:.0f} s', flush=True)
        for m in (1, 2):
            t0 = time.time()
            R.warm(m)
            print(f'{key}: reference kd-tree for method {m} built in {time.time() - t0:.0f} s', flush=True)
        for tag, name, mix, first, count in specs:
            w = synth.workload(name, None, mix=mix)
            q = np.ascontiguousarray(w.queries(first, count))
            out = {'workload': name, 'mix': mix, 'first': np.int64(first), 'count': np.int64(count),
                   'vdim': np.int32(w.vdim)}
            idx, flg, nrm = R.find(q)
            i0, _ = R.ndpolate(q, 0, 0, threads=threads)
            out['none_nan'] = np.packbits(np.isnan(i0[:, 0]))
            for m in (1, 2):
                t0 = time.time()
                interps, dists = R.ndpolate(q, m, 0, threads=threads)
                # in-grid rows do not depend on the method: check, then store them once (method 2 in full)
                ing = ~np.isnan(i0[:, 0])
                assert np.array_equal(interps[ing].view(np.uint64), i0[ing].view(np.uint64))
                if m == 2:
                    out['interps_m2'] = interps
                else:
                    out['interps_m1_off'] = interps[~ing]          # nearest: only the extrapolated rows differ
                out[f'dists_m{m}'] = dists
                cuts = np.linspace(0, count, threads + 1).astype(np.int64)
                from concurrent.futures import ThreadPoolExecutor
                coords = np.empty((count, w.naxes), np.int32)
                ndist = np.empty(count, np.float64)

                def run(k):
                    lo, hi = int(cuts[k]), int(cuts[k + 1])
                    c, d = R.find_nearest(idx[lo:hi], flg[lo:hi], nrm[lo:hi], m, 0)
                    coords[lo:hi], ndist[lo:hi] = c, d

                with ThreadPoolExecutor(threads) as ex:
                    list(ex.map(run, range(threads)))
                out[f'coords_m{m}'] = coords.astype(np.int16)
                out[f'ndist_m{m}'] = ndist
                print(f'{tag}: method {m}: {count} queries in {time.time() - t0:.0f} s '
                      f'({(~ing).mean() * 100:.1f} % off-grid)', flush=True)
            path = os.path.join(HERE, f'full_{tag}.npz')
            np.savez_compressed(path, **out)
            print(f'{tag}: -> {os.path.getsize(path) / 1e6:.2f} MB', flush=True)
        R.close()


if __name__ == '__main__':
    main(sys.argv[1:] or ['C3', 'C5', 'C4'])
