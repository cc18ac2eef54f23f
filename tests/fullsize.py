"""Loader / checker for the full-size golden vectors (tests/golden/full_*.npz, produced by
tests/golden/make_golden_fullsize.py from the UNMODIFIED reference at BASELINE.json's full
grid sizes).  Grid and queries are regenerated from the seeded synthetic stream; only the
reference's outputs are stored."""
import glob
import os

import numpy as np

from cases import assert_same_bits

FULL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'full_*.npz')))


def load_full(path):
    from ndpolator_b200 import synth
    z = np.load(path)
    w = synth.workload(str(z['workload']), None, mix=str(z['mix']))
    return z, w, int(z['first']), int(z['count'])


def check_full(z, ndpolate, find_nearest, methods=(0, 1, 2)):
    """ndpolate(method) -> (interps (n, vdim), dists (n, 1)) and find_nearest(method) -> (coords, dist),
    both as numpy arrays for the golden query range, kd search."""
    n = int(z['count'])
    none_nan = np.unpackbits(z['none_nan'])[:n].astype(bool)
    want2 = z['interps_m2']
    if 0 in methods:
        i0, d0 = ndpolate(0)
        assert np.array_equal(np.isnan(i0[:, 0]), none_nan), 'NaN placement with extrapolation none'
        assert_same_bits(i0[~none_nan], want2[~none_nan], 'in-grid interps')
        assert not d0.any()
    if 2 in methods:
        i2, d2 = ndpolate(2)
        assert_same_bits(i2, want2, 'interps linear')
        assert_same_bits(d2.reshape(-1, 1), z['dists_m2'], 'dists linear')
    if 1 in methods:
        i1, d1 = ndpolate(1)
        assert_same_bits(i1[~none_nan], want2[~none_nan], 'in-grid interps (nearest)')
        assert_same_bits(i1[none_nan], z['interps_m1_off'], 'interps nearest')
        assert_same_bits(d1.reshape(-1, 1), z['dists_m1'], 'dists nearest')
    if find_nearest is not None:
        for m in (1, 2):
            if m in methods:
                c, d = find_nearest(m)
                assert np.array_equal(np.asarray(c, dtype=np.int16), z[f'coords_m{m}']), f'nearest coords, method {m}'
                assert_same_bits(np.asarray(d).reshape(-1), z[f'ndist_m{m}'], f'nearest dist, method {m}')
