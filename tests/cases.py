"""Shared parity cases: small tables that exercise every branch of the hot path
(voids, associated axes, vdim > 1, non-uniform ticks, >3-D, the runtime-n path)
and adversarial query sets (exact ticks, tick +- 5e-7, cell mid-points => exact
distance ties, half a cell outside either end), plus an Ndpolator-shaped adapter
over the CPU oracle so that known-answer tests run unchanged on every engine."""
from __future__ import annotations

import numpy as np

from oracle.oracle import Oracle


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64) if a.dtype == np.float64 else a


def same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def assert_same_bits(a, b, what=''):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f'{what}: shape {a.shape} != {b.shape}'
    if not np.array_equal(bits(a), bits(b)):
        bad = np.argwhere(bits(a) != bits(b))
        i = tuple(bad[0])
        extra = ''
        if a.dtype == np.float64:
            ours = a[tuple(bad.T)]
            extra = (f'; of the differing elements {int(np.isnan(ours).sum())} are NaN on the left, '
                     f'{int(np.isnan(b[tuple(bad.T)]).sum())} on the right; rows affected: {len(set(bad[:, 0].tolist()))}')
        raise AssertionError(f'{what}: {len(bad)} of {a.size} elements differ; first at {i}: {a[i]!r} vs {b[i]!r}{extra}')


def adversarial_queries(axes, n_rand, rng):
    lo = np.array([a[0] for a in axes])
    hi = np.array([a[-1] for a in axes])
    out = [lo - 0.3 * (hi - lo) + 1.6 * (hi - lo) * rng.random((n_rand, len(axes)))]
    special = []
    for a in axes:
        d = np.diff(a)
        s = list(a) + list(a + 5e-7) + list(a - 5e-7) + list((a[1:] + a[:-1]) / 2)
        s += list(a[:-1] + 5e-7 * d) + list(a[1:] - 5e-7 * d)
        s += [a[0] - 0.5 * d[0], a[-1] + 0.5 * d[-1], a[0] - 1.5 * d[0], a[-1] + 2.0 * d[-1]]
        special.append(np.array(s))
    m = max(64, n_rand // 4)
    out.append(np.stack([rng.choice(s, m) for s in special], axis=1))
    r = lo - 0.1 * (hi - lo) + 1.2 * (hi - lo) * rng.random((m, len(axes)))
    sp = np.stack([rng.choice(s, m) for s in special], axis=1)
    out.append(np.where(rng.random((m, len(axes))) < 0.5, sp, r))
    return np.ascontiguousarray(np.concatenate(out))


def build_cases(seed=2):
    """-> list of (name, axes tuple, nbasic, grid)"""
    rng = np.random.default_rng(seed)
    cases = []
    ax = (np.linspace(1, 5, 5), np.linspace(10, 50, 5), np.linspace(100, 500, 5))
    g = (ax[0][:, None, None] + ax[1][None, :, None] + ax[2][None, None, :])[..., None]
    cases.append(('readme', ax, 3, g))
    ax = (np.linspace(0, 5, 6), np.linspace(-1, 1, 5), np.linspace(0, 12, 7))
    g = rng.random((6, 5, 7, 1)); g[rng.random((6, 5, 7)) < 0.2] = np.nan
    cases.append(('holed3d', ax, 3, g))
    ax = (np.array([0, 0.1, 0.5, 2, 3, 7.]), np.array([1., 2, 4, 8, 16]))
    g = rng.random((6, 5, 3)); g[rng.random((6, 5)) < 0.3] = np.nan
    cases.append(('nonuniform2d_v3', ax, 2, g))
    ax = (np.linspace(0, 4, 5), np.linspace(0, 3, 4), np.linspace(0, 1, 6))
    g = rng.random((5, 4, 6, 2)); g[rng.random((5, 4)) < 0.2] = np.nan
    cases.append(('2basic_1assoc_v2', ax, 2, g))
    ax = (np.linspace(0, 9, 10), np.linspace(0, 3, 4), np.linspace(0, 2, 3))
    g = rng.random((10, 4, 3, 2)); g[rng.random((10,)) < 0.2] = np.nan
    cases.append(('1basic_2assoc', ax, 1, g))
    ax = tuple(np.linspace(0, 3, 4) for _ in range(4))
    g = rng.random((4, 4, 4, 4, 1)); g[rng.random((4, 4, 4, 4)) < 0.06] = np.nan
    cases.append(('4d', ax, 4, g))
    ax = (np.linspace(0, 1e-4, 11), np.linspace(0, 1, 5))
    g = rng.random((11, 5, 1))
    cases.append(('finetick', ax, 2, g))
    ax = tuple(np.linspace(0, 5, 6) for _ in range(5))
    g = rng.random((6,) * 5 + (1,))
    i0, i1 = np.meshgrid(np.arange(6), np.arange(6), indexing='ij')
    g[(i0 + i1) / 5 > 1.3] = np.nan
    cases.append(('5d_corner_void', ax, 5, g))
    ax = tuple(np.linspace(0, 2, 3) for _ in range(7))
    g = rng.random((3,) * 7 + (2,)); g[rng.random((3,) * 7) < 0.01] = np.nan
    cases.append(('7d_runtime_n', ax, 7, g))
    ax = (np.linspace(0, 12, 13), np.linspace(0, 10, 11))
    g = rng.random((13, 11, 1)); g[rng.random((13, 11)) < 0.45] = np.nan
    cases.append(('2d_sparse', ax, 2, g))
    ax = (np.geomspace(3500, 50000, 12), np.linspace(0, 5, 6), np.linspace(-2.5, 0.5, 4), np.linspace(0.02, 1, 7))
    g = rng.random((12, 6, 4, 7, 16))
    i0, i1 = np.meshgrid(np.arange(12), np.arange(6), indexing='ij')
    g[(i0 / 11 + (1 - i1 / 5)) > 1.2254] = np.nan
    cases.append(('c3_like_v16', ax, 3, g))
    # scalar tables with associated axes: voids are whole basic vertices (the fused CUDA path applies) ...
    ax = (np.linspace(0, 7, 8), np.linspace(0, 5, 6), np.linspace(0, 1, 5), np.linspace(-1, 1, 4))
    g = rng.random((8, 6, 5, 4, 1)); g[rng.random((8, 6)) < 0.25] = np.nan
    cases.append(('2basic_2assoc_v1', ax, 2, g))
    # ... and a grid whose NaNs depend on the associated index too (ndp_types.c:179-184 looks at associated
    # index 0 only, ndpolator.c:491 at the corners actually used): the mask bit does not decide "fully defined"
    ax = (np.linspace(0, 6, 7), np.linspace(0, 4, 5), np.linspace(0, 1, 4))
    g = rng.random((7, 5, 4, 1)); g[rng.random((7, 5)) < 0.2] = np.nan; g[rng.random((7, 5, 4)) < 0.05] = np.nan
    cases.append(('assoc_dependent_nan_v1', ax, 2, g))
    return cases


class OracleNdpolator:
    """The reference's Python surface (ndpolator/ndpolator.py) over a CPU oracle."""

    def __init__(self, basic_axes, kind='port'):
        self.axes = basic_axes
        self.kind = kind
        self.table = {}

    def register(self, name, associated_axes, grid):
        self.table[name] = {'associated_axes': associated_axes, 'grid': grid, 'capsule': None}

    def _o(self, name):
        e = self.table[name]
        if e['capsule'] is None:
            axes = self.axes if e['associated_axes'] is None else self.axes + e['associated_axes']
            e['capsule'] = Oracle(self.kind, axes, e['grid'], len(self.axes))
        return e['capsule']

    def import_query_pts(self, name, query_pts):
        return self._o(name).find(query_pts)

    def find_hypercubes(self, name, normed_query_pts, indices, flags, associated_axes=None):
        o = self._o(name)
        dim, fdhc, v = o.hypercubes(indices, flags, normed_query_pts)
        return tuple(v[i, :(1 << int(d)) * o.vdim].reshape((2,) * int(d) + (o.vdim,)) for i, d in enumerate(dim))

    def distance(self, name, query_pts):
        return self._o(name).distance(query_pts)

    def ndpolate(self, name, query_pts, extrapolation_method='none', search_algorithm='kdtree'):
        methods = {'none': 0, 'nearest': 1, 'linear': 2}
        searches = {'kdtree': 0, 'linear': 1}
        if extrapolation_method not in methods:
            raise ValueError('bad extrapolation_method')
        if search_algorithm not in searches:
            raise ValueError('bad search_algorithm')
        interps, dists = self._o(name).ndpolate(query_pts, methods[extrapolation_method], searches[search_algorithm])
        return {'interps': interps} if extrapolation_method == 'none' else {'interps': interps, 'dists': dists}


def bisector_queries(w, n, seed=77):
    """C4-like workload `w` (mask depends on axes 0 and 1): queries a hair (2^-18 .. 2^-44, and exactly 0) away from
    the bisector between two lattice points on the mask-invariant axes 2..4, half of them far outside on axis 0."""
    rng = np.random.default_rng(seed)
    q = w.queries(0, n).copy()
    top = float(w.axes[0][-1])
    q[:, 0] = np.where(rng.random(n) < 0.5, top + rng.random(n) * 20.0, q[:, 0])
    eps = np.array([0.0] + [2.0 ** -k for k in (18, 20, 21, 24, 28, 31, 32, 33, 36, 40, 44)])
    for a in (2, 3, 4):
        pick = rng.random(n) < 0.6
        base = rng.integers(0, int(top), n) + np.where(rng.random(n) < 0.5, 0.5, 0.0)   # vertex / cell-centre bisectors
        q[:, a] = np.where(pick, base + rng.choice([-1.0, 1.0], n) * rng.choice(eps, n), q[:, a])
    return q
