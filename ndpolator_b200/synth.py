"""Deterministic synthetic tables and query batches for the configurations of
BASELINE.json: C1..C5 and scaled-down variants for tests.

Everything comes from one counter-based stream

    u(seed, k) = (splitmix64(seed * 0x9E3779B97F4A7C15 + k) >> 11) * 2**-53

so that any sub-range of any array can be generated independently and
identically with numpy (host, oracle side) and torch (device, bench side).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_GOLD = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_MASK = (1 << 64) - 1


def _s64(x: int) -> int:
    """Python int (mod 2**64) -> the same bit pattern as a signed 64-bit value."""
    x &= _MASK
    return x - (1 << 64) if x >= (1 << 63) else x


def u01_np(seed: int, first: int, count: int) -> np.ndarray:
    """u(seed, k) for k in [first, first+count), float64 in [0, 1)."""
    with np.errstate(over='ignore'):
        k = np.arange(first, first + count, dtype=np.uint64)
        z = k + np.uint64((seed * _GOLD) & _MASK) + np.uint64(_GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def u01_torch(seed: int, first: int, count: int, device):
    """Same stream on a torch device (int64 arithmetic wraps like uint64; logical
    shifts are emulated by masking the sign extension)."""
    import torch
    k = torch.arange(first, first + count, dtype=torch.int64, device=device)
    z = k + _s64(seed * _GOLD) + _s64(_GOLD)

    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)

    z = (z ^ lsr(z, 30)) * _s64(_M1)
    z = (z ^ lsr(z, 27)) * _s64(_M2)
    z = z ^ lsr(z, 31)
    return lsr(z, 11).to(torch.float64) * (1.0 / (1 << 53))


def _u01(seed, first, count, device=None):
    return u01_np(seed, first, count) if device is None else u01_torch(seed, first, count, device)


@dataclass
class Workload:
    name: str
    basic_axes: tuple
    associated_axes: tuple | None
    vdim: int
    method: str
    n_queries: int
    grid_seed: int
    query_seed: int
    mix: str = 'near'
    extra: dict = field(default_factory=dict)

    @property
    def axes(self):
        return self.basic_axes + (self.associated_axes or ())

    @property
    def naxes(self):
        return len(self.axes)

    @property
    def shape(self):
        return tuple(len(a) for a in self.axes) + (self.vdim,)

    @property
    def bytes_per_query(self):
        """Algorithmic bytes per query: query + 2^n corners + interps + dist."""
        return self.naxes * 8 + (1 << self.naxes) * self.vdim * 8 + self.vdim * 8 + 8

    # -- table -----------------------------------------------------------------
    def grid(self, device=None):
        """The (len..., vdim) float64 grid, NaN on void vertices (component 0..vdim-1 all NaN)."""
        builder = _GRIDS[self.name.split('@')[0].split('/')[0]]
        return builder(self, device)

    # -- queries ---------------------------------------------------------------
    def queries(self, first: int, count: int, device=None):
        """Rows [first, first+count) of the (n_queries, naxes) batch."""
        gen = _QUERIES[self.name.split('@')[0].split('/')[0]]
        return gen(self, first, count, device)


def _xp(device):
    if device is None:
        return np
    import torch
    return torch


def _as(device, a):
    if device is None:
        return np.asarray(a, dtype=np.float64)
    import torch
    return torch.as_tensor(np.asarray(a, dtype=np.float64), device=device)


def _lattice_index(w: Workload, nb: int, device):
    """Per-basic-axis index arrays broadcastable over the basic lattice."""
    xp = _xp(device)
    out = []
    for a in range(nb):
        n = len(w.axes[a])
        shp = [1] * nb
        shp[a] = n
        if device is None:
            out.append(np.arange(n).reshape(shp))
        else:
            out.append(xp.arange(n, device=device).reshape(shp))
    return out


def _apply_void(grid, void, nb):
    """void: boolean over the basic lattice -> NaN in every component/associated slot."""
    extra = grid.ndim - nb
    v = void.reshape(tuple(void.shape) + (1,) * extra)
    if isinstance(grid, np.ndarray):
        grid[np.broadcast_to(v, grid.shape)] = np.nan
    else:
        import torch
        grid.masked_fill_(v, float('nan'))
    return grid


def _grid_sum(w: Workload, device):
    """grid[i,j,k,0] = a1[i] + a2[j] + a3[k] (+...), the README / benchmark field."""
    nb = len(w.basic_axes)
    axes = [_as(device, a) for a in w.axes]
    g = None
    for a, ax in enumerate(axes):
        shp = [1] * w.naxes
        shp[a] = len(ax)
        term = ax.reshape(shp)
        g = term if g is None else g + term
    g = g.reshape(tuple(g.shape) + (1,))
    if w.vdim > 1:
        g = g.repeat(*([1] * w.naxes + [w.vdim])) if device is not None else np.repeat(g, w.vdim, axis=-1)
    g = g.clone() if device is not None else np.array(g, dtype=np.float64)
    frac = w.extra.get('iid_void')
    if frac:
        nverts = int(np.prod([len(a) for a in w.basic_axes]))
        void = (_u01(w.extra['void_seed'], 0, nverts, device) < frac).reshape([len(a) for a in w.basic_axes])
        _apply_void(g, void, nb)
    return g


def _grid_random(w: Workload, device):
    """grid = u(grid_seed, flat element); structured corner void or i.i.d. void."""
    nb = len(w.basic_axes)
    total = int(np.prod(w.shape))
    g = _u01(w.grid_seed, 0, total, device).reshape(w.shape)
    cut = w.extra.get('corner_cut')
    if cut is not None:
        i = _lattice_index(w, nb, device)
        l0, l1 = len(w.axes[0]) - 1, len(w.axes[1]) - 1
        if w.extra.get('invert_axis1'):
            void = (i[0] / l0 + (1.0 - i[1] / l1)) > cut
        elif l0 == l1:
            void = (i[0] + i[1]) / float(l0) > cut
        else:
            void = (i[0] / l0 + i[1] / l1) > cut
        shape = [len(a) for a in w.basic_axes]
        void = void.expand(*shape) if device is not None else np.broadcast_to(void, shape)
        _apply_void(g, void, nb)
    frac = w.extra.get('iid_void')
    if frac:
        nverts = int(np.prod([len(a) for a in w.basic_axes]))
        void = (_u01(w.extra['void_seed'], 0, nverts, device) < frac).reshape([len(a) for a in w.basic_axes])
        _apply_void(g, void, nb)
    return g


def _span(w: Workload, device):
    lo = _as(device, [a[0] for a in w.axes])
    hi = _as(device, [a[-1] for a in w.axes])
    return lo, hi


def _q_box(w: Workload, first, count, device):
    """Each axis uniform in [min - f*range, max + f*range] (C2: f = 0.3), or in explicit
    per-axis [lo, hi] (C1: README ranges)."""
    n = w.naxes
    u = _u01(w.query_seed, first * n, count * n, device).reshape(count, n)
    if 'box' in w.extra:
        lo = _as(device, [b[0] for b in w.extra['box']])
        hi = _as(device, [b[1] for b in w.extra['box']])
    else:
        f = w.extra.get('margin', 0.3)
        a0, a1 = _span(w, device)
        lo, hi = a0 - f * (a1 - a0), a1 + f * (a1 - a0)
    return lo + (hi - lo) * u


def _q_corner(w: Workload, first, count, device):
    """C3/C4: u_a ~ U(-0.05, 1.05) of each basic axis' range; 'near' mix reflects points
    that fall deep inside the void corner back through the centre; associated axes are
    uniform inside their range."""
    xp = _xp(device)
    n, nb = w.naxes, len(w.basic_axes)
    u = _u01(w.query_seed, first * n, count * n, device).reshape(count, n)
    m = w.extra.get('margin', 0.05)
    ub = -m + (1.0 + 2.0 * m) * u[:, :nb]
    cut = w.extra.get('corner_cut')
    if cut is not None and w.mix == 'near':
        u0 = ub[:, 0]
        u1 = (1.0 - ub[:, 1]) if w.extra.get('invert_axis1') else ub[:, 1]
        deep = (u0 + u1) > (cut + 0.10)
        r0 = xp.where(deep, 1.0 - u0, u0)
        r1 = xp.where(deep, 1.0 - u1, u1)
        if w.extra.get('invert_axis1'):
            r1 = 1.0 - r1
        if device is None:
            ub = ub.copy()
        else:
            ub = ub.clone()
        ub[:, 0] = r0
        ub[:, 1] = r1
    lo, hi = _span(w, device)
    q = xp.empty((count, n), dtype=xp.float64) if device is None else xp.empty((count, n), dtype=xp.float64, device=device)
    q[:, :nb] = lo[:nb] + (hi[:nb] - lo[:nb]) * ub
    if n > nb:
        q[:, nb:] = lo[nb:] + (hi[nb:] - lo[nb:]) * u[:, nb:]
    return q


def _q_mixed(w: Workload, first, count, device):
    """C5: one third in-grid (void-corner points reflected), two thirds with >= 1 randomly
    chosen axis pushed outside the box by U(0, 0.3 * range)."""
    xp = _xp(device)
    n = w.naxes
    u = _u01(w.query_seed, first * n, count * n, device).reshape(count, n)
    sel = _u01(w.query_seed + 1, first * n, count * n, device).reshape(count, n)
    amt = _u01(w.query_seed + 2, first * n, count * n, device).reshape(count, n)
    kind = _u01(w.query_seed + 3, first, count, device)
    cut = w.extra.get('corner_cut')
    base = u.copy() if device is None else u.clone()
    if cut is not None:
        deep = (base[:, 0] + base[:, 1]) > cut
        base[:, 0] = xp.where(deep, 1.0 - base[:, 0], base[:, 0])
        base[:, 1] = xp.where(deep, 1.0 - base[:, 1], base[:, 1])
    off = kind >= (1.0 / 3.0)
    push = sel < 0.25
    # make sure at least one axis is pushed for every off-grid row: the axis with the smallest sel
    if device is None:
        amin = np.argmin(sel, axis=1)
        push[np.arange(count), amin] = True
    else:
        amin = xp.argmin(sel, dim=1)
        push[xp.arange(count, device=device), amin] = True
    push = push & off.reshape(count, 1)
    side_hi = amt >= 0.5
    mag = 0.3 * xp.where(side_hi, 2.0 * (amt - 0.5), 2.0 * amt)
    out_u = xp.where(side_hi, 1.0 + mag, -mag)
    uu = xp.where(push, out_u, base)
    lo, hi = _span(w, device)
    return lo + (hi - lo) * uu


_GRIDS = {'C1': _grid_sum, 'C2': _grid_sum, 'C3': _grid_random, 'C4': _grid_random, 'C5': _grid_random}
_QUERIES = {'C1': _q_box, 'C2': _q_box, 'C3': _q_corner, 'C4': _q_corner, 'C5': _q_mixed}


def workload(name: str, scale: int | None = None, mix: str = 'near', n_queries: int | None = None,
             spacing: float | None = None, variant: str | None = None) -> Workload:
    """name in {'C1','C2','C3','C4','C5'}.  `scale` shrinks the per-axis tick count for
    tests (None = the size BASELINE.json names); the generated name records it ('C4@12').
    `spacing` (C2, C4, C5): ticks 1.7 + spacing * i instead of the integers 0..L-1 -- a spacing that is not a
    power of two takes the kernels' general index search (IEEE division) instead of the exact-multiply one."""
    tag = name if scale is None else f'{name}@{scale}'
    if variant is not None:
        # SURVEY 8d variants: 'iid' (C4): i.i.d. voids u(405, v) < 0.3 instead of the structured corner -- 0.7^32 = 1e-5 of
        # the hypercubes are fully defined, ~100 % of the queries extrapolate and the mask depends on all five axes;
        # 'allbasic' (C3): mu taken as a fourth basic axis (nbasic = 4)
        w = workload(name, scale, mix, n_queries, spacing)
        if variant == 'iid' and name == 'C4':
            w.extra = {'iid_void': 0.3, 'void_seed': 405}
        elif variant == 'allbasic' and name == 'C3':
            w.basic_axes = w.basic_axes + w.associated_axes
            w.associated_axes = None
        else:
            raise ValueError(f'variant {variant!r} does not apply to {name}')
        w.name = f'{w.name}/{variant}'
        return w
    if spacing is not None:
        w = workload(name, scale, mix, n_queries)
        if name not in ('C2', 'C4', 'C5'):
            raise ValueError('spacing applies to C2, C4 and C5')
        w.basic_axes = tuple(1.7 + spacing * np.arange(len(a), dtype=np.float64) for a in w.basic_axes)
        w.name = f'{w.name}/dx{spacing:g}'
        return w
    if name == 'C1':
        # README 3-D 5x5x5 linear scalar field, fully populated (reference README.md:100-127)
        ax = (np.linspace(1, 5, 5), np.linspace(10, 50, 5), np.linspace(100, 500, 5))
        return Workload(tag, ax, None, 1, 'nearest', n_queries or 10**6, 0, 101,
                        extra={'box': [(-5, 5), (-50, 50), (-500, 500)]})
    if name == 'C2':
        # benchmark_kdtree_speedup shape (reference benchmarks/benchmark_kdtree_speedup.py:18-49)
        L = scale or 50
        ax = tuple(np.linspace(0, L - 1, L) for _ in range(3))
        return Workload(tag, ax, None, 1, 'nearest', n_queries or 10**7, 0, 201,
                        extra={'iid_void': 0.2, 'void_seed': 202, 'margin': 0.3})
    if name == 'C3':
        s = scale or 80
        teff = np.geomspace(3500, 50000, s)
        logg = np.linspace(0, 5, max(4, s // 4))
        abun = np.linspace(-2.5, 0.5, max(3, s // 8))
        mu = np.linspace(0.02, 1, max(5, (s * 5) // 8))
        return Workload(tag, (teff, logg, abun), (mu,), 16, 'linear', n_queries or 10**8, 303, 301, mix=mix,
                        extra={'corner_cut': 1.2254, 'invert_axis1': True})
    if name == 'C4':
        L = scale or 40
        ax = tuple(np.linspace(0, L - 1, L) for _ in range(5))
        return Workload(tag, ax, None, 1, 'linear', n_queries or 10**9, 404, 401, mix=mix,
                        extra={'corner_cut': 1.2254})
    if name == 'C5':
        L = scale or 16
        ax = tuple(np.linspace(0, L - 1, L) for _ in range(6))
        return Workload(tag, ax, None, 1, 'linear', n_queries or 10**6, 504, 501, mix=mix,
                        extra={'corner_cut': 1.5528})
    raise ValueError(f'unknown workload {name}')
