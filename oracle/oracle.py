"""ctypes driver for the two CPU parity checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle('port')``      -> oracle/libndp_oracle.so   (our restatement, oracle/ndp_oracle.c)
* ``Oracle('reference')`` -> oracle/_ref/libndp_ref.so (the unmodified reference C core
  behind oracle/ref_shim.c; present only where oracle/Makefile could see /root/reference,
  or on the GPU box via the repo snapshot)

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of bench.py import this module.  The product package ``ndpolator_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, 'libndp_oracle.so')
REF_LIB = os.path.join(HERE, '_ref', 'libndp_ref.so')

NONE, NEAREST, LINEAR = 0, 1, 2          # reference src/ndpolator.h:28-32
KDTREE, LINEAR_SEARCH = 0, 1             # reference src/ndpolator.h:44-47


def build(quiet: bool = True) -> None:
    """(Re)builds both checkers with oracle/Makefile (building the checker is not using it)."""
    subprocess.run(['make', '-C', HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have(kind: str) -> bool:
    return os.path.exists(PORT_LIB if kind == 'port' else REF_LIB)


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def _p(a, t):
    return a.ctypes.data_as(t)


class Oracle:
    """One registered table on one of the CPU checkers (flat C entry points, GIL released)."""

    def __init__(self, kind, axes, grid, nbasic=None):
        self.kind = kind
        path, self.pfx = (PORT_LIB, 'ndpo_') if kind == 'port' else (REF_LIB, 'ndpref_')
        if not os.path.exists(path) and kind == 'port':
            build()
        self.lib = C.CDLL(path)
        self.axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
        self.naxes = len(self.axes)
        self.nbasic = self.naxes if not nbasic else int(nbasic)
        self.axlen = np.array([len(a) for a in self.axes], dtype=np.int32)
        self.ticks = np.concatenate(self.axes)
        if grid is None:
            self.grid, self.vdim = None, 1
        else:
            self.grid = np.ascontiguousarray(grid, dtype=np.float64)
            self.vdim = int(self.grid.shape[-1])
        f = self._fn('table_new', C.c_void_p, [C.c_int, C.c_int, _ip, _dp, C.c_int, _dp])
        self.h = f(self.naxes, self.nbasic, _p(self.axlen, _ip), _p(self.ticks, _dp), self.vdim,
                   _p(self.grid, _dp) if self.grid is not None else None)
        if not self.h:
            raise RuntimeError('oracle table_new failed')

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, self.pfx + name)
        f.restype, f.argtypes = restype, argtypes
        return f

    def close(self):
        if getattr(self, 'h', None):
            self._fn('table_free', None, [C.c_void_p])(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- register-time products -------------------------------------------------
    @property
    def nverts(self):
        return int(self._fn('table_nverts', C.c_int64, [C.c_void_p])(self.h))

    def masks(self):
        nv = self.nverts
        vm, hm = np.empty(nv, np.int32), np.empty(nv, np.int32)
        self._fn('table_masks', None, [C.c_void_p, _ip, _ip])(self.h, _p(vm, _ip), _p(hm, _ip))
        return vm, hm

    def warm(self, method):
        self._fn('table_warm', None, [C.c_void_p, C.c_int])(self.h, int(method))

    def tree_topology(self, method):
        """(parent vertex id, depth) per vertex id of the insertion-ordered kd-tree: the port's own
        index-linked tree, or -- kind 'reference' -- a walk over the reference's real struct kdnode."""
        nv = self.nverts
        parent = np.full(nv, -2, np.int64)
        depth = np.full(nv, -1, np.int32)
        self._fn('tree_topology', None, [C.c_void_p, C.c_int, _lp, _ip])(self.h, int(method), _p(parent, _lp), _p(depth, _ip))
        return parent, depth

    # -- per-query entry points -------------------------------------------------
    def _q(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.ndim == 2 and q.shape[1] == self.naxes
        return q

    def find(self, q):
        q = self._q(q)
        n = q.shape[0]
        idx, flg = np.empty(q.shape, np.int32), np.empty(q.shape, np.int32)
        nrm = np.empty(q.shape, np.float64)
        self._fn('import', None, [C.c_void_p, C.c_int64, _dp, _ip, _ip, _dp])(
            self.h, n, _p(q, _dp), _p(idx, _ip), _p(flg, _ip), _p(nrm, _dp))
        return idx, flg, nrm

    def hypercubes(self, idx, flg, nrm):
        n = idx.shape[0]
        slot = (1 << self.naxes) * self.vdim
        dim, fdhc = np.empty(n, np.int32), np.empty(n, np.int32)
        v = np.full((n, slot), np.nan)
        self._fn('hypercubes', None, [C.c_void_p, C.c_int64, _ip, _ip, _dp, _ip, _ip, _dp])(
            self.h, n, _p(np.ascontiguousarray(idx, np.int32), _ip), _p(np.ascontiguousarray(flg, np.int32), _ip),
            _p(np.ascontiguousarray(nrm, np.float64), _dp), _p(dim, _ip), _p(fdhc, _ip), _p(v, _dp))
        return dim, fdhc, v

    def find_nearest(self, idx, flg, nrm, method, search):
        n = idx.shape[0]
        coords = np.empty((n, self.naxes), np.int32)
        dist = np.empty(n, np.float64)
        self._fn('find_nearest', None, [C.c_void_p, C.c_int64, _ip, _ip, _dp, C.c_int, C.c_int, _ip, _dp])(
            self.h, n, _p(np.ascontiguousarray(idx, np.int32), _ip), _p(np.ascontiguousarray(flg, np.int32), _ip),
            _p(np.ascontiguousarray(nrm, np.float64), _dp), int(method), int(search), _p(coords, _ip), _p(dist, _dp))
        return coords, dist

    def ndpolate(self, q, method=NONE, search=KDTREE, threads=1):
        """(interps (N,vdim), dists (N,1)).  threads>1 shards the batch over host threads
        (ctypes drops the GIL; the lazily built tree is warmed first so workers only read)."""
        q = self._q(q)
        n = q.shape[0]
        interps = np.empty((n, self.vdim), np.float64)
        dists = np.empty((n, 1), np.float64)
        f = self._fn('ndpolate', C.c_int, [C.c_void_p, C.c_int64, _dp, C.c_int, C.c_int, _dp, _dp])

        def run(lo, hi):
            if hi > lo:
                rc = f(self.h, hi - lo, _p(q[lo:hi], _dp), int(method), int(search),
                       _p(interps[lo:hi], _dp), _p(dists[lo:hi], _dp))
                if rc:
                    raise ValueError('invalid extrapolation method')

        if threads <= 1 or n < 2 * threads:
            run(0, n)
        else:
            if search == KDTREE and method != NONE:
                self.warm(method)
            cuts = np.linspace(0, n, threads + 1).astype(np.int64)
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(lambda k: run(int(cuts[k]), int(cuts[k + 1])), range(threads)))
        return interps, dists

    def distance(self, q):
        q = self._q(q)
        n = q.shape[0]
        d = np.empty((n, 1), np.float64)
        self._fn('distance', None, [C.c_void_p, C.c_int64, _dp, _dp])(self.h, n, _p(q, _dp), _p(d, _dp))
        return d
