"""ctypes driver of tests/hostsim/hostsim.cpp -- TEST SCAFFOLDING (see that file)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# NDPB_HOSTSIM_FLAGS: extra -D switches of ndp_core.cuh under test (e.g. -DNDP_FAST_MARGIN), built beside the default
FLAGS = os.environ.get('NDPB_HOSTSIM_FLAGS', '').split()
LIB = os.path.join(HERE, 'libhostsim' + ''.join(f.replace('-D', '_').replace('=', '') for f in FLAGS) + '.so')
SRC = os.path.join(HERE, 'hostsim.cpp')
CSRC = os.path.join(HERE, '..', '..', 'ndpolator_b200', 'csrc')
CORE = [os.path.join(CSRC, f) for f in ('ndp_core.cuh', 'ndp_nsm.cuh', 'ndp_fused.cuh')]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build():
    if (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in [SRC] + CORE)):
        subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC'] + FLAGS + ['-o', LIB, SRC], check=True)
    lib = C.CDLL(LIB)
    lib.hs_create.restype = C.c_void_p
    lib.hs_create.argtypes = [C.c_int, C.c_int, _ip, _dp, C.c_int, _dp, C.c_int]
    return lib


def _p(a, t):
    return a.ctypes.data_as(t)


class HostSim:
    def __init__(self, axes, grid, nbasic):
        self.lib = build()
        self.axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
        self.n = len(axes)
        self.axlen = np.array([len(a) for a in axes], np.int32)
        self.ticks = np.concatenate(self.axes)
        self.grid = np.ascontiguousarray(grid, dtype=np.float64)
        self.vdim = grid.shape[-1]
        self.h = C.c_void_p(self.lib.hs_create(self.n, nbasic, _p(self.axlen, _ip), _p(self.ticks, _dp), self.vdim, _p(self.grid, _dp), 1))
        self.nverts = int(np.prod(self.axlen[:nbasic]))

    def __del__(self):
        try:
            self.lib.hs_free(self.h)
        except Exception:
            pass

    def masks(self):
        v, h = np.empty(self.nverts, np.int32), np.empty(self.nverts, np.int32)
        self.lib.hs_masks(self.h, _p(v, _ip), _p(h, _ip))
        return v, h

    def topology(self, method):
        p, d = np.empty(self.nverts, np.int32), np.empty(self.nverts, np.int32)
        self.lib.hs_topology(self.h, int(method), _p(p, _ip), _p(d, _ip))
        return p, d

    def find(self, q):
        i, f, n = np.empty(q.shape, np.int32), np.empty(q.shape, np.int32), np.empty(q.shape)
        self.lib.hs_find(self.h, C.c_longlong(len(q)), _p(q, _dp), _p(i, _ip), _p(f, _ip), _p(n, _dp))
        return i, f, n

    def find_nearest(self, q, method, search):
        c, d = np.empty(q.shape, np.int32), np.empty(len(q))
        self.lib.hs_find_nearest(self.h, C.c_longlong(len(q)), _p(q, _dp), int(method), int(search), _p(c, _ip), _p(d, _dp))
        return c, d

    def ndpolate(self, q, method, search, variant):
        i, d = np.empty((len(q), self.vdim)), np.empty((len(q), 1))
        self.lib.hs_ndpolate(self.h, C.c_longlong(len(q)), _p(q, _dp), int(method), int(search), int(variant), _p(i, _dp), _p(d, _dp))
        return i, d

    def set_block_order(self, order):
        """Blocked copy of the grid (ndp_fused.cuh: NdpBlocks) for the fused path; 0 on tables it does not apply to."""
        return int(self.lib.hs_set_block_order(self.h, int(order)))

    def ndpolate_fused(self, q, method, search):
        """The fused path's logic (plan from masks + nearest-site map, one gather); None when the
        table or mode does not qualify for it."""
        i, d = np.empty((len(q), self.vdim)), np.empty((len(q), 1))
        ok = self.lib.hs_ndpolate_fused(self.h, C.c_longlong(len(q)), _p(q, _dp), int(method), int(search), _p(i, _dp), _p(d, _dp))
        return (i, d) if ok else None

    def stats(self):
        o = (C.c_longlong * 5)()
        self.lib.hs_stats(self.h, o)
        return {'offgrid': o[0], 'hard': o[1], 'ties': o[2], 'slow': o[3], 'mapped': o[4]}

    def nsm_info(self):
        o = (C.c_longlong * 12)()
        self.lib.hs_nsm_info(self.h, o)
        return [{'valid': o[4 * g], 'S': o[4 * g + 1], 'regions': o[4 * g + 2], 'entries': o[4 * g + 3]} for g in range(3)]
