"""Drop-in for the reference's CPython extension module ``cndpolator``
(reference src/ndp_py.c:448-456): the same four entry points with the same
argument names and return shapes, running on a B200 through libndpb200.so.

    ndpolate(capsule=None, query_pts=None, axes=None, grid=None, nbasic=0,
             extrapolation_method=0, search_algorithm=0)   src/ndp_py.c:343-409
    find(axes, query_pts, nbasic)                            src/ndp_py.c:131-193
    hypercubes(normed_query_pts, indices, axes, flags, grid, nbasic=0)
                                                             src/ndp_py.c:254-318
    distance(query_pts, capsule=None, axes=None, grid=None, nbasic=0)
                                                             src/ndp_py.c:196-251

The "capsule" is a :class:`Table` (device-resident axes, grid, masks, cell-major
grid copy and lazily built kd topology) instead of a PyCapsule around an
``ndp_table``.  Query points may be numpy arrays (results come back as numpy
arrays, staged through pipelined H2D/D2H chunks) or CUDA tensors / anything with
``__cuda_array_interface__`` (results stay on the device as torch tensors).
"""
from __future__ import annotations

import ctypes as C
import enum
import weakref

import numpy as np

from . import _lib


class ExtrapolationMethod(enum.IntEnum):      # src/ndp_py.c:433-437, src/ndpolator.h:28-32
    NONE = 0
    NEAREST = 1
    LINEAR = 2


class SearchAlgorithm(enum.IntEnum):          # src/ndp_py.c:439-442, src/ndpolator.h:44-47
    KDTREE = 0
    LINEAR = 1


def _is_cuda(x) -> bool:
    return hasattr(x, '__cuda_array_interface__') or (hasattr(x, 'is_cuda') and x.is_cuda)


def _ptr(x) -> int:
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, 'data_ptr'):
        return x.data_ptr()
    return x.__cuda_array_interface__['data'][0]


def _host_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _prep_queries(query_pts, naxes, device_index=None):
    """-> (array kept alive, pointer, n, on_device)"""
    if _is_cuda(query_pts):
        import torch
        q = query_pts if isinstance(query_pts, torch.Tensor) else torch.as_tensor(query_pts, device=f'cuda:{device_index or 0}')
        if device_index is not None and q.device.index != device_index:
            raise ValueError(f'query_pts lives on {q.device}, the table on cuda:{device_index}')
        if q.dtype != torch.float64 or not q.is_contiguous():
            q = q.to(torch.float64).contiguous()
        if q.dim() == 1:
            q = q.reshape(-1, naxes)
        if q.dim() != 2 or q.shape[1] != naxes:
            raise ValueError(f'query_pts must have shape (N, {naxes})')
        return q, q.data_ptr(), q.shape[0], True
    q = _host_f64(query_pts)
    if q.ndim == 1:
        q = q.reshape(-1, naxes)
    if q.ndim != 2 or q.shape[1] != naxes:
        raise ValueError(f'query_pts must have shape (N, {naxes})')
    return q, q.ctypes.data, q.shape[0], False


def _empty(shape, dtype, like_device, device_index):
    if like_device:
        import torch
        tdt = {np.float64: torch.float64, np.int32: torch.int32}[dtype]
        return torch.empty(shape, dtype=tdt, device=f'cuda:{device_index}')
    return np.empty(shape, dtype=dtype)


class Table:
    """Device-resident registered table; stands where the reference keeps its
    PyCapsule(ndp_table) (src/ndp_py.c:370)."""

    def __init__(self, axes, grid, nbasic=0, device=None):
        L = _lib.lib()
        self.device = _lib.default_device() if device is None else int(device)
        self.axes = tuple(_host_f64(a) for a in axes)
        self.naxes = len(self.axes)
        self.nbasic = int(nbasic) if nbasic else self.naxes
        axlen = (C.c_int32 * self.naxes)(*[len(a) for a in self.axes])
        axptr = (C.c_void_p * self.naxes)(*[a.ctypes.data for a in self.axes])
        if grid is None:
            gptr, self.vdim, self._grid = None, 1, None
        elif _is_cuda(grid):
            import torch
            g = grid if isinstance(grid, torch.Tensor) else torch.as_tensor(grid, device='cuda')
            g = g.to(torch.float64).contiguous()
            gptr, self.vdim, self._grid = g.data_ptr(), int(g.shape[-1]), g
        else:
            # like PyArray_FROM_OTF(grid, NPY_DOUBLE, NPY_ARRAY_CARRAY), src/ndp_py.c:120
            g = _host_f64(grid)
            gptr, self.vdim, self._grid = g.ctypes.data, int(g.shape[-1]), g
        if grid is not None:
            want = int(np.prod([len(a) for a in self.axes])) * self.vdim
            have = int(np.prod(self._grid.shape))
            if want != have:
                raise ValueError(f'grid has {have} elements, axes x vdim need {want}')
        h = C.c_void_p()
        _lib.check(L.ndpb_table_create(axptr, axlen, self.naxes, self.nbasic, gptr, self.vdim, self.device, C.byref(h)))
        self._h = h
        self._grid = None          # the library made its own device copy
        self._finalizer = weakref.finalize(self, L.ndpb_table_destroy, h)

    # -- bookkeeping -----------------------------------------------------------
    def close(self):
        self._finalizer()

    def set_option(self, key: str, value: int):
        _lib.check(_lib.lib().ndpb_table_set_option(self._h, key.encode(), int(value)))

    def stat(self, key: str) -> int:
        return int(_lib.lib().ndpb_table_stat(self._h, key.encode()))

    def sync(self):
        """Completes stream-ordered calls (set_option('async', 1): calls on device tensors only enqueue their
        kernels -- on the table's stream, or on the caller's with set_option('stream', ...) -- and return)."""
        _lib.check(_lib.lib().ndpb_table_sync(self._h))

    @property
    def nverts(self) -> int:
        return int(_lib.lib().ndpb_table_nverts(self._h))

    def masks(self):
        vm = np.empty(self.nverts, np.int32)
        hm = np.empty(self.nverts, np.int32)
        _lib.check(_lib.lib().ndpb_table_masks(self._h, vm.ctypes.data, hm.ctypes.data))
        return vm, hm

    def tree_topology(self, method):
        parent = np.empty(self.nverts, np.int32)
        depth = np.empty(self.nverts, np.int32)
        _lib.check(_lib.lib().ndpb_table_tree_topology(self._h, int(method), parent.ctypes.data, depth.ctypes.data))
        return parent, depth

    def _order_after_producer(self, on_device):
        """Device-resident inputs: make the table's stream wait for what the caller's current torch
        stream has enqueued so far (ndpb_table_wait_stream; host inputs need no ordering)."""
        if on_device:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().ndpb_table_wait_stream(self._h, stream))

    def _check_out(self, out, n, on_device):
        interps, dists = out
        for arr, shape, what in ((interps, (n, self.vdim), 'interps'), (dists, (n, 1), 'dists')):
            if _is_cuda(arr) != on_device:
                raise ValueError(f'out[{what}] must live where query_pts lives (device or host)')
            if on_device:
                import torch
                if arr.device.index != self.device or arr.dtype != torch.float64 or not arr.is_contiguous():
                    raise ValueError(f'out[{what}] must be a contiguous float64 tensor on cuda:{self.device}')
            elif arr.dtype != np.float64 or not arr.flags['C_CONTIGUOUS']:
                raise ValueError(f'out[{what}] must be a C-contiguous float64 array')
            if tuple(arr.shape) != shape and not (what == 'dists' and tuple(arr.shape) == (n,)):
                raise ValueError(f'out[{what}] must have shape {shape}')

    # -- hot path ----------------------------------------------------------------
    def find(self, query_pts):
        q, qp, n, dev = _prep_queries(query_pts, self.naxes, self.device)
        self._order_after_producer(dev)
        idx = _empty((n, self.naxes), np.int32, dev, self.device)
        flg = _empty((n, self.naxes), np.int32, dev, self.device)
        nrm = _empty((n, self.naxes), np.float64, dev, self.device)
        _lib.check(_lib.lib().ndpb_find(self._h, qp, n, _ptr(idx), _ptr(flg), _ptr(nrm)))
        return idx, flg, nrm

    def _cached_inputs(self, normed, indices, flags):
        """-> (normed, indices, flags, n, on_device): contiguous float64 / int32 / int32, all on the host or all on
        this table's GPU (device tensors stay where they are: the cache path never round-trips through the host)."""
        dev = _is_cuda(normed)
        if dev != _is_cuda(indices) or dev != _is_cuda(flags):
            raise ValueError('normed_query_pts, indices and flags must all live on the host or all on the device')
        if dev:
            import torch
            nrm = normed.to(torch.float64).contiguous()
            idx = indices.to(torch.int32).contiguous()
            flg = flags.to(torch.int32).contiguous()
            for t in (nrm, idx, flg):
                if t.device.index != self.device:
                    raise ValueError(f'cached import products live on {t.device}, the table on cuda:{self.device}')
            self._order_after_producer(True)
        else:
            nrm = _host_f64(normed)
            idx = np.ascontiguousarray(indices, dtype=np.int32)
            flg = np.ascontiguousarray(flags, dtype=np.int32)
        if tuple(nrm.shape) != tuple(idx.shape) or tuple(nrm.shape) != tuple(flg.shape) or nrm.shape[-1] != self.naxes:
            raise ValueError(f'cached import products must all have shape (N, {self.naxes})')
        return nrm, idx, flg, int(idx.shape[0]), dev

    def hypercubes_raw(self, normed, indices, flags):
        """-> (dim (n,), fdhc (n,), v (n, 2^naxes * vdim)): one packed block per output instead of N arrays; the
        first 2^dim[i] * vdim values of row i are the hypercube of query i.  Device inputs give device outputs."""
        nrm, idx, flg, n, dev = self._cached_inputs(normed, indices, flags)
        slot = (1 << self.naxes) * self.vdim
        dim = _empty((n,), np.int32, dev, self.device)
        fdhc = _empty((n,), np.int32, dev, self.device)
        v = _empty((n, slot), np.float64, dev, self.device)
        _lib.check(_lib.lib().ndpb_hypercubes(self._h, _ptr(nrm), _ptr(idx), _ptr(flg), n, _ptr(dim), _ptr(fdhc), _ptr(v)))
        return dim, fdhc, v

    def ndpolate_cached(self, normed, indices, flags, method=0, search=0):
        """ndpolate from cached import products (what find() returned) -- the hypercube-caching path of the
        reference's README, device-resident when the cache is.  Same results as ndpolate() on the original points."""
        nrm, idx, flg, n, dev = self._cached_inputs(normed, indices, flags)
        interps = _empty((n, self.vdim), np.float64, dev, self.device)
        dists = _empty((n, 1), np.float64, dev, self.device)
        _lib.check(_lib.lib().ndpb_ndpolate_cached(self._h, _ptr(nrm), _ptr(idx), _ptr(flg), n, int(method), int(search),
                                                    _ptr(interps), _ptr(dists)))
        return interps, dists

    def find_nearest(self, query_pts, method, search):
        q, qp, n, dev = _prep_queries(query_pts, self.naxes, self.device)
        self._order_after_producer(dev)
        coords = _empty((n, self.naxes), np.int32, dev, self.device)
        dist = _empty((n,), np.float64, dev, self.device)
        _lib.check(_lib.lib().ndpb_find_nearest(self._h, qp, n, int(method), int(search), _ptr(coords), _ptr(dist)))
        return coords, dist

    def ndpolate(self, query_pts, method=0, search=0, out=None):
        q, qp, n, dev = _prep_queries(query_pts, self.naxes, self.device)
        self._order_after_producer(dev)
        if out is None:
            interps = _empty((n, self.vdim), np.float64, dev, self.device)
            dists = _empty((n, 1), np.float64, dev, self.device)
        else:
            self._check_out(out, n, dev)
            interps, dists = out
        _lib.check(_lib.lib().ndpb_ndpolate(self._h, qp, n, int(method), int(search), _ptr(interps), _ptr(dists)))
        return interps, dists

    def distance(self, query_pts):
        q, qp, n, dev = _prep_queries(query_pts, self.naxes, self.device)
        self._order_after_producer(dev)
        dists = _empty((n, 1), np.float64, dev, self.device)
        _lib.check(_lib.lib().ndpb_distance(self._h, qp, n, _ptr(dists)))
        return dists


class PackedHypercubes:
    """Device-resident result of find_hypercubes(): `values` (N, 2^naxes * vdim) holds the hypercube of query i
    in its first 2^dim[i] * vdim entries (corner-major, first free axis = most significant bit, as in the
    reference); `dim` (N,) and `fdhc` (N,) are device tensors too.  hc[i] gives what the reference's tuple
    element i is, as a device tensor of shape (2,)*dim[i] + (vdim,); len(hc) == N."""

    def __init__(self, dim, fdhc, values, vdim):
        self.dim, self.fdhc, self.values, self.vdim = dim, fdhc, values, vdim
        self._dim_host = None

    def __len__(self):
        return int(self.dim.shape[0])

    def __getitem__(self, i):
        if self._dim_host is None:
            self._dim_host = self.dim.cpu().numpy()
        d = int(self._dim_host[i])
        return self.values[i, :(1 << d) * self.vdim].reshape((2,) * d + (self.vdim,))

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class Group:
    """One registered table replicated on several GPUs of this process (ndpb_group): a host batch is
    cut into contiguous query ranges, one per device, and every device copies its results into its
    slice of the output arrays -- no collective.  Same results as a single :class:`Table`.
    Stands where the reference keeps its PyCapsule, like Table; host (numpy) buffers only."""

    def __init__(self, axes, grid, nbasic=0, devices=(0,)):
        L = _lib.lib()
        self.devices = [int(d) for d in devices]
        self.device = self.devices[0]
        self.axes = tuple(_host_f64(a) for a in axes)
        self.naxes = len(self.axes)
        self.nbasic = int(nbasic) if nbasic else self.naxes
        g = _host_f64(grid)
        self.vdim = int(g.shape[-1])
        want = int(np.prod([len(a) for a in self.axes])) * self.vdim
        if want != g.size:
            raise ValueError(f'grid has {g.size} elements, axes x vdim need {want}')
        axlen = (C.c_int32 * self.naxes)(*[len(a) for a in self.axes])
        axptr = (C.c_void_p * self.naxes)(*[a.ctypes.data for a in self.axes])
        devs = (C.c_int32 * len(self.devices))(*self.devices)
        h = C.c_void_p()
        _lib.check(L.ndpb_group_create(axptr, axlen, self.naxes, self.nbasic, g.ctypes.data, self.vdim,
                                       devs, len(self.devices), C.byref(h)))
        self._h = h
        self._finalizer = weakref.finalize(self, L.ndpb_group_destroy, h)

    def close(self):
        self._finalizer()

    def set_option(self, key: str, value: int):
        _lib.check(_lib.lib().ndpb_group_set_option(self._h, key.encode(), int(value)))

    def stat(self, key: str) -> int:
        return int(_lib.lib().ndpb_group_stat(self._h, key.encode()))

    def _host_queries(self, query_pts):
        if _is_cuda(query_pts):
            raise ValueError('a multi-GPU table takes host (numpy) query arrays; a device tensor belongs to one GPU')
        return _prep_queries(query_pts, self.naxes)

    def ndpolate(self, query_pts, method=0, search=0, out=None):
        q, qp, n, _ = self._host_queries(query_pts)
        if out is None:
            interps, dists = np.empty((n, self.vdim), np.float64), np.empty((n, 1), np.float64)
        else:
            interps, dists = out
            for arr, what in ((interps, 'interps'), (dists, 'dists')):
                if not isinstance(arr, np.ndarray) or arr.dtype != np.float64 or not arr.flags['C_CONTIGUOUS']:
                    raise ValueError(f'out[{what}] must be a C-contiguous float64 numpy array')
            if interps.size != n * self.vdim or dists.size != n:
                raise ValueError('out arrays have the wrong size')
        _lib.check(_lib.lib().ndpb_group_ndpolate(self._h, qp, n, int(method), int(search), _ptr(interps), _ptr(dists)))
        return interps, dists

    def distance(self, query_pts):
        q, qp, n, _ = self._host_queries(query_pts)
        dists = np.empty((n, 1), np.float64)
        _lib.check(_lib.lib().ndpb_group_distance(self._h, qp, n, _ptr(dists)))
        return dists

    def member(self, i=0):
        """Borrowed :class:`Table` view of member i (taps such as find / masks / find_nearest)."""
        t = Table.__new__(Table)
        t.device, t.axes, t.naxes, t.nbasic, t.vdim = self.devices[i], self.axes, self.naxes, self.nbasic, self.vdim
        t._h = C.c_void_p(_lib.lib().ndpb_group_table(self._h, i))
        t._grid = None
        t._owner = self                       # keeps the group alive; the member is not destroyed separately
        t._finalizer = lambda: None
        return t

    def find(self, query_pts):
        return self.member(0).find(query_pts)

    def hypercubes_raw(self, normed, indices, flags):
        return self.member(0).hypercubes_raw(normed, indices, flags)


# ---------------------------------------------------------------------------
# module-level functions with the reference's signatures
# ---------------------------------------------------------------------------

def ndpolate(capsule=None, query_pts=None, axes=None, grid=None, nbasic=0,
             extrapolation_method=0, search_algorithm=0):
    """-> (interps, dists) with a valid capsule, else (interps, dists, capsule)."""
    if isinstance(capsule, (Table, Group)):
        interps, dists = capsule.ndpolate(query_pts, extrapolation_method, search_algorithm)
        return interps, dists
    if query_pts is not None and axes is not None and grid is not None:
        table = Table(axes, grid, nbasic)
        interps, dists = table.ndpolate(query_pts, extrapolation_method, search_algorithm)
        return interps, dists, table
    # the reference returns NULL without an exception here (src/ndp_py.c:372-374) -> SystemError
    raise SystemError('ndpolate() needs either a capsule or query_pts, axes and grid')


def find(axes, query_pts, nbasic):
    """-> (indices int32, flags int32, normed float64), each shaped like query_pts."""
    table = Table(axes, None, nbasic)
    try:
        return table.find(query_pts)
    finally:
        table.close()


def hypercubes(normed_query_pts, indices, axes, flags, grid, nbasic=0, capsule=None):
    """-> tuple of N arrays of shape (2,)*dim + (vdim,)."""
    table = capsule if isinstance(capsule, (Table, Group)) else Table(axes, grid, nbasic)
    try:
        dim, fdhc, v = table.hypercubes_raw(normed_query_pts, indices, flags)
    finally:
        if table is not capsule:
            table.close()
    vdim = table.vdim
    if _is_cuda(v):
        # device-resident cache: one packed block (PackedHypercubes) instead of N arrays; index it like the tuple
        return PackedHypercubes(dim, fdhc, v, vdim)
    return tuple(v[i, :(1 << int(d)) * vdim].reshape((2,) * int(d) + (vdim,)).copy() for i, d in enumerate(dim))


def distance(query_pts, capsule=None, axes=None, grid=None, nbasic=0):
    """-> (N, 1) squared index-space distances to the nearest fully defined hypercube."""
    if isinstance(capsule, (Table, Group)):
        return capsule.distance(query_pts)
    if axes is None or grid is None:
        raise ValueError('Either capsule must be valid or axes and grid must be provided')   # src/ndp_py.c:212-216
    table = Table(axes, grid, nbasic if nbasic else np.shape(grid)[0])   # src/ndp_py.c:224 (sic)
    try:
        return table.distance(query_pts)
    finally:
        table.close()
