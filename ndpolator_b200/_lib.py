"""ctypes binding of libndpb200.so (include/ndpb200.h).

The library is loaded lazily and loudly: a missing .so, a missing symbol or a
missing CUDA device raises -- there is no CPU path behind this module."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NDPB_LIB: a tuning variant built by build.build(extra_flags=..., lib_out=...); same ABI, same results
LIB_PATH = os.environ.get('NDPB_LIB') or os.path.join(HERE, 'libndpb200.so')

# every symbol include/ndpb200.h declares (tests check the export list against the header)
SYMBOLS = [
    'ndpb_table_create', 'ndpb_table_destroy', 'ndpb_find', 'ndpb_hypercubes', 'ndpb_find_nearest',
    'ndpb_ndpolate', 'ndpb_ndpolate_cached', 'ndpb_distance', 'ndpb_table_nverts', 'ndpb_table_masks', 'ndpb_table_tree_topology',
    'ndpb_table_set_option', 'ndpb_table_stat', 'ndpb_table_wait_stream', 'ndpb_table_sync', 'ndpb_last_error', 'ndpb_version',
    'ndpb_group_create', 'ndpb_group_destroy', 'ndpb_group_size', 'ndpb_group_table', 'ndpb_group_ndpolate',
    'ndpb_group_distance', 'ndpb_group_set_option', 'ndpb_group_stat',
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

_lib = None


class NdpbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'libndpb200 error {code}: {msg}')
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build it with `python -m ndpolator_b200.build` '
            '(nvcc, sm_100a).  ndpolator_b200 has no CPU fallback.')
    L = C.CDLL(LIB_PATH)
    L.ndpb_table_create.restype = C.c_int
    L.ndpb_table_create.argtypes = [C.POINTER(_vp), _ip, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.POINTER(_vp)]
    L.ndpb_table_destroy.restype = C.c_int
    L.ndpb_table_destroy.argtypes = [_vp]
    L.ndpb_find.restype = C.c_int
    L.ndpb_find.argtypes = [_vp, _vp, C.c_int64, _vp, _vp, _vp]
    L.ndpb_hypercubes.restype = C.c_int
    L.ndpb_hypercubes.argtypes = [_vp, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp]
    L.ndpb_find_nearest.restype = C.c_int
    L.ndpb_find_nearest.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]
    L.ndpb_ndpolate.restype = C.c_int
    L.ndpb_ndpolate.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]
    L.ndpb_ndpolate_cached.restype = C.c_int
    L.ndpb_ndpolate_cached.argtypes = [_vp, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]
    L.ndpb_distance.restype = C.c_int
    L.ndpb_distance.argtypes = [_vp, _vp, C.c_int64, _vp]
    L.ndpb_table_nverts.restype = C.c_int64
    L.ndpb_table_nverts.argtypes = [_vp]
    L.ndpb_table_masks.restype = C.c_int
    L.ndpb_table_masks.argtypes = [_vp, _vp, _vp]
    L.ndpb_table_tree_topology.restype = C.c_int
    L.ndpb_table_tree_topology.argtypes = [_vp, C.c_int, _vp, _vp]
    L.ndpb_table_set_option.restype = C.c_int
    L.ndpb_table_set_option.argtypes = [_vp, C.c_char_p, C.c_int64]
    L.ndpb_table_wait_stream.restype = C.c_int
    L.ndpb_table_wait_stream.argtypes = [_vp, C.c_uint64]
    L.ndpb_table_sync.restype = C.c_int
    L.ndpb_table_sync.argtypes = [_vp]
    L.ndpb_table_stat.restype = C.c_int64
    L.ndpb_table_stat.argtypes = [_vp, C.c_char_p]
    L.ndpb_group_create.restype = C.c_int
    L.ndpb_group_create.argtypes = [C.POINTER(_vp), _ip, C.c_int, C.c_int, _vp, C.c_int, _ip, C.c_int, C.POINTER(_vp)]
    L.ndpb_group_destroy.restype = C.c_int
    L.ndpb_group_destroy.argtypes = [_vp]
    L.ndpb_group_size.restype = C.c_int
    L.ndpb_group_size.argtypes = [_vp]
    L.ndpb_group_table.restype = _vp
    L.ndpb_group_table.argtypes = [_vp, C.c_int]
    L.ndpb_group_ndpolate.restype = C.c_int
    L.ndpb_group_ndpolate.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]
    L.ndpb_group_distance.restype = C.c_int
    L.ndpb_group_distance.argtypes = [_vp, _vp, C.c_int64, _vp]
    L.ndpb_group_set_option.restype = C.c_int
    L.ndpb_group_set_option.argtypes = [_vp, C.c_char_p, C.c_int64]
    L.ndpb_group_stat.restype = C.c_int64
    L.ndpb_group_stat.argtypes = [_vp, C.c_char_p]
    L.ndpb_last_error.restype = C.c_char_p
    L.ndpb_version.restype = C.c_char_p
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise NdpbError(rc, lib().ndpb_last_error().decode())


def default_device() -> int:
    """One process per GPU: NDPB_DEVICE, else LOCAL_RANK (torchrun), else 0."""
    for var in ('NDPB_DEVICE', 'LOCAL_RANK'):
        if os.environ.get(var, '') != '':
            return int(os.environ[var])
    return 0
