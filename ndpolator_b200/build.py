"""In-tree build of libndpb200.so (the CUDA kernels + C ABI) for sm_100a.

nvcc cross-compiles without a GPU.  -fmad=false is part of the numerics contract
(no FMA contraction: every lerp keeps the reference's three roundings)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libndpb200.so')
# (source, object stem, extra defines): ndp_tu_fused.cu is compiled once per axis count
SOURCES = [('ndp_api.cu', 'ndp_api', []), ('ndp_tu_main_strided.cu', 'ndp_tu_main_strided', []),
           ('ndp_tu_main_cellmajor.cu', 'ndp_tu_main_cellmajor', []), ('ndp_tu_finish.cu', 'ndp_tu_finish', []),
           ('ndp_tu_search.cu', 'ndp_tu_search', []), ('ndp_tu_staged.cu', 'ndp_tu_staged', [])] + \
          [('ndp_tu_fused.cu', f'ndp_tu_fused_{n}', [f'-DNDP_FUSED_N={n}']) for n in (5, 6, 4, 3, 2)]
HEADERS = ['ndp_core.cuh', 'ndp_nsm.cuh', 'ndp_fused.cuh', 'ndp_kernels.cuh', 'ndp_fused_kernels.cuh', 'ndp_table.cuh',
           'ndp_launch.h', os.path.join('..', '..', 'include', 'ndpb200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17', '--extended-lambda',
    '-fmad=false',
    '-Xcompiler', '-fPIC',
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libndpb200.so cannot be built (there is no CPU fallback)')


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in [src for src, _, _ in SOURCES] + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, extra_flags=(), lib_out: str | None = None) -> str:
    """Compiles csrc/*.cu (one object per kernel family, in parallel) and links
    ndpolator_b200/libndpb200.so, if missing or out of date.  extra_flags + lib_out build a
    tuning variant (e.g. -DNDP_HARD_OCC=3) beside it; NDPB_LIB selects it at load time."""
    if lib_out is None and not force and not stale():
        return LIB
    out = lib_out or LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    objdir = os.path.join(HERE, 'build') if lib_out is None else os.path.splitext(out)[0] + '_obj'
    os.makedirs(objdir, exist_ok=True)

    def compile_one(item):
        src, stem, defines = item
        obj = os.path.join(objdir, stem + '.o')
        cmd = [nvcc] + NVCC_FLAGS + defines + list(extra_flags) + ['-c', '-o', obj, os.path.join(CSRC, src)]
        if verbose:
            print(' '.join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n' + res.stdout + res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', out] + objs
    if verbose:
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    return out


if __name__ == '__main__':
    print(build(force=True, verbose=True))
