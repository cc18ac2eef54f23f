"""CPU: the C-ABI library loads, exports every symbol include/ndpb200.h declares,
and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'ndpb200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ndpb_[a-z_0-9]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib_path():
    from ndpolator_b200 import build
    return build.build()


def test_header_and_binding_agree():
    from ndpolator_b200 import _lib
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib_path):
    lib = C.CDLL(lib_path)
    for sym in declared_symbols():
        assert hasattr(lib, sym), f'{sym} declared in include/ndpb200.h but not exported'
    lib.ndpb_version.restype = C.c_char_p
    assert b'ndpb200' in lib.ndpb_version()


def test_library_is_sm100a_only(lib_path):
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-lelf', lib_path], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    assert not re.search(r'sm_(?!100a)\d+', out)


def test_does_not_link_the_oracle(lib_path):
    import subprocess
    out = subprocess.run(['ldd', lib_path], capture_output=True, text=True).stdout
    assert 'oracle' not in out and 'ndp_ref' not in out and 'hostsim' not in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from ndpolator_b200 import Ndpolator
    from ndpolator_b200._lib import NdpbError
    ndp = Ndpolator((np.linspace(0, 1, 3),))
    ndp.register('t', None, np.zeros((3, 1)))
    with pytest.raises(NdpbError) as e:
        ndp.ndpolate('t', np.array([[0.5]]))
    assert e.value.code == 2 and 'no CPU fallback' in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'ndpolator_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f
                assert 'libndp_oracle' not in text and 'libndp_ref' not in text and 'libhostsim' not in text, f


def test_packaging_builds_and_ships_the_library(tmp_path):
    """pyproject.toml / setup.py (reference setup.py:4-21): build_py compiles the CUDA library through
    ndpolator_b200/build.py and the built package carries libndpb200.so, the ctypes mirror and the API class."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / 'lib'
    res = subprocess.run([sys.executable, 'setup.py', '-q', 'egg_info', '--egg-base', str(tmp_path), 'build_py', '--build-lib', str(out)], cwd=root,
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    pkg = out / 'ndpolator_b200'
    for name in ('libndpb200.so', 'cndpolator.py', 'ndpolator.py', '_lib.py', 'build.py'):
        assert (pkg / name).exists(), f'{name} missing from the built package'
    assert (pkg / 'csrc' / 'ndp_api.cu').exists()
