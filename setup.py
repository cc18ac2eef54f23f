"""setup.py -- compiles the CUDA library in-tree before the package is collected.

The reference builds its C extension through setuptools' Extension (reference setup.py:4-21); here
the native piece is a plain C-ABI shared library (include/ndpb200.h) loaded with ctypes, so the build
step is ndpolator_b200/build.py (nvcc, -gencode arch=compute_100a,code=sm_100a, -fmad=false) hooked
into build_py.  nvcc cross-compiles without a GPU; a missing nvcc fails the build loudly."""
import os
import sys

from setuptools import setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


class build_py_with_cuda(build_py):
    def run(self):
        sys.path.insert(0, HERE)
        from ndpolator_b200 import build as nb
        print('building', nb.build(verbose=bool(os.environ.get('NDPB_VERBOSE_BUILD'))))
        super().run()


setup(cmdclass={'build_py': build_py_with_cuda})
