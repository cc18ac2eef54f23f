#!/usr/bin/env python
"""Builds tuning variants of libndpb200.so beside the default one (tools/bin/libndpb200_<tag>.so; same ABI,
same results, selected at load time with NDPB_LIB).  Usage: python tools/build_variants.py tag=-DFLAG=V,-DFLAG2=W ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndpolator_b200 import build  # noqa: E402

os.makedirs(os.path.join(ROOT, 'tools', 'bin'), exist_ok=True)
for spec in sys.argv[1:]:
    tag, flags = spec.split('=', 1)
    out = os.path.join(ROOT, 'tools', 'bin', f'libndpb200_{tag}.so')
    print(build.build(extra_flags=[f for f in flags.split(',') if f], lib_out=out))
