"""Generates tests/golden/ref_<case>.npz by running the UNMODIFIED reference C
core (oracle/_ref/libndp_ref.so, built in place from /root/reference by
oracle/Makefile) on the seeded cases of tests/cases.py.

Run in the build container (the only place /root/reference exists):

    make -C oracle && python tests/golden/make_golden.py

The fixtures pin both the CPU restatement (oracle/ndp_oracle.c, CPU tests) and
the CUDA path (GPU tests) to outputs of the reference itself; nothing reads
/root/reference at test time."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import adversarial_queries, build_cases  # noqa: E402
from oracle.oracle import Oracle, have  # noqa: E402


def main(only=()):
    if not have('reference'):
        raise SystemExit('oracle/_ref/libndp_ref.so missing: run `make -C oracle` where /root/reference is mounted')
    for ci, (name, axes, nbasic, grid) in enumerate(build_cases()):
        if only and name not in only:
            continue
        rng = np.random.default_rng(1000 + ci)
        q = adversarial_queries(axes, 400, rng)
        R = Oracle('reference', axes, grid, nbasic)
        out = {'nbasic': np.int32(nbasic), 'grid': grid, 'q': q}
        for a, ax in enumerate(axes):
            out[f'axis{a}'] = ax
        out['vmask'], out['hcmask'] = R.masks()
        idx, flg, nrm = R.find(q)
        out.update(indices=idx, flags=flg, normed=nrm)
        dim, fdhc, v = R.hypercubes(idx, flg, nrm)
        out.update(hc_dim=dim, hc_fdhc=fdhc)
        if v.size <= 200_000:
            out['hc_v'] = v
        for m in (0, 1, 2):
            for s in (0, 1):
                interps, dists = R.ndpolate(q, m, s)
                out[f'interps_m{m}_s{s}'] = interps
                out[f'dists_m{m}_s{s}'] = dists
                if m:
                    coords, dist = R.find_nearest(idx, flg, nrm, m, s)
                    out[f'coords_m{m}_s{s}'] = coords
                    out[f'ndist_m{m}_s{s}'] = dist
        out['distance'] = R.distance(q)
        path = os.path.join(HERE, f'ref_{name}.npz')
        np.savez_compressed(path, **out)
        print(f'{name}: {len(q)} queries -> {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main(sys.argv[1:])      # optional: case names to (re)generate; default all
