"""CPU: the per-query device logic (ndpolator_b200/csrc/ndp_core.cuh, the code the
CUDA kernels wrap) compiled for the host by tests/hostsim and diffed bit for bit
against the oracle: masks, import, the rebuilt kd topology, nearest search in all
four modes (fast path, pyramid descent, tie resolution), and ndpolate through the
strided, cell-major and runtime-n gather variants.  This is a development check of
the kernel logic in the GPU-less container; the GPU parity suite is test_gpu_parity.py."""
import shutil

import os

import numpy as np
import pytest

from cases import adversarial_queries, assert_same_bits, build_cases
from oracle.oracle import Oracle

pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason='g++ not available')


@pytest.mark.parametrize('case', build_cases(seed=3), ids=lambda c: c[0])
def test_device_logic_matches_oracle(case):
    from hostsim.driver import HostSim
    name, axes, nbasic, grid = case
    q = adversarial_queries(axes, 700, np.random.default_rng(9))
    O, H = Oracle('port', axes, grid, nbasic), HostSim(axes, grid, nbasic)
    for a, b, what in zip(O.masks(), H.masks(), ('vmask', 'hcmask')):
        assert_same_bits(a, b, what)
    fo = O.find(q)
    for a, b, what in zip(fo, H.find(q), ('indices', 'flags', 'normed')):
        assert_same_bits(a, b, what)
    for m in (1, 2):
        po, do = O.tree_topology(m)
        ph, dh = H.topology(m)
        assert np.array_equal(po, ph.astype(np.int64)), f'kd parent, method {m}'
        assert np.array_equal(do, dh), f'kd depth, method {m}'
        for s in (0, 1):
            a, b = O.find_nearest(*fo, m, s), H.find_nearest(q, m, s)
            assert_same_bits(a[0], b[0], f'coords {m}/{s}')
            assert_same_bits(a[1], b[1], f'ndist {m}/{s}')
    for m in (0, 1, 2):
        for s in (0, 1):
            a = O.ndpolate(q, m, s)
            for variant in (0, 1, 2):
                b = H.ndpolate(q, m, s, variant)
                assert_same_bits(a[0], b[0], f'interps {m}/{s} variant {variant}')
                assert_same_bits(a[1], b[1], f'dists {m}/{s} variant {variant}')
    st = H.stats()
    assert st['hard'] > 0 and st['ties'] > 0      # the adversarial set must reach the slow paths


@pytest.mark.parametrize('case', build_cases(seed=3), ids=lambda c: c[0])
def test_non_finite_queries(case):
    """Non-finite query coordinates.  +-inf is an ordinary out-of-bounds coordinate for the
    reference (flag 2, normed +-inf: src/ndpolator.c:35,:379-381) and must match bit for bit in
    every mode.  A NaN coordinate imports as index 1, flag 0, normed NaN (:27-39); with
    extrapolation 'none' that is fully defined behaviour (NaN interps) and must match too.  NaN
    together with an extrapolating method is outside the contract: the reference's kd traversal
    compares against NaN, and round(NaN) indexes out of bounds on associated axes (it segfaults on
    the 2basic_1assoc case here), so there is nothing to be identical to."""
    from hostsim.driver import HostSim
    name, axes, nbasic, grid = case
    q = adversarial_queries(axes, 300, np.random.default_rng(5))
    qi = q.copy()
    qi[::4, -1] = np.inf
    qi[1::4, 0] = -np.inf
    qn = q.copy()
    qn[::3, 0] = np.nan
    qn[1::7, :] = np.nan
    O, H = Oracle('port', axes, grid, nbasic), HostSim(axes, grid, nbasic)
    for a, b, what in zip(O.find(qi), H.find(qi), ('indices', 'flags', 'normed')):
        assert_same_bits(a, b, what + ' (inf)')
    for a, b, what in zip(O.find(qn), H.find(qn), ('indices', 'flags', 'normed')):
        assert_same_bits(a, b, what + ' (nan)')
    for m in (0, 1, 2):
        for s in (0, 1):
            a = O.ndpolate(qi, m, s)
            for variant in (0, 1, 2):
                b = H.ndpolate(qi, m, s, variant)
                assert_same_bits(a[0], b[0], f'interps {m}/{s} variant {variant} (inf)')
                assert_same_bits(a[1], b[1], f'dists {m}/{s} variant {variant} (inf)')
    a = O.ndpolate(qn, 0, 0)
    for variant in (0, 1, 2):
        b = H.ndpolate(qn, 0, 0, variant)
        assert_same_bits(a[0], b[0], f'interps none variant {variant} (nan)')
    # NaN rows that sit in a fully defined hypercube never reach the search: identical in every mode
    # (all-basic tables only: with associated axes the x86 oracle indexes with (int) NaN and may fault)
    if nbasic != len(axes):
        return
    fo = O.find(qn)
    _, fdhc, _ = O.hypercubes(*fo)
    rows = np.nonzero(np.asarray(fdhc).ravel() == 1)[0]
    assert np.isnan(qn[rows]).any()
    for m in (1, 2):
        for s in (0, 1):
            a, b = O.ndpolate(qn, m, s), H.ndpolate(qn, m, s, 2)
            assert_same_bits(a[0][rows], b[0][rows], f'interps {m}/{s} (nan, in a fully defined hypercube)')
            assert_same_bits(a[1][rows], b[1][rows], f'dists {m}/{s} (nan, in a fully defined hypercube)')


@pytest.mark.skipif(bool(__import__('os').environ.get('NDPB_HOSTSIM_FLAGS')), reason='already inside the variant run')
def test_margin_shortcut_variant_is_bit_exact_too():
    """ndp_search_fast's margin shortcut (-DNDP_FAST_MARGIN, off in the shipped library until it has been
    measured on a B200) must give the same bits: rerun this file against a host build with it on."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, NDPB_HOSTSIM_FLAGS='-DNDP_FAST_MARGIN')
    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-x', '-q', '-p', 'no:cacheprovider'],
                         env=env, cwd=os.path.dirname(here), capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]


def _far_queries(axes, n, rng):
    """Points far outside the box: beyond the nearest-site map's outermost band (32 cells) and beyond
    its total-distance guard (2^20), mixed with ordinary ones."""
    lo = np.array([a[0] for a in axes]); hi = np.array([a[-1] for a in axes])
    step = (hi - lo) / np.array([len(a) - 1 for a in axes])
    q = lo + (hi - lo) * rng.random((n, len(axes)))
    k = rng.integers(0, len(axes), n)
    far = rng.choice([3.0, 17.0, 31.5, 33.0, 70.0, 2000.0, 1e7], n) * rng.choice([-1.0, 1.0], n)
    q[np.arange(n), k] = np.where(far > 0, hi[k], lo[k]) + far * step[k]
    return q


@pytest.mark.parametrize('case', build_cases(seed=3), ids=lambda c: c[0])
def test_fused_plan_matches_oracle(case):
    """The fused path's per-query logic (ndp_fused.cuh: hypercube-mask slice + nearest-site map, one cell
    gather; what it cannot settle goes through the generic path) against the oracle, bit for bit, all
    methods and both search algorithms -- on the tables that qualify for it (scalar, 2..6 axes, mask bit
    decides 'fully defined')."""
    from hostsim.driver import HostSim
    name, axes, nbasic, grid = case
    rng = np.random.default_rng(31)
    q = np.concatenate([adversarial_queries(axes, 1500, rng), _far_queries(axes, 600, rng)])
    O, H = Oracle('port', axes, grid, nbasic), HostSim(axes, grid, nbasic)
    qualifies = grid.shape[-1] == 1 and 2 <= len(axes) <= 6 and name != 'assoc_dependent_nan_v1'
    for m in (0, 1, 2):
        for s in (0, 1):
            b = H.ndpolate_fused(q, m, s)
            assert (b is not None) == qualifies, f'{name}: fused path eligibility'
            if b is None:
                continue
            a = O.ndpolate(q, m, s)
            assert_same_bits(a[0], b[0], f'interps {m}/{s}')
            assert_same_bits(a[1], b[1], f'dists {m}/{s}')
    if qualifies:
        st = H.stats()
        assert st['mapped'] > 0 and st['slow'] > 0      # both the map and the exact fall-back were exercised


@pytest.mark.parametrize('spec', [('C4', 12, 'near'), ('C4', 12, 'deep'), ('C5', 6, 'near'), ('C2', 12, 'near'), ('C1', None, 'near')],
                         ids=lambda s: f'{s[0]}@{s[1]}-{s[2]}')
def test_fused_plan_on_baseline_mixes(spec):
    """The same on scaled BASELINE.json configurations with their own query mixes: the map must settle
    practically everything there (that is the point of it), still bit for bit."""
    from hostsim.driver import HostSim
    from ndpolator_b200 import synth
    name, scale, mix = spec
    n = 60000
    w = synth.workload(name, scale, mix=mix, n_queries=n)
    grid, q = w.grid(), w.queries(0, n)
    nb = len(w.basic_axes)
    O, H = Oracle('port', w.axes, grid, nb), HostSim(w.axes, grid, nb)
    for m in (0, 1, 2):
        for s in (0, 1):
            qq = q if s == 0 else q[:1500]
            a, b = O.ndpolate(qq, m, s, threads=4), H.ndpolate_fused(qq, m, s)
            assert b is not None
            assert_same_bits(a[0], b[0], f'interps {m}/{s}')
            assert_same_bits(a[1], b[1], f'dists {m}/{s}')
    st = H.stats()
    assert st['slow'] < 1e-3 * st['mapped']


def test_site_map_margins_next_to_bisectors():
    """Queries a hair (2^-18 .. 2^-44, and exactly 0) away from the bisector between two lattice points on the
    mask-invariant axes: the map settles them while the margin exceeds what rounding can bridge (2^-20 for any
    total, 2^-32 for totals below 2^8), the exact lattice search takes the rest -- the oracle's bits either way."""
    from hostsim.driver import HostSim
    from ndpolator_b200 import synth
    from cases import bisector_queries
    w = synth.workload('C4', 12, mix='near', n_queries=1)
    grid = w.grid()
    q = bisector_queries(w, 24000)
    O, H = Oracle('port', w.axes, grid, 5), HostSim(w.axes, grid, 5)
    for m in (1, 2):
        for s in (0, 1):
            qq = q if s == 0 else q[:3000]
            a, b = O.ndpolate(qq, m, s, threads=4), H.ndpolate_fused(qq, m, s)
            assert b is not None
            assert_same_bits(a[0], b[0], f'interps {m}/{s}')
            assert_same_bits(a[1], b[1], f'dists {m}/{s}')
    st = H.stats()
    assert st['mapped'] > 0 and st['slow'] > 0


def _absorbed_tie_case():
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fuzz_absorbed_tie_case.npz'))
    return [z['ax0'], z['ax1'], z['ax2']], int(z['nbasic']), z['grid'], z['q']


def test_absorbed_tie_beside_a_closed_form_axis():
    """Found by tools/fuzz_hostsim.py: a query within 1e-8 of a vertex on two basic axes and far outside on the
    associated one -- the tiny basic-axis terms are absorbed by the associated-axis term, so cells on BOTH sides of the
    mask-invariant axis tie with the cell the query sits in; the full scan takes the lowest id among them, which lies
    across that axis.  The closed-form treatment of the invariant axis must give way to the full lattice then."""
    from hostsim.driver import HostSim
    axes, nbasic, grid, q = _absorbed_tie_case()
    O, H = Oracle('port', axes, grid, nbasic), HostSim(axes, grid, nbasic)
    from oracle.oracle import have
    R = Oracle('reference', axes, grid, nbasic) if have('reference') else None
    fo = O.find(q)
    for m in (1, 2):
        for s in (0, 1):
            a, b, f = O.ndpolate(q, m, s), H.ndpolate(q, m, s, 0), H.ndpolate_fused(q, m, s)
            if R is not None:                                # what is expected is what the reference itself returns
                r = R.ndpolate(q, m, s)
                assert_same_bits(r[0], a[0], f'port vs reference interps {m}/{s}')
                assert_same_bits(r[1], a[1], f'port vs reference dists {m}/{s}')
            assert_same_bits(a[0], b[0], f'interps {m}/{s}')
            assert_same_bits(a[1], b[1], f'dists {m}/{s}')
            if f is not None:
                assert_same_bits(a[0], f[0], f'fused interps {m}/{s}')
                assert_same_bits(a[1], f[1], f'fused dists {m}/{s}')
            a, b = O.find_nearest(*fo, m, s), H.find_nearest(q, m, s)
            assert_same_bits(a[0], b[0], f'coords {m}/{s}')
            assert_same_bits(a[1], b[1], f'ndist {m}/{s}')


def test_fused_plan_with_the_finest_site_map(monkeypatch):
    """16 sub-cells per unit cell (the library's default resolution of the nearest-site maps; the rest of this
    suite builds them at 4 to keep the host-side build short): same bits."""
    from hostsim.driver import HostSim
    from ndpolator_b200 import synth
    monkeypatch.setenv('NDPB_HOSTSIM_NSM_S', '16')
    w = synth.workload('C4', 10, mix='near', n_queries=40000)
    grid, q = w.grid(), w.queries(0, 40000)
    O, H = Oracle('port', w.axes, grid, 5), HostSim(w.axes, grid, 5)
    assert H.nsm_info()[1]['S'] == 16
    for m in (1, 2):
        a, b = O.ndpolate(q, m, 0, threads=4), H.ndpolate_fused(q, m, 0)
        assert_same_bits(a[0], b[0], f'interps {m}')
        assert_same_bits(a[1], b[1], f'dists {m}')
    st = H.stats()
    assert st['slow'] < 1e-3 * st['mapped']


@pytest.mark.parametrize('case', [c for c in build_cases(seed=3) if c[0] in ('holed3d', '4d', '5d_corner_void', '2basic_2assoc_v1')],
                         ids=lambda c: c[0])
def test_blocked_layout_gives_the_same_cells(case):
    """Every block order of the blocked grid copy (2^K corners per contiguous block, K = 1 .. naxes; K = naxes is the
    cell-major copy) assembles the same cell image, hence the same bits as the oracle."""
    from hostsim.driver import HostSim
    name, axes, nbasic, grid = case
    q = adversarial_queries(axes, 1200, np.random.default_rng(41))
    O, H = Oracle('port', axes, grid, nbasic), HostSim(axes, grid, nbasic)
    want = O.ndpolate(q, 2, 0)
    for order in range(1, len(axes) + 1):
        assert H.set_block_order(order) == 1
        got = H.ndpolate_fused(q, 2, 0)
        assert_same_bits(want[0], got[0], f'interps, block order {order}')
        assert_same_bits(want[1], got[1], f'dists, block order {order}')
