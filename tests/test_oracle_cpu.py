"""CPU: the oracle restatement is pinned (a) bit-for-bit to golden vectors that
were produced by the reference itself (tests/golden/make_golden.py) and (b), when
oracle/_ref/libndp_ref.so is present, bit-for-bit to the live reference on fresh
random + adversarial inputs."""
import glob
import os

import numpy as np
import pytest

from cases import adversarial_queries, assert_same_bits, build_cases
from oracle.oracle import Oracle, have

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'ref_*.npz')))


def load_golden(path):
    z = np.load(path)
    axes = []
    while f'axis{len(axes)}' in z:
        axes.append(z[f'axis{len(axes)}'])
    return z, tuple(axes), int(z['nbasic'])


def check_engine_against_golden(z, find, hypercubes, ndpolate, find_nearest, distance, masks):
    """Shared by the CPU and GPU suites: every recorded reference output, bit for bit."""
    q = z['q']
    vm, hm = masks()
    assert_same_bits(vm, z['vmask'], 'vmask')
    assert_same_bits(hm, z['hcmask'], 'hcmask')
    idx, flg, nrm = find(q)
    assert_same_bits(idx, z['indices'], 'indices')
    assert_same_bits(flg, z['flags'], 'flags')
    assert_same_bits(nrm, z['normed'], 'normed')
    dim, fdhc, v = hypercubes(z['indices'], z['flags'], z['normed'])
    assert_same_bits(dim, z['hc_dim'], 'hypercube dim')
    assert_same_bits(fdhc, z['hc_fdhc'], 'hypercube fdhc')
    if 'hc_v' in z:
        ref = z['hc_v']
        vdim = z['grid'].shape[-1]
        for i in range(len(q)):       # only the first 2^dim corner vectors of a slot are defined
            m = (1 << int(dim[i])) * vdim
            assert_same_bits(v[i, :m], ref[i, :m], f'hypercube values of query {i}')
    for m in (0, 1, 2):
        for s in (0, 1):
            interps, dists = ndpolate(q, m, s)
            assert_same_bits(interps, z[f'interps_m{m}_s{s}'], f'interps method={m} search={s}')
            assert_same_bits(np.asarray(dists).reshape(-1, 1), z[f'dists_m{m}_s{s}'], f'dists method={m} search={s}')
            if m:
                coords, dist = find_nearest(q, z['indices'], z['flags'], z['normed'], m, s)
                assert_same_bits(coords, z[f'coords_m{m}_s{s}'], f'nearest coords method={m} search={s}')
                assert_same_bits(dist, z[f'ndist_m{m}_s{s}'], f'nearest dist method={m} search={s}')
    assert_same_bits(np.asarray(distance(q)).reshape(-1, 1), z['distance'], 'distance()')


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 10


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[4:-4] for p in GOLDEN])
def test_port_matches_reference_golden(path):
    z, axes, nbasic = load_golden(path)
    O = Oracle('port', axes, z['grid'], nbasic)
    check_engine_against_golden(
        z, O.find, O.hypercubes, O.ndpolate,
        lambda q, idx, flg, nrm, m, s: O.find_nearest(idx, flg, nrm, m, s),
        O.distance, O.masks)


@pytest.mark.skipif(not have('reference'), reason='oracle/_ref/libndp_ref.so not built (needs /root/reference)')
@pytest.mark.parametrize('case', build_cases(seed=11), ids=lambda c: c[0])
def test_port_matches_live_reference(case):
    name, axes, nbasic, grid = case
    rng = np.random.default_rng(77)
    q = adversarial_queries(axes, 800, rng)
    P, R = Oracle('port', axes, grid, nbasic), Oracle('reference', axes, grid, nbasic)
    for a, b, what in zip(P.masks(), R.masks(), ('vmask', 'hcmask')):
        assert_same_bits(a, b, what)
    fp, fr = P.find(q), R.find(q)
    for a, b, what in zip(fp, fr, ('indices', 'flags', 'normed')):
        assert_same_bits(a, b, what)
    hp, hr = P.hypercubes(*fp), R.hypercubes(*fr)
    assert_same_bits(hp[0], hr[0], 'dim')
    assert_same_bits(hp[1], hr[1], 'fdhc')
    for m in (0, 1, 2):
        for s in (0, 1):
            a, b = P.ndpolate(q, m, s), R.ndpolate(q, m, s)
            assert_same_bits(a[0], b[0], f'interps {m}/{s}')
            assert_same_bits(a[1], b[1], f'dists {m}/{s}')
            if m:
                a, b = P.find_nearest(*fp, m, s), R.find_nearest(*fr, m, s)
                assert_same_bits(a[0], b[0], f'coords {m}/{s}')
                assert_same_bits(a[1], b[1], f'ndist {m}/{s}')
    assert_same_bits(P.distance(q), R.distance(q), 'distance')


@pytest.mark.skipif(not have('reference'), reason='oracle/_ref/libndp_ref.so not built (needs /root/reference)')
@pytest.mark.parametrize('spec', [('C4', 14, 'iid'), ('C3', 24, 'allbasic'), ('C2', 14, None), ('C5', 6, None)],
                         ids=lambda s: f'{s[0]}@{s[1]}-{s[2]}')
def test_port_matches_live_reference_on_survey_variants(spec):
    """The oracle port against the live reference on the SURVEY 8d workloads and variants the GPU tests check against
    the port: C4 with i.i.d. voids (sparse hypercube mask; a scale where it is not empty), C3 all-basic, scaled C2 / C5."""
    from ndpolator_b200 import synth
    name, scale, variant = spec
    n = 3000
    w = synth.workload(name, scale, n_queries=n, variant=variant)
    grid, q = w.grid(), w.queries(0, n)
    nb = len(w.basic_axes)
    P, R = Oracle('port', w.axes, grid, nb), Oracle('reference', w.axes, grid, nb)
    for a, b, what in zip(P.masks(), R.masks(), ('vmask', 'hcmask')):
        assert_same_bits(a, b, what)
    empty_hc = not R.masks()[1].any()          # the reference crashes on an empty kd-tree (kd_res_free(NULL)): skip 'linear' there
    for m in (0, 1) + (() if empty_hc else (2,)):
        a, b = P.ndpolate(q, m, 0), R.ndpolate(q, m, 0)
        assert_same_bits(a[0], b[0], f'{w.name} interps {m}')
        assert_same_bits(a[1], b[1], f'{w.name} dists {m}')


def test_threaded_oracle_equals_serial():
    name, axes, nbasic, grid = build_cases()[1]
    q = adversarial_queries(axes, 4000, np.random.default_rng(5))
    O = Oracle('port', axes, grid, nbasic)
    a = O.ndpolate(q, 2, 0, threads=1)
    b = O.ndpolate(q, 2, 0, threads=4)
    assert_same_bits(a[0], b[0])
    assert_same_bits(a[1], b[1])


def test_port_matches_reference_golden_fullsize_c3():
    """C3 at BASELINE.json's full size (80x20x10x50, vdim 16): the port against outputs of the reference
    itself (the C4 / C5 full-size fixtures are checked on the GPU only: the port's insertion-built tree for
    them takes minutes, like the reference's)."""
    import fullsize
    path = [p for p in fullsize.FULL if p.endswith('full_c3.npz')]
    assert path, 'tests/golden/full_c3.npz missing'
    z, w, first, count = fullsize.load_full(path[0])
    q = w.queries(first, count)
    O = Oracle('port', w.axes, w.grid(), len(w.basic_axes))
    idx, flg, nrm = O.find(q)
    fullsize.check_full(z, lambda m: O.ndpolate(q, m, 0, threads=os.cpu_count() or 1),
                        lambda m: O.find_nearest(idx, flg, nrm, m, 0))


@pytest.mark.skipif(not have('reference'), reason='oracle/_ref/libndp_ref.so not built (needs /root/reference)')
@pytest.mark.parametrize('case', build_cases(seed=11), ids=lambda c: c[0])
def test_port_tree_shape_is_the_references(case):
    """parent / depth of the port's index-linked kd-tree against a walk over the reference's own
    struct kdnode (oracle/ref_shim.c: ndpref_tree_topology)."""
    name, axes, nbasic, grid = case
    P, R = Oracle('port', axes, grid, nbasic), Oracle('reference', axes, grid, nbasic)
    for m in (1, 2):
        pp, dp = P.tree_topology(m)
        pr, dr = R.tree_topology(m)
        assert np.array_equal(pp, pr), f'kd parent, method {m}'
        assert np.array_equal(dp, dr), f'kd depth, method {m}'
