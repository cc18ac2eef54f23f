"""GPU: parity of the CUDA path, called through the C ABI (ndpolator_b200.cndpolator.Table
-> libndpb200.so), against (a) golden vectors produced by the reference itself and
(b) the CPU oracle on fresh seeded inputs.  Bar (BASELINE.json north_star): indices,
flags, hypercube corner ids / fdhc / dim, nearest coords (ties included) and NaN
placement bit-exact; interps, dists, normed within 1e-13 relative -- the tests demand
bit-exactness for those too (compiled with -fmad=false), which implies the 1e-13 bar."""
import os

import numpy as np
import pytest

from cases import adversarial_queries, assert_same_bits, build_cases
from oracle.oracle import Oracle
from fullsize import FULL, check_full, load_full
from test_oracle_cpu import GOLDEN, check_engine_against_golden, load_golden

pytestmark = pytest.mark.gpu


def _table(axes, grid, nbasic, **opts):
    from ndpolator_b200.cndpolator import Table
    t = Table(axes, grid, nbasic)
    for k, v in opts.items():
        t.set_option(k, v)
    return t


def test_extension_is_loaded_and_gpu_present():
    import torch
    assert torch.cuda.is_available()
    from ndpolator_b200 import _lib
    assert b'sm_100a' in _lib.lib().ndpb_version()


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[4:-4] for p in GOLDEN])
@pytest.mark.parametrize('variant', [(1, 0, 0), (2, 0, 0), (2, 0, 1), (2, 1, 1)],
                         ids=['strided', 'cellmajor-fused', 'cellmajor-staged', 'cellmajor-direct'])
def test_cuda_matches_reference_golden(path, variant):
    gather, staged_off, fused_off = variant
    z, axes, nbasic = load_golden(path)
    T = _table(axes, z['grid'], nbasic, gather=gather, staged=staged_off, fused=fused_off)
    check_engine_against_golden(
        z, T.find, lambda idx, flg, nrm: T.hypercubes_raw(nrm, idx, flg), T.ndpolate,
        lambda q, idx, flg, nrm, m, s: T.find_nearest(q, m, s), T.distance, T.masks)
    if gather == 2 and len(axes) <= 6 and z['grid'].shape[-1] * 8 < 128:
        assert T.stat('cellmajor') == 1
    scalar = z['grid'].shape[-1] == 1 and 2 <= len(axes) <= 6
    vector = z['grid'].shape[-1] > 1 and 2 <= len(axes) <= 6
    if not fused_off and ((gather == 2 and scalar) or vector):
        # the last call was distance(); run one ndpolate to read which path served it
        T.ndpolate(z['q'], 2, 0)
        assert T.stat('fused') == (0 if 'assoc_dependent_nan' in path else 1)
        assert T.stat('use_hc') == (0 if 'assoc_dependent_nan' in path else 1)
    T.close()


@pytest.mark.parametrize('path', FULL, ids=[os.path.basename(p)[5:-4] for p in FULL])
def test_cuda_matches_reference_golden_fullsize(path):
    """BASELINE.json's configurations AT FULL GRID SIZE (C4 40^5 with both query mixes, C5 16^6, C3
    80x20x10x50 v16) against outputs of the reference itself on sub-ranges of the very batches bench.py
    times (tests/golden/make_golden_fullsize.py; the reference needed ~20 min of kd-tree building for C4)."""
    import torch
    z, w, first, count = load_full(path)
    dev = 'cuda'
    from ndpolator_b200.cndpolator import Table
    T = Table(w.axes, w.grid(dev), len(w.basic_axes))
    q = w.queries(first, count, dev)

    def ndp(m):
        gi, gd = T.ndpolate(q, m, 0)
        return gi.cpu().numpy(), gd.cpu().numpy()

    def near(m):
        c, d = T.find_nearest(q, m, 0)
        return c.cpu().numpy(), d.cpu().numpy()

    check_full(z, ndp, near)
    if w.vdim == 1:
        ndp(2)
        assert T.stat('fused') == 1 and T.stat('slow') < 100      # the benchmarked path is the one checked
        T.set_option('fused', 1)                                   # and the staged / two-pass path as well
        check_full(z, ndp, None, methods=(1, 2))
    # host buffers through the chunked H2D / D2H pipeline give the same bits
    T.set_option('fused', 0)
    qh = q.cpu().numpy()
    check_full(z, lambda m: T.ndpolate(qh, m, 0), None, methods=(2,))
    T.close()


def test_device_inputs_are_ordered_after_their_producer():
    """A query tensor produced on torch's current stream immediately before the call (here behind a long
    fill kernel) must be complete when the table's own stream reads it (ndpb_table_wait_stream)."""
    import torch
    from ndpolator_b200 import synth
    w = synth.workload('C4', 12)
    T = _table(w.axes, w.grid('cuda'), 5)
    n = 1 << 20
    q_ready = w.queries(0, n, 'cuda')
    want = T.ndpolate(q_ready, 2, 0)
    want = (want[0].clone(), want[1].clone())
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(side):
            junk = torch.empty(1 << 28, dtype=torch.float64, device='cuda')
            junk.fill_(1.0)                                   # ~2 GB of writes ahead of the producer on this stream
            q = torch.full((n, 5), float('nan'), dtype=torch.float64, device='cuda')
            q.copy_(q_ready)                                   # the producer: still queued when ndpolate is called
            got = T.ndpolate(q, 2, 0)
        assert torch.equal(got[0].view(torch.int64), want[0].view(torch.int64))
        assert torch.equal(got[1].view(torch.int64), want[1].view(torch.int64))
        del junk
    T.close()


def test_device_grid_registration_is_ordered():
    """Registering a table from a DEVICE-resident grid while other work is still queued on the producer's stream: the
    library's copy of the grid must be complete before its (non-blocking-stream) kernels derive the masks from it.
    (A plain cudaMemcpy of device memory returns early and is ordered on the legacy default stream only: masks were
    computed from whatever the recycled pages held -- one flaky suite run in 17 before the fix.)"""
    import torch
    from ndpolator_b200 import synth
    w = synth.workload('C3', None)
    nb = len(w.basic_axes)
    vm, hm = Oracle('port', w.axes, w.grid(), nb).masks()
    grid_d = w.grid('cuda')
    torch.cuda.synchronize()
    for it in range(6):
        ones = torch.ones_like(grid_d)                  # pages full of finite values go back to the driver ...
        del ones
        torch.cuda.empty_cache()
        pending = [torch.full((1 << 26,), float('nan'), dtype=torch.float64, device='cuda') for _ in range(4)]   # ... 2 GB of writes queued
        T = _table(w.axes, grid_d, nb)
        a, b = T.masks()
        assert_same_bits(vm, a, f'vmask, iteration {it}')
        assert_same_bits(hm, b, f'hcmask, iteration {it}')
        T.close()
        del pending


def test_bad_device_arguments_raise():
    import torch
    name, axes, nbasic, grid = build_cases()[1]
    T = _table(axes, grid, nbasic)
    q = torch.zeros((8, 3), dtype=torch.float64, device='cuda')
    with pytest.raises(ValueError):
        T.ndpolate(q, 0, 0, out=(torch.zeros((8, 1), dtype=torch.float32, device='cuda'), torch.zeros((8, 1), dtype=torch.float64, device='cuda')))
    with pytest.raises(ValueError):
        T.ndpolate(q, 0, 0, out=(np.zeros((8, 1)), np.zeros((8, 1))))
    with pytest.raises(ValueError):
        T.ndpolate(q, 0, 0, out=(torch.zeros((7, 1), dtype=torch.float64, device='cuda'), torch.zeros((8, 1), dtype=torch.float64, device='cuda')))
    T.close()


@pytest.mark.parametrize('case', build_cases(seed=21), ids=lambda c: c[0])
def test_cuda_matches_oracle_fresh_inputs(case):
    name, axes, nbasic, grid = case
    q = adversarial_queries(axes, 20000, np.random.default_rng(123))
    O = Oracle('port', axes, grid, nbasic)
    T = _table(axes, grid, nbasic)
    for a, b, what in zip(O.masks(), T.masks(), ('vmask', 'hcmask')):
        assert_same_bits(a, b, what)
    fo = O.find(q)
    for a, b, what in zip(fo, T.find(q), ('indices', 'flags', 'normed')):
        assert_same_bits(a, b, what)
    dim_o, fdhc_o, _ = O.hypercubes(*fo)
    dim_t, fdhc_t, _ = T.hypercubes_raw(fo[2], fo[0], fo[1])
    assert_same_bits(dim_o, dim_t, 'dim')
    assert_same_bits(fdhc_o, fdhc_t, 'fdhc')
    for m in (1, 2):
        po, do = O.tree_topology(m)
        pt, dt = T.tree_topology(m)
        assert np.array_equal(po, pt.astype(np.int64)), f'kd parent, method {m}'
        assert np.array_equal(do, dt), f'kd depth, method {m}'
    ties = 0
    for m in (0, 1, 2):
        for s in (0, 1):
            a, b = O.ndpolate(q, m, s), T.ndpolate(q, m, s)
            assert_same_bits(a[0], b[0], f'interps {m}/{s}')
            assert_same_bits(a[1], b[1], f'dists {m}/{s}')
            ties += T.stat('ties')
            if m:
                a, b = O.find_nearest(*fo, m, s), T.find_nearest(q, m, s)
                assert_same_bits(a[0], b[0], f'coords {m}/{s}')
                assert_same_bits(a[1], b[1], f'ndist {m}/{s}')
    assert_same_bits(O.distance(q), T.distance(q), 'distance')
    assert T.stat('launches') > 0
    T.close()


def test_tie_pass_runs_before_topology_exists():
    """Exact mid-point queries on a fresh table: the first call meets ties with no kd
    topology yet, builds it, and resolves them like the reference's traversal."""
    name, axes, nbasic, grid = build_cases()[1]
    rng = np.random.default_rng(3)
    mids = [(a[1:] + a[:-1]) / 2 for a in axes]
    q = np.stack([rng.choice(m, 5000) for m in mids], axis=1) + np.array([7.0, 0, 0])    # outside along axis 0
    O = Oracle('port', axes, grid, nbasic)
    T = _table(axes, grid, nbasic)
    a, b = O.ndpolate(q, 1, 0), T.ndpolate(q, 1, 0)
    assert T.stat('ties') > 0
    assert_same_bits(a[0], b[0], 'interps')
    assert_same_bits(a[1], b[1], 'dists')
    a, b = O.ndpolate(q, 2, 0), T.ndpolate(q, 2, 0)
    assert_same_bits(a[0], b[0], 'interps linear')
    T.close()


@pytest.mark.parametrize('fused_off', [0, 1], ids=['fused', 'two-pass'])
def test_stream_ordered_calls_async_option(fused_off):
    """set_option('async', 1): calls on device tensors only enqueue (ndpb_table_sync completes them); exact kd
    distance ties are settled inside the kernels because the topology is built ahead of the first such call.
    Several calls in flight on the caller's stream, distinct outputs, identical bits to the oracle."""
    import torch
    name, axes, nbasic, grid = build_cases()[1]
    rng = np.random.default_rng(5)
    mids = [(a[1:] + a[:-1]) / 2 for a in axes]
    ties = np.stack([rng.choice(m, 3000) for m in mids], axis=1) + np.array([7.0, 0, 0])
    q = np.concatenate([adversarial_queries(axes, 20000, np.random.default_rng(17)), ties])
    O = Oracle('port', axes, grid, nbasic)
    T = _table(axes, grid, nbasic, fused=fused_off)
    stream = torch.cuda.Stream()
    T.set_option('stream', stream.cuda_stream)
    T.set_option('async', 1)
    with torch.cuda.stream(stream):
        qd = torch.as_tensor(q, device='cuda')
        outs = {}
        for rep in range(2):
            for m in (0, 1, 2):
                for sr in (0, 1):
                    outs[(rep, m, sr)] = T.ndpolate(qd, m, sr)          # returns at once; nothing is read back here
        near = {(m, sr): T.find_nearest(qd, m, sr) for m in (1, 2) for sr in (0, 1)}
        dist = T.distance(qd)
    T.sync()
    assert T.stat('launches') > 0
    fo = O.find(q)
    for (rep, m, sr), (gi, gd) in outs.items():
        want = O.ndpolate(q, m, sr)
        assert_same_bits(want[0], gi.cpu().numpy(), f'async interps {m}/{sr} rep {rep}')
        assert_same_bits(want[1], gd.cpu().numpy(), f'async dists {m}/{sr} rep {rep}')
    for (m, sr), (gc, gd) in near.items():
        want = O.find_nearest(*fo, m, sr)
        assert_same_bits(want[0], gc.cpu().numpy(), f'async coords {m}/{sr}')
        assert_same_bits(want[1], gd.cpu().numpy(), f'async ndist {m}/{sr}')
    assert_same_bits(O.distance(q), dist.cpu().numpy(), 'async distance')
    # host buffers stay synchronous in async mode
    hi, hd = T.ndpolate(q, 2, 0)
    assert_same_bits(O.ndpolate(q, 2, 0)[0], hi, 'host call in async mode')
    T.set_option('async', 0)
    T.set_option('stream', 0)
    T.close()


def test_site_map_margins_next_to_bisectors():
    """GPU twin of the host-build test of the same name: queries 0 .. 2^-18 away from a bisector on the
    mask-invariant axes (the map settles them down to a 2^-32 margin for totals below 2^8, the exact search the rest)."""
    from cases import bisector_queries
    from ndpolator_b200 import synth
    w = synth.workload('C4', 12, mix='near', n_queries=1)
    grid = w.grid()
    q = bisector_queries(w, 60000)
    O = Oracle('port', w.axes, grid, 5)
    T = _table(w.axes, grid, 5)
    for m in (1, 2):
        for s in (0, 1):
            qq = q if s == 0 else q[:3000]
            a, b = O.ndpolate(qq, m, s, threads=os.cpu_count() or 1), T.ndpolate(qq, m, s)
            assert T.stat('fused') == 1 and T.stat('slow') > 0
            assert_same_bits(a[0], b[0], f'interps {m}/{s}')
            assert_same_bits(a[1], b[1], f'dists {m}/{s}')
    T.close()


def test_absorbed_tie_beside_a_closed_form_axis():
    """GPU twin of the host-build test of the same name (a case tools/fuzz_hostsim.py found): exact ties created by
    absorbed terms across a mask-invariant axis, every path -- fused + k_slow, three-pass, find_nearest."""
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fuzz_absorbed_tie_case.npz'))
    axes, nbasic, grid, q = [z['ax0'], z['ax1'], z['ax2']], int(z['nbasic']), z['grid'], z['q']
    O = Oracle('port', axes, grid, nbasic)
    fo = O.find(q)
    for fused_off in (0, 1):
        T = _table(axes, grid, nbasic, fused=fused_off)
        for m in (1, 2):
            for s in (0, 1):
                a, b = O.ndpolate(q, m, s), T.ndpolate(q, m, s)
                assert_same_bits(a[0], b[0], f'interps {m}/{s} fused_off={fused_off}')
                assert_same_bits(a[1], b[1], f'dists {m}/{s} fused_off={fused_off}')
                a, b = O.find_nearest(*fo, m, s), T.find_nearest(q, m, s)
                assert_same_bits(a[0], b[0], f'coords {m}/{s}')
                assert_same_bits(a[1], b[1], f'ndist {m}/{s}')
        T.close()


def test_host_chunking_device_tensors_and_search_variants_agree():
    import torch
    name, axes, nbasic, grid = build_cases()[7]          # 5-D with a structured void
    q = adversarial_queries(axes, 30000, np.random.default_rng(8))
    O = Oracle('port', axes, grid, nbasic)
    want = {(m, s): O.ndpolate(q, m, s) for m in (0, 1, 2) for s in (0, 1)}
    T = _table(axes, grid, nbasic, chunk=4096)           # many chunks through the 2-slot pipeline
    for (m, s), w in want.items():
        got = T.ndpolate(q, m, s)
        assert_same_bits(w[0], got[0], f'chunked interps {m}/{s}')
        assert_same_bits(w[1], got[1], f'chunked dists {m}/{s}')
    assert T.stat('h2d_bytes') == q.nbytes
    qd = torch.as_tensor(q, device='cuda')
    for (m, s), w in want.items():
        gi, gd = T.ndpolate(qd, m, s)
        assert gi.is_cuda and gd.is_cuda
        assert_same_bits(w[0], gi.cpu().numpy(), f'device interps {m}/{s}')
        assert_same_bits(w[1], gd.cpu().numpy(), f'device dists {m}/{s}')
    assert T.stat('h2d_bytes') == 0
    for impl in (1, 2):                                  # literal warp scan vs lattice descent
        T.set_option('linear_search', impl)
        for m in (1, 2):
            got = T.ndpolate(q, m, 1)
            assert_same_bits(want[(m, 1)][0], got[0], f'linear_search={impl} interps {m}')
            assert_same_bits(want[(m, 1)][1], got[1], f'linear_search={impl} dists {m}')
    T.close()


@pytest.mark.parametrize('fused_off', [0, 1], ids=['fused', 'two-pass'])
@pytest.mark.parametrize('spec', [('C1', None, 200000), ('C2', 50, 200000), ('C3', 24, 60000),
                                  ('C4', 12, 100000), ('C4deep', 12, 100000), ('C5', 6, 100000)], ids=lambda s: f'{s[0]}@{s[1]}')
def test_baseline_configs_scaled_vs_oracle(spec, fused_off):
    """The BASELINE.json configurations (C1/C2 at full grid size, C3-C5 scaled so that the
    oracle's O(N*depth) tree build finishes in seconds), every extrapolation method."""
    from ndpolator_b200 import synth
    name, scale, n = spec
    w = synth.workload(name[:2], scale, mix='deep' if name.endswith('deep') else 'near', n_queries=n)
    grid, q = w.grid(), w.queries(0, n)
    nb = len(w.basic_axes)
    O = Oracle('port', w.axes, grid, nb)
    T = _table(w.axes, grid, nb, fused=fused_off)
    for m in (0, 1, 2):
        a, b = O.ndpolate(q, m, 0, threads=os.cpu_count() or 1), T.ndpolate(q, m, 0)
        assert_same_bits(a[0], b[0], f'{name} interps method {m}')
        assert_same_bits(a[1], b[1], f'{name} dists method {m}')
    sub = q[:300]
    for m in (1, 2):
        a, b = O.ndpolate(sub, m, 1), T.ndpolate(sub, m, 1)
        assert_same_bits(a[0], b[0], f'{name} interps method {m}, search linear')
        assert_same_bits(a[1], b[1], f'{name} dists method {m}, search linear')
    T.close()


@pytest.mark.parametrize('spec', [('C4', 14, 'iid', 40000), ('C3', 24, 'allbasic', 40000)], ids=lambda s: f'{s[0]}@{s[1]}-{s[2]}')
def test_survey_variants_vs_oracle(spec):
    """SURVEY 8d variants: C4 with i.i.d. 30 % voids (the mask depends on all five axes -- no nearest-site map, the
    three-pass path with the exact lattice search; ~100 % of the queries extrapolate, the hypercube mask is nearly
    empty) and C3 with mu as a fourth basic axis."""
    from ndpolator_b200 import synth
    name, scale, variant, n = spec
    w = synth.workload(name, scale, n_queries=n, variant=variant)
    grid, q = w.grid(), w.queries(0, n)
    nb = len(w.basic_axes)
    O = Oracle('port', w.axes, grid, nb)
    T = _table(w.axes, grid, nb)
    for m in (0, 1, 2):
        a, b = O.ndpolate(q, m, 0, threads=os.cpu_count() or 1), T.ndpolate(q, m, 0)
        assert_same_bits(a[0], b[0], f'{w.name} interps method {m}')
        assert_same_bits(a[1], b[1], f'{w.name} dists method {m}')
    if variant == 'iid':
        assert T.stat('fused') == 0 and T.stat('variant_axes_nearest') == 5
    T.close()


@pytest.mark.parametrize('void', [0.25, 0.45], ids=['146-cells', 'almost-empty'])
def test_sparse_hypercube_mask_search(void):
    """A hypercube mask with a handful of fully defined cells (i.i.d. voids: 0.75^16 = 1 % of the 4-D cells, or almost
    none): 'linear' scores the set cells themselves (k_search_sparse) -- kd and full-scan tie rules, exact ticks and
    mid-points among the queries, topology missing on the first call and present on the second, stream-ordered too."""
    import torch
    rng = np.random.default_rng(11)
    axes = [np.arange(12, dtype=np.float64) * s + o for s, o in ((1.0, 0.0), (0.5, 2.0), (2.0, -3.0), (1.0, 5.0))]
    grid = rng.random((12, 12, 12, 12, 1))
    grid[rng.random((12, 12, 12, 12)) < void] = np.nan
    q = np.concatenate([adversarial_queries(axes, 20000, np.random.default_rng(12)),
                        np.stack([rng.choice(a, 3000) for a in axes], axis=1) + rng.choice([-0.0, 7.0], (3000, 1)) * np.array([1.0, 0, 0, 0])])
    O = Oracle('port', axes, grid, 4)
    T = _table(axes, grid, 4, fused=1)                     # the three-pass path: its search is what is under test
    assert T.stat('mask_count_linear') * 64 <= T.nverts
    for rep in range(2):                                   # rep 0 meets kd ties without the topology, rep 1 with it
        for s in (0, 1):
            qq = q if s == 0 else q[:4000]
            a, b = O.ndpolate(qq, 2, s, threads=os.cpu_count() or 1), T.ndpolate(qq, 2, s)
            assert T.stat('sparse_search_linear') == 1 and T.stat('fused') == 0
            assert_same_bits(a[0], b[0], f'interps linear/{s} rep {rep}')
            assert_same_bits(a[1], b[1], f'dists linear/{s} rep {rep}')
    T.set_option('fused', 0)                               # and the fused path on the same table (map + k_slow)
    a, b = O.ndpolate(q, 2, 0, threads=os.cpu_count() or 1), T.ndpolate(q, 2, 0)
    assert_same_bits(a[0], b[0], 'interps linear, fused'); assert_same_bits(a[1], b[1], 'dists linear, fused')
    T.set_option('fused', 1)
    fo = O.find(q)
    a, b = O.find_nearest(*fo, 2, 0), T.find_nearest(q, 2, 0)
    assert_same_bits(a[0], b[0], 'coords'); assert_same_bits(a[1], b[1], 'ndist')
    assert_same_bits(O.distance(q), T.distance(q), 'distance')
    T.set_option('async', 1)
    gi, gd = T.ndpolate(torch.as_tensor(q, device='cuda'), 2, 0)
    T.sync()
    a = O.ndpolate(q, 2, 0, threads=os.cpu_count() or 1)
    assert_same_bits(a[0], gi.cpu().numpy(), 'async interps'); assert_same_bits(a[1], gd.cpu().numpy(), 'async dists')
    T.close()


def test_c4_full_size_properties():
    """BASELINE.json's headline configuration at full size (40^5 grid, 819 MB; cell-major
    copy 23 GB): size-independent properties, plus a small sub-sample pinned to the oracle
    through the tree-free full scan (the oracle's kd-tree for this grid takes ~15 min to build)."""
    import torch
    from ndpolator_b200 import synth
    w = synth.workload('C4', None)
    dev = 'cuda'
    grid_d = w.grid(dev)
    nb = len(w.basic_axes)
    from ndpolator_b200.cndpolator import Table
    T = Table(w.axes, grid_d, nb)
    assert T.stat('cellmajor') == 1
    n = 1 << 22
    q = w.queries(0, n, dev)
    # (1) strided and cell-major gathers agree bit for bit; NaN only where no extrapolation is possible
    res = {}
    for m in (0, 1, 2):
        gi, gd = T.ndpolate(q, m, 0)
        res[m] = (gi.clone(), gd.clone())
    # the software pipeline (TMA-refilled query buffers, cp.async cell slots) is free of races: repeated runs agree
    # bit for bit -- method 'none' has the shortest plan, hence the tightest timing between a read and its refill
    for rep in range(6):
        for m in (0, 2):
            gi, gd = T.ndpolate(q, m, 0)
            assert torch.equal(gi.view(torch.int64), res[m][0].view(torch.int64)), f'run {rep} of method {m} differs'
            assert torch.equal(gd.view(torch.int64), res[m][1].view(torch.int64))
    T.set_option('gather', 1)
    for m in (0, 2):
        gi, gd = T.ndpolate(q, m, 0)
        assert torch.equal(gi.view(torch.int64), res[m][0].view(torch.int64))
        assert torch.equal(gd.view(torch.int64), res[m][1].view(torch.int64))
    T.set_option('gather', 2)
    none_nan = torch.isnan(res[0][0][:, 0])
    frac_in = 1.0 - none_nan.double().mean().item()
    assert 0.45 < frac_in < 0.65                       # ~56 % of the near mix sits in a full hypercube
    for m in (1, 2):
        assert not torch.isnan(res[m][0]).any()
        assert torch.equal(res[m][0][~none_nan], res[0][0][~none_nan])     # in-grid points do not depend on the method
        assert (res[m][1][~none_nan] == 0).all() and (res[m][1] >= 0).all()
        assert (res[m][1][none_nan] > 0).float().mean() > 0.99
    # (2) nearest: the value returned is the grid value at the coords the search reports, kd == full scan
    c_kd, d_kd = T.find_nearest(q[:200000], 1, 0)
    c_ls, d_ls = T.find_nearest(q[:200000], 1, 1)
    assert torch.equal(c_kd, c_ls) and torch.equal(d_kd.view(torch.int64), d_ls.view(torch.int64))
    flat = (c_kd.long() * torch.tensor([40 ** 4, 40 ** 3, 40 ** 2, 40, 1], device=dev)).sum(1)
    vals = grid_d.reshape(-1)[flat]
    off = none_nan[:200000]
    assert torch.equal(vals[off], res[1][0][:200000, 0][off])
    # (3) sub-sample vs the CPU oracle (full scan: no tree build)
    grid_h = grid_d.cpu().numpy()
    O = Oracle('port', w.axes, grid_h, nb)
    idx = torch.nonzero(none_nan)[:12, 0]
    sub = q[idx].cpu().numpy()
    a = O.ndpolate(sub, 1, 1)
    assert_same_bits(a[0], res[1][0][idx].cpu().numpy(), 'C4 full size, nearest vs oracle scan')
    assert_same_bits(a[1], res[1][1][idx].cpu().numpy(), 'C4 full size, nearest dists vs oracle scan')
    inside = torch.nonzero(~none_nan)[:20000, 0]
    a = O.ndpolate(q[inside].cpu().numpy(), 0, 0)
    assert_same_bits(a[0], res[0][0][inside].cpu().numpy(), 'C4 full size, in-grid vs oracle')
    T.close()


def _group_case():
    from ndpolator_b200 import synth
    w = synth.workload('C4', 12, n_queries=300000)
    return w, w.grid(), w.queries(0, 300000)


@pytest.mark.parametrize('ndev', [1, 2, 8])
def test_group_one_call_several_gpus(ndev):
    """ndpb_group: one host batch cut into contiguous ranges over `ndev` replicated tables, each device copying
    its results into its slice of the caller's arrays -- bit for bit what one table returns (and the oracle)."""
    import torch
    if torch.cuda.device_count() < ndev:
        pytest.skip(f'needs {ndev} GPUs')
    from ndpolator_b200.cndpolator import Group
    w, grid, q = _group_case()
    O = Oracle('port', w.axes, grid, 5)
    G = Group(w.axes, grid, 5, devices=list(range(ndev)))
    G.set_option('min_queries_per_device', 1000)
    for m in (0, 1, 2):
        want = O.ndpolate(q, m, 0, threads=os.cpu_count() or 1)
        got = G.ndpolate(q, m, 0)
        assert_same_bits(want[0], got[0], f'group interps method {m}')
        assert_same_bits(want[1], got[1], f'group dists method {m}')
        assert G.stat('used_devices') == ndev
    assert_same_bits(O.distance(q[:5000]), G.distance(q[:5000]), 'group distance')
    small = G.ndpolate(q[:10], 2, 0)                       # a batch too small to split runs on one device
    assert G.stat('used_devices') == 1
    assert_same_bits(small[0], O.ndpolate(q[:10], 2, 0)[0], 'small batch')
    with pytest.raises(ValueError):
        G.ndpolate(torch.zeros((4, 5), dtype=torch.float64, device='cuda'), 0, 0)
    G.close()


def test_ndpolator_class_with_device_list():
    import torch
    from ndpolator_b200 import Ndpolator
    w, grid, q = _group_case()
    devices = list(range(min(2, torch.cuda.device_count())))
    ndp = Ndpolator(w.basic_axes, devices=devices)
    ndp.register('t', None, grid)
    got = ndp.ndpolate('t', q[:50000], extrapolation_method='linear')
    O = Oracle('port', w.axes, grid, 5)
    want = O.ndpolate(q[:50000], 2, 0, threads=os.cpu_count() or 1)
    assert_same_bits(want[0], got['interps'], 'Ndpolator(devices=...) interps')
    assert_same_bits(want[1], got['dists'], 'Ndpolator(devices=...) dists')


@pytest.mark.parametrize('case', [c for c in build_cases(seed=21) if c[0] in ('holed3d', '2basic_1assoc_v2', '5d_corner_void', 'c3_like_v16')],
                         ids=lambda c: c[0])
def test_device_resident_cache_path(case):
    """import_query_pts -> find_hypercubes -> ndpolate_cached on DEVICE tensors (reference README
    'Hypercube caching', ndpolator/ndpolator.py:110-187): nothing round-trips through the host, and every product
    equals the oracle's / the uncached call's, bit for bit."""
    import torch
    from ndpolator_b200 import Ndpolator
    name, axes, nbasic, grid = case
    q = adversarial_queries(axes, 5000, np.random.default_rng(77))
    O = Oracle('port', axes, grid, nbasic)
    ndp = Ndpolator(tuple(axes[:nbasic]))
    ndp.register('t', tuple(axes[nbasic:]) or None, grid)
    qd = torch.as_tensor(q, device='cuda')
    ndp.ndpolate('t', qd)                                    # creates the device table
    idx, flg, nrm = ndp.import_query_pts('t', qd)
    assert idx.is_cuda and flg.is_cuda and nrm.is_cuda
    fo = O.find(q)
    for a, b, what in zip(fo, (idx, flg, nrm), ('indices', 'flags', 'normed')):
        assert_same_bits(a, b.cpu().numpy(), what)
    hc = ndp.find_hypercubes('t', nrm, idx, flg, tuple(axes[nbasic:]) or None)
    assert hc.values.is_cuda and len(hc) == len(q)
    dim_o, fdhc_o, v_o = O.hypercubes(*fo)
    assert_same_bits(dim_o, hc.dim.cpu().numpy(), 'dim')
    assert_same_bits(fdhc_o, hc.fdhc.cpu().numpy(), 'fdhc')
    vdim = grid.shape[-1]
    for i in (0, 17, len(q) - 1):
        m = (1 << int(dim_o[i])) * vdim
        assert_same_bits(v_o[i, :m].reshape(hc[i].shape), hc[i].cpu().numpy(), f'hypercube {i}')
    for method, m in (('none', 0), ('nearest', 1), ('linear', 2)):
        got = ndp.ndpolate_cached('t', nrm, idx, flg, extrapolation_method=method)
        want = O.ndpolate(q, m, 0)
        if name in ('holed3d', '5d_corner_void', 'c3_like_v16'):       # served by k_fused / k_fused_vec on the cached products
            assert ndp.table['t']['capsule'].stat('fused') == 1
        assert got['interps'].is_cuda
        assert_same_bits(want[0], got['interps'].cpu().numpy(), f'cached interps {method}')
        if m:
            assert_same_bits(want[1], got['dists'].cpu().numpy(), f'cached dists {method}')
    # host cache gives host results
    got = ndp.ndpolate_cached('t', fo[2], fo[0], fo[1], extrapolation_method='linear')
    assert isinstance(got['interps'], np.ndarray)
    assert_same_bits(O.ndpolate(q, 2, 0)[0], got['interps'], 'cached interps from host arrays')


@pytest.mark.parametrize('case', [c for c in build_cases(seed=21) if c[0] in ('holed3d', '4d', '5d_corner_void', '2basic_2assoc_v1')],
                         ids=lambda c: c[0])
def test_blocked_copy_every_order(case):
    """The blocked grid copy k_fused gathers from (2^K corners per contiguous block; K = naxes is the cell-major
    copy, smaller K serves tables whose full copy would not fit): every order gives the oracle's bits."""
    name, axes, nbasic, grid = case
    q = adversarial_queries(axes, 20000, np.random.default_rng(99))
    O = Oracle('port', axes, grid, nbasic)
    want = {m: O.ndpolate(q, m, 0) for m in (0, 1, 2)}
    T = _table(axes, grid, nbasic)
    for order in range(1, len(axes) + 1):
        T.set_option('block_order', order)
        assert T.stat('block_order') == order and T.stat('cellmajor') == (1 if order == len(axes) else 0)
        for m in (0, 1, 2):
            got = T.ndpolate(q, m, 0)
            assert T.stat('fused') == 1
            assert_same_bits(want[m][0], got[0], f'interps, method {m}, block order {order}')
            assert_same_bits(want[m][1], got[1], f'dists, method {m}, block order {order}')
    T.close()
