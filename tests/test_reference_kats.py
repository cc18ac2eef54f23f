"""Known-answer tests restated from the reference's own test-suite
(reference tests/test_interp.py, test_extrapolation.py, test_ndp_distance.py,
test_kdtree_integration.py, test_exceptions.py), written against the
``Ndpolator`` surface and run on every engine:

  * the CPU oracle restatement            (not gpu)
  * the compiled reference, when present  (not gpu)
  * ndpolator_b200 on the B200            (gpu)
"""
import numpy as np
import pytest

from cases import OracleNdpolator
from oracle.oracle import have


def _make(engine, basic_axes):
    if engine == 'b200':
        from ndpolator_b200 import Ndpolator
        return Ndpolator(basic_axes=basic_axes)
    return OracleNdpolator(basic_axes, kind=engine)


ENGINES = [
    pytest.param('port', id='oracle-port'),
    pytest.param('reference', id='oracle-reference',
                 marks=pytest.mark.skipif(not have('reference'), reason='oracle/_ref not built')),
    pytest.param('b200', id='b200', marks=pytest.mark.gpu),
]


def _np(x):
    return x.cpu().numpy() if hasattr(x, 'cpu') else np.asarray(x)


def _axes3():
    return (np.linspace(1000, 5000, 5), np.linspace(1, 5, 5), np.linspace(0.1, 0.5, 5))


def _fv(p):
    return p[0] / 1000 + p[1] + p[2] * 10


def _linear_grid(axes):
    g = np.empty(tuple(len(a) for a in axes) + (1,))
    for i, x in enumerate(axes[0]):
        for j, y in enumerate(axes[1]):
            for k, z in enumerate(axes[2]):
                g[i, j, k, 0] = _fv((x, y, z))
    return g


@pytest.mark.parametrize('engine', ENGINES)
def test_import_query_pts_kat(engine):
    # reference tests/test_interp.py:46-82
    axes = _axes3()
    ndp = _make(engine, axes)
    ndp.register('main', None, np.random.default_rng(0).normal(1.0, 0.1, (5, 5, 5, 1)))
    q = np.array([[750, 2.25, 0.35], [1250, 2.5, 0.375], [2500, 3.75, 0.4], [2750, 4.0, 0.5], [3000, 4.25, 0.525]])
    indices, flags, normed = (_np(x) for x in ndp.import_query_pts('main', q))
    assert np.array_equal(indices, [[1, 2, 3], [1, 2, 3], [2, 3, 4], [2, 4, 4], [3, 4, 4]])
    assert np.array_equal(flags, [[2, 0, 0], [0, 0, 0], [0, 0, 1], [0, 1, 1], [1, 0, 2]])
    assert np.allclose(normed, [[-0.25, 0.25, 0.50], [0.25, 0.50, 0.75], [0.50, 0.75, 0.00],
                                [0.75, 0.00, 1.00], [0.00, 0.25, 1.25]])


@pytest.mark.parametrize('engine', ENGINES)
def test_find_hypercubes_shapes(engine):
    # reference tests/test_interp.py:96-112
    axes = _axes3()
    ndp = _make(engine, axes)
    ndp.register('main', None, np.random.default_rng(0).normal(1.0, 0.1, (5, 5, 5, 1)))
    q = np.array([[750, 2.25, 0.35], [1250, 2.5, 0.375], [2500, 3.75, 0.4], [2750, 4.0, 0.5],
                  [3000, 4.25, 0.525], [3000, 4.0, 0.425], [3000, 4.0, 0.4]])
    indices, flags, normed = (_np(x) for x in ndp.import_query_pts('main', q))
    hcs = ndp.find_hypercubes('main', normed_query_pts=normed, indices=indices, flags=flags)
    expected = [(2, 2, 2, 1), (2, 2, 2, 1), (2, 2, 1), (2, 1), (2, 2, 1), (2, 1), (1,)]
    assert [h.shape for h in hcs] == expected


@pytest.mark.parametrize('engine', ENGINES)
def test_distances_kat(engine):
    # reference tests/test_interp.py:134-179: 27 points around a 5^3 grid, 0.25 cells outside
    axes = _axes3()
    ndp = _make(engine, axes)
    ndp.register('main', None, _linear_grid(axes))
    xs, ys, zs = [750, 2750, 5250], [0.75, 2.75, 5.25], [0.075, 0.275, 0.525]
    q = np.array([[x, y, z] for x in xs for y in ys for z in zs], dtype=float)
    outside = np.array([(i != 1) + (j != 1) + (k != 1) for i in range(3) for j in range(3) for k in range(3)])
    res = ndp.ndpolate('main', q, extrapolation_method='nearest')
    expected_nearest = np.full(27, 3 * 0.25 ** 2)
    expected_nearest[13] = 0.0
    assert np.allclose(_np(res['dists'])[:, 0], expected_nearest)
    res = ndp.ndpolate('main', q, extrapolation_method='linear')
    assert np.allclose(_np(res['dists'])[:, 0], outside * 0.25 ** 2)


@pytest.mark.parametrize('engine', ENGINES)
def test_ndpolate_linear_field(engine):
    # reference tests/test_interp.py:201-233 (seeded here)
    axes = _axes3()
    ndp = _make(engine, axes)
    ndp.register('main', None, _linear_grid(axes))
    rng = np.random.default_rng(42)
    q = np.ascontiguousarray(np.vstack((rng.uniform(1000, 5000, 1000), rng.uniform(1, 5, 1000), rng.uniform(0.1, 0.5, 1000))).T)
    expected = np.array([_fv(p) for p in q])
    assert np.allclose(_np(ndp.ndpolate('main', q, extrapolation_method='none')['interps'])[:, 0], expected, rtol=1e-5)

    q = np.ascontiguousarray(np.vstack((rng.uniform(500, 5500, 1000), rng.uniform(0.5, 5.5, 1000), rng.uniform(0.05, 0.55, 1000))).T)
    expected = np.array([_fv(p) for p in q])
    oob = ((q[:, 0] < axes[0][0]) | (q[:, 0] > axes[0][-1]) | (q[:, 1] < axes[1][0]) | (q[:, 1] > axes[1][-1]) |
           (q[:, 2] < axes[2][0]) | (q[:, 2] > axes[2][-1]))
    got = _np(ndp.ndpolate('main', q, extrapolation_method='none')['interps'])[:, 0]
    assert np.array_equal(np.isnan(got), oob)

    nearest = np.array([_fv((axes[0][np.argmin(np.abs(axes[0] - p[0]))], axes[1][np.argmin(np.abs(axes[1] - p[1]))],
                             axes[2][np.argmin(np.abs(axes[2] - p[2]))])) for p in q])
    rounded = np.where(oob, nearest, expected)
    got = _np(ndp.ndpolate('main', q, extrapolation_method='nearest')['interps'])[:, 0]
    assert np.allclose(got, rounded, rtol=1e-5)

    got = _np(ndp.ndpolate('main', q, extrapolation_method='linear')['interps'])[:, 0]
    assert np.allclose(got, expected, rtol=1e-5)


@pytest.mark.parametrize('engine', ENGINES)
def test_all_vertices_and_one_on_vertex_point(engine):
    # reference tests/test_interp.py:255-259, :281-290
    axes = _axes3()
    ndp = _make(engine, axes)
    ndp.register('main', None, _linear_grid(axes))
    q = np.array([(x, y, z) for x in axes[0] for y in axes[1] for z in axes[2]])
    got = _np(ndp.ndpolate('main', q, extrapolation_method='none')['interps'])[:, 0]
    assert np.allclose(got, [_fv(p) for p in q], rtol=1e-5)
    q = np.array([[2500, 5.0, 0.25]])
    got = _np(ndp.ndpolate('main', q, extrapolation_method='none')['interps'])[:, 0]
    assert np.allclose(got, [_fv(q[0])], rtol=1e-5)


@pytest.mark.parametrize('engine', ENGINES)
def test_extrapolation_2d_and_1d(engine):
    # reference tests/test_extrapolation.py:31-134: f = x + 2y on a 3x3 grid; 1-D boundaries
    ax = (np.array([0., 1., 2.]), np.array([0., 1., 2.]))
    grid = np.array([[[x + 2 * y] for y in ax[1]] for x in ax[0]])
    ndp = _make(engine, ax)
    ndp.register('t', None, grid)
    inside = np.array([[0.5, 0.5], [1.5, 1.0], [0.25, 1.75]])
    assert np.allclose(_np(ndp.ndpolate('t', inside, 'none')['interps'])[:, 0], inside[:, 0] + 2 * inside[:, 1], atol=1e-10)
    out = np.array([[-0.5, 0.5], [2.5, 1.0], [1.0, 3.0], [-1.0, -1.0]])
    assert np.all(np.isnan(_np(ndp.ndpolate('t', out, 'none')['interps'])))
    got = _np(ndp.ndpolate('t', out, 'nearest')['interps'])[:, 0]
    assert np.all(np.isfinite(got))
    got = _np(ndp.ndpolate('t', out, 'linear')['interps'])[:, 0]
    assert np.allclose(got, out[:, 0] + 2 * out[:, 1], atol=1e-10)

    ax1 = (np.array([0., 1., 2., 3.]),)
    g1 = np.array([[0.], [10.], [20.], [30.]])
    ndp1 = _make(engine, ax1)
    ndp1.register('t', None, g1)
    q = np.array([[-1.0], [0.0], [3.0], [4.0], [100.0]])
    got = _np(ndp1.ndpolate('t', q, 'nearest')['interps'])[:, 0]
    assert np.allclose(got, [0, 0, 30, 30, 30])
    got = _np(ndp1.ndpolate('t', q, 'linear')['interps'])[:, 0]
    assert np.allclose(got, [-10, 0, 30, 40, 1000])
    res = ndp1.ndpolate('t', np.array([[5.0]]), 'linear')
    assert _np(res['interps']).shape == (1, 1) and _np(res['dists']).shape == (1, 1)
    with pytest.raises(ValueError):
        ndp1.ndpolate('t', q, extrapolation_method='cubic')
    with pytest.raises(ValueError):
        ndp1.ndpolate('t', q, search_algorithm='octree')


@pytest.mark.parametrize('engine', ENGINES)
def test_distance_api(engine):
    # reference tests/test_ndp_distance.py:42-100 (non-uniform scales) and :135-195 (NaN-holed grid)
    ax = (np.array([0., 10., 20., 30.]), np.array([0., 1., 2., 3.]), np.array([0., 0.1, 0.2, 0.3]))
    ndp = _make(engine, ax)
    ndp.register('t', None, np.ones((4, 4, 4, 1)))
    q = np.array([[15, 1.5, 0.15], [35, 1.5, 0.15], [-5, 1.5, 0.15], [15, 3.5, 0.15], [15, 1.5, 0.35],
                  [35, 3.5, 0.15], [35, 3.5, 0.35], [-10, -1, -0.1], [30, 3, 0.3], [0, 0, 0]])
    d = _np(ndp.distance('t', q))[:, 0]
    expected = [0, 0.25, 0.25, 0.25, 0.25, 0.5, 0.75, 3.0, 0, 0]
    assert np.allclose(d, expected, atol=1e-9)

    grid = np.ones((4, 4, 4, 1))
    grid[3, 3, 3, 0] = np.nan
    grid[0, 0, 0, 0] = np.nan
    ndp2 = _make(engine, ax)
    ndp2.register('t', None, grid)
    q = np.array([[15, 1.5, 0.15], [25, 2.5, 0.25], [5, 0.5, 0.05], [29, 2.9, 0.29], [35, 3.5, 0.35]])
    d = _np(ndp2.distance('t', q))[:, 0]
    # the two corner cells are not fully defined: points inside them are >0 from the nearest full cell
    assert d[0] == 0 and d[1] > 0 and d[2] > 0 and d[3] > 0 and d[4] > d[3]


@pytest.mark.parametrize('engine', ENGINES)
def test_vertex_payload_mapping_and_determinism(engine):
    # reference tests/test_kdtree_integration.py:38-53, :74-88, :123-138, :236-256
    ax = (np.linspace(0, 4, 5), np.linspace(0, 3, 4), np.linspace(0, 2, 3))
    grid = np.empty((5, 4, 3, 1))
    for i in range(5):
        for j in range(4):
            for k in range(3):
                grid[i, j, k, 0] = 100 * i + 10 * j + k
    ndp = _make(engine, ax)
    ndp.register('t', None, grid)
    verts = np.array([[i, j, k] for i in range(5) for j in range(4) for k in range(3)], dtype=float)
    shifted = verts + np.array([4.3, 3.3, 2.3])       # far outside: nearest vertex is the top corner
    got = _np(ndp.ndpolate('t', shifted, 'nearest')['interps'])[:, 0]
    assert np.all(got == grid[4, 3, 2, 0])
    outside = verts.copy()
    outside[:, 0] = -0.4                               # nearest vertex keeps (j, k), i = 0
    got = _np(ndp.ndpolate('t', outside, 'nearest')['interps'])[:, 0]
    assert np.array_equal(got, 10 * verts[:, 1] + verts[:, 2])
    centres = np.array([[i + 0.5, j + 0.5, k + 0.5] for i in range(4) for j in range(3) for k in range(2)])
    a = _np(ndp.ndpolate('t', centres - [5, 0, 0], 'linear')['interps'])
    b = _np(ndp.ndpolate('t', centres - [5, 0, 0], 'linear')['interps'])
    assert np.array_equal(a, b)
    far = np.array([[1e3, -1e3, 1e3], [-1e4, 1e4, 0.5]])
    assert np.all(np.isfinite(_np(ndp.ndpolate('t', far, 'linear')['interps'])))


def test_exceptions_b200_api():
    # reference tests/test_exceptions.py:11-39 -- argument checks need no GPU
    from ndpolator_b200 import Ndpolator
    ax = np.linspace(0, 1, 3)
    with pytest.raises(TypeError):
        Ndpolator(basic_axes=[ax])
    with pytest.raises(TypeError):
        Ndpolator(basic_axes=(list(ax),))
    with pytest.raises(ValueError):
        Ndpolator(basic_axes=(np.array([1.0]),))
    ndp = Ndpolator(basic_axes=(ax,))
    with pytest.raises(TypeError):
        ndp.register(5, None, np.zeros((3, 1)))
    with pytest.raises(TypeError):
        ndp.register('t', [ax], np.zeros((3, 3, 1)))
    with pytest.raises(TypeError):
        ndp.register('t', None, [[0.0], [1.0], [2.0]])
    ndp.register('t', None, np.zeros((3, 1)))
    assert ndp.tables == ['t'] and repr(ndp) == '<Ndpolator N=1, 1 tables>'
