"""GPU twin of tests/test_device_logic_cpu.py::test_non_finite_queries: +-inf query coordinates
bit-exact in every mode, NaN with extrapolation 'none' (src/ndpolator.c:27-39, :379-381, :551-554).
Runs through the fused kernel where the table qualifies and through the two-pass path otherwise."""
import numpy as np
import pytest

from cases import adversarial_queries, assert_same_bits, build_cases
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('fused_off', [0, 1], ids=['fused', 'two-pass'])
@pytest.mark.parametrize('case', build_cases(seed=3), ids=lambda c: c[0])
def test_non_finite_queries_gpu(case, fused_off):
    from ndpolator_b200.cndpolator import Table
    name, axes, nbasic, grid = case
    q = adversarial_queries(axes, 5000, np.random.default_rng(5))
    qi = q.copy()
    qi[::4, -1] = np.inf
    qi[1::4, 0] = -np.inf
    qn = q.copy()
    qn[::3, 0] = np.nan
    qn[1::7, :] = np.nan
    O, T = Oracle('port', axes, grid, nbasic), Table(axes, grid, nbasic)
    T.set_option('fused', fused_off)
    for a, b, what in zip(O.find(qi), T.find(qi), ('indices', 'flags', 'normed')):
        assert_same_bits(a, b, what + ' (inf)')
    for a, b, what in zip(O.find(qn), T.find(qn), ('indices', 'flags', 'normed')):
        assert_same_bits(a, b, what + ' (nan)')
    for m in (0, 1, 2):
        for s in (0, 1):
            a, b = O.ndpolate(qi, m, s), T.ndpolate(qi, m, s)
            assert_same_bits(a[0], b[0], f'interps {m}/{s} (inf)')
            assert_same_bits(a[1], b[1], f'dists {m}/{s} (inf)')
    a, b = O.ndpolate(qn, 0, 0), T.ndpolate(qn, 0, 0)
    assert_same_bits(a[0], b[0], 'interps none (nan)')
    T.close()
