"""ndpolator_b200 -- B200 (sm_100a) implementation of ndpolator's batched
``ndpolate`` hot path behind the reference's own Python API.

    from ndpolator_b200 import Ndpolator          # same class, same methods
    ndpolator_b200.install_as_ndpolator()         # `import ndpolator` now resolves here
"""
__version__ = '1.3.0+b200.1'

from .cndpolator import ExtrapolationMethod, SearchAlgorithm, Table  # noqa: F401
from .ndpolator import Ndpolator  # noqa: F401


def install_as_ndpolator():
    """Registers this package under the names the reference uses (``ndpolator``,
    ``cndpolator``) so that code written against the reference -- including its own
    test-suite -- runs unmodified on the GPU path."""
    import sys
    from . import cndpolator as _c
    sys.modules['ndpolator'] = sys.modules[__name__]
    sys.modules['cndpolator'] = _c
