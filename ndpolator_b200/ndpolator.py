"""The user-facing class of ndpolator (reference ndpolator/ndpolator.py), kept
verbatim at the API level -- same constructor, ``register``, ``import_query_pts``,
``find_hypercubes``, ``distance``, ``ndpolate``, ``tables``, same exceptions, same
result dictionary -- with the native calls going to the B200 implementation in
:mod:`ndpolator_b200.cndpolator` instead of the reference's C extension."""
from __future__ import annotations

import numpy as np

from . import cndpolator
from .cndpolator import ExtrapolationMethod, SearchAlgorithm


class Ndpolator():
    """N-dimensional interpolation and extrapolation on a (possibly NaN-holed)
    rectilinear grid.  See reference ndpolator/ndpolator.py:7-21.

    Attributes
    ----------
        * axes : tuple[ndarray]   basic axes that span the grid
        * table : dict            name -> {'associated_axes', 'grid', 'capsule'}
    """

    def __init__(self, basic_axes: tuple, device: int | None = None, devices=None) -> None:
        """`device`: GPU ordinal of this process' tables (default: NDPB_DEVICE / LOCAL_RANK / 0).
        `devices`: list of GPU ordinals -- every registered table is replicated on all of them and each
        ndpolate() call on a host array is split across them by contiguous query ranges (one call, several
        GPUs; see cndpolator.Group).  Both are extensions: the reference's constructor takes basic_axes only."""
        # reference ndpolator/ndpolator.py:35-44
        if not isinstance(basic_axes, tuple):
            raise TypeError('parameter `basic_axes` must be a tuple of ndarrays')
        for ti, basic_axis in enumerate(basic_axes):
            if not isinstance(basic_axis, np.ndarray):
                raise TypeError(f'the `basic_axes[{ti}]` element must be a ndarray')
            if len(basic_axis) <= 1:
                raise ValueError('each basic axis must have more than one element')

        self.axes = basic_axes
        self.table = dict()
        self.device = device
        self.devices = list(devices) if devices is not None else None

    def __repr__(self) -> str:
        return f'<Ndpolator N={len(self.axes)}, {len(self.table)} tables>'

    def __str__(self) -> str:
        return f'<Ndpolator N={len(self.axes)}, {len(self.table)} tables>'

    @property
    def tables(self) -> list:
        return list(self.table.keys())

    def register(self, name: str, associated_axes: tuple, grid: np.ndarray) -> None:
        # reference ndpolator/ndpolator.py:91-108
        if not isinstance(name, str):
            raise TypeError('parameter `name` must be a string')
        if associated_axes:
            if not isinstance(associated_axes, tuple):
                raise TypeError('parameter `associated_axes` must be a tuple of ndarrays')
            for ti, associated_axis in enumerate(associated_axes):
                if not isinstance(associated_axis, np.ndarray):
                    raise TypeError(f'the `associated_axes[{ti}]` element must be a ndarray')
                if len(associated_axis) <= 1:
                    raise ValueError('each associated axis must have more than one element')
        if not isinstance(grid, np.ndarray):
            raise TypeError('parameter `grid` must be a ndarray')

        self.table[name] = {
            'associated_axes': associated_axes,
            'grid': grid,
            'capsule': None
        }

    def _all_axes(self, name):
        associated_axes = self.table[name]['associated_axes']
        return self.axes if associated_axes is None else self.axes + associated_axes

    def _capsule(self, name):
        """The device table of `name`, created on first use and cached where the
        reference caches its PyCapsule (ndpolator/ndpolator.py:317-339)."""
        entry = self.table[name]
        if entry['capsule'] is None:
            if self.devices is not None and len(self.devices) > 1:
                entry['capsule'] = cndpolator.Group(self._all_axes(name), entry['grid'], len(self.axes), devices=self.devices)
            else:
                device = self.devices[0] if self.devices else self.device
                entry['capsule'] = cndpolator.Table(self._all_axes(name), entry['grid'], len(self.axes), device=device)
        return entry['capsule']

    def import_query_pts(self, name: str, query_pts: np.ndarray) -> tuple:
        # reference ndpolator/ndpolator.py:144-150
        if not cndpolator._is_cuda(query_pts):
            query_pts = np.ascontiguousarray(query_pts)
        capsule = self.table[name]['capsule']
        if capsule:
            return capsule.find(query_pts)
        indices, flags, normed_query_pts = cndpolator.find(axes=self._all_axes(name), query_pts=query_pts, nbasic=len(self.axes))
        return indices, flags, normed_query_pts

    def find_hypercubes(self, name: str, normed_query_pts: np.ndarray, indices: np.ndarray, flags: np.ndarray, associated_axes: tuple = None) -> np.ndarray:
        # reference ndpolator/ndpolator.py:184-187 (associated axes are an argument here, not looked up)
        axes = self.axes if associated_axes is None else self.axes + associated_axes
        grid = self.table[name]['grid']
        registered = self._all_axes(name)
        same_axes = len(axes) == len(registered) and all(a is b for a, b in zip(axes, registered))
        capsule = self._capsule(name) if same_axes else None
        return cndpolator.hypercubes(normed_query_pts=normed_query_pts, indices=indices, axes=axes, flags=flags, grid=grid, capsule=capsule)

    def ndpolate_cached(self, name: str, normed_query_pts, indices, flags, extrapolation_method: str = 'none',
                        search_algorithm: str = 'kdtree') -> dict:
        """Extension (not in the reference): ndpolate from cached import products -- what import_query_pts()
        returned, kept on the device if they were computed there -- e.g. for several tables on the same axes.
        Same result dictionary as ndpolate()."""
        methods = {'none': ExtrapolationMethod.NONE, 'nearest': ExtrapolationMethod.NEAREST, 'linear': ExtrapolationMethod.LINEAR}
        searches = {'kdtree': SearchAlgorithm.KDTREE, 'linear': SearchAlgorithm.LINEAR}
        if extrapolation_method not in methods:
            raise ValueError(f"extrapolation_method={extrapolation_method} is not valid; it must be one of {methods.keys()}.")
        if search_algorithm not in searches:
            raise ValueError(f"search_algorithm={search_algorithm} is not valid; it must be one of {searches.keys()}.")
        capsule = self._capsule(name)
        table = capsule.member(0) if isinstance(capsule, cndpolator.Group) else capsule
        interps, dists = table.ndpolate_cached(normed_query_pts, indices, flags, methods[extrapolation_method], searches[search_algorithm])
        return {'interps': interps} if extrapolation_method == 'none' else {'interps': interps, 'dists': dists}

    def distance(self, name: str, query_pts: np.ndarray) -> np.ndarray:
        # reference ndpolator/ndpolator.py:229-251 (here the table is cached on first use as well)
        if not cndpolator._is_cuda(query_pts):
            query_pts = np.ascontiguousarray(query_pts)
        return cndpolator.distance(query_pts, self._capsule(name), None, None, len(self.axes))

    def ndpolate(self, name: str, query_pts: np.ndarray, extrapolation_method: str = 'none', search_algorithm: str = 'kdtree') -> dict:
        # reference ndpolator/ndpolator.py:293-349
        extrapolation_methods = {
            'none': ExtrapolationMethod.NONE,
            'nearest': ExtrapolationMethod.NEAREST,
            'linear': ExtrapolationMethod.LINEAR
        }

        search_algorithms = {
            'kdtree': SearchAlgorithm.KDTREE,
            'linear': SearchAlgorithm.LINEAR
        }

        if extrapolation_method not in extrapolation_methods.keys():
            raise ValueError(f"extrapolation_method={extrapolation_method} is not valid; it must be one of {extrapolation_methods.keys()}.")
        extrapolation_method = extrapolation_methods[extrapolation_method]

        if search_algorithm not in search_algorithms.keys():
            raise ValueError(f"search_algorithm={search_algorithm} is not valid; it must be one of {search_algorithms.keys()}.")
        search_algorithm = search_algorithms[search_algorithm]

        if not cndpolator._is_cuda(query_pts):
            query_pts = np.ascontiguousarray(query_pts)

        interps, dists = cndpolator.ndpolate(
            capsule=self._capsule(name),
            query_pts=query_pts,
            nbasic=len(self.axes),
            extrapolation_method=extrapolation_method,
            search_algorithm=search_algorithm
        )

        if extrapolation_method == ExtrapolationMethod.NONE:
            return {
                'interps': interps
            }
        else:
            return {
                'interps': interps,
                'dists': dists
            }
