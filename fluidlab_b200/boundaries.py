"""Domain boundaries (parameter holders).

Mirrors the constructor surface of fluidlab/fluidengine/boundaries/boundaries.py (`create_boundary`,
`CubeBoundary` :95-104, `CylinderBoundary` :26-37).  The @ti.func bodies (`impose_x_v`, `impose_x`) are
evaluated inside the CUDA kernels (csrc/fmpm_common.cuh boundary_v, csrc/fmpm_io.cu effector_impose_x);
the objects here only carry the f32-rounded parameters, exactly as the reference rounds them.
"""
import numpy as np
from .macros import DTYPE_NP


def _tup(v):
    return tuple(eval(v)) if isinstance(v, str) else tuple(v)  # the reference evals strings (utils/misc.py:20-24)


class Boundary:
    type_id = -1

    def __init__(self, restitution=0.0, lock_dims=()):
        self.restitution = float(restitution)
        self.lock_dims = list(lock_dims)

    @property
    def lock_mask(self):
        m = 0
        for d in self.lock_dims:
            m |= 1 << int(d)
        return m


class CubeBoundary(Boundary):
    type_id = 0

    def __init__(self, lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95), **kwargs):
        super().__init__(**kwargs)
        self.upper = np.array(_tup(upper), dtype=DTYPE_NP)
        self.lower = np.array(_tup(lower), dtype=DTYPE_NP)
        assert (self.upper >= self.lower).all()
        self.xz_center = np.array((0.5, 0.5), dtype=DTYPE_NP)
        self.xz_radius = 0.0


class CylinderBoundary(Boundary):
    type_id = 1

    def __init__(self, y_range=(0.05, 0.95), xz_center=(0.5, 0.5), xz_radius=0.45, **kwargs):
        super().__init__(**kwargs)
        y_range = np.array(_tup(y_range), dtype=DTYPE_NP)
        self.lower = np.array([0.0, y_range[0], 0.0], dtype=DTYPE_NP)
        self.upper = np.array([1.0, y_range[1], 1.0], dtype=DTYPE_NP)
        self.xz_center = np.array(_tup(xz_center), dtype=DTYPE_NP)
        self.xz_radius = float(xz_radius)


def create_boundary(type='cube', **kwargs):
    if type == 'cylinder':
        return CylinderBoundary(**kwargs)
    if type == 'cube':
        return CubeBoundary(**kwargs)
    assert False, f'unknown boundary type {type}'
