"""Particle samplers (input generator for the hot path).

Restates the NumPy samplers of fluidlab/fluidengine/bodies/bodies.py that the shipped envs use:
`add_nowhere` (:109-111), `add_cube` (:87-107,113-126), `add_cylinder` (:128-155), `add_ball` (:157-185) with the
random / grid / natural fillings, the rotation of a body about its centre of mass (`euler`, :74-77) and `get` (:212-235).
The global NumPy seed is fixed to 0 around every body like the reference does (:26-28,46), so particle sets are
reproducible draw for draw.  Mesh bodies (:187-210) need trimesh voxelisation and are out of scope.
"""
import numpy as np
from scipy.spatial.transform import Rotation
from .macros import RHO, NOWHERE


class Bodies:
    def __init__(self, dim=3, particle_density=1e6):
        self.dim = dim
        self.particle_density = particle_density
        self._x, self._mat, self._used, self._rho, self._bid = [], [], [], [], []

    def __len__(self):
        return len(self._x)

    def _n_for_volume(self, volume):
        return round(volume * self.particle_density)

    def _n_for_length(self, length):
        return round(length * np.cbrt(self.particle_density))

    def _sample_box(self, lower, upper, filling):
        size = upper - lower
        if filling == 'random':
            n = self._n_for_volume(np.prod(size))
            return np.random.uniform(low=lower, high=upper, size=(n, self.dim))
        if filling == 'grid':
            axes = [np.linspace(lower[d], upper[d], self._n_for_length(size[d]) + 1) for d in range(3)]
            return np.stack(np.meshgrid(*axes, indexing='ij'), -1).reshape((-1, 3))
        raise NotImplementedError(f'Unsupported filling type: {filling}.')

    def _ring(self, r, y, c):
        """one horizontal ring of the natural fillings: round(2 pi r * density^(1/3)) (at least one) equally spaced points"""
        n = max(self._n_for_length(2 * np.pi * r), 1)
        ang = np.linspace(0, 2 * np.pi, n + 1)[:-1]
        return np.stack([np.cos(ang) * r + c[0], np.full(n, y), np.sin(ang) * r + c[2]], 1)

    def _push(self, pts, material, used, euler=(0.0, 0.0, 0.0)):
        pts = np.asarray(pts, dtype=np.float64)
        # every body is rotated about its centre of mass, also by the identity (bodies.py:74-77): (R (p - c)) + c in this order, so that
        # the floating-point values match the reference's
        Rm = Rotation.from_euler('zyx', np.array(euler, dtype=np.float64)[::-1], degrees=True).as_matrix()
        com = pts.mean(0)
        pts = (Rm @ (pts - com).T).T + com
        n = len(pts)
        self._bid.append(np.full(n, len(self._x)))
        self._x.append(np.asarray(pts, dtype=np.float64))
        self._mat.append(np.full(n, material))
        self._used.append(np.full(n, used))
        self._rho.append(np.full(n, RHO[material]))

    def add_body(self, type, filling='random', **kw):
        assert filling in ['random', 'grid', 'natural'], f'Unsupported filling type: {filling}.'
        state = np.random.get_state()
        np.random.seed(0)
        material = kw['material']
        euler = kw.get('euler', (0.0, 0.0, 0.0))   # `color` (rendering only) is accepted and ignored
        if type == 'nowhere':
            self._push(np.tile(np.array(NOWHERE), (kw['n_particles'], 1)), material, False, euler)
        elif type == 'cube':
            lower = np.array(kw['lower'], dtype=np.float64)
            upper = lower + np.array(kw['size']) if kw.get('size') is not None else np.array(kw['upper'], dtype=np.float64)
            assert (upper >= lower).all()
            self._push(self._sample_box(lower, upper, 'grid' if filling == 'natural' else filling), material, True, euler)
        elif type == 'cylinder':
            c, r, h = np.array(kw['center'], dtype=np.float64), float(kw['radius']), float(kw['height'])
            if filling == 'natural':   # layers in y, concentric rings in each layer (bodies.py:132-145)
                rings = [self._ring(rr, yy, c) for yy in np.linspace(c[1] - h / 2, c[1] + h / 2, self._n_for_length(h) + 1)
                         for rr in np.linspace(0, r, self._n_for_length(r) + 1)]
                self._push(np.concatenate(rings), material, True, euler)
            else:
                pts = self._sample_box(np.array([c[0] - r, c[1] - h / 2.0, c[2] - r]), np.array([c[0] + r, c[1] + h / 2.0, c[2] + r]), filling)
                self._push(pts[np.linalg.norm(pts[:, [0, 2]] - c[[0, 2]], axis=1) <= r], material, True, euler)
        elif type == 'ball':
            c, r = np.array(kw['center'], dtype=np.float64), float(kw['radius'])
            if filling == 'natural':   # concentric spheres, latitude circles on each (bodies.py:160-174)
                rings = []
                for rs in np.linspace(0, r, self._n_for_length(r) + 1):
                    for lat in np.linspace(-np.pi / 2, np.pi / 2, self._n_for_length(rs * np.pi) + 1):
                        y = c[1] + np.sin(lat) * rs
                        rings.append(self._ring(np.sqrt(max(rs ** 2 - (c[1] - y) ** 2, 0)), y, c))
                self._push(np.concatenate(rings), material, True, euler)
            else:
                pts = self._sample_box(c - r, c + r, filling)
                self._push(pts[np.linalg.norm(pts - c, axis=1) <= r], material, True, euler)
        elif type == 'mesh':
            raise NotImplementedError('mesh bodies (bodies.py:187-210) need the trimesh voxeliser, which is outside this hot path: '
                                      'sample the mesh offline and pass the points through a custom body')
        else:
            raise NotImplementedError(f'Unsupported body type: {type}.')
        np.random.set_state(state)

    def get(self):
        if not self._x:
            return None
        body_id = np.concatenate(self._bid)
        out = {
            'x': np.concatenate(self._x), 'mat': np.concatenate(self._mat), 'used': np.concatenate(self._used),
            'rho': np.concatenate(self._rho), 'body_id': body_id,
            'bodies': {'n': len(self._x), 'n_particles': [len(b) for b in self._x],
                       'particle_ids': [np.sort(np.where(body_id == i)[0]) for i in range(len(self._x))]},
        }
        return out
