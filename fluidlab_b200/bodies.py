"""Particle samplers (input generator for the hot path).

Restates the NumPy samplers of fluidlab/fluidengine/bodies/bodies.py that the shipped envs use:
`add_nowhere` (:109-111), `add_cube` random/grid filling (:87-107,113-126), `add_cylinder` random
filling (:128-155), `add_ball` random filling (:157-187) and `get` (:212-235).  The global NumPy
seed is fixed to 0 around every body like the reference does (:26-28,46), so particle sets are
reproducible draw for draw.  Mesh bodies need trimesh voxelisation and are out of scope.
"""
import numpy as np
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

    def _push(self, pts, material, used):
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
        if type == 'nowhere':
            self._push(np.tile(np.array(NOWHERE), (kw['n_particles'], 1)), material, False)
        elif type == 'cube':
            lower = np.array(kw['lower'], dtype=np.float64)
            upper = lower + np.array(kw['size']) if kw.get('size') is not None else np.array(kw['upper'], dtype=np.float64)
            assert (upper >= lower).all()
            self._push(self._sample_box(lower, upper, 'grid' if filling == 'natural' else filling), material, True)
        elif type == 'cylinder':
            if filling == 'natural':
                raise NotImplementedError('natural cylinder filling is not restated')
            c, r, h = np.array(kw['center'], dtype=np.float64), float(kw['radius']), float(kw['height'])
            pts = self._sample_box(np.array([c[0] - r, c[1] - h / 2.0, c[2] - r]), np.array([c[0] + r, c[1] + h / 2.0, c[2] + r]), filling)
            self._push(pts[np.linalg.norm(pts[:, [0, 2]] - c[[0, 2]], axis=1) <= r], material, True)
        elif type == 'ball':
            if filling == 'natural':
                raise NotImplementedError('natural ball filling is not restated')
            c, r = np.array(kw['center'], dtype=np.float64), float(kw['radius'])
            pts = self._sample_box(c - r, c + r, filling)
            self._push(pts[np.linalg.norm(pts - c, axis=1) <= r], material, True)
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
