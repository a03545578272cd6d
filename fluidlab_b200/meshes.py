"""SDF mesh colliders (parameter/data holders).  Mirrors fluidlab/fluidengine/meshes/mesh.py (`Mesh.__init__` :16-39,
`load_file` :41-66 for the SDF part, `init_transform` :97-127), meshes/static.py (`Static`), meshes/dynamic.py (`Dynamic`) and
meshes/statics.py (`Statics`).  Only what the simulation needs is kept: the baked SDF volume
`{'voxels': float32[res,res,res], 'T_mesh_to_voxels': float64[4,4]}` (utils/mesh.py:63-87; the reference ships them as
`assets/meshes/processed/<name>-128.sdf` pickles), the initial transform and the contact parameters.  Vertices / normals /
colours are rendering data and out of scope.  The @ti.func bodies (`sdf_`, `normal_`, `collide`, `collider_v`) run inside the
CUDA kernels (csrc/fmpm_sdf.cuh)."""
import os
import pickle as pkl
import numpy as np
import torch
from scipy.spatial.transform import Rotation

from .boundaries import _tup
from . import macros as _macros
from .macros import FRICTION, DTYPE_NP


def _trans_quat_scale_to_T(pos, euler, scale):
    """T_init = trans_quat_to_T(pos, quat) @ scale_to_T(scale) (meshes/mesh.py:97-103, utils/geom.py:34-53)."""
    T = np.eye(4, dtype=DTYPE_NP)
    T[:3, :3] = Rotation.from_euler('zyx', np.array(euler, dtype=np.float64)[::-1], degrees=True).as_matrix().astype(DTYPE_NP)
    T[:3, 3] = np.array(pos, dtype=DTYPE_NP)
    S = np.eye(4, dtype=DTYPE_NP); S[[0, 1, 2], [0, 1, 2]] = np.array(scale, dtype=DTYPE_NP)
    return T @ S


class Mesh:
    def __init__(self, file=None, material=None, file_vis=None, sdf_res=128, pos=(0.0, 0.0, 0.0), euler=(0.0, 0.0, 0.0), scale=(1.0, 1.0, 1.0),
                 softness=0, has_dynamics=False, sdf=None, assets_dir=None):
        self.pos, self.euler, self.scale = _tup(pos), _tup(euler), _tup(scale)
        self.raw_file, self.sdf_res = file, sdf_res
        # the reference's yaml configs name the collider material as a string (`material: PLATE`, configs/macros.py:19-30); resolve it in macros
        self.material = getattr(_macros, material) if isinstance(material, str) else material
        self.has_dynamics = has_dynamics
        self.softness = float(softness)
        self.friction = 0.0
        if self.has_dynamics:
            self.friction = float(FRICTION[self.material])
            if sdf is None:  # the reference's baked pickle (meshes/mesh.py:60-66)
                assets_dir = assets_dir or os.environ.get('FLUIDLAB_ASSETS')
                assert assets_dir is not None, 'pass sdf={voxels,T_mesh_to_voxels} or set FLUIDLAB_ASSETS to fluidlab/assets'
                name = os.path.splitext(os.path.basename(file))[0]
                sdf = pkl.load(open(os.path.join(assets_dir, 'meshes', 'processed', f'{name}-{sdf_res}.sdf'), 'rb'))
            self.sdf_voxels_np = np.ascontiguousarray(sdf['voxels'], dtype=DTYPE_NP)
            self.sdf_voxels_res = self.sdf_voxels_np.shape[0]
            T_init = _trans_quat_scale_to_T(self.pos, self.euler, self.scale)
            # meshes/mesh.py:121-127: T_mesh_to_voxels <- T_mesh_to_voxels @ inv(T_init), in DTYPE_NP
            self.T_mesh_to_voxels_np = (np.asarray(sdf['T_mesh_to_voxels']).astype(DTYPE_NP) @ np.linalg.inv(T_init)).astype(DTYPE_NP)
            self._vox_dev = None

    def device_struct(self, lib_mod, device):
        if self._vox_dev is None or self._vox_dev.device != device:
            self._vox_dev = torch.from_numpy(self.sdf_voxels_np).to(device)
        m = lib_mod.FmpmSdfMesh()
        m.voxels = self._vox_dev.data_ptr(); m.res = int(self.sdf_voxels_res)
        m.T_mesh_to_voxels = (lib_mod.C.c_float * 16)(*[float(v) for v in self.T_mesh_to_voxels_np.reshape(-1)])
        m.friction, m.softness = self.friction, self.softness
        return m


class Static(Mesh):
    """static mesh-based object (meshes/static.py); collides on the grid when has_dynamics (MPM:388-390)."""


class Dynamic(Mesh):
    """mesh posed by its effector (`container`) (meshes/dynamic.py); always has dynamics when owned by a Rigid."""

    def __init__(self, container, **kwargs):
        self.container = container
        super().__init__(**kwargs)


class Statics(list):
    def add_static(self, **kwargs):
        self.append(Static(**kwargs))
