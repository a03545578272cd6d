import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    if os.environ.get("FLUIDLAB_CUDA_EMU") == "1":
        # Development aid (never set by the driver): run `-m gpu` tests of SMALL scenes on the CUDA execution-model shim of tests/cuda_emu/, e.g.
        #   FLUIDLAB_CUDA_EMU=1 python -m pytest tests/test_gpu_parity.py -m gpu -k "reference_kernels or rigid"
        # Simulators built without an explicit device land on CPU tensors and the emulated library; anything that touches CUDA directly fails.
        sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
        import harness
        import torch
        harness.enable()
        torch.cuda.is_available = lambda: True
        from fluidlab_b200 import simulator
        init = simulator.MPMSimulator.__init__

        def emu_init(self, *a, device=None, **k):
            init(self, *a, device="cpu" if device is None else device, **k)
            self.use_graphs = False
        simulator.MPMSimulator.__init__ = emu_init


# `-m gpu` tests that have not yet passed on a B200 are collected AFTER the hardware-validated ones, so that with `-x` a first-run surprise in a new
# path does not hide the state of the paths that have already been measured.  Remove a pattern once its tests have passed on hardware.  (Round 2:
# everything round 1 had left here has run green on hardware — profiles/README.md; the re-conditioned C4 tests and the staged-upload test in r02x / r02final.)
FIRST_HARDWARE_RUN_PENDING = ()


def pytest_collection_modifyitems(config, items):
    def pending(item):
        return item.get_closest_marker('gpu') is not None and any(p in item.nodeid for p in FIRST_HARDWARE_RUN_PENDING)
    items.sort(key=pending)   # stable: the order inside each group is unchanged


def make_particles(x, mat_ids, n_grid, used=None, rho=None):
    """Particle dict in the format OracleSim / the CUDA simulator build() expect (MPM:136-175)."""
    from fluidlab_b200.macros import MU, LAMDA, RHO, MAT_CLASS
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    mat = np.broadcast_to(np.asarray(mat_ids, dtype=np.int32), (n,)).copy()
    dx = 1.0 / n_grid
    p_vol = (dx * 0.5) ** 2
    def table(tab, dtype):   # vectorised material-table lookup (8M-particle scenes)
        ids, inv = np.unique(mat, return_inverse=True)
        return np.array([tab[int(m)] for m in ids], dtype=dtype)[inv.reshape(-1)]
    rho = table(RHO, np.float64) if rho is None else np.asarray(rho, dtype=np.float64)
    return dict(
        x=x, mat=mat, used=np.ones(n, dtype=np.int32) if used is None else np.asarray(used, dtype=np.int32),
        cls=table(MAT_CLASS, np.int32),
        mu=table(MU, np.float32).astype(np.float64),
        lam=table(LAMDA, np.float32).astype(np.float64),
        rho=rho, body_id=np.zeros(n, dtype=np.int32), bodies={'n': 1},
        mass=(np.float32(p_vol) * rho.astype(np.float32)).astype(np.float64),  # MPM:174 evaluates in f32
    )


@pytest.fixture
def particles_factory():
    return make_particles


def sphere_sdf(radius, half_extent, res=32):
    """synthetic baked SDF volume in the reference's pickle format (utils/mesh.py:63-87): voxels[res^3] of a sphere of
    `radius` centred at the mesh origin; the volume spans [-half_extent, half_extent]^3 in mesh coordinates."""
    ax = np.linspace(-half_extent, half_extent, res)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    vox = (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - radius).astype(np.float32)
    s = (res - 1) / (2 * half_extent)
    T = np.eye(4); T[0, 0] = T[1, 1] = T[2, 2] = s; T[:3, 3] = s * half_extent
    return vox, T


def box_sdf(half, half_extent, res=32):
    """SDF of an axis-aligned box with half sizes `half` (3,)"""
    ax = np.linspace(-half_extent, half_extent, res)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    q = np.stack([np.abs(X) - half[0], np.abs(Y) - half[1], np.abs(Z) - half[2]], -1)
    vox = (np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)).astype(np.float32)
    s = (res - 1) / (2 * half_extent)
    T = np.eye(4); T[0, 0] = T[1, 1] = T[2, 2] = s; T[:3, 3] = s * half_extent
    return vox, T


def cone_sdf(radius, height, half_extent, res=48):
    """SDF volume of a solid cone (apex at +z = height / 2, base disc of `radius` at z = -height / 2, axis = mesh z), in the reference's baked
    format — an analytic stand-in for assets/meshes/processed/cone_tip-128.sdf, which cannot travel to the GPU box (8 MB, not the repo's to ship)."""
    ax = np.linspace(-half_extent, half_extent, res)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    r = np.sqrt(X ** 2 + Y ** 2)
    zt = Z + height / 2                     # height above the base
    # distance to the slanted side (line from (radius, 0) to (0, height) in the (r, z) half plane), negative inside
    nrm = np.array([height, radius]) / np.hypot(height, radius)
    side = (r - radius) * nrm[0] + zt * nrm[1]
    vox = np.maximum(side, -zt).astype(np.float32)   # intersection of the side's half space and z >= base
    s = (res - 1) / (2 * half_extent)
    T = np.eye(4); T[0, 0] = T[1, 1] = T[2, 2] = s; T[:3, 3] = s * half_extent
    return vox, T
