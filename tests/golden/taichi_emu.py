"""A tiny NumPy emulation of the part of the Taichi 1.1 Python API that the hot-path files of zhouxian/FluidLab use
(`fluidengine/simulators/mpm_simulator.py`, `boundaries/`, `effectors/`, `meshes/static.py|dynamic.py`, `agents/`, `utils/geom.py`).

TEST INFRASTRUCTURE, build container only.  Taichi itself cannot be installed here, but its kernels are plain Python syntax over
`ti.Vector` / `ti.Matrix` / fields: with this module registered as `taichi`, the reference's UNMODIFIED kernel source executes
eagerly, particle by particle and node by node (`tests/golden/make_reference_run.py`), which turns the reference's own forward code into
a (slow) executable specification for small scenes.  What is emulated rather than real — and therefore still not pinned:
  * `ti.svd` (NumPy SVD + the convention assumed in DESIGN.md §2: U, V rotations, sign on the smallest singular value);
  * fp32 arithmetic is NumPy float32 (same IEEE operations, possibly another summation order inside 3x3 products);
  * reverse-mode autodiff (`kernel.grad`) does not exist here: only forward kernels can run;
  * value semantics of local vectors: plain `a = b` assignments in the reference source are compiled as `a = copy(b)`
    (`install_value_semantics`), everything else of the source is untouched.
"""
import itertools
import sys
import types

import numpy as np

f32, f64, i32 = np.float32, np.float64, np.int32
default_fp = np.float32


def _fix(a):
    """everything the reference computes is DTYPE_TI = f32: float64 intermediates are NumPy promotion artefacts"""
    if isinstance(a, np.ndarray) and a.dtype == np.float64 and default_fp is np.float32:
        return a.astype(np.float32).view(TiArr)
    if isinstance(a, np.float64) and default_fp is np.float32:
        return np.float32(a)
    return a


class TiArr(np.ndarray):
    """ti.Vector / ti.Matrix value"""

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        ins = [np.asarray(x) if isinstance(x, TiArr) else x for x in inputs]
        if out is not None:
            outs = tuple(np.asarray(o) if isinstance(o, TiArr) else o for o in out)
            res = getattr(ufunc, method)(*ins, out=outs, **kw)
            return out[0] if len(out) == 1 else out
        res = getattr(ufunc, method)(*ins, **kw)
        if isinstance(res, np.ndarray):
            res = res.view(TiArr)
        return _fix(res)

    def __matmul__(self, o):
        return _fix((np.asarray(self) @ np.asarray(o)).view(TiArr))

    def __rmatmul__(self, o):
        return _fix((np.asarray(o) @ np.asarray(self)).view(TiArr))

    def __getitem__(self, k):
        if isinstance(k, tuple):
            k = tuple(int(v) if isinstance(v, (np.integer,)) else v for v in k)
        r = np.ndarray.__getitem__(self, k)
        return r

    # component access, smoke_field.py:247-258
    x = property(lambda self: self[0], lambda self, v: self.__setitem__(0, v))
    y = property(lambda self: self[1], lambda self, v: self.__setitem__(1, v))
    z = property(lambda self: self[2], lambda self, v: self.__setitem__(2, v))
    w = property(lambda self: self[3], lambda self, v: self.__setitem__(3, v))

    # --- ti.Vector / ti.Matrix methods used by the reference
    def norm(self, eps=0.0):
        a = np.asarray(self)
        return _fix(np.sqrt((a * a).sum(dtype=a.dtype) + a.dtype.type(eps)))

    def normalized(self, eps=0.0):
        return self / self.norm(eps)

    def dot(self, o):
        return _fix((np.asarray(self) * np.asarray(o)).sum(dtype=np.asarray(self).dtype))

    def cross(self, o):
        return _fix(np.cross(np.asarray(self), np.asarray(o)).view(TiArr))

    def outer_product(self, o):
        return _fix(np.outer(np.asarray(self), np.asarray(o)).view(TiArr))

    def cast(self, dt):
        if dt is int:
            dt = np.int32
        if dt is float:
            dt = default_fp
        a = np.asarray(self)
        if np.issubdtype(np.dtype(dt), np.integer):
            return np.trunc(a).astype(dt).view(TiArr)
        return a.astype(dt).view(TiArr)

    def transpose(self):
        return np.asarray(self).T.copy().view(TiArr)

    def determinant(self):
        a = np.asarray(self)
        if a.shape == (3, 3):   # written out like a compiler would (no LAPACK)
            return _fix(a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
                        + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))
        return _fix(np.linalg.det(a.astype(np.float64)))

    def inverse(self):
        a = np.asarray(self)
        return np.linalg.inv(a.astype(np.float64)).astype(a.dtype).view(TiArr)

    def sum(self, *a, **k):
        return _fix(np.asarray(self).sum(*a, **k))

    def any(self):
        return bool(np.asarray(self).any())

    def all(self):
        return bool(np.asarray(self).all())

    def fill(self, v):
        np.asarray(self)[...] = v


def _mk(data, dt=None):
    a = np.array([np.asarray(v) for v in data] if isinstance(data, (list, tuple)) else data)
    if dt is not None:
        if dt is int:
            dt = np.int32
        a = a.astype(dt)
    elif a.dtype == np.float64:
        a = a.astype(default_fp)
    elif a.dtype == np.int64:
        a = a.astype(np.int32)
    return a.view(TiArr)


class _VectorNS:
    def __call__(self, data, dt=None):
        return _mk(data, dt)

    @staticmethod
    def zero(dt, n):
        return np.zeros(n, dtype=np.int32 if dt is int else dt).view(TiArr)

    @staticmethod
    def field(n, dtype, shape=None, needs_grad=False, **kw):
        return Field((n,), dtype, shape, needs_grad)


class _MatrixNS:
    def __call__(self, data, dt=None):
        return _mk(data, dt)

    @staticmethod
    def zero(dt, n, m):
        return np.zeros((n, m), dtype=dt).view(TiArr)

    @staticmethod
    def identity(dt, n):
        return np.eye(n, dtype=dt).view(TiArr)

    @staticmethod
    def field(n, m, dtype, shape=None, needs_grad=False, **kw):
        return Field((n, m), dtype, shape, needs_grad)


def _index(idx):
    if not isinstance(idx, tuple):
        idx = (idx,)
    out = []
    for i in idx:
        if i is None:
            continue
        if isinstance(i, np.ndarray) and i.ndim == 1:
            out.extend(int(v) for v in i)
        else:
            out.append(int(i))
    return tuple(out)


class Field:
    """ti.field / ti.Vector.field / ti.Matrix.field"""

    def __init__(self, tail, dtype, shape=None, needs_grad=False):
        self.tail, self.dtype = tuple(tail), (np.int32 if dtype is int else dtype)
        self.needs_grad = needs_grad
        self.arr = None
        self.grad = Field(tail, dtype, None, False) if needs_grad else None
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
        self.shape = shape
        self.arr = np.zeros(shape + self.tail, dtype=self.dtype)
        if self.grad is not None:
            self.grad._alloc(shape)

    def __getitem__(self, idx):
        r = self.arr[_index(idx)]
        return r.view(TiArr) if isinstance(r, np.ndarray) else r

    def __setitem__(self, idx, v):
        self.arr[_index(idx)] = np.asarray(v, dtype=self.dtype)

    def fill(self, v):
        self.arr[...] = v

    def from_numpy(self, a):
        self.arr[...] = np.asarray(a).reshape(self.arr.shape)

    def to_numpy(self):
        return self.arr.copy()

    def copy_from(self, other):
        self.arr[...] = other.arr

    @property
    def n(self):          # ti.Vector.field(n, ...).n (smoke_field.py:334)
        return self.arr.shape[-1]

    def __iter__(self):   # struct-for over a field iterates its INDICES
        if len(self.shape) == 1:
            return iter(range(self.shape[0]))
        return (np.array(t, dtype=np.int32).view(TiArr) for t in itertools.product(*[range(n) for n in self.shape]))


class _StructProxy:
    def __init__(self, sf, idx):
        object.__setattr__(self, '_sf', sf); object.__setattr__(self, '_idx', idx)

    def __getattr__(self, k):
        r = self._sf.data[k][self._idx]
        return r.view(TiArr) if isinstance(r, np.ndarray) else r

    def __setattr__(self, k, v):
        self._sf.data[k][self._idx] = np.asarray(v, dtype=self._sf.data[k].dtype)

    def fill(self, v):
        for a in self._sf.data.values():
            a[self._idx] = v


class StructField:
    def __init__(self, members, shape, needs_grad):
        shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
        self.shape = shape
        self.data = {}
        for name, t in members.items():
            tail, dt = t if isinstance(t, tuple) else ((), t)
            self.data[name] = np.zeros(shape + tuple(tail), dtype=np.int32 if dt is int else dt)
        self.grad = StructField(members, shape, False) if needs_grad else None

    def __getitem__(self, idx):
        return _StructProxy(self, _index(idx))

    def __getattr__(self, k):   # field-level member access: particles.x -> a field-like view (used by losses / to_numpy)
        if k in ('data', 'shape', 'grad'):
            raise AttributeError(k)
        f = Field.__new__(Field)
        f.arr, f.dtype, f.grad, f.tail = self.data[k], self.data[k].dtype, None, ()
        if self.__dict__.get('grad') is not None:
            g = Field.__new__(Field)
            g.arr, g.dtype, g.grad, g.tail = self.grad.data[k], self.grad.data[k].dtype, None, ()
            f.grad = g
        return f

    def fill(self, v):
        for a in self.data.values():
            a[...] = v


class _StructType:
    def __init__(self, **members):
        self.members = members

    def field(self, shape, needs_grad=False, layout=None, **kw):
        return StructField(self.members, shape, needs_grad)


class _Place:
    def __init__(self, shape):
        self.shape = shape

    def place(self, *fields):
        for f in fields:
            if f.arr is None:
                f._alloc(self.shape)


def svd(A, dt=None):
    """ti.svd(A, dt) -> U, Sigma (as a diagonal MATRIX), V with the convention assumed for Taichi 1.1"""
    a = np.asarray(A, dtype=np.float64)
    U, S, Vh = np.linalg.svd(a)
    V = Vh.T
    if np.linalg.det(U) < 0:
        U[:, 2] = -U[:, 2]; S[2] = -S[2]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]; S[2] = -S[2]
    dt = dt or default_fp
    return U.astype(dt).view(TiArr), np.diag(S).astype(dt).view(TiArr), V.astype(dt).view(TiArr)


def _cast(x, dt):
    if dt is int:
        dt = np.int32
    if dt is float:
        dt = default_fp
    if isinstance(x, TiArr):
        return x.cast(dt)
    if np.issubdtype(np.dtype(dt), np.integer):
        return np.dtype(dt).type(np.trunc(x))
    return np.dtype(dt).type(x)


def _floor(x, dt=None):
    r = np.floor(np.asarray(x))
    if dt is not None:
        r = r.astype(np.int32 if dt is int else dt)
    return r.view(TiArr) if isinstance(r, np.ndarray) and r.ndim else r[()]


def _ndrange(*args):
    rs = [range(int(a)) if isinstance(a, (int, np.integer)) else range(int(a[0]), int(a[1])) for a in args]
    return itertools.product(*rs)


def _grouped(it):
    for t in it:
        yield np.array(t, dtype=np.int32).view(TiArr)


def _un(fn):
    return lambda x, *a: _fix(fn(np.asarray(x) if isinstance(x, TiArr) else x, *a)) if not isinstance(x, TiArr) else _fix(fn(np.asarray(x), *a).view(TiArr))


def install():
    """register this module's namespace as `taichi`"""
    ti = _AnyModule('taichi')   # anything not emulated below (smoke / render-only API) resolves to a permissive dummy
    ti.f32, ti.f64, ti.i32 = f32, f64, i32
    ident = lambda f=None, **kw: f if f is not None else (lambda g: g)
    ti.kernel = ti.func = ti.data_oriented = ident
    ti.static = lambda x, *a: x
    ti.ndrange, ti.grouped = _ndrange, _grouped
    ti.Vector, ti.Matrix = _VectorNS(), _MatrixNS()
    ti.field = lambda dtype, shape=None, needs_grad=False, **kw: Field((), dtype, shape, needs_grad)
    tt = types.SimpleNamespace(ndarray=lambda *a, **k: None, vector=lambda n, dt: ((n,), dt), matrix=lambda n, m, dt: ((n, m), dt), struct=_StructType)
    ti.types = tt
    ti.Layout = types.SimpleNamespace(SOA=None, AOS=None)
    ti.root = types.SimpleNamespace(dense=lambda axes, shape: _Place(shape))
    for ax in ('i', 'j', 'k', 'ij', 'ijk'):
        setattr(ti, ax, ax)
    ti.cast, ti.floor, ti.svd = _cast, _floor, svd
    ti.exp, ti.sqrt, ti.sin, ti.cos, ti.abs = _un(np.exp), _un(np.sqrt), _un(np.sin), _un(np.cos), _un(np.abs)
    ti.pow = lambda a, b: _fix(np.power(np.asarray(a) if isinstance(a, TiArr) else a, b).view(TiArr) if isinstance(a, TiArr) else np.power(a, b))
    ti.min = lambda a, b: _fix(np.minimum(np.asarray(a), np.asarray(b)).view(TiArr)) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else min(a, b)
    ti.max = lambda a, b: _fix(np.maximum(np.asarray(a), np.asarray(b)).view(TiArr)) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else max(a, b)
    ti.ad = types.SimpleNamespace(grad_for=lambda *_: (lambda g: g), grad_replaced=lambda f: f)
    ti.init = lambda *a, **k: None
    ti.template = lambda *a, **k: None
    ti.cpu = ti.gpu = ti.cuda = 'arch'
    sys.modules['taichi'] = ti
    return ti


class _Anything:
    """permissive stand-in for objects of the optional / rendering dependencies the reference imports (yacs, gym, trimesh, ...)"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and callable(a[0]) and not isinstance(a[0], _Anything):
            return a[0]
        return _Anything()

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        o = _Anything(); object.__setattr__(self, k, o); return o

    def __getitem__(self, k):
        return _Anything()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())


class _AnyModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        o = _Anything(); setattr(self, k, o); return o


def _value_copy(v):
    return v.copy() if isinstance(v, np.ndarray) else v


def install_value_semantics(ref_root):
    """Taichi vectors / matrices have VALUE semantics: `inc = pos_voxels; inc[i] += delta` (meshes/dynamic.py:72-76) must not touch
    `pos_voxels`.  Python names alias, so the reference's source is compiled with every plain `name = other_name` / `name = field[...]` /
    `name = obj.attr` assignment rewritten to `name = copy(...)` (a no-op for anything that is not an array).  Nothing else of the source
    is changed."""
    import ast
    import builtins
    from importlib.machinery import SourceFileLoader
    builtins.__ti_value_copy__ = _value_copy

    class T(ast.NodeTransformer):
        def visit_Assign(self, node):
            self.generic_visit(node)
            # `name = other_name`, and loads such as `vl = self.grid.v_tmp[s, I]` / `x = self.particles[f, p].x` (a field load yields a VALUE in
            # Taichi; smoke_field.py:238-248 then overwrites components of the local)
            if isinstance(node.value, (ast.Name, ast.Subscript, ast.Attribute)) and all(isinstance(t, ast.Name) for t in node.targets):
                node.value = ast.Call(func=ast.Name(id='__ti_value_copy__', ctx=ast.Load()), args=[node.value], keywords=[])
            return node
    orig = SourceFileLoader.source_to_code

    def source_to_code(self, data, path, *, _optimize=-1):
        if str(path).startswith(ref_root):
            tree = ast.fix_missing_locations(T().visit(ast.parse(data, path)))
            return compile(tree, path, 'exec', dont_inherit=True, optimize=_optimize)
        return orig(self, data, path, _optimize=_optimize)
    SourceFileLoader.source_to_code = source_to_code
    sys.dont_write_bytecode = True


def stub_optional_dependencies():
    import importlib
    for name in ('trimesh', 'yacs', 'yacs.config', 'gym', 'gym.spaces', 'mesh_to_sdf', 'skimage', 'skimage.measure', 'matplotlib', 'matplotlib.pyplot',
                 'imageio', 'pyrender', 'open3d', 'cv2', 'OpenGL', 'OpenGL.GL', 'pyglet'):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _AnyModule(name)
    sys.modules['fluidlab.fluidengine.renderers.gl_renderer_src'] = _AnyModule('fluidlab.fluidengine.renderers.gl_renderer_src')   # compiled FleX binding
