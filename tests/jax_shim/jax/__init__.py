"""NumPy-backed stand-in for the parts of `jax` that PyPIC3D's hot path uses.  Test infrastructure only (see ../README.md)."""
import functools
import itertools
import threading
import types

import numpy as _np

from . import numpy as numpy          # noqa: F401  (jax.numpy)
from .numpy import ShimArray, _wrap
from . import lax, sharding, tree_util, scipy   # noqa: F401

__version__ = "0.0-numpy-shim"
Array = ShimArray


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def device_put(x, sharding=None):
    return x


def device_get(x):
    return x


def block_until_ready(x):
    return x


class _Debug:
    @staticmethod
    def print(fmt, **kw):
        print(fmt.format(**kw))


debug = _Debug()


# ---------------------------------------------------------------------------------------------- pytrees (tuples, lists, dicts, NamedTuples)
def _is_leaf(x):
    return not isinstance(x, (tuple, list, dict))


def _tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: _tree_map(f, tree[k], *[r[k] for r in rest]) for k in tree}
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*[_tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree)])
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    return f(tree, *rest)


def _leaves(tree):
    if isinstance(tree, dict):
        return [l for k in tree for l in _leaves(tree[k])]
    if isinstance(tree, (tuple, list)):
        return [l for t in tree for l in _leaves(t)]
    return [tree]


# ---------------------------------------------------------------------------------------------- vmap: a Python loop
def vmap(fun, in_axes=0, out_axes=0):
    def take(arg, axis, i):
        if axis is None:
            return arg
        if isinstance(axis, (tuple, list, dict)) and not _is_leaf(arg):
            if isinstance(arg, dict):
                return {k: take(arg[k], axis[k], i) for k in arg}
            vals = [take(a, ax, i) for a, ax in zip(arg, axis)]
            return type(arg)(*vals) if hasattr(arg, "_fields") else type(arg)(vals)
        return _tree_map(lambda leaf: _wrap(_np.take(_np.asarray(leaf), i, axis=axis)) if leaf is not None else None, arg)

    def size_of(arg, axis):
        if axis is None:
            return None
        if isinstance(axis, (tuple, list)) and not _is_leaf(arg):
            for a, ax in zip(arg, axis):
                n = size_of(a, ax)
                if n is not None:
                    return n
            return None
        for leaf in _leaves(arg):
            if leaf is not None:
                return _np.asarray(leaf).shape[axis]
        return None

    @functools.wraps(fun)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError(f"vmap: {len(args)} arguments but in_axes has {len(axes)} entries")
        n = None
        for a, ax in zip(args, axes):
            m = size_of(a, ax)
            if m is not None:
                if n is not None and m != n:
                    raise ValueError(f"vmap: inconsistent mapped sizes {n} vs {m}")
                n = m
        if n is None:
            raise ValueError("vmap: nothing to map over")
        outs = [fun(*[take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        first = outs[0]

        def stack(*leaves, axis):
            return _wrap(_np.stack([_np.asarray(l) for l in leaves], axis=axis))
        if isinstance(out_axes, (tuple, list)) and not _is_leaf(first):
            vals = [_tree_map(lambda *ls, ax=ax: stack(*ls, axis=ax), first[j], *[o[j] for o in outs[1:]]) for j, ax in enumerate(out_axes)]
            return type(first)(*vals) if hasattr(first, "_fields") else type(first)(vals)
        ax = out_axes if not isinstance(out_axes, (tuple, list)) else out_axes[0]
        return _tree_map(lambda *ls: stack(*ls, axis=ax), first, *outs[1:])
    return mapped


# ---------------------------------------------------------------------------------------------- devices, meshes, shard_map
class _Device:
    def __init__(self, i):
        self.id = i
        self.platform = "cpu"

    def __repr__(self):
        return f"ShimDevice({self.id})"


_N_DEVICES = 64
_DEVICES = [_Device(i) for i in range(_N_DEVICES)]


def devices(backend=None):
    if backend not in (None, "cpu"):
        raise RuntimeError(f"no {backend} devices in the NumPy shim")
    return list(_DEVICES)


def device_count():
    return _N_DEVICES


def default_device(dev):
    import contextlib
    return contextlib.nullcontext()


_ctx = threading.local()      # .coords (dict axis name -> index), .group (shared exchange state)


class _Group:
    def __init__(self, mesh):
        self.mesh = mesh
        self.shape = tuple(mesh.devices.shape)
        self.n = int(_np.prod(self.shape))
        self.barrier = threading.Barrier(self.n)
        self.box = {}
        self.error = None


def _spec_names(spec, ndim):
    names = list(spec) if spec is not None else []
    return names + [None] * (ndim - len(names))


def shard_map(f, mesh=None, in_specs=None, out_specs=None, check_vma=None, check_rep=None, **kw):
    names = tuple(mesh.axis_names)
    shape = tuple(mesh.devices.shape)

    def shard(leaf, spec, coord):
        if leaf is None or spec is None:
            return leaf
        a = _np.asarray(leaf)
        sl = [slice(None)] * a.ndim
        for d, nm in enumerate(_spec_names(spec, a.ndim)):
            if nm is None:
                continue
            k = shape[names.index(nm)]
            if a.shape[d] % k:
                raise ValueError(f"shard_map: dimension {d} of size {a.shape[d]} is not divisible by mesh axis {nm} = {k}")
            w = a.shape[d] // k
            sl[d] = slice(coord[names.index(nm)] * w, (coord[names.index(nm)] + 1) * w)
        return _wrap(a[tuple(sl)].copy())

    def apply_specs(args, specs, coord):
        specs = specs if isinstance(specs, (tuple, list)) and not isinstance(specs, sharding.PartitionSpec) else (specs,) * len(args)
        out = []
        for a, s in zip(args, specs):
            if s is None or (isinstance(s, sharding.PartitionSpec) and len(s) == 0):
                out.append(a)
            else:
                out.append(_tree_map(lambda leaf: shard(leaf, s, coord), a))
        return out

    def assemble(per_device, spec):
        coords = list(itertools.product(*[range(k) for k in shape]))
        first = per_device[coords[0]]
        if spec is None or (isinstance(spec, sharding.PartitionSpec) and len(spec) == 0):
            return first                                         # replicated result

        def cat(*leaves):
            arr = {c: _np.asarray(l) for c, l in zip(coords, leaves)}
            nd = arr[coords[0]].ndim
            mapped = _spec_names(spec, nd)
            full = None
            # concatenate mesh axis by mesh axis (innermost last)
            def build(prefix, axis_i):
                if axis_i == len(names):
                    return arr[tuple(prefix)]
                parts = [build(prefix + [i], axis_i + 1) for i in range(shape[axis_i])]
                if names[axis_i] in mapped:
                    return _np.concatenate(parts, axis=mapped.index(names[axis_i]))
                return parts[0]
            full = build([], 0)
            return _wrap(full)
        return _tree_map(cat, first, *[per_device[c] for c in coords[1:]])

    def run(*args):
        group = _Group(mesh)
        coords = list(itertools.product(*[range(k) for k in shape]))
        results = {}

        def worker(coord):
            _ctx.coords = {nm: coord[i] for i, nm in enumerate(names)}
            _ctx.group = group
            _ctx.coord = coord
            _ctx.op = 0
            try:
                results[coord] = f(*apply_specs(args, in_specs, coord))
            except BaseException as exc:     # noqa: BLE001 -- re-raised on the caller's thread
                group.error = exc
                group.barrier.abort()
        if len(coords) == 1:
            worker(coords[0])
        else:
            threads = [threading.Thread(target=worker, args=(c,)) for c in coords]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if group.error is not None:
            raise group.error
        first = results[coords[0]]
        if isinstance(out_specs, (tuple, list)) and not isinstance(out_specs, sharding.PartitionSpec):
            vals = [assemble({c: results[c][j] for c in coords}, s) for j, s in enumerate(out_specs)]
            return type(first)(*vals) if hasattr(first, "_fields") else tuple(vals)
        return assemble(results, out_specs)
    return run
