"""`jax.lax`: cond, conv_general_dilated (VALID, stride 1), and the collectives of `shard_map` over the fake-device mesh."""
import numpy as _np

from .numpy import _wrap


def cond(pred, true_fun, false_fun, *operands, operand=_np._NoValue):
    if operand is not _np._NoValue:          # legacy keyword form: cond(pred, tf, ff, operand=x)
        operands = (operand,)
    return true_fun(*operands) if bool(_np.asarray(pred)) else false_fun(*operands)


def select(pred, a, b):
    return _wrap(_np.where(pred, a, b))


def stop_gradient(x):
    return x


def dynamic_slice(x, start, sizes):
    sl = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start, sizes))
    return _wrap(_np.asarray(x)[sl])


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None, rhs_dilation=None, dimension_numbers=None, **kw):
    """N-D cross-correlation, batch/feature layout (N, C, spatial...) x (O, I, spatial...), stride 1, VALID padding: what
    utilities/filters.py asks for."""
    lhs, rhs = _np.asarray(lhs), _np.asarray(rhs)
    if dimension_numbers is not None:
        dn = tuple(dimension_numbers)
        ok = all(isinstance(s, str) for s in dn) and dn[0][:2] == "NC" and dn[1][:2] == "OI" and dn[2][:2] == "NC"
        if not ok:
            raise NotImplementedError(f"conv_general_dilated shim: dimension_numbers {dimension_numbers}")
    if any(int(s) != 1 for s in window_strides) or str(padding).upper() != "VALID":
        raise NotImplementedError("conv_general_dilated shim: stride 1, VALID only")
    n, c = lhs.shape[:2]
    o, i = rhs.shape[:2]
    if c != i:
        raise ValueError("conv_general_dilated shim: channel mismatch")
    ks = rhs.shape[2:]
    out_sp = tuple(l - k + 1 for l, k in zip(lhs.shape[2:], ks))
    out = _np.zeros((n, o) + out_sp, dtype=_np.result_type(lhs, rhs))
    for off in _np.ndindex(*ks):
        sl = tuple(slice(a, a + m) for a, m in zip(off, out_sp))
        patch = lhs[(slice(None), slice(None)) + sl]                       # (n, c, spatial)
        w = rhs[(slice(None), slice(None)) + off]                          # (o, i)
        out += _np.einsum("nc...,oc->no...", patch, w)
    return _wrap(out)


# ---------------------------------------------------------------------------------------------- collectives
def _ctx():
    from . import _ctx as ctx
    if not hasattr(ctx, "group"):
        raise RuntimeError("collective used outside shard_map")
    return ctx


def axis_index(axis_name):
    c = _ctx()
    if isinstance(axis_name, (tuple, list)):
        idx = 0
        for nm in axis_name:
            idx = idx * c.group.shape[c.group.mesh.axis_names.index(nm)] + c.coords[nm]
        return _wrap(_np.asarray(idx))
    return _wrap(_np.asarray(c.coords[axis_name]))


def _exchange(value):
    """Every device posts `value`; returns the dict coord -> value after all have posted."""
    c = _ctx()
    g = c.group
    key = c.op
    c.op += 1
    g.box.setdefault(key, {})[c.coord] = value
    if g.n > 1:
        g.barrier.wait()
    snapshot = dict(g.box[key])
    if g.n > 1:
        g.barrier.wait()
        if c.coord == min(snapshot):
            g.box.pop(key, None)
    else:
        g.box.pop(key, None)
    return snapshot


def ppermute(x, axis_name, perm):
    from . import _tree_map
    c = _ctx()
    names = tuple(c.group.mesh.axis_names)
    ax = names.index(axis_name)
    me = c.coord
    posted = _exchange(x)
    src = None
    for s, d in perm:
        if int(d) == me[ax]:
            src = int(s)
    if src is None:
        return _tree_map(lambda leaf: _wrap(_np.zeros_like(_np.asarray(leaf))), x)
    from_coord = tuple(src if i == ax else v for i, v in enumerate(me))
    return _tree_map(lambda leaf: _wrap(_np.array(_np.asarray(leaf), copy=True)), posted[from_coord])


def _reduce(x, axis_name, fn):
    from . import _tree_map
    c = _ctx()
    names = tuple(c.group.mesh.axis_names)
    axes = [names.index(n) for n in (axis_name if isinstance(axis_name, (tuple, list)) else (axis_name,))]
    posted = _exchange(x)
    me = c.coord
    members = [co for co in posted if all(co[i] == me[i] for i in range(len(names)) if i not in axes)]
    vals = [posted[co] for co in sorted(members)]
    return _tree_map(lambda *leaves: _wrap(fn(_np.stack([_np.asarray(l) for l in leaves]), axis=0)), vals[0], *vals[1:])


def pmax(x, axis_name):
    return _reduce(x, axis_name, _np.max)


def pmin(x, axis_name):
    return _reduce(x, axis_name, _np.min)


def psum(x, axis_name):
    return _reduce(x, axis_name, _np.sum)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(_np.asarray(cond_fun(val))):
        val = body_fun(val)
    return val


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def scan(f, init, xs, length=None):
    carry = init
    ys = []
    n = length if xs is None else len(xs)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    return carry, ys
