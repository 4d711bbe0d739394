"""`jax.numpy` on NumPy (float64 / int64 defaults, like jax with x64 enabled).  Arrays are an ndarray subclass with `.at`."""
import numpy as _np

newaxis = None
pi = _np.pi
inf = _np.inf
nan = _np.nan
float64, float32, int32, int64, bool_, uint8, int8, complex128 = (_np.float64, _np.float32, _np.int32, _np.int64, _np.bool_, _np.uint8,
                                                                   _np.int8, _np.complex128)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


def _normalise(arr, idx, mode):
    """jax indexing for scatter/gather with integer index arrays: negative indices wrap once, then out-of-range entries are
    dropped (mode='drop', jax's default for scatters is also drop) -- returns (index tuple, keep mask or None)."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    if any(isinstance(i, (slice, type(Ellipsis))) or i is None for i in idx) or any(_np.asarray(i).dtype == _np.bool_ for i in idx):
        return idx, None
    arrs = _np.broadcast_arrays(*[_np.asarray(i) for i in idx])
    keep = _np.ones(arrs[0].shape, dtype=bool)
    out = []
    for d, a in enumerate(arrs):
        n = arr.shape[d]
        a = _np.where(a < 0, a + n, a)
        keep &= (a >= 0) & (a < n)
        out.append(a)
    return tuple(out), keep


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _apply(self, values, op, mode):
        out = _np.array(self.arr, copy=True)
        idx, keep = _normalise(out, self.idx, mode)
        if keep is None:
            if op == "add":
                _np.add.at(out, idx, values)
            elif op == "set":
                out[idx] = values
            elif op == "mul":
                _np.multiply.at(out, idx, values)
            elif op == "min":
                _np.minimum.at(out, idx, values)
            elif op == "max":
                _np.maximum.at(out, idx, values)
            return out.view(ShimArray)
        v = _np.broadcast_to(_np.asarray(values, dtype=out.dtype), keep.shape + out.shape[len(idx):])
        sel = tuple(a[keep] for a in idx)
        if op == "add":
            _np.add.at(out, sel, v[keep])
        elif op == "set":
            out[sel] = v[keep]
        elif op == "mul":
            _np.multiply.at(out, sel, v[keep])
        elif op == "min":
            _np.minimum.at(out, sel, v[keep])
        elif op == "max":
            _np.maximum.at(out, sel, v[keep])
        return out.view(ShimArray)

    def add(self, values, mode=None, **kw):
        return self._apply(values, "add", mode)

    def set(self, values, mode=None, **kw):
        return self._apply(values, "set", mode)

    def multiply(self, values, mode=None, **kw):
        return self._apply(values, "mul", mode)

    def min(self, values, mode=None, **kw):
        return self._apply(values, "min", mode)

    def max(self, values, mode=None, **kw):
        return self._apply(values, "max", mode)

    def get(self, mode=None, fill_value=None, **kw):
        idx, keep = _normalise(self.arr, self.idx, mode)
        if keep is None:
            return _wrap(_np.asarray(self.arr)[idx])
        safe = tuple(_np.where(keep, a, 0) for a in idx)
        got = _np.asarray(self.arr)[safe]
        if mode == "fill" or fill_value is not None:
            got = _np.where(keep, got, 0 if fill_value is None else fill_value)
        return _wrap(got)


class ShimArray(_np.ndarray):
    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self

    def __hash__(self):
        return id(self)


ndarray = ShimArray


def _wrap(x):
    if isinstance(x, _np.ndarray):
        return x.view(ShimArray)
    if isinstance(x, (_np.generic,)):
        return _np.asarray(x).view(ShimArray)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _lift(fn):
    def f(*a, **k):
        return _wrap(fn(*a, **k))
    f.__name__ = getattr(fn, "__name__", "f")
    return f


def asarray(x, dtype=None, **kw):
    return _wrap(_np.asarray(x, dtype=dtype))


def array(x, dtype=None, copy=True, **kw):
    return _wrap(_np.array(x, dtype=dtype, copy=True))


def round(x, decimals=0):          # noqa: A001 -- jnp.round == round-half-to-even
    return _wrap(_np.round(x, decimals))


def astype(x, dtype):
    return _wrap(_np.asarray(x).astype(dtype))


def __getattr__(name):
    fn = getattr(_np, name)
    if callable(fn) and not isinstance(fn, type):
        return _lift(fn)
    return fn
