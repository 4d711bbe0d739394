"""Drop-ins for PyPIC3D/utilities/filters.py:73-155 (27-point VALID convolution on the tile interior)."""
import torch

from .. import ops


def _filter_params(phi, num_guard_cells):
    p = ops._lib.PicParams()
    p.dtype = 0 if phi.dtype == ops.F32 else 1
    p.g = int(num_guard_cells)
    lead = 1
    for n in phi.shape[:-3]:
        lead *= int(n)
    p.mesh[0], p.mesh[1], p.mesh[2] = lead, 1, 1
    for a in range(3):
        p.gmesh[a] = p.mesh[a]
        p.tile[a] = int(phi.shape[-3 + a]) - 2 * p.g
    return p


def bilinear_filter(phi, num_guard_cells=1, _sp=None):
    phi = ops._chk(phi, "phi")
    return ops.filter27(_filter_params(phi, num_guard_cells), "bilinear", 0.0, phi)


def digital_filter(phi, alpha, num_guard_cells=1, _sp=None):
    phi = ops._chk(phi, "phi")
    return ops.filter27(_filter_params(phi, num_guard_cells), "digital", float(ops._lib._scalar(alpha)), phi)


def _vector(fn, field, *args, **kw):
    if isinstance(field, torch.Tensor):
        return torch.stack([fn(field[c].contiguous(), *args, **kw) for c in range(3)], dim=0)
    return tuple(fn(c, *args, **kw) for c in field)


def bilinear_filter_vector(field, num_guard_cells=1, _sp=None):
    return _vector(bilinear_filter, field, num_guard_cells=num_guard_cells)


def digital_filter_vector(field, alpha, num_guard_cells=1, _sp=None):
    return _vector(digital_filter, field, alpha, num_guard_cells=num_guard_cells)
