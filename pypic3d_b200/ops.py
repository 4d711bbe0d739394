"""Low-level operators: torch CUDA tensors in, hand-written sm_100a kernels (C ABI) underneath.

Everything here is functional (inputs are never mutated) unless the name ends in `_`.  No CPU path exists: a tensor
that is not on a CUDA device raises `PicError`.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import PicError, check, make_params

F32, F64 = torch.float32, torch.float64


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, name, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise PicError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise PicError(f"{name}: tensor is on {t.device}; pypic3d_b200 has no CPU path (CUDA tensors only)")
    if not t.is_contiguous():
        raise PicError(f"{name}: tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise PicError(f"{name}: dtype {t.dtype} != {dtype}")
    return t


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _v(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def np_dtype(t):
    return np.float32 if t.dtype == F32 else np.float64


def params_for(static_parameters, dynamic_parameters, species_config, like, **kw):
    if like.dtype not in (F32, F64):
        raise PicError(f"unsupported dtype {like.dtype}")
    return make_params(static_parameters, dynamic_parameters, species_config, np_dtype(like), **kw)


def _check_particles(p, x, u, active):
    _chk(x, "particles.x"); _chk(u, "particles.u", x.dtype); _chk(active, "particles.active")
    if active.dtype not in (torch.bool, torch.uint8):
        raise PicError("particles.active must be bool")
    mesh = tuple(p.mesh)
    if tuple(x.shape[:3]) != mesh or x.shape[-1] != 3 or x.shape[3] != p.n_species or tuple(active.shape) != tuple(x.shape[:-1]):
        raise ValueError("Tiled particle communication requires one logical particle tile per device: "
                         f"particle tile topology {tuple(x.shape[:3])} does not match device mesh {mesh}.")
    return int(x.shape[4])


def _check_field(p, f, name, dtype):
    _chk(f, name, dtype)
    want = tuple(p.mesh) + tuple(p.tile[a] + 2 * p.g for a in range(3))
    if tuple(f.shape) != want:
        raise ValueError("Tiled field communication requires one logical tile per device: "
                         f"field tile topology {tuple(f.shape)} does not match {want}.")


# ------------------------------------------------------------------------------------------------ particles
def push(p, x, u, active, E, B):
    cap = _check_particles(p, x, u, active)
    for i, f in enumerate(tuple(E) + tuple(B)):
        _check_field(p, f, f"field[{i}]", x.dtype)
    out = torch.empty_like(u)
    check(_lib.lib().pic_push(ctypes.byref(p), _p(x), _p(u), _p(out), _p(active), cap, _v(E), _v(B), _stream()), "pic_push")
    return out


def deposit(p, mode, x, u, active, like_field):
    """mode 'esirkepov' | 'direct' -> 3 new tiled arrays with raw (unfolded) deposits; 'rho' -> 1 array."""
    cap = _check_particles(p, x, u, active)
    _check_field(p, like_field, "J template", x.dtype)
    L = _lib.lib()
    if mode == "rho":
        rho = torch.zeros_like(like_field)
        check(L.pic_deposit_rho(ctypes.byref(p), _p(x), _p(active), cap, _p(rho), _stream()), "pic_deposit_rho")
        return rho
    J = [torch.zeros_like(like_field) for _ in range(3)]
    fn = L.pic_deposit_esirkepov if mode == "esirkepov" else L.pic_deposit_direct
    check(fn(ctypes.byref(p), _p(x), _p(u), _p(active), cap, _v(J), _stream()), f"pic_deposit_{mode}")
    return tuple(J)


def move(p, x, u, active, dt):
    cap = _check_particles(p, x, u, active)
    out = torch.empty_like(x)
    check(_lib.lib().pic_move(ctypes.byref(p), _p(x), _p(out), _p(u), _p(active), cap, float(dt), _stream()), "pic_move")
    return out


def retile(p, x, u, active):
    cap = _check_particles(p, x, u, active)
    xo, uo = torch.empty_like(x), torch.empty_like(u)
    ao = torch.empty_like(active)
    scratch = torch.empty(2 * active.numel() + 1, dtype=torch.int32, device=x.device)
    overflow = torch.zeros(1, dtype=torch.int32, device=x.device)
    check(_lib.lib().pic_retile(ctypes.byref(p), _p(x), _p(u), _p(active), _p(xo), _p(uo), _p(ao), cap, _p(scratch), _p(overflow),
                                _stream()), "pic_retile")
    return xo, uo, ao, overflow[0] != 0


def particle_energy(p, u, active):
    cap = int(u.shape[4])
    out = torch.zeros(2, dtype=F64, device=u.device)
    check(_lib.lib().pic_particle_energy(ctypes.byref(p), _p(_chk(u, "u")), _p(_chk(active, "active")), cap, _p(out), _stream()),
          "pic_particle_energy")
    return out


# ------------------------------------------------------------------------------------------------ fields
def update_E_(p, E, B, J):
    check(_lib.lib().pic_update_E(ctypes.byref(p), _v(E), _v(B), _v(J), _stream()), "pic_update_E")


def update_B_(p, B, E):
    check(_lib.lib().pic_update_B(ctypes.byref(p), _v(B), _v(E), _stream()), "pic_update_B")


def filter27(p, kind, alpha, f, out=None):
    out = torch.empty_like(f) if out is None else out
    check(_lib.lib().pic_filter(ctypes.byref(p), 0 if kind == "digital" else 1, float(alpha), _p(f), _p(out), _stream()), "pic_filter")
    return out


def halo_refresh_(p, fields, bcs):
    """In-place axis-sequential refresh x -> y -> z (ghost_cells.py:199-215) over the local tile mesh."""
    L = _lib.lib()
    ptrs = _v(fields)
    for axis in range(3):
        check(L.pic_halo_refresh_axis(ctypes.byref(p), axis, int(bcs[axis]), len(fields), ptrs, _stream()), "pic_halo_refresh_axis")


def halo_refresh_axis_(p, axis, bc, fields):
    check(_lib.lib().pic_halo_refresh_axis(ctypes.byref(p), int(axis), int(bc), len(fields), _v(fields), _stream()), "pic_halo_refresh_axis")


def halo_fold_axis_(p, axis, bc, fields):
    check(_lib.lib().pic_halo_fold_axis(ctypes.byref(p), int(axis), int(bc), len(fields), _v(fields), _stream()), "pic_halo_fold_axis")


def halo_fold_(p, fields, bcs):
    L = _lib.lib()
    ptrs = _v(fields)
    for axis in range(3):
        check(L.pic_halo_fold_axis(ctypes.byref(p), axis, int(bcs[axis]), len(fields), ptrs, _stream()), "pic_halo_fold_axis")


def zero_wall_(p, f, axis):
    check(_lib.lib().pic_zero_wall(ctypes.byref(p), int(axis), _p(f), _stream()), "pic_zero_wall")


def div_residual(p, F, a=None, ca=0.0, b=None, cb=0.0, out=None):
    """out[interior] = div_backward(F) + ca*a + cb*b (pic_div_residual); ghosts of `out` are zero."""
    if out is None:
        out = torch.zeros_like(F[0])
    check(_lib.lib().pic_div_residual(ctypes.byref(p), _v(F), _p(a) if a is not None else None, float(ca),
                                      _p(b) if b is not None else None, float(cb), _p(out), _stream()), "pic_div_residual")
    return out


def sum_squares_interior(p, f, out):
    check(_lib.lib().pic_sum_squares_interior(ctypes.byref(p), _p(f), _p(out), _stream()), "pic_sum_squares_interior")


def pack_planes(p, axis, start, nplanes, fields, buf):
    check(_lib.lib().pic_pack_planes(ctypes.byref(p), axis, start, nplanes, len(fields), _v(fields), _p(buf), _stream()), "pic_pack_planes")


def _box_arrays(boxes):
    n = len(boxes)
    lo = (ctypes.c_int32 * (3 * n))(*[int(v) for b in boxes for v in b[0]])
    sz = (ctypes.c_int32 * (3 * n))(*[int(v) for b in boxes for v in b[1]])
    return n, lo, sz


def pack_boxes(p, boxes, fields, buf):
    """boxes: [((lo_x, lo_y, lo_z), (n_x, n_y, n_z))] in array indices of the ghosted tile; buf receives them back to back."""
    n, lo, sz = _box_arrays(boxes)
    check(_lib.lib().pic_pack_boxes(ctypes.byref(p), n, lo, sz, len(fields), _v(fields), _p(buf), _stream()), "pic_pack_boxes")


def unpack_boxes_(p, boxes, fields, buf, mode):
    n, lo, sz = _box_arrays(boxes)
    check(_lib.lib().pic_unpack_boxes(ctypes.byref(p), n, lo, sz, len(fields), _v(fields), _p(buf), int(mode), _stream()), "pic_unpack_boxes")


def unpack_planes_(p, axis, start, nplanes, fields, buf, mode):
    check(_lib.lib().pic_unpack_planes(ctypes.byref(p), axis, start, nplanes, len(fields), _v(fields), _p(buf), int(mode), _stream()),
          "pic_unpack_planes")
