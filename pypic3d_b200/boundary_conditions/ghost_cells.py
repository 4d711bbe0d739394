"""Drop-ins for the six ghost-cell wrappers of PyPIC3D/boundary_conditions/ghost_cells.py:615-736.

The reference maps one tile per JAX device and exchanges faces with `lax.ppermute` inside `shard_map`; here every tile
of the local mesh is resident on one GPU and the exchange is a per-axis CUDA kernel (x -> y -> z, full transverse
extent, so edges/corners propagate exactly as in ghost_cells.py:199-215).  Across GPUs see `pypic3d_b200.distributed`."""
import torch

from .. import ops
from ..parameters import make_field_mesh  # noqa: F401  (re-export, ghost_cells.py:111)
from .grid_and_stencil import BC_CONDUCTING

BC_TYPE_FIELD = 0
BC_TYPE_PARTICLE = 1


def _boundary_conditions_for_type(static_parameters, bc_type):
    bc_type = int(bc_type)
    if bc_type == BC_TYPE_FIELD:
        return tuple(int(v) for v in static_parameters.boundary_conditions)
    if bc_type == BC_TYPE_PARTICLE:
        return tuple(int(v) for v in static_parameters.particle_boundary_conditions)
    raise ValueError("bc_type must be 0 for field boundaries or 1 for particle boundaries.")


def _halo_params(field, static_parameters, g):
    """Minimal parameter block for the grid kernels (geometry only)."""
    p = ops._lib.PicParams()
    if field.dtype not in (ops.F32, ops.F64):
        raise ops.PicError(f"unsupported dtype {field.dtype}")
    p.dtype = 0 if field.dtype == ops.F32 else 1
    p.g = int(g)
    tile = [int(w) for w in static_parameters.tile_shape]
    mesh = getattr(static_parameters, "field_mesh", None)
    mesh = tuple(int(v) for v in mesh) if isinstance(mesh, (tuple, list)) else tuple(int(v) for v in field.shape[:3])
    if field.ndim != 6 or tuple(field.shape[:3]) != mesh:
        raise ValueError("Tiled field communication requires one logical tile per device: "
                         f"field tile topology {tuple(field.shape[:3])} does not match device mesh {mesh}.")
    for a in range(3):
        p.mesh[a] = p.gmesh[a] = mesh[a]
        p.tile[a] = tile[a]
        if field.shape[3 + a] != tile[a] + 2 * p.g:
            raise ValueError(f"tile extent {tuple(field.shape[3:])} does not match tile_shape {tuple(tile)} with {p.g} guard cells")
    return p


def _is_stacked(field_tiles):
    return isinstance(field_tiles, torch.Tensor) and field_tiles.ndim == 7 and int(field_tiles.shape[0]) == 3


def _components(field_tiles, inplace):
    if _is_stacked(field_tiles):
        t = ops._chk(field_tiles, "field") if inplace else ops._chk(field_tiles, "field").clone()
        return [t[0], t[1], t[2]], t
    comps = [ops._chk(c, "field") if inplace else ops._chk(c, "field").clone() for c in field_tiles]
    return comps, None


def update_tiled_ghost_cells(field_tiles, static_parameters, num_guard_cells=2, bc_type=BC_TYPE_FIELD, _inplace=False, _dyn=None):
    f = ops._chk(field_tiles, "field") if _inplace else ops._chk(field_tiles, "field").clone()
    p = _halo_params(f, static_parameters, num_guard_cells)
    ops.halo_refresh_(p, [f], _boundary_conditions_for_type(static_parameters, bc_type))
    return f


def update_tiled_vector_ghost_cells(field_tiles, static_parameters, num_guard_cells=2, bc_type=BC_TYPE_FIELD, _inplace=False, _dyn=None):
    comps, stacked = _components(field_tiles, _inplace)
    p = _halo_params(comps[0], static_parameters, num_guard_cells)
    ops.halo_refresh_(p, comps, _boundary_conditions_for_type(static_parameters, bc_type))
    return stacked if stacked is not None else tuple(comps)


def fold_tiled_ghost_cells(field_tiles, static_parameters, num_guard_cells=2, bc_type=BC_TYPE_FIELD, _inplace=False, _dyn=None):
    f = ops._chk(field_tiles, "field") if _inplace else ops._chk(field_tiles, "field").clone()
    p = _halo_params(f, static_parameters, num_guard_cells)
    ops.halo_fold_(p, [f], _boundary_conditions_for_type(static_parameters, bc_type))
    return f


def fold_tiled_vector_ghost_cells(field_tiles, static_parameters, num_guard_cells=2, bc_type=BC_TYPE_FIELD, _inplace=False, _dyn=None):
    comps, stacked = _components(field_tiles, _inplace)
    p = _halo_params(comps[0], static_parameters, num_guard_cells)
    ops.halo_fold_(p, comps, _boundary_conditions_for_type(static_parameters, bc_type))
    return stacked if stacked is not None else tuple(comps)


def apply_tiled_zero_boundary(field_tiles, static_parameters, axis, num_guard_cells=2, _inplace=False):
    """ghost_cells.py:653-672: zero the global conducting-wall planes of `axis`, then refresh."""
    axis = int(axis)
    f = ops._chk(field_tiles, "field") if _inplace else ops._chk(field_tiles, "field").clone()
    if int(static_parameters.boundary_conditions[axis]) == BC_CONDUCTING:
        p = _halo_params(f, static_parameters, num_guard_cells)
        ops.zero_wall_(p, f, axis)
    return update_tiled_ghost_cells(f, static_parameters, num_guard_cells, _inplace=True)
