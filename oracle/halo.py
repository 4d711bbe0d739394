"""Guard-cell refresh / fold / wall helpers on an in-process tile mesh (test infrastructure).
Follows PyPIC3D/boundary_conditions/ghost_cells.py:42-49,72-95,142-386,615-736.  `lax.ppermute` over the
named mesh axis is emulated by shifting along the leading tile axes; chain ends of a non-periodic axis
receive zeros (ghost_cells.py:189-191)."""
import numpy as np

from .stencil import BC_PERIODIC, BC_CONDUCTING
from .params import mesh_shape

BC_TYPE_FIELD = 0
BC_TYPE_PARTICLE = 1


def boundary_conditions_for_type(sp, bc_type):
    # ghost_cells.py:33-39
    if int(bc_type) == BC_TYPE_FIELD:
        return tuple(int(v) for v in sp.boundary_conditions)
    if int(bc_type) == BC_TYPE_PARTICLE:
        return tuple(int(v) for v in sp.particle_boundary_conditions)
    raise ValueError("bc_type must be 0 for field boundaries or 1 for particle boundaries.")


def reduced_axes(tile_shape, mesh):
    return tuple(int(tile_shape[a]) == 1 and int(mesh[a]) == 1 for a in range(3))  # :42-49


def _shift_tiles(values, tile_axis, direction, periodic):
    """Emulates lax.ppermute: result[dest] = values[source], dest = source + direction."""
    out = np.roll(values, direction, axis=tile_axis)
    if not periodic:
        idx = [slice(None)] * values.ndim
        idx[tile_axis] = 0 if direction > 0 else -1
        out[tuple(idx)] = 0.0
    return out


def _sl(axis, s):
    idx = [slice(None)] * 6
    idx[3 + axis] = s
    return tuple(idx)


def refresh(field_tiles, tile_shape, bcs, g):
    """Scalar refresh on (ntx,nty,ntz,Lx,Ly,Lz).  ghost_cells.py:142-215, x -> y -> z."""
    t = np.array(field_tiles, dtype=np.float64, copy=True)
    mesh = t.shape[:3]
    red = reduced_axes(tile_shape, mesh)
    for axis in range(3):
        bc = bcs[axis]
        lo_g, hi_g = _sl(axis, slice(0, g)), _sl(axis, slice(-g, None))
        if red[axis]:                                              # :142-159
            if bc == BC_PERIODIC:
                interior = t[_sl(axis, slice(g, g + 1))]
                t[lo_g] = np.broadcast_to(interior, t[lo_g].shape)
                t[hi_g] = np.broadcast_to(interior, t[hi_g].shape)
            else:
                t[lo_g] = 0.0
                t[hi_g] = 0.0
            continue
        lo_i, hi_i = _sl(axis, slice(g, 2 * g)), _sl(axis, slice(-2 * g, -g))
        lower_values = _shift_tiles(t[hi_i], axis, +1, bc == BC_PERIODIC)   # :187
        upper_values = _shift_tiles(t[lo_i], axis, -1, bc == BC_PERIODIC)   # :191
        t[lo_g] = lower_values
        t[hi_g] = upper_values
    return t


def fold(field_tiles, tile_shape, bcs, g):
    """Scalar fold-add on (ntx,nty,ntz,Lx,Ly,Lz).  ghost_cells.py:218-316."""
    t = np.array(field_tiles, dtype=np.float64, copy=True)
    mesh = t.shape[:3]
    red = reduced_axes(tile_shape, mesh)
    for axis in range(3):
        bc = bcs[axis]
        lo_g, hi_g = _sl(axis, slice(0, g)), _sl(axis, slice(-g, None))
        if red[axis]:                                              # :218-235
            ghost_sum = np.sum(t[lo_g], axis=3 + axis, keepdims=True) + np.sum(t[hi_g], axis=3 + axis, keepdims=True)
            interior = _sl(axis, slice(g, g + 1))
            if bc == BC_PERIODIC:
                t[interior] += ghost_sum
            elif bc == BC_CONDUCTING:
                t[interior] -= ghost_sum
            t[lo_g] = 0.0
            t[hi_g] = 0.0
            continue
        lo_i, hi_i = _sl(axis, slice(g, 2 * g)), _sl(axis, slice(-2 * g, -g))
        lower_values = t[lo_g].copy()
        upper_values = t[hi_g].copy()
        from_positive = _shift_tiles(lower_values, axis, -1, bc == BC_PERIODIC)   # :271
        from_negative = _shift_tiles(upper_values, axis, +1, bc == BC_PERIODIC)   # :272
        t[hi_i] += from_positive
        t[lo_i] += from_negative
        if bc == BC_CONDUCTING:                                     # :238-260 (global-wall tiles only)
            first = [slice(None)] * 6
            first[axis] = slice(0, 1)
            last = [slice(None)] * 6
            last[axis] = slice(mesh[axis] - 1, mesh[axis])
            tgt = list(lo_i); tgt[axis] = slice(0, 1)
            src = list(lo_g); src[axis] = slice(0, 1)
            t[tuple(tgt)] -= lower_values[tuple(first)]
            tgt = list(hi_i); tgt[axis] = slice(mesh[axis] - 1, mesh[axis])
            t[tuple(tgt)] -= upper_values[tuple(last)]
        t[lo_g] = 0.0
        t[hi_g] = 0.0
    return t


def update_tiled_ghost_cells(field_tiles, sp, num_guard_cells=2, bc_type=BC_TYPE_FIELD):
    return refresh(field_tiles, sp.tile_shape, boundary_conditions_for_type(sp, bc_type), int(num_guard_cells))  # :615


def update_tiled_vector_ghost_cells(field_tiles, sp, num_guard_cells=2, bc_type=BC_TYPE_FIELD):
    return tuple(update_tiled_ghost_cells(c, sp, num_guard_cells, bc_type) for c in field_tiles)  # :637


def fold_tiled_ghost_cells(field_tiles, sp, num_guard_cells=2, bc_type=BC_TYPE_FIELD):
    return fold(field_tiles, sp.tile_shape, boundary_conditions_for_type(sp, bc_type), int(num_guard_cells))  # :703


def fold_tiled_vector_ghost_cells(field_tiles, sp, num_guard_cells=2, bc_type=BC_TYPE_FIELD):
    return tuple(fold_tiled_ghost_cells(c, sp, num_guard_cells, bc_type) for c in field_tiles)  # :723


def apply_tiled_zero_boundary(field_tiles, sp, axis, num_guard_cells=2):
    """ghost_cells.py:653-672 (+ :344-362): zero planes g and -g-1 on global-wall tiles, then refresh."""
    axis = int(axis)
    g = int(num_guard_cells)
    if int(sp.boundary_conditions[axis]) != BC_CONDUCTING:
        return update_tiled_ghost_cells(field_tiles, sp, g)
    t = np.array(field_tiles, dtype=np.float64, copy=True)
    lo = [slice(None)] * 6; lo[axis] = 0; lo[3 + axis] = g
    hi = [slice(None)] * 6; hi[axis] = t.shape[axis] - 1; hi[3 + axis] = -g - 1
    t[tuple(lo)] = 0.0
    t[tuple(hi)] = 0.0
    return update_tiled_ghost_cells(t, sp, g)


def apply_tiled_constant_boundary(field_tiles, sp, axis, num_guard_cells=2):
    """ghost_cells.py:675-700 (+ :365-386)."""
    axis = int(axis)
    g = int(num_guard_cells)
    t = update_tiled_ghost_cells(field_tiles, sp, g)
    if int(sp.boundary_conditions[axis]) != BC_CONDUCTING:
        return t
    lo_g = [slice(None)] * 6; lo_g[axis] = slice(0, 1); lo_g[3 + axis] = slice(0, g)
    lo_i = [slice(None)] * 6; lo_i[axis] = slice(0, 1); lo_i[3 + axis] = slice(g, g + 1)
    n = t.shape[axis]
    hi_g = [slice(None)] * 6; hi_g[axis] = slice(n - 1, n); hi_g[3 + axis] = slice(-g, None)
    hi_i = [slice(None)] * 6; hi_i[axis] = slice(n - 1, n); hi_i[3 + axis] = slice(-g - 1, -g)
    t[tuple(lo_g)] = np.broadcast_to(t[tuple(lo_i)], t[tuple(lo_g)].shape)
    t[tuple(hi_g)] = np.broadcast_to(t[tuple(hi_i)], t[tuple(hi_g)].shape)
    return t
