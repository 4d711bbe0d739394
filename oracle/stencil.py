"""Anchor / offset / wrap helpers (test infrastructure).
Follows PyPIC3D/boundary_conditions/grid_and_stencil.py:15-239."""
import numpy as np

BC_PERIODIC = 0
BC_CONDUCTING = 1


def wrap_periodic_position(x, wind):
    # grid_and_stencil.py:31-35
    x = np.asarray(x, dtype=np.float64)
    h = 0.5 * wind
    w = np.mod(x + h, wind) - h
    return np.where((w == -h) & (x >= h), h, w)


def axis_has_active_cells(axis_size, ghost_cells=False):
    return axis_size > (3 if ghost_cells else 1)  # :54


def inactive_axis_index(axis_size, ghost_cells=False):
    return 1 if ghost_cells and axis_size >= 3 else axis_size // 2  # :73


def uniform_axis_spacing(axis):
    return axis[1] - axis[0] if len(axis) > 1 else 1.0  # :89


def compute_particle_anchor(position, grid_axis, shape_factor):
    # :109-118 ; floor for CIC, round-half-even for TSC (jnp.round == np.rint)
    s = uniform_axis_spacing(grid_axis)
    o = grid_axis[0]
    q = (np.asarray(position) - o) / s
    return (np.floor(q) if int(shape_factor) == 1 else np.rint(q)).astype(np.int64)


def particle_axis_offset(position, anchor, grid_axis):
    s = uniform_axis_spacing(grid_axis)  # :137-138
    return position - (anchor * s + grid_axis[0])


def build_axis_stencil_points(anchor, axis_size, bc, offsets):
    st = np.asarray(anchor)[None, ...] + np.asarray(offsets)[:, None]  # :160-166
    return np.mod(st, axis_size) if bc == BC_PERIODIC else st


def collapse_axis_stencil(points, weights, axis_size, ghost_cells=False):
    if axis_has_active_cells(axis_size, ghost_cells):  # :189-201
        return points, weights
    idx = inactive_axis_index(axis_size, ghost_cells)
    return (np.full((1, points.shape[1]), idx, dtype=points.dtype),
            np.sum(weights, axis=0, keepdims=True))


def prepare_particle_axis_stencil(position, grid_axis, axis_size, shape_factor, bc, wind=None, ghost_cells=False):
    anchor = compute_particle_anchor(position, grid_axis, shape_factor)  # :235-239
    offset = particle_axis_offset(position, anchor, grid_axis)
    points = build_axis_stencil_points(anchor, axis_size, bc, np.array([-1, 0, 1]))
    return position, anchor, offset, points
