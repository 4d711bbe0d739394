"""Drop-ins for PyPIC3D/solvers/first_order_yee.py:12 update_E and :96 update_B (pml_state must be None: the PML branch
is out of scope)."""
from .. import ops
from ..boundary_conditions import ghost_cells
from ..boundary_conditions.grid_and_stencil import BC_CONDUCTING
from ..utilities.filters import digital_filter_vector


def _yee_params(field, static_parameters, dynamic_parameters):
    p = ghost_cells._halo_params(field, static_parameters, static_parameters.guard_cells)
    sc = ops._lib._scalar
    p.dt, p.dx, p.dy, p.dz = float(sc(dynamic_parameters.dt)), float(sc(dynamic_parameters.dx)), float(sc(dynamic_parameters.dy)), float(sc(dynamic_parameters.dz))
    p.C, p.eps = float(sc(dynamic_parameters.C)), float(sc(dynamic_parameters.eps))
    return p


def update_E(E_tiles, B_tiles, J_tiles, static_parameters, dynamic_parameters, pml_state=None):
    if pml_state is not None:
        raise NotImplementedError("PML is outside the hot path of pypic3d_b200")
    g = int(static_parameters.guard_cells)
    alpha = float(ops._lib._scalar(dynamic_parameters.alpha))
    B = ghost_cells.update_tiled_vector_ghost_cells(tuple(B_tiles), static_parameters, g)            # :31
    E = [ops._chk(c, "E").clone() for c in E_tiles]
    p = _yee_params(E[0], static_parameters, dynamic_parameters)
    ops.update_E_(p, E, list(B), [ops._chk(c, "J", E[0].dtype) for c in J_tiles])                      # :42-72
    E = ghost_cells.update_tiled_vector_ghost_cells(tuple(E), static_parameters, g, _inplace=True)   # :74
    if alpha != 1.0:                                                                                 # :78 (identity when alpha == 1)
        E = digital_filter_vector(E, alpha, num_guard_cells=g)
    Ex, Ey, Ez = E
    bc_x, bc_y, bc_z = (int(b) for b in static_parameters.boundary_conditions)
    z = lambda f, axis: ghost_cells.apply_tiled_zero_boundary(f, static_parameters, axis=axis, num_guard_cells=g, _inplace=True)
    if bc_x == BC_CONDUCTING:                                                                        # :80-89
        Ey, Ez = z(Ey, 0), z(Ez, 0)
    if bc_y == BC_CONDUCTING:
        Ex, Ez = z(Ex, 1), z(Ez, 1)
    if bc_z == BC_CONDUCTING:
        Ex, Ey = z(Ex, 2), z(Ey, 2)
    return ghost_cells.update_tiled_vector_ghost_cells((Ex, Ey, Ez), static_parameters, g, _inplace=True), pml_state


def update_B(E_tiles, B_tiles, static_parameters, dynamic_parameters, pml_state=None, do_filter=False):
    if pml_state is not None:
        raise NotImplementedError("PML is outside the hot path of pypic3d_b200")
    g = int(static_parameters.guard_cells)
    alpha = float(ops._lib._scalar(dynamic_parameters.alpha))
    E = ghost_cells.update_tiled_vector_ghost_cells(tuple(E_tiles), static_parameters, g)            # :115
    B = [ops._chk(c, "B").clone() for c in B_tiles]
    p = _yee_params(B[0], static_parameters, dynamic_parameters)
    ops.update_B_(p, B, list(E))                                                                     # :116-142 (dt/2)
    B = tuple(B)
    if do_filter and alpha != 1.0:                                                                   # :145-159
        B = ghost_cells.update_tiled_vector_ghost_cells(B, static_parameters, g, _inplace=True)
        B = digital_filter_vector(B, alpha, num_guard_cells=g)
    return ghost_cells.update_tiled_vector_ghost_cells(B, static_parameters, g, _inplace=True), pml_state
