"""Drop-in for PyPIC3D/deposition/J_from_rhov.py:32 J_from_rhov (rho*v deposit + optional bilinear/digital filter)."""
from .. import ops
from ..boundary_conditions.ghost_cells import fold_tiled_vector_ghost_cells, update_tiled_vector_ghost_cells
from ..utilities.filters import bilinear_filter_vector, digital_filter_vector


def J_from_rhov(particles, species_config, J, static_parameters, dynamic_parameters):
    p = ops.params_for(static_parameters, dynamic_parameters, species_config, particles.x)
    g = int(static_parameters.guard_cells)
    kw = dict(_inplace=True, _dyn=dynamic_parameters)
    Jt = ops.deposit(p, "direct", particles.x, particles.u, particles.active, J[0])
    Jt = fold_tiled_vector_ghost_cells(Jt, static_parameters, g, bc_type=1, **kw)
    Jt = update_tiled_vector_ghost_cells(Jt, static_parameters, g, bc_type=1, **kw)
    if static_parameters.current_filter == "bilinear":
        Jt = bilinear_filter_vector(Jt, num_guard_cells=g, _sp=static_parameters)
        Jt = update_tiled_vector_ghost_cells(Jt, static_parameters, num_guard_cells=g, bc_type=1, **kw)
    elif static_parameters.current_filter == "digital":
        Jt = digital_filter_vector(Jt, dynamic_parameters.alpha, num_guard_cells=g, _sp=static_parameters)
        Jt = update_tiled_vector_ghost_cells(Jt, static_parameters, num_guard_cells=g, bc_type=1, **kw)
    return Jt
