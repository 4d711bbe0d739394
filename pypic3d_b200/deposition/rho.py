"""Drop-in for PyPIC3D/deposition/rho.py:30 compute_rho."""
from .. import ops
from ..boundary_conditions.ghost_cells import fold_tiled_ghost_cells, update_tiled_ghost_cells
from ..utilities.filters import digital_filter


def compute_rho(particles, species_config, rho, static_parameters, dynamic_parameters):
    p = ops.params_for(static_parameters, dynamic_parameters, species_config, particles.x)
    g = int(static_parameters.guard_cells)
    kw = dict(_inplace=True, _dyn=dynamic_parameters)
    r = ops.deposit(p, "rho", particles.x, particles.x, particles.active, rho)
    r = fold_tiled_ghost_cells(r, static_parameters, g, bc_type=1, **kw)
    r = update_tiled_ghost_cells(r, static_parameters, g, bc_type=1, **kw)
    if static_parameters.current_filter == "digital":
        r = digital_filter(r, dynamic_parameters.alpha, num_guard_cells=g, _sp=static_parameters)
        r = update_tiled_ghost_cells(r, static_parameters, g, bc_type=1, **kw)
    return r
