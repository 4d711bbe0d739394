"""Drop-in for PyPIC3D/deposition/Esirkepov.py:49 Esirkepov_current."""
from .. import ops
from ..boundary_conditions.ghost_cells import fold_tiled_vector_ghost_cells, update_tiled_vector_ghost_cells


def Esirkepov_current(particles, species_config, J, static_parameters, dynamic_parameters):
    p = ops.params_for(static_parameters, dynamic_parameters, species_config, particles.x)
    g = int(static_parameters.guard_cells)
    Jt = ops.deposit(p, "esirkepov", particles.x, particles.u, particles.active, J[0])
    Jt = fold_tiled_vector_ghost_cells(Jt, static_parameters, num_guard_cells=g, bc_type=1, _inplace=True, _dyn=dynamic_parameters)
    return update_tiled_vector_ghost_cells(Jt, static_parameters, num_guard_cells=g, bc_type=1, _inplace=True, _dyn=dynamic_parameters)
