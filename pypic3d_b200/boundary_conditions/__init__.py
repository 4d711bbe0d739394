from .ghost_cells import (  # noqa: F401
    update_tiled_ghost_cells, update_tiled_vector_ghost_cells, fold_tiled_ghost_cells, fold_tiled_vector_ghost_cells,
    apply_tiled_zero_boundary, make_field_mesh, BC_TYPE_FIELD, BC_TYPE_PARTICLE,
)
from .grid_and_stencil import BC_PERIODIC, BC_CONDUCTING  # noqa: F401
