"""Host-side grid builders (PyPIC3D/utilities/grids.py:42-165).  Pure index math; the kernels recompute the per-tile
origin/spacing from `grids.center[a][0]`, `grids.vertex[a][0]` and dx,dy,dz."""
import numpy as np

from ..boundary_conditions.grid_and_stencil import build_collocated_axis, build_staggered_axis


def build_collocated_grid(dp):
    grid = (build_collocated_axis(-dp.x_wind / 2, dp.dx, dp.Nx), build_collocated_axis(-dp.y_wind / 2, dp.dy, dp.Ny),
            build_collocated_axis(-dp.z_wind / 2, dp.dz, dp.Nz))
    return grid, grid


def build_yee_grid(dp):
    center = (build_collocated_axis(-dp.x_wind / 2, dp.dx, dp.Nx), build_collocated_axis(-dp.y_wind / 2, dp.dy, dp.Ny),
              build_collocated_axis(-dp.z_wind / 2, dp.dz, dp.Nz))
    vertex = (build_staggered_axis(-dp.x_wind / 2, dp.dx, dp.Nx), build_staggered_axis(-dp.y_wind / 2, dp.dy, dp.Ny),
              build_staggered_axis(-dp.z_wind / 2, dp.dz, dp.Nz))
    return center, vertex


def build_tiled_yee_grids(static_parameters, dynamic_parameters):
    g = int(static_parameters.guard_cells)
    ds = (dynamic_parameters.dx, dynamic_parameters.dy, dynamic_parameters.dz)
    out = []
    for grid in (dynamic_parameters.grids.center, dynamic_parameters.grids.vertex):
        axes = []
        for a in range(3):
            n = int(grid[a].shape[0]) - 2
            w = int(static_parameters.tile_shape[a])
            if n % w != 0:
                raise ValueError("Shared tile sizes must divide the physical grid dimensions exactly.")
            off = np.arange(w + 2 * g, dtype=np.float64)
            t = np.arange(n // w, dtype=np.float64)
            axes.append(grid[a][0] + (off[None, :] + t[:, None] * w - (g - 1)) * ds[a])
        out.append(tuple(axes))
    return out[0], out[1]
