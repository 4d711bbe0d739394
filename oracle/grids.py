"""Grid construction (test infrastructure).  Follows PyPIC3D/boundary_conditions/grid_and_stencil.py:242-286
(build_collocated_axis / build_staggered_axis) and PyPIC3D/utilities/grids.py:42-165."""
import numpy as np


def build_collocated_axis(minimum_physical, spacing, count):
    # grid_and_stencil.py:263-265
    start = minimum_physical - spacing
    stop = minimum_physical + count * spacing
    return np.linspace(start, stop, int(count) + 2)


def build_staggered_axis(minimum_physical, spacing, count):
    # grid_and_stencil.py:284-286
    start = minimum_physical - 0.5 * spacing
    stop = minimum_physical + (count + 0.5) * spacing
    return np.linspace(start, stop, int(count) + 2)


def build_yee_grid(dp):
    # utilities/grids.py:42-77 ("center" = collocated nodes, "vertex" = staggered half-cell)
    center = (build_collocated_axis(-dp.x_wind / 2, dp.dx, dp.Nx),
              build_collocated_axis(-dp.y_wind / 2, dp.dy, dp.Ny),
              build_collocated_axis(-dp.z_wind / 2, dp.dz, dp.Nz))
    vertex = (build_staggered_axis(-dp.x_wind / 2, dp.dx, dp.Nx),
              build_staggered_axis(-dp.y_wind / 2, dp.dy, dp.Ny),
              build_staggered_axis(-dp.z_wind / 2, dp.dz, dp.Nz))
    return center, vertex


def build_collocated_grid(dp):
    # utilities/grids.py:14-40
    grid = (build_collocated_axis(-dp.x_wind / 2, dp.dx, dp.Nx),
            build_collocated_axis(-dp.y_wind / 2, dp.dy, dp.Ny),
            build_collocated_axis(-dp.z_wind / 2, dp.dz, dp.Nz))
    return grid, grid


def tile_grid_axis(global_axis_grid, d, tile_width, tile_count, g):
    """(tile_count, W+2g) coordinate lines.  utilities/grids.py:80-111:
    line[t, l] = grid[0] + (l + t*W - (g-1)) * d."""
    offsets = np.arange(tile_width + 2 * g, dtype=np.float64)
    tiles = np.arange(tile_count, dtype=np.float64)
    return global_axis_grid[0] + (offsets[None, :] + tiles[:, None] * tile_width - (g - 1)) * d


def build_tiled_yee_grids(sp, dp):
    """Returns (tiled_center, tiled_vertex); each a 3-tuple of (nt_axis, L_axis) arrays.
    (The reference broadcasts these to (ntx,nty,ntz,L); only the own-axis tile index matters,
    utilities/grids.py:107-111.)"""
    g = int(sp.guard_cells)
    ds = (dp.dx, dp.dy, dp.dz)
    out = []
    for grid in (dp.grids.center, dp.grids.vertex):
        axes = []
        for a in range(3):
            n = int(grid[a].shape[0]) - 2
            w = int(sp.tile_shape[a])
            if n % w != 0:
                raise ValueError("Shared tile sizes must divide the physical grid dimensions exactly.")
            axes.append(tile_grid_axis(grid[a], ds[a], w, n // w, g))
        out.append(tuple(axes))
    return out[0], out[1]
