"""Parameter / state containers with the reference's field names (test infrastructure).

Follows PyPIC3D/parameters.py:7-53 (GridParameters, StaticParameters, DynamicParameters)
and PyPIC3D/particles/particle_class.py:5-31 (SpeciesConfig, TiledParticles).  Leaves are
NumPy arrays / Python scalars; `field_mesh` is just the tile-grid shape tuple because the
oracle emulates the one-tile-per-device mesh inside one process.
"""
from typing import NamedTuple


class GridParameters(NamedTuple):
    vertex: tuple
    center: tuple
    tiled_vertex_grid: tuple
    tiled_center_grid: tuple


class StaticParameters(NamedTuple):
    name: str
    output_dir: str
    Nt: int
    verbose: bool
    GPUs: bool
    benchmark: bool
    solver: str
    electrostatic: bool
    relativistic: bool
    particle_pusher: str
    current_deposition: str
    current_filter: str
    shape_factor: int
    guard_cells: int
    tile_shape: tuple
    particle_tile_capacity_factor: float
    pml_active: bool
    boundary_conditions: tuple
    particle_boundary_conditions: tuple
    field_mesh: object


class DynamicParameters(NamedTuple):
    dt: float
    dx: float
    dy: float
    dz: float
    Nx: int
    Ny: int
    Nz: int
    x_wind: float
    y_wind: float
    z_wind: float
    C: float
    eps: float
    mu: float
    kb: float
    alpha: float
    grids: GridParameters


class SpeciesConfig(NamedTuple):
    charge: object    # (S,)
    mass: object      # (S,)
    weight: object    # (S,)
    update_x: object  # (S,3) bool
    update_u: object  # (S,3) bool


class TiledParticles(NamedTuple):
    x: object       # (ntx,nty,ntz,S,cap,3)
    u: object       # (ntx,nty,ntz,S,cap,3)  -- the *velocity* v, not gamma*v (particle_push.py:50-52)
    active: object  # (ntx,nty,ntz,S,cap) bool


def mesh_shape(static_parameters, dynamic_parameters=None, like=None):
    """Tile-grid shape (ntx,nty,ntz).  Reference: ghost_cells.py:98-116 (mesh.devices.shape)."""
    if like is not None:
        return tuple(int(v) for v in like.shape[:3])
    fm = static_parameters.field_mesh
    if isinstance(fm, (tuple, list)):
        return tuple(int(v) for v in fm)
    tw = static_parameters.tile_shape
    return (int(dynamic_parameters.Nx) // int(tw[0]),
            int(dynamic_parameters.Ny) // int(tw[1]),
            int(dynamic_parameters.Nz) // int(tw[2]))
