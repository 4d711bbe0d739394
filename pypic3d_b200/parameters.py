"""Parameter pytrees with the reference's names and fields (PyPIC3D/parameters.py:7-203).

Leaves are Python scalars / NumPy arrays (host metadata); `field_mesh` is the tile-mesh shape tuple -- there is no JAX
device mesh here: on one GPU every tile of the mesh is resident, across GPUs `pypic3d_b200.distributed` maps one tile per
rank (the reference's one-tile-per-device contract, ghost_cells.py:98-116)."""
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


def _axis_tuple(axis_values):
    if isinstance(axis_values, (tuple, list)):
        return tuple(int(v) for v in axis_values)
    return (int(axis_values["x"]), int(axis_values["y"]), int(axis_values["z"]))


def _tile_shape(static_config):
    if "tile_shape" in static_config:
        return tuple(int(w) for w in static_config["tile_shape"])
    return (int(static_config["particle_tile_nx"]), int(static_config["particle_tile_ny"]), int(static_config["particle_tile_nz"]))


def make_field_mesh(tile_grid_shape):
    """ghost_cells.py:111-116.  Returns the validated tile-mesh shape."""
    shape = tuple(int(w) for w in tile_grid_shape)
    if len(shape) != 3 or any(w < 1 for w in shape):
        raise ValueError(f"invalid tile topology {shape}")
    return shape


def build_static_parameters(static_config):
    """parameters.py:89-123."""
    tile_shape = _tile_shape(static_config)
    if static_config.get("field_mesh") is not None:
        mesh = static_config["field_mesh"]
    else:
        mesh = make_field_mesh((int(static_config["Nx"]) // tile_shape[0], int(static_config["Ny"]) // tile_shape[1],
                                int(static_config["Nz"]) // tile_shape[2]))
    return StaticParameters(
        name=static_config.get("name", "Default Simulation"), output_dir=static_config.get("output_dir", "."),
        Nt=int(static_config.get("Nt", 0)), verbose=bool(static_config.get("verbose", False)),
        GPUs=bool(static_config.get("GPUs", False)), benchmark=bool(static_config.get("benchmark", False)),
        solver=static_config.get("solver", "electrodynamic_yee"), electrostatic=bool(static_config.get("electrostatic", False)),
        relativistic=bool(static_config.get("relativistic", True)), particle_pusher=static_config.get("particle_pusher", "boris"),
        current_deposition=static_config.get("current_deposition", "direct"), current_filter=static_config.get("current_filter", "none"),
        shape_factor=int(static_config["shape_factor"]), guard_cells=int(static_config["guard_cells"]), tile_shape=tile_shape,
        particle_tile_capacity_factor=float(static_config.get("particle_tile_capacity_factor", 1.0)),
        pml_active=bool(static_config.get("pml_active", False)),
        boundary_conditions=_axis_tuple(static_config["boundary_conditions"]),
        particle_boundary_conditions=_axis_tuple(static_config.get("particle_boundary_conditions", {"x": 0, "y": 0, "z": 0})),
        field_mesh=mesh)


def build_dynamic_parameters(dynamic_config, extra_dynamic_config=None):
    """parameters.py:126-161."""
    extra = extra_dynamic_config or {}
    grids = dynamic_config["grids"]
    if hasattr(grids, "_asdict"):
        grids = grids._asdict()
    grids = GridParameters(vertex=grids["vertex"], center=grids["center"], tiled_vertex_grid=grids["tiled_vertex_grid"],
                           tiled_center_grid=grids["tiled_center_grid"])
    g = lambda k: float(dynamic_config.get(k, extra.get(k, 1.0)))
    return DynamicParameters(
        dt=float(dynamic_config["dt"]), dx=float(dynamic_config["dx"]), dy=float(dynamic_config["dy"]), dz=float(dynamic_config["dz"]),
        Nx=int(dynamic_config["Nx"]), Ny=int(dynamic_config["Ny"]), Nz=int(dynamic_config["Nz"]),
        x_wind=float(dynamic_config["x_wind"]), y_wind=float(dynamic_config["y_wind"]), z_wind=float(dynamic_config["z_wind"]),
        C=g("C"), eps=g("eps"), mu=g("mu"), kb=g("kb"), alpha=g("alpha"), grids=grids)


def _output_value(value):
    if hasattr(value, "tolist"):
        return value.tolist()
    if isinstance(value, tuple):
        return tuple(_output_value(v) for v in value)
    if isinstance(value, list):
        return [_output_value(v) for v in value]
    if isinstance(value, dict):
        return {k: _output_value(v) for k, v in value.items()}
    return value


def static_parameters_for_output(static_parameters):
    return {k: _output_value(v) for k, v in static_parameters._asdict().items() if k != "field_mesh"}


def dynamic_parameters_for_output(dynamic_parameters):
    return {k: _output_value(v) for k, v in dynamic_parameters._asdict().items() if k != "grids"}


def boundary_dict(static_parameters):
    bx, by, bz = static_parameters.boundary_conditions
    return {"x": bx, "y": by, "z": bz}


def particle_boundary_dict(static_parameters):
    bx, by, bz = static_parameters.particle_boundary_conditions
    return {"x": bx, "y": by, "z": bz}
