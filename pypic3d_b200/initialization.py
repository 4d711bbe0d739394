"""`initialize_simulation(config)` with the reference's TOML schema, defaults and return tuple
(PyPIC3D/initialization.py:138-488, utils.py:621-653, :540-618).  Host-side NumPy set-up; the returned particles / fields
are torch CUDA tensors.  Not reproduced (outside the hot path): PML,
openPMD / matplotlib output -- the corresponding keys are accepted and ignored with a notice."""
import os
from types import SimpleNamespace

import numpy as np
import torch

from .boundary_conditions.grid_and_stencil import BC_CONDUCTING, BC_PERIODIC
from .boundary_conditions.ghost_cells import update_tiled_vector_ghost_cells
from .evolve import time_loop_electrodynamic, time_loop_electrostatic
from .parameters import build_dynamic_parameters, build_static_parameters, make_field_mesh
from .particles.particle_class import TiledParticles
from .particles.particle_initialization import load_particles_from_toml
from .utilities.grids import build_collocated_grid, build_tiled_yee_grids, build_yee_grid
from .utils import courant_condition


def default_parameters():
    """initialization.py:138-208 (verbatim defaults)."""
    plotting = {"plotting": True, "save_data": False, "plotfields": False, "plotpositions": False, "plotvelocities": False,
                "plotenergy": True, "plotcurrent": False, "plasmaFreq": False, "plot_phasespace": False, "plot_errors": False,
                "plot_dispersion": False, "plot_chargeconservation": False, "plot_openpmd_particles": False,
                "plot_openpmd_fields": False, "plotting_interval": 10, "openpmd_field_queue_size": 2,
                "openpmd_particle_queue_size": 2, "dump_particles": False, "dump_fields": False}
    static = {"name": "Default Simulation", "output_dir": os.getcwd(), "solver": "electrodynamic_yee", "particle_x_bc": "periodic",
              "particle_y_bc": "periodic", "particle_z_bc": "periodic", "x_bc": "periodic", "y_bc": "periodic", "z_bc": "periodic",
              "Nt": None, "relativistic": True, "particle_pusher": "boris", "benchmark": False, "verbose": False, "GPUs": False,
              "cfl": 1.0, "ds_per_debye": None, "shape_factor": 1, "guard_cells": 2, "particle_tile_nx": None,
              "particle_tile_ny": None, "particle_tile_nz": None, "particle_tile_capacity_factor": 1.0,
              "current_calculation": "j_from_rhov", "filter_j": "bilinear"}
    dynamic = {"Nx": 30, "Ny": 30, "Nz": 30, "x_wind": 1e-2, "y_wind": 1e-2, "z_wind": 1e-2, "t_wind": 1e-12, "dt": None,
               "eps": 8.85418782e-12, "mu": 1.25663706e-6, "C": 2.99792458e8, "kb": 1.380649e-23, "alpha": 1.0}
    return plotting, static, dynamic


def update_parameters_from_toml(config, static_parameters, dynamic_parameters, plotting_parameters):
    """utils.py:621-653: keys are only taken if they already exist in the defaults; `[constants]` is ignored."""
    for key, value in config.get("simulation_parameters", {}).items():
        if key in static_parameters:
            static_parameters[key] = value
        if key in dynamic_parameters:
            dynamic_parameters[key] = value
    for key, value in config.get("static_parameters", {}).items():
        if key in static_parameters:
            static_parameters[key] = value
    for key, value in config.get("dynamic_parameters", {}).items():
        if key in dynamic_parameters:
            dynamic_parameters[key] = value
    for key, value in config.get("plotting", {}).items():
        if key in plotting_parameters:
            plotting_parameters[key] = value
    return static_parameters, dynamic_parameters, plotting_parameters


def _encode_field_bc(name):
    codes = {"periodic": BC_PERIODIC, "conducting": BC_CONDUCTING}
    if name not in codes:
        raise ValueError(f"Unsupported field boundary condition: {name}")
    return codes[name]


def _encode_particle_bc(name):
    codes = {"periodic": 0, "reflecting": 1, "absorbing": 2}
    if name not in codes:
        raise ValueError(f"Unsupported particle boundary condition: {name}")
    return codes[name]


def _validate(static_config, dynamic_config):
    """initialization.py:42-124."""
    if static_config["solver"] not in ("electrodynamic_yee", "electrostatic"):
        raise ValueError(f"Unsupported solver: {static_config['solver']}. Use 'electrodynamic_yee' or 'electrostatic'.")
    if static_config["current_calculation"] not in ("j_from_rhov", "esirkepov"):
        raise ValueError("Unsupported current_calculation. Use 'j_from_rhov' or 'esirkepov'.")
    if static_config["current_calculation"] == "esirkepov" and static_config["filter_j"] != "none":
        raise ValueError("Esirkepov current filtering is not supported; use filter_j='none'.")
    if static_config["filter_j"] not in ("none", "bilinear", "digital"):
        raise ValueError(f"Unsupported filter_j: {static_config['filter_j']}")
    if static_config["particle_pusher"] not in ("boris", "higuera_cary"):
        raise ValueError(f"Unknown particle_pusher: {static_config['particle_pusher']}")
    if int(static_config["shape_factor"]) not in (1, 2):
        raise ValueError("shape_factor must be 1 or 2")
    for n, w in zip(("Nx", "Ny", "Nz"), ("particle_tile_nx", "particle_tile_ny", "particle_tile_nz")):
        if int(dynamic_config[n]) % int(static_config[w]) != 0:
            raise ValueError("Shared tile sizes must divide the physical grid dimensions exactly.")


def _add_external_field(component, ext, sp, dp, name):
    """utils.py:508-537: add a physical (Nx,Ny,Nz) array into the tile interiors."""
    w = [int(v) for v in sp.tile_shape]
    g = int(sp.guard_cells)
    shape = (int(dp.Nx), int(dp.Ny), int(dp.Nz))
    if tuple(ext.shape) != shape:
        raise ValueError(f"Shape mismatch for field '{name}': external field shape {ext.shape} does not match expected interior shape {shape}")
    nt = component.shape[:3]
    for tx in range(nt[0]):
        for ty in range(nt[1]):
            for tz in range(nt[2]):
                component[tx, ty, tz, g:g + w[0], g:g + w[1], g:g + w[2]] += ext[tx * w[0]:(tx + 1) * w[0], ty * w[1]:(ty + 1) * w[1], tz * w[2]:(tz + 1) * w[2]]
    return component


def load_external_fields_from_toml(fields, external_fields, config, sp, dp):
    """utils.py:540-618.  `fields` = [Ex,Ey,Ez,Bx,By,Bz,Jx,Jy,Jz] NumPy tiles; type 0-8; evolve=false -> external-only."""
    ext_E, ext_B = [list(external_fields[0]), list(external_fields[1])]
    for key in [k for k in config.keys() if k[:5] == "field"]:
        blk = config[key]
        arr = np.asarray(np.load(blk["path"]), dtype=np.float64)
        ftype, evolve = int(blk["type"]), blk.get("evolve", True)
        if not evolve and (ftype < 0 or ftype > 5):
            raise ValueError("External-only fields must be electric or magnetic field components with type 0 through 5")
        if evolve:
            fields[ftype] = _add_external_field(fields[ftype], arr, sp, dp, blk["name"])
        elif ftype < 3:
            ext_E[ftype] = _add_external_field(ext_E[ftype], arr, sp, dp, blk["name"])
        else:
            ext_B[ftype - 3] = _add_external_field(ext_B[ftype - 3], arr, sp, dp, blk["name"])
    return fields, (tuple(ext_E), tuple(ext_B))


def initialize_simulation(toml_file, device=None, dtype=torch.float64, verbose=True):
    """Returns the reference's 8-tuple: (loop, particles, fields, static_parameters, dynamic_parameters,
    plotting_parameters, plasma_parameters, species_config)  (initialization.py:479-488)."""
    config = {} if toml_file is None else toml_file
    plotting, static_config, dynamic_config = default_parameters()
    static_config, dynamic_config, plotting = update_parameters_from_toml(config, static_config, dynamic_config, plotting)
    if config.get("pml"):
        raise NotImplementedError("[pml] is outside the hot path of pypic3d_b200 (SURVEY.md section 2 row 15)")
    Nx, Ny, Nz = dynamic_config["Nx"], dynamic_config["Ny"], dynamic_config["Nz"]
    electrostatic = static_config.get("solver") == "electrostatic"
    for n, w in zip((Nx, Ny, Nz), ("particle_tile_nx", "particle_tile_ny", "particle_tile_nz")):
        if static_config[w] is None or electrostatic:                                        # :243-254: one tile when electrostatic
            static_config[w] = int(n)
    static_config["guard_cells"] = max(int(static_config["guard_cells"]), 2)                  # :256
    _validate(static_config, dynamic_config)
    os.makedirs(os.path.join(static_config["output_dir"], "data"), exist_ok=True)
    dx, dy, dz = dynamic_config["x_wind"] / Nx, dynamic_config["y_wind"] / Ny, dynamic_config["z_wind"] / Nz
    dynamic_config.update(dx=dx, dy=dy, dz=dz)
    if dynamic_config["dt"] is None:                                                          # :267-272
        dynamic_config["dt"] = courant_condition(static_config["cfl"], dx, dy, dz, SimpleNamespace(**dynamic_config))
    dt = dynamic_config["dt"]
    nt_given = static_config["Nt"] is not None
    static_config["Nt"] = int(static_config["Nt"]) if nt_given else int(dynamic_config["t_wind"] / dt)
    if nt_given:                                   # initialization.py:279-283: an explicit Nt defines the time window
        dynamic_config["t_wind"] = dt * static_config["Nt"]
    static_config["electrostatic"] = electrostatic                                            # :237-238
    static_config["current_deposition"] = "esirkepov" if static_config["current_calculation"] == "esirkepov" else "direct"
    static_config["current_filter"] = static_config["filter_j"]
    static_config["boundary_conditions"] = {a: _encode_field_bc(static_config[f"{a}_bc"]) for a in "xyz"}
    static_config["particle_boundary_conditions"] = {a: _encode_particle_bc(static_config[f"particle_{a}_bc"]) for a in "xyz"}
    static_config["Nx"], static_config["Ny"], static_config["Nz"] = Nx, Ny, Nz
    static_config["pml_active"] = False
    static_config["field_mesh"] = make_field_mesh((Nx // static_config["particle_tile_nx"], Ny // static_config["particle_tile_ny"],
                                                   Nz // static_config["particle_tile_nz"]))
    ns = SimpleNamespace(**dynamic_config)
    center, vertex = build_collocated_grid(ns) if electrostatic else build_yee_grid(ns)      # :310-313
    dynamic_config["grids"] = {"center": center, "vertex": vertex, "tiled_center_grid": (), "tiled_vertex_grid": ()}
    sp = build_static_parameters(static_config)
    dp = build_dynamic_parameters(dynamic_config)
    tc, tv = build_tiled_yee_grids(sp, dp)
    dp = dp._replace(grids=dp.grids._replace(tiled_center_grid=tc, tiled_vertex_grid=tv))

    particles_np, species_np, names, meta = load_particles_from_toml(config, sp, dp, verbose=verbose)
    plotting["particle_species_names"] = names
    plotting["particle_species_metadata"] = meta                                             # :359-360
    w, g = [int(v) for v in sp.tile_shape], int(sp.guard_cells)
    shape = tuple(sp.field_mesh) + (w[0] + 2 * g, w[1] + 2 * g, w[2] + 2 * g)
    fields_np = [np.zeros(shape) for _ in range(9)]
    ext = (tuple(np.zeros(shape) for _ in range(3)), tuple(np.zeros(shape) for _ in range(3)))
    fields_np, ext = load_external_fields_from_toml(fields_np, ext, config, sp, dp)

    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("pypic3d_b200 needs a CUDA device: there is no CPU path")
        device = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)
    particles = TiledParticles(x=t(particles_np.x), u=t(particles_np.u), active=torch.from_numpy(particles_np.active).to(device))
    E = update_tiled_vector_ghost_cells(tuple(t(a) for a in fields_np[0:3]), sp, g)                       # :414-419
    B = update_tiled_vector_ghost_cells(tuple(t(a) for a in fields_np[3:6]), sp, g)
    J = tuple(t(a) for a in fields_np[6:9])
    ext_E = update_tiled_vector_ghost_cells(tuple(t(a) for a in ext[0]), sp, g)
    ext_B = update_tiled_vector_ghost_cells(tuple(t(a) for a in ext[1]), sp, g)
    rho, phi = t(np.zeros(shape)), t(np.zeros(shape))
    fields = (E, B, J, rho, phi, (ext_E, ext_B), None, torch.tensor(False, device=device))
    from .utils import build_plasma_parameters_dict, particle_sanity_check, print_stats
    plasma_parameters = build_plasma_parameters_dict(sp, dp, meta[0]) if meta else {}       # :381-384 (first species)
    plasma_parameters["species"] = meta
    particle_sanity_check(particles)                                                          # :386
    if verbose:
        print_stats(sp, dp)                                                                   # :379
    if verbose:
        print(f"Initializing Simulation: {sp.name}\nUsing tiled Yee storage with tile shape: {sp.tile_shape}; dt = {dt}; Nt = {sp.Nt}")
    loop = time_loop_electrostatic if electrostatic else time_loop_electrodynamic             # :466-470
    return loop, particles, fields, sp, dp, plotting, plasma_parameters, species_np
