"""Drop-ins for the step-adjacent helpers of PyPIC3D/utils.py: add_external_fields:205, compute_energy:108,
compute_total_momentum:190, courant_condition:761, dump_parameters_to_toml:655; `load_external_fields_from_toml` (:540) and
`update_parameters_from_toml` (:621) are implemented in `initialization.py` and re-exported here under the reference's path."""
import os
from datetime import datetime

import numpy as np
import toml
import torch

from . import ops
from .parameters import static_parameters_for_output, dynamic_parameters_for_output


def add_external_fields(E, B, external_fields):
    """Particles see evolved + external fields (utils.py:205-216).  Elementwise add: torch plumbing, not a hot op
    (the resident fast path gathers the external fields inside K1 instead of materialising the sum)."""
    external_E, external_B = external_fields
    return tuple(e + x for e, x in zip(E, external_E)), tuple(b + x for b, x in zip(B, external_B))


def compute_energy(particles, E, B, static_parameters, dynamic_parameters, species_config=None):
    from .boundary_conditions.ghost_cells import _halo_params
    p = _halo_params(E[0], static_parameters, static_parameters.guard_cells)
    sc = ops._lib._scalar
    acc = torch.zeros(2, dtype=torch.float64, device=E[0].device)
    for c in range(3):
        ops.sum_squares_interior(p, ops._chk(E[c], "E"), acc[0:1])
        ops.sum_squares_interior(p, ops._chk(B[c], "B"), acc[1:2])
    dV = float(sc(dynamic_parameters.dx)) * float(sc(dynamic_parameters.dy)) * float(sc(dynamic_parameters.dz))
    e_energy = 0.5 * float(sc(dynamic_parameters.eps)) * acc[0] * dV
    b_energy = 0.5 / float(sc(dynamic_parameters.mu)) * acc[1] * dV
    pp = ops.params_for(static_parameters, dynamic_parameters, species_config, particles.x,
                        mesh=tuple(particles.x.shape[:3]), gmesh=tuple(particles.x.shape[:3]))
    ke = ops.particle_energy(pp, particles.u, particles.active)[0] if particles.x.shape[4] > 0 else torch.zeros((), dtype=torch.float64, device=E[0].device)
    return e_energy, b_energy, ke


def compute_total_momentum(particles, species_config=None, static_parameters=None, dynamic_parameters=None):
    p = ops._lib.PicParams()
    x = particles.x
    p.dtype = 0 if x.dtype == ops.F32 else 1
    for a in range(3):
        p.mesh[a] = p.gmesh[a] = int(x.shape[a])
    S = int(x.shape[3])
    p.n_species = S
    p.C = 1.0
    mass = ops._lib._to_numpy(species_config.mass).reshape(-1)
    weight = ops._lib._to_numpy(species_config.weight).reshape(-1)
    for s in range(S):
        p.mass[s], p.weight[s] = float(mass[s]), float(weight[s])
    return ops.particle_energy(p, particles.u, particles.active)[1]


def courant_condition(courant_number, dx, dy, dz, dynamic_parameters):
    inv = sum(1 / d for d, n in zip((dx, dy, dz), (dynamic_parameters.Nx, dynamic_parameters.Ny, dynamic_parameters.Nz)) if n > 1)
    return courant_number / (dynamic_parameters.C * inv)


def make_dir(path):
    if not os.path.exists(path):                                                       # utils.py:218-226
        os.makedirs(path)


def vth_to_T(vth, m, kb):
    return m * vth ** 2 / kb                                                           # utils.py:229-241


def T_to_vth(T, m, kb):
    return float(np.sqrt(kb * T / m))                                                  # utils.py:243-255


def particle_sanity_check(particles):
    """Shape contract of the tiled storage (utils.py:299-312)."""
    assert tuple(particles.x.shape) == tuple(particles.u.shape)
    assert particles.x.shape[-1] == 3
    assert tuple(particles.active.shape) == tuple(particles.x.shape[:-1])


def print_stats(static_parameters, dynamic_parameters):
    sp, dp = static_parameters, dynamic_parameters                                     # utils.py:315-339
    print(f"\ntime window: {sp.Nt * dp.dt} s with {sp.Nt} time steps of {dp.dt} s")
    print(f"x window: {dp.x_wind} m with dx: {dp.dx} m")
    print(f"y window: {dp.y_wind} m with dy: {dp.dy} m")
    print(f"z window: {dp.z_wind} m with dz: {dp.dz} m\n")


def build_plasma_parameters_dict(static_parameters, dynamic_parameters, electrons):
    """Plasma frequency, Debye length and thermal velocity of the FIRST species' metadata (utils.py:384-423; the reference passes
    `particle_metadata[0]` whatever that species is)."""
    dp = dynamic_parameters
    me, Te, q, weight = (np.float64(electrons[k]) for k in ("mass", "temperature", "charge", "weight"))
    N = electrons["N_particles"]
    with np.errstate(all="ignore"):           # a neutral or weightless first species gives inf / nan like the reference, not an exception
        density = weight * N / np.float64(dp.x_wind * dp.y_wind * dp.z_wind)
        debye = float(np.sqrt(dp.eps * dp.kb * Te / (density * q ** 2)))
        wp = float(np.sqrt(density) * abs(q) / np.sqrt(dp.eps * me))
        vth = float(np.sqrt(3 * dp.kb * Te / me))
    Te = float(Te)
    return {"Theoretical Plasma Frequency": wp,
            "Debye Length": debye,
            "Thermal Velocity": vth,
            "Number of Electrons": N,
            "Temperature of Electrons": Te,
            "dx per debye length": debye / dp.dx, "dy per debye length": debye / dp.dy, "dz per debye length": debye / dp.dz}


def check_stability(plasma_parameters, dt):
    """The reference's start-up report and its two warnings (utils.py:341-381)."""
    wp, debye = plasma_parameters["Theoretical Plasma Frequency"], plasma_parameters["Debye Length"]
    per = plasma_parameters["dx per debye length"]
    if wp * dt > 2.0:
        print("# of Electrons is Low and may introduce numerical stability")
    if per < 1:
        print("Debye Length is less than the spatial resolution, this may introduce numerical instability")
    print(f"Theoretical Plasma Frequency: {wp} Hz")
    print(f"Debye Length: {debye} m")
    print(f"Thermal Velocity: {plasma_parameters['Thermal Velocity']}")
    print(f"Dx Per Debye Length: {per}")
    print(f"Number of Electrons: {plasma_parameters['Number of Electrons']}\n")


def _toml_ready(value):
    """Plain TOML-encodable data: tuples -> lists, arrays / tensors / NumPy scalars -> Python values, None entries dropped (TOML has
    no null; the `toml` encoder would otherwise write a tuple of dicts as the list of their keys)."""
    if isinstance(value, dict):
        return {str(k): _toml_ready(v) for k, v in value.items() if v is not None}
    if isinstance(value, (list, tuple)):
        return [_toml_ready(v) for v in value if v is not None]
    if hasattr(value, "tolist"):
        return _toml_ready(value.tolist())
    return value


def dump_parameters_to_toml(simulation_stats, static_parameters, dynamic_parameters, plasma_parameters, plotting_parameters, particles):
    """`<output_dir>/data/output.toml` with the reference's sections (utils.py:655-737): simulation_stats, static/dynamic parameters
    (without mesh / grids), plasma_parameters, plotting (without the per-species bookkeeping keys), one `[[particles]]` summary per
    species (its metadata + storage / active_particles / tile_shape), version and package_versions."""
    from . import __version__
    drop = ("particle_species_names", "particle_species_metadata")
    config = {"simulation_stats": simulation_stats,
              "static_parameters": static_parameters_for_output(static_parameters),
              "dynamic_parameters": dynamic_parameters_for_output(dynamic_parameters),
              "plasma_parameters": plasma_parameters,
              "plotting": {k: v for k, v in plotting_parameters.items() if k not in drop},
              "particles": []}
    names = plotting_parameters.get("particle_species_names")
    metadata = plotting_parameters.get("particle_species_metadata")
    tile_shape = [int(w) for w in static_parameters.tile_shape]
    active = torch.as_tensor(particles.active)
    for s in range(int(active.shape[3])):
        entry = {"name": f"species_{s}" if names is None else names[s]} if metadata is None else dict(metadata[s])
        entry["storage"] = "tiled"
        entry["active_particles"] = int(active[:, :, :, s, :].sum())
        entry["tile_shape"] = tile_shape
        config["particles"].append(entry)
    config["version"] = {"PyPIC3D_version": f"pypic3d_b200 {__version__}", "date": datetime.now().strftime("%Y-%m-%d")}
    config["package_versions"] = {"torch": torch.__version__, "numpy": np.__version__, "toml": toml.__version__}
    with open(os.path.join(static_parameters.output_dir, "data/output.toml"), "w") as f:
        toml.dump(_toml_ready(config), f)


def __getattr__(name):      # reference module path for the two TOML helpers that live in initialization.py (import cycle otherwise)
    if name in ("load_external_fields_from_toml", "update_parameters_from_toml"):
        from . import initialization
        return getattr(initialization, name)
    raise AttributeError(name)
