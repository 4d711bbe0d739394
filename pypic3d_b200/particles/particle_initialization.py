"""Host-side particle loading with the reference's TOML schema (PyPIC3D/particles/particle_initialization.py:217-427).

Differences that cannot be avoided without JAX: positions come from `numpy.random.default_rng(i)` instead of
`jax.random.uniform(jax.random.key(i))` (:67-69, :244); velocities use `np.random.normal` like the reference (:74-76,
unseeded there -- call `np.random.seed` first for reproducible runs).  `.npy` overrides (`initial_x = "file.npy"`) load
the same arrays the reference would.  Everything else (weight carry-over between `[particleN]` blocks :240,338-343,
update flags, tile packing with `particle_tile_capacity_factor` :140-149) follows the reference."""
import math

import numpy as np

from .particle_class import SpeciesConfig, TiledParticles


def grab_particle_keys(config):
    return [k for k in config.keys() if k[:8] == "particle"]   # particle_initialization.py:10-16


def _load_positions(param, block, default, n, ds, ns, rng):
    if param in block:                                        # :26-37
        if isinstance(block[param], str):
            return np.asarray(np.load(block[param]), dtype=np.float64)
        val = block[param]
        if ns == 1:
            return val * np.ones(n)
        return rng.uniform(val - ds / 2, val + ds / 2, n)
    return default


def _load_velocities(param, block, default, n):
    if param in block:                                        # :40-46
        if isinstance(block[param], str):
            return np.asarray(np.load(block[param]), dtype=np.float64)
        return np.full(n, block[param]) + default
    return default


def pack_species_arrays_into_tiles(species_arrays, dynamic_parameters, tile_shape, capacity_factor):
    """particle_initialization.py:140-214: stable sort by flat tile, capacity = ceil(max count * factor) (>= 1)."""
    dp = dynamic_parameters
    w = [int(v) for v in tile_shape]
    N = (int(dp.Nx), int(dp.Ny), int(dp.Nz))
    nt = [int(math.ceil(N[a] / w[a])) for a in range(3)]
    S = len(species_arrays)
    counts = np.zeros((nt[0], nt[1], nt[2], S), dtype=int)
    data = []
    for s, (x, u, active) in enumerate(species_arrays):
        x = np.asarray(x, dtype=np.float64); u = np.asarray(u, dtype=np.float64); active = np.asarray(active, dtype=bool)
        cells = [np.clip(np.floor((x[:, a] + wind / 2) / d).astype(int), 0, N[a] - 1)
                 for a, (wind, d) in enumerate(((dp.x_wind, dp.dx), (dp.y_wind, dp.dy), (dp.z_wind, dp.dz)))]
        t = [cells[a] // w[a] for a in range(3)]
        flat = (t[0] * nt[1] + t[1]) * nt[2] + t[2]
        idx = np.nonzero(active)[0]
        counts[:, :, :, s] = np.bincount(flat[idx], minlength=nt[0] * nt[1] * nt[2]).reshape(nt)
        data.append((x, u, t, flat, idx))
    cap = int(np.max(counts)) if counts.size else 0
    cap = max(1, int(math.ceil(cap * float(capacity_factor))))
    xt = np.zeros((nt[0], nt[1], nt[2], S, cap, 3)); ut = np.zeros_like(xt)
    at = np.zeros((nt[0], nt[1], nt[2], S, cap), dtype=bool)
    for s, (x, u, t, flat, idx) in enumerate(data):
        order = idx[np.argsort(flat[idx], kind="stable")]
        fc = counts[:, :, :, s].reshape(-1)
        starts = np.cumsum(fc) - fc
        slots = np.arange(order.size) - starts[flat[order]]
        xt[t[0][order], t[1][order], t[2][order], s, slots] = x[order]
        ut[t[0][order], t[1][order], t[2][order], s, slots] = u[order]
        at[t[0][order], t[1][order], t[2][order], s, slots] = True
    return xt, ut, at, counts


def load_particles_from_toml(config, static_parameters, dynamic_parameters, verbose=True):
    """Returns (TiledParticles[NumPy], SpeciesConfig[NumPy], species_names, species_metadata)."""
    config = config or {}
    dp = dynamic_parameters
    N = (int(dp.Nx), int(dp.Ny), int(dp.Nz))
    winds = (dp.x_wind, dp.y_wind, dp.z_wind)
    ds = (dp.dx, dp.dy, dp.dz)
    species_arrays, meta = [], []
    weight = 1.0                                               # carries over between blocks (:240)
    i = 0
    for key in grab_particle_keys(config):
        blk = config[key]
        rngs = [np.random.default_rng(i + k) for k in range(3)]
        i += 3
        name, charge, mass = blk["name"], blk["charge"], blk["mass"]
        if "N_particles" in blk:
            n = int(blk["N_particles"]); n_per_cell = n / (N[0] * N[1] * N[2])
        elif "N_per_cell" in blk:
            n_per_cell = blk["N_per_cell"]; n = int(n_per_cell * N[0] * N[1] * N[2])
        else:
            raise ValueError(f"[{key}] needs N_particles or N_per_cell")
        if "temperature" in blk:
            T = blk["temperature"]; vth = math.sqrt(dp.kb * T / mass)
        elif "vth" in blk:
            vth = blk["vth"]; T = mass * vth ** 2 / dp.kb
        else:
            T = 1.0; vth = math.sqrt(dp.kb * T / mass)
        Ts = [blk.get(k, T) for k in ("Tx", "Ty", "Tz")]
        lo = [blk.get(k, -w_ / 2) for k, w_ in zip(("xmin", "ymin", "zmin"), winds)]
        hi = [blk.get(k, w_ / 2) for k, w_ in zip(("xmax", "ymax", "zmax"), winds)]
        pos = [rngs[a].uniform(lo[a], hi[a], n) for a in range(3)]                      # :67-69
        vel = [np.random.normal(0, math.sqrt(dp.kb * Ts[a] / mass), n) for a in range(3)]  # :71-76
        for a, k in enumerate(("initial_x", "initial_y", "initial_z")):
            pos[a] = _load_positions(k, blk, pos[a], n, ds[a], N[a], rngs[a])
        for a, k in enumerate(("initial_vx", "initial_vy", "initial_vz")):
            vel[a] = _load_velocities(k, blk, vel[a], n)
        for k in ("x_bc", "y_bc", "z_bc"):                      # metadata only (:303-322)
            if k in blk and blk[k] not in ("periodic", "reflecting", "absorbing"):
                raise AssertionError(f"Invalid {k[0]} boundary condition: {blk[k]}")
        if "temperature" not in blk:
            T = (mass / (3 * dp.kb * n)) * float(sum(np.sum(v ** 2) for v in vel))
        if "weight" in blk:
            weight = blk["weight"]
        elif "number_density" in blk:
            weight = (blk["number_density"] / n_per_cell) * (dp.dx * dp.dy * dp.dz)
        g = lambda k: bool(blk.get(k, True))
        upd_x = (g("update_pos") and g("update_x"), g("update_pos") and g("update_y"), g("update_pos") and g("update_z"))
        upd_u = (g("update_v") and g("update_vx"), g("update_v") and g("update_vy"), g("update_v") and g("update_vz"))
        species_arrays.append((np.stack(pos, -1), np.stack(vel, -1), np.ones(n, dtype=bool)))
        meta.append({"name": name, "N_particles": n, "N_per_cell": n_per_cell, "charge": charge, "mass": mass, "temperature": T,
                     "thermal_velocity": vth, "weight": weight, "x_bc": blk.get("x_bc", "periodic"), "y_bc": blk.get("y_bc", "periodic"),
                     "z_bc": blk.get("z_bc", "periodic"), "update_x": upd_x, "update_u": upd_u})
        if verbose:
            print(f"\nInitializing particle species: {name}\nNumber of particles: {n}\nCharge: {charge}\nMass: {mass}\n"
                  f"Thermal Velocity: {vth}\nParticle Weight: {weight}")
    xt, ut, at, _ = pack_species_arrays_into_tiles(species_arrays, dp, static_parameters.tile_shape,
                                                   static_parameters.particle_tile_capacity_factor)
    sc = SpeciesConfig(charge=np.asarray([m["charge"] for m in meta], dtype=np.float64),
                       mass=np.asarray([m["mass"] for m in meta], dtype=np.float64),
                       weight=np.asarray([m["weight"] for m in meta], dtype=np.float64),
                       update_x=np.asarray([m["update_x"] for m in meta], dtype=bool).reshape((-1, 3)),
                       update_u=np.asarray([m["update_u"] for m in meta], dtype=bool).reshape((-1, 3)))
    return TiledParticles(xt, ut, at), sc, tuple(m["name"] for m in meta), tuple(meta)
