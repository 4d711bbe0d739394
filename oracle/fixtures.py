"""Parameter / particle / field builders used by the parity tests (test infrastructure).
Follows /root/reference/tests/kernel_fixtures.py:29-153 (kernel_parameters), :315-345 (empty tiles),
:376-416 (field_tiles_from_global), :419-526 (particle_species), :529-594 (build_tiled_particles)."""
import math

import numpy as np

from .params import StaticParameters, DynamicParameters, GridParameters, SpeciesConfig, TiledParticles
from .grids import build_yee_grid, build_collocated_grid, build_tiled_yee_grids
from .halo import update_tiled_ghost_cells


def kernel_parameters(*, Nx=8, Ny=6, Nz=4, x_wind=4.0, y_wind=3.0, z_wind=2.0, dx=None, dy=None, dz=None, dt=0.05,
                      tile_shape=None, guard_cells=2, shape_factor=1, boundary_conditions=(0, 0, 0),
                      particle_boundary_conditions=(0, 0, 0), solver="electrodynamic_yee", electrostatic=False,
                      relativistic=True, particle_pusher="boris", current_deposition="direct", current_filter="none",
                      C=1.0, eps=1.0, mu=1.0, kb=1.0, alpha=1.0, name="test", output_dir=".", Nt=1, verbose=False,
                      GPUs=False, benchmark=False, pml_active=False, particle_tile_capacity_factor=1.0):
    dx = x_wind / Nx if dx is None else dx
    dy = y_wind / Ny if dy is None else dy
    dz = z_wind / Nz if dz is None else dz
    tile_shape = (Nx, Ny, Nz) if tile_shape is None else tuple(int(w) for w in tile_shape)
    mesh = (int(Nx) // tile_shape[0], int(Ny) // tile_shape[1], int(Nz) // tile_shape[2])
    sp = StaticParameters(name=name, output_dir=output_dir, Nt=int(Nt), verbose=bool(verbose), GPUs=bool(GPUs),
                          benchmark=bool(benchmark), solver=solver, electrostatic=bool(electrostatic),
                          relativistic=bool(relativistic), particle_pusher=particle_pusher,
                          current_deposition=current_deposition, current_filter=current_filter,
                          shape_factor=int(shape_factor), guard_cells=int(guard_cells), tile_shape=tile_shape,
                          particle_tile_capacity_factor=float(particle_tile_capacity_factor), pml_active=bool(pml_active),
                          boundary_conditions=tuple(int(v) for v in boundary_conditions),
                          particle_boundary_conditions=tuple(int(v) for v in particle_boundary_conditions),
                          field_mesh=mesh)
    dp = DynamicParameters(dt=float(dt), dx=float(dx), dy=float(dy), dz=float(dz), Nx=int(Nx), Ny=int(Ny), Nz=int(Nz),
                           x_wind=float(x_wind), y_wind=float(y_wind), z_wind=float(z_wind), C=float(C), eps=float(eps),
                           mu=float(mu), kb=float(kb), alpha=float(alpha),
                           grids=GridParameters((), (), (), ()))
    center, vertex = build_collocated_grid(dp) if electrostatic else build_yee_grid(dp)
    dp = dp._replace(grids=GridParameters(vertex=vertex, center=center, tiled_vertex_grid=(), tiled_center_grid=()))
    tc, tv = build_tiled_yee_grids(sp, dp)
    dp = dp._replace(grids=GridParameters(vertex=vertex, center=center, tiled_vertex_grid=tv, tiled_center_grid=tc))
    return sp, dp


def retiled_parameters(sp, dp, tile_shape, guard_cells=None, **static_updates):
    """kernel_fixtures.py:348-373."""
    tile_shape = tuple(int(w) for w in tile_shape)
    g = int(sp.guard_cells) if guard_cells is None else int(guard_cells)
    mesh = (int(dp.Nx) // tile_shape[0], int(dp.Ny) // tile_shape[1], int(dp.Nz) // tile_shape[2])
    sp = sp._replace(tile_shape=tile_shape, guard_cells=g, field_mesh=mesh, **static_updates)
    dp = dp._replace(grids=GridParameters(dp.grids.vertex, dp.grids.center, (), ()))
    tc, tv = build_tiled_yee_grids(sp, dp)
    dp = dp._replace(grids=GridParameters(dp.grids.vertex, dp.grids.center, tv, tc))
    return sp, dp


def empty_tiled_scalar(sp, dp):
    w = [int(v) for v in sp.tile_shape]
    g = int(sp.guard_cells)
    return np.zeros((int(dp.Nx) // w[0], int(dp.Ny) // w[1], int(dp.Nz) // w[2], w[0] + 2 * g, w[1] + 2 * g, w[2] + 2 * g))


def empty_tiled_vector(sp, dp):
    return tuple(empty_tiled_scalar(sp, dp) for _ in range(3))


def field_tiles_from_global(field, sp, dp, num_guard_cells=None):
    field = np.asarray(field, dtype=np.float64)
    w = [int(v) for v in sp.tile_shape]
    g = int(sp.guard_cells if num_guard_cells is None else num_guard_cells)
    N = (int(dp.Nx), int(dp.Ny), int(dp.Nz))
    nt = [N[a] // w[a] for a in range(3)]
    sg = [(field.shape[a] - N[a]) // 2 for a in range(3)]
    interior = field[sg[0]:sg[0] + N[0], sg[1]:sg[1] + N[1], sg[2]:sg[2] + N[2]]
    it = interior.reshape(nt[0], w[0], nt[1], w[1], nt[2], w[2]).transpose(0, 2, 4, 1, 3, 5)
    out = np.zeros((nt[0], nt[1], nt[2], w[0] + 2 * g, w[1] + 2 * g, w[2] + 2 * g))
    out[:, :, :, g:-g, g:-g, g:-g] = it
    return update_tiled_ghost_cells(out, sp, g)


def vector_tiles_from_global(field, sp, dp, num_guard_cells=None):
    return tuple(field_tiles_from_global(c, sp, dp, num_guard_cells) for c in field)


def particle_species(name, charge, mass, *, weight=1.0, x1, x2=None, x3=None, u1=None, u2=None, u3=None,
                     active_mask=None, update_x=(True, True, True), update_u=(True, True, True)):
    x1 = np.asarray(x1, dtype=np.float64)
    n = x1.shape[0]
    z = np.zeros(n)
    cols = lambda a: z if a is None else np.asarray(a, dtype=np.float64)
    if isinstance(update_x, bool):
        update_x = (update_x,) * 3
    if isinstance(update_u, bool):
        update_u = (update_u,) * 3
    return {"name": name, "charge": charge, "mass": mass, "weight": weight,
            "x": np.stack((x1, cols(x2), cols(x3)), axis=-1), "u": np.stack((cols(u1), cols(u2), cols(u3)), axis=-1),
            "active": np.ones(n, dtype=bool) if active_mask is None else np.asarray(active_mask, dtype=bool),
            "update_x": tuple(update_x), "update_u": tuple(update_u)}


def build_tiled_particles(species, sp, dp, capacity_factor=None):
    w = [int(v) for v in sp.tile_shape]
    if capacity_factor is None:
        capacity_factor = sp.particle_tile_capacity_factor
    nt = [int(math.ceil(int(n) / w[a])) for a, n in enumerate((dp.Nx, dp.Ny, dp.Nz))]
    S = len(species)
    counts = np.zeros((nt[0], nt[1], nt[2], S), dtype=int)
    data = []
    for s, sd in enumerate(species):
        x = np.asarray(sd["x"], dtype=np.float64)
        cells = [np.clip(np.floor((x[:, a] + wind / 2.0) / d).astype(int), 0, int(n) - 1)
                 for a, (wind, d, n) in enumerate(((dp.x_wind, dp.dx, dp.Nx), (dp.y_wind, dp.dy, dp.Ny), (dp.z_wind, dp.dz, dp.Nz)))]
        t = [cells[a] // w[a] for a in range(3)]
        flat = (t[0] * nt[1] + t[1]) * nt[2] + t[2]
        counts[:, :, :, s] = np.bincount(flat, minlength=nt[0] * nt[1] * nt[2]).reshape(nt)
        data.append((x, np.asarray(sd["u"], dtype=np.float64), np.asarray(sd["active"], dtype=bool), t, flat))
    cap = int(np.max(counts)) if counts.size else 0
    cap = max(1, int(math.ceil(cap * float(capacity_factor))))
    xt = np.zeros((nt[0], nt[1], nt[2], S, cap, 3)); ut = np.zeros_like(xt)
    at = np.zeros((nt[0], nt[1], nt[2], S, cap), dtype=bool)
    for s, (x, u, active, t, flat) in enumerate(data):
        order = np.argsort(flat, kind="stable")
        fc = counts[:, :, :, s].reshape(-1)
        starts = np.cumsum(fc) - fc
        slots = np.arange(order.size) - starts[flat[order]]
        xt[t[0][order], t[1][order], t[2][order], s, slots] = x[order]
        ut[t[0][order], t[1][order], t[2][order], s, slots] = u[order]
        at[t[0][order], t[1][order], t[2][order], s, slots] = active[order]
    sc = SpeciesConfig(charge=np.asarray([sd["charge"] for sd in species], dtype=np.float64),
                       mass=np.asarray([sd["mass"] for sd in species], dtype=np.float64),
                       weight=np.asarray([sd["weight"] for sd in species], dtype=np.float64),
                       update_x=np.asarray([sd["update_x"] for sd in species], dtype=bool),
                       update_u=np.asarray([sd["update_u"] for sd in species], dtype=bool))
    return TiledParticles(xt, ut, at), sc


def thermal_plasma(sp, dp, ppc_per_species=8, vth=(0.05, 0.05 / math.sqrt(1836.0)), seed=1234,
                   charge=(-1.0, 1.0), mass=(1.0, 1836.0), weight=1.0):
    """Synthetic uniform thermal plasma (SURVEY.md section 8d input 4), in units of the run's C."""
    rng = np.random.default_rng(seed)
    n = int(dp.Nx) * int(dp.Ny) * int(dp.Nz) * int(ppc_per_species)
    species = []
    for s in range(len(charge)):
        pos = [rng.uniform(-w / 2, w / 2, n) if N > 1 else np.zeros(n)
               for w, N in ((dp.x_wind, dp.Nx), (dp.y_wind, dp.Ny), (dp.z_wind, dp.Nz))]
        vel = [rng.normal(0.0, vth[s] * dp.C, n) for _ in range(3)]
        species.append(particle_species(f"s{s}", charge[s], mass[s], weight=weight, x1=pos[0], x2=pos[1], x3=pos[2],
                                        u1=vel[0], u2=vel[1], u3=vel[2]))
    return build_tiled_particles(species, sp, dp)
