"""Shared seeded problem builders for the parity tests (CPU oracle side)."""
import numpy as np

from oracle import fixtures as fx

CASES = [((8, 6, 4), (2, 3, 2)), ((8, 6, 4), (8, 6, 4)), ((8, 1, 1), (2, 1, 1)), ((6, 6, 1), (3, 2, 1)), ((1, 6, 4), (1, 3, 2))]


def make_case(N, tile, sf, *, n=60, seed=11, vmax=0.4, dt=0.05, **kw):
    sp, dp = fx.kernel_parameters(Nx=N[0], Ny=N[1], Nz=N[2], x_wind=4.0 if N[0] > 1 else 1.0, y_wind=3.0 if N[1] > 1 else 1.0,
                                  z_wind=2.0 if N[2] > 1 else 1.0, shape_factor=sf, dt=dt, tile_shape=tile,
                                  particle_tile_capacity_factor=kw.pop("capacity", 2.0), **kw)
    rng = np.random.default_rng(seed)
    species = []
    for s, (q, m) in enumerate(((-1.0, 1.0), (2.0, 5.0))):
        pos = [rng.uniform(-w / 2, w / 2, n) if NN > 1 else np.zeros(n) for w, NN in ((dp.x_wind, dp.Nx), (dp.y_wind, dp.Ny), (dp.z_wind, dp.Nz))]
        vel = [rng.uniform(-vmax, vmax, n) for _ in range(3)]
        species.append(fx.particle_species(f"s{s}", q, m, weight=0.5 + s, x1=pos[0], x2=pos[1], x3=pos[2], u1=vel[0], u2=vel[1], u3=vel[2],
                                           update_u=(True, s == 0, True)))
    tp, sc = fx.build_tiled_particles(species, sp, dp)
    Eg = tuple(rng.normal(size=(N[0] + 2, N[1] + 2, N[2] + 2)) for _ in range(3))
    Bg = tuple(rng.normal(size=(N[0] + 2, N[1] + 2, N[2] + 2)) for _ in range(3))
    E, B = fx.vector_tiles_from_global(Eg, sp, dp), fx.vector_tiles_from_global(Bg, sp, dp)
    return sp, dp, tp, sc, E, B


def make_fields(sp, dp, E=None, B=None, scale=0.05, seed=5):
    """The reference `fields` 8-tuple; E/B default to small smooth random fields with consistent halos."""
    rng = np.random.default_rng(seed)
    N = (dp.Nx, dp.Ny, dp.Nz)
    if E is None:
        E = fx.vector_tiles_from_global(tuple(scale * rng.normal(size=(N[0] + 2, N[1] + 2, N[2] + 2)) for _ in range(3)), sp, dp)
    if B is None:
        B = fx.vector_tiles_from_global(tuple(scale * rng.normal(size=(N[0] + 2, N[1] + 2, N[2] + 2)) for _ in range(3)), sp, dp)
    z = fx.empty_tiled_vector
    return (E, B, z(sp, dp), fx.empty_tiled_scalar(sp, dp), fx.empty_tiled_scalar(sp, dp), (z(sp, dp), z(sp, dp)), None, False)
