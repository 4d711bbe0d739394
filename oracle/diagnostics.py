"""Energy / momentum / assembly / conservation checkers (test infrastructure).
Follows PyPIC3D/utils.py:108-202,761-801 and PyPIC3D/diagnostics/output_adapters.py:40-83.
`gauss_residual` and `continuity_residual` are not in the reference; they follow the
reference's own continuity test (tests/code_tests/esirkepov_test.py:700-744)."""
import numpy as np


def compute_energy(particles, E, B, sp, dp, species_config):
    g = int(sp.guard_cells)                                             # utils.py:155-187
    I = (slice(None),) * 3 + (slice(g, -g),) * 3
    dV = dp.dx * dp.dy * dp.dz
    E2 = np.sum(np.asarray(E[0])[I] ** 2 + np.asarray(E[1])[I] ** 2 + np.asarray(E[2])[I] ** 2) * dV
    B2 = np.sum(np.asarray(B[0])[I] ** 2 + np.asarray(B[1])[I] ** 2 + np.asarray(B[2])[I] ** 2) * dV
    e_energy = 0.5 * dp.eps * E2
    b_energy = 0.5 / dp.mu * B2
    C = dp.C
    u = np.asarray(particles.u, dtype=np.float64)
    v2 = u[..., 0] ** 2 + u[..., 1] ** 2 + u[..., 2] ** 2
    active = np.asarray(particles.active).astype(np.float64)
    sm = np.asarray(species_config.mass, dtype=np.float64) * np.asarray(species_config.weight, dtype=np.float64)
    mass = np.broadcast_to(sm.reshape((1, 1, 1, sm.shape[0], 1)), active.shape)
    with np.errstate(all="ignore"):
        gamma = 1.0 / np.sqrt(1 - v2 / C ** 2)
        p2 = np.square(mass * gamma) * v2
        ke = np.sum(active * (np.sqrt(p2 * C ** 2 + mass ** 2 * C ** 4) - mass * C ** 2))
    return e_energy, b_energy, ke


def compute_total_momentum(particles, species_config):
    u = np.asarray(particles.u, dtype=np.float64)                       # utils.py:190-202
    vmag = np.sqrt(u[..., 0] ** 2 + u[..., 1] ** 2 + u[..., 2] ** 2)
    active = np.asarray(particles.active).astype(np.float64)
    sm = np.asarray(species_config.mass, dtype=np.float64) * np.asarray(species_config.weight, dtype=np.float64)
    mass = np.broadcast_to(sm.reshape((1, 1, 1, sm.shape[0], 1)), active.shape)
    return np.sum(active * vmag * mass)


def courant_condition(courant_number, dx, dy, dz, dp):
    inv = sum(1 / d for d, n in zip((dx, dy, dz), (dp.Nx, dp.Ny, dp.Nz)) if n > 1)   # utils.py:761-801
    return courant_number / (dp.C * inv)


def assemble_tiled_scalar_field(field_tiles, tile_shape, num_guard_cells=2):
    """output_adapters.py:40-75: tiles -> one global array with ONE ghost layer."""
    ft = np.asarray(field_tiles)
    w = [int(v) for v in tile_shape]
    g = int(num_guard_cells)
    nt = ft.shape[:3]
    out = np.zeros((nt[0] * w[0] + 2, nt[1] * w[1] + 2, nt[2] * w[2] + 2), dtype=ft.dtype)
    for tx in range(nt[0]):
        for ty in range(nt[1]):
            for tz in range(nt[2]):
                t = ft[tx, ty, tz, g - 1:g + w[0] + 1, g - 1:g + w[1] + 1, g - 1:g + w[2] + 1]
                out[tx * w[0]:tx * w[0] + w[0] + 2, ty * w[1]:ty * w[1] + w[1] + 2, tz * w[2]:tz * w[2] + w[2] + 2] = t
    return out


def assemble_tiled_vector_field(field_tiles, tile_shape, num_guard_cells=2):
    return tuple(assemble_tiled_scalar_field(c, tile_shape, num_guard_cells) for c in field_tiles)


def continuity_residual(rho_old, rho_new, J, sp, dp):
    """(rho_new-rho_old)/dt + div J on interior nodes, backward differences (J_c lives on face i+1/2 at index i).
    Periodic global assembly; inactive axes contribute no divergence."""
    g = int(sp.guard_cells)
    w = sp.tile_shape

    def glob(f):
        return assemble_tiled_scalar_field(f, w, g)[1:-1, 1:-1, 1:-1]
    r0, r1 = glob(rho_old), glob(rho_new)
    Jx, Jy, Jz = (glob(c) for c in J)
    div = np.zeros_like(r0)
    if r0.shape[0] > 1:
        div += (Jx - np.roll(Jx, 1, axis=0)) / dp.dx
    if r0.shape[1] > 1:
        div += (Jy - np.roll(Jy, 1, axis=1)) / dp.dy
    if r0.shape[2] > 1:
        div += (Jz - np.roll(Jz, 1, axis=2)) / dp.dz
    return (r1 - r0) / dp.dt + div


def gauss_residual(E, rho, sp, dp):
    """div E - rho/eps on interior nodes of a periodic domain."""
    g = int(sp.guard_cells)
    w = sp.tile_shape

    def glob(f):
        return assemble_tiled_scalar_field(f, w, g)[1:-1, 1:-1, 1:-1]
    Ex, Ey, Ez = (glob(c) for c in E)
    r = glob(rho)
    div = np.zeros_like(r)
    if r.shape[0] > 1:
        div += (Ex - np.roll(Ex, 1, axis=0)) / dp.dx
    if r.shape[1] > 1:
        div += (Ey - np.roll(Ey, 1, axis=1)) / dp.dy
    if r.shape[2] > 1:
        div += (Ez - np.roll(Ez, 1, axis=2)) / dp.dz
    return div - r / dp.eps
