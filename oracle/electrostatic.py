"""Electrostatic field solve (test infrastructure).  Follows PyPIC3D/solvers/electrostatic_yee.py:
_apply_tiled_phi_constant_boundaries :20-37, solve_poisson_with_conjugate_gradient :71-156,
_centered_tiled_electrostatic_gradient :212-246, calculate_tiled_electrostatic_fields :249-276.
Pinned by the reference's single-mode Poisson test (tests/physics_tests/electrostatic_yee_test.py:92-119)."""
import numpy as np

from .deposition import compute_rho
from .filters import digital_filter
from .halo import BC_CONDUCTING, apply_tiled_constant_boundary, update_tiled_ghost_cells, update_tiled_vector_ghost_cells


def apply_tiled_phi_constant_boundaries(field_tiles, sp, g):
    """:20-37 -- per conducting axis: ghost refresh, then constant exterior ghosts; plain refresh when no axis conducts.
    (Each conducting axis re-runs the full refresh, which zeroes the exterior ghosts of every non-periodic axis, so with two
    conducting axes only the last one keeps constant-potential ghosts.  Reference behaviour, followed as is.)"""
    applied = False
    for axis in range(3):
        if int(sp.boundary_conditions[axis]) == BC_CONDUCTING:
            field_tiles = apply_tiled_constant_boundary(field_tiles, sp, axis, g)
            applied = True
    if not applied:
        field_tiles = update_tiled_ghost_cells(field_tiles, sp, g)
    return field_tiles


def _refresh_single_tile_scalar(field, sp, g, apply_conducting=False):
    """:40-47"""
    tiles = np.asarray(field, dtype=np.float64)[None, None, None]
    tiles = apply_tiled_phi_constant_boundaries(tiles, sp, g) if apply_conducting else update_tiled_ghost_cells(tiles, sp, g)
    return np.asarray(tiles)[0, 0, 0]


def solve_poisson_with_conjugate_gradient(rho, phi, sp, dp, tol=1e-12, max_iter=5000, return_iterations=False):
    """:71-156 on (Nx+2g, Ny+2g, Nz+2g) arrays."""
    dx, dy, dz, eps = float(dp.dx), float(dp.dy), float(dp.dz), float(dp.eps)
    g = int(sp.guard_cells)
    A = slice(g, -g)
    F = slice(g + 1, None if g == 1 else -g + 1)
    Bk = slice(g - 1, -g - 1)

    def lapl(f):                                                                                        # :101-105
        d2x = (f[F, A, A] + f[Bk, A, A] - 2.0 * f[A, A, A]) / (dx * dx)
        d2y = (f[A, F, A] + f[A, Bk, A] - 2.0 * f[A, A, A]) / (dy * dy)
        d2z = (f[A, A, F] + f[A, A, Bk] - 2.0 * f[A, A, A]) / (dz * dz)
        return d2x + d2y + d2z

    def apply_bc(f):                                                                                    # :107-108
        return _refresh_single_tile_scalar(f, sp, g, apply_conducting=True)

    rho = np.asarray(rho, dtype=np.float64)
    phi = apply_bc(np.array(phi, dtype=np.float64, copy=True))                                          # :137
    r = rho[A, A, A] / eps + lapl(phi)                                                                  # :138
    p = np.zeros_like(phi)
    p[A, A, A] = r
    p = apply_bc(p)                                                                                     # :139-141
    k = 0
    while k < max_iter and float(np.sum(r * r)) > tol ** 2:                                             # :125-128
        lapl_p = -lapl(p)                                                                               # :113
        alpha = np.sum(r * r) / np.sum(p[A, A, A] * lapl_p)                                             # :114
        phi = phi.copy()
        phi[A, A, A] += alpha * p[A, A, A]
        phi = apply_bc(phi)                                                                             # :116-117
        r_next = r - alpha * lapl_p                                                                     # :119
        beta = np.sum(r_next * r_next) / np.sum(r * r)                                                  # :120
        p = p.copy()
        p[A, A, A] = r_next + beta * p[A, A, A]
        p = apply_bc(p)                                                                                 # :121-122
        r = r_next
        k += 1
    phi = apply_bc(phi)                                                                                 # :156
    return (phi, k) if return_iterations else phi


def centered_tiled_electrostatic_gradient(phi_tiles, sp, dp, g):
    """:212-246"""
    g = int(g)
    A = slice(g, -g)
    F = slice(g + 1, None if g == 1 else -g + 1)
    Bk = slice(g - 1, -g - 1)
    phi_tiles = np.asarray(apply_tiled_phi_constant_boundaries(phi_tiles, sp, g))
    E = [np.zeros_like(phi_tiles) for _ in range(3)]
    E[0][:, :, :, A, A, A] = -1.0 * (phi_tiles[:, :, :, F, A, A] - phi_tiles[:, :, :, Bk, A, A]) / (2.0 * dp.dx)
    E[1][:, :, :, A, A, A] = -1.0 * (phi_tiles[:, :, :, A, F, A] - phi_tiles[:, :, :, A, Bk, A]) / (2.0 * dp.dy)
    E[2][:, :, :, A, A, A] = -1.0 * (phi_tiles[:, :, :, A, A, F] - phi_tiles[:, :, :, A, A, Bk]) / (2.0 * dp.dz)
    return update_tiled_vector_ghost_cells(tuple(E), sp, g)


def calculate_tiled_electrostatic_fields(sp, dp, particles, species_config, rho_tiles, phi_tiles):
    """:249-276"""
    g = int(sp.guard_cells)
    rho_tiles = np.asarray(compute_rho(particles, species_config, rho_tiles, sp, dp))                  # :260
    phi = solve_poisson_with_conjugate_gradient(rho_tiles[0, 0, 0], np.asarray(phi_tiles)[0, 0, 0], sp, dp)   # :264
    phi_tiles = phi[None, None, None]
    phi_tiles = apply_tiled_phi_constant_boundaries(phi_tiles, sp, g)                                   # :266
    phi_tiles = digital_filter(np.asarray(phi_tiles), dp.alpha, num_guard_cells=g)                      # :270
    phi_tiles = apply_tiled_phi_constant_boundaries(phi_tiles, sp, g)                                   # :271
    E_tiles = centered_tiled_electrostatic_gradient(phi_tiles, sp, dp, g)                               # :274
    return E_tiles, np.asarray(phi_tiles), rho_tiles
