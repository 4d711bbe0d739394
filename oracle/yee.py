"""First-order Yee E / B(half) updates (test infrastructure).  Follows PyPIC3D/solvers/first_order_yee.py:12-162
(pml_state is None on this path)."""
import numpy as np

from . import halo
from .filters import digital_filter_vector
from .stencil import BC_CONDUCTING


def update_E(E_tiles, B_tiles, J_tiles, sp, dp):
    g = int(sp.guard_cells)
    A = slice(g, -g)
    Bk = slice(g - 1, -g - 1)
    Ex, Ey, Ez = (np.array(c, dtype=np.float64, copy=True) for c in E_tiles)
    Bx, By, Bz = halo.update_tiled_vector_ghost_cells(B_tiles, sp, g)               # :31
    Jx, Jy, Jz = J_tiles
    dt, dx, dy, dz, C, eps = dp.dt, dp.dx, dp.dy, dp.dz, dp.C, dp.eps
    T = (slice(None),) * 3
    dBz_dy = (Bz[T + (A, A, A)] - Bz[T + (A, Bk, A)]) / dy                          # :42-47
    dBy_dz = (By[T + (A, A, A)] - By[T + (A, A, Bk)]) / dz
    dBx_dz = (Bx[T + (A, A, A)] - Bx[T + (A, A, Bk)]) / dz
    dBx_dy = (Bx[T + (A, A, A)] - Bx[T + (A, Bk, A)]) / dy
    dBz_dx = (Bz[T + (A, A, A)] - Bz[T + (Bk, A, A)]) / dx
    dBy_dx = (By[T + (A, A, A)] - By[T + (Bk, A, A)]) / dx
    curl_x = dBz_dy - dBy_dz
    curl_y = dBx_dz - dBz_dx
    curl_z = dBy_dx - dBx_dy
    I = T + (A, A, A)
    Ex[I] = Ex[I] + (C ** 2 * curl_x - np.asarray(Jx)[I] / eps) * dt                # :61-72
    Ey[I] = Ey[I] + (C ** 2 * curl_y - np.asarray(Jy)[I] / eps) * dt
    Ez[I] = Ez[I] + (C ** 2 * curl_z - np.asarray(Jz)[I] / eps) * dt
    Ex, Ey, Ez = halo.update_tiled_vector_ghost_cells((Ex, Ey, Ez), sp, g)          # :74
    Ex, Ey, Ez = digital_filter_vector((Ex, Ey, Ez), dp.alpha, num_guard_cells=g)   # :78
    bcx, bcy, bcz = (int(b) for b in sp.boundary_conditions)
    if bcx == BC_CONDUCTING:                                                        # :80-89
        Ey = halo.apply_tiled_zero_boundary(Ey, sp, 0, g); Ez = halo.apply_tiled_zero_boundary(Ez, sp, 0, g)
    if bcy == BC_CONDUCTING:
        Ex = halo.apply_tiled_zero_boundary(Ex, sp, 1, g); Ez = halo.apply_tiled_zero_boundary(Ez, sp, 1, g)
    if bcz == BC_CONDUCTING:
        Ex = halo.apply_tiled_zero_boundary(Ex, sp, 2, g); Ey = halo.apply_tiled_zero_boundary(Ey, sp, 2, g)
    return halo.update_tiled_vector_ghost_cells((Ex, Ey, Ez), sp, g)                # :93


def update_B(E_tiles, B_tiles, sp, dp, do_filter=False):
    g = int(sp.guard_cells)
    A = slice(g, -g)
    F = slice(g + 1, None if g == 1 else -g + 1)
    Bx, By, Bz = (np.array(c, dtype=np.float64, copy=True) for c in B_tiles)
    Ex, Ey, Ez = halo.update_tiled_vector_ghost_cells(E_tiles, sp, g)               # :115
    dt = dp.dt / 2                                                                  # :116  (half step!)
    dx, dy, dz = dp.dx, dp.dy, dp.dz
    T = (slice(None),) * 3
    dEz_dy = (Ez[T + (A, F, A)] - Ez[T + (A, A, A)]) / dy                           # :121-126
    dEy_dz = (Ey[T + (A, A, F)] - Ey[T + (A, A, A)]) / dz
    dEx_dz = (Ex[T + (A, A, F)] - Ex[T + (A, A, A)]) / dz
    dEx_dy = (Ex[T + (A, F, A)] - Ex[T + (A, A, A)]) / dy
    dEz_dx = (Ez[T + (F, A, A)] - Ez[T + (A, A, A)]) / dx
    dEy_dx = (Ey[T + (F, A, A)] - Ey[T + (A, A, A)]) / dx
    curl_x = dEz_dy - dEy_dz
    curl_y = dEx_dz - dEz_dx
    curl_z = dEy_dx - dEx_dy
    I = T + (A, A, A)
    Bx[I] = Bx[I] - dt * curl_x                                                     # :140-142
    By[I] = By[I] - dt * curl_y
    Bz[I] = Bz[I] - dt * curl_z
    if do_filter:                                                                   # :145-159
        Bx, By, Bz = halo.update_tiled_vector_ghost_cells((Bx, By, Bz), sp, g)
        Bx, By, Bz = digital_filter_vector((Bx, By, Bz), dp.alpha, num_guard_cells=g)
    return halo.update_tiled_vector_ghost_cells((Bx, By, Bz), sp, g)                # :162
