"""Pins the NumPy oracle against the known-answer tests held by the reference's own test-suite
(SURVEY.md section 8c.1).  Each test names the reference test file:line it ports; literal inputs and
expected values are the reference's."""
import numpy as np
import pytest

from oracle import stencil, grids, halo, particles as opart, diagnostics, yee, fixtures as fx, evolve
from oracle.params import TiledParticles, SpeciesConfig

BC_P, BC_C = stencil.BC_PERIODIC, stencil.BC_CONDUCTING


# ---- tests/code_tests/grid_and_stencil_test.py:27-110 ---------------------------------------------
def test_wrap_periodic_position():
    w = stencil.wrap_periodic_position(np.array([-2.5, -2.0, 1.75, 2.0, 2.25, 6.5]), 4.0)
    assert np.allclose(w, [1.5, -2.0, 1.75, 2.0, -1.75, -1.5])


def test_axis_activity_and_inactive_index():
    assert stencil.axis_has_active_cells(6, ghost_cells=True)
    assert not stencil.axis_has_active_cells(3, ghost_cells=True)
    assert stencil.inactive_axis_index(3, ghost_cells=True) == 1
    assert stencil.inactive_axis_index(1, ghost_cells=False) == 0


def test_axis_spacing_and_anchor_helpers():
    axis = grids.build_collocated_axis(-1.0, 0.5, 4)
    pos = np.array([0.1, 0.6])
    a = stencil.compute_particle_anchor(pos, axis, 1)
    assert float(stencil.uniform_axis_spacing(axis)) == pytest.approx(0.5)
    assert np.array_equal(a, [3, 4])
    assert np.allclose(stencil.particle_axis_offset(pos, a, axis), [0.1, 0.1])


def test_build_axis_stencil_points_periodic_and_conducting():
    anchor, off = np.array([0, 5]), np.array([-1, 0, 1])
    p = stencil.build_axis_stencil_points(anchor, 6, BC_P, off)
    c = stencil.build_axis_stencil_points(anchor, 6, BC_C, off)
    assert np.array_equal(p[:, 0], [5, 0, 1]) and np.array_equal(p[:, 1], [4, 5, 0])
    assert np.array_equal(c[:, 0], [-1, 0, 1]) and np.array_equal(c[:, 1], [4, 5, 6])


def test_prepare_particle_axis_stencil_wraps_periodic_indices():
    axis = grids.build_collocated_axis(-1.0, 0.5, 4)
    _, a, off, pts = stencil.prepare_particle_axis_stencil(np.array([1.1]), axis, 6, 1, BC_P, ghost_cells=True)
    assert np.array_equal(a, [5]) and np.allclose(off, [0.1]) and np.array_equal(pts[:, 0], [4, 5, 0])


def test_collapse_axis_stencil_for_inactive_axis():
    pts = np.array([[0, 1], [1, 1], [2, 1]])
    w = np.array([[0.2, 0.1], [0.5, 0.6], [0.3, 0.3]])
    cp, cw = stencil.collapse_axis_stencil(pts, w, 3, ghost_cells=True)
    assert np.array_equal(cp, [[1, 1]]) and np.allclose(cw, [[1.0, 1.0]])


def test_build_axis_helpers_include_ghost_cells():
    assert np.allclose(grids.build_collocated_axis(-1.0, 0.5, 4), [-1.5, -1.0, -0.5, 0.0, 0.5, 1.0])
    assert np.allclose(grids.build_staggered_axis(-1.0, 0.5, 4), [-1.25, -0.75, -0.25, 0.25, 0.75, 1.25])


# ---- tests/code_tests/distributed_ghost_cells_test.py:238-500 (g=1 KATs with hand-set ghosts) -----
class _SP:
    def __init__(self, bcs, tile_shape, pbcs=(0, 0, 0)):
        self.boundary_conditions, self.tile_shape, self.particle_boundary_conditions = bcs, tile_shape, pbcs


def _coordinate_tiles(mesh, tile, g=1):
    t = np.zeros(tuple(mesh) + tuple(w + 2 * g for w in tile))
    ii, jj, kk = np.meshgrid(*(np.arange(w, dtype=float) for w in tile), indexing="ij")
    for tx in range(mesh[0]):
        for ty in range(mesh[1]):
            for tz in range(mesh[2]):
                t[tx, ty, tz, g:-g, g:-g, g:-g] = 100.0 * tx + 10.0 * ty + tz + 0.01 * ii + 0.001 * jj + 0.0001 * kk
    return t


def test_one_device_periodic_refresh_self_exchange():
    tiles = _coordinate_tiles((1, 1, 1), (3, 2, 2))
    out = halo.update_tiled_ghost_cells(tiles, _SP((0, 0, 0), (3, 2, 2)), 1)
    assert np.allclose(out[0, 0, 0, 0, 1:-1, 1:-1], tiles[0, 0, 0, -2, 1:-1, 1:-1])
    assert np.allclose(out[0, 0, 0, -1, 1:-1, 1:-1], tiles[0, 0, 0, 1, 1:-1, 1:-1])


def _ghost_deposits(shape=(5, 4, 4)):
    t = np.zeros((1, 1, 1) + shape)
    t[0, 0, 0, 0, 1:-1, 1:-1] = 3.0
    t[0, 0, 0, -1, 1:-1, 1:-1] = 5.0
    return t


def test_one_device_periodic_fold():
    out = halo.fold_tiled_ghost_cells(_ghost_deposits(), _SP((0, 0, 0), (3, 2, 2)), 1)
    assert np.allclose(out[0, 0, 0, -2, 1:-1, 1:-1], 3.0) and np.allclose(out[0, 0, 0, 1, 1:-1, 1:-1], 5.0)
    assert np.allclose(out[0, 0, 0, 0], 0.0) and np.allclose(out[0, 0, 0, -1], 0.0)


def test_one_device_conducting_fold_sign():
    out = halo.fold_tiled_ghost_cells(_ghost_deposits(), _SP((1, 0, 0), (3, 2, 2)), 1)
    assert np.allclose(out[0, 0, 0, 1, 1:-1, 1:-1], -3.0) and np.allclose(out[0, 0, 0, -2, 1:-1, 1:-1], -5.0)
    assert np.allclose(out[0, 0, 0, 0], 0.0) and np.allclose(out[0, 0, 0, -1], 0.0)


def test_reduced_axis_refresh_uses_single_interior_cell():
    tiles = _coordinate_tiles((1, 1, 1), (1, 3, 2))
    tiles[0, 0, 0, 0] = -100.0
    tiles[0, 0, 0, -1] = 100.0
    out = halo.update_tiled_ghost_cells(tiles, _SP((0, 0, 0), (1, 3, 2)), 1)
    assert np.allclose(out[0, 0, 0, 0], out[0, 0, 0, 1]) and np.allclose(out[0, 0, 0, -1], out[0, 0, 0, 1])
    tiles = _coordinate_tiles((2, 2, 1), (2, 2, 1))
    out = halo.update_tiled_ghost_cells(tiles, _SP((0, 0, 0), (2, 2, 1)), 1)
    assert np.allclose(out[..., 0], out[..., 1]) and np.allclose(out[..., -1], out[..., 1])


def test_bc_type_selects_field_or_particle_boundaries():
    tiles = _coordinate_tiles((1, 1, 1), (3, 2, 2))
    sp = _SP((0, 0, 0), (3, 2, 2), pbcs=(2, 0, 0))
    f = halo.update_tiled_ghost_cells(tiles, sp, 1, bc_type=0)
    p = halo.update_tiled_ghost_cells(tiles, sp, 1, bc_type=1)
    assert np.allclose(f[0, 0, 0, 0, 1:-1, 1:-1], tiles[0, 0, 0, -2, 1:-1, 1:-1])
    assert np.allclose(p[0, 0, 0, 0], 0.0) and np.allclose(p[0, 0, 0, -1], 0.0)
    ff = halo.fold_tiled_ghost_cells(_ghost_deposits(), sp, 1, bc_type=0)
    pf = halo.fold_tiled_ghost_cells(_ghost_deposits(), sp, 1, bc_type=1)
    assert np.allclose(ff[0, 0, 0, -2, 1:-1, 1:-1], 3.0) and np.allclose(ff[0, 0, 0, 1, 1:-1, 1:-1], 5.0)
    assert np.allclose(pf[0, 0, 0, 1:-1, 1:-1, 1:-1], 0.0) and np.allclose(pf[0, 0, 0, 0], 0.0)


def test_reduced_absorbing_particle_axis_discards_ghosts():
    sp = _SP((0, 0, 0), (1, 2, 2), pbcs=(2, 0, 0))
    r = halo.update_tiled_ghost_cells(_coordinate_tiles((1, 1, 1), (1, 2, 2)), sp, 1, bc_type=1)
    assert np.allclose(r[0, 0, 0, 0], 0.0) and np.allclose(r[0, 0, 0, -1], 0.0)
    f = halo.fold_tiled_ghost_cells(_ghost_deposits((3, 4, 4)), sp, 1, bc_type=1)
    assert np.allclose(f[0, 0, 0], 0.0)


# ---- tests/code_tests/particle_refresh_test.py:80-183 ---------------------------------------------
def _refresh_params(pbc=(0, 0, 0)):
    return fx.kernel_parameters(Nx=4, Ny=1, Nz=1, x_wind=4.0, y_wind=1.0, z_wind=1.0, dt=1.0, tile_shape=(2, 1, 1),
                                particle_boundary_conditions=pbc)


def _moving_species(x1, v1, active=None, update_x=True):
    return fx.particle_species("moving", 2.0, 3.0, weight=4.0, x1=x1, u1=v1, active_mask=active, update_x=update_x)


def _active_rows(tp):
    a = tp.active.reshape(-1)
    x, u = tp.x.reshape(-1, 3)[a], tp.u.reshape(-1, 3)[a]
    o = np.argsort(x[:, 0], kind="stable")
    return x[o], u[o]


def test_update_positions_respects_active_and_update_flags():
    sp, dp = _refresh_params()
    tp, sc = fx.build_tiled_particles([_moving_species([-1.5, -0.5, 0.5], [0.25, 0.5, 0.75], active=[True, False, True])], sp, dp)
    moved = opart.update_tiled_particle_positions(tp, sc, 1.0)
    x, _ = _active_rows(moved)
    assert np.allclose(x[:, 0], [-1.25, 1.25])
    assert np.allclose(moved.x[~tp.active], tp.x[~tp.active])
    tp, sc = fx.build_tiled_particles([_moving_species([-1.5], [0.25], update_x=False)], sp, dp)
    assert np.allclose(opart.update_tiled_particle_positions(tp, sc, 1.0).x, tp.x)


def test_adjacent_tile_offset_periodic_edges():
    assert np.array_equal(opart.adjacent_tile_offset(np.array([0, 1, 0, 1]), np.array([0, 0, 1, 1]), 2), [0, 1, -1, 0])


def test_refresh_moves_to_neighbor_static_shape():
    sp, dp = _refresh_params()
    tp, sc = fx.build_tiled_particles([_moving_species([-1.5, -0.25, 0.25], [0.0] * 3)], sp, dp)
    x = tp.x.copy(); x[0, 0, 0, 0, 1, 0] = 0.25
    r, ovf = opart.refresh_tiled_particle_tiles(tp._replace(x=x), sp, dp)
    assert r.x.shape == tp.x.shape and not ovf
    assert r.active[0, 0, 0, 0].sum() == 1 and r.active[1, 0, 0, 0].sum() == 2
    xs, us = _active_rows(r)
    assert np.allclose(xs[:, 0], [-1.5, 0.25, 0.25]) and np.allclose(us[:, 0], 0.0)


def test_refresh_wraps_periodic():
    sp, dp = _refresh_params()
    tp, sc = fx.build_tiled_particles([_moving_species([1.75], [0.0])], sp, dp)
    x = tp.x.copy(); x[1, 0, 0, 0, 0, 0] = 2.25
    r, ovf = opart.refresh_tiled_particle_tiles(tp._replace(x=x), sp, dp)
    assert not ovf and r.active[0, 0, 0, 0, 0] and np.isclose(r.x[0, 0, 0, 0, 0, 0], -1.75)


def test_refresh_reflects():
    sp, dp = _refresh_params((1, 0, 0))
    tp, sc = fx.build_tiled_particles([_moving_species([1.75, -1.75], [0.5, -0.25])], sp, dp)
    x = tp.x.copy(); x[1, 0, 0, 0, 0, 0] = 2.25; x[0, 0, 0, 0, 0, 0] = -2.10
    r, ovf = opart.refresh_tiled_particle_tiles(tp._replace(x=x), sp, dp)
    xs, us = _active_rows(r)
    assert not ovf and np.allclose(xs[:, 0], [-1.90, 1.75]) and np.allclose(us[:, 0], [0.25, -0.5])


def test_refresh_absorbs():
    sp, dp = _refresh_params((2, 0, 0))
    tp, sc = fx.build_tiled_particles([_moving_species([1.75, -0.25], [0.5, 0.0])], sp, dp)
    x = tp.x.copy(); x[1, 0, 0, 0, 0, 0] = 2.25
    r, ovf = opart.refresh_tiled_particle_tiles(tp._replace(x=x), sp, dp)
    xs, us = _active_rows(r)
    assert not ovf and r.active.sum() == 1 and np.allclose(xs[:, 0], [-0.25]) and np.allclose(us[:, 0], [0.0])


def test_refresh_reports_overflow():
    sp, dp = _refresh_params()
    tp, sc = fx.build_tiled_particles([_moving_species([-1.5, 0.5, 1.5], [0.0] * 3)], sp, dp)
    x = tp.x.copy(); x[0, 0, 0, 0, 0, 0] = 0.25
    r, ovf = opart.refresh_tiled_particle_tiles(tp._replace(x=x), sp, dp)
    assert r.x.shape == tp.x.shape and ovf and r.active.sum() == 2


# ---- tests/code_tests/distributed_particle_refresh_test.py:105-169 --------------------------------
def _dist_params(mesh, tile):
    n = [mesh[a] * tile[a] for a in range(3)]
    return fx.kernel_parameters(Nx=n[0], Ny=n[1], Nz=n[2], x_wind=float(n[0]), y_wind=float(n[1]), z_wind=float(n[2]),
                                tile_shape=tile)


def _empty_particles(mesh, slots=2):
    return TiledParticles(np.zeros(mesh + (1, slots, 3)), np.zeros(mesh + (1, slots, 3)), np.zeros(mesh + (1, slots), dtype=bool))


def _put(p, tile, slot, x):
    p.x[tile + (0, slot)] = x
    p.active[tile + (0, slot)] = True
    return p


def test_dist_refresh_x_neighbor():
    sp, dp = _dist_params((2, 1, 1), (2, 1, 1))
    r, ovf = opart.refresh_tiled_particle_tiles(_put(_empty_particles((2, 1, 1)), (0, 0, 0), 0, (0.25, 0, 0)), sp, dp)
    assert not ovf and r.active[0, 0, 0, 0].sum() == 0 and r.active[1, 0, 0, 0].sum() == 1
    assert np.isclose(r.x[1, 0, 0, 0, 0, 0], 0.25)


def test_dist_refresh_periodic_edge():
    sp, dp = _dist_params((2, 1, 1), (2, 1, 1))
    r, ovf = opart.refresh_tiled_particle_tiles(_put(_empty_particles((2, 1, 1)), (1, 0, 0), 0, (2.25, 0, 0)), sp, dp)
    assert not ovf and r.active[1, 0, 0, 0].sum() == 0 and r.active[0, 0, 0, 0].sum() == 1
    assert np.isclose(r.x[0, 0, 0, 0, 0, 0], -1.75)


def test_dist_refresh_diagonal():
    sp, dp = _dist_params((2, 2, 1), (2, 2, 1))
    r, ovf = opart.refresh_tiled_particle_tiles(_put(_empty_particles((2, 2, 1)), (0, 0, 0), 0, (0.25, 0.25, 0)), sp, dp)
    assert not ovf and r.active[0, 0, 0, 0].sum() == 0 and r.active[1, 1, 0, 0].sum() == 1
    assert np.allclose(r.x[1, 1, 0, 0, 0, :2], [0.25, 0.25])


def test_dist_refresh_capacity_overflow():
    sp, dp = _dist_params((2, 1, 1), (2, 1, 1))
    p = _put(_put(_empty_particles((2, 1, 1), 1), (0, 0, 0), 0, (0.25, 0, 0)), (1, 0, 0), 0, (1.25, 0, 0))
    r, ovf = opart.refresh_tiled_particle_tiles(p, sp, dp)
    assert ovf and r.x.shape == p.x.shape and r.active.sum() == 1


# ---- tests/code_tests/utils_test.py:286-346 (energy KATs) -----------------------------------------
def _one_species():
    return SpeciesConfig(np.array([1.0]), np.array([1.0]), np.array([1.0]), np.ones((1, 3), bool), np.ones((1, 3), bool))


def test_energy_kat():
    sp, dp = fx.kernel_parameters(Nx=1, Ny=1, Nz=1, x_wind=1.0, y_wind=1.0, z_wind=1.0, eps=2.0, mu=4.0, C=10.0)
    E, B = fx.empty_tiled_vector(sp, dp), fx.empty_tiled_vector(sp, dp)
    g = 2
    I = (0, 0, 0, slice(g, g + 1), slice(g, g + 1), slice(g, g + 1))
    E[0][I] = 2.0
    B[1][I] = 3.0
    tp = TiledParticles(np.zeros((1, 1, 1, 1, 0, 3)), np.zeros((1, 1, 1, 1, 0, 3)), np.zeros((1, 1, 1, 1, 0), bool))
    e, b, k = diagnostics.compute_energy(tp, E, B, sp, dp, _one_species())
    assert np.isclose(e, 4.0) and np.isclose(b, 1.125) and k == 0.0


def test_energy_ignores_inactive():
    sp, dp = fx.kernel_parameters(Nx=1, Ny=1, Nz=1, x_wind=1.0, y_wind=1.0, z_wind=1.0, C=10.0)
    E, B = fx.empty_tiled_vector(sp, dp), fx.empty_tiled_vector(sp, dp)
    tp = TiledParticles(np.array([[[[[[0.6, 0, 0]]]]]]), np.array([[[[[[1.0, 0, 0]]]]]]), np.array([[[[[False]]]]]))
    assert np.isclose(diagnostics.compute_energy(tp, E, B, sp, dp, _one_species())[2], 0.0)


# ---- tests/code_tests/esirkepov_test.py:510-529 (update_E KAT) ------------------------------------
def test_update_E_reads_two_guard_current_interior():
    sp, dp = fx.kernel_parameters(Nx=4, Ny=1, Nz=1, x_wind=4.0, y_wind=1.0, z_wind=1.0, dt=0.25, tile_shape=(2, 1, 1), C=1.0, eps=2.0)
    E, B, J = fx.empty_tiled_vector(sp, dp), fx.empty_tiled_vector(sp, dp), fx.empty_tiled_vector(sp, dp)
    J[0][:, :, :, 2:-2, 2:-2, 2:-2] = 4.0
    Ea = yee.update_E(E, B, J, sp, dp)
    assert np.allclose(Ea[0][:, :, :, 2:-2, 2:-2, 2:-2], -0.5)


# ---- code-is-authority pin for the half-step B update (SURVEY.md section 4 "stale pins") -----------
def test_update_B_is_a_half_step():
    # first_order_yee.py:116 sets dt = dt/2; yee_convergence_test.py:69-75 (expects dt*cos x) is stale.
    n = 16
    sp, dp = fx.kernel_parameters(Nx=n, Ny=1, Nz=1, x_wind=2 * np.pi, y_wind=1.0, z_wind=1.0, dt=0.1)
    xc = -np.pi + np.arange(n) * dp.dx
    Ez = np.zeros((n + 2, 3, 3)); Ez[1:-1, 1, 1] = np.sin(xc)
    z = np.zeros_like(Ez)
    E = fx.vector_tiles_from_global((z, z, Ez), sp, dp)
    B = update = yee.update_B(E, fx.empty_tiled_vector(sp, dp), sp, dp)
    By = diagnostics.assemble_tiled_scalar_field(B[1], sp.tile_shape, 2)[1:-1, 1, 1]
    dEz = (np.roll(np.sin(xc), -1) - np.sin(xc)) / dp.dx
    assert np.allclose(By, +0.5 * dp.dt * dEz, atol=1e-14)   # By -= (dt/2) * (dEx/dz - dEz/dx)


# ---- tests/code_tests/ghost_cells_test.py:41-165 (tile_shape (2,2,2), g = 1; literal inputs and expectations) -----
def test_gc_periodic_refresh_between_two_tiles():
    """ghost_cells_test.py:41-57"""
    sp = _SP((0, 0, 0), (2, 2, 2))
    f = np.zeros((2, 1, 1, 4, 4, 4))
    f[0, 0, 0, 1:3, 1:3, 1:3] = 1.0
    f[1, 0, 0, 1:3, 1:3, 1:3] = 2.0
    r = halo.update_tiled_ghost_cells(f, sp, 1)
    assert np.all(r[0, 0, 0, -1, 1:3, 1:3] == 2.0) and np.all(r[1, 0, 0, 0, 1:3, 1:3] == 1.0)


def test_gc_periodic_fold_adds_to_owner_tile():
    """ghost_cells_test.py:75-97"""
    sp = _SP((0, 0, 0), (2, 2, 2))
    f = np.zeros((2, 1, 1, 4, 4, 4))
    f[0, 0, 0, -1, 2, 2] = 3.0
    f[1, 0, 0, 0, 2, 2] = 5.0
    r = halo.fold_tiled_ghost_cells(f, sp, 1)
    assert r[1, 0, 0, 1, 2, 2] == pytest.approx(3.0) and r[0, 0, 0, -2, 2, 2] == pytest.approx(5.0)
    assert r[0, 0, 0, -1, 2, 2] == 0.0 and r[1, 0, 0, 0, 2, 2] == 0.0


def test_gc_zero_boundary_zeros_global_tangential_faces():
    """ghost_cells_test.py:99-128"""
    sp = _SP((1, 1, 1), (2, 2, 2))
    Ex, Ey, Ez = (np.ones((1, 1, 1, 4, 4, 4)) for _ in range(3))
    Ey = halo.apply_tiled_zero_boundary(Ey, sp, axis=0, num_guard_cells=1)
    Ez = halo.apply_tiled_zero_boundary(Ez, sp, axis=0, num_guard_cells=1)
    Ex = halo.apply_tiled_zero_boundary(Ex, sp, axis=1, num_guard_cells=1)
    Ez = halo.apply_tiled_zero_boundary(Ez, sp, axis=1, num_guard_cells=1)
    Ex = halo.apply_tiled_zero_boundary(Ex, sp, axis=2, num_guard_cells=1)
    Ey = halo.apply_tiled_zero_boundary(Ey, sp, axis=2, num_guard_cells=1)
    assert np.all(Ey[0, 0, 0, 1, :, :] == 0.0) and np.all(Ez[0, 0, 0, 1, :, :] == 0.0)
    assert np.all(Ex[0, 0, 0, :, 1, :] == 0.0) and np.all(Ez[0, 0, 0, :, 1, :] == 0.0)
    assert np.all(Ex[0, 0, 0, :, :, 1] == 0.0) and np.all(Ey[0, 0, 0, :, :, 1] == 0.0)


def test_gc_zero_and_constant_boundary_keep_a_periodic_field():
    """ghost_cells_test.py:130-139 and :157-165"""
    sp = _SP((0, 0, 0), (2, 2, 2))
    field = np.arange(64, dtype=float).reshape((1, 1, 1, 4, 4, 4))
    refreshed = halo.update_tiled_ghost_cells(field, sp, 1)
    assert np.allclose(halo.apply_tiled_zero_boundary(refreshed, sp, axis=0, num_guard_cells=1), refreshed)
    assert np.allclose(halo.apply_tiled_constant_boundary(field, sp, axis=0, num_guard_cells=1), refreshed)


def test_gc_constant_boundary_copies_adjacent_interior_to_global_ghosts():
    """ghost_cells_test.py:141-155"""
    sp = _SP((1, 0, 0), (2, 2, 2))
    field = np.arange(64, dtype=float).reshape((1, 1, 1, 4, 4, 4))
    r = halo.apply_tiled_constant_boundary(field, sp, axis=0, num_guard_cells=1)
    assert np.allclose(r[0, 0, 0, 0, :, :], r[0, 0, 0, 1, :, :]) and np.allclose(r[0, 0, 0, -1, :, :], r[0, 0, 0, -2, :, :])
    assert np.allclose(r[0, 0, 0, 1:-1, 1:-1, 1:-1], field[0, 0, 0, 1:-1, 1:-1, 1:-1])


# ---- tests/code_tests/direct_deposition_test.py:535-620 (fold KATs with hand-set ghost deposits) -----
def test_dd_fold_periodic_adds_current_deposits_to_neighbors():
    """direct_deposition_test.py:535-548 (Nx=4, tile (2,1,1), g = 1)"""
    sp = _SP((0, 0, 0), (2, 1, 1))
    t = np.zeros((2, 1, 1, 4, 3, 3))
    t[0, 0, 0, -1, 1, 1] = 2.0
    t[1, 0, 0, 0, 1, 1] = 3.0
    f = halo.fold_tiled_ghost_cells(t, sp, 1)
    assert f[1, 0, 0, 1, 1, 1] == 2.0 and f[0, 0, 0, -2, 1, 1] == 3.0
    assert np.allclose(f[:, :, :, 0, :, :], 0.0) and np.allclose(f[:, :, :, -1, :, :], 0.0)


def test_dd_fold_two_guard_layers_adds_deposits_to_neighbors():
    """direct_deposition_test.py:550-570 (Nx=8, Ny=Nz=4, tile (4,4,4), g = 2)"""
    sp = _SP((0, 0, 0), (4, 4, 4))
    t = np.zeros((2, 1, 1, 8, 8, 8))
    t[1, 0, 0, 0, 2, 2] = 2.0
    t[1, 0, 0, 1, 2, 2] = 3.0
    t[0, 0, 0, -2, 2, 2] = 5.0
    t[0, 0, 0, -1, 2, 2] = 7.0
    f = halo.fold_tiled_ghost_cells(t, sp, 2)
    assert f[0, 0, 0, 4, 2, 2] == 2.0 and f[0, 0, 0, 5, 2, 2] == 3.0 and f[1, 0, 0, 2, 2, 2] == 5.0 and f[1, 0, 0, 3, 2, 2] == 7.0
    assert np.allclose(f[:, :, :, :2, :, :], 0.0) and np.allclose(f[:, :, :, -2:, :, :], 0.0)


def test_dd_fold_two_guard_reduced_axis_folds_to_single_active_cell():
    """direct_deposition_test.py:572-590 (Nx=8, Ny=Nz=1, tile (4,1,1), g = 2)"""
    sp = _SP((0, 0, 0), (4, 1, 1))
    t = np.zeros((2, 1, 1, 8, 5, 5))
    t[0, 0, 0, 2, 0, 2] = 1.0
    t[0, 0, 0, 2, 1, 2] = 2.0
    t[0, 0, 0, 2, 3, 2] = 3.0
    t[0, 0, 0, 2, 4, 2] = 4.0
    f = halo.fold_tiled_ghost_cells(t, sp, 2)
    assert f[0, 0, 0, 2, 2, 2] == 10.0
    assert np.allclose(f[:, :, :, :, :2, :], 0.0) and np.allclose(f[:, :, :, :, -2:, :], 0.0)


def test_dd_fold_conducting_reflects_exterior_deposits():
    """direct_deposition_test.py:592-616 (x conducting, tile (2,1,1), g = 1)"""
    sp = _SP((1, 0, 0), (2, 1, 1))
    t = np.zeros((2, 1, 1, 4, 3, 3))
    t[0, 0, 0, 0, 1, 1] = 4.0
    t[-1, 0, 0, -1, 1, 1] = 7.0
    t[0, 0, 0, -1, 1, 1] = 2.0
    t[1, 0, 0, 0, 1, 1] = 3.0
    f = halo.fold_tiled_ghost_cells(t, sp, 1)
    assert f[0, 0, 0, 1, 1, 1] == -4.0 and f[-1, 0, 0, -2, 1, 1] == -7.0 and f[1, 0, 0, 1, 1, 1] == 2.0 and f[0, 0, 0, -2, 1, 1] == 3.0
    assert np.allclose(f[:, :, :, 0, :, :], 0.0) and np.allclose(f[:, :, :, -1, :, :], 0.0)


def test_dd_fold_matches_global_fold_for_mixed_boundaries():
    """direct_deposition_test.py:618-663: tiled fold + refresh == the test's own global one-ghost fold + refresh
    (`_fold_ghost_cells:105-127`, `_update_ghost_cells:82-102`), x periodic / y conducting / z periodic, tiles (2,2,1) of 4x4x2."""
    def update_global(f, bcs):                                                                     # :82-102
        f = f.copy()
        for a, bc in enumerate(bcs):
            lo = [slice(None)] * 3; hi = [slice(None)] * 3; lo[a] = 0; hi[a] = -1
            il = [slice(None)] * 3; ih = [slice(None)] * 3; il[a] = 1; ih[a] = -2
            if bc == 0:
                f[tuple(lo)], f[tuple(hi)] = f[tuple(ih)].copy(), f[tuple(il)].copy()
            else:
                f[tuple(lo)] = 0.0; f[tuple(hi)] = 0.0
        return f

    def fold_global(f, bcs):                                                                       # :105-127
        f = f.copy()
        for a, bc in enumerate(bcs):
            lo = [slice(None)] * 3; hi = [slice(None)] * 3; lo[a] = 0; hi[a] = -1
            il = [slice(None)] * 3; ih = [slice(None)] * 3; il[a] = 1; ih[a] = -2
            sgn = 1.0 if bc == 0 else -1.0
            if bc == 0:
                f[tuple(il)] += f[tuple(hi)]; f[tuple(ih)] += f[tuple(lo)]
            else:
                f[tuple(il)] += sgn * f[tuple(lo)]; f[tuple(ih)] += sgn * f[tuple(hi)]
            f[tuple(lo)] = 0.0; f[tuple(hi)] = 0.0
        return f

    bcs, tile = (0, 1, 0), (2, 2, 1)
    sp = _SP(bcs, tile)
    field = np.zeros((6, 6, 4))
    tiles = np.zeros((2, 2, 2, 4, 4, 3))
    field[0, 2, 1] = 1.5;   tiles[0, 0, 0, 0, 2, 1] = 1.5
    field[-1, 3, 2] = -0.5; tiles[-1, 1, 1, -1, 1, 1] = -0.5
    field[2, 0, 1] = 3.0;   tiles[0, 0, 0, 2, 0, 1] = 3.0
    field[3, -1, 2] = -4.0; tiles[1, -1, 1, 1, -1, 1] = -4.0
    field[2, 2, 0] = 2.0;   tiles[0, 0, 0, 2, 2, 0] = 2.0
    field[3, 3, -1] = 5.0;  tiles[1, 1, -1, 1, 1, -1] = 5.0
    folded = halo.update_tiled_ghost_cells(halo.fold_tiled_ghost_cells(tiles, sp, 1), sp, 1)
    got = diagnostics.assemble_tiled_scalar_field(folded, tile, 1)
    want = update_global(fold_global(field, bcs), bcs)
    assert np.allclose(got, want, rtol=1e-15, atol=1e-15)


# ---- tests/code_tests/evolve_test.py:75-293 (one-particle end-to-end steps with literal expectations) -----
def _one_particle_step(N, x, u, ext_Ex=0.0, pbc=(0, 0, 0)):
    sp, dp = fx.kernel_parameters(Nx=N[0], Ny=N[1], Nz=N[2], x_wind=1.0, y_wind=1.0, z_wind=1.0, dt=0.1, shape_factor=1,
                                  current_deposition="direct", current_filter="none", relativistic=False, particle_boundary_conditions=pbc,
                                  C=1.0, eps=1.0, mu=1.0, alpha=1.0, guard_cells=2)
    tp = TiledParticles(x=np.asarray(x, dtype=float).reshape((1, 1, 1, 1, 1, 3)), u=np.asarray(u, dtype=float).reshape((1, 1, 1, 1, 1, 3)),
                        active=np.ones((1, 1, 1, 1, 1), dtype=bool))
    sc = SpeciesConfig(charge=np.array([1.0]), mass=np.array([1.0]), weight=np.array([1.0]), update_x=np.ones((1, 3), dtype=bool),
                       update_u=np.ones((1, 3), dtype=bool))
    z = fx.empty_tiled_vector
    ext_E = list(z(sp, dp)); ext_E[0] = np.full_like(ext_E[0], ext_Ex)
    fields = (z(sp, dp), z(sp, dp), z(sp, dp), fx.empty_tiled_scalar(sp, dp), fx.empty_tiled_scalar(sp, dp), (tuple(ext_E), z(sp, dp)), None, False)
    return evolve.time_loop_electrodynamic(tp, sc, fields, sp, dp), fields


def test_evolve_external_E_pushes_particle_without_evolving_it():
    """evolve_test.py:75-157"""
    (tp, f), f0 = _one_particle_step((3, 3, 3), [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], ext_Ex=1.0)
    assert f[6] is None and not bool(f[7])
    assert tp.u[0, 0, 0, 0, 0, 0] > 0.0 and tp.u[0, 0, 0, 0, 0, 1] == 0.0 and tp.u[0, 0, 0, 0, 0, 2] == 0.0
    assert np.allclose(f[5][0][0], f0[5][0][0])


def test_evolve_step_moves_the_particle():
    """evolve_test.py:159-215"""
    (tp, f), _ = _one_particle_step((3, 1, 1), [0.0, 0.0, 0.0], [0.05, 0.0, 0.0])
    assert not bool(f[-1]) and np.all(tp.active) and tp.x[0, 0, 0, 0, 0, 0] > 0.0


def test_evolve_absorbing_particle_mask():
    """evolve_test.py:217-289: x = 0.49, u = 0.2, dt = 0.1 crosses the absorbing wall at 0.5 -> inactive"""
    (tp, f), _ = _one_particle_step((3, 1, 1), [0.49, 0.0, 0.0], [0.2, 0.0, 0.0], pbc=(2, 0, 0))
    assert tp.x[:, :, :, 0, :, 0].reshape(-1).shape[0] == 1
    assert np.array_equal(tp.active[:, :, :, 0, :].reshape(-1), np.array([False]))


def test_build_yee_grid_reference_pins():
    """utils_test.py:144-170: "center" is the collocated grid (one ghost node each side), "vertex" the half-cell staggered one;
    the same for the package's host-side builder."""
    from types import SimpleNamespace
    from pypic3d_b200.utilities.grids import build_yee_grid as pkg_build
    ps = SimpleNamespace(Nx=8, Ny=6, Nz=4, x_wind=4.0, y_wind=3.0, z_wind=2.0, dx=0.5, dy=0.5, dz=0.5)
    for build in (grids.build_yee_grid, pkg_build):
        center, vertex = build(ps)
        assert len(center) == 3 and len(vertex) == 3
        assert [len(c) for c in center] == [10, 8, 6] and [len(v) for v in vertex] == [10, 8, 6]
        assert np.allclose(center[0][1:-1], -ps.x_wind / 2 + ps.dx * np.arange(ps.Nx))
        assert np.allclose(vertex[0][1:-1], center[0][1:-1] + 0.5 * ps.dx)


@pytest.mark.parametrize("layout", ["one tile", (8, 1, 1), (16, 1, 1)])
def test_update_B_second_order_convergence(layout):
    """physics_tests/yee_convergence_test.py:56-101 with its literal set-up (E_z = sin x on [.., 2 pi), dt = 1e-3, g = 2, Nx = 32 / 64 /
    128, one-tile and multi-tile layouts, observed order > 1.8) -- against (dt/2) cos x, the half step the code takes
    (first_order_yee.py:116), not the stale dt cos x of the test."""
    errors = []
    for n in (32, 64, 128):
        tile = (n, 1, 1) if layout == "one tile" else layout
        sp, dp = fx.kernel_parameters(Nx=n, Ny=1, Nz=1, x_wind=2 * np.pi, y_wind=1.0, z_wind=1.0, dt=1.0e-3, tile_shape=tile, guard_cells=2)
        center, vertex = grids.build_tiled_yee_grids(sp, dp)
        A = slice(2, -2)
        E = [np.array(c, copy=True) for c in fx.empty_tiled_vector(sp, dp)]
        E[2][:, :, :, A, A, A] = np.sin(center[0][:, A])[:, None, None, :, None, None]
        B = yee.update_B(tuple(E), fx.empty_tiled_vector(sp, dp), sp, dp)
        exact = 0.5 * dp.dt * np.cos(vertex[0][:, A])[:, None, None, :, None, None]
        errors.append(float(np.sqrt(np.mean((B[1][:, :, :, A, A, A] - exact) ** 2))))
    assert np.log2(errors[0] / errors[1]) > 1.8 and np.log2(errors[1] / errors[2]) > 1.8


# ---- tests/physics_tests/boris_test.py:133-247 (single-particle pusher KATs, C = 10) -----
def test_boris_single_particle_gyroradius():
    """boris_test.py:133-183: v = (1,0,0), B = (0,1,0), q = m = 1, dt = 0.1, 5000 steps; mean |(x, z)| ~ 1.28 +- 0.5."""
    from oracle import pusher as opush
    v, x = (1.0, 0.0, 0.0), np.zeros(3)
    r = []
    for _ in range(5000):
        v = opush.boris(v, (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 1.0, 1.0, 0.1, None)
        x = x + np.asarray(v) * 0.1
        r.append(np.hypot(x[0], x[2]))
    assert abs(np.mean(r) - 1.28) <= 0.5


def test_higuera_cary_single_particle_kats():
    """boris_test.py:185-247"""
    from oracle import pusher as opush
    C = 10.0
    assert np.allclose(opush.higuera_cary((0.2, -0.1, 0.05), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), 1.0, 1.0, 0.1, C), [0.2, -0.1, 0.05])
    v = opush.higuera_cary((0.2, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 1.0, 1.0, 0.1, C)
    assert np.isclose(np.sqrt(sum(c * c for c in v)), 0.2, rtol=1e-12, atol=1e-12)
    v = opush.higuera_cary((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 0.0), 1.0, 1.0, 0.1, C)
    assert v[0] > 0.0 and np.allclose(v[1:], 0.0)
    z2 = np.zeros(2)
    v = opush.higuera_cary((z2, z2, z2), (np.ones(2), z2, z2), (z2, z2, z2), np.array([1.0, 2.0]), np.array([1.0, 1.0]), 0.1, C)
    assert v[0].shape == (2,) and v[0][1] > v[0][0] and np.allclose(v[1], 0.0) and np.allclose(v[2], 0.0)


@pytest.mark.parametrize("shape_factor", (1, 2))
@pytest.mark.parametrize("case", ("wave", "polynomial", "oscillatory"))
def test_interpolation_order_reference_pins(shape_factor, case):
    """boris_test.py:249-451: interpolate_field_to_particles on the three analytic fields, offset test points, the reference's
    regression (utils.py:73-105: slope of log(err) + 3 log(dx) against log(dx), must exceed 1.9) -- on nx = 30, 50, 70, 90 of the
    reference's 30..220 to keep the NumPy gather cheap."""
    from oracle import pusher as opush
    spec = {"wave": (2 * np.pi, 3.0, lambda X, Y, Z: np.sin(X) * np.cos(Y) * np.sin(Z)),
            "polynomial": (1.0, 2.0, lambda X, Y, Z: X ** 3 * Y + Y ** 3 * Z + Z ** 3 * X),
            "oscillatory": (np.pi, 4.0, lambda X, Y, Z: np.cos(2 * X) * np.sin(3 * Y) * np.cos(4 * Z))}[case]
    wind, frac, f = spec
    dxs, errs = [], []
    for nx in (30, 50, 70, 90):
        g = np.linspace(0, wind, nx)
        dx = wind / (nx - 1)
        X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
        t = np.linspace(dx / frac, wind - dx / frac, nx)
        Xt, Yt, Zt = (a.ravel() for a in np.meshgrid(t, t, t, indexing="ij"))
        got = opush.interpolate_field_to_particles(f(X, Y, Z), Xt, Yt, Zt, (g, g, g), shape_factor)
        errs.append(float(np.mean(np.abs(got - f(Xt, Yt, Zt)))))
        dxs.append(dx)
    slope = abs(np.polyfit(np.log(dxs), np.log(errs) + 3 * np.log(dxs), 1)[0])
    assert slope > 1.9
    assert abs(np.polyfit(np.log(dxs), np.log(errs), 1)[0]) > 1.8      # and the interpolation itself is second order


# ---- tests/physics_tests/pusher_schmitz_test.py:69-211 (q = m = C = 1; the reference's literal thresholds) -----
def _gamma(v):
    return 1.0 / np.sqrt(1.0 - sum(c * c for c in v))


def _v_from_u(u):
    g = np.sqrt(1.0 + sum(c * c for c in u))
    return tuple(c / g for c in u)


def _advance(fn, v, E, B, dt, steps):
    for _ in range(steps):
        v = tuple(float(c) for c in fn(v, E, B, 1.0, 1.0, dt, 1.0))
    return v


def test_schmitz_constant_B_preserves_energy():
    """pusher_schmitz_test.py:69-84"""
    from oracle import pusher as opush
    for gamma0 in (1.001, 10.0):
        v0 = _v_from_u((np.sqrt(gamma0 ** 2 - 1.0), 0.0, 0.0))
        for fn in (opush.relativistic_boris, opush.higuera_cary):
            v = _advance(fn, v0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 0.2 * np.pi * gamma0, 10)
            assert abs(_gamma(v) - gamma0) <= 2.0e-12


def test_schmitz_force_free_crossed_fields():
    """pusher_schmitz_test.py:86-123: Higuera-Cary keeps the force-free trajectory exact, relativistic Boris does not."""
    from oracle import pusher as opush
    gamma0 = 10.0
    v0 = np.sqrt(gamma0 ** 2 - 1.0) / gamma0
    E, B, dt = (0.0, v0, 0.0), (0.0, 0.0, 1.0), 0.2 * np.pi * gamma0
    worst = {}
    for name, fn in (("hc", opush.higuera_cary), ("boris", opush.relativistic_boris)):
        v, angle, err = (v0, 0.0, 0.0), 0.0, 0.0
        for _ in range(20):
            v = _advance(fn, v, E, B, dt, 1)
            angle = max(angle, abs(np.arctan2(v[1], v[0])))
            err = max(err, abs((_gamma(v) - gamma0) / (gamma0 - 1.0)))
        worst[name] = (angle, err)
    assert worst["hc"][0] < 1.0e-13 and worst["hc"][1] < 1.0e-12
    assert 1.0e-3 < worst["boris"][0] < 2.0e-1 and worst["boris"][1] > 1.0e-2


def test_schmitz_lorentz_boosted_gyration():
    """pusher_schmitz_test.py:125-183"""
    from oracle import pusher as opush
    gp, gf = 10.0, 5.0
    bf = np.sqrt(1.0 - 1.0 / gf ** 2)
    vp = np.sqrt(1.0 - 1.0 / gp ** 2)
    omega = 1.0 / gp
    radius = vp / omega
    dt, steps = 1.0e-3 * 2.0 * np.pi * gp * gf, 20
    lab_time = steps * dt
    tm = lab_time / gf
    for _ in range(8):
        res = gf * (tm + bf * radius * (1.0 - np.cos(omega * tm))) - lab_time
        tm -= res / (gf * (1.0 + bf * vp * np.sin(omega * tm)))
    vx_m, vy_m = vp * np.sin(omega * tm), vp * np.cos(omega * tm)
    den = 1.0 + bf * vx_m
    exact = ((vx_m + bf) / den, vy_m / (gf * den), 0.0)
    E, B, v0 = (0.0, bf * gf, 0.0), (0.0, 0.0, gf), (bf, vp / gf, 0.0)
    err = lambda v: np.sqrt(sum((v[i] - exact[i]) ** 2 for i in range(3)))
    assert err(_advance(opush.higuera_cary, v0, E, B, dt, steps)) < 1.0e-6
    assert err(_advance(opush.relativistic_boris, v0, E, B, dt, steps)) < 1.0e-3


def test_schmitz_oscillating_parallel_E_returns_to_initial_energy():
    """pusher_schmitz_test.py:185-211"""
    from oracle import pusher as opush
    gamma_perp = 1.1
    v0 = _v_from_u((np.sqrt(gamma_perp ** 2 - 1.0), 0.0, 0.0))
    omega0 = 0.5 * (1.0 / gamma_perp)
    E0 = 10.0 * omega0
    dt = (2.0 * np.pi / omega0) / 100
    for fn in (opush.relativistic_boris, opush.higuera_cary):
        v = v0
        for step in range(500):
            v = _advance(fn, v, (0.0, 0.0, E0 * np.cos(omega0 * (step + 0.5) * dt)), (0.0, 0.0, 1.0), dt, 1)
        assert abs(v[2]) <= 1.0e-10 and abs(_gamma(v) - gamma_perp) <= 1.0e-10


# ---------------------------------------------------------------- tiling fixture (tests/kernel_fixtures.py:529-594)
def test_fixture_tiled_particles_preserve_inactive_slots_and_metadata():
    """particle_initialization_test.py:27-94, literal inputs: the tile-major packing every parity test builds its particles with."""
    sp, dp = fx.kernel_parameters(Nx=4, Ny=2, Nz=1, x_wind=4.0, y_wind=2.0, z_wind=1.0, dx=1.0, dy=1.0, dz=1.0, dt=0.1, tile_shape=(2, 1, 1))
    ions = fx.particle_species("ions", 2.0, 3.0, weight=4.0, x1=[-1.5, 0.5, 1.5], x2=[-0.5, 0.5, 0.5], x3=[0.0, 0.0, 0.0],
                               u1=[0.1, 0.2, 0.3], u2=[0.0, 0.0, 0.0], u3=[1.0, 2.0, 3.0], update_x=(True, False, True),
                               update_u=(False, True, True), active_mask=[True, False, True])
    tp, sc = fx.build_tiled_particles([ions], sp, dp)
    assert tp.x.shape == (2, 2, 1, 1, 2, 3) and int(tp.active.sum()) == 2
    assert not hasattr(tp, "charge") and not hasattr(tp, "mass") and not hasattr(tp, "weight")
    assert np.allclose(tp.x[0, 0, 0, 0, 0], [-1.5, -0.5, 0.0]) and np.allclose(tp.x[1, 1, 0, 0, 0], [0.5, 0.5, 0.0])
    assert np.allclose(tp.x[1, 1, 0, 0, 1], [1.5, 0.5, 0.0]) and np.allclose(tp.u[1, 1, 0, 0, 1], [0.3, 0.0, 3.0])
    assert sc.charge.shape == sc.mass.shape == sc.weight.shape == (1,) and sc.update_x.shape == sc.update_u.shape == (1, 3)
    assert np.allclose(sc.charge, [2.0]) and np.allclose(sc.mass, [3.0]) and np.allclose(sc.weight, [4.0])
    assert list(sc.update_x[0]) == [True, False, True] and list(sc.update_u[0]) == [False, True, True]


def test_fixture_species_metadata_is_not_slot_shaped_and_capacity_headroom():
    """particle_initialization_test.py:94-186."""
    sp, dp = fx.kernel_parameters(Nx=4, Ny=1, Nz=1, x_wind=4.0, y_wind=1.0, z_wind=1.0, dx=1.0, dy=1.0, dz=1.0, dt=0.1, tile_shape=(1, 1, 1))
    sps = [fx.particle_species("electrons", -1.0, 2.0, weight=0.5, x1=[-1.5, -0.5]), fx.particle_species("ions", 2.0, 5.0, weight=0.25, x1=[0.5, 1.5])]
    tp, sc = fx.build_tiled_particles(sps, sp, dp)
    assert tp.x.shape[:4] == (4, 1, 1, 2)
    assert sc.charge.shape == sc.mass.shape == sc.weight.shape == (2,) and sc.update_x.shape == sc.update_u.shape == (2, 3)
    sp, dp = fx.kernel_parameters(Nx=4, Ny=1, Nz=1, x_wind=4.0, y_wind=1.0, z_wind=1.0, dx=1.0, dy=1.0, dz=1.0, dt=0.1, tile_shape=(2, 1, 1),
                                  particle_tile_capacity_factor=3.0)
    tp, sc = fx.build_tiled_particles([fx.particle_species("ions", 1.0, 1.0, x1=[-1.5, -0.5, 1.5])], sp, dp)
    assert tp.active.shape[-1] == 6 and int(tp.active.sum()) == 3 and sc.charge.shape == (1,)
