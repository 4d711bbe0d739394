"""Oracle cross-checks ported from the reference's invariants:
manual single-particle stencils, tile invariance (1e-12), discrete continuity, total charge, pusher physics."""
import itertools

import numpy as np
import pytest

from oracle import fixtures as fx, deposition as dep, particles as opart, pusher, yee, halo, diagnostics as diag, evolve
from tests import independent as ind

RT = dict(rtol=1e-12, atol=1e-12)


def _params_1d(sf, dep_mode="direct", tile=(4, 1, 1), **kw):
    # single_particle_pipeline_test.py:31-66
    return fx.kernel_parameters(Nx=8, Ny=1, Nz=1, x_wind=4.0, y_wind=1.0, z_wind=1.0, tile_shape=tile, guard_cells=2,
                                shape_factor=sf, current_deposition=dep_mode, relativistic=False, dt=0.05, **kw)


def _one(sp, dp, x, u):
    return fx.build_tiled_particles([fx.particle_species("single", -1.0, 1.0, weight=0.5, x1=[x[0]], x2=[x[1]], x3=[x[2]],
                                                         u1=[u[0]], u2=[u[1]], u3=[u[2]])], sp, dp)


POS = [(-1.32, 0.0, 0.0), (-0.03, 0.0, 0.0), (1.97, 0.0, 0.0)]   # interior / tile face / global boundary


@pytest.mark.parametrize("sf,x", list(itertools.product((1, 2), POS)))
def test_single_particle_rho(sf, x):
    # single_particle_pipeline_test.py:408-431
    sp, dp = _params_1d(sf)
    p, sc = _one(sp, dp, x, (0.0, 0.2, 0.0))
    rho = dep.compute_rho(p, sc, fx.empty_tiled_scalar(sp, dp), sp, dp)
    glob = diag.assemble_tiled_scalar_field(rho, sp.tile_shape, 2)[1:-1, 1, 1]
    assert np.allclose(glob, ind.manual_rho_1d(x[0], -0.5, 8, 4.0, sf), **RT)
    assert glob.sum() * dp.dx == pytest.approx(-0.5, abs=1e-12)
    # ghosts of the tiled result are consistent with a refresh
    assert np.allclose(rho, halo.update_tiled_ghost_cells(rho, sp, 2, bc_type=1), **RT)


@pytest.mark.parametrize("sf,x", list(itertools.product((1, 2), POS)))
def test_single_particle_direct_current(sf, x):
    # single_particle_pipeline_test.py:433-455
    u = (0.11, -0.17, 0.07)
    sp, dp = _params_1d(sf)
    p, sc = _one(sp, dp, x, u)
    J = dep.J_from_rhov(p, sc, fx.empty_tiled_vector(sp, dp), sp, dp)
    exp = ind.manual_direct_J_1d(x[0], u, -0.5, 8, 4.0, sf)
    for c in range(3):
        assert np.allclose(diag.assemble_tiled_scalar_field(J[c], sp.tile_shape, 2)[1:-1, 1, 1], exp[c], **RT)


@pytest.mark.parametrize("sf,x", list(itertools.product((1, 2), POS)))
def test_single_particle_esirkepov_1d(sf, x):
    # single_particle_pipeline_test.py:457-485
    u = (0.08, -0.17, 0.07)
    sp, dp = _params_1d(sf, "esirkepov")
    p, sc = _one(sp, dp, x, u)
    J = dep.Esirkepov_current(p, sc, fx.empty_tiled_vector(sp, dp), sp, dp)
    exp = ind.manual_esirkepov_J_1d(x[0], u, -0.5, 8, 4.0, sf, dp.dt)
    for c in range(3):
        assert np.allclose(diag.assemble_tiled_scalar_field(J[c], sp.tile_shape, 2)[1:-1, 1, 1], exp[c], **RT)


def test_esirkepov_step_E_equals_minus_dt_J_over_eps():
    # single_particle_pipeline_test.py:543-586 (E pin is valid; the B pins at :588-599 are stale)
    sp, dp = _params_1d(1, "esirkepov")
    x0, u0 = (-1.32, 0.0, 0.0), (0.08, 0.2, 0.0)
    p, sc = _one(sp, dp, x0, u0)
    z = fx.empty_tiled_vector
    fields = (z(sp, dp), z(sp, dp), z(sp, dp), fx.empty_tiled_scalar(sp, dp), fx.empty_tiled_scalar(sp, dp), (z(sp, dp), z(sp, dp)), None, False)
    pa, fa = evolve.time_loop_electrodynamic(p, sc, fields, sp, dp)
    Jexp = dep.Esirkepov_current(p, sc, z(sp, dp), sp, dp)
    for c in range(3):
        assert np.allclose(fa[2][c], Jexp[c], **RT)
        Eg = diag.assemble_tiled_scalar_field(fa[0][c], sp.tile_shape, 2)[1:-1, 1:-1, 1:-1]
        Jg = diag.assemble_tiled_scalar_field(Jexp[c], sp.tile_shape, 2)[1:-1, 1:-1, 1:-1]
        assert np.allclose(Eg, -dp.dt * Jg / dp.eps, **RT)
    assert np.allclose(pa.x[..., 0][pa.active], [x0[0] + u0[0] * dp.dt], **RT)
    assert not fa[-1]
    # code-is-authority B: two half steps with E_old=0 then E_new  =>  Bz = -(dt/2) dEy/dx
    Ey = diag.assemble_tiled_scalar_field(fa[0][1], sp.tile_shape, 2)
    Bz = diag.assemble_tiled_scalar_field(fa[1][2], sp.tile_shape, 2)
    assert np.allclose(Bz[1:-1, 1, 1], -(dp.dt / 2) * (Ey[2:, 1, 1] - Ey[1:-1, 1, 1]) / dp.dx, **RT)


# ---- halo: axis-sequential oracle == one-shot global-index formulation --------------------------------
@pytest.mark.parametrize("mesh,tile,g", [((1, 1, 1), (3, 2, 2), 1), ((2, 2, 1), (2, 2, 1), 1), ((2, 1, 2), (2, 3, 2), 2),
                                         ((4, 2, 2), (2, 3, 2), 2), ((1, 1, 1), (1, 3, 2), 2)])
@pytest.mark.parametrize("bcs", [(0, 0, 0), (1, 0, 0), (0, 1, 2), (2, 2, 1), (1, 1, 1)])
def test_halo_global_index_equivalence(mesh, tile, g, bcs):
    rng = np.random.default_rng(7)
    t = rng.normal(size=tuple(mesh) + tuple(w + 2 * g for w in tile))
    assert np.allclose(halo.refresh(t, tile, bcs, g), ind.global_index_refresh(t, tile, bcs, g), **RT)
    assert np.allclose(halo.fold(t, tile, bcs, g), ind.global_index_fold(t, tile, bcs, g), **RT)


# ---- tile invariance at 1e-12 (yee_test.py:467-675, pusher_test.py:199-438, esirkepov_test.py:396-451,
#      direct_deposition_test.py:362-449, rho_test.py:246-253) ------------------------------------------
def _det_vector_field(n, seed):
    rng = np.random.default_rng(seed)
    return tuple(rng.normal(size=(n[0] + 2, n[1] + 2, n[2] + 2)) for _ in range(3))


def _plasma(sp, dp, seed=3, n=40, vmax=0.3):
    rng = np.random.default_rng(seed)
    sps = []
    for s, (q, m) in enumerate(((-1.0, 1.0), (2.0, 5.0))):
        pos = [rng.uniform(-w / 2, w / 2, n) if N > 1 else np.zeros(n) for w, N in ((dp.x_wind, dp.Nx), (dp.y_wind, dp.Ny), (dp.z_wind, dp.Nz))]
        vel = [rng.uniform(-vmax, vmax, n) for _ in range(3)]
        sps.append(fx.particle_species(f"s{s}", q, m, weight=0.5 + s, x1=pos[0], x2=pos[1], x3=pos[2], u1=vel[0], u2=vel[1], u3=vel[2]))
    return sps


CASES = [((8, 6, 4), (2, 3, 2)), ((8, 6, 4), (4, 2, 2)), ((8, 1, 1), (2, 1, 1)), ((6, 6, 1), (3, 2, 1)), ((1, 6, 4), (1, 3, 2))]


def _assemble_v(F, sp):
    return [diag.assemble_tiled_scalar_field(c, sp.tile_shape, sp.guard_cells) for c in F]


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
def test_tile_invariance_all_kernels(N, tile, sf):
    kw = dict(Nx=N[0], Ny=N[1], Nz=N[2], x_wind=4.0 if N[0] > 1 else 1.0, y_wind=3.0 if N[1] > 1 else 1.0,
              z_wind=2.0 if N[2] > 1 else 1.0, shape_factor=sf, dt=0.05, particle_tile_capacity_factor=3.0)
    sp1, dp1 = fx.kernel_parameters(**kw)
    spt, dpt = fx.kernel_parameters(tile_shape=tile, **kw)
    species = _plasma(sp1, dp1)
    Eg, Bg = _det_vector_field(N, 1), _det_vector_field(N, 2)
    res = []
    for sp, dp in ((sp1, dp1), (spt, dpt)):
        p, sc = fx.build_tiled_particles(species, sp, dp)
        E, B = fx.vector_tiles_from_global(Eg, sp, dp), fx.vector_tiles_from_global(Bg, sp, dp)
        pushed = pusher.particle_push(p, sc, E, B, sp, dp)
        Jz = fx.empty_tiled_vector(sp, dp)
        out = {"Je": _assemble_v(dep.Esirkepov_current(pushed, sc, Jz, sp, dp), sp),
               "Jd": _assemble_v(dep.J_from_rhov(pushed, sc, Jz, sp, dp), sp),
               "rho": [diag.assemble_tiled_scalar_field(dep.compute_rho(pushed, sc, fx.empty_tiled_scalar(sp, dp), sp, dp), sp.tile_shape, 2)],
               "B": _assemble_v(yee.update_B(E, B, sp, dp), sp)}
        out["E"] = _assemble_v(yee.update_E(E, B, fx.vector_tiles_from_global(Eg, sp, dp), sp, dp), sp)
        a = pushed.active.reshape(-1)
        key = np.lexsort(pushed.x.reshape(-1, 3)[a].T)
        out["u"] = [pushed.u.reshape(-1, 3)[a][key]]
        res.append(out)
    for k in res[0]:
        for a, b in zip(res[0][k], res[1][k]):
            assert np.allclose(a, b, **RT), k


# ---- continuity (esirkepov_test.py:700-744) in 1-D, 2-D and 3-D, both shapes, multi-tile --------------
@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
def test_esirkepov_continuity(N, tile, sf):
    kw = dict(Nx=N[0], Ny=N[1], Nz=N[2], x_wind=4.0 if N[0] > 1 else 1.0, y_wind=3.0 if N[1] > 1 else 1.0,
              z_wind=2.0 if N[2] > 1 else 1.0, shape_factor=sf, dt=0.05, tile_shape=tile, particle_tile_capacity_factor=3.0,
              current_deposition="esirkepov")
    sp, dp = fx.kernel_parameters(**kw)
    dmin = min(dp.dx, dp.dy, dp.dz)
    p, sc = fx.build_tiled_particles(_plasma(sp, dp, vmax=0.45 * dmin / dp.dt), sp, dp)
    rho0 = dep.compute_rho(p, sc, fx.empty_tiled_scalar(sp, dp), sp, dp)
    J = dep.Esirkepov_current(p, sc, fx.empty_tiled_vector(sp, dp), sp, dp)
    p1 = opart.update_tiled_particle_positions(p, sc, dp.dt)
    p1, ovf = opart.refresh_tiled_particle_tiles(p1, sp, dp)
    assert not ovf
    rho1 = dep.compute_rho(p1, sc, fx.empty_tiled_scalar(sp, dp), sp, dp)
    res = diag.continuity_residual(rho0, rho1, J, sp, dp)
    scale = max(1.0, np.abs(rho1 - rho0).max() / dp.dt)
    assert np.abs(res).max() <= 1e-11 * scale


# ---- pusher physics (physics_tests/pusher_schmitz_test.py:69-123) -------------------------------------
@pytest.mark.parametrize("fn", (pusher.relativistic_boris, pusher.higuera_cary, pusher.boris))
def test_pure_B_conserves_speed(fn):
    v = (np.array([0.3]), np.array([-0.2]), np.array([0.1]))
    z = np.zeros(1)
    for _ in range(50):
        v = fn(v, (z, z, z), (z + 0.3, z - 0.4, z + 1.2), -1.0, 1.0, 0.1, 1.0)
    assert np.isclose(v[0] ** 2 + v[1] ** 2 + v[2] ** 2, 0.14, rtol=2e-12)


def test_force_free_ExB_drift():
    # E = -v x B  => velocity unchanged (relativistic Boris is not exactly force-free; HC/Boris keep it to O(dt^2))
    v = (np.array([0.0]), np.array([0.1]), np.array([0.0]))
    B = (np.zeros(1), np.zeros(1), np.ones(1))
    E = (-(v[1] * B[2]), np.zeros(1), np.zeros(1))
    out = pusher.boris(v, E, B, 1.0, 1.0, 0.01, 1.0)
    assert np.allclose([o[0] for o in out], [0.0, 0.1, 0.0], atol=1e-6)


# ---- rho properties (tests/code_tests/rho_test.py:255-345) ---------------------------------------------
def _rho_case(sf=2, **kw):
    from tests.cases import make_case
    sp, dp, tp, sc, E, B = make_case((8, 6, 4), (4, 3, 2), sf, **kw)
    return sp, dp, tp, sc


def test_rho_uses_current_positions_not_velocities():
    """rho_test.py:329-369: the same positions with zeroed velocities deposit the same rho."""
    sp, dp, tp, sc = _rho_case()
    z = fx.empty_tiled_scalar(sp, dp)
    a = dep.compute_rho(tp, sc, z, sp, dp)
    b = dep.compute_rho(tp._replace(u=np.zeros_like(tp.u)), sc, z, sp, dp)
    assert np.allclose(a, b, rtol=1e-12, atol=1e-12)


def test_rho_ghost_folding_uses_the_particle_boundary_conditions():
    """rho_test.py:301-327: periodic vs absorbing particle BC on x change the folded rho (the field BCs are identical)."""
    sp0, dp0, tp, sc = _rho_case(sf=1, particle_boundary_conditions=(0, 0, 0))
    sp1, dp1, _, _ = _rho_case(sf=1, particle_boundary_conditions=(2, 0, 0))
    a = dep.compute_rho(tp, sc, fx.empty_tiled_scalar(sp0, dp0), sp0, dp0)
    b = dep.compute_rho(tp, sc, fx.empty_tiled_scalar(sp1, dp1), sp1, dp1)
    assert float(np.abs(np.asarray(a) - np.asarray(b)).max()) > 1e-12


def test_rho_digital_filter_depends_on_alpha():
    """rho_test.py:255-299: current_filter = "digital": rho(alpha = 0.55) == digital_filter(rho(alpha = 1), 0.55) + ghost refresh."""
    from oracle import filters as ofil, halo as ohalo
    sp, dp, tp, sc = _rho_case(current_filter="digital", alpha=1.0)
    z = fx.empty_tiled_scalar(sp, dp)
    r10 = np.asarray(dep.compute_rho(tp, sc, z, sp, dp))
    r055 = np.asarray(dep.compute_rho(tp, sc, z, sp, dp._replace(alpha=0.55)))
    g = int(sp.guard_cells)
    want = ohalo.update_tiled_ghost_cells(ofil.digital_filter(r10, 0.55, num_guard_cells=g), sp, g, bc_type=1)
    assert np.allclose(r055, want, rtol=1e-13, atol=1e-13) and np.abs(r055 - r10).max() > 1e-6
