"""-m gpu parity tests: every CUDA operator, called through the C ABI (libpic_b200.so) from the reference-shaped
Python API, against the NumPy oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star / BASELINE.md section 4):
  float64  single-kernel and few-step results: 1e-12 relative to the field scale (the reference's own tile-invariance
           tolerance, tests/code_tests/yee_test.py:489)
  float32  2e-5 relative to the field scale after a few steps (north_star: "<=1e-5 fp32 ... looser after divergence")
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import fixtures as fx, pusher as opush, deposition as odep, particles as opart, yee as oyee, halo as ohalo, filters as ofil
from oracle import evolve as oevolve, diagnostics as odiag
from tests.cases import CASES, make_case, make_fields
from tests import gpu_util as gu
from tests import independent as ind

pytestmark = pytest.mark.gpu

F64, F32 = torch.float64, torch.float32
TOL = {F64: 1e-12, F32: 2e-5}


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    gu.require_cuda()
    from pypic3d_b200 import _lib
    assert _lib.lib().pic_version()  # the product library is loaded (no fallback exists)


# ------------------------------------------------------------------------------------------------ push
@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("pn,rel", [("boris", True), ("boris", False), ("higuera_cary", True)])
@pytest.mark.parametrize("dtype", (F64, F32))
def test_particle_push(N, tile, sf, pn, rel, dtype):
    from pypic3d_b200.pusher.particle_push import particle_push
    sp, dp, tp, sc, E, B = make_case(N, tile, sf, particle_pusher=pn, relativistic=rel)
    ref = opush.particle_push(tp, sc, E, B, sp, dp)
    ps, pd = gu.to_pkg_params(sp, dp)
    out = particle_push(gu.particles_to_gpu(tp, dtype), gu.species_to_pkg(sc), gu.vec_to_gpu(E, dtype), gu.vec_to_gpu(B, dtype), ps, pd)
    gu.assert_close(out.u, ref.u, TOL[dtype] * (1 if dtype == F64 else 5), "u")
    assert torch.equal(out.active, gu.tt(tp.active))


# ------------------------------------------------------------------------------------------------ deposits
@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("dtype", (F64, F32))
def test_esirkepov_current(N, tile, sf, dtype):
    from pypic3d_b200.deposition.Esirkepov import Esirkepov_current
    sp, dp, tp, sc, E, B = make_case(N, tile, sf, current_deposition="esirkepov")
    z = fx.empty_tiled_vector(sp, dp)
    ref = odep.Esirkepov_current(tp, sc, z, sp, dp)
    ps, pd = gu.to_pkg_params(sp, dp)
    J = Esirkepov_current(gu.particles_to_gpu(tp, dtype), gu.species_to_pkg(sc), gu.vec_to_gpu(z, dtype), ps, pd)
    scale = max(np.abs(r).max() for r in ref)
    for c in range(3):
        gu.assert_close(J[c] / scale, ref[c] / scale, TOL[dtype], f"J{c}")


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("filt", ("none", "bilinear", "digital"))
def test_J_from_rhov(N, tile, sf, filt):
    from pypic3d_b200.deposition.J_from_rhov import J_from_rhov
    sp, dp, tp, sc, E, B = make_case(N, tile, sf, current_filter=filt, alpha=0.6)
    z = fx.empty_tiled_vector(sp, dp)
    ref = odep.J_from_rhov(tp, sc, z, sp, dp)
    ps, pd = gu.to_pkg_params(sp, dp)
    J = J_from_rhov(gu.particles_to_gpu(tp), gu.species_to_pkg(sc), gu.vec_to_gpu(z), ps, pd)
    scale = max(np.abs(r).max() for r in ref)
    for c in range(3):
        gu.assert_close(J[c] / scale, ref[c] / scale, 1e-12, f"J{c}")


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("filt", ("none", "digital"))
def test_compute_rho(N, tile, sf, filt):
    from pypic3d_b200.deposition.rho import compute_rho
    sp, dp, tp, sc, E, B = make_case(N, tile, sf, current_filter=filt, alpha=0.7)
    ref = odep.compute_rho(tp, sc, fx.empty_tiled_scalar(sp, dp), sp, dp)
    ps, pd = gu.to_pkg_params(sp, dp)
    rho = compute_rho(gu.particles_to_gpu(tp), gu.species_to_pkg(sc), gu.tt(fx.empty_tiled_scalar(sp, dp)), ps, pd)
    gu.assert_close(rho / np.abs(ref).max(), ref / np.abs(ref).max(), 1e-12, "rho")


@pytest.mark.parametrize("N,tile", CASES[:3])
@pytest.mark.parametrize("sf", (1, 2))
def test_direct_deposits_float32(N, tile, sf):
    """J_from_rhov (+ bilinear filter), compute_rho and the position update in the throughput dtype: same operators, float32 leaves,
    against the float64 oracle at the float32 tolerance of the module header."""
    from pypic3d_b200.deposition.J_from_rhov import J_from_rhov
    from pypic3d_b200.deposition.rho import compute_rho
    from pypic3d_b200.particles.particle_tile_communication import update_tiled_particle_positions
    sp, dp, tp, sc, E, B = make_case(N, tile, sf, current_filter="bilinear", alpha=0.6)
    ps, pd = gu.to_pkg_params(sp, dp)
    gp, gs = gu.particles_to_gpu(tp, F32), gu.species_to_pkg(sc)
    z = fx.empty_tiled_vector(sp, dp)
    ref = odep.J_from_rhov(tp, sc, z, sp, dp)
    J = J_from_rhov(gp, gs, gu.vec_to_gpu(z, F32), ps, pd)
    assert J[0].dtype == F32
    scale = max(np.abs(r).max() for r in ref)
    for c in range(3):
        gu.assert_close(J[c] / scale, ref[c] / scale, TOL[F32], f"J{c} f32")
    rref = odep.compute_rho(tp, sc, fx.empty_tiled_scalar(sp, dp), sp, dp)
    rho = compute_rho(gp, gs, gu.tt(fx.empty_tiled_scalar(sp, dp), F32), ps, pd)
    gu.assert_close(rho / np.abs(rref).max(), rref / np.abs(rref).max(), TOL[F32], "rho f32")
    moved = update_tiled_particle_positions(gp, gs, 1.0)
    gu.assert_close(moved.x, opart.update_tiled_particle_positions(tp, sc, 1.0).x, 1e-6, "move f32")


@pytest.mark.parametrize("sf,x", list(itertools.product((1, 2), [(-1.32, 0.0, 0.0), (-0.03, 0.0, 0.0), (1.97, 0.0, 0.0)])))
def test_single_particle_manual_stencils(sf, x):
    """tests/code_tests/single_particle_pipeline_test.py:408-485 on the GPU path (independent scalar restatement)."""
    from pypic3d_b200.deposition import Esirkepov_current, J_from_rhov, compute_rho
    u = (0.08, -0.17, 0.07)
    sp, dp = fx.kernel_parameters(Nx=8, Ny=1, Nz=1, x_wind=4.0, y_wind=1.0, z_wind=1.0, tile_shape=(4, 1, 1), shape_factor=sf,
                                  relativistic=False, dt=0.05, current_deposition="esirkepov")
    tp, sc = fx.build_tiled_particles([fx.particle_species("single", -1.0, 1.0, weight=0.5, x1=[x[0]], x2=[x[1]], x3=[x[2]],
                                                           u1=[u[0]], u2=[u[1]], u3=[u[2]])], sp, dp)
    ps, pd = gu.to_pkg_params(sp, dp)
    gp, gs, z = gu.particles_to_gpu(tp), gu.species_to_pkg(sc), gu.vec_to_gpu(fx.empty_tiled_vector(sp, dp))
    glob = lambda f: odiag.assemble_tiled_scalar_field(gu.npy(f), sp.tile_shape, 2)[1:-1, 1, 1]
    assert np.allclose(glob(compute_rho(gp, gs, z[0], ps, pd)), ind.manual_rho_1d(x[0], -0.5, 8, 4.0, sf), rtol=1e-12, atol=1e-12)
    Jd, Je = J_from_rhov(gp, gs, z, ps, pd), Esirkepov_current(gp, gs, z, ps, pd)
    ed, ee = ind.manual_direct_J_1d(x[0], u, -0.5, 8, 4.0, sf), ind.manual_esirkepov_J_1d(x[0], u, -0.5, 8, 4.0, sf, dp.dt)
    for c in range(3):
        assert np.allclose(glob(Jd[c]), ed[c], rtol=1e-12, atol=1e-12)
        assert np.allclose(glob(Je[c]), ee[c], rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------------------------------------ halo / filters / Yee
@pytest.mark.parametrize("mesh,tile,g", [((1, 1, 1), (3, 2, 2), 1), ((2, 2, 1), (2, 2, 1), 1), ((2, 1, 2), (2, 3, 2), 2),
                                         ((4, 2, 2), (2, 3, 2), 2), ((1, 1, 1), (1, 3, 2), 2), ((1, 1, 1), (16, 1, 1), 2)])
@pytest.mark.parametrize("bcs", [(0, 0, 0), (1, 0, 0), (0, 1, 2), (2, 2, 1), (1, 1, 1)])
@pytest.mark.parametrize("dtype", (F64, F32))
def test_halo_refresh_and_fold(mesh, tile, g, bcs, dtype):
    from pypic3d_b200.boundary_conditions import ghost_cells as gc

    class SP:
        tile_shape, boundary_conditions, particle_boundary_conditions, field_mesh = tile, bcs, bcs, mesh
    t = np.random.default_rng(7).normal(size=tuple(mesh) + tuple(w + 2 * g for w in tile))
    if dtype == F32:
        t = t.astype(np.float32).astype(np.float64)
    gu.assert_close(gc.update_tiled_ghost_cells(gu.tt(t, dtype), SP, g), ohalo.refresh(t, tile, bcs, g), 1e-15 if dtype == F64 else 1e-7, "refresh")
    gu.assert_close(gc.fold_tiled_ghost_cells(gu.tt(t, dtype), SP, g), ohalo.fold(t, tile, bcs, g), 1e-14 if dtype == F64 else 1e-6, "fold")
    v = gc.update_tiled_vector_ghost_cells(tuple(gu.tt(t + k, dtype) for k in range(3)), SP, g, bc_type=1)
    for k in range(3):
        gu.assert_close(v[k], ohalo.refresh(t + k, tile, bcs, g), 1e-15 if dtype == F64 else 1e-6, "vector refresh")


def test_halo_kats_from_reference():
    """distributed_ghost_cells_test.py:276-310 literal KATs (g=1)."""
    from pypic3d_b200.boundary_conditions import ghost_cells as gc

    class SP:
        tile_shape, boundary_conditions, particle_boundary_conditions, field_mesh = (3, 2, 2), (0, 0, 0), (2, 0, 0), (1, 1, 1)
    t = np.zeros((1, 1, 1, 5, 4, 4)); t[0, 0, 0, 0, 1:-1, 1:-1] = 3.0; t[0, 0, 0, -1, 1:-1, 1:-1] = 5.0
    out = gu.npy(gc.fold_tiled_ghost_cells(gu.tt(t), SP, 1))
    assert np.allclose(out[0, 0, 0, -2, 1:-1, 1:-1], 3.0) and np.allclose(out[0, 0, 0, 1, 1:-1, 1:-1], 5.0) and np.allclose(out[0, 0, 0, 0], 0)
    out = gu.npy(gc.fold_tiled_ghost_cells(gu.tt(t), SP, 1, bc_type=1))
    assert np.allclose(out, 0.0)
    SP.boundary_conditions = (1, 0, 0)
    out = gu.npy(gc.fold_tiled_ghost_cells(gu.tt(t), SP, 1))
    assert np.allclose(out[0, 0, 0, 1, 1:-1, 1:-1], -3.0) and np.allclose(out[0, 0, 0, -2, 1:-1, 1:-1], -5.0)


@pytest.mark.parametrize("g", (1, 2))
@pytest.mark.parametrize("dtype", (F64, F32))
def test_filters(g, dtype):
    from pypic3d_b200.utilities.filters import bilinear_filter, digital_filter, digital_filter_vector
    t = np.random.default_rng(3).normal(size=(2, 1, 2, 5 + 2 * g, 4 + 2 * g, 3 + 2 * g))
    tol = 1e-14 if dtype == F64 else 1e-6
    gu.assert_close(bilinear_filter(gu.tt(t, dtype), g), ofil.bilinear_filter(t, g), tol, "bilinear")
    gu.assert_close(digital_filter(gu.tt(t, dtype), 0.6, g), ofil.digital_filter(t, 0.6, g), tol, "digital")
    v = digital_filter_vector(tuple(gu.tt(t + k, dtype) for k in range(3)), 0.5, g)
    for k in range(3):
        gu.assert_close(v[k], ofil.digital_filter(t + k, 0.5, g), tol, "digital vector")


@pytest.mark.parametrize("N,tile", CASES + [((16, 12, 8), (16, 12, 8))])
@pytest.mark.parametrize("bcs", [(0, 0, 0), (1, 0, 1)])
@pytest.mark.parametrize("alpha", (1.0, 0.8))
@pytest.mark.parametrize("dtype", (F64, F32))
def test_yee_update_E_and_B(N, tile, bcs, alpha, dtype):
    from pypic3d_b200.solvers.first_order_yee import update_E, update_B
    sp, dp, tp, sc, E, B = make_case(N, tile, 1, boundary_conditions=bcs, alpha=alpha, C=1.3, eps=0.7)
    J = tuple(0.3 * c for c in B)
    ps, pd = gu.to_pkg_params(sp, dp)
    Eg, Bg, Jg = gu.vec_to_gpu(E, dtype), gu.vec_to_gpu(B, dtype), gu.vec_to_gpu(J, dtype)
    tol = TOL[dtype]
    out, pml = update_E(Eg, Bg, Jg, ps, pd)
    assert pml is None
    for a, b in zip(out, oyee.update_E(E, B, J, sp, dp)):
        gu.assert_close(a, b, tol, "update_E")
    for flt in (False, True):
        out, _ = update_B(Eg, Bg, ps, pd, None, do_filter=flt)
        for a, b in zip(out, oyee.update_B(E, B, sp, dp, do_filter=flt)):
            gu.assert_close(a, b, tol, "update_B")


@pytest.mark.parametrize("N", [(8, 6, 4), (8, 1, 4), (36, 9, 5), (4, 20, 70)])
@pytest.mark.parametrize("bcs", [(0, 0, 0), (1, 0, 1), (0, 1, 0)])
@pytest.mark.parametrize("dtype", (F64, F32))
def test_fused_yee_equals_the_three_sweeps(N, bcs, dtype):
    """pic_yee_fused (B half -> E -> B half in one pass) against (a) the CUDA sweeps + guard-cell refreshes it replaces -- bit for
    bit, guard cells included -- and (b) the oracle's update_B / update_E / update_B sequence (evolve.py:88-96).  Grids with a
    reduced axis, widths that are not multiples of the 4 x 8 x 32 tile, periodic and conducting walls."""
    import ctypes
    from pypic3d_b200 import _lib, ops
    from pypic3d_b200.solvers.first_order_yee import update_E, update_B
    sp, dp, tp, sc, E, B = make_case(N, N, 1, boundary_conditions=bcs, alpha=1.0, C=1.3, eps=0.7)
    rng = np.random.default_rng(5)
    J = tuple(rng.normal(size=c.shape) for c in B)
    # guard cells as the step finds them: refreshed E, B (field BCs); J folded + refreshed with the particle BCs (periodic here)
    E = ohalo.update_tiled_vector_ghost_cells(E, sp, 2); B = ohalo.update_tiled_vector_ghost_cells(B, sp, 2)
    J = ohalo.update_tiled_vector_ghost_cells(J, sp, 2, bc_type=1)
    ps, pd = gu.to_pkg_params(sp, dp)
    Eg, Bg, Jg = gu.vec_to_gpu(E, dtype), gu.vec_to_gpu(B, dtype), gu.vec_to_gpu(J, dtype)
    p = ops.params_for(ps, pd, None, Eg[0])
    E2 = [torch.zeros_like(c) for c in Eg]; B2 = [torch.zeros_like(c) for c in Bg]
    _lib.check(_lib.lib().pic_yee_fused(ctypes.byref(p), ops._v(Eg), ops._v(Bg), ops._v(Jg), ops._v(E2), ops._v(B2), ops._stream()), "pic_yee_fused")
    for axis in range(3):            # a reduced periodic axis is read from its guard cells and refreshed by the caller
        if N[axis] < 2 and bcs[axis] == 0:
            ops.halo_refresh_axis_(p, axis, 0, E2 + B2)
    # (a) the sweeps of the drop-in path
    Bh, _ = update_B(Eg, Bg, ps, pd, None, do_filter=False)
    En, _ = update_E(Eg, Bh, Jg, ps, pd)
    Bn, _ = update_B(En, Bh, ps, pd, None, do_filter=False)
    for a, b in zip(E2 + B2, tuple(En) + tuple(Bn)):
        assert torch.equal(a, b)
    # (b) the oracle
    oBh = oyee.update_B(E, B, sp, dp, do_filter=False)
    oEn = oyee.update_E(E, oBh, J, sp, dp)
    oBn = oyee.update_B(oEn, oBh, sp, dp, do_filter=False)
    for a, b in zip(E2 + B2, tuple(oEn) + tuple(oBn)):
        gu.assert_close(a, b, TOL[dtype], "fused Yee")


def test_update_E_kat_from_reference():
    """esirkepov_test.py:510-529: J=4, dt=.25, eps=2 -> E=-0.5."""
    from pypic3d_b200.solvers.first_order_yee import update_E
    sp, dp = fx.kernel_parameters(Nx=4, Ny=1, Nz=1, x_wind=4.0, y_wind=1.0, z_wind=1.0, dt=0.25, tile_shape=(2, 1, 1), C=1.0, eps=2.0)
    z = fx.empty_tiled_vector(sp, dp)
    J = fx.empty_tiled_vector(sp, dp); J[0][:, :, :, 2:-2, 2:-2, 2:-2] = 4.0
    ps, pd = gu.to_pkg_params(sp, dp)
    out, _ = update_E(gu.vec_to_gpu(z), gu.vec_to_gpu(z), gu.vec_to_gpu(J), ps, pd)
    assert np.allclose(gu.npy(out[0])[:, :, :, 2:-2, 2:-2, 2:-2], -0.5)


# ------------------------------------------------------------------------------------------------ move / retile
@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("pbc", [(0, 0, 0), (1, 0, 2), (2, 1, 0)])
@pytest.mark.parametrize("capacity", (2.0, 1.0))
def test_move_and_retile(N, tile, pbc, capacity):
    from pypic3d_b200.particles.particle_tile_communication import update_tiled_particle_positions, refresh_tiled_particle_tiles
    sp, dp, tp, sc, E, B = make_case(N, tile, 1, particle_boundary_conditions=pbc, capacity=capacity, vmax=0.35)
    ps, pd = gu.to_pkg_params(sp, dp)
    gp, gs = gu.particles_to_gpu(tp), gu.species_to_pkg(sc)
    moved_ref = opart.update_tiled_particle_positions(tp, sc, 1.0)
    moved = update_tiled_particle_positions(gp, gs, 1.0)
    gu.assert_close(moved.x, moved_ref.x, 1e-15, "move")
    ref, ovf = opart.refresh_tiled_particle_tiles(moved_ref, sp, dp)
    out, govf = refresh_tiled_particle_tiles(moved, ps, pd)
    assert bool(govf.item()) == bool(ovf)
    assert np.array_equal(gu.npy(out.active), ref.active)           # slot-exact (k-th incoming -> k-th free slot)
    gu.assert_close(out.x, ref.x, 1e-14, "retile x")
    gu.assert_close(out.u, ref.u, 1e-14, "retile u")


def test_retile_kats_from_reference():
    """particle_refresh_test.py:110-183 + distributed_particle_refresh_test.py:105-169."""
    from pypic3d_b200.particles.particle_tile_communication import refresh_tiled_particle_tiles
    import pypic3d_b200 as pp

    def run(mesh, tile, entries, slots=2, pbc=(0, 0, 0)):
        n = [mesh[a] * tile[a] for a in range(3)]
        sp, dp = fx.kernel_parameters(Nx=n[0], Ny=n[1], Nz=n[2], x_wind=float(n[0]), y_wind=float(n[1]), z_wind=float(n[2]),
                                      tile_shape=tile, particle_boundary_conditions=pbc)
        x = np.zeros(mesh + (1, slots, 3)); u = np.zeros_like(x); a = np.zeros(mesh + (1, slots), bool)
        for t, s, pos, vel in entries:
            x[t + (0, s)] = pos; u[t + (0, s)] = vel; a[t + (0, s)] = True
        ps, pd = gu.to_pkg_params(sp, dp)
        out, ovf = refresh_tiled_particle_tiles(pp.TiledParticles(gu.tt(x), gu.tt(u), gu.tt(a)), ps, pd)
        return gu.npy(out.x), gu.npy(out.u), gu.npy(out.active), bool(ovf.item())
    x, u, a, o = run((2, 1, 1), (2, 1, 1), [((0, 0, 0), 0, (0.25, 0, 0), (0, 0, 0))])
    assert not o and a[0].sum() == 0 and a[1].sum() == 1 and np.isclose(x[1, 0, 0, 0, 0, 0], 0.25)
    x, u, a, o = run((2, 1, 1), (2, 1, 1), [((1, 0, 0), 0, (2.25, 0, 0), (0, 0, 0))])
    assert not o and a[1].sum() == 0 and a[0].sum() == 1 and np.isclose(x[0, 0, 0, 0, 0, 0], -1.75)
    x, u, a, o = run((2, 2, 1), (2, 2, 1), [((0, 0, 0), 0, (0.25, 0.25, 0), (0, 0, 0))])
    assert not o and a[0, 0].sum() == 0 and a[1, 1].sum() == 1 and np.allclose(x[1, 1, 0, 0, 0, :2], [0.25, 0.25])
    x, u, a, o = run((2, 1, 1), (2, 1, 1), [((0, 0, 0), 0, (0.25, 0, 0), (0, 0, 0)), ((1, 0, 0), 0, (1.25, 0, 0), (0, 0, 0))], slots=1)
    assert o and a.sum() == 1
    x, u, a, o = run((2, 1, 1), (2, 1, 1), [((1, 0, 0), 0, (2.25, 0, 0), (0.5, 0, 0)), ((0, 0, 0), 0, (-2.10, 0, 0), (-0.25, 0, 0))], pbc=(1, 0, 0))
    assert not o and np.allclose(np.sort(x[a][:, 0]), [-1.90, 1.75]) and np.allclose(np.sort(u[a][:, 0]), [-0.5, 0.25])
    x, u, a, o = run((2, 1, 1), (2, 1, 1), [((1, 0, 0), 0, (2.25, 0, 0), (0.5, 0, 0)), ((0, 0, 0), 0, (-0.25, 0, 0), (0, 0, 0))], pbc=(2, 0, 0))
    assert not o and a.sum() == 1 and np.allclose(x[a][:, 0], [-0.25])


# ------------------------------------------------------------------------------------------------ energy
def test_energy_kats_from_reference():
    """utils_test.py:286-346."""
    from pypic3d_b200.utils import compute_energy, compute_total_momentum
    import pypic3d_b200 as pp
    sp, dp = fx.kernel_parameters(Nx=1, Ny=1, Nz=1, x_wind=1.0, y_wind=1.0, z_wind=1.0, eps=2.0, mu=4.0, C=10.0)
    E, B = fx.empty_tiled_vector(sp, dp), fx.empty_tiled_vector(sp, dp)
    E[0][0, 0, 0, 2, 2, 2] = 2.0; B[1][0, 0, 0, 2, 2, 2] = 3.0
    ps, pd = gu.to_pkg_params(sp, dp)
    sc = pp.SpeciesConfig(np.array([1.0]), np.array([1.0]), np.array([1.0]), np.ones((1, 3), bool), np.ones((1, 3), bool))
    tp = pp.TiledParticles(gu.tt(np.array([[[[[[0.6, 0, 0]]]]]])), gu.tt(np.array([[[[[[1.0, 0, 0]]]]]])), gu.tt(np.array([[[[[False]]]]])))
    e, b, k = compute_energy(tp, gu.vec_to_gpu(E), gu.vec_to_gpu(B), ps, pd, sc)
    assert np.isclose(float(e), 4.0) and np.isclose(float(b), 1.125) and float(k) == 0.0


@pytest.mark.parametrize("N,tile", CASES[:2])
def test_energy_matches_oracle(N, tile):
    from pypic3d_b200.utils import compute_energy, compute_total_momentum
    sp, dp, tp, sc, E, B = make_case(N, tile, 1)
    ps, pd = gu.to_pkg_params(sp, dp)
    got = compute_energy(gu.particles_to_gpu(tp), gu.vec_to_gpu(E), gu.vec_to_gpu(B), ps, pd, gu.species_to_pkg(sc))
    for a, b in zip(got, odiag.compute_energy(tp, E, B, sp, dp, sc)):
        assert np.isclose(float(a), b, rtol=1e-12)
    assert np.isclose(float(compute_total_momentum(gu.particles_to_gpu(tp), gu.species_to_pkg(sc))), odiag.compute_total_momentum(tp, sc), rtol=1e-12)


# ------------------------------------------------------------------------------------------------ the whole step
STEP_CASES = [
    dict(N=(8, 6, 4), tile=(2, 3, 2), sf=1, dep="esirkepov"),
    dict(N=(8, 6, 4), tile=(4, 3, 4), sf=2, dep="esirkepov"),
    dict(N=(8, 6, 4), tile=(8, 6, 4), sf=2, dep="direct", filt="bilinear"),
    dict(N=(8, 6, 4), tile=(2, 3, 2), sf=1, dep="direct", filt="none", fbc=(1, 0, 0)),
    dict(N=(16, 1, 1), tile=(4, 1, 1), sf=2, dep="direct", filt="bilinear"),
    dict(N=(6, 6, 1), tile=(3, 2, 1), sf=2, dep="esirkepov", fbc=(0, 1, 0), pbc=(0, 1, 0)),
    dict(N=(8, 6, 4), tile=(8, 6, 4), sf=1, dep="esirkepov", pbc=(2, 0, 1), alpha=0.9),
]


def _step_setup(c, capacity=3.0):
    sp, dp, tp, sc, E, B = make_case(c["N"], c["tile"], c["sf"], current_deposition=c["dep"], current_filter=c.get("filt", "none"),
                                     boundary_conditions=c.get("fbc", (0, 0, 0)), particle_boundary_conditions=c.get("pbc", (0, 0, 0)),
                                     alpha=c.get("alpha", 1.0), capacity=capacity, vmax=0.3, dt=0.04)
    return sp, dp, tp, sc, make_fields(sp, dp)


@pytest.mark.parametrize("c", STEP_CASES)
@pytest.mark.parametrize("dtype", (F64, F32))
def test_time_loop_electrodynamic_matches_oracle(c, dtype):
    """Drop-in evolve.time_loop_electrodynamic, 3 steps, slot-exact particles and fields vs the oracle."""
    from pypic3d_b200.evolve import time_loop_electrodynamic
    sp, dp, tp, sc, fields = _step_setup(c)
    ps, pd = gu.to_pkg_params(sp, dp)
    gp, gs, gf = gu.particles_to_gpu(tp, dtype), gu.species_to_pkg(sc), gu.fields_to_gpu(fields, dtype)
    for _ in range(3):
        tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
        gp, gf = time_loop_electrodynamic(gp, gs, gf, ps, pd)
    tol = TOL[dtype] * (10 if dtype == F64 else 5)
    assert np.array_equal(gu.npy(gp.active), tp.active)
    gu.assert_close(gp.x, tp.x, tol, "x"); gu.assert_close(gp.u, tp.u, tol, "u")
    for k in range(3):
        for a, b in zip(gf[k], fields[k]):
            gu.assert_close(a, b, tol, "EBJ"[k])
    assert bool(gf[7].item()) == bool(fields[7])


@pytest.mark.parametrize("c", [s for s in STEP_CASES])
@pytest.mark.parametrize("dtype", (F64, F32))
@pytest.mark.parametrize("sort_interval", (1, 2, 0))
def test_resident_simulation_matches_oracle(c, dtype, sort_interval):
    """The fused/sorted throughput path (K1 + K2 + Yee) on ONE tile: import -> 4 steps -> export == oracle, slot-exact."""
    from pypic3d_b200.simulation import Simulation
    c = dict(c, tile=c["N"])
    sp, dp, tp, sc, fields = _step_setup(c)
    ps, pd = gu.to_pkg_params(sp, dp)
    sim = Simulation(gu.particles_to_gpu(tp, dtype), gu.species_to_pkg(sc), gu.fields_to_gpu(fields, dtype), ps, pd, sort_interval=sort_interval)
    for _ in range(4):
        tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
    sim.step(4)
    gp, gf = sim.export_state()
    tol = TOL[dtype] * (10 if dtype == F64 else 5)
    assert np.array_equal(gu.npy(gp.active), tp.active)
    gu.assert_close(gp.x, tp.x, tol, "x"); gu.assert_close(gp.u, tp.u, tol, "u")
    for k in range(3):
        for a, b in zip(gf[k], fields[k]):
            gu.assert_close(a, b, tol, "EBJ"[k])
    assert sim.overflow() == bool(fields[7])


@pytest.mark.parametrize("c", STEP_CASES[:3])
def test_resident_without_slot_ids(c):
    """track_ids=False (the bench configuration): particles come back compacted in cell order; same set of particles."""
    from pypic3d_b200.simulation import Simulation
    c = dict(c, tile=c["N"])
    sp, dp, tp, sc, fields = _step_setup(c)
    ps, pd = gu.to_pkg_params(sp, dp)
    sim = Simulation(gu.particles_to_gpu(tp), gu.species_to_pkg(sc), gu.fields_to_gpu(fields), ps, pd, sort_interval=2, track_ids=False)
    for _ in range(3):
        tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
    sim.step(3)
    gp, gf = sim.export_state()
    for s in range(tp.x.shape[3]):
        got = gu.sorted_active(gu.npy(gp.x)[:, :, :, s], gu.npy(gp.u)[:, :, :, s], gu.npy(gp.active)[:, :, :, s])
        want = gu.sorted_active(tp.x[:, :, :, s], tp.u[:, :, :, s], tp.active[:, :, :, s])
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-10
    for k in range(3):
        for a, b in zip(gf[k], fields[k]):
            gu.assert_close(a, b, 1e-11, "EBJ"[k])
    # and the exported state can be loaded back
    sim.load_state(gp, (gf[0], gf[1], gf[2]))
    sim.step(1)
    assert not sim.overflow()


@pytest.mark.parametrize("sf", (1, 2))
def test_resident_with_external_fields(sf):
    from pypic3d_b200.simulation import Simulation
    sp, dp, tp, sc, fields = _step_setup(dict(N=(8, 6, 4), tile=(8, 6, 4), sf=sf, dep="esirkepov"))
    ext = make_fields(sp, dp, seed=9, scale=0.2)
    fields = fields[:5] + ((ext[0], ext[1]),) + fields[6:]
    ps, pd = gu.to_pkg_params(sp, dp)
    sim = Simulation(gu.particles_to_gpu(tp), gu.species_to_pkg(sc), gu.fields_to_gpu(fields), ps, pd)
    for _ in range(2):
        tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
    sim.step(2)
    gp, gf = sim.export_state()
    gu.assert_close(gp.u, tp.u, 1e-11, "u")
    for a, b in zip(gf[0], fields[0]):
        gu.assert_close(a, b, 1e-11, "E")


# ------------------------------------------------------------------------------------------------ K1 v9 (supercell tiles)
TILE_CASES = [
    dict(N=(8, 8, 4), pbc=(0, 0, 0), rel=True),
    dict(N=(12, 8, 8), pbc=(2, 0, 1), rel=True),
    dict(N=(4, 8, 8), pbc=(0, 0, 0), rel=False),
]


@pytest.mark.parametrize("c", TILE_CASES)
@pytest.mark.parametrize("dtype", (F64, F32))
@pytest.mark.parametrize("sort_interval", (1, 5, 0))
@pytest.mark.parametrize("variant,jtile", [("pair", "0"), ("pair", "scan"), ("pair", "group"), ("tile", "0"), ("tile", "1"), ("tile", "group")])
def test_resident_tile_kernel_matches_oracle(c, dtype, sort_interval, variant, jtile, monkeypatch):
    """K1 v10 (pic_fused_pair3d) and K1 v9 (pic_fused_tile3d): E/B gathered from shared-memory supercell tiles.  Fast particles
    (up to 0.16 cells per step) and sort_interval 5 / never let particles drift into the tile margin and beyond it, so the tile
    gather, its global-memory fallback and the deferred cell-crossers are all exercised; 12 steps (dt inside the Yee CFL limit),
    slot-exact against the oracle."""
    from pypic3d_b200.simulation import Simulation
    monkeypatch.setenv("PIC_K1_VARIANT", variant)
    # "1": same-cell currents through shared-memory J tiles + TMA reduce (v9, f32 only); "group": match-any group reduction; "0": scan
    monkeypatch.setenv("PIC_K9_JTILE", "1" if jtile == "1" else "0")
    monkeypatch.setenv("PIC_K9_GROUPRED", "1" if jtile == "group" else "0")
    monkeypatch.setenv("PIC_K10_RED", "scan" if jtile == "scan" else "smem")      # (pair, float32: shared-memory reduction unless "scan")
    N = c["N"]
    sp, dp, tp, sc, E, B = make_case(N, N, 1, current_deposition="esirkepov", relativistic=c["rel"],
                                     particle_boundary_conditions=c["pbc"], capacity=3.0, vmax=4.0, C=10.0, dt=0.015, n=200)
    fields = make_fields(sp, dp)
    ps, pd = gu.to_pkg_params(sp, dp)
    sim = Simulation(gu.particles_to_gpu(tp, dtype), gu.species_to_pkg(sc), gu.fields_to_gpu(fields, dtype), ps, pd, sort_interval=sort_interval)
    assert sim.k1_variant == variant
    for _ in range(12):
        tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
    assert np.isfinite(tp.x[tp.active]).all()
    sim.step(12)
    gp, gf = sim.export_state()
    tol = TOL[dtype] * (10 if dtype == F64 else 5)
    assert np.array_equal(gu.npy(gp.active), tp.active)
    gu.assert_close(gp.x, tp.x, tol, "x"); gu.assert_close(gp.u, tp.u, tol, "u")
    for k in range(3):
        for a, b in zip(gf[k], fields[k]):
            gu.assert_close(a, b, tol, "EBJ"[k])
    assert sim.overflow() == bool(fields[7])
    if sort_interval == 0:      # never re-sorted: some particles must have left their tile and taken the global-memory gather
        assert int(sim.flags[2].item()) > 0


@pytest.mark.parametrize("dtype,tol", [(F64, 1e-11), (F32, 2e-4)])
@pytest.mark.parametrize("n,ppc", [(32, 8), (16, 24)])
@pytest.mark.parametrize("tiled,jtile", [("pair", "0"), ("pair", "group"), ("pair", "auto"), ("tile", "0"), ("tile", "1"), ("tile", "group"), ("tile", "auto")])
def test_tile_and_global_k1_variants_agree(dtype, tol, n, ppc, tiled, jtile, monkeypatch):
    """Same thermal plasma, 12 steps with a sort every 5: the supercell-tile K1 and the global-gather K1 differ only by the
    order of the floating-point atomics.  32^3 x 16 ppc is the bench's density (512 particles per supercell and species, inside
    the 640-slot particle stage); 16^3 x 48 ppc puts 1536 particles into every supercell, so most chunks lie beyond the staged
    slice and are read from global memory."""
    from pypic3d_b200.simulation import Simulation
    sp, dp = fx.kernel_parameters(Nx=n, Ny=n, Nz=n, x_wind=float(n), y_wind=float(n), z_wind=float(n), shape_factor=1, dt=0.3,
                                  current_deposition="esirkepov", particle_tile_capacity_factor=1.0)
    tp, sc = fx.thermal_plasma(sp, dp, ppc_per_species=ppc, vth=(0.15, 0.02), seed=5)
    fields = make_fields(sp, dp, scale=0.02)
    ps, pd = gu.to_pkg_params(sp, dp)
    out = {}
    monkeypatch.setenv("PIC_K9_JTILE", "1" if jtile == "1" else "0")
    monkeypatch.setenv("PIC_K9_GROUPRED", {"group": "1", "auto": "auto"}.get(jtile, "0"))
    for variant in (tiled, "global"):
        monkeypatch.setenv("PIC_K1_VARIANT", variant)
        sim = Simulation(gu.particles_to_gpu(tp, dtype), gu.species_to_pkg(sc), gu.fields_to_gpu(fields, dtype), ps, pd, sort_interval=5)
        assert sim.k1_variant == variant
        sim.step(12)
        out[variant] = sim.export_state()
        assert not sim.overflow()
    (pa, fa), (pb, fb) = out[tiled], out["global"]
    assert torch.equal(pa.active, pb.active)
    gu.assert_close(pa.x, gu.npy(pb.x), tol, "x"); gu.assert_close(pa.u, gu.npy(pb.u), tol, "u")
    for k in range(3):
        for a, b in zip(fa[k], fb[k]):
            gu.assert_close(a, gu.npy(b), tol, "EBJ"[k])


# ------------------------------------------------------------------------------------------------ conservation at size
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("dtype,tol", [(F64, 1e-11), (F32, 5e-4)])
def test_charge_conservation_resident_64cubed(sf, dtype, tol):
    """Discrete continuity (rho_new - rho_old)/dt + div J = 0 (esirkepov_test.py:700-744) and Gauss-residual drift on a
    48^3 x 8 ppc thermal plasma run through the resident path; all reductions through CUDA kernels + torch plumbing."""
    from pypic3d_b200.simulation import Simulation
    from pypic3d_b200.deposition.rho import compute_rho
    n = 48
    sp, dp = fx.kernel_parameters(Nx=n, Ny=n, Nz=n, x_wind=float(n), y_wind=float(n), z_wind=float(n), shape_factor=sf, dt=0.3,
                                  current_deposition="esirkepov", particle_tile_capacity_factor=1.0)
    tp, sc = fx.thermal_plasma(sp, dp, ppc_per_species=4, vth=(0.08, 0.01), seed=3)
    fields = make_fields(sp, dp, E=fx.empty_tiled_vector(sp, dp), B=fx.empty_tiled_vector(sp, dp))
    ps, pd = gu.to_pkg_params(sp, dp)
    gp, gs = gu.particles_to_gpu(tp, dtype), gu.species_to_pkg(sc)
    zero = gu.tt(fx.empty_tiled_scalar(sp, dp), dtype)
    sim = Simulation(gp, gs, gu.fields_to_gpu(fields, dtype), ps, pd, sort_interval=2)
    rho_old = compute_rho(gp, gs, zero, ps, pd)
    I = (0, 0, 0, slice(2, -2), slice(2, -2), slice(2, -2))

    def div(F):
        bx = (0, 0, 0, slice(1, -3), slice(2, -2), slice(2, -2)); by = (0, 0, 0, slice(2, -2), slice(1, -3), slice(2, -2)); bz = (0, 0, 0, slice(2, -2), slice(2, -2), slice(1, -3))
        return (F[0][I] - F[0][bx]) / dp.dx + (F[1][I] - F[1][by]) / dp.dy + (F[2][I] - F[2][bz]) / dp.dz
    gauss0 = None
    for step in range(3):
        sim.step(1)
        gp2, gf = sim.export_state()
        rho_new = compute_rho(gp2, gs, zero, ps, pd)
        res = (rho_new[I] - rho_old[I]) / dp.dt + div(gf[2])
        scale = float((rho_new[I] - rho_old[I]).abs().max() / dp.dt) + 1e-30
        assert float(res.abs().max()) <= tol * scale, (step, float(res.abs().max()), scale)
        gauss = div(gf[0]) - rho_new[I] / dp.eps
        if gauss0 is not None:
            assert float((gauss - gauss0).abs().max()) <= tol * float(rho_new[I].abs().max()) * 10
        gauss0 = gauss
        rho_old = rho_new
    assert not sim.overflow()
    assert int(gp2.active.sum()) == int(tp.active.sum())
