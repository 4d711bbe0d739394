"""-m gpu parity for the electrostatic field solve: CUDA CG / wall treatment / gradient / full
electrostatic step against the oracle restatement of PyPIC3D/solvers/electrostatic_yee.py and evolve.py:106-161."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx, electrostatic as oes, evolve as oevolve
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu
F64, F32 = torch.float64, torch.float32


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    gu.require_cuda()


def _params(n, bcs=(0, 0, 0), alpha=1.0, sf=1, dt=0.1, wind=None):
    w = float(n) if wind is None else wind
    return fx.kernel_parameters(Nx=n[0] if isinstance(n, tuple) else n, Ny=n[1] if isinstance(n, tuple) else n,
                                Nz=n[2] if isinstance(n, tuple) else n, x_wind=w, y_wind=w, z_wind=w, guard_cells=2, shape_factor=sf,
                                boundary_conditions=bcs, eps=1.0, alpha=alpha, electrostatic=True, solver="electrostatic", dt=dt)


def test_cg_single_mode_reference_pin_on_gpu():
    """electrostatic_yee_test.py:92-119 through the CUDA solver, the reference's literal tolerances."""
    from pypic3d_b200.solvers.electrostatic_yee import solve_poisson_with_conjugate_gradient
    n, w, g = 16, 2 * np.pi, 2
    sp, dp = _params(n, wind=w)
    ps, pd = gu.to_pkg_params(sp, dp)
    x = np.linspace(0, w, n, endpoint=False)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    phi_true = np.sin(X + Y + Z)
    lap = sum((np.roll(phi_true, 1, a) + np.roll(phi_true, -1, a) - 2 * phi_true) / dp.dx ** 2 for a in range(3))
    A = slice(g, -g)
    rho = np.zeros((n + 4,) * 3); rho[A, A, A] = -lap * dp.eps
    phi, iters = solve_poisson_with_conjugate_gradient(gu.tt(rho), gu.tt(np.zeros_like(rho)), ps, pd, tol=1e-10, max_iter=4000,
                                                       return_iterations=True)
    num = gu.npy(phi)[A, A, A]
    num = num - num.mean()
    assert np.allclose(num, phi_true - phi_true.mean(), atol=1e-7, rtol=1e-6)
    _, it_ref = oes.solve_poisson_with_conjugate_gradient(rho, np.zeros_like(rho), sp, dp, tol=1e-10, max_iter=4000, return_iterations=True)
    assert abs(iters - it_ref) <= 1


@pytest.mark.parametrize("n,bcs", [((12, 10, 8), (0, 0, 0)), ((12, 10, 8), (1, 0, 0)), ((8, 8, 8), (0, 1, 1)), ((16, 16, 16), (1, 1, 1))])
@pytest.mark.parametrize("check_every", (1, 16))
def test_cg_matches_oracle(n, bcs, check_every):
    from pypic3d_b200.solvers.electrostatic_yee import solve_poisson_with_conjugate_gradient
    sp, dp = _params(n, bcs=bcs, wind=5.0)
    ps, pd = gu.to_pkg_params(sp, dp)
    g = 2
    A = slice(g, -g)
    rng = np.random.default_rng(7)
    rho = np.zeros(tuple(v + 4 for v in n)); rho[A, A, A] = rng.normal(size=n)
    rho[A, A, A] -= rho[A, A, A].mean()
    phi0 = np.zeros_like(rho); phi0[A, A, A] = 0.1 * rng.normal(size=n)
    want, it_ref = oes.solve_poisson_with_conjugate_gradient(rho, phi0, sp, dp, return_iterations=True)
    got, iters = solve_poisson_with_conjugate_gradient(gu.tt(rho), gu.tt(phi0), ps, pd, return_iterations=True, check_every=check_every)
    assert abs(iters - it_ref) <= 2, (iters, it_ref)
    scale = np.abs(want).max()
    assert np.abs(gu.npy(got) - want).max() <= 1e-9 * scale


@pytest.mark.parametrize("bcs", [(0, 0, 0), (1, 0, 1)])
def test_phi_boundaries_and_gradient_match_oracle(bcs):
    from pypic3d_b200.solvers.electrostatic_yee import apply_tiled_phi_constant_boundaries, _centered_tiled_electrostatic_gradient
    n = (8, 6, 4)
    sp, dp = _params(n, bcs=bcs, wind=3.0)
    ps, pd = gu.to_pkg_params(sp, dp)
    rng = np.random.default_rng(3)
    phi = rng.normal(size=(1, 1, 1, 12, 10, 8))
    want = np.asarray(oes.apply_tiled_phi_constant_boundaries(phi, sp, 2))
    gu.assert_close(apply_tiled_phi_constant_boundaries(gu.tt(phi), ps, pd, 2), want, 0.0, "phi bc")
    E = _centered_tiled_electrostatic_gradient(gu.tt(phi), ps, pd, 2)
    for a, b in zip(E, oes.centered_tiled_electrostatic_gradient(phi, sp, dp, 2)):
        gu.assert_close(a, np.asarray(b), 1e-13, "E")


@pytest.mark.parametrize("sf,alpha,bcs,pbc", [(1, 1.0, (0, 0, 0), (0, 0, 0)), (2, 0.9, (0, 0, 0), (0, 0, 0)), (2, 1.0, (0, 0, 0), (0, 0, 0))])
def test_time_loop_electrostatic_matches_oracle(sf, alpha, bcs, pbc):
    """Drop-in evolve.time_loop_electrostatic, 3 steps, slot-exact particles and fields against the oracle (f64).
    Periodic only: with a conducting wall the reference's rho deposit leaves a net interior charge (the wall node's share is
    not folded back), the constant-potential Poisson problem is then inconsistent and its CG diverges (phi ~ 1e16 after the
    5000 iterations) -- there is nothing meaningful to compare; the wall treatment itself is covered on consistent right-hand
    sides by test_cg_matches_oracle and test_phi_boundaries_and_gradient_match_oracle."""
    from pypic3d_b200.evolve import time_loop_electrostatic
    n = 8
    sp, dp = fx.kernel_parameters(Nx=n, Ny=n, Nz=n, x_wind=float(n), y_wind=float(n), z_wind=float(n), tile_shape=(n, n, n), dt=0.1,
                                  shape_factor=sf, electrostatic=True, solver="electrostatic", alpha=alpha, boundary_conditions=bcs,
                                  particle_boundary_conditions=pbc)
    tp, sc = fx.thermal_plasma(sp, dp, ppc_per_species=2, vth=(0.1, 0.01), seed=4)
    z = fx.empty_tiled_vector
    fields = (z(sp, dp), z(sp, dp), z(sp, dp), fx.empty_tiled_scalar(sp, dp), fx.empty_tiled_scalar(sp, dp), (z(sp, dp), z(sp, dp)), None, False)
    ps, pd = gu.to_pkg_params(sp, dp)
    gp, gs, gf = gu.particles_to_gpu(tp), gu.species_to_pkg(sc), gu.fields_to_gpu(fields)
    for _ in range(3):
        tp, fields = oevolve.time_loop_electrostatic(tp, sc, fields, sp, dp)
        gp, gf = time_loop_electrostatic(gp, gs, gf, ps, pd)
    assert np.array_equal(gu.npy(gp.active), tp.active)
    gu.assert_close(gp.x, tp.x, 1e-9, "x"); gu.assert_close(gp.u, tp.u, 1e-9, "u")
    for a, b in zip(gf[0], fields[0]):
        gu.assert_close(a, np.asarray(b), 1e-8, "E")
    gu.assert_close(gf[3], np.asarray(fields[3]), 1e-11, "rho")
    gu.assert_close(gf[4], np.asarray(fields[4]), 1e-8, "phi")
    assert bool(gf[7].item()) == bool(fields[7])


def test_cg_float32_converges_to_float_accuracy():
    from pypic3d_b200.solvers.electrostatic_yee import solve_poisson_with_conjugate_gradient
    n = (16, 16, 16)
    sp, dp = _params(n, wind=4.0)
    ps, pd = gu.to_pkg_params(sp, dp)
    A = slice(2, -2)
    rng = np.random.default_rng(5)
    rho = np.zeros((20, 20, 20)); rho[A, A, A] = rng.normal(size=n); rho[A, A, A] -= rho[A, A, A].mean()
    want = oes.solve_poisson_with_conjugate_gradient(rho, np.zeros_like(rho), sp, dp)
    got = solve_poisson_with_conjugate_gradient(gu.tt(rho, F32), gu.tt(np.zeros_like(rho), F32), ps, pd, tol=1e-5, max_iter=500)
    w = want[A, A, A] - want[A, A, A].mean()
    gt = gu.npy(got)[A, A, A]; gt = gt - gt.mean()
    assert np.abs(gt - w).max() <= 2e-4 * np.abs(w).max()
