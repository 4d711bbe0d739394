"""Oracle pins for the electrostatic field solve, from the reference's own physics test
tests/physics_tests/electrostatic_yee_test.py: the single-mode Poisson solve (:92-119) and the centred-gradient contract
(:121-153); plus conducting-wall and step-level sanity for the restatement itself."""
import numpy as np
import pytest

from oracle import fixtures as fx, electrostatic as oes, evolve as oevolve


def _setup(n=16, bcs=(0, 0, 0), alpha=1.0):
    w = 2 * np.pi
    sp, dp = fx.kernel_parameters(Nx=n, Ny=n, Nz=n, x_wind=w, y_wind=w, z_wind=w, tile_shape=(n, n, n), guard_cells=2,
                                  shape_factor=1, boundary_conditions=bcs, eps=1.0, alpha=alpha, electrostatic=True,
                                  solver="electrostatic")
    x = np.linspace(0, w, n, endpoint=False)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    return sp, dp, X, Y, Z


def _neg_lapl_periodic(f, dx, dy, dz):
    lx = (np.roll(f, 1, 0) + np.roll(f, -1, 0) - 2.0 * f) / (dx * dx)
    ly = (np.roll(f, 1, 1) + np.roll(f, -1, 1) - 2.0 * f) / (dy * dy)
    lz = (np.roll(f, 1, 2) + np.roll(f, -1, 2) - 2.0 * f) / (dz * dz)
    return -(lx + ly + lz)


def test_cg_single_mode_reference_pin():
    """electrostatic_yee_test.py:92-119 with the reference's literal tolerances."""
    sp, dp, X, Y, Z = _setup()
    g = 2
    A = slice(g, -g)
    phi_true = np.sin(X + Y + Z)
    rho_int = _neg_lapl_periodic(phi_true, dp.dx, dp.dy, dp.dz) * dp.eps
    rho = np.zeros((20, 20, 20)); rho[A, A, A] = rho_int
    phi = oes.solve_poisson_with_conjugate_gradient(rho, np.zeros_like(rho), sp, dp, tol=1e-10, max_iter=4000)
    num = phi[A, A, A] - phi[A, A, A].mean()
    assert np.allclose(num, phi_true - phi_true.mean(), atol=1e-7, rtol=1e-6)
    assert np.abs(_neg_lapl_periodic(num, dp.dx, dp.dy, dp.dz) - rho_int / dp.eps).max() < 1e-6
    # ghosts are the periodic images of the interior
    assert np.array_equal(phi[:g, A, A], phi[-2 * g:-g, A, A]) and np.array_equal(phi[A, A, -g:], phi[A, A, g:2 * g])


def test_gradient_matches_centered_difference_of_solved_phi():
    """electrostatic_yee_test.py:121-153: E = -grad(phi) of the CG solution, centred differences, periodic."""
    sp, dp, X, Y, Z = _setup()
    g = 2
    A = slice(g, -g)
    tp, sc = fx.build_tiled_particles([fx.particle_species("test", 1.0, 1.0, x1=np.array([0.1]), x2=np.array([0.2]), x3=np.array([0.3]),
                                                           u1=np.zeros(1), u2=np.zeros(1), u3=np.zeros(1))], sp, dp)
    E, phi, rho = oes.calculate_tiled_electrostatic_fields(sp, dp, tp, sc, fx.empty_tiled_scalar(sp, dp), fx.empty_tiled_scalar(sp, dp))
    want_phi = oes.solve_poisson_with_conjugate_gradient(rho[0, 0, 0], np.zeros_like(rho[0, 0, 0]), sp, dp)
    pin = -want_phi[A, A, A]
    for c, (axis, d) in enumerate(((0, dp.dx), (1, dp.dy), (2, dp.dz))):
        want = (np.roll(pin, -1, axis) - np.roll(pin, 1, axis)) / (2.0 * d)
        assert np.allclose(np.asarray(E[c])[0, 0, 0][A, A, A], want, rtol=1e-10, atol=1e-12)
    assert abs(rho[0, 0, 0][A, A, A].sum() * dp.dx * dp.dy * dp.dz - 1.0) < 1e-12      # one unit charge deposited


@pytest.mark.parametrize("bcs", [(1, 0, 0), (0, 1, 1)])
def test_cg_conducting_walls_residual_and_constant_ghosts(bcs):
    sp, dp, X, Y, Z = _setup(n=12, bcs=bcs)
    g = 2
    A = slice(g, -g)
    rng = np.random.default_rng(2)
    rho = np.zeros((16, 16, 16)); rho[A, A, A] = rng.normal(size=(12, 12, 12))
    rho[A, A, A] -= rho[A, A, A].mean()     # periodic / constant-potential walls: the operator has the constant null space
    phi, iters = oes.solve_poisson_with_conjugate_gradient(rho, np.zeros_like(rho), sp, dp, return_iterations=True)
    assert 0 < iters < 5000
    F, Bk = slice(g + 1, -g + 1), slice(g - 1, -g - 1)
    lap = ((phi[F, A, A] + phi[Bk, A, A] - 2 * phi[A, A, A]) / dp.dx ** 2 + (phi[A, F, A] + phi[A, Bk, A] - 2 * phi[A, A, A]) / dp.dy ** 2
           + (phi[A, A, F] + phi[A, A, Bk] - 2 * phi[A, A, A]) / dp.dz ** 2)
    assert np.abs(lap + rho[A, A, A] / dp.eps).max() < 1e-9
    # Reference quirk kept by the restatement: every conducting axis first re-runs the full ghost refresh (which zeroes the
    # exterior ghosts of all non-periodic axes, electrostatic_yee.py:24-33 + ghost_cells.py:686), so only the LAST conducting
    # axis ends up with constant-potential ghosts; earlier conducting axes keep zero ghosts.
    cond = [a for a in range(3) if bcs[a] == 1]
    for axis in cond:
        lo = [A, A, A]; lo[axis] = slice(0, g)
        first = [A, A, A]; first[axis] = slice(g, g + 1)
        if axis == cond[-1]:
            assert np.array_equal(phi[tuple(lo)], np.broadcast_to(phi[tuple(first)], phi[tuple(lo)].shape))
        else:
            assert not phi[tuple(lo)].any()


def test_electrostatic_step_runs_and_conserves_particles():
    n = 8
    sp, dp = fx.kernel_parameters(Nx=n, Ny=n, Nz=n, x_wind=float(n), y_wind=float(n), z_wind=float(n), tile_shape=(n, n, n), dt=0.1,
                                  shape_factor=2, electrostatic=True, solver="electrostatic", alpha=0.9)
    tp, sc = fx.thermal_plasma(sp, dp, ppc_per_species=2, vth=(0.1, 0.01), seed=4)
    z = fx.empty_tiled_vector
    fields = (z(sp, dp), z(sp, dp), z(sp, dp), fx.empty_tiled_scalar(sp, dp), fx.empty_tiled_scalar(sp, dp), (z(sp, dp), z(sp, dp)), None, False)
    n0 = int(tp.active.sum())
    for _ in range(2):
        tp, fields = oevolve.time_loop_electrostatic(tp, sc, fields, sp, dp)
    assert int(tp.active.sum()) == n0 and not fields[7]
    assert np.isfinite(np.asarray(fields[0][0])).all() and np.abs(np.asarray(fields[0][0])).max() > 0
