"""CPU check of the CUDA kernels' per-particle arithmetic (no GPU needed).

The __host__ __device__ bodies in pypic3d_b200/csrc/pic_slots.cuh (the exact code the __global__ kernels execute)
are compiled for the host by g++ into a TEST-ONLY library and compared with the oracle.  This is not a product
path: the package never loads this library."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import fixtures as fx, pusher, deposition as dep, particles as opart
from pypic3d_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
OUT = os.path.join(HERE, "hostcheck", "_build", "libhostcheck.so")


@pytest.fixture(scope="module")
def hc():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", OUT], check=True)
    L = ctypes.CDLL(OUT)
    assert L.hc_params_size() == ctypes.sizeof(_lib.PicParams)
    return L


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _v3(arrs):
    return (ctypes.c_void_p * 3)(*[a.ctypes.data for a in arrs])


from tests.cases import CASES


def _setup(N, tile, sf, pusher_name="boris", rel=True, dtype=np.float64, **kw):
    sp, dp = fx.kernel_parameters(Nx=N[0], Ny=N[1], Nz=N[2], x_wind=4.0 if N[0] > 1 else 1.0, y_wind=3.0 if N[1] > 1 else 1.0,
                                  z_wind=2.0 if N[2] > 1 else 1.0, shape_factor=sf, dt=0.05, tile_shape=tile,
                                  particle_tile_capacity_factor=2.0, particle_pusher=pusher_name, relativistic=rel, **kw)
    rng = np.random.default_rng(11)
    n = 60
    species = []
    for s, (q, m) in enumerate(((-1.0, 1.0), (2.0, 5.0))):
        pos = [rng.uniform(-w / 2, w / 2, n) if NN > 1 else np.zeros(n) for w, NN in ((dp.x_wind, dp.Nx), (dp.y_wind, dp.Ny), (dp.z_wind, dp.Nz))]
        vel = [rng.uniform(-0.4, 0.4, n) for _ in range(3)]
        species.append(fx.particle_species(f"s{s}", q, m, weight=0.5 + s, x1=pos[0], x2=pos[1], x3=pos[2], u1=vel[0], u2=vel[1], u3=vel[2],
                                           update_u=(True, s == 0, True)))
    tp, sc = fx.build_tiled_particles(species, sp, dp)
    Eg = tuple(rng.normal(size=(N[0] + 2, N[1] + 2, N[2] + 2)) for _ in range(3))
    Bg = tuple(rng.normal(size=(N[0] + 2, N[1] + 2, N[2] + 2)) for _ in range(3))
    E, B = fx.vector_tiles_from_global(Eg, sp, dp), fx.vector_tiles_from_global(Bg, sp, dp)
    return sp, dp, tp, sc, E, B


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("pn,rel", [("boris", True), ("boris", False), ("higuera_cary", True)])
def test_push_body_matches_oracle(hc, N, tile, sf, pn, rel):
    sp, dp, tp, sc, E, B = _setup(N, tile, sf, pn, rel)
    ref = pusher.particle_push(tp, sc, E, B, sp, dp)
    p = _lib.make_params(sp, dp, sc, np.float64)
    x = np.ascontiguousarray(tp.x); u = np.ascontiguousarray(tp.u); a = np.ascontiguousarray(tp.active.astype(np.uint8))
    out = np.zeros_like(u)
    Ec = [np.ascontiguousarray(c) for c in E]; Bc = [np.ascontiguousarray(c) for c in B]
    hc.hc_push(ctypes.byref(p), _ptr(x), _ptr(u), _ptr(out), _ptr(a), ctypes.c_int64(tp.x.shape[4]), _v3(Ec), _v3(Bc))
    assert np.allclose(out, ref.u, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("mode", (0, 1, 2))
def test_deposit_body_matches_oracle(hc, N, tile, sf, mode):
    sp, dp, tp, sc, E, B = _setup(N, tile, sf)
    z = fx.empty_tiled_vector(sp, dp)
    if mode == 0:
        ref = dep.Esirkepov_current(tp, sc, z, sp, dp, fold=False)
    elif mode == 1:
        ref = dep.J_from_rhov(tp, sc, z, sp, dp, fold=False)
    else:
        r = dep.compute_rho(tp, sc, z[0], sp, dp, fold=False)
        ref = (r, np.zeros_like(r), np.zeros_like(r))
    p = _lib.make_params(sp, dp, sc, np.float64)
    x = np.ascontiguousarray(tp.x); u = np.ascontiguousarray(tp.u); a = np.ascontiguousarray(tp.active.astype(np.uint8))
    J = [np.zeros_like(z[0]) for _ in range(3)]
    hc.hc_deposit(ctypes.byref(p), mode, _ptr(x), _ptr(u), _ptr(a), ctypes.c_int64(tp.x.shape[4]), _v3(J))
    scale = max(1.0, max(np.abs(r).max() for r in ref))
    for c in range(3 if mode < 2 else 1):
        assert np.allclose(J[c], ref[c], rtol=1e-12, atol=1e-12 * scale), (mode, c)


@pytest.mark.parametrize("pbc", [(0, 0, 0), (1, 0, 2), (2, 1, 0)])
def test_retile_classify_body_matches_oracle(hc, pbc):
    sp, dp, tp, sc, E, B = _setup((8, 6, 4), (2, 3, 2), 1, particle_boundary_conditions=pbc)
    moved = opart.update_tiled_particle_positions(tp, sc, 1.0)   # dt=1: up to 0.4 -> crosses tiles and walls
    ref, ovf = opart.refresh_tiled_particle_tiles(moved, sp, dp)
    p = _lib.make_params(sp, dp, sc, np.float64)
    x = np.ascontiguousarray(moved.x); u = np.ascontiguousarray(moved.u); a = np.ascontiguousarray(moved.active.astype(np.uint8))
    xo, uo, ao = np.zeros_like(x), np.zeros_like(u), np.zeros_like(a)
    code = np.zeros(a.size, dtype=np.int32); flag = np.zeros(1, dtype=np.int32)
    hc.hc_retile_classify(ctypes.byref(p), _ptr(x), _ptr(u), _ptr(a), _ptr(xo), _ptr(uo), _ptr(ao), ctypes.c_int64(x.shape[4]), _ptr(code), _ptr(flag))
    # staying particles are identical; movers+stayers == oracle's active count (unless overflow)
    stay = ao.astype(bool)
    assert np.allclose(xo[stay], ref.x[stay], rtol=1e-13, atol=1e-13) and np.all(ref.active[stay])
    if not ovf:
        assert (code >= 1).sum() == ref.active.sum()


@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("depmode", (0, 1))
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
@pytest.mark.parametrize("N", [(8, 6, 4), (8, 1, 1), (6, 6, 1)])
def test_fused_body_matches_oracle_step(hc, sf, depmode, dtype, tol, N):
    """K1 body on a single tile == oracle push -> deposit -> move -> retile (evolve.py:33-79)."""
    sp, dp, tp, sc, E, B = _setup(N, N, sf, current_deposition="esirkepov" if depmode == 0 else "direct")
    pushed = pusher.particle_push(tp, sc, E, B, sp, dp)
    z = fx.empty_tiled_vector(sp, dp)
    if depmode == 0:
        Jref = dep.Esirkepov_current(pushed, sc, z, sp, dp, fold=False)
        moved, _ = opart.refresh_tiled_particle_tiles(opart.update_tiled_particle_positions(pushed, sc, dp.dt), sp, dp)
    else:
        half, _ = opart.refresh_tiled_particle_tiles(opart.update_tiled_particle_positions(pushed, sc, dp.dt / 2), sp, dp)
        Jref = dep.J_from_rhov(half, sc, z, sp, dp, fold=False)
        moved, _ = opart.refresh_tiled_particle_tiles(opart.update_tiled_particle_positions(half, sc, dp.dt / 2), sp, dp)
    p = _lib.make_params(sp, dp, sc, dtype)
    Ec = [np.ascontiguousarray(c[0, 0, 0], dtype=dtype) for c in E]; Bc = [np.ascontiguousarray(c[0, 0, 0], dtype=dtype) for c in B]
    J = [np.zeros_like(Ec[0]) for _ in range(3)]
    flags = np.zeros(1, dtype=np.int32)
    for s in range(2):
        act = tp.active[0, 0, 0, s]
        comp = [np.ascontiguousarray(tp.x[0, 0, 0, s][act][:, c], dtype=dtype) for c in range(3)] + \
               [np.ascontiguousarray(tp.u[0, 0, 0, s][act][:, c], dtype=dtype) for c in range(3)]
        cp = (ctypes.c_void_p * 6)(*[a.ctypes.data for a in comp])
        hc.hc_fused(ctypes.byref(p), s, depmode, cp, ctypes.c_int64(int(act.sum())), _v3(Ec), _v3(Bc), _v3(J), None, ctypes.c_int64(0), None, _ptr(flags))
        xr = moved.x[0, 0, 0, s][act]; ur = moved.u[0, 0, 0, s][act]
        for c in range(3):
            assert np.allclose(comp[c], xr[:, c], rtol=tol, atol=tol * 4), ("x", s, c)
            assert np.allclose(comp[3 + c], ur[:, c], rtol=tol, atol=tol), ("u", s, c)
    scale = max(np.abs(r).max() for r in Jref)
    for c in range(3):
        assert np.allclose(J[c], Jref[c][0, 0, 0], rtol=tol, atol=tol * scale), ("J", c)
    assert flags[0] == 0


@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("rel", (True, False))
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 3e-5)])
@pytest.mark.parametrize("pbc", [(0, 0, 0), (1, 0, 2)])
@pytest.mark.parametrize("vmax", (0.4, 3.0))
def test_fused3d_fast_body_matches_oracle_step(hc, sf, rel, dtype, tol, pbc, vmax):
    """K1 v2 (specialised 3-D Esirkepov body) == oracle push -> Esirkepov -> move -> retile on one tile.
    vmax=3.0 (|v| dt up to 0.15 = 0.3 dx) forces frequent cell crossings: every branch of the union stencil is exercised."""
    from tests.cases import make_case
    N = (8, 6, 4)
    sp, dp, tp, sc, E, B = make_case(N, N, sf, current_deposition="esirkepov", relativistic=rel, particle_boundary_conditions=pbc,
                                     vmax=vmax, C=10.0)
    pushed = pusher.particle_push(tp, sc, E, B, sp, dp)
    z = fx.empty_tiled_vector(sp, dp)
    Jref = dep.Esirkepov_current(pushed, sc, z, sp, dp, fold=False)
    moved, _ = opart.refresh_tiled_particle_tiles(opart.update_tiled_particle_positions(pushed, sc, dp.dt), sp, dp)
    p = _lib.make_params(sp, dp, sc, dtype)
    Ec = [np.ascontiguousarray(c[0, 0, 0], dtype=dtype) for c in E]; Bc = [np.ascontiguousarray(c[0, 0, 0], dtype=dtype) for c in B]
    J = [np.zeros_like(Ec[0]) for _ in range(3)]
    flags = np.zeros(1, dtype=np.int32)
    for s in range(2):
        act = tp.active[0, 0, 0, s]
        comp = [np.ascontiguousarray(tp.x[0, 0, 0, s][act][:, c], dtype=dtype) for c in range(3)] + \
               [np.ascontiguousarray(tp.u[0, 0, 0, s][act][:, c], dtype=dtype) for c in range(3)]
        cp = (ctypes.c_void_p * 6)(*[a.ctypes.data for a in comp])
        hc.hc_fused3d(ctypes.byref(p), s, cp, ctypes.c_int64(int(act.sum())), _v3(Ec), _v3(Bc), _v3(J), None, ctypes.c_int64(0), None, _ptr(flags))
        alive = moved.active[0, 0, 0, s][act]
        xr = moved.x[0, 0, 0, s][act]; ur = moved.u[0, 0, 0, s][act]
        assert np.array_equal(~np.isnan(comp[0]), alive)
        for c in range(3):
            assert np.allclose(comp[c][alive], xr[alive][:, c], rtol=tol, atol=tol * 4), ("x", s, c)
            assert np.allclose(comp[3 + c][alive], ur[alive][:, c], rtol=tol, atol=tol * 10), ("u", s, c)
    scale = max(np.abs(r).max() for r in Jref)
    for c in range(3):
        assert np.allclose(J[c], Jref[c][0, 0, 0], rtol=tol, atol=tol * scale), ("J", c)
    assert flags[0] == 0


@pytest.mark.parametrize("rel", (True, False))
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 3e-5)])
@pytest.mark.parametrize("pbc", [(0, 0, 0), (1, 0, 2)])
@pytest.mark.parametrize("shift", (0, 1, -1))
def test_tile3d_body_matches_oracle_step(hc, rel, dtype, tol, pbc, shift):
    """K1 v9 body (E/B gathered from a supercell tile, global-memory fallback outside it) == oracle step on one tile.
    shift != 0 hands every particle the tile of a neighbouring supercell along one axis, so particles land in the tile's
    one-cell margin or outside it: the tile gather and the fallback must give the same numbers."""
    from tests.cases import make_case
    N = (8, 8, 4)
    sp, dp, tp, sc, E, B = make_case(N, N, 1, current_deposition="esirkepov", relativistic=rel, particle_boundary_conditions=pbc,
                                     vmax=3.0, C=10.0, n=120)
    assert int(sp.guard_cells) == 2
    pushed = pusher.particle_push(tp, sc, E, B, sp, dp)
    z = fx.empty_tiled_vector(sp, dp)
    Jref = dep.Esirkepov_current(pushed, sc, z, sp, dp, fold=False)
    moved, _ = opart.refresh_tiled_particle_tiles(opart.update_tiled_particle_positions(pushed, sc, dp.dt), sp, dp)
    p = _lib.make_params(sp, dp, sc, dtype)
    Ec = [np.ascontiguousarray(c[0, 0, 0], dtype=dtype) for c in E]; Bc = [np.ascontiguousarray(c[0, 0, 0], dtype=dtype) for c in B]
    J = [np.zeros_like(Ec[0]) for _ in range(3)]
    flags = np.zeros(4, dtype=np.int32)
    for s in range(2):
        act = tp.active[0, 0, 0, s]
        comp = [np.ascontiguousarray(tp.x[0, 0, 0, s][act][:, c], dtype=dtype) for c in range(3)] + \
               [np.ascontiguousarray(tp.u[0, 0, 0, s][act][:, c], dtype=dtype) for c in range(3)]
        cp = (ctypes.c_void_p * 6)(*[a.ctypes.data for a in comp])
        hc.hc_tile3d(ctypes.byref(p), s, cp, ctypes.c_int64(int(act.sum())), _v3(Ec), _v3(Bc), _v3(J), ctypes.c_int(shift), _ptr(flags))
        alive = moved.active[0, 0, 0, s][act]
        xr = moved.x[0, 0, 0, s][act]; ur = moved.u[0, 0, 0, s][act]
        assert np.array_equal(~np.isnan(comp[0]), alive)
        for c in range(3):
            assert np.allclose(comp[c][alive], xr[alive][:, c], rtol=tol, atol=tol * 4), ("x", s, c)
            assert np.allclose(comp[3 + c][alive], ur[alive][:, c], rtol=tol, atol=tol * 10), ("u", s, c)
    scale = max(np.abs(r).max() for r in Jref)
    for c in range(3):
        assert np.allclose(J[c], Jref[c][0, 0, 0], rtol=tol, atol=tol * scale), ("J", c)
    assert flags[0] == 0


@pytest.mark.parametrize("rel", (True, False))
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 3e-5)])
@pytest.mark.parametrize("pbc", [(0, 0, 0), (1, 0, 2)])
@pytest.mark.parametrize("shift", (0, 1, -1))
@pytest.mark.parametrize("W", (1, 2))
@pytest.mark.parametrize("vmax", (0.4, 3.0))
def test_pair3d_body_matches_oracle_step(hc, rel, dtype, tol, pbc, shift, W, vmax):
    """K1 v10 body (pic_pair.cuh pair_advance + crosser_finish + the scalar fallback, finished as kernels_pair.cu does) == oracle
    push -> Esirkepov -> move -> retile on one tile.  W = 2 hands the body two particles at a time that need not share a cell or
    even a supercell; shift != 0 gives every group the tile of a neighbouring supercell, so particles land in the margin or
    outside it (scalar global-memory body); vmax = 3 makes a third of the particles change cell; the (1, 0, 2) case puts
    reflecting and absorbing walls on the rim supercells (the `edge` path)."""
    from tests.cases import make_case
    N = (8, 8, 4)
    sp, dp, tp, sc, E, B = make_case(N, N, 1, current_deposition="esirkepov", relativistic=rel, particle_boundary_conditions=pbc,
                                     vmax=vmax, C=10.0, n=120)
    assert int(sp.guard_cells) == 2
    pushed = pusher.particle_push(tp, sc, E, B, sp, dp)
    z = fx.empty_tiled_vector(sp, dp)
    Jref = dep.Esirkepov_current(pushed, sc, z, sp, dp, fold=False)
    moved, _ = opart.refresh_tiled_particle_tiles(opart.update_tiled_particle_positions(pushed, sc, dp.dt), sp, dp)
    p = _lib.make_params(sp, dp, sc, dtype)
    Ec = [np.ascontiguousarray(c[0, 0, 0], dtype=dtype) for c in E]; Bc = [np.ascontiguousarray(c[0, 0, 0], dtype=dtype) for c in B]
    J = [np.zeros_like(Ec[0]) for _ in range(3)]
    flags = np.zeros(4, dtype=np.int32)
    for s in range(2):
        act = tp.active[0, 0, 0, s]
        comp = [np.ascontiguousarray(tp.x[0, 0, 0, s][act][:, c], dtype=dtype) for c in range(3)] + \
               [np.ascontiguousarray(tp.u[0, 0, 0, s][act][:, c], dtype=dtype) for c in range(3)]
        cp = (ctypes.c_void_p * 6)(*[a.ctypes.data for a in comp])
        hc.hc_pair3d(ctypes.byref(p), s, cp, ctypes.c_int64(int(act.sum())), _v3(Ec), _v3(Bc), _v3(J), ctypes.c_int(W), ctypes.c_int(shift), _ptr(flags))
        alive = moved.active[0, 0, 0, s][act]
        xr = moved.x[0, 0, 0, s][act]; ur = moved.u[0, 0, 0, s][act]
        assert np.array_equal(~np.isnan(comp[0]), alive)
        for c in range(3):
            assert np.allclose(comp[c][alive], xr[alive][:, c], rtol=tol, atol=tol * 4), ("x", s, c)
            assert np.allclose(comp[3 + c][alive], ur[alive][:, c], rtol=tol, atol=tol * 10), ("u", s, c)
    scale = max(np.abs(r).max() for r in Jref)
    for c in range(3):
        assert np.allclose(J[c], Jref[c][0, 0, 0], rtol=tol, atol=tol * scale), ("J", c)
    assert flags[0] == 0
    if shift == 0 and W == 1:
        assert flags[2] == 0        # every particle is covered by the tile of its own supercell
    if shift != 0:
        assert flags[2] > 0         # some particles are not: they took the scalar global-memory body


# ---- the kernel's push body (gather + pusher, as the CUDA kernels execute it) through the reference's pusher physics tests -----
def _kernel_push(hc, pusher_name, v, E, B, dt, steps, E_of_step=None):
    """One particle at the origin of a 1x1x1 domain with uniform fields, q = m = C = 1, pushed `steps` times by slot_push."""
    sp, dp = fx.kernel_parameters(Nx=1, Ny=1, Nz=1, x_wind=1.0, y_wind=1.0, z_wind=1.0, C=1.0, dt=dt, particle_pusher=pusher_name,
                                  relativistic=True, shape_factor=1)
    tp, sc = fx.build_tiled_particles([fx.particle_species("p", 1.0, 1.0, x1=np.zeros(1), x2=np.zeros(1), x3=np.zeros(1),
                                                            u1=np.array([v[0]]), u2=np.array([v[1]]), u3=np.array([v[2]]))], sp, dp)
    p = _lib.make_params(sp, dp, sc, np.float64)
    x = np.ascontiguousarray(tp.x); u = np.ascontiguousarray(tp.u); a = np.ascontiguousarray(tp.active.astype(np.uint8))
    shape = fx.empty_tiled_scalar(sp, dp).shape
    out = np.zeros_like(u)
    for step in range(steps):
        Es = E if E_of_step is None else E_of_step(step)
        Ec = [np.full(shape, c, dtype=np.float64) for c in Es]; Bc = [np.full(shape, c, dtype=np.float64) for c in B]
        hc.hc_push(ctypes.byref(p), _ptr(x), _ptr(u), _ptr(out), _ptr(a), ctypes.c_int64(tp.x.shape[4]), _v3(Ec), _v3(Bc))
        u, out = out, u
    return tuple(float(c) for c in u.reshape(-1, 3)[0])


def _g(v):
    return 1.0 / np.sqrt(1.0 - sum(c * c for c in v))


@pytest.mark.parametrize("pn", ("boris", "higuera_cary"))
def test_kernel_push_body_passes_schmitz_constant_B(hc, pn):
    """pusher_schmitz_test.py:69-84 through the CUDA kernels' own push body (host build)."""
    for gamma0 in (1.001, 10.0):
        u0 = np.sqrt(gamma0 ** 2 - 1.0)
        v = _kernel_push(hc, pn, (u0 / gamma0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 0.2 * np.pi * gamma0, 10)
        assert abs(_g(v) - gamma0) <= 2.0e-12 * max(1.0, gamma0)


def test_kernel_push_body_passes_schmitz_force_free(hc):
    """pusher_schmitz_test.py:86-123: Higuera-Cary exact, relativistic Boris off by 1e-3 .. 2e-1 rad after 20 steps."""
    gamma0 = 10.0
    v0 = np.sqrt(gamma0 ** 2 - 1.0) / gamma0
    E, B, dt = (0.0, v0, 0.0), (0.0, 0.0, 1.0), 0.2 * np.pi * gamma0
    v = _kernel_push(hc, "higuera_cary", (v0, 0.0, 0.0), E, B, dt, 20)
    assert abs(np.arctan2(v[1], v[0])) < 1.0e-12 and abs((_g(v) - gamma0) / (gamma0 - 1.0)) < 1.0e-11
    worst = 0.0
    for steps in range(1, 21):
        v = _kernel_push(hc, "boris", (v0, 0.0, 0.0), E, B, dt, steps)
        worst = max(worst, abs(np.arctan2(v[1], v[0])))
    assert 1.0e-3 < worst < 2.0e-1


@pytest.mark.parametrize("pn", ("boris", "higuera_cary"))
def test_kernel_push_body_passes_schmitz_oscillating_E(hc, pn):
    """pusher_schmitz_test.py:185-211"""
    gamma_perp = 1.1
    u = np.sqrt(gamma_perp ** 2 - 1.0)
    omega0 = 0.5 / gamma_perp
    E0, dt = 10.0 * omega0, (2.0 * np.pi / omega0) / 100
    v = _kernel_push(hc, pn, (u / gamma_perp, 0.0, 0.0), None, (0.0, 0.0, 1.0), dt, 500,
                     E_of_step=lambda s: (0.0, 0.0, E0 * np.cos(omega0 * (s + 0.5) * dt)))
    assert abs(v[2]) <= 1.0e-9 and abs(_g(v) - gamma_perp) <= 1.0e-9


def test_owner_offset_is_consistent_with_the_wrap_at_the_seam(hc):
    """Distributed ownership on a split periodic axis (top rank of two, box [0, h)) in float32: a particle whose x + h rounds up to
    `wind` is wrapped to -h and must be handed to the +1 neighbour (not kept with a position outside the local tile); a particle
    exactly on +h keeps the reference's alias only on the top rank -- handed across the seam it travels as -h; inside the box
    nothing changes.  (The charge-conservation check of bench.py at 128^3 cells per rank found one stranded electron per step.)"""
    f = np.float32
    wind = f(256 * 2.66e-4); h = f(0.5) * wind
    fn = hc.hc_owner_offset_f32
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_float, ctypes.c_float, ctypes.c_float]
    def call(x, lo, hi):
        v = ctypes.c_float(float(x))
        off = fn(ctypes.byref(v), ctypes.c_float(float(lo)), ctypes.c_float(float(hi)), ctypes.c_float(float(wind)))
        return off, f(v.value)
    # (ii) exactly on the upper wall of the top rank: leaves upward, arrives as -h
    off, w = call(h, f(0), h)
    assert off == 1 and w == -h
    # (i) the largest float below +h: x + h rounds to wind
    x = np.nextafter(h, f(0))
    assert f(x + h) == wind
    off, w = call(x, f(0), h)
    assert off == 1 and w == -h
    # interior particle of the top rank: untouched
    x = f(0.25) * h
    off, w = call(x, f(0), h)
    assert off == 0 and w == f(f(x + h) - h)
    # bottom rank [-h, 0): leaving downward wraps to just below +h (or the +h alias, which the top rank can hold)
    x = f(-h - f(1e-6))
    off, w = call(x, -h, f(0))
    assert off == -1 and f(0) < w <= h
    # an axis that is not split: the wrap stays on this rank
    off, w = call(f(h + f(1e-5)), f(-np.inf), f(np.inf))
    assert off == 0 and w < f(0)
