"""The oracle against the REFERENCE'S OWN SOURCE, executed.

/root/reference/PyPIC3D is pure Python on jax; jax is not installable here, so tests/jax_shim provides a NumPy-backed `jax`
(README there) and tests/ref_loader.py imports the reference's hot-path modules unmodified on top of it.  Every test below calls a
reference function and the oracle function that restates it on the same inputs and compares the results to a few ulp: this is
what pins `oracle/` (and through it the CUDA kernels) to the reference's code rather than to a reading of it.

The tests skip where /root/reference does not exist (the GPU box); tests/golden/ref_*.npz carry the reference's outputs there
(tests/golden/make_ref_golden.py, tests/test_golden.py)."""
import numpy as np
import pytest

from oracle import fixtures as fx, pusher as opush, deposition as odep, particles as opart, yee as oyee, halo as ohalo, filters as ofil
from oracle import evolve as oevolve, shapes as oshapes, stencil as ostencil
from tests import ref_loader
from tests.cases import CASES, make_case, make_fields

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present on this host")

RTOL = 2e-14      # a few ulp of float64: the only differences allowed are summation orders


@pytest.fixture(scope="module")
def ref():
    return ref_loader.reference()


def to_ref(ref, sp, dp, tp=None, sc=None):
    """Oracle pytrees -> the reference's own NamedTuples (same field names; field_mesh becomes the reference's device mesh)."""
    import jax.numpy as jnp
    P = ref.parameters
    mesh = ref.ghost_cells.make_field_mesh(tuple(int(v) for v in sp.field_mesh))
    rsp = P.StaticParameters(**{**sp._asdict(), "field_mesh": mesh})
    g = dp.grids
    conv = lambda t: tuple(jnp.asarray(a) for a in t)
    grids = P.GridParameters(vertex=conv(g.vertex), center=conv(g.center), tiled_vertex_grid=(), tiled_center_grid=())
    rdp = P.DynamicParameters(**{**dp._asdict(), "grids": grids})
    # the tiled coordinate lines come from the reference's own builder (utilities/grids.py:114-165); the oracle keeps only the
    # own-axis tile index of the same lines
    tc, tv = ref.grids.build_tiled_yee_grids(rsp, rdp)
    for a in range(3):
        own = [0, 0, 0]
        for t in range(np.asarray(tc[a]).shape[a]):
            own[a] = t
            assert np.array_equal(np.asarray(tc[a])[tuple(own)], np.asarray(g.tiled_center_grid[a])[t])
            assert np.array_equal(np.asarray(tv[a])[tuple(own)], np.asarray(g.tiled_vertex_grid[a])[t])
    rdp = rdp._replace(grids=P.GridParameters(vertex=conv(g.vertex), center=conv(g.center), tiled_vertex_grid=tv, tiled_center_grid=tc))
    out = [rsp, rdp]
    if tp is not None:
        C = ref.particle_class
        out.append(C.TiledParticles(x=jnp.asarray(tp.x), u=jnp.asarray(tp.u), active=jnp.asarray(tp.active)))
        out.append(C.SpeciesConfig(*[jnp.asarray(v) for v in sc]))
    return out


def vec(ref, F):
    import jax.numpy as jnp
    return tuple(jnp.asarray(c) for c in F)


def arr(a):
    import jax.numpy as jnp
    return jnp.asarray(a)


def close(a, b, what, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} vs {rtol:.1e} * {scale:.3e}"


# ------------------------------------------------------------------------------------------------ small pieces
def test_shape_weights(ref):
    d = np.linspace(-0.49, 0.49, 23)
    for a, b in zip(ref.shapes.get_first_order_weights(arr(d + 0.5), arr(d + 0.5), arr(d + 0.5), 1.3, 0.7, 2.0)[0], oshapes.first_order_weights(d + 0.5, 1.3)):
        close(a, b, "CIC weights", 0)
    for a, b in zip(ref.shapes.get_second_order_weights(arr(d), arr(d), arr(d), 1.3, 0.7, 2.0)[1], oshapes.second_order_weights(d, 0.7)):
        close(a, b, "TSC weights", 0)


@pytest.mark.parametrize("sf", (1, 2))
def test_anchor_offset_and_wrap(ref, sf):
    G = ref.grid_and_stencil
    axis = -2.0 - 0.5 + 0.5 * np.arange(12)
    pos = np.random.default_rng(0).uniform(-2.0, 2.0, 200)
    a_ref = G.compute_particle_anchor(arr(pos), arr(axis), sf)
    a_or = ostencil.compute_particle_anchor(pos, axis, sf)
    assert np.array_equal(np.asarray(a_ref), a_or)
    close(G.particle_axis_offset(arr(pos), a_ref, arr(axis)), ostencil.particle_axis_offset(pos, a_or, axis), "offset", 0)
    x = np.array([-2.0, 2.0, 2.0000001, -2.0000001, 5.3, -7.9, 0.0, 1.999999])
    close(G.wrap_periodic_position(arr(x), 4.0), ostencil.wrap_periodic_position(x, 4.0), "wrap", 0)


@pytest.mark.parametrize("name", ("boris", "relativistic_boris", "higuera_cary"))
def test_single_particle_pushers(ref, name):
    rng = np.random.default_rng(3)
    fn = {"boris": ref.boris.boris_single_particle, "relativistic_boris": ref.boris.relativistic_boris_single_particle,
          "higuera_cary": ref.higuera_cary.higuera_cary_single_particle}[name]
    ofn = {"boris": opush.boris, "relativistic_boris": opush.relativistic_boris, "higuera_cary": opush.higuera_cary}[name]
    sp, dp = fx.kernel_parameters(C=3.0)
    _, rdp = to_ref(ref, sp, dp)
    for _ in range(20):
        v = rng.uniform(-1.2, 1.2, 3); E = rng.normal(size=3); B = rng.normal(size=3)
        q, m, dt = rng.choice([-1.0, 2.0]), rng.uniform(0.5, 3.0), 0.07
        got = fn(v[0], v[1], v[2], E[0], E[1], E[2], B[0], B[1], B[2], q, m, dt, rdp)
        want = ofn(v, E, B, q, m, dt, 3.0)
        close(np.array([float(c) for c in got]), np.array(want), name, 5e-16)


@pytest.mark.parametrize("alpha", (1.0, 0.6))
def test_filters(ref, alpha):
    t = np.random.default_rng(3).normal(size=(2, 1, 2, 9, 8, 7))
    close(ref.filters.digital_filter(arr(t), alpha, 2), ofil.digital_filter(t, alpha, 2), "digital", 1e-15)
    close(ref.filters.bilinear_filter(arr(t), 2), ofil.bilinear_filter(t, 2), "bilinear", 1e-15)


# ------------------------------------------------------------------------------------------------ operators on tiled state
@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("bcs", [(0, 0, 0), (1, 0, 1)])
def test_ghost_cell_refresh_and_fold(ref, N, tile, bcs):
    """boundary_conditions/ghost_cells.py:615-736 on 1..8 fake devices (shard_map + ppermute through the shim's mailbox)."""
    sp, dp, tp, sc, E, B = make_case(N, tile, 1, boundary_conditions=bcs, particle_boundary_conditions=bcs)
    rsp, rdp = to_ref(ref, sp, dp)
    G = ref.ghost_cells
    t = np.random.default_rng(7).normal(size=E[0].shape)
    for bc_type in (0, 1):
        close(G.update_tiled_ghost_cells(arr(t), rsp, 2, bc_type=bc_type), ohalo.update_tiled_ghost_cells(t, sp, 2, bc_type), "refresh", 0)
        close(G.fold_tiled_ghost_cells(arr(t), rsp, 2, bc_type=bc_type), ohalo.fold_tiled_ghost_cells(t, sp, 2, bc_type), "fold", 4e-16)


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("pn,rel", [("boris", True), ("boris", False), ("higuera_cary", True)])
def test_particle_push(ref, N, tile, sf, pn, rel):
    sp, dp, tp, sc, E, B = make_case(N, tile, sf, particle_pusher=pn, relativistic=rel, C=2.5)
    rsp, rdp, rtp, rsc = to_ref(ref, sp, dp, tp, sc)
    got = ref.particle_push.particle_push(rtp, rsc, vec(ref, E), vec(ref, B), rsp, rdp)
    want = opush.particle_push(tp, sc, E, B, sp, dp)
    close(got.u, want.u, "pushed u")
    close(got.x, want.x, "x untouched", 0)


@pytest.mark.parametrize("N,tile", CASES + [((6, 6, 6), (6, 6, 6))])
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("pbc", [(0, 0, 0), (1, 0, 2)])
def test_esirkepov_current(ref, N, tile, sf, pbc):
    """deposition/Esirkepov.py:49-362 end to end (deposit, fold, refresh) -- the headline kernel's reference, executed."""
    sp, dp, tp, sc, E, B = make_case(N, tile, sf, current_deposition="esirkepov", particle_boundary_conditions=pbc, vmax=0.9)
    rsp, rdp, rtp, rsc = to_ref(ref, sp, dp, tp, sc)
    z = fx.empty_tiled_vector(sp, dp)
    got = ref.Esirkepov.Esirkepov_current(rtp, rsc, vec(ref, z), rsp, rdp)
    want = odep.Esirkepov_current(tp, sc, z, sp, dp)
    for c in range(3):
        close(got[c], want[c], f"J{c}")


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("flt", ("none", "bilinear", "digital"))
def test_direct_current_and_rho(ref, N, tile, sf, flt):
    sp, dp, tp, sc, E, B = make_case(N, tile, sf, current_deposition="direct", current_filter=flt, alpha=0.7)
    rsp, rdp, rtp, rsc = to_ref(ref, sp, dp, tp, sc)
    z = fx.empty_tiled_vector(sp, dp)
    got = ref.J_from_rhov.J_from_rhov(rtp, rsc, vec(ref, z), rsp, rdp)
    want = odep.J_from_rhov(tp, sc, z, sp, dp)
    for c in range(3):
        close(got[c], want[c], f"J{c}")
    zr = fx.empty_tiled_scalar(sp, dp)
    close(ref.rho.compute_rho(rtp, rsc, arr(zr), rsp, rdp), odep.compute_rho(tp, sc, zr, sp, dp), "rho")


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("bcs", [(0, 0, 0), (1, 0, 1)])
@pytest.mark.parametrize("alpha", (1.0, 0.8))
def test_yee_updates(ref, N, tile, bcs, alpha):
    sp, dp, tp, sc, E, B = make_case(N, tile, 1, boundary_conditions=bcs, alpha=alpha, C=1.3, eps=0.7)
    rsp, rdp = to_ref(ref, sp, dp)
    J = tuple(0.3 * c for c in B)
    got, pml = ref.first_order_yee.update_E(vec(ref, E), vec(ref, B), vec(ref, J), rsp, rdp)
    for a, b in zip(got, oyee.update_E(E, B, J, sp, dp)):
        close(a, b, "update_E", 1e-15)
    for flt in (False, True):
        got, _ = ref.first_order_yee.update_B(vec(ref, E), vec(ref, B), rsp, rdp, None, do_filter=flt)
        for a, b in zip(got, oyee.update_B(E, B, sp, dp, do_filter=flt)):
            close(a, b, "update_B", 1e-15)


@pytest.mark.parametrize("N,tile", CASES)
@pytest.mark.parametrize("pbc", [(0, 0, 0), (1, 0, 2), (2, 1, 0)])
@pytest.mark.parametrize("capacity", (2.0, 1.0))
def test_move_and_retile(ref, N, tile, pbc, capacity):
    """particles/particle_tile_communication.py:82-99, 292-453: slot-exact, including the overflow flag, on up to 8 fake devices."""
    sp, dp, tp, sc, E, B = make_case(N, tile, 1, particle_boundary_conditions=pbc, capacity=capacity, vmax=0.35)
    rsp, rdp, rtp, rsc = to_ref(ref, sp, dp, tp, sc)
    C = ref.particle_tile_communication
    moved_r = C.update_tiled_particle_positions(rtp, rsc, dp.dt)
    moved_o = opart.update_tiled_particle_positions(tp, sc, dp.dt)
    close(moved_r.x, moved_o.x, "moved x", 0)
    got, ovf = C.refresh_tiled_particle_tiles(moved_r, rsp, rdp)
    want, oovf = opart.refresh_tiled_particle_tiles(moved_o, sp, dp)
    assert np.array_equal(np.asarray(got.active), want.active)
    close(got.x, want.x, "retiled x", 0); close(got.u, want.u, "retiled u", 0)
    assert bool(np.asarray(ovf)) == bool(oovf)


# ------------------------------------------------------------------------------------------------ the whole step, several steps
STEP_CASES = [
    dict(N=(8, 6, 4), tile=(8, 6, 4), sf=1, dep="esirkepov", flt="none", bcs=(0, 0, 0), pbc=(0, 0, 0)),
    dict(N=(8, 6, 4), tile=(2, 3, 2), sf=2, dep="esirkepov", flt="none", bcs=(0, 0, 0), pbc=(0, 0, 0)),
    dict(N=(8, 8, 4), tile=(8, 8, 4), sf=1, dep="esirkepov", flt="none", bcs=(0, 0, 0), pbc=(2, 0, 1)),
    dict(N=(8, 1, 1), tile=(2, 1, 1), sf=2, dep="direct", flt="bilinear", bcs=(0, 0, 0), pbc=(0, 0, 0)),
    dict(N=(6, 6, 1), tile=(3, 2, 1), sf=2, dep="direct", flt="bilinear", bcs=(0, 0, 0), pbc=(0, 0, 0)),
    dict(N=(8, 1, 4), tile=(8, 1, 4), sf=2, dep="direct", flt="bilinear", bcs=(0, 0, 1), pbc=(0, 0, 0)),
]


@pytest.mark.parametrize("c", STEP_CASES)
def test_time_loop_electrodynamic_multi_step(ref, c):
    """evolve.py:16-103 for 5 steps, the reference's code against the oracle: a multi-step trajectory pinned to the reference."""
    sp, dp, tp, sc, E, B = make_case(c["N"], c["tile"], c["sf"], current_deposition=c["dep"], current_filter=c["flt"],
                                     boundary_conditions=c["bcs"], particle_boundary_conditions=c["pbc"], capacity=3.0, vmax=0.5, dt=0.04)
    fields = make_fields(sp, dp)
    rsp, rdp, rtp, rsc = to_ref(ref, sp, dp, tp, sc)
    import jax.numpy as jnp
    Er, Br, Jr, rho, phi, ext, pml, ovf = fields
    rfields = (vec(ref, Er), vec(ref, Br), vec(ref, Jr), jnp.asarray(rho), jnp.asarray(phi), (vec(ref, ext[0]), vec(ref, ext[1])), None,
               jnp.asarray(bool(ovf)))
    for step in range(5):
        rtp, rfields = ref.evolve.time_loop_electrodynamic(rtp, rsc, rfields, rsp, rdp)
        tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
        assert np.array_equal(np.asarray(rtp.active), tp.active), step
        close(rtp.x, tp.x, f"x step {step}", 1e-13); close(rtp.u, tp.u, f"u step {step}", 1e-13)
        for k, nm in enumerate("EBJ"):
            for a, b in zip(rfields[k], fields[k]):
                close(a, b, f"{nm} step {step}", 1e-13)
        assert bool(np.asarray(rfields[7])) == bool(fields[7])
