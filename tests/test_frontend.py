"""Config front end: TOML schema / defaults / particle loading on the host (CPU tests), and the
packaged two_stream / weibel style configurations run end-to-end on the GPU against the oracle from identical initial state."""
import os

import numpy as np
import pytest

from pypic3d_b200.initialization import default_parameters, update_parameters_from_toml
from pypic3d_b200.particles.particle_initialization import load_particles_from_toml, pack_species_arrays_into_tiles
from oracle import fixtures as fx

TWO_STREAM = {
    "simulation_parameters": {"name": "two stream", "t_wind": 5.413e-8, "solver": "electrodynamic_yee", "Nx": 100, "Ny": 1, "Nz": 1,
                              "particle_tile_nx": 100, "particle_tile_ny": 1, "particle_tile_nz": 1, "particle_tile_capacity_factor": 1.0,
                              "x_wind": 1, "y_wind": 1, "z_wind": 1, "verbose": False, "cfl": 0.99, "shape_factor": 2,
                              "current_calculation": "j_from_rhov", "alpha": 1.0, "relativistic": True, "bc": "periodic"},
    "plotting": {"plotting_interval": 8},
    "particle1": {"name": "electron1", "N_particles": 1500, "charge": -1.602e-19, "mass": 9.1093837e-31, "vth": 14989622.9,
                  "number_density": 1.0e15, "initial_vx": 0.5e8},
    "particle2": {"name": "electron2", "N_particles": 1500, "charge": -1.602e-19, "mass": 9.1093837e-31, "vth": 14989622.9, "initial_vx": -0.5e8},
    "particle3": {"name": "ion1", "N_particles": 3000, "charge": 1.602e-19, "mass": 1.67e-27, "vth": 34982.78402326697},
}


def test_defaults_and_toml_merge():
    plotting, static, dynamic = default_parameters()
    assert static["filter_j"] == "bilinear" and static["current_calculation"] == "j_from_rhov" and static["guard_cells"] == 2
    assert dynamic["C"] == 2.99792458e8 and dynamic["alpha"] == 1.0 and plotting["plotting_interval"] == 10
    cfg = {"simulation_parameters": {"Nx": 7, "cfl": 0.5, "bc": "periodic", "alpha": 0.9}, "constants": {"C": 1.0}, "plotting": {"plotting_interval": 3}}
    static, dynamic, plotting = update_parameters_from_toml(cfg, static, dynamic, plotting)
    assert dynamic["Nx"] == 7 and static["cfl"] == 0.5 and dynamic["alpha"] == 0.9 and "bc" not in static and dynamic["C"] == 2.99792458e8
    assert plotting["plotting_interval"] == 3


def test_default_parameters_reference_pins():
    """tests/code_tests/initialization_test.py:27-44, literal expectations."""
    plotting, sim, dynamic = default_parameters()
    assert "Nx" in dynamic and "particle_pusher" in sim and sim["particle_pusher"] == "boris"
    assert sim["solver"] == "electrodynamic_yee" and "electrostatic" not in sim and "fast_backend" not in sim
    assert sim["particle_x_bc"] == sim["particle_y_bc"] == sim["particle_z_bc"] == "periodic" and sim["guard_cells"] == 2
    for k in ("plot_vtk_particles", "plot_vtk_scalars", "plot_vtk_vectors"):
        assert k not in plotting
    assert "eps" in dynamic and "plotfields" in plotting


@pytest.mark.parametrize("solver", ("old_solver", "spectral"))
def test_initialize_simulation_rejects_unknown_solver(tmp_path, solver):
    """initialization_test.py:276-301: ValueError("Unsupported solver ...") before anything is allocated (no GPU needed)."""
    from pypic3d_b200.initialization import initialize_simulation
    cfg = {"simulation_parameters": {"name": "unknown solver test", "output_dir": str(tmp_path), "solver": solver, "Nx": 4, "Ny": 1,
                                     "Nz": 1, "x_wind": 1.0, "y_wind": 1.0, "z_wind": 1.0, "Nt": 1, "dt": 1.0e-10},
           "plotting": {"plotting": False}}
    with pytest.raises(ValueError, match="Unsupported solver"):
        initialize_simulation(cfg, verbose=False)


def test_initialize_simulation_rejects_bad_options_on_the_host(tmp_path):
    """initialization.py:97-124 validation (current_calculation / filter / tile divisibility), all raised on the host."""
    from pypic3d_b200.initialization import initialize_simulation
    base = {"name": "v", "output_dir": str(tmp_path), "Nx": 4, "Ny": 1, "Nz": 1, "x_wind": 1.0, "y_wind": 1.0, "z_wind": 1.0, "Nt": 1, "dt": 1e-10}
    for extra, msg in ((dict(current_calculation="villasenor"), "current_calculation"),
                       (dict(current_calculation="esirkepov", filter_j="bilinear"), "Esirkepov current filtering"),
                       (dict(filter_j="boxcar"), "filter_j"),
                       (dict(particle_tile_nx=3), "divide the physical grid")):
        with pytest.raises(ValueError, match=msg):
            initialize_simulation({"simulation_parameters": dict(base, **extra), "plotting": {"plotting": False}}, verbose=False)


def test_particle_loading_weight_carry_over_and_packing():
    sp, dp = fx.kernel_parameters(Nx=100, Ny=1, Nz=1, x_wind=1.0, y_wind=1.0, z_wind=1.0, tile_shape=(50, 1, 1), kb=1.380649e-23)
    np.random.seed(0)
    tp, sc, names, meta = load_particles_from_toml(TWO_STREAM, sp, dp, verbose=False)
    assert names == ("electron1", "electron2", "ion1")
    w = (1.0e15 / (1500 / 100)) * (0.01 * 1 * 1)
    assert np.allclose(sc.weight, [w, w, w])                      # weight leaks to later blocks (particle_initialization.py:240)
    assert tp.x.shape[:4] == (2, 1, 1, 3) and tp.active.sum() == 6000
    assert abs(tp.u[..., 0][tp.active[..., :]][:10].mean()) > 0    # drifts applied
    # every particle sits in the tile that owns it
    cell = np.floor((tp.x[..., 0] + 0.5) / 0.01).astype(int) // 50
    for t in range(2):
        assert np.all(cell[t][tp.active[t]] == t)
    # capacity = ceil(max count * factor)
    x = np.zeros((10, 3)); x[:, 0] = np.linspace(-0.49, 0.49, 10)
    xt, ut, at, counts = pack_species_arrays_into_tiles([(x, x * 0, np.ones(10, bool))], dp, (50, 1, 1), 1.5)
    assert at.shape[-1] == int(np.ceil(counts.max() * 1.5))


@pytest.mark.gpu
@pytest.mark.parametrize("tile_nx,dep", [(100, "j_from_rhov"), (50, "j_from_rhov"), (100, "esirkepov")])
def test_two_stream_config_matches_oracle(tmp_path, tile_nx, dep):
    """demos/two_stream/two_stream.toml (as packaged, and split in two tiles, and with Esirkepov): 20 steps from the
    identical initial state, energies and fields vs the oracle."""
    import torch
    from tests import gpu_util as gu
    gu.require_cuda()
    from pypic3d_b200.initialization import initialize_simulation
    from pypic3d_b200.__main__ import run_PyPIC3D
    from oracle import evolve as oevolve, diagnostics as odiag
    from oracle.params import StaticParameters as OS, DynamicParameters as OD, GridParameters as OG, TiledParticles as OT, SpeciesConfig as OC
    cfg = {k: dict(v) for k, v in TWO_STREAM.items()}
    cfg["simulation_parameters"].update(output_dir=str(tmp_path), particle_tile_nx=tile_nx, particle_tile_capacity_factor=1.5, Nt=20,
                                        current_calculation=dep, filter_j="bilinear" if dep == "j_from_rhov" else "none")
    np.random.seed(0)
    loop, particles, fields, sp, dp, plotting, plasma, species = initialize_simulation(cfg, verbose=False)
    # oracle from the identical initial state
    osp = OS(**sp._asdict()); odp = OD(**{**dp._asdict(), "grids": OG(**dp.grids._asdict())})
    otp = OT(gu.npy(particles.x), gu.npy(particles.u), gu.npy(particles.active))
    osc = OC(*[np.asarray(v) for v in species])
    n = lambda F: tuple(gu.npy(c) for c in F)
    of = (n(fields[0]), n(fields[1]), n(fields[2]), gu.npy(fields[3]), gu.npy(fields[4]), (n(fields[5][0]), n(fields[5][1])), None, False)
    for _ in range(20):
        otp, of = oevolve.time_loop_electrodynamic(otp, osc, of, osp, odp)
    np.random.seed(0)
    sp2, dp2, plotting2, plasma2, gp, gf, species2 = run_PyPIC3D(cfg, verbose=False)
    assert np.array_equal(gu.npy(gp.active), otp.active)
    gu.assert_close(gp.x, otp.x, 1e-10, "x")
    scale_u = np.abs(otp.u).max()
    assert np.abs(gu.npy(gp.u) - otp.u).max() <= 1e-10 * scale_u
    for k in range(3):
        for a, b in zip(gf[k], of[k]):
            gu.assert_close(a, b, 1e-9, "EBJ"[k])
    # energy history file written with the reference's row format, and it matches the oracle's energies at t=0
    rows = open(os.path.join(str(tmp_path), "data", "total_energy.txt")).read().strip().splitlines()
    assert len(rows) == 3 and rows[0].startswith("0.0, ")


@pytest.mark.gpu
def test_electrostatic_config_matches_oracle(tmp_path):
    """solver = "electrostatic" (two-stream style plasma on 16x4x4): one tile, collocated grids (initialization.py:251-254,
    310-313), evolve.time_loop_electrostatic selected; 5 steps from the identical initial state against the oracle."""
    from tests import gpu_util as gu
    gu.require_cuda()
    from pypic3d_b200.initialization import initialize_simulation
    from pypic3d_b200.evolve import time_loop_electrostatic
    from pypic3d_b200.__main__ import run_PyPIC3D
    from oracle import evolve as oevolve
    from oracle.params import StaticParameters as OS, DynamicParameters as OD, GridParameters as OG, TiledParticles as OT, SpeciesConfig as OC
    cfg = {k: dict(v) for k, v in TWO_STREAM.items()}
    cfg["simulation_parameters"].update(output_dir=str(tmp_path), solver="electrostatic", Nx=16, Ny=4, Nz=4, particle_tile_nx=8,
                                        particle_tile_ny=4, particle_tile_nz=4, particle_tile_capacity_factor=1.5, Nt=5, shape_factor=1)
    for k in ("particle1", "particle2", "particle3"):
        cfg[k]["N_particles"] = cfg[k]["N_particles"] // 5
    # The reference's stopping rule is absolute (sum r^2 <= 1e-24 with r = rho/eps0 in SI units).  At laboratory densities it is
    # unreachable: every solve then runs its 5000 iterations at round-off stagnation, the constant null-space mode of the
    # periodic operator drifts and the result depends on the summation order at the 1e-8..1e-3 level -- nothing to pin.  A
    # tenuous plasma keeps the rule reachable (a dozen iterations), so the comparison is meaningful.
    cfg["particle1"]["number_density"] = 1.0e-3
    np.random.seed(0)
    loop, particles, fields, sp, dp, plotting, plasma, species = initialize_simulation(cfg, verbose=False)
    assert loop is time_loop_electrostatic and sp.electrostatic and tuple(sp.tile_shape) == (16, 4, 4)
    assert np.array_equal(np.asarray(dp.grids.center[0]), np.asarray(dp.grids.vertex[0]))
    osp = OS(**sp._asdict()); odp = OD(**{**dp._asdict(), "grids": OG(**dp.grids._asdict())})
    otp = OT(gu.npy(particles.x), gu.npy(particles.u), gu.npy(particles.active))
    osc = OC(*[np.asarray(v) for v in species])
    n = lambda F: tuple(gu.npy(c) for c in F)
    of = (n(fields[0]), n(fields[1]), n(fields[2]), gu.npy(fields[3]), gu.npy(fields[4]), (n(fields[5][0]), n(fields[5][1])), None, False)
    for _ in range(5):
        otp, of = oevolve.time_loop_electrostatic(otp, osc, of, osp, odp)
    np.random.seed(0)
    sp2, dp2, plotting2, plasma2, gp, gf, species2 = run_PyPIC3D(cfg, verbose=False)
    assert np.array_equal(gu.npy(gp.active), otp.active)
    gu.assert_close(gp.x, otp.x, 1e-10, "x")
    assert np.abs(gu.npy(gp.u) - otp.u).max() <= 1e-9 * np.abs(otp.u).max()
    scale = max(np.abs(np.asarray(c)).max() for c in of[0])
    assert scale > 0
    for a, b in zip(gf[0], of[0]):       # (a one-iteration difference at the stopping threshold moves E by ~1e-3 of its scale)
        assert np.abs(gu.npy(a) - np.asarray(b)).max() <= 1e-2 * scale


WEIBEL = {   # /root/reference/demos/weibel/weibel.toml as packaged (1-D, 201 cells, TSC, j_from_rhov + bilinear filter by default)
    "simulation_parameters": {"name": "Demo of a 1D Weibel instability with 3,000 electrons", "Nt": 6000, "bc": "periodic",
                              "solver": "electrodynamic_yee", "Nx": 201, "Ny": 1, "Nz": 1, "x_wind": 3e-1, "y_wind": 1, "z_wind": 1,
                              "verbose": False, "cfl": 0.99, "shape_factor": 2, "relativistic": True},
    "plotting": {"plot_errors": False, "phaseSpace": False, "plotfields": False, "plotting_interval": 30},
    "particle1": {"name": "electron1", "N_per_cell": 40, "charge": -1.602e-19, "mass": 9.1093837e-31, "number_density": 1.0e15,
                  "Tx": 296494.82871735515, "Tz": 29649482.87173552, "Ty": 0},
    "particle2": {"name": "ion1", "N_per_cell": 40, "charge": 1.602e-19, "mass": 1.67e-27, "Tx": 296494.82871735515,
                  "Tz": 29649482.87173552, "Ty": 0},
}


@pytest.mark.gpu
@pytest.mark.parametrize("dep", ["j_from_rhov", "esirkepov"])
def test_weibel_config_matches_oracle(tmp_path, dep):
    """demos/weibel/weibel.toml as packaged (SURVEY section 8d input 2), and its Esirkepov variant: 30 steps from the identical
    initial state through `run_PyPIC3D` (resident path) against the oracle; the anisotropy-driven B_y growth is compared too."""
    from tests import gpu_util as gu
    gu.require_cuda()
    from pypic3d_b200.initialization import initialize_simulation
    from pypic3d_b200.__main__ import run_PyPIC3D
    from oracle import evolve as oevolve
    from oracle.params import StaticParameters as OS, DynamicParameters as OD, GridParameters as OG, TiledParticles as OT, SpeciesConfig as OC
    cfg = {k: dict(v) for k, v in WEIBEL.items()}
    cfg["simulation_parameters"].update(output_dir=str(tmp_path), Nt=30, current_calculation=dep,
                                        filter_j="bilinear" if dep == "j_from_rhov" else "none")
    np.random.seed(0)
    loop, particles, fields, sp, dp, plotting, plasma, species = initialize_simulation(cfg, verbose=False)
    assert tuple(sp.tile_shape) == (201, 1, 1) and int(sp.shape_factor) == 2 and int(particles.active.sum()) == 2 * 40 * 201
    osp = OS(**sp._asdict()); odp = OD(**{**dp._asdict(), "grids": OG(**dp.grids._asdict())})
    otp = OT(gu.npy(particles.x), gu.npy(particles.u), gu.npy(particles.active))
    osc = OC(*[np.asarray(v) for v in species])
    n = lambda F: tuple(gu.npy(c) for c in F)
    of = (n(fields[0]), n(fields[1]), n(fields[2]), gu.npy(fields[3]), gu.npy(fields[4]), (n(fields[5][0]), n(fields[5][1])), None, False)
    for _ in range(30):
        otp, of = oevolve.time_loop_electrodynamic(otp, osc, of, osp, odp)
    np.random.seed(0)
    sp2, dp2, plotting2, plasma2, gp, gf, species2 = run_PyPIC3D(cfg, verbose=False)
    assert np.array_equal(gu.npy(gp.active), otp.active)
    gu.assert_close(gp.x, otp.x, 1e-10, "x")
    assert np.abs(gu.npy(gp.u) - otp.u).max() <= 1e-9 * np.abs(otp.u).max()
    for k in range(3):
        scale = max(max(np.abs(np.asarray(c)).max() for c in of[k]), 1e-300)
        for a, b in zip(gf[k], of[k]):
            assert np.abs(gu.npy(a) - np.asarray(b)).max() <= 1e-8 * scale, "EBJ"[k]
    by_gpu = float((gu.npy(gf[1][1]) ** 2).sum()); by_ref = float((np.asarray(of[1][1]) ** 2).sum())
    assert by_ref > 0 and abs(by_gpu - by_ref) <= 1e-7 * by_ref


@pytest.mark.gpu
def test_harris_sheet_config_matches_oracle(tmp_path):
    """demos/reconnection_2d/harris_current.toml scaled down (48 x 1 x 48 instead of 500 x 1 x 500; SURVEY section 8d input 3):
    x periodic / z conducting fields, `[field1]` / `[field2]` .npy magnetic field (types 3 and 5), four species of which two
    come from .npy position / velocity files (the demo's initial_conditions.py recipe, seeded, with the sheet profile
    sampled properly), TSC, j_from_rhov + bilinear,
    cfl 0.1.  The per-species `z_bc = "reflecting"` keys are metadata only in the reference (particle BCs stay periodic).
    6 steps from the identical initial state against the oracle."""
    from tests import gpu_util as gu
    gu.require_cuda()
    from pypic3d_b200.initialization import initialize_simulation
    from pypic3d_b200.__main__ import run_PyPIC3D
    from oracle import evolve as oevolve
    from oracle.params import StaticParameters as OS, DynamicParameters as OD, GridParameters as OG, TiledParticles as OT, SpeciesConfig as OC
    # ---- the demo's initial_conditions.py on a smaller grid, seeded
    eps, me, c, q = 8.854e-12, 9.10938356e-31, 2.99792458e8, 1.602e-19
    vth = 0.05 * c
    n_peak, n_bg = 4000, 1200
    n0 = 1e22
    nb = 0.3 * n0
    di = c / (q * np.sqrt(n0 / me / eps))
    wind = 15 * di
    nx = nz = 48
    lam, B0 = 0.5 * di, 0.2
    x = np.linspace(-wind / 2, wind / 2, nx); z = np.linspace(-wind / 2, wind / 2, nz)
    X, Z = np.meshgrid(x, z, indexing="ij")
    Bx = B0 * np.tanh(Z / lam) + 100 * B0 * np.cos(2 * np.pi * X / wind) * np.sin(np.pi * Z / wind)
    Bz = -100 * B0 * np.sin(2 * np.pi * X / wind) * np.cos(np.pi * Z / wind)
    rng = np.random.default_rng(7)
    files = {"Bx": Bx[:, None, :], "Bz": Bz[:, None, :]}
    for sp_name in ("electron", "ion"):
        files[f"{sp_name}_x"] = rng.uniform(-wind / 2, wind / 2, n_peak)
        files[f"{sp_name}_y"] = np.zeros(n_peak)
        # sech^2 sheet by inverse CDF (the demo's own line stores the PROFILE VALUE 1/cosh^2 as the coordinate, which puts every
        # particle up to a metre outside the 0.8 mm box; the sheet it means to load is this one)
        files[f"{sp_name}_z"] = np.clip(lam * np.arctanh(rng.uniform(-0.999, 0.999, n_peak)), -0.45 * wind, 0.45 * wind)
        for a in "xyz":
            files[f"{sp_name}_v{a}"] = rng.normal(0, vth, n_peak)
    zbar = files["electron_z"] / lam
    files["electron_vy"] = files["electron_vy"] + (-c * B0 / (4 * np.pi * q * lam)) / np.cosh(zbar) ** 2 / (n0 * np.cosh(zbar) ** 2 + nb)
    for k, v in files.items():
        np.save(tmp_path / f"{k}.npy", v)
    path = lambda k: str(tmp_path / f"{k}.npy")
    weight = n0 / n_peak * wind * wind
    species = lambda name, n, charge, pre=None: dict(
        {"name": name, "N_particles": n, "charge": charge, "mass": 9.1093837e-31, "vth": 14989622.9, "weight": weight,
         "x_bc": "periodic", "z_bc": "reflecting", "y_bc": "periodic"},
        **({} if pre is None else {f"initial_{a}": path(f"{pre}_{a}") for a in ("x", "y", "z", "vx", "vy", "vz")}))
    cfg = {
        "simulation_parameters": {"name": "harris", "Nt": 6, "x_bc": "periodic", "z_bc": "conducting", "solver": "electrodynamic_yee",
                                  "Nx": nx, "Ny": 1, "Nz": nz, "x_wind": wind, "y_wind": 1, "z_wind": wind, "verbose": False,
                                  "cfl": 0.1, "shape_factor": 2, "relativistic": True, "output_dir": str(tmp_path),
                                  "particle_tile_capacity_factor": 1.5},
        "plotting": {"plotting_interval": 50},
        "field1": {"name": "Bx field", "type": 3, "path": path("Bx")},
        "field2": {"name": "Bz field", "type": 5, "path": path("Bz")},
        "particle1": species("peak electrons", n_peak, -1.602e-19, "electron"),
        "particle2": species("peak ions", n_peak, 1.602e-19, "ion"),
        "particle3": species("background electrons", n_bg, -1.602e-19),
        "particle4": species("background ions", n_bg, 1.602e-19),
    }
    np.random.seed(0)
    loop, particles, fields, sp, dp, plotting, plasma, spc = initialize_simulation(cfg, verbose=False)
    assert tuple(sp.boundary_conditions) == (0, 0, 1) and tuple(sp.particle_boundary_conditions) == (0, 0, 0)
    assert int(particles.active.sum()) == 2 * (n_peak + n_bg)
    g = int(sp.guard_cells)
    assert np.allclose(gu.npy(fields[1][0])[0, 0, 0, g:-g, g, g:-g], Bx)          # [field1] landed in Bx's interior
    osp = OS(**sp._asdict()); odp = OD(**{**dp._asdict(), "grids": OG(**dp.grids._asdict())})
    otp = OT(gu.npy(particles.x), gu.npy(particles.u), gu.npy(particles.active))
    osc = OC(*[np.asarray(v) for v in spc])
    n = lambda F: tuple(gu.npy(c) for c in F)
    of = (n(fields[0]), n(fields[1]), n(fields[2]), gu.npy(fields[3]), gu.npy(fields[4]), (n(fields[5][0]), n(fields[5][1])), None, False)
    for _ in range(6):
        otp, of = oevolve.time_loop_electrodynamic(otp, osc, of, osp, odp)
    np.random.seed(0)
    sp2, dp2, plotting2, plasma2, gp, gf, spc2 = run_PyPIC3D(cfg, verbose=False)
    assert np.array_equal(gu.npy(gp.active), otp.active)
    gu.assert_close(gp.x, otp.x, 1e-10, "x")
    assert np.abs(gu.npy(gp.u) - otp.u).max() <= 1e-9 * np.abs(otp.u).max()
    for k in range(3):
        scale = max(max(np.abs(np.asarray(c)).max() for c in of[k]), 1e-300)
        for a, b in zip(gf[k], of[k]):
            assert np.abs(gu.npy(a) - np.asarray(b)).max() <= 1e-8 * scale, "EBJ"[k]


@pytest.mark.gpu
def test_initialize_simulation_reference_pins(tmp_path):
    """initialization_test.py:138-274 with the reference's literal configurations: the Courant dt is computed when no dt is
    given (:138-183); the GLOBAL particle boundary conditions are encoded as (1, 2, 0) while the per-species x_bc stays
    metadata (:185-231); an electrostatic configuration gets the electrostatic loop, 6-D field tiles and collocated grids
    (:233-274)."""
    from tests import gpu_util as gu
    gu.require_cuda()
    from pypic3d_b200.initialization import initialize_simulation
    from pypic3d_b200.evolve import time_loop_electrostatic
    from pypic3d_b200.particles.particle_class import TiledParticles
    zeros4, x4, zeros1 = str(tmp_path / "zeros4.npy"), str(tmp_path / "x4.npy"), str(tmp_path / "zeros1.npy")
    np.save(x4, np.array([-0.375, -0.125, 0.125, 0.375])); np.save(zeros4, np.zeros(4)); np.save(zeros1, np.zeros(1))
    sp_block = lambda n, z, x=None: {"name": "electrons", "N_particles": n, "charge": -1.0, "mass": 1.0, "temperature": 1.0,
                                     "initial_x": x or z, "initial_y": z, "initial_z": z, "initial_vx": z, "initial_vy": z, "initial_vz": z}
    cfg = {"simulation_parameters": {"name": "courant dt tiled runtime test", "output_dir": str(tmp_path), "Nx": 4, "Ny": 1, "Nz": 1,
                                     "x_wind": 1.0, "y_wind": 1.0, "z_wind": 1.0, "Nt": 1, "particle_tile_nx": 4, "particle_tile_ny": 1,
                                     "particle_tile_nz": 1, "filter_j": "none"},
           "plotting": {"plotting": False}, "particle1": sp_block(4, zeros4, x4)}
    assert float(initialize_simulation(cfg, verbose=False)[4].dt) > 0.0
    cfg = {"simulation_parameters": {"name": "global particle bc test", "output_dir": str(tmp_path), "solver": "electrodynamic_yee",
                                     "Nx": 1, "Ny": 1, "Nz": 1, "x_wind": 1.0, "y_wind": 1.0, "z_wind": 1.0, "Nt": 1, "dt": 1.0e-10,
                                     "particle_x_bc": "reflecting", "particle_y_bc": "absorbing", "particle_z_bc": "periodic"},
           "plotting": {"plotting": False}, "particle1": dict(sp_block(1, zeros1), x_bc="absorbing")}
    _, particles, _, sp, *_ = initialize_simulation(cfg, verbose=False)
    assert tuple(sp.particle_boundary_conditions) == (1, 2, 0) and isinstance(particles, TiledParticles)
    cfg = {"simulation_parameters": {"name": "electrostatic collocated grid test", "output_dir": str(tmp_path), "solver": "electrostatic",
                                     "Nx": 4, "Ny": 2, "Nz": 1, "x_wind": 1.0, "y_wind": 1.0, "z_wind": 1.0, "Nt": 1, "dt": 1.0e-10},
           "plotting": {"plotting": False}, "particle1": sp_block(1, zeros1)}
    loop, particles, fields, sp, dp, *_ = initialize_simulation(cfg, verbose=False)
    assert loop is time_loop_electrostatic and isinstance(particles, TiledParticles) and fields[0][0].ndim == 6
    for v, c in zip(dp.grids.vertex, dp.grids.center):
        assert np.allclose(np.asarray(v), np.asarray(c))


# ---- tests/code_tests/particle_initialization_test.py:188-424 (literal loader KATs; host-side NumPy, no GPU) -----
def _loader_params(N, wind, tile):
    return fx.kernel_parameters(Nx=N[0], Ny=N[1], Nz=N[2], x_wind=wind[0], y_wind=wind[1], z_wind=wind[2], dx=1.0, dy=1.0, dz=1.0, dt=0.1,
                                tile_shape=tile, kb=1.0, eps=1.0, shape_factor=1)


def _npy(tmp_path, name, values):
    path = str(tmp_path / name)
    np.save(path, np.asarray(values, dtype=float))
    return path


def test_loader_uses_tile_axes_before_species(tmp_path):
    """particle_initialization_test.py:188-261"""
    sp, dp = _loader_params((4, 2, 1), (4.0, 2.0, 1.0), (2, 1, 1))
    cfg = {"particle1": {"name": "electrons", "N_particles": 3, "charge": -1.0, "mass": 2.0, "weight": 4.0, "temperature": 1.0,
                         "initial_x": _npy(tmp_path, "x.npy", [-1.5, 0.5, 1.5]), "initial_y": _npy(tmp_path, "y.npy", [-0.5, 0.5, 0.5]),
                         "initial_z": _npy(tmp_path, "z.npy", [0.0, 0.0, 0.0]), "initial_vx": _npy(tmp_path, "vx.npy", [0.1, 0.2, 0.3]),
                         "initial_vy": _npy(tmp_path, "vy.npy", [0.0, 0.0, 0.0]), "initial_vz": _npy(tmp_path, "vz.npy", [1.0, 2.0, 3.0])}}
    tp, sc, names, meta = load_particles_from_toml(cfg, sp, dp, verbose=False)
    assert names == ("electrons",) and meta[0]["name"] == "electrons"
    assert tp.x.shape == (2, 2, 1, 1, 2, 3) and tp.u.shape == (2, 2, 1, 1, 2, 3)
    assert tp.active[0, 0, 0, 0, 0] and tp.active[1, 1, 0, 0, 0] and tp.active[1, 1, 0, 0, 1] and int(tp.active.sum()) == 3
    assert np.allclose(tp.x[0, 0, 0, 0, 0], [-1.5, -0.5, 0.0]) and np.allclose(tp.x[1, 1, 0, 0, 0], [0.5, 0.5, 0.0])
    assert np.allclose(tp.x[1, 1, 0, 0, 1], [1.5, 0.5, 0.0]) and np.allclose(tp.u[1, 1, 0, 0, 1], [0.3, 0.0, 3.0])
    assert np.allclose(sc.charge, [-1.0]) and np.allclose(sc.mass, [2.0]) and np.allclose(sc.weight, [4.0])


def test_loader_preserves_interleaved_tile_order(tmp_path):
    """particle_initialization_test.py:263-350"""
    sp, dp = _loader_params((4, 1, 1), (4.0, 1.0, 1.0), (2, 1, 1))
    y = _npy(tmp_path, "y.npy", [0.0] * 4); z = _npy(tmp_path, "z.npy", [0.0] * 4)
    vy = _npy(tmp_path, "vy.npy", [0.0] * 4); vz = _npy(tmp_path, "vz.npy", [1.0, 2.0, 3.0, 4.0])
    blk = lambda name, q, m, w, x, vx: {"name": name, "N_particles": 4, "charge": q, "mass": m, "weight": w, "temperature": 1.0,
                                        "initial_x": x, "initial_y": y, "initial_z": z, "initial_vx": vx, "initial_vy": vy, "initial_vz": vz}
    cfg = {"particle1": blk("electrons", -1.0, 2.0, 4.0, _npy(tmp_path, "ex.npy", [-1.5, 0.5, -0.5, 1.5]), _npy(tmp_path, "evx.npy", [10.0, 20.0, 30.0, 40.0])),
           "particle2": blk("ions", 1.0, 3.0, 5.0, _npy(tmp_path, "ix.npy", [1.5, -1.5, 0.5, -0.5]), _npy(tmp_path, "ivx.npy", [100.0, 200.0, 300.0, 400.0]))}
    tp, sc, names, meta = load_particles_from_toml(cfg, sp, dp, verbose=False)
    assert names == ("electrons", "ions") and tuple(m["name"] for m in meta) == ("electrons", "ions")
    assert tp.x.shape == (2, 1, 1, 2, 2, 3) and int(tp.active.sum()) == 8
    assert np.allclose(tp.x[0, 0, 0, 0, :, 0], [-1.5, -0.5]) and np.allclose(tp.u[0, 0, 0, 0, :, 0], [10.0, 30.0])
    assert np.allclose(tp.x[1, 0, 0, 0, :, 0], [0.5, 1.5]) and np.allclose(tp.u[1, 0, 0, 0, :, 0], [20.0, 40.0])
    assert np.allclose(tp.x[0, 0, 0, 1, :, 0], [-1.5, -0.5]) and np.allclose(tp.u[0, 0, 0, 1, :, 0], [200.0, 400.0])
    assert np.allclose(tp.x[1, 0, 0, 1, :, 0], [1.5, 0.5]) and np.allclose(tp.u[1, 0, 0, 1, :, 0], [100.0, 300.0])
    assert np.allclose(sc.charge, [-1.0, 1.0]) and np.allclose(sc.mass, [2.0, 3.0]) and np.allclose(sc.weight, [4.0, 5.0])


def test_loader_maps_update_flags(tmp_path):
    """particle_initialization_test.py:357-420"""
    sp, dp = _loader_params((1, 1, 1), (1.0, 1.0, 1.0), (1, 1, 1))
    zero = _npy(tmp_path, "zero.npy", [0.0])
    cfg = {"particle1": {"name": "partly fixed", "N_particles": 1, "charge": 1.0, "mass": 1.0, "temperature": 1.0, "initial_x": zero,
                         "initial_y": zero, "initial_z": zero, "initial_vx": 0.0, "initial_vy": 0.0, "initial_vz": 0.0, "update_pos": True,
                         "update_x": True, "update_y": False, "update_z": True, "update_v": True, "update_vx": False, "update_vy": True,
                         "update_vz": False}}
    tp, sc, names, meta = load_particles_from_toml(cfg, sp, dp, verbose=False)
    assert bool(sc.update_x[0, 0]) and not bool(sc.update_x[0, 1]) and bool(sc.update_x[0, 2])
    assert not bool(sc.update_u[0, 0]) and bool(sc.update_u[0, 1]) and not bool(sc.update_u[0, 2])


# ---- tests/code_tests/utils_test.py:187-284 (external-field loading; host-side NumPy, no GPU) -----
def _ext_setup(tmp_path, name, value_shape, value):
    sp, dp = fx.kernel_parameters(Nx=2, Ny=2, Nz=2, x_wind=1.0, y_wind=1.0, z_wind=1.0, guard_cells=2)
    shape = (1, 1, 1, 6, 6, 6)
    fields = [np.zeros(shape) for _ in range(9)]
    ext = (tuple(np.zeros(shape) for _ in range(3)), tuple(np.zeros(shape) for _ in range(3)))
    path = str(tmp_path / name)
    np.save(path, np.ones(value_shape) * value)
    return sp, dp, fields, ext, path, (0, 0, 0, slice(2, -2), slice(2, -2), slice(2, -2))


def test_add_external_fields_adds_components():
    """utils_test.py:187-200 (oracle restatement of utils.py:205-216)"""
    from oracle.evolve import add_external_fields
    E = tuple(np.ones((2, 2, 2)) * v for v in (1, 2, 3)); B = tuple(np.ones((2, 2, 2)) * v for v in (4, 5, 6))
    xE = tuple(np.ones((2, 2, 2)) * v for v in (10, 20, 30)); xB = tuple(np.ones((2, 2, 2)) * v for v in (40, 50, 60))
    tE, tB = add_external_fields(E, B, (xE, xB))
    for got, want in zip(tE + tB, (11.0, 22.0, 33.0, 44.0, 55.0, 66.0)):
        assert np.allclose(got, want)


def test_load_external_fields_reference_pins(tmp_path):
    """utils_test.py:202-284: default / evolve=true -> evolved fields; evolve=false -> external-only; external currents and
    wrong shapes are rejected."""
    from pypic3d_b200.initialization import load_external_fields_from_toml
    sp, dp, fields, ext, path, I = _ext_setup(tmp_path, "ex.npy", (2, 2, 2), 1.0)
    fields, ext = load_external_fields_from_toml(fields, ext, {"field1": {"name": "Ex", "type": 0, "path": path}}, sp, dp)
    assert np.allclose(fields[0][I], 1.0) and np.allclose(ext[0][0], 0.0)
    sp, dp, fields, ext, path, I = _ext_setup(tmp_path, "by.npy", (2, 2, 2), 3.0)
    fields, ext = load_external_fields_from_toml(fields, ext, {"field1": {"name": "By", "type": 4, "path": path, "evolve": True}}, sp, dp)
    assert np.allclose(fields[4][I], 3.0) and np.allclose(ext[1][1], 0.0)
    sp, dp, fields, ext, path, I = _ext_setup(tmp_path, "bz.npy", (2, 2, 2), 5.0)
    fields, ext = load_external_fields_from_toml(fields, ext, {"field1": {"name": "external Bz", "type": 5, "path": path, "evolve": False}}, sp, dp)
    assert np.allclose(fields[5][I], 0.0) and np.allclose(ext[1][2][I], 5.0)
    sp, dp, fields, ext, path, I = _ext_setup(tmp_path, "jx.npy", (2, 2, 2), 1.0)
    with pytest.raises(ValueError, match="External-only fields must be electric or magnetic"):
        load_external_fields_from_toml(fields, ext, {"field1": {"name": "external Jx", "type": 6, "path": path, "evolve": False}}, sp, dp)
    sp, dp, fields, ext, path, I = _ext_setup(tmp_path, "wrong.npy", (3, 2, 2), 1.0)
    with pytest.raises(ValueError, match="Shape mismatch"):
        load_external_fields_from_toml(fields, ext, {"field1": {"name": "wrong Ex", "type": 0, "path": path, "evolve": False}}, sp, dp)


@pytest.mark.gpu
def test_two_stream_growth_rate_and_energy_history(tmp_path):
    """The observable the north star names: demos/two_stream/two_stream.toml as packaged, 120 steps (the whole linear phase: the
    field energy grows ~20x and starts to saturate).  The electric-field energy history written by `run_PyPIC3D` (reference file
    format) equals the oracle's from the identical initial state, the fitted growth rates agree, and both sit in the band the
    warm-beam (vth / v0 = 0.3), 1500-particle-per-beam run can reach of the cold-beam rate w_p / (2 sqrt 2) (0.40 - 0.46 over seeds)."""
    from tests import gpu_util as gu
    gu.require_cuda()
    from pypic3d_b200.initialization import initialize_simulation
    from pypic3d_b200.__main__ import run_PyPIC3D
    from oracle import evolve as oevolve, diagnostics as odiag
    from oracle.params import StaticParameters as OS, DynamicParameters as OD, GridParameters as OG, TiledParticles as OT, SpeciesConfig as OC
    cfg = {k: dict(v) for k, v in TWO_STREAM.items()}
    cfg["simulation_parameters"].update(output_dir=str(tmp_path), Nt=121, particle_tile_capacity_factor=1.5)
    np.random.seed(0)
    loop, particles, fields, sp, dp, plotting, plasma, species = initialize_simulation(cfg, verbose=False)
    osp = OS(**sp._asdict()); odp = OD(**{**dp._asdict(), "grids": OG(**dp.grids._asdict())})
    otp = OT(gu.npy(particles.x), gu.npy(particles.u), gu.npy(particles.active))
    osc = OC(*[np.asarray(v) for v in species])
    n = lambda F: tuple(gu.npy(c) for c in F)
    of = (n(fields[0]), n(fields[1]), n(fields[2]), gu.npy(fields[3]), gu.npy(fields[4]), (n(fields[5][0]), n(fields[5][1])), None, False)
    t_ref, e_ref = [], []
    for step in range(121):
        if step % 8 == 0:
            e, b, k = odiag.compute_energy(otp, of[0], of[1], osp, odp, osc)
            t_ref.append(step * float(dp.dt)); e_ref.append(float(e))
        otp, of = oevolve.time_loop_electrodynamic(otp, osc, of, osp, odp)
    np.random.seed(0)
    run_PyPIC3D(cfg, verbose=False)
    rows = [r.split(",") for r in open(os.path.join(str(tmp_path), "data", "electric_field_energy.txt")).read().strip().splitlines()]
    t_gpu = np.array([float(r[0]) for r in rows]); e_gpu = np.array([float(r[1]) for r in rows])
    t_ref, e_ref = np.array(t_ref), np.array(e_ref)
    assert t_gpu.shape == t_ref.shape == (16,) and np.allclose(t_gpu, t_ref, rtol=1e-12, atol=0)
    assert e_gpu[0] == 0.0 and np.allclose(e_gpu[1:], e_ref[1:], rtol=1e-6)
    assert e_ref[-1] > 10 * e_ref[1]                                   # the instability did grow
    sel = slice(2, 13)                                                 # steps 16 .. 96
    rate = lambda t, e: np.polyfit(t[sel], np.log(e[sel]), 1)[0] / 2
    g_gpu, g_ref = rate(t_gpu, e_gpu), rate(t_ref, e_ref)
    wp_total = np.sqrt(2 * 1.0e15 * 1.602e-19 ** 2 / (float(dp.eps) * 9.1093837e-31))
    assert abs(g_gpu - g_ref) <= 1e-5 * abs(g_ref)
    assert 0.25 <= g_ref / (wp_total / (2 * np.sqrt(2))) <= 0.7


@pytest.mark.gpu
@pytest.mark.parametrize("dep", ["j_from_rhov", "esirkepov"])
def test_weibel_growth_rate_and_magnetic_energy_history(tmp_path, dep):
    """The north star's second observable: demos/weibel/weibel.toml as packaged (1-D, 201 cells, TSC, 40 + 40 particles per cell,
    Tz / Tx = 100), 1200 steps, with the packaged current deposition (j_from_rhov + bilinear filter) and with Esirkepov.  The noise
    of the initial load rings as a standing electromagnetic mode (its magnetic energy swings between ~2e-5 and ~1e-7 with a period
    of 200 steps) on top of which the anisotropy-driven field grows; sampled every 200 steps -- at the minima of the ringing -- the
    magnetic energy rises 4x - 6x between steps 200 and 1200 (seed 0; ~20x for other seeds) (a 6000-step oracle run of the packaged demo carries on to 4e-4).  The
    history written by `run_PyPIC3D` (reference file format) equals the oracle's from the identical initial state, and the growth
    rates fitted to both agree."""
    from tests import gpu_util as gu
    gu.require_cuda()
    from pypic3d_b200.initialization import initialize_simulation
    from pypic3d_b200.__main__ import run_PyPIC3D
    from oracle import evolve as oevolve, diagnostics as odiag
    from oracle.params import StaticParameters as OS, DynamicParameters as OD, GridParameters as OG, TiledParticles as OT, SpeciesConfig as OC
    NT, every = 1201, 200
    cfg = {k: dict(v) for k, v in WEIBEL.items()}
    cfg["simulation_parameters"].update(output_dir=str(tmp_path), Nt=NT, current_calculation=dep,
                                        filter_j="bilinear" if dep == "j_from_rhov" else "none")
    cfg["plotting"]["plotting_interval"] = every
    np.random.seed(0)
    loop, particles, fields, sp, dp, plotting, plasma, species = initialize_simulation(cfg, verbose=False)
    osp = OS(**sp._asdict()); odp = OD(**{**dp._asdict(), "grids": OG(**dp.grids._asdict())})
    otp = OT(gu.npy(particles.x), gu.npy(particles.u), gu.npy(particles.active))
    osc = OC(*[np.asarray(v) for v in species])
    n = lambda F: tuple(gu.npy(c) for c in F)
    of = (n(fields[0]), n(fields[1]), n(fields[2]), gu.npy(fields[3]), gu.npy(fields[4]), (n(fields[5][0]), n(fields[5][1])), None, False)
    t_ref, b_ref = [], []
    for step in range(NT):
        if step % every == 0:
            e, b, k = odiag.compute_energy(otp, of[0], of[1], osp, odp, osc)
            t_ref.append(step * float(dp.dt)); b_ref.append(float(b))
        otp, of = oevolve.time_loop_electrodynamic(otp, osc, of, osp, odp)
    np.random.seed(0)
    run_PyPIC3D(cfg, verbose=False)
    rows = [r.split(",") for r in open(os.path.join(str(tmp_path), "data", "magnetic_field_energy.txt")).read().strip().splitlines()]
    t_gpu = np.array([float(r[0]) for r in rows]); b_gpu = np.array([float(r[1]) for r in rows])
    t_ref, b_ref = np.array(t_ref), np.array(b_ref)
    print("weibel", dep, "B energy every 200 steps (oracle):", b_ref.tolist())
    assert t_gpu.shape == t_ref.shape == (7,) and np.allclose(t_gpu, t_ref, rtol=1e-12, atol=0)
    assert b_gpu[0] == 0.0 and np.allclose(b_gpu[1:], b_ref[1:], rtol=1e-4)
    assert b_ref[-1] > 3 * b_ref[1]                                    # the anisotropy did drive the magnetic field up (4.4x - 5.8x)
    sel = slice(1, 7)                                                  # steps 200 .. 1200
    rate = lambda t, e: np.polyfit(t[sel], np.log(e[sel]), 1)[0] / 2
    g_gpu, g_ref = rate(t_gpu, b_gpu), rate(t_ref, b_ref)
    print("weibel", dep, "growth rate", g_ref, g_gpu)
    assert g_ref > 0 and abs(g_gpu - g_ref) <= 1e-3 * abs(g_ref)
    # the order of magnitude the anisotropy allows: w_pe * v_hot / c = 1.3e8 1/s for the packaged temperatures (the fit over the
    # first 1200 steps, start-up ringing included, gives 1.6e8 for this seed)
    wpe = np.sqrt(1.0e15 * 1.602e-19 ** 2 / (float(dp.eps) * 9.1093837e-31))
    vz = np.sqrt(1.380649e-23 * 29649482.87173552 / 9.1093837e-31)
    assert 0.1 < g_ref / (wpe * vz / 299792458.0) < 10


def test_front_end_host_logic_without_a_gpu(tmp_path, monkeypatch, capsys):
    """Everything `initialize_simulation` does on the host for the packaged two-stream configuration -- parameters, particle loading,
    species metadata, plasma parameters, the start-up report -- and the `output.toml` the driver writes from it.  The only device
    operation on that path (the initial guard-cell refresh of zero fields) is replaced by the identity, so this runs on CPU tensors."""
    import toml
    import torch
    import pypic3d_b200.initialization as ini
    from pypic3d_b200.utils import dump_parameters_to_toml
    monkeypatch.setattr(ini, "update_tiled_vector_ghost_cells", lambda v, sp, g: v)
    cfg = {k: dict(v) for k, v in TWO_STREAM.items()}
    cfg["simulation_parameters"].update(output_dir=str(tmp_path))
    np.random.seed(0)
    loop, particles, fields, sp, dp, plotting, plasma, species = ini.initialize_simulation(cfg, device=torch.device("cpu"), verbose=True)
    out = capsys.readouterr().out
    assert "time window:" in out and "x window: 1.0 m with dx: 0.01 m" in out
    assert loop is ini.time_loop_electrodynamic and tuple(particles.x.shape[:4]) == (1, 1, 1, 3)
    assert plotting["particle_species_names"] == ("electron1", "electron2", "ion1") and len(plotting["particle_species_metadata"]) == 3
    e1 = plotting["particle_species_metadata"][0]
    n = e1["weight"] * e1["N_particles"] / 1.0                      # first species: the reference's "electrons" (initialization.py:381-384)
    assert np.isclose(plasma["Theoretical Plasma Frequency"], np.sqrt(n) * abs(e1["charge"]) / np.sqrt(dp.eps * e1["mass"]))
    assert np.isclose(plasma["dx per debye length"], plasma["Debye Length"] / dp.dx) and plasma["Number of Electrons"] == 1500
    os.makedirs(os.path.join(tmp_path, "data"), exist_ok=True)
    dump_parameters_to_toml({"total_time": 1.0, "total_iterations": sp.Nt}, sp, dp, plasma, plotting, particles)
    c = toml.load(os.path.join(tmp_path, "data/output.toml"))
    assert [p_["name"] for p_ in c["particles"]] == ["electron1", "electron2", "ion1"]
    assert [p_["active_particles"] for p_ in c["particles"]] == [1500, 1500, 3000] and c["particles"][0]["storage"] == "tiled"
    assert c["static_parameters"]["current_deposition"] == "direct" and c["dynamic_parameters"]["Nx"] == 100     # j_from_rhov (initialization.py:97-102)
    assert np.isclose(c["plasma_parameters"]["Debye Length"], plasma["Debye Length"])
