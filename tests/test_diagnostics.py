"""The I/O-boundary helpers (pypic3d_b200/diagnostics): the cases of the reference's tests/code_tests/particle_diagnostics_test.py
(:75-160, same literal inputs) and the tile assembly of output_adapters.py:40-83 against the oracle's.  CPU tensors: these helpers
are plain data movement and device-agnostic."""
import os

import numpy as np
import pytest
import torch

from oracle import diagnostics as odiag, fixtures as fx
from pypic3d_b200 import TiledParticles
from pypic3d_b200.diagnostics import (assemble_tiled_scalar_field, assemble_tiled_vector_field, fields_for_output,
                                      particles_for_output, scalar_field_for_output, vector_field_for_output, write_data)


def _params(pbc=(0, 0, 0)):
    return fx.kernel_parameters(Nx=4, Ny=2, Nz=1, x_wind=4.0, y_wind=2.0, z_wind=1.0, dx=1.0, dy=1.0, dz=1.0, dt=0.2,
                                tile_shape=(2, 1, 1), particle_boundary_conditions=pbc)


def _species():
    ions = fx.particle_species("ions", 2.0, 3.0, weight=4.0, x1=[-1.5, -0.5, 0.5, 1.5], x2=[-0.5, -0.5, 0.5, 0.5], x3=[0.0] * 4,
                               u1=[0.1, 0.2, 0.3, 0.4], u2=[1.0, 1.1, 1.2, 1.3], u3=[2.0, 2.1, 2.2, 2.3],
                               active_mask=[True, False, True, True])
    electrons = fx.particle_species("electrons", -1.0, 0.5, weight=8.0, x1=[-1.25, 0.25, 1.25], x2=[0.25, -0.25, 0.25], x3=[0.0] * 3,
                                    u1=[-0.1, -0.2, -0.3], u2=[-1.0, -1.1, -1.2], u3=[-2.0, -2.1, -2.2],
                                    active_mask=[False, True, True])
    return [ions, electrons]


def _tiled(sp, dp):
    tp, sc = fx.build_tiled_particles(_species(), sp, dp)
    return TiledParticles(torch.from_numpy(tp.x), torch.from_numpy(tp.u), torch.from_numpy(tp.active)), sc


def _sorted_rows(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


def test_flatten_matches_the_active_particles_of_every_species():
    sp, dp = _params()
    tp, sc = _tiled(sp, dp)
    flat = particles_for_output(tp, species_config=sc)
    assert len(flat) == 2
    for s, (orig, rec) in enumerate(zip(_species(), flat)):
        act = orig["active"]
        assert rec.species_index == s and rec.name == f"species_{s}"
        assert np.allclose(_sorted_rows(rec.x), _sorted_rows(orig["x"][act]))
        assert np.allclose(_sorted_rows(rec.u), _sorted_rows(orig["u"][act]))
        n = int(act.sum())
        for got, want in ((rec.charge, orig["charge"]), (rec.mass, orig["mass"]), (rec.weight, orig["weight"])):
            assert tuple(got.shape) == (n,) and np.allclose(got, want)


def test_inactive_and_absorbed_slots_are_not_flattened():
    sp, dp = _params()
    tp, sc = _tiled(sp, dp)
    flat = particles_for_output(tp, species_config=sc)
    assert flat[0].x.shape[0] == 3 and flat[1].x.shape[0] == 2
    assert not np.isclose(flat[0].x[:, 0], -0.5).any() and not np.isclose(flat[1].x[:, 0], -1.25).any()
    # absorb the ion at x = 1.5 (tile (1,1,0), species 0): it disappears from the output
    slot = int(torch.nonzero(torch.isclose(tp.x[1, 1, 0, 0, :, 0], torch.tensor(1.5, dtype=tp.x.dtype)) & tp.active[1, 1, 0, 0])[0])
    active = tp.active.clone()
    active[1, 1, 0, 0, slot] = False
    flat = particles_for_output(tp._replace(active=active), species_config=sc)
    assert flat[0].x.shape[0] == 2 and not np.isclose(flat[0].x[:, 0], 1.5).any()


def test_diagnostic_positions_are_half_a_step_back_and_rewrapped():
    sp, dp = _params()
    tp, sc = _tiled(sp, dp)
    flat = particles_for_output(tp, species_config=sc, static_parameters=sp, dynamic_parameters=dp)
    for orig, rec in zip(_species(), flat):
        act = orig["active"]
        want = orig["x"][act] - 0.5 * orig["u"][act] * dp.dt
        for a, wind in enumerate((dp.x_wind, dp.y_wind, dp.z_wind)):     # periodic axes: back inside [-wind/2, wind/2]
            want[:, a] = np.where(want[:, a] > wind / 2, want[:, a] - wind, np.where(want[:, a] < -wind / 2, want[:, a] + wind, want[:, a]))
        assert np.allclose(_sorted_rows(rec.x_diagnostic), _sorted_rows(want))
    # z: u3 * dt / 2 = 0.2 .. 0.23 stays inside +-0.5; force a wrap on x with a fast particle and check the non-periodic branch too
    x = torch.tensor([[[[[[1.9, 0.0, 0.0]]]]]], dtype=torch.float64)
    u = torch.tensor([[[[[[-2.0, 0.0, 0.0]]]]]], dtype=torch.float64)
    one = TiledParticles(x, u, torch.ones((1, 1, 1, 1, 1), dtype=torch.bool))
    sc1 = sc._replace(charge=sc.charge[:1], mass=sc.mass[:1], weight=sc.weight[:1])
    rec = particles_for_output(one, species_config=sc1, static_parameters=sp, dynamic_parameters=dp)[0]
    assert np.isclose(float(rec.x_diagnostic[0, 0]), 1.9 + 0.2 - 4.0)
    sp2, dp2 = _params(pbc=(2, 0, 0))
    rec = particles_for_output(one, species_config=sc1, static_parameters=sp2, dynamic_parameters=dp2)[0]
    assert np.isclose(float(rec.x_diagnostic[0, 0]), 2.1)


def test_species_names_and_type_check():
    sp, dp = _params()
    tp, sc = _tiled(sp, dp)
    flat = particles_for_output(tp, species_config=sc, species_names=["ions", "electrons"])
    assert [r.name for r in flat] == ["ions", "electrons"]
    with pytest.raises(TypeError):
        particles_for_output((tp.x, tp.u, tp.active), species_config=sc)


@pytest.mark.parametrize("tile,g", [((2, 1, 1), 2), ((4, 2, 1), 2), ((2, 2, 1), 1), ((1, 1, 1), 3)])
def test_tile_assembly_matches_the_oracle_and_round_trips(tile, g):
    sp, dp = fx.kernel_parameters(Nx=4, Ny=2, Nz=1, x_wind=4.0, y_wind=2.0, z_wind=1.0, tile_shape=tile, guard_cells=g)
    rng = np.random.default_rng(7)
    glob = rng.normal(size=(4 + 2, 2 + 2, 1 + 2))
    tiles = fx.field_tiles_from_global(glob, sp, dp)
    out = assemble_tiled_scalar_field(torch.from_numpy(tiles), sp, tile, num_guard_cells=g)
    assert np.array_equal(out.numpy(), odiag.assemble_tiled_scalar_field(tiles, tile, g))
    assert np.array_equal(out.numpy()[1:-1, 1:-1, 1:-1], glob[1:-1, 1:-1, 1:-1])        # interiors survive the round trip
    # unrefreshed guards: the later tile's guard layer overwrites the earlier tile's edge cell, as in the reference's loop
    noisy = rng.normal(size=tiles.shape)
    assert np.array_equal(assemble_tiled_scalar_field(torch.from_numpy(noisy), sp, tile, g).numpy(),
                          odiag.assemble_tiled_scalar_field(noisy, tile, g))
    vec = assemble_tiled_vector_field(tuple(torch.from_numpy(tiles * k) for k in (1.0, 2.0, 3.0)), sp, tile, g)
    assert len(vec) == 3 and np.array_equal(vec[2].numpy(), 3.0 * out.numpy())
    assert np.array_equal(scalar_field_for_output(torch.from_numpy(tiles), sp).numpy(), out.numpy())
    assert scalar_field_for_output(out, sp) is out                                       # already global: passed through
    v3 = tuple(torch.from_numpy(tiles) for _ in range(3))
    assert np.array_equal(vector_field_for_output(v3, sp)[1].numpy(), out.numpy())


def test_fields_for_output_drops_the_overflow_flag():
    sp, dp = fx.kernel_parameters(Nx=4, Ny=2, Nz=1, x_wind=4.0, y_wind=2.0, z_wind=1.0, tile_shape=(2, 1, 1))
    t = lambda: torch.from_numpy(fx.empty_tiled_scalar(sp, dp))
    v = lambda: (t(), t(), t())
    out = fields_for_output((v(), v(), v(), t(), t(), (v(), v()), None, torch.tensor(False)), sp)
    assert len(out) == 7 and out[6] is None                                              # pml_state kept, overflow dropped
    assert tuple(out[0][0].shape) == (6, 4, 3) and tuple(out[3].shape) == (6, 4, 3) and tuple(out[5][1][2].shape) == (6, 4, 3)
    assert len(fields_for_output((v(), v(), v(), t(), t(), (v(), v())), sp)) == 6


def test_write_data_row_format(tmp_path):
    f = os.path.join(tmp_path, "total_energy.txt")
    write_data(f, 0, torch.tensor(4.0, dtype=torch.float64))
    write_data(f, 0.25, 1.125)
    assert open(f).read() == "0.0, 4.0\n0.25, 1.125\n"


def test_dump_parameters_to_toml_writes_tiled_species_summaries(tmp_path):
    """utils_test.py:348-421 with its literal inputs."""
    import toml
    from pypic3d_b200.parameters import build_static_parameters, build_dynamic_parameters
    from pypic3d_b200.utils import dump_parameters_to_toml
    active = torch.tensor([[[[[True, False, True], [False, True, False]]]]])
    zeros3 = torch.zeros(tuple(active.shape) + (3,), dtype=torch.float64)
    particles = TiledParticles(x=zeros3, u=zeros3, active=active)
    os.makedirs(os.path.join(tmp_path, "data"))
    sp = build_static_parameters({"name": "dump test", "output_dir": str(tmp_path), "shape_factor": 1, "guard_cells": 2, "tile_shape": (2, 1, 1),
                                  "boundary_conditions": {"x": 0, "y": 0, "z": 0}, "particle_boundary_conditions": {"x": 0, "y": 0, "z": 0},
                                  "field_mesh": object()})
    z = np.zeros
    dp = build_dynamic_parameters({"dt": 0.1, "dx": 1.0, "dy": 1.0, "dz": 1.0, "Nx": 2, "Ny": 1, "Nz": 1, "x_wind": 2.0, "y_wind": 1.0, "z_wind": 1.0,
                                   "grids": {"vertex": (z(4), z(3), z(3)), "center": (z(4), z(3), z(3)),
                                             "tiled_vertex_grid": (z((1, 1, 1, 6)),) * 3, "tiled_center_grid": (z((1, 1, 1, 6)),) * 3}}, {})
    plotting = {"particle_species_names": ("electrons", "ions"),
                "particle_species_metadata": ({"name": "electrons", "charge": -1.0}, {"name": "ions", "charge": 1.0}),
                "plotting_interval": 10}
    dump_parameters_to_toml({"total_time": 1.0}, sp, dp, {}, plotting, particles)
    cfg = toml.load(os.path.join(tmp_path, "data/output.toml"))
    for key in ("particle_species_names", "particle_species_metadata"):
        assert key not in cfg["static_parameters"] and key not in cfg.get("plotting", {})
    assert cfg["plotting"]["plotting_interval"] == 10 and cfg["simulation_stats"]["total_time"] == 1.0
    assert "field_mesh" not in cfg["static_parameters"] and "grids" not in cfg["dynamic_parameters"]
    e, i = cfg["particles"]
    assert (e["name"], e["charge"], e["storage"], e["active_particles"], e["tile_shape"]) == ("electrons", -1.0, "tiled", 2, [2, 1, 1])
    assert (i["name"], i["charge"], i["storage"], i["active_particles"]) == ("ions", 1.0, "tiled", 1)
    assert "date" in cfg["version"] and "torch" in cfg["package_versions"]
    # without metadata: default species names
    dump_parameters_to_toml({"total_time": 1.0}, sp, dp, {}, {}, particles)
    cfg = toml.load(os.path.join(tmp_path, "data/output.toml"))
    assert [p_["name"] for p_ in cfg["particles"]] == ["species_0", "species_1"]


def test_utils_exposes_the_reference_paths_of_the_toml_helpers():
    from pypic3d_b200 import utils, initialization
    assert utils.load_external_fields_from_toml is initialization.load_external_fields_from_toml
    assert utils.update_parameters_from_toml is initialization.update_parameters_from_toml
    with pytest.raises(AttributeError):
        utils.no_such_helper


def test_dump_parameters_to_toml_encodes_the_front_ends_metadata(tmp_path):
    """What `python -m pypic3d_b200` hands over: tuples of per-species dicts with tuple / NumPy / None leaves must survive the TOML
    round trip (the encoder alone would write a tuple of dicts as the list of their keys)."""
    import toml
    from pypic3d_b200.utils import dump_parameters_to_toml
    sp, dp = fx.kernel_parameters(Nx=4, Ny=2, Nz=1, x_wind=4.0, y_wind=2.0, z_wind=1.0, tile_shape=(2, 1, 1))
    sp = sp._replace(output_dir=str(tmp_path))
    os.makedirs(os.path.join(tmp_path, "data"))
    tp, sc = _tiled(sp, dp)
    meta = ({"name": "ions", "N_particles": 4, "N_per_cell": None, "charge": np.float64(2.0), "temperature": np.float32(1.5),
             "update_x": (True, np.bool_(False), True), "x_bc": "periodic"},
            {"name": "electrons", "N_particles": 3, "N_per_cell": 1.5, "charge": -1.0, "update_x": (True, True, True)})
    plotting = {"plotting_interval": np.int64(5), "particle_species_names": ("ions", "electrons"), "particle_species_metadata": meta}
    dump_parameters_to_toml({"total_time": 2.0, "total_iterations": 7}, sp, dp, {"species": meta}, plotting, tp)
    cfg = toml.load(os.path.join(tmp_path, "data/output.toml"))
    assert cfg["plasma_parameters"]["species"][0]["update_x"] == [True, False, True]
    assert "N_per_cell" not in cfg["plasma_parameters"]["species"][0] and cfg["plasma_parameters"]["species"][1]["N_per_cell"] == 1.5
    assert cfg["particles"][0]["active_particles"] == 3 and cfg["particles"][1]["active_particles"] == 2
    assert cfg["particles"][0]["temperature"] == 1.5 and cfg["plotting"]["plotting_interval"] == 5
    assert cfg["static_parameters"]["tile_shape"] == [2, 1, 1] and cfg["dynamic_parameters"]["Nx"] == 4


def test_plasma_parameter_helpers(capsys):
    """utils.py:229-255, 299-423 and utils_test.py:172-185."""
    from pypic3d_b200 import utils
    assert utils.vth_to_T(2.0, 3.0, 4.0) == 3.0 and np.isclose(utils.T_to_vth(3.0, 3.0, 4.0), 2.0)
    sp, dp = fx.kernel_parameters(Nx=4, Ny=2, Nz=1, x_wind=4.0, y_wind=2.0, z_wind=1.0, dx=1.0, dy=1.0, dz=1.0, eps=2.0, kb=0.5)
    e = {"mass": 2.0, "temperature": 8.0, "N_particles": 16, "charge": -1.0, "weight": 0.5}
    pl = utils.build_plasma_parameters_dict(sp, dp, e)
    n = 0.5 * 16 / 8.0
    assert np.isclose(pl["Theoretical Plasma Frequency"], np.sqrt(n) / np.sqrt(2.0 * 2.0))
    assert np.isclose(pl["Debye Length"], np.sqrt(2.0 * 0.5 * 8.0 / n)) and np.isclose(pl["Thermal Velocity"], np.sqrt(3 * 0.5 * 8.0 / 2.0))
    assert pl["Number of Electrons"] == 16 and np.isclose(pl["dx per debye length"], pl["Debye Length"] / 1.0)
    utils.check_stability({"Theoretical Plasma Frequency": 1.0, "Debye Length": 1.0, "Thermal Velocity": 1.0, "Number of Electrons": 100,
                           "dx per debye length": 0.5}, 3.0)
    out = capsys.readouterr().out
    assert "# of Electrons is Low" in out and "Debye Length is less than the spatial resolution" in out and "Number of Electrons: 100" in out
    tp, sc = _tiled(sp._replace(tile_shape=(2, 1, 1)), dp)
    utils.particle_sanity_check(tp)
    with pytest.raises(AssertionError):
        utils.particle_sanity_check(tp._replace(active=tp.active[..., :1]))
    utils.print_stats(sp, dp)
    assert "x window: 4.0 m with dx: 1.0 m" in capsys.readouterr().out
