"""The drop-in boundary (SURVEY.md section 8 b): the reference's calling convention and parameter pytrees, pinned by its own
tests -- parameter_cutover_test.py:13-24 (exact parameter names of the two loops) and parameters_test.py:17-66 (the
Static/Dynamic split).  Host-side only."""
import inspect

import numpy as np
import pytest

LOOP_PARAMETERS = ["particles", "species_config", "fields", "static_parameters", "dynamic_parameters"]


def test_loops_take_only_the_split_parameter_contract():
    """parameter_cutover_test.py:13-24"""
    from pypic3d_b200.evolve import time_loop_electrodynamic, time_loop_electrostatic
    assert list(inspect.signature(time_loop_electrodynamic).parameters) == LOOP_PARAMETERS
    assert list(inspect.signature(time_loop_electrostatic).parameters) == LOOP_PARAMETERS
    from oracle.evolve import time_loop_electrodynamic as o1, time_loop_electrostatic as o2
    assert list(inspect.signature(o1).parameters) == LOOP_PARAMETERS and list(inspect.signature(o2).parameters) == LOOP_PARAMETERS


def test_sub_entry_points_keep_the_reference_argument_names():
    """The operator-level drop-ins mirror the reference's signatures; literals from
    pusher/particle_push.py:13, deposition/Esirkepov.py:49, J_from_rhov.py:32, rho.py:30, solvers/first_order_yee.py:12,96,
    particles/particle_tile_communication.py:82,440."""
    from pypic3d_b200.pusher.particle_push import particle_push
    from pypic3d_b200.deposition.Esirkepov import Esirkepov_current
    from pypic3d_b200.deposition.J_from_rhov import J_from_rhov
    from pypic3d_b200.deposition.rho import compute_rho
    from pypic3d_b200.solvers.first_order_yee import update_E, update_B
    from pypic3d_b200.particles.particle_tile_communication import update_tiled_particle_positions, refresh_tiled_particle_tiles
    names = lambda f: list(inspect.signature(f).parameters)
    assert names(particle_push) == ["particles", "species_config", "E_tiles", "B_tiles", "static_parameters", "dynamic_parameters"]
    for f in (Esirkepov_current, J_from_rhov):
        assert names(f)[:5] == ["particles", "species_config", "J", "static_parameters", "dynamic_parameters"]
    assert names(compute_rho) == ["particles", "species_config", "rho", "static_parameters", "dynamic_parameters"]
    assert names(update_E) == ["E_tiles", "B_tiles", "J_tiles", "static_parameters", "dynamic_parameters", "pml_state"]
    assert names(update_B) == ["E_tiles", "B_tiles", "static_parameters", "dynamic_parameters", "pml_state", "do_filter"]
    assert names(update_tiled_particle_positions)[:3] == ["tiled_particles", "species_config", "dt"]
    assert names(refresh_tiled_particle_tiles) == ["tiled_particles", "static_parameters", "dynamic_parameters"]
    # diagnostics boundary: diagnostics/output_adapters.py:40,78,154; diagnostics/plotting.py:258
    from pypic3d_b200.diagnostics.output_adapters import assemble_tiled_scalar_field, assemble_tiled_vector_field, particles_for_output
    from pypic3d_b200.diagnostics.plotting import write_data
    for f in (assemble_tiled_scalar_field, assemble_tiled_vector_field):
        assert names(f) == ["field_tiles", "static_parameters", "tile_shape", "num_guard_cells"]
    assert names(particles_for_output) == ["particles", "species_config", "species_names", "static_parameters", "dynamic_parameters"]
    assert names(write_data) == ["filename", "time", "data"]


def test_static_and_dynamic_parameters_split_kernel_contract():
    """parameters_test.py:17-66 with its literal values, on the package's NamedTuples (built through the oracle fixture, which
    uses the same field names)."""
    from oracle import fixtures as fx
    from pypic3d_b200.parameters import StaticParameters, DynamicParameters, GridParameters
    osp, odp = fx.kernel_parameters(dt=0.1, dx=0.25, dy=0.5, dz=1.0, Nx=4, Ny=2, Nz=1, x_wind=1.0, y_wind=1.0, z_wind=1.0, shape_factor=1,
                                    guard_cells=2, tile_shape=(4, 2, 1), current_deposition="direct", current_filter="none",
                                    particle_boundary_conditions=(0, 1, 2), relativistic=False, C=1.0, eps=2.0, mu=3.0, kb=4.0, alpha=0.5)
    sp = StaticParameters(**osp._asdict())
    dp = DynamicParameters(**{**odp._asdict(), "grids": GridParameters(**odp.grids._asdict())})
    assert sp.current_deposition == "direct" and sp.current_filter == "none" and sp.particle_pusher == "boris"
    assert sp.tile_shape == (4, 2, 1) and sp.boundary_conditions == (0, 0, 0) and sp.particle_boundary_conditions == (0, 1, 2)
    assert "particle_species_names" not in sp._asdict() and "particle_species_metadata" not in sp._asdict()
    assert isinstance(hash(sp), int)
    with pytest.raises(TypeError):
        sp["current_deposition"]
    assert isinstance(dp.grids, GridParameters)
    for k in ("current_deposition", "current_filter", "field_mesh"):
        assert k not in dp._asdict()
    assert float(dp.dt) == pytest.approx(0.1) and float(dp.C) == pytest.approx(1.0)
    with pytest.raises(TypeError):
        dp["dt"]
    # same field names, in the same order, as the oracle's restatement of parameters.py:7-53
    assert StaticParameters._fields == type(osp)._fields and DynamicParameters._fields == type(odp)._fields
