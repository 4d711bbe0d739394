"""Drop-in for PyPIC3D/evolve.py:16 time_loop_electrodynamic (and :106 time_loop_electrostatic) -- same signature, same `fields` 8-tuple, same order of
operations, torch CUDA tensors as leaves.  This composition mirrors the reference call-for-call (slot identity kept);
`pypic3d_b200.simulation.Simulation` is the resident, fused throughput path for long runs."""
from .deposition.Esirkepov import Esirkepov_current
from .deposition.J_from_rhov import J_from_rhov
from .particles.particle_tile_communication import refresh_tiled_particle_tiles, update_tiled_particle_positions
from .pusher.particle_push import particle_push
from .solvers.electrostatic_yee import calculate_tiled_electrostatic_fields
from .solvers.first_order_yee import update_B, update_E
from .utils import add_external_fields

__all__ = ["time_loop_electrodynamic", "time_loop_electrostatic"]


def time_loop_electrodynamic(particles, species_config, fields, static_parameters, dynamic_parameters):
    """Advance a tiled electrodynamic PIC system by one time step."""
    E, B, J, rho, phi, external_fields, pml_state, overflow_previous = fields
    dt = dynamic_parameters.dt
    push_E, push_B = add_external_fields(E, B, external_fields)
    particles = particle_push(particles, species_config, push_E, push_B, static_parameters, dynamic_parameters)

    if static_parameters.current_deposition == "esirkepov":
        J = Esirkepov_current(particles, species_config, J, static_parameters, dynamic_parameters)
        particles = update_tiled_particle_positions(particles, species_config, dt)
        particles, overflow = refresh_tiled_particle_tiles(particles, static_parameters, dynamic_parameters)
        overflow = overflow_previous | overflow
    else:
        particles = update_tiled_particle_positions(particles, species_config, dt / 2)
        particles, overflow = refresh_tiled_particle_tiles(particles, static_parameters, dynamic_parameters)
        overflow = overflow_previous | overflow
        J = J_from_rhov(particles, species_config, J, static_parameters, dynamic_parameters)
        particles = update_tiled_particle_positions(particles, species_config, dt / 2)
        particles, overflow = refresh_tiled_particle_tiles(particles, static_parameters, dynamic_parameters)
        overflow = overflow_previous | overflow

    B, pml_state = update_B(E, B, static_parameters, dynamic_parameters, pml_state, do_filter=False)
    E, pml_state = update_E(E, B, J, static_parameters, dynamic_parameters, pml_state)
    B, pml_state = update_B(E, B, static_parameters, dynamic_parameters, pml_state, do_filter=True)
    return particles, (E, B, J, rho, phi, external_fields, pml_state, overflow)


def time_loop_electrostatic(particles, species_config, fields, static_parameters, dynamic_parameters):
    """Advance a tiled electrostatic PIC system by one time step (evolve.py:106-161): push -> move -> retile -> rho -> CG
    Poisson solve -> E = -grad(phi).  B, J, external fields and pml_state pass through."""
    E, B, J, rho, phi, external_fields, pml_state, overflow_previous = fields
    dt = dynamic_parameters.dt
    push_E, push_B = add_external_fields(E, B, external_fields)
    particles = particle_push(particles, species_config, push_E, push_B, static_parameters, dynamic_parameters)
    particles = update_tiled_particle_positions(particles, species_config, dt)
    particles, overflow = refresh_tiled_particle_tiles(particles, static_parameters, dynamic_parameters)
    overflow = overflow_previous | overflow
    E, phi, rho = calculate_tiled_electrostatic_fields(static_parameters, dynamic_parameters, particles, species_config, rho, phi)
    return particles, (E, B, J, rho, phi, external_fields, pml_state, overflow)
