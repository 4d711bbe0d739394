"""One electrodynamic PIC step (test infrastructure).  Follows PyPIC3D/evolve.py:16-103."""
import numpy as np

from .pusher import particle_push
from .deposition import Esirkepov_current, J_from_rhov
from .particles import update_tiled_particle_positions, refresh_tiled_particle_tiles
from .yee import update_E, update_B


def add_external_fields(E, B, external_fields):
    ext_E, ext_B = external_fields                       # utils.py:205-216
    return (tuple(np.asarray(e) + np.asarray(x) for e, x in zip(E, ext_E)),
            tuple(np.asarray(b) + np.asarray(x) for b, x in zip(B, ext_B)))


def time_loop_electrodynamic(particles, species_config, fields, static_parameters, dynamic_parameters):
    E, B, J, rho, phi, external_fields, pml_state, overflow_previous = fields          # :27
    sp, dp = static_parameters, dynamic_parameters
    dt = dp.dt
    push_E, push_B = add_external_fields(E, B, external_fields)                       # :33
    particles = particle_push(particles, species_config, push_E, push_B, sp, dp)      # :36
    if sp.current_deposition == "esirkepov":                                          # :82
        J = Esirkepov_current(particles, species_config, J, sp, dp)
        particles = update_tiled_particle_positions(particles, species_config, dt)
        particles, overflow = refresh_tiled_particle_tiles(particles, sp, dp)
        overflow = bool(overflow_previous) | overflow
    else:                                                                             # :46-67
        particles = update_tiled_particle_positions(particles, species_config, dt / 2)
        particles, overflow = refresh_tiled_particle_tiles(particles, sp, dp)
        overflow = bool(overflow_previous) | overflow
        J = J_from_rhov(particles, species_config, J, sp, dp)
        particles = update_tiled_particle_positions(particles, species_config, dt / 2)
        particles, overflow = refresh_tiled_particle_tiles(particles, sp, dp)
        overflow = bool(overflow_previous) | overflow                                 # (sic: 2nd overwrite, :66)
    B = update_B(E, B, sp, dp, do_filter=False)                                       # :88
    E = update_E(E, B, J, sp, dp)                                                     # :92
    B = update_B(E, B, sp, dp, do_filter=True)                                        # :96
    return particles, (E, B, J, rho, phi, external_fields, pml_state, overflow)


def time_loop_electrostatic(particles, species_config, fields, static_parameters, dynamic_parameters):
    """PyPIC3D/evolve.py:106-161."""
    from .electrostatic import calculate_tiled_electrostatic_fields
    E, B, J, rho, phi, external_fields, pml_state, overflow_previous = fields                          # :122
    sp, dp = static_parameters, dynamic_parameters
    push_E, push_B = add_external_fields(E, B, external_fields)                                       # :128
    particles = particle_push(particles, species_config, push_E, push_B, sp, dp)                      # :131
    particles = update_tiled_particle_positions(particles, species_config, dp.dt)                     # :141
    particles, overflow = refresh_tiled_particle_tiles(particles, sp, dp)                             # :144
    overflow = bool(overflow_previous) | overflow
    E, phi, rho = calculate_tiled_electrostatic_fields(sp, dp, particles, species_config, rho, phi)   # :148
    return particles, (E, B, J, rho, phi, external_fields, pml_state, overflow)
