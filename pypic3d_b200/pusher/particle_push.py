"""Drop-in for PyPIC3D/pusher/particle_push.py:13 (gather + Boris / relativistic Boris / Higuera-Cary)."""
from .. import ops
from ..particles.particle_class import TiledParticles


def particle_push(particles, species_config, E_tiles, B_tiles, static_parameters, dynamic_parameters):
    p = ops.params_for(static_parameters, dynamic_parameters, species_config, particles.x)
    new_u = ops.push(p, particles.x, particles.u, particles.active, tuple(E_tiles), tuple(B_tiles))
    return TiledParticles(x=particles.x, u=new_u, active=particles.active)
