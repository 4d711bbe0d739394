"""State pytrees with the reference's names (PyPIC3D/particles/particle_class.py:5-31).  Leaves are torch CUDA tensors
(species metadata may be NumPy/host)."""
from typing import NamedTuple


class SpeciesConfig(NamedTuple):
    charge: object    # (species,)
    mass: object      # (species,)
    weight: object    # (species,)
    update_x: object  # (species, 3) bool
    update_u: object  # (species, 3) bool


class TiledParticles(NamedTuple):
    x: object        # (ntx, nty, ntz, species, max_particles_per_tile, 3) positions
    u: object        # (ntx, nty, ntz, species, max_particles_per_tile, 3) velocities v (not gamma*v)
    active: object   # (ntx, nty, ntz, species, max_particles_per_tile) bool
