"""pypic3d_b200 -- B200-native (sm_100a CUDA) implementation of PyPIC3D's electrodynamic PIC step.

Module paths mirror the reference (`PyPIC3D.evolve`, `.pusher.particle_push`, `.deposition.Esirkepov`, ...), so a
caller switches by changing the package name and handing torch CUDA tensors instead of JAX arrays.
"""
from .parameters import StaticParameters, DynamicParameters, GridParameters  # noqa: F401
from .particles.particle_class import SpeciesConfig, TiledParticles  # noqa: F401

__version__ = "0.1.0"
