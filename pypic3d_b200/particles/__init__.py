from .particle_class import SpeciesConfig, TiledParticles  # noqa: F401
