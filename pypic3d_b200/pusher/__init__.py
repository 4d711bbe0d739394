from .particle_push import particle_push  # noqa: F401
