"""Drop-in for PyPIC3D/particles/particle_tile_communication.py:82-99 and :440-453."""
from .. import ops
from .particle_class import TiledParticles


def update_tiled_particle_positions(tiled_particles, species_config, dt, static_parameters=None, dynamic_parameters=None):
    """x += active * update_x * u * dt (reference signature: (tiled_particles, species_config, dt))."""
    p = ops._lib.PicParams()
    x = tiled_particles.x
    p.dtype = 0 if x.dtype == ops.F32 else 1
    for a in range(3):
        p.mesh[a] = p.gmesh[a] = int(x.shape[a])
    S = int(x.shape[3])
    p.n_species = S
    ux = ops._lib._to_numpy(species_config.update_x).astype(bool).reshape(S, 3)
    for s in range(S):
        for c in range(3):
            p.update_x[s][c] = int(ux[s, c])
    return tiled_particles._replace(x=ops.move(p, x, tiled_particles.u, tiled_particles.active, float(ops._lib._scalar(dt))))


def refresh_tiled_particle_tiles(tiled_particles, static_parameters, dynamic_parameters):
    """Global particle BCs + re-ownership + neighbour migration with fixed slot capacity; returns (particles, overflow)."""
    x = tiled_particles.x
    S = int(x.shape[3])

    class _SC:  # retile only needs the species count
        charge = [0.0] * S; mass = [1.0] * S; weight = [1.0] * S
        update_x = [[True] * 3] * S; update_u = [[True] * 3] * S
    p = ops.params_for(static_parameters, dynamic_parameters, _SC, x)
    if tuple(x.shape[:3]) != tuple(p.mesh):
        raise ValueError("Tiled particle communication requires one logical particle tile per device: "
                         f"particle tile topology {tuple(x.shape[:3])} does not match device mesh {tuple(p.mesh)}.")
    xo, uo, ao, overflow = ops.retile(p, x, tiled_particles.u, tiled_particles.active)
    return TiledParticles(x=xo, u=uo, active=ao), overflow
