"""Drop-ins for PyPIC3D/diagnostics/output_adapters.py: the boundary between the tile-major runtime state and the ordinary
global arrays that tests, energy histories and file writers look at.

    assemble_tiled_scalar_field / assemble_tiled_vector_field    output_adapters.py:40-83
    scalar_field_for_output / vector_field_for_output / fields_for_output    :86-151
    particles_for_output (+ ParticleOutputRecord)                :154-228

These run at `plotting_interval`, not on the step, and are pure data movement: torch slicing on whatever device the tensors live on
(a multi-GPU run gathers its tiles first, `distributed.gather_tiles`)."""
from typing import NamedTuple

import torch

from ..particles.particle_class import TiledParticles


class ParticleOutputRecord(NamedTuple):
    name: str
    species_index: int
    x: object
    x_diagnostic: object
    u: object
    charge: object
    mass: object
    weight: object


def _is_tiled_scalar(field):
    return getattr(field, "ndim", 0) == 6


def _is_tiled_vector(field):
    return isinstance(field, (list, tuple)) and len(field) == 3 and _is_tiled_scalar(field[0])


def assemble_tiled_scalar_field(field_tiles, static_parameters, tile_shape, num_guard_cells=2):
    """Tiles (ntx, nty, ntz, Lx, Ly, Lz) -> one global array with ONE ghost layer, (Nx+2, Ny+2, Nz+2).  Every tile contributes its
    interior plus one guard layer and tiles are written in (tx, ty, tz) order, so where two tiles overlap the later one wins --
    the reference's loop, which only matters for fields whose guards have not been refreshed."""
    field_tiles = torch.as_tensor(field_tiles)
    w = [int(v) for v in tile_shape]
    g = int(num_guard_cells)
    nt = [int(v) for v in field_tiles.shape[:3]]
    out = torch.zeros((nt[0] * w[0] + 2, nt[1] * w[1] + 2, nt[2] * w[2] + 2), dtype=field_tiles.dtype, device=field_tiles.device)
    for tx in range(nt[0]):
        for ty in range(nt[1]):
            for tz in range(nt[2]):
                out[tx * w[0]:tx * w[0] + w[0] + 2, ty * w[1]:ty * w[1] + w[1] + 2, tz * w[2]:tz * w[2] + w[2] + 2] = \
                    field_tiles[tx, ty, tz, g - 1:g + w[0] + 1, g - 1:g + w[1] + 1, g - 1:g + w[2] + 1]
    return out


def assemble_tiled_vector_field(field_tiles, static_parameters, tile_shape, num_guard_cells=2):
    return tuple(assemble_tiled_scalar_field(c, static_parameters, tile_shape, num_guard_cells) for c in field_tiles)


def scalar_field_for_output(field, static_parameters):
    if not _is_tiled_scalar(field):
        return field
    return assemble_tiled_scalar_field(field, static_parameters, tuple(int(v) for v in static_parameters.tile_shape),
                                       num_guard_cells=int(static_parameters.guard_cells))


def vector_field_for_output(field, static_parameters):
    if not _is_tiled_vector(field):
        return field
    return assemble_tiled_vector_field(field, static_parameters, tuple(int(v) for v in static_parameters.tile_shape),
                                       num_guard_cells=int(static_parameters.guard_cells))


def fields_for_output(fields, static_parameters):
    """(E, B, J, rho, phi, (ext_E, ext_B)[, pml_state]) as global arrays; the overflow flag is a driver diagnostic and is dropped
    (output_adapters.py:120-151)."""
    E, B, J, rho, phi, external_fields, *rest = fields
    external_E, external_B = external_fields
    out = (vector_field_for_output(E, static_parameters), vector_field_for_output(B, static_parameters),
           vector_field_for_output(J, static_parameters), scalar_field_for_output(rho, static_parameters),
           scalar_field_for_output(phi, static_parameters),
           (vector_field_for_output(external_E, static_parameters), vector_field_for_output(external_B, static_parameters)))
    return out if not rest else out + (rest[0],)


def _axis_diagnostic_position(x, u, dt, wind, bc):
    # positions are stored at integer steps after the move; diagnostics want them half a step back, re-wrapped on periodic axes
    xd = x - u * dt / 2
    if int(bc) == 0:
        h = wind / 2
        xd = torch.where(xd > h, xd - wind, torch.where(xd < -h, xd + wind, xd))
    return xd


def _scalar(v):
    return float(v.item()) if hasattr(v, "item") else float(v)


def particles_for_output(particles, species_config=None, species_names=None, static_parameters=None, dynamic_parameters=None):
    """Flatten the fixed-capacity tiled storage into one record per species holding only the active particles, tile-major in slot
    order.  With both parameter sets given, `x_diagnostic` is the half-step-back position (output_adapters.py:154-228)."""
    if not isinstance(particles, TiledParticles):
        raise TypeError("Particle output requires TiledParticles runtime storage.")
    half_step = static_parameters is not None and dynamic_parameters is not None
    x_all, u_all, a_all = torch.as_tensor(particles.x), torch.as_tensor(particles.u), torch.as_tensor(particles.active)
    records = []
    for s in range(int(a_all.shape[3])):
        active = a_all[:, :, :, s, :].reshape(-1).to(torch.bool)
        x = x_all[:, :, :, s, :, :].reshape(-1, 3)[active]
        u = u_all[:, :, :, s, :, :].reshape(-1, 3)[active]
        n = int(x.shape[0])
        full = lambda v: torch.full((n,), _scalar(v[s]), dtype=x.dtype, device=x.device)
        xd = x
        if half_step:
            dp, bc = dynamic_parameters, static_parameters.particle_boundary_conditions
            dt = _scalar(dp.dt)
            xd = torch.stack([_axis_diagnostic_position(x[:, a], u[:, a], dt, _scalar(wind), _scalar(bc[a]))
                              for a, wind in enumerate((dp.x_wind, dp.y_wind, dp.z_wind))], dim=-1)
        records.append(ParticleOutputRecord(name=f"species_{s}" if species_names is None else species_names[s], species_index=s,
                                            x=x, x_diagnostic=xd, u=u, charge=full(species_config.charge),
                                            mass=full(species_config.mass), weight=full(species_config.weight)))
    return records
