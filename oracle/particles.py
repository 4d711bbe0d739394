"""Particle move + retile/migration on an in-process tile mesh (test infrastructure).
Follows PyPIC3D/particles/particle_tile_communication.py:41-99,145-230,233-453."""
import numpy as np

from .params import TiledParticles
from .stencil import wrap_periodic_position


def update_tiled_particle_positions(tp, species_config, dt):
    """particle_tile_communication.py:82-99."""
    x = np.array(tp.x, dtype=np.float64, copy=True)
    u = np.asarray(tp.u, dtype=np.float64)
    act = np.asarray(tp.active, dtype=bool)
    af = act.astype(np.float64)
    upx = np.asarray(species_config.update_x, dtype=bool)
    upx = upx.reshape((1, 1, 1, upx.shape[0], 1, 3))
    for c in range(3):
        step = af * u[..., c] * dt
        x[..., c] = np.where(act & upx[..., c], x[..., c] + step, x[..., c])
    return tp._replace(x=x)


def apply_axis_boundary(x, u, active, wind, bc):
    """particle_tile_communication.py:41-59.  bc: 0 periodic wrap, 1 reflect, 2 absorb."""
    h = 0.5 * wind
    if bc == 0:
        return wrap_periodic_position(x, wind), u, active
    if bc == 1:
        rx = np.where(x > h, 2.0 * h - x, np.where(x < -h, -2.0 * h - x, x))
        ru = np.where((x >= h) | (x <= -h), -u, u)
        return rx, ru, active
    if bc == 2:
        return x, u, active & (x <= h) & (x >= -h)
    return x, u, active


def particle_tile_indices(x, y, z, sp, dp, tile_counts):
    """particle_tile_communication.py:62-79."""
    out = []
    for p, wind, d, n, w, nt in ((x, dp.x_wind, dp.dx, dp.Nx, sp.tile_shape[0], tile_counts[0]),
                                 (y, dp.y_wind, dp.dy, dp.Ny, sp.tile_shape[1], tile_counts[1]),
                                 (z, dp.z_wind, dp.dz, dp.Nz, sp.tile_shape[2], tile_counts[2])):
        cell = np.floor((p + 0.5 * wind) / d).astype(np.int64)
        cell = np.clip(cell, 0, int(n) - 1)
        out.append(np.clip(cell // int(w), 0, int(nt) - 1))
    return tuple(out)


def adjacent_tile_offset(dest, source, tile_count):
    """particle_tile_communication.py:145-165."""
    dest = np.asarray(dest)
    if tile_count == 1:
        return np.zeros_like(dest)
    off = dest - source
    if tile_count == 2:
        return off
    off = np.where(off == tile_count - 1, -1, off)
    off = np.where(off == -(tile_count - 1), 1, off)
    return off


def _movement_offsets(count):
    return (0,) if count == 1 else (1, 0, -1)        # :102-105


def _neighbor_source(t, off, n, bc):
    """Source tile whose stream with offset `off` lands on tile t (ppermute perms :108-135); None if no sender."""
    if n == 1 or off == 0:
        return t
    s = t - off
    if bc == 0:
        return s % n
    return s if 0 <= s < n else None


def refresh_tiled_particle_tiles(tp, sp, dp):
    """particle_tile_communication.py:292-453.  Returns (TiledParticles, overflow: bool)."""
    x = np.asarray(tp.x, dtype=np.float64)
    u = np.asarray(tp.u, dtype=np.float64)
    act = np.asarray(tp.active, dtype=bool)
    mesh = x.shape[:3]
    S, cap = act.shape[3], act.shape[4]
    pbc = tuple(int(b) for b in sp.particle_boundary_conditions)
    winds = (dp.x_wind, dp.y_wind, dp.z_wind)
    bx = x.copy(); bu = u.copy(); bact = act.copy()
    for c in range(3):                                             # :233-272
        bx[..., c], bu[..., c], bact = apply_axis_boundary(bx[..., c], bu[..., c], bact, winds[c], pbc[c])
    dest = particle_tile_indices(bx[..., 0], bx[..., 1], bx[..., 2], sp, dp, mesh)
    tidx = np.meshgrid(np.arange(mesh[0]), np.arange(mesh[1]), np.arange(mesh[2]), indexing="ij")
    offs = [adjacent_tile_offset(dest[c], tidx[c][..., None, None], mesh[c]) for c in range(3)]
    nonlocal_ = (offs[0] != 0) | (offs[1] != 0) | (offs[2] != 0)
    invalid = (np.abs(offs[0]) > 1) | (np.abs(offs[1]) > 1) | (np.abs(offs[2]) > 1)
    moving = bact & nonlocal_ & ~invalid
    stay = bact & ~moving & ~invalid
    new_x = np.where(stay[..., None], bx, 0.0)
    new_u = np.where(stay[..., None], bu, 0.0)
    new_act = stay.copy()
    overflow = bool(np.any(bact & invalid))
    streams = [(ox, oy, oz) for ox in _movement_offsets(mesh[0]) for oy in _movement_offsets(mesh[1])
               for oz in _movement_offsets(mesh[2]) if not (ox == 0 and oy == 0 and oz == 0)]
    if not streams:
        return TiledParticles(new_x, new_u, new_act), overflow
    for tx in range(mesh[0]):
        for ty in range(mesh[1]):
            for tz in range(mesh[2]):
                for s in range(S):
                    inc_x, inc_u = [], []
                    for (ox, oy, oz) in streams:                    # stream order :326-365, slot order inside
                        sx = _neighbor_source(tx, ox, mesh[0], pbc[0])
                        sy = _neighbor_source(ty, oy, mesh[1], pbc[1])
                        sz = _neighbor_source(tz, oz, mesh[2], pbc[2])
                        if sx is None or sy is None or sz is None:
                            continue
                        m = (moving[sx, sy, sz, s] & (offs[0][sx, sy, sz, s] == ox) & (offs[1][sx, sy, sz, s] == oy)
                             & (offs[2][sx, sy, sz, s] == oz))
                        inc_x.append(bx[sx, sy, sz, s][m]); inc_u.append(bu[sx, sy, sz, s][m])
                    inc_x = np.concatenate(inc_x, axis=0) if inc_x else np.zeros((0, 3))
                    inc_u = np.concatenate(inc_u, axis=0) if inc_u else np.zeros((0, 3))
                    free = np.flatnonzero(~stay[tx, ty, tz, s])      # :169-230 k-th incoming -> k-th free slot
                    nfit = min(len(free), inc_x.shape[0])
                    if inc_x.shape[0] > len(free):
                        overflow = True
                    sl = free[:nfit]
                    new_x[tx, ty, tz, s, sl] = inc_x[:nfit]
                    new_u[tx, ty, tz, s, sl] = inc_u[:nfit]
                    new_act[tx, ty, tz, s, sl] = True
    return TiledParticles(new_x, new_u, new_act), overflow
