"""Field gather + velocity push (test infrastructure).
Follows PyPIC3D/pusher/particle_push.py:13-175, pusher/boris.py:15-258, pusher/higuera_cary.py:6-113."""
import numpy as np

from .params import TiledParticles
from .stencil import BC_PERIODIC, prepare_particle_axis_stencil, axis_has_active_cells, inactive_axis_index
from .shapes import weights as shape_weights


def interpolate_field_to_particles(field, x, y, z, grid, shape_factor, ghost_cells=False,
                                   active_axes=None, inactive_axis_indices=None):
    """boris.py:128-258.  3-point-per-axis gather with CIC/TSC weights; inactive axes collapse."""
    n = (len(grid[0]), len(grid[1]), len(grid[2]))
    if active_axes is None:
        active_axes = tuple(axis_has_active_cells(k, ghost_cells) for k in n)
    if inactive_axis_indices is None:
        inactive_axis_indices = (None, None, None)
    pts, wts = [], []
    for a, p in enumerate((x, y, z)):
        _, _, delta, points = prepare_particle_axis_stencil(p, grid[a], n[a], shape_factor, BC_PERIODIC,
                                                            ghost_cells=ghost_cells)
        d = grid[a][1] - grid[a][0] if n[a] > 1 else 1.0          # boris.py:173-175
        w = np.stack(shape_weights(delta, d, shape_factor), axis=0)
        if not active_axes[a]:                                     # boris.py:192-234
            idx = inactive_axis_indices[a]
            if idx is None:
                idx = inactive_axis_index(n[a], ghost_cells)
            points = np.full((1, points.shape[1]), int(idx), dtype=points.dtype)
            w = np.sum(w, axis=0, keepdims=True)
        pts.append(points)
        wts.append(w)
    out = np.zeros(np.shape(x), dtype=np.float64)
    for i in range(pts[0].shape[0]):
        for j in range(pts[1].shape[0]):
            for k in range(pts[2].shape[0]):
                out = out + field[pts[0][i], pts[1][j], pts[2][k]] * wts[0][i] * wts[1][j] * wts[2][k]
    return out


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def boris(v, E, B, q, m, dt, C):
    """boris.py:41-55 (non-relativistic)."""
    h = q * dt / (2 * m)
    vm = tuple(v[c] + h * E[c] for c in range(3))
    t = tuple(h * B[c] for c in range(3))
    cr = _cross(vm, t)
    vp = tuple(vm[c] + cr[c] for c in range(3))
    den = 1 + t[0] ** 2 + t[1] ** 2 + t[2] ** 2
    s = tuple(2 * t[c] / den for c in range(3))
    cr = _cross(vp, s)
    vplus = tuple(vm[c] + cr[c] for c in range(3))
    return tuple(vplus[c] + h * E[c] for c in range(3))


def relativistic_boris(v, E, B, q, m, dt, C):
    """boris.py:96-121.  State is v; converted v -> u = gamma v -> v every step."""
    h = q * dt / (2 * m)
    gamma = 1 / np.sqrt(1 - ((v[0] ** 2 + v[1] ** 2 + v[2] ** 2) / C ** 2))
    um = tuple(v[c] * gamma + h * E[c] for c in range(3))
    gm = np.sqrt(1 + ((um[0] ** 2 + um[1] ** 2 + um[2] ** 2) / C ** 2))
    t = tuple(h * B[c] / gm for c in range(3))
    cr = _cross(um, t)
    up = tuple(um[c] + cr[c] for c in range(3))
    den = 1 + t[0] ** 2 + t[1] ** 2 + t[2] ** 2
    s = tuple(2 * t[c] / den for c in range(3))
    cr = _cross(up, s)
    uplus = tuple(um[c] + cr[c] for c in range(3))
    nu = tuple(uplus[c] + h * E[c] for c in range(3))
    ng = np.sqrt(1 + ((nu[0] ** 2 + nu[1] ** 2 + nu[2] ** 2) / C ** 2))
    return tuple(nu[c] / ng for c in range(3))


def higuera_cary(v, E, B, q, m, dt, C):
    """higuera_cary.py:58-113."""
    gamma = 1.0 / np.sqrt(1.0 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / C ** 2)
    u = tuple(gamma * v[c] for c in range(3))
    h = q * dt / (2.0 * m)
    eps = tuple(h * E[c] for c in range(3))
    beta = tuple(h * B[c] for c in range(3))
    ue = tuple(u[c] + eps[c] for c in range(3))
    beta2 = beta[0] * beta[0] + beta[1] * beta[1] + beta[2] * beta[2]
    ustar = (ue[0] * beta[0] + ue[1] * beta[1] + ue[2] * beta[2]) / C
    g_ue = np.sqrt(1.0 + (ue[0] * ue[0] + ue[1] * ue[1] + ue[2] * ue[2]) / C ** 2)
    sigma = g_ue ** 2 - beta2
    gnext = np.sqrt((sigma + np.sqrt(sigma ** 2 + 4.0 * (beta2 + ustar ** 2))) / 2.0)
    t = tuple(beta[c] / gnext for c in range(3))
    s = 1.0 / (1.0 + (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]))
    uet = ue[0] * t[0] + ue[1] * t[1] + ue[2] * t[2]
    cr = _cross(ue, t)
    um = tuple(s * (ue[c] + uet * t[c] + cr[c]) for c in range(3))
    cr = _cross(um, t)
    nu = tuple(um[c] + eps[c] + cr[c] for c in range(3))
    ng = np.sqrt(1.0 + (nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]) / C ** 2)
    return tuple(nu[c] / ng for c in range(3))


def component_grids(cx, cy, cz, vx, vy, vz):
    """particle_push.py:63-68: Ex(v,c,c) Ey(c,v,c) Ez(c,c,v) Bx(c,v,v) By(v,c,v) Bz(v,v,c)."""
    return ((vx, cy, cz), (cx, vy, cz), (cx, cy, vz), (cx, vy, vz), (vx, cy, vz), (vx, vy, cz))


def particle_push(particles, species_config, E_tiles, B_tiles, sp, dp):
    """particle_push.py:13-175."""
    tile_shape = tuple(int(w) for w in sp.tile_shape)
    g = int(sp.guard_cells)
    ntx, nty, ntz = particles.x.shape[:3]
    active_axes = (ntx * tile_shape[0] > 1, nty * tile_shape[1] > 1, ntz * tile_shape[2] > 1)   # :38-42
    tc, tv = dp.grids.tiled_center_grid, dp.grids.tiled_vertex_grid
    charge = np.asarray(species_config.charge, dtype=np.float64)
    mass = np.asarray(species_config.mass, dtype=np.float64)
    upd = np.asarray(species_config.update_u, dtype=bool)
    new_u = np.array(particles.u, dtype=np.float64, copy=True)
    if sp.particle_pusher == "boris":
        pusher = relativistic_boris if sp.relativistic else boris
    elif sp.particle_pusher == "higuera_cary":
        pusher = higuera_cary
    else:
        raise ValueError(f"Unknown particle_pusher: {sp.particle_pusher}")
    for tx in range(ntx):
        for ty in range(nty):
            for tz in range(ntz):
                xt = np.asarray(particles.x[tx, ty, tz], dtype=np.float64)
                ut = np.asarray(particles.u[tx, ty, tz], dtype=np.float64)
                at = np.asarray(particles.active[tx, ty, tz], dtype=bool)
                x, y, z = (xt[..., c].reshape(-1) for c in range(3))
                v = tuple(ut[..., c].reshape(-1) for c in range(3))
                q = np.broadcast_to(charge[:, None], at.shape).reshape(-1)
                m = np.broadcast_to(mass[:, None], at.shape).reshape(-1)
                grids = component_grids(tc[0][tx], tc[1][ty], tc[2][tz], tv[0][tx], tv[1][ty], tv[2][tz])
                comps = (E_tiles[0], E_tiles[1], E_tiles[2], B_tiles[0], B_tiles[1], B_tiles[2])
                F = [interpolate_field_to_particles(np.asarray(comps[c][tx, ty, tz], dtype=np.float64), x, y, z,
                                                    grids[c], sp.shape_factor, ghost_cells=True,
                                                    active_axes=active_axes, inactive_axis_indices=(g, g, g))
                     for c in range(6)]
                with np.errstate(all="ignore"):
                    nv = pusher(v, F[0:3], F[3:6], q, m, dp.dt, dp.C)
                act = at.reshape(-1)
                for c in range(3):                                         # :134-142
                    mask = act & np.broadcast_to(upd[:, c, None], at.shape).reshape(-1)
                    new_u[tx, ty, tz, ..., c] = np.where(mask, nv[c], v[c]).reshape(at.shape)
    return TiledParticles(x=particles.x, u=new_u, active=particles.active)
