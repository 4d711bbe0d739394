"""Current / charge deposition (test infrastructure).
Follows PyPIC3D/deposition/Esirkepov.py:17-504, deposition/J_from_rhov.py:23-260, deposition/rho.py:22-196."""
import numpy as np

from .stencil import compute_particle_anchor, particle_axis_offset, prepare_particle_axis_stencil, collapse_axis_stencil
from .shapes import weights as shape_weights
from .halo import fold_tiled_vector_ghost_cells, update_tiled_vector_ghost_cells, fold_tiled_ghost_cells, update_tiled_ghost_cells
from .filters import bilinear_filter_vector, digital_filter_vector, digital_filter


def _scatter_add(tile, ix, iy, iz, values):
    """`tile.at[ix,iy,iz].add(values, mode="drop")` (Esirkepov.py:238): negative indices are first
    normalised Python-style (JAX), then anything still out of range is dropped."""
    shape = tile.shape
    ix = np.where(ix < 0, ix + shape[0], ix)
    iy = np.where(iy < 0, iy + shape[1], iy)
    iz = np.where(iz < 0, iz + shape[2], iz)
    ok = (ix >= 0) & (ix < shape[0]) & (iy >= 0) & (iy < shape[1]) & (iz >= 0) & (iz < shape[2])
    np.add.at(tile, (ix[ok], iy[ok], iz[ok]), values[ok])


def shift_old_stencil(old_w_list, shift):
    """Esirkepov.py:17-25: per-particle cyclic roll of the 5 old weights by -shift."""
    old_w = np.stack(old_w_list, axis=0)                  # (5, N)
    idx = (np.arange(5)[:, None] + shift[None, :]) % 5    # rolled[i] = w[(i + s) mod 5]
    rolled = np.take_along_axis(old_w, idx, axis=0)
    return [rolled[i] for i in range(5)]


def _collapse_redundant_axis(points, cur, old, axis_active, extended):
    # Esirkepov.py:28-45
    if axis_active:
        return points, cur, old
    zero = np.zeros_like(cur[0])
    ct = cur[0] + cur[1] + cur[2] + cur[3] + cur[4]
    ot = old[0] + old[1] + old[2] + old[3] + old[4]
    return np.full(points.shape, extended // 2, dtype=points.dtype), [zero, zero, ct, zero, zero], [zero, zero, ot, zero, zero]


def esirkepov_deposit_tile(x_tile, u_tile, active_tile, species_config, grids_xyz, sp, dp, local_shape, active_axes):
    """One tile of Esirkepov.py:105-331 (no fold)."""
    dx, dy, dz, dt = dp.dx, dp.dy, dp.dz, dp.dt
    sf = int(sp.shape_factor)
    xt = np.asarray(x_tile, dtype=np.float64); ut = np.asarray(u_tile, dtype=np.float64)
    at = np.asarray(active_tile, dtype=bool)
    old = [xt[..., c].reshape(-1) for c in range(3)]
    v = [ut[..., c].reshape(-1) for c in range(3)]
    active = at.reshape(-1).astype(np.float64)
    qw = np.asarray(species_config.charge, dtype=np.float64) * np.asarray(species_config.weight, dtype=np.float64)
    q = np.broadcast_to(qw[:, None], at.shape).reshape(-1)
    n = old[0].shape[0]
    upx = np.asarray(species_config.update_x, dtype=bool)
    new = [old[c] + np.where(np.broadcast_to(upx[:, c, None], at.shape).reshape(-1), v[c] * dt, 0.0) for c in range(3)]  # :125-127
    d = (dx, dy, dz)
    pts, W, OW = [], [], []
    offsets = np.array([-2, -1, 0, 1, 2])
    for c in range(3):
        a_new = compute_particle_anchor(new[c], grids_xyz[c], sf)
        a_old = compute_particle_anchor(old[c], grids_xyz[c], sf)
        del_new = particle_axis_offset(new[c], a_new, grids_xyz[c])
        del_old = particle_axis_offset(old[c], a_old, grids_xyz[c])
        shift = a_new - a_old
        p = a_new[None, :] + offsets[:, None]
        tmp = np.zeros(n)
        w3 = shape_weights(del_new, d[c], sf)
        o3 = shape_weights(del_old, d[c], sf)
        w5 = [tmp, w3[0], w3[1], w3[2], tmp]
        o5 = shift_old_stencil([tmp, o3[0], o3[1], o3[2], tmp], shift)
        p, w5, o5 = _collapse_redundant_axis(p, w5, o5, active_axes[c], local_shape[c])
        pts.append(p); W.append(w5); OW.append(o5)
    xa, ya, za = active_axes
    dJ = [active * (-(q / (dy * dz)) / dt) if xa else active * q * v[0] / (dx * dy * dz),      # :197-214
          active * (-(q / (dx * dz)) / dt) if ya else active * q * v[1] / (dx * dy * dz),
          active * (-(q / (dx * dy)) / dt) if za else active * q * v[2] / (dx * dy * dz)]
    J = [np.zeros(local_shape), np.zeros(local_shape), np.zeros(local_shape)]
    xw, yw, zw = W
    ox, oy, oz = OW
    xp, yp, zp = pts
    nact = int(xa) + int(ya) + int(za)
    if nact == 3:
        Wx = np.zeros((5, 5, 5, n)); Wy = np.zeros_like(Wx); Wz = np.zeros_like(Wx)
        for i in range(5):
            for j in range(5):
                for k in range(5):                                   # :381-406
                    Wx[i, j, k] = (xw[i] - ox[i]) * (1 / 3 * (yw[j] * zw[k] + oy[j] * oz[k]) + 1 / 6 * (yw[j] * oz[k] + oy[j] * zw[k]))
                    Wy[i, j, k] = (yw[j] - oy[j]) * (1 / 3 * (xw[i] * zw[k] + ox[i] * oz[k]) + 1 / 6 * (xw[i] * oz[k] + ox[i] * zw[k]))
                    Wz[i, j, k] = (zw[k] - oz[k]) * (1 / 3 * (xw[i] * yw[j] + ox[i] * oy[j]) + 1 / 6 * (xw[i] * oy[j] + ox[i] * yw[j]))
        Jx_loc = np.cumsum(dJ[0] * Wx, axis=0)
        Jy_loc = np.cumsum(dJ[1] * Wy, axis=1)
        Jz_loc = np.cumsum(dJ[2] * Wz, axis=2)
        for i in range(5):
            for j in range(5):
                for k in range(5):
                    _scatter_add(J[0], xp[i], yp[j], zp[k], Jx_loc[i, j, k])
                    _scatter_add(J[1], xp[i], yp[j], zp[k], Jy_loc[i, j, k])
                    _scatter_add(J[2], xp[i], yp[j], zp[k], Jz_loc[i, j, k])
    elif nact == 2:
        null = 0 if not xa else (1 if not ya else 2)
        # in-plane axes (a, b) in increasing order; Esirkepov.py:438-504
        a, b = [c for c in range(3) if c != null]
        wa, wb, oa, ob = W[a], W[b], OW[a], OW[b]
        Wn = np.zeros((5, 5, n)); Wa = np.zeros_like(Wn); Wb = np.zeros_like(Wn)
        for i in range(5):
            for j in range(5):
                Wn[i, j] = 1 / 3 * (wa[i] * wb[j] + oa[i] * ob[j]) + 1 / 6 * (wa[i] * ob[j] + oa[i] * wb[j])
                Wa[i, j] = 1 / 2 * (wa[i] - oa[i]) * (wb[j] + ob[j])
                Wb[i, j] = 1 / 2 * (wb[j] - ob[j]) * (wa[i] + oa[i])
        Fn = dJ[null] * Wn
        Ja = np.cumsum(dJ[a] * Wa, axis=0)
        Jb = np.cumsum(dJ[b] * Wb, axis=1)
        for i in range(5):
            for j in range(5):
                idx = [None, None, None]
                idx[null] = pts[null][2]
                idx[a] = pts[a][i]
                idx[b] = pts[b][j]
                _scatter_add(J[null], idx[0], idx[1], idx[2], Fn[i, j])
                _scatter_add(J[a], idx[0], idx[1], idx[2], Ja[i, j])
                _scatter_add(J[b], idx[0], idx[1], idx[2], Jb[i, j])
    else:
        dim = 0 if xa else (1 if ya else 2)                        # :288-329, weights :412-435
        wd, od = W[dim], OW[dim]
        Wd = np.stack([wd[i] - od[i] for i in range(5)], axis=0)
        Wo = np.stack([(wd[i] + od[i]) / 2 for i in range(5)], axis=0)
        Jd = np.cumsum(dJ[dim] * Wd, axis=0)
        others = [c for c in range(3) if c != dim]
        for i in range(5):
            idx = [pts[0][2], pts[1][2], pts[2][2]]
            idx[dim] = pts[dim][i]
            _scatter_add(J[dim], idx[0], idx[1], idx[2], Jd[i])
            for c in others:
                _scatter_add(J[c], idx[0], idx[1], idx[2], (dJ[c] * Wo)[i])
    return J


def Esirkepov_current(particles, species_config, J, sp, dp, fold=True):
    """Esirkepov.py:49-362."""
    g = int(sp.guard_cells)
    tw = tuple(int(w) for w in sp.tile_shape)
    ntx, nty, ntz = J[0].shape[:3]
    local = (tw[0] + 2 * g, tw[1] + 2 * g, tw[2] + 2 * g)
    active_axes = (ntx * tw[0] > 1, nty * tw[1] > 1, ntz * tw[2] > 1)
    tc = dp.grids.tiled_center_grid
    out = [np.zeros(J[0].shape), np.zeros(J[0].shape), np.zeros(J[0].shape)]
    for tx in range(ntx):
        for ty in range(nty):
            for tz in range(ntz):
                Jt = esirkepov_deposit_tile(particles.x[tx, ty, tz], particles.u[tx, ty, tz], particles.active[tx, ty, tz],
                                            species_config, (tc[0][tx], tc[1][ty], tc[2][tz]), sp, dp, local, active_axes)
                for c in range(3):
                    out[c][tx, ty, tz] = Jt[c]
    out = tuple(out)
    if not fold:
        return out
    out = fold_tiled_vector_ghost_cells(out, sp, num_guard_cells=g, bc_type=1)      # :357
    return update_tiled_vector_ghost_cells(out, sp, num_guard_cells=g, bc_type=1)    # :359


def _collapse_tiled(points, weights, local_n, reduced, g):
    # J_from_rhov.py:23-28 / rho.py:22-27
    if reduced:
        return np.full((1, points.shape[1]), int(g), dtype=points.dtype), np.sum(weights, axis=0, keepdims=True)
    return collapse_axis_stencil(points, weights, local_n, ghost_cells=True)


def _node_face_stencils(pos, grids_xyz, local, reduced, g, sf, d, want_face):
    pts, wn, wf = [], [], []
    for c in range(3):
        _, a, dn, p = prepare_particle_axis_stencil(pos[c], grids_xyz[c], local[c], sf, 2, ghost_cells=True)
        w_node = np.stack(shape_weights(dn, d[c], sf), axis=0)
        pc, w_node = _collapse_tiled(p, w_node, local[c], reduced[c], g)
        if want_face:
            df = (pos[c] - grids_xyz[c][0]) - (a + 0.5) * d[c]          # J_from_rhov.py:139-141
            w_face = np.stack(shape_weights(df, d[c], sf), axis=0)
            _, w_face = _collapse_tiled(pc if pc.shape[0] == 3 else p, w_face, local[c], reduced[c], g)
            wf.append(w_face)
        pts.append(pc); wn.append(w_node)
    return pts, wn, wf


def J_from_rhov(particles, species_config, J, sp, dp, fold=True):
    """J_from_rhov.py:32-260."""
    g = int(sp.guard_cells)
    tw = tuple(int(w) for w in sp.tile_shape)
    ntx, nty, ntz = J[0].shape[:3]
    local = (tw[0] + 2 * g, tw[1] + 2 * g, tw[2] + 2 * g)
    reduced = (tw[0] == 1 and ntx == 1, tw[1] == 1 and nty == 1, tw[2] == 1 and ntz == 1)
    d = (dp.dx, dp.dy, dp.dz)
    sf = int(sp.shape_factor)
    tc = dp.grids.tiled_center_grid
    qw = np.asarray(species_config.charge, dtype=np.float64) * np.asarray(species_config.weight, dtype=np.float64)
    out = [np.zeros(J[0].shape), np.zeros(J[0].shape), np.zeros(J[0].shape)]
    for tx in range(ntx):
        for ty in range(nty):
            for tz in range(ntz):
                xt = np.asarray(particles.x[tx, ty, tz], dtype=np.float64)
                ut = np.asarray(particles.u[tx, ty, tz], dtype=np.float64)
                at = np.asarray(particles.active[tx, ty, tz], dtype=bool)
                pos = [xt[..., c].reshape(-1) for c in range(3)]
                v = [ut[..., c].reshape(-1) for c in range(3)]
                active = at.reshape(-1).astype(np.float64)
                dq = np.broadcast_to(qw[:, None], at.shape).reshape(-1) / (d[0] * d[1] * d[2])
                pts, wn, wf = _node_face_stencils(pos, (tc[0][tx], tc[1][ty], tc[2][tz]), local, reduced, g, sf, d, True)
                Jt = [np.zeros(local), np.zeros(local), np.zeros(local)]
                for i in range(pts[0].shape[0]):
                    for j in range(pts[1].shape[0]):
                        for k in range(pts[2].shape[0]):                 # :181-198
                            ix, iy, iz = pts[0][i], pts[1][j], pts[2][k]
                            _scatter_add(Jt[0], ix, iy, iz, active * dq * v[0] * wf[0][i] * wn[1][j] * wn[2][k])
                            _scatter_add(Jt[1], ix, iy, iz, active * dq * v[1] * wn[0][i] * wf[1][j] * wn[2][k])
                            _scatter_add(Jt[2], ix, iy, iz, active * dq * v[2] * wn[0][i] * wn[1][j] * wf[2][k])
                for c in range(3):
                    out[c][tx, ty, tz] = Jt[c]
    out = tuple(out)
    if not fold:
        return out
    out = fold_tiled_vector_ghost_cells(out, sp, g, bc_type=1)          # :226
    out = update_tiled_vector_ghost_cells(out, sp, g, bc_type=1)        # :228
    if sp.current_filter == "bilinear":                                 # :234-255
        out = bilinear_filter_vector(out, num_guard_cells=g)
        out = update_tiled_vector_ghost_cells(out, sp, num_guard_cells=g, bc_type=1)
    elif sp.current_filter == "digital":
        out = digital_filter_vector(out, dp.alpha, num_guard_cells=g)
        out = update_tiled_vector_ghost_cells(out, sp, num_guard_cells=g, bc_type=1)
    return out


def compute_rho(particles, species_config, rho, sp, dp, fold=True):
    """rho.py:30-196."""
    g = int(sp.guard_cells)
    tw = tuple(int(w) for w in sp.tile_shape)
    ntx, nty, ntz = rho.shape[:3]
    local = (tw[0] + 2 * g, tw[1] + 2 * g, tw[2] + 2 * g)
    reduced = (tw[0] == 1 and ntx == 1, tw[1] == 1 and nty == 1, tw[2] == 1 and ntz == 1)
    d = (dp.dx, dp.dy, dp.dz)
    sf = int(sp.shape_factor)
    tc = dp.grids.tiled_center_grid
    qv = np.asarray(species_config.charge, dtype=np.float64) * np.asarray(species_config.weight, dtype=np.float64) / (d[0] * d[1] * d[2])
    out = np.zeros(rho.shape)
    for tx in range(ntx):
        for ty in range(nty):
            for tz in range(ntz):
                xt = np.asarray(particles.x[tx, ty, tz], dtype=np.float64)
                at = np.asarray(particles.active[tx, ty, tz], dtype=bool)
                pos = [xt[..., c].reshape(-1) for c in range(3)]
                active = at.reshape(-1).astype(np.float64)
                q = np.broadcast_to(qv[:, None], at.shape).reshape(-1)
                pts, wn, _ = _node_face_stencils(pos, (tc[0][tx], tc[1][ty], tc[2][tz]), local, reduced, g, sf, d, False)
                rt = np.zeros(local)
                for i in range(pts[0].shape[0]):
                    for j in range(pts[1].shape[0]):
                        for k in range(pts[2].shape[0]):
                            _scatter_add(rt, pts[0][i], pts[1][j], pts[2][k], active * q * wn[0][i] * wn[1][j] * wn[2][k])
                out[tx, ty, tz] = rt
    if not fold:
        return out
    out = fold_tiled_ghost_cells(out, sp, g, bc_type=1)
    out = update_tiled_ghost_cells(out, sp, g, bc_type=1)
    if sp.current_filter == "digital":                                  # rho.py:180-191
        out = digital_filter(out, dp.alpha, num_guard_cells=g)
        out = update_tiled_ghost_cells(out, sp, g, bc_type=1)
    return out
