"""Independent scalar restatements used to cross-check the oracle (and, on the GPU box, the CUDA path).

* `manual_*`: single-particle rho / direct-J / 1-D Esirkepov on a GLOBAL periodic grid in pure Python floats --
  the counterpart of the reference's `_manual_rho_tiles`, `_manual_direct_current_tiles`,
  `_manual_esirkepov_current_tiles_1d` (tests/code_tests/single_particle_pipeline_test.py:229-320), but built
  without tiles / ghost fold so it also checks the fold+refresh machinery.
* `global_index_refresh/fold`: one-shot global-index formulation of the axis-sequential halo refresh / fold
  (ghost_cells.py:142-316).  This is the formulation the CUDA halo kernels use.
"""
import math

import numpy as np


def _w(delta_over_d, sf):
    r = delta_over_d
    if sf == 1:
        return [0.0, 1.0 - r, r]
    return [0.5 * (0.5 - r) ** 2, 0.75 - r * r, 0.5 * (0.5 + r) ** 2]


def _anchor(xi, sf):
    if sf == 1:
        return math.floor(xi)
    a = math.floor(xi + 0.5)           # round half to even
    if xi + 0.5 == a and a % 2 == 1:
        a -= 1
    return a


def manual_rho_1d(x, qw, n, wind, sf):
    d = wind / n
    xi = (x + wind / 2) / d
    a = _anchor(xi, sf)
    w = _w(xi - a, sf)
    rho = np.zeros(n)
    for k in range(3):
        rho[(a - 1 + k) % n] += qw / d * w[k]
    return rho


def manual_direct_J_1d(x, u, qw, n, wind, sf):
    d = wind / n
    xi = (x + wind / 2) / d
    a = _anchor(xi, sf)
    wn = _w(xi - a, sf)
    wf = _w(xi - a - 0.5, sf)
    J = np.zeros((3, n))
    for k in range(3):
        i = (a - 1 + k) % n
        J[0, i] += qw / d * u[0] * wf[k]
        J[1, i] += qw / d * u[1] * wn[k]
        J[2, i] += qw / d * u[2] * wn[k]
    return J


def manual_esirkepov_J_1d(x, u, qw, n, wind, sf, dt):
    d = wind / n
    xo, xn = (x + wind / 2) / d, (x + u[0] * dt + wind / 2) / d
    ao, an = _anchor(xo, sf), _anchor(xn, sf)
    s1 = [0.0] * 5
    s0 = [0.0] * 5
    wn, wo = _w(xn - an, sf), _w(xo - ao, sf)
    for k in range(3):
        s1[1 + k] = wn[k]
        j = 1 + k + (ao - an)        # old weights expressed in the new-anchor frame
        assert 0 <= j < 5
        s0[j] = wo[k]
    J = np.zeros((3, n))
    acc = 0.0
    for i in range(5):
        acc += (-(qw) / dt) * (s1[i] - s0[i])          # dy = dz = 1
        node = (an - 2 + i) % n
        J[0, node] += acc
        J[1, node] += qw * u[1] / d * 0.5 * (s1[i] + s0[i])
        J[2, node] += qw * u[2] / d * 0.5 * (s1[i] + s0[i])
    return J


def _axis_target(t, l, W, g, nt, bc, reduced, mode):
    """Map local index l of tile t on one axis to its owner.  Returns (tile, local, sign) or None (drop/zero)."""
    if g <= l < g + W:
        return t, l, 1.0
    if reduced:
        if bc == 0:
            return t, g, 1.0
        if mode == "fold" and bc == 1:
            return t, g, -1.0
        return None
    G = t * W + l - g
    N = nt * W
    if 0 <= G < N:
        return G // W, G % W + g, 1.0
    if bc == 0:
        G %= N
        return G // W, G % W + g, 1.0
    if mode == "fold" and bc == 1:
        return (t, l + g, -1.0) if l < g else (t, l - g, -1.0)
    return None


def global_index_refresh(tiles, tile_shape, bcs, g):
    tiles = np.asarray(tiles, dtype=np.float64)
    mesh, L = tiles.shape[:3], tiles.shape[3:]
    red = [tile_shape[a] == 1 and mesh[a] == 1 for a in range(3)]
    out = tiles.copy()
    for idx in np.ndindex(*tiles.shape):
        t, l = idx[:3], idx[3:]
        if all(g <= l[a] < g + tile_shape[a] for a in range(3)):
            continue
        src = [_axis_target(t[a], l[a], tile_shape[a], g, mesh[a], bcs[a], red[a], "refresh") for a in range(3)]
        if any(s is None for s in src):
            out[idx] = 0.0
        else:
            out[idx] = tiles[src[0][0], src[1][0], src[2][0], src[0][1], src[1][1], src[2][1]]
    return out


def global_index_fold(tiles, tile_shape, bcs, g):
    tiles = np.asarray(tiles, dtype=np.float64)
    mesh = tiles.shape[:3]
    red = [tile_shape[a] == 1 and mesh[a] == 1 for a in range(3)]
    out = np.zeros_like(tiles)
    for idx in np.ndindex(*tiles.shape):
        t, l = idx[:3], idx[3:]
        if all(g <= l[a] < g + tile_shape[a] for a in range(3)):
            out[idx] += tiles[idx]
            continue
        if tiles[idx] == 0.0:
            continue
        dst = [_axis_target(t[a], l[a], tile_shape[a], g, mesh[a], bcs[a], red[a], "fold") for a in range(3)]
        if any(s is None for s in dst):
            continue
        out[dst[0][0], dst[1][0], dst[2][0], dst[0][1], dst[1][1], dst[2][1]] += dst[0][2] * dst[1][2] * dst[2][2] * tiles[idx]
    return out
