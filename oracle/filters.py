"""27-point digital / bilinear filters (test infrastructure).  Follows PyPIC3D/utilities/filters.py:6-155."""
import numpy as np


def _valid_conv3(phi, kernel):
    # filters.py:6-43: VALID 3x3x3 correlation over the last three axes (lax.conv does not flip the kernel)
    out = np.zeros(phi.shape[:-3] + tuple(s - 2 for s in phi.shape[-3:]), dtype=np.float64)
    nx, ny, nz = out.shape[-3:]
    for i in range(3):
        for j in range(3):
            for k in range(3):
                if kernel[i, j, k] != 0.0:
                    out = out + kernel[i, j, k] * phi[..., i:i + nx, j:j + ny, k:k + nz]
    return out


def _apply(phi, kernel, g):
    g = int(g)
    phi = np.array(phi, dtype=np.float64, copy=True)
    st = slice(g - 1, None if g == 1 else -g + 1)     # filters.py:50-52
    act = slice(g, -g)
    phi[..., act, act, act] = _valid_conv3(phi[..., st, st, st], kernel)
    return phi


def bilinear_filter(phi, num_guard_cells=1):
    k1 = np.array([1.0, 2.0, 1.0])                    # filters.py:73-94
    kernel = (k1[:, None, None] * k1[None, :, None] * k1[None, None, :]) / 64.0
    return _apply(phi, kernel, num_guard_cells)


def digital_filter(phi, alpha, num_guard_cells=1):
    w = (1.0 - alpha) / 6.0                           # filters.py:98-129
    kernel = np.zeros((3, 3, 3))
    kernel[1, 1, 1] = alpha
    kernel[0, 1, 1] = kernel[2, 1, 1] = w
    kernel[1, 0, 1] = kernel[1, 2, 1] = w
    kernel[1, 1, 0] = kernel[1, 1, 2] = w
    return _apply(phi, kernel, num_guard_cells)


def bilinear_filter_vector(field, num_guard_cells=1):
    return tuple(bilinear_filter(c, num_guard_cells) for c in field)


def digital_filter_vector(field, alpha, num_guard_cells=1):
    return tuple(digital_filter(c, alpha, num_guard_cells) for c in field)
