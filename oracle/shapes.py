"""Particle shape weights (test infrastructure).  Follows PyPIC3D/deposition/shapes.py:6-54."""
import numpy as np


def first_order_weights(delta, d):
    # shapes.py:36-54  -> [0, 1-delta/d, delta/d]
    s0 = 1 - delta / d
    s1 = delta / d
    return [np.zeros_like(s0), s0, s1]


def second_order_weights(delta, d):
    # shapes.py:6-32
    r = delta / d
    s0 = (3 / 4) - r ** 2
    s1 = (1 / 2) * ((1 / 2) + r) ** 2
    sm = (1 / 2) * ((1 / 2) - r) ** 2
    return [sm, s0, s1]


def weights(delta, d, shape_factor):
    return first_order_weights(delta, d) if int(shape_factor) == 1 else second_order_weights(delta, d)
