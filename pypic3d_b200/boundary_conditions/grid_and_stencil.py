"""Boundary codes of PyPIC3D/boundary_conditions/grid_and_stencil.py:10-11 and the host-side axis builders (:242-286)."""
import numpy as np

BC_PERIODIC = 0
BC_CONDUCTING = 1


def build_collocated_axis(minimum_physical, spacing, count):
    return np.linspace(minimum_physical - spacing, minimum_physical + count * spacing, int(count) + 2)


def build_staggered_axis(minimum_physical, spacing, count):
    return np.linspace(minimum_physical - 0.5 * spacing, minimum_physical + (count + 0.5) * spacing, int(count) + 2)
