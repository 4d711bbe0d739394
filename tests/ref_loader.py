"""Load the reference's hot-path modules UNMODIFIED from /root/reference on top of tests/jax_shim (TEST INFRASTRUCTURE ONLY).

`reference()` returns a namespace of the reference modules, or None when /root/reference is absent (the GPU box): callers skip.
The reference package's own `__init__.py` is not executed (it imports matplotlib, plotly, openPMD, tqdm, ...): a bare package
object with the right `__path__` stands in for it, so `from PyPIC3D.deposition.shapes import ...` inside the reference resolves to
the files under /root/reference/PyPIC3D.  Modules with heavy optional imports (utils.py) are not loaded."""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("PYPIC3D_REFERENCE", "/root/reference")
_NS = None


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "PyPIC3D"))


def reference():
    global _NS
    if _NS is not None:
        return _NS
    if not available():
        return None
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jax_shim")
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__version__", "").endswith("numpy-shim"):
        raise RuntimeError("a real jax is already imported; the shim-backed reference loader must not mix with it")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    pkg = types.ModuleType("PyPIC3D")
    pkg.__path__ = [os.path.join(REF_ROOT, "PyPIC3D")]
    pkg.__package__ = "PyPIC3D"
    sys.modules["PyPIC3D"] = pkg
    # sub-packages with an __init__ that pulls in more than the hot path are stubbed the same way
    for sub in ("boundary_conditions", "pusher", "deposition", "particles", "utilities", "solvers", "diagnostics"):
        m = types.ModuleType(f"PyPIC3D.{sub}")
        m.__path__ = [os.path.join(REF_ROOT, "PyPIC3D", sub)]
        m.__package__ = f"PyPIC3D.{sub}"
        sys.modules[f"PyPIC3D.{sub}"] = m
        setattr(pkg, sub, m)
    # (PyPIC3D.utils imports plotly at module level without using it on the hot path: tests/jax_shim/plotly is an empty stand-in)
    ns = types.SimpleNamespace()
    names = {
        "shapes": "PyPIC3D.deposition.shapes",
        "grid_and_stencil": "PyPIC3D.boundary_conditions.grid_and_stencil",
        "ghost_cells": "PyPIC3D.boundary_conditions.ghost_cells",
        "boris": "PyPIC3D.pusher.boris",
        "higuera_cary": "PyPIC3D.pusher.higuera_cary",
        "particle_class": "PyPIC3D.particles.particle_class",
        "parameters": "PyPIC3D.parameters",
        "grids": "PyPIC3D.utilities.grids",
        "filters": "PyPIC3D.utilities.filters",
        "particle_push": "PyPIC3D.pusher.particle_push",
        "Esirkepov": "PyPIC3D.deposition.Esirkepov",
        "J_from_rhov": "PyPIC3D.deposition.J_from_rhov",
        "rho": "PyPIC3D.deposition.rho",
        "first_order_yee": "PyPIC3D.solvers.first_order_yee",
        "particle_tile_communication": "PyPIC3D.particles.particle_tile_communication",
        "utils": "PyPIC3D.utils",
        "electrostatic_yee": "PyPIC3D.solvers.electrostatic_yee",
        "evolve": "PyPIC3D.evolve",
    }
    for short, full in names.items():
        setattr(ns, short, importlib.import_module(full))
    _NS = ns
    return ns
