"""Generates tests/golden/*.npz with the NumPy oracle (the reference itself cannot run here: no jax).
These are regression pins of multi-step trajectories: initial state + state after `steps` steps, float64.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import evolve as oevolve, fixtures as fx  # noqa: E402
from tests.cases import make_case, make_fields  # noqa: E402

GOLDEN = {
    "esirkepov_tsc_3d": dict(N=(8, 6, 4), tile=(8, 6, 4), sf=2, steps=3, kw=dict(current_deposition="esirkepov")),
    "esirkepov_cic_3d_tiled": dict(N=(8, 6, 4), tile=(4, 3, 2), sf=1, steps=3, kw=dict(current_deposition="esirkepov")),
    "direct_bilinear_1d_two_stream_like": dict(N=(16, 1, 1), tile=(8, 1, 1), sf=2, steps=4,
                                              kw=dict(current_deposition="direct", current_filter="bilinear")),
    "direct_bilinear_2d_conducting_z": dict(N=(8, 1, 6), tile=(8, 1, 6), sf=2, steps=3,
                                           kw=dict(current_deposition="direct", current_filter="bilinear", boundary_conditions=(0, 0, 1))),
    # tile widths that are multiples of 4 with g = 2: the resident path runs the supercell tile kernel (K1 v9) on this one
    "esirkepov_cic_3d_supercells": dict(N=(8, 8, 4), tile=(8, 8, 4), sf=1, steps=4, kw=dict(current_deposition="esirkepov")),
    # evolve.py:106 time_loop_electrostatic on a neutral thermal plasma (a consistent periodic Poisson problem), eps = 1
    "electrostatic_periodic_tsc": dict(N=(8, 8, 8), tile=(8, 8, 8), sf=2, steps=3, loop="electrostatic",
                                      kw=dict(solver="electrostatic", electrostatic=True, alpha=0.9)),
}


def build(name):
    c = GOLDEN[name]
    if c.get("loop") == "electrostatic":
        N = c["N"]
        sp, dp = fx.kernel_parameters(Nx=N[0], Ny=N[1], Nz=N[2], x_wind=float(N[0]), y_wind=float(N[1]), z_wind=float(N[2]),
                                      tile_shape=c["tile"], dt=0.1, shape_factor=c["sf"], **c["kw"])
        tp, sc = fx.thermal_plasma(sp, dp, ppc_per_species=2, vth=(0.1, 0.01), seed=4)
        z = fx.empty_tiled_vector
        fields = (z(sp, dp), z(sp, dp), z(sp, dp), fx.empty_tiled_scalar(sp, dp), fx.empty_tiled_scalar(sp, dp), (z(sp, dp), z(sp, dp)),
                  None, False)
        return c, sp, dp, tp, sc, fields
    sp, dp, tp, sc, E, B = make_case(c["N"], c["tile"], c["sf"], capacity=3.0, vmax=0.3, dt=0.04, **c["kw"])
    return c, sp, dp, tp, sc, make_fields(sp, dp)


def loop_of(c):
    return oevolve.time_loop_electrostatic if c.get("loop") == "electrostatic" else oevolve.time_loop_electrodynamic


def main():
    for name in GOLDEN:
        c, sp, dp, tp, sc, fields = build(name)
        tp1, f1 = tp, fields
        for _ in range(c["steps"]):
            tp1, f1 = loop_of(c)(tp1, sc, f1, sp, dp)
        out = {"x": tp1.x, "u": tp1.u, "active": tp1.active, "overflow": np.array(bool(f1[7])), "rho": np.asarray(f1[3]), "phi": np.asarray(f1[4])}
        for k, nm in enumerate("EBJ"):
            for comp in range(3):
                out[f"{nm}{comp}"] = np.asarray(f1[k][comp])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k in ("x", "E0")})


if __name__ == "__main__":
    main()
