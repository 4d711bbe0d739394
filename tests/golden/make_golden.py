"""Generates tests/golden/*.npz with the NumPy oracle (the reference itself cannot run here: no jax).
These are regression pins of multi-step trajectories: initial state + state after `steps` steps, float64.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import evolve as oevolve  # noqa: E402
from tests.cases import make_case, make_fields  # noqa: E402

GOLDEN = {
    "esirkepov_tsc_3d": dict(N=(8, 6, 4), tile=(8, 6, 4), sf=2, steps=3, kw=dict(current_deposition="esirkepov")),
    "esirkepov_cic_3d_tiled": dict(N=(8, 6, 4), tile=(4, 3, 2), sf=1, steps=3, kw=dict(current_deposition="esirkepov")),
    "direct_bilinear_1d_two_stream_like": dict(N=(16, 1, 1), tile=(8, 1, 1), sf=2, steps=4,
                                              kw=dict(current_deposition="direct", current_filter="bilinear")),
    "direct_bilinear_2d_conducting_z": dict(N=(8, 1, 6), tile=(8, 1, 6), sf=2, steps=3,
                                           kw=dict(current_deposition="direct", current_filter="bilinear", boundary_conditions=(0, 0, 1))),
}


def build(name):
    c = GOLDEN[name]
    sp, dp, tp, sc, E, B = make_case(c["N"], c["tile"], c["sf"], capacity=3.0, vmax=0.3, dt=0.04, **c["kw"])
    return c, sp, dp, tp, sc, make_fields(sp, dp)


def main():
    for name in GOLDEN:
        c, sp, dp, tp, sc, fields = build(name)
        tp1, f1 = tp, fields
        for _ in range(c["steps"]):
            tp1, f1 = oevolve.time_loop_electrodynamic(tp1, sc, f1, sp, dp)
        out = {"x": tp1.x, "u": tp1.u, "active": tp1.active, "overflow": np.array(bool(f1[7]))}
        for k, nm in enumerate("EBJ"):
            for comp in range(3):
                out[f"{nm}{comp}"] = f1[k][comp]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k in ("x", "E0")})


if __name__ == "__main__":
    main()
