"""Generates tests/golden/ref_*.npz by running the REFERENCE'S OWN CODE (/root/reference/PyPIC3D, unmodified) on the NumPy-backed
jax shim of tests/jax_shim -- `evolve.time_loop_electrodynamic` / `time_loop_electrostatic` for a few steps from the same seeded
initial states as tests/golden/make_golden.py.  Only runs where /root/reference exists (this build container); the .npz files are
committed so that the GPU box, which has no reference, can compare the CUDA paths with the reference's outputs.

    python tests/golden/make_ref_golden.py

What "reference output" means here: the reference's source lines executed with NumPy float64 arrays underneath instead of XLA
(tests/jax_shim/README.md).  Element-wise arithmetic is IEEE-identical; reduction orders are NumPy's."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.make_golden import GOLDEN, build  # noqa: E402

REF_GOLDEN = [n for n in GOLDEN]


def run_reference(name):
    from tests import ref_loader
    ref = ref_loader.reference()          # (puts the shim on sys.path)
    from tests.test_reference_code import to_ref, vec
    import jax.numpy as jnp
    c, sp, dp, tp, sc, fields = build(name)
    rsp, rdp, rtp, rsc = to_ref(ref, sp, dp, tp, sc)
    E, B, J, rho, phi, ext, pml, ovf = fields
    rf = (vec(ref, E), vec(ref, B), vec(ref, J), jnp.asarray(rho), jnp.asarray(phi), (vec(ref, ext[0]), vec(ref, ext[1])), None, jnp.asarray(bool(ovf)))
    loop = ref.evolve.time_loop_electrostatic if c.get("loop") == "electrostatic" else ref.evolve.time_loop_electrodynamic
    for _ in range(c["steps"]):
        rtp, rf = loop(rtp, rsc, rf, rsp, rdp)
    out = {"x": np.asarray(rtp.x), "u": np.asarray(rtp.u), "active": np.asarray(rtp.active), "overflow": np.array(bool(np.asarray(rf[7]))),
           "rho": np.asarray(rf[3]), "phi": np.asarray(rf[4])}
    for k, nm in enumerate("EBJ"):
        for comp in range(3):
            out[f"{nm}{comp}"] = np.asarray(rf[k][comp])
    return out


def main():
    for name in REF_GOLDEN:
        out = run_reference(name)
        np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), **out)
        print("ref_" + name, {k: v.shape for k, v in out.items() if k in ("x", "E0")})


if __name__ == "__main__":
    main()
