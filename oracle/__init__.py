"""
oracle/ -- CPU restatement of the PyPIC3D electrodynamic PIC step (TEST INFRASTRUCTURE ONLY).

This package is a NumPy float64 restatement of the reference algorithm
(`/root/reference/PyPIC3D`, v0.1.3) for the hot path named by BASELINE.json:
push (gather + Boris), Esirkepov / direct (rho*v) / rho deposition, particle move +
retile/migration, first-order Yee E/B update, digital/bilinear filters and the guard-cell
refresh/fold.  Every function cites the reference file:line it follows.

It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  The product package
(`pypic3d_b200/`) never imports, links or executes anything in here, and has no CPU
fallback: it raises if the CUDA library is missing.

Pinning: the reference is pure Python/JAX and `jax`/`jaxlib` are not installed in this
image (no wheel in /opt/wheelhouse, no network), so the reference itself cannot be run
here.  The oracle is instead pinned against every known-answer test, manual
single-particle restatement and invariant that the reference's own test-suite holds for
this path; those are ported one-for-one in `tests/test_oracle_*.py`
with the reference test file:line next to each.  The reference ships no golden vectors;
multi-step observables (growth rates, energy histories) are therefore "parity unpinned
against reference outputs" -- they are compared CUDA-vs-oracle on identical initial
state only.
"""

from . import params, grids, stencil, shapes, pusher, halo, filters, deposition, particles, yee, evolve, diagnostics, fixtures  # noqa: F401
