"""ms per step of BASELINE.json's configurations 1-3 -- the packaged demos -- on the resident CUDA path, with the oracle (NumPy
port of the reference, one host core) timed beside it on the same initial state.

    python tools/demo_bench.py [--steps 200] [--oracle-steps 5] [--dtype f64] > profiles/r02_demo_bench.json

  two_stream       demos/two_stream/two_stream.toml as packaged: 100 x 1 x 1 cells, 6000 particles, TSC, j_from_rhov
  weibel           demos/weibel/weibel.toml as packaged: 201 x 1 x 1 cells, 16080 particles, TSC, j_from_rhov + bilinear
  reconnection_2d  demos/reconnection_2d/harris_current.toml at its packaged size: 500 x 1 x 500 cells, 260000 particles in four
                   species, TSC, j_from_rhov + bilinear, x periodic / z conducting, cfl 0.1; the .npy initial conditions are
                   produced by the demo's recipe (initial_conditions.py) with a seeded generator, the sheet profile sampled by
                   inverse CDF (tests/test_frontend.py::test_harris_sheet_config_matches_oracle explains why)

These are launch-latency-bound problems (6 k - 260 k particles): the number that matters is device time per step; the oracle
column is what the reference's algorithm costs in NumPy on one core (the reference itself needs jax, which this image lacks).
Not a bench.py line: bench.py's contract is the 256^3 workload (BASELINE.json configs[3])."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def harris_config(tmp, nx=500, n_peak=100000, n_bg=30000, seed=7):
    eps, me, c, q = 8.854e-12, 9.10938356e-31, 2.99792458e8, 1.602e-19
    vth = 0.05 * c
    n0 = 1e22
    nb = 0.3 * n0
    di = c / (q * np.sqrt(n0 / me / eps))
    wind = 15 * di
    lam, B0 = 0.5 * di, 0.2
    x = np.linspace(-wind / 2, wind / 2, nx)
    X, Z = np.meshgrid(x, x, indexing="ij")
    files = {"Bx": (B0 * np.tanh(Z / lam) + 100 * B0 * np.cos(2 * np.pi * X / wind) * np.sin(np.pi * Z / wind))[:, None, :],
             "Bz": (-100 * B0 * np.sin(2 * np.pi * X / wind) * np.cos(np.pi * Z / wind))[:, None, :]}
    rng = np.random.default_rng(seed)
    for name in ("electron", "ion"):
        files[f"{name}_x"] = rng.uniform(-wind / 2, wind / 2, n_peak)
        files[f"{name}_y"] = np.zeros(n_peak)
        files[f"{name}_z"] = np.clip(lam * np.arctanh(rng.uniform(-0.999, 0.999, n_peak)), -0.45 * wind, 0.45 * wind)
        for a in "xyz":
            files[f"{name}_v{a}"] = rng.normal(0, vth, n_peak)
    zbar = files["electron_z"] / lam
    files["electron_vy"] = files["electron_vy"] + (-c * B0 / (4 * np.pi * q * lam)) / np.cosh(zbar) ** 2 / (n0 * np.cosh(zbar) ** 2 + nb)
    for k, v in files.items():
        np.save(os.path.join(tmp, f"{k}.npy"), v)
    path = lambda k: os.path.join(tmp, f"{k}.npy")
    weight = n0 / n_peak * wind * wind
    species = lambda name, n, charge, pre=None: dict(
        {"name": name, "N_particles": n, "charge": charge, "mass": 9.1093837e-31, "vth": 14989622.9, "weight": weight,
         "x_bc": "periodic", "z_bc": "reflecting", "y_bc": "periodic"},
        **({} if pre is None else {f"initial_{a}": path(f"{pre}_{a}") for a in ("x", "y", "z", "vx", "vy", "vz")}))
    return {
        "simulation_parameters": {"name": "harris", "Nt": 1500, "x_bc": "periodic", "z_bc": "conducting", "solver": "electrodynamic_yee",
                                  "Nx": nx, "Ny": 1, "Nz": nx, "x_wind": wind, "y_wind": 1, "z_wind": wind, "verbose": False,
                                  "cfl": 0.1, "shape_factor": 2, "relativistic": True, "output_dir": tmp,
                                  "particle_tile_capacity_factor": 1.5},
        "plotting": {"plotting_interval": 50},
        "field1": {"name": "Bx field", "type": 3, "path": path("Bx")},
        "field2": {"name": "Bz field", "type": 5, "path": path("Bz")},
        "particle1": species("peak electrons", n_peak, -1.602e-19, "electron"),
        "particle2": species("peak ions", n_peak, 1.602e-19, "ion"),
        "particle3": species("background electrons", n_bg, -1.602e-19),
        "particle4": species("background ions", n_bg, 1.602e-19),
    }


def run_one(name, cfg, steps, oracle_steps, dtype):
    import torch
    from pypic3d_b200.initialization import initialize_simulation
    from pypic3d_b200.simulation import Simulation
    from pypic3d_b200.utils import compute_energy, add_external_fields
    from oracle import evolve as oevolve
    from oracle.params import StaticParameters as OS, DynamicParameters as OD, GridParameters as OG, TiledParticles as OT, SpeciesConfig as OC
    npy = lambda t: t.detach().cpu().numpy()
    np.random.seed(0)
    loop, particles, fields, sp, dp, plotting, plasma, species = initialize_simulation(cfg, dtype=dtype, verbose=False)
    n_particles = int(particles.active.sum().item())
    # ---- oracle (float64, one core) on the identical initial state
    osp = OS(**sp._asdict()); odp = OD(**{**dp._asdict(), "grids": OG(**dp.grids._asdict())})
    otp = OT(npy(particles.x).astype(np.float64), npy(particles.u).astype(np.float64), npy(particles.active))
    osc = OC(*[np.asarray(v) for v in species])
    n = lambda F: tuple(npy(c).astype(np.float64) for c in F)
    of = (n(fields[0]), n(fields[1]), n(fields[2]), npy(fields[3]).astype(np.float64), npy(fields[4]).astype(np.float64),
          (n(fields[5][0]), n(fields[5][1])), None, False)
    otp, of = oevolve.time_loop_electrodynamic(otp, osc, of, osp, odp)          # warm-up
    t0 = time.perf_counter()
    for _ in range(oracle_steps):
        otp, of = oevolve.time_loop_electrodynamic(otp, osc, of, osp, odp)
    oracle_ms = (time.perf_counter() - t0) / oracle_steps * 1e3
    # ---- resident CUDA path
    sim = Simulation(particles, species, fields, sp, dp)
    sim.step(10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    sim.step(steps)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) / steps * 1e3
    dev_ms = e0.elapsed_time(e1) / steps
    parts, f2 = sim.export_state()
    tE, tB = add_external_fields(f2[0], f2[1], f2[5])
    e, b, k = compute_energy(parts, tE, tB, sp, dp, species_config=species)
    return {"workload": name, "cells": [int(dp.Nx), int(dp.Ny), int(dp.Nz)], "particles": n_particles, "shape_factor": int(sp.shape_factor),
            "current_deposition": sp.current_deposition, "current_filter": sp.current_filter, "k1_variant": sim.k1_variant,
            "dtype": str(dtype).replace("torch.", ""), "steps": steps, "ms_per_step_device": dev_ms, "ms_per_step_wall": wall_ms,
            "particle_steps_per_s": n_particles / (dev_ms * 1e-3), "oracle_ms_per_step_1core": oracle_ms, "oracle_steps": oracle_steps,
            "speedup_vs_oracle": oracle_ms / dev_ms, "overflow": bool(sim.overflow()),
            "energies_after": {"electric": float(e), "magnetic": float(b), "kinetic": float(k)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--oracle-steps", type=int, default=5)
    ap.add_argument("--dtype", default="f64", choices=("f32", "f64"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    from tests.test_frontend import TWO_STREAM, WEIBEL
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        cfgs = {"two_stream": {k: dict(v) for k, v in TWO_STREAM.items()}, "weibel": {k: dict(v) for k, v in WEIBEL.items()},
                "reconnection_2d": harris_config(tmp)}
        for name, cfg in cfgs.items():
            if args.only and name != args.only:
                continue
            cfg["simulation_parameters"]["output_dir"] = tmp
            out.append(run_one(name, cfg, args.steps, args.oracle_steps if name != "reconnection_2d" else max(1, args.oracle_steps // 2), dtype))
            print(json.dumps(out[-1]), file=sys.stderr)
    print(json.dumps({"what": "BASELINE.json configs 1-3 (packaged demos) on the resident CUDA path, one B200; oracle = NumPy port of the "
                              "reference on one host core, same initial state", "results": out}))


if __name__ == "__main__":
    main()
