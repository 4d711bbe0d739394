"""`python -m pypic3d_b200 --config x.toml` -- the reference's `PyPIC3D --config` workflow (PyPIC3D/__main__.py:38-237) on the
CUDA path: initialise from TOML, step Nt times, write the same `data/*.txt` energy/momentum histories and `data/output.toml`.
A single-tile configuration runs on the resident fused path (`Simulation`); a multi-tile one on the drop-in composition."""
import argparse
import time

import toml
import torch

from .diagnostics.plotting import write_data
from .initialization import initialize_simulation
from .simulation import Simulation
from .utils import add_external_fields, compute_energy, compute_total_momentum, dump_parameters_to_toml


def _raise_if_overflowed(flag):
    if bool(flag):                                                   # __main__.py:29-35
        raise RuntimeError("tiled particle tile capacity overflowed during periodic retile")


def run_PyPIC3D(config_file, dtype=torch.float64, resident=None, verbose=True):
    loop, particles, fields, sp, dp, plotting, plasma, species = initialize_simulation(config_file, dtype=dtype, verbose=verbose)
    dt, Nt, out = dp.dt, sp.Nt, sp.output_dir
    single_tile = tuple(particles.x.shape[:3]) == (1, 1, 1)
    resident = single_tile if resident is None else (resident and single_tile)
    resident = resident and not sp.electrostatic       # the resident fused path is the electrodynamic step
    sim = Simulation(particles, species, fields, sp, dp) if resident else None

    def energies(particles, fields):
        E, B, J, rho, phi, ext, *rest = fields
        tE, tB = add_external_fields(E, B, ext)
        e, b, k = compute_energy(particles, tE, tB, sp, dp, species_config=species)
        return float(e), float(b), float(k)

    initial_energy = sum(energies(particles, fields))
    interval = int(plotting["plotting_interval"])
    for t in range(Nt):
        if t % interval == 0:
            if sim is not None:
                particles, fields = sim.export_state()
            e, b, k = energies(particles, fields)
            tot = e + b + k
            write_data(f"{out}/data/total_energy.txt", t * dt, tot)
            write_data(f"{out}/data/energy_error.txt", t * dt, abs(initial_energy - tot) / max(initial_energy, 1e-10))
            write_data(f"{out}/data/electric_field_energy.txt", t * dt, e)
            write_data(f"{out}/data/magnetic_field_energy.txt", t * dt, b)
            write_data(f"{out}/data/kinetic_energy.txt", t * dt, k)
            write_data(f"{out}/data/total_momentum.txt", t * dt, float(compute_total_momentum(particles, species_config=species)))
        if sim is not None:
            sim.step(1)
            if (t + 1) % interval == 0 or t == Nt - 1:
                _raise_if_overflowed(sim.overflow())
        else:
            particles, fields = loop(particles, species, fields, sp, dp)
            _raise_if_overflowed(fields[7].item())
    if sim is not None:
        particles, fields = sim.export_state()
    return sp, dp, plotting, plasma, particles, fields, species


def main(argv=None):
    parser = argparse.ArgumentParser(description="PyPIC3D electrodynamic PIC step on B200 (sm_100a CUDA)")
    parser.add_argument("--config", type=str, required=True, help="Path to the configuration file")
    parser.add_argument("--dtype", default="float64", choices=("float32", "float64"))
    args = parser.parse_args(argv)
    print(f"Using Configuration File: {args.config}")
    cfg = toml.load(args.config)
    start = time.time()
    sp, dp, plotting, plasma, particles, fields, species = run_PyPIC3D(cfg, dtype=getattr(torch, args.dtype))
    torch.cuda.synchronize()
    duration = time.time() - start
    E, B, J, rho, phi, ext, *rest = fields
    tE, tB = add_external_fields(E, B, ext)
    e, b, k = compute_energy(particles, tE, tB, sp, dp, species_config=species)
    print(f"Final Electric Field Energy: {float(e)}\nFinal Magnetic Field Energy: {float(b)}\nFinal Kinetic Energy: {float(k)}")
    print(f"Total Final Energy: {float(e) + float(b) + float(k)}\n")
    stats = {"total_time": duration, "total_iterations": sp.Nt, "time_per_iteration": duration / max(sp.Nt, 1)}
    dump_parameters_to_toml(stats, sp, dp, plasma, plotting, particles)                # __main__.py:226-233
    print(f"\nSimulation Complete\nTotal Simulation Time: {duration} s\nTime Per Iteration: {duration / max(sp.Nt, 1)} s")


if __name__ == "__main__":
    main()
