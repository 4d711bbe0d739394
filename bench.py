#!/usr/bin/env python
"""bench.py -- particle-steps/s of the PIC hot path (push + deposit + Yee + guard cells) on N B200s.

A "step" is one electrodynamic PIC step (PyPIC3D/evolve.py:16-103) over the whole synthetic thermal plasma:
BASELINE.json configs[3] at N=1 (256^3 cells x 16 ppc, Esirkepov + Yee, periodic) and configs[4] for N>1 (the same box
per GPU, weak scaling, NCCL halo exchange + particle migration).

    python bench.py --gpus 1 --steps 20 --warmup 5            # ours (CUDA, C ABI)
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm: NumPy restatement of the reference (oracle), one tile per core
    torchrun ... bench.py --gpus N ...                         # N>1: one rank per GPU

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

C_LIGHT = 299792458.0
EPS0 = 8.8541878128e-12
MU0 = 1.25663706212e-6
QE = 1.602176634e-19
ME = 9.1093837015e-31
MP = 1.67262192369e-27


def physical_setup(n_local, mesh, ppc, shape_factor, dtype_name):
    """SURVEY.md section 8d input 4: thermal e-/p+ plasma, n0 = 1e18 m^-3, dx = lambda_D, vth_e = 0.05 c, dt = 0.99 dx/(3c)."""
    vth_e = 0.05 * C_LIGHT
    n0 = 1e18
    kT = ME * vth_e ** 2
    lam = math.sqrt(EPS0 * kT / (n0 * QE ** 2))
    dx = lam
    N = [n_local * mesh[a] for a in range(3)]
    dt = 0.99 / (C_LIGHT * 3.0 / dx)                        # utils.py:761-801 courant_condition, 3 active axes
    weight = n0 * dx ** 3 / (ppc // 2)
    return dict(N=N, dx=dx, dt=dt, weight=weight, vth=(vth_e, vth_e * math.sqrt(ME / MP)), charge=(-QE, QE), mass=(ME, MP),
                wind=[N[a] * dx for a in range(3)])


def make_params(cfg, n_local, mesh, shape_factor, deposition="esirkepov", current_filter="none"):
    from pypic3d_b200.parameters import StaticParameters, DynamicParameters, GridParameters
    from pypic3d_b200.utilities.grids import build_yee_grid
    N, dx = cfg["N"], cfg["dx"]
    sp = StaticParameters(name="bench", output_dir=".", Nt=0, verbose=False, GPUs=True, benchmark=True, solver="electrodynamic_yee",
                          electrostatic=False, relativistic=True, particle_pusher="boris", current_deposition=deposition,
                          current_filter=current_filter, shape_factor=shape_factor, guard_cells=2, tile_shape=(n_local,) * 3,
                          particle_tile_capacity_factor=1.25, pml_active=False, boundary_conditions=(0, 0, 0),
                          particle_boundary_conditions=(0, 0, 0), field_mesh=tuple(mesh))
    dp = DynamicParameters(dt=cfg["dt"], dx=dx, dy=dx, dz=dx, Nx=N[0], Ny=N[1], Nz=N[2], x_wind=cfg["wind"][0], y_wind=cfg["wind"][1],
                           z_wind=cfg["wind"][2], C=C_LIGHT, eps=EPS0, mu=MU0, kb=1.380649e-23, alpha=1.0,
                           grids=GridParameters((), (), (), ()))
    center, vertex = build_yee_grid(dp)
    dp = dp._replace(grids=GridParameters(vertex=vertex, center=center, tiled_vertex_grid=(), tiled_center_grid=()))
    return sp, dp


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][2]), "reasons": sorted(reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_reference(steps, warmup, shape_factor, n=16, ppc=16, seed=1234):
    """The reference's algorithm on the host cores: NumPy float64 restatement (oracle/), bounded sample n^3 x ppc."""
    from oracle import fixtures as fx, evolve as oevolve
    cfg = physical_setup(n, (1, 1, 1), ppc, shape_factor, "f64")
    sp, dp = fx.kernel_parameters(Nx=n, Ny=n, Nz=n, x_wind=cfg["wind"][0], y_wind=cfg["wind"][1], z_wind=cfg["wind"][2], dt=cfg["dt"],
                                  shape_factor=shape_factor, current_deposition="esirkepov", C=C_LIGHT, eps=EPS0, mu=MU0)
    tp, sc = fx.thermal_plasma(sp, dp, ppc_per_species=ppc // 2, vth=(0.05, 0.05 * math.sqrt(ME / MP)), seed=seed,
                               charge=cfg["charge"], mass=cfg["mass"], weight=cfg["weight"])
    z = fx.empty_tiled_vector
    fields = (z(sp, dp), z(sp, dp), z(sp, dp), fx.empty_tiled_scalar(sp, dp), fx.empty_tiled_scalar(sp, dp), (z(sp, dp), z(sp, dp)), None, False)
    npart = int(tp.active.sum())
    for _ in range(warmup):
        tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
    t0 = time.perf_counter()
    for _ in range(steps):
        tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
    el = time.perf_counter() - t0
    return {"value": npart * steps / el, "unit": "particle-steps/s", "cores": 1, "kind": "port", "particles": npart,
            "sample": f"{n}^3 cells x {ppc} ppc ({npart} particles), {steps} steps, NumPy float64 restatement of the reference "
                      f"(jax is not installable here); NumPy scatter-add is single-threaded"}, el / max(steps, 1)


def host_cores():
    """CPUs this process may actually use: the affinity mask, capped by a cgroup CPU quota when the container has one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path, parse in (("/sys/fs/cgroup/cpu.max", lambda t: t.split()),                                  # cgroup v2: "<quota|max> <period>"
                        ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", lambda t: (t.strip(), None))):            # cgroup v1
        try:
            quota, period = parse(open(path).read())
            if period is None:
                period = open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read().strip()
            if quota not in ("max", "-1") and int(quota) > 0 and int(period) > 0:
                n = min(n, max(1, int(quota) // int(period)))
            break
        except (OSError, ValueError):
            continue
    return max(1, n)


def cpu_reference_all_cores(steps, warmup, shape_factor, n=16, ppc=16, workers=0):
    """The CPU arm on every host core.  NumPy runs the restatement on one thread, so the cores are used the way the reference uses
    devices: one tile per worker (its `field_mesh`, ghost_cells.py:181-316) -- `workers` oracle processes, each stepping its own
    n^3 tile of the same plasma (seed 1234 + worker).  The tiles do not exchange halos, so this is an upper bound for a
    `workers`-tile CPU run.  value = all particles x steps / the slowest worker's time."""
    workers = int(workers) if workers else min(host_cores(), 128)
    if workers <= 1:
        return cpu_reference(steps, warmup, shape_factor, n=n, ppc=ppc)
    try:
        return _cpu_reference_workers(steps, warmup, shape_factor, n, ppc, workers)
    except Exception as exc:                     # a host that cannot spawn the workers still gets the one-process baseline
        base, sec = cpu_reference(steps, warmup, shape_factor, n=n, ppc=ppc)
        base["sample"] += f" [all-core run failed ({type(exc).__name__}: {str(exc)[:120]}); one process instead]"
        return base, sec


def _cpu_reference_workers(steps, warmup, shape_factor, n, ppc, workers):
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", "--steps", str(steps), "--warmup", str(warmup),
           "--shape-factor", str(shape_factor), "--cpu-n", str(n), "--ppc", str(ppc)]
    procs = [subprocess.Popen(cmd + ["--seed", str(1234 + w)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for w in range(workers)]
    results = []
    try:
        for pr in procs:
            out, err = pr.communicate(timeout=600)
            line = next((l for l in reversed(out.splitlines()) if l.startswith("{")), None)
            if pr.returncode != 0 or line is None:
                raise RuntimeError(f"cpu worker failed (rc={pr.returncode}): {err[-500:]}")
            results.append(json.loads(line))
    finally:
        for pr in procs:                          # only the processes started here, by handle
            if pr.poll() is None:
                pr.kill()
    npart = sum(r["particles"] for r in results)
    slowest = max(r["seconds"] for r in results)
    return {"value": npart * steps / slowest, "unit": "particle-steps/s", "cores": workers, "kind": "port",
            "sample": f"{workers} tiles of {n}^3 cells x {ppc} ppc ({npart} particles in all), {steps} steps, one process per host core, "
                      f"each the NumPy float64 restatement of the reference on its own tile without halo exchange (jax is not "
                      f"installable here; NumPy itself runs the step on one thread); slowest worker {slowest:.2f} s"}, slowest / max(steps, 1)


def run_cpu_worker(args):
    base, sec = cpu_reference(args.steps, args.warmup, args.shape_factor, n=args.cpu_n, ppc=args.ppc, seed=args.seed)
    _emit({"particles": base["particles"], "seconds": sec * args.steps})


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms = cpu_reference_all_cores(args.steps, args.warmup, args.shape_factor, n=args.cpu_n, ppc=args.ppc, workers=args.cpu_workers)
    line = {"impl": "reference", "metric": "particle-steps/sec (push+deposit+Yee)", "value": base["value"], "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": reference_config(args, base), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def workload_config(args, n_gpus):
    mesh = mesh_for(n_gpus)
    return {"workload": f"synthetic 3D periodic thermal plasma, {args.n}^3 cells per GPU x {args.ppc} ppc (2 species), "
                        f"{'Esirkepov' if getattr(args, 'deposition', 'esirkepov') == 'esirkepov' else 'j_from_rhov (' + args.current_filter + ' filter)'}"
                        f" + first-order Yee, shape_factor={args.shape_factor}, relativistic Boris, g=2",
            "cells_per_gpu": args.n ** 3, "ppc": args.ppc, "mesh": list(mesh), "sort_interval": args.sort_interval,
            "cache": "inputs >> L2 (particles 6.4 GB f32 at 256^3 x 16 ppc); no L2 flush needed"}


def reference_config(args, base):
    """What the CPU arm really runs: the same plasma (density, temperature, dx = lambda_D, dt, species, Esirkepov + Yee) on a bounded
    sample -- one n^3-cell tile per host core -- because the reference's formulation cannot hold the 256^3 case."""
    c = workload_config(args, args.gpus)
    c["workload"] = (f"bounded sample of the same synthetic 3D periodic thermal plasma: {base['cores']} independent tile(s) of {args.cpu_n}^3 cells "
                     f"x {args.ppc} ppc (2 species), Esirkepov + first-order Yee, shape_factor={args.shape_factor}, relativistic Boris, g=2, float64; "
                     f"the GPU arm runs {args.n}^3 cells per GPU")
    c["cells_per_gpu"] = None
    c["cells_per_tile"] = args.cpu_n ** 3
    c["tiles"] = base["cores"]
    c["mesh"] = None
    c["cache"] = "host run"
    return c


def mesh_for(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(n, (n, 1, 1))


# ------------------------------------------------------------------------------------------------ device workload
def device_plasma(cfg, sp, dp, n_local, moff, ppc, dtype, device, seed, cap_factor=1.0):
    """Synthetic thermal plasma generated directly on the device in the reference TiledParticles layout (one tile)."""
    import torch
    import pypic3d_b200 as pp
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n_per = n_local ** 3 * (ppc // 2)
    cap = int(math.ceil(n_per * cap_factor))          # reference layout: fixed slot capacity, inactive tail (particle_class.py:17-31)
    x = torch.zeros((1, 1, 1, 2, cap, 3), dtype=dtype, device=device)
    u = torch.zeros_like(x)
    for s in range(2):
        for a in range(3):
            lo = -cfg["wind"][a] / 2 + moff[a] * n_local * cfg["dx"]
            x[0, 0, 0, s, :n_per, a] = (lo + torch.rand(n_per, generator=g, device=device, dtype=torch.float64) * (n_local * cfg["dx"])).to(dtype)
            u[0, 0, 0, s, :n_per, a] = (torch.randn(n_per, generator=g, device=device, dtype=torch.float32) * cfg["vth"][s]).to(dtype)
    # keep positions strictly inside the local box after rounding to dtype
    for a in range(3):
        lo = -cfg["wind"][a] / 2 + moff[a] * n_local * cfg["dx"]
        hi = lo + n_local * cfg["dx"]
        x[..., :n_per, a].clamp_(min=lo, max=float(np.nextafter(np.float32(hi), np.float32(lo))) if dtype == torch.float32 else float(np.nextafter(hi, lo)))
    active = torch.zeros((1, 1, 1, 2, cap), dtype=torch.bool, device=device)
    active[..., :n_per] = True
    species = pp.SpeciesConfig(charge=np.array(cfg["charge"]), mass=np.array(cfg["mass"]), weight=np.array([cfg["weight"]] * 2),
                               update_x=np.ones((2, 3), bool), update_u=np.ones((2, 3), bool))
    return pp.TiledParticles(x=x, u=u, active=active), species


def zero_fields(n_local, dtype, device):
    import torch
    L = n_local + 4
    z = lambda: torch.zeros((1, 1, 1, L, L, L), dtype=dtype, device=device)
    v = lambda: (z(), z(), z())
    return (v(), v(), v(), z(), z(), (v(), v()), None, torch.tensor(False, device=device))


_REAL_STDOUT = None


def _capture_stdout():
    """Route everything libraries print on fd 1 (e.g. the NCCL version banner) to stderr; the JSON line is written to the
    real stdout at the end, so stdout carries exactly one line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def k1_traffic(args, dtype_name, variant):
    """DRAM bytes per K1 launch from the committed ncu capture of this configuration (profiles/r0N_k1_traffic*.json), or None."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not (f.endswith(".json") and "k1_traffic" in f):
            continue
        try:
            tj = json.load(open(os.path.join(pdir, f)))
            c = tj["config"]
        except (ValueError, KeyError, OSError):
            continue
        if (c.get("n"), c.get("ppc"), c.get("dtype"), c.get("shape_factor"), c.get("k1_variant", "global")) == (args.n, args.ppc, dtype_name, args.shape_factor, variant):
            best = (f, tj)
    return best


def run_leg(args, dtype_name, device, world, rank, local_rank, dist, *, e2e, check, clocks):
    """One dtype of the workload: build the plasma, warm up, time `steps` steps on the device, roofline of K1, optional
    end-to-end run and conservation check.  Returns the leg's dictionary (rank 0's view; timings are max over ranks)."""
    import torch
    from pypic3d_b200 import _lib
    from pypic3d_b200.simulation import Simulation
    n_gpus = world
    mesh = mesh_for(n_gpus)
    dtype = torch.float32 if dtype_name == "f32" else torch.float64
    cfg = physical_setup(args.n, mesh, args.ppc, args.shape_factor, dtype_name)
    sp, dp = make_params(cfg, args.n, mesh, args.shape_factor, deposition="direct" if args.deposition == "j_from_rhov" else "esirkepov",
                         current_filter=args.current_filter)
    moff = (rank // (mesh[1] * mesh[2]), (rank // mesh[2]) % mesh[1], rank % mesh[2])
    particles, species = device_plasma(cfg, sp, dp, args.n, moff, args.ppc, dtype, device, seed=1234 + rank,
                                       cap_factor=1.02 if world > 1 else 1.0)
    fields = zero_fields(args.n, dtype, device)
    if world > 1:
        from pypic3d_b200.distributed import DistributedHalo
        halo_factory = lambda p: DistributedHalo(p, dist.group.WORLD, device)
    else:
        halo_factory = None
    # track_ids=False: exported particles come back compacted in cell order instead of their original slots (any slot
    # order is a valid TiledParticles state; the reference itself re-slots particles when they migrate), so the state a
    # caller feeds back in the e2e loop is already nearly sorted.
    sim = Simulation(particles, species, fields, sp, dp, sort_interval=args.sort_interval, gmesh=mesh, moff=moff,
                     halo=halo_factory, capacity_factor=1.25 if world > 1 else 1.02, track_ids=False)
    n_local_particles = sim.n_particles()
    del particles, fields
    torch.cuda.empty_cache()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        sim.step(1)
    barrier()
    sampler = ClockSampler(local_rank) if (rank == 0 and clocks) else None
    if sampler:
        sampler.start()
    sim.k1_events = []
    sim.phase_events = {} if os.environ.get("PIC_BENCH_PHASES", "1") == "1" else None
    _lib.LAUNCHES = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    sim.step(args.steps)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = _lib.LAUNCHES
    k1_ms = [a.elapsed_time(b) for a, b in sim.k1_events]
    sim.k1_events = None
    phases = sim.phase_summary(args.steps) if sim.phase_events is not None else None
    sim.phase_events = None
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    # periodic boundaries conserve the global particle number, so the count taken right after construction (exact: freshly
    # sorted, dead slots excluded) is the number of particles advanced in every timed step
    npart = torch.tensor([float(n_local_particles)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(npart, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    total_particles = float(npart.item())
    value = total_particles * args.steps / (ms_total * 1e-3)
    overflow = sim.overflow()

    # ---- roofline of the dominant kernel (K1, one launch per species per step) and of the whole step
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    real = 4 if dtype_name == "f32" else 8
    k1_bytes_pp = 12 * real + (9.0 / args.ppc) * real        # r/w x,u + gather E,B (6) + J write (3) per cell
    step_bytes_pp = 12 * real + 1 + (24.0 / args.ppc) * real  # 55 B (f32) / 109 B (f64) at 16 ppc
    k1_avg_ms = float(np.mean(k1_ms)) if k1_ms else None
    per_launch_particles = n_local_particles / 2
    traffic, traffic_src = None, None
    tr = k1_traffic(args, dtype_name, sim.k1_variant)
    if tr is not None:
        tj = tr[1]
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * per_launch_particles / tj["particles_per_launch"]
        traffic_src = f"profiles/{tr[0]} (ncu dram__bytes_read.sum + dram__bytes_write.sum, per launch)"
    roof = None
    if k1_avg_ms:
        achieved = k1_bytes_pp * per_launch_particles / (k1_avg_ms * 1e-3) / 1e9
        kname = {"pair": "k_pair3d + k_pair_fixup (K1 v10: supercell E/B tiles in shared memory, two particles per thread in packed f32x2; "
                         "cell changers finished by the fix-up pass -- both launches are inside the timed launch pair",
                 "tile": "k_tile3d (K1 v9: supercell E/B tiles in shared memory"}.get(sim.k1_variant, "k_fused3d (K1 v8: global gather")
        roof = {"bound": "hbm", "kernel": kname + "; gather+push+deposit+move+BC, one species per launch)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": k1_bytes_pp * per_launch_particles, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": k1_bytes_pp, "avg_launch_ms": k1_avg_ms,
                "k1_share_of_step": float(np.sum(k1_ms)) / ms_total if ms_total else None,
                "avg_launch_ms_by_species": [float(np.mean(k1_ms[i::sim.S])) for i in range(sim.S)]}
    step_achieved = step_bytes_pp * (total_particles / n_gpus) * args.steps / (ms_total * 1e-3) / 1e9
    roof_step = {"bytes_per_particle_step": step_bytes_pp, "achieved": step_achieved, "peak": peak, "unit": "GB/s", "frac": step_achieved / peak}
    leg = {"dtype": dtype_name, "value": value, "unit": "particle-steps/s", "ms_per_step": ms_total / args.steps, "particles": total_particles,
           "overflow": overflow, "roofline": roof, "roofline_step": roof_step, "gpu_launches": launches, "k1_variant": sim.k1_variant,
           "k1_options_now": [sim._k1_options(i) for i in range(sim.S)], "k1_global_fallback_particles": int(sim.flags[2].item()),
           "clocks": sampler.summary() if sampler else None,
           "phases_ms": phases}      # rank 0's device time per step and phase (CUDA events between the phases of Simulation._step_once)

    # ---- end-to-end through the public API with HOST buffers (rank-local): H2D -> load_state -> step -> export -> D2H
    if e2e:
        leg["e2e"] = run_e2e(sim, args, device, world, dist if world > 1 else None)

    # ---- charge conservation on the measured configuration, at this size and this number of GPUs (one more step)
    if check:
        keys = ("continuity_residual_max", "continuity_scale", "gauss_drift_max", "gauss_scale")
        try:
            c = sim.conservation_step()
        except Exception as exc:          # the diagnostic must not take the measurement down with it (collectives stay matched below)
            c = {k: float("nan") for k in keys}
            c["error"] = repr(exc)[:300]
        if world > 1:
            for k in keys:
                tt = torch.tensor([c[k]], dtype=torch.float64, device=device)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                c[k] = float(tt.item())
        c["continuity_relative"] = c["continuity_residual_max"] / c["continuity_scale"] if c["continuity_scale"] else None
        c["gauss_drift_relative"] = c["gauss_drift_max"] / c["gauss_scale"] if c["gauss_scale"] else None
        c["what"] = ("one more step after the timed region: max |(rho_new - rho_old)/dt + div J| / max |(rho_new - rho_old)/dt| and "
                     "max |change of (div E - rho/eps)| / max(|rho|/eps), maxima over all ranks; rho, div J and div E are evaluated in "
                     "float64 from the state of the run (positions, J, E as the run holds them in its own dtype)")
        leg["check"] = c
    del sim
    torch.cuda.empty_cache()
    return leg


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--n", "--cells-per-axis", dest="n", type=int, default=256, help="cells per axis per GPU (use --cells-per-axis under torchrun: its parser claims --n)")
    ap.add_argument("--ppc", type=int, default=16)
    ap.add_argument("--shape-factor", type=int, default=1)
    ap.add_argument("--dtype", default="f32", choices=("f32", "f64"), help="dtype of the headline leg (the other one is reported under legs)")
    ap.add_argument("--deposition", default="esirkepov", choices=("esirkepov", "j_from_rhov"),
                    help="secondary configurations only (the headline metric is Esirkepov): the reference's default is j_from_rhov + bilinear")
    ap.add_argument("--current-filter", default="none", choices=("none", "bilinear"))
    ap.add_argument("--sort-interval", type=int, default=10)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-n", type=int, default=16)
    ap.add_argument("--cpu-workers", type=int, default=0, help="CPU arm: oracle processes (0 = one per host core, at most 128)")
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--seed", type=int, default=1234, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the charge-conservation check of the measured configuration")
    ap.add_argument("--no-second-leg", action="store_true", help="only the headline dtype (A/B runs)")
    ap.add_argument("--check", action="store_true", help=argparse.SUPPRESS)      # (round-1 flag: the check is on by default now)
    args = ap.parse_args()
    if args.cpu_worker:
        return run_cpu_worker(args)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pypic3d_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL's version/debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world
    other = "f64" if args.dtype == "f32" else "f32"
    head = run_leg(args, args.dtype, device, world, rank, local_rank, dist, e2e=not args.no_e2e, check=not args.no_check, clocks=True)
    legs = {args.dtype: head}
    if not args.no_second_leg:
        legs[other] = run_leg(args, other, device, world, rank, local_rank, dist, e2e=False, check=not args.no_check, clocks=True)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_all_cores(3, 1, args.shape_factor, n=args.cpu_n, ppc=args.ppc, workers=args.cpu_workers)

    if rank == 0:
        line = {"metric": "particle-steps/sec (push+deposit+Yee)", "value": head["value"], "unit": "particle-steps/s", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": workload_config(args, n_gpus), "particles": head["particles"], "overflow": head["overflow"],
                "roofline": head["roofline"], "roofline_step": head["roofline_step"], "cpu_baseline": cpu_base, "e2e": head.get("e2e"),
                "gpu_launches": head["gpu_launches"], "k1_variant": head["k1_variant"], "k1_options_now": head["k1_options_now"],
                "k1_global_fallback_particles": head["k1_global_fallback_particles"], "clocks": head["clocks"],
                "check": head.get("check"),
                "legs": {k: {kk: vv for kk, vv in v.items() if kk != "e2e"} for k, v in legs.items()},
                "legs_note": "the reference computes in float64 (PyPIC3D/__main__.py:186); legs.f64 is the same workload in the reference's dtype "
                             "(109 B per particle-step), legs.f32 the throughput mode (55 B); `value` is legs[dtype]"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(sim, args, device, world, dist):
    """Same metric through the reference-facing API with host buffers: every step copies the TiledParticles + E,B,J pytrees
    from pinned host memory, runs `load_state -> sort -> step -> export_state`, and copies the result pytrees back."""
    import torch
    parts, fields = sim.export_state()
    host = {"x": parts.x.cpu().pin_memory(), "u": parts.u.cpu().pin_memory(), "a": parts.active.cpu().pin_memory(),
            "F": [[c.cpu().pin_memory() for c in fields[k]] for k in range(3)]}
    dev = {"x": parts.x, "u": parts.u, "a": parts.active, "F": [[c for c in fields[k]] for k in range(3)]}
    h2d = host["x"].numel() * host["x"].element_size() * 2 + host["a"].numel() + sum(c.numel() * c.element_size() for k in host["F"] for c in k)
    d2h = h2d
    import pypic3d_b200 as pp

    def one():
        dev["x"].copy_(host["x"], non_blocking=True); dev["u"].copy_(host["u"], non_blocking=True); dev["a"].copy_(host["a"], non_blocking=True)
        for k in range(3):
            for c in range(3):
                dev["F"][k][c].copy_(host["F"][k][c], non_blocking=True)
        sim.load_state(pp.TiledParticles(dev["x"], dev["u"], dev["a"]), (tuple(dev["F"][0]), tuple(dev["F"][1]), tuple(dev["F"][2])))
        sim.step(1)
        p2, f2 = sim.export_state(out=(dev["x"], dev["u"], dev["a"]))
        host["x"].copy_(p2.x, non_blocking=True); host["u"].copy_(p2.u, non_blocking=True); host["a"].copy_(p2.active, non_blocking=True)
        for k in range(3):
            for c in range(3):
                host["F"][k][c].copy_(f2[k][c], non_blocking=True)
    one()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        one()
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    npart = torch.tensor([float(parts.active.sum().item())], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.all_reduce(npart, op=dist.ReduceOp.SUM)
    return {"value": float(npart.item()) * args.e2e_steps / float(el.item()), "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps, "ms_per_step": float(el.item()) / args.e2e_steps * 1e3,
            "path": "pinned host TiledParticles+E,B,J -> H2D -> Simulation.load_state -> sort -> step -> export_state -> D2H"}


if __name__ == "__main__":
    main()
