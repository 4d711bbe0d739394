"""Debug: where is the continuity residual of Simulation.conservation_step large?  (bench workload, one GPU)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pypic3d_b200.simulation import Simulation
from pypic3d_b200 import ops

n = int(sys.argv[1]); steps = int(sys.argv[2]); dtype_name = sys.argv[3]
dtype = torch.float64 if dtype_name == "f64" else torch.float32
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
mesh = bench.mesh_for(world)
halo = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    from pypic3d_b200.distributed import DistributedHalo
    halo = lambda p: DistributedHalo(p, dist.group.WORLD, dev)
moff = (rank // (mesh[1] * mesh[2]), (rank // mesh[2]) % mesh[1], rank % mesh[2])
cfg = bench.physical_setup(n, mesh, 16, 1, dtype_name)
sp, dp = bench.make_params(cfg, n, mesh, 1)
particles, species = bench.device_plasma(cfg, sp, dp, n, moff, 16, dtype, dev, seed=1234 + rank, cap_factor=1.02)
fields = bench.zero_fields(n, dtype, dev)
sim = Simulation(particles, species, fields, sp, dp, sort_interval=10, capacity_factor=1.25, track_ids=False, gmesh=mesh, moff=moff, halo=halo)
del particles, fields
sim.step(steps)
p = sim.p
rho0 = sim.charge_density()
a0 = [int(torch.isnan(s_.buf[s_.cur][0][:int(s_.n_dev.item())]).sum().item()) for s_ in sim.species]
n0 = [int(s_.n_dev.item()) for s_ in sim.species]
g0 = sim.gauss_residual(rho0)
sim.step(1)
rho1 = sim.charge_density()
if sim._J_ghosts_stale:
    sim.halo.refresh_(sim.J, tuple(p.particle_bc)); sim._J_ghosts_stale = False
inv_dt = 1.0 / float(p.dt)
pd, _, cast = sim._diag()
cont = ops.div_residual(pd, [cast(c) for c in sim.J], rho1, inv_dt, rho0, -inv_dt)
g1 = sim.gauss_residual(rho1)
g = int(p.g)
I = (0, 0, 0, slice(g, -g), slice(g, -g), slice(g, -g))
c = cont[I].abs(); scale = ((rho1[I] - rho0[I]) * inv_dt).abs().max().item()
print("rank", rank, "n", n, "steps", steps, dtype_name, "slots", n0, "dead", a0)
print("cont max", c.max().item(), "scale", scale, "rel", c.max().item() / scale)
print("gauss drift max", (g1[I] - g0[I]).abs().max().item(), "implied cont", (g1[I] - g0[I]).abs().max().item() * float(p.eps) / float(p.dt))
bad = (c > 1e-3 * scale).nonzero()
print("bad cells", bad.shape[0], "of", c.numel())
if bad.shape[0]:
    print("bad min idx", bad.min(0).values.tolist(), "max idx", bad.max(0).values.tolist())
    print("first bad", bad[:8].tolist())
    dj = ops.div_residual(pd, [cast(c) for c in sim.J], rho1, 0.0, rho0, 0.0)
    for b_ in bad[:4].tolist():
        ix, iy, iz = b_[0] + g, b_[1] + g, b_[2] + g
        print(" cell", b_, "cont", cont[0, 0, 0, ix, iy, iz].item(), "drho/dt", ((rho1 - rho0) * inv_dt)[0, 0, 0, ix, iy, iz].item(), "divJ", dj[0, 0, 0, ix, iy, iz].item(),
              "rho0", rho0[0, 0, 0, ix, iy, iz].item(), "rho1", rho1[0, 0, 0, ix, iy, iz].item(), "g0", g0[0, 0, 0, ix, iy, iz].item(), "g1", g1[0, 0, 0, ix, iy, iz].item())
    # particles near the bad cell before/after
    parts, _ = sim.export_state(fields=False)
    x = parts.x[0, 0, 0]; act = parts.active[0, 0, 0]
    dx_ = float(p.dx)
    lo = [-float(p.wind[a]) / 2 + moff[a] * n * dx_ for a in range(3)]
    b0 = bad[0].tolist()
    for s_ in range(2):
        xs = x[s_][act[s_]]
        cell = torch.stack([torch.floor((xs[:, a] - lo[a]) / dx_) for a in range(3)], 1)
        near = ((cell[:, 0] <= 0) | (cell[:, 0] >= n - 1)) & ((cell[:, 1] - b0[1]).abs() <= 1) & ((cell[:, 2] - b0[2]).abs() <= 1)
        xn = xs[near]
        print(" species", s_, "near", xn.shape[0], [((xn[i, 0].item() - lo[0]) / dx_, (xn[i, 1].item() - lo[1]) / dx_, (xn[i, 2].item() - lo[2]) / dx_) for i in range(min(xn.shape[0], 12))])
print("sum rho0", rho0[I].sum().item(), "sum rho1", rho1[I].sum().item(), "sum |rho|", rho0[I].abs().sum().item())

if world > 1:
    dist.barrier(); dist.destroy_process_group()
