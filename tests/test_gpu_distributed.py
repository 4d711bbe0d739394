"""-m gpu multi-GPU parity: the resident path on a (2,1,1) / (2,2,1) rank mesh (NCCL halo exchange + particle migration)
against the oracle's in-process tile mesh.  Needs >= 2 (or 4) CUDA devices; skipped (hardware absent) otherwise -- run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_distributed.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mesh, sf, pbc, steps, q, tile=(8, 6, 4), expect_variant=None, dtype=None):
    import torch.distributed as dist
    from oracle import evolve as oevolve
    from tests import gpu_util as gu
    from tests.cases import make_case, make_fields
    from pypic3d_b200.simulation import Simulation
    from pypic3d_b200.distributed import DistributedHalo, coords_of
    import pypic3d_b200 as pp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
      try:
          N = tuple(mesh[a] * tile[a] for a in range(3))
          sp, dp, tp, sc, E, B = make_case(N, tile, sf, current_deposition="esirkepov", particle_boundary_conditions=pbc, n=400,
                                           capacity=4.0, vmax=0.3, dt=0.04)
          fields = make_fields(sp, dp)
          c = coords_of(rank, mesh)
          sl = (slice(c[0], c[0] + 1), slice(c[1], c[1] + 1), slice(c[2], c[2] + 1))
          ps, pd = gu.to_pkg_params(sp, dp)
          t = lambda a: gu.tt(np.ascontiguousarray(a[sl]), dtype, dev=dev) if (dtype is not None and a.dtype.kind == "f") else gu.tt(np.ascontiguousarray(a[sl]), dev=dev)
          v = lambda F: tuple(t(x) for x in F)
          parts = pp.TiledParticles(t(tp.x), t(tp.u), t(tp.active))
          f8 = (v(fields[0]), v(fields[1]), v(fields[2]), t(fields[3]), t(fields[4]), (v(fields[5][0]), v(fields[5][1])), None,
                torch.tensor(False, device=dev))
          sim = Simulation(parts, gu.species_to_pkg(sc), f8, ps, pd, sort_interval=2, gmesh=mesh, moff=c, capacity_factor=4.0,
                           halo=lambda p: DistributedHalo(p, None, dev))
          if expect_variant is not None:
              assert sim.k1_variant == expect_variant, sim.k1_variant
          for _ in range(steps):
              tp, fields = oevolve.time_loop_electrodynamic(tp, sc, fields, sp, dp)
          sim.step(steps)
          gp, gf = sim.export_state()
          err = 0.0
          for k in range(3):
              for a, b in zip(gf[k], fields[k]):
                  ref = b[sl]
                  err = max(err, float(np.abs(gu.npy(a) - ref).max()) / max(1.0, float(np.abs(ref).max())))
          got = gu.sorted_active(gu.npy(gp.x), gu.npy(gp.u), gu.npy(gp.active))
          want = gu.sorted_active(tp.x[sl], tp.u[sl], tp.active[sl])
          perr = float(np.abs(got - want).max()) if got.shape == want.shape and got.size else (0.0 if got.shape == want.shape else 1e9)
          q.put((rank, err, perr, got.shape[0], want.shape[0], sim.overflow(), bool(fields[7])))
      except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mesh", [(2, 1, 1), (1, 1, 2), (2, 2, 1)])
@pytest.mark.parametrize("sf", (1, 2))
@pytest.mark.parametrize("pbc", [(0, 0, 0), (2, 0, 1)])
def test_distributed_resident_matches_oracle(mesh, sf, pbc):
    world = mesh[0] * mesh[1] * mesh[2]
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mesh, sf, pbc, 4, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for r in res:
        assert len(r) == 7, r[1]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for rank, err, perr, n_got, n_want, ovf, oovf in res:
        assert n_got == n_want, (rank, n_got, n_want)
        assert err < 1e-11, (rank, err)
        assert perr < 1e-11, (rank, perr)
        assert ovf == oovf


@pytest.mark.parametrize("mesh", [(2, 1, 1), (1, 1, 2)])
@pytest.mark.parametrize("pbc", [(0, 0, 0), (2, 0, 1)])
@pytest.mark.parametrize("variant", ("pair", "tile"))
def test_distributed_tile_kernel_matches_oracle(mesh, pbc, variant, monkeypatch):
    """K1 v10 / v9 (supercell tiles) on a split domain: leaver packets, the appended-slot tail pass and the non-periodic move."""
    monkeypatch.setenv("PIC_K1_VARIANT", variant)
    world = mesh[0] * mesh[1] * mesh[2]
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mesh, 1, pbc, 5, q, (8, 8, 4), variant)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for r in res:
        assert len(r) == 7, r[1]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for rank, err, perr, n_got, n_want, ovf, oovf in res:
        assert n_got == n_want, (rank, n_got, n_want)
        assert err < 1e-11, (rank, err)
        assert perr < 1e-11, (rank, perr)
        assert ovf == oovf


@pytest.mark.parametrize("dtype", ("f64", "f32"))
def test_distributed_2x2x2_matches_oracle(dtype, monkeypatch):
    """The mesh the 8-GPU scaling run uses: every rank has a neighbour on every axis (faces, edges and corners all cross ranks), K1
    v10 with leaver packets in all 26 directions, in the reference's dtype and in the throughput dtype."""
    mesh, world = (2, 2, 2), 8
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    td = torch.float64 if dtype == "f64" else torch.float32
    procs = [ctx.Process(target=_worker, args=(r, world, port, mesh, 1, (0, 0, 0), 5, q, (8, 8, 4), "pair", td)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for r in res:
        assert len(r) == 7, r[1]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    tol = 1e-11 if dtype == "f64" else 2e-4
    for rank, err, perr, n_got, n_want, ovf, oovf in res:
        assert n_got == n_want, (rank, n_got, n_want)
        assert err < tol, (rank, err)
        assert perr < tol, (rank, perr)
        assert ovf == oovf
