"""Drop-ins for the step-adjacent helpers of PyPIC3D/utils.py: add_external_fields:205, compute_energy:108,
compute_total_momentum:190, courant_condition:761."""
import torch

from . import ops


def add_external_fields(E, B, external_fields):
    """Particles see evolved + external fields (utils.py:205-216).  Elementwise add: torch plumbing, not a hot op
    (the resident fast path gathers the external fields inside K1 instead of materialising the sum)."""
    external_E, external_B = external_fields
    return tuple(e + x for e, x in zip(E, external_E)), tuple(b + x for b, x in zip(B, external_B))


def compute_energy(particles, E, B, static_parameters, dynamic_parameters, species_config=None):
    from .boundary_conditions.ghost_cells import _halo_params
    p = _halo_params(E[0], static_parameters, static_parameters.guard_cells)
    sc = ops._lib._scalar
    acc = torch.zeros(2, dtype=torch.float64, device=E[0].device)
    for c in range(3):
        ops.sum_squares_interior(p, ops._chk(E[c], "E"), acc[0:1])
        ops.sum_squares_interior(p, ops._chk(B[c], "B"), acc[1:2])
    dV = float(sc(dynamic_parameters.dx)) * float(sc(dynamic_parameters.dy)) * float(sc(dynamic_parameters.dz))
    e_energy = 0.5 * float(sc(dynamic_parameters.eps)) * acc[0] * dV
    b_energy = 0.5 / float(sc(dynamic_parameters.mu)) * acc[1] * dV
    pp = ops.params_for(static_parameters, dynamic_parameters, species_config, particles.x,
                        mesh=tuple(particles.x.shape[:3]), gmesh=tuple(particles.x.shape[:3]))
    ke = ops.particle_energy(pp, particles.u, particles.active)[0] if particles.x.shape[4] > 0 else torch.zeros((), dtype=torch.float64, device=E[0].device)
    return e_energy, b_energy, ke


def compute_total_momentum(particles, species_config=None, static_parameters=None, dynamic_parameters=None):
    p = ops._lib.PicParams()
    x = particles.x
    p.dtype = 0 if x.dtype == ops.F32 else 1
    for a in range(3):
        p.mesh[a] = p.gmesh[a] = int(x.shape[a])
    S = int(x.shape[3])
    p.n_species = S
    p.C = 1.0
    mass = ops._lib._to_numpy(species_config.mass).reshape(-1)
    weight = ops._lib._to_numpy(species_config.weight).reshape(-1)
    for s in range(S):
        p.mass[s], p.weight[s] = float(mass[s]), float(weight[s])
    return ops.particle_energy(p, particles.u, particles.active)[1]


def courant_condition(courant_number, dx, dy, dz, dynamic_parameters):
    inv = sum(1 / d for d, n in zip((dx, dy, dz), (dynamic_parameters.Nx, dynamic_parameters.Ny, dynamic_parameters.Nz)) if n > 1)
    return courant_number / (dynamic_parameters.C * inv)
