"""Drop-ins for PyPIC3D/solvers/electrostatic_yee.py: the conjugate-gradient Poisson solve, the
constant-potential wall treatment and the centred electrostatic gradient, as CUDA kernels behind the C ABI
(csrc/kernels_poisson.cu).  Like the reference (initialization.py:251-254), electrostatic runs use ONE tile covering the
whole domain."""
import ctypes

import torch

from .. import _lib, ops
from .._lib import check
from ..boundary_conditions.ghost_cells import update_tiled_vector_ghost_cells
from ..deposition.rho import compute_rho
from ..utilities.filters import digital_filter

__all__ = ["solve_poisson_with_conjugate_gradient", "calculate_tiled_electrostatic_fields", "apply_tiled_phi_constant_boundaries"]


def _tile_params(field, static_parameters, dynamic_parameters):
    """PicParams of the single ghosted tile `field` (..., Lx, Ly, Lz)."""
    p = ops.params_for(static_parameters, dynamic_parameters, None, field)
    L = tuple(int(v) for v in field.shape[-3:])
    g = int(static_parameters.guard_cells)
    want = tuple(int(n) + 2 * g for n in (dynamic_parameters.Nx, dynamic_parameters.Ny, dynamic_parameters.Nz))
    if L != want or any(int(m) != 1 for m in p.mesh):
        raise ValueError(f"the electrostatic solve acts on one tile of shape {want}; got {L} with tile mesh {tuple(p.mesh)}")
    return p


def solve_poisson_with_conjugate_gradient(rho, phi, static_parameters, dynamic_parameters, tol=1e-12, max_iter=5000,
                                          return_iterations=False, check_every=16):
    """electrostatic_yee.py:71-156.  rho, phi: (Nx+2g, Ny+2g, Nz+2g) CUDA tensors; returns the new phi (inputs untouched)."""
    rho = ops._chk(rho, "rho")
    phi = ops._chk(phi, "phi", rho.dtype).clone()
    p = _tile_params(phi, static_parameters, dynamic_parameters)
    work = [torch.empty_like(phi) for _ in range(3)]
    scal = torch.zeros(8, dtype=torch.float64, device=phi.device)
    iters = ctypes.c_int(0)
    check(_lib.lib().pic_poisson_cg(ctypes.byref(p), ops._p(rho), ops._p(phi), ops._p(work[0]), ops._p(work[1]), ops._p(work[2]),
                                    ops._p(scal), float(tol), int(max_iter), int(check_every), ctypes.byref(iters), ops._stream()),
          "pic_poisson_cg")
    return (phi, iters.value) if return_iterations else phi


def apply_tiled_phi_constant_boundaries(field_tiles, static_parameters, dynamic_parameters, g=None):
    """electrostatic_yee.py:20-37 on a (1,1,1,Lx,Ly,Lz) scalar tile: constant-potential exterior ghosts on conducting walls,
    a plain field-BC ghost refresh otherwise.  Returns a new tensor."""
    out = ops._chk(field_tiles, "field_tiles").clone()
    p = _tile_params(out, static_parameters, dynamic_parameters)
    check(_lib.lib().pic_phi_boundaries(ctypes.byref(p), ops._p(out), ops._stream()), "pic_phi_boundaries")
    return out


def _centered_tiled_electrostatic_gradient(phi_tiles, static_parameters, dynamic_parameters, g):
    """electrostatic_yee.py:212-246: E = -grad(phi) on the interior, then the vector ghost refresh."""
    phi_tiles = apply_tiled_phi_constant_boundaries(phi_tiles, static_parameters, dynamic_parameters, g)
    p = _tile_params(phi_tiles, static_parameters, dynamic_parameters)
    E = [torch.zeros_like(phi_tiles) for _ in range(3)]
    check(_lib.lib().pic_gradient_neg(ctypes.byref(p), ops._p(phi_tiles), ops._v(E), ops._stream()), "pic_gradient_neg")
    return update_tiled_vector_ghost_cells(tuple(E), static_parameters, int(g), _inplace=True, _dyn=dynamic_parameters)


def calculate_tiled_electrostatic_fields(static_parameters, dynamic_parameters, particles, species_config, rho_tiles, phi_tiles):
    """electrostatic_yee.py:249-276: rho deposition -> CG Poisson -> wall treatment -> digital filter -> E = -grad(phi)."""
    g = int(static_parameters.guard_cells)
    rho_tiles = compute_rho(particles, species_config, rho_tiles, static_parameters, dynamic_parameters)
    phi = solve_poisson_with_conjugate_gradient(rho_tiles[0, 0, 0], phi_tiles[0, 0, 0], static_parameters, dynamic_parameters)
    phi_tiles = phi[None, None, None]
    phi_tiles = apply_tiled_phi_constant_boundaries(phi_tiles, static_parameters, dynamic_parameters, g)
    phi_tiles = digital_filter(phi_tiles, dynamic_parameters.alpha, num_guard_cells=g)
    phi_tiles = apply_tiled_phi_constant_boundaries(phi_tiles, static_parameters, dynamic_parameters, g)
    E_tiles = _centered_tiled_electrostatic_gradient(phi_tiles, static_parameters, dynamic_parameters, g)
    return E_tiles, phi_tiles, rho_tiles
