"""Helpers shared by the -m gpu parity tests: oracle (NumPy) pytrees <-> pypic3d_b200 (torch CUDA) pytrees."""
import numpy as np
import torch

import pypic3d_b200 as pp
from pypic3d_b200.parameters import StaticParameters, DynamicParameters, GridParameters


def require_cuda():
    # -m gpu tests must FAIL (not skip) without a device: a silent skip would hide a CPU fallback.
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda:0")


def to_pkg_params(sp, dp):
    s = StaticParameters(**sp._asdict())
    d = DynamicParameters(**{**dp._asdict(), "grids": GridParameters(**dp.grids._asdict())})
    return s, d


def tt(a, dtype=torch.float64, dev="cuda:0"):
    a = np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        return torch.from_numpy(a).to(dev)
    return torch.from_numpy(a).to(dev).to(dtype).contiguous()


def particles_to_gpu(tp, dtype=torch.float64):
    return pp.TiledParticles(x=tt(tp.x, dtype), u=tt(tp.u, dtype), active=tt(tp.active))


def species_to_pkg(sc):
    return pp.SpeciesConfig(*[np.asarray(v) for v in sc])


def vec_to_gpu(F, dtype=torch.float64):
    return tuple(tt(c, dtype) for c in F)


def npy(t):
    return t.detach().cpu().numpy().astype(np.float64) if t.dtype != torch.bool else t.detach().cpu().numpy()


def fields_to_gpu(fields, dtype=torch.float64):
    E, B, J, rho, phi, ext, pml, ovf = fields
    return (vec_to_gpu(E, dtype), vec_to_gpu(B, dtype), vec_to_gpu(J, dtype), tt(rho, dtype), tt(phi, dtype),
            (vec_to_gpu(ext[0], dtype), vec_to_gpu(ext[1], dtype)), None, torch.tensor(bool(ovf), device="cuda:0"))


def assert_close(actual, expected, tol, what=""):
    a = npy(actual) if isinstance(actual, torch.Tensor) else np.asarray(actual)
    e = np.asarray(expected, dtype=np.float64)
    scale = max(1.0, float(np.abs(e).max())) if e.size else 1.0
    err = float(np.abs(a - e).max()) if e.size else 0.0
    assert err <= tol * scale, f"{what}: max abs err {err:.3e} > {tol:.1e} * {scale:.3e}"


def sorted_active(tp_x, tp_u, tp_active):
    """Active particles as rows sorted lexicographically (for layouts that do not keep slot identity)."""
    a = np.asarray(tp_active).reshape(-1)
    x = np.asarray(tp_x).reshape(-1, 3)[a]
    u = np.asarray(tp_u).reshape(-1, 3)[a]
    rows = np.concatenate([x, u], axis=1)
    key = np.lexsort(np.round(rows, 9).T[::-1])
    return rows[key]
