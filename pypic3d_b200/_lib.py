"""ctypes binding of libpic_b200.so (the C ABI declared in include/pic_b200.h).

There is NO CPU fallback: if the shared library cannot be built/loaded, or a tensor is not on a CUDA device, the
operators raise.  PyTorch is used only for device memory and streams; every computation on the hot path is a
hand-written sm_100a kernel behind the C ABI.
"""
import ctypes
import hashlib
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("PIC_B200_LIB", os.path.join(_HERE, "libpic_b200.so"))   # override: A/B builds for profiling
SOURCES = ["kernels_particles_ref.cu", "kernels_fields.cu", "kernels_fast.cu", "kernels_pair.cu", "kernels_poisson.cu", "microbench.cu"]
HEADERS = ["pic_common.cuh", "pic_math.cuh", "pic_slots.cuh", "pic_pair.cuh", "pic_tma.cuh"]
MAX_SPECIES = 16

VERSION_SOURCE = "kernels_fields.cu"      # the unit that defines pic_version(): compiled with -DPIC_SOURCE_HASH=...
COMPILE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]
NVCC_FLAGS = COMPILE_FLAGS + ["-shared"]   # (one-shot form: nvcc NVCC_FLAGS -o lib.so *.cu)


class PicParams(ctypes.Structure):
    """Mirror of `struct PicParams` (include/pic_b200.h)."""
    _fields_ = [
        ("dtype", ctypes.c_int32), ("shape_factor", ctypes.c_int32), ("pusher", ctypes.c_int32), ("g", ctypes.c_int32),
        ("mesh", ctypes.c_int32 * 3), ("gmesh", ctypes.c_int32 * 3), ("moff", ctypes.c_int32 * 3),
        ("tile", ctypes.c_int32 * 3), ("field_bc", ctypes.c_int32 * 3), ("particle_bc", ctypes.c_int32 * 3),
        ("n_species", ctypes.c_int32), ("pad0", ctypes.c_int32),
        ("dt", ctypes.c_double), ("dx", ctypes.c_double), ("dy", ctypes.c_double), ("dz", ctypes.c_double),
        ("wind", ctypes.c_double * 3),
        ("C", ctypes.c_double), ("eps", ctypes.c_double), ("mu", ctypes.c_double), ("alpha", ctypes.c_double),
        ("center0", ctypes.c_double * 3), ("vertex0", ctypes.c_double * 3),
        ("charge", ctypes.c_double * MAX_SPECIES), ("mass", ctypes.c_double * MAX_SPECIES),
        ("weight", ctypes.c_double * MAX_SPECIES),
        ("update_x", (ctypes.c_uint8 * 3) * MAX_SPECIES), ("update_u", (ctypes.c_uint8 * 3) * MAX_SPECIES),
    ]


class PicSoA(ctypes.Structure):
    """Mirror of `struct PicSoA`."""
    _fields_ = [("comp", ctypes.c_void_p * 6), ("id", ctypes.c_void_p), ("cap", ctypes.c_int64), ("n", ctypes.c_int64),
                ("n_dev", ctypes.c_void_p)]


class PicLeave(ctypes.Structure):
    """Mirror of `struct PicLeave`."""
    _fields_ = [("buf", ctypes.c_void_p), ("row_off", ctypes.c_int32 * 27), ("cap", ctypes.c_int32 * 27)]


def source_hash():
    """sha256 over every CUDA source, header and the compile flags: identifies the build the loaded library must come from."""
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(_HERE, "csrc", f), "rb") as fh:
            h.update(f.encode()); h.update(fh.read())
    with open(os.path.join(_ROOT, "include", "pic_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def built_hash(path=None):
    """Source hash recorded next to a built library (written by build(); checked against pic_version() when it is loaded)."""
    try:
        with open((path or LIB_PATH) + ".hash") as fh:
            return fh.read().strip()
    except OSError:
        return None


def needs_build():
    if "PIC_B200_LIB" in os.environ:
        return False
    return not os.path.exists(LIB_PATH) or built_hash() != source_hash()


def _includes(fname, seen=None):
    """Project headers a translation unit pulls in (recursive scan of #include "..." inside csrc/)."""
    seen = set() if seen is None else seen
    path = os.path.join(_HERE, "csrc", fname)
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if line.startswith("#include \""):
                inc = os.path.basename(line.split("\"")[1])
                if inc not in seen and os.path.exists(os.path.join(_HERE, "csrc", inc)):
                    seen.add(inc)
                    _includes(inc, seen)
    return seen


def _tu_hash(src, extra_flags=()):
    h = hashlib.sha256()
    for f in [src] + sorted(_includes(src)):
        with open(os.path.join(_HERE, "csrc", f), "rb") as fh:
            h.update(f.encode()); h.update(fh.read())
    with open(os.path.join(_ROOT, "include", "pic_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(list(COMPILE_FLAGS) + list(extra_flags)).encode())
    return h.hexdigest()[:16]


def build(force=False, verbose=False, extra_flags=(), out=None, ptxas_verbose=False):
    """Compile every CUDA translation unit for sm_100a into pypic3d_b200/libpic_b200.so (in-tree).

    One nvcc process per translation unit, in parallel; objects are cached under csrc/_obj keyed by the hash of the unit, the
    shared headers and the flags, so editing one kernel file recompiles one unit.  The hash of ALL sources is compiled into
    pic_version() and written beside the library; lib() refuses a library whose hash differs from the sources it sits next to.
    Returns the library path (and, with ptxas_verbose, the ptxas -v log as second value)."""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not needs_build() and not extra_flags:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objdir = os.path.join(_HERE, "csrc", "_obj")
    os.makedirs(objdir, exist_ok=True)
    shash = source_hash() + ("+" + hashlib.sha256(" ".join(extra_flags).encode()).hexdigest()[:6] if extra_flags else "")
    jobs, objs, logs = [], [], {}
    for src in SOURCES:
        flags = list(extra_flags)
        if src == VERSION_SOURCE:
            flags = flags + [f'-DPIC_SOURCE_HASH="{shash}"']
        obj = os.path.join(objdir, f"{os.path.splitext(src)[0]}.{_tu_hash(src, flags)}.o")
        objs.append(obj)
        if force or ptxas_verbose or not os.path.exists(obj):
            cmd = [nvcc] + COMPILE_FLAGS + (["-Xptxas", "-v"] if ptxas_verbose else []) + flags + ["-c", os.path.join(_HERE, "csrc", src), "-o", obj]
            jobs.append((src, cmd))
    if verbose:
        for _, cmd in jobs:
            print(" ".join(cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr[-4000:]}")
        return src, r.stderr
    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for src, log in ex.map(run, jobs):
                logs[src] = log
    r = subprocess.run([nvcc] + LINK_FLAGS + ["-o", out] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    with open(out + ".hash", "w") as fh:
        fh.write(shash + "\n")
    # drop cached objects of older source versions (keep the ones just linked and anything built in the last day by an A/B)
    keep = set(objs)
    for f in os.listdir(objdir):
        fp = os.path.join(objdir, f)
        if fp not in keep and f.endswith(".o") and time.time() - os.path.getmtime(fp) > 86400:
            os.remove(fp)
    if ptxas_verbose:
        return out, "\n".join(logs.get(s, "") for s in SOURCES)
    return out


_LIB = None
LAUNCHES = 0   # kernels launched through the C ABI since last reset (bench.py "gpu_launches")
KERNELS_PER_CALL = {"pic_poisson_cg": 0, "pic_phi_boundaries": 3, "pic_halo_fold_axis": 2, "pic_sort_scan": 3, "pic_sort_blocked_offsets": 5, "pic_sort_blocked_finish": 2, "pic_retile": 2, "pic_microbench": 0, "pic_params_size": 0,
                    "pic_version": 0}


class _Counted:
    __slots__ = ("fn", "k")

    def __init__(self, fn, k):
        self.fn, self.k = fn, k

    def __call__(self, *args):
        global LAUNCHES
        LAUNCHES += self.k
        return self.fn(*args)


class _Namespace:
    pass

_VP = ctypes.c_void_p
_PP = ctypes.POINTER(PicParams)
_I64 = ctypes.c_int64
_INT = ctypes.c_int
_DBL = ctypes.c_double
_V3 = ctypes.POINTER(ctypes.c_void_p)
_SOA = ctypes.POINTER(PicSoA)
_LEAVE = ctypes.POINTER(PicLeave)

# name -> argtypes; every symbol declared in include/pic_b200.h
SIGNATURES = {
    "pic_push": [_PP, _VP, _VP, _VP, _VP, _I64, _V3, _V3, _VP],
    "pic_deposit_esirkepov": [_PP, _VP, _VP, _VP, _I64, _V3, _VP],
    "pic_deposit_direct": [_PP, _VP, _VP, _VP, _I64, _V3, _VP],
    "pic_deposit_rho": [_PP, _VP, _VP, _I64, _VP, _VP],
    "pic_move": [_PP, _VP, _VP, _VP, _VP, _I64, _DBL, _VP],
    "pic_retile": [_PP, _VP, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _VP],
    "pic_update_E": [_PP, _V3, _V3, _V3, _VP],
    "pic_update_B": [_PP, _V3, _V3, _VP],
    "pic_yee_fused": [_PP, _V3, _V3, _V3, _V3, _V3, _VP],
    "pic_filter": [_PP, _INT, _DBL, _VP, _VP, _VP],
    "pic_halo_refresh_axis": [_PP, _INT, _INT, _INT, _V3, _VP],
    "pic_halo_fold_axis": [_PP, _INT, _INT, _INT, _V3, _VP],
    "pic_zero_wall": [_PP, _INT, _VP, _VP],
    "pic_poisson_cg": [_PP, _VP, _VP, _VP, _VP, _VP, _VP, _DBL, _INT, _INT, ctypes.POINTER(ctypes.c_int), _VP],
    "pic_phi_boundaries": [_PP, _VP, _VP],
    "pic_constant_wall": [_PP, _INT, _VP, _VP],
    "pic_gradient_neg": [_PP, _VP, _V3, _VP],
    "pic_pack_planes": [_PP, _INT, _INT, _INT, _INT, _V3, _VP, _VP],
    "pic_unpack_planes": [_PP, _INT, _INT, _INT, _INT, _V3, _VP, _INT, _VP],
    "pic_div_residual": [_PP, _V3, _VP, _DBL, _VP, _DBL, _VP, _VP],
    "pic_sum_squares_interior": [_PP, _VP, _VP, _VP],
    "pic_particle_energy": [_PP, _VP, _VP, _I64, _VP, _VP],
    "pic_soa_import": [_PP, _INT, _VP, _VP, _VP, _I64, _SOA, _VP, _VP],
    "pic_soa_export": [_PP, _INT, _SOA, _VP, _VP, _VP, _I64, _VP, _VP],
    "pic_sort_histogram": [_PP, _SOA, _VP, _VP],
    "pic_sort_scan": [_I64, _VP, _VP, _VP, _VP],
    "pic_sort_scatter": [_PP, _SOA, _SOA, _VP, _VP, _VP],
    "pic_fused_push_deposit": [_PP, _INT, _INT, _SOA, _V3, _V3, _V3, _V3, _V3, _LEAVE, _VP, _VP],
    "pic_fused_tile3d": [_PP, _INT, _SOA, _VP, _INT, _INT, _V3, _V3, _V3, _LEAVE, _VP, _VP],
    "pic_pack_boxes": [_PP, _INT, _VP, _VP, _INT, _V3, _VP, _VP],
    "pic_unpack_boxes": [_PP, _INT, _VP, _VP, _INT, _V3, _VP, _INT, _VP],
    "pic_fused_pair3d": [_PP, _INT, _SOA, _VP, _INT, _INT, _V3, _V3, _V3, _LEAVE, _VP, _VP, _I64, _VP],
    "pic_pair_work_bytes": [_PP, _I64],
    "pic_sort_blocked_offsets": [_PP, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP],
    "pic_sort_blocked_finish": [_PP, _VP, _VP, _VP, _SOA, _VP],
    "pic_packets_reset": [_PP, _LEAVE, _VP],
    "pic_soa_append_packets": [_PP, _SOA, _LEAVE, _VP, _VP],
    "pic_microbench": [_INT, _INT, ctypes.POINTER(ctypes.c_float)],
    "pic_params_size": [],
    "pic_version": [],
}


def lib():
    """Load (building first if sources are newer) the CUDA library.  Raises if unavailable: no CPU fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if needs_build():
        build()
    try:
        L = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover
        raise RuntimeError(f"pypic3d_b200: cannot load {LIB_PATH}: {exc}; the CUDA extension is mandatory") from exc
    ns = _Namespace()
    for name, argtypes in SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_char_p if name == "pic_version" else (ctypes.c_int64 if name == "pic_pair_work_bytes" else ctypes.c_int)
        setattr(ns, name, _Counted(fn, KERNELS_PER_CALL.get(name, 1)))
    if L.pic_params_size() != ctypes.sizeof(PicParams):
        raise RuntimeError("pypic3d_b200: PicParams layout mismatch between Python and libpic_b200.so")
    version = ns.pic_version().decode()
    if "PIC_B200_LIB" not in os.environ and source_hash() not in version:
        raise RuntimeError(f"pypic3d_b200: {LIB_PATH} ({version}) was not built from the sources beside it "
                           f"(source hash {source_hash()}); run pypic3d_b200._lib.build(force=True)")
    ns._cdll = L
    _LIB = ns
    return ns


class PicError(RuntimeError):
    pass


PIC_EINVAL, PIC_EUNSUPPORTED = -1, -2


def check(code, what):
    if code != 0:
        kind = {-1: "invalid argument", -2: "unsupported configuration"}.get(code, f"CUDA error {code}")
        raise PicError(f"{what}: {kind}")


PUSHERS = {("boris", False): 0, ("boris", True): 1, ("higuera_cary", True): 2, ("higuera_cary", False): 2}


def _scalar(v):
    if hasattr(v, "item"):
        return v.item()
    return v


def make_params(static_parameters, dynamic_parameters, species_config=None, dtype=np.float64, mesh=None, gmesh=None,
                moff=(0, 0, 0)):
    """POD parameter block from the reference-shaped pytrees (parameters.py:14-53, particle_class.py:5-14)."""
    sp, dp = static_parameters, dynamic_parameters
    p = PicParams()
    p.dtype = 0 if np.dtype(dtype) == np.float32 else 1
    p.shape_factor = int(getattr(sp, "shape_factor", 1))
    key = (getattr(sp, "particle_pusher", "boris"), bool(getattr(sp, "relativistic", True)))
    if key not in PUSHERS:
        raise ValueError(f"Unknown particle_pusher: {key[0]}")
    p.pusher = PUSHERS[key]
    p.g = int(sp.guard_cells)
    tile = [int(w) for w in sp.tile_shape]
    N = [int(_scalar(dp.Nx)), int(_scalar(dp.Ny)), int(_scalar(dp.Nz))]
    full_mesh = [N[a] // tile[a] for a in range(3)]
    gmesh = full_mesh if gmesh is None else [int(v) for v in gmesh]
    mesh = gmesh if mesh is None else [int(v) for v in mesh]
    d = [float(_scalar(dp.dx)), float(_scalar(dp.dy)), float(_scalar(dp.dz))]
    wind = [float(_scalar(dp.x_wind)), float(_scalar(dp.y_wind)), float(_scalar(dp.z_wind))]
    for a in range(3):
        p.mesh[a], p.gmesh[a], p.moff[a], p.tile[a] = mesh[a], gmesh[a], int(moff[a]), tile[a]
        p.field_bc[a] = int(getattr(sp, "boundary_conditions", (0, 0, 0))[a])
        p.particle_bc[a] = int(getattr(sp, "particle_boundary_conditions", (0, 0, 0))[a])
        p.wind[a] = wind[a]
        grids = getattr(dp, "grids", None)
        center = getattr(grids, "center", ()) if grids is not None else ()
        vertex = getattr(grids, "vertex", ()) if grids is not None else ()
        p.center0[a] = float(center[a][0]) if len(center) == 3 else -wind[a] / 2 - d[a]
        p.vertex0[a] = float(vertex[a][0]) if len(vertex) == 3 else -wind[a] / 2 - 0.5 * d[a]
    p.dt, p.dx, p.dy, p.dz = float(_scalar(getattr(dp, "dt", 0.0))), d[0], d[1], d[2]
    p.C, p.eps, p.mu, p.alpha = (float(_scalar(getattr(dp, k, 1.0))) for k in ("C", "eps", "mu", "alpha"))
    if species_config is not None:
        charge = np.asarray(_to_numpy(species_config.charge), dtype=np.float64).reshape(-1)
        S = charge.shape[0]
        if S > MAX_SPECIES:
            raise ValueError(f"at most {MAX_SPECIES} species are supported")
        mass = np.asarray(_to_numpy(species_config.mass), dtype=np.float64).reshape(-1)
        weight = np.asarray(_to_numpy(species_config.weight), dtype=np.float64).reshape(-1)
        ux = np.asarray(_to_numpy(species_config.update_x)).astype(bool).reshape(S, 3)
        uu = np.asarray(_to_numpy(species_config.update_u)).astype(bool).reshape(S, 3)
        p.n_species = S
        for s in range(S):
            p.charge[s], p.mass[s], p.weight[s] = charge[s], mass[s], weight[s]
            for c in range(3):
                p.update_x[s][c] = int(ux[s, c])
                p.update_u[s][c] = int(uu[s, c])
    return p


def _to_numpy(a):
    if hasattr(a, "detach"):
        return a.detach().cpu().numpy()
    return np.asarray(a)
