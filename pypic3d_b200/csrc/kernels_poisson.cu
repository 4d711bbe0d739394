// kernels_poisson.cu -- the electrostatic field solve of PyPIC3D/solvers/electrostatic_yee.py on one ghosted tile:
//   pic_poisson_cg      matrix-free conjugate gradient for  -lapl(phi) = rho / eps   (electrostatic_yee.py:71-156)
//   pic_constant_wall   constant-potential exterior ghosts on a conducting wall      (ghost_cells.py:365-386, 675-700)
//   pic_gradient_neg    E = -grad(phi), centred differences on the tile interior     (electrostatic_yee.py:212-246)
// The CG iteration is the reference's (same update order, same stopping rule  k < max_iter && sum r^2 > tol^2); its scalars
// (sum r^2, p.Ap, the iteration count and the "done" flag) live on the device, every kernel is a no-op once "done" is set, and
// the host looks at the flag only every `check_every` iterations -- so the result is the one the reference's while_loop
// produces, without a host round trip per iteration.  Dot products are accumulated in double for both dtypes.
#include "pic_common.cuh"

namespace pic {

struct PoissonDims {
    int L[3], W[3], g;
    int64_t n_int;       // interior points
    double inv_d2[3];    // 1/dx^2, 1/dy^2, 1/dz^2
    double inv_2d[3];    // 1/(2 dx) ...
    double inv_eps;
};

static PoissonDims poisson_dims(const PicParams* p) {
    PoissonDims d;
    const double dd[3] = {p->dx, p->dy, p->dz};
    d.g = p->g;
    d.n_int = 1;
    for (int a = 0; a < 3; ++a) {
        d.W[a] = p->tile[a];
        d.L[a] = p->tile[a] + 2 * p->g;
        d.n_int *= d.W[a];
        d.inv_d2[a] = 1.0 / (dd[a] * dd[a]);
        d.inv_2d[a] = 1.0 / (2.0 * dd[a]);
    }
    d.inv_eps = 1.0 / p->eps;
    return d;
}

// scalars: [0] sum r^2 (current), [1] p.Ap, [2] sum r^2 (next), [3] done flag, [4] iterations done
enum { S_RR = 0, S_PAP = 1, S_RRN = 2, S_DONE = 3, S_ITER = 4 };

__device__ __forceinline__ int64_t interior_index(const PoissonDims& d, int64_t i) {
    const int z = (int)(i % d.W[2]); i /= d.W[2];
    const int y = (int)(i % d.W[1]); i /= d.W[1];
    const int x = (int)i;
    return ((int64_t)(x + d.g) * d.L[1] + (y + d.g)) * d.L[2] + (z + d.g);
}

// lapl(f) at ghosted index c  (electrostatic_yee.py:101-105: second differences divided by the squared spacing, summed x+y+z)
template <typename T>
__device__ __forceinline__ T lapl_at(const PoissonDims& d, const T* __restrict__ f, int64_t c) {
    const int64_t sx = (int64_t)d.L[1] * d.L[2], sy = d.L[2];
    const T fc = f[c];
    const T ax = (f[c + sx] + f[c - sx] - (T)2 * fc) * (T)d.inv_d2[0];
    const T ay = (f[c + sy] + f[c - sy] - (T)2 * fc) * (T)d.inv_d2[1];
    const T az = (f[c + 1] + f[c - 1] - (T)2 * fc) * (T)d.inv_d2[2];
    return ax + ay + az;
}

__device__ __forceinline__ void block_add(double acc, double* out) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(out, acc);
}

// residual = rho/eps + lapl(phi);  p0 = 0 with the interior set to the residual  (electrostatic_yee.py:137-141)
template <typename T>
__global__ void __launch_bounds__(256) k_cg_init(PoissonDims d, const T* __restrict__ rho, const T* __restrict__ phi, T* __restrict__ r,
                                                 T* __restrict__ pv, double* scal) {
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n_int; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = interior_index(d, i);
        const T res = rho[c] * (T)d.inv_eps + lapl_at<T>(d, phi, c);
        r[c] = res;
        pv[c] = res;
        acc += (double)res * (double)res;
    }
    block_add(acc, scal + S_RR);
}

// lp = -lapl(p) on the interior, p.Ap accumulated  (electrostatic_yee.py:113-114)
template <typename T>
__global__ void __launch_bounds__(256) k_cg_ap(PoissonDims d, const T* __restrict__ pv, T* __restrict__ lp, double* scal) {
    if (scal[S_DONE] != 0.0) return;
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n_int; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = interior_index(d, i);
        const T v = -lapl_at<T>(d, pv, c);
        lp[c] = v;
        acc += (double)pv[c] * (double)v;
    }
    block_add(acc, scal + S_PAP);
}

// alpha = sum r^2 / p.Ap;  phi += alpha p;  r -= alpha lp;  next sum r^2  (electrostatic_yee.py:114-119)
template <typename T>
__global__ void __launch_bounds__(256) k_cg_xr(PoissonDims d, T* __restrict__ phi, T* __restrict__ r, const T* __restrict__ pv,
                                               const T* __restrict__ lp, double* scal) {
    if (scal[S_DONE] != 0.0) return;
    const T alpha = (T)(scal[S_RR] / scal[S_PAP]);
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n_int; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = interior_index(d, i);
        phi[c] += alpha * pv[c];
        const T rn = r[c] - alpha * lp[c];
        r[c] = rn;
        acc += (double)rn * (double)rn;
    }
    block_add(acc, scal + S_RRN);
}

// beta = next sum r^2 / sum r^2;  p = r + beta p on the interior  (electrostatic_yee.py:120-122)
template <typename T>
__global__ void __launch_bounds__(256) k_cg_p(PoissonDims d, const T* __restrict__ r, T* __restrict__ pv, const double* scal) {
    if (scal[S_DONE] != 0.0) return;
    const T beta = (T)(scal[S_RRN] / scal[S_RR]);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n_int; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = interior_index(d, i);
        pv[c] = r[c] + beta * pv[c];
    }
}

// roll the scalars to the next iteration and evaluate the reference's loop condition (electrostatic_yee.py:125-128)
__global__ void k_cg_next(double* scal, double tol2, int max_iter, int first) {
    if (scal[S_DONE] != 0.0) return;
    if (!first) {
        scal[S_RR] = scal[S_RRN];
        scal[S_ITER] += 1.0;
    }
    scal[S_RRN] = 0.0;
    scal[S_PAP] = 0.0;
    if (!(scal[S_ITER] < (double)max_iter && scal[S_RR] > tol2)) scal[S_DONE] = 1.0;
}

// conducting wall on `axis`: every exterior ghost plane takes the value of the adjacent interior plane
template <typename T>
__global__ void __launch_bounds__(256) k_constant_wall(PoissonDims d, int axis, T* __restrict__ f) {
    const int ua = axis == 0 ? 1 : 0, va = axis == 2 ? 1 : 2;
    const int64_t plane = (int64_t)d.L[ua] * d.L[va];
    const int64_t total = plane * d.g * 2;
    const int64_t st[3] = {(int64_t)d.L[1] * d.L[2], d.L[2], 1};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int v = (int)(r % d.L[va]); r /= d.L[va];
        const int u = (int)(r % d.L[ua]); r /= d.L[ua];
        const int k = (int)(r % d.g); r /= d.g;
        const bool hi = r != 0;
        const int ghost = hi ? d.L[axis] - d.g + k : k;
        const int src = hi ? d.L[axis] - d.g - 1 : d.g;
        const int64_t base = (int64_t)u * st[ua] + (int64_t)v * st[va];
        f[base + ghost * st[axis]] = f[base + src * st[axis]];
    }
}

// E = -(phi[+1] - phi[-1]) / (2 d) on the interior  (electrostatic_yee.py:233-243)
template <typename T>
__global__ void __launch_bounds__(256) k_gradient_neg(PoissonDims d, const T* __restrict__ phi, T* __restrict__ Ex, T* __restrict__ Ey,
                                                      T* __restrict__ Ez) {
    const int64_t sx = (int64_t)d.L[1] * d.L[2], sy = d.L[2];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n_int; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = interior_index(d, i);
        Ex[c] = (T)-1 * (phi[c + sx] - phi[c - sx]) * (T)d.inv_2d[0];
        Ey[c] = (T)-1 * (phi[c + sy] - phi[c - sy]) * (T)d.inv_2d[1];
        Ez[c] = (T)-1 * (phi[c + 1] - phi[c - 1]) * (T)d.inv_2d[2];
    }
}

}  // namespace pic

using namespace pic;

extern "C" int pic_halo_refresh_axis(const PicParams* p, int axis, int bc, int ncomp, void* const* fields, void* stream);

namespace pic {

template <typename T>
static int launch_constant_wall(const PicParams* p, int axis, void* f, cudaStream_t st) {
    const PoissonDims d = poisson_dims(p);
    const int ua = axis == 0 ? 1 : 0, va = axis == 2 ? 1 : 2;
    k_constant_wall<T><<<grid_for((int64_t)d.L[ua] * d.L[va] * d.g * 2, 256), 256, 0, st>>>(d, axis, (T*)f);
    PIC_LAUNCH_RET();
}

// _apply_tiled_phi_constant_boundaries (electrostatic_yee.py:20-37) for one tile: per conducting axis a full field-BC ghost
// refresh followed by the constant wall fill; a plain refresh when no axis is conducting.
template <typename T>
static int apply_phi_bc(const PicParams* p, void* f, cudaStream_t st) {
    void* one[1] = {f};
    bool applied = false;
    for (int axis = 0; axis < 3; ++axis) {
        if (p->field_bc[axis] != PIC_BC_CONDUCTING) continue;
        for (int a = 0; a < 3; ++a) {
            const int rc = pic_halo_refresh_axis(p, a, p->field_bc[a], 1, one, st);
            if (rc) return rc;
        }
        const int rc = launch_constant_wall<T>(p, axis, f, st);
        if (rc) return rc;
        applied = true;
    }
    if (!applied)
        for (int a = 0; a < 3; ++a) {
            const int rc = pic_halo_refresh_axis(p, a, p->field_bc[a], 1, one, st);
            if (rc) return rc;
        }
    return 0;
}

template <typename T>
static int run_cg(const PicParams* p, const void* rho, void* phi, void* r, void* pv, void* lp, double* scal, double tol, int max_iter,
                  int check_every, int* iters_out, cudaStream_t st) {
    const PoissonDims d = poisson_dims(p);
    const int grid = grid_for(d.n_int, 256, 4);
    const size_t bytes = (size_t)d.L[0] * d.L[1] * d.L[2] * sizeof(T);
    int rc;
    if ((rc = apply_phi_bc<T>(p, phi, st))) return rc;                                   // phi = apply_bc(phi)            :137
    cudaMemsetAsync(scal, 0, 8 * sizeof(double), st);
    cudaMemsetAsync(pv, 0, bytes, st);                                                   // p0 = zeros_like(phi)            :139
    cudaMemsetAsync(r, 0, bytes, st);
    cudaMemsetAsync(lp, 0, bytes, st);
    k_cg_init<T><<<grid, 256, 0, st>>>(d, (const T*)rho, (const T*)phi, (T*)r, (T*)pv, scal);
    if ((rc = apply_phi_bc<T>(p, pv, st))) return rc;                                    // p0 = apply_bc(p0)               :141
    k_cg_next<<<1, 1, 0, st>>>(scal, tol * tol, max_iter, 1);
    if (check_every < 1) check_every = 1;
    double host[8];
    for (int it = 0; it < max_iter;) {
        for (int k = 0; k < check_every && it < max_iter; ++k, ++it) {
            k_cg_ap<T><<<grid, 256, 0, st>>>(d, (const T*)pv, (T*)lp, scal);
            k_cg_xr<T><<<grid, 256, 0, st>>>(d, (T*)phi, (T*)r, (const T*)pv, (const T*)lp, scal);
            k_cg_p<T><<<grid, 256, 0, st>>>(d, (const T*)r, (T*)pv, scal);
            if ((rc = apply_phi_bc<T>(p, pv, st))) return rc;                            // p_next = apply_bc(p_next)       :122
            k_cg_next<<<1, 1, 0, st>>>(scal, tol * tol, max_iter, 0);
        }
        cudaMemcpyAsync(host, scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st);
        const cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return (int)e;
        if (host[S_DONE] != 0.0) break;
    }
    cudaMemcpyAsync(host, scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    if (iters_out) *iters_out = (int)host[S_ITER];
    // the reference refreshes phi's ghosts every iteration (:117); they are never read inside the loop, so one refresh of the
    // final phi gives the same array
    if ((rc = apply_phi_bc<T>(p, phi, st))) return rc;                                   // return apply_bc(phi)            :156
    PIC_LAUNCH_RET();
}

template <typename T>
static int launch_gradient(const PicParams* p, const void* phi, void* const E[3], cudaStream_t st) {
    const PoissonDims d = poisson_dims(p);
    k_gradient_neg<T><<<grid_for(d.n_int, 256, 4), 256, 0, st>>>(d, (const T*)phi, (T*)E[0], (T*)E[1], (T*)E[2]);
    PIC_LAUNCH_RET();
}

}  // namespace pic

extern "C" {

int pic_poisson_cg(const PicParams* p, const void* rho, void* phi, void* work_r, void* work_p, void* work_lp, double* scal, double tol,
                   int max_iter, int check_every, int* iters_out, void* stream) {
    PIC_CHECK_ARG(p && rho && phi && work_r && work_p && work_lp && scal && max_iter >= 0 && tol >= 0.0);
    PIC_CHECK_ARG(p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1 && p->g >= 1);
    PIC_DISPATCH_T(p, run_cg, p, rho, phi, work_r, work_p, work_lp, scal, tol, max_iter, check_every, iters_out, (cudaStream_t)stream);
}

int pic_phi_boundaries(const PicParams* p, void* field, void* stream) {
    PIC_CHECK_ARG(p && field && p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    PIC_DISPATCH_T(p, apply_phi_bc, p, field, (cudaStream_t)stream);
}

int pic_constant_wall(const PicParams* p, int axis, void* field, void* stream) {
    PIC_CHECK_ARG(p && field && axis >= 0 && axis <= 2 && p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    PIC_DISPATCH_T(p, launch_constant_wall, p, axis, field, (cudaStream_t)stream);
}

int pic_gradient_neg(const PicParams* p, const void* phi, void* const E[3], void* stream) {
    PIC_CHECK_ARG(p && phi && E && E[0] && E[1] && E[2] && p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1 && p->g >= 1);
    PIC_DISPATCH_T(p, launch_gradient, p, phi, E, (cudaStream_t)stream);
}

}  // extern "C"
