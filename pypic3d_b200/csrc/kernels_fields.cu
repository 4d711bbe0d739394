// kernels_fields.cu -- grid kernels: Yee E / B(half) updates, 27-point filters, guard-cell refresh / fold,
// wall planes, face pack / unpack for NCCL halo exchange, energy reductions.
// All are HBM-bandwidth-bound streaming/stencil kernels: z (fastest axis) maps to threadIdx.x for coalescing.
#include "pic_common.cuh"

namespace pic {

struct Dims {
    int L[3];
    int W[3];
    int g;
    int64_t ntiles;
    size_t tile_elems;
};

static inline Dims dims_of(const PicParams* p) {
    Dims d;
    for (int a = 0; a < 3; ++a) { d.W[a] = p->tile[a]; d.L[a] = p->tile[a] + 2 * p->g; }
    d.g = p->g;
    d.ntiles = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2];
    d.tile_elems = (size_t)d.L[0] * d.L[1] * d.L[2];
    return d;
}

// Interior-cell launch geometry: block (x=z cells, y=y cells), grid.x over (tile, x cell, y/z blocks).
struct CellGrid {
    dim3 grid, block;
    int nbz, nby;
};
static inline CellGrid cell_grid(const Dims& d) {
    CellGrid c;
    int bx = d.W[2] >= 128 ? 128 : (d.W[2] >= 64 ? 64 : 32);
    int by = 256 / bx;
    if (by > d.W[1]) by = d.W[1] < 1 ? 1 : d.W[1];
    c.block = dim3(bx, by, 1);
    c.nbz = (d.W[2] + bx - 1) / bx;
    c.nby = (d.W[1] + by - 1) / by;
    const int64_t blocks = d.ntiles * d.W[0] * c.nby * c.nbz;
    c.grid = dim3((unsigned)blocks, 1, 1);
    return c;
}
__device__ __forceinline__ bool cell_of_block(const Dims& d, int nby, int nbz, int64_t& tile, int& ix, int& iy, int& iz) {
    int64_t b = blockIdx.x;
    const int bz = (int)(b % nbz); b /= nbz;
    const int by = (int)(b % nby); b /= nby;
    ix = (int)(b % d.W[0]) + d.g; b /= d.W[0];
    tile = b;
    iy = by * blockDim.y + threadIdx.y;
    iz = bz * blockDim.x + threadIdx.x;
    if (iy >= d.W[1] || iz >= d.W[2]) return false;
    iy += d.g; iz += d.g;
    return true;
}

// ---------------------------------------------------------------- Yee (first_order_yee.py:42-72, :121-142)
template <typename T>
__global__ void __launch_bounds__(256) k_update_B(Dims d, int nby, int nbz, T* __restrict__ Bx, T* __restrict__ By,
                                                  T* __restrict__ Bz, const T* __restrict__ Ex, const T* __restrict__ Ey,
                                                  const T* __restrict__ Ez, T hdt, T dx, T dy, T dz) {
    int64_t tile; int ix, iy, iz;
    if (!cell_of_block(d, nby, nbz, tile, ix, iy, iz)) return;
    const size_t sx = (size_t)d.L[1] * d.L[2], sy = d.L[2];
    const size_t i = tile * d.tile_elems + ix * sx + iy * sy + iz;
    const T ex = Ex[i], ey = Ey[i], ez = Ez[i];
    const T dEz_dy = (Ez[i + sy] - ez) * dy;   // forward differences (dx, dy, dz are the RECIPROCAL spacings here)
    const T dEy_dz = (Ey[i + 1] - ey) * dz;
    const T dEx_dz = (Ex[i + 1] - ex) * dz;
    const T dEx_dy = (Ex[i + sy] - ex) * dy;
    const T dEz_dx = (Ez[i + sx] - ez) * dx;
    const T dEy_dx = (Ey[i + sx] - ey) * dx;
    Bx[i] = Bx[i] - hdt * (dEz_dy - dEy_dz);
    By[i] = By[i] - hdt * (dEx_dz - dEz_dx);
    Bz[i] = Bz[i] - hdt * (dEy_dx - dEx_dy);
}

template <typename T>
__global__ void __launch_bounds__(256) k_update_E(Dims d, int nby, int nbz, T* __restrict__ Ex, T* __restrict__ Ey,
                                                  T* __restrict__ Ez, const T* __restrict__ Bx, const T* __restrict__ By,
                                                  const T* __restrict__ Bz, const T* __restrict__ Jx, const T* __restrict__ Jy,
                                                  const T* __restrict__ Jz, T dt, T dx, T dy, T dz, T C2, T eps) {
    int64_t tile; int ix, iy, iz;
    if (!cell_of_block(d, nby, nbz, tile, ix, iy, iz)) return;
    const size_t sx = (size_t)d.L[1] * d.L[2], sy = d.L[2];
    const size_t i = tile * d.tile_elems + ix * sx + iy * sy + iz;
    const T bx = Bx[i], by = By[i], bz = Bz[i];
    const T dBz_dy = (bz - Bz[i - sy]) * dy;   // backward differences (dx, dy, dz, eps are RECIPROCALS here)
    const T dBy_dz = (by - By[i - 1]) * dz;
    const T dBx_dz = (bx - Bx[i - 1]) * dz;
    const T dBx_dy = (bx - Bx[i - sy]) * dy;
    const T dBz_dx = (bz - Bz[i - sx]) * dx;
    const T dBy_dx = (by - By[i - sx]) * dx;
    Ex[i] = Ex[i] + (C2 * (dBz_dy - dBy_dz) - Jx[i] * eps) * dt;
    Ey[i] = Ey[i] + (C2 * (dBx_dz - dBz_dx) - Jy[i] * eps) * dt;
    Ez[i] = Ez[i] + (C2 * (dBy_dx - dBx_dy) - Jz[i] * eps) * dt;
}

template <typename T>
static int launch_update_B(const PicParams* p, void* const B[3], const void* const E[3], cudaStream_t st) {
    const Dims d = dims_of(p);
    const CellGrid cg = cell_grid(d);
    k_update_B<T><<<cg.grid, cg.block, 0, st>>>(d, cg.nby, cg.nbz, (T*)B[0], (T*)B[1], (T*)B[2], (const T*)E[0], (const T*)E[1],
                                                (const T*)E[2], (T)(p->dt / 2), (T)(1.0 / p->dx), (T)(1.0 / p->dy), (T)(1.0 / p->dz));
    PIC_LAUNCH_RET();
}
template <typename T>
static int launch_update_E(const PicParams* p, void* const E[3], const void* const B[3], const void* const J[3], cudaStream_t st) {
    const Dims d = dims_of(p);
    const CellGrid cg = cell_grid(d);
    k_update_E<T><<<cg.grid, cg.block, 0, st>>>(d, cg.nby, cg.nbz, (T*)E[0], (T*)E[1], (T*)E[2], (const T*)B[0], (const T*)B[1],
                                                (const T*)B[2], (const T*)J[0], (const T*)J[1], (const T*)J[2], (T)p->dt, (T)(1.0 / p->dx),
                                                (T)(1.0 / p->dy), (T)(1.0 / p->dz), (T)(p->C * p->C), (T)(1.0 / p->eps));
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- 27-point filter (filters.py:73-129)
template <typename T>
__global__ void __launch_bounds__(256) k_filter(Dims d, int nby, int nbz, const T* __restrict__ in, T* __restrict__ out,
                                                int kind, T alpha) {
    int64_t tile; int ix, iy, iz;
    if (!cell_of_block(d, nby, nbz, tile, ix, iy, iz)) return;
    const ptrdiff_t sx = (ptrdiff_t)d.L[1] * d.L[2], sy = d.L[2];
    const size_t i = tile * d.tile_elems + ix * sx + iy * sy + iz;
    T acc = (T)0;
    if (kind == PIC_FILTER_DIGITAL) {
        const T nw = ((T)1 - alpha) / (T)6;
        // same accumulation order as a row-major 3x3x3 kernel walk
        acc += nw * in[i - sx];
        acc += nw * in[i - sy];
        acc += nw * in[i - 1];
        acc += alpha * in[i];
        acc += nw * in[i + 1];
        acc += nw * in[i + sy];
        acc += nw * in[i + sx];
    } else {
        const T k1[3] = {(T)1, (T)2, (T)1};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c)
                    acc += ((k1[a] * k1[b] * k1[c]) / (T)64) * in[i + (a - 1) * sx + (b - 1) * sy + (c - 1)];
    }
    out[i] = acc;
}

template <typename T>
static int launch_filter(const PicParams* p, int kind, double alpha, const void* in, void* out, cudaStream_t st) {
    const Dims d = dims_of(p);
    cudaError_t e = cudaMemcpyAsync(out, in, d.ntiles * d.tile_elems * sizeof(T), cudaMemcpyDeviceToDevice, st);  // ghosts unchanged
    if (e != cudaSuccess) return (int)e;
    const CellGrid cg = cell_grid(d);
    k_filter<T><<<cg.grid, cg.block, 0, st>>>(d, cg.nby, cg.nbz, (const T*)in, (T*)out, kind, (T)alpha);
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- per-axis halo passes over the local tile mesh
struct Ptrs8 {
    void* f[8];
};

// Decompose a flat index over (comp, tile, plane, u, v) where (u, v) span the FULL transverse extent (ghosts included,
// ghost_cells.py:162-178) and v is the fastest-varying memory axis available.
struct AxisView {
    int axis, ua, va;      // axis and the two transverse axes (ua slower, va faster in memory)
    int64_t stride[3];
};
__host__ __device__ inline AxisView axis_view(const Dims& d, int axis) {
    AxisView v;
    v.axis = axis;
    v.ua = axis == 0 ? 1 : 0;
    v.va = axis == 2 ? 1 : 2;
    v.stride[0] = (int64_t)d.L[1] * d.L[2];
    v.stride[1] = d.L[2];
    v.stride[2] = 1;
    return v;
}

// refresh (ghost_cells.py:142-196): ghost_lo[0:g] <- neighbour(-1).[-2g:-g]; ghost_hi[-g:] <- neighbour(+1).[g:2g];
// zeros at a non-periodic chain end; reduced axis: both ghosts <- the single interior plane (periodic) or 0.
template <typename T>
__global__ void __launch_bounds__(256) k_refresh_axis(Dims d, int axis, int bc, int reduced, int mesh0, int mesh1, int mesh2,
                                                      int ncomp, Ptrs8 F) {
    const AxisView av = axis_view(d, axis);
    const int g = d.g, La = d.L[axis], Lu = d.L[av.ua], Lv = d.L[av.va];
    const int mesh[3] = {mesh0, mesh1, mesh2};
    const int64_t per_tile = (int64_t)2 * g * Lu * Lv;
    const int64_t total = (int64_t)ncomp * d.ntiles * per_tile;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int v = (int)(r % Lv); r /= Lv;
        const int u = (int)(r % Lu); r /= Lu;
        const int pl = (int)(r % (2 * g)); r /= (2 * g);
        const int64_t tile = r % d.ntiles;
        const int comp = (int)(r / d.ntiles);
        T* f = (T*)F.f[comp];
        const bool lower = pl < g;
        const int dst_l = lower ? pl : La - 2 * g + pl;           // [0,g) or [La-g, La)
        const int64_t tv = (int64_t)u * av.stride[av.ua] + (int64_t)v * av.stride[av.va];
        T val = (T)0;
        if (reduced) {
            if (bc == PIC_BC_PERIODIC) val = f[tile * d.tile_elems + (int64_t)g * av.stride[axis] + tv];
        } else {
            TileCoord tc = tile_coord(tile, mesh);
            int t[3] = {tc.tx, tc.ty, tc.tz};
            int nt = t[axis] + (lower ? -1 : 1);
            bool have = true;
            if (nt < 0 || nt >= mesh[axis]) {
                if (bc == PIC_BC_PERIODIC) nt = (nt + mesh[axis]) % mesh[axis];
                else have = false;
            }
            if (have) {
                t[axis] = nt;
                const int64_t ntile = ((int64_t)t[0] * mesh[1] + t[1]) * mesh[2] + t[2];
                const int src_l = lower ? (La - 2 * g + pl) : (g + (pl - g));   // [-2g:-g] or [g:2g]
                val = f[ntile * d.tile_elems + (int64_t)src_l * av.stride[axis] + tv];
            }
        }
        f[tile * d.tile_elems + (int64_t)dst_l * av.stride[axis] + tv] = val;
    }
}

// fold, phase 1 (ghost_cells.py:218-289): every interior plane gathers the ghost deposits it owns.
// plane l in [g, g+W): if l < 2g it receives neighbour(-1).ghost_hi[l-g] (+) and, on a conducting lower wall, -own
// ghost_lo[l-g]; if l >= W it receives neighbour(+1).ghost_lo[l-W] (+) and, on a conducting upper wall, -own ghost_hi.
template <typename T>
__global__ void __launch_bounds__(256) k_fold_axis(Dims d, int axis, int bc, int reduced, int mesh0, int mesh1, int mesh2,
                                                   int ncomp, Ptrs8 F) {
    const AxisView av = axis_view(d, axis);
    const int g = d.g, La = d.L[axis], W = d.W[axis], Lu = d.L[av.ua], Lv = d.L[av.va];
    const int mesh[3] = {mesh0, mesh1, mesh2};
    const int npl = reduced ? 1 : (W < 2 * g ? W : 2 * g);        // interior planes that can receive anything
    const int64_t per_tile = (int64_t)npl * Lu * Lv;
    const int64_t total = (int64_t)ncomp * d.ntiles * per_tile;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int v = (int)(r % Lv); r /= Lv;
        const int u = (int)(r % Lu); r /= Lu;
        const int pl = (int)(r % npl); r /= npl;
        const int64_t tile = r % d.ntiles;
        const int comp = (int)(r / d.ntiles);
        T* f = (T*)F.f[comp];
        const int64_t tv = (int64_t)u * av.stride[av.ua] + (int64_t)v * av.stride[av.va];
        const int64_t base = tile * d.tile_elems + tv;
        if (reduced) {
            T gs = (T)0;
            for (int k = 0; k < g; ++k) gs += f[base + (int64_t)k * av.stride[axis]];
            T gs2 = (T)0;
            for (int k = 0; k < g; ++k) gs2 += f[base + (int64_t)(La - g + k) * av.stride[axis]];
            gs = gs + gs2;
            T* tgt = f + base + (int64_t)g * av.stride[axis];
            if (bc == PIC_BC_PERIODIC) *tgt += gs;
            else if (bc == PIC_BC_CONDUCTING) *tgt -= gs;
            continue;
        }
        // map pl to an interior plane: first min(g, .) planes from the bottom, the rest from the top
        int l;
        if (W >= 2 * g) l = pl < g ? g + pl : (La - 2 * g) + (pl - g);
        else l = g + pl;
        TileCoord tc = tile_coord(tile, mesh);
        const int t0[3] = {tc.tx, tc.ty, tc.tz};
        T* tgt = f + base + (int64_t)l * av.stride[axis];
        T acc = *tgt;
        if (l >= La - 2 * g) {      // upper interior <- lower ghost of neighbour(+1)   (ghost_cells.py:271,274)
            const int k = l - (La - 2 * g);
            int nt = t0[axis] + 1;
            bool have = true;
            if (nt >= mesh[axis]) { if (bc == PIC_BC_PERIODIC) nt = 0; else have = false; }
            if (have) {
                int t[3] = {t0[0], t0[1], t0[2]};
                t[axis] = nt;
                const int64_t ntile = ((int64_t)t[0] * mesh[1] + t[1]) * mesh[2] + t[2];
                acc += f[ntile * d.tile_elems + tv + (int64_t)k * av.stride[axis]];
            }
        }
        if (l < 2 * g) {            // lower interior <- upper ghost of neighbour(-1)   (ghost_cells.py:272,275)
            const int k = l - g;
            int nt = t0[axis] - 1;
            bool have = true;
            if (nt < 0) { if (bc == PIC_BC_PERIODIC) nt = mesh[axis] - 1; else have = false; }
            if (have) {
                int t[3] = {t0[0], t0[1], t0[2]};
                t[axis] = nt;
                const int64_t ntile = ((int64_t)t[0] * mesh[1] + t[1]) * mesh[2] + t[2];
                acc += f[ntile * d.tile_elems + tv + (int64_t)(La - g + k) * av.stride[axis]];
            }
        }
        if (bc == PIC_BC_CONDUCTING) {  // _add_exterior_conducting_fold (ghost_cells.py:238-260): wall tiles only
            if (t0[axis] == 0 && l < 2 * g) acc -= f[base + (int64_t)(l - g) * av.stride[axis]];
            if (t0[axis] == mesh[axis] - 1 && l >= La - 2 * g) acc -= f[base + (int64_t)(La - g + (l - (La - 2 * g))) * av.stride[axis]];
        }
        *tgt = acc;
    }
}

// fold, phase 2: ghosts of this axis <- 0 (ghost_cells.py:232-233, 286-287).  Also used stand-alone.
template <typename T>
__global__ void __launch_bounds__(256) k_zero_ghost_axis(Dims d, int axis, int ncomp, Ptrs8 F) {
    const AxisView av = axis_view(d, axis);
    const int g = d.g, La = d.L[axis], Lu = d.L[av.ua], Lv = d.L[av.va];
    const int64_t per_tile = (int64_t)2 * g * Lu * Lv;
    const int64_t total = (int64_t)ncomp * d.ntiles * per_tile;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int v = (int)(r % Lv); r /= Lv;
        const int u = (int)(r % Lu); r /= Lu;
        const int pl = (int)(r % (2 * g)); r /= (2 * g);
        const int64_t tile = r % d.ntiles;
        const int comp = (int)(r / d.ntiles);
        const int l = pl < g ? pl : La - 2 * g + pl;
        ((T*)F.f[comp])[tile * d.tile_elems + (int64_t)l * av.stride[axis] + (int64_t)u * av.stride[av.ua] + (int64_t)v * av.stride[av.va]] = (T)0;
    }
}

static inline int axis_reduced(const PicParams* p, int axis) { return (p->tile[axis] == 1 && p->gmesh[axis] == 1) ? 1 : 0; }

template <typename T>
static int launch_refresh(const PicParams* p, int axis, int bc, int ncomp, void* const* fields, cudaStream_t st) {
    const Dims d = dims_of(p);
    Ptrs8 F;
    for (int c = 0; c < ncomp; ++c) F.f[c] = fields[c];
    const AxisView av = axis_view(d, axis);
    const int64_t total = (int64_t)ncomp * d.ntiles * 2 * d.g * d.L[av.ua] * d.L[av.va];
    k_refresh_axis<T><<<grid_for(total, 256), 256, 0, st>>>(d, axis, bc, axis_reduced(p, axis), p->mesh[0], p->mesh[1], p->mesh[2], ncomp, F);
    PIC_LAUNCH_RET();
}

template <typename T>
static int launch_fold(const PicParams* p, int axis, int bc, int ncomp, void* const* fields, cudaStream_t st) {
    const Dims d = dims_of(p);
    Ptrs8 F;
    for (int c = 0; c < ncomp; ++c) F.f[c] = fields[c];
    const AxisView av = axis_view(d, axis);
    const int red = axis_reduced(p, axis);
    const int npl = red ? 1 : (d.W[axis] < 2 * d.g ? d.W[axis] : 2 * d.g);
    const int64_t total = (int64_t)ncomp * d.ntiles * npl * d.L[av.ua] * d.L[av.va];
    k_fold_axis<T><<<grid_for(total, 256), 256, 0, st>>>(d, axis, bc, red, p->mesh[0], p->mesh[1], p->mesh[2], ncomp, F);
    const int64_t tz = (int64_t)ncomp * d.ntiles * 2 * d.g * d.L[av.ua] * d.L[av.va];
    k_zero_ghost_axis<T><<<grid_for(tz, 256), 256, 0, st>>>(d, axis, ncomp, F);
    PIC_LAUNCH_RET();
}

// wall planes (ghost_cells.py:344-362): plane g on the first tile and plane -g-1 on the last tile of `axis`.
template <typename T>
__global__ void __launch_bounds__(256) k_zero_wall(Dims d, int axis, int mesh0, int mesh1, int mesh2, int lo_wall, int hi_wall, T* f) {
    const AxisView av = axis_view(d, axis);
    const int Lu = d.L[av.ua], Lv = d.L[av.va];
    const int mesh[3] = {mesh0, mesh1, mesh2};
    const int64_t total = d.ntiles * 2 * Lu * Lv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int v = (int)(r % Lv); r /= Lv;
        const int u = (int)(r % Lu); r /= Lu;
        const int side = (int)(r % 2); r /= 2;
        const int64_t tile = r;
        const TileCoord tc = tile_coord(tile, mesh);
        const int t[3] = {tc.tx, tc.ty, tc.tz};
        if (side == 0 && !(lo_wall && t[axis] == 0)) continue;
        if (side == 1 && !(hi_wall && t[axis] == mesh[axis] - 1)) continue;
        const int l = side == 0 ? d.g : d.L[axis] - d.g - 1;
        f[tile * d.tile_elems + (int64_t)l * av.stride[axis] + (int64_t)u * av.stride[av.ua] + (int64_t)v * av.stride[av.va]] = (T)0;
    }
}

template <typename T>
static int launch_zero_wall(const PicParams* p, int axis, void* field, cudaStream_t st) {
    const Dims d = dims_of(p);
    const AxisView av = axis_view(d, axis);
    const int64_t total = d.ntiles * 2 * d.L[av.ua] * d.L[av.va];
    const int lo_wall = (p->moff[axis] == 0), hi_wall = (p->moff[axis] + p->mesh[axis] == p->gmesh[axis]);
    k_zero_wall<T><<<grid_for(total, 256), 256, 0, st>>>(d, axis, p->mesh[0], p->mesh[1], p->mesh[2], lo_wall, hi_wall, (T*)field);
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- face pack / unpack (multi-GPU halo exchange)
// buffer layout [comp][tile][plane][u][v]; planes start..start+nplanes-1 along `axis`, full transverse extent.
template <typename T, int DIR /*0 pack, 1 unpack*/>
__global__ void __launch_bounds__(256) k_planes(Dims d, int axis, int start, int nplanes, int ncomp, Ptrs8 F, T* buf, int mode) {
    const AxisView av = axis_view(d, axis);
    const int Lu = d.L[av.ua], Lv = d.L[av.va];
    const int64_t total = (int64_t)ncomp * d.ntiles * nplanes * Lu * Lv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int v = (int)(r % Lv); r /= Lv;
        const int u = (int)(r % Lu); r /= Lu;
        const int pl = (int)(r % nplanes); r /= nplanes;
        const int64_t tile = r % d.ntiles;
        const int comp = (int)(r / d.ntiles);
        T* f = (T*)F.f[comp] + tile * d.tile_elems + (int64_t)(start + pl) * av.stride[axis] + (int64_t)u * av.stride[av.ua] + (int64_t)v * av.stride[av.va];
        if (DIR == 0) buf[i] = *f;
        else if (mode == PIC_HALO_SET) *f = buf[i];
        else if (mode == PIC_HALO_ADD) *f += buf[i];
        else *f -= buf[i];
    }
}

template <typename T>
static int launch_planes(const PicParams* p, int dir, int axis, int start, int nplanes, int ncomp, void* const* fields, void* buf,
                         int mode, cudaStream_t st) {
    const Dims d = dims_of(p);
    Ptrs8 F;
    for (int c = 0; c < ncomp; ++c) F.f[c] = fields[c];
    const AxisView av = axis_view(d, axis);
    const int64_t total = (int64_t)ncomp * d.ntiles * nplanes * d.L[av.ua] * d.L[av.va];
    if (total == 0) return 0;
    if (dir == 0) k_planes<T, 0><<<grid_for(total, 256), 256, 0, st>>>(d, axis, start, nplanes, ncomp, F, (T*)buf, mode);
    else k_planes<T, 1><<<grid_for(total, 256), 256, 0, st>>>(d, axis, start, nplanes, ncomp, F, (T*)buf, mode);
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- box pack / unpack (one-shot 26-neighbour exchange)
// Up to 26 axis-aligned boxes of one ghosted tile <-> one packed buffer, in ONE launch for all components: box b occupies
// buf[off[b] .. off[b] + ncomp * nx * ny * nz) as [comp][x][y][z].  The boxes are the faces, edges and corners a rank exchanges
// with its neighbours when all split axes are handled in one round instead of x -> y -> z (distributed.py exchange_boxes_).
struct BoxTable {
    int n;
    int lo[26][3];
    int sz[26][3];
    long long off[27];       // element offset of every box in the buffer; off[n] = total
};
template <typename T, int DIR /*0 pack, 1 unpack*/>
__global__ void __launch_bounds__(256) k_boxes(Dims d, int ncomp, Ptrs8 F, const __grid_constant__ BoxTable bt, T* buf, int mode) {
    const long long total = bt.off[bt.n];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int b = 0;
        while (b + 1 < bt.n && i >= bt.off[b + 1]) ++b;                      // (<= 26 boxes, sorted offsets)
        long long r = i - bt.off[b];
        const int nz = bt.sz[b][2], ny = bt.sz[b][1], nx = bt.sz[b][0];
        const int z = (int)(r % nz); r /= nz;
        const int y = (int)(r % ny); r /= ny;
        const int x = (int)(r % nx);
        const int comp = (int)(r / nx);
        T* f = (T*)F.f[comp] + ((size_t)(bt.lo[b][0] + x) * d.L[1] + (bt.lo[b][1] + y)) * d.L[2] + (bt.lo[b][2] + z);
        // (the boxes of a sum-exchange OVERLAP in the array -- a face slab spans the edge and corner regions too -- so the
        // accumulating modes must be atomic; SET boxes are disjoint)
        if (DIR == 0) buf[i] = *f;
        else if (mode == PIC_HALO_SET) *f = buf[i];
        else if (mode == PIC_HALO_ADD) atomicAdd(f, buf[i]);
        else atomicAdd(f, -buf[i]);
    }
}
template <typename T>
static int launch_boxes(const PicParams* p, int dir, int nbox, const int32_t* lo, const int32_t* sz, int ncomp, void* const* fields, void* buf,
                        int mode, cudaStream_t st) {
    const Dims d = dims_of(p);
    Ptrs8 F;
    for (int c = 0; c < ncomp; ++c) F.f[c] = fields[c];
    BoxTable bt;
    bt.n = nbox;
    long long off = 0;
    for (int b = 0; b < nbox; ++b) {
        for (int a = 0; a < 3; ++a) {
            bt.lo[b][a] = lo[3 * b + a]; bt.sz[b][a] = sz[3 * b + a];
            if (lo[3 * b + a] < 0 || sz[3 * b + a] < 0 || lo[3 * b + a] + sz[3 * b + a] > d.L[a]) return PIC_EINVAL;
        }
        bt.off[b] = off;
        off += (long long)ncomp * sz[3 * b] * sz[3 * b + 1] * sz[3 * b + 2];
    }
    bt.off[nbox] = off;
    if (off == 0) return 0;
    if (dir == 0) k_boxes<T, 0><<<grid_for(off, 256), 256, 0, st>>>(d, ncomp, F, bt, (T*)buf, mode);
    else k_boxes<T, 1><<<grid_for(off, 256), 256, 0, st>>>(d, ncomp, F, bt, (T*)buf, mode);
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- sum of squares over tile interiors (utils.py:160-166)
template <typename T>
__global__ void __launch_bounds__(256) k_sumsq(Dims d, const T* __restrict__ f, double* out) {
    const int64_t per_tile = (int64_t)d.W[0] * d.W[1] * d.W[2];
    const int64_t total = d.ntiles * per_tile;
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int z = (int)(r % d.W[2]); r /= d.W[2];
        const int y = (int)(r % d.W[1]); r /= d.W[1];
        const int x = (int)(r % d.W[0]); r /= d.W[0];
        const double v = (double)f[r * d.tile_elems + ((size_t)(x + d.g) * d.L[1] + (y + d.g)) * d.L[2] + (z + d.g)];
        acc += v * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

template <typename T>
static int launch_sumsq(const PicParams* p, const void* f, double* out, cudaStream_t st) {
    const Dims d = dims_of(p);
    const int64_t total = d.ntiles * d.W[0] * d.W[1] * d.W[2];
    k_sumsq<T><<<grid_for(total, 256, 4), 256, 0, st>>>(d, (const T*)f, out);
    PIC_LAUNCH_RET();
}


// ---------------------------------------------------------------- fused Yee step (first_order_yee.py:12-162, evolve.py:88-96)
// B(half, E_old) -> E(full, B', J) -> B(half, E_new) for one tile in ONE pass over the fields: 15 reals per cell (read E, B, J,
// write E, B) instead of 30 for the three sweeps plus their guard-cell refreshes.  A CTA stages the E box [-1, +2] and the B box
// [-1, +1] around its TX x TY x TZ cells in shared memory, recomputes the intermediate B' on [-1, +1] and E_new on [0, +1] (the
// one-cell redundancy the stencils need), and writes E_new, B_new of its own cells to the OUTPUT arrays (never in place: other CTAs
// still read the old values).  Same expressions as k_update_B / k_update_E, so the results are bit-identical to the three sweeps.
// How an axis side is closed (per side, YeeSides):
//   WRAP  single-rank periodic axis: indices outside the interior are wrapped; the epilogue also writes the guard copies;
//   HALO  the guard cells hold the neighbour rank's values (refreshed E, B two deep, J one deep): intermediates are recomputed;
//   WALL  non-periodic global boundary: the refresh the reference performs leaves zeros in the exterior guards, so the
//         intermediates there are zero; conducting walls zero the tangential E on the first / last interior plane.
constexpr int YEE_WRAP = 0, YEE_HALO = 1, YEE_WALL = 2;
struct YeeSides {
    int lo[3], hi[3];             // YEE_* per axis side
    int cond_lo[3], cond_hi[3];   // conducting wall on this rank's low / high side of the axis
};
constexpr int YTY = 8, YTZ = 32;  // tile = TX x 8 x 32 cells, 256 threads = 8 (y) x 32 (z); a thread owns <= 2 x 2 (y, z) columns of a box

// array index along axis a -> where to read it (WRAP: the interior cell it mirrors) / is it an exterior wall guard cell
struct YeeAxis {
    int src;      // source array index
    int wall;     // exterior guard cell of a non-periodic wall: intermediates are zero there
    int cond;     // first / last interior plane of a conducting wall
};
__device__ __forceinline__ YeeAxis yee_axis(const Dims& d, const YeeSides& sd, int a, int i) {
    YeeAxis r;
    const int g = d.g, L = d.L[a];
    i = i < 0 ? 0 : (i > L - 1 ? L - 1 : i);                     // (partial tiles at the upper end read in-range garbage, never stored)
    r.src = i;
    r.wall = 0;
    if (i < g) {
        if (sd.lo[a] == YEE_WRAP) r.src = i + d.W[a];
        r.wall = sd.lo[a] == YEE_WALL;
    } else if (i >= L - g) {
        if (sd.hi[a] == YEE_WRAP) r.src = i - d.W[a];
        r.wall = sd.hi[a] == YEE_WALL;
    }
    r.cond = (sd.cond_lo[a] && i == g) || (sd.cond_hi[a] && i == L - g - 1);
    return r;
}

template <typename T, int TX>
__global__ void __launch_bounds__(256) k_yee_fused(Dims d, YeeSides sd, const T* __restrict__ Ex, const T* __restrict__ Ey,
                                                   const T* __restrict__ Ez, const T* __restrict__ Bx, const T* __restrict__ By,
                                                   const T* __restrict__ Bz, const T* __restrict__ Jx, const T* __restrict__ Jy,
                                                   const T* __restrict__ Jz, T* __restrict__ Ex2, T* __restrict__ Ey2, T* __restrict__ Ez2,
                                                   T* __restrict__ Bx2, T* __restrict__ By2, T* __restrict__ Bz2, T dt, T hdt, T idx, T idy,
                                                   T idz, T C2, T ieps, int nby, int nbz) {
    constexpr int EX = TX + 3, EY = YTY + 3, EZ = YTZ + 3;       // E box: cells [-1, T + 2)
    constexpr int BX = TX + 2, BY = YTY + 2, BZ = YTZ + 2;       // B box: cells [-1, T + 1)
    constexpr int ESX = EY * EZ, ESY = EZ, BSX = BY * BZ, BSY = BZ;
    extern __shared__ __align__(16) unsigned char yee_smem[];
    T* sEx = reinterpret_cast<T*>(yee_smem);                     // [EX][EY][EZ] x 3
    T* sEy = sEx + EX * EY * EZ;
    T* sEz = sEy + EX * EY * EZ;
    T* sBx = sEz + EX * EY * EZ;                                 // [BX][BY][BZ] x 3
    T* sBy = sBx + BX * BY * BZ;
    T* sBz = sBy + BX * BY * BZ;
    int b = blockIdx.x;
    const int bz = b % nbz; b /= nbz;
    const int by = b % nby; b /= nby;
    const int o0 = d.g + b * TX, o1 = d.g + by * YTY, o2 = d.g + bz * YTZ;    // array index of the tile's first cell
    const int sx = d.L[1] * d.L[2], sy = d.L[2];
    const int ty = threadIdx.x >> 5, tz = threadIdx.x & 31;
    // the thread's columns of the boxes: box y = ty (+8), box z = tz (+32); box coordinate c <-> array index o - 1 + c
    YeeAxis ay[2], az[2];
    int cy[2], cz[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        cy[k] = ty + 8 * k; cz[k] = tz + 32 * k;
        ay[k] = yee_axis(d, sd, 1, o1 - 1 + cy[k]);
        az[k] = yee_axis(d, sd, 2, o2 - 1 + cz[k]);
    }
    // ---- phase 0: stage E_old on [-1, +2] and B_old on [-1, +1]
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
        if (cy[ky] >= EY) continue;
#pragma unroll
        for (int kz = 0; kz < 2; ++kz) {
            if (cz[kz] >= EZ) continue;
            const int gyz = ay[ky].src * sy + az[kz].src;
            const bool inB = cy[ky] < BY && cz[kz] < BZ;
#pragma unroll
            for (int lx = 0; lx < EX; ++lx) {
                const YeeAxis ax = yee_axis(d, sd, 0, o0 - 1 + lx);
                const size_t gi = (size_t)ax.src * sx + gyz;
                const int e = lx * ESX + cy[ky] * ESY + cz[kz];
                sEx[e] = Ex[gi]; sEy[e] = Ey[gi]; sEz[e] = Ez[gi];
                if (inB && lx < BX) {
                    const int bb = lx * BSX + cy[ky] * BSY + cz[kz];
                    sBx[bb] = Bx[gi]; sBy[bb] = By[gi]; sBz[bb] = Bz[gi];
                }
            }
        }
    }
    __syncthreads();
    // ---- phase 1: B' = B - (dt/2) curl_forward(E_old) on [-1, +1]  (k_update_B; exterior wall guards: zero)
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
        if (cy[ky] >= BY) continue;
#pragma unroll
        for (int kz = 0; kz < 2; ++kz) {
            if (cz[kz] >= BZ) continue;
            const bool wyz = ay[ky].wall || az[kz].wall;
#pragma unroll
            for (int lx = 0; lx < BX; ++lx) {
                const bool w = wyz || yee_axis(d, sd, 0, o0 - 1 + lx).wall;
                const int e = lx * ESX + cy[ky] * ESY + cz[kz];
                const int bb = lx * BSX + cy[ky] * BSY + cz[kz];
                const T ex = sEx[e], ey = sEy[e], ez = sEz[e];
                const T dEz_dy = (sEz[e + ESY] - ez) * idy;
                const T dEy_dz = (sEy[e + 1] - ey) * idz;
                const T dEx_dz = (sEx[e + 1] - ex) * idz;
                const T dEx_dy = (sEx[e + ESY] - ex) * idy;
                const T dEz_dx = (sEz[e + ESX] - ez) * idx;
                const T dEy_dx = (sEy[e + ESX] - ey) * idx;
                sBx[bb] = w ? (T)0 : sBx[bb] - hdt * (dEz_dy - dEy_dz);
                sBy[bb] = w ? (T)0 : sBy[bb] - hdt * (dEx_dz - dEz_dx);
                sBz[bb] = w ? (T)0 : sBz[bb] - hdt * (dEy_dx - dEx_dy);
            }
        }
    }
    __syncthreads();
    // ---- phase 2: E_new = E + dt (C^2 curl_backward(B') - J / eps) on [0, +1]: box coordinates 1 .. T + 1
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
        if (cy[ky] < 1 || cy[ky] > YTY + 1) continue;
#pragma unroll
        for (int kz = 0; kz < 2; ++kz) {
            if (cz[kz] < 1 || cz[kz] > YTZ + 1) continue;
            const bool wyz = ay[ky].wall || az[kz].wall;
            const int gyz = ay[ky].src * sy + az[kz].src;
#pragma unroll
            for (int lx = 1; lx <= TX + 1; ++lx) {
                const YeeAxis ax = yee_axis(d, sd, 0, o0 - 1 + lx);
                const bool w = wyz || ax.wall;
                const int e = lx * ESX + cy[ky] * ESY + cz[kz];
                const int bb = lx * BSX + cy[ky] * BSY + cz[kz];
                const size_t gi = (size_t)ax.src * sx + gyz;
                const T bx_ = sBx[bb], by_ = sBy[bb], bz_ = sBz[bb];
                const T dBz_dy = (bz_ - sBz[bb - BSY]) * idy;
                const T dBy_dz = (by_ - sBy[bb - 1]) * idz;
                const T dBx_dz = (bx_ - sBx[bb - 1]) * idz;
                const T dBx_dy = (bx_ - sBx[bb - BSY]) * idy;
                const T dBz_dx = (bz_ - sBz[bb - BSX]) * idx;
                const T dBy_dx = (by_ - sBy[bb - BSX]) * idx;
                T ex = sEx[e] + (C2 * (dBz_dy - dBy_dz) - Jx[gi] * ieps) * dt;
                T ey = sEy[e] + (C2 * (dBx_dz - dBz_dx) - Jy[gi] * ieps) * dt;
                T ez = sEz[e] + (C2 * (dBy_dx - dBx_dy) - Jz[gi] * ieps) * dt;
                // conducting walls: tangential E vanishes on the first / last interior plane (first_order_yee.py:80-89)
                if (ay[ky].cond || az[kz].cond || w) ex = (T)0;
                if (ax.cond || az[kz].cond || w) ey = (T)0;
                if (ax.cond || ay[ky].cond || w) ez = (T)0;
                sEx[e] = ex; sEy[e] = ey; sEz[e] = ez;
            }
        }
    }
    __syncthreads();
    // ---- phase 3: B_new = B' - (dt/2) curl_forward(E_new) on the tile's own cells (box 1 .. T); store E_new, B_new and, along
    // WRAP axes, their guard copies (a cell within g of an end also lives W further out on the other side)
    const int iy = o1 + ty, iz = o2 + tz;
    if (iy < d.L[1] - d.g && iz < d.L[2] - d.g) {
        int ny = 1, nz = 1, offy[3] = {0, 0, 0}, offz[3] = {0, 0, 0};
        if (sd.lo[1] == YEE_WRAP) {
            if (iy - d.W[1] >= 0) offy[ny++] = -d.W[1] * sy;
            if (iy + d.W[1] < d.L[1]) offy[ny++] = d.W[1] * sy;
        }
        if (sd.lo[2] == YEE_WRAP) {
            if (iz - d.W[2] >= 0) offz[nz++] = -d.W[2];
            if (iz + d.W[2] < d.L[2]) offz[nz++] = d.W[2];
        }
#pragma unroll
        for (int lx = 1; lx <= TX; ++lx) {
            const int ix = o0 - 1 + lx;
            if (ix >= d.L[0] - d.g) break;
            const int e = lx * ESX + (ty + 1) * ESY + (tz + 1);
            const int bb = lx * BSX + (ty + 1) * BSY + (tz + 1);
            const T ex = sEx[e], ey = sEy[e], ez = sEz[e];
            const T dEz_dy = (sEz[e + ESY] - ez) * idy;
            const T dEy_dz = (sEy[e + 1] - ey) * idz;
            const T dEx_dz = (sEx[e + 1] - ex) * idz;
            const T dEx_dy = (sEx[e + ESY] - ex) * idy;
            const T dEz_dx = (sEz[e + ESX] - ez) * idx;
            const T dEy_dx = (sEy[e + ESX] - ey) * idx;
            const T bxn = sBx[bb] - hdt * (dEz_dy - dEy_dz);
            const T byn = sBy[bb] - hdt * (dEx_dz - dEz_dx);
            const T bzn = sBz[bb] - hdt * (dEy_dx - dEx_dy);
            int nx = 1, offx[3] = {0, 0, 0};
            if (sd.lo[0] == YEE_WRAP) {
                if (ix - d.W[0] >= 0) offx[nx++] = -d.W[0];
                if (ix + d.W[0] < d.L[0]) offx[nx++] = d.W[0];
            }
            for (int i0 = 0; i0 < nx; ++i0)
                for (int i1 = 0; i1 < ny; ++i1)
                    for (int i2 = 0; i2 < nz; ++i2) {
                        const size_t gi = (size_t)(ix + offx[i0]) * sx + (size_t)(iy * sy + offy[i1]) + (iz + offz[i2]);
                        Ex2[gi] = ex; Ey2[gi] = ey; Ez2[gi] = ez;
                        Bx2[gi] = bxn; By2[gi] = byn; Bz2[gi] = bzn;
                    }
        }
    }
}

template <typename T, int TX>
static int launch_yee_fused_tx(const PicParams* p, const Dims& d, const YeeSides& sd, const void* const E[3], const void* const B[3],
                               const void* const J[3], void* const E2[3], void* const B2[3], cudaStream_t st) {
    const int nbx = (d.W[0] + TX - 1) / TX, nby = (d.W[1] + YTY - 1) / YTY, nbz = (d.W[2] + YTZ - 1) / YTZ;
    const size_t smem = (size_t)(3 * (TX + 3) * (YTY + 3) * (YTZ + 3) + 3 * (TX + 2) * (YTY + 2) * (YTZ + 2)) * sizeof(T);
    static size_t attr = 0;
    if (attr < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_yee_fused<T, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = smem;
    }
    k_yee_fused<T, TX><<<nbx * nby * nbz, 256, smem, st>>>(d, sd, (const T*)E[0], (const T*)E[1], (const T*)E[2], (const T*)B[0], (const T*)B[1],
                                                           (const T*)B[2], (const T*)J[0], (const T*)J[1], (const T*)J[2], (T*)E2[0], (T*)E2[1],
                                                           (T*)E2[2], (T*)B2[0], (T*)B2[1], (T*)B2[2], (T)p->dt, (T)(p->dt / 2), (T)(1.0 / p->dx),
                                                           (T)(1.0 / p->dy), (T)(1.0 / p->dz), (T)(p->C * p->C), (T)(1.0 / p->eps), nby, nbz);
    PIC_LAUNCH_RET();
}

// Streaming form of the fused Yee step: a CTA owns a TY x TZ column of cells and MARCHES along x.  Every thread owns one (y, z)
// point of the widest box (E_old on [-1, +2]) and keeps its column's rolling x window in registers -- E_old(p), E_old(p + 1),
// B'(p - 1), B'(p), E_new(p - 1), E_new(p) -- while the y / z neighbours of the current plane are exchanged through three small
// shared-memory planes.  Every plane of E, B, J is therefore read once per (y, z) tile (+ a 1.5-cell rim), the loads of plane p + 2
// are in flight while plane p is computed, and the kernel is bound by HBM instead of by the tile kernel's 3.8x redundant L2 reads.
// Same expressions as k_update_B / k_update_E / k_yee_fused (bit-identical results); same YeeSides semantics.
constexpr int YS_TY = 8, YS_TZ = 32, YS_NY = YS_TY + 3, YS_NZ = YS_TZ + 3, YS_PTS = YS_NY * YS_NZ, YS_THREADS = 416;
#ifndef PIC_YEE_XC
#define PIC_YEE_XC 32          /* planes per CTA along x */
#endif
template <typename T>
__global__ void __launch_bounds__(YS_THREADS) k_yee_stream(Dims d, YeeSides sd, const T* __restrict__ Ex, const T* __restrict__ Ey,
                                                           const T* __restrict__ Ez, const T* __restrict__ Bx, const T* __restrict__ By,
                                                           const T* __restrict__ Bz, const T* __restrict__ Jx, const T* __restrict__ Jy,
                                                           const T* __restrict__ Jz, T* __restrict__ Ex2, T* __restrict__ Ey2,
                                                           T* __restrict__ Ez2, T* __restrict__ Bx2, T* __restrict__ By2, T* __restrict__ Bz2,
                                                           T dt, T hdt, T idx, T idy, T idz, T C2, T ieps, int nby, int nbz, int xc) {
    __shared__ T sEo[3][YS_PTS];          // E_old of the current plane
    __shared__ T sBp[3][YS_PTS];          // B' of the current plane
    __shared__ T sEn[2][3][YS_PTS];       // E_new of the current and the previous plane
    int b = blockIdx.x;
    const int bz = b % nbz; b /= nbz;
    const int by = b % nby; b /= nby;
    const int X0 = d.g + b * xc;                                  // first own plane (array index)
    int X1 = X0 + xc;                                             // one past the last own plane
    if (X1 > d.L[0] - d.g) X1 = d.L[0] - d.g;
    const int o1 = d.g + by * YS_TY, o2 = d.g + bz * YS_TZ;
    const int sx = d.L[1] * d.L[2], sy = d.L[2];
    const int t = threadIdx.x;
    const bool in_box = t < YS_PTS;
    const int cy = in_box ? t / YS_NZ : 0, cz = in_box ? t % YS_NZ : 0;      // box coordinate c <-> array index o - 1 + c
    const YeeAxis ay = yee_axis(d, sd, 1, o1 - 1 + cy), az = yee_axis(d, sd, 2, o2 - 1 + cz);
    const int gyz = ay.src * sy + az.src;
    const bool wyz = ay.wall || az.wall;
    const bool boxB = in_box && cy <= YS_TY + 1 && cz <= YS_TZ + 1;          // B' lives on [-1, +1]
    const bool boxE = in_box && cy >= 1 && cy <= YS_TY + 1 && cz >= 1 && cz <= YS_TZ + 1;   // E_new on [0, +1]
    const int iy = o1 - 1 + cy, iz = o2 - 1 + cz;
    const bool own = in_box && cy >= 1 && cy <= YS_TY && cz >= 1 && cz <= YS_TZ && iy < d.L[1] - d.g && iz < d.L[2] - d.g;
    // guard copies of an own cell along WRAP axes
    int ny = 1, nz = 1, offy[3] = {0, 0, 0}, offz[3] = {0, 0, 0};
    if (own) {
        if (sd.lo[1] == YEE_WRAP) {
            if (iy - d.W[1] >= 0) offy[ny++] = -d.W[1] * sy;
            if (iy + d.W[1] < d.L[1]) offy[ny++] = d.W[1] * sy;
        }
        if (sd.lo[2] == YEE_WRAP) {
            if (iz - d.W[2] >= 0) offz[nz++] = -d.W[2];
            if (iz + d.W[2] < d.L[2]) offz[nz++] = d.W[2];
        }
    }
    auto store6 = [&](T* __restrict__ a0, T* __restrict__ a1, T* __restrict__ a2, int ix, T v0, T v1, T v2) {
        int nx = 1, offx[3] = {0, 0, 0};
        if (sd.lo[0] == YEE_WRAP) {
            if (ix - d.W[0] >= 0) offx[nx++] = -d.W[0];
            if (ix + d.W[0] < d.L[0]) offx[nx++] = d.W[0];
        }
        for (int i0 = 0; i0 < nx; ++i0)
            for (int i1 = 0; i1 < ny; ++i1)
                for (int i2 = 0; i2 < nz; ++i2) {
                    const size_t gi = (size_t)(ix + offx[i0]) * sx + (size_t)(iy * sy + offy[i1]) + (iz + offz[i2]);
                    a0[gi] = v0; a1[gi] = v1; a2[gi] = v2;
                }
    };
    auto plane = [&](int p) { return yee_axis(d, sd, 0, p); };
    // rolling window: eo = E_old(p), eo1 = E_old(p + 1), eo2 = E_old(p + 2) (in flight); bo = B_old(p), bo1 = B_old(p + 1) (in flight);
    // jc = J(p), j1 = J(p + 1) (in flight)
    T eo[3] = {0, 0, 0}, eo1[3] = {0, 0, 0}, eo2[3] = {0, 0, 0}, bo[3] = {0, 0, 0}, bo1[3] = {0, 0, 0}, jc[3] = {0, 0, 0}, j1[3] = {0, 0, 0};
    T bp_prev[3] = {0, 0, 0}, en_prev[3] = {0, 0, 0};
    int p = X0 - 1;
    if (in_box) {
        const size_t g0 = (size_t)plane(p).src * sx + gyz, g1 = (size_t)plane(p + 1).src * sx + gyz;
        eo[0] = Ex[g0]; eo[1] = Ey[g0]; eo[2] = Ez[g0];
        bo[0] = Bx[g0]; bo[1] = By[g0]; bo[2] = Bz[g0];
        eo1[0] = Ex[g1]; eo1[1] = Ey[g1]; eo1[2] = Ez[g1];
    }
    for (; p <= X1; ++p) {
        const YeeAxis ax = plane(p);
        // ---- prefetch: E_old(p + 2), B_old(p + 1), J(p + 1)
        if (in_box && p < X1) {
            const size_t g1 = (size_t)plane(p + 1).src * sx + gyz, g2 = (size_t)plane(p + 2).src * sx + gyz;
            eo2[0] = Ex[g2]; eo2[1] = Ey[g2]; eo2[2] = Ez[g2];
            bo1[0] = Bx[g1]; bo1[1] = By[g1]; bo1[2] = Bz[g1];
            j1[0] = Jx[g1]; j1[1] = Jy[g1]; j1[2] = Jz[g1];
        }
        if (in_box) { sEo[0][t] = eo[0]; sEo[1][t] = eo[1]; sEo[2][t] = eo[2]; }
        __syncthreads();
        // ---- B'(p) = B_old(p) - (dt/2) curl_forward(E_old)(p) on [-1, +1]; exterior wall guards: zero
        T bp[3] = {0, 0, 0};
        if (boxB) {
            const bool w = wyz || ax.wall;
            const T ex = eo[0], ey = eo[1], ez = eo[2];
            const T dEz_dy = (sEo[2][t + YS_NZ] - ez) * idy;
            const T dEy_dz = (sEo[1][t + 1] - ey) * idz;
            const T dEx_dz = (sEo[0][t + 1] - ex) * idz;
            const T dEx_dy = (sEo[0][t + YS_NZ] - ex) * idy;
            const T dEz_dx = (eo1[2] - ez) * idx;
            const T dEy_dx = (eo1[1] - ey) * idx;
            bp[0] = w ? (T)0 : bo[0] - hdt * (dEz_dy - dEy_dz);
            bp[1] = w ? (T)0 : bo[1] - hdt * (dEx_dz - dEz_dx);
            bp[2] = w ? (T)0 : bo[2] - hdt * (dEy_dx - dEx_dy);
            sBp[0][t] = bp[0]; sBp[1][t] = bp[1]; sBp[2][t] = bp[2];
        }
        __syncthreads();
        // ---- E_new(p) = E_old(p) + dt (C^2 curl_backward(B')(p) - J(p) / eps) on [0, +1], planes X0 .. X1
        T en[3] = {0, 0, 0};
        const int q = p & 1;
        if (boxE && p >= X0) {
            const bool w = wyz || ax.wall;
            const T bx_ = bp[0], by_ = bp[1], bz_ = bp[2];
            const T dBz_dy = (bz_ - sBp[2][t - YS_NZ]) * idy;
            const T dBy_dz = (by_ - sBp[1][t - 1]) * idz;
            const T dBx_dz = (bx_ - sBp[0][t - 1]) * idz;
            const T dBx_dy = (bx_ - sBp[0][t - YS_NZ]) * idy;
            const T dBz_dx = (bz_ - bp_prev[2]) * idx;
            const T dBy_dx = (by_ - bp_prev[1]) * idx;
            T ex = eo[0] + (C2 * (dBz_dy - dBy_dz) - jc[0] * ieps) * dt;
            T ey = eo[1] + (C2 * (dBx_dz - dBz_dx) - jc[1] * ieps) * dt;
            T ez = eo[2] + (C2 * (dBy_dx - dBx_dy) - jc[2] * ieps) * dt;
            if (ay.cond || az.cond || w) ex = (T)0;
            if (ax.cond || az.cond || w) ey = (T)0;
            if (ax.cond || ay.cond || w) ez = (T)0;
            en[0] = ex; en[1] = ey; en[2] = ez;
            sEn[q][0][t] = ex; sEn[q][1][t] = ey; sEn[q][2][t] = ez;
            if (own && p < X1) store6(Ex2, Ey2, Ez2, p, ex, ey, ez);
        }
        __syncthreads();
        // ---- B_new(p - 1) = B'(p - 1) - (dt/2) curl_forward(E_new)(p - 1) on the own cells, planes X0 .. X1 - 1
        if (own && p >= X0 + 1) {
            const int r = q ^ 1;
            const T ex = en_prev[0], ey = en_prev[1], ez = en_prev[2];
            const T dEz_dy = (sEn[r][2][t + YS_NZ] - ez) * idy;
            const T dEy_dz = (sEn[r][1][t + 1] - ey) * idz;
            const T dEx_dz = (sEn[r][0][t + 1] - ex) * idz;
            const T dEx_dy = (sEn[r][0][t + YS_NZ] - ex) * idy;
            const T dEz_dx = (en[2] - ez) * idx;
            const T dEy_dx = (en[1] - ey) * idx;
            const T bxn = bp_prev[0] - hdt * (dEz_dy - dEy_dz);
            const T byn = bp_prev[1] - hdt * (dEx_dz - dEz_dx);
            const T bzn = bp_prev[2] - hdt * (dEy_dx - dEx_dy);
            store6(Bx2, By2, Bz2, p - 1, bxn, byn, bzn);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            bp_prev[c] = bp[c]; en_prev[c] = en[c];
            eo[c] = eo1[c]; eo1[c] = eo2[c]; bo[c] = bo1[c]; jc[c] = j1[c];
        }
    }
}

template <typename T>
static int launch_yee_stream(const PicParams* p, const Dims& d, const YeeSides& sd, const void* const E[3], const void* const B[3],
                             const void* const J[3], void* const E2[3], void* const B2[3], cudaStream_t st) {
    // planes per CTA: enough CTAs to fill the chip several times over, few enough redundant rim planes
    int xc = PIC_YEE_XC;
    const int nby = (d.W[1] + YS_TY - 1) / YS_TY, nbz = (d.W[2] + YS_TZ - 1) / YS_TZ;
    while (xc > 8 && (int64_t)((d.W[0] + xc - 1) / xc) * nby * nbz < 4 * 148) xc /= 2;
    const int nbx = (d.W[0] + xc - 1) / xc;
    k_yee_stream<T><<<nbx * nby * nbz, YS_THREADS, 0, st>>>(d, sd, (const T*)E[0], (const T*)E[1], (const T*)E[2], (const T*)B[0], (const T*)B[1],
                                                            (const T*)B[2], (const T*)J[0], (const T*)J[1], (const T*)J[2], (T*)E2[0], (T*)E2[1],
                                                            (T*)E2[2], (T*)B2[0], (T*)B2[1], (T*)B2[2], (T)p->dt, (T)(p->dt / 2), (T)(1.0 / p->dx),
                                                            (T)(1.0 / p->dy), (T)(1.0 / p->dz), (T)(p->C * p->C), (T)(1.0 / p->eps), nby, nbz, xc);
    PIC_LAUNCH_RET();
}

template <typename T>
static int launch_yee_fused(const PicParams* p, const void* const E[3], const void* const B[3], const void* const J[3], void* const E2[3],
                            void* const B2[3], cudaStream_t st) {
    const Dims d = dims_of(p);
    YeeSides sd;
    for (int a = 0; a < 3; ++a) {
        const bool split = p->gmesh[a] != p->mesh[a];
        const bool periodic = p->field_bc[a] == PIC_BC_PERIODIC;
        const bool at_lo = p->moff[a] == 0, at_hi = p->moff[a] + p->mesh[a] == p->gmesh[a];
        // a periodic axis narrower than the guard depth (a reduced axis, W = 1) is served from its guard cells like a split axis;
        // the caller refreshes them afterwards (pic_halo_refresh_axis) -- Simulation does
        const bool wrap = periodic && !split && d.W[a] >= d.g;
        sd.lo[a] = periodic ? (wrap ? YEE_WRAP : YEE_HALO) : (at_lo ? YEE_WALL : YEE_HALO);
        sd.hi[a] = periodic ? (wrap ? YEE_WRAP : YEE_HALO) : (at_hi ? YEE_WALL : YEE_HALO);
        sd.cond_lo[a] = (p->field_bc[a] == PIC_BC_CONDUCTING && at_lo) ? 1 : 0;
        sd.cond_hi[a] = (p->field_bc[a] == PIC_BC_CONDUCTING && at_hi) ? 1 : 0;
    }
#ifndef PIC_YEE_TX
#define PIC_YEE_TX 2        /* x cells per tile (float): the kernel is latency-bound, more resident CTAs beat less redundancy */
#endif
#ifndef PIC_YEE_STREAM
#define PIC_YEE_STREAM 1    /* 1: the x-marching kernel (HBM-bound); 0: the tile kernel k_yee_fused (A/B control) */
#endif
    if (PIC_YEE_STREAM) return launch_yee_stream<T>(p, d, sd, E, B, J, E2, B2, st);
    if (sizeof(T) == 4 && d.W[0] >= 2 * PIC_YEE_TX) return launch_yee_fused_tx<T, PIC_YEE_TX>(p, d, sd, E, B, J, E2, B2, st);
    return launch_yee_fused_tx<T, 2>(p, d, sd, E, B, J, E2, B2, st);
}

// ---------------------------------------------------------------- divergence residuals (conservation diagnostics)
// out[A] = div_backward(F)[A] + ca * a[A] + cb * b[A] on every tile interior: with F = E, a = rho, ca = -1/eps this is the Gauss
// residual div E - rho/eps; with F = J, a = rho_new, b = rho_old, ca = -cb = 1/dt the discrete continuity residual
// (rho_new - rho_old)/dt + div J that Esirkepov's deposition keeps at round-off (the reference checks it in
// tests/code_tests/esirkepov_test.py:700-744).  Backward differences: E/J components sit half a cell above the nodes of rho.
template <typename T>
__global__ void __launch_bounds__(256) k_div_residual(Dims d, int nby, int nbz, const T* __restrict__ Fx, const T* __restrict__ Fy,
                                                      const T* __restrict__ Fz, const T* __restrict__ a, T ca, const T* __restrict__ b,
                                                      T cb, T* __restrict__ out, T dx, T dy, T dz) {
    int64_t tile; int ix, iy, iz;
    if (!cell_of_block(d, nby, nbz, tile, ix, iy, iz)) return;
    const size_t sx = (size_t)d.L[1] * d.L[2], sy = d.L[2];
    const size_t i = tile * d.tile_elems + ix * sx + iy * sy + iz;
    T r = (Fx[i] - Fx[i - sx]) / dx + (Fy[i] - Fy[i - sy]) / dy + (Fz[i] - Fz[i - 1]) / dz;
    if (a) r += ca * a[i];
    if (b) r += cb * b[i];
    out[i] = r;
}
template <typename T>
static int launch_div_residual(const PicParams* p, const void* const F[3], const void* a, double ca, const void* b, double cb, void* out,
                               cudaStream_t st) {
    const Dims d = dims_of(p);
    const CellGrid cg = cell_grid(d);
    k_div_residual<T><<<cg.grid, cg.block, 0, st>>>(d, cg.nby, cg.nbz, (const T*)F[0], (const T*)F[1], (const T*)F[2], (const T*)a, (T)ca,
                                                    (const T*)b, (T)cb, (T*)out, (T)p->dx, (T)p->dy, (T)p->dz);
    PIC_LAUNCH_RET();
}

}  // namespace pic

using namespace pic;

static bool halo_args_ok(const PicParams* p, int axis, int ncomp, void* const* fields) {
    if (!p || !fields || axis < 0 || axis > 2 || ncomp < 1 || ncomp > 8) return false;
    for (int c = 0; c < ncomp; ++c)
        if (!fields[c]) return false;
    return true;
}

extern "C" {

#ifndef PIC_SOURCE_HASH
#define PIC_SOURCE_HASH "unhashed"
#endif
// the hash of every CUDA source + header + flag this library was built from (pypic3d_b200/_lib.py source_hash()): the loader
// refuses a library that does not match the sources it sits next to
const char* pic_version(void) { return "pic_b200 0.2 (sm_100a) src " PIC_SOURCE_HASH; }
int pic_params_size(void) { return (int)sizeof(PicParams); }

int pic_update_E(const PicParams* p, void* const E[3], const void* const B[3], const void* const J[3], void* stream) {
    PIC_CHECK_ARG(p && E && B && J);
    PIC_DISPATCH_T(p, launch_update_E, p, E, B, J, (cudaStream_t)stream);
}

int pic_update_B(const PicParams* p, void* const B[3], const void* const E[3], void* stream) {
    PIC_CHECK_ARG(p && E && B);
    PIC_DISPATCH_T(p, launch_update_B, p, B, E, (cudaStream_t)stream);
}

int pic_filter(const PicParams* p, int kind, double alpha, const void* in, void* out, void* stream) {
    PIC_CHECK_ARG(p && in && out && in != out && (kind == PIC_FILTER_DIGITAL || kind == PIC_FILTER_BILINEAR) && p->g >= 1);
    PIC_DISPATCH_T(p, launch_filter, p, kind, alpha, in, out, (cudaStream_t)stream);
}

int pic_halo_refresh_axis(const PicParams* p, int axis, int bc, int ncomp, void* const* fields, void* stream) {
    PIC_CHECK_ARG(halo_args_ok(p, axis, ncomp, fields));
    PIC_DISPATCH_T(p, launch_refresh, p, axis, bc, ncomp, fields, (cudaStream_t)stream);
}

int pic_halo_fold_axis(const PicParams* p, int axis, int bc, int ncomp, void* const* fields, void* stream) {
    PIC_CHECK_ARG(halo_args_ok(p, axis, ncomp, fields));
    PIC_DISPATCH_T(p, launch_fold, p, axis, bc, ncomp, fields, (cudaStream_t)stream);
}

int pic_zero_wall(const PicParams* p, int axis, void* field, void* stream) {
    PIC_CHECK_ARG(p && field && axis >= 0 && axis <= 2);
    PIC_DISPATCH_T(p, launch_zero_wall, p, axis, field, (cudaStream_t)stream);
}

int pic_pack_planes(const PicParams* p, int axis, int start, int nplanes, int ncomp, const void* const* fields, void* buf,
                    void* stream) {
    PIC_CHECK_ARG(halo_args_ok(p, axis, ncomp, (void* const*)fields) && buf && start >= 0 && nplanes >= 0 &&
                  start + nplanes <= p->tile[axis] + 2 * p->g);
    PIC_DISPATCH_T(p, launch_planes, p, 0, axis, start, nplanes, ncomp, (void* const*)fields, buf, 0, (cudaStream_t)stream);
}

int pic_unpack_planes(const PicParams* p, int axis, int start, int nplanes, int ncomp, void* const* fields, const void* buf,
                      int mode, void* stream) {
    PIC_CHECK_ARG(halo_args_ok(p, axis, ncomp, fields) && buf && start >= 0 && nplanes >= 0 &&
                  start + nplanes <= p->tile[axis] + 2 * p->g && mode >= 0 && mode <= 2);
    PIC_DISPATCH_T(p, launch_planes, p, 1, axis, start, nplanes, ncomp, fields, (void*)buf, mode, (cudaStream_t)stream);
}

int pic_pack_boxes(const PicParams* p, int nbox, const int32_t* lo, const int32_t* size, int ncomp, const void* const* fields, void* buf,
                   void* stream) {
    PIC_CHECK_ARG(p && lo && size && fields && buf && nbox >= 0 && nbox <= 26 && ncomp >= 1 && ncomp <= 8);
    PIC_CHECK_ARG(p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    PIC_DISPATCH_T(p, launch_boxes, p, 0, nbox, lo, size, ncomp, (void* const*)fields, buf, 0, (cudaStream_t)stream);
}

int pic_unpack_boxes(const PicParams* p, int nbox, const int32_t* lo, const int32_t* size, int ncomp, void* const* fields, const void* buf,
                     int mode, void* stream) {
    PIC_CHECK_ARG(p && lo && size && fields && buf && nbox >= 0 && nbox <= 26 && ncomp >= 1 && ncomp <= 8 && mode >= 0 && mode <= 2);
    PIC_CHECK_ARG(p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    PIC_DISPATCH_T(p, launch_boxes, p, 1, nbox, lo, size, ncomp, fields, (void*)buf, mode, (cudaStream_t)stream);
}

int pic_yee_fused(const PicParams* p, const void* const E[3], const void* const B[3], const void* const J[3], void* const E_out[3],
                  void* const B_out[3], void* stream) {
    PIC_CHECK_ARG(p && E && B && J && E_out && B_out);
    PIC_CHECK_ARG(p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    if (p->g < 2 || p->alpha != 1.0) return PIC_EUNSUPPORTED;   // the E box reaches two guard cells; the digital filter needs the sweeps
    for (int c = 0; c < 3; ++c) PIC_CHECK_ARG(E[c] != E_out[c] && B[c] != B_out[c]);
    PIC_DISPATCH_T(p, launch_yee_fused, p, E, B, J, E_out, B_out, (cudaStream_t)stream);
}

int pic_div_residual(const PicParams* p, const void* const F[3], const void* a, double ca, const void* b, double cb, void* out,
                     void* stream) {
    PIC_CHECK_ARG(p && F && out && p->g >= 1);
    PIC_DISPATCH_T(p, launch_div_residual, p, F, a, ca, b, cb, out, (cudaStream_t)stream);
}

int pic_sum_squares_interior(const PicParams* p, const void* field, double* out, void* stream) {
    PIC_CHECK_ARG(p && field && out);
    PIC_DISPATCH_T(p, launch_sumsq, p, field, out, (cudaStream_t)stream);
}

}  // extern "C"
