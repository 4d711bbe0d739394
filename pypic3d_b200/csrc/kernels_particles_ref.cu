// kernels_particles_ref.cu -- reference-layout particle operators (one thread per TiledParticles slot).
// These are the drop-ins for the reference's sub-entry points; they work on any local tile mesh and keep
// slot identity exactly as the reference does.  The throughput path is kernels_fast.cu.
#include "pic_common.cuh"

namespace pic {

// ---------------------------------------------------------------- push (particle_push.py:45-144)
template <typename T, int SF>
__global__ void __launch_bounds__(256) k_push_ref(const __grid_constant__ PicParams p, const T* __restrict__ x,
                                                  const T* __restrict__ u_in, T* __restrict__ u_out,
                                                  const uint8_t* __restrict__ active, int64_t cap, Field6<T> F) {
    const int64_t total = (int64_t)p.mesh[0] * p.mesh[1] * p.mesh[2] * p.n_species * cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        slot_push<T, SF>(p, i, x, u_in, u_out, active, cap, F);
}

template <typename T, int SF>
static int launch_push(const PicParams* p, const void* x, const void* u_in, void* u_out, const uint8_t* active,
                       int64_t cap, const void* const E[3], const void* const B[3], cudaStream_t st) {
    Field6<T> F;
    for (int c = 0; c < 3; ++c) { F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c]; }
    const int64_t total = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2] * p->n_species * cap;
    if (total == 0) return 0;
    k_push_ref<T, SF><<<grid_for(total, 256), 256, 0, st>>>(*p, (const T*)x, (const T*)u_in, (T*)u_out, active, cap, F);
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- deposits; mode 0: Esirkepov, 1: direct J, 2: rho
template <typename T, int SF, int MODE>
__global__ void __launch_bounds__(256) k_deposit_ref(const __grid_constant__ PicParams p, const T* __restrict__ x,
                                                     const T* __restrict__ u, const uint8_t* __restrict__ active,
                                                     int64_t cap, Field3W<T> J) {
    const int64_t total = (int64_t)p.mesh[0] * p.mesh[1] * p.mesh[2] * p.n_species * cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        slot_deposit<T, SF, MODE>(p, i, x, u, active, cap, J);
}

template <typename T, int SF>
static int launch_deposit(const PicParams* p, int mode, const void* x, const void* u, const uint8_t* active, int64_t cap,
                          void* const J[3], int ncomp, cudaStream_t st) {
    Field3W<T> F;
    for (int c = 0; c < 3; ++c) F.f[c] = (T*)J[c < ncomp ? c : 0];
    const int64_t total = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2] * p->n_species * cap;
    if (total == 0) return 0;
    const int grid = grid_for(total, 256);
    if (mode == 0) k_deposit_ref<T, SF, 0><<<grid, 256, 0, st>>>(*p, (const T*)x, (const T*)u, active, cap, F);
    else if (mode == 1) k_deposit_ref<T, SF, 1><<<grid, 256, 0, st>>>(*p, (const T*)x, (const T*)u, active, cap, F);
    else k_deposit_ref<T, SF, 2><<<grid, 256, 0, st>>>(*p, (const T*)x, (const T*)u, active, cap, F);
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- move (particle_tile_communication.py:82-99)
template <typename T>
__global__ void __launch_bounds__(256) k_move_ref(const __grid_constant__ PicParams p, const T* __restrict__ x_in,
                                                  T* __restrict__ x_out, const T* __restrict__ u,
                                                  const uint8_t* __restrict__ active, int64_t cap, T dt) {
    const int64_t total = (int64_t)p.mesh[0] * p.mesh[1] * p.mesh[2] * p.n_species * cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        slot_move<T>(p, i, x_in, x_out, u, active, cap, dt);
}

template <typename T>
static int launch_move(const PicParams* p, const void* x_in, void* x_out, const void* u, const uint8_t* active, int64_t cap,
                       double dt, cudaStream_t st) {
    const int64_t total = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2] * p->n_species * cap;
    if (total == 0) return 0;
    k_move_ref<T><<<grid_for(total, 256), 256, 0, st>>>(*p, (const T*)x_in, (T*)x_out, (const T*)u, active, cap, (T)dt);
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- retile (particle_tile_communication.py:233-426)
template <typename T>
__global__ void __launch_bounds__(256) k_retile_classify(const __grid_constant__ PicParams p, const T* __restrict__ x_in,
                                                         const T* __restrict__ u_in, const uint8_t* __restrict__ active_in,
                                                         T* __restrict__ x_out, T* __restrict__ u_out,
                                                         uint8_t* __restrict__ active_out, int64_t cap,
                                                         int32_t* __restrict__ code, int32_t* overflow) {
    const int64_t total = (int64_t)p.mesh[0] * p.mesh[1] * p.mesh[2] * p.n_species * cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        slot_retile_classify<T>(p, i, x_in, u_in, active_in, x_out, u_out, active_out, cap, code, overflow);
}

// exclusive block scan of one int per thread (blockDim.x == 256); returns the prefix, total via *total.
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* smem /*>= 9 ints*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();  // protect smem reuse across calls
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < (int)(blockDim.x >> 5)) smem[lane] = winc - w;
        if (lane == 31) smem[8] = winc;
    }
    __syncthreads();
    *total = smem[8];
    return smem[warp] + inc - v;
}

// One CTA per destination (tile, species): k-th incoming particle (stream order, then source slot order) goes to the
// k-th free slot (free = not *staying*); surplus sets the overflow flag (_fill_incoming_particles, :169-230).
template <typename T>
__global__ void __launch_bounds__(256) k_retile_fill(const __grid_constant__ PicParams p, const T* __restrict__ x_in,
                                                     const T* __restrict__ u_in, T* __restrict__ x_out, T* __restrict__ u_out,
                                                     uint8_t* __restrict__ active_out, int64_t cap,
                                                     const int32_t* __restrict__ code, int32_t* __restrict__ free_list,
                                                     int32_t* overflow) {
    __shared__ int sm[9];
    const int S = p.n_species;
    const int64_t dest_tile_ = blockIdx.x / S;
    const int s = blockIdx.x % S;
    const TileCoord tc = tile_coord(dest_tile_, p.mesh);
    const int dt_[3] = {tc.tx, tc.ty, tc.tz};
    const int64_t dbase = (dest_tile_ * S + s) * cap;
    // 1. ordered list of free slots of the destination
    int n_free = 0;
    for (int64_t c0 = 0; c0 < cap; c0 += blockDim.x) {
        const int64_t slot = c0 + threadIdx.x;
        const int fr = (slot < cap && code[dbase + slot] != 1) ? 1 : 0;
        int tot;
        const int pre = block_excl_scan(fr, &tot, sm);
        if (fr) free_list[dbase + n_free + pre] = (int32_t)slot;
        n_free += tot;
    }
    __syncthreads();
    // 2. incoming streams in the reference's order
    int rank0 = 0;
    for (int sx = 0; sx < 3; ++sx)
        for (int sy = 0; sy < 3; ++sy)
            for (int sz = 0; sz < 3; ++sz) {
                const int o[3] = {1 - sx, 1 - sy, 1 - sz};
                if (o[0] == 0 && o[1] == 0 && o[2] == 0) continue;
                int src[3];
                bool ok = true;
                for (int c = 0; c < 3; ++c) {
                    const int n = p.mesh[c];
                    if (n == 1) { if (o[c] != 0) ok = false; src[c] = dt_[c]; continue; }   // _movement_offsets(1) == (0,)
                    int sc = dt_[c] - o[c];
                    if (o[c] != 0) {
                        if (p.particle_bc[c] == PIC_BC_PERIODIC) sc = ((sc % n) + n) % n;
                        else if (sc < 0 || sc >= n) ok = false;
                    }
                    src[c] = sc;
                }
                if (!ok) continue;
                const int stream_code = 2 + (sx * 3 + sy) * 3 + sz;
                const int64_t sbase = ((((int64_t)src[0] * p.mesh[1] + src[1]) * p.mesh[2] + src[2]) * S + s) * cap;
                for (int64_t c0 = 0; c0 < cap; c0 += blockDim.x) {
                    const int64_t slot = c0 + threadIdx.x;
                    const int mv = (slot < cap && code[sbase + slot] == stream_code) ? 1 : 0;
                    int tot;
                    const int pre = block_excl_scan(mv, &tot, sm);
                    if (mv) {
                        const int rank = rank0 + pre;
                        if (rank < n_free) {
                            const int64_t j = dbase + free_list[dbase + rank];
                            const int64_t i = sbase + slot;
                            T pos[3], vel[3];
                            for (int c = 0; c < 3; ++c) { pos[c] = x_in[3 * i + c]; vel[c] = u_in[3 * i + c]; }
                            bounded_state<T>(p, pos, vel);
                            for (int c = 0; c < 3; ++c) { x_out[3 * j + c] = pos[c]; u_out[3 * j + c] = vel[c]; }
                            active_out[j] = 1;
                        } else {
                            atomicOr(overflow, 1);
                        }
                    }
                    rank0 += tot;
                }
            }
}

template <typename T>
static int launch_retile(const PicParams* p, const void* x_in, const void* u_in, const uint8_t* active_in, void* x_out,
                         void* u_out, uint8_t* active_out, int64_t cap, int32_t* scratch, int32_t* overflow, cudaStream_t st) {
    const int64_t ntiles = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2];
    const int64_t total = ntiles * p->n_species * cap;
    if (total == 0) return 0;
    int32_t* code = scratch;
    int32_t* free_list = scratch + total;
    k_retile_classify<T><<<grid_for(total, 256), 256, 0, st>>>(*p, (const T*)x_in, (const T*)u_in, active_in, (T*)x_out,
                                                               (T*)u_out, active_out, cap, code, overflow);
    if (ntiles > 1) {
        k_retile_fill<T><<<(unsigned)(ntiles * p->n_species), 256, 0, st>>>(*p, (const T*)x_in, (const T*)u_in, (T*)x_out,
                                                                           (T*)u_out, active_out, cap, code, free_list, overflow);
    }
    PIC_LAUNCH_RET();
}

// ---------------------------------------------------------------- particle energy / momentum (utils.py:170-202)
template <typename T>
__global__ void __launch_bounds__(256) k_particle_energy(const __grid_constant__ PicParams p, const T* __restrict__ u,
                                                         const uint8_t* __restrict__ active, int64_t cap, double* out) {
    const int64_t total = (int64_t)p.mesh[0] * p.mesh[1] * p.mesh[2] * p.n_species * cap;
    double ke = 0.0, mom = 0.0;
    const double C = p.C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (!active[i]) continue;
        const int s = (int)((i / cap) % p.n_species);
        const double m = p.mass[s] * p.weight[s];
        const double vx = (double)u[3 * i], vy = (double)u[3 * i + 1], vz = (double)u[3 * i + 2];
        const double v2 = vx * vx + vy * vy + vz * vz;
        const double gamma = 1.0 / sqrt(1.0 - v2 / (C * C));
        const double p2 = (m * gamma) * (m * gamma) * v2;
        ke += sqrt(p2 * C * C + m * m * C * C * C * C) - m * C * C;
        mom += sqrt(v2) * m;
    }
    ke = warp_sum(ke);
    mom = warp_sum(mom);
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, ke); atomicAdd(out + 1, mom); }
}

}  // namespace pic

using namespace pic;

template <typename T, int SF>
static int dep_esir(const PicParams* p, const void* x, const void* u, const uint8_t* a, int64_t cap, void* const J[3], cudaStream_t st) {
    return launch_deposit<T, SF>(p, 0, x, u, a, cap, J, 3, st);
}
template <typename T, int SF>
static int dep_direct(const PicParams* p, const void* x, const void* u, const uint8_t* a, int64_t cap, void* const J[3], cudaStream_t st) {
    return launch_deposit<T, SF>(p, 1, x, u, a, cap, J, 3, st);
}
template <typename T, int SF>
static int dep_rho(const PicParams* p, const void* x, const uint8_t* a, int64_t cap, void* rho, cudaStream_t st) {
    void* J[3] = {rho, rho, rho};
    return launch_deposit<T, SF>(p, 2, x, x, a, cap, J, 1, st);
}

template <typename T>
static int launch_penergy(const PicParams* p, const void* u, const uint8_t* active, int64_t cap, double* out, cudaStream_t st) {
    const int64_t total = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2] * p->n_species * cap;
    if (total == 0) return 0;
    k_particle_energy<T><<<grid_for(total, 256, 4), 256, 0, st>>>(*p, (const T*)u, active, cap, out);
    PIC_LAUNCH_RET();
}

extern "C" {

int pic_push(const PicParams* p, const void* x, const void* u_in, void* u_out, const uint8_t* active, int64_t cap,
             const void* const E[3], const void* const B[3], void* stream) {
    PIC_CHECK_ARG(p && x && u_in && u_out && active && E && B && cap >= 0 && p->n_species <= PIC_MAX_SPECIES);
    PIC_DISPATCH_T_SF(p, launch_push, p, x, u_in, u_out, active, cap, E, B, (cudaStream_t)stream);
}

int pic_deposit_esirkepov(const PicParams* p, const void* x, const void* u, const uint8_t* active, int64_t cap,
                          void* const J[3], void* stream) {
    PIC_CHECK_ARG(p && x && u && active && J && cap >= 0 && p->n_species <= PIC_MAX_SPECIES);
    PIC_DISPATCH_T_SF(p, dep_esir, p, x, u, active, cap, J, (cudaStream_t)stream);
}

int pic_deposit_direct(const PicParams* p, const void* x, const void* u, const uint8_t* active, int64_t cap,
                       void* const J[3], void* stream) {
    PIC_CHECK_ARG(p && x && u && active && J && cap >= 0 && p->n_species <= PIC_MAX_SPECIES);
    PIC_DISPATCH_T_SF(p, dep_direct, p, x, u, active, cap, J, (cudaStream_t)stream);
}

int pic_deposit_rho(const PicParams* p, const void* x, const uint8_t* active, int64_t cap, void* rho, void* stream) {
    PIC_CHECK_ARG(p && x && active && rho && cap >= 0 && p->n_species <= PIC_MAX_SPECIES);
    PIC_DISPATCH_T_SF(p, dep_rho, p, x, active, cap, rho, (cudaStream_t)stream);
}

int pic_move(const PicParams* p, const void* x_in, void* x_out, const void* u, const uint8_t* active, int64_t cap, double dt,
             void* stream) {
    PIC_CHECK_ARG(p && x_in && x_out && u && active && cap >= 0 && p->n_species <= PIC_MAX_SPECIES);
    PIC_DISPATCH_T(p, launch_move, p, x_in, x_out, u, active, cap, dt, (cudaStream_t)stream);
}

int pic_retile(const PicParams* p, const void* x_in, const void* u_in, const uint8_t* active_in, void* x_out, void* u_out,
               uint8_t* active_out, int64_t cap, int32_t* scratch, int32_t* overflow, void* stream) {
    PIC_CHECK_ARG(p && x_in && u_in && active_in && x_out && u_out && active_out && scratch && overflow && cap >= 0);
    PIC_CHECK_ARG(x_in != x_out && u_in != u_out && p->n_species <= PIC_MAX_SPECIES);
    for (int c = 0; c < 3; ++c) PIC_CHECK_ARG(p->mesh[c] == p->gmesh[c] && p->moff[c] == 0);
    PIC_DISPATCH_T(p, launch_retile, p, x_in, u_in, active_in, x_out, u_out, active_out, cap, scratch, overflow,
                   (cudaStream_t)stream);
}

int pic_particle_energy(const PicParams* p, const void* u, const uint8_t* active, int64_t cap, double* out, void* stream) {
    PIC_CHECK_ARG(p && u && active && out && cap >= 0 && p->n_species <= PIC_MAX_SPECIES);
    PIC_DISPATCH_T(p, launch_penergy, p, u, active, cap, out, (cudaStream_t)stream);
}

}  // extern "C"
