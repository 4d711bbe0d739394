// microbench.cu -- small design-evidence microbenchmarks (results are committed under profiles/).
// They answer the questions the K1 design hinges on: how fast are shared-memory / global floating-point
// reductions for the address patterns a cell-sorted particle stream produces, and how fast is an L1-resident gather.
#include "pic_common.cuh"

namespace pic {

constexpr int MB_TILE = 11 * 11 * 11 * 3;  // J accumulator of an 8^3 supercell with CIC reach (3993 reals)
constexpr int MB_OPS = 64;

__device__ __forceinline__ uint32_t mb_hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// pattern 0: every lane its own pseudo-random address; 1: groups of 8 consecutive lanes share an address (8 ppc run);
// 2: whole warp one address.
__device__ __forceinline__ uint32_t mb_addr(int pattern, int it, uint32_t range) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t key = pattern == 0 ? t : (pattern == 1 ? (t >> 3) : (t >> 5));
    return mb_hash(key * 977u + it * 131071u) % range;
}

template <typename T>
__global__ void __launch_bounds__(256) k_mb_smem_atomic(int pattern, T* out) {
    __shared__ T tile[MB_TILE];
    for (int i = threadIdx.x; i < MB_TILE; i += blockDim.x) tile[i] = (T)0;
    __syncthreads();
    for (int it = 0; it < MB_OPS; ++it) atomicAdd(&tile[mb_addr(pattern, it, MB_TILE)], (T)1);
    __syncthreads();
    T acc = 0;
    for (int i = threadIdx.x; i < MB_TILE; i += blockDim.x) acc += tile[i];
    if (acc == (T)-1) out[0] = acc;
}

// global reductions into a window of `range` reals per CTA (a cell-sorted stream touches a small moving window)
template <typename T>
__global__ void __launch_bounds__(256) k_mb_gmem_atomic(int pattern, T* buf, uint32_t range, uint32_t nwin) {
    T* win = buf + (size_t)(blockIdx.x % nwin) * range;
    for (int it = 0; it < MB_OPS; ++it) atomicAdd(&win[mb_addr(pattern, it, range)], (T)1);
}

// segmented warp reduction over runs of 8 lanes followed by one atomic per run (candidate replacement for pattern 1)
template <typename T>
__global__ void __launch_bounds__(256) k_mb_shfl_then_atomic(T* buf, uint32_t range, uint32_t nwin) {
    T* win = buf + (size_t)(blockIdx.x % nwin) * range;
    for (int it = 0; it < MB_OPS; ++it) {
        T v = (T)1;
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if ((threadIdx.x & 7) == 0) atomicAdd(&win[mb_addr(1, it, range)], v);
    }
}

// L1-resident gather: 48 dependent-address loads per thread from a 12^3 x 6 window (the CIC E/B stencil of a supercell)
template <typename T>
__global__ void __launch_bounds__(256) k_mb_gather(const T* __restrict__ buf, uint32_t range, uint32_t nwin, T* out) {
    const T* win = buf + (size_t)(blockIdx.x % nwin) * range;
    T acc = 0;
    for (int it = 0; it < MB_OPS; ++it) acc += __ldg(&win[mb_addr(1, it, range)]);
    if (acc == (T)-1) out[0] = acc;
}

template <typename F>
static float time_launches(int iters, F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return ms;
}

}  // namespace pic

using namespace pic;

extern "C" int pic_microbench(int which, int iters, float* ms_out) {
    if (!ms_out || iters < 1) return PIC_EINVAL;
    const int grid = num_sms() * 16, block = 256;
    const uint32_t range = 12 * 12 * 12 * 6, nwin = 4096;
    void* buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)range * nwin * sizeof(double));
    if (e != cudaSuccess) return (int)e;
    cudaMemset(buf, 0, (size_t)range * nwin * sizeof(double));
    float ms = -1.f;
    switch (which) {
        case 0: case 1: case 2: ms = time_launches(iters, [&] { k_mb_smem_atomic<float><<<grid, block>>>(which, (float*)buf); }); break;
        case 3: case 4: case 5: ms = time_launches(iters, [&] { k_mb_smem_atomic<double><<<grid, block>>>(which - 3, (double*)buf); }); break;
        case 6: case 7: case 8: ms = time_launches(iters, [&] { k_mb_gmem_atomic<float><<<grid, block>>>(which - 6, (float*)buf, range, nwin); }); break;
        case 9: case 10: case 11: ms = time_launches(iters, [&] { k_mb_gmem_atomic<double><<<grid, block>>>(which - 9, (double*)buf, range, nwin); }); break;
        case 12: ms = time_launches(iters, [&] { k_mb_shfl_then_atomic<float><<<grid, block>>>((float*)buf, range, nwin); }); break;
        case 13: ms = time_launches(iters, [&] { k_mb_gather<float><<<grid, block>>>((const float*)buf, range, nwin, (float*)buf); }); break;
        default: cudaFree(buf); return PIC_EINVAL;
    }
    e = cudaDeviceSynchronize();
    cudaFree(buf);
    *ms_out = ms;
    return (int)e;
}
