// pic_common.cuh -- launch helpers shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pic_math.cuh"
#include "pic_slots.cuh"

#define PIC_CHECK_ARG(cond) \
    do {                    \
        if (!(cond)) return PIC_EINVAL; \
    } while (0)

#define PIC_LAUNCH_RET()                         \
    do {                                         \
        cudaError_t e__ = cudaPeekAtLastError(); \
        return (int)e__;                         \
    } while (0)

// Dispatch a templated launcher on (dtype, shape_factor).
#define PIC_DISPATCH_T_SF(p, FN, ...)                                                   \
    do {                                                                                \
        if ((p)->dtype == PIC_F32) {                                                    \
            if ((p)->shape_factor == 1) return FN<float, 1>(__VA_ARGS__);               \
            if ((p)->shape_factor == 2) return FN<float, 2>(__VA_ARGS__);               \
        } else if ((p)->dtype == PIC_F64) {                                             \
            if ((p)->shape_factor == 1) return FN<double, 1>(__VA_ARGS__);              \
            if ((p)->shape_factor == 2) return FN<double, 2>(__VA_ARGS__);              \
        }                                                                               \
        return PIC_EUNSUPPORTED;                                                        \
    } while (0)

#define PIC_DISPATCH_T(p, FN, ...)                                    \
    do {                                                              \
        if ((p)->dtype == PIC_F32) return FN<float>(__VA_ARGS__);     \
        if ((p)->dtype == PIC_F64) return FN<double>(__VA_ARGS__);    \
        return PIC_EUNSUPPORTED;                                      \
    } while (0)

namespace pic {

static inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// Grid-stride launch geometry: enough CTAs to cover `n`, capped at a multiple of the SM count.
static inline int grid_for(int64_t n, int block, int ctas_per_sm = 8) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pic
