// pic_tma.cuh -- mbarrier / TMA (cp.async.bulk[.tensor]) helpers shared by the supercell-tile kernels (K1 v9 k_tile3d, K1 v10
// k_pair3d) and the tensor-map encoder both launchers use.  Device-only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pic_slots.cuh"

namespace pic {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// mbarrier (shared::cta) + TMA helpers for the tile pipeline
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, int bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "MBAR_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra MBAR_DONE;\n"
        " bra MBAR_WAIT;\n"
        "MBAR_DONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one bounded wait: true once the phase with `parity` has completed; otherwise the hardware may suspend the thread for up to
// `hint_ns` before returning false (no issue slots burnt while waiting)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, int parity, int hint_ns) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        " selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, int parity) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n"
        " mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// TMA: one [x 8][y TILE_NY][z 8] box of a field component -> shared memory, completion signalled on `bar` (complete_tx)
__device__ __forceinline__ void tma_load_box(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int z, int y, int x) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(z), "r"(y), "r"(x)
                 : "memory");
}
// TMA bulk copy of `bytes` contiguous bytes (16-byte aligned, multiple of 16) global -> shared, completion on `bar`
__device__ __forceinline__ void tma_load_bytes(void* smem_dst, const void* gmem_src, int bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 eviction-priority policy for data that is touched once (particle slices)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_bytes_hint(void* smem_dst, const void* gmem_src, int bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st_f2_hint(float* ptr, float a, float b, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;\n" ::"l"(ptr), "f"(a), "f"(b), "l"(pol) : "memory");
}
// TMA reduce: global J box (8 x 8 x 8 nodes at z, y, x) += the shared-memory tile (element-wise add performed in L2)
__device__ __forceinline__ void tma_reduce_add_box(const CUtensorMap* map, const void* smem_src, int z, int y, int x) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];\n"
                 ::"l"((uint64_t)map), "r"(smem_u32(smem_src)), "r"(z), "r"(y), "r"(x)
                 : "memory");
}
__device__ __forceinline__ void red_shared_add(float* p, float v) {
    asm volatile("red.shared.add.f32 [%0], %1;\n" ::"r"(smem_u32(p)), "f"(v) : "memory");
}
__device__ __forceinline__ void red_shared_add(double* p, double v) {
    asm volatile("red.shared.add.f64 [%0], %1;\n" ::"r"(smem_u32(p)), "d"(v) : "memory");
}
struct TileMaps {
    CUtensorMap m[6];      // Ex Ey Ez Bx By Bz, each the ghosted (Lx, Ly, Lz) tile with an 8 x TILE_NY x 8 box
    CUtensorMap j[3];      // Jx Jy Jz with an 8 x 8 x 8 box (TMA reduce target of the shared-memory J tiles)
};


typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline PFN_tensorMapEncodeTiled tensor_map_encoder() {
    static PFN_tensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return nullptr;
        encode = (PFN_tensorMapEncodeTiled)fn;
    }
    return encode;
}

// TMA descriptors of the six ghosted E/B tiles (box TILE_N x TILE_NY x TILE_N) and the three J tiles (box TILE_N^3); false when the
// driver has no tensor-map entry point or rejects the layout
template <typename T>
static inline bool make_tile_maps(const int L[3], const T* const F[6], T* const J[3], TileMaps& tm) {
    PFN_tensorMapEncodeTiled encode = tensor_map_encoder();
    if (!encode) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)L[2], (cuuint64_t)L[1], (cuuint64_t)L[0]};
    const cuuint64_t strides[2] = {(cuuint64_t)L[2] * sizeof(T), (cuuint64_t)L[2] * L[1] * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)TILE_N, (cuuint32_t)TILE_NY, (cuuint32_t)TILE_N};
    const cuuint32_t jbox[3] = {(cuuint32_t)TILE_N, (cuuint32_t)TILE_N, (cuuint32_t)TILE_N};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    for (int c = 0; c < 6; ++c)
        if (encode(&tm.m[c], dt, 3, const_cast<T*>(F[c]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    for (int c = 0; c < 3; ++c)
        if (encode(&tm.j[c], dt, 3, J[c], dims, strides, jbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    return true;
}

}  // namespace pic
