// kernels_fast.cu -- the resident throughput path: one tile per GPU, compact cell-sorted SoA particles.
//   K1  k_fused        gather E,B -> push -> deposit -> move -> particle BC (+ leaver extraction), one pass over
//                      particle memory: reads x,y,z,vx,vy,vz once, writes them once (48 B f32 / 96 B f64 per particle)
//   K2  k_sort_*       counting sort by local cell (histogram, exclusive scan, scatter), run every few steps
//   import / export    TiledParticles (reference AoS + active mask) <-> compact SoA
#include <stdlib.h>
#include <cuda.h>

#include "pic_common.cuh"
#include "pic_tma.cuh"

namespace pic {

// ---------------------------------------------------------------- import / export
template <typename T>
__global__ void __launch_bounds__(256) k_import(int species, int n_species, const T* __restrict__ x, const T* __restrict__ u,
                                                const uint8_t* __restrict__ active, int64_t cap_ref, SoAView<T> s, int32_t* count) {
    for (int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; slot < cap_ref; slot += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = (int64_t)species * cap_ref + slot;
        if (!active[i]) continue;
        const int64_t j = atomicAdd(count, 1);
        if (j >= s.cap) continue;  // caller checks count <= cap
        for (int c = 0; c < 3; ++c) { s.c[c][j] = x[3 * i + c]; s.c[3 + c][j] = u[3 * i + c]; }
        if (s.id) s.id[j] = (int32_t)slot;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_export(int species, SoAView<T> s, T* __restrict__ x, T* __restrict__ u,
                                                uint8_t* __restrict__ active, int64_t cap_ref, int32_t* count) {
    const int64_t n = s.count();
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const T px = s.c[0][j];
        if (pic_isnan(px)) continue;
        int64_t slot;
        if (s.id && s.id[j] >= 0) { slot = s.id[j]; atomicAdd(count, 1); }
        else slot = atomicAdd(count, 1);
        if (slot >= cap_ref) continue;
        const int64_t i = (int64_t)species * cap_ref + slot;
        for (int c = 0; c < 3; ++c) { x[3 * i + c] = s.c[c][j]; u[3 * i + c] = s.c[3 + c][j]; }
        active[i] = 1;
    }
}

// ---------------------------------------------------------------- counting sort by local cell
// The stream is nearly sorted when it is re-sorted (a few per cent of the particles have changed cell), so the 32 consecutive
// particles of a warp form a handful of RUNS of equal cell: the head lane of a run issues ONE atomic for the whole run (the
// histogram's add; the scatter's returning cursor add, whose result the run shares by shuffle) and the members of a run write
// consecutive destination slots.  Runs are found with one shuffle and one ballot (neighbour compare) -- the __match_any_sync
// grouping tried in round 1 cost more than the atomics it saved.  PIC_SORT_RUNS=0 keeps the one-atomic-per-particle form.
#ifndef PIC_SORT_RUNS
#define PIC_SORT_RUNS 1
#endif
struct WarpRun {
    int head_lane;     // lane of the first particle of my run
    int len;           // (head lanes) particles in the run
    bool head;
};
__device__ __forceinline__ WarpRun warp_run(int cell, bool valid, int lane) {
    const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);          // (valid lanes are a prefix of the warp)
    WarpRun r;
    r.head = valid && (lane == 0 || cell != prev);
    const unsigned heads = __ballot_sync(0xffffffffu, r.head);
    const unsigned below = heads & (0xffffffffu >> (31 - lane));       // heads at or below my lane
    r.head_lane = below ? 31 - __clz(below) : 0;
    const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
    r.len = above ? __ffs(above) : __popc(vmask) - lane;
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256) k_sort_hist(const __grid_constant__ PicParams p, SoAView<T> s, int32_t* __restrict__ count) {
    const int64_t n = s.count();
#if PIC_SORT_RUNS
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x - lane; b < n; b += stride) {
        const int64_t j = b + lane;
        const bool valid = j < n;
        const int cell = valid ? local_cell<T>(p, s.c[0][j], s.c[1][j], s.c[2][j]) : -1;
        const WarpRun r = warp_run(cell, valid, lane);
        if (r.head) atomicAdd(&count[cell], r.len);
    }
#else
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&count[local_cell<T>(p, s.c[0][j], s.c[1][j], s.c[2][j])], 1);
#endif
}

template <typename T>
__global__ void __launch_bounds__(256) k_sort_scatter(const __grid_constant__ PicParams p, SoAView<T> s, SoAView<T> d,
                                                      const int32_t* __restrict__ offset, int32_t* __restrict__ cursor) {
    const int64_t n = s.count();
#if PIC_SORT_RUNS
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x - lane; b < n; b += stride) {
        const int64_t j = b + lane;
        const bool valid = j < n;
        T px = (T)0, py = (T)0, pz = (T)0;
        int cell = -1;
        if (valid) {
            px = s.c[0][j]; py = s.c[1][j]; pz = s.c[2][j];
            cell = local_cell<T>(p, px, py, pz);
        }
        const WarpRun r = warp_run(cell, valid, lane);
        int base = 0;
        if (r.head) base = atomicAdd(&cursor[cell], r.len);
        base = __shfl_sync(0xffffffffu, base, r.head_lane);
        if (!valid) continue;
        const int64_t dst = (int64_t)offset[cell] + base + (lane - r.head_lane);
        if (dst >= d.cap) continue;
        d.c[0][dst] = px; d.c[1][dst] = py; d.c[2][dst] = pz;
        d.c[3][dst] = s.c[3][j]; d.c[4][dst] = s.c[4][j]; d.c[5][dst] = s.c[5][j];
        if (s.id && d.id) d.id[dst] = s.id[j];
    }
#else
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const T px = s.c[0][j], py = s.c[1][j], pz = s.c[2][j];
        const int cell = local_cell<T>(p, px, py, pz);
        const int64_t dst = (int64_t)offset[cell] + atomicAdd(&cursor[cell], 1);
        if (dst >= d.cap) continue;
        d.c[0][dst] = px; d.c[1][dst] = py; d.c[2][dst] = pz;
        d.c[3][dst] = s.c[3][j]; d.c[4][dst] = s.c[4][j]; d.c[5][dst] = s.c[5][j];
        if (s.id && d.id) d.id[dst] = s.id[j];
    }
#endif
}

// exclusive scan of int32: 2048 elements per CTA (256 threads x 8), block sums scanned by one CTA, then added back.
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = 256 * SCAN_ITEMS;

__device__ __forceinline__ int block_scan_excl_256(int v, int* total, int* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < 8 ? sm[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < 8) sm[lane] = winc - w;
        if (lane == 7) sm[8] = winc;
    }
    __syncthreads();
    *total = sm[8];
    return sm[warp] + inc - v;
}

__global__ void __launch_bounds__(256) k_scan_local(int64_t n, const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                    int32_t* __restrict__ block_sums) {
    __shared__ int sm[9];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    int total;
    int pre = block_scan_excl_256(sum, &total, sm);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = pre;
        pre += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_scan_sums(int64_t nblocks, int32_t* __restrict__ block_sums) {
    __shared__ int sm[9];
    int carry = 0;
    for (int64_t c0 = 0; c0 < nblocks; c0 += 256) {
        const int64_t i = c0 + threadIdx.x;
        const int v = i < nblocks ? block_sums[i] : 0;
        int total;
        const int pre = block_scan_excl_256(v, &total, sm);
        if (i < nblocks) block_sums[i] = carry + pre;
        carry += total;
    }
}

__global__ void __launch_bounds__(256) k_scan_add(int64_t n, int32_t* __restrict__ out, const int32_t* __restrict__ block_sums) {
    const int add = block_sums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// ---------------------------------------------------------------- K1: fused gather + push + deposit + move + BC
template <typename T, int SF, int DEP, bool ALL3D>
__global__ void __launch_bounds__(256) k_fused(const __grid_constant__ PicParams p, int species, SoAView<T> s, Field6<T> F,
                                               Field6<T> X, int has_ext, Field3W<T> J, LeaveBuf leave, int32_t* flags) {
    Geom<T> gm;
    make_geom<T>(p, 0, 0, 0, gm);
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = J.f[c]; sink.L[c] = gm.L[c]; }
    sink.off = 0;
    sink.flags = flags;
    bool distributed = false;
    for (int c = 0; c < 3; ++c) distributed |= (p.gmesh[c] != p.mesh[c]);
    const int64_t n = s.count();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        fused_particle<T, SF, DEP, ALL3D>(p, species, gm, i, s, F, X, has_ext, sink, leave, distributed, flags);
}

// K1 v3: specialised 3-D Esirkepov kernel (bodies: fast3d_advance / union_deposit in pic_slots.cuh).
//  * geometry and per-species constants are computed on the host and live in the constant bank;
//  * particles are cell-sorted, so neighbouring lanes usually deposit to the SAME nodes: the same-cell current values are
//    combined with a segmented warp scan (key = stencil base index) and only the last lane of each run issues the global
//    RED -- measured on B200, global fp32 RED sustains ~0.7 lane-ops/clk/SM (profiles/r01_microbench_atomics.jsonl), which
//    made the un-aggregated kernel L2-atomic-bound (profiles/r01_k1_versions.md);
//  * the few particles that change anchor cell are queued in shared memory and deposited by full warps through the
//    union-stencil body, so the common path stays branch-free.
#ifndef PIC_K1_CTAS
#define PIC_K1_CTAS 4   /* 64 registers; measured 5.31 ms vs 5.56 (3 CTAs) vs 6.22 (5, spills) per launch */
#endif
#ifndef PIC_DEPOSIT_V
#define PIC_DEPOSIT_V 0   /* CIC same-cell deposit: 0 = segmented warp scan (4.88 ms); 1 = run-cooperative reduction through shared memory (9.0 ms: the per-run work split diverges badly when cell-changers cut the runs) */
#endif
#ifndef PIC_K1_PREFETCH
#define PIC_K1_PREFETCH 0   /* software prefetch of the next particle measured neutral-to-worse (5.37 vs 5.31 ms) */
#endif
template <typename T, int SF, int PUSHER, int STEPS>
__global__ void __launch_bounds__(256, (SF == 1 ? PIC_K1_CTAS : 2)) k_fused3d(const __grid_constant__ PicParams p, int species, const __grid_constant__ Geom<T> gm,
                                                    const __grid_constant__ FastConst<T> k, SoAView<T> s, Field6<T> F, Field3W<T> J,
                                                    LeaveBuf leave, int distributed, int32_t* flags) {
    constexpr int NV = SameCell<SF>::NV, NN = SameCell<SF>::NN;
    // STEPS = segmented-scan depth: runs are cut at groups of 1 << STEPS lanes (fewer steps = fewer shuffles, more REDs)
    constexpr int G = 1 << STEPS;
    constexpr int QW = 64;                            // per-warp queue of anchor-changing particles (flushed at >= 32)
    constexpr int NWARP = 8;
    __shared__ T q_old[NWARP][3][QW];
    __shared__ T q_new[NWARP][3][QW];
    constexpr bool COOP = (SF == 1) && (PIC_DEPOSIT_V == 1) && (sizeof(T) == 4);   // f64 would exceed the 48 KB static smem
    constexpr int NVP = NV + 1;                       // odd row stride: conflict-free for both the writes and the run reads
    __shared__ T vbuf[COOP ? NWARP : 1][COOP ? 32 : 1][COOP ? NVP : 1];
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = J.f[c]; sink.L[c] = gm.L[c]; }
    sink.off = 0;
    sink.flags = flags;
    Field6<T> X;
    for (int c = 0; c < 6; ++c) X.f[c] = nullptr;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = lane & (G - 1);
    int qn = 0;                                       // warp-uniform queue fill
    // Each CTA walks a CONTIGUOUS range of the cell-sorted particle stream, 256 particles (8 warps x 32) per iteration:
    // consecutive iterations touch neighbouring cells, so the E/B stencil rows stay in L1 between iterations and across
    // the warps of the CTA.  Warps never synchronise with each other (the deferred queue is warp-private).
    const int64_t n_total = s.count();
    const int64_t per_block = ((n_total + (int64_t)gridDim.x * 256 - 1) / ((int64_t)gridDim.x * 256)) * 256;
    const int64_t b_begin = (int64_t)blockIdx.x * per_block;
    const int64_t w_end = (b_begin + per_block < n_total) ? b_begin + per_block : n_total;
#if PIC_K1_PREFETCH
    // software prefetch: the six particle words of the NEXT iteration are requested before the current particle is
    // processed, so their DRAM latency overlaps ~1000 instructions of work
    T cur[6];
    {
        const int64_t i0 = b_begin + warp * 32 + lane;
#pragma unroll
        for (int c = 0; c < 6; ++c) cur[c] = (i0 < w_end) ? s.c[c][i0] : pic_nan<T>();
    }
#endif
    for (int64_t base = b_begin + warp * 32; base < w_end; base += 256) {
        const int64_t i = base + lane;
        T vals[NV], po[3], xn[3], v[3];
        int key = 0, kind = 0;
#if PIC_K1_PREFETCH
        T nxt[6];
        {
            const int64_t in = i + 256;
#pragma unroll
            for (int c = 0; c < 6; ++c) nxt[c] = (in < w_end) ? s.c[c][in] : pic_nan<T>();
        }
        if (i < w_end) kind = fast3d_advance<T, SF, PUSHER, false>(p, species, k, i, s, F, X, leave, distributed != 0, flags, po, xn, v, key, vals, cur);
#pragma unroll
        for (int c = 0; c < 6; ++c) cur[c] = nxt[c];
#else
        if (i < w_end) kind = fast3d_advance<T, SF, PUSHER, false>(p, species, k, i, s, F, X, leave, distributed != 0, flags, po, xn, v, key, vals);
#endif
        // ---- deferred anchor-changing particles: warp-private queue
        const unsigned defer = __ballot_sync(0xffffffffu, kind == 2);
        if (defer) {
            if (kind == 2) {
                const int slot = qn + __popc(defer & ((1u << lane) - 1u));
#pragma unroll
                for (int a = 0; a < 3; ++a) { q_old[warp][a][slot] = po[a]; q_new[warp][a][slot] = xn[a]; }
            }
            qn += __popc(defer);
            __syncwarp();
        }
        if (kind != 1) {
            key = -1 - lane;
#pragma unroll
            for (int n = 0; n < NV; ++n) vals[n] = (T)0;
        }
#if defined(PIC_ABLATE) && PIC_ABLATE == 1   /* profiling build: no scan, no RED (keeps vals alive through one store) */
        {
            T acc = (T)0;
            for (int n = 0; n < NV; ++n) acc += vals[n];
            if (acc == (T)1.2345e30) sink.J[0][0] = acc;
            continue;
        }
#endif
        if (COOP) {
            // ---- run-cooperative reduction: every lane parks its same-cell values in shared memory; the lanes of one run
            // (consecutive particles of the same cell) then split the NV sums of that run between them: lane `pos` of a run of
            // length `len` sums values pos, pos+len, ... over the run and issues those REDs.  ~NV loads+adds per lane whatever
            // the run length, instead of 3 shuffle steps x NV values.
#pragma unroll
            for (int n = 0; n < NV; ++n) vbuf[warp][lane][n] = vals[n];
            const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
            const bool head = (lane == 0) || (key != key_prev);
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            __syncwarp();
            if (key >= 0) {
                const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
                const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
                const int end = above ? lane + __ffs(above) - 1 : 31;
                const int len = end - start + 1;
                for (int n = lane - start; n < NV; n += len) {
                    T sum = (T)0;
                    for (int l = start; l <= end; ++l) sum += vbuf[warp][l][n];
                    const int c = n / (NV / 3), r = n % (NV / 3);
                    const int f = r / (NN * NN), m1 = (r / NN) % NN, m2 = r % NN;
                    atomicAdd(sink.J[c] + key + SameCell<SF>::offset(c, f, m1, m2, k.sx, k.sy), sum);
                }
            }
            __syncwarp();
        } else {
        // ---- segmented inclusive scan over lanes with equal key (flag = "a segment head lies in (lane-d, lane]")
            const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
            const bool head = (gl == 0) || (key != key_prev);
            int flag = head ? 1 : 0;
#pragma unroll
            for (int d = 1; d < G; d <<= 1) {
                const int fo = __shfl_up_sync(0xffffffffu, flag, d);
                const bool take = (gl >= d) && (flag == 0);
#pragma unroll
                for (int n = 0; n < NV; ++n) {
                    const T o = __shfl_up_sync(0xffffffffu, vals[n], d);
                    if (take) vals[n] += o;      // predicated add (one instruction instead of select + add)
                }
                if (take) flag |= fo;
            }
            const int head_next = __shfl_down_sync(0xffffffffu, head ? 1 : 0, 1);
            const bool tail = (gl == G - 1) || (head_next != 0);
            if (tail && key >= 0) {
                int n = 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    T* Jc = sink.J[c] + key;
#pragma unroll
                    for (int f = 0; f < NN - 1; ++f)
#pragma unroll
                        for (int m1 = 0; m1 < NN; ++m1)
#pragma unroll
                            for (int m2 = 0; m2 < NN; ++m2) atomicAdd(Jc + SameCell<SF>::offset(c, f, m1, m2, k.sx, k.sy), vals[n++]);
                }
            }
        }
        // ---- flush the warp queue with a (nearly) full warp
        if (qn >= 32 || (base + 256 >= w_end && qn > 0)) {
            for (int e = lane; e < qn; e += 32) {
                const T o3[3] = {q_old[warp][0][e], q_old[warp][1][e], q_old[warp][2][e]};
                const T n3[3] = {q_new[warp][0][e], q_new[warp][1][e], q_new[warp][2][e]};
                const T v3[3] = {(T)0, (T)0, (T)0};   // velocities only enter the deposit on inactive axes (none here)
                union_deposit<T, SF>(p, species, gm, k, o3, n3, v3, sink);
            }
            qn = 0;
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------- K1 v9: supercell tile kernel (CIC)
// The blocked sort order keeps the particles of one 4x4x4-cell supercell contiguous; `blk_off` (from the last sort) delimits
// them.  A CTA walks a contiguous range of supercells.  For each one it stages the 8x8x8-node E/B neighbourhood of all six
// components in shared memory with 16-byte cp.async copies (double-buffered: the next supercell's tile is in flight while the
// current one is processed), and the per-particle gather becomes 48 shared-memory loads with constant offsets from one
// 32-bit base per component -- no 64-bit address arithmetic, no L1 misses on the critical path.  Particles drift between
// sorts; the tile carries a one-cell margin, and a particle that left it (kind 3) takes the global-memory body, so the
// result never depends on how well the stream is sorted.  Push, same-cell deposit (segmented warp scan + RED), deferred
// cell-crossers, move, BCs and leaver packets are the v8 code.
#ifndef PIC_K9_CTAS
#define PIC_K9_CTAS 2       /* CTAs per SM (sets the register cap): 2 x 12 warps at 80 registers; 64 registers spill the per-supercell state */
#endif
#ifndef PIC_K9_NW
#define PIC_K9_NW 12        /* warps per CTA */
#endif
#ifndef PIC_K9_DEAL
#define PIC_K9_DEAL 1       /* how a supercell's 32-particle chunks reach the warps: 0 = static round-robin, 1 = shared-memory counter */
#endif
#ifndef PIC_K9_PREFETCH
#define PIC_K9_PREFETCH 0   /* L2 prefetch of the particles NW chunks ahead: measured 5.08 vs 5.00 ms without */
#endif
#ifndef PIC_K9_PPT
/* particles per thread and chunk-loop iteration: 1 = the measured kernel.  2 = lane l advances the ADJACENT particles 2l and 2l+1
   of a 64-particle chunk: dealing, descriptor reads, queue bookkeeping and -- because neighbours in the sorted stream mostly share
   their cell -- the warp reduction are paid once per pair (B's values are added to A's when the cells match, sent out directly
   when they do not).  Needs ~100 registers: build with -DPIC_K9_PPT=2 -DPIC_K9_NW=10.  Compiles; not yet run on a GPU
   (profiles/r01_k1_instruction_budget.md, candidate 1). */
#define PIC_K9_PPT 1
#endif
#ifndef PIC_K9_JT_SYNC
/* J-tile (MODE 1) hand-over between warps: 0 = __threadfence_block() + atomicAdd -- the build every J-tile number of round 1 was
   measured with; its fence.sc makes ptxas turn ALL 126 global REDG of that kernel into returning ATOMG (cuobjdump, see
   tests/test_sass_evidence.py).  1 = one atom.acq_rel.cta on the arrival counter: same ordering guarantee for the tile, the global
   adds stay REDG.  Compiles, not yet run on a GPU -- to be A/B'd before it becomes the default. */
#define PIC_K9_JT_SYNC 0
#endif

// One level of the segmented scan: vals[n] += vals[n] of the lane d below, for the lanes that take.  float: the twelve adds are
// six packed FADD2 (Blackwell f32x2); the shuffles move the two halves of each register pair separately.
#ifndef PIC_SCAN_F32X2
#define PIC_SCAN_F32X2 1
#endif
template <typename T, int NV>
__device__ __forceinline__ void scan_level_add(T* vals, int d, bool take) {
    if constexpr (PIC_SCAN_F32X2 && sizeof(T) == 4 && NV % 2 == 0) {
#pragma unroll
        for (int n = 0; n < NV; n += 2) {
            const float2 o = make_float2(__shfl_up_sync(0xffffffffu, vals[n], d), __shfl_up_sync(0xffffffffu, vals[n + 1], d));
            if (take) {
                const float2 r = __fadd2_rn(make_float2(vals[n], vals[n + 1]), o);
                vals[n] = r.x; vals[n + 1] = r.y;
            }
        }
    } else {
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const T o = __shfl_up_sync(0xffffffffu, vals[n], d);
            if (take) vals[n] += o;      // predicated add (one instruction instead of select + add)
        }
    }
}

// Segmented inclusive scan (depth STEPS) of the same-cell current values over lanes with equal key; the last lane of every run
// issues the REDs.  (flag = "a segment head lies in (lane-d, lane]")
template <typename T, int SF, int STEPS>
__device__ __forceinline__ void same_cell_scan_red(T* vals, int key, int lane, const TileSink<T>& sink, int sx, int sy) {
    constexpr int NV = SameCell<SF>::NV, NN = SameCell<SF>::NN;
    constexpr int G = 1 << STEPS;
    const int gl = lane & (G - 1);
    const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (gl == 0) || (key != key_prev);
    int flag = head ? 1 : 0;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const int fo = __shfl_up_sync(0xffffffffu, flag, d);
        const bool take = (gl >= d) && (flag == 0);
        scan_level_add<T, NV>(vals, d, take);
        if (take) flag |= fo;
    }
    const int head_next = __shfl_down_sync(0xffffffffu, head ? 1 : 0, 1);
    const bool tail = (gl == G - 1) || (head_next != 0);
    if (tail && key >= 0) {
        int n = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T* Jc = sink.J[c] + key;
#pragma unroll
            for (int f = 0; f < NN - 1; ++f)
#pragma unroll
                for (int m1 = 0; m1 < NN; ++m1)
#pragma unroll
                    for (int m2 = 0; m2 < NN; ++m2) atomicAdd(Jc + SameCell<SF>::offset(c, f, m1, m2, sx, sy), vals[n++]);
        }
    }
}

// Alternative to the segmented scan: reduce over ALL lanes of the warp that deposit to the same cell, contiguous or not.
// `__match_any_sync` yields each lane's group; the groups are linked lists over the lanes and are summed by pointer doubling
// (log2 of the longest group many rounds of 12 index shuffles); the lowest lane of every group issues the REDs.  A stale sort
// fragments the same-cell runs (cell changers sit between them), which costs the scan one RED set per fragment but this
// reduction nothing: REDs per warp = distinct cells per warp.
template <typename T, int SF>
__device__ __forceinline__ void same_cell_group_red(T* vals, int key, int lane, const TileSink<T>& sink, int sx, int sy) {
    constexpr int NV = SameCell<SF>::NV, NN = SameCell<SF>::NN;
    const unsigned group = __match_any_sync(0xffffffffu, key);
    const unsigned above = group & (0xfffffffeu << lane);
    int next = above ? __ffs(above) - 1 : 32;
    while (__any_sync(0xffffffffu, next < 32)) {
        const bool has = next < 32;
        const int src = has ? next : lane;
        if constexpr (sizeof(T) == 4 && NV % 2 == 0) {
#pragma unroll
            for (int n = 0; n < NV; n += 2) {
                const float2 o = make_float2(__shfl_sync(0xffffffffu, vals[n], src), __shfl_sync(0xffffffffu, vals[n + 1], src));
                if (has) {
                    const float2 r = __fadd2_rn(make_float2(vals[n], vals[n + 1]), o);
                    vals[n] = r.x; vals[n + 1] = r.y;
                }
            }
        } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const T o = __shfl_sync(0xffffffffu, vals[n], src);
                if (has) vals[n] += o;
            }
        }
        const int nn = __shfl_sync(0xffffffffu, next, src);
        next = has ? nn : 32;
    }
    if (key >= 0 && lane == __ffs(group) - 1) {
        int n = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T* Jc = sink.J[c] + key;
#pragma unroll
            for (int f = 0; f < NN - 1; ++f)
#pragma unroll
                for (int m1 = 0; m1 < NN; ++m1)
#pragma unroll
                    for (int m2 = 0; m2 < NN; ++m2) atomicAdd(Jc + SameCell<SF>::offset(c, f, m1, m2, sx, sy), vals[n++]);
        }
    }
}

// Same reduction, but the run tails add into the supercell's shared-memory J tile (`jt`: [3][8][8][8], origin = the E/B tile's)
// when their stencil is covered by it (srel >= 0), and into global memory otherwise.
template <typename T, int STEPS>
__device__ __forceinline__ void same_cell_scan_red_tile(T* vals, int key, int srel, int lane, const TileSink<T>& sink, int sx, int sy, T* jt) {
    constexpr int SF = 1;
    constexpr int NV = SameCell<SF>::NV, NN = SameCell<SF>::NN;
    constexpr int G = 1 << STEPS;
    const int gl = lane & (G - 1);
    const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (gl == 0) || (key != key_prev);
    int flag = head ? 1 : 0;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const int fo = __shfl_up_sync(0xffffffffu, flag, d);
        const bool take = (gl >= d) && (flag == 0);
        scan_level_add<T, NV>(vals, d, take);
        if (take) flag |= fo;
    }
    const int head_next = __shfl_down_sync(0xffffffffu, head ? 1 : 0, 1);
    const bool tail = (gl == G - 1) || (head_next != 0);
    if (tail && key >= 0) {
        if (srel >= 0) {
            int n = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T* Jc = jt + c * (TILE_N * TILE_N * TILE_N) + srel;
#pragma unroll
                for (int m1 = 0; m1 < NN; ++m1)
#pragma unroll
                    for (int m2 = 0; m2 < NN; ++m2) red_shared_add(Jc + SameCell<SF>::offset(c, 0, m1, m2, TILE_N * TILE_N, TILE_N), vals[n++]);
            }
        } else {
            int n = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T* Jc = sink.J[c] + key;
#pragma unroll
                for (int m1 = 0; m1 < NN; ++m1)
#pragma unroll
                    for (int m2 = 0; m2 < NN; ++m2) atomicAdd(Jc + SameCell<SF>::offset(c, 0, m1, m2, sx, sy), vals[n++]);
            }
        }
    }
}

// Particle staging: the elected thread also bulk-copies the supercell's slice of the six particle arrays (TMA 1-D copies on the
// same barrier as the field tile), so the hot loop reads x,y,z,vx,vy,vz from shared memory and no warp ever waits for DRAM on
// its critical path.  PCAP slots per stage cover a supercell of up to ~PCAP particles (mean 512 at 8 ppc per species); the
// slots beyond it, and arrays that are not 16-byte aligned, are read from global memory as before.
constexpr int K9_PCAP = 576;
constexpr int K9_QW = 48;        // per-warp queue of anchor-changing particles, flushed once 16 are waiting (so <= 47 ever queue up)
// JT = true: the same-cell currents are accumulated in a shared-memory J tile per supercell ([3][8][8][8], two tiles in flight)
// with shared-memory atomics and flushed to global memory by ONE TMA reduce per component when the last warp leaves the
// supercell -- instead of 12 global REDs per run of same-cell particles.  It pays for species whose sort is stale (cell
// changers fragment the runs: 5.3 RED sectors per particle for electrons 8 steps after a sort, 3.5 right after it).
template <typename T, int PUSHER, int STEPS, int NW, bool PER1, int MODE>
__global__ void __launch_bounds__(NW * 32, (sizeof(T) == 8 ? 1 : PIC_K9_CTAS)) k_tile3d(const __grid_constant__ PicParams p, int species, const __grid_constant__ Geom<T> gm,
                                                   const __grid_constant__ FastConst<T> k, const __grid_constant__ SoAView<T> s, const __grid_constant__ Field6<T> F, Field3W<T> J,
                                                   LeaveBuf leave, int distributed, int32_t* flags, const __grid_constant__ TileMaps tm,
                                                   const int32_t* __restrict__ blk_off, int nblk, int nby, int nbz, int stage_particles) {
    constexpr int SF = 1;
    constexpr bool JT = (MODE == 1);      // same-cell reduction: 0 segmented scan + RED, 1 shared-memory J tile, 2 match-any groups + RED
    constexpr int NV = SameCell<SF>::NV;
    constexpr int QW = K9_QW;
    constexpr int NSTAGE = 3;                                // ring: the next supercell is in flight while the current one is
                                                             // processed, and a warp may run one supercell ahead of the slowest
    constexpr int PCAP = K9_PCAP;
    constexpr int TILE_ALL = 6 * TILE_ELEMS;
    constexpr int STAGE_ELEMS = TILE_ALL + 6 * PCAP;         // field tile + particle slice
    constexpr int TILE_BYTES = TILE_ALL * (int)sizeof(T);
    constexpr int AL = 16 / (int)sizeof(T);                  // elements per 16 bytes
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // full[NSTAGE]: tile + particles have landed (TMA complete_tx)
    uint64_t* empty = full + NSTAGE;                         // empty[NSTAGE]: every warp is done reading the stage
    int* requested = reinterpret_cast<int*>(empty + NSTAGE); // last supercell whose stage has been requested
    int* chunk_ctr = requested + 1;                          // [NSTAGE + 1] next undealt chunk of the supercell in each ring slot (+ tail pass)
    int* jdone = chunk_ctr + NSTAGE + 1;                     // [2] warps that have left the supercell whose currents sit in J tile 0 / 1
    uint64_t* jfree = reinterpret_cast<uint64_t*>(smem_raw + 96);   // [2] J tile flushed and zeroed again
    int* desc = reinterpret_cast<int*>(smem_raw + 128);      // [NSTAGE][8] what the requester knows about the supercell in each slot:
                                                             // first slot, end slot, end of the staged slice, tile origin x, y, z
    T* stages = reinterpret_cast<T*>(smem_raw + 256);        // [NSTAGE][6][8][9][8] + [6][PCAP]
    T* q_old = stages + NSTAGE * STAGE_ELEMS;                // [NW][3][QW]
    T* q_new = q_old + NW * 3 * QW;
    constexpr int JT_ELEMS = 3 * TILE_N * TILE_N * TILE_N;
    T* jtiles = q_new + NW * 3 * QW;                         // [2][3][8][8][8] (JT only; 128-byte aligned by construction)
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = J.f[c]; sink.L[c] = gm.L[c]; }
    sink.off = 0;
    sink.flags = flags;
    Field6<T> X;
    for (int c = 0; c < 6; ++c) X.f[c] = nullptr;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int per_cta = (nblk + (int)gridDim.x - 1) / (int)gridDim.x;
    const int b0 = (int)blockIdx.x * per_cta;
    int b1 = (b0 + per_cta < nblk) ? b0 + per_cta : nblk;
    if (b1 < b0) b1 = b0;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, NW); }
        mbar_init(jfree, 1); mbar_init(jfree + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        *requested = b0 - 1;
        for (int i = 0; i <= NSTAGE; ++i) chunk_ctr[i] = 0;
        jdone[0] = jdone[1] = 0;
    }
    if (JT)
        for (int i = tid; i < 2 * JT_ELEMS; i += NW * 32) jtiles[i] = (T)0;
    __syncthreads();
    T* qo = q_old + warp * 3 * QW;
    T* qn_ = q_new + warp * 3 * QW;
    int qn = 0;                                              // warp-uniform queue fill
    const int n_live = (int)s.count();
    const int cap_al = (int)(s.cap < 0x7fffffff ? s.cap : 0x7fffffff) & ~(AL - 1);
    // staged slice of supercell [beg, end): slots [beg_al, beg_al + n) with beg_al = beg rounded down to 16 bytes
    auto slice_len = [&](int beg, int end) {
        if (!stage_particles || end <= beg) return 0;
        const int beg_al = beg & ~(AL - 1);
        int n = (end - beg_al + AL - 1) & ~(AL - 1);
        if (n > PCAP) n = PCAP;
        if (beg_al + n > cap_al) n = cap_al - beg_al;
        return n > 0 ? n : 0;
    };
    // The ring is fed by single threads: six TMA box copies (g == 2: the tile's first node is the supercell's first cell; the
    // spare y rows beyond the 8th are padding -- zero-filled where they leave the array, never read) and six bulk copies of
    // the particle slice per supercell.  Requests are made in order; `requested` is the last supercell asked for.  Any warp's
    // lane 0 may make the next one (compare-and-swap) as soon as the ring slot it needs has been released by every warp --
    // it never blocks for that: if the slot is still in use the attempt is dropped and repeated by the next warp that enters
    // a supercell or waits for a stage, so a straggler delays nobody but the warps that actually need its slot.
    auto try_request = [&](int upto) {
        const int r = *reinterpret_cast<volatile int*>(requested) + 1;
        if (r > upto || r >= b1) return;
        const int jr = r - b0;
        const int sr = jr % NSTAGE;
        if (jr >= NSTAGE && !mbar_test(empty + sr, (jr / NSTAGE - 1) & 1)) return;     // previous occupant: supercell r - NSTAGE
        if (atomicCAS(requested, r - 1, r) != r - 1) return;
        const int bz = r % nbz, by = (r / nbz) % nby, bx = r / (nbz * nby);
        const int beg = blk_off[r];
        int end = blk_off[r + 1];
        if (end > n_live) end = n_live;
        const int n = slice_len(beg, end);
        T* st = stages + sr * STAGE_ELEMS;
        chunk_ctr[sr] = 0;           // (published to the other warps, like the descriptor, by the release of the expect_tx arrive)
        int* d = desc + sr * 8;
        d[0] = beg; d[1] = end; d[2] = (beg & ~(AL - 1)) + n;
        d[4] = bx * TILE_B; d[5] = by * TILE_B; d[6] = bz * TILE_B;
        mbar_arrive_expect_tx(full + sr, TILE_BYTES + 6 * n * (int)sizeof(T));
#pragma unroll
        for (int c = 0; c < 6; ++c)
            tma_load_box(st + c * TILE_ELEMS, &tm.m[c], full + sr, bz * TILE_B, by * TILE_B, bx * TILE_B);
        if (n > 0) {
            const int beg_al = beg & ~(AL - 1);
#pragma unroll
            for (int c = 0; c < 6; ++c) tma_load_bytes(st + TILE_ALL + c * PCAP, s.c[c] + beg_al, n * (int)sizeof(T), full + sr);
        }
    };
    int slot = 0, par = 0;       // ring slot of the supercell being processed and the parity of its use count
    int jsel = 0, juse = 0;      // J tile of the supercell being processed (alternates) and how often that tile has been used
#if PIC_K9_DEAL == 0
    int rot = 0;                 // static dealing: chunks go to the warps round-robin, continuing across supercells, so every warp
                                 // gets the same number of chunks (+-1) whatever the supercell populations are
#endif
    // iteration b == b1 is the tail pass: this CTA's share of the slots appended since the last sort (particles received from
    // neighbour ranks, ~1e-4 of the stream per step).  They are not binned; an impossible tile origin sends them through
    // the global-memory gather of the same body.
    for (int b = b0; b <= b1; ++b) {
        const bool tail_pass = (b == b1);
        const T* st = stages + slot * STAGE_ELEMS;
        TileSrc<T> ts;
        ts.t = st;
        int p_beg, p_end, i_staged_end = 0;
        if (!tail_pass) {
            if (lane == 0) try_request(b + 1);
            while (!mbar_try_wait(full + slot, par, 1000)) {
                if (lane == 0) try_request(b + 1);       // the stage we wait for may not even have been requested yet
            }
            {   // everything else about this supercell was worked out once, by the thread that requested its stage
                const int4 d0 = *reinterpret_cast<const int4*>(desc + slot * 8);
                const int4 d1 = *reinterpret_cast<const int4*>(desc + slot * 8 + 4);
                p_beg = d0.x; p_end = d0.y; i_staged_end = d0.z;
                ts.o[0] = d1.x; ts.o[1] = d1.y; ts.o[2] = d1.z;
            }
            if (JT && juse > 0) {                        // this J tile last held supercell b - 2: wait for its flush
                while (!mbar_try_wait(jfree + jsel, (juse - 1) & 1, 1000)) {}
            }
            __syncwarp();
        } else {
            const int tail0 = blk_off[nblk];
            const int ntail = n_live > tail0 ? n_live - tail0 : 0;
            const int per = ((ntail + (int)gridDim.x * 32 - 1) / ((int)gridDim.x * 32)) * 32;
            ts.o[0] = ts.o[1] = ts.o[2] = -(1 << 24);
            p_beg = tail0 + (int)blockIdx.x * per;
            p_end = (p_beg + per < n_live) ? p_beg + per : n_live;
        }
#if PIC_K9_PPT == 2
        const int p_even = p_beg & ~1;                           // pairs start on even slots (the staged slice does too)
        const int nchunk = p_end > p_beg ? (p_end - p_even + 63) >> 6 : 0;
#else
        const int nchunk = p_end > p_beg ? (p_end - p_beg + 31) >> 5 : 0;
#endif
        const T* pst = st + TILE_ALL - (p_beg & ~(AL - 1));      // staged particle i of array c sits at pst[c * PCAP + i]
#if PIC_K9_DEAL == 1
        // a warp that finishes early takes the next undealt chunk, or moves on to the next supercell: the warps stay within
        // one chunk of each other, so ring slots are released promptly
        int* ctr = chunk_ctr + (tail_pass ? NSTAGE : slot);
        for (;;) {
            int ch = 0;
            if (lane == 0) ch = atomicAdd(ctr, 1);
            ch = __shfl_sync(0xffffffffu, ch, 0);
            if (ch >= nchunk) break;
#else
        for (int ch = (warp - rot + NW * 64) % NW; ch < nchunk; ch += NW) {
#endif
#if PIC_K9_PPT == 2
            // ---- two adjacent particles per lane: A = slot i0 (even), B = i0 + 1
            auto flush_queue = [&]() {
                if (qn >= QW - 32) {
                    for (int e = lane; e < qn; e += 32) {
                        const T o3[3] = {qo[e], qo[QW + e], qo[2 * QW + e]};
                        const T n3[3] = {qn_[e], qn_[QW + e], qn_[2 * QW + e]};
                        const T v3[3] = {(T)0, (T)0, (T)0};
                        union_deposit<T, SF>(p, species, gm, k, o3, n3, v3, sink);
                    }
                    qn = 0;
                    __syncwarp();
                }
            };
            auto advance_one = [&](int i, bool valid, T* vals, int& key, int dead_key) {
                T po[3], xn[3], v[3], cur[6];
                int kind = 0;
                key = 0;
                if (valid) {
                    if (i < i_staged_end) {
#pragma unroll
                        for (int c = 0; c < 6; ++c) cur[c] = pst[c * PCAP + i];
                    } else {
#pragma unroll
                        for (int c = 0; c < 6; ++c) cur[c] = s.c[c][i];
                    }
                    kind = fast3d_advance<T, SF, PUSHER, false, true, PER1>(p, species, k, i, s, F, X, leave, distributed != 0, flags, po, xn, v, key, vals, cur, &ts, nullptr);
                }
                const unsigned defer = __ballot_sync(0xffffffffu, kind == 2);
                if (defer) {
                    if (kind == 2) {
                        const int slot_q = qn + __popc(defer & ((1u << lane) - 1u));
#pragma unroll
                        for (int a = 0; a < 3; ++a) { qo[a * QW + slot_q] = po[a]; qn_[a * QW + slot_q] = xn[a]; }
                    }
                    qn += __popc(defer);
                    __syncwarp();
                }
                if (kind != 1) {
                    key = dead_key;
#pragma unroll
                    for (int n = 0; n < NV; ++n) vals[n] = (T)0;
                }
                return kind;
            };
            const int i0 = p_even + ch * 64 + 2 * lane;
            T vals[NV], valsB[NV];
            int key, keyB;
            const int kindA = advance_one(i0, i0 >= p_beg && i0 < p_end, vals, key, -1 - lane);
            flush_queue();                                   // <= 47 queued before B may add another 32
            const int kindB = advance_one(i0 + 1, i0 + 1 < p_end, valsB, keyB, -33 - lane);
            {
                const bool b_live = (kindB == 1);
                const bool join = b_live && (kindA != 1 || key == keyB);     // A's slot is empty, or A and B share the cell
                if (b_live && !join) {                       // neighbours in different cells (a run boundary): B goes out by itself
                    int n = 0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        T* Jc = sink.J[c] + keyB;
#pragma unroll
                        for (int m1 = 0; m1 < SameCell<SF>::NN; ++m1)
#pragma unroll
                            for (int m2 = 0; m2 < SameCell<SF>::NN; ++m2) atomicAdd(Jc + SameCell<SF>::offset(c, 0, m1, m2, k.sx, k.sy), valsB[n++]);
                    }
                }
#pragma unroll
                for (int n = 0; n < NV; ++n) vals[n] += join ? valsB[n] : (T)0;     // (A's values are zero when its slot is empty)
                if (join) key = keyB;
            }
            if (JT) same_cell_scan_red_tile<T, STEPS>(vals, key, -1, lane, sink, k.sx, k.sy, jtiles + jsel * JT_ELEMS);
            else if (MODE == 2) same_cell_group_red<T, SF>(vals, key, lane, sink, k.sx, k.sy);
            else same_cell_scan_red<T, SF, STEPS>(vals, key, lane, sink, k.sx, k.sy);
            flush_queue();
#else
            const int i = p_beg + ch * 32 + lane;
            T vals[NV], po[3], xn[3], v[3], cur[6];
            int key = 0, kind = 0, srel = -1;
            if (i < p_end) {
                if (i < i_staged_end) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) cur[c] = pst[c * PCAP + i];
                } else {
#pragma unroll
                    for (int c = 0; c < 6; ++c) cur[c] = s.c[c][i];
                }
                kind = fast3d_advance<T, SF, PUSHER, false, true, PER1>(p, species, k, i, s, F, X, leave, distributed != 0, flags, po, xn, v, key, vals, cur, &ts, JT ? &srel : nullptr);
            }
            // ---- deferred anchor-changing particles: warp-private queue
            const unsigned defer = __ballot_sync(0xffffffffu, kind == 2);
            if (defer) {
                if (kind == 2) {
                    const int slot_q = qn + __popc(defer & ((1u << lane) - 1u));
#pragma unroll
                    for (int a = 0; a < 3; ++a) { qo[a * QW + slot_q] = po[a]; qn_[a * QW + slot_q] = xn[a]; }
                }
                qn += __popc(defer);
                __syncwarp();
            }
            if (kind != 1) {
                key = -1 - lane;
#pragma unroll
                for (int n = 0; n < NV; ++n) vals[n] = (T)0;
            }
            if (JT) same_cell_scan_red_tile<T, STEPS>(vals, key, tail_pass ? -1 : srel, lane, sink, k.sx, k.sy, jtiles + jsel * JT_ELEMS);
            else if (MODE == 2) same_cell_group_red<T, SF>(vals, key, lane, sink, k.sx, k.sy);
            else same_cell_scan_red<T, SF, STEPS>(vals, key, lane, sink, k.sx, k.sy);
            // ---- flush the warp queue once half a warp of them is waiting
            if (qn >= QW - 32) {
                for (int e = lane; e < qn; e += 32) {
                    const T o3[3] = {qo[e], qo[QW + e], qo[2 * QW + e]};
                    const T n3[3] = {qn_[e], qn_[QW + e], qn_[2 * QW + e]};
                    const T v3[3] = {(T)0, (T)0, (T)0};   // velocities only enter the deposit on inactive axes (none here)
                    union_deposit<T, SF>(p, species, gm, k, o3, n3, v3, sink);
                }
                qn = 0;
                __syncwarp();
            }
#endif
        }
#if PIC_K9_DEAL == 0
        rot = (rot + nchunk) % NW;
#endif
        if (!tail_pass) {
            __syncwarp();
            if (JT) {
                // the last warp to leave the supercell adds its J tile to global memory with one TMA reduce per component
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // our shared-memory adds -> visible to the TMA
                int last = 0;
                if (lane == 0) {
#if PIC_K9_JT_SYNC == 0
                    __threadfence_block();
                    last = (atomicAdd(jdone + jsel, 1) == NW - 1) ? 1 : 0;
#else
                    unsigned old_;
                    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;\n"
                                 : "=r"(old_) : "r"(smem_u32(jdone + jsel)) : "memory");
                    last = (old_ == (unsigned)(NW - 1)) ? 1 : 0;
#endif
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    T* jt = jtiles + jsel * JT_ELEMS;
#if PIC_K9_JT_SYNC == 0
                    __threadfence_block();
#endif
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            tma_reduce_add_box(&tm.j[c], jt + c * (TILE_N * TILE_N * TILE_N), ts.o[2], ts.o[1], ts.o[0]);
                        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // the tile has been read
                    }
                    __syncwarp();
                    for (int i = lane; i < JT_ELEMS; i += 32) jt[i] = (T)0;
                    __syncwarp();
                    if (lane == 0) {
                        jdone[jsel] = 0;
#if PIC_K9_JT_SYNC == 0
                        __threadfence_block();
#endif
                        mbar_arrive(jfree + jsel);               // release: orders the zeroing above before the tile's reuse
                    }
                }
                if (jsel == 1) ++juse;
                jsel ^= 1;
            }
            if (lane == 0) mbar_arrive(empty + slot);        // this warp no longer reads the stage in `slot`
            if (++slot == NSTAGE) { slot = 0; par ^= 1; }
        }
    }
    if (qn > 0) {
        for (int e = lane; e < qn; e += 32) {
            const T o3[3] = {qo[e], qo[QW + e], qo[2 * QW + e]};
            const T n3[3] = {qn_[e], qn_[QW + e], qn_[2 * QW + e]};
            const T v3[3] = {(T)0, (T)0, (T)0};
            union_deposit<T, SF>(p, species, gm, k, o3, n3, v3, sink);
        }
    }
}

// Append every received packet (PicLeave layout: header row with the int32 count, then rows x,y,z,vx,vy,vz,species) at the
// SoA tail; the device counter n_dev advances atomically, so no host round trip is needed between exchange and append.
template <typename T>
__global__ void __launch_bounds__(256) k_append_packets(SoAView<T> s, LeaveBuf recv, int32_t* flags) {
    for (int d = 0; d < 27; ++d) {
        const int cap = recv.cap[d];
        if (cap == 0) continue;
        const T* base = (const T*)recv.buf + (int64_t)recv.row_off[d] * 7;
        int n_in = *(const int32_t*)base;
        if (n_in > cap) n_in = cap;     // the sender flagged the overflow
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_in; j += gridDim.x * blockDim.x) {
            const int64_t dst = atomicAdd(s.n_dev, 1);
            if (dst >= s.cap) { atomicOr(flags, 2); continue; }
            const T* pk = base + (int64_t)(1 + j) * 7;
            for (int c = 0; c < 6; ++c) s.c[c][dst] = pk[c];
            if (s.id) s.id[dst] = -1;
        }
    }
}

__global__ void k_packets_reset(LeaveBuf l, int real_bytes) {
    const int d = threadIdx.x;
    if (d < 27 && l.cap[d] > 0) {
        char* h = (char*)l.buf + (int64_t)l.row_off[d] * 7 * real_bytes;
        for (int b = 0; b < 7 * real_bytes; ++b) h[b] = 0;
    }
}

__global__ void k_set_count(int32_t* n_dev, const int32_t* src) { *n_dev = *src; }

// ---------------------------------------------------------------- launchers
template <typename T>
static int launch_import(const PicParams* p, int species, const void* x, const void* u, const uint8_t* active, int64_t cap_ref,
                         const PicSoA* soa, int32_t* count, cudaStream_t st) {
    if (cap_ref == 0) return 0;
    k_import<T><<<grid_for(cap_ref, 256), 256, 0, st>>>(species, p->n_species, (const T*)x, (const T*)u, active, cap_ref, view_of<T>(soa), count);
    PIC_LAUNCH_RET();
}
template <typename T>
static int launch_export(const PicParams* p, int species, const PicSoA* soa, void* x, void* u, uint8_t* active, int64_t cap_ref,
                         int32_t* count, cudaStream_t st) {
    if (soa->n == 0 && !soa->n_dev) return 0;
    k_export<T><<<grid_for(soa->n, 256), 256, 0, st>>>(species, view_of<T>(soa), (T*)x, (T*)u, active, cap_ref, count);
    PIC_LAUNCH_RET();
}
template <typename T>
static int launch_hist(const PicParams* p, const PicSoA* src, int32_t* count, cudaStream_t st) {
    if (src->n == 0 && !src->n_dev) return 0;
    k_sort_hist<T><<<grid_for(src->n, 256), 256, 0, st>>>(*p, view_of<T>(src), count);
    PIC_LAUNCH_RET();
}
template <typename T>
static int launch_scatter(const PicParams* p, const PicSoA* src, const PicSoA* dst, const int32_t* offset, int32_t* cursor, cudaStream_t st) {
    if (src->n == 0 && !src->n_dev) return 0;
    k_sort_scatter<T><<<grid_for(src->n, 256), 256, 0, st>>>(*p, view_of<T>(src), view_of<T>(dst), offset, cursor);
    if (src->n_dev)   // live particles precede the trash bin: the new slot count is offset[ncells]
        k_set_count<<<1, 1, 0, st>>>(src->n_dev, offset + (int64_t)p->tile[0] * p->tile[1] * p->tile[2]);
    PIC_LAUNCH_RET();
}

template <typename T, int SF>
static int launch_fused(const PicParams* p, int species, int deposition, const PicSoA* soa, const void* const E[3],
                        const void* const B[3], const void* const extE[3], const void* const extB[3], void* const J[3],
                        const PicLeave* leave, int32_t* flags, cudaStream_t st) {
    if (soa->n == 0 && !soa->n_dev) return 0;
    Field6<T> F, X;
    Field3W<T> Jw;
    const int has_ext = (extE && extB) ? 1 : 0;
    for (int c = 0; c < 3; ++c) {
        F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c];
        X.f[c] = has_ext ? (const T*)extE[c] : nullptr; X.f[3 + c] = has_ext ? (const T*)extB[c] : nullptr;
        Jw.f[c] = (T*)J[c];
    }
    const LeaveBuf lb = leave_of(leave);
    const bool all3d = (p->gmesh[0] * p->tile[0] > 1) && (p->gmesh[1] * p->tile[1] > 1) && (p->gmesh[2] * p->tile[2] > 1) && p->g >= 2;
    const int grid = grid_for(soa->n, 256, 8);
    const SoAView<T> sv = view_of<T>(soa);
    if (all3d && deposition == 0 && !has_ext && p->pusher != PIC_PUSHER_HC && !getenv("PIC_K1_GENERIC")) {
        Geom<T> gm;
        make_geom<T>(*p, 0, 0, 0, gm);
        FastConst<T> k;
        make_fast_const<T>(*p, species, gm, k);
        int distributed = 0;
        for (int c = 0; c < 3; ++c) distributed |= (p->gmesh[c] != p->mesh[c]);
        static int steps = 0;
        if (!steps) {
            const char* e = getenv("PIC_K1_SCAN_STEPS");
            steps = e ? atoi(e) : 3;   // measured best on B200 (profiles/r01_k1_versions.md)
            if (steps < 3 || steps > 5) steps = 3;
        }
        const int gridk = grid_for(soa->n, 256, SF == 1 ? 3 * PIC_K1_CTAS : 8);
#define PIC_LAUNCH_K1(PUSH, ST) k_fused3d<T, SF, PUSH, ST><<<gridk, 256, 0, st>>>(*p, species, gm, k, sv, F, Jw, lb, distributed, flags)
        if (SF == 2 || steps == 3) {
            if (p->pusher == PIC_PUSHER_BORIS) PIC_LAUNCH_K1(PIC_PUSHER_BORIS, 3); else PIC_LAUNCH_K1(PIC_PUSHER_BORIS_REL, 3);
        } else if (steps == 4) {
            if (p->pusher == PIC_PUSHER_BORIS) PIC_LAUNCH_K1(PIC_PUSHER_BORIS, 4); else PIC_LAUNCH_K1(PIC_PUSHER_BORIS_REL, 4);
        } else {
            if (p->pusher == PIC_PUSHER_BORIS) PIC_LAUNCH_K1(PIC_PUSHER_BORIS, 5); else PIC_LAUNCH_K1(PIC_PUSHER_BORIS_REL, 5);
        }
#undef PIC_LAUNCH_K1
        PIC_LAUNCH_RET();
    }
    if (deposition == 0) {
        if (all3d) k_fused<T, SF, 0, true><<<grid, 256, 0, st>>>(*p, species, sv, F, X, has_ext, Jw, lb, flags);
        else k_fused<T, SF, 0, false><<<grid, 256, 0, st>>>(*p, species, sv, F, X, has_ext, Jw, lb, flags);
    } else {
        if (all3d) k_fused<T, SF, 1, true><<<grid, 256, 0, st>>>(*p, species, sv, F, X, has_ext, Jw, lb, flags);
        else k_fused<T, SF, 1, false><<<grid, 256, 0, st>>>(*p, species, sv, F, X, has_ext, Jw, lb, flags);
    }
    PIC_LAUNCH_RET();
}

// K1 v9 launcher: returns PIC_EUNSUPPORTED when the configuration is outside what the tile kernel was built for (the caller then
// uses pic_fused_push_deposit).
template <typename T>
static int launch_tile3d(const PicParams* p, int species, const PicSoA* soa, const int32_t* blk_off, int nblk, int options, const void* const E[3],
                         const void* const B[3], void* const J[3], const PicLeave* leave, int32_t* flags, cudaStream_t st) {
    static_assert(PIC_SORT_BLOCK == TILE_B, "the tile kernel walks the supercells of the blocked sort order");
    if (p->shape_factor != 1 || p->g != 2 || (p->pusher != PIC_PUSHER_BORIS && p->pusher != PIC_PUSHER_BORIS_REL)) return PIC_EUNSUPPORTED;
    for (int a = 0; a < 3; ++a)
        if (p->tile[a] % TILE_B != 0 || p->gmesh[a] * p->tile[a] <= 1) return PIC_EUNSUPPORTED;
    const int nbx = p->tile[0] / TILE_B, nby = p->tile[1] / TILE_B, nbz = p->tile[2] / TILE_B;
    if (nblk != nbx * nby * nbz) return PIC_EINVAL;
    for (int c = 0; c < 3; ++c)
        if (((uintptr_t)E[c] | (uintptr_t)B[c]) & 15) return PIC_EUNSUPPORTED;     // TMA global addresses are 16-byte aligned
    if (soa->n == 0 && !soa->n_dev) return 0;
    Field6<T> F;
    Field3W<T> Jw;
    for (int c = 0; c < 3; ++c) { F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c]; Jw.f[c] = (T*)J[c]; }
    Geom<T> gm;
    make_geom<T>(*p, 0, 0, 0, gm);
    FastConst<T> k;
    make_fast_const<T>(*p, species, gm, k);
    int distributed = 0;
    for (int c = 0; c < 3; ++c) distributed |= (p->gmesh[c] != p->mesh[c]);
    constexpr int NW = PIC_K9_NW, QW = K9_QW;
    const bool jt = (options & 1) && sizeof(T) == 4;          // shared-memory J tiles: f32 only (TMA reduce type)
    const bool grp = !jt && (options & 2);                    // match-any group reduction instead of the segmented scan
    const size_t smem = 256 + (size_t)(3 * (6 * TILE_ELEMS + 6 * K9_PCAP) + 2 * NW * 3 * QW + (jt ? 2 * 3 * TILE_N * TILE_N * TILE_N : 0)) * sizeof(T);
    if (smem > 227 * 1024) return PIC_EUNSUPPORTED;
    int grid = num_sms() * (sizeof(T) == 8 ? 1 : PIC_K9_CTAS);
    int stage_particles = 1;     // TMA bulk copies need 16-byte aligned sources
    for (int c = 0; c < 6; ++c)
        if ((uintptr_t)soa->comp[c] & 15) stage_particles = 0;
    if (grid > nblk) grid = nblk;
    const SoAView<T> sv = view_of<T>(soa);
    const LeaveBuf lb = leave_of(leave);
    // TMA descriptors of the six ghosted field tiles: dims (z, y, x) = (Lz, Ly, Lx), box 8 x TILE_NY x 8
    TileMaps tm;
    if (!make_tile_maps<T>(gm.L, F.f, Jw.f, tm)) return PIC_EUNSUPPORTED;
    bool per1 = !distributed;
    for (int a = 0; a < 3; ++a) per1 = per1 && (p->particle_bc[a] == PIC_BC_PERIODIC);
#define PIC_LAUNCH_K9(PUSH, PER, JTV)                                                                                    \
    do {                                                                                                                 \
        static size_t attr_smem = 0;                                                                                     \
        if (attr_smem < smem) {                                                                                          \
            cudaError_t e = cudaFuncSetAttribute(k_tile3d<T, PUSH, 3, NW, PER, JTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                                         \
            attr_smem = smem;                                                                                            \
        }                                                                                                                \
        k_tile3d<T, PUSH, 3, NW, PER, JTV><<<grid, NW * 32, smem, st>>>(*p, species, gm, k, sv, F, Jw, lb, distributed, flags, tm, blk_off, nblk, nby, nbz, stage_particles); \
    } while (0)
#define PIC_LAUNCH_K9_J(PUSH, PER) do { if (jt) PIC_LAUNCH_K9(PUSH, PER, 1); else if (grp) PIC_LAUNCH_K9(PUSH, PER, 2); else PIC_LAUNCH_K9(PUSH, PER, 0); } while (0)
    if (p->pusher == PIC_PUSHER_BORIS) { if (per1) PIC_LAUNCH_K9_J(PIC_PUSHER_BORIS, true); else PIC_LAUNCH_K9_J(PIC_PUSHER_BORIS, false); }
    else { if (per1) PIC_LAUNCH_K9_J(PIC_PUSHER_BORIS_REL, true); else PIC_LAUNCH_K9_J(PIC_PUSHER_BORIS_REL, false); }
#undef PIC_LAUNCH_K9_J
#undef PIC_LAUNCH_K9
    PIC_LAUNCH_RET();
}

template <typename T>
static int launch_append_packets(const PicParams* p, const PicSoA* soa, const PicLeave* recv, int32_t* flags, cudaStream_t st) {
    int maxcap = 0;
    for (int d = 0; d < 27; ++d) maxcap = recv->cap[d] > maxcap ? recv->cap[d] : maxcap;
    if (maxcap == 0) return 0;
    k_append_packets<T><<<grid_for(maxcap, 256, 2), 256, 0, st>>>(view_of<T>(soa), leave_of(recv), flags);
    PIC_LAUNCH_RET();
}

}  // namespace pic

using namespace pic;

extern "C" {

int pic_soa_import(const PicParams* p, int species, const void* x, const void* u, const uint8_t* active, int64_t cap_ref,
                   const PicSoA* soa, int32_t* d_count, void* stream) {
    PIC_CHECK_ARG(p && x && u && active && soa && d_count && species >= 0 && species < p->n_species);
    PIC_CHECK_ARG(p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    PIC_DISPATCH_T(p, launch_import, p, species, x, u, active, cap_ref, soa, d_count, (cudaStream_t)stream);
}

int pic_soa_export(const PicParams* p, int species, const PicSoA* soa, void* x, void* u, uint8_t* active, int64_t cap_ref,
                   int32_t* d_count, void* stream) {
    PIC_CHECK_ARG(p && x && u && active && soa && d_count && species >= 0 && species < p->n_species);
    PIC_DISPATCH_T(p, launch_export, p, species, soa, x, u, active, cap_ref, d_count, (cudaStream_t)stream);
}

int pic_sort_histogram(const PicParams* p, const PicSoA* src, int32_t* cell_count, void* stream) {
    PIC_CHECK_ARG(p && src && cell_count);
    PIC_DISPATCH_T(p, launch_hist, p, src, cell_count, (cudaStream_t)stream);
}

int pic_sort_scan(int64_t n, const int32_t* in, int32_t* out, int32_t* block_scratch, void* stream) {
    PIC_CHECK_ARG(in && out && block_scratch && n >= 0);
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_local<<<(unsigned)nblocks, 256, 0, st>>>(n, in, out, block_scratch);
    if (nblocks > 1) {
        k_scan_sums<<<1, 256, 0, st>>>(nblocks, block_scratch);
        k_scan_add<<<(unsigned)nblocks, 256, 0, st>>>(n, out, block_scratch);
    }
    PIC_LAUNCH_RET();
}

int pic_sort_scatter(const PicParams* p, const PicSoA* src, const PicSoA* dst, const int32_t* cell_offset, int32_t* cell_cursor,
                     void* stream) {
    PIC_CHECK_ARG(p && src && dst && cell_offset && cell_cursor);
    PIC_DISPATCH_T(p, launch_scatter, p, src, dst, cell_offset, cell_cursor, (cudaStream_t)stream);
}

int pic_fused_push_deposit(const PicParams* p, int species, int deposition, const PicSoA* soa, const void* const E[3],
                           const void* const B[3], const void* const extE[3], const void* const extB[3], void* const J[3],
                           const PicLeave* leave, int32_t* flags, void* stream) {
    PIC_CHECK_ARG(p && soa && E && B && J && flags && species >= 0 && species < p->n_species && (deposition == 0 || deposition == 1));
    PIC_CHECK_ARG(p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    bool distributed = false;
    for (int c = 0; c < 3; ++c) distributed |= (p->gmesh[c] != p->mesh[c]);
    if (distributed) {
        PIC_CHECK_ARG(leave && leave->buf);
        if (deposition == 1) return PIC_EUNSUPPORTED;  // centred deposit across ranks needs a mid-step migration
    }
    PIC_DISPATCH_T_SF(p, launch_fused, p, species, deposition, soa, E, B, extE, extB, J, leave, flags, (cudaStream_t)stream);
}

int pic_fused_tile3d(const PicParams* p, int species, const PicSoA* soa, const int32_t* blk_off, int nblk, int options,
                     const void* const E[3], const void* const B[3], void* const J[3], const PicLeave* leave, int32_t* flags, void* stream) {
    PIC_CHECK_ARG(p && soa && blk_off && E && B && J && flags && species >= 0 && species < p->n_species && nblk > 0);
    PIC_CHECK_ARG(p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    bool distributed = false;
    for (int c = 0; c < 3; ++c) distributed |= (p->gmesh[c] != p->mesh[c]);
    if (distributed) PIC_CHECK_ARG(leave && leave->buf);
    PIC_DISPATCH_T(p, launch_tile3d, p, species, soa, blk_off, nblk, options, E, B, J, leave, flags, (cudaStream_t)stream);
}

int pic_packets_reset(const PicParams* p, const PicLeave* leave, void* stream) {
    PIC_CHECK_ARG(p && leave && leave->buf);
    k_packets_reset<<<1, 32, 0, (cudaStream_t)stream>>>(leave_of(leave), p->dtype == PIC_F32 ? 4 : 8);
    PIC_LAUNCH_RET();
}

int pic_soa_append_packets(const PicParams* p, const PicSoA* soa, const PicLeave* recv, int32_t* flags, void* stream) {
    PIC_CHECK_ARG(p && soa && soa->n_dev && recv && recv->buf && flags);
    PIC_DISPATCH_T(p, launch_append_packets, p, soa, recv, flags, (cudaStream_t)stream);
}

}  // extern "C"
