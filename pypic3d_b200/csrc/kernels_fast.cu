// kernels_fast.cu -- the resident throughput path: one tile per GPU, compact cell-sorted SoA particles.
//   K1  k_fused        gather E,B -> push -> deposit -> move -> particle BC (+ leaver extraction), one pass over
//                      particle memory: reads x,y,z,vx,vy,vz once, writes them once (48 B f32 / 96 B f64 per particle)
//   K2  k_sort_*       counting sort by local cell (histogram, exclusive scan, scatter), run every few steps
//   import / export    TiledParticles (reference AoS + active mask) <-> compact SoA
#include <stdlib.h>

#include "pic_common.cuh"

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
template <typename T>
__global__ void __launch_bounds__(256) k_sort_hist(const __grid_constant__ PicParams p, SoAView<T> s, int32_t* __restrict__ count) {
    const int64_t n = s.count();
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&count[local_cell<T>(p, s.c[0][j], s.c[1][j], s.c[2][j])], 1);
}

template <typename T>
__global__ void __launch_bounds__(256) k_sort_scatter(const __grid_constant__ PicParams p, SoAView<T> s, SoAView<T> d,
                                                      const int32_t* __restrict__ offset, int32_t* __restrict__ cursor) {
    const int64_t n = s.count();
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const T px = s.c[0][j], py = s.c[1][j], pz = s.c[2][j];
        const int cell = local_cell<T>(p, px, py, pz);
        const int64_t dst = (int64_t)offset[cell] + atomicAdd(&cursor[cell], 1);
        if (dst >= d.cap) continue;
        d.c[0][dst] = px; d.c[1][dst] = py; d.c[2][dst] = pz;
        d.c[3][dst] = s.c[3][j]; d.c[4][dst] = s.c[4][j]; d.c[5][dst] = s.c[5][j];
        if (s.id && d.id) d.id[dst] = s.id[j];
    }
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
