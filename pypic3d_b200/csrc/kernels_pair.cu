// kernels_pair.cu -- K1 v10: the supercell tile kernel with the pair body (pic_pair.cuh), and the blocked, padded cell sort that
// feeds it.
//
//   k_pair3d      gather E,B -> push -> Esirkepov deposit -> move -> particle BC for one species, one pass over particle memory.
//                 Same walk as K1 v9 (kernels_fast.cu k_tile3d): a CTA takes a contiguous range of 4x4x4-cell supercells of the
//                 blocked sort order; per supercell the 8x9x8-node E/B box of all six components (six cp.async.bulk.tensor.3d) and
//                 the supercell's slice of x,y,z,vx,vy,vz (six cp.async.bulk) land in a shared-memory ring.  New in v10:
//                   * the per-thread body advances W particles at once (float: W = 2, packed f32x2 arithmetic) -- pic_pair.cuh;
//                   * one dedicated producer warp feeds the ring; the consumer warps never touch the request logic;
//                   * supercell slices start on 16-byte boundaries (padded sort below), so a thread's pair is one 8-byte
//                     shared-memory load per array and one 8-byte global store per array, and no pair straddles two supercells;
//                   * chunks of 32 W particles are dealt round-robin, continuing across supercells (no dealing atomics);
//                   * everything that is not "stays in its cell, covered by the tile" is DEFERRED, not handled in line: particles that
//                     change cell (union-stencil deposit, wrap / reflect / absorb / ownership), particles whose stencil the tile
//                     does not cover (they drifted more than a cell out of their supercell since the last sort) and slots appended
//                     since the sort.  The tile kernel leaves such a slot untouched and writes its index into the CTA's segment of a
//                     work list; k_pair_fixup then runs the scalar global-memory body of K1 v8 (fused_particle_fast3d) over the
//                     list.  The result never depends on how fresh the sort is, and the hot kernel carries none of the rare
//                     paths' code or registers (measured: their mere presence cost the proton launch 10 %, the in-line handling
//                     of 4 % cell changers cost the electron launch 1.1 ms of 4.8 -- profiles/r02_k1_ablation.md).
//   k_sortb_*     counting sort by cell in the 4^3-blocked order with every supercell's slice padded to a multiple of 4 slots
//                 (NaN = dead slot, the convention of the resident layout), producing blk_off for K1 directly.
#include <stdlib.h>
#include <cuda.h>

#include "pic_common.cuh"
#include "pic_tma.cuh"
#include "pic_pair.cuh"

namespace pic {

#ifndef PIC_K10_NWC
#define PIC_K10_NWC 7          /* float: consumer warps per CTA (+ 1 producer warp) */
#endif
#ifndef PIC_K10_CTAS
#define PIC_K10_CTAS 2         /* float: CTAs per SM -> 128 registers per thread at 2 x 8 warps */
#endif
#ifndef PIC_K10_NWC64
#define PIC_K10_NWC64 15       /* double: one CTA per SM (the ring is twice as large) */
#endif
#ifndef PIC_K10_NSTAGE
#define PIC_K10_NSTAGE 3       /* float: ring stages per CTA */
#endif
#ifndef PIC_K10_NSTAGE64
#define PIC_K10_NSTAGE64 3     /* double */
#endif
#ifndef PIC_K10_STRIDED
#define PIC_K10_STRIDED 0      /* supercells -> CTAs: 0 = one contiguous range per CTA; G > 0 = groups of G consecutive supercells dealt
                                  round-robin, so the CTAs that run at the same time work on neighbouring supercells (L2 reuse of the
                                  overlapping E/B boxes and of J) */
#endif
#ifndef PIC_K10_EVICT
#define PIC_K10_EVICT 0        /* 1: particle slices are loaded and stored with an L2 evict-first policy (they are touched once per launch;
                                  the fields and J are what should stay in L2) */
#endif
#ifndef PIC_K10_CLAIM_ASM
#define PIC_K10_CLAIM_ASM 1    /* chunk claim: 1 = predicated atom.shared by lane 0 (no divergence), 0 = if (lane == 0) atomicAdd */
#endif
template <typename T> struct K10Ring { static constexpr int NSTAGE = sizeof(T) == 4 ? PIC_K10_NSTAGE : PIC_K10_NSTAGE64; };
#ifndef PIC_K10_DEAL   /* (historical: the kernel now always deals through the counter) */
#define PIC_K10_DEAL 1         /* chunks of a supercell reach the warps 0: round-robin continuing across supercells (4.20 ms per launch),
                                  1: through a shared-memory counter (3.91 ms: a warp held up by a queue flush no longer delays the
                                  release of its ring slot) -- profiles/r02_ab3_run.log */
#endif
#ifndef PIC_ABL10
#define PIC_ABL10 0           /* profiling builds only (tools/ab.py): WRONG RESULTS, each removes one cost centre so that its share of the
                                  launch can be read off a timing -- 1: tiles are loaded only for the first parts of a CTA (TMA tile fill);
                                  2: the x vertex anchor follows the centre anchor (gather bank conflicts); 3: no REDs; 4: no reduction and no
                                  REDs; 5: no particle stores; 6: cell changers are treated as stayers (queue, union deposit, wrap);
                                  7: all lanes gather from one address (no conflicts at all) */
#endif
#ifndef PIC_K10_W
#define PIC_K10_W 2            /* particles per thread in float (1 = scalar control) */
#endif
// staged particle slots per supercell and array (mean 512 at 8 ppc per species); double with the wider tile: 576, to fit 227 KB
template <typename T> struct K10Cap { static constexpr int PCAP = (sizeof(T) == 4 || PIC_TILE_YPAD > 1) ? 576 : 640; };

template <typename T, int W>
__device__ __forceinline__ Vec<T, W> ld_vec(const T* ptr) {
    Vec<T, W> r;
    if constexpr (W == 2 && sizeof(T) == 4) {
        const float2 f = *reinterpret_cast<const float2*>(ptr);
        r.v[0] = f.x; r.v[1] = f.y;
    } else if constexpr (W == 2 && sizeof(T) == 8) {
        const double2 f = *reinterpret_cast<const double2*>(ptr);
        r.v[0] = f.x; r.v[1] = f.y;
    } else {
#pragma unroll
        for (int j = 0; j < W; ++j) r.v[j] = ptr[j];
    }
    return r;
}
template <typename T, int W>
__device__ __forceinline__ void st_vec(T* ptr, const Vec<T, W>& a) {
    if constexpr (W == 2 && sizeof(T) == 4) *reinterpret_cast<float2*>(ptr) = make_float2(a.v[0], a.v[1]);
    else if constexpr (W == 2 && sizeof(T) == 8) *reinterpret_cast<double2*>(ptr) = make_double2(a.v[0], a.v[1]);
    else {
#pragma unroll
        for (int j = 0; j < W; ++j) ptr[j] = a.v[j];
    }
}

// Particle state of a chunk, read from the staged slice in shared memory every time it is asked (pic_pair.cuh ArrayLoader): the
// slice stays valid until the warp releases the stage, so x need not live in registers across the gather nor v before the push.
template <typename T, int W, int PCAP>
struct ChunkLoader {
    const T* st;            // staged slot of this thread's first particle: array c at st[c * PCAP]
    __device__ __forceinline__ Vec<T, W> pos(int a) const { return ld_vec<T, W>(st + a * PCAP); }
    __device__ __forceinline__ Vec<T, W> vel(int a) const { return ld_vec<T, W>(st + (3 + a) * PCAP); }
};

// the 12 same-cell currents of one cell -> global J (fire-and-forget REDs)
template <typename T>
__device__ __forceinline__ void red_cell(const TileSink<T>& sink, int base, int sx, int sy, const T* v) {
#if PIC_ABL10 == 3 || PIC_ABL10 == 4
    if (sx != 0x7fffffff) return;            // (always taken at run time; the values stay alive for the compiler)
#endif
    int n = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T* Jc = sink.J[c] + base;
#pragma unroll
        for (int m1 = 0; m1 < 2; ++m1)
#pragma unroll
            for (int m2 = 0; m2 < 2; ++m2) atomicAdd(Jc + SameCell<1>::offset(c, 0, m1, m2, sx, sy), v[n++]);
    }
}

// one level of the segmented scan (float: packed adds)
template <typename T>
__device__ __forceinline__ void pair_scan_level(T* v, int d, bool take) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int n = 0; n < 12; n += 2) {
            const float2 o = make_float2(__shfl_up_sync(0xffffffffu, v[n], d), __shfl_up_sync(0xffffffffu, v[n + 1], d));
            if (take) {
                const float2 r = __fadd2_rn(make_float2(v[n], v[n + 1]), o);
                v[n] = r.x; v[n + 1] = r.y;
            }
        }
    } else {
#pragma unroll
        for (int n = 0; n < 12; ++n) {
            const T o = __shfl_up_sync(0xffffffffu, v[n], d);
            if (take) v[n] += o;
        }
    }
}

// Segmented inclusive scan (runs of up to 1 << STEPS lanes with equal key); run tails issue the REDs.  key < 0: nothing to add.
template <typename T, int STEPS>
__device__ __forceinline__ void pair_scan_red(T* v, int key, int lane, const TileSink<T>& sink, int key0, int sx, int sy) {
    constexpr int G = 1 << STEPS;
    const int gl = lane & (G - 1);
    const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (gl == 0) || (key != key_prev);
    int flag = head ? 1 : 0;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const int fo = __shfl_up_sync(0xffffffffu, flag, d);
        const bool take = (gl >= d) && (flag == 0);
        pair_scan_level<T>(v, d, take);
        if (take) flag |= fo;
    }
    const int head_next = __shfl_down_sync(0xffffffffu, head ? 1 : 0, 1);
    const bool tail = (gl == G - 1) || (head_next != 0);
    if (tail && key >= 0) red_cell<T>(sink, key0 + (key >> 6) * sx + ((key >> 3) & 7) * sy + (key & 7), sx, sy, v);
}

// All lanes of the warp with the same key are summed, contiguous or not (match-any groups, pointer doubling); the lowest lane of
// every group issues the REDs.  For a species whose sort is stale (cell changers fragment the runs).
template <typename T>
__device__ __forceinline__ void pair_group_red(T* v, int key, int lane, const TileSink<T>& sink, int key0, int sx, int sy) {
    const unsigned group = __match_any_sync(0xffffffffu, key);
    const unsigned above = group & (0xfffffffeu << lane);
    int next = above ? __ffs(above) - 1 : 32;
    while (__any_sync(0xffffffffu, next < 32)) {
        const bool has = next < 32;
        const int src = has ? next : lane;
        if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int n = 0; n < 12; n += 2) {
                const float2 o = make_float2(__shfl_sync(0xffffffffu, v[n], src), __shfl_sync(0xffffffffu, v[n + 1], src));
                if (has) {
                    const float2 r = __fadd2_rn(make_float2(v[n], v[n + 1]), o);
                    v[n] = r.x; v[n + 1] = r.y;
                }
            }
        } else {
#pragma unroll
            for (int n = 0; n < 12; ++n) {
                const T o = __shfl_sync(0xffffffffu, v[n], src);
                if (has) v[n] += o;
            }
        }
        const int nn = __shfl_sync(0xffffffffu, next, src);
        next = has ? nn : 32;
    }
    if (key >= 0 && lane == __ffs(group) - 1) red_cell<T>(sink, key0 + (key >> 6) * sx + ((key >> 3) & 7) * sy + (key & 7), sx, sy, v);
}

template <typename T> struct K10Stage { static constexpr int ELEMS = 6 * TILE_ELEMS + 6 * K10Cap<T>::PCAP; };
constexpr int K10_HDR = 512;    // barriers + descriptors in front of the ring

// Same-cell reduction through shared memory (float): every lane parks its 12 values in the warp's scratch rows; the contiguous
// runs of equal keys (<= 32) are found with one ballot; then lane t sums quad q = t % 3 (the four values of current component q)
// of run t / 3 over the run's rows -- one LDS.128 + 4 adds per row, independent accumulators -- and issues that component's four
// REDs.  Against the segmented scan: ~70 instead of ~160 instructions per chunk, 4 instead of 12 RED instructions, no shuffle
// chains on the critical path; 10 runs per pass (a freshly sorted chunk of 64 particles has 8 - 9).
constexpr int K10_RED_ROW = 12;                       // floats per scratch row
constexpr int K10_RED_BYTES = 32 * K10_RED_ROW * 4 + 32 * 8;    // scratch rows + (start, len, key) per run
__device__ __forceinline__ void pair_smem_red(const float* lv, int key, int lane, const TileSink<float>& sink, int key0, int sx, int sy,
                                              float* scratch, int2* runs) {
    float4* row = reinterpret_cast<float4*>(scratch + lane * K10_RED_ROW);
    row[0] = make_float4(lv[0], lv[1], lv[2], lv[3]);
    row[1] = make_float4(lv[4], lv[5], lv[6], lv[7]);
    row[2] = make_float4(lv[8], lv[9], lv[10], lv[11]);
    const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (lane == 0) || (key != key_prev);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    if (head) {
        const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
        const int len = above ? __ffs(above) : 32 - lane;
        runs[__popc(heads & ((1u << lane) - 1u))] = make_int2(lane | (len << 8), key);
    }
    __syncwarp();
    const int nrun = __popc(heads);
    const int t3 = lane / 3, q = lane - 3 * t3;
    for (int base = 0; base < nrun; base += 10) {
        const int r = base + t3;
        if (lane < 30 && r < nrun) {
            const int2 info = runs[r];
            if (info.y >= 0) {
                const int start = info.x & 0xff, len = info.x >> 8;
                const float4* src = reinterpret_cast<const float4*>(scratch + start * K10_RED_ROW) + q;
                float4 acc = src[0];
                for (int i = 1; i < len; ++i) {
                    const float4 v = src[i * (K10_RED_ROW / 4)];
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                const int k = info.y;
                float* J = (q == 0 ? sink.J[0] : (q == 1 ? sink.J[1] : sink.J[2])) + (key0 + (k >> 6) * sx + ((k >> 3) & 7) * sy + (k & 7));
                const int o1 = (q == 2) ? sy : 1, o2 = (q == 0) ? sy : sx;      // SameCell<1>::offset(c = q, 0, m1, m2)
                atomicAdd(J, acc.x); atomicAdd(J + o1, acc.y); atomicAdd(J + o2, acc.z); atomicAdd(J + o1 + o2, acc.w);
            }
        }
    }
    __syncwarp();
}

template <typename T, int W, int NWC, int MODE>
struct PairSmem {
    static constexpr size_t bytes = (size_t)K10_HDR + (size_t)K10Ring<T>::NSTAGE * K10Stage<T>::ELEMS * sizeof(T)
                                    + ((sizeof(T) == 4 && MODE == 3) ? (size_t)NWC * K10_RED_BYTES : 0);
};

// Work list of the deferred slots (see the header): CTA c of the tile kernel appends to the segment that starts at the first slot of
// its supercell range (a segment is as long as the range holds slots, so it cannot overflow) and leaves the entry count in
// cnt[c].  An entry is the slot index plus, for a cell changer, its position BEFORE the move (the tile kernel has already stored
// the new velocity and the new, not yet wrapped, position in the slot); bit 31 of the index marks a slot the tile did not cover:
// untouched, the fix-up pass advances it from scratch (counted in flags[2]).
template <typename T>
struct DeferList {
    int32_t* cnt;      // [K10_MAX_GRID]
    int32_t* idx;      // [cap]
    T* pos[3];         // [cap] each
};
constexpr int K10_DEFER_UNCOVERED = (int)0x80000000u;
constexpr int K10_MAX_GRID = 1024;      // CTAs of the tile kernel = leading entries of the work array (include/pic_b200.h PIC_PAIR_WORK_HEAD)

// next chunk of the part in stage `slot`: one shared-memory atomic by lane 0 (predicated, the warp does not diverge)
__device__ __forceinline__ int claim_chunk(int* counter, int lane) {
    int v = 0;
#if !PIC_K10_CLAIM_ASM
    if (lane == 0) v = atomicAdd(counter, 1);
    return v;
#endif
    asm volatile("{\n .reg .pred p;\n setp.eq.s32 p, %2, 0;\n @p atom.shared.add.u32 %0, [%1], 1;\n}\n"
                 : "+r"(v) : "r"(smem_u32(counter)), "r"(lane) : "memory");
    return v;
}

// MODE: 0 = segmented scan + RED, 2 = match-any groups + RED (same numbering as K1 v9), 3 = shared-memory segmented reduction
template <typename T, int W, int PUSHER, int NWC, int CTAS, bool PER1, int MODE>
__global__ void __launch_bounds__((NWC + 1) * 32, CTAS)
k_pair3d(const __grid_constant__ FastConst<T> k, const __grid_constant__ PairConst<T> pc, const __grid_constant__ SoAView<T> s, Field3W<T> J,
         const __grid_constant__ DeferList<T> defer, int32_t* flags, const __grid_constant__ TileMaps tm,
         const int32_t* __restrict__ blk_off, int nblk, int nbx, int nby, int nbz, int g, int seg_cap) {
    constexpr int NSTAGE = K10Ring<T>::NSTAGE;
    constexpr int PCAP = K10Cap<T>::PCAP;
    constexpr int TILE_ALL = 6 * TILE_ELEMS;
    constexpr int STAGE_ELEMS = K10Stage<T>::ELEMS;
    constexpr int AL = 16 / (int)sizeof(T);
    constexpr int CH = 32 * W;                               // particles per chunk (one warp iteration)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [NSTAGE] tile + particles have landed (TMA complete_tx)
    uint64_t* empty = full + NSTAGE;                         // [NSTAGE] every consumer warp is done with the stage
    int* desc = reinterpret_cast<int*>(smem_raw + 64);       // [NSTAGE][8]: slice begin, end, end-of-stream, edge flag, J key of the tile origin, chunk counter
    int* ctl = reinterpret_cast<int*>(smem_raw + 192);       // [0] entries in this CTA's work-list segment, [1] consumer warps that have finished
    T* dx0 = reinterpret_cast<T*>(smem_raw + 256);           // [NSTAGE][4]: position of the tile's first centre-line node
    T* stages = reinterpret_cast<T*>(smem_raw + K10_HDR);    // [NSTAGE]([6][8][TILE_NY][8] + [6][PCAP])
    unsigned char* red_raw0 = smem_raw + K10_HDR + (size_t)NSTAGE * STAGE_ELEMS * sizeof(T);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_cta = (nblk + (int)gridDim.x - 1) / (int)gridDim.x;
    const int b0 = (int)blockIdx.x * per_cta < nblk ? (int)blockIdx.x * per_cta : nblk;
    const int b1 = (b0 + per_cta < nblk) ? b0 + per_cta : nblk;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, NWC); }
        ctl[0] = 0; ctl[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int n_live = (int)s.count();
    const int cap_al = (int)(s.cap < 0x7fffffff ? s.cap : 0x7fffffff) & ~(AL - 1);

    // ------------------------------------------------------------------ producer warp: one thread feeds the ring
    if (warp == NWC) {
        if (lane == 0) {
            // The ring carries a stream of PARTS: a supercell with up to PCAP particles is one part, a denser one is cut into parts of
            // PCAP slots (same tile, next slice), an empty one is skipped; a final part with the `done` mark ends the stream.
            // Slice bounds are read two supercells ahead of their use, so the producer never sits on a DRAM round trip between a
            // stage being released and its refill being issued.
            int j = 0;                                           // parts issued so far
#if PIC_K10_EVICT
            const uint64_t pol_stream = l2_policy_evict_first();
#endif
            auto wait_slot = [&](int sr) {                       // previous occupant: part j - NSTAGE (suspended wait, no spinning)
                if (j >= NSTAGE) while (!mbar_try_wait(empty + sr, ((j / NSTAGE) - 1) & 1, 20000)) {}
            };
#if PIC_K10_STRIDED > 0
            {
                constexpr int G = PIC_K10_STRIDED;
                const int ngrp = (nblk + G - 1) / G;
                auto blk_of = [&](int t) { const int grp = (t / G) * (int)gridDim.x + (int)blockIdx.x; return grp < ngrp ? grp * G + t % G : nblk; };
                int bn = blk_of(0);
                int beg_n = blk_off[bn < nblk ? bn : nblk], end_n = blk_off[bn < nblk ? bn + 1 : nblk];
                for (int t = 0; bn < nblk; ++t) {
                    const int b = bn;
                    int beg = beg_n, end = end_n;
                    bn = blk_of(t + 1);
                    if (bn < nblk) { beg_n = blk_off[bn]; end_n = blk_off[bn + 1]; }     // (consumed in the next iteration)
#else
            if (b1 > b0) {
                int beg = blk_off[b0];
                int end_next = blk_off[b0 + 1];
                int end_next2 = (b0 + 2 <= nblk) ? blk_off[b0 + 2] : end_next;
                for (int b = b0; b < b1; ++b) {
                    int end = end_next;
                    const int beg_following = end;
                    end_next = end_next2;
                    if (b + 3 <= nblk) end_next2 = blk_off[b + 3];
#endif
                    if (end > n_live) end = n_live;
                    if (beg & (AL - 1)) { atomicOr(flags, 8); end = beg; }   // contract: slices come from the padded sort (pic_sort_blocked_*)
                    const int bz = b % nbz, by = (b / nbz) % nby, bx = b / (nbz * nby);
                    const int ox = bx * TILE_B + g - 2, oy = by * TILE_B + g - 2, oz = bz * TILE_B + g - 2;
                    for (int pb = beg; pb < end; pb += PCAP, ++j) {
                        const int sr = j % NSTAGE;
                        const int pe = (pb + PCAP < end) ? pb + PCAP : end;
                        int n = (pe - pb + AL - 1) & ~(AL - 1);
                        if (pb + n > cap_al) n = cap_al - pb;
                        wait_slot(sr);
                        int* d = desc + sr * 8;
                        d[0] = pb; d[1] = pe; d[2] = 0;
                        d[3] = (bx == 0 || bx == nbx - 1 || by == 0 || by == nby - 1 || bz == 0 || bz == nbz - 1) ? 1 : 0;
                        d[4] = ox * k.sx + oy * k.sy + oz;
                        d[5] = 0;                                    // next undealt chunk of this part
                        T* x0 = dx0 + sr * 4;
                        x0[0] = pic_fma((T)ox, k.sc[0], k.oc[0]); x0[1] = pic_fma((T)oy, k.sc[1], k.oc[1]); x0[2] = pic_fma((T)oz, k.sc[2], k.oc[2]);
                        T* st = stages + sr * STAGE_ELEMS;
#if PIC_ABL10 == 1
                        const bool fill_tile = j < NSTAGE;
#else
                        const bool fill_tile = true;
#endif
                        mbar_arrive_expect_tx(full + sr, (fill_tile ? TILE_ALL * (int)sizeof(T) : 0) + 6 * n * (int)sizeof(T));
                        if (fill_tile) {
#pragma unroll
                            for (int c = 0; c < 6; ++c) tma_load_box(st + c * TILE_ELEMS, &tm.m[c], full + sr, oz, oy, ox);
                        }
#pragma unroll
                        for (int c = 0; c < 6; ++c) {
#if PIC_K10_EVICT
                            tma_load_bytes_hint(st + TILE_ALL + c * PCAP, s.c[c] + pb, n * (int)sizeof(T), full + sr, pol_stream);
#else
                            tma_load_bytes(st + TILE_ALL + c * PCAP, s.c[c] + pb, n * (int)sizeof(T), full + sr);
#endif
                        }
                    }
#if PIC_K10_STRIDED == 0
                    beg = beg_following;
#endif
                }
            }
            const int sr = j % NSTAGE;                           // end of the stream
            wait_slot(sr);
            desc[sr * 8 + 2] = 1;
            mbar_arrive(full + sr);
        }
        return;
    }

    // ------------------------------------------------------------------ consumer warps
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = J.f[c]; sink.L[c] = k.L[c]; }
    sink.off = 0;
    sink.flags = flags;
    unsigned char* red_raw = red_raw0 + (size_t)warp * K10_RED_BYTES;   // (MODE 3 only)
#if PIC_K10_STRIDED > 0
    const int dq0 = (int)blockIdx.x * seg_cap;             // equal shares of the list (overflow: flags |= 16)
#else
    const int dq0 = b0 < nblk ? blk_off[b0] : 0;          // first entry of this CTA's work-list segment
#endif
#if PIC_K10_EVICT
    const uint64_t pol_store = l2_policy_evict_first();
#endif
    int slot = 0, par = 0;
    for (;;) {
        while (!mbar_try_wait(full + slot, par, 1000)) {}
        const int4 d0 = *reinterpret_cast<const int4*>(desc + slot * 8);
        if (d0.z) break;                                         // end of this CTA's stream
        const int key0 = desc[slot * 8 + 4];
        const int p_beg = d0.x, p_end = d0.y;
        const bool edge = d0.w != 0;
        const T x0[3] = {dx0[slot * 4], dx0[slot * 4 + 1], dx0[slot * 4 + 2]};
        const T* tile = stages + slot * STAGE_ELEMS;
        const T* pst = tile + TILE_ALL;
        const int nchunk = (p_end - p_beg + CH - 1) / CH;
        // chunks are dealt through the part's counter; the NEXT chunk is claimed before the current one is processed, so the
        // shared-memory atomic's round trip hides behind the body (a warp over-claims one chunk past the end: harmless)
        int ch_claim = claim_chunk(desc + slot * 8 + 5, lane);
        for (;;) {
            const int ch = __shfl_sync(0xffffffffu, ch_claim, 0);
            if (ch >= nchunk) break;
            ch_claim = claim_chunk(desc + slot * 8 + 5, lane);
            const int i0 = p_beg + ch * CH + W * lane;
            const ChunkLoader<T, W, PCAP> ld{pst + (i0 - p_beg)};
            bool live[W];
            {
                const Vec<T, W> px = ld.pos(0);                  // (lanes past p_end read stage memory that is never used)
#pragma unroll
                for (int j = 0; j < W; ++j) live[j] = (i0 + j < p_end) && !pic_isnan(px.v[j]);
            }
            Vec<T, W> pos_out[3], vel_out[3], xraw[3], vals[12];
            int kind[W], cid[W];
            pair_advance_ld<T, W, PUSHER, PER1, ChunkLoader<T, W, PCAP>>(k, pc, tile, x0, edge, ld, live, pos_out, vel_out, xraw, kind, cid, vals);
            // ---- store: stayers are finished; a cell changer leaves its new velocity and its raw new position (k_pair_fixup deposits
            //      it over the union stencil and applies the boundary conditions); an uncovered particle keeps its old state
            {
                bool stv[W];
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    stv[j] = (kind[j] == PAIR_SAME) || (kind[j] == PAIR_CROSS);
#pragma unroll
                    for (int a = 0; a < 3; ++a) pos_out[a].v[j] = (kind[j] == PAIR_CROSS) ? xraw[a].v[j] : pos_out[a].v[j];
                }
                bool all = true;
#pragma unroll
                for (int j = 0; j < W; ++j) all = all && stv[j];
#if PIC_ABL10 == 5
                if (nblk != 0x7fffffff) all = false, stv[0] = stv[W - 1] = false;
#endif
                if (all) {
#if PIC_K10_EVICT
                    if constexpr (W == 2 && sizeof(T) == 4) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            st_f2_hint(s.c[c] + i0, pos_out[c].v[0], pos_out[c].v[1], pol_store);
                            st_f2_hint(s.c[3 + c] + i0, vel_out[c].v[0], vel_out[c].v[1], pol_store);
                        }
                    } else
#endif
#pragma unroll
                    for (int c = 0; c < 3; ++c) { st_vec<T, W>(s.c[c] + i0, pos_out[c]); st_vec<T, W>(s.c[3 + c] + i0, vel_out[c]); }
                } else {
#pragma unroll
                    for (int j = 0; j < W; ++j)
                        if (stv[j]) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) { s.c[c][i0 + j] = pos_out[c].v[j]; s.c[3 + c][i0 + j] = vel_out[c].v[j]; }
                        }
                }
            }
            // ---- defer the cell changers and the uncovered particles to the CTA's work-list segment
            {
                unsigned m[W];
                unsigned any = 0;
#pragma unroll
                for (int j = 0; j < W; ++j) { m[j] = __ballot_sync(0xffffffffu, kind[j] >= PAIR_CROSS); any |= m[j]; }
                if (any) {
                    int total = 0;
#pragma unroll
                    for (int j = 0; j < W; ++j) total += __popc(m[j]);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(ctl, total);
                    base = __shfl_sync(0xffffffffu, base, 0);
#if PIC_K10_STRIDED > 0
                    if (base + total > seg_cap) { if (lane == 0) atomicOr(flags, 16); total = 0; for (int j = 0; j < W; ++j) m[j] = 0, kind[j] = PAIR_NONE; }
#endif
                    base += dq0;
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        if (kind[j] >= PAIR_CROSS) {
                            const int e = base + __popc(m[j] & ((1u << lane) - 1u));
                            defer.idx[e] = (i0 + j) | (kind[j] == PAIR_SLOW ? K10_DEFER_UNCOVERED : 0);
#pragma unroll
                            for (int a = 0; a < 3; ++a) defer.pos[a][e] = ld.pos(a).v[j];
                        }
                        base += __popc(m[j]);
                    }
                }
            }
            // ---- same-cell currents: join the thread's particles, reduce over the lanes of one cell, RED at the run tails
            {
                T lv[12];
                int key;
                if constexpr (W == 2) {
                    const bool dA = kind[0] == PAIR_SAME, dB = kind[1] == PAIR_SAME;
                    const bool solo = dA && dB && (cid[0] != cid[1]);      // the pair straddles a cell boundary: B goes out by itself
                    const T mB = solo ? (T)0 : (T)1;                       // (B's values are exact zeros unless it deposits: one FMA per value)
#pragma unroll
                    for (int n = 0; n < 12; ++n) lv[n] = pic_fma(vals[n].v[1], mB, vals[n].v[0]);
                    key = dA ? cid[0] : (dB ? cid[1] : -1 - lane);
                    if (solo) {
                        T bv[12];
#pragma unroll
                        for (int n = 0; n < 12; ++n) bv[n] = vals[n].v[1];
                        red_cell<T>(sink, key0 + (cid[1] >> 6) * k.sx + ((cid[1] >> 3) & 7) * k.sy + (cid[1] & 7), k.sx, k.sy, bv);
                    }
                } else {
#pragma unroll
                    for (int n = 0; n < 12; ++n) lv[n] = vals[n].v[0];
                    key = (kind[0] == PAIR_SAME) ? cid[0] : -1 - lane;
                }
#if PIC_ABL10 == 4
                if (k.sx == 0x7fffffff)
#endif
                if constexpr (MODE == 3 && sizeof(T) == 4)
                    pair_smem_red(lv, key, lane, sink, key0, k.sx, k.sy, reinterpret_cast<float*>(red_raw),
                                  reinterpret_cast<int2*>(red_raw + 32 * K10_RED_ROW * 4));
                else if (MODE == 2) pair_group_red<T>(lv, key, lane, sink, key0, k.sx, k.sy);
                else pair_scan_red<T, 3>(lv, key, lane, sink, key0, k.sx, k.sy);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);            // this warp no longer reads the stage in `slot`
        if (++slot == NSTAGE) { slot = 0; par ^= 1; }
    }
    // ---- the last consumer warp to finish publishes the length of the CTA's work-list segment
    __syncwarp();
    if (lane == 0) {
        // (an acq_rel atomic, not __threadfence_block(): with a fence in the kernel ptxas turns every fire-and-forget REDG of the
        // deposit into a returning ATOMG -- measured 3.5 instead of 3.0 ms per launch)
        unsigned done, n_def;
        asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;\n" : "=r"(done) : "r"(smem_u32(ctl + 1)) : "memory");
        if (done == NWC - 1) {
            asm volatile("ld.acquire.cta.shared.u32 %0, [%1];\n" : "=r"(n_def) : "r"(smem_u32(ctl)) : "memory");
            defer.cnt[blockIdx.x] = (int)n_def;
        }
    }
}

// Second pass of K1 v10: the deferred slots of every CTA's work-list segment -- cell changers: union-stencil Esirkepov deposit of
// the move (old position from the list, new position from the slot), particle boundary conditions, ownership, store
// (crosser_finish); uncovered slots: the whole step through the scalar global-memory body -- then the slots appended since the
// last sort (particles received from neighbour ranks), also through the scalar body.
template <typename T, int PUSHER, bool PER1>
__global__ void __launch_bounds__(256)
k_pair_fixup(const __grid_constant__ PicParams p, int species, const __grid_constant__ Geom<T> gm, const __grid_constant__ FastConst<T> k,
             const __grid_constant__ SoAView<T> s, const __grid_constant__ Field6<T> F, Field3W<T> J, const __grid_constant__ LeaveBuf leave,
             int distributed, int32_t* flags, const __grid_constant__ DeferList<T> defer, const int32_t* __restrict__ blk_off, int nblk,
             int tile_grid, int sub, int seg_cap) {
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = J.f[c]; sink.L[c] = gm.L[c]; }
    sink.off = 0;
    sink.flags = flags;
    Field6<T> X;
    for (int c = 0; c < 6; ++c) X.f[c] = nullptr;
    // `sub` CTAs of this kernel share the segment of one CTA of the tile kernel
    const int seg = (int)blockIdx.x / sub, part = (int)blockIdx.x % sub;
    const int per_cta = (nblk + tile_grid - 1) / tile_grid;
    const int b0 = seg * per_cta < nblk ? seg * per_cta : nblk;
#if PIC_K10_STRIDED > 0
    const int dq0 = seg * seg_cap;
    (void)b0;
#else
    const int dq0 = blk_off[b0];
#endif
    const int n = defer.cnt[seg];
    int uncovered = 0;
    for (int e = part * 256 + (int)threadIdx.x; e < n; e += sub * 256) {
        const int v = defer.idx[dq0 + e];
        const int i = v & 0x7fffffff;
        if (v < 0) {
            ++uncovered;
            fused_particle_fast3d<T, 1, PUSHER, false>(p, species, gm, k, (int64_t)i, s, F, X, sink, leave, distributed != 0, flags);
        } else {
            const T po[3] = {defer.pos[0][dq0 + e], defer.pos[1][dq0 + e], defer.pos[2][dq0 + e]};
            const T xn[3] = {s.c[0][i], s.c[1][i], s.c[2][i]};
            crosser_finish<T, PER1>(p, species, gm, k, (int64_t)i, s, po, xn, sink, leave, distributed != 0, flags);
        }
    }
    uncovered = __reduce_add_sync(0xffffffffu, uncovered);
    if ((threadIdx.x & 31) == 0 && uncovered) atomic_add_i32(flags + 2, uncovered);
    const int n_live = (int)s.count();
    for (int i = blk_off[nblk] + (int)blockIdx.x * 256 + (int)threadIdx.x; i < n_live; i += (int)gridDim.x * 256)
        fused_particle_fast3d<T, 1, PUSHER, false>(p, species, gm, k, (int64_t)i, s, F, X, sink, leave, distributed != 0, flags);
}

// ---------------------------------------------------------------- blocked, padded counting sort
// Cells are ordered supercell-major (local_cell, PIC_SORT_BLOCK = 4); every supercell's slice starts on a multiple of 4 slots.
// k_sortb_blocks: padded particle count per supercell.  (exclusive scan: pic_sort_scan)  k_sortb_cells: first slot of every cell;
// k_sortb_pad: the padding slots are dead (x = NaN).  The scatter is kernels_fast.cu k_sort_scatter.
__global__ void __launch_bounds__(256) k_sortb_blocks(int nblk, const int32_t* __restrict__ cell_count, int32_t* __restrict__ blk_padded) {
    // one warp per supercell: 64 cell counts -> padded sum; entry nblk is 0 (scan total lands there)
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w > nblk) return;
    int v = 0;
    if (w < nblk) v = cell_count[w * 64 + lane] + cell_count[w * 64 + 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) blk_padded[w] = (w < nblk) ? ((v + 3) & ~3) : 0;
}

__global__ void __launch_bounds__(256) k_sortb_cells(int nblk, const int32_t* __restrict__ cell_count, const int32_t* __restrict__ blk_off,
                                                     int32_t* __restrict__ cell_offset, int64_t cap, int32_t* flags) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nblk) {
        if (w == nblk && lane == 0) {
            cell_offset[nblk * 64] = blk_off[nblk];          // the trash bin (dead particles) starts after the last padded slice
            if ((int64_t)blk_off[nblk] > cap) atomicOr(flags, 2);
        }
        return;
    }
    const int c0 = cell_count[w * 64 + 2 * lane], c1 = cell_count[w * 64 + 2 * lane + 1];
    int inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    const int excl = blk_off[w] + inc - (c0 + c1);
    cell_offset[w * 64 + 2 * lane] = excl;
    cell_offset[w * 64 + 2 * lane + 1] = excl + c0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_sortb_pad(int nblk, const int32_t* __restrict__ cell_offset, const int32_t* __restrict__ cell_count,
                                                   const int32_t* __restrict__ blk_off, SoAView<T> d) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int last = b * 64 + 63;
    const int used_end = cell_offset[last] + cell_count[last];
    const int end = blk_off[b + 1];
    for (int i = used_end; i < end && i < d.cap; ++i) {
        d.c[0][i] = pic_nan<T>();
        for (int c = 1; c < 6; ++c) d.c[c][i] = (T)0;
        if (d.id) d.id[i] = -1;
    }
}

__global__ void k_set_count_pair(int32_t* n_dev, const int32_t* src, int64_t cap) {
    const int v = *src;
    *n_dev = (int64_t)v < cap ? v : (int32_t)cap;
}

// ---------------------------------------------------------------- launchers
}  // namespace pic
extern "C" int64_t pic_pair_work_bytes(const PicParams* p, int64_t cap);
namespace pic {
template <typename T>
static int launch_pair3d(const PicParams* p, int species, const PicSoA* soa, const int32_t* blk_off, int nblk, int options, const void* const E[3],
                         const void* const B[3], void* const J[3], const PicLeave* leave, int32_t* flags, void* work, int64_t work_len,
                         cudaStream_t st) {
    static_assert(PIC_SORT_BLOCK == TILE_B, "the tile kernel walks the supercells of the blocked sort order");
    if (p->shape_factor != 1 || p->g != 2 || (p->pusher != PIC_PUSHER_BORIS && p->pusher != PIC_PUSHER_BORIS_REL)) return PIC_EUNSUPPORTED;
    for (int a = 0; a < 3; ++a) {
        if (p->tile[a] % TILE_B != 0 || p->gmesh[a] * p->tile[a] <= 1) return PIC_EUNSUPPORTED;
    }
    const int nbx = p->tile[0] / TILE_B, nby = p->tile[1] / TILE_B, nbz = p->tile[2] / TILE_B;
    if (nblk != nbx * nby * nbz) return PIC_EINVAL;
    for (int c = 0; c < 3; ++c)
        if (((uintptr_t)E[c] | (uintptr_t)B[c]) & 15) return PIC_EUNSUPPORTED;     // TMA global addresses are 16-byte aligned
    for (int c = 0; c < 6; ++c)
        if ((uintptr_t)soa->comp[c] & 15) return PIC_EUNSUPPORTED;
    if (soa->n == 0 && !soa->n_dev) return 0;
    Field6<T> F;
    Field3W<T> Jw;
    for (int c = 0; c < 3; ++c) { F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c]; Jw.f[c] = (T*)J[c]; }
    Geom<T> gm;
    make_geom<T>(*p, 0, 0, 0, gm);
    FastConst<T> k;
    make_fast_const<T>(*p, species, gm, k);
    PairConst<T> pc;
    make_pair_const<T>(k, pc);
    // the pair body takes the vertex line to be the centre line shifted up by half a cell (utilities/grids.py): verify
    for (int a = 0; a < 3; ++a)
        if (fabs((double)(gm.ov[a] - gm.oc[a]) - 0.5 * (double)gm.d[a]) > 1e-4 * (double)gm.d[a]) return PIC_EUNSUPPORTED;
    int distributed = 0;
    for (int c = 0; c < 3; ++c) distributed |= (p->gmesh[c] != p->mesh[c]);
    const bool grp = (options & 2) != 0;
    const bool smr = !grp && (options & 4) != 0 && sizeof(T) == 4;      // shared-memory segmented reduction (float)
    TileMaps tm;
    if (!make_tile_maps<T>(gm.L, F.f, Jw.f, tm)) return PIC_EUNSUPPORTED;
    bool per1 = !distributed;
    for (int a = 0; a < 3; ++a) per1 = per1 && (p->particle_bc[a] == PIC_BC_PERIODIC);
    constexpr bool F32 = sizeof(T) == 4;
    constexpr int W = F32 ? PIC_K10_W : 1;
    constexpr int NWC = F32 ? PIC_K10_NWC : PIC_K10_NWC64;
    constexpr int CTAS = F32 ? PIC_K10_CTAS : 1;
    int grid = num_sms() * CTAS;
    if (grid > nblk) grid = nblk;
    if (grid > K10_MAX_GRID) grid = K10_MAX_GRID;
    // work: [K10_MAX_GRID] segment lengths, [cap] slot indices, 3 x [cap] old positions (T)
    const int64_t cap_w = (soa->cap + 3) & ~(int64_t)3;
    if (work_len < (int64_t)pic_pair_work_bytes(p, soa->cap)) return PIC_EINVAL;
    DeferList<T> defer;
    defer.cnt = (int32_t*)work;
    defer.idx = (int32_t*)work + K10_MAX_GRID;
    for (int a = 0; a < 3; ++a) defer.pos[a] = (T*)((int32_t*)work + K10_MAX_GRID + cap_w) + a * cap_w;
    const SoAView<T> sv = view_of<T>(soa);
    const LeaveBuf lb = leave_of(leave);
    const int seg_cap = (int)((cap_w / grid) & ~(int64_t)3);
#define PIC_LAUNCH_K10(PUSH, PER, MD)                                                                                    \
    do {                                                                                                                 \
        constexpr size_t smem = PairSmem<T, W, NWC, MD>::bytes;                                                          \
        static_assert(smem <= 227 * 1024, "ring exceeds the shared memory of one CTA");                                 \
        static bool attr_set = false;                                                                                    \
        if (!attr_set) {                                                                                                 \
            cudaError_t e = cudaFuncSetAttribute(k_pair3d<T, W, PUSH, NWC, CTAS, PER, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                                         \
            attr_set = true;                                                                                             \
        }                                                                                                                \
        k_pair3d<T, W, PUSH, NWC, CTAS, PER, MD><<<grid, (NWC + 1) * 32, smem, st>>>(k, pc, sv, Jw, defer, flags, tm, blk_off, nblk, nbx, nby, nbz, p->g, seg_cap); \
    } while (0)
#define PIC_LAUNCH_K10_M(PUSH, PER) do { if (grp) PIC_LAUNCH_K10(PUSH, PER, 2); else if (smr) PIC_LAUNCH_K10(PUSH, PER, 3); else PIC_LAUNCH_K10(PUSH, PER, 0); } while (0)
    constexpr int SUB = 4;
    if (p->pusher == PIC_PUSHER_BORIS) {
        if (per1) PIC_LAUNCH_K10_M(PIC_PUSHER_BORIS, true); else PIC_LAUNCH_K10_M(PIC_PUSHER_BORIS, false);
        if (per1) k_pair_fixup<T, PIC_PUSHER_BORIS, true><<<grid * SUB, 256, 0, st>>>(*p, species, gm, k, sv, F, Jw, lb, distributed, flags, defer, blk_off, nblk, grid, SUB, seg_cap);
        else k_pair_fixup<T, PIC_PUSHER_BORIS, false><<<grid * SUB, 256, 0, st>>>(*p, species, gm, k, sv, F, Jw, lb, distributed, flags, defer, blk_off, nblk, grid, SUB, seg_cap);
    } else {
        if (per1) PIC_LAUNCH_K10_M(PIC_PUSHER_BORIS_REL, true); else PIC_LAUNCH_K10_M(PIC_PUSHER_BORIS_REL, false);
        if (per1) k_pair_fixup<T, PIC_PUSHER_BORIS_REL, true><<<grid * SUB, 256, 0, st>>>(*p, species, gm, k, sv, F, Jw, lb, distributed, flags, defer, blk_off, nblk, grid, SUB, seg_cap);
        else k_pair_fixup<T, PIC_PUSHER_BORIS_REL, false><<<grid * SUB, 256, 0, st>>>(*p, species, gm, k, sv, F, Jw, lb, distributed, flags, defer, blk_off, nblk, grid, SUB, seg_cap);
    }
#undef PIC_LAUNCH_K10_M
#undef PIC_LAUNCH_K10
    PIC_LAUNCH_RET();
}

template <typename T>
static int launch_sortb_pad(const PicParams* p, int nblk, const int32_t* cell_offset, const int32_t* cell_count, const int32_t* blk_off,
                            const PicSoA* dst, cudaStream_t st) {
    k_sortb_pad<T><<<(nblk + 255) / 256, 256, 0, st>>>(nblk, cell_offset, cell_count, blk_off, view_of<T>(dst));
    PIC_LAUNCH_RET();
}

}  // namespace pic

using namespace pic;

extern "C" {

int pic_fused_pair3d(const PicParams* p, int species, const PicSoA* soa, const int32_t* blk_off, int nblk, int options,
                     const void* const E[3], const void* const B[3], void* const J[3], const PicLeave* leave, int32_t* flags,
                     void* work, int64_t work_len, void* stream) {
    PIC_CHECK_ARG(p && soa && blk_off && E && B && J && flags && work && species >= 0 && species < p->n_species && nblk > 0);
    PIC_CHECK_ARG(p->mesh[0] == 1 && p->mesh[1] == 1 && p->mesh[2] == 1);
    bool distributed = false;
    for (int c = 0; c < 3; ++c) distributed |= (p->gmesh[c] != p->mesh[c]);
    if (distributed) PIC_CHECK_ARG(leave && leave->buf);
    PIC_DISPATCH_T(p, launch_pair3d, p, species, soa, blk_off, nblk, options, E, B, J, leave, flags, work, work_len, (cudaStream_t)stream);
}

// Size in bytes of the `work` array pic_fused_pair3d needs for an SoA of capacity `cap` in the dtype of `p`.
int64_t pic_pair_work_bytes(const PicParams* p, int64_t cap) {
    if (!p || cap < 0) return -1;
    const int64_t cap_w = (cap + 3) & ~(int64_t)3;
    const int64_t real = (p->dtype == PIC_F32) ? 4 : 8;
    return 4 * ((int64_t)K10_MAX_GRID + cap_w) + 3 * cap_w * real;
}

// Offsets of the blocked, padded sort from the per-cell histogram (pic_sort_histogram): blk_off[nblk + 1] (first slot of every
// supercell, multiples of 4) and cell_offset[ncells + 1] (first slot of every cell; the last entry is the trash bin = total padded
// length).  blk_work: nblk + 1 ints; scan_scratch as for pic_sort_scan.  flags |= 2 when the padded stream exceeds `cap`.
int pic_sort_blocked_offsets(const PicParams* p, const int32_t* cell_count, int32_t* cell_offset, int32_t* blk_work, int32_t* blk_off,
                             int32_t* scan_scratch, int64_t cap, int32_t* flags, void* stream) {
    PIC_CHECK_ARG(p && cell_count && cell_offset && blk_work && blk_off && scan_scratch && flags);
    for (int a = 0; a < 3; ++a) PIC_CHECK_ARG(p->tile[a] % TILE_B == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (p->tile[0] / TILE_B) * (p->tile[1] / TILE_B) * (p->tile[2] / TILE_B);
    const int warps = nblk + 1;
    k_sortb_blocks<<<(warps * 32 + 255) / 256, 256, 0, st>>>(nblk, cell_count, blk_work);
    const int rc = pic_sort_scan(nblk + 1, blk_work, blk_off, scan_scratch, stream);
    if (rc) return rc;
    k_sortb_cells<<<(warps * 32 + 255) / 256, 256, 0, st>>>(nblk, cell_count, blk_off, cell_offset, cap, flags);
    PIC_LAUNCH_RET();
}

// After pic_sort_scatter with the offsets above: mark the padding slots of every supercell dead and set the live slot count
// (= total padded length) on the device.
int pic_sort_blocked_finish(const PicParams* p, const int32_t* cell_offset, const int32_t* cell_count, const int32_t* blk_off,
                            const PicSoA* dst, void* stream) {
    PIC_CHECK_ARG(p && cell_offset && cell_count && blk_off && dst);
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (p->tile[0] / TILE_B) * (p->tile[1] / TILE_B) * (p->tile[2] / TILE_B);
    if (dst->n_dev) k_set_count_pair<<<1, 1, 0, st>>>(dst->n_dev, blk_off + nblk, dst->cap);
    PIC_DISPATCH_T(p, launch_sortb_pad, p, nblk, cell_offset, cell_count, blk_off, dst, st);
}

}  // extern "C"
