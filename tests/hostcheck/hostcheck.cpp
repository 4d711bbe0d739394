// hostcheck.cpp -- TEST-ONLY harness: compiles the __host__ __device__ particle bodies of pypic3d_b200/csrc/pic_slots.cuh
// for the CPU so the kernel arithmetic can be checked against the oracle in a container without a GPU.
// It is built by tests/test_hostcheck_math.py into tests/hostcheck/_build/ and is never loaded by the product package.
#include <vector>
#include <cmath>
#include "../../pypic3d_b200/csrc/pic_slots.cuh"
#include "../../pypic3d_b200/csrc/pic_pair.cuh"

using namespace pic;

template <typename T, int SF>
static void t_push(const PicParams* p, const void* x, const void* u_in, void* u_out, const uint8_t* active, int64_t cap,
                   const void* const E[3], const void* const B[3]) {
    Field6<T> F;
    for (int c = 0; c < 3; ++c) { F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c]; }
    const int64_t total = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2] * p->n_species * cap;
    for (int64_t i = 0; i < total; ++i) slot_push<T, SF>(*p, i, (const T*)x, (const T*)u_in, (T*)u_out, active, cap, F);
}

template <typename T, int SF>
static void t_deposit(const PicParams* p, int mode, const void* x, const void* u, const uint8_t* active, int64_t cap, void* const J[3]) {
    Field3W<T> F;
    for (int c = 0; c < 3; ++c) F.f[c] = (T*)J[c];
    const int64_t total = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2] * p->n_species * cap;
    for (int64_t i = 0; i < total; ++i) {
        if (mode == 0) slot_deposit<T, SF, 0>(*p, i, (const T*)x, (const T*)u, active, cap, F);
        else if (mode == 1) slot_deposit<T, SF, 1>(*p, i, (const T*)x, (const T*)u, active, cap, F);
        else slot_deposit<T, SF, 2>(*p, i, (const T*)x, (const T*)u, active, cap, F);
    }
}

template <typename T, int SF>
static void t_fused(const PicParams* p, int species, int dep, void* const comp[6], int64_t n, const void* const E[3],
                    const void* const B[3], void* const J[3], void* leave, int64_t leave_cap, int32_t* leave_count, int32_t* flags) {
    Field6<T> F, X;
    SoAView<T> s;
    for (int c = 0; c < 6; ++c) { s.c[c] = (T*)comp[c]; X.f[c] = nullptr; }
    s.id = nullptr; s.cap = n; s.n = n; s.n_dev = nullptr;
    for (int c = 0; c < 3; ++c) { F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c]; }
    Geom<T> gm;
    make_geom<T>(*p, 0, 0, 0, gm);
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = (T*)J[c]; sink.L[c] = gm.L[c]; }
    sink.off = 0;
    LeaveBuf lb = leave_of(nullptr); (void)leave; (void)leave_cap; (void)leave_count;
    bool distributed = false;
    for (int c = 0; c < 3; ++c) distributed |= (p->gmesh[c] != p->mesh[c]);
    const bool all3d = (p->gmesh[0] * p->tile[0] > 1) && (p->gmesh[1] * p->tile[1] > 1) && (p->gmesh[2] * p->tile[2] > 1) && p->g >= 2;
    for (int64_t i = 0; i < n; ++i) {
        if (dep == 0) {
            if (all3d) fused_particle<T, SF, 0, true>(*p, species, gm, i, s, F, X, 0, sink, lb, distributed, flags);
            else fused_particle<T, SF, 0, false>(*p, species, gm, i, s, F, X, 0, sink, lb, distributed, flags);
        } else {
            if (all3d) fused_particle<T, SF, 1, true>(*p, species, gm, i, s, F, X, 0, sink, lb, distributed, flags);
            else fused_particle<T, SF, 1, false>(*p, species, gm, i, s, F, X, 0, sink, lb, distributed, flags);
        }
    }
}

template <typename T, int SF>
static void t_fused3d(const PicParams* p, int species, void* const comp[6], int64_t n, const void* const E[3], const void* const B[3],
                      void* const J[3], void* leave, int64_t leave_cap, int32_t* leave_count, int32_t* flags) {
    Field6<T> F, X;
    SoAView<T> s;
    for (int c = 0; c < 6; ++c) { s.c[c] = (T*)comp[c]; X.f[c] = nullptr; }
    s.id = nullptr; s.cap = n; s.n = n; s.n_dev = nullptr;
    for (int c = 0; c < 3; ++c) { F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c]; }
    Geom<T> gm;
    make_geom<T>(*p, 0, 0, 0, gm);
    FastConst<T> k;
    make_fast_const<T>(*p, species, gm, k);
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = (T*)J[c]; sink.L[c] = gm.L[c]; }
    sink.off = 0;
    LeaveBuf lb = leave_of(nullptr); (void)leave; (void)leave_cap; (void)leave_count;
    bool distributed = false;
    for (int c = 0; c < 3; ++c) distributed |= (p->gmesh[c] != p->mesh[c]);
    for (int64_t i = 0; i < n; ++i) {
        if (p->pusher == PIC_PUSHER_BORIS) fused_particle_fast3d<T, SF, PIC_PUSHER_BORIS, false>(*p, species, gm, k, i, s, F, X, sink, lb, distributed, flags);
        else fused_particle_fast3d<T, SF, PIC_PUSHER_BORIS_REL, false>(*p, species, gm, k, i, s, F, X, sink, lb, distributed, flags);
    }
}

// Host emulation of K1 v9's gather source: for every particle, the 8x8x8-node tile of the supercell that contains it (or, with
// shift != 0, of a neighbouring supercell -- the particle then sits in the tile's margin or outside it and must take the
// global-memory fallback) is copied out of the global arrays and handed to fast3d_advance<TILE = true>.
template <typename T, int SF>
static void t_tile3d(const PicParams* p, int species, void* const comp[6], int64_t n, const void* const E[3], const void* const B[3],
                     void* const J[3], int shift, int32_t* flags) {
    if (SF != 1) return;
    Field6<T> F, X;
    SoAView<T> s;
    for (int c = 0; c < 6; ++c) { s.c[c] = (T*)comp[c]; X.f[c] = nullptr; }
    s.id = nullptr; s.cap = n; s.n = n; s.n_dev = nullptr;
    for (int c = 0; c < 3; ++c) { F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c]; }
    Geom<T> gm;
    make_geom<T>(*p, 0, 0, 0, gm);
    FastConst<T> k;
    make_fast_const<T>(*p, species, gm, k);
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = (T*)J[c]; sink.L[c] = gm.L[c]; }
    sink.off = 0;
    LeaveBuf lb = leave_of(nullptr);
    std::vector<T> tile(6 * TILE_ELEMS);
    // shift == 0: every particle sits inside its own supercell's tile, so the global arrays must never be read -- hand the
    // body NaN-filled ones to prove it
    std::vector<T> poison((size_t)gm.L[0] * gm.L[1] * gm.L[2], (T)NAN);
    Field6<T> Fbody = F;
    if (shift == 0)
        for (int c = 0; c < 6; ++c) Fbody.f[c] = poison.data();
    const double dd[3] = {p->dx, p->dy, p->dz};
    for (int64_t i = 0; i < n; ++i) {
        TileSrc<T> ts;
        ts.t = tile.data();
        bool dead = false;
        for (int a = 0; a < 3; ++a) {
            const double x = (double)s.c[a][i];
            if (x != x) { dead = true; ts.o[a] = 0; continue; }
            int cell = (int)std::floor((x + 0.5 * p->wind[a]) / dd[a]);
            cell = cell < 0 ? 0 : (cell > p->tile[a] - 1 ? p->tile[a] - 1 : cell);
            int blk = cell / TILE_B + (a == (int)(i % 3) ? shift : 0);
            const int nb = p->tile[a] / TILE_B;
            blk = blk < 0 ? 0 : (blk > nb - 1 ? nb - 1 : blk);
            ts.o[a] = blk * TILE_B + p->g - 2;
        }
        if (!dead)
            for (int c = 0; c < 6; ++c)
                for (int x = 0; x < TILE_N; ++x)
                    for (int y = 0; y < TILE_N; ++y)
                        for (int z = 0; z < TILE_N; ++z)
                            tile[c * TILE_ELEMS + x * TILE_SX + y * TILE_N + z] =
                                F.f[c][((size_t)(ts.o[0] + x) * gm.L[1] + (ts.o[1] + y)) * gm.L[2] + (ts.o[2] + z)];
        T po[3], xn[3], v[3], vals[SameCell<1>::NV];
        int key = 0, kind;
        if (p->pusher == PIC_PUSHER_BORIS) kind = fast3d_advance<T, 1, PIC_PUSHER_BORIS, false, true>(*p, species, k, i, s, Fbody, X, lb, false, flags, po, xn, v, key, vals, nullptr, &ts);
        else kind = fast3d_advance<T, 1, PIC_PUSHER_BORIS_REL, false, true>(*p, species, k, i, s, Fbody, X, lb, false, flags, po, xn, v, key, vals, nullptr, &ts);
        if (kind == 1) {
            int m = 0;
            for (int c = 0; c < 3; ++c)
                for (int f = 0; f < 1; ++f)
                    for (int m1 = 0; m1 < 2; ++m1)
                        for (int m2 = 0; m2 < 2; ++m2) sink.add_unchecked(sink.J[c] + key + SameCell<1>::offset(c, f, m1, m2, k.sx, k.sy), vals[m++]);
        } else if (kind == 2) {
            union_deposit<T, 1>(*p, species, gm, k, po, xn, v, sink);
        }
    }
}

// Host emulation of K1 v10 (kernels_pair.cu k_pair3d): particles are taken W at a time; the tile is the one of the supercell that
// contains the FIRST particle of the group (shift != 0: a neighbouring supercell), so the other one may or may not be covered by it.
// Every outcome of pair_advance is finished the way the kernel does: same-cell values are added, cell crossers go through
// crosser_finish, uncovered particles through the scalar global-memory body.  `edge` is set like the producer warp sets it.
template <typename T, int SF, int W>
static void t_pair3d_w(const PicParams* p, int species, void* const comp[6], int64_t n, const void* const E[3], const void* const B[3],
                       void* const J[3], int shift, int32_t* flags) {
    if (SF != 1) return;
    Field6<T> F, X;
    SoAView<T> s;
    for (int c = 0; c < 6; ++c) { s.c[c] = (T*)comp[c]; X.f[c] = nullptr; }
    s.id = nullptr; s.cap = n; s.n = n; s.n_dev = nullptr;
    for (int c = 0; c < 3; ++c) { F.f[c] = (const T*)E[c]; F.f[3 + c] = (const T*)B[c]; }
    Geom<T> gm;
    make_geom<T>(*p, 0, 0, 0, gm);
    FastConst<T> k;
    make_fast_const<T>(*p, species, gm, k);
    PairConst<T> pc;
    make_pair_const<T>(k, pc);
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = (T*)J[c]; sink.L[c] = gm.L[c]; }
    sink.off = 0;
    LeaveBuf lb = leave_of(nullptr);
    bool per1 = true;
    for (int a = 0; a < 3; ++a) per1 = per1 && (p->particle_bc[a] == PIC_BC_PERIODIC) && (p->gmesh[a] == p->mesh[a]);
    std::vector<T> tile(6 * TILE_ELEMS);
    const double dd[3] = {p->dx, p->dy, p->dz};
    for (int64_t i0 = 0; i0 < n; i0 += W) {
        int o[3] = {0, 0, 0};
        bool edge = false;
        {
            int64_t lead = -1;
            for (int j = 0; j < W && lead < 0; ++j)
                if (i0 + j < n && !pic_isnan(s.c[0][i0 + j])) lead = i0 + j;
            if (lead >= 0)
                for (int a = 0; a < 3; ++a) {
                    const double x = (double)s.c[a][lead];
                    int cell = (int)std::floor((x + 0.5 * p->wind[a]) / dd[a]);
                    cell = cell < 0 ? 0 : (cell > p->tile[a] - 1 ? p->tile[a] - 1 : cell);
                    int blk = cell / TILE_B + (a == (int)((i0 / W) % 3) ? shift : 0);
                    const int nb = p->tile[a] / TILE_B;
                    blk = blk < 0 ? 0 : (blk > nb - 1 ? nb - 1 : blk);
                    o[a] = blk * TILE_B + p->g - 2;
                    edge = edge || blk == 0 || blk == nb - 1;
                }
        }
        for (int c = 0; c < 6; ++c)
            for (int x = 0; x < TILE_N; ++x)
                for (int y = 0; y < TILE_NY; ++y)
                    for (int z = 0; z < TILE_N; ++z) {
                        const int gx = o[0] + x, gy = o[1] + y, gz = o[2] + z;
                        const bool in = gx < gm.L[0] && gy < gm.L[1] && gz < gm.L[2];       // TMA zero-fills outside the array
                        tile[c * TILE_ELEMS + x * TILE_SX + y * TILE_N + z] = in ? F.f[c][((size_t)gx * gm.L[1] + gy) * gm.L[2] + gz] : (T)0;
                    }
        const T x0[3] = {pic_fma((T)o[0], k.sc[0], k.oc[0]), pic_fma((T)o[1], k.sc[1], k.oc[1]), pic_fma((T)o[2], k.sc[2], k.oc[2])};
        const int key0 = o[0] * k.sx + o[1] * k.sy + o[2];
        Vec<T, W> pos[3], vel[3], pos_out[3], vel_out[3], xraw[3], vals[12];
        bool live[W];
        int kind[W], cid[W];
        for (int j = 0; j < W; ++j) {
            const bool in = i0 + j < n;
            for (int c = 0; c < 3; ++c) { pos[c].v[j] = in ? s.c[c][i0 + j] : pic_nan<T>(); vel[c].v[j] = in ? s.c[3 + c][i0 + j] : (T)0; }
            live[j] = in && !pic_isnan(pos[0].v[j]);
        }
        if (p->pusher == PIC_PUSHER_BORIS) {
            if (per1) pair_advance<T, W, PIC_PUSHER_BORIS, true>(k, pc, tile.data(), x0, edge, pos, vel, live, pos_out, vel_out, xraw, kind, cid, vals);
            else pair_advance<T, W, PIC_PUSHER_BORIS, false>(k, pc, tile.data(), x0, edge, pos, vel, live, pos_out, vel_out, xraw, kind, cid, vals);
        } else {
            if (per1) pair_advance<T, W, PIC_PUSHER_BORIS_REL, true>(k, pc, tile.data(), x0, edge, pos, vel, live, pos_out, vel_out, xraw, kind, cid, vals);
            else pair_advance<T, W, PIC_PUSHER_BORIS_REL, false>(k, pc, tile.data(), x0, edge, pos, vel, live, pos_out, vel_out, xraw, kind, cid, vals);
        }
        for (int j = 0; j < W; ++j) {
            const int64_t i = i0 + j;
            if (kind[j] == PAIR_SLOW) {
                flags[2] += 1;
                if (p->pusher == PIC_PUSHER_BORIS) fused_particle_fast3d<T, 1, PIC_PUSHER_BORIS, false>(*p, species, gm, k, i, s, F, X, sink, lb, false, flags);
                else fused_particle_fast3d<T, 1, PIC_PUSHER_BORIS_REL, false>(*p, species, gm, k, i, s, F, X, sink, lb, false, flags);
                continue;
            }
            if (kind[j] == PAIR_NONE) continue;
            for (int c = 0; c < 3; ++c) { s.c[c][i] = pos_out[c].v[j]; s.c[3 + c][i] = vel_out[c].v[j]; }
            if (kind[j] == PAIR_SAME) {
                const int base = key0 + (cid[j] >> 6) * k.sx + ((cid[j] >> 3) & 7) * k.sy + (cid[j] & 7);
                int m = 0;
                for (int c = 0; c < 3; ++c)
                    for (int m1 = 0; m1 < 2; ++m1)
                        for (int m2 = 0; m2 < 2; ++m2) sink.add_unchecked(sink.J[c] + base + SameCell<1>::offset(c, 0, m1, m2, k.sx, k.sy), vals[m++].v[j]);
            } else {
                const T po[3] = {pos[0].v[j], pos[1].v[j], pos[2].v[j]};
                const T xn[3] = {xraw[0].v[j], xraw[1].v[j], xraw[2].v[j]};
                if (per1) crosser_finish<T, true>(*p, species, gm, k, i, s, po, xn, sink, lb, false, flags);
                else crosser_finish<T, false>(*p, species, gm, k, i, s, po, xn, sink, lb, false, flags);
            }
        }
    }
}
template <typename T, int SF>
static void t_pair3d(const PicParams* p, int species, void* const comp[6], int64_t n, const void* const E[3], const void* const B[3],
                     void* const J[3], int W, int shift, int32_t* flags) {
    if (W == 2) t_pair3d_w<T, SF, 2>(p, species, comp, n, E, B, J, shift, flags);
    else t_pair3d_w<T, SF, 1>(p, species, comp, n, E, B, J, shift, flags);
}

#define HC_DISPATCH(p, FN, ...)                                                   \
    do {                                                                          \
        if ((p)->dtype == PIC_F32) {                                              \
            if ((p)->shape_factor == 1) FN<float, 1>(__VA_ARGS__); else FN<float, 2>(__VA_ARGS__);   \
        } else {                                                                  \
            if ((p)->shape_factor == 1) FN<double, 1>(__VA_ARGS__); else FN<double, 2>(__VA_ARGS__); \
        }                                                                         \
    } while (0)

extern "C" {
int hc_params_size(void) { return (int)sizeof(PicParams); }
void hc_push(const PicParams* p, const void* x, const void* u_in, void* u_out, const uint8_t* active, int64_t cap,
             const void* const E[3], const void* const B[3]) { HC_DISPATCH(p, t_push, p, x, u_in, u_out, active, cap, E, B); }
void hc_deposit(const PicParams* p, int mode, const void* x, const void* u, const uint8_t* active, int64_t cap, void* const J[3]) {
    HC_DISPATCH(p, t_deposit, p, mode, x, u, active, cap, J);
}
void hc_retile_classify(const PicParams* p, const void* x_in, const void* u_in, const uint8_t* a_in, void* x_out, void* u_out,
                        uint8_t* a_out, int64_t cap, int32_t* code, int32_t* overflow) {
    const int64_t total = (int64_t)p->mesh[0] * p->mesh[1] * p->mesh[2] * p->n_species * cap;
    for (int64_t i = 0; i < total; ++i) {
        if (p->dtype == PIC_F32) slot_retile_classify<float>(*p, i, (const float*)x_in, (const float*)u_in, a_in, (float*)x_out, (float*)u_out, a_out, cap, code, overflow);
        else slot_retile_classify<double>(*p, i, (const double*)x_in, (const double*)u_in, a_in, (double*)x_out, (double*)u_out, a_out, cap, code, overflow);
    }
}
void hc_fused3d(const PicParams* p, int species, void* const comp[6], int64_t n, const void* const E[3], const void* const B[3],
                void* const J[3], void* leave, int64_t leave_cap, int32_t* leave_count, int32_t* flags) {
    HC_DISPATCH(p, t_fused3d, p, species, comp, n, E, B, J, leave, leave_cap, leave_count, flags);
}
void hc_tile3d(const PicParams* p, int species, void* const comp[6], int64_t n, const void* const E[3], const void* const B[3],
               void* const J[3], int shift, int32_t* flags) {
    HC_DISPATCH(p, t_tile3d, p, species, comp, n, E, B, J, shift, flags);
}
void hc_pair3d(const PicParams* p, int species, void* const comp[6], int64_t n, const void* const E[3], const void* const B[3],
               void* const J[3], int W, int shift, int32_t* flags) {
    HC_DISPATCH(p, t_pair3d, p, species, comp, n, E, B, J, W, shift, flags);
}
void hc_fused(const PicParams* p, int species, int dep, void* const comp[6], int64_t n, const void* const E[3],
              const void* const B[3], void* const J[3], void* leave, int64_t leave_cap, int32_t* leave_count, int32_t* flags) {
    HC_DISPATCH(p, t_fused, p, species, dep, comp, n, E, B, J, leave, leave_cap, leave_count, flags);
}
// seam-consistent owner of a particle on a split periodic axis (pic_slots.cuh owner_offset_periodic), float32
int hc_owner_offset_f32(float* pos, float box_lo, float box_hi, float wind) { return owner_offset_periodic<float>(*pos, box_lo, box_hi, wind); }
}
