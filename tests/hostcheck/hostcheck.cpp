// hostcheck.cpp -- TEST-ONLY harness: compiles the __host__ __device__ particle bodies of pypic3d_b200/csrc/pic_slots.cuh
// for the CPU so the kernel arithmetic can be checked against the oracle in a container without a GPU.
// It is built by tests/test_hostcheck_math.py into tests/hostcheck/_build/ and is never loaded by the product package.
#include <vector>
#include <cmath>
#include "../../pypic3d_b200/csrc/pic_slots.cuh"

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
void hc_fused(const PicParams* p, int species, int dep, void* const comp[6], int64_t n, const void* const E[3],
              const void* const B[3], void* const J[3], void* leave, int64_t leave_cap, int32_t* leave_count, int32_t* flags) {
    HC_DISPATCH(p, t_fused, p, species, dep, comp, n, E, B, J, leave, leave_cap, leave_count, flags);
}
}
