// hostcheck.cpp -- TEST-ONLY harness: compiles the __host__ __device__ particle bodies of pypic3d_b200/csrc/pic_slots.cuh
// for the CPU so the kernel arithmetic can be checked against the oracle in a container without a GPU.
// It is built by tests/test_hostcheck_math.py into tests/hostcheck/_build/ and is never loaded by the product package.
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
void hc_fused(const PicParams* p, int species, int dep, void* const comp[6], int64_t n, const void* const E[3],
              const void* const B[3], void* const J[3], void* leave, int64_t leave_cap, int32_t* leave_count, int32_t* flags) {
    HC_DISPATCH(p, t_fused, p, species, dep, comp, n, E, B, J, leave, leave_cap, leave_count, flags);
}
}
