// pic_slots.cuh -- per-slot (reference layout) and per-particle (resident SoA) bodies of the particle kernels,
// written as __host__ __device__ functions: the __global__ kernels are thin grid-stride loops around them, and the
// test-only host-check harness (tests/hostcheck/) runs the very same bodies on the CPU in a GPU-less container.
#pragma once
#include "pic_math.cuh"

namespace pic {

template <typename T>
struct Field6 {
    const T* f[6];
};
template <typename T>
struct Field3W {
    T* f[3];
};

PIC_HD int atomic_add_i32(int32_t* p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    int o = *p; *p += v; return o;
#endif
}
PIC_HD void atomic_or_i32(int32_t* p, int v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
template <typename T>
PIC_HD T ld_ro(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

struct TileCoord {
    int tx, ty, tz;
};
PIC_HD TileCoord tile_coord(int64_t tile, const int mesh[3]) {
    TileCoord t;
    t.tz = (int)(tile % mesh[2]);
    t.ty = (int)((tile / mesh[2]) % mesh[1]);
    t.tx = (int)(tile / ((int64_t)mesh[2] * mesh[1]));
    return t;
}

PIC_HD void slot_decode(const PicParams& p, int64_t i, int64_t cap, int64_t& tile, int& s) {
    const int64_t per_tile = (int64_t)p.n_species * cap;
    tile = i / per_tile;
    s = (int)((i / cap) % p.n_species);
}
PIC_HD size_t tile_elems_of(const PicParams& p) {
    return (size_t)(p.tile[0] + 2 * p.g) * (p.tile[1] + 2 * p.g) * (p.tile[2] + 2 * p.g);
}

// ---------------------------------------------------------------- push (particle_push.py:45-144)
template <typename T, int SF>
PIC_HD void slot_push(const PicParams& p, int64_t i, const T* x, const T* u_in, T* u_out, const uint8_t* active, int64_t cap,
                      const Field6<T>& F) {
    T v[3] = {u_in[3 * i], u_in[3 * i + 1], u_in[3 * i + 2]};
    T out[3] = {v[0], v[1], v[2]};
    if (active[i]) {
        int64_t tile; int s;
        slot_decode(p, i, cap, tile, s);
        const TileCoord tc = tile_coord(tile, p.mesh);
        Geom<T> gm;
        make_geom<T>(p, tc.tx, tc.ty, tc.tz, gm);
        const T pos[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
        T EB[6];
        gather6<T, SF>(F.f, tile * tile_elems_of(p), gm, pos, EB);
        T nv[3];
        push_velocity<T>(p.pusher, v, EB, EB + 3, (T)p.charge[s], (T)p.mass[s], (T)p.dt, (T)p.C, nv);
        for (int c = 0; c < 3; ++c)
            if (p.update_u[s][c]) out[c] = nv[c];  // active & update_u (particle_push.py:134-142)
    }
    u_out[3 * i] = out[0]; u_out[3 * i + 1] = out[1]; u_out[3 * i + 2] = out[2];
}

// ---------------------------------------------------------------- deposits; MODE 0: Esirkepov, 1: direct J, 2: rho
template <typename T, int SF, int MODE>
PIC_HD void slot_deposit(const PicParams& p, int64_t i, const T* x, const T* u, const uint8_t* active, int64_t cap,
                         const Field3W<T>& J) {
    if (!active[i]) return;
    int64_t tile; int s;
    slot_decode(p, i, cap, tile, s);
    const TileCoord tc = tile_coord(tile, p.mesh);
    Geom<T> gm;
    make_geom<T>(p, tc.tx, tc.ty, tc.tz, gm);
    TileSink<T> sink;
    for (int c = 0; c < 3; ++c) { sink.J[c] = J.f[c]; sink.L[c] = gm.L[c]; }
    sink.off = tile * tile_elems_of(p);
    const T pos[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    const T qw = (T)(p.charge[s] * p.weight[s]);  // species_weighted_charge, Esirkepov.py:102
    if (MODE == 2) {
        const T v0[3] = {0, 0, 0};
        node_face_deposit<T, SF, false>(gm, pos, v0, qw, sink);
    } else {
        const T v[3] = {u[3 * i], u[3 * i + 1], u[3 * i + 2]};
        if (MODE == 1) {
            node_face_deposit<T, SF, true>(gm, pos, v, qw, sink);
        } else {
            T xn[3];
            for (int c = 0; c < 3; ++c) xn[c] = pos[c] + (p.update_x[s][c] ? v[c] * (T)p.dt : (T)0);  // Esirkepov.py:125-127
            esirkepov_deposit<T, SF>(gm, pos, xn, v, qw, (T)p.dt, sink);
        }
    }
}

// ---------------------------------------------------------------- move (particle_tile_communication.py:82-99)
template <typename T>
PIC_HD void slot_move(const PicParams& p, int64_t i, const T* x_in, T* x_out, const T* u, const uint8_t* active, int64_t cap, T dt) {
    const int s = (int)((i / cap) % p.n_species);
    const bool act = active[i] != 0;
    for (int c = 0; c < 3; ++c) {
        const T xi = x_in[3 * i + c];
        x_out[3 * i + c] = (act && p.update_x[s][c]) ? xi + u[3 * i + c] * dt : xi;
    }
}

// ---------------------------------------------------------------- retile classification (particle_tile_communication.py:233-320)
template <typename T>
PIC_HD bool bounded_state(const PicParams& p, T pos[3], T vel[3]) {
    bool alive = true;
    for (int c = 0; c < 3; ++c) alive = apply_axis_bc<T>(pos[c], vel[c], (T)p.wind[c], p.particle_bc[c]) && alive;
    return alive;
}

// code: 0 = empty/dropped, 1 = stays, 2 + stream = moves with stream index in the reference's loop order
// (ox, oy, oz each over (1, 0, -1); particle_tile_communication.py:326-331).
template <typename T>
PIC_HD void slot_retile_classify(const PicParams& p, int64_t i, const T* x_in, const T* u_in, const uint8_t* active_in, T* x_out,
                                 T* u_out, uint8_t* active_out, int64_t cap, int32_t* code, int32_t* overflow) {
    const T dd[3] = {(T)p.dx, (T)p.dy, (T)p.dz};
    int cd = 0;
    T pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0};
    if (active_in[i]) {
        for (int c = 0; c < 3; ++c) { pos[c] = x_in[3 * i + c]; vel[c] = u_in[3 * i + c]; }
        if (bounded_state<T>(p, pos, vel)) {
            int64_t tile; int s;
            slot_decode(p, i, cap, tile, s);
            const TileCoord tc = tile_coord(tile, p.mesh);
            const int src[3] = {tc.tx, tc.ty, tc.tz};
            int off[3];
            bool invalid = false, nonlocal_ = false;
            for (int c = 0; c < 3; ++c) {
                const int N = p.gmesh[c] * p.tile[c];
                const int dt_ = dest_tile<T>(pos[c], (T)p.wind[c], dd[c], N, p.tile[c], p.gmesh[c]);
                off[c] = adjacent_offset(dt_, src[c] + p.moff[c], p.gmesh[c]);
                invalid |= (off[c] > 1 || off[c] < -1);
                nonlocal_ |= (off[c] != 0);
            }
            if (invalid) { atomic_or_i32(overflow, 1); cd = 0; }
            else if (nonlocal_) cd = 2 + ((1 - off[0]) * 3 + (1 - off[1])) * 3 + (1 - off[2]);
            else cd = 1;
        }
    }
    code[i] = cd;
    const bool stay = (cd == 1);
    for (int c = 0; c < 3; ++c) {
        x_out[3 * i + c] = stay ? pos[c] : (T)0;  // non-staying slots are zeroed (:319-320)
        u_out[3 * i + c] = stay ? vel[c] : (T)0;
    }
    active_out[i] = stay ? 1 : 0;
}

// ================================================================ resident SoA path
template <typename T>
struct SoAView {
    T* c[6];
    int32_t* id;
    int64_t cap, n;
    int32_t* n_dev;
    PIC_HD int64_t count() const {   // slots in use: device counter when present (clamped to the capacity), else the host value
        if (!n_dev) return n;
        const int64_t m = *n_dev;
        return m < cap ? m : cap;
    }
};
template <typename T>
static inline SoAView<T> view_of(const PicSoA* s) {
    SoAView<T> v;
    for (int k = 0; k < 6; ++k) v.c[k] = (T*)s->comp[k];
    v.id = s->id;
    v.cap = s->cap;
    v.n = s->n;
    v.n_dev = s->n_dev;
    return v;
}

// Device view of PicLeave (include/pic_b200.h): per-direction packets with an int32 row count in the header row.
struct LeaveBuf {
    void* buf;
    int32_t row_off[27];
    int32_t cap[27];
    template <typename T>
    PIC_HD void push(int dir, const T pos[3], const T v[3], int species, int32_t* flags) const {
        T* base = (T*)buf + (int64_t)row_off[dir] * 7;
        const int slot = atomic_add_i32((int32_t*)base, 1);
        if (slot < cap[dir]) {
            T* pk = base + (int64_t)(1 + slot) * 7;
            for (int c = 0; c < 3; ++c) { pk[c] = pos[c]; pk[3 + c] = v[c]; }
            pk[6] = (T)species;
        } else {
            atomic_or_i32(flags, 2);
        }
    }
};
static inline LeaveBuf leave_of(const PicLeave* l) {
    LeaveBuf b;
    b.buf = l ? l->buf : nullptr;
    for (int d = 0; d < 27; ++d) { b.row_off[d] = l ? l->row_off[d] : 0; b.cap[d] = l ? l->cap[d] : 0; }
    return b;
}

// Sort key of a particle.  PIC_SORT_BLOCK > 1 orders cells in B x B x B blocks (block-major, row-major inside a block) so that a
// CTA's 256 consecutive particles span a compact 3-D neighbourhood and the E/B stencil is reused in all three directions from
// L1/L2; requires every tile width to be a multiple of B, otherwise plain row-major (z fastest) order is used.  Both orders keep
// all particles of one cell contiguous, which is all the deposit aggregation needs.
#ifndef PIC_SORT_BLOCK
#define PIC_SORT_BLOCK 4
#endif
template <typename T>
PIC_HD int local_cell(const PicParams& p, T px, T py, T pz) {
    if (pic_isnan(px)) return p.tile[0] * p.tile[1] * p.tile[2];  // dead -> trash bin at the end
    const T pos[3] = {px, py, pz};
    const T dd[3] = {(T)p.dx, (T)p.dy, (T)p.dz};
    int c[3];
    for (int a = 0; a < 3; ++a) {
        int cell = (int)pic_floor((pos[a] + (T)0.5 * (T)p.wind[a]) / dd[a]) - p.moff[a] * p.tile[a];
        c[a] = cell < 0 ? 0 : (cell > p.tile[a] - 1 ? p.tile[a] - 1 : cell);
    }
    constexpr int B = PIC_SORT_BLOCK;
    if (B > 1 && p.tile[0] % B == 0 && p.tile[1] % B == 0 && p.tile[2] % B == 0) {
        const int nby = p.tile[1] / B, nbz = p.tile[2] / B;
        const int blk = ((c[0] / B) * nby + (c[1] / B)) * nbz + (c[2] / B);
        return blk * (B * B * B) + ((c[0] % B) * B + (c[1] % B)) * B + (c[2] % B);
    }
    return (c[0] * p.tile[1] + c[1]) * p.tile[2] + c[2];
}

// Fast gather: identical arithmetic to gather6 but for the all-axes-active case with clamped (not wrapped) indices and
// zero-weight CIC points skipped; requires g >= 2 so owned particles never touch the array edge.
template <typename T, int SF>
PIC_HD void gather6_fast(const Field6<T>& F, const Geom<T>& gm, const T pos[3], T out[6]) {
    constexpr int K0 = (SF == 1) ? 1 : 0;
    int idx[2][3][3];
    T w[2][3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int gt = 0; gt < 2; ++gt) {
            int an;
            axis_stencil<T, SF>(pos[a], gt ? gm.ov[a] : gm.oc[a], gt ? gm.sv[a] : gm.sc[a], gm.d[a], an, w[gt][a]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int i = an - 1 + k;
                i = i < 0 ? 0 : (i > gm.L[a] - 1 ? gm.L[a] - 1 : i);
                idx[gt][a][k] = i;
            }
        }
    }
    const int GT[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const int gx = GT[c][0], gy = GT[c][1], gz = GT[c][2];
        const T* f = F.f[c];
        T acc = (T)0;
#pragma unroll
        for (int i = K0; i < 3; ++i) {
            T ai = (T)0;
#pragma unroll
            for (int j = K0; j < 3; ++j) {
                const size_t row = ((size_t)idx[gx][0][i] * gm.L[1] + idx[gy][1][j]) * gm.L[2];
                T aj = (T)0;
#pragma unroll
                for (int k = K0; k < 3; ++k) aj += ld_ro(f + row + idx[gz][2][k]) * w[gz][2][k];
                ai += aj * w[gy][1][j];
            }
            acc += ai * w[gx][0][i];
        }
        out[c] = acc;
    }
}

// K1 body: gather E,B (+ext) -> push -> deposit -> move -> particle BC (+ leaver extraction) for particle i.
// == evolve.py:33-79 for one particle of one local tile.  DEP 0 = Esirkepov, 1 = direct (centred).
template <typename T, int SF, int DEP, bool ALL3D>
PIC_HD void fused_particle(const PicParams& p, int species, const Geom<T>& gm, int64_t i, const SoAView<T>& s, const Field6<T>& F,
                           const Field6<T>& X, int has_ext, const TileSink<T>& sink, const LeaveBuf& leave, bool distributed,
                           int32_t* flags) {
    T pos[3] = {s.c[0][i], s.c[1][i], s.c[2][i]};
    if (pic_isnan(pos[0])) return;
    T v[3] = {s.c[3][i], s.c[4][i], s.c[5][i]};
    const T q = (T)p.charge[species], m = (T)p.mass[species];
    const T qw = (T)(p.charge[species] * p.weight[species]);
    const T dt = (T)p.dt, C = (T)p.C;
    // ---- gather (+ external fields, utils.py:205-216)
    T EB[6];
    if (ALL3D) gather6_fast<T, SF>(F, gm, pos, EB);
    else gather6<T, SF>(F.f, 0, gm, pos, EB);
    if (has_ext) {
        T EX[6];
        if (ALL3D) gather6_fast<T, SF>(X, gm, pos, EX);
        else gather6<T, SF>(X.f, 0, gm, pos, EX);
        for (int c = 0; c < 6; ++c) EB[c] += EX[c];
    }
    // ---- push
    T nv[3];
    push_velocity<T>(p.pusher, v, EB, EB + 3, q, m, dt, C, nv);
    for (int c = 0; c < 3; ++c)
        if (p.update_u[species][c]) v[c] = nv[c];
    // ---- deposit + move + global particle BCs
    bool alive = true;
    if (DEP == 0) {  // evolve.py:70-79
        T xn[3];
        for (int c = 0; c < 3; ++c) xn[c] = pos[c] + (p.update_x[species][c] ? v[c] * dt : (T)0);
        esirkepov_deposit<T, SF>(gm, pos, xn, v, qw, dt, sink);
        for (int c = 0; c < 3; ++c) { pos[c] = xn[c]; alive = apply_axis_bc<T>(pos[c], v[c], (T)p.wind[c], p.particle_bc[c]) && alive; }
    } else {         // evolve.py:46-67: half move, BC, deposit at the centred position, half move, BC
        const T hdt = dt / (T)2;
        for (int c = 0; c < 3; ++c) {
            if (p.update_x[species][c]) pos[c] = pos[c] + v[c] * hdt;
            alive = apply_axis_bc<T>(pos[c], v[c], (T)p.wind[c], p.particle_bc[c]) && alive;
        }
        if (alive) {
            node_face_deposit<T, SF, true>(gm, pos, v, qw, sink);
            for (int c = 0; c < 3; ++c) {
                if (p.update_x[species][c]) pos[c] = pos[c] + v[c] * hdt;
                alive = apply_axis_bc<T>(pos[c], v[c], (T)p.wind[c], p.particle_bc[c]) && alive;
            }
        }
    }
    // ---- ownership (distributed axes only): leavers go to the per-direction packet buffers
    if (alive && distributed) {
        const T dd[3] = {(T)p.dx, (T)p.dy, (T)p.dz};
        int off[3];
        bool invalid = false, nonlocal_ = false;
        for (int c = 0; c < 3; ++c) {
            const int N = p.gmesh[c] * p.tile[c];
            const int dtile = dest_tile<T>(pos[c], (T)p.wind[c], dd[c], N, p.tile[c], p.gmesh[c]);
            off[c] = adjacent_offset(dtile, p.moff[c], p.gmesh[c]);
            invalid |= (off[c] > 1 || off[c] < -1);
            nonlocal_ |= (off[c] != 0);
        }
        if (invalid) { atomic_or_i32(flags, 1); alive = false; }
        else if (nonlocal_) {
            const int dir = ((1 - off[0]) * 3 + (1 - off[1])) * 3 + (1 - off[2]);
            leave.push<T>(dir, pos, v, species, flags);
            alive = false;
        }
    }
    if (!alive) pos[0] = pic_nan<T>();
    for (int c = 0; c < 3; ++c) { s.c[c][i] = pos[c]; s.c[3 + c][i] = v[c]; }
}


// ================================================================ K1 v2: specialised 3-D Esirkepov body
// Same arithmetic as fused_particle<T,SF,0,true> (gather6_fast + push_velocity + esirkepov_deposit + move + BC), restructured
// for instruction count: reciprocal multiplies instead of IEEE divisions, the old-position stencil shared between the gather
// and the deposit, a branch-free union stencil of NS = SF+2 nodes per axis based at b = min(a_old, a_new) (- 1 for TSC), and
// the transverse Esirkepov factor written as T[j][k] = P_j * S1_k + Q_j * S0_k with P = S1/3 + S0/6, Q = S0/3 + S1/6.
// Requirements (checked by the launcher): all three axes active, g >= 2, periodic-or-any particle BCs, Esirkepov deposition.
template <typename T>
struct FastConst {
    T oc[3], sc[3], inv_sc[3], ov[3], sv[3], inv_sv[3], d[3], inv_d[3], wind[3];
    T dJ[3];          // -(q w / (d_a d_b)) / dt per axis
    T h;              // q dt / (2 m)
    T dt, C2, inv_C2;
    int L[3];
    int sx, sy;       // element strides of the x and y axes
    int pbc[3];
    int upd_x[3], upd_u[3];
    T box_lo[3], box_hi[3];   // local subdomain [lo, hi) on axes split across ranks (+-inf elsewhere)
    long long rowoff[3][3];   // byte offset of stencil row (a, b): (a * sx + b * sy) * sizeof(T)
};

template <typename T>
PIC_HD void make_fast_const(const PicParams& p, int species, const Geom<T>& gm, FastConst<T>& k) {
    for (int a = 0; a < 3; ++a) {
        k.oc[a] = gm.oc[a]; k.sc[a] = gm.sc[a]; k.inv_sc[a] = (T)1 / gm.sc[a];
        k.ov[a] = gm.ov[a]; k.sv[a] = gm.sv[a]; k.inv_sv[a] = (T)1 / gm.sv[a];
        k.d[a] = gm.d[a]; k.inv_d[a] = (T)1 / gm.d[a];
        k.wind[a] = (T)p.wind[a];
        k.L[a] = gm.L[a];
        k.pbc[a] = p.particle_bc[a];
        k.upd_x[a] = p.update_x[species][a];
        k.upd_u[a] = p.update_u[species][a];
    }
    for (int a = 0; a < 3; ++a) {
        const bool split = p.gmesh[a] != p.mesh[a];
        const double lo = -0.5 * p.wind[a] + (double)(p.moff[a] * p.tile[a]) * (a == 0 ? p.dx : (a == 1 ? p.dy : p.dz));
        const double hi = -0.5 * p.wind[a] + (double)((p.moff[a] + 1) * p.tile[a]) * (a == 0 ? p.dx : (a == 1 ? p.dy : p.dz));
        k.box_lo[a] = split ? (T)lo : (T)(-INFINITY);
        k.box_hi[a] = split ? (T)hi : (T)(INFINITY);
    }
    const T qw = (T)(p.charge[species] * p.weight[species]);
    const T dt = (T)p.dt;
    k.dJ[0] = -(qw / (gm.d[1] * gm.d[2])) / dt;
    k.dJ[1] = -(qw / (gm.d[2] * gm.d[0])) / dt;
    k.dJ[2] = -(qw / (gm.d[0] * gm.d[1])) / dt;
    k.h = (T)p.charge[species] * dt / ((T)2 * (T)p.mass[species]);
    k.dt = dt;
    k.C2 = (T)p.C * (T)p.C;
    k.inv_C2 = (T)1 / k.C2;
    k.sx = gm.L[1] * gm.L[2];
    k.sy = gm.L[2];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) k.rowoff[a][b] = ((long long)a * k.sx + (long long)b * k.sy) * (long long)sizeof(T);
}

// 1/sqrt(x) and a/b for the push: f32 on the device uses the SFU approximations (<= 2 ulp; the f32 parity tolerance is
// 2e-5), f64 and the host keep IEEE operations.
#ifndef PIC_FASTMATH
#define PIC_FASTMATH 0   /* measured slower on B200 (profiles/r01_k1_versions.md): the kernel is latency-, not issue-bound */
#endif
#ifndef PIC_GATHER_F32X2
#define PIC_GATHER_F32X2 1   /* tile gather (K1 v9): packed f32x2 arithmetic for the z / y interpolation of the two x planes */
#endif
#ifndef PIC_GATHER_V
#define PIC_GATHER_V 2   /* 2: 64-bit row offsets from the constant bank (4.88 ms); 0: pointer + int row (5.15 ms); 1: u32 offsets (slower) */
#endif
PIC_HD float pic_rsqrt(float x) {
#if defined(__CUDA_ARCH__) && PIC_FASTMATH
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
PIC_HD double pic_rsqrt(double x) { return 1.0 / sqrt(x); }
PIC_HD float pic_fdiv(float a, float b) {
#if defined(__CUDA_ARCH__) && PIC_FASTMATH
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
PIC_HD double pic_fdiv(double a, double b) { return a / b; }

// anchor + the three shape weights with reciprocal multiplies (<= 1 ulp from axis_stencil)
template <typename T, int SF>
PIC_HD void axis_stencil_rcp(T pos, T o, T s, T inv_s, T inv_d, int& a, T w[3]) {
    const T q = (pos - o) * inv_s;
    const T fa = (SF == 1) ? pic_floor(q) : pic_rint(q);
    a = (int)fa;
    const T r = (pos - (fa * s + o)) * inv_d;
    if (SF == 1) {
        w[0] = (T)0; w[1] = (T)1 - r; w[2] = r;
    } else {
        const T hm = (T)0.5 - r, hp = (T)0.5 + r;
        w[0] = (T)0.5 * hm * hm; w[1] = (T)0.75 - r * r; w[2] = (T)0.5 * hp * hp;
    }
}

#ifndef PIC_WRAP_V
#define PIC_WRAP_V 1   /* measured 5.15 vs 5.31 ms per K1 launch */
#endif
template <typename T>
PIC_HD T wrap_periodic_fast(T x, T wind) {
    // bit-identical to wrap_periodic for -wind <= x + h < 2 wind (fmod is exact there); general fallback otherwise
    const T h = (T)0.5 * wind;
    T t = x + h;
#if PIC_WRAP_V == 1
    // one rarely-taken branch: interior particles only pay (x + h) - h, exactly what the reference's mod() leaves them with
    if (t >= wind || t < (T)0) {
        if (t >= (T)2 * wind || t < -wind) return wrap_periodic(x, wind);
        t = (t >= wind) ? t - wind : t + wind;     // (a tiny negative t rounds t + wind to wind: mod() leaves +h there too)
        const T w = t - h;
        return (w == -h && x >= h) ? h : w;
    }
    return t - h;
#else
    if (t >= (T)2 * wind || t < -wind) return wrap_periodic(x, wind);   // never taken for |v| dt < wind
    const T up = t - wind, dn = t + wind;
    t = (t >= wind) ? up : ((t < (T)0) ? ((dn >= wind) ? dn - wind : dn) : t);
    T w = t - h;
    w = (w == -h && x >= h) ? h : w;
    return w;
#endif
}

// Owner of a particle on a periodic axis that is split across ranks: the physical crossing direction of the local box [lo, hi),
// taken before the wrap (first -> last tile is -1, last -> first is +1, particle_tile_communication.py:145-165), made CONSISTENT
// with what the wrap returns at the seam of the global domain.  Two round-off cases would otherwise strand a particle on a rank
// whose tile does not contain it (one electron per step at 128^3 cells per rank in float32, visible as a charge-conservation
// violation in a single node): (i) x + h rounds up to `wind` although x < h: the wrap lands on -h, so the particle has crossed to
// the +1 neighbour; (ii) x == +h exactly: the wrap keeps the reference's +h alias (grid_and_stencil.py:31-35), which only the top
// rank can hold -- a particle handed across the seam travels as -h.
template <typename T>
PIC_HD int owner_offset_periodic(T& pos, T box_lo, T box_hi, T wind) {
    const T raw = pos, half = (T)0.5 * wind;
    int off = (raw >= box_hi) ? 1 : ((raw < box_lo) ? -1 : 0);
    T w = wrap_periodic_fast<T>(raw, wind);
    if (box_hi < (T)INFINITY) {          // (an axis that is not split has box = (-inf, inf): the wrap stays on this rank)
        if (off == 0) {
            if (w < raw - half) off = 1;
            else if (w > raw + half) off = -1;
        }
        if (off > 0 && w >= half) w = -half;
    }
    pos = w;
    return off;
}

// Esirkepov deposit on the union stencil of NS = SF+2 nodes per axis based at min(a_old, a_new) (- 1 for TSC): handles any
// anchor shift in {-1, 0, +1}; anything else (or a stencil leaving the ghosted tile) goes to the bounds-checked general body.
template <typename T, int SF>
PIC_HD void union_deposit(const PicParams& p, int species, const Geom<T>& gm, const FastConst<T>& k, const T pos[3], const T xn[3],
                          const T v[3], const TileSink<T>& sink) {
    constexpr int NS = SF + 2;
    constexpr int K0 = (SF == 1) ? 1 : 0;
    T S0[3][NS], S1[3][NS];
    int b[3];
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int an, ao;
        T wn[3], wo[3];
        axis_stencil_rcp<T, SF>(xn[a], k.oc[a], k.sc[a], k.inv_sc[a], k.inv_d[a], an, wn);
        axis_stencil_rcp<T, SF>(pos[a], k.oc[a], k.sc[a], k.inv_sc[a], k.inv_d[a], ao, wo);
        const int sh = an - ao;
        ok = ok && (sh >= -1) && (sh <= 1);
        const int lo = (sh < 0 ? an : ao) - (SF == 1 ? 0 : 1);   // first node of the union stencil
        b[a] = lo;
        ok = ok && (lo >= 0) && (lo + NS - 1 < k.L[a]);
        const bool old_first = (sh >= 0);   // old stencil starts at the union base
        const bool new_first = (sh <= 0);
        // CIC: weights (w[1], w[2]) at nodes (a, a+1); TSC: (w[0], w[1], w[2]) at (a-1, a, a+1)
#pragma unroll
        for (int m = 0; m < NS; ++m) {
            const int k0 = m + K0, k1 = m + K0 - 1;     // weight index if the stencil starts at the base / one node later
            const T o_a = (k0 <= 2) ? wo[k0] : (T)0;
            const T o_b = (k1 >= K0 && k1 <= 2) ? wo[k1] : (T)0;
            const T n_a = (k0 <= 2) ? wn[k0] : (T)0;
            const T n_b = (k1 >= K0 && k1 <= 2) ? wn[k1] : (T)0;
            S0[a][m] = old_first ? o_a : o_b;
            S1[a][m] = new_first ? n_a : n_b;
        }
    }
    if (!ok) {
        const T qw = (T)(p.charge[species] * p.weight[species]);
        esirkepov_deposit<T, SF>(gm, pos, xn, v, qw, k.dt, sink);
        return;
    }
    const T third = (T)(1.0 / 3.0), sixth = (T)(1.0 / 6.0);
    T P[2][NS], Q[2][NS];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int m = 0; m < NS; ++m) {
            P[a][m] = third * S1[a][m] + sixth * S0[a][m];
            Q[a][m] = third * S0[a][m] + sixth * S1[a][m];
        }
    const int base = b[0] * k.sx + b[1] * k.sy + b[2];
    {   // Jx: cumsum over i of (S1x - S0x); T_yz[j][k] = P_y[j] S1z[k] + Q_y[j] S0z[k]
        T* J = sink.J[0] + base;
        T cum = (T)0;
#pragma unroll
        for (int ii = 0; ii < NS - 1; ++ii) {
            cum += S1[0][ii] - S0[0][ii];
            const T fc = k.dJ[0] * cum;
#pragma unroll
            for (int jj = 0; jj < NS; ++jj)
#pragma unroll
                for (int kk = 0; kk < NS; ++kk) {
                    const T val = fc * (P[1][jj] * S1[2][kk] + Q[1][jj] * S0[2][kk]);
                    if (val != (T)0) sink.add_unchecked(J + ii * k.sx + jj * k.sy + kk, val);
                }
        }
    }
    {   // Jy: cumsum over j; T_xz[i][k] = P_x[i] S1z[k] + Q_x[i] S0z[k]
        T* J = sink.J[1] + base;
        T cum = (T)0;
#pragma unroll
        for (int jj = 0; jj < NS - 1; ++jj) {
            cum += S1[1][jj] - S0[1][jj];
            const T fc = k.dJ[1] * cum;
#pragma unroll
            for (int ii = 0; ii < NS; ++ii)
#pragma unroll
                for (int kk = 0; kk < NS; ++kk) {
                    const T val = fc * (P[0][ii] * S1[2][kk] + Q[0][ii] * S0[2][kk]);
                    if (val != (T)0) sink.add_unchecked(J + ii * k.sx + jj * k.sy + kk, val);
                }
        }
    }
    {   // Jz: cumsum over k; T_xy[i][j] = P_x[i] S1y[j] + Q_x[i] S0y[j]
        T* J = sink.J[2] + base;
        T cum = (T)0;
#pragma unroll
        for (int kk = 0; kk < NS - 1; ++kk) {
            cum += S1[2][kk] - S0[2][kk];
            const T fc = k.dJ[2] * cum;
#pragma unroll
            for (int ii = 0; ii < NS; ++ii)
#pragma unroll
                for (int jj = 0; jj < NS; ++jj) {
                    const T val = fc * (P[0][ii] * S1[1][jj] + Q[0][ii] * S0[1][jj]);
                    if (val != (T)0) sink.add_unchecked(J + ii * k.sx + jj * k.sy + kk, val);
                }
        }
    }
}

// Number of per-particle current values of the same-cell (no anchor shift) Esirkepov stencil: 3 components x (NN-1) faces x
// NN x NN transverse nodes, NN = SF + 1 nodes per axis.  CIC: 12, TSC: 54.
template <int SF>
struct SameCell {
    static constexpr int NN = SF + 1;
    static constexpr int NV = 3 * (NN - 1) * NN * NN;
    // element offset of value (c, f, m1, m2) relative to the stencil base node
    PIC_HD static int offset(int c, int f, int m1, int m2, int sx, int sy) {
        return c == 0 ? f * sx + m1 * sy + m2 : (c == 1 ? m1 * sx + f * sy + m2 : m1 * sx + m2 * sy + f);
    }
};

// Global-memory gather: Ex(v,c,c) Ey(c,v,c) Ez(c,c,v) Bx(c,v,v) By(v,c,v) Bz(v,v,c).
// One 64-bit base pointer per component; the 3x3 stencil rows are reached by adding 64-bit byte offsets that live in the
// constant bank (2 integer instructions per row), the z neighbours by immediate offsets.
template <typename T, int SF, bool HAS_EXT>
PIC_HD void gather_rows(const FastConst<T>& k, const Field6<T>& F, const Field6<T>& X, const int ac[3], const int av[3],
                        const T wc[3][3], const T wv[3][3], T EB[6]) {
    constexpr int K0 = (SF == 1) ? 1 : 0;
        int bc_[3], bv_[3];   // clamped first stencil index (memory safety only; owned particles never clamp for g >= 2)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int hi = k.L[a] - 3;
            int c0 = ac[a] - 1, v0 = av[a] - 1;
            c0 = c0 < hi ? c0 : hi; c0 = c0 > 0 ? c0 : 0;
            v0 = v0 < hi ? v0 : hi; v0 = v0 > 0 ? v0 : 0;
            bc_[a] = c0; bv_[a] = v0;
        }
        const int GT[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int gx = GT[c][0], gy = GT[c][1], gz = GT[c][2];
            const T* wx = gx ? wv[0] : wc[0];
            const T* wy = gy ? wv[1] : wc[1];
            const T* wz = gz ? wv[2] : wc[2];
            const int base = ((gx ? bv_[0] : bc_[0]) * k.sx) + ((gy ? bv_[1] : bc_[1]) * k.sy) + (gz ? bv_[2] : bc_[2]);
            const char* f = (const char*)(F.f[c] + base);
            const char* fx = HAS_EXT ? (const char*)(X.f[c] + base) : nullptr;
            T acc = (T)0;
#pragma unroll
            for (int a_ = K0; a_ < 3; ++a_) {
                T ai = (T)0;
#pragma unroll
                for (int b_ = K0; b_ < 3; ++b_) {
                    const T* r = (const T*)(f + k.rowoff[a_][b_]);
                    const T* rx = HAS_EXT ? (const T*)(fx + k.rowoff[a_][b_]) : nullptr;
                    T aj = (T)0;
#pragma unroll
                    for (int c_ = K0; c_ < 3; ++c_) {
                        T val = ld_ro(r + c_);
                        if (HAS_EXT) val += ld_ro(rx + c_);
                        aj += val * wz[c_];
                    }
                    ai += aj * wy[b_];
                }
                acc += ai * wx[a_];
            }
            EB[c] = acc;
        }
}

// K1 v9: shared-memory E/B tile of one supercell (TILE_B^3 cells of the blocked sort order).  TILE_N = 8 nodes per axis starting
// at array index o[a] = (first cell of the supercell) + g - 2 cover both Yee lines of every CIC particle that sits in the supercell
// or at most one cell outside it: centre anchors o+1 .. o+6 and vertex anchors o .. o+6, each reading nodes (a, a+1).
constexpr int TILE_B = 4;
constexpr int TILE_N = 8;
#ifndef PIC_TILE_YPAD
#define PIC_TILE_YPAD 1   /* spare y rows per x plane.  1: plane stride 72 words, so the centre / vertex anchors of one cell (one plane
                             apart) fall into different banks.  2 (stride 80, also conflict-free for a warp that straddles a y step)
                             was measured: 4.209 vs 4.213 ms per launch -- no difference, so the smaller tile stays. */
#endif
constexpr int TILE_NY = TILE_N + PIC_TILE_YPAD;
constexpr int TILE_SX = TILE_NY * TILE_N;                // x-plane stride (the tile is the dense TMA box [x 8][y TILE_NY][z 8])
constexpr int TILE_ELEMS = TILE_N * TILE_SX;             // per component
template <typename T>
struct TileSrc {
    const T* t;      // [6][TILE_N][TILE_NY][TILE_N], z fastest -- same component order as Field6
    int o[3];        // array index of the tile's first node on each axis
};

// K1 v3 per-particle body: load -> stencils -> gather -> push -> new position -> store (move + BC + ownership).
// Returns kind: 0 = nothing to deposit (dead slot), 1 = same-cell deposit: `vals` (SameCell<SF>::NV values) to be added at
// J_c[key + offset(c, f, m1, m2)] -- these are what the kernel reduces across lanes of the same cell before the RED;
// 2 = the particle changed anchor on some axis (or its stencil leaves the tile): deposit through union_deposit(old, new).
// TILE = true (K1 v9, CIC only): E and B are read from the supercell tile `ts` instead of global memory; a particle whose stencil
// is not covered by the tile (it drifted more than one cell out of its supercell since the last sort) gathers from global
// memory with the same arithmetic, so the result never depends on how stale the sort is.
// PERIODIC1 = true: all three particle BCs periodic on a single rank (the launcher checks) -- the move is just the wrap.
template <typename T, int SF, int PUSHER, bool HAS_EXT, bool TILE = false, bool PERIODIC1 = false>
PIC_HD int fast3d_advance(const PicParams& p, int species, const FastConst<T>& k, int64_t i, const SoAView<T>& s, const Field6<T>& F,
                          const Field6<T>& X, const LeaveBuf& leave, bool distributed, int32_t* flags, T pos_old[3], T xn[3], T vout[3],
                          int& key, T* vals, const T* pre = nullptr, const TileSrc<T>* ts = nullptr, int* srel = nullptr) {
    static_assert(!TILE || (SF == 1 && !HAS_EXT), "the tile gather is built for CIC without external fields");
    // srel (TILE only, may be null): element offset of the particle's first deposit node inside an 8x8x8 shared-memory J tile
    // with the E/B tile's origin, or -1 when the particle's stencil is not covered by the tile
    int srel_ = -1;
    constexpr int NN = SF + 1;
    constexpr int K0 = (SF == 1) ? 1 : 0;
    // `pre`: x,y,z,vx,vy,vz of particle i already loaded by the caller (software prefetch of the next iteration)
    T pos[3] = {pre ? pre[0] : s.c[0][i], pre ? pre[1] : s.c[1][i], pre ? pre[2] : s.c[2][i]};
    if (pic_isnan(pos[0])) return 0;
    T v[3] = {pre ? pre[3] : s.c[3][i], pre ? pre[4] : s.c[4][i], pre ? pre[5] : s.c[5][i]};
    // ---- stencils of the old position on the center and vertex lines
    int ac[3], av[3];
    T wc[3][3], wv[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        axis_stencil_rcp<T, SF>(pos[a], k.oc[a], k.sc[a], k.inv_sc[a], k.inv_d[a], ac[a], wc[a]);
        axis_stencil_rcp<T, SF>(pos[a], k.ov[a], k.sv[a], k.inv_sv[a], k.inv_d[a], av[a], wv[a]);
    }
    T EB[6];
    if (TILE) {
        // ---- gather from the supercell tile: one 32-bit element offset per component, the 8 corners by constant offsets
        int rc[3], rv[3];
        bool in_tile = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            rc[a] = ac[a] - ts->o[a];
            rv[a] = av[a] - ts->o[a];
            in_tile = in_tile && ((unsigned)rc[a] <= (unsigned)(TILE_N - 2)) && ((unsigned)rv[a] <= (unsigned)(TILE_N - 2));
        }
        if (in_tile) srel_ = (rc[0] * TILE_N + rc[1]) * TILE_N + rc[2];
        if (!in_tile) {      // drifted out of the tile's one-cell margin (rare): same arithmetic from global memory
            gather_rows<T, SF, HAS_EXT>(k, F, X, ac, av, wc, wv, EB);
            atomic_add_i32(flags + 2, 1);    // diagnostic: flags[2] counts the particles that took this path
        } else {
        const int GT[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int gx = GT[c][0], gy = GT[c][1], gz = GT[c][2];
            const T* wx = gx ? wv[0] : wc[0];
            const T* wy = gy ? wv[1] : wc[1];
            const T* wz = gz ? wv[2] : wc[2];
            const T* f = ts->t + c * TILE_ELEMS + ((gx ? rv[0] : rc[0]) * TILE_SX + (gy ? rv[1] : rc[1]) * TILE_N + (gz ? rv[2] : rc[2]));
#if defined(__CUDA_ARCH__) && PIC_GATHER_F32X2
            if constexpr (sizeof(T) == 4) {
                // Blackwell packed FP32: the two x planes of the stencil travel through the z and y interpolation side by side in
                // one FMUL2 / FFMA2 each (8 floating-point instructions per component instead of 14; same roundings as below)
                const float2 z0y0 = make_float2(f[0], f[TILE_SX]), z1y0 = make_float2(f[1], f[TILE_SX + 1]);
                const float2 z0y1 = make_float2(f[TILE_N], f[TILE_SX + TILE_N]), z1y1 = make_float2(f[TILE_N + 1], f[TILE_SX + TILE_N + 1]);
                const float2 wz1 = make_float2(wz[1], wz[1]), wz2 = make_float2(wz[2], wz[2]);
                const float2 wy1 = make_float2(wy[1], wy[1]), wy2 = make_float2(wy[2], wy[2]);
                const float2 ay0 = __ffma2_rn(z1y0, wz2, __fmul2_rn(z0y0, wz1));
                const float2 ay1 = __ffma2_rn(z1y1, wz2, __fmul2_rn(z0y1, wz1));
                const float2 ax = __ffma2_rn(ay1, wy2, __fmul2_rn(ay0, wy1));
                EB[c] = ax.x * wx[1] + ax.y * wx[2];
                continue;
            }
#endif
            T acc = (T)0;
#pragma unroll
            for (int a_ = 0; a_ < 2; ++a_) {
                T ai = (T)0;
#pragma unroll
                for (int b_ = 0; b_ < 2; ++b_) {
                    const T* r = f + a_ * TILE_SX + b_ * TILE_N;
                    const T aj = r[0] * wz[1] + r[1] * wz[2];
                    ai += aj * wy[1 + b_];
                }
                acc += ai * wx[1 + a_];
            }
            EB[c] = acc;
        }
        }
    } else
#if PIC_GATHER_V == 2
    { gather_rows<T, SF, HAS_EXT>(k, F, X, ac, av, wc, wv, EB); }
#elif PIC_GATHER_V == 1
    // ---- gather: Ex(v,c,c) Ey(c,v,c) Ez(c,c,v) Bx(c,v,v) By(v,c,v) Bz(v,v,c)
    // Offsets are unsigned 32-bit element indices from the component base pointers (which live in the constant bank), so
    // each load is one "uniform base + 32-bit offset" LDG instead of a 64-bit address computation per row.
    {
        unsigned oc_[3], ov_[3];   // clamped first stencil index (memory safety only; owned particles never clamp for g >= 2)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int c0 = ac[a] - 1, v0 = av[a] - 1;
            c0 = c0 < 0 ? 0 : (c0 > k.L[a] - 3 ? k.L[a] - 3 : c0);
            v0 = v0 < 0 ? 0 : (v0 > k.L[a] - 3 ? k.L[a] - 3 : v0);
            const unsigned stride = a == 0 ? (unsigned)k.sx : (a == 1 ? (unsigned)k.sy : 1u);
            oc_[a] = (unsigned)c0 * stride;
            ov_[a] = (unsigned)v0 * stride;
        }
        const int GT[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int gx = GT[c][0], gy = GT[c][1], gz = GT[c][2];
            const T* wx = gx ? wv[0] : wc[0];
            const T* wy = gy ? wv[1] : wc[1];
            const T* wz = gz ? wv[2] : wc[2];
            const unsigned base = (gx ? ov_[0] : oc_[0]) + (gy ? ov_[1] : oc_[1]) + (gz ? ov_[2] : oc_[2]);
            const T* f = F.f[c];
            const T* fx = HAS_EXT ? X.f[c] : nullptr;
            T acc = (T)0;
#pragma unroll
            for (int a_ = K0; a_ < 3; ++a_) {
                T ai = (T)0;
#pragma unroll
                for (int b_ = K0; b_ < 3; ++b_) {
                    const unsigned row = base + (unsigned)a_ * (unsigned)k.sx + (unsigned)b_ * (unsigned)k.sy;
                    T aj = (T)0;
#pragma unroll
                    for (int c_ = K0; c_ < 3; ++c_) {
                        T val = ld_ro(f + (row + (unsigned)c_));
                        if (HAS_EXT) val += ld_ro(fx + (row + (unsigned)c_));
                        aj += val * wz[c_];
                    }
                    ai += aj * wy[b_];
                }
                acc += ai * wx[a_];
            }
            EB[c] = acc;
        }
    }
#else
    // ---- gather: Ex(v,c,c) Ey(c,v,c) Ez(c,c,v) Bx(c,v,v) By(v,c,v) Bz(v,v,c)
#if defined(PIC_ABLATE) && PIC_ABLATE == 2   /* profiling build: no gather */
    for (int c = 0; c < 6; ++c) EB[c] = (T)(ac[c % 3] + av[c % 3]) * wc[0][1];
    if (false)
#endif
    {
        int bc_[3], bv_[3];   // clamped first stencil index (memory safety only; owned particles never clamp for g >= 2)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int c0 = ac[a] - 1, v0 = av[a] - 1;
            bc_[a] = c0 < 0 ? 0 : (c0 > k.L[a] - 3 ? k.L[a] - 3 : c0);
            bv_[a] = v0 < 0 ? 0 : (v0 > k.L[a] - 3 ? k.L[a] - 3 : v0);
        }
        const int GT[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int gx = GT[c][0], gy = GT[c][1], gz = GT[c][2];
            const T* wx = gx ? wv[0] : wc[0];
            const T* wy = gy ? wv[1] : wc[1];
            const T* wz = gz ? wv[2] : wc[2];
            const int base = ((gx ? bv_[0] : bc_[0]) * k.sx) + ((gy ? bv_[1] : bc_[1]) * k.sy) + (gz ? bv_[2] : bc_[2]);
            const T* f = F.f[c] + base;
            const T* fx = HAS_EXT ? X.f[c] + base : nullptr;
            T acc = (T)0;
#pragma unroll
            for (int a_ = K0; a_ < 3; ++a_) {
                T ai = (T)0;
#pragma unroll
                for (int b_ = K0; b_ < 3; ++b_) {
                    const int row = a_ * k.sx + b_ * k.sy;
                    T aj = (T)0;
#pragma unroll
                    for (int c_ = K0; c_ < 3; ++c_) {
                        T val = ld_ro(f + row + c_);
                        if (HAS_EXT) val += ld_ro(fx + row + c_);
                        aj += val * wz[c_];
                    }
                    ai += aj * wy[b_];
                }
                acc += ai * wx[a_];
            }
            EB[c] = acc;
        }
    }
#endif
    // ---- push (same formulas as push_velocity; reciprocals hoisted)
    {
        const T h = k.h;
        T um[3], t[3], cr[3], up[3], nu[3];
        if (PUSHER == PIC_PUSHER_BORIS) {
            for (int c = 0; c < 3; ++c) { um[c] = v[c] + h * EB[c]; t[c] = h * EB[3 + c]; }
            cross3(um, t, cr);
            for (int c = 0; c < 3; ++c) up[c] = um[c] + cr[c];
            const T f2 = pic_fdiv((T)2, (T)1 + t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
            T sv_[3] = {t[0] * f2, t[1] * f2, t[2] * f2};
            cross3(up, sv_, cr);
            for (int c = 0; c < 3; ++c) nu[c] = (um[c] + cr[c]) + h * EB[c];
        } else if (PUSHER == PIC_PUSHER_BORIS_REL) {
            const T gamma = pic_rsqrt((T)1 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * k.inv_C2);
            for (int c = 0; c < 3; ++c) um[c] = v[c] * gamma + h * EB[c];
            const T inv_gm = pic_rsqrt((T)1 + (um[0] * um[0] + um[1] * um[1] + um[2] * um[2]) * k.inv_C2);
            for (int c = 0; c < 3; ++c) t[c] = h * EB[3 + c] * inv_gm;
            cross3(um, t, cr);
            for (int c = 0; c < 3; ++c) up[c] = um[c] + cr[c];
            const T f2 = pic_fdiv((T)2, (T)1 + t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
            T sv_[3] = {t[0] * f2, t[1] * f2, t[2] * f2};
            cross3(up, sv_, cr);
            for (int c = 0; c < 3; ++c) nu[c] = (um[c] + cr[c]) + h * EB[c];
            const T inv_ng = pic_rsqrt((T)1 + (nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]) * k.inv_C2);
            for (int c = 0; c < 3; ++c) nu[c] *= inv_ng;
        } else {
            push_velocity<T>(PIC_PUSHER_HC, v, EB, EB + 3, (T)p.charge[species], (T)p.mass[species], k.dt, (T)p.C, nu);
        }
        for (int c = 0; c < 3; ++c)
            if (k.upd_u[c]) v[c] = nu[c];
    }
    // ---- new position and its stencil; same-cell test
    bool same = true;
    T wn[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        xn[a] = pos[a] + (k.upd_x[a] ? v[a] * k.dt : (T)0);
        int an;
        axis_stencil_rcp<T, SF>(xn[a], k.oc[a], k.sc[a], k.inv_sc[a], k.inv_d[a], an, wn[a]);
        const int lo = ac[a] - (SF == 1 ? 0 : 1);
        same = same && (an == ac[a]) && (lo >= 0) && (lo + NN - 1 < k.L[a]);
        pos_old[a] = pos[a];
        vout[a] = v[a];
    }
    int kind = 2;
    if (same) {
        kind = 1;
        if (srel) *srel = srel_;
        key = (ac[0] - (SF == 1 ? 0 : 1)) * k.sx + (ac[1] - (SF == 1 ? 0 : 1)) * k.sy + (ac[2] - (SF == 1 ? 0 : 1));
        const T third = (T)(1.0 / 3.0), sixth = (T)(1.0 / 6.0);
        T P[2][NN], Q[2][NN], cum[3][NN - 1];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            T run = (T)0;
#pragma unroll
            for (int m = 0; m < NN - 1; ++m) {
                run += wn[a][m + K0] - wc[a][m + K0];
                cum[a][m] = k.dJ[a] * run;
            }
            if (a < 2) {
#pragma unroll
                for (int m = 0; m < NN; ++m) {
                    P[a][m] = third * wn[a][m + K0] + sixth * wc[a][m + K0];
                    Q[a][m] = third * wc[a][m + K0] + sixth * wn[a][m + K0];
                }
            }
        }
        int n = 0;
        // component x: transverse (y, z); y: (x, z); z: (x, y) -- order matches SameCell<SF>::offset(c, f, m1, m2)
#pragma unroll
        for (int f = 0; f < NN - 1; ++f)
#pragma unroll
            for (int m1 = 0; m1 < NN; ++m1)
#pragma unroll
                for (int m2 = 0; m2 < NN; ++m2) vals[n++] = cum[0][f] * (P[1][m1] * wn[2][m2 + K0] + Q[1][m1] * wc[2][m2 + K0]);
#pragma unroll
        for (int f = 0; f < NN - 1; ++f)
#pragma unroll
            for (int m1 = 0; m1 < NN; ++m1)
#pragma unroll
                for (int m2 = 0; m2 < NN; ++m2) vals[n++] = cum[1][f] * (P[0][m1] * wn[2][m2 + K0] + Q[0][m1] * wc[2][m2 + K0]);
#pragma unroll
        for (int f = 0; f < NN - 1; ++f)
#pragma unroll
            for (int m1 = 0; m1 < NN; ++m1)
#pragma unroll
                for (int m2 = 0; m2 < NN; ++m2) vals[n++] = cum[2][f] * (P[0][m1] * wn[1][m2 + K0] + Q[0][m1] * wc[1][m2 + K0]);
    }
    // ---- move + global particle BCs + ownership, then store
    // Ownership on split axes is the physical crossing direction of the local box [lo, hi): taken BEFORE the periodic wrap
    // (first -> last tile is -1, last -> first is +1, particle_tile_communication.py:145-165) and after reflect/absorb.
    if (PERIODIC1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { s.c[c][i] = wrap_periodic_fast<T>(xn[c], k.wind[c]); s.c[3 + c][i] = v[c]; }
        return kind;
    }
    bool alive = true;
    int dir = 13;   // ((1-0)*3 + (1-0))*3 + (1-0): stays
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        pos[a] = xn[a];
        int off;
        if (k.pbc[a] == PIC_BC_PERIODIC) {
            off = owner_offset_periodic<T>(pos[a], k.box_lo[a], k.box_hi[a], k.wind[a]);
        } else {
            alive = apply_axis_bc<T>(pos[a], v[a], k.wind[a], k.pbc[a]) && alive;
            off = (pos[a] >= k.box_hi[a]) ? 1 : ((pos[a] < k.box_lo[a]) ? -1 : 0);
        }
        dir -= off * (a == 0 ? 9 : (a == 1 ? 3 : 1));
    }
    if (alive && distributed && dir != 13) {
        leave.push<T>(dir, pos, v, species, flags);
        alive = false;
    }
    if (!alive) pos[0] = pic_nan<T>();
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.c[c][i] = pos[c]; s.c[3 + c][i] = v[c]; }
    return kind;
}

// Scalar composition of the K1 v3 body (what one lane does when nothing is shared with its neighbours); used by the host
// check and as the specification of the kernel's warp-aggregated deposit.
template <typename T, int SF, int PUSHER, bool HAS_EXT>
PIC_HD void fused_particle_fast3d(const PicParams& p, int species, const Geom<T>& gm, const FastConst<T>& k, int64_t i,
                                  const SoAView<T>& s, const Field6<T>& F, const Field6<T>& X, const TileSink<T>& sink,
                                  const LeaveBuf& leave, bool distributed, int32_t* flags) {
    constexpr int NN = SameCell<SF>::NN;
    T po[3], xn[3], v[3], vals[SameCell<SF>::NV];
    int key = 0;
    const int kind = fast3d_advance<T, SF, PUSHER, HAS_EXT>(p, species, k, i, s, F, X, leave, distributed, flags, po, xn, v, key, vals);
    if (kind == 1) {
        int n = 0;
        for (int c = 0; c < 3; ++c)
            for (int f = 0; f < NN - 1; ++f)
                for (int m1 = 0; m1 < NN; ++m1)
                    for (int m2 = 0; m2 < NN; ++m2) {
                        const T val = vals[n++];
                        if (val != (T)0) sink.add_unchecked(sink.J[c] + key + SameCell<SF>::offset(c, f, m1, m2, k.sx, k.sy), val);
                    }
    } else if (kind == 2) {
        union_deposit<T, SF>(p, species, gm, k, po, xn, v, sink);
    }
}

}  // namespace pic
