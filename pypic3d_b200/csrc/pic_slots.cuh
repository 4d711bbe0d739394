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
};
template <typename T>
static inline SoAView<T> view_of(const PicSoA* s) {
    SoAView<T> v;
    for (int k = 0; k < 6; ++k) v.c[k] = (T*)s->comp[k];
    v.id = s->id;
    v.cap = s->cap;
    v.n = s->n;
    return v;
}

struct LeaveBuf {
    void* buf;        // [27][leave_cap][7]
    int64_t cap;
    int32_t* count;   // [27]
};

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
            const int64_t slot = atomic_add_i32(&leave.count[dir], 1);
            if (slot < leave.cap) {
                T* pk = (T*)leave.buf + ((int64_t)dir * leave.cap + slot) * 7;
                for (int c = 0; c < 3; ++c) { pk[c] = pos[c]; pk[3 + c] = v[c]; }
                pk[6] = (T)species;
            } else {
                atomic_or_i32(flags, 2);
            }
            alive = false;
        }
    }
    if (!alive) pos[0] = pic_nan<T>();
    for (int c = 0; c < 3; ++c) { s.c[c][i] = pos[c]; s.c[3 + c][i] = v[c]; }
}

}  // namespace pic
