// pic_pair.cuh -- K1 v10 per-thread body ("pair body"): W particles per thread advanced together from a supercell tile.
//
// What it computes is the step of pic_slots.cuh fast3d_advance (gather E,B -> Boris -> Esirkepov same-cell currents -> move) for
// CIC, three active axes, g = 2 -- PyPIC3D pusher/particle_push.py:45-144, pusher/boris.py:15-258, deposition/Esirkepov.py:105-406,
// particles/particle_tile_communication.py:82-99 -- restructured around what limits the kernel on B200 (profiles/r01_k1_*):
// instruction issue and shared-memory wavefronts, not HBM.
//   * W = 2 (float): the two particles of a thread travel through EVERY floating-point instruction side by side in Blackwell's
//     packed f32x2 forms (FADD2 / FMUL2 / FFMA2, incl. the round-down FADD2.RM): one issue slot per two particles.
//     W = 1 (double, or float as a control): same code, scalar.
//   * floor() without the conversion pipe: t = q + 1.5 * 2^23 rounded DOWN leaves floor(q) in the low mantissa bits
//     (exact for |q| < 2^22); fa = t - magic is floor(q) as a float, the integer falls out of the bit pattern.
//   * tile-local coordinates: q = (x - x0) / d with x0 the position of the tile's first node, so the cell offset r = q - floor(q)
//     is an exact subtraction and needs no second pass through the global origin (the reference's delta = x - (a s + o):
//     same quantity, rounding differences of one ulp of x; f64 tolerance 1e-12, f32 2e-5);
//     the vertex (staggered) line is the centre line shifted up by half a cell (utilities/grids.py:42-132), so q_v = q_c - 1/2.
//   * gather as seven lerps per component (z, y, x) on the eight corner values read from shared memory.
//   * same-cell Esirkepov weights in mean/difference form: with S0 = (1 - r0, r0), S1 = (1 - r1, r1) the reference's
//     1/3 (S1a S1b + S0a S0b) + 1/6 (S1a S0b + S0a S1b)  ==  Sm_a Sm_b + (1/12) dS_a dS_b,  Sm = (S0 + S1) / 2, dS = S1 - S0
//     (Esirkepov.py:381-406 expanded): 10 instead of 19 operations per current component.
//   * the periodic wrap, reflect / absorb and the ownership test only matter for particles that change cell; those are handed
//     back to the kernel (kind 2), which defers them to the fix-up pass (k_pair_fixup -> crosser_finish); everybody else just rounds
//     (x + h) - h like the reference's mod().
// Everything here is __host__ __device__: tests/hostcheck compiles it with g++ and checks it against the oracle without a GPU.
#pragma once
#include <string.h>

#include "pic_slots.cuh"

namespace pic {

// ---------------------------------------------------------------- W-wide values
template <typename T, int W>
struct Vec {
    T v[W];
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float2 pk(const Vec<float, 2>& a) { return make_float2(a.v[0], a.v[1]); }
__device__ __forceinline__ Vec<float, 2> upk(float2 f) {
    Vec<float, 2> r;
    r.v[0] = f.x; r.v[1] = f.y;
    return r;
}
#define PIC_PACKED(T, W) (W == 2 && sizeof(T) == 4)
#else
#define PIC_PACKED(T, W) false
#endif

PIC_HD float pic_abs(float x) { return fabsf(x); }
PIC_HD double pic_abs(double x) { return fabs(x); }
PIC_HD float pic_fma(float a, float b, float c) { return fmaf(a, b, c); }
PIC_HD double pic_fma(double a, double b, double c) { return fma(a, b, c); }

template <typename T, int W>
PIC_HD Vec<T, W> vsplat(T s) {
    Vec<T, W> r;
#pragma unroll
    for (int j = 0; j < W; ++j) r.v[j] = s;
    return r;
}
template <typename T, int W>
PIC_HD Vec<T, W> vadd(const Vec<T, W>& a, const Vec<T, W>& b) {
#if defined(__CUDA_ARCH__)
    if constexpr (PIC_PACKED(T, W)) return upk(__fadd2_rn(pk(a), pk(b)));
    else
#endif
    {
        Vec<T, W> r;
#pragma unroll
        for (int j = 0; j < W; ++j) r.v[j] = a.v[j] + b.v[j];
        return r;
    }
}
template <typename T, int W>
PIC_HD Vec<T, W> vsub(const Vec<T, W>& a, const Vec<T, W>& b) {
#if defined(__CUDA_ARCH__)
    if constexpr (PIC_PACKED(T, W)) return upk(__fadd2_rn(pk(a), make_float2(-b.v[0], -b.v[1])));   // negation is an operand modifier
    else
#endif
    {
        Vec<T, W> r;
#pragma unroll
        for (int j = 0; j < W; ++j) r.v[j] = a.v[j] - b.v[j];
        return r;
    }
}
template <typename T, int W>
PIC_HD Vec<T, W> vmul(const Vec<T, W>& a, const Vec<T, W>& b) {
#if defined(__CUDA_ARCH__)
    if constexpr (PIC_PACKED(T, W)) return upk(__fmul2_rn(pk(a), pk(b)));
    else
#endif
    {
        Vec<T, W> r;
#pragma unroll
        for (int j = 0; j < W; ++j) r.v[j] = a.v[j] * b.v[j];
        return r;
    }
}
// a * b + c, one rounding
template <typename T, int W>
PIC_HD Vec<T, W> vfma(const Vec<T, W>& a, const Vec<T, W>& b, const Vec<T, W>& c) {
#if defined(__CUDA_ARCH__)
    if constexpr (PIC_PACKED(T, W)) return upk(__ffma2_rn(pk(a), pk(b), pk(c)));
    else
#endif
    {
        Vec<T, W> r;
#pragma unroll
        for (int j = 0; j < W; ++j) r.v[j] = pic_fma(a.v[j], b.v[j], c.v[j]);
        return r;
    }
}
// c - a * b, one rounding
template <typename T, int W>
PIC_HD Vec<T, W> vfnma(const Vec<T, W>& a, const Vec<T, W>& b, const Vec<T, W>& c) {
#if defined(__CUDA_ARCH__)
    if constexpr (PIC_PACKED(T, W)) return upk(__ffma2_rn(make_float2(-a.v[0], -a.v[1]), pk(b), pk(c)));
    else
#endif
    {
        Vec<T, W> r;
#pragma unroll
        for (int j = 0; j < W; ++j) r.v[j] = pic_fma(-a.v[j], b.v[j], c.v[j]);
        return r;
    }
}
template <typename T, int W>
PIC_HD Vec<T, W> vneg(const Vec<T, W>& a) {      // (folds into the operand modifier of the consuming instruction)
    Vec<T, W> r;
#pragma unroll
    for (int j = 0; j < W; ++j) r.v[j] = -a.v[j];
    return r;
}
template <typename T, int W>
PIC_HD Vec<T, W> vmuls(const Vec<T, W>& a, T s) { return vmul(a, vsplat<T, W>(s)); }
template <typename T, int W>
PIC_HD Vec<T, W> vadds(const Vec<T, W>& a, T s) { return vadd(a, vsplat<T, W>(s)); }
// a + r * (b - a)
template <typename T, int W>
PIC_HD Vec<T, W> vlerp(const Vec<T, W>& a, const Vec<T, W>& b, const Vec<T, W>& r) { return vfma(r, vsub(b, a), a); }

// ---- floor through the magic constant: t = q + MAGIC rounded towards -inf; floor(q) = t - MAGIC, (int)floor(q) = magic_int(t)
template <typename T> struct Magic;
template <> struct Magic<float> { static constexpr float value = 12582912.0f; };                  // 1.5 * 2^23
template <> struct Magic<double> { static constexpr double value = 6755399441055744.0; };         // 1.5 * 2^52
PIC_HD int magic_int(float t) {
#if defined(__CUDA_ARCH__)
    int v = __float_as_int(t) - 0x4B400000;
    asm("" : "+r"(v));      // keep the small integer a value of its own: folded into the address arithmetic, the 0x4B400000 no
    return v;               // longer fits the load instructions' immediate offsets and costs one add PER LOAD
#else
    int32_t b;
    memcpy(&b, &t, 4);
    return b - 0x4B400000;
#endif
}
PIC_HD int magic_int(double t) {
#if defined(__CUDA_ARCH__)
    return __double2loint(t);
#else
    int64_t b;
    memcpy(&b, &t, 8);
    return (int32_t)(uint32_t)(b & 0xffffffffll);
#endif
}
template <typename T, int W>
PIC_HD Vec<T, W> vmagic_floor(const Vec<T, W>& q) {
    constexpr T M = Magic<T>::value;
#if defined(__CUDA_ARCH__)
    if constexpr (PIC_PACKED(T, W)) return upk(__fadd2_rd(pk(q), make_float2(M, M)));
    else {
        Vec<T, W> r;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if constexpr (sizeof(T) == 4) r.v[j] = __fadd_rd(q.v[j], M);
            else r.v[j] = __dadd_rd(q.v[j], M);
        }
        return r;
    }
#else
    Vec<T, W> r;
    for (int j = 0; j < W; ++j) r.v[j] = M + pic_floor(q.v[j]);     // exact: |floor(q)| < 2^22
    return r;
#endif
}

// 1/sqrt(x) and 1/x inside the push.  float on the device: the SFU approximations (2 ulp; the f32 tolerance is 2e-5 and the
// arguments are 1 + O(v^2/c^2)); double: rsqrt() (1 ulp) / IEEE division.  Host: IEEE.
PIC_HD float pair_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
PIC_HD double pair_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
PIC_HD float pair_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
PIC_HD double pair_rcp(double x) { return 1.0 / x; }
template <typename T, int W>
PIC_HD Vec<T, W> vrsqrt(const Vec<T, W>& a) {
    Vec<T, W> r;
#pragma unroll
    for (int j = 0; j < W; ++j) r.v[j] = pair_rsqrt(a.v[j]);
    return r;
}
template <typename T, int W>
PIC_HD Vec<T, W> vrcp(const Vec<T, W>& a) {
    Vec<T, W> r;
#pragma unroll
    for (int j = 0; j < W; ++j) r.v[j] = pair_rcp(a.v[j]);
    return r;
}

// ---------------------------------------------------------------- the body
// kinds handed back per particle
constexpr int PAIR_NONE = 0;      // nothing to do (dead slot, slot outside the chunk)
constexpr int PAIR_SAME = 1;      // advanced; stays in its cell: `vals` hold its 12 same-cell currents (SameCell<1> order)
constexpr int PAIR_CROSS = 2;     // advanced; changes cell (or, in an edge supercell, needs boundary work): the caller finishes it
constexpr int PAIR_SLOW = 3;      // NOT advanced: its stencil is not covered by the tile -> scalar global-memory path

// Per-launch constants of the pair body that FastConst does not already carry.
template <typename T>
struct PairConst {
    T dt_inv_d[3];     // dt / d_a   (0 on an axis this species does not move along: SpeciesConfig.update_x)
    T dt_move[3];      // dt         (0 on such an axis)
    T ndJ[3];          // -dJ_a = (q w / (d_b d_c)) / dt: current per unit of (r0 - r1) ... sign folded, see pair_advance
    T half[3];         // wind_a / 2
};
template <typename T>
PIC_HD void make_pair_const(const FastConst<T>& k, PairConst<T>& pc) {
    for (int a = 0; a < 3; ++a) {
        pc.dt_inv_d[a] = k.upd_x[a] ? k.dt * k.inv_d[a] : (T)0;
        pc.dt_move[a] = k.upd_x[a] ? k.dt : (T)0;
        pc.ndJ[a] = -k.dJ[a];
        pc.half[a] = (T)0.5 * k.wind[a];
    }
}

// tile: [6][TILE_N][TILE_NY][TILE_N] (z fastest), component order Ex Ey Ez Bx By Bz; x0[a] = position of the tile's first
// centre-line node.  pos/vel in: state at t; out: pos = (x_new + h) - h (not wrapped), vel = v_new.  xraw = x + v dt.
// Where the body gets the particle state from.  The kernel's loader reads the staged shared-memory slice EVERY time it is asked,
// so x is not kept in registers across the gather and v is not live before the push: fewer registers at the widest point.
template <typename T, int W>
struct ArrayLoader {
    const Vec<T, W>* p;
    const Vec<T, W>* v;
    PIC_HD Vec<T, W> pos(int a) const { return p[a]; }
    PIC_HD Vec<T, W> vel(int a) const { return v[a]; }
};

template <typename T, int W, int PUSHER, bool PER1, class Ld>
PIC_HD void pair_advance_ld(const FastConst<T>& k, const PairConst<T>& pc, const T* tile, const T x0[3], bool edge, const Ld& ld,
                            const bool live[W], Vec<T, W> pos_out[3], Vec<T, W> vel_out[3], Vec<T, W> xraw[3], int kind[W], int cid[W],
                            Vec<T, W> vals[12]);

template <typename T, int W, int PUSHER, bool PER1>
PIC_HD void pair_advance(const FastConst<T>& k, const PairConst<T>& pc, const T* tile, const T x0[3], bool edge,
                         const Vec<T, W> pos[3], const Vec<T, W> vel[3], const bool live[W], Vec<T, W> pos_out[3],
                         Vec<T, W> vel_out[3], Vec<T, W> xraw[3], int kind[W], int cid[W], Vec<T, W> vals[12]) {
    const ArrayLoader<T, W> ld{pos, vel};
    pair_advance_ld<T, W, PUSHER, PER1, ArrayLoader<T, W>>(k, pc, tile, x0, edge, ld, live, pos_out, vel_out, xraw, kind, cid, vals);
}

template <typename T, int W, int PUSHER, bool PER1, class Ld>
PIC_HD void pair_advance_ld(const FastConst<T>& k, const PairConst<T>& pc, const T* tile, const T x0[3], bool edge, const Ld& ld,
                            const bool live[W], Vec<T, W> pos_out[3], Vec<T, W> vel_out[3], Vec<T, W> xraw[3], int kind[W], int cid[W],
                            Vec<T, W> vals[12]) {
    typedef Vec<T, W> V;
    // ---- tile coordinates, coverage test (NaN fails it), safe substitution for everything that is not advanced here
    V q[3];
    bool ok[W];
#pragma unroll
    for (int j = 0; j < W; ++j) ok[j] = live[j];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        q[a] = vmuls(vsub(ld.pos(a), vsplat<T, W>(x0[a])), k.inv_d[a]);
        const V off = vadds(q[a], (T)-3.75);
#pragma unroll
        for (int j = 0; j < W; ++j) ok[j] = ok[j] && (pic_abs(off.v[j]) < (T)3.2499);       // q in (0.5001, 6.9999): both Yee lines covered
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < W; ++j) q[a].v[j] = ok[j] ? q[a].v[j] : (T)3.75;
    // ---- anchors and offsets on the centre and the vertex line
    V tc[3], r0[3], rv[3];
    int ic[3][W], iv[3][W];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        tc[a] = vmagic_floor(q[a]);
        r0[a] = vsub(q[a], vadds(tc[a], -Magic<T>::value));
        const V qv = vadds(q[a], (T)-0.5);
        const V tv = vmagic_floor(qv);
        rv[a] = vsub(qv, vadds(tv, -Magic<T>::value));
#pragma unroll
        for (int j = 0; j < W; ++j) { ic[a][j] = magic_int(tc[a].v[j]); iv[a][j] = magic_int(tv.v[j]); }
#if defined(PIC_ABL10) && PIC_ABL10 == 2
        if (a == 0)
            for (int j = 0; j < W; ++j) iv[a][j] = ic[a][j];
#endif
#if defined(PIC_ABL10) && PIC_ABL10 == 7
        for (int j = 0; j < W; ++j) iv[a][j] = ic[a][j] = (ic[a][j] > 100) ? 1 : 3;
#endif
    }
    // ---- gather: Ex(v,c,c) Ey(c,v,c) Ez(c,c,v) Bx(c,v,v) By(v,c,v) Bz(v,v,c), eight corners -> lerp z, y, x
    V EB[6];
    {
        int ox[2][W], oy[2][W], oz[2][W];      // element offsets of the anchor on [centre, vertex] line
#pragma unroll
        for (int j = 0; j < W; ++j) {
            ox[0][j] = ic[0][j] * TILE_SX; ox[1][j] = iv[0][j] * TILE_SX;
            oy[0][j] = ic[1][j] * TILE_N;  oy[1][j] = iv[1][j] * TILE_N;
            oz[0][j] = ic[2][j];           oz[1][j] = iv[2][j];
        }
        const int GT[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int gx = GT[c][0], gy = GT[c][1], gz = GT[c][2];
            V f000, f001, f010, f011, f100, f101, f110, f111;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const T* f = tile + c * TILE_ELEMS + (ox[gx][j] + oy[gy][j] + oz[gz][j]);
                f000.v[j] = f[0];                 f001.v[j] = f[1];
                f010.v[j] = f[TILE_N];            f011.v[j] = f[TILE_N + 1];
                f100.v[j] = f[TILE_SX];           f101.v[j] = f[TILE_SX + 1];
                f110.v[j] = f[TILE_SX + TILE_N];  f111.v[j] = f[TILE_SX + TILE_N + 1];
            }
            const V& rx = gx ? rv[0] : r0[0];
            const V& ry = gy ? rv[1] : r0[1];
            const V& rz = gz ? rv[2] : r0[2];
            const V a00 = vlerp(f000, f001, rz), a01 = vlerp(f010, f011, rz);
            const V a10 = vlerp(f100, f101, rz), a11 = vlerp(f110, f111, rz);
            const V b0 = vlerp(a00, a01, ry), b1 = vlerp(a10, a11, ry);
            EB[c] = vlerp(b0, b1, rx);
        }
    }
    // ---- push (boris.py:41-55 / 96-121): v is the velocity, u = gamma v only inside the relativistic rotation
    V vn[3];
    const V vel[3] = {ld.vel(0), ld.vel(1), ld.vel(2)};
    {
        const V one = vsplat<T, W>((T)1);
        V hE[3], um[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) hE[c] = vmuls(EB[c], k.h);
        V hb = vsplat<T, W>(k.h);                  // h / gamma^- (relativistic), h otherwise
        if (PUSHER == PIC_PUSHER_BORIS_REL) {
            const V v2 = vfma(vel[2], vel[2], vfma(vel[1], vel[1], vmul(vel[0], vel[0])));
            const V gamma = vrsqrt(vfnma(v2, vsplat<T, W>(k.inv_C2), one));
#pragma unroll
            for (int c = 0; c < 3; ++c) um[c] = vfma(vel[c], gamma, hE[c]);
            const V u2 = vfma(um[2], um[2], vfma(um[1], um[1], vmul(um[0], um[0])));
            hb = vmuls(vrsqrt(vfma(u2, vsplat<T, W>(k.inv_C2), one)), k.h);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) um[c] = vadd(vel[c], hE[c]);
        }
        V t[3], up[3], s[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] = vmul(EB[3 + c], hb);
        // u' = u^- + u^- x t
        up[0] = vfnma(um[2], t[1], vfma(um[1], t[2], um[0]));
        up[1] = vfnma(um[0], t[2], vfma(um[2], t[0], um[1]));
        up[2] = vfnma(um[1], t[0], vfma(um[0], t[1], um[2]));
        const V den = vfma(t[2], t[2], vfma(t[1], t[1], vfma(t[0], t[0], one)));
        const V f2 = vmuls(vrcp(den), (T)2);
#pragma unroll
        for (int c = 0; c < 3; ++c) s[c] = vmul(t[c], f2);
        // u^+ = u^- + u' x s;  u = u^+ + h E
        V nu[3];
        nu[0] = vadd(vfnma(up[2], s[1], vfma(up[1], s[2], um[0])), hE[0]);
        nu[1] = vadd(vfnma(up[0], s[2], vfma(up[2], s[0], um[1])), hE[1]);
        nu[2] = vadd(vfnma(up[1], s[0], vfma(up[0], s[1], um[2])), hE[2]);
        if (PUSHER == PIC_PUSHER_BORIS_REL) {
            const V n2 = vfma(nu[2], nu[2], vfma(nu[1], nu[1], vmul(nu[0], nu[0])));
            const V ig = vrsqrt(vfma(n2, vsplat<T, W>(k.inv_C2), one));
#pragma unroll
            for (int c = 0; c < 3; ++c) vn[c] = vmul(nu[c], ig);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) vn[c] = nu[c];
        }
    }
    // ---- move: new tile coordinate, same-cell test; global position for the store
    V r1[3];
    bool same[W];
#pragma unroll
    for (int j = 0; j < W; ++j) same[j] = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (!k.upd_u[a]) vn[a] = vel[a];                           // frozen velocity component of this species (uniform branch)
        vel_out[a] = vn[a];
        const V qn = vfma(vn[a], vsplat<T, W>(pc.dt_inv_d[a]), q[a]);      // (frozen axis: dt_inv_d = dt_move = 0)
        const V tn = vmagic_floor(qn);
        r1[a] = vsub(qn, vadds(tn, -Magic<T>::value));
#pragma unroll
        for (int j = 0; j < W; ++j) same[j] = same[j] && (magic_int(tn.v[j]) == ic[a][j]);
        xraw[a] = vfma(vn[a], vsplat<T, W>(pc.dt_move[a]), ld.pos(a));
        const V th = vadds(xraw[a], pc.half[a]);
        pos_out[a] = vadds(th, -pc.half[a]);                       // what mod(x + h, wind) - h leaves an interior particle with
        if (edge) {
            // supercell at the rim of the local box: a particle may need boundary work without changing cell (it sits exactly
            // on the upper wall, grid_and_stencil.py:31-35) -- send everything outside [-h, h) / the local box through the crosser path
#pragma unroll
            for (int j = 0; j < W; ++j) {
                bool inside = (th.v[j] >= (T)0) && (th.v[j] < k.wind[a]);
                if (!PER1) inside = inside && (xraw[a].v[j] >= k.box_lo[a]) && (xraw[a].v[j] < k.box_hi[a]);
                same[j] = same[j] && inside;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
#if defined(PIC_ABL10) && PIC_ABL10 == 6
        same[j] = true;
#endif
        kind[j] = ok[j] ? (same[j] ? PAIR_SAME : PAIR_CROSS) : (live[j] ? PAIR_SLOW : PAIR_NONE);
        cid[j] = (ic[0][j] << 6) | (ic[1][j] << 3) | ic[2][j];
    }
    // ---- same-cell Esirkepov currents (zero for every particle that is not PAIR_SAME)
    V dr[3], m0[3], m1[3], cum[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        dr[a] = vsub(r1[a], r0[a]);
#pragma unroll
        for (int j = 0; j < W; ++j) dr[a].v[j] = (kind[j] == PAIR_SAME) ? dr[a].v[j] : (T)0;
        m1[a] = vfma(dr[a], vsplat<T, W>((T)0.5), r0[a]);          // mean offset: weight of node a+1 averaged over the path
        m0[a] = vsub(vsplat<T, W>((T)1), m1[a]);
        cum[a] = vmuls(dr[a], pc.ndJ[a]);                          // dJ_a * (S1 - S0)[node a] = dJ_a * (-(r1 - r0))
    }
    const V s0 = vmuls(dr[0], (T)(1.0 / 12.0)), s1 = vmuls(dr[1], (T)(1.0 / 12.0));
    const V c12 = vmul(s1, dr[2]), c02 = vmul(s0, dr[2]), c01 = vmul(s0, dr[1]);
    // component x: transverse (y, z); y: (x, z); z: (x, y) -- order of SameCell<1>::offset(c, 0, m1, m2)
    const V n12 = vneg(c12), n02 = vneg(c02), n01 = vneg(c01);
    vals[0] = vmul(cum[0], vfma(m0[1], m0[2], c12));
    vals[1] = vmul(cum[0], vfma(m0[1], m1[2], n12));
    vals[2] = vmul(cum[0], vfma(m1[1], m0[2], n12));
    vals[3] = vmul(cum[0], vfma(m1[1], m1[2], c12));
    vals[4] = vmul(cum[1], vfma(m0[0], m0[2], c02));
    vals[5] = vmul(cum[1], vfma(m0[0], m1[2], n02));
    vals[6] = vmul(cum[1], vfma(m1[0], m0[2], n02));
    vals[7] = vmul(cum[1], vfma(m1[0], m1[2], c02));
    vals[8] = vmul(cum[2], vfma(m0[0], m0[1], c01));
    vals[9] = vmul(cum[2], vfma(m0[0], m1[1], n01));
    vals[10] = vmul(cum[2], vfma(m1[0], m0[1], n01));
    vals[11] = vmul(cum[2], vfma(m1[0], m1[1], c01));
}

// Finish a particle that pair_advance handed back as PAIR_CROSS: Esirkepov deposit over the union stencil of its old and new
// cell (global REDs), then the global particle boundary conditions, ownership and the store -- the tail of fast3d_advance.
// v has already been stored by the fast path; it is re-read only where a boundary can change it.
template <typename T, bool PER1>
PIC_HD void crosser_finish(const PicParams& p, int species, const Geom<T>& gm, const FastConst<T>& k, int64_t i, const SoAView<T>& s,
                           const T po[3], const T xn[3], const TileSink<T>& sink, const LeaveBuf& leave, bool distributed, int32_t* flags) {
    const T v0[3] = {(T)0, (T)0, (T)0};            // velocities only enter the deposit on inactive axes (none here)
    union_deposit<T, 1>(p, species, gm, k, po, xn, v0, sink);
    if (PER1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s.c[c][i] = wrap_periodic_fast<T>(xn[c], k.wind[c]);
        return;
    }
    T pos[3] = {xn[0], xn[1], xn[2]};
    // the velocity is only needed where a wall can change it (reflect) or when the particle leaves this rank (it travels in the
    // packet); on periodic axes it stays what the tile kernel stored -- most cell changers cost three stores, not six + three loads
    const bool walls = (k.pbc[0] != PIC_BC_PERIODIC) || (k.pbc[1] != PIC_BC_PERIODIC) || (k.pbc[2] != PIC_BC_PERIODIC);
    T v[3] = {(T)0, (T)0, (T)0};
    if (walls) { v[0] = s.c[3][i]; v[1] = s.c[4][i]; v[2] = s.c[5][i]; }
    bool alive = true;
    int dir = 13;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
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
        if (!walls) { v[0] = s.c[3][i]; v[1] = s.c[4][i]; v[2] = s.c[5][i]; }
        leave.push<T>(dir, pos, v, species, flags);
        alive = false;
    }
    if (!alive) pos[0] = pic_nan<T>();
#pragma unroll
    for (int c = 0; c < 3; ++c) s.c[c][i] = pos[c];
    if (walls) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s.c[3 + c][i] = v[c];
    }
}

}  // namespace pic
