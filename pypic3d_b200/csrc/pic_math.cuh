// pic_math.cuh -- per-particle PIC arithmetic shared by every kernel (reference-layout and resident fast path).
// All functions are __host__ __device__ so the same code can be compiled for the CPU by the test-only
// host-check harness (tests/hostcheck/) in a container without a GPU; the product never runs them on the CPU.
//
// What each block restates (reference = /root/reference/PyPIC3D):
//   tile geometry     utilities/grids.py:80-111, boundary_conditions/grid_and_stencil.py:242-286
//   anchor / weights  boundary_conditions/grid_and_stencil.py:92-138, deposition/shapes.py:6-54
//   gather            pusher/boris.py:128-258, pusher/particle_push.py:63-94
//   pushers           pusher/boris.py:15-124, pusher/higuera_cary.py:58-113
//   Esirkepov         deposition/Esirkepov.py:105-331,365-504
//   direct J / rho    deposition/J_from_rhov.py:83-200, deposition/rho.py:66-150
//   particle BCs      particles/particle_tile_communication.py:41-79, grid_and_stencil.py:15-35
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/pic_b200.h"

#if defined(__CUDACC__)
#define PIC_HD __host__ __device__ __forceinline__
#else
#define PIC_HD inline
#endif

namespace pic {

// ------------------------------------------------------------------------------------------------ helpers
PIC_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);  // no FMA contraction: keep the tile origins bit-identical to the NumPy oracle
#else
    volatile double r = a * b;
    return r;
#endif
}
PIC_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
PIC_HD float pic_floor(float x) { return floorf(x); }
PIC_HD double pic_floor(double x) { return floor(x); }
PIC_HD float pic_rint(float x) { return rintf(x); }   // round-half-even == jnp.round
PIC_HD double pic_rint(double x) { return rint(x); }
PIC_HD float pic_sqrt(float x) { return sqrtf(x); }
PIC_HD double pic_sqrt(double x) { return sqrt(x); }
PIC_HD float pic_fmod(float a, float b) { return fmodf(a, b); }
PIC_HD double pic_fmod(double a, double b) { return fmod(a, b); }
template <typename T> PIC_HD bool pic_isnan(T x) { return !(x == x); }
template <typename T> PIC_HD T pic_nan() { return (T)(NAN); }

// ------------------------------------------------------------------------------------------------ geometry
template <typename T>
struct Geom {
    T oc[3], sc[3];  // tiled center (collocated-node) line: origin = line[0], spacing = line[1]-line[0]
    T ov[3], sv[3];  // tiled vertex (staggered) line
    T d[3];          // dx, dy, dz (the *parameter*, used by the shape weights)
    int L[3];        // local extents W + 2g
    int active[3];   // global N > 1  (particle_push.py:38-42, Esirkepov.py:92-94)
    int g;
};

// line[t, l] = grid[0] + (l + t*W - (g-1)) * d   (utilities/grids.py:98-100); t is the GLOBAL tile coordinate.
template <typename T>
PIC_HD void make_geom(const PicParams& p, int tx, int ty, int tz, Geom<T>& gm) {
    const double dd[3] = {p.dx, p.dy, p.dz};
    const int t[3] = {tx + p.moff[0], ty + p.moff[1], tz + p.moff[2]};
    for (int a = 0; a < 3; ++a) {
        const double k0 = (double)(t[a] * p.tile[a] - (p.g - 1));
        const double c0 = add_rn(p.center0[a], mul_rn(k0, dd[a]));
        const double c1 = add_rn(p.center0[a], mul_rn(k0 + 1.0, dd[a]));
        const double v0 = add_rn(p.vertex0[a], mul_rn(k0, dd[a]));
        const double v1 = add_rn(p.vertex0[a], mul_rn(k0 + 1.0, dd[a]));
        gm.oc[a] = (T)c0;
        gm.sc[a] = (T)add_rn(c1, -c0);
        gm.ov[a] = (T)v0;
        gm.sv[a] = (T)add_rn(v1, -v0);
        gm.d[a] = (T)dd[a];
        gm.L[a] = p.tile[a] + 2 * p.g;
        gm.active[a] = (p.gmesh[a] * p.tile[a] > 1) ? 1 : 0;
    }
    gm.g = p.g;
}

// anchor (floor for CIC, round-half-even for TSC), offset delta and the three shape weights at a-1, a, a+1.
template <typename T, int SF>
PIC_HD void axis_stencil(T pos, T origin, T spacing, T d, int& a, T w[3]) {
    const T q = (pos - origin) / spacing;
    const T fa = (SF == 1) ? pic_floor(q) : pic_rint(q);
    a = (int)fa;
    const T delta = pos - (fa * spacing + origin);
    const T r = delta / d;
    if (SF == 1) {
        w[0] = (T)0;
        w[1] = (T)1 - r;
        w[2] = r;
    } else {
        w[0] = (T)0.5 * ((T)0.5 - r) * ((T)0.5 - r);
        w[1] = (T)0.75 - r * r;
        w[2] = (T)0.5 * ((T)0.5 + r) * ((T)0.5 + r);
    }
}

PIC_HD int wrap_index(int i, int L) {
    i %= L;
    return i < 0 ? i + L : i;
}

// ------------------------------------------------------------------------------------------------ gather
// Component grids (particle_push.py:63-68): 1 = vertex (staggered), 0 = center, per axis.
//   Ex(v,c,c) Ey(c,v,c) Ez(c,c,v) Bx(c,v,v) By(v,c,v) Bz(v,v,c)
template <typename T, int SF>
PIC_HD void gather6(const T* const F[6], size_t tile_off, const Geom<T>& gm, const T pos[3], T out[6]) {
    int idx[2][3][3];
    T w[2][3][3];
    int n[3];
    for (int a = 0; a < 3; ++a) {
        for (int gt = 0; gt < 2; ++gt) {
            int an;
            T ww[3];
            axis_stencil<T, SF>(pos[a], gt ? gm.ov[a] : gm.oc[a], gt ? gm.sv[a] : gm.sc[a], gm.d[a], an, ww);
            if (gm.active[a]) {
                for (int k = 0; k < 3; ++k) {
                    idx[gt][a][k] = wrap_index(an - 1 + k, gm.L[a]);  // jnp.mod(pts, axis_size), boris.py:163-171
                    w[gt][a][k] = ww[k];
                }
            } else {  // collapsed inactive axis: index g, summed weight (boris.py:192-234)
                idx[gt][a][0] = gm.g;
                w[gt][a][0] = (ww[0] + ww[1]) + ww[2];
            }
        }
        n[a] = gm.active[a] ? 3 : 1;
    }
    const int GT[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
    const int k0x = (SF == 1 && n[0] == 3) ? 1 : 0;  // CIC: w[-1] == 0
    const int k0y = (SF == 1 && n[1] == 3) ? 1 : 0;
    const int k0z = (SF == 1 && n[2] == 3) ? 1 : 0;
    for (int c = 0; c < 6; ++c) {
        const int gx = GT[c][0], gy = GT[c][1], gz = GT[c][2];
        const T* f = F[c] + tile_off;
        T acc = (T)0;
        for (int i = k0x; i < n[0]; ++i) {
            T ai = (T)0;
            for (int j = k0y; j < n[1]; ++j) {
                T aj = (T)0;
                const size_t row = ((size_t)idx[gx][0][i] * gm.L[1] + idx[gy][1][j]) * gm.L[2];
                for (int k = k0z; k < n[2]; ++k) aj += f[row + idx[gz][2][k]] * w[gz][2][k];
                ai += aj * w[gy][1][j];
            }
            acc += ai * w[gx][0][i];
        }
        out[c] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ pushers
template <typename T>
PIC_HD void cross3(const T a[3], const T b[3], T o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// v is the velocity (not gamma v); q, m are the raw species charge and mass (particle_push.py:53-54).
template <typename T>
PIC_HD void push_velocity(int pusher, const T v[3], const T E[3], const T B[3], T q, T m, T dt, T C, T out[3]) {
    const T h = q * dt / ((T)2 * m);
    T um[3], t[3], cr[3], up[3], s[3];
    if (pusher == PIC_PUSHER_BORIS) {  // boris.py:41-55
        for (int c = 0; c < 3; ++c) { um[c] = v[c] + h * E[c]; t[c] = h * B[c]; }
        cross3(um, t, cr);
        for (int c = 0; c < 3; ++c) up[c] = um[c] + cr[c];
        const T den = (T)1 + t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
        for (int c = 0; c < 3; ++c) s[c] = (T)2 * t[c] / den;
        cross3(up, s, cr);
        for (int c = 0; c < 3; ++c) out[c] = (um[c] + cr[c]) + h * E[c];
        return;
    }
    const T C2 = C * C;
    const T gamma = (T)1 / pic_sqrt((T)1 - ((v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / C2));
    if (pusher == PIC_PUSHER_BORIS_REL) {  // boris.py:96-121
        for (int c = 0; c < 3; ++c) um[c] = v[c] * gamma + h * E[c];
        const T gm_ = pic_sqrt((T)1 + ((um[0] * um[0] + um[1] * um[1] + um[2] * um[2]) / C2));
        for (int c = 0; c < 3; ++c) t[c] = h * B[c] / gm_;
        cross3(um, t, cr);
        for (int c = 0; c < 3; ++c) up[c] = um[c] + cr[c];
        const T den = (T)1 + t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
        for (int c = 0; c < 3; ++c) s[c] = (T)2 * t[c] / den;
        cross3(up, s, cr);
        T nu[3];
        for (int c = 0; c < 3; ++c) nu[c] = (um[c] + cr[c]) + h * E[c];
        const T ng = pic_sqrt((T)1 + ((nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]) / C2));
        for (int c = 0; c < 3; ++c) out[c] = nu[c] / ng;
        return;
    }
    // Higuera-Cary (higuera_cary.py:58-113)
    T u[3], eps[3], beta[3], ue[3];
    for (int c = 0; c < 3; ++c) { u[c] = gamma * v[c]; eps[c] = h * E[c]; beta[c] = h * B[c]; ue[c] = u[c] + eps[c]; }
    const T beta2 = beta[0] * beta[0] + beta[1] * beta[1] + beta[2] * beta[2];
    const T ustar = (ue[0] * beta[0] + ue[1] * beta[1] + ue[2] * beta[2]) / C;
    const T gue = pic_sqrt((T)1 + (ue[0] * ue[0] + ue[1] * ue[1] + ue[2] * ue[2]) / C2);
    const T sigma = gue * gue - beta2;
    const T gnext = pic_sqrt((sigma + pic_sqrt(sigma * sigma + (T)4 * (beta2 + ustar * ustar))) / (T)2);
    for (int c = 0; c < 3; ++c) t[c] = beta[c] / gnext;
    const T sc = (T)1 / ((T)1 + (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]));
    const T uet = ue[0] * t[0] + ue[1] * t[1] + ue[2] * t[2];
    cross3(ue, t, cr);
    T umid[3];
    for (int c = 0; c < 3; ++c) umid[c] = sc * (ue[c] + uet * t[c] + cr[c]);
    cross3(umid, t, cr);
    T nu[3];
    for (int c = 0; c < 3; ++c) nu[c] = umid[c] + eps[c] + cr[c];
    const T ng = pic_sqrt((T)1 + (nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]) / C2);
    for (int c = 0; c < 3; ++c) out[c] = nu[c] / ng;
}

// ------------------------------------------------------------------------------------------------ deposit sinks
// A sink adds `val` to component c of the ghosted local tile at (ix,iy,iz); out-of-range is dropped
// (`.at[].add(mode="drop")`, Esirkepov.py:238).
template <typename T>
struct TileSink {
    T* J[3];
    size_t off;
    int L[3];
    int32_t* flags = nullptr;   // when set: flags[0] |= 4 if a particle moved more than one cell in a step (its current is NOT deposited)
    PIC_HD void add(int c, int ix, int iy, int iz, T val) const {
        if ((unsigned)ix >= (unsigned)L[0] || (unsigned)iy >= (unsigned)L[1] || (unsigned)iz >= (unsigned)L[2]) return;
        T* ptr = J[c] + off + ((size_t)ix * L[1] + iy) * L[2] + iz;
#if defined(__CUDA_ARCH__)
        atomicAdd(ptr, val);
#else
        *ptr += val;
#endif
    }
    PIC_HD void report_cfl() const {
        if (!flags) return;
#if defined(__CUDA_ARCH__)
        atomicOr(flags, 4);
#else
        *flags |= 4;
#endif
    }
    PIC_HD void add_unchecked(T* ptr, T val) const {
#if defined(__CUDA_ARCH__)
        atomicAdd(ptr, val);
#else
        *ptr += val;
#endif
    }
};

// ------------------------------------------------------------------------------------------------ Esirkepov
// Charge-conserving current for one particle moving xo -> xn inside one tile.  Factorised form of
// Esirkepov.py:381-406 (3-D), :452-502 (2-D), :422-433 (1-D): for an active axis c
//     J_c[i,j,k] = dJ_c * cumsum_i(S1_c - S0_c)[i] * T_ab[j,k]
// with T_ab = 1/3 (S1a S1b + S0a S0b) + 1/6 (S1a S0b + S0a S1b) (both others active), 1/2 (S1a + S0a) (one other
// active) or 1; an inactive axis deposits dJ_c * T (no cumsum) with dJ_c = q w v_c / dV (Esirkepov.py:197-214).
// Only the slots that can be non-zero are visited: SF+2 per axis (|anchor shift| <= 1, guaranteed by |v| dt < d).
template <typename T, int SF, typename Sink>
PIC_HD void esirkepov_deposit(const Geom<T>& gm, const T xo[3], const T xn[3], const T v[3], T qw, T dt, const Sink& sink) {
    constexpr int NS = SF + 2;
    int n[3], base_idx[3];
    T S1[3][NS], S0[3][NS];
    for (int a = 0; a < 3; ++a) {
        int an, ao;
        T wn[3], wo[3];
        axis_stencil<T, SF>(xn[a], gm.oc[a], gm.sc[a], gm.d[a], an, wn);
        axis_stencil<T, SF>(xo[a], gm.oc[a], gm.sc[a], gm.d[a], ao, wo);
        if (!gm.active[a]) {  // collapse_redundant_axis, Esirkepov.py:28-45 (index = L//2 = g)
            n[a] = 1;
            base_idx[a] = gm.L[a] / 2;
            S1[a][0] = (wn[0] + wn[1]) + wn[2];
            S0[a][0] = (wo[0] + wo[1]) + wo[2];
            continue;
        }
        const int s = an - ao;  // shift_old_stencil, Esirkepov.py:17-25
        if (s > 1 || s < -1) {        // > 1 cell per step violates the Courant limit; the reference result is undefined
            sink.report_cfl();
            return;
        }
        n[a] = NS;
        const int b = ((SF == 1) ? 2 : 1) - (s > 0 ? s : 0);  // first visited slot of the 5-slot frame
        base_idx[a] = an - 2 + b;
        for (int m = 0; m < NS; ++m) {
            const int kn = b + m - 1;       // index into the new 3-weight stencil
            const int ko = b + m + s - 1;   // old weights rolled into the new-anchor frame
            S1[a][m] = (kn >= 0 && kn <= 2) ? wn[kn] : (T)0;
            S0[a][m] = (ko >= 0 && ko <= 2) ? wo[ko] : (T)0;
        }
    }
    const T dV = gm.d[0] * gm.d[1] * gm.d[2];
    const T third = (T)(1.0 / 3.0), sixth = (T)(1.0 / 6.0);
    for (int c = 0; c < 3; ++c) {
        const int a = (c + 1) % 3, b = (c + 2) % 3;
        T dJ;
        if (gm.active[c]) dJ = -(qw / (gm.d[a] * gm.d[b])) / dt;
        else dJ = qw * v[c] / dV;
        const int nc = gm.active[c] ? n[c] - 1 : 1;  // the last cumsum entry is sum(S1-S0) == 0
        T cum = (T)0;
        for (int mc = 0; mc < nc; ++mc) {
            T fc;
            if (gm.active[c]) { cum += S1[c][mc] - S0[c][mc]; fc = dJ * cum; }
            else fc = dJ;
            if (fc == (T)0) continue;
            for (int ma = 0; ma < n[a]; ++ma) {
                for (int mb = 0; mb < n[b]; ++mb) {
                    T t;
                    if (gm.active[a] && gm.active[b])
                        t = third * (S1[a][ma] * S1[b][mb] + S0[a][ma] * S0[b][mb]) + sixth * (S1[a][ma] * S0[b][mb] + S0[a][ma] * S1[b][mb]);
                    else if (gm.active[a]) t = (T)0.5 * (S1[a][ma] + S0[a][ma]);
                    else if (gm.active[b]) t = (T)0.5 * (S1[b][mb] + S0[b][mb]);
                    else t = (T)1;
                    const T val = fc * t;
                    if (val == (T)0) continue;
                    int ii[3];
                    ii[c] = base_idx[c] + mc;
                    ii[a] = base_idx[a] + ma;
                    ii[b] = base_idx[b] + mb;
                    sink.add(c, ii[0], ii[1], ii[2], val);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ direct J / rho
// J_c += (q w / dV) v_c Wface_c (x) Wnode_others at the 3x3x3 stencil around the node anchor (J_from_rhov.py:139-198).
// Reduced axes collapse to index g with summed weights (J_from_rhov.py:23-28).
template <typename T, int SF, bool WITH_J, typename Sink>
PIC_HD void node_face_deposit(const Geom<T>& gm, const T pos[3], const T v[3], T qw, const Sink& sink) {
    int n[3], i0[3];
    T wn[3][3], wf[3][3];
    for (int a = 0; a < 3; ++a) {
        int an;
        T w[3];
        axis_stencil<T, SF>(pos[a], gm.oc[a], gm.sc[a], gm.d[a], an, w);
        T f[3] = {0, 0, 0};
        if (WITH_J) {
            const T dface = (pos[a] - gm.oc[a]) - ((T)an + (T)0.5) * gm.d[a];
            const T r = dface / gm.d[a];
            if (SF == 1) { f[0] = (T)0; f[1] = (T)1 - r; f[2] = r; }
            else { f[0] = (T)0.5 * ((T)0.5 - r) * ((T)0.5 - r); f[1] = (T)0.75 - r * r; f[2] = (T)0.5 * ((T)0.5 + r) * ((T)0.5 + r); }
        }
        if (gm.active[a]) {
            n[a] = 3;
            i0[a] = an - 1;
            for (int k = 0; k < 3; ++k) { wn[a][k] = w[k]; wf[a][k] = f[k]; }
        } else {
            n[a] = 1;
            i0[a] = gm.g;
            wn[a][0] = (w[0] + w[1]) + w[2];
            wf[a][0] = (f[0] + f[1]) + f[2];
        }
    }
    const T dq = qw / (gm.d[0] * gm.d[1] * gm.d[2]);
    for (int i = 0; i < n[0]; ++i)
        for (int j = 0; j < n[1]; ++j)
            for (int k = 0; k < n[2]; ++k) {
                const int ix = i0[0] + i, iy = i0[1] + j, iz = i0[2] + k;
                if (WITH_J) {
                    const T jx = dq * v[0] * wf[0][i] * wn[1][j] * wn[2][k];
                    const T jy = dq * v[1] * wn[0][i] * wf[1][j] * wn[2][k];
                    const T jz = dq * v[2] * wn[0][i] * wn[1][j] * wf[2][k];
                    if (jx != (T)0) sink.add(0, ix, iy, iz, jx);
                    if (jy != (T)0) sink.add(1, ix, iy, iz, jy);
                    if (jz != (T)0) sink.add(2, ix, iy, iz, jz);
                } else {
                    const T r = dq * wn[0][i] * wn[1][j] * wn[2][k];
                    if (r != (T)0) sink.add(0, ix, iy, iz, r);
                }
            }
}

// ------------------------------------------------------------------------------------------------ particle BCs
// wrap_periodic_position (grid_and_stencil.py:31-35): jnp.mod has the sign of the divisor.
template <typename T>
PIC_HD T wrap_periodic(T x, T wind) {
    const T h = (T)0.5 * wind;
    T r = pic_fmod(x + h, wind);
    if (r < (T)0) r += wind;
    T w = r - h;
    if (w == -h && x >= h) w = h;
    return w;
}

// _apply_tiled_axis_boundary (particle_tile_communication.py:41-59).  Returns false when absorbed.
template <typename T>
PIC_HD bool apply_axis_bc(T& x, T& u, T wind, int bc) {
    const T h = (T)0.5 * wind;
    if (bc == PIC_BC_PERIODIC) { x = wrap_periodic(x, wind); return true; }
    if (bc == PIC_BC_CONDUCTING) {  // reflecting
        const T x0 = x;
        if (x0 > h) x = (T)2 * h - x0;
        else if (x0 < -h) x = -(T)2 * h - x0;
        if (x0 >= h || x0 <= -h) u = -u;
        return true;
    }
    if (bc == PIC_BC_ABSORBING) return (x <= h) && (x >= -h);
    return true;
}

// _particle_tile_indices (particle_tile_communication.py:62-79) for one axis.
template <typename T>
PIC_HD int dest_tile(T x, T wind, T d, int N, int W, int nt) {
    int cell = (int)pic_floor((x + (T)0.5 * wind) / d);
    cell = cell < 0 ? 0 : (cell > N - 1 ? N - 1 : cell);
    int t = cell / W;
    return t < 0 ? 0 : (t > nt - 1 ? nt - 1 : t);
}

// _adjacent_tile_offset (particle_tile_communication.py:145-165)
PIC_HD int adjacent_offset(int dest, int src, int nt) {
    if (nt == 1) return 0;
    int off = dest - src;
    if (nt == 2) return off;
    if (off == nt - 1) off = -1;
    else if (off == -(nt - 1)) off = 1;
    return off;
}

}  // namespace pic
