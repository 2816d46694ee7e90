// nb_sweep_f32.cuh — the FP32 pair sweep (NB_PREC_F32 / NB_PREC_F32_COMP).
//
// Replaces nbody_forces_kernel_float4 / nbody_forces_kahan_kernel_float4 /
// nbody_potential_kernel_float4 / nbody_potential_kahan_kernel_float4
// (reference nbody_streams/cuda_kernels.py:90-179, :254-366, :439-518, :591-677).
//
// B200 design (not a translation of the reference's 1-i-per-thread tile loop):
//   * each thread owns NP packed PAIRS of i-particles held in 64-bit registers; all
//     pair arithmetic is issued as Blackwell packed-FP32 instructions
//     (FADD2 / FFMA2 / FMUL2, sm_100+) with the j-particle as a scalar-broadcast
//     operand: one warp instruction advances 64 interactions;
//   * j-tiles (structure-of-arrays, 256 particles) are streamed into a 3-deep
//     shared-memory ring by the TMA engine (cp.async.bulk + mbarrier, one elected
//     thread); the inner loop fetches four j-particles per broadcast LDS.128;
//   * the grid is persistent (SMs x resident CTAs) and splits the (i-block x j-tile)
//     work stream-K style, so every CTA does the same number of tiles whatever N is;
//   * summation is blocked: FFMA partial sums over one tile are folded once per tile
//     into running sums kept in shared memory — float for 'float32', DOUBLE for
//     'float32_kahan' (compensated summation at <1 % of the cost of the reference's
//     12 extra FADDs per interaction; the reference's own CPU path also keeps f64
//     accumulators, fields.py:221,248-250);
//   * nb_pack tags every tile whose 256 particles share one mass and/or one softening;
//     such tiles skip the per-interaction FMNMX (max-softening rule hoisted per tile)
//     and the mass multiply (applied once per tile in the fold);
//   * the self term and padding are masked only in the tiles that can contain them;
//     the cubic spline is a clamped Newtonian term for every pair plus an exact
//     correction in the rare (r < h) case.
#pragma once
#include <type_traits>

#include "nb_sweep.cuh"

namespace nb {

// ---------------------------------------------------------------------------
// one (pair of i) x (one j) interaction, packed f32x2
//   xi,yi,zi : i positions (pair);  xj.. the j particle (scalar-broadcast operand)
//   UM : tile has one mass  -> sums are accumulated unweighted, m applied in the fold
//   UC : tile has one softening term -> cpair = max(c_tile, c_i) precomputed per tile
//   sp : accumulates +m/r-like terms; phi = -G * sp at the end
// ---------------------------------------------------------------------------
template <int KID, int MODE, bool MASK, bool UM, bool UC>
__device__ __forceinline__ void pair_f32(const float xj, const float yj, const float zj, const float mjs,
                                         const float cj, const f2 xi, const f2 yi, const f2 zi, const float ci0,
                                         const float ci1, const f2 cpair, f2& ax, f2& ay, f2& az, f2& sp,
                                         const float eps2, bool& inside, const bool v0, const bool v1) {
    const f2 dx = sub2(splat(xj), xi);
    const f2 dy = sub2(splat(yj), yi);
    const f2 dz = sub2(splat(zj), zi);
    constexpr bool kUsesC = (KID == NB_KERNEL_PLUMMER || KID == NB_KERNEL_DEHNEN_K1 || KID == NB_KERNEL_DEHNEN_K2);
    // h2: softening term of the pair (max rule); c: start value of the r^2 FFMA chain
    f2 h2, c;
    if (KID == NB_KERNEL_NEWTONIAN) {
        h2 = c = splat(eps2);
    } else {
        h2 = UC ? cpair : pk(fmaxf(cj, ci0), fmaxf(cj, ci1));
        c  = kUsesC ? h2 : splat(eps2);
    }
    f2 mj = splat(mjs);
    if (MASK) {   // invalid lanes (self pair, padding): zero mass, harmless distance, never "inside h"
        float m0 = mjs, m1 = mjs, c0, c1, g0, g1;
        upk(c, c0, c1);
        upk(h2, g0, g1);
        if (!v0) { m0 = 0.f; c0 = 1.f; g0 = 0.f; }
        if (!v1) { m1 = 0.f; c1 = 1.f; g1 = 0.f; }
        mj = pk(m0, m1);
        c  = pk(c0, c1);
        if (KID == NB_KERNEL_SPLINE) h2 = pk(g0, g1);
    }
    const f2 d = fma2(dx, dx, fma2(dy, dy, fma2(dz, dz, c)));
    float d0, d1;
    upk(d, d0, d1);
    if (KID == NB_KERNEL_SPLINE) {
        float h20, h21;
        upk(h2, h20, h21);
        inside |= (d0 < h20) | (d1 < h21);
        d0 = fmaxf(d0, h20);   // clamp: Newtonian value at r = h for inside pairs,
        d1 = fmaxf(d1, h21);   // the exact spline correction is added by spline_fix()
    }
    const f2 rinv = pk(rsqrt_approx(d0), rsqrt_approx(d1));

    if (KID == NB_KERNEL_DEHNEN_K1 || KID == NB_KERNEL_DEHNEN_K2) {
        const f2 mr    = UM ? rinv : mul2(mj, rinv);
        const f2 rinv2 = mul2(rinv, rinv);
        // the packed c holds h^2 + eps2 (it seeds the r^2 chain); the correction terms want the bare h^2
        const f2 t     = mul2(add2(h2, splat(-eps2)), rinv2);
        if (MODE & NB_MODE_POT) {
            f2 pf;
            if (KID == NB_KERNEL_DEHNEN_K1) pf = fma2(splat(0.5f), t, splat(1.f));
            else pf = fma2(t, fma2(splat(0.375f), t, splat(0.5f)), splat(1.f));
            sp = fma2(mr, pf, sp);
        }
        if (MODE & NB_MODE_FORCE) {
            f2 ff;
            if (KID == NB_KERNEL_DEHNEN_K1) ff = fma2(splat(1.5f), t, splat(1.f));
            else ff = fma2(t, fma2(splat(3.75f), t, splat(1.5f)), splat(1.f));
            const f2 f = mul2(mul2(mr, rinv2), ff);
            ax = fma2(f, dx, ax);
            ay = fma2(f, dy, ay);
            az = fma2(f, dz, az);
        }
    } else {
        if (MODE == NB_MODE_POT) {
            sp = UM ? add2(sp, rinv) : fma2(mj, rinv, sp);
        } else {
            const f2 mr = UM ? rinv : mul2(mj, rinv);
            if (MODE & NB_MODE_POT) sp = add2(sp, mr);
            const f2 rinv2 = mul2(rinv, rinv);
            const f2 f     = mul2(mr, rinv2);
            ax = fma2(f, dx, ax);
            ay = fma2(f, dy, ay);
            az = fma2(f, dz, az);
        }
    }
}

// exact spline correction for one (i, j) with r < h: replaces the clamped Newtonian term the
// packed path added by the true spline value (UM: unweighted, like the packed path).
template <int MODE, bool UM>
__device__ __forceinline__ void spline_fix(const float xj, const float yj, const float zj, const float mj_in,
                                           const float cj, const float xi, const float yi, const float zi,
                                           const float ci, float& ax, float& ay, float& az, float& sp,
                                           const float eps2, const bool valid) {
    const float dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const float d  = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, eps2)));
    const float h2 = fmaxf(cj, ci);
    if (valid && d < h2) {
        const float mj = UM ? 1.f : mj_in;
        const float rc = rsqrt_approx(h2);   // what the packed path used: rsqrt(max(d, h2)); also 1/h
        const float mr = UM ? rc : mj * rc;
        float w, p;
        spline_inside<float>(d, rsqrt_approx(d), rc, w, p);
        if (MODE & NB_MODE_FORCE) {
            const float f_main = mr * (rc * rc);
            const float f      = fmaf(mj, w, -f_main);
            ax = fmaf(f, dx, ax);
            ay = fmaf(f, dy, ay);
            az = fmaf(f, dz, az);
        }
        if (MODE & NB_MODE_POT) sp += fmaf(mj, p, -mr);
    }
}

// per-thread i-state: NP packed pairs + their per-tile partial sums
template <int NP>
struct IState {
    f2 x[NP], y[NP], z[NP];        // positions (i0, i1)
    float c0[NP], c1[NP];          // softening terms
    f2 ax[NP], ay[NP], az[NP], sp[NP];
};

// ---------------------------------------------------------------------------
// one j-tile against this thread's NP pairs
// ---------------------------------------------------------------------------
template <int KID, int MODE, bool MASK, bool UM, bool UC, int NP, int JUF>
__device__ __forceinline__ void tile_f32(const unsigned char* __restrict__ stage, IState<NP>& S, const float eps2,
                                         const float c_tile, const long long jg0, const long long ig0,
                                         const long long n_j, const int jb = 0, const int je = TJ) {
    const float* __restrict__ sX = reinterpret_cast<const float*>(stage) + COMP_X * TJ;
    const float* __restrict__ sY = reinterpret_cast<const float*>(stage) + COMP_Y * TJ;
    const float* __restrict__ sZ = reinterpret_cast<const float*>(stage) + COMP_Z * TJ;
    const float* __restrict__ sM = reinterpret_cast<const float*>(stage) + COMP_M * TJ;
    const float* __restrict__ sC = reinterpret_cast<const float*>(stage) + COMP_C * TJ;
    constexpr int JU = JUF;   // j-particles per loop iteration (4 or 8)
    constexpr bool kNeedC = (KID != NB_KERNEL_NEWTONIAN) && !UC;

    f2 cpair[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) cpair[q] = UC ? pk(fmaxf(c_tile, S.c0[q]), fmaxf(c_tile, S.c1[q])) : 0ull;
    // masked tiles: validity in 32-bit arithmetic.  rel = (global index of the tile's first j) - (global
    // index of this thread's first i); the self pair of i-pair q is j_local + rel == 2*THREADS*q (+1);
    // nv = number of real (non-padding) particles in this tile.
    int rel = 0, nv = TJ;
    if (MASK) {
        const long long D = jg0 - ig0;
        rel = (D > (1LL << 29) || D < -(1LL << 29)) ? (1 << 29) : static_cast<int>(D);
        const long long left = n_j - jg0;
        nv = left >= TJ ? TJ : (left <= 0 ? 0 : static_cast<int>(left));
    }

#pragma unroll 1
    for (int j0 = jb; j0 < je; j0 += JU) {   // [jb, je): the whole tile, or one part of it (sub-tile work units)
        float xs[8], ys[8], zs[8], ms[8], cs[8];
        if (JU >= 4) {   // broadcast LDS.128: four j-particles per load
#pragma unroll
            for (int g = 0; g < JU / 4; ++g) {
                const float4 x4 = *reinterpret_cast<const float4*>(sX + j0 + 4 * g);
                const float4 y4 = *reinterpret_cast<const float4*>(sY + j0 + 4 * g);
                const float4 z4 = *reinterpret_cast<const float4*>(sZ + j0 + 4 * g);
                xs[4 * g] = x4.x; xs[4 * g + 1] = x4.y; xs[4 * g + 2] = x4.z; xs[4 * g + 3] = x4.w;
                ys[4 * g] = y4.x; ys[4 * g + 1] = y4.y; ys[4 * g + 2] = y4.z; ys[4 * g + 3] = y4.w;
                zs[4 * g] = z4.x; zs[4 * g + 1] = z4.y; zs[4 * g + 2] = z4.z; zs[4 * g + 3] = z4.w;
                if (!UM) {
                    const float4 m4 = *reinterpret_cast<const float4*>(sM + j0 + 4 * g);
                    ms[4 * g] = m4.x; ms[4 * g + 1] = m4.y; ms[4 * g + 2] = m4.z; ms[4 * g + 3] = m4.w;
                } else {
                    ms[4 * g] = ms[4 * g + 1] = ms[4 * g + 2] = ms[4 * g + 3] = 1.f;
                }
                if (kNeedC) {
                    const float4 c4 = *reinterpret_cast<const float4*>(sC + j0 + 4 * g);
                    cs[4 * g] = c4.x; cs[4 * g + 1] = c4.y; cs[4 * g + 2] = c4.z; cs[4 * g + 3] = c4.w;
                } else {
                    cs[4 * g] = cs[4 * g + 1] = cs[4 * g + 2] = cs[4 * g + 3] = c_tile;
                }
            }
        } else if (JU == 2) {   // broadcast LDS.64: two j-particles per load
            const float2 x2 = *reinterpret_cast<const float2*>(sX + j0);
            const float2 y2 = *reinterpret_cast<const float2*>(sY + j0);
            const float2 z2 = *reinterpret_cast<const float2*>(sZ + j0);
            xs[0] = x2.x; xs[1] = x2.y; ys[0] = y2.x; ys[1] = y2.y; zs[0] = z2.x; zs[1] = z2.y;
            ms[0] = ms[1] = 1.f;
            cs[0] = cs[1] = c_tile;
            if (!UM) { const float2 m2 = *reinterpret_cast<const float2*>(sM + j0); ms[0] = m2.x; ms[1] = m2.y; }
            if (kNeedC) { const float2 c2 = *reinterpret_cast<const float2*>(sC + j0); cs[0] = c2.x; cs[1] = c2.y; }
        } else {
            xs[0] = sX[j0]; ys[0] = sY[j0]; zs[0] = sZ[j0]; ms[0] = sM[j0]; cs[0] = sC[j0];
        }
        bool inside = false;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
#pragma unroll
            for (int jj = 0; jj < JU; ++jj) {
                bool v0 = true, v1 = true;
                if (MASK) {
                    const int jl = j0 + jj;
                    const int dd = jl + rel - 2 * THREADS * q;
                    const bool jv = jl < nv;
                    v0 = jv && (dd != 0);
                    v1 = jv && (dd != 1);
                }
                pair_f32<KID, MODE, MASK, UM, UC>(xs[jj], ys[jj], zs[jj], ms[jj], cs[jj], S.x[q], S.y[q], S.z[q],
                                                  S.c0[q], S.c1[q], cpair[q], S.ax[q], S.ay[q], S.az[q], S.sp[q],
                                                  eps2, inside, v0, v1);
            }
        }
        if (KID == NB_KERNEL_SPLINE) {
            if (inside) {
#pragma unroll 1
                for (int jj = 0; jj < JU; ++jj) {
                    const float xj = sX[j0 + jj], yj = sY[j0 + jj], zj = sZ[j0 + jj], mj = sM[j0 + jj],
                                cj = sC[j0 + jj];
                    const int jl = j0 + jj;
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        bool v0 = true, v1 = true;
                        if (MASK) {
                            const int dd = jl + rel - 2 * THREADS * q;
                            const bool jv = jl < nv;
                            v0 = jv && (dd != 0);
                            v1 = jv && (dd != 1);
                        }
                        float x0, x1, y0, y1, z0, z1, ax0, ax1, ay0, ay1, az0, az1, sp0, sp1;
                        upk(S.x[q], x0, x1); upk(S.y[q], y0, y1); upk(S.z[q], z0, z1);
                        upk(S.ax[q], ax0, ax1); upk(S.ay[q], ay0, ay1); upk(S.az[q], az0, az1);
                        upk(S.sp[q], sp0, sp1);
                        spline_fix<MODE, UM>(xj, yj, zj, mj, cj, x0, y0, z0, S.c0[q], ax0, ay0, az0, sp0, eps2, v0);
                        spline_fix<MODE, UM>(xj, yj, zj, mj, cj, x1, y1, z1, S.c1[q], ax1, ay1, az1, sp1, eps2, v1);
                        S.ax[q] = pk(ax0, ax1); S.ay[q] = pk(ay0, ay1); S.az[q] = pk(az0, az1);
                        S.sp[q] = pk(sp0, sp1);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Cold path kept OUT of the hot loops' register allocation: masked tiles (<= 5 per i-block) run in a
// non-inlined function that works on a COPY of the thread state (so the state itself never has its address
// taken and stays in registers in the hot loops).  ptxas schedules the hot loop noticeably worse (measured:
// -4.6 %) when that rare code shares its register allocation.  The spline's exact groups, spline_exact(), are
// NOT treated that way here: they are inlined into tile_f32_spline behind a branch of the hot loop (the spline
// kernels carry 28-56 B of spills for it).  The pair-symmetric sweep (nb_sym_f32.cuh), which handles every
// large-N call, does put its exact step behind a call on a state copy.
// ---------------------------------------------------------------------------
template <int KID, int MODE, int NP, int JUF>
__device__ __noinline__ void tile_f32_masked(const unsigned char* stage, IState<NP>* T, const float eps2,
                                             const long long jg0, const long long ig0, const long long n_j,
                                             const int jb = 0, const int je = TJ) {
    tile_f32<KID, MODE, true, false, false, NP, JUF>(stage, *T, eps2, 0.f, jg0, ig0, n_j, jb, je);
}

// ---------------------------------------------------------------------------
// Cubic spline, unmasked tiles: two-phase groups.
// The spline is exactly Newtonian for r >= h (cuda_kernels.py:64-67) and for realistic inputs ~1e-6 of
// the pairs are inside h.  Per group of 2 j-particles x NP i-pairs: phase 1 computes the separations
// and r^2 and keeps a running minimum (one FMNMX3 per pair of interactions); if the minimum is not
// below the largest softening term that can apply (tile max of c_j, thread max of c_i) the group
// finishes on the pure Newtonian path — no clamp, no per-pair compare; otherwise the whole group is
// redone by exact_group() with the reference's piecewise polynomial.
// ---------------------------------------------------------------------------
template <int MODE, bool UM>
__device__ __forceinline__ void spline_exact(const float xj, const float yj, const float zj, const float mj_in,
                                             const float cj, const float xi, const float yi, const float zi,
                                             const float ci, float& ax, float& ay, float& az, float& sp,
                                             const float eps2) {
    const float dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const float d  = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, eps2)));
    const float h2 = fmaxf(cj, ci);
    const float mj = UM ? 1.f : mj_in;
    float w, p;
    const float ri = rsqrt_approx(d);
    if (d < h2) {
        spline_inside<float>(d, ri, rsqrt_approx(h2), w, p);
    } else {
        p = ri;
        w = ri * (ri * ri);
    }
    if (MODE & NB_MODE_FORCE) {
        const float f = mj * w;
        ax = fmaf(f, dx, ax);
        ay = fmaf(f, dy, ay);
        az = fmaf(f, dz, az);
    }
    if (MODE & NB_MODE_POT) sp = fmaf(mj, p, sp);
}

template <int MODE, bool UM, int NP>
__device__ __forceinline__ void tile_f32_spline(const unsigned char* __restrict__ stage, IState<NP>& S,
                                                const float eps2, const float c_max_tile, const int jb = 0,
                                                const int je = TJ) {
    const float* __restrict__ sX = reinterpret_cast<const float*>(stage) + COMP_X * TJ;
    const float* __restrict__ sY = reinterpret_cast<const float*>(stage) + COMP_Y * TJ;
    const float* __restrict__ sZ = reinterpret_cast<const float*>(stage) + COMP_Z * TJ;
    const float* __restrict__ sM = reinterpret_cast<const float*>(stage) + COMP_M * TJ;
    const float* __restrict__ sC = reinterpret_cast<const float*>(stage) + COMP_C * TJ;
    float h2max = c_max_tile;
#pragma unroll
    for (int q = 0; q < NP; ++q) h2max = fmaxf(h2max, fmaxf(S.c0[q], S.c1[q]));
    const f2 ceps = splat(eps2);

#pragma unroll 1
    for (int j0 = jb; j0 < je; j0 += 4) {
        const float4 x4 = *reinterpret_cast<const float4*>(sX + j0);
        const float4 y4 = *reinterpret_cast<const float4*>(sY + j0);
        const float4 z4 = *reinterpret_cast<const float4*>(sZ + j0);
        float4 m4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (!UM) m4 = *reinterpret_cast<const float4*>(sM + j0);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w},
                    zs[4] = {z4.x, z4.y, z4.z, z4.w}, ms[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int g = 0; g < 2; ++g) {   // two groups of 2 j-particles
            f2 dx[2][NP], dy[2][NP], dz[2][NP], d[2][NP];
            float dmin = 3.0e38f;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    dx[jj][q] = sub2(splat(xs[2 * g + jj]), S.x[q]);
                    dy[jj][q] = sub2(splat(ys[2 * g + jj]), S.y[q]);
                    dz[jj][q] = sub2(splat(zs[2 * g + jj]), S.z[q]);
                    d[jj][q]  = fma2(dx[jj][q], dx[jj][q], fma2(dy[jj][q], dy[jj][q], fma2(dz[jj][q], dz[jj][q], ceps)));
                    float d0, d1;
                    upk(d[jj][q], d0, d1);
                    dmin = fminf(dmin, fminf(d0, d1));
                }
            if (dmin < h2max) {   // rare: some pair of this group may be inside its softening length
#pragma unroll 1
                for (int jj = 0; jj < 2; ++jj) {
                    const int j = j0 + 2 * g + jj;
                    const float xj = sX[j], yj = sY[j], zj = sZ[j], mj = sM[j], cj = sC[j];
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        float x0, x1, y0, y1, z0, z1, ax0, ax1, ay0, ay1, az0, az1, sp0, sp1;
                        upk(S.x[q], x0, x1); upk(S.y[q], y0, y1); upk(S.z[q], z0, z1);
                        upk(S.ax[q], ax0, ax1); upk(S.ay[q], ay0, ay1); upk(S.az[q], az0, az1);
                        upk(S.sp[q], sp0, sp1);
                        spline_exact<MODE, UM>(xj, yj, zj, mj, cj, x0, y0, z0, S.c0[q], ax0, ay0, az0, sp0, eps2);
                        spline_exact<MODE, UM>(xj, yj, zj, mj, cj, x1, y1, z1, S.c1[q], ax1, ay1, az1, sp1, eps2);
                        S.ax[q] = pk(ax0, ax1); S.ay[q] = pk(ay0, ay1); S.az[q] = pk(az0, az1);
                        S.sp[q] = pk(sp0, sp1);
                    }
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        float d0, d1;
                        upk(d[jj][q], d0, d1);
                        const f2 rinv = pk(rsqrt_approx(d0), rsqrt_approx(d1));
                        if (MODE == NB_MODE_POT) {
                            S.sp[q] = UM ? add2(S.sp[q], rinv) : fma2(splat(ms[2 * g + jj]), rinv, S.sp[q]);
                        } else {
                            const f2 mr = UM ? rinv : mul2(splat(ms[2 * g + jj]), rinv);
                            if (MODE & NB_MODE_POT) S.sp[q] = add2(S.sp[q], mr);
                            const f2 f = mul2(mr, mul2(rinv, rinv));
                            S.ax[q] = fma2(f, dx[jj][q], S.ax[q]);
                            S.ay[q] = fma2(f, dy[jj][q], S.ay[q]);
                            S.az[q] = fma2(f, dz[jj][q], S.az[q]);
                        }
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// the sweep kernel
//   ACC = 0: per-tile partials folded into FLOAT running sums  (NB_PREC_F32)
//   ACC = 1: per-tile partials folded into DOUBLE running sums (NB_PREC_F32_COMP, 'float32_kahan')
// Running sums live in shared memory ([4][2*NP][THREADS], conflict-free) so the
// register file is left to the pair arithmetic.
// ---------------------------------------------------------------------------
template <int NP, int ACC>
struct SweepF32Smem {
    static constexpr int kStageBytes = STAGES * TILE_BYTES_F32;
    static constexpr int kBarOff     = (kStageBytes + 15) & ~15;
    static constexpr int kRunOff     = kBarOff + 64;
    static constexpr int kRunBytes   = 4 * 2 * NP * THREADS * (ACC ? 8 : 4);
    static constexpr int kTotal      = kRunOff + kRunBytes;
};

//   DEC = false: one __syncthreads per tile frees the previous stage, thread 0 refills it
//   DEC = true : warps run decoupled — every warp arrives on the stage's "empty" mbarrier when it is
//                done with a tile and the LAST warp out (pending count 1) issues the TMA refill, so no
//                thread ever waits for another warp except through data availability
//   SUB = true : sub-tile work units (a.split_log2 > 0, small N); SUB = false: whole tiles, constant loop bounds
template <int KID, int MODE, int ACC, int NP, int MINB, int JUF = 4, bool DEC = false, bool SUB = false>
__global__ void __launch_bounds__(THREADS, MINB) sweep_f32_kernel(const SweepArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    using L = SweepF32Smem<NP, ACC>;
    using run_t = typename std::conditional<ACC != 0, double, float>::type;
    constexpr int IB = THREADS * 2 * NP;
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* empty_bar = full_bar + STAGES;
    run_t* run          = reinterpret_cast<run_t*>(smem + L::kRunOff);   // [4][2*NP][THREADS]

    const int tid = threadIdx.x;
    const int cta = blockIdx.x, grid = gridDim.x;
    const long long u0 = unit_begin(a.total, cta, grid);
    const long long u1 = unit_begin(a.total, cta + 1, grid);
    const int n_units  = static_cast<int>(u1 - u0);
    const float eps2   = static_cast<float>(a.eps2);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int sl  = SUB ? a.split_log2 : 0;
    const int upb = a.n_jt << sl;               // work units per i-block
    const int plen = TJ >> sl;                  // j-particles per unit
    auto issue = [&](int k) {   // elected thread: TMA bulk load of unit k's j-tile
        const long long u = u0 + k;
        const int jl      = static_cast<int>(u % upb) >> sl;
        const int gt      = local_to_global_tile(a, jl);
        uint64_t* bar     = &full_bar[k % STAGES];
        mbar_expect_tx(bar, TILE_BYTES_F32);
        bulk_g2s(smem + (k % STAGES) * TILE_BYTES_F32, a.packed + static_cast<size_t>(gt) * TILE_BYTES_F32,
                 TILE_BYTES_F32, bar);
    };
    if (tid == 0) {
        for (int k = 0; k < (DEC ? STAGES : STAGES - 1) && k < n_units; ++k) issue(k);
    }

    long long u = u0;
    int k = 0;
    while (u < u1) {
        const int ib  = static_cast<int>(u / upb);
        const int jl0 = static_cast<int>(u - static_cast<long long>(ib) * upb);   // first unit of the segment
        const long long rem = u1 - u;
        const int jl1 = static_cast<int>(min(static_cast<long long>(upb), jl0 + rem));

        // ---- this thread's i-pairs (read straight from the packed records) ----
        const long long ig_block = a.i_begin + static_cast<long long>(ib) * IB;
        const long long ig0      = ig_block + 2LL * tid;
        IState<NP> S;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const long long ig = ig0 + 2LL * THREADS * q;   // even; ig+1 is in the same tile
            S.x[q] = S.y[q] = S.z[q] = 0ull;
            S.c0[q] = S.c1[q] = 0.f;   // dummy i beyond the buffer: results are discarded
            if (ig < a.n_pad) {   // aligned 64-bit loads give (i, i+1) register pairs directly
                const long long t = ig / TJ;
                const int kk      = static_cast<int>(ig - t * TJ);
                const float* tb   = reinterpret_cast<const float*>(a.packed + static_cast<size_t>(t) * TILE_BYTES_F32);
                S.x[q] = *reinterpret_cast<const f2*>(tb + COMP_X * TJ + kk);
                S.y[q] = *reinterpret_cast<const f2*>(tb + COMP_Y * TJ + kk);
                S.z[q] = *reinterpret_cast<const f2*>(tb + COMP_Z * TJ + kk);
                const float2 c2 = *reinterpret_cast<const float2*>(tb + COMP_C * TJ + kk);
                S.c0[q] = c2.x;
                S.c1[q] = c2.y;
            }
            S.ax[q] = S.ay[q] = S.az[q] = S.sp[q] = 0ull;
        }
#pragma unroll
        for (int e = 0; e < 4 * 2 * NP; ++e) run[e * THREADS + tid] = run_t(0);

        for (int v = jl0; v < jl1; ++v, ++k) {
            const int jl = v >> sl;                       // tile of this unit
            const int jb = (v - (jl << sl)) * plen;       // and the part of it: j-particles [jb, jb + plen)
            const int je = jb + plen;
            if (!DEC) {
                __syncthreads();   // everyone is done with tile k-1 -> its stage may be refilled
                if (tid == 0 && k + STAGES - 1 < n_units) issue(k + STAGES - 1);
            }
            mbar_wait(&full_bar[k % STAGES], (k / STAGES) & 1);

            const int gt        = local_to_global_tile(a, jl);
            const long long jg0 = static_cast<long long>(gt) * TJ;
            const bool masked   = (jg0 + TJ > a.n_j) || (jg0 < ig_block + IB && jg0 + TJ > ig_block);
            const unsigned char* stage = smem + (k % STAGES) * TILE_BYTES_F32;
            const float4 meta = *reinterpret_cast<const float4*>(stage + 5 * TJ * 4);   // {c_uni|NaN, m_uni|NaN, c_max, -}
            const bool uc = (meta.x == meta.x), um = (meta.y == meta.y);
            float m_fold = 1.f;

            // Whole tiles (the only case from a few 10^4 particles on) call the tile loops with CONSTANT bounds: the hot
            // loops are sensitive to code generation (a run-time trip count cost 4-5 % at N = 1M, a second inlined
            // copy in the same kernel up to 13 %), so the sub-tile case is a separate kernel instantiation (SUB).
#define NB_TILE_DISPATCH_F32(JB, JE)                                                                                  \
    if (KID == NB_KERNEL_SPLINE && !masked) {                                                                         \
        if (um) {                                                                                                     \
            m_fold = meta.y;                                                                                          \
            tile_f32_spline<MODE, true, NP>(stage, S, eps2, meta.z, JB, JE);                                          \
        } else {                                                                                                      \
            tile_f32_spline<MODE, false, NP>(stage, S, eps2, meta.z, JB, JE);                                         \
        }                                                                                                             \
    } else if (masked) {                                                                                              \
        IState<NP> T = S;                                                                                             \
        tile_f32_masked<KID, MODE, NP, JUF>(stage, &T, eps2, jg0, ig0, a.n_j, jb, je);                                \
        S = T;                                                                                                        \
    } else if (um && uc) {                                                                                            \
        m_fold = meta.y;                                                                                              \
        tile_f32<KID, MODE, false, true, true, NP, JUF>(stage, S, eps2, meta.x, jg0, ig0, a.n_j, JB, JE);             \
    } else if (um) {                                                                                                  \
        m_fold = meta.y;                                                                                              \
        tile_f32<KID, MODE, false, true, false, NP, JUF>(stage, S, eps2, 0.f, jg0, ig0, a.n_j, JB, JE);               \
    } else {                                                                                                          \
        tile_f32<KID, MODE, false, false, false, NP, JUF>(stage, S, eps2, 0.f, jg0, ig0, a.n_j, JB, JE);              \
    }
            if constexpr (!SUB) {
                NB_TILE_DISPATCH_F32(0, TJ)
            } else {
                NB_TILE_DISPATCH_F32(jb, je)
            }
#undef NB_TILE_DISPATCH_F32

            // fold this tile's f32 partials (x tile mass) into the running sums
            const run_t mf = static_cast<run_t>(m_fold);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                auto fold = [&](int comp, f2 v) {
                    float lo, hi;
                    upk(v, lo, hi);
                    run_t* r0 = &run[((comp * 2 * NP) + 2 * q) * THREADS + tid];
                    run_t* r1 = &run[((comp * 2 * NP) + 2 * q + 1) * THREADS + tid];
                    *r0 = fma(static_cast<run_t>(lo), mf, *r0);
                    *r1 = fma(static_cast<run_t>(hi), mf, *r1);
                };
                if (MODE & NB_MODE_FORCE) {
                    fold(0, S.ax[q]);
                    fold(1, S.ay[q]);
                    fold(2, S.az[q]);
                }
                if (MODE & NB_MODE_POT) fold(3, S.sp[q]);
                S.ax[q] = S.ay[q] = S.az[q] = S.sp[q] = 0ull;
            }
            if (DEC) {   // release the stage; the last warp out refills it with the tile STAGES ahead
                __syncwarp();
                if ((tid & 31) == 0) {
                    const uint32_t pending = mbar_arrive_pending(&empty_bar[k % STAGES]);
                    if (pending == 1 && k + STAGES < n_units) issue(k + STAGES);
                }
            }
        }

        // ---- write this segment's partial sums ----
        const bool is_full = (jl0 == 0 && jl1 == upb);
        const size_t base  = is_full ? static_cast<size_t>(ib) * IB
                                     : (static_cast<size_t>(cta) * 2 + (u == u0 ? 0 : 1)) * IB;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const size_t li = base + 2 * (tid + THREADS * q);
            auto R = [&](int comp, int half) { return run[((comp * 2 * NP) + 2 * q + half) * THREADS + tid]; };
            if (ACC) {
                double4* dst = reinterpret_cast<double4*>(is_full ? a.full : a.part) + li;
                dst[0] = make_double4(R(0, 0), R(1, 0), R(2, 0), R(3, 0));
                dst[1] = make_double4(R(0, 1), R(1, 1), R(2, 1), R(3, 1));
            } else {
                float4* dst = reinterpret_cast<float4*>(is_full ? a.full : a.part) + li;
                dst[0] = make_float4(R(0, 0), R(1, 0), R(2, 0), R(3, 0));
                dst[1] = make_float4(R(0, 1), R(1, 1), R(2, 1), R(3, 1));
            }
        }
        u += jl1 - jl0;
    }
}

}  // namespace nb
