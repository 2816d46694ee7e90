// nb_sym_f32.cuh — pair-symmetric FP32 sweep (NB_PREC_F32 / NB_PREC_F32_COMP), the large-N path.
//
// Same job as nb_sweep_f32.cuh (replaces nbody_forces_kernel_float4 / ..._kahan_kernel_float4 /
// nbody_potential_kernel_float4 / ..._kahan..., reference cuda_kernels.py:90-179, :254-366, :439-518,
// :591-677) but every unordered pair is evaluated ONCE and applied to both particles (see nb_sym.cuh).
//
// Data flow inside a CTA (128 threads = 4 warps, one (i-block x j-tile) unit at a time):
//   * i side: each thread owns NP = 4 packed pairs of i-particles (8 i, 1024 per CTA) in 64-bit registers,
//     exactly like the one-sided sweep; partial sums over one tile are folded into float running sums in
//     shared memory and those are added to the double workspace every SYM_FLUSH_TILES tiles (blocked
//     summation: 256-term float partials, 8-term float second level, double above that);
//   * j side: the tile (TMA bulk copy into the ring) is processed in rounds of 32 j-particles, ONE PER LANE.
//     In step s of a round lane l works on particle (l + s) mod 32: it reads that particle's record with a
//     conflict-free LDS, evaluates it against its 8 i-particles with packed FADD2/FFMA2/FMUL2, accumulates
//     +K dx into its own a_i and into the particle's travelling accumulator, and hands that accumulator to
//     lane l-1 with SHFL (a systolic rotation).  After 32 steps every particle has met all 256 i-particles of
//     the warp and its accumulator is back in lane (j mod 32), which stores it to the warp's slice of a
//     shared buffer.  After the tile the four slices are added in double and added atomically to ws.
//   * tiles that overlap the CTA's own i-block ("diagonal") run the one-sided masked path of
//     nb_sweep_f32.cuh (self pairs excluded) and contribute to the i side only.
// Bound: FP32 pipe, 14 (force) / 16 (force+potential) packed instructions + 2 MUFU per 128 interactions.
#pragma once
#include "nb_sweep_f32.cuh"
#include "nb_sym.cuh"

namespace nb {

#ifndef NB_SYM_NP
#define NB_SYM_NP 4
#endif
constexpr int SYM_NP          = NB_SYM_NP;               // packed i-pairs per thread
constexpr int SYM_IB_F32      = THREADS * 2 * SYM_NP;    // 1024 i-particles per CTA
constexpr int SYM_TPB_F32     = SYM_IB_F32 / TJ;         // 4 tiles per i-block
constexpr int SYM_FLUSH_TILES = 8;                       // i-side: float running sums -> double workspace
#ifndef NB_SYM_UNROLL
#define NB_SYM_UNROLL 8
#endif
constexpr int SYM_UNROLL = NB_SYM_UNROLL;                 // steps per iteration of the hot loop
// The spline's force and force+potential sweeps use their own non-pipelined loop (sym_tile_f32_spline: 19.75 / 22.9 ms at
// N = 262144 against 21.8-22.4 / 22.9-23.8 ms in the pipelined loop at 1, 2 or 4 steps per iteration); only its
// potential-only sweep is faster pipelined (12.25 against 13.7 ms, 4 steps per iteration).
constexpr int SYM_UNROLL_SPLINE = 4;

__device__ __forceinline__ f2 shfl2(const f2 v, const int src) {
    float lo, hi;
    upk(v, lo, hi);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return pk(lo, hi);
}
__device__ __forceinline__ float sum2(const f2 v) {
    float lo, hi;
    upk(v, lo, hi);
    return lo + hi;
}

// the travelling accumulators of one j-particle: packed (contributions of the even i, of the odd i)
struct JAcc32 {
    f2 x, y, z, p;
};

// exact cubic-spline evaluation of one scalar pair, both sides (the rare r < h case; reference
// cuda_kernels.py:61-87 / :405-436 restated in spline_inside)
template <int MODE>
__device__ __forceinline__ void sym_spline_exact(const float xj, const float yj, const float zj, const float mj,
                                                 const float cj, const float xi, const float yi, const float zi,
                                                 const float mi, const float ci, float& ax, float& ay, float& az,
                                                 float& sp, float& jx, float& jy, float& jz, float& jp,
                                                 const float eps2) {
    const float dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const float d  = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, eps2)));
    const float h2 = fmaxf(cj, ci);
    float w, p;
    const float ri = rsqrt_approx(d);
    if (d < h2) {
        spline_inside<float>(d, ri, rsqrt_approx(h2), w, p);
    } else {
        p = ri;
        w = ri * (ri * ri);
    }
    if (MODE & NB_MODE_FORCE) {
        const float fi = mj * w, fj = mi * w;
        ax = fmaf(fi, dx, ax);
        ay = fmaf(fi, dy, ay);
        az = fmaf(fi, dz, az);
        jx = fmaf(fj, dx, jx);
        jy = fmaf(fj, dy, jy);
        jz = fmaf(fj, dz, jz);
    }
    if (MODE & NB_MODE_POT) {
        sp = fmaf(mj, p, sp);
        jp = fmaf(mi, p, jp);
    }
}

// Hot per-thread i-state: positions and per-tile partial sums only.  The softening terms and masses of the
// i-particles are needed by the rare paths only (general tiles, diagonal tiles, exact spline steps) and are
// re-read from the packed records there, so they do not occupy registers in the hot loop.
template <int NP>
struct SymHot32 {
    f2 x[NP], y[NP], z[NP];
    f2 ax[NP], ay[NP], az[NP], sp[NP];
};

// thread state handed to the non-inlined cold paths (kept out of the hot loop's register allocation)
template <int NP>
struct SymCold32 {
    IState<NP> S;
    f2 m[NP];
    JAcc32 J;
};

// i-particle masses and softening terms of this thread's pairs, straight from the packed records
// (dummy i beyond the i-range: massless)
template <int NP>
__device__ __forceinline__ void sym_load_mc(const unsigned char* __restrict__ packed, const long long ig0,
                                            const long long i_end, SymCold32<NP>& T) {
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const long long ig = ig0 + 2LL * THREADS * q;
        T.m[q] = 0ull;
        T.S.c0[q] = T.S.c1[q] = 0.f;
        if (ig < i_end) {
            const long long t = ig / TJ;
            const int kk      = static_cast<int>(ig - t * TJ);
            const float* tb   = reinterpret_cast<const float*>(packed + static_cast<size_t>(t) * TILE_BYTES_F32);
            T.m[q] = *reinterpret_cast<const f2*>(tb + COMP_M * TJ + kk);
            const float2 c2 = *reinterpret_cast<const float2*>(tb + COMP_C * TJ + kk);
            T.S.c0[q] = c2.x;
            T.S.c1[q] = c2.y;
        }
    }
}
template <int NP>
__device__ __forceinline__ void sym_hot_to_cold(const SymHot32<NP>& H, IState<NP>& S) {
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        S.x[q] = H.x[q]; S.y[q] = H.y[q]; S.z[q] = H.z[q];
        S.ax[q] = H.ax[q]; S.ay[q] = H.ay[q]; S.az[q] = H.az[q]; S.sp[q] = H.sp[q];
    }
}
template <int NP>
__device__ __forceinline__ void sym_cold_to_hot(const IState<NP>& S, SymHot32<NP>& H) {
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        H.ax[q] = S.ax[q]; H.ay[q] = S.ay[q]; H.az[q] = S.az[q]; H.sp[q] = S.sp[q];
    }
}

// ---------------------------------------------------------------------------
// one symmetric j-tile against this thread's NP pairs; jw = this warp's slice [4][TJ] of the j-side buffer.
//   ST = SymHot32 (U: uniform fast path) or IState (general path; ci0/ci1/mi = per-particle terms)
// The 256 steps of a tile (8 rounds x 32 rotation steps) form ONE software-pipelined loop: iteration t
// finishes step t (kernel factors + accumulates from the separations and reciprocal square roots computed one
// iteration earlier), rotates the j-accumulators, and then computes the separations, r^2 and MUFU.RSQ of step
// t+1.  Neither the MUFU latency nor the LDS latency (records are fetched two steps ahead) nor the SHFL
// latency (the rotation overlaps the next step's separations) sits on a dependency chain inside an
// iteration — that matters because only 4 warps share an SM sub-partition at 128 registers per thread.
// ---------------------------------------------------------------------------
//   SUB: the unit is a part of the tile, steps [t0, t1) (multiples of 32: whole rotation rounds); !SUB: [0, TJ)
template <int KID, int MODE, bool U, int NP, typename ST, bool SUB>
__device__ __forceinline__ void sym_tile_f32(const unsigned char* __restrict__ stage, ST& S, const f2* mi,
                                             const float* ci0, const float* ci1, const float cmax_i,
                                             const float c_pair, const float eps2, float* __restrict__ jw,
                                             const int lane, const int t0_in, const int t1_in) {
    const int t0 = SUB ? t0_in : 0, t1 = SUB ? t1_in : TJ;
    const float* __restrict__ sX = reinterpret_cast<const float*>(stage) + COMP_X * TJ;
    const float* __restrict__ sY = reinterpret_cast<const float*>(stage) + COMP_Y * TJ;
    const float* __restrict__ sZ = reinterpret_cast<const float*>(stage) + COMP_Z * TJ;
    const float* __restrict__ sM = reinterpret_cast<const float*>(stage) + COMP_M * TJ;
    const float* __restrict__ sC = reinterpret_cast<const float*>(stage) + COMP_C * TJ;
    constexpr bool kUsesC  = (KID == NB_KERNEL_PLUMMER || KID == NB_KERNEL_DEHNEN_K1 || KID == NB_KERNEL_DEHNEN_K2);
    constexpr bool kDehnen = (KID == NB_KERNEL_DEHNEN_K1 || KID == NB_KERNEL_DEHNEN_K2);
    const f2 cc   = splat(c_pair);
    const f2 ceps = splat(eps2);
    const int src = (lane + 1) & 31;

    // record of step t: particle 32 (t / 32) + (lane + t) mod 32 of the tile.  Its BYTE offset inside a component
    // array is ((4 lane + 4 t) & 124) | (4 t & 0x380): the first term comes from a running per-thread counter (one add
    // with an immediate + one and-or per step), the second is uniform and changes once per rotation round.
    auto rec_off = [&](int t) { return ((4 * lane + 4 * t) & 124) | ((4 * t) & (4 * (TJ - 32))); };
    auto ldf = [&](const float* base, int off) { return *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(base) + off); };
    float xj, yj, zj, mj = 1.f, cj = c_pair;     // step t
    float xn, yn, zn, mn = 1.f, cn = c_pair;     // step t + 1 (prefetched)
    {
        const int j0 = rec_off(t0), j1 = rec_off(t0 + 1);
        xj = ldf(sX, j0); yj = ldf(sY, j0); zj = ldf(sZ, j0);
        xn = ldf(sX, j1); yn = ldf(sY, j1); zn = ldf(sZ, j1);
        if (!U) {
            mj = ldf(sM, j0); cj = ldf(sC, j0);
            mn = ldf(sM, j1); cn = ldf(sC, j1);
        }
    }
    f2 dx[NP], dy[NP], dz[NP], rinv[NP], h2[NP];
    bool exact = false;
    // separations, softening term, r^2 and reciprocal square roots of the step whose record is (xj.., cj)
    auto prepare = [&]() {
        float dmin = 3.0e38f;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            dx[q] = sub2(splat(xj), S.x[q]);
            dy[q] = sub2(splat(yj), S.y[q]);
            dz[q] = sub2(splat(zj), S.z[q]);
            f2 c;
            if (KID == NB_KERNEL_NEWTONIAN) {
                h2[q] = c = ceps;
            } else {
                h2[q] = U ? cc : pk(fmaxf(cj, ci0[q]), fmaxf(cj, ci1[q]));
                c     = kUsesC ? h2[q] : ceps;
            }
            const f2 d = fma2(dx[q], dx[q], fma2(dy[q], dy[q], fma2(dz[q], dz[q], c)));
            float d0, d1;
            upk(d, d0, d1);
            if (KID == NB_KERNEL_SPLINE) dmin = fminf(dmin, fminf(d0, d1));
            rinv[q] = pk(rsqrt_approx(d0), rsqrt_approx(d1));
        }
        if (KID == NB_KERNEL_SPLINE) exact = dmin < (U ? c_pair : fmaxf(cj, cmax_i));
    };
    prepare();

    JAcc32 J;
    J.x = J.y = J.z = J.p = 0ull;
    constexpr int kUnroll = SUB ? 2 : ((KID == NB_KERNEL_SPLINE) ? SYM_UNROLL_SPLINE : SYM_UNROLL);
    static_assert(32 % kUnroll == 0 && kUnroll >= 2, "chunks must tile a rotation round; records are fetched 2 steps ahead");
    int lane_t4 = 4 * lane + 4 * t0;     // 4 (lane + tb), unmasked
#pragma unroll 1
    for (int tb = t0; tb < t1; tb += kUnroll, lane_t4 += 4 * kUnroll) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int t = tb + u;
        if (KID == NB_KERNEL_SPLINE && exact) {
            // exact step, in place (see sym_tile_f32_spline)
            float jx0, jx1, jy0, jy1, jz0, jz1, jp0, jp1;
            upk(J.x, jx0, jx1); upk(J.y, jy0, jy1); upk(J.z, jz0, jz1); upk(J.p, jp0, jp1);
            const float mje = U ? 1.f : mj;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                float x0, x1, y0, y1, z0, z1, m0 = 1.f, m1 = 1.f, ax0, ax1, ay0, ay1, az0, az1, sp0, sp1;
                upk(S.x[q], x0, x1); upk(S.y[q], y0, y1); upk(S.z[q], z0, z1);
                upk(S.ax[q], ax0, ax1); upk(S.ay[q], ay0, ay1); upk(S.az[q], az0, az1); upk(S.sp[q], sp0, sp1);
                if (!U) upk(mi[q], m0, m1);
                sym_spline_exact<MODE>(xj, yj, zj, mje, cj, x0, y0, z0, m0, U ? c_pair : ci0[q], ax0, ay0, az0, sp0,
                                       jx0, jy0, jz0, jp0, eps2);
                sym_spline_exact<MODE>(xj, yj, zj, mje, cj, x1, y1, z1, m1, U ? c_pair : ci1[q], ax1, ay1, az1, sp1,
                                       jx1, jy1, jz1, jp1, eps2);
                S.ax[q] = pk(ax0, ax1); S.ay[q] = pk(ay0, ay1); S.az[q] = pk(az0, az1); S.sp[q] = pk(sp0, sp1);
            }
            J.x = pk(jx0, jx1); J.y = pk(jy0, jy1); J.z = pk(jz0, jz1); J.p = pk(jp0, jp1);
        } else {
            // kernel factors (mass-less: w ~ 1/r^3, p ~ 1/r) applied to both sides
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const f2 rinv2 = mul2(rinv[q], rinv[q]);
                f2 w = 0ull, p = rinv[q];
                if (kDehnen) {
                    const f2 tt = mul2(add2(h2[q], splat(-eps2)), rinv2);   // the terms use the bare h^2
                    if (MODE & NB_MODE_POT) {
                        const f2 pf = (KID == NB_KERNEL_DEHNEN_K1)
                                          ? fma2(splat(0.5f), tt, splat(1.f))
                                          : fma2(tt, fma2(splat(0.375f), tt, splat(0.5f)), splat(1.f));
                        p = mul2(rinv[q], pf);
                    }
                    if (MODE & NB_MODE_FORCE) {
                        const f2 ff = (KID == NB_KERNEL_DEHNEN_K1)
                                          ? fma2(splat(1.5f), tt, splat(1.f))
                                          : fma2(tt, fma2(splat(3.75f), tt, splat(1.5f)), splat(1.f));
                        w = mul2(mul2(rinv[q], rinv2), ff);
                    }
                } else if (MODE & NB_MODE_FORCE) {
                    w = mul2(rinv[q], rinv2);
                }
                if (U) {
                    if (MODE & NB_MODE_POT) {
                        S.sp[q] = add2(S.sp[q], p);
                        J.p     = add2(J.p, p);
                    }
                    if (MODE & NB_MODE_FORCE) {
                        S.ax[q] = fma2(w, dx[q], S.ax[q]);
                        S.ay[q] = fma2(w, dy[q], S.ay[q]);
                        S.az[q] = fma2(w, dz[q], S.az[q]);
                        J.x     = fma2(w, dx[q], J.x);
                        J.y     = fma2(w, dy[q], J.y);
                        J.z     = fma2(w, dz[q], J.z);
                    }
                } else {
                    const f2 mjs = splat(mj);
                    if (MODE & NB_MODE_POT) {
                        S.sp[q] = fma2(mjs, p, S.sp[q]);
                        J.p     = fma2(mi[q], p, J.p);
                    }
                    if (MODE & NB_MODE_FORCE) {
                        const f2 fi = mul2(mjs, w), fj = mul2(mi[q], w);
                        S.ax[q] = fma2(fi, dx[q], S.ax[q]);
                        S.ay[q] = fma2(fi, dy[q], S.ay[q]);
                        S.az[q] = fma2(fi, dz[q], S.az[q]);
                        J.x     = fma2(fj, dx[q], J.x);
                        J.y     = fma2(fj, dy[q], J.y);
                        J.z     = fma2(fj, dz[q], J.z);
                    }
                }
            }
        }
        // systolic rotation: the accumulators follow their j-particle to the lane that meets it next
        if (MODE & NB_MODE_FORCE) {
            J.x = shfl2(J.x, src);
            J.y = shfl2(J.y, src);
            J.z = shfl2(J.z, src);
        }
        if (MODE & NB_MODE_POT) J.p = shfl2(J.p, src);
        if (u == kUnroll - 1 && (tb & 31) == 32 - kUnroll) {   // round complete: lane l holds the sums of particle (t - 31) + l
            const int jb = t - 31;
            if (MODE & NB_MODE_FORCE) {
                jw[0 * TJ + jb + lane] = sum2(J.x);
                jw[1 * TJ + jb + lane] = sum2(J.y);
                jw[2 * TJ + jb + lane] = sum2(J.z);
            }
            if (MODE & NB_MODE_POT) jw[3 * TJ + jb + lane] = sum2(J.p);
            J.x = J.y = J.z = J.p = 0ull;
        }
        // advance the record pipeline and prepare step t + 1 (the last iteration prepares a step that is not used)
        xj = xn; yj = yn; zj = zn; mj = mn; cj = cn;
        {
            // step t + 2 = tb + kUnroll * ((u + 2) / kUnroll) + (u + 2) % kUnroll; tb is a multiple of kUnroll (<= 32), so
            // the round bits of t + 2 are those of the chunk start
            const int ua    = u + 2;
            const int chunk = tb + (ua / kUnroll) * kUnroll;
            const int jn    = (lane_t4 + 4 * ua) & 124;
            const int gr    = chunk & (TJ - 32);          // uniform: folded into the base address
            xn = ldf(sX + gr, jn); yn = ldf(sY + gr, jn); zn = ldf(sZ + gr, jn);
            if (!U) {
                mn = ldf(sM + gr, jn);
                cn = ldf(sC + gr, jn);
            }
        }
        prepare();
      }
    }
}

// ---------------------------------------------------------------------------
// Cubic spline (the API default, sim.py:419): the same tile sweep, NOT software-pipelined.  A step computes
// the separations and r^2 of its NP pairs and their minimum (one FMNMX3 per packed pair); the spline is
// exactly Newtonian for r >= h (cuda_kernels.py:64-67), so unless that minimum lies below the largest
// softening term that can apply the step finishes on the branch-free Newtonian path; otherwise (rare) the
// whole step is redone in place with the reference's piecewise polynomial (sym_spline_exact).  Measured best
// with one step per iteration and the three phases in program order (round-2 experiments, N = 262144 force:
// this loop with the exact step in place 19.75 ms; with the exact step behind a call on a state copy 21.4 ms;
// with the exact step moved out of line through workspace atomics 24.8 ms; the software-pipelined loop
// 21.8-22.4 ms with the in-place step, 23.3-24.8 ms with the out-of-line ones).
// ---------------------------------------------------------------------------
template <int MODE, bool U, int NP, typename ST, bool SUB>
__device__ __forceinline__ void sym_tile_f32_spline(const unsigned char* __restrict__ stage, ST& S, const f2* mi,
                                                    const float* ci0, const float* ci1, const float cmax_i,
                                                    const float c_pair, const float eps2, float* __restrict__ jw,
                                                    const int lane, const int t0_in, const int t1_in) {
    const int r0 = SUB ? (t0_in >> 5) : 0, r1 = SUB ? (t1_in >> 5) : TJ / 32;   // rotation rounds of this unit
    const float* __restrict__ sX = reinterpret_cast<const float*>(stage) + COMP_X * TJ;
    const float* __restrict__ sY = reinterpret_cast<const float*>(stage) + COMP_Y * TJ;
    const float* __restrict__ sZ = reinterpret_cast<const float*>(stage) + COMP_Z * TJ;
    const float* __restrict__ sM = reinterpret_cast<const float*>(stage) + COMP_M * TJ;
    const float* __restrict__ sC = reinterpret_cast<const float*>(stage) + COMP_C * TJ;
    const f2 ceps = splat(eps2);
    const int src = (lane + 1) & 31;
    // steps per loop iteration: measured at N = 262144 — force 19.8 ms (1 step) / 21.1 ms (2); force+potential 22.9 / 22.0
    constexpr int kSteps = (MODE == NB_MODE_BOTH) ? 2 : 1;

#pragma unroll 1
    for (int rho = r0; rho < r1; ++rho) {
        JAcc32 J;
        J.x = J.y = J.z = J.p = 0ull;
        const int jb = 32 * rho;
        // the record of the NEXT step is fetched one step ahead (LDS latency off the critical path)
        float xn = sX[jb + lane], yn = sY[jb + lane], zn = sZ[jb + lane], mn = 1.f, cn = c_pair;
        if (!U) {
            mn = sM[jb + lane];
            cn = sC[jb + lane];
        }
#pragma unroll kSteps
        for (int s = 0; s < 32; ++s) {
            const float xj = xn, yj = yn, zj = zn, mj = mn, cj = cn;
            {
                // (a running byte-offset counter as in sym_tile_f32 costs more MOVs than it saves address arithmetic in this
                // non-unrolled loop: 97 against 91 instructions on the Newtonian path of a step)
                const int jj = jb + ((lane + s + 1) & 31);   // s = 31 wraps to this lane's first record (unused)
                xn = sX[jj];
                yn = sY[jj];
                zn = sZ[jj];
                if (!U) {
                    mn = sM[jj];
                    cn = sC[jj];
                }
            }
            f2 dx[NP], dy[NP], dz[NP], d[NP];
            float dmin = 3.0e38f;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                dx[q] = sub2(splat(xj), S.x[q]);
                dy[q] = sub2(splat(yj), S.y[q]);
                dz[q] = sub2(splat(zj), S.z[q]);
                d[q]  = fma2(dx[q], dx[q], fma2(dy[q], dy[q], fma2(dz[q], dz[q], ceps)));
                float d0, d1;
                upk(d[q], d0, d1);
                dmin = fminf(dmin, fminf(d0, d1));
            }
            if (dmin < (U ? c_pair : fmaxf(cj, cmax_i))) {
                // exact step, in place: every pair of the step with the reference's piecewise polynomial.  (Moving
                // this behind a call on a state copy was tried first: 8 % slower at h/a = 0.01 and 2-4x slower once a few
                // percent of the steps take this path, h/a >= 0.2 — profiles/r02_spline_rate_vs_softening.md.)
                float jx0, jx1, jy0, jy1, jz0, jz1, jp0, jp1;
                upk(J.x, jx0, jx1); upk(J.y, jy0, jy1); upk(J.z, jz0, jz1); upk(J.p, jp0, jp1);
                const float mje = U ? 1.f : mj;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    float x0, x1, y0, y1, z0, z1, m0 = 1.f, m1 = 1.f, ax0, ax1, ay0, ay1, az0, az1, sp0, sp1;
                    upk(S.x[q], x0, x1); upk(S.y[q], y0, y1); upk(S.z[q], z0, z1);
                    upk(S.ax[q], ax0, ax1); upk(S.ay[q], ay0, ay1); upk(S.az[q], az0, az1); upk(S.sp[q], sp0, sp1);
                    if (!U) upk(mi[q], m0, m1);
                    sym_spline_exact<MODE>(xj, yj, zj, mje, cj, x0, y0, z0, m0, U ? c_pair : ci0[q], ax0, ay0, az0, sp0,
                                           jx0, jy0, jz0, jp0, eps2);
                    sym_spline_exact<MODE>(xj, yj, zj, mje, cj, x1, y1, z1, m1, U ? c_pair : ci1[q], ax1, ay1, az1, sp1,
                                           jx1, jy1, jz1, jp1, eps2);
                    S.ax[q] = pk(ax0, ax1); S.ay[q] = pk(ay0, ay1); S.az[q] = pk(az0, az1); S.sp[q] = pk(sp0, sp1);
                }
                J.x = pk(jx0, jx1); J.y = pk(jy0, jy1); J.z = pk(jz0, jz1); J.p = pk(jp0, jp1);
            } else {
                f2 rinv[NP];
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    float d0, d1;
                    upk(d[q], d0, d1);
                    rinv[q] = pk(rsqrt_approx(d0), rsqrt_approx(d1));
                }
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const f2 w = (MODE & NB_MODE_FORCE) ? mul2(rinv[q], mul2(rinv[q], rinv[q])) : 0ull;
                    if (U) {
                        if (MODE & NB_MODE_POT) {
                            S.sp[q] = add2(S.sp[q], rinv[q]);
                            J.p     = add2(J.p, rinv[q]);
                        }
                        if (MODE & NB_MODE_FORCE) {
                            S.ax[q] = fma2(w, dx[q], S.ax[q]);
                            S.ay[q] = fma2(w, dy[q], S.ay[q]);
                            S.az[q] = fma2(w, dz[q], S.az[q]);
                            J.x     = fma2(w, dx[q], J.x);
                            J.y     = fma2(w, dy[q], J.y);
                            J.z     = fma2(w, dz[q], J.z);
                        }
                    } else {
                        const f2 mjs = splat(mj);
                        if (MODE & NB_MODE_POT) {
                            S.sp[q] = fma2(mjs, rinv[q], S.sp[q]);
                            J.p     = fma2(mi[q], rinv[q], J.p);
                        }
                        if (MODE & NB_MODE_FORCE) {
                            const f2 fi = mul2(mjs, w), fj = mul2(mi[q], w);
                            S.ax[q] = fma2(fi, dx[q], S.ax[q]);
                            S.ay[q] = fma2(fi, dy[q], S.ay[q]);
                            S.az[q] = fma2(fi, dz[q], S.az[q]);
                            J.x     = fma2(fj, dx[q], J.x);
                            J.y     = fma2(fj, dy[q], J.y);
                            J.z     = fma2(fj, dz[q], J.z);
                        }
                    }
                }
            }
            if (MODE & NB_MODE_FORCE) {
                J.x = shfl2(J.x, src);
                J.y = shfl2(J.y, src);
                J.z = shfl2(J.z, src);
            }
            if (MODE & NB_MODE_POT) J.p = shfl2(J.p, src);
        }
        if (MODE & NB_MODE_FORCE) {
            jw[0 * TJ + jb + lane] = sum2(J.x);
            jw[1 * TJ + jb + lane] = sum2(J.y);
            jw[2 * TJ + jb + lane] = sum2(J.z);
        }
        if (MODE & NB_MODE_POT) jw[3 * TJ + jb + lane] = sum2(J.p);
    }
}

// kernel-id dispatch of the two loop styles
template <int KID, int MODE, bool U, int NP, typename ST, bool SUB>
__device__ __forceinline__ void sym_tile_f32_any(const unsigned char* __restrict__ stage, ST& S, const f2* mi,
                                                 const float* ci0, const float* ci1, const float cmax_i,
                                                 const float c_pair, const float eps2, float* __restrict__ jw,
                                                 const int lane, const int t0, const int t1) {
    if constexpr (KID == NB_KERNEL_SPLINE && MODE != NB_MODE_POT)
        sym_tile_f32_spline<MODE, U, NP, ST, SUB>(stage, S, mi, ci0, ci1, cmax_i, c_pair, eps2, jw, lane, t0, t1);
    else
        sym_tile_f32<KID, MODE, U, NP, ST, SUB>(stage, S, mi, ci0, ci1, cmax_i, c_pair, eps2, jw, lane, t0, t1);
}

// general (per-particle mass / softening) symmetric tile: rare (species boundaries, per-particle softening
// arrays, padding) -> non-inlined, re-reads the i-particles' masses and softening terms from the packed
// records and works on a copy of the hot state
template <int KID, int MODE, int NP, bool SUB>
__device__ __noinline__ void sym_tile_f32_general(const unsigned char* stage, SymHot32<NP>* H,
                                                  const unsigned char* packed, const long long ig0,
                                                  const long long i_end, const float eps2, float* jw,
                                                  const int lane, const int t0, const int t1) {
    SymCold32<NP> T;
    sym_hot_to_cold<NP>(*H, T.S);
    sym_load_mc<NP>(packed, ig0, i_end, T);
    float cmax_i = 0.f;
#pragma unroll
    for (int q = 0; q < NP; ++q) cmax_i = fmaxf(cmax_i, fmaxf(T.S.c0[q], T.S.c1[q]));
    sym_tile_f32_any<KID, MODE, false, NP, IState<NP>, SUB>(stage, T.S, T.m, T.S.c0, T.S.c1, cmax_i, 0.f, eps2, jw, lane,
                                                            t0, t1);
    sym_cold_to_hot<NP>(T.S, *H);
}

// diagonal tile (overlaps the CTA's own i-block): one-sided masked path of nb_sweep_f32.cuh
template <int KID, int MODE, int NP>
__device__ __noinline__ void sym_tile_f32_diag(const unsigned char* stage, SymHot32<NP>* H,
                                               const unsigned char* packed, const long long ig0,
                                               const long long i_end, const float eps2, const long long jg0,
                                               const long long n_j, const int t0, const int t1) {
    SymCold32<NP> T;
    sym_hot_to_cold<NP>(*H, T.S);
    sym_load_mc<NP>(packed, ig0, i_end, T);
    tile_f32<KID, MODE, true, false, false, NP, 4>(stage, T.S, eps2, 0.f, jg0, ig0, n_j, t0, t1);
    sym_cold_to_hot<NP>(T.S, *H);
}

template <int NP>
struct SymF32Smem {
    static constexpr int kStageBytes = STAGES * TILE_BYTES_F32;
    static constexpr int kBarOff     = (kStageBytes + 63) & ~63;
    static constexpr int kRunOff     = kBarOff + 128;   // mbarriers + the TMA producer's unit cursor
    static constexpr int kRunBytes   = 4 * 2 * NP * THREADS * 4;          // float running sums [4][2NP][THREADS]
    static constexpr int kJOff       = kRunOff + kRunBytes;
    static constexpr int kJBytes     = (THREADS / 32) * 4 * TJ * 4;       // j-side slices [warps][4][TJ]
    static constexpr int kTotal      = kJOff + kJBytes;
};

//   SUB = true: sub-tile work units (a.split_log2 > 0, small problems), a separate instantiation so that the
//   whole-tile kernels keep constant loop bounds
template <int KID, int MODE, int NP, int MINB, bool SUB = false>
__global__ void __launch_bounds__(THREADS, MINB) sym_f32_kernel(const SymArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    using L = SymF32Smem<NP>;
    constexpr int IB  = THREADS * 2 * NP;
    constexpr int TPB = IB / TJ;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    float* run         = reinterpret_cast<float*>(smem + L::kRunOff);   // [4][2*NP][THREADS]
    float* jbuf        = reinterpret_cast<float*>(smem + L::kJOff);     // [warps][4][TJ]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, grid = gridDim.x;
    const SymRuns R(a, cta, grid);            // this CTA's units: round-robin chunks + a stream-K tail share
    const int n_units = static_cast<int>(R.n_units());
    if (n_units <= 0) return;
    const float eps2 = static_cast<float>(a.eps2);
    const long long i_end = static_cast<long long>(a.it0 + a.nT) * TJ;   // first slot past the i-range

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // unit cursor of the consumer (all threads, registers) and of the TMA producer (thread 0 only, kept in
    // shared memory so that it costs no registers in the hot loop); the producer runs STAGES-1 units ahead
    SymCursor<TPB> cons;
    cons.open(a, R, 0);
    SymCursor<TPB>* prod = reinterpret_cast<SymCursor<TPB>*>(smem + L::kBarOff + 64);
    int issued = 0;
    auto issue_next = [&]() {
        prod->settle(a);
        bool dg;
        const int gt  = sym_tile<TPB>(a, prod->ib, prod->jl >> (SUB ? a.split_log2 : 0), dg);
        uint64_t* bar = &full_bar[issued % STAGES];
        mbar_expect_tx(bar, TILE_BYTES_F32);
        bulk_g2s(smem + (issued % STAGES) * TILE_BYTES_F32, a.packed + static_cast<size_t>(gt) * TILE_BYTES_F32, TILE_BYTES_F32, bar);
        ++issued;
        prod->advance(a, R, 1);
    };
    if (tid == 0) {
        prod->open(a, R, 0);
        for (int k = 0; k < STAGES - 1 && k < n_units; ++k) issue_next();
    }

    // j side of the previous symmetric tile: summed over the warps and added to ws after the next barrier
    bool pend = false;
    int pend_gt = 0, pend_t0 = 0, pend_t1 = TJ;
    float pend_scale = 1.f;
    auto flush_j = [&]() {
#pragma unroll
        for (int jj = (SUB ? pend_t0 : 0) + tid; jj < (SUB ? pend_t1 : TJ); jj += THREADS) {
            double* dst = a.ws + (static_cast<size_t>(pend_gt) * TJ + jj) * 4;
#pragma unroll
            for (int comp = 0; comp < 4; ++comp) {
                if (comp < 3 ? !(MODE & NB_MODE_FORCE) : !(MODE & NB_MODE_POT)) continue;
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < THREADS / 32; ++w) v += static_cast<double>(jbuf[(w * 4 + comp) * TJ + jj]);
                // a_j += -m_i K dx  (dx = r_j - r_i);  s_phi_j += +m_i / r
                atomicAdd(dst + comp, (comp < 3 ? -v : v) * static_cast<double>(pend_scale));
            }
        }
    };

    int k = 0;
    while (k < n_units) {
        cons.settle(a);
        const int ib  = cons.ib;
        const int jl0 = cons.jl;
        const int seg = static_cast<int>(min(static_cast<long long>(cons.cnt - cons.jl), cons.u_end - cons.u));

        // ---- this thread's i-pairs (read straight from the packed records) ----
        const int tile0          = a.it0 + TPB * ib;
        const long long ig_block = static_cast<long long>(tile0) * TJ;
        const long long ig0      = ig_block + 2LL * tid;
        SymHot32<NP> S;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const long long ig = ig0 + 2LL * THREADS * q;   // even; ig+1 is in the same tile
            // dummy i beyond the i-range: far away (and massless, see sym_load_mc) -> contributes nothing
            S.x[q] = S.y[q] = S.z[q] = splat(NB_PAD_POS_F32);
            if (ig < i_end) {
                const long long t = ig / TJ;
                const int kk      = static_cast<int>(ig - t * TJ);
                const float* tb   = reinterpret_cast<const float*>(a.packed + static_cast<size_t>(t) * TILE_BYTES_F32);
                S.x[q] = *reinterpret_cast<const f2*>(tb + COMP_X * TJ + kk);
                S.y[q] = *reinterpret_cast<const f2*>(tb + COMP_Y * TJ + kk);
                S.z[q] = *reinterpret_cast<const f2*>(tb + COMP_Z * TJ + kk);
            }
            S.ax[q] = S.ay[q] = S.az[q] = S.sp[q] = 0ull;
        }
        // is the whole i-block one species (same mass and softening, no padding, all TPB tiles present)?
        bool blk_uni = (tile0 + TPB <= a.it0 + a.nT);
        float cI = 0.f, mI = 0.f;
        if (blk_uni) {
#pragma unroll
            for (int t = 0; t < TPB; ++t) {
                const float4 meta = *reinterpret_cast<const float4*>(
                    a.packed + static_cast<size_t>(tile0 + t) * TILE_BYTES_F32 + 5 * TJ * 4);
                if (t == 0) {
                    cI = meta.x;
                    mI = meta.y;
                }
                blk_uni = blk_uni && (meta.x == cI) && (meta.y == mI);   // NaN (non-uniform tile) fails both
            }
        }
#pragma unroll
        for (int e = 0; e < 4 * 2 * NP; ++e) run[e * THREADS + tid] = 0.f;
        int tiles_in_run = 0;

        auto flush_i = [&]() {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const long long ig = ig0 + 2LL * THREADS * q + half;
                    double* dst = a.ws + static_cast<size_t>(ig) * 4;
#pragma unroll
                    for (int comp = 0; comp < 4; ++comp) {
                        if (comp < 3 ? !(MODE & NB_MODE_FORCE) : !(MODE & NB_MODE_POT)) continue;
                        float* r = &run[((comp * 2 * NP) + 2 * q + half) * THREADS + tid];
                        if (ig < i_end) atomicAdd(dst + comp, static_cast<double>(*r));
                        *r = 0.f;
                    }
                }
            }
        };

        for (int t = 0; t < seg; ++t, ++k) {
            const int jl = jl0 + t;
            __syncthreads();   // everyone is done with tile k-1: its stage may be refilled, its j-slices are complete
            if (tid == 0 && issued < n_units) issue_next();
            if (pend) {
                flush_j();
                __syncthreads();   // slices are free again
                pend = false;
            }
            mbar_wait(&full_bar[k % STAGES], (k / STAGES) & 1);

            bool diag;
            const int sl = SUB ? a.split_log2 : 0;
            const int gt = sym_tile<TPB>(a, ib, jl >> sl, diag);
            const int t0 = (jl & ((1 << sl) - 1)) * (TJ >> sl), t1 = t0 + (TJ >> sl);   // this unit's steps of the tile
            const unsigned char* stage = smem + (k % STAGES) * TILE_BYTES_F32;
            const float4 meta = *reinterpret_cast<const float4*>(stage + 5 * TJ * 4);   // {c_uni|NaN, m_uni|NaN, c_max, -}
            float m_fold = 1.f;
            if (diag) {
                sym_tile_f32_diag<KID, MODE, NP>(stage, &S, a.packed, ig0, i_end, eps2, static_cast<long long>(gt) * TJ,
                                                 a.n_j, SUB ? t0 : 0, SUB ? t1 : TJ);
            } else {
                float* jw = jbuf + warp * 4 * TJ;
                if (blk_uni && meta.x == meta.x && meta.y == meta.y) {
                    m_fold     = meta.y;
                    pend_scale = mI;
                    sym_tile_f32_any<KID, MODE, true, NP, SymHot32<NP>, SUB>(stage, S, nullptr, nullptr, nullptr, 0.f,
                                                                              fmaxf(cI, meta.x), eps2, jw, lane, t0, t1);
                } else {
                    pend_scale = 1.f;
                    sym_tile_f32_general<KID, MODE, NP, SUB>(stage, &S, a.packed, ig0, i_end, eps2, jw, lane, t0, t1);
                }
                pend    = true;
                pend_gt = gt;
                pend_t0 = t0;
                pend_t1 = t1;
            }

            // fold this tile's partial sums (x tile mass) into the float running sums
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                auto fold = [&](int comp, f2 v) {
                    float lo, hi;
                    upk(v, lo, hi);
                    float* r0 = &run[((comp * 2 * NP) + 2 * q) * THREADS + tid];
                    float* r1 = &run[((comp * 2 * NP) + 2 * q + 1) * THREADS + tid];
                    *r0 = fmaf(lo, m_fold, *r0);
                    *r1 = fmaf(hi, m_fold, *r1);
                };
                if (MODE & NB_MODE_FORCE) {
                    fold(0, S.ax[q]);
                    fold(1, S.ay[q]);
                    fold(2, S.az[q]);
                }
                if (MODE & NB_MODE_POT) fold(3, S.sp[q]);
                S.ax[q] = S.ay[q] = S.az[q] = S.sp[q] = 0ull;
            }
            if (++tiles_in_run == SYM_FLUSH_TILES) {
                flush_i();
                tiles_in_run = 0;
            }
        }
        if (tiles_in_run) flush_i();
        cons.advance(a, R, seg);
    }
    __syncthreads();
    if (pend) flush_j();
}

}  // namespace nb
