// nb_sweep_f64.cuh — the FP64 pair sweep (NB_PREC_F64).
//
// Replaces nbody_forces_kernel<double> / nbody_potential_kernel<double>
// (reference nbody_streams/cuda_kernels.py:755-864, :1299-1387).
//
// Same persistent stream-K / TMA-ring structure as the f32 sweep; the pair
// arithmetic is scalar DFMA with NI i-particles register-blocked per thread and
// a hand-rolled full-precision rsqrt (MUFU.RSQ64H seed + one third-order step,
// 5 FP64 ops; force-only sweeps take r^-3 directly from the seed, 6 ops instead
// of 5 + 2) instead of the library's longer sequence.  Bound: FP64 pipe.
#pragma once
#include "nb_sweep.cuh"

namespace nb {

//   UM : tile has one mass (sums accumulated unweighted, m applied in the per-tile fold)
//   UC : tile has one softening term; cmax = max(c_tile, c_i) precomputed per tile
template <int KID, int MODE, bool MASK, bool UM, bool UC>
__device__ __forceinline__ void pair_f64(const double xj, const double yj, const double zj, double mj,
                                         const double cj, const double xi, const double yi, const double zi,
                                         const double ci, const double cmax, double& ax, double& ay, double& az,
                                         double& sp, const double eps2, const bool valid) {
    const double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    double c, h2;
    if (KID == NB_KERNEL_NEWTONIAN) {
        c = h2 = eps2;
    } else if (KID == NB_KERNEL_SPLINE) {
        c  = eps2;
        h2 = UC ? cmax : fmax(cj, ci);
    } else {
        c = h2 = UC ? cmax : fmax(cj, ci);
    }
    if (UM) mj = 1.0;
    if (MASK && !valid) {   // self pair / padding: zero mass, harmless distance, never "inside h"
        mj = 0.0;
        c  = 1.0;
        if (KID == NB_KERNEL_SPLINE) h2 = 0.0;
    }
    double d = fma(dx, dx, fma(dy, dy, fma(dz, dz, c)));
    bool inside = false;
    if (KID == NB_KERNEL_SPLINE) {
        inside = (d < h2) && (!MASK || valid);
        d = fmax(d, h2);
    }
    if (MODE == NB_MODE_FORCE && (KID == NB_KERNEL_NEWTONIAN || KID == NB_KERNEL_PLUMMER)) {   // only r^-3 is needed
        const double w = rsqrt3_f64(d);
        const double f = (UM && !MASK) ? w : mj * w;
        ax = fma(f, dx, ax);
        ay = fma(f, dy, ay);
        az = fma(f, dz, az);
        return;
    }
    const double rinv = rsqrt_f64(d);
    const double mr   = (UM && !MASK) ? rinv : mj * rinv;
    if (KID == NB_KERNEL_DEHNEN_K1 || KID == NB_KERNEL_DEHNEN_K2) {
        const double rinv2 = rinv * rinv;
        const double t     = (h2 - eps2) * rinv2;   // c = h^2 + eps2 seeds the r^2 chain; the terms use bare h^2
        if (MODE & NB_MODE_POT) {
            const double pf = (KID == NB_KERNEL_DEHNEN_K1) ? fma(0.5, t, 1.0) : fma(t, fma(0.375, t, 0.5), 1.0);
            sp = fma(mr, pf, sp);
        }
        if (MODE & NB_MODE_FORCE) {
            const double ff = (KID == NB_KERNEL_DEHNEN_K1) ? fma(1.5, t, 1.0) : fma(t, fma(3.75, t, 1.5), 1.0);
            const double f  = (mr * rinv2) * ff;
            ax = fma(f, dx, ax);
            ay = fma(f, dy, ay);
            az = fma(f, dz, az);
        }
    } else {
        double f = 0.0;
        if (MODE & NB_MODE_FORCE) f = mr * (rinv * rinv);
        double p = mr;
        if (KID == NB_KERNEL_SPLINE) {
            if (inside) {   // exact Monaghan spline instead of the clamped Newtonian value
                const double dr = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
                double w, ph;
                spline_inside<double>(dr, rsqrt_f64(dr), rinv /* = 1/h: d was clamped to h2 */, w, ph);
                if (MODE & NB_MODE_FORCE) f = mj * w;
                if (MODE & NB_MODE_POT) p = mj * ph;
            }
        }
        if (MODE & NB_MODE_POT) sp += p;
        if (MODE & NB_MODE_FORCE) {
            ax = fma(f, dx, ax);
            ay = fma(f, dy, ay);
            az = fma(f, dz, az);
        }
    }
}

template <int NI>
struct IState64 {
    double x[NI], y[NI], z[NI], c[NI];
    double ax[NI], ay[NI], az[NI], sp[NI];
};

template <int KID, int MODE, bool MASK, bool UM, bool UC, int NI>
__device__ __forceinline__ void tile_f64(const unsigned char* __restrict__ stage, IState64<NI>& S,
                                         const double eps2, const double c_tile, const long long jg0,
                                         const long long ig0, const long long n_j, const int jb = 0,
                                         const int je = TJ) {
    const double* __restrict__ sX = reinterpret_cast<const double*>(stage) + COMP_X * TJ;
    const double* __restrict__ sY = reinterpret_cast<const double*>(stage) + COMP_Y * TJ;
    const double* __restrict__ sZ = reinterpret_cast<const double*>(stage) + COMP_Z * TJ;
    const double* __restrict__ sM = reinterpret_cast<const double*>(stage) + COMP_M * TJ;
    const double* __restrict__ sC = reinterpret_cast<const double*>(stage) + COMP_C * TJ;
    double cmax[NI];
#pragma unroll
    for (int q = 0; q < NI; ++q) cmax[q] = UC ? fmax(c_tile, S.c[q]) : 0.0;
#pragma unroll 1
    for (int j0 = jb; j0 < je; j0 += 2) {   // [jb, je): the whole tile, or one part of it (sub-tile work units)
        const double2 x2 = *reinterpret_cast<const double2*>(sX + j0);
        const double2 y2 = *reinterpret_cast<const double2*>(sY + j0);
        const double2 z2 = *reinterpret_cast<const double2*>(sZ + j0);
        double2 m2 = make_double2(1.0, 1.0);
        if (!UM) m2 = *reinterpret_cast<const double2*>(sM + j0);
        double2 c2 = make_double2(0.0, 0.0);
        if (KID != NB_KERNEL_NEWTONIAN && !UC) c2 = *reinterpret_cast<const double2*>(sC + j0);
#pragma unroll
        for (int q = 0; q < NI; ++q) {
            bool v0 = true, v1 = true;
            if (MASK) {
                const long long ig = ig0 + static_cast<long long>(THREADS) * q;
                const long long jg = jg0 + j0;
                v0 = (jg < n_j) && (jg != ig);
                v1 = (jg + 1 < n_j) && (jg + 1 != ig);
            }
            pair_f64<KID, MODE, MASK, UM, UC>(x2.x, y2.x, z2.x, m2.x, c2.x, S.x[q], S.y[q], S.z[q], S.c[q], cmax[q],
                                              S.ax[q], S.ay[q], S.az[q], S.sp[q], eps2, v0);
            pair_f64<KID, MODE, MASK, UM, UC>(x2.y, y2.y, z2.y, m2.y, c2.y, S.x[q], S.y[q], S.z[q], S.c[q], cmax[q],
                                              S.ax[q], S.ay[q], S.az[q], S.sp[q], eps2, v1);
        }
    }
}

// ---------------------------------------------------------------------------
// Cubic spline, unmasked tiles, two-phase groups (same idea as tile_f32_spline): phase 1 computes the
// separations and r^2 of a group (2 j x NI i) and tracks the minimum r^2 through the HIGH 32-bit word of
// the doubles (monotonic for positive values; one integer min per interaction instead of an emulated
// double fmin); if that minimum cannot be below the largest applicable h^2 the group finishes as pure
// Newtonian, otherwise it is redone with the exact piecewise polynomial.
// ---------------------------------------------------------------------------
template <int MODE, bool UM>
__device__ __forceinline__ void spline_exact_f64(const double xj, const double yj, const double zj,
                                                 const double mj_in, const double cj, const double xi,
                                                 const double yi, const double zi, const double ci, double& ax,
                                                 double& ay, double& az, double& sp, const double eps2) {
    const double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const double d  = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
    const double h2 = fmax(cj, ci);
    const double mj = UM ? 1.0 : mj_in;
    double w, p;
    const double ri = rsqrt_f64(d);
    if (d < h2) {
        spline_inside<double>(d, ri, rsqrt_f64(h2), w, p);
    } else {
        p = ri;
        w = ri * (ri * ri);
    }
    if (MODE & NB_MODE_FORCE) {
        const double f = mj * w;
        ax = fma(f, dx, ax);
        ay = fma(f, dy, ay);
        az = fma(f, dz, az);
    }
    if (MODE & NB_MODE_POT) sp = fma(mj, p, sp);
}

template <int MODE, bool UM, int NI>
__device__ __forceinline__ void tile_f64_spline(const unsigned char* __restrict__ stage, IState64<NI>& S,
                                                const double eps2, const double c_max_tile, const int jb = 0,
                                                const int je = TJ) {
    const double* __restrict__ sX = reinterpret_cast<const double*>(stage) + COMP_X * TJ;
    const double* __restrict__ sY = reinterpret_cast<const double*>(stage) + COMP_Y * TJ;
    const double* __restrict__ sZ = reinterpret_cast<const double*>(stage) + COMP_Z * TJ;
    const double* __restrict__ sM = reinterpret_cast<const double*>(stage) + COMP_M * TJ;
    const double* __restrict__ sC = reinterpret_cast<const double*>(stage) + COMP_C * TJ;
    double h2max = c_max_tile;
#pragma unroll
    for (int q = 0; q < NI; ++q) h2max = fmax(h2max, S.c[q]);
    const int hi_lim = __double2hiint(h2max);   // d may be inside only if hiword(d) <= hiword(h2max)

#pragma unroll 1
    for (int j0 = jb; j0 < je; j0 += 2) {
        const double2 x2 = *reinterpret_cast<const double2*>(sX + j0);
        const double2 y2 = *reinterpret_cast<const double2*>(sY + j0);
        const double2 z2 = *reinterpret_cast<const double2*>(sZ + j0);
        double2 m2 = make_double2(1.0, 1.0);
        if (!UM) m2 = *reinterpret_cast<const double2*>(sM + j0);
        const double xs[2] = {x2.x, x2.y}, ys[2] = {y2.x, y2.y}, zs[2] = {z2.x, z2.y}, ms[2] = {m2.x, m2.y};
        double dx[2][NI], dy[2][NI], dz[2][NI], d[2][NI];
        int hi_min = 0x7fffffff;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int q = 0; q < NI; ++q) {
                dx[jj][q] = xs[jj] - S.x[q];
                dy[jj][q] = ys[jj] - S.y[q];
                dz[jj][q] = zs[jj] - S.z[q];
                d[jj][q]  = fma(dx[jj][q], dx[jj][q], fma(dy[jj][q], dy[jj][q], fma(dz[jj][q], dz[jj][q], eps2)));
                hi_min = min(hi_min, __double2hiint(d[jj][q]));
            }
        if (hi_min <= hi_lim) {   // rare: some pair may be inside its softening length -> exact path
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {   // re-read the j-particle from shared memory (keeps xs.. in registers)
                const int j = j0 + jj;
                const double xj = sX[j], yj = sY[j], zj = sZ[j], mj = sM[j], cj = sC[j];
#pragma unroll
                for (int q = 0; q < NI; ++q)
                    spline_exact_f64<MODE, UM>(xj, yj, zj, mj, cj, S.x[q], S.y[q], S.z[q], S.c[q], S.ax[q], S.ay[q],
                                               S.az[q], S.sp[q], eps2);
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int q = 0; q < NI; ++q) {
                    double f;
                    if (MODE == NB_MODE_FORCE) {   // only r^-3 is needed
                        const double w = rsqrt3_f64(d[jj][q]);
                        f = UM ? w : ms[jj] * w;
                    } else {
                        const double rinv = rsqrt_f64(d[jj][q]);
                        const double mr   = UM ? rinv : ms[jj] * rinv;
                        S.sp[q] += mr;
                        f = mr * (rinv * rinv);
                    }
                    if (MODE & NB_MODE_FORCE) {
                        S.ax[q] = fma(f, dx[jj][q], S.ax[q]);
                        S.ay[q] = fma(f, dy[jj][q], S.ay[q]);
                        S.az[q] = fma(f, dz[jj][q], S.az[q]);
                    }
                }
        }
    }
}

template <int KID, int MODE, int NI, int MINB, bool SUB = false>
__global__ void __launch_bounds__(THREADS, MINB) sweep_f64_kernel(const SweepArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int IB = THREADS * NI;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * TILE_BYTES_F64);
    double* run        = reinterpret_cast<double*>(smem + STAGES * TILE_BYTES_F64 + 64);   // [4][NI][THREADS]

    const int tid = threadIdx.x;
    const int cta = blockIdx.x, grid = gridDim.x;
    const long long u0 = unit_begin(a.total, cta, grid);
    const long long u1 = unit_begin(a.total, cta + 1, grid);
    const int n_units  = static_cast<int>(u1 - u0);
    const double eps2  = a.eps2;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int sl   = SUB ? a.split_log2 : 0;   // SUB: sub-tile work units (small N), a separate instantiation
    const int upb  = a.n_jt << sl;              // work units per i-block
    const int plen = TJ >> sl;                  // j-particles per unit
    auto issue = [&](int k) {
        const long long u = u0 + k;
        const int jl      = static_cast<int>(u % upb) >> sl;
        const int gt      = local_to_global_tile(a, jl);
        uint64_t* bar     = &full_bar[k % STAGES];
        mbar_expect_tx(bar, TILE_BYTES_F64);
        bulk_g2s(smem + (k % STAGES) * TILE_BYTES_F64, a.packed + static_cast<size_t>(gt) * TILE_BYTES_F64,
                 TILE_BYTES_F64, bar);
    };
    if (tid == 0) {
        for (int k = 0; k < STAGES - 1 && k < n_units; ++k) issue(k);
    }

    long long u = u0;
    int k = 0;
    while (u < u1) {
        const int ib  = static_cast<int>(u / upb);
        const int jl0 = static_cast<int>(u - static_cast<long long>(ib) * upb);
        const long long rem = u1 - u;
        const int jl1 = static_cast<int>(min(static_cast<long long>(upb), jl0 + rem));

        const long long ig_block = a.i_begin + static_cast<long long>(ib) * IB;
        const long long ig0      = ig_block + tid;
        IState64<NI> S;
#pragma unroll
        for (int q = 0; q < NI; ++q) {
            const long long ig = ig0 + static_cast<long long>(THREADS) * q;
            S.x[q] = S.y[q] = S.z[q] = 0.0;
            S.c[q] = 0.0;   // dummy i beyond the buffer: results are discarded
            if (ig < a.n_pad) {
                const long long t = ig / TJ;
                const int kk      = static_cast<int>(ig - t * TJ);
                const double* tb  = reinterpret_cast<const double*>(a.packed + static_cast<size_t>(t) * TILE_BYTES_F64);
                S.x[q] = tb[COMP_X * TJ + kk];
                S.y[q] = tb[COMP_Y * TJ + kk];
                S.z[q] = tb[COMP_Z * TJ + kk];
                S.c[q] = tb[COMP_C * TJ + kk];
            }
            S.ax[q] = S.ay[q] = S.az[q] = S.sp[q] = 0.0;
        }
#pragma unroll
        for (int e = 0; e < 4 * NI; ++e) run[e * THREADS + tid] = 0.0;

        for (int v = jl0; v < jl1; ++v, ++k) {
            const int jl = v >> sl;                       // tile of this unit
            const int jb = (v - (jl << sl)) * plen;       // and the part of it: j-particles [jb, jb + plen)
            const int je = jb + plen;
            __syncthreads();
            if (tid == 0 && k + STAGES - 1 < n_units) issue(k + STAGES - 1);
            mbar_wait(&full_bar[k % STAGES], (k / STAGES) & 1);

            const int gt        = local_to_global_tile(a, jl);
            const long long jg0 = static_cast<long long>(gt) * TJ;
            const bool masked   = (jg0 + TJ > a.n_j) || (jg0 < ig_block + IB && jg0 + TJ > ig_block);
            const unsigned char* stage = smem + (k % STAGES) * TILE_BYTES_F64;
            const double2 meta = *reinterpret_cast<const double2*>(stage + 5 * TJ * 8);   // {c_uni, m_uni} or NaN
            const bool uc = (meta.x == meta.x), um = (meta.y == meta.y);
            double m_fold = 1.0;
            // constant bounds for whole tiles, a second inlined copy for sub-tile units (see nb_sweep_f32.cuh)
#define NB_TILE_DISPATCH_F64(JB, JE)                                                                                  \
    if (KID == NB_KERNEL_SPLINE && !masked && uc) {                                                                   \
        /* two-phase spline needs an upper bound of the tile's h^2: available when the tile is uniform */             \
        if (um) {                                                                                                     \
            m_fold = meta.y;                                                                                          \
            tile_f64_spline<MODE, true, NI>(stage, S, eps2, meta.x, JB, JE);                                          \
        } else {                                                                                                      \
            tile_f64_spline<MODE, false, NI>(stage, S, eps2, meta.x, JB, JE);                                         \
        }                                                                                                             \
    } else if (masked) {                                                                                              \
        tile_f64<KID, MODE, true, false, false, NI>(stage, S, eps2, 0.0, jg0, ig0, a.n_j, JB, JE);                    \
    } else if (um && uc) {                                                                                            \
        m_fold = meta.y;                                                                                              \
        tile_f64<KID, MODE, false, true, true, NI>(stage, S, eps2, meta.x, jg0, ig0, a.n_j, JB, JE);                  \
    } else if (um) {                                                                                                  \
        m_fold = meta.y;                                                                                              \
        tile_f64<KID, MODE, false, true, false, NI>(stage, S, eps2, 0.0, jg0, ig0, a.n_j, JB, JE);                    \
    } else {                                                                                                          \
        tile_f64<KID, MODE, false, false, false, NI>(stage, S, eps2, 0.0, jg0, ig0, a.n_j, JB, JE);                   \
    }
            if constexpr (!SUB) {
                NB_TILE_DISPATCH_F64(0, TJ)
            } else {
                NB_TILE_DISPATCH_F64(jb, je)
            }
#undef NB_TILE_DISPATCH_F64
            // fold the tile's partial sums (x tile mass) into the running sums in shared memory
#pragma unroll
            for (int q = 0; q < NI; ++q) {
                if (MODE & NB_MODE_FORCE) {
                    run[(0 * NI + q) * THREADS + tid] = fma(S.ax[q], m_fold, run[(0 * NI + q) * THREADS + tid]);
                    run[(1 * NI + q) * THREADS + tid] = fma(S.ay[q], m_fold, run[(1 * NI + q) * THREADS + tid]);
                    run[(2 * NI + q) * THREADS + tid] = fma(S.az[q], m_fold, run[(2 * NI + q) * THREADS + tid]);
                }
                if (MODE & NB_MODE_POT)
                    run[(3 * NI + q) * THREADS + tid] = fma(S.sp[q], m_fold, run[(3 * NI + q) * THREADS + tid]);
                S.ax[q] = S.ay[q] = S.az[q] = S.sp[q] = 0.0;
            }
        }

        const bool is_full = (jl0 == 0 && jl1 == upb);
        const size_t base  = is_full ? static_cast<size_t>(ib) * IB
                                     : (static_cast<size_t>(cta) * 2 + (u == u0 ? 0 : 1)) * IB;
        double4* dst = reinterpret_cast<double4*>(is_full ? a.full : a.part) + base;
#pragma unroll
        for (int q = 0; q < NI; ++q)
            dst[tid + THREADS * q] = make_double4(run[(0 * NI + q) * THREADS + tid], run[(1 * NI + q) * THREADS + tid],
                                                  run[(2 * NI + q) * THREADS + tid], run[(3 * NI + q) * THREADS + tid]);
        u += jl1 - jl0;
    }
}

}  // namespace nb
