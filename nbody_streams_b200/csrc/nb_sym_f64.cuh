// nb_sym_f64.cuh — pair-symmetric FP64 sweep (NB_PREC_F64), the large-N path.
//
// Same job as nb_sweep_f64.cuh (replaces nbody_forces_kernel<double> / nbody_potential_kernel<double>,
// reference cuda_kernels.py:755-864, :1299-1387) with every unordered pair evaluated once (nb_sym.cuh) and
// the same systolic structure as nb_sym_f32.cuh: NI = 4 i-particles per thread in registers, one j-particle
// per lane per step, travelling j-accumulators rotated with SHFL, per-warp j-slices in shared memory added
// to the double workspace after every tile.  Per pair: 3 DADD + 3 DFMA + rsqrt (MUFU.RSQ64H + 5) + 2 DMUL +
// 6 DFMA = 19 FP64 instructions for TWO interactions (the one-sided sweep: 16 per interaction).
// The i-side sums stay in double for the whole segment (registers per tile, shared-memory running sums
// across tiles because uniform tiles are accumulated unweighted and scaled by the tile mass per tile).
#pragma once
#include "nb_sweep_f64.cuh"
#include "nb_sym.cuh"

namespace nb {

#ifndef NB_SYM_NI_F64
#define NB_SYM_NI_F64 4
#endif
#ifndef NB_SYM_F64_DIRECT
#define NB_SYM_F64_DIRECT 1
#endif
constexpr int SYM_NI_F64     = NB_SYM_NI_F64;           // i-particles per thread
constexpr int SYM_IB_F64     = THREADS * SYM_NI_F64;    // 512 per CTA
constexpr int SYM_TPB_F64    = SYM_IB_F64 / TJ;         // 2 tiles per i-block
// j-side sums: DIRECT = every warp adds its 32 travelling accumulators to the workspace after each rotation round
// (RED.E.ADD.F64 from registers); otherwise per-warp slices in shared memory, reduced and added once per tile
constexpr bool SYM_F64_DIRECT = NB_SYM_F64_DIRECT != 0;
constexpr int SYM_STAGES_F64 = 2;                       // a tile takes ~40k cycles to sweep: one tile of prefetch is plenty
#ifndef NB_SYM_UNROLL_F64
#define NB_SYM_UNROLL_F64 2
#endif
constexpr int SYM_UNROLL_F64 = NB_SYM_UNROLL_F64;       // rotation steps per loop iteration
// force-only sweeps (18 FP64 instructions per pair, fewer live values) gain from 4 steps per iteration, the sweeps with the
// potential do not (N = 262144, ms, 2 -> 4 steps: plummer force 44.55 -> 43.47, spline force 47.72 -> 47.39;
// plummer force+potential 53.07 -> 53.49, spline force+potential 54.37 -> 54.30; profiles/r02_ubench_f64.md)
#ifndef NB_SYM_UNROLL_F64_FORCE
#define NB_SYM_UNROLL_F64_FORCE 4
#endif
constexpr int SYM_UNROLL_F64_FORCE = NB_SYM_UNROLL_F64_FORCE;

__device__ __forceinline__ double shfl_d(const double v, const int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}

struct JAcc64 {
    double x, y, z, p;
};

// mass-less kernel factors of one pair from d = r^2 + softening term: w (1/r^3-like), p (1/r-like)
template <int KID, int MODE>
__device__ __forceinline__ void sym_factors_f64(const double d, const double h2, const double eps2, double& w,
                                                double& p) {
    constexpr bool kDehnen = (KID == NB_KERNEL_DEHNEN_K1 || KID == NB_KERNEL_DEHNEN_K2);
    if (!kDehnen && MODE == NB_MODE_FORCE) {   // only r^-3 is needed
        w = rsqrt3_f64(d);
        p = 0.0;
        return;
    }
    const double rinv  = rsqrt_f64(d);
    const double rinv2 = rinv * rinv;
    p = rinv;
    w = 0.0;
    if (kDehnen) {
        const double t = (h2 - eps2) * rinv2;   // c = h^2 + eps2 seeds the r^2 chain; the terms use bare h^2
        if (MODE & NB_MODE_POT)
            p = rinv * ((KID == NB_KERNEL_DEHNEN_K1) ? fma(0.5, t, 1.0) : fma(t, fma(0.375, t, 0.5), 1.0));
        if (MODE & NB_MODE_FORCE)
            w = (rinv * rinv2) * ((KID == NB_KERNEL_DEHNEN_K1) ? fma(1.5, t, 1.0) : fma(t, fma(3.75, t, 1.5), 1.0));
    } else if (MODE & NB_MODE_FORCE) {
        w = rinv * rinv2;
    }
}

// acc += f * d for both sides of a pair.  A DFMA that reads three different registers takes 3 issue cycles, 2 when
// operands come from the reuse cache (nb_ubench 17-20); ptxas pairs the two accumulates that share f and d whatever
// the source order is (an asm block with the six in the ideal order produced the same SASS).
__device__ __forceinline__ void sym_acc6_f64(const double fi, const double fj, const double dx, const double dy,
                                             const double dz, double& ax, double& ay, double& az, JAcc64& J) {
    ax  = fma(dx, fi, ax);
    J.x = fma(dx, fj, J.x);
    ay  = fma(dy, fi, ay);
    J.y = fma(dy, fj, J.y);
    az  = fma(dz, fi, az);
    J.z = fma(dz, fj, J.z);
}

template <int MODE, bool U>
__device__ __forceinline__ void sym_apply_f64(const double w, const double p, const double dx, const double dy,
                                              const double dz, const double mj, const double mi, double& ax,
                                              double& ay, double& az, double& sp, JAcc64& J) {
    if (U) {
        if (MODE & NB_MODE_POT) {
            sp += p;
            J.p += p;
        }
        if (MODE & NB_MODE_FORCE) sym_acc6_f64(w, w, dx, dy, dz, ax, ay, az, J);
    } else {
        if (MODE & NB_MODE_POT) {
            sp  = fma(mj, p, sp);
            J.p = fma(mi, p, J.p);
        }
        if (MODE & NB_MODE_FORCE) sym_acc6_f64(mj * w, mi * w, dx, dy, dz, ax, ay, az, J);
    }
}

template <int NI>
struct SymCold64 {
    IState64<NI> S;
    double m[NI];
    JAcc64 J;
};

// one symmetric j-tile; jw = this warp's slice [4][TJ] of the j-side buffer
//   SUB: the unit is a part of the tile, steps [t0, t1) (whole rotation rounds); !SUB: the whole tile
template <int KID, int MODE, bool U, int NI, bool SUB>
__device__ __forceinline__ void sym_tile_f64(const unsigned char* __restrict__ stage, IState64<NI>& S,
                                             const double (&mi)[NI], const double cmax_i, const double c_pair,
                                             const double eps2, double* __restrict__ jw, const int lane,
                                             const int t0_in, const int t1_in, double* __restrict__ wsj = nullptr,
                                             const double scale = 1.0) {
    const int r0 = SUB ? (t0_in >> 5) : 0, r1 = SUB ? (t1_in >> 5) : TJ / 32;   // rotation rounds of this unit
    const double* __restrict__ sX = reinterpret_cast<const double*>(stage) + COMP_X * TJ;
    const double* __restrict__ sY = reinterpret_cast<const double*>(stage) + COMP_Y * TJ;
    const double* __restrict__ sZ = reinterpret_cast<const double*>(stage) + COMP_Z * TJ;
    const double* __restrict__ sM = reinterpret_cast<const double*>(stage) + COMP_M * TJ;
    const double* __restrict__ sC = reinterpret_cast<const double*>(stage) + COMP_C * TJ;
    constexpr bool kUsesC = (KID == NB_KERNEL_PLUMMER || KID == NB_KERNEL_DEHNEN_K1 || KID == NB_KERNEL_DEHNEN_K2);
    // (2 steps also for the spline: +3 % force+potential, +7 % potential over 1)
    constexpr int kUnroll = (MODE == NB_MODE_FORCE) ? SYM_UNROLL_F64_FORCE : SYM_UNROLL_F64;
    const int src = (lane + 1) & 31;

#pragma unroll 1
    for (int rho = r0; rho < r1; ++rho) {
        JAcc64 J;
        J.x = J.y = J.z = J.p = 0.0;
        const int jb = 32 * rho;
#pragma unroll kUnroll
        for (int s = 0; s < 32; ++s) {
            const int jj    = jb + ((lane + s) & 31);
            const double xj = sX[jj], yj = sY[jj], zj = sZ[jj];
            double mj = 1.0, cj = c_pair;
            if (!U) {
                mj = sM[jj];
                cj = sC[jj];
            }
            if (KID != NB_KERNEL_SPLINE) {
#pragma unroll
                for (int q = 0; q < NI; ++q) {
                    const double dx = xj - S.x[q], dy = yj - S.y[q], dz = zj - S.z[q];
                    double h2, c;
                    if (KID == NB_KERNEL_NEWTONIAN) {
                        h2 = c = eps2;
                    } else {
                        h2 = U ? c_pair : fmax(cj, S.c[q]);
                        c  = kUsesC ? h2 : eps2;
                    }
                    const double d = fma(dx, dx, fma(dy, dy, fma(dz, dz, c)));
                    double w, p;
                    sym_factors_f64<KID, MODE>(d, h2, eps2, w, p);
                    sym_apply_f64<MODE, U>(w, p, dx, dy, dz, mj, mi[q], S.ax[q], S.ay[q], S.az[q], S.sp[q], J);
                }
            } else {
                // two-phase: r^2 of the step first, minimum tracked through the high words (monotonic for
                // positive doubles); Newtonian finish unless some pair may lie inside its softening length
                double dx[NI], dy[NI], dz[NI], d[NI];
                int hi_min = 0x7fffffff;
#pragma unroll
                for (int q = 0; q < NI; ++q) {
                    dx[q]  = xj - S.x[q];
                    dy[q]  = yj - S.y[q];
                    dz[q]  = zj - S.z[q];
                    d[q]   = fma(dx[q], dx[q], fma(dy[q], dy[q], fma(dz[q], dz[q], eps2)));
                    hi_min = min(hi_min, __double2hiint(d[q]));
                }
                const double h2lim = U ? c_pair : fmax(cj, cmax_i);
                if (hi_min <= __double2hiint(h2lim)) {
                    // exact step, in place: every pair of the step with the reference's piecewise polynomial
                    // (cuda_kernels.py:687-752, :1228-1297 restated in spline_inside)
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        const double h2 = U ? c_pair : fmax(cj, S.c[q]);
                        const double ri = rsqrt_f64(d[q]);
                        double w, p;
                        if (d[q] < h2) {
                            spline_inside<double>(d[q], ri, rsqrt_f64(h2), w, p);
                        } else {
                            p = ri;
                            w = ri * (ri * ri);
                        }
                        sym_apply_f64<MODE, U>(w, p, dx[q], dy[q], dz[q], mj, mi[q], S.ax[q], S.ay[q], S.az[q], S.sp[q], J);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        double w, p;
                        sym_factors_f64<NB_KERNEL_NEWTONIAN, MODE>(d[q], 0.0, 0.0, w, p);
                        sym_apply_f64<MODE, U>(w, p, dx[q], dy[q], dz[q], mj, mi[q], S.ax[q], S.ay[q], S.az[q], S.sp[q], J);
                    }
                }
            }
            // systolic rotation: the accumulators follow their j-particle to the lane that meets it next
            if (MODE & NB_MODE_FORCE) {
                J.x = shfl_d(J.x, src);
                J.y = shfl_d(J.y, src);
                J.z = shfl_d(J.z, src);
            }
            if (MODE & NB_MODE_POT) J.p = shfl_d(J.p, src);
        }
        if (SYM_F64_DIRECT) {
            double* dst = wsj + (jb + lane) * 4;   // a_j += -m_i K dx ; s_phi_j += m_i / r
            if (MODE & NB_MODE_FORCE) {
                atomicAdd(dst + 0, -J.x * scale);
                atomicAdd(dst + 1, -J.y * scale);
                atomicAdd(dst + 2, -J.z * scale);
            }
            if (MODE & NB_MODE_POT) atomicAdd(dst + 3, J.p * scale);
        } else {
            if (MODE & NB_MODE_FORCE) {
                jw[0 * TJ + jb + lane] = J.x;
                jw[1 * TJ + jb + lane] = J.y;
                jw[2 * TJ + jb + lane] = J.z;
            }
            if (MODE & NB_MODE_POT) jw[3 * TJ + jb + lane] = J.p;
        }
    }
}

template <int KID, int MODE, int NI, bool SUB>
__device__ __noinline__ void sym_tile_f64_general(const unsigned char* stage, SymCold64<NI>* T, const double cmax_i,
                                                  const double eps2, double* jw, const int lane, const int t0,
                                                  const int t1, double* wsj) {
    IState64<NI> S = T->S;   // work on registers, not on the caller's local-memory copy
    double m[NI];
#pragma unroll
    for (int q = 0; q < NI; ++q) m[q] = T->m[q];
    sym_tile_f64<KID, MODE, false, NI, SUB>(stage, S, m, cmax_i, 0.0, eps2, jw, lane, t0, t1, wsj, 1.0);
    T->S = S;
}

// diagonal tile (overlaps the CTA's own i-block): one-sided masked path of nb_sweep_f64.cuh on a state copy
template <int KID, int MODE, int NI>
__device__ __noinline__ void sym_tile_f64_diag(const unsigned char* stage, IState64<NI>* T, const double eps2,
                                               const long long jg0, const long long ig0, const long long n_j,
                                               const int t0, const int t1) {
    tile_f64<KID, MODE, true, false, false, NI>(stage, *T, eps2, 0.0, jg0, ig0, n_j, t0, t1);
}

template <int NI>
struct SymF64Smem {
    static constexpr int kStageBytes = SYM_STAGES_F64 * TILE_BYTES_F64;
    static constexpr int kBarOff     = (kStageBytes + 63) & ~63;
    static constexpr int kRunOff     = kBarOff + 128;   // mbarriers + the TMA producer's unit cursor
    static constexpr int kRunBytes   = 4 * NI * THREADS * 8;              // double running sums [4][NI][THREADS]
    static constexpr int kJOff       = kRunOff + kRunBytes;
    static constexpr int kJBytes     = SYM_F64_DIRECT ? 0 : (THREADS / 32) * 4 * TJ * 8;   // j-side slices [warps][4][TJ]
    static constexpr int kTotal      = kJOff + kJBytes;
};

template <int KID, int MODE, int NI, int MINB, bool SUB = false>
__global__ void __launch_bounds__(THREADS, MINB) sym_f64_kernel(const SymArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    using L = SymF64Smem<NI>;
    constexpr int NSTG = SYM_STAGES_F64;
    constexpr int IB   = THREADS * NI;
    constexpr int TPB  = IB / TJ;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    double* run        = reinterpret_cast<double*>(smem + L::kRunOff);   // [4][NI][THREADS]
    double* jbuf       = reinterpret_cast<double*>(smem + L::kJOff);     // [warps][4][TJ]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, grid = gridDim.x;
    const SymRuns R(a, cta, grid);            // this CTA's units: round-robin chunks + a stream-K tail share
    const int n_units = static_cast<int>(R.n_units());
    if (n_units <= 0) return;
    const double eps2     = a.eps2;
    const long long i_end = static_cast<long long>(a.it0 + a.nT) * TJ;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTG; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // unit cursor of the consumer (all threads, registers) and of the TMA producer (thread 0 only, kept in
    // shared memory so that it costs no registers in the hot loop); the producer runs NSTG-1 units ahead
    SymCursor<TPB> cons;
    cons.open(a, R, 0);
    SymCursor<TPB>* prod = reinterpret_cast<SymCursor<TPB>*>(smem + L::kBarOff + 64);
    int issued = 0;
    auto issue_next = [&]() {
        prod->settle(a);
        bool dg;
        const int gt  = sym_tile<TPB>(a, prod->ib, prod->jl >> (SUB ? a.split_log2 : 0), dg);
        uint64_t* bar = &full_bar[issued % NSTG];
        mbar_expect_tx(bar, TILE_BYTES_F64);
        bulk_g2s(smem + (issued % NSTG) * TILE_BYTES_F64, a.packed + static_cast<size_t>(gt) * TILE_BYTES_F64, TILE_BYTES_F64, bar);
        ++issued;
        prod->advance(a, R, 1);
    };
    if (tid == 0) {
        prod->open(a, R, 0);
        for (int k = 0; k < NSTG - 1 && k < n_units; ++k) issue_next();
    }

    bool pend = false;
    int pend_gt = 0, pend_t0 = 0, pend_t1 = TJ;
    double pend_scale = 1.0;
    auto flush_j = [&]() {
#pragma unroll
        for (int jj = (SUB ? pend_t0 : 0) + tid; jj < (SUB ? pend_t1 : TJ); jj += THREADS) {
            double* dst = a.ws + (static_cast<size_t>(pend_gt) * TJ + jj) * 4;
#pragma unroll
            for (int comp = 0; comp < 4; ++comp) {
                if (comp < 3 ? !(MODE & NB_MODE_FORCE) : !(MODE & NB_MODE_POT)) continue;
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < THREADS / 32; ++w) v += jbuf[(w * 4 + comp) * TJ + jj];
                atomicAdd(dst + comp, (comp < 3 ? -v : v) * pend_scale);   // a_j += -m_i K dx ; s_phi_j += m_i / r
            }
        }
    };

    int k = 0;
    while (k < n_units) {
        cons.settle(a);
        const int ib  = cons.ib;
        const int jl0 = cons.jl;
        const int seg = static_cast<int>(min(static_cast<long long>(cons.cnt - cons.jl), cons.u_end - cons.u));

        const int tile0          = a.it0 + TPB * ib;
        const long long ig_block = static_cast<long long>(tile0) * TJ;
        const long long ig0      = ig_block + tid;
        IState64<NI> S;
        double mi[NI];
        double cmax_i = 0.0;
#pragma unroll
        for (int q = 0; q < NI; ++q) {
            const long long ig = ig0 + static_cast<long long>(THREADS) * q;
            S.x[q] = S.y[q] = S.z[q] = NB_PAD_POS_F64;   // dummy i beyond the i-range: far away, massless
            S.c[q] = 0.0;
            mi[q]  = 0.0;
            if (ig < i_end) {
                const long long t = ig / TJ;
                const int kk      = static_cast<int>(ig - t * TJ);
                const double* tb  = reinterpret_cast<const double*>(a.packed + static_cast<size_t>(t) * TILE_BYTES_F64);
                S.x[q] = tb[COMP_X * TJ + kk];
                S.y[q] = tb[COMP_Y * TJ + kk];
                S.z[q] = tb[COMP_Z * TJ + kk];
                mi[q]  = tb[COMP_M * TJ + kk];
                S.c[q] = tb[COMP_C * TJ + kk];
            }
            cmax_i = fmax(cmax_i, S.c[q]);
            S.ax[q] = S.ay[q] = S.az[q] = S.sp[q] = 0.0;
        }
        bool blk_uni = (tile0 + TPB <= a.it0 + a.nT);
        double cI = 0.0, mI = 0.0;
        if (blk_uni) {
#pragma unroll
            for (int t = 0; t < TPB; ++t) {
                const double2 meta = *reinterpret_cast<const double2*>(
                    a.packed + static_cast<size_t>(tile0 + t) * TILE_BYTES_F64 + 5 * TJ * 8);
                if (t == 0) {
                    cI = meta.x;
                    mI = meta.y;
                }
                blk_uni = blk_uni && (meta.x == cI) && (meta.y == mI);   // NaN (non-uniform tile) fails both
            }
        }
#pragma unroll
        for (int e = 0; e < 4 * NI; ++e) run[e * THREADS + tid] = 0.0;

        for (int t = 0; t < seg; ++t, ++k) {
            const int jl = jl0 + t;
            __syncthreads();   // everyone is done with tile k-1: its stage may be refilled, its j-slices are complete
            if (tid == 0 && issued < n_units) issue_next();
            if (!SYM_F64_DIRECT && pend) {
                flush_j();
                __syncthreads();
                pend = false;
            }
            mbar_wait(&full_bar[k % NSTG], (k / NSTG) & 1);

            bool diag;
            const int sl = SUB ? a.split_log2 : 0;
            const int gt = sym_tile<TPB>(a, ib, jl >> sl, diag);
            const int t0 = (jl & ((1 << sl) - 1)) * (TJ >> sl), t1 = t0 + (TJ >> sl);   // this unit's steps of the tile
            const unsigned char* stage = smem + (k % NSTG) * TILE_BYTES_F64;
            const double2 meta = *reinterpret_cast<const double2*>(stage + 5 * TJ * 8);   // {c_uni, m_uni} or NaN
            double m_fold = 1.0;
            if (diag) {
                IState64<NI> T = S;
                sym_tile_f64_diag<KID, MODE, NI>(stage, &T, eps2, static_cast<long long>(gt) * TJ, ig0, a.n_j,
                                                 SUB ? t0 : 0, SUB ? t1 : TJ);
                S = T;
            } else {
                double* jw  = jbuf + warp * 4 * TJ;
                double* wsj = a.ws + static_cast<size_t>(gt) * TJ * 4;
                if (blk_uni && meta.x == meta.x && meta.y == meta.y) {
                    m_fold     = meta.y;
                    pend_scale = mI;
                    sym_tile_f64<KID, MODE, true, NI, SUB>(stage, S, mi, cmax_i, fmax(cI, meta.x), eps2, jw, lane, t0, t1,
                                                           wsj, mI);
                } else {
                    pend_scale = 1.0;
                    SymCold64<NI> T;
                    T.S = S;
#pragma unroll
                    for (int q = 0; q < NI; ++q) T.m[q] = mi[q];
                    sym_tile_f64_general<KID, MODE, NI, SUB>(stage, &T, cmax_i, eps2, jw, lane, t0, t1, wsj);
                    S = T.S;
                }
                pend    = !SYM_F64_DIRECT;
                pend_gt = gt;
                pend_t0 = t0;
                pend_t1 = t1;
            }
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
        // segment done: add this block's i-side sums to the workspace
#pragma unroll
        for (int q = 0; q < NI; ++q) {
            const long long ig = ig0 + static_cast<long long>(THREADS) * q;
            if (ig < i_end) {
                double* dst = a.ws + static_cast<size_t>(ig) * 4;
#pragma unroll
                for (int comp = 0; comp < 4; ++comp) {
                    if (comp < 3 ? !(MODE & NB_MODE_FORCE) : !(MODE & NB_MODE_POT)) continue;
                    atomicAdd(dst + comp, run[(comp * NI + q) * THREADS + tid]);
                }
            }
        }
        cons.advance(a, R, seg);
    }
    __syncthreads();
    if (pend) flush_j();
}

}  // namespace nb
