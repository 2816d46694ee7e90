// nb_common.cuh — shared constants, PTX helpers and the softened pair kernels.
//
// Hand-written for sm_100a (B200).  No tensor cores on purpose: the r^-3 pair
// sum is not a dense contraction; the bound is the FP32 / FP64 CUDA-core pipe.
//
// Reference maths restated here (never copied): the five softening kernels of
//   nbody_streams/cuda_kernels.py:22-87   (force, float)   / :687-752  (double)
//   nbody_streams/cuda_kernels.py:373-436 (potential)      / :1228-1297 (double)
// and the max-softening pair rule of cuda_kernels.py:150.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nbody_sm100.h"

namespace nb {

constexpr int TJ      = 256;   // particles per j-tile (one bulk async copy)
constexpr int THREADS = 128;   // threads per sweep CTA
constexpr int STAGES  = 3;     // shared-memory ring depth for j-tiles

constexpr int I_F32 = 8;       // i-particles per thread, f32 (4 packed f32x2 pairs)
constexpr int I_F64 = 4;       // i-particles per thread, f64
constexpr int IB_F32 = THREADS * I_F32;   // i-particles per work unit
constexpr int IB_F64 = THREADS * I_F64;

// One tile of the packed buffer is structure-of-arrays: [x TJ][y TJ][z TJ][m TJ][c TJ].
// SoA lets the inner loop fetch 4 (f32) / 2 (f64) j-particles per LDS.128 and lets the
// i-pairs be fetched as aligned 64-bit (x_i, x_i+1) register pairs for FADD2/FFMA2.
// Each tile ends with 16 bytes of metadata written by nb_pack: {c_uni, m_uni} (f32: 2 floats + pad,
// f64: 2 doubles) = the common softening term / mass of the tile's 256 particles, or NaN when they
// differ.  Tiles with uniform mass / softening take sweep paths without the per-interaction FMNMX
// and mass multiply.  The metadata rides along with the tile's bulk copy (and the all-gather).
constexpr int TILE_META_BYTES = 16;
constexpr int TILE_BYTES_F32 = TJ * 20 + TILE_META_BYTES;
constexpr int TILE_BYTES_F64 = TJ * 40 + TILE_META_BYTES;
constexpr int COMP_X = 0, COMP_Y = 1, COMP_Z = 2, COMP_M = 3, COMP_C = 4;

// Padding records (index >= n) and dummy i-particles: zero mass, zero softening and a position far away
// from everything, so that an unmasked pair with a real particle evaluates to finite values times m = 0
// (r^2 ~ 3e36 / 3e300 stays inside the exponent range; a pad sitting ON a real particle would give inf * 0).
constexpr float NB_PAD_POS_F32  = 1.0e18f;
constexpr double NB_PAD_POS_F64 = 1.0e150;

// ---------------------------------------------------------------------------
// PTX: mbarrier + 1-D bulk async copy (TMA engine; SASS UBLKCP / SYNCS)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// arrive and learn how many arrivals were still pending BEFORE this one (1 => we completed the phase)
__device__ __forceinline__ uint32_t mbar_arrive_pending(uint64_t* bar) {
    uint64_t state;
    uint32_t cnt;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(cnt) : "l"(state));
    return cnt;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "NB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra NB_DONE;\n"
        "bra NB_WAIT;\n"
        "NB_DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU.RSQ
    return y;
}
// Full-precision double rsqrt: MUFU.RSQ64H seed y (~2^-20) + one third-order (Halley-type) correction
//   x^-1/2 = y (1 - e)^-1/2 = y (1 + e/2 + 3/8 e^2 + O(e^3)),  e = 1 - x y^2;  error ~ (5/16) e^3 < 2^-60.
// Written so that no DFMA reads three different registers (measured with nb_ubench 17-20: such a DFMA takes
// 3.05 issue cycles per warp instead of 2.05).
__device__ __forceinline__ double rsqrt_f64(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double xy = x * y;
    const double e  = fma(-xy, y, 1.0);
    const double t  = fma(0.375, e, 0.5);
    const double s  = fma(e, t, 1.0);
    return y * s;
}
// x^-3/2 directly from the same seed (one instruction less than rsqrt_f64 followed by two multiplies):
//   x^-3/2 = y^3 (1 - e)^-3/2 = y^3 (1 + 3/2 e + 15/8 e^2 + O(e^3)),  error ~ (35/16) e^3 < 2^-57.
__device__ __forceinline__ double rsqrt3_f64(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double y2 = y * y;
    const double e  = fma(-x, y2, 1.0);
    const double y3 = y2 * y;
    const double t  = fma(1.875, e, 1.5);
    const double s  = fma(e, t, 1.0);
    return y3 * s;
}

// ---------------------------------------------------------------------------
// Packed FP32x2 (Blackwell FADD2 / FMUL2 / FFMA2).  A pair lives in ONE 64-bit
// register (type f2) so ptxas keeps it in an aligned register pair for its whole
// life; going through float2 + the __f*2_rn intrinsics makes ptxas rebuild
// loop-invariant pairs with MOVs on every use (measured in SASS).
// ---------------------------------------------------------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 splat(float v) { return pk(v, v); }   // becomes a .F32 broadcast operand
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float lo_of(f2 v) { float a, b; upk(v, a, b); return a; }
__device__ __forceinline__ float hi_of(f2 v) { float a, b; upk(v, a, b); return b; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---------------------------------------------------------------------------
// Cubic-spline (Monaghan 1992) pieces, evaluated only for pairs with r < h.
// Restates cuda_kernels.py:61-87 (force) and :405-436 (potential); T = float or double.
// ---------------------------------------------------------------------------
// Division-free form for the sweeps' exact paths: given d = r^2 (> 0 or 0), rinv = 1/r and hinv = 1/h
// (both from the kernels' reciprocal square roots) it returns w (the 1/r^3-equivalent force factor) and p
// (the +1/r-equivalent, i.e. minus the potential factor).  The reference's polynomials in q = r/h; its
// -(1/15)/q^3 and (1/15)/q terms of the outer branch are written as (1/15) rinv^3 and (1/15) rinv.
// d == 0 (coincident particles with eps2 = 0) takes the q -> 0 limit like the reference's q < 1e-8 case.
template <typename T>
__device__ __forceinline__ void spline_inside(const T d, const T rinv, const T hinv, T& w, T& p) {
    const T r     = d > T(0) ? d * rinv : T(0);
    const T q     = r * hinv;
    const T q2    = q * q;
    const T h3inv = hinv * hinv * hinv;
    if (q <= T(0.5)) {
        w = h3inv * (T(10.666666666666666) + q2 * (T(32.0) * q - T(38.4)));
        p = (T(2.8) - q2 * (T(5.333333333333333) + (q2 * q2) * (T(6.4) * q - T(9.6)))) * hinv;
    } else {
        const T k = T(0.06666666666666667) * rinv;
        w = h3inv * (T(21.333333333333333) + q * (T(-48.0) + q * (T(38.4) - T(10.666666666666667) * q))) -
            k * (rinv * rinv);
        p = (T(3.2) - q2 * (T(10.666666666666666) + q * (T(-16.0) + q * (T(9.6) - T(2.1333333333333333) * q)))) * hinv - k;
    }
}

// integer helpers for the stream-K split: CTA c owns units [unit_begin(c), unit_begin(c+1))
__host__ __device__ __forceinline__ long long unit_begin(long long total, int c, int grid) {
    return (total * (long long)c) / (long long)grid;
}

}  // namespace nb
