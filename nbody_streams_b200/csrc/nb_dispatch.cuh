// nb_dispatch.cuh — host-side launch / occupancy-query plumbing shared by the
// per-kernel-id translation units (one .cu per softening kernel so the template
// instantiations compile in parallel).
#pragma once
#include "nb_sweep_f32.cuh"
#include "nb_sweep_f64.cuh"

namespace nb {

enum { OP_LAUNCH = 0, OP_QUERY = 1 };

// tuning variants: (i-pairs per thread NP, min resident CTAs per SM)
struct VariantF32 { int np, minb, ju; bool dec; };
struct VariantF64 { int ni, minb; };
constexpr int NV_F32 = 4;
constexpr int NV_F64 = 3;
__host__ __device__ constexpr VariantF32 variant_f32(int v) {
    // 0: default (2 i-pairs/thread, 5 CTAs/SM, 4 j per iteration)   1: 4 CTAs/SM, 124 registers
    // 2: 1 i-pair/thread, 8 CTAs/SM (best for force-only and potential-only plummer)
    // 3: decoupled warps (no per-tile __syncthreads)
    // measured on B200 (profiles/r01_variant_sweep.log): within 5 % of each other; nb_capi.cu picks per (kernel, mode).
    // Also tried: 1 i-pair at 6 and 7 CTAs/SM (80 / 64 registers) — no better than 0 or 2 in any mode.
    return v == 1 ? VariantF32{2, 4, 4, false} : v == 2 ? VariantF32{1, 8, 4, false}
         : v == 3 ? VariantF32{2, 5, 4, true} : VariantF32{2, 5, 4, false};
}
__host__ __device__ constexpr VariantF64 variant_f64(int v) {
    return v == 1 ? VariantF64{4, 3} : v == 2 ? VariantF64{2, 5} : VariantF64{2, 4};
}

template <typename K>
inline int run_kernel(K kern, int smem, int op, const SweepArgs* a, int grid, int* out, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    if (op == OP_QUERY) {
        int nb_per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_per_sm, kern, THREADS, smem);
        if (e != cudaSuccess) return (int)e;
        *out = nb_per_sm;
        return 0;
    }
    kern<<<grid, THREADS, smem, st>>>(*a);
    e = cudaGetLastError();
    return (int)e;
}

template <int KID, int MODE, int ACC, int V>
inline int run_f32(int op, const SweepArgs* a, int grid, int* out, cudaStream_t st) {
    constexpr VariantF32 v = variant_f32(V);
    if (op == OP_LAUNCH && a->split_log2 > 0)      // sub-tile work units: separate instantiation (same blocking)
        return run_kernel(sweep_f32_kernel<KID, MODE, ACC, v.np, v.minb, v.ju, v.dec, true>,
                          SweepF32Smem<v.np, ACC>::kTotal, op, a, grid, out, st);
    return run_kernel(sweep_f32_kernel<KID, MODE, ACC, v.np, v.minb, v.ju, v.dec, false>, SweepF32Smem<v.np, ACC>::kTotal,
                      op, a, grid, out, st);
}
template <int KID, int MODE, int V>
inline int run_f64(int op, const SweepArgs* a, int grid, int* out, cudaStream_t st) {
    constexpr VariantF64 v = variant_f64(V);
    constexpr int smem = STAGES * TILE_BYTES_F64 + 64 + 4 * v.ni * THREADS * 8;
    if (op == OP_LAUNCH && a->split_log2 > 0)
        return run_kernel(sweep_f64_kernel<KID, MODE, v.ni, v.minb, true>, smem, op, a, grid, out, st);
    return run_kernel(sweep_f64_kernel<KID, MODE, v.ni, v.minb, false>, smem, op, a, grid, out, st);
}

template <int KID, int MODE, int ACC, bool ALLV>
inline int pick_variant_f32(int variant, int op, const SweepArgs* a, int grid, int* out, cudaStream_t st) {
    if constexpr (ALLV) {
        switch (variant) {
            case 1: return run_f32<KID, MODE, ACC, 1>(op, a, grid, out, st);
            case 2: return run_f32<KID, MODE, ACC, 2>(op, a, grid, out, st);
            case 3: return run_f32<KID, MODE, ACC, 3>(op, a, grid, out, st);
            default: break;
        }
    }
    return run_f32<KID, MODE, ACC, 0>(op, a, grid, out, st);
}

template <int KID, bool ALLV>
inline int dispatch_f32(int mode, int acc, int variant, int op, const SweepArgs* a, int grid, int* out,
                        cudaStream_t st) {
    if (acc) {
        switch (mode) {
            case NB_MODE_FORCE: return pick_variant_f32<KID, NB_MODE_FORCE, 1, ALLV>(variant, op, a, grid, out, st);
            case NB_MODE_POT: return pick_variant_f32<KID, NB_MODE_POT, 1, ALLV>(variant, op, a, grid, out, st);
            case NB_MODE_BOTH: return pick_variant_f32<KID, NB_MODE_BOTH, 1, ALLV>(variant, op, a, grid, out, st);
        }
    } else {
        switch (mode) {
            case NB_MODE_FORCE: return pick_variant_f32<KID, NB_MODE_FORCE, 0, ALLV>(variant, op, a, grid, out, st);
            case NB_MODE_POT: return pick_variant_f32<KID, NB_MODE_POT, 0, ALLV>(variant, op, a, grid, out, st);
            case NB_MODE_BOTH: return pick_variant_f32<KID, NB_MODE_BOTH, 0, ALLV>(variant, op, a, grid, out, st);
        }
    }
    return NB_E_INVALID_ARG;
}

template <int KID, int MODE, bool ALLV>
inline int pick_variant_f64(int variant, int op, const SweepArgs* a, int grid, int* out, cudaStream_t st) {
    if constexpr (ALLV) {
        switch (variant) {
            case 1: return run_f64<KID, MODE, 1>(op, a, grid, out, st);
            case 2: return run_f64<KID, MODE, 2>(op, a, grid, out, st);
            default: break;
        }
    }
    return run_f64<KID, MODE, 0>(op, a, grid, out, st);
}

template <int KID, bool ALLV>
inline int dispatch_f64(int mode, int variant, int op, const SweepArgs* a, int grid, int* out, cudaStream_t st) {
    switch (mode) {
        case NB_MODE_FORCE: return pick_variant_f64<KID, NB_MODE_FORCE, ALLV>(variant, op, a, grid, out, st);
        case NB_MODE_POT: return pick_variant_f64<KID, NB_MODE_POT, ALLV>(variant, op, a, grid, out, st);
        case NB_MODE_BOTH: return pick_variant_f64<KID, NB_MODE_BOTH, ALLV>(variant, op, a, grid, out, st);
    }
    return NB_E_INVALID_ARG;
}

}  // namespace nb

// one pair of entry points per softening kernel id, defined in nb_sweep_k<K>.cu
#define NB_DECLARE_KID(K)                                                                                     \
    int nb_dispatch_f32_k##K(int mode, int acc, int variant, int op, const nb::SweepArgs* a, int grid, int* out, \
                             cudaStream_t st);                                                                \
    int nb_dispatch_f64_k##K(int mode, int variant, int op, const nb::SweepArgs* a, int grid, int* out,       \
                             cudaStream_t st);
NB_DECLARE_KID(0)
NB_DECLARE_KID(1)
NB_DECLARE_KID(2)
NB_DECLARE_KID(3)
NB_DECLARE_KID(4)

#define NB_DEFINE_KID(K, ALLV)                                                                                \
    int nb_dispatch_f32_k##K(int mode, int acc, int variant, int op, const nb::SweepArgs* a, int grid, int* out, \
                             cudaStream_t st) {                                                               \
        return nb::dispatch_f32<K, ALLV>(mode, acc, variant, op, a, grid, out, st);                           \
    }                                                                                                         \
    int nb_dispatch_f64_k##K(int mode, int variant, int op, const nb::SweepArgs* a, int grid, int* out,       \
                             cudaStream_t st) {                                                               \
        return nb::dispatch_f64<K, ALLV>(mode, variant, op, a, grid, out, st);                                \
    }
