// nb_sym_dispatch.cuh — host-side launch / occupancy plumbing of the pair-symmetric sweeps
// (one translation unit per softening kernel id: nb_sym_k<K>.cu).
#pragma once
#include "nb_sym_f32.cuh"
#include "nb_sym_f64.cuh"

namespace nb {

enum { SYM_OP_LAUNCH = 0, SYM_OP_QUERY = 1 };

#ifndef NB_SYM_MINB_F32
#define NB_SYM_MINB_F32 4
#endif
constexpr int SYM_MINB_F32 = NB_SYM_MINB_F32;   // resident CTAs / SM the f32 kernel is compiled for (4: <= 128 registers)
#ifndef NB_SYM_MINB_F64
#define NB_SYM_MINB_F64 4
#endif
constexpr int SYM_MINB_F64 = NB_SYM_MINB_F64;

// cudaFuncSetAttribute is needed once per (kernel, device), not per launch
template <auto KERN>
inline int sym_run(int smem, int op, const SymArgs* a, int grid, int* out, cudaStream_t st) {
    static unsigned long long configured = 0ull;   // bit per device ordinal (< 64)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured & bit)) {
        e = cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured |= bit;
    }
    if (op == SYM_OP_QUERY) {
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, KERN, THREADS, smem);
        if (e != cudaSuccess) return (int)e;
        *out = per_sm;
        return 0;
    }
    KERN<<<grid, THREADS, smem, st>>>(*a);
    return (int)cudaGetLastError();
}

template <int KID, bool SUB>
inline int dispatch_sym_sub(int prec, int mode, int op, const SymArgs* a, int grid, int* out, cudaStream_t st) {
    if (prec == NB_PREC_F64) {
        constexpr int smem = SymF64Smem<SYM_NI_F64>::kTotal;
        switch (mode) {
            case NB_MODE_FORCE: return sym_run<sym_f64_kernel<KID, NB_MODE_FORCE, SYM_NI_F64, SYM_MINB_F64, SUB>>(smem, op, a, grid, out, st);
            case NB_MODE_POT: return sym_run<sym_f64_kernel<KID, NB_MODE_POT, SYM_NI_F64, SYM_MINB_F64, SUB>>(smem, op, a, grid, out, st);
            case NB_MODE_BOTH: return sym_run<sym_f64_kernel<KID, NB_MODE_BOTH, SYM_NI_F64, SYM_MINB_F64, SUB>>(smem, op, a, grid, out, st);
        }
    } else {
        constexpr int smem = SymF32Smem<SYM_NP>::kTotal;
        switch (mode) {
            case NB_MODE_FORCE: return sym_run<sym_f32_kernel<KID, NB_MODE_FORCE, SYM_NP, SYM_MINB_F32, SUB>>(smem, op, a, grid, out, st);
            case NB_MODE_POT: return sym_run<sym_f32_kernel<KID, NB_MODE_POT, SYM_NP, SYM_MINB_F32, SUB>>(smem, op, a, grid, out, st);
            case NB_MODE_BOTH: return sym_run<sym_f32_kernel<KID, NB_MODE_BOTH, SYM_NP, SYM_MINB_F32, SUB>>(smem, op, a, grid, out, st);
        }
    }
    return NB_E_INVALID_ARG;
}

// sub-tile work units (a->split_log2 > 0, small problems) are a separate instantiation; the occupancy query uses the
// whole-tile kernel (the grid only has to be a sensible size: there is no inter-CTA synchronisation)
template <int KID>
inline int dispatch_sym(int prec, int mode, int op, const SymArgs* a, int grid, int* out, cudaStream_t st) {
    if (op == SYM_OP_LAUNCH && a->split_log2 > 0) return dispatch_sym_sub<KID, true>(prec, mode, op, a, grid, out, st);
    return dispatch_sym_sub<KID, false>(prec, mode, op, a, grid, out, st);
}

}  // namespace nb

#define NB_SYM_DECLARE_KID(K) \
    int nb_dispatch_sym_k##K(int prec, int mode, int op, const nb::SymArgs* a, int grid, int* out, cudaStream_t st);
NB_SYM_DECLARE_KID(0)
NB_SYM_DECLARE_KID(1)
NB_SYM_DECLARE_KID(2)
NB_SYM_DECLARE_KID(3)
NB_SYM_DECLARE_KID(4)

#define NB_SYM_DEFINE_KID(K)                                                                                  \
    int nb_dispatch_sym_k##K(int prec, int mode, int op, const nb::SymArgs* a, int grid, int* out,           \
                             cudaStream_t st) {                                                               \
        return nb::dispatch_sym<K>(prec, mode, op, a, grid, out, st);                                         \
    }
