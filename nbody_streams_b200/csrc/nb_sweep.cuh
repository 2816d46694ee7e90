// nb_sweep.cuh — argument block + stream-K bookkeeping shared by the f32 and f64
// sweep kernels and by the finish (fix-up) kernel.
#pragma once
#include "nb_common.cuh"

namespace nb {

// One sweep = i-particles [i_begin, i_begin+n_i) x j-tiles (range A then range B).
// Work units are (i-block, j-tile) pairs, linearised i-block-major; CTA c of a
// persistent grid owns the contiguous unit range [unit_begin(c), unit_begin(c+1)).
// A CTA therefore walks "segments" (i-block, [jl0, jl1)); a segment that covers
// the whole j-range is written to full[i], a partial one to part[cta][slot][i]
// (slot 0: the CTA's first segment, slot 1: its last).  nb_finish re-derives the
// same split and adds the pieces in j-ascending order, so results do not depend
// on scheduling.
struct SweepArgs {
    const unsigned char* packed;  // tile-major particle records
    long long n_j;                // real particles in packed (>= n_j is padding)
    long long n_pad;              // padded particle count (whole tiles)
    int jtA0, lenA, jtB0, lenB;   // j-tile ranges; local tile jl -> jl<lenA ? jtA0+jl : jtB0+jl-lenA
    int n_jt;                     // lenA + lenB
    int split_log2;               // a tile is swept in 2^split_log2 parts, one work unit each (small N: more units
                                  // than CTAs); units per i-block upb = n_jt << split_log2
    int n_ib;                     // number of i-blocks
    long long i_begin, n_i;       // global index of first i, count
    long long total;              // n_ib * upb work units
    double eps2;
    unsigned char* full;          // [n_ib * IB] partial elements
    unsigned char* part;          // [grid * 2 * IB] partial elements
};

// elements per i in the partial buffers: 4 accumulators (ax, ay, az, s_phi);
// float for NB_PREC_F32, double otherwise.
__host__ __device__ inline size_t partial_elem_bytes(int prec) { return prec == NB_PREC_F32 ? 16 : 32; }
__host__ __device__ inline int block_i(int prec) { return prec == NB_PREC_F64 ? IB_F64 : IB_F32; }

// Sub-tile work units for small problems: with fewer than two units per CTA the persistent grid is mostly idle
// (N = 10k in float32: 800 units for 740 CTAs, makespan 2 against a mean of 1.08), so tiles are cut into 2 or 4
// parts (N = 10k: 3200 quarter-tile units, makespan 5 against a mean of 4.3).  nb_sweep and nb_finish derive the same split from (n_ib, n_jt, grid).
__host__ __device__ inline int pick_split_log2(long long n_ib, long long n_jt, int grid) {
    const long long units = n_ib * n_jt;        // aim at >= 4 units per CTA: the tail is then <= 1/4 of the sweep
    if (units >= 4LL * grid) return 0;
    if (2 * units >= 4LL * grid) return 1;
    return 2;
}

__device__ __forceinline__ int local_to_global_tile(const SweepArgs& a, int jl) {
    return jl < a.lenA ? a.jtA0 + jl : a.jtB0 + (jl - a.lenA);
}

}  // namespace nb
