// nb_sym.cuh — argument block + work-unit enumeration of the PAIR-SYMMETRIC sweeps.
//
// The one-sided sweeps (nb_sweep_f32.cuh / nb_sweep_f64.cuh) evaluate every ordered pair (i, j): N^2 kernel
// evaluations, like the reference kernels (cuda_kernels.py:141-168).  The kernel factor of a pair,
// K(r_ij^2, max(h_i, h_j)), is symmetric in i and j, so the symmetric sweeps evaluate every UNORDERED pair
// once and apply it to both particles (Newton's third law): a_i += m_j K (r_j - r_i), a_j -= m_i K (r_j - r_i).
// Per pair that is 3 more FMAs but no second distance / rsqrt / kernel evaluation: 14 instead of 2 x 11
// packed FP32 instructions per two interactions.
//
// Work decomposition: i-blocks of TPB tiles (the particles a CTA keeps in registers) x j-tiles.  For an i-range
// of tiles [it0, it0+nT) a block ib sweeps
//     - (tri) the tiles of its own range from its first tile on: the first TPB of them overlap the block itself
//       ("diagonal": one-sided masked path, self pairs excluded), the rest are symmetric;
//     - the rectangle ranges A and B (other ranks' shards, symmetric);
//     - range C if ibC0 <= ib < ibC1 (the half-shard rule that splits the opposite shard of an even ring).
// Units are linearised block-major and cut into equal contiguous pieces per CTA (stream-K); partial sums of
// both sides are added atomically (double) into ws[slot][4], so pieces need no fix-up pass.
#pragma once
#include "nb_common.cuh"

namespace nb {

struct SymArgs {
    const unsigned char* packed;  // tile-major particle records
    long long n_j;                // real particles in `packed` (index >= n_j is padding)
    int it0, nT;                  // i-range: tiles [it0, it0 + nT)
    int n_ib;                     // i-blocks = ceil(nT / TPB)
    int tri;                      // sweep the triangle inside the i-range
    int rA0, lenA, rB0, lenB;     // rectangle j-tile ranges (every i-block)
    int rC0, lenC, ibC0, ibC1;    // extra range for i-blocks [ibC0, ibC1)
    int split_log2;               // a tile is swept in 2^split_log2 parts, one work unit each (small problems: more
                                  // units than CTAs); unit indices below are in these SUB-units
    long long total;              // work units (sub-units)
    double eps2;
    double* ws;                   // [slots][4] doubles: sum of ax, ay, az, s_phi (atomically accumulated)
};

// units of the i-blocks before ib
template <int TPB>
__host__ __device__ inline long long sym_prefix(const SymArgs& a, long long ib) {
    long long p = ib * static_cast<long long>(a.lenA + a.lenB);
    if (a.tri) p += ib * a.nT - static_cast<long long>(TPB) * (ib * (ib - 1) / 2);
    long long c = ib - a.ibC0;
    const long long cmax = a.ibC1 - a.ibC0;
    if (c > cmax) c = cmax;
    if (c > 0) p += c * a.lenC;
    return p;
}
template <int TPB>
__host__ __device__ inline int sym_count(const SymArgs& a, int ib) {
    int n = a.lenA + a.lenB;
    if (a.tri) n += a.nT - TPB * ib;
    if (ib >= a.ibC0 && ib < a.ibC1) n += a.lenC;
    return n;
}
// the i-block holding unit u: largest ib with prefix(ib) <= u
template <int TPB>
__host__ __device__ inline int sym_find_block(const SymArgs& a, long long u) {
    int lo = 0, hi = a.n_ib - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sym_prefix<TPB>(a, mid) <= u) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
// global tile of local unit jl of block ib; diag = the tile overlaps the block's own particles
template <int TPB>
__host__ __device__ inline int sym_tile(const SymArgs& a, int ib, int jl, bool& diag) {
    const int triA = a.tri ? a.nT - TPB * ib : 0;
    if (jl < triA) {
        diag = jl < TPB;
        return a.it0 + TPB * ib + jl;
    }
    diag = false;
    jl -= triA;
    if (jl < a.lenA) return a.rA0 + jl;
    jl -= a.lenA;
    if (jl < a.lenB) return a.rB0 + jl;
    return a.rC0 + (jl - a.lenB);
}

// ---------------------------------------------------------------------------
// Which units a CTA sweeps.  Units cost different amounts (uniform tiles take the fast path, tiles with
// per-particle masses / softening or padding the general one, diagonal tiles the masked one), and such tiles
// come in contiguous stretches of the unit list, so a purely contiguous split would leave whole stretches of
// slow units to single CTAs.  The list is therefore dealt out in CHUNKS of SYM_CHUNK consecutive units,
// round-robin over the CTAs (chunk c, c+grid, c+2 grid, ...), for as many complete rounds as there are; the
// remainder (< grid chunks) is cut into equal contiguous shares like stream-K, so every CTA still gets the
// same number of units +-1.  Static and arithmetic: no atomics, and the TMA producer can run ahead.
// A "run" is one chunk or the tail share; SymRuns enumerates the runs of one CTA.
// ---------------------------------------------------------------------------
constexpr int SYM_CHUNK = 8;

// Sub-tile work units for small problems (same idea as pick_split_log2 of the one-sided sweeps): with fewer than four
// units per CTA a tile is swept in 2 or 4 parts of 4 / 2 rotation rounds.  `units` = tile-units of the launch.
__host__ __device__ inline int sym_pick_split_log2(long long units, int grid) {
    if (units >= 4LL * grid) return 0;
    if (2 * units >= 4LL * grid) return 1;
    return 2;
}

struct SymRuns {
    long long rounds;     // complete round-robin rounds of chunks
    long long tail_b, tail_e;
    int cta, grid;
    __device__ SymRuns(const SymArgs& a, int cta_, int grid_) : cta(cta_), grid(grid_) {
        rounds = (a.total / SYM_CHUNK) / grid;
        const long long base = rounds * grid * SYM_CHUNK;
        const long long rem  = a.total - base;
        tail_b = base + unit_begin(rem, cta, grid);
        tail_e = base + unit_begin(rem, cta + 1, grid);
    }
    __device__ long long n_units() const { return rounds * SYM_CHUNK + (tail_e - tail_b); }
    __device__ long long n_runs() const { return rounds + (tail_e > tail_b ? 1 : 0); }
    __device__ void run(long long r, long long& b, long long& e) const {
        if (r < rounds) {
            b = (r * grid + cta) * SYM_CHUNK;
            e = b + SYM_CHUNK;
        } else {
            b = tail_b;
            e = tail_e;
        }
    }
};

// cursor over the units of one CTA in sweep order, decoded to (i-block, local unit)
template <int TPB>
struct SymCursor {
    long long r, u, u_end;   // current run, current unit, end of the run
    int ib, jl, cnt;
    // (jl and cnt count sub-units: unit jl of block ib is part jl & (S-1) of the block's tile-unit jl >> split_log2)
    __device__ void open(const SymArgs& a, const SymRuns& R, long long run) {
        r = run;
        R.run(run, u, u_end);
        ib  = sym_find_block<TPB>(a, u >> a.split_log2);
        jl  = static_cast<int>(u - (sym_prefix<TPB>(a, ib) << a.split_log2));
        cnt = sym_count<TPB>(a, ib) << a.split_log2;
    }
    // make (ib, jl) valid for the current unit (skips exhausted / empty blocks)
    __device__ void settle(const SymArgs& a) {
        while (jl >= cnt) {
            ++ib;
            jl  = 0;
            cnt = sym_count<TPB>(a, ib) << a.split_log2;
        }
    }
    // advance by n units inside the current run; opens the next run when this one is exhausted
    __device__ void advance(const SymArgs& a, const SymRuns& R, int n) {
        u += n;
        jl += n;
        if (u >= u_end && r + 1 < R.n_runs()) open(a, R, r + 1);
    }
};

}  // namespace nb
