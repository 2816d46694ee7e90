// nb_misc.cu — O(N) kernels around the sweep: pack, finish (stream-K fix-up),
// fused KDK kick/drift/pack, energy reductions.
//
// Reference behaviour replaced (all were separate CuPy array expressions, one
// launch each): fields.py:596-598,601-614 (column split, scalar broadcast),
// fields.py:832-836 (float4 pack), fields.py:855-858 ((N,3) assembly),
// run.py:188,195 (dtype casts), run.py:672,675,695 (kick / drift),
// run.py:619-623 (energies).
#include "nb_sweep.cuh"
#include "nb_sym.cuh"

namespace nb {

// ---------------------------------------------------------------------------
// pack
// ---------------------------------------------------------------------------
__device__ __forceinline__ double load_as_double(const void* p, int dt, long long idx) {
    return dt == NB_DT_F64 ? static_cast<const double*>(p)[idx] : static_cast<double>(static_cast<const float*>(p)[idx]);
}

// one CTA of TJ threads per tile: SoA records + the tile's uniformity metadata
template <typename T>
__global__ void __launch_bounds__(TJ) pack_kernel(unsigned char* __restrict__ tiles, const void* __restrict__ pos,
                                                   int pos_dt, const void* __restrict__ mass, int mass_dt,
                                                   double mass_scalar, const void* __restrict__ h, int h_dt,
                                                   double h_scalar, long long n, double eps2, int add_eps) {
    constexpr int TB = (sizeof(T) == 8) ? TILE_BYTES_F64 : TILE_BYTES_F32;
    const long long t = blockIdx.x;
    const int k       = threadIdx.x;
    const long long idx = t * TJ + k;
    T* tb = reinterpret_cast<T*>(tiles + static_cast<size_t>(t) * TB);
    // padding: zero mass, zero softening, parked far away (see NB_PAD_POS_*)
    const T pad = (sizeof(T) == 8) ? static_cast<T>(NB_PAD_POS_F64) : static_cast<T>(NB_PAD_POS_F32);
    T x = pad, y = pad, z = pad, m = T(0), c = T(0);
    if (idx < n) {
        // cast to the kernel dtype FIRST (reference: pos.astype(dtype), run.py:188), then square h
        x = static_cast<T>(load_as_double(pos, pos_dt, 3 * idx + 0));
        y = static_cast<T>(load_as_double(pos, pos_dt, 3 * idx + 1));
        z = static_cast<T>(load_as_double(pos, pos_dt, 3 * idx + 2));
        m = static_cast<T>(mass ? load_as_double(mass, mass_dt, idx) : mass_scalar);
        const T hh = static_cast<T>(h ? load_as_double(h, h_dt, idx) : h_scalar);
        c = hh * hh;
        if (add_eps) c += static_cast<T>(eps2);
    }
    tb[COMP_X * TJ + k] = x;
    tb[COMP_Y * TJ + k] = y;
    tb[COMP_Z * TJ + k] = z;
    tb[COMP_M * TJ + k] = m;
    tb[COMP_C * TJ + k] = c;
    __shared__ T first[2];
    if (k == 0) {
        first[0] = c;
        first[1] = m;
    }
    __syncthreads();
    const int c_same = __syncthreads_and(c == first[0]);
    const int m_same = __syncthreads_and(m == first[1]);
    // tile maximum of c (float tiles only: 16 B of metadata hold {c_uni, m_uni, c_max, 0})
    __shared__ float cmax_w[TJ / 32];
    float cm = static_cast<float>(c);
    if (sizeof(T) == 4) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, off));
        if ((k & 31) == 0) cmax_w[k >> 5] = cm;
        __syncthreads();
    }
    if (k == 0) {
        const T nan = static_cast<T>(__longlong_as_double(0x7ff8000000000000LL));
        T* meta = tb + 5 * TJ;
        meta[0] = c_same ? first[0] : nan;
        meta[1] = m_same ? first[1] : nan;
        if (sizeof(T) == 4) {
            float mx = cmax_w[0];
#pragma unroll
            for (int w = 1; w < TJ / 32; ++w) mx = fmaxf(mx, cmax_w[w]);
            meta[2] = static_cast<T>(mx);
            meta[3] = T(0);
        }
    }
}

// ---------------------------------------------------------------------------
// finish: add the stream-K pieces of every sweep slot in fixed order
// ---------------------------------------------------------------------------
struct FinishArgs {
    const unsigned char* ws;     // workspace base
    size_t slot_bytes;           // bytes per sweep slot
    size_t part_off;             // offset of part[] inside a slot
    int n_sweeps;
    int n_jt[4];
    int grid;
    int IB;
    int n_ib;
    long long n_i;
    double G;
};

// CTA that owns unit u of `total` units split over `grid` CTAs: the largest c with unit_begin(c) <= u.
// 32-bit arithmetic whenever the products fit (always in the small-N regime this kernel matters for).
__device__ __forceinline__ int cta_of_unit(long long u, long long total, int grid, bool small) {
    if (small) {
        const unsigned t = static_cast<unsigned>(total);
        return static_cast<int>((static_cast<unsigned>(u + 1) * static_cast<unsigned>(grid) + t - 1u) / t) - 1;
    }
    return static_cast<int>(((u + 1) * grid + total - 1) / total) - 1;
}
__device__ __forceinline__ long long unit_begin_fast(long long total, int c, int grid, bool small) {
    if (small) return static_cast<long long>((static_cast<unsigned>(total) * static_cast<unsigned>(c)) / static_cast<unsigned>(grid));
    return unit_begin(total, c, grid);
}

// TPI threads share one i-particle: thread r takes the pieces k = r, r + TPI, ... (ascending) and the TPI
// partial sums are combined in a fixed order, so the result does not depend on scheduling.  At small N an
// i-block is cut into dozens of pieces (one per CTA that touched it); walking them with one thread per i was
// 10-35 us of pure latency.
template <typename PT /*float4|double4*/, typename OT /*float|double*/, int TPI>
__global__ void finish_kernel(const FinishArgs a, OT* __restrict__ acc, OT* __restrict__ phi) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long i = t / TPI;
    const int r = static_cast<int>(t - i * TPI);
    const bool live = i < a.n_i;
    const long long ii = live ? i : a.n_i - 1;            // idle lanes shadow a valid i (shuffles need every lane)
    const int ib = static_cast<int>(ii / a.IB);
    const int li = static_cast<int>(ii - static_cast<long long>(ib) * a.IB);
    double sx = 0.0, sy = 0.0, sz = 0.0, sp = 0.0;
    for (int s = 0; s < a.n_sweeps; ++s) {
        const int n_jt = a.n_jt[s];                       // work units per i-block of this slot
        if (n_jt <= 0) continue;
        const unsigned char* slot = a.ws + static_cast<size_t>(s) * a.slot_bytes;
        const PT* full = reinterpret_cast<const PT*>(slot);
        const PT* part = reinterpret_cast<const PT*>(slot + a.part_off);
        const long long total = static_cast<long long>(a.n_ib) * n_jt;
        const bool small = (total + 1) * a.grid < (1LL << 31);
        const long long ub = static_cast<long long>(ib) * n_jt, ue = ub + n_jt;
        // pieces of [ub, ue) in ascending unit (= ascending CTA) order.  dense: every CTA owns at least one unit,
        // so piece k belongs to CTA c0 + k; otherwise every unit is its own piece (most CTAs are empty)
        const bool dense = total >= a.grid;
        const int c0 = cta_of_unit(ub, total, a.grid, small);
        const int n_pieces = dense ? cta_of_unit(ue - 1, total, a.grid, small) - c0 + 1 : n_jt;
        for (int k = r; k < n_pieces; k += TPI) {
            const int c = dense ? c0 + k : cta_of_unit(ub + k, total, a.grid, small);
            const long long cb = unit_begin_fast(total, c, a.grid, small);
            const long long ce = unit_begin_fast(total, c + 1, a.grid, small);
            const long long sb = cb > ub ? cb : ub, se = ce < ue ? ce : ue;
            PT v;
            if (sb == ub && se == ue) v = full[static_cast<size_t>(ib) * a.IB + li];
            else v = part[(static_cast<size_t>(c) * 2 + (sb == cb ? 0 : 1)) * a.IB + li];
            sx += static_cast<double>(v.x);
            sy += static_cast<double>(v.y);
            sz += static_cast<double>(v.z);
            sp += static_cast<double>(v.w);
        }
    }
    if (TPI > 1) {
#pragma unroll
        for (int off = 1; off < TPI; off <<= 1) {          // fixed combination order: (0+1) + (2+3)
            sx += __shfl_xor_sync(0xffffffffu, sx, off);
            sy += __shfl_xor_sync(0xffffffffu, sy, off);
            sz += __shfl_xor_sync(0xffffffffu, sz, off);
            sp += __shfl_xor_sync(0xffffffffu, sp, off);
        }
    }
    if (!live || r != 0) return;
    if (acc) {
        acc[3 * i + 0] = static_cast<OT>(a.G * sx);
        acc[3 * i + 1] = static_cast<OT>(a.G * sy);
        acc[3 * i + 2] = static_cast<OT>(a.G * sz);
    }
    if (phi) phi[i] = static_cast<OT>(-a.G * sp);
}

// finish of the pair-symmetric sweeps: the workspace already holds the complete double sums of every slot
template <typename OT>
__global__ void sym_finish_kernel(const double4* __restrict__ ws, long long first, long long n_i, double G,
                                  OT* __restrict__ acc, OT* __restrict__ phi) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_i) return;
    const double4 v = ws[first + i];
    if (acc) {
        acc[3 * i + 0] = static_cast<OT>(G * v.x);
        acc[3 * i + 1] = static_cast<OT>(G * v.y);
        acc[3 * i + 2] = static_cast<OT>(G * v.z);
    }
    if (phi) phi[i] = static_cast<OT>(-G * v.w);
}

// Multi-GPU finish fused with the reduction of the partial sums (exchange = 'p2p'): every rank's symmetric sweep
// leaves partial sums for ALL particle slots in its own workspace; the owner of a shard adds the rows of its shard
// from every rank's workspace — its own and, over NVLink peer memory, the others' — in fixed rank order, scales and
// writes acc / phi.  Replaces reduce-scatter + finish by ONE kernel; the caller brackets it with symmetric-memory
// barriers (all sweeps done before, all reads done before the workspaces are cleared again).
struct PeerWs {
    const double4* p[8];
    int count;
};

template <typename OT>
__global__ void sym_finish_peers_kernel(const PeerWs peers, long long first, long long n_i, double G,
                                        OT* __restrict__ acc, OT* __restrict__ phi) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_i) return;
    double4 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)       // all loads in flight before the first add (peer latency ~2 us)
        if (r < peers.count) v[r] = peers.p[r][first + i];
    double sx = 0.0, sy = 0.0, sz = 0.0, sp = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (r < peers.count) {
            sx += v[r].x;
            sy += v[r].y;
            sz += v[r].z;
            sp += v[r].w;
        }
    }
    if (acc) {
        acc[3 * i + 0] = static_cast<OT>(G * sx);
        acc[3 * i + 1] = static_cast<OT>(G * sy);
        acc[3 * i + 2] = static_cast<OT>(G * sz);
    }
    if (phi) phi[i] = static_cast<OT>(-G * sp);
}

// ---------------------------------------------------------------------------
// KDK pieces
// ---------------------------------------------------------------------------
template <typename AT /*acc dtype*/, typename T /*packed dtype*/>
__global__ void kick_drift_pack_kernel(double* __restrict__ pos, double* __restrict__ vel,
                                       const AT* __restrict__ acc, const double* __restrict__ acc_ext,
                                       T* __restrict__ tiles, long long first, long long n, double dt_half,
                                       double dt) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double a = static_cast<double>(acc[3 * i + d]);
        if (acc_ext) a += acc_ext[3 * i + d];
        // same operation order as the reference: vel += acc*(dt/2); pos += vel*dt  (run.py:672,675)
        const double v = vel[3 * i + d] + a * dt_half;
        vel[3 * i + d] = v;
        p[d] = pos[3 * i + d] + v * dt;
        pos[3 * i + d] = p[d];
    }
    if (tiles) {
        constexpr int TB = (sizeof(T) == 8) ? TILE_BYTES_F64 : TILE_BYTES_F32;
        const long long g = first + i;
        const long long t = g / TJ;
        const int k       = static_cast<int>(g - t * TJ);
        T* tb = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(tiles) + static_cast<size_t>(t) * TB);
        tb[COMP_X * TJ + k] = static_cast<T>(p[0]);
        tb[COMP_Y * TJ + k] = static_cast<T>(p[1]);
        tb[COMP_Z * TJ + k] = static_cast<T>(p[2]);
    }
}

// ---------------------------------------------------------------------------
// Fused kick + drift + repack + EXCHANGE over NVLink peer memory (multi-GPU, exchange='p2p'):
// the new x,y,z of this rank's particles are stored straight into the packed buffer of EVERY rank —
// either one multimem.st per value to the NVSwitch multicast address (the switch replicates it to all
// GPUs) or one plain store per peer through its mapped peer pointer — so no separate all-gather runs.
// The caller brackets steps with symmetric-memory barriers (see engine.py).
// ---------------------------------------------------------------------------
struct PeerPtrs {
    unsigned char* p[8];
    int count;
};

__device__ __forceinline__ void multimem_st(float* addr, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_st(double* addr, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

template <typename AT, typename T>
__global__ void kick_drift_push_kernel(double* __restrict__ pos, double* __restrict__ vel,
                                       const AT* __restrict__ acc, const double* __restrict__ acc_ext,
                                       const PeerPtrs peers, unsigned char* __restrict__ mc, long long first,
                                       long long n, double dt_half, double dt) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double a = static_cast<double>(acc[3 * i + d]);
        if (acc_ext) a += acc_ext[3 * i + d];
        const double v = vel[3 * i + d] + a * dt_half;
        vel[3 * i + d] = v;
        p[d] = pos[3 * i + d] + v * dt;
        pos[3 * i + d] = p[d];
    }
    constexpr int TB = (sizeof(T) == 8) ? TILE_BYTES_F64 : TILE_BYTES_F32;
    const long long g = first + i;
    const long long t = g / TJ;
    const int k       = static_cast<int>(g - t * TJ);
    const size_t off  = static_cast<size_t>(t) * TB;
    if (mc) {
        T* tb = reinterpret_cast<T*>(mc + off);
        multimem_st(tb + COMP_X * TJ + k, static_cast<T>(p[0]));
        multimem_st(tb + COMP_Y * TJ + k, static_cast<T>(p[1]));
        multimem_st(tb + COMP_Z * TJ + k, static_cast<T>(p[2]));
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r >= peers.count) break;
            T* tb = reinterpret_cast<T*>(peers.p[r] + off);
            tb[COMP_X * TJ + k] = static_cast<T>(p[0]);
            tb[COMP_Y * TJ + k] = static_cast<T>(p[1]);
            tb[COMP_Z * TJ + k] = static_cast<T>(p[2]);
        }
    }
}

template <typename AT>
__global__ void kick_kernel(double* __restrict__ vel, const AT* __restrict__ acc,
                            const double* __restrict__ acc_ext, long long n3, double dt_half) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n3) return;
    double a = static_cast<double>(acc[e]);
    if (acc_ext) a += acc_ext[e];
    vel[e] += a * dt_half;
}

// ---------------------------------------------------------------------------
// energy / momentum reduction (deterministic two-stage)
// ---------------------------------------------------------------------------
constexpr int ERED_THREADS = 256;
constexpr int ERED_PER_BLOCK = 4096;

__global__ void energy_stage1(const double* __restrict__ vel, const void* __restrict__ mass, int mass_dt,
                              const void* __restrict__ phi, int phi_dt, const void* __restrict__ acc, int acc_dt,
                              long long n, double* __restrict__ partials) {
    __shared__ double sh[8][ERED_THREADS / 32];
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long base = static_cast<long long>(blockIdx.x) * ERED_PER_BLOCK;
    for (int o = threadIdx.x; o < ERED_PER_BLOCK; o += ERED_THREADS) {
        const long long i = base + o;
        if (i >= n) break;
        const double m  = load_as_double(mass, mass_dt, i);
        const double vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        v[0] += 0.5 * m * (vx * vx + vy * vy + vz * vz);
        if (phi) v[1] += 0.5 * m * load_as_double(phi, phi_dt, i);
        v[2] += m * vx;
        v[3] += m * vy;
        v[4] += m * vz;
        if (acc) {
            v[5] += m * load_as_double(acc, acc_dt, 3 * i);
            v[6] += m * load_as_double(acc, acc_dt, 3 * i + 1);
            v[7] += m * load_as_double(acc, acc_dt, 3 * i + 2);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], off);
        if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double s = 0.0;
        for (int w = 0; w < ERED_THREADS / 32; ++w) s += sh[threadIdx.x][w];
        partials[static_cast<size_t>(blockIdx.x) * 8 + threadIdx.x] = s;
    }
}

__global__ void energy_stage2(const double* __restrict__ partials, int n_blocks, double* __restrict__ out) {
    // 8 warps, one per quantity; each warp adds its column in fixed order
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < n_blocks; b += 32) s += partials[static_cast<size_t>(b) * 8 + q];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) out[q] = s;
}

}  // namespace nb

// ===========================================================================
// host-side launchers (called from nb_capi.cu)
// ===========================================================================
using namespace nb;

#define NB_CUDA_OK(expr)                    \
    do {                                    \
        cudaError_t e_ = (expr);            \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

int nb_launch_pack(int prec, int kernel_id, void* packed, long long first, const void* pos, int pos_dt,
                   const void* mass, int mass_dt, double mass_scalar, const void* h, int h_dt, double h_scalar,
                   long long n, long long n_pad, double eps2, cudaStream_t st) {
    if (n_pad <= 0) return 0;
    const int add_eps = (kernel_id == NB_KERNEL_PLUMMER || kernel_id == NB_KERNEL_DEHNEN_K1 ||
                         kernel_id == NB_KERNEL_DEHNEN_K2);
    const unsigned blocks = static_cast<unsigned>(n_pad / TJ);
    const long long t0 = first / TJ;
    const bool f64 = (prec == NB_PREC_F64);
    unsigned char* tiles = static_cast<unsigned char*>(packed) + static_cast<size_t>(t0) * (f64 ? TILE_BYTES_F64 : TILE_BYTES_F32);
    if (f64)
        pack_kernel<double><<<blocks, TJ, 0, st>>>(tiles, pos, pos_dt, mass, mass_dt, mass_scalar, h, h_dt, h_scalar,
                                                   n, eps2, add_eps);
    else
        pack_kernel<float><<<blocks, TJ, 0, st>>>(tiles, pos, pos_dt, mass, mass_dt, mass_scalar, h, h_dt, h_scalar,
                                                  n, eps2, add_eps);
    NB_CUDA_OK(cudaGetLastError());
    return 0;
}

int nb_launch_finish(int prec, long long n_i, int grid, int IB, int n_sweeps, const int* n_jt, const void* ws,
                     size_t slot_bytes, size_t part_off, double G, void* acc, void* phi, cudaStream_t st) {
    if (n_i <= 0) return 0;
    FinishArgs a;
    a.ws = static_cast<const unsigned char*>(ws);
    a.slot_bytes = slot_bytes;
    a.part_off = part_off;
    a.n_sweeps = n_sweeps;
    for (int s = 0; s < 4; ++s) a.n_jt[s] = s < n_sweeps ? n_jt[s] : 0;
    a.grid = grid;
    a.IB = IB;
    a.n_ib = static_cast<int>((n_i + IB - 1) / IB);
    a.n_i = n_i;
    a.G = G;
    // small N: 4 threads per i and small blocks (latency-bound piece walk); large N: one thread per i
    const bool small = n_i < 148 * 256;
    const int threads = small ? 64 : 256;
    const int tpi = small ? 4 : 1;
    const unsigned blocks = static_cast<unsigned>((n_i * tpi + threads - 1) / threads);
#define NB_FINISH(PT, OT)                                                                                         \
    do {                                                                                                          \
        if (small) finish_kernel<PT, OT, 4><<<blocks, threads, 0, st>>>(a, static_cast<OT*>(acc), static_cast<OT*>(phi)); \
        else finish_kernel<PT, OT, 1><<<blocks, threads, 0, st>>>(a, static_cast<OT*>(acc), static_cast<OT*>(phi));       \
    } while (0)
    if (prec == NB_PREC_F32) NB_FINISH(float4, float);
    else if (prec == NB_PREC_F64) NB_FINISH(double4, double);
    else NB_FINISH(double4, float);
#undef NB_FINISH
    NB_CUDA_OK(cudaGetLastError());
    return 0;
}

int nb_launch_sym_finish(int prec, const void* ws, long long first, long long n_i, double G, void* acc, void* phi,
                         cudaStream_t st) {
    if (n_i <= 0) return 0;
    const int threads = n_i >= 148 * 256 ? 256 : 64;
    const unsigned blocks = static_cast<unsigned>((n_i + threads - 1) / threads);
    const double4* w = static_cast<const double4*>(ws);
    if (prec == NB_PREC_F64)
        sym_finish_kernel<double><<<blocks, threads, 0, st>>>(w, first, n_i, G, static_cast<double*>(acc),
                                                              static_cast<double*>(phi));
    else
        sym_finish_kernel<float><<<blocks, threads, 0, st>>>(w, first, n_i, G, static_cast<float*>(acc),
                                                             static_cast<float*>(phi));
    NB_CUDA_OK(cudaGetLastError());
    return 0;
}

int nb_launch_sym_finish_peers(int prec, const void* const* peer_ws, int n_peers, long long first, long long n_i,
                               double G, void* acc, void* phi, cudaStream_t st) {
    if (n_i <= 0) return 0;
    PeerWs pw;
    pw.count = n_peers;
    for (int r = 0; r < 8; ++r) pw.p[r] = r < n_peers ? static_cast<const double4*>(peer_ws[r]) : nullptr;
    const int threads = 128;
    const unsigned blocks = static_cast<unsigned>((n_i + threads - 1) / threads);
    if (prec == NB_PREC_F64)
        sym_finish_peers_kernel<double><<<blocks, threads, 0, st>>>(pw, first, n_i, G, static_cast<double*>(acc),
                                                                    static_cast<double*>(phi));
    else
        sym_finish_peers_kernel<float><<<blocks, threads, 0, st>>>(pw, first, n_i, G, static_cast<float*>(acc),
                                                                   static_cast<float*>(phi));
    NB_CUDA_OK(cudaGetLastError());
    return 0;
}

int nb_launch_kick_drift_pack(int prec, double* pos, double* vel, const void* acc, int acc_dt,
                              const double* acc_ext, void* packed, long long first, long long n, double dt_half,
                              double dt, cudaStream_t st) {
    if (n <= 0) return 0;
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((n + threads - 1) / threads);
    if (prec == NB_PREC_F64) {
        double* tiles = static_cast<double*>(packed);
        if (acc_dt == NB_DT_F64)
            kick_drift_pack_kernel<double, double><<<blocks, threads, 0, st>>>(
                pos, vel, static_cast<const double*>(acc), acc_ext, tiles, first, n, dt_half, dt);
        else
            kick_drift_pack_kernel<float, double><<<blocks, threads, 0, st>>>(
                pos, vel, static_cast<const float*>(acc), acc_ext, tiles, first, n, dt_half, dt);
    } else {
        float* tiles = static_cast<float*>(packed);
        if (acc_dt == NB_DT_F64)
            kick_drift_pack_kernel<double, float><<<blocks, threads, 0, st>>>(
                pos, vel, static_cast<const double*>(acc), acc_ext, tiles, first, n, dt_half, dt);
        else
            kick_drift_pack_kernel<float, float><<<blocks, threads, 0, st>>>(
                pos, vel, static_cast<const float*>(acc), acc_ext, tiles, first, n, dt_half, dt);
    }
    NB_CUDA_OK(cudaGetLastError());
    return 0;
}

int nb_launch_kick_drift_push(int prec, double* pos, double* vel, const void* acc, int acc_dt,
                              const double* acc_ext, void* const* peer_packed, int n_peers, void* mc_packed,
                              long long first, long long n, double dt_half, double dt, cudaStream_t st) {
    if (n <= 0) return 0;
    PeerPtrs pp;
    pp.count = n_peers;
    for (int r = 0; r < 8; ++r) pp.p[r] = r < n_peers ? static_cast<unsigned char*>(peer_packed[r]) : nullptr;
    unsigned char* mc = static_cast<unsigned char*>(mc_packed);
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((n + threads - 1) / threads);
    const bool f64 = (prec == NB_PREC_F64), a64 = (acc_dt == NB_DT_F64);
    if (f64 && a64)
        kick_drift_push_kernel<double, double><<<blocks, threads, 0, st>>>(pos, vel, static_cast<const double*>(acc),
                                                                           acc_ext, pp, mc, first, n, dt_half, dt);
    else if (f64)
        kick_drift_push_kernel<float, double><<<blocks, threads, 0, st>>>(pos, vel, static_cast<const float*>(acc),
                                                                          acc_ext, pp, mc, first, n, dt_half, dt);
    else if (a64)
        kick_drift_push_kernel<double, float><<<blocks, threads, 0, st>>>(pos, vel, static_cast<const double*>(acc),
                                                                          acc_ext, pp, mc, first, n, dt_half, dt);
    else
        kick_drift_push_kernel<float, float><<<blocks, threads, 0, st>>>(pos, vel, static_cast<const float*>(acc),
                                                                         acc_ext, pp, mc, first, n, dt_half, dt);
    NB_CUDA_OK(cudaGetLastError());
    return 0;
}

int nb_launch_kick(double* vel, const void* acc, int acc_dt, const double* acc_ext, long long n, double dt_half,
                   cudaStream_t st) {
    if (n <= 0) return 0;
    const long long n3 = 3 * n;
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((n3 + threads - 1) / threads);
    if (acc_dt == NB_DT_F64)
        kick_kernel<double><<<blocks, threads, 0, st>>>(vel, static_cast<const double*>(acc), acc_ext, n3, dt_half);
    else
        kick_kernel<float><<<blocks, threads, 0, st>>>(vel, static_cast<const float*>(acc), acc_ext, n3, dt_half);
    NB_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t nb_energy_ws_bytes(long long n) {
    const long long blocks = (n + ERED_PER_BLOCK - 1) / ERED_PER_BLOCK;
    return static_cast<size_t>(blocks > 0 ? blocks : 1) * 8 * sizeof(double);
}

int nb_launch_energy(const double* vel, const void* mass, int mass_dt, const void* phi, int phi_dt,
                     const void* acc, int acc_dt, long long n, double* partials, double* out, cudaStream_t st) {
    const int blocks = static_cast<int>((n + ERED_PER_BLOCK - 1) / ERED_PER_BLOCK);
    if (blocks > 0)
        energy_stage1<<<blocks, ERED_THREADS, 0, st>>>(vel, mass, mass_dt, phi, phi_dt, acc, acc_dt, n, partials);
    energy_stage2<<<1, 256, 0, st>>>(partials, blocks, out);
    NB_CUDA_OK(cudaGetLastError());
    return 0;
}
