// nb_capi.cu — extern "C" entry points of libnbody_sm100.so (see include/nbody_sm100.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nb_dispatch.cuh"
#include "nb_sym_dispatch.cuh"

// launchers defined in nb_misc.cu / nb_ubench.cu
int nb_launch_pack(int prec, int kernel_id, void* packed, long long first, const void* pos, int pos_dt,
                   const void* mass, int mass_dt, double mass_scalar, const void* h, int h_dt, double h_scalar,
                   long long n, long long n_pad, double eps2, cudaStream_t st);
int nb_launch_finish(int prec, long long n_i, int grid, int IB, int n_sweeps, const int* n_jt, const void* ws,
                     size_t slot_bytes, size_t part_off, double G, void* acc, void* phi, cudaStream_t st);
int nb_launch_sym_finish(int prec, const void* ws, long long first, long long n_i, double G, void* acc, void* phi,
                         cudaStream_t st);
int nb_launch_sym_finish_peers(int prec, const void* const* peer_ws, int n_peers, long long first, long long n_i,
                               double G, void* acc, void* phi, cudaStream_t st);
int nb_launch_kick_drift_pack(int prec, double* pos, double* vel, const void* acc, int acc_dt,
                              const double* acc_ext, void* packed, long long first, long long n, double dt_half,
                              double dt, cudaStream_t st);
int nb_launch_kick(double* vel, const void* acc, int acc_dt, const double* acc_ext, long long n, double dt_half,
                   cudaStream_t st);
int nb_launch_kick_drift_push(int prec, double* pos, double* vel, const void* acc, int acc_dt,
                              const double* acc_ext, void* const* peer_packed, int n_peers, void* mc_packed,
                              long long first, long long n, double dt_half, double dt, cudaStream_t st);
size_t nb_energy_ws_bytes(long long n);
int nb_launch_energy(const double* vel, const void* mass, int mass_dt, const void* phi, int phi_dt,
                     const void* acc, int acc_dt, long long n, double* partials, double* out, cudaStream_t st);
int nb_ubench_impl(int which, int iters, double* lane_ops_per_s, double* ms_out);

using namespace nb;

namespace {

constexpr int VARIANT_UNREAD = -2;   // NB_F32_VARIANT / NB_F64_VARIANT not looked at yet
constexpr int VARIANT_AUTO   = -1;   // pick per (kernel, mode): the measured best (see eff_variant)
int g_variant_f32 = VARIANT_UNREAD;
int g_variant_f64 = VARIANT_UNREAD;

int env_variant(const char* name, int nmax) {
    const char* s = std::getenv(name);
    if (!s || !*s) return VARIANT_AUTO;
    const int v = std::atoi(s);
    return (v >= 0 && v < nmax) ? v : VARIANT_AUTO;
}
int cur_variant_f32() {
    if (g_variant_f32 == VARIANT_UNREAD) g_variant_f32 = env_variant("NB_F32_VARIANT", NV_F32);
    return g_variant_f32;
}
int cur_variant_f64() {
    if (g_variant_f64 == VARIANT_UNREAD) g_variant_f64 = env_variant("NB_F64_VARIANT", NV_F64);
    return g_variant_f64;
}

bool valid_prec(int p) { return p == NB_PREC_F32 || p == NB_PREC_F32_COMP || p == NB_PREC_F64; }
bool valid_mode(int m) { return m == NB_MODE_FORCE || m == NB_MODE_POT || m == NB_MODE_BOTH; }
bool valid_kid(int k) { return k >= 0 && k <= 4; }

// tuning variants other than 0 exist only for plummer and spline
int eff_variant(int prec, int kid, int mode) {
    const bool allv = (kid == NB_KERNEL_PLUMMER || kid == NB_KERNEL_SPLINE);
    if (!allv) return 0;
    const int v = prec == NB_PREC_F64 ? cur_variant_f64() : cur_variant_f32();
    if (v != VARIANT_AUTO) return v;
    // measured on B200 with this build (profiles/r01_variant_sweep.log): variant 0 everywhere except
    //   float32 plummer force-only / potential-only: 1 i-pair per thread at 8 CTAs/SM, +4.5 % / +4 %
    //   (the fused force+potential loop needs the registers of the 2-pair blocking and stays on variant 0).
    // float64: the three variants are within 2-3 % of each other and their order changes between builds
    // (spline: variant 1 was +3 % before the division-free exact path, -2 % after) -> variant 0.
    if (prec == NB_PREC_F64) return 0;
    return (kid == NB_KERNEL_PLUMMER && mode != NB_MODE_BOTH) ? 2 : 0;
}
int block_i_for(int prec, int kid, int mode) {
    const int v = eff_variant(prec, kid, mode);
    return prec == NB_PREC_F64 ? THREADS * variant_f64(v).ni : THREADS * 2 * variant_f32(v).np;
}
// workspace sizing must not depend on the softening kernel: use the largest i-block of the precision
int block_i_max(int prec) {
    int m = 0;
    if (prec == NB_PREC_F64) {
        for (int v = 0; v < NV_F64; ++v) m = max(m, THREADS * variant_f64(v).ni);
    } else {
        for (int v = 0; v < NV_F32; ++v) m = max(m, THREADS * 2 * variant_f32(v).np);
    }
    return m;
}

int dispatch(int prec, int kid, int mode, int op, const SweepArgs* a, int grid, int* out, cudaStream_t st) {
    const int v = eff_variant(prec, kid, mode);
    if (prec == NB_PREC_F64) {
        switch (kid) {
            case 0: return nb_dispatch_f64_k0(mode, v, op, a, grid, out, st);
            case 1: return nb_dispatch_f64_k1(mode, v, op, a, grid, out, st);
            case 2: return nb_dispatch_f64_k2(mode, v, op, a, grid, out, st);
            case 3: return nb_dispatch_f64_k3(mode, v, op, a, grid, out, st);
            case 4: return nb_dispatch_f64_k4(mode, v, op, a, grid, out, st);
        }
    } else {
        const int acc = (prec == NB_PREC_F32_COMP) ? 1 : 0;
        switch (kid) {
            case 0: return nb_dispatch_f32_k0(mode, acc, v, op, a, grid, out, st);
            case 1: return nb_dispatch_f32_k1(mode, acc, v, op, a, grid, out, st);
            case 2: return nb_dispatch_f32_k2(mode, acc, v, op, a, grid, out, st);
            case 3: return nb_dispatch_f32_k3(mode, acc, v, op, a, grid, out, st);
            case 4: return nb_dispatch_f32_k4(mode, acc, v, op, a, grid, out, st);
        }
    }
    return NB_E_INVALID_ARG;
}

int dispatch_sym(int prec, int kid, int mode, int op, const SymArgs* a, int grid, int* out, cudaStream_t st) {
    switch (kid) {
        case 0: return nb_dispatch_sym_k0(prec, mode, op, a, grid, out, st);
        case 1: return nb_dispatch_sym_k1(prec, mode, op, a, grid, out, st);
        case 2: return nb_dispatch_sym_k2(prec, mode, op, a, grid, out, st);
        case 3: return nb_dispatch_sym_k3(prec, mode, op, a, grid, out, st);
        case 4: return nb_dispatch_sym_k4(prec, mode, op, a, grid, out, st);
    }
    return NB_E_INVALID_ARG;
}
int sym_tpb(int prec) { return prec == NB_PREC_F64 ? SYM_TPB_F64 : SYM_TPB_F32; }

constexpr long SYM_MIN_N_DEFAULT = 8192;   // below this the one-sided sweep's finer work units win (r02_n_sweep.md)
long g_sym_min_n = -1;

struct SlotLayout {
    size_t full_bytes, part_off, slot_bytes;
};
SlotLayout slot_layout(int prec, long long n_i, int grid) {
    const size_t eb = partial_elem_bytes(prec);
    const size_t IBm = static_cast<size_t>(block_i_max(prec));
    // full[] is indexed by ib*IB+li with the ACTUAL IB; rounding n_i up to the largest IB covers every variant
    const size_t n_full = ((static_cast<size_t>(n_i) + IBm - 1) / IBm) * IBm;
    SlotLayout L;
    L.full_bytes = n_full * eb;
    L.part_off   = (L.full_bytes + 255) & ~static_cast<size_t>(255);
    L.slot_bytes = L.part_off + ((static_cast<size_t>(grid) * 2 * IBm * eb + 255) & ~static_cast<size_t>(255));
    return L;
}

}  // namespace

extern "C" {

int nb_version(void) { return 10000 * 0 + 100 * 1 + 0; }

const char* nb_error_string(int err) {
    if (err == 0) return "ok";
    if (err > 0) return cudaGetErrorString(static_cast<cudaError_t>(err));
    switch (err) {
        case NB_E_INVALID_ARG: return "nbody_sm100: invalid argument";
        case NB_E_WORKSPACE: return "nbody_sm100: workspace too small";
        case NB_E_UNSUPPORTED: return "nbody_sm100: combination not compiled into this build";
        case NB_E_ARCH: return "nbody_sm100: device is not sm_100 (B200)";
    }
    return "nbody_sm100: unknown error";
}

int nb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* mem_total, size_t* mem_free,
                   int* clock_khz, char* name, int name_len) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
        *clock_khz = khz;
    }
    if (mem_total || mem_free) {
        size_t f = 0, t = 0;
        e = cudaMemGetInfo(&f, &t);
        if (e != cudaSuccess) return (int)e;
        if (mem_total) *mem_total = t;
        if (mem_free) *mem_free = f;
    }
    if (name && name_len > 0) {
        std::strncpy(name, p.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return 0;
}

/* tuning (not part of the reference-facing surface): select a compiled variant */
int nb_set_variant(int prec, int variant) {
    if (prec == NB_PREC_F64) {
        if (variant < VARIANT_AUTO || variant >= NV_F64) return NB_E_INVALID_ARG;
        g_variant_f64 = variant;
    } else {
        if (variant < VARIANT_AUTO || variant >= NV_F32) return NB_E_INVALID_ARG;
        g_variant_f32 = variant;
    }
    return 0;
}
int nb_get_variant(int prec) { return prec == NB_PREC_F64 ? cur_variant_f64() : cur_variant_f32(); }

int nb_tile_j(void) { return TJ; }
int nb_block_i(int prec) { return block_i_max(prec); }
long nb_padded_count(long n) { return ((n + TJ - 1) / TJ) * TJ; }
size_t nb_packed_bytes(int prec, long n) {
    const size_t nt = static_cast<size_t>(nb_padded_count(n)) / TJ;
    return nt * (prec == NB_PREC_F64 ? TILE_BYTES_F64 : TILE_BYTES_F32);
}

int nb_pack(int prec, int kernel_id, void* packed, long first, const void* pos, int pos_dt, const void* mass,
            int mass_dt, double mass_scalar, const void* h, int h_dt, double h_scalar, long n, long n_pad,
            double eps2, void* stream) {
    if (!valid_prec(prec) || !valid_kid(kernel_id) || !packed || n < 0 || n_pad < n) return NB_E_INVALID_ARG;
    if (first % TJ != 0 || n_pad % TJ != 0) return NB_E_INVALID_ARG;
    if (n > 0 && !pos) return NB_E_INVALID_ARG;
    return nb_launch_pack(prec, kernel_id, packed, first, pos, pos_dt, mass, mass_dt, mass_scalar, h, h_dt,
                          h_scalar, n, n_pad, eps2, static_cast<cudaStream_t>(stream));
}

int nb_sweep_grid(int prec, int kernel_id, int mode, int* grid) {
    if (!valid_prec(prec) || !valid_kid(kernel_id) || !valid_mode(mode) || !grid) return NB_E_INVALID_ARG;
    int dev = 0, sms = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return NB_E_ARCH;
    int per_sm = 0;
    const int rc = dispatch(prec, kernel_id, mode, OP_QUERY, nullptr, 0, &per_sm, 0);
    if (rc) return rc;
    if (per_sm < 1) per_sm = 1;
    *grid = sms * per_sm;
    return 0;
}

size_t nb_workspace_bytes(int prec, long n_i, int grid, int n_sweeps) {
    if (!valid_prec(prec) || n_i < 0 || grid <= 0 || n_sweeps <= 0) return 0;
    return slot_layout(prec, n_i, grid).slot_bytes * static_cast<size_t>(n_sweeps);
}

int nb_sweep(int prec, int kernel_id, int mode, const void* packed, long n_j, int jt_begin, int jt_end,
             int jt2_begin, int jt2_end, long i_begin, long n_i, double eps2, void* workspace,
             size_t workspace_bytes, int slot, int grid, void* stream) {
    if (!valid_prec(prec) || !valid_kid(kernel_id) || !valid_mode(mode)) return NB_E_INVALID_ARG;
    if (!packed || !workspace || n_j < 0 || n_i < 0 || i_begin < 0 || grid <= 0 || slot < 0 || slot >= 4)
        return NB_E_INVALID_ARG;
    if (jt_end < jt_begin || jt2_end < jt2_begin || jt_begin < 0 || jt2_begin < 0) return NB_E_INVALID_ARG;
    // Tile ranges may extend past the last real particle (multi-rank buffers are padded to equal shards);
    // such tiles hold zero-mass padding written by nb_pack and are skipped pair by pair.  The caller
    // guarantees that every tile index lies inside the packed buffer it allocated.
    const long long n_pad = nb_padded_count(n_j);
    const SlotLayout L = slot_layout(prec, n_i, grid);
    if (workspace_bytes < L.slot_bytes * static_cast<size_t>(slot + 1)) return NB_E_WORKSPACE;
    if (n_i == 0) return 0;

    const int IB = block_i_for(prec, kernel_id, mode);
    SweepArgs a;
    a.packed = static_cast<const unsigned char*>(packed);
    a.n_j = n_j;
    a.n_pad = n_pad;
    a.jtA0 = jt_begin;
    a.lenA = jt_end - jt_begin;
    a.jtB0 = jt2_begin;
    a.lenB = jt2_end - jt2_begin;
    a.n_jt = a.lenA + a.lenB;
    a.n_ib = static_cast<int>((n_i + IB - 1) / IB);
    a.i_begin = i_begin;
    a.n_i = n_i;
    a.split_log2 = pick_split_log2(a.n_ib, a.n_jt, grid);
    a.total = static_cast<long long>(a.n_ib) * (static_cast<long long>(a.n_jt) << a.split_log2);
    a.eps2 = eps2;
    unsigned char* base = static_cast<unsigned char*>(workspace) + L.slot_bytes * static_cast<size_t>(slot);
    a.full = base;
    a.part = base + L.part_off;
    if (a.n_jt == 0) return 0;   // nothing to sweep; nb_finish skips slots with n_jt == 0
    return dispatch(prec, kernel_id, mode, OP_LAUNCH, &a, grid, nullptr, static_cast<cudaStream_t>(stream));
}

/* kernel_id is needed because the i-block size of a slot depends on the tuning variant in force for it */
int nb_finish(int prec, int kernel_id, int mode, long n_i, int grid, int n_sweeps, const int* n_jt,
              const void* workspace, double G, void* acc, void* phi, void* stream) {
    if (!valid_prec(prec) || !valid_kid(kernel_id) || !valid_mode(mode) || n_i < 0 || grid <= 0 || n_sweeps < 1 ||
        n_sweeps > 4 || !n_jt || !workspace)
        return NB_E_INVALID_ARG;
    if (!(mode & NB_MODE_FORCE)) acc = nullptr;
    if (!(mode & NB_MODE_POT)) phi = nullptr;
    const SlotLayout L = slot_layout(prec, n_i, grid);
    // units per i-block of every slot, with the same sub-tile split nb_sweep chose for it
    const int IB = block_i_for(prec, kernel_id, mode);
    const long long n_ib = (n_i + IB - 1) / IB;
    int upb[4] = {0, 0, 0, 0};
    for (int s = 0; s < n_sweeps; ++s) upb[s] = n_jt[s] > 0 ? n_jt[s] << pick_split_log2(n_ib, n_jt[s], grid) : 0;
    return nb_launch_finish(prec, n_i, grid, IB, n_sweeps, upb, workspace, L.slot_bytes, L.part_off, G, acc, phi,
                            static_cast<cudaStream_t>(stream));
}

int nb_direct(int prec, int kernel_id, int mode, const void* pos, int pos_dt, const void* mass, int mass_dt,
              double mass_scalar, const void* h, int h_dt, double h_scalar, long n, double G, double eps2,
              void* packed, void* workspace, size_t workspace_bytes, int grid, void* acc, void* phi,
              void* stream) {
    if (n < 0) return NB_E_INVALID_ARG;
    if (n == 0) return 0;
    const long n_pad = nb_padded_count(n);
    int rc = nb_pack(prec, kernel_id, packed, 0, pos, pos_dt, mass, mass_dt, mass_scalar, h, h_dt, h_scalar, n,
                     n_pad, eps2, stream);
    if (rc) return rc;
    const int n_tiles = static_cast<int>(n_pad / TJ);
    rc = nb_sweep(prec, kernel_id, mode, packed, n, 0, n_tiles, 0, 0, 0, n, eps2, workspace, workspace_bytes, 0,
                  grid, stream);
    if (rc) return rc;
    return nb_finish(prec, kernel_id, mode, n, grid, 1, &n_tiles, workspace, G, acc, phi, stream);
}

/* ---- pair-symmetric sweeps ------------------------------------------------ */

long nb_sym_min_n(void) {
    if (g_sym_min_n < 0) {
        const char* s = std::getenv("NB_SYM_MIN_N");
        g_sym_min_n = (s && *s) ? std::atol(s) : SYM_MIN_N_DEFAULT;
        if (g_sym_min_n < 0) g_sym_min_n = 0;
    }
    return g_sym_min_n;
}
int nb_set_sym_min_n(long n) {
    if (n < 0) return NB_E_INVALID_ARG;
    g_sym_min_n = n;
    return 0;
}
int nb_sym_block_tiles(int prec) { return valid_prec(prec) ? sym_tpb(prec) : 0; }

int nb_sym_grid(int prec, int kernel_id, int mode, int* grid) {
    if (!valid_prec(prec) || !valid_kid(kernel_id) || !valid_mode(mode) || !grid) return NB_E_INVALID_ARG;
    int dev = 0, sms = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return NB_E_ARCH;
    int per_sm = 0;
    const int rc = dispatch_sym(prec, kernel_id, mode, SYM_OP_QUERY, nullptr, 0, &per_sm, 0);
    if (rc) return rc;
    if (per_sm < 1) per_sm = 1;
    *grid = sms * per_sm;
    return 0;
}

size_t nb_sym_workspace_bytes(long n_slots) { return n_slots > 0 ? static_cast<size_t>(n_slots) * 32 : 0; }

int nb_sym_sweep(int prec, int kernel_id, int mode, const void* packed, long n_j, int it_begin, int it_end, int tri,
                 int ra_begin, int ra_end, int rb_begin, int rb_end, int rc_begin, int rc_end, int ibc_begin,
                 int ibc_end, double eps2, void* workspace, size_t workspace_bytes, long n_slots, int zero_first,
                 int grid, void* stream) {
    if (!valid_prec(prec) || !valid_kid(kernel_id) || !valid_mode(mode)) return NB_E_INVALID_ARG;
    if (!packed || !workspace || n_j < 0 || grid <= 0 || n_slots < 0) return NB_E_INVALID_ARG;
    if (it_begin < 0 || it_end < it_begin || ra_begin < 0 || ra_end < ra_begin || rb_begin < 0 || rb_end < rb_begin ||
        rc_begin < 0 || rc_end < rc_begin || ibc_begin < 0 || ibc_end < ibc_begin)
        return NB_E_INVALID_ARG;
    if (workspace_bytes < nb_sym_workspace_bytes(n_slots)) return NB_E_WORKSPACE;
    // every tile touched must have workspace rows; tiles must hold real particles or sentinel padding
    const long max_tile = n_slots / TJ;
    if (it_end > max_tile || ra_end > max_tile || rb_end > max_tile || rc_end > max_tile) return NB_E_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (zero_first) {
        cudaError_t e = cudaMemsetAsync(workspace, 0, nb_sym_workspace_bytes(n_slots), st);
        if (e != cudaSuccess) return (int)e;
    }
    const int TPB = sym_tpb(prec);
    SymArgs a;
    a.packed = static_cast<const unsigned char*>(packed);
    a.n_j  = n_j;
    a.it0  = it_begin;
    a.nT   = it_end - it_begin;
    a.n_ib = (a.nT + TPB - 1) / TPB;
    a.tri  = tri ? 1 : 0;
    a.split_log2 = 0;
    a.rA0 = ra_begin; a.lenA = ra_end - ra_begin;
    a.rB0 = rb_begin; a.lenB = rb_end - rb_begin;
    a.rC0 = rc_begin; a.lenC = rc_end - rc_begin;
    a.ibC0 = ibc_begin < a.n_ib ? ibc_begin : a.n_ib;
    a.ibC1 = ibc_end < a.n_ib ? ibc_end : a.n_ib;
    if (a.ibC1 < a.ibC0) a.ibC1 = a.ibC0;
    a.eps2 = eps2;
    a.ws   = static_cast<double*>(workspace);
    if (a.n_ib == 0) return 0;
    const long long tile_units = prec == NB_PREC_F64 ? sym_prefix<SYM_TPB_F64>(a, a.n_ib) : sym_prefix<SYM_TPB_F32>(a, a.n_ib);
    if (tile_units <= 0) return 0;
    a.split_log2 = sym_pick_split_log2(tile_units, grid);
    a.total      = tile_units << a.split_log2;
    return dispatch_sym(prec, kernel_id, mode, SYM_OP_LAUNCH, &a, grid, nullptr, st);
}

int nb_sym_finish(int prec, int mode, const void* workspace, long first, long n_i, double G, void* acc, void* phi,
                  void* stream) {
    if (!valid_prec(prec) || !valid_mode(mode) || !workspace || first < 0 || n_i < 0) return NB_E_INVALID_ARG;
    if (!(mode & NB_MODE_FORCE)) acc = nullptr;
    if (!(mode & NB_MODE_POT)) phi = nullptr;
    return nb_launch_sym_finish(prec, workspace, first, n_i, G, acc, phi, static_cast<cudaStream_t>(stream));
}

int nb_sym_finish_peers(int prec, int mode, const void* const* peer_workspaces, int n_peers, long first, long n_i,
                        double G, void* acc, void* phi, void* stream) {
    if (!valid_prec(prec) || !valid_mode(mode) || !peer_workspaces || n_peers < 1 || n_peers > 8 || first < 0 || n_i < 0)
        return NB_E_INVALID_ARG;
    for (int r = 0; r < n_peers; ++r)
        if (!peer_workspaces[r]) return NB_E_INVALID_ARG;
    if (!(mode & NB_MODE_FORCE)) acc = nullptr;
    if (!(mode & NB_MODE_POT)) phi = nullptr;
    return nb_launch_sym_finish_peers(prec, peer_workspaces, n_peers, first, n_i, G, acc, phi,
                                      static_cast<cudaStream_t>(stream));
}

int nb_direct_sym(int prec, int kernel_id, int mode, const void* pos, int pos_dt, const void* mass, int mass_dt,
                  double mass_scalar, const void* h, int h_dt, double h_scalar, long n, double G, double eps2,
                  void* packed, void* workspace, size_t workspace_bytes, int grid, void* acc, void* phi,
                  void* stream) {
    if (n < 0) return NB_E_INVALID_ARG;
    if (n == 0) return 0;
    const long n_pad = nb_padded_count(n);
    int rc = nb_pack(prec, kernel_id, packed, 0, pos, pos_dt, mass, mass_dt, mass_scalar, h, h_dt, h_scalar, n,
                     n_pad, eps2, stream);
    if (rc) return rc;
    const int n_tiles = static_cast<int>(n_pad / TJ);
    rc = nb_sym_sweep(prec, kernel_id, mode, packed, n, 0, n_tiles, 1, 0, 0, 0, 0, 0, 0, 0, 0, eps2, workspace,
                      workspace_bytes, n_pad, 1, grid, stream);
    if (rc) return rc;
    return nb_sym_finish(prec, mode, workspace, 0, n, G, acc, phi, stream);
}

int nb_kick_drift_pack(int prec, double* pos, double* vel, const void* acc, int acc_dt, const double* acc_ext,
                       void* packed, long first, long n, double dt_half, double dt, void* stream) {
    if (!valid_prec(prec) || !pos || !vel || !acc || n < 0 || first < 0) return NB_E_INVALID_ARG;
    return nb_launch_kick_drift_pack(prec, pos, vel, acc, acc_dt, acc_ext, packed, first, n, dt_half, dt,
                                     static_cast<cudaStream_t>(stream));
}

int nb_kick_drift_push(int prec, double* pos, double* vel, const void* acc, int acc_dt, const double* acc_ext,
                       void* const* peer_packed, int n_peers, void* mc_packed, long first, long n, double dt_half,
                       double dt, void* stream) {
    if (!valid_prec(prec) || !pos || !vel || !acc || n < 0 || first < 0) return NB_E_INVALID_ARG;
    if (!mc_packed && (!peer_packed || n_peers < 1 || n_peers > 8)) return NB_E_INVALID_ARG;
    return nb_launch_kick_drift_push(prec, pos, vel, acc, acc_dt, acc_ext, peer_packed, n_peers, mc_packed, first, n,
                                     dt_half, dt, static_cast<cudaStream_t>(stream));
}

int nb_kick(double* vel, const void* acc, int acc_dt, const double* acc_ext, long n, double dt_half,
            void* stream) {
    if (!vel || !acc || n < 0) return NB_E_INVALID_ARG;
    return nb_launch_kick(vel, acc, acc_dt, acc_ext, n, dt_half, static_cast<cudaStream_t>(stream));
}

size_t nb_energy_workspace_bytes(long n) { return nb_energy_ws_bytes(n); }

int nb_energy(const double* vel, const void* mass, int mass_dt, const void* phi, int phi_dt, const void* acc,
              int acc_dt, long n, double* partials, double* out, void* stream) {
    if (!vel || !mass || !partials || !out || n < 0) return NB_E_INVALID_ARG;
    return nb_launch_energy(vel, mass, mass_dt, phi, phi_dt, acc, acc_dt, n, partials, out,
                            static_cast<cudaStream_t>(stream));
}

int nb_ubench(int which, int iters, double* lane_ops_per_s, double* ms) {
    return nb_ubench_impl(which, iters, lane_ops_per_s, ms);
}

}  // extern "C"
