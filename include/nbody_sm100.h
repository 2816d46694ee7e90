/*
 * nbody_sm100.h — C ABI of libnbody_sm100.so
 *
 * B200 (sm_100a) direct-summation N-body force/potential engine.  This is the
 * drop-in boundary for the hot path of appy2806/Nbody_streams: every entry
 * point below replaces one piece of the reference's CuPy RawKernel path and is
 * what a ctypes/cffi binding on the reference side would bind (see
 * INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer unless the comment says "host";
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns int: 0 = ok, >0 = cudaError_t, <0 = NB_E_* below;
 *     nothing is printed, nothing is allocated behind the caller's back and no
 *     call synchronises the device (the nb_ubench_* calls are the exception:
 *     they time themselves and therefore synchronise);
 *   - the caller owns all memory; scratch is a caller-provided workspace whose
 *     size comes from nb_workspace_bytes().
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   nbody_streams/cuda_kernels.py:90-101   nbody_forces_kernel_float4
 *   nbody_streams/cuda_kernels.py:254-265  nbody_forces_kahan_kernel_float4
 *   nbody_streams/cuda_kernels.py:439-448  nbody_potential_kernel_float4
 *   nbody_streams/cuda_kernels.py:591-600  nbody_potential_kahan_kernel_float4
 *   nbody_streams/cuda_kernels.py:755-769  nbody_forces_kernel<double>
 *   nbody_streams/cuda_kernels.py:1299-1310 nbody_potential_kernel<double>
 *   nbody_streams/fields.py:567-616        _prepare_gpu_arrays (SoA split, scalar broadcast)
 *   nbody_streams/fields.py:828-858        float4 packing + (N,3) output assembly
 *   nbody_streams/run.py:672-695           KDK kick / drift array expressions
 *   nbody_streams/run.py:619-623           energy reductions
 */
#ifndef NBODY_SM100_H
#define NBODY_SM100_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums (plain ints in the ABI) -------------------------------------- */

/* precision: arithmetic of the pair kernel (reference: fields.py:104-115) */
#define NB_PREC_F32       0  /* 'float32'        : f32 pair math, f32 running sums            */
#define NB_PREC_F32_COMP  1  /* 'float32_kahan'  : f32 pair math, per-tile f32 partials folded
                                                   into f64 running sums (compensated)          */
#define NB_PREC_F64       2  /* 'float64'        : f64 everywhere                               */

/* softening kernel ids (reference: fields.py:94-100) */
#define NB_KERNEL_NEWTONIAN 0
#define NB_KERNEL_PLUMMER   1
#define NB_KERNEL_DEHNEN_K1 2
#define NB_KERNEL_DEHNEN_K2 3
#define NB_KERNEL_SPLINE    4

/* what a sweep accumulates */
#define NB_MODE_FORCE 1
#define NB_MODE_POT   2
#define NB_MODE_BOTH  3

/* element type codes for caller arrays */
#define NB_DT_F32 0
#define NB_DT_F64 1

/* library error codes (negative; positive values are cudaError_t) */
#define NB_E_INVALID_ARG   (-1)
#define NB_E_WORKSPACE     (-2)  /* workspace too small                      */
#define NB_E_UNSUPPORTED   (-3)  /* combination not compiled                 */
#define NB_E_ARCH          (-4)  /* device is not sm_100                     */

/* ---- introspection ------------------------------------------------------ */

int         nb_version(void);                 /* 10000*major + 100*minor + patch */
const char *nb_error_string(int err);         /* host string, never NULL         */

/* Device properties of the CURRENT device (cudaGetDevice).  Replaces the CuPy
 * queries of fields.py:1149-1185 (get_gpu_info).  name: host buffer. */
int nb_device_info(int *sm_count, int *cc_major, int *cc_minor,
                   size_t *mem_total, size_t *mem_free,
                   int *clock_khz, char *name, int name_len);

/* ---- geometry of the packed particle buffer ----------------------------- */

/* Particles live on the device in a tile-major "packed" buffer that is both
 * the j-source the sweeps stream through shared memory (one bulk async copy
 * per tile) and the place i-particles are read from.  One tile holds
 * nb_tile_j() = 256 particles, structure-of-arrays inside the tile, followed by
 * 16 bytes of tile metadata:
 *     f32 precisions: [x][y][z][m][c] x 256 floats  + {c_uni, m_uni, c_max, 0}  = 5136 B
 *     f64 precision : [x][y][z][m][c] x 256 doubles + {c_uni, m_uni}            = 10256 B
 * c = h^2 (+ eps2 for kernels 1..3): the softening term the pair kernel takes
 * the max of (max-softening rule, reference cuda_kernels.py:150).  c_uni / m_uni
 * are the common c / mass of the tile or NaN when its particles differ (tiles
 * with uniform values take faster sweep paths).
 * Padding particles (index >= n) have m = 0, c = 0 and are parked at x = y = z =
 * 1e18 (f32) / 1e150 (f64), far from everything, so unmasked pair evaluations
 * against them are finite and multiply a zero mass. */
int    nb_tile_j(void);
int    nb_block_i(int prec);                       /* i-particles per CTA work unit  */
size_t nb_packed_bytes(int prec, long n);          /* bytes for n particles (padded) */
long   nb_padded_count(long n);                    /* n rounded up to whole tiles    */

/* Fill packed records [first, first + n_pad) (first must be tile aligned and
 * n_pad a multiple of the tile) from caller arrays holding n <= n_pad particles:
 *   pos  : (n,3) row-major, dtype pos_dt
 *   mass : (n) dtype mass_dt, or NULL -> mass_scalar broadcast (fields.py:601-602)
 *   h    : (n) dtype h_dt,    or NULL -> h_scalar broadcast    (fields.py:609-610)
 * Casts to the kernel dtype on the fly (fuses run.py:188 + fields.py:596-598,832-836). */
int nb_pack(int prec, int kernel_id, void *packed, long first,
            const void *pos, int pos_dt,
            const void *mass, int mass_dt, double mass_scalar,
            const void *h, int h_dt, double h_scalar,
            long n, long n_pad, double eps2, void *stream);

/* ---- the O(N^2) sweep --------------------------------------------------- */

/* Number of CTAs a sweep of this variant wants: SM count x resident CTAs/SM. */
int nb_sweep_grid(int prec, int kernel_id, int mode, int *grid);

/* Workspace for n_sweeps sweeps over n_i i-particles with the given grid. */
size_t nb_workspace_bytes(int prec, long n_i, int grid, int n_sweeps);

/* One sweep: for the i-particles with global indices [i_begin, i_begin+n_i)
 * accumulate the pair sums over the j-tiles [jt_begin, jt_end) U
 * [jt2_begin, jt2_end) of `packed` (the second range may be empty; used to
 * sweep "everything but my own shard" while an all-gather is in flight).
 * n_j is the number of REAL particles in `packed` (j >= n_j is padding; tile
 * ranges may include all-padding tiles as long as they lie inside the buffer
 * the caller allocated and nb_pack filled).
 * Partial sums go to slot `slot` of the workspace; nb_finish() combines slots.
 * Work is split stream-K style over `grid` persistent CTAs. */
int nb_sweep(int prec, int kernel_id, int mode,
             const void *packed, long n_j,
             int jt_begin, int jt_end, int jt2_begin, int jt2_end,
             long i_begin, long n_i, double eps2,
             void *workspace, size_t workspace_bytes, int slot, int grid,
             void *stream);

/* Combine the partial sums of n_sweeps slots in fixed (j-ascending, slot
 * ascending) order, scale by G and write
 *     acc (n_i,3) row-major and/or phi (n_i)   (either may be NULL)
 * in the kernel dtype (f32 for NB_PREC_F32*, f64 for NB_PREC_F64).
 * n_jt[k] (host array) = number of j-tiles slot k swept. */
int nb_finish(int prec, int kernel_id, int mode, long n_i, int grid, int n_sweeps,
              const int *n_jt, const void *workspace, double G,
              void *acc, void *phi, void *stream);

/* Convenience: pack + one full sweep + finish for n particles given plain
 * (n,3)/(n) device arrays — the call shape of the reference kernels
 * (cuda_kernels.py:90-101 etc.).  `packed` must hold nb_packed_bytes(prec,n),
 * `workspace` nb_workspace_bytes(prec,n,grid,1). */
int nb_direct(int prec, int kernel_id, int mode,
              const void *pos, int pos_dt,
              const void *mass, int mass_dt, double mass_scalar,
              const void *h, int h_dt, double h_scalar,
              long n, double G, double eps2,
              void *packed, void *workspace, size_t workspace_bytes, int grid,
              void *acc, void *phi, void *stream);

/* ---- the pair-symmetric O(N^2/2) sweep (large N) -------------------------- */

/* The kernel factor of a pair, K(r_ij^2, max(h_i,h_j)), is symmetric in i and j
 * (reference cuda_kernels.py:144-167 evaluates it once per ORDERED pair).  The
 * symmetric sweep evaluates every unordered pair once and applies it to both
 * particles: a_i += m_j K (r_j-r_i), a_j -= m_i K (r_j-r_i).  Work units are
 * (i-block of nb_sym_block_tiles() tiles) x (j-tile); both sides' partial sums
 * are added atomically, in double, into a workspace of 4 doubles per particle
 * SLOT (ax, ay, az, s_phi), so results are reproducible to rounding of those
 * double additions only (not bitwise; use nb_sweep for bitwise reproducibility).
 * NB_PREC_F32 and NB_PREC_F32_COMP share one implementation here (float pair
 * arithmetic, <= 256-term float partial sums, double above).
 * With fewer than 4 units per CTA a tile is swept in 2 or 4 parts, one unit each.
 * nb_sym_min_n(): particle count from which callers should prefer this path
 * (env NB_SYM_MIN_N, default 8192: measured crossover on one B200; below it the
 * one-sided sweep's finer work units fill the GPU better). */
long   nb_sym_min_n(void);
int    nb_set_sym_min_n(long n);
int    nb_sym_block_tiles(int prec);
int    nb_sym_grid(int prec, int kernel_id, int mode, int *grid);
size_t nb_sym_workspace_bytes(long n_slots);       /* 32 bytes per particle slot */

/* One symmetric sweep for the i-tiles [it_begin, it_end) of `packed`:
 *   tri != 0 : all pairs INSIDE the i-range (each once);
 *   [ra_begin,ra_end) U [rb_begin,rb_end): j-tile ranges OUTSIDE the i-range,
 *              swept by every i-block, results applied to both sides (other
 *              ranks' shards in a multi-GPU run);
 *   [rc_begin,rc_end): one more outside range, swept only by the i-blocks
 *              [ibc_begin, ibc_end) (splits the opposite shard of an even ring).
 * Tiles must exist in `packed` and hold real particles or nb_pack padding.
 * Sums are ADDED into workspace rows of the particles' global slots
 * (n_slots rows; zero_first != 0 clears all rows first, on the stream). */
int nb_sym_sweep(int prec, int kernel_id, int mode,
                 const void *packed, long n_j,
                 int it_begin, int it_end, int tri,
                 int ra_begin, int ra_end, int rb_begin, int rb_end,
                 int rc_begin, int rc_end, int ibc_begin, int ibc_end,
                 double eps2, void *workspace, size_t workspace_bytes,
                 long n_slots, int zero_first, int grid, void *stream);

/* acc (n_i,3) / phi (n_i) of the slots [first, first+n_i) from the workspace
 * sums (x G; kernel dtype).  Either output may be NULL. */
int nb_sym_finish(int prec, int mode, const void *workspace, long first, long n_i,
                  double G, void *acc, void *phi, void *stream);

/* Multi-GPU: finish fused with the reduction of the partial sums over NVLink peer memory.  Every rank's
 * nb_sym_sweep leaves partial sums for all particle slots in its own workspace; this call, on the owner of the
 * slots [first, first+n_i), adds those rows from the n_peers (<= 8) workspaces in the HOST array
 * peer_workspaces (mapped peer pointers, the caller's own included; fixed order = rank order), scales by G and
 * writes acc / phi.  Replaces reduce-scatter + nb_sym_finish by one kernel.  The caller orders the ranks
 * (barrier before: every rank's sweeps are done; barrier before the workspaces are cleared again). */
int nb_sym_finish_peers(int prec, int mode, const void *const *peer_workspaces, int n_peers,
                        long first, long n_i, double G, void *acc, void *phi, void *stream);

/* pack + symmetric sweep of all pairs + finish; same arguments as nb_direct,
 * workspace of nb_sym_workspace_bytes(nb_padded_count(n)), grid from nb_sym_grid. */
int nb_direct_sym(int prec, int kernel_id, int mode,
                  const void *pos, int pos_dt,
                  const void *mass, int mass_dt, double mass_scalar,
                  const void *h, int h_dt, double h_scalar,
                  long n, double G, double eps2,
                  void *packed, void *workspace, size_t workspace_bytes, int grid,
                  void *acc, void *phi, void *stream);

/* ---- fused KDK leapfrog pieces (reference run.py:666-695) ---------------- */

/* vel += (acc [+ acc_ext]) * dt_half ; pos += vel * dt ; then refresh x,y,z of
 * packed records [first, first+n) from the new positions (cast to the kernel
 * dtype).  pos, vel: (n,3) f64.  acc: (n,3) in acc_dt.  acc_ext: (n,3) f64 or
 * NULL.  packed may be NULL (no refresh). */
int nb_kick_drift_pack(int prec, double *pos, double *vel,
                       const void *acc, int acc_dt, const double *acc_ext,
                       void *packed, long first, long n,
                       double dt_half, double dt, void *stream);

/* Multi-GPU variant of nb_kick_drift_pack that fuses the exchange: the refreshed x,y,z are stored
 * straight into the packed buffer of every rank over NVLink — through `mc_packed`, the NVSwitch
 * multicast address of the (symmetric) packed buffers (one multimem.st per value, replicated by the
 * switch), or, when mc_packed is NULL, through the n_peers (<= 8) mapped peer pointers in the HOST
 * array peer_packed (which includes this rank's own buffer).  Replaces kick_drift_pack + all-gather.
 * The caller must order steps across ranks (barrier before: peers are done reading the old records;
 * barrier after: every rank's stores have landed). */
int nb_kick_drift_push(int prec, double *pos, double *vel,
                       const void *acc, int acc_dt, const double *acc_ext,
                       void *const *peer_packed, int n_peers, void *mc_packed,
                       long first, long n, double dt_half, double dt, void *stream);

/* vel += (acc [+ acc_ext]) * dt_half */
int nb_kick(double *vel, const void *acc, int acc_dt, const double *acc_ext,
            long n, double dt_half, void *stream);

/* out[0] = KE = 1/2 sum m v^2, out[1] = PE = 1/2 sum m phi (0 if phi NULL),
 * out[2..4] = sum m v, out[5..7] = sum m a (0 if acc NULL).  out: 8 doubles on
 * the device, overwritten.  Deterministic two-stage reduction; `partials`
 * must hold nb_energy_workspace_bytes(n). (reference run.py:619-623) */
size_t nb_energy_workspace_bytes(long n);
int nb_energy(const double *vel, const void *mass, int mass_dt,
              const void *phi, int phi_dt, const void *acc, int acc_dt,
              long n, double *partials, double *out, void *stream);

/* ---- tuning (not part of the reference-facing surface) --------------------- */

/* Select one of the compiled register-blocking variants of the plummer/spline
 * sweeps (f32: 0..3, f64: 0..2) or -1 = automatic, the default: the measured
 * best variant per (kernel, mode) — 0 except float32 plummer force-only /
 * potential-only (2).  NB_F32_VARIANT / NB_F64_VARIANT in the
 * environment set the initial value.  Changes nb_sweep_grid() results.
 * nb_get_variant returns the current setting (-1 for automatic). */
int nb_set_variant(int prec, int variant);
int nb_get_variant(int prec);

/* ---- peak microbenchmarks (roofline denominators) ------------------------ */

/* which: 0 FFMA (scalar), 1 FFMA2 (packed f32x2), 2 FFMA2 with three fresh
 * operands, 3 FADD2, 4 FMUL2, 5 MUFU.RSQ, 6 packed pair-kernel instruction mix,
 * 7 DFMA, 8 scalar pair-kernel mix.  Runs on the current device, synchronises,
 * returns lane-operations per second (an FFMA2 counts 2 per lane) and the
 * elapsed milliseconds of the timed launch. */
int nb_ubench(int which, int iters, double *lane_ops_per_s, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_SM100_H */
