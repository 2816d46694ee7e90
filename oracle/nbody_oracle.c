/*
 * nbody_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's Numba CPU direct-summation path, used
 * only as the parity checker (tests/, __graft_entry__.smoke()) and as the timed
 * CPU baseline (bench.py cpu_baseline / --impl reference).  Nothing under
 * nbody_streams_b200/ imports, links or executes this file.
 *
 * Pinned against the reference itself: tests/golden/ holds outputs of the
 * reference's compute_nbody_{forces,potential}_cpu generated in the build
 * container by tests/golden/make_golden.py; tests/test_oracle.py checks this
 * restatement against them (float64: <= 1e-13 of |a|, float32: <= 2e-6).
 *
 * What is restated (paths relative to the reference checkout):
 *   _get_force_kernel       nbody_streams/fields.py:122-189
 *   _compute_forces_cpu     nbody_streams/fields.py:191-256
 *   _get_potential_kernel   nbody_streams/fields.py:258-314
 *   _compute_potential_cpu  nbody_streams/fields.py:317-369
 * The wrappers' r_eps = dtype(1e-15) and the G multiply (fields.py:1083-1086,
 * :1141-1146) are applied by oracle/__init__.py.
 *
 * Type semantics follow Numba's typing of the reference source: in the
 * "float32" instantiation positions, masses, softenings, dx, r2 and sqrt() are
 * float, but every `1.0 / (...)`, the spline polynomial and the running sums
 * fx, fy, fz, pot_i are double (Python float literals are float64 in Numba),
 * and the result is rounded to float on store.
 *
 * Build: see oracle/Makefile.  -DNBO_FAST builds the throughput variant
 * (-O3 -ffast-math, AVX2/AVX-512 clones, self-skip hoisted out of the j loop),
 * numerically equivalent within rounding, used only for CPU timing.
 */
#include <math.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(NBO_FAST) && defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define NBO_CLONES __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define NBO_CLONES
#endif

#ifdef NBO_FAST
#define NBO_NAME(x) x##_fast
#else
#define NBO_NAME(x) x
#endif

/* ---- force kernels: fields.py:122-189 ---------------------------------- */

static inline double force_kernel_f64(double r2, double h, int kernel_id) {
    if (kernel_id == 0) return 1.0 / (r2 * sqrt(r2));
    if (kernel_id == 1) {
        const double denom = r2 + h * h;
        return 1.0 / (denom * sqrt(denom));
    }
    if (kernel_id == 2) {
        const double denom = r2 + h * h;
        const double sqrt_denom = sqrt(denom);
        const double term1 = 1.0 / (denom * sqrt_denom);
        const double term2 = 1.5 * h * h / (denom * denom * sqrt_denom);
        return term1 + term2;
    }
    if (kernel_id == 3) {
        const double denom = r2 + h * h;
        const double sqrt_denom = sqrt(denom);
        const double hh = h * h;
        const double term1 = 1.0 / (denom * sqrt_denom);
        const double term2 = 1.5 * hh / (denom * denom * sqrt_denom);
        const double term3 = 3.75 * hh * hh / (denom * denom * denom * sqrt_denom);
        return term1 + term2 + term3;
    }
    const double r = sqrt(r2);
    if (r > h) return 1.0 / (r * r * r);
    const double hinv = 1.0 / h;
    const double h3inv = hinv * hinv * hinv;
    const double q = r * hinv;
    const double q2 = q * q;
    if (kernel_id == 4) {
        if (q <= 0.5) return h3inv * (10.666666666666666 + q2 * (-38.4 + 32.0 * q));
        return h3inv * (21.333333333333333 - 48.0 * q + 38.4 * q2 - 10.666666666666667 * q2 * q -
                        0.06666666666666667 / (q2 * q));
    }
    return 0.0;
}

/* float32 instantiation: float inputs, double results (see header comment) */
static inline double force_kernel_f32(float r2, float h, int kernel_id) {
    if (kernel_id == 0) return 1.0 / (double)(r2 * sqrtf(r2));
    if (kernel_id == 1) {
        const float denom = r2 + h * h;
        return 1.0 / (double)(denom * sqrtf(denom));
    }
    if (kernel_id == 2) {
        const float denom = r2 + h * h;
        const float sqrt_denom = sqrtf(denom);
        const double term1 = 1.0 / (double)(denom * sqrt_denom);
        const double term2 = 1.5 * (double)h * (double)h / (double)(denom * denom * sqrt_denom);
        return term1 + term2;
    }
    if (kernel_id == 3) {
        const float denom = r2 + h * h;
        const float sqrt_denom = sqrtf(denom);
        const float hh = h * h;
        const double term1 = 1.0 / (double)(denom * sqrt_denom);
        const double term2 = 1.5 * (double)hh / (double)(denom * denom * sqrt_denom);
        const double term3 = 3.75 * (double)hh * (double)hh / (double)(denom * denom * denom * sqrt_denom);
        return term1 + term2 + term3;
    }
    const float r = sqrtf(r2);
    if (r > h) return 1.0 / (double)(r * r * r);
    const double hinv = 1.0 / (double)h;
    const double h3inv = hinv * hinv * hinv;
    const double q = (double)r * hinv;
    const double q2 = q * q;
    if (kernel_id == 4) {
        if (q <= 0.5) return h3inv * (10.666666666666666 + q2 * (-38.4 + 32.0 * q));
        return h3inv * (21.333333333333333 - 48.0 * q + 38.4 * q2 - 10.666666666666667 * q2 * q -
                        0.06666666666666667 / (q2 * q));
    }
    return 0.0;
}

/* ---- potential kernels: fields.py:258-314 -------------------------------- */

static inline double spline_phi(double r, double h) {
    if (h == 0.0 || r >= h) return -1.0 / r;
    const double hinv = 1.0 / h;
    const double q = r * hinv;
    if (q < 1e-8) return -2.8 * hinv;
    const double q2 = q * q;
    if (q <= 0.5) return (-2.8 + q2 * (5.33333333333333333 + q2 * q2 * (6.4 * q - 9.6))) * hinv;
    if (q <= 1.0)
        return (-3.2 + 0.066666666666666666666 / q +
                q2 * (10.666666666666666666666 + q * (-16.0 + q * (9.6 - 2.1333333333333333333333 * q)))) *
               hinv;
    return -1.0 / r;
}

static inline double potential_kernel_f64(double r2, double h, int kernel_id) {
    const double r = sqrt(r2);
    if (kernel_id == 0) return r > 0 ? -1.0 / r : 0.0;
    if (kernel_id == 1) return -1.0 / sqrt(r2 + h * h);
    if (kernel_id == 2) {
        const double h2 = h * h;
        const double denom = r2 + h2;
        const double sqrt_denom = sqrt(denom);
        const double inv_sqrt = 1.0 / sqrt_denom;
        const double inv_d32 = inv_sqrt / denom;
        return -inv_sqrt - 0.5 * h2 * inv_d32;
    }
    if (kernel_id == 3) {   /* NB: the reference CPU line :287 gives d^-2, not d^-5/2 */
        const double h2 = h * h;
        const double h4 = h2 * h2;
        const double denom = r2 + h2;
        const double sqrt_denom = sqrt(denom);
        const double inv_sqrt = 1.0 / sqrt_denom;
        const double inv_d32 = inv_sqrt / denom;
        const double inv_d52 = inv_d32 / sqrt_denom;
        return -inv_sqrt - 0.5 * h2 * inv_d32 - 0.375 * h4 * inv_d52;
    }
    if (kernel_id == 4) return spline_phi(r, h);
    return 0.0;
}

static inline double potential_kernel_f32(float r2, float h, int kernel_id) {
    const float r = sqrtf(r2);
    if (kernel_id == 0) return r > 0 ? -1.0 / (double)r : 0.0;
    if (kernel_id == 1) return -1.0 / (double)sqrtf(r2 + h * h);
    if (kernel_id == 2) {
        const float h2 = h * h;
        const float denom = r2 + h2;
        const float sqrt_denom = sqrtf(denom);
        const double inv_sqrt = 1.0 / (double)sqrt_denom;
        const double inv_d32 = inv_sqrt / (double)denom;
        return -inv_sqrt - 0.5 * (double)h2 * inv_d32;
    }
    if (kernel_id == 3) {
        const float h2 = h * h;
        const float h4 = h2 * h2;
        const float denom = r2 + h2;
        const float sqrt_denom = sqrtf(denom);
        const double inv_sqrt = 1.0 / (double)sqrt_denom;
        const double inv_d32 = inv_sqrt / (double)denom;
        const double inv_d52 = inv_d32 / (double)sqrt_denom;
        return -inv_sqrt - 0.5 * (double)h2 * inv_d32 - 0.375 * (double)h4 * inv_d52;
    }
    if (kernel_id == 4) return spline_phi((double)r, (double)h);
    return 0.0;
}

/* the dehnen_k2 potential as the reference GPU kernel computes it
 * (cuda_kernels.py:394-402, :1251-1259): the third term uses d^-5/2. */
static inline double potential_kernel_k2_gpu(double r2, double h) {
    const double h2 = h * h, h4 = h2 * h2;
    const double denom = r2 + h2;
    const double inv_sqrt = 1.0 / sqrt(denom);
    const double inv_d32 = inv_sqrt * inv_sqrt * inv_sqrt;
    const double inv_d52 = inv_d32 * inv_sqrt * inv_sqrt;
    return -inv_sqrt - 0.5 * h2 * inv_d32 - 0.375 * h4 * inv_d52;
}

/* ---- pair loops: fields.py:191-256, :317-369 ------------------------------ */
/* Rows [i0, i1) against all n particles; out has (i1-i0) rows.  The Numba
 * kernels are independent per i (prange), so a row subset is exact. */

#ifdef NBO_FAST
#define NBO_JLOOP(body)                       \
    for (long j = 0; j < i; ++j) { body }      \
    for (long j = i + 1; j < n; ++j) { body }
#else
#define NBO_JLOOP(body)            \
    for (long j = 0; j < n; ++j) { \
        if (i == j) continue;      \
        body                       \
    }
#endif

NBO_CLONES
void NBO_NAME(nbo_forces_f64)(const double* pos, const double* mass, const double* h, int kernel_id, double r_eps,
                              long n, long i0, long i1, double* out) {
#pragma omp parallel for schedule(static)
    for (long i = i0; i < i1; ++i) {
        double fx = 0.0, fy = 0.0, fz = 0.0;
        const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        const double hi = h[i];
        NBO_JLOOP(
            const double mj = mass[j]; const double hj = h[j];
            const double dx = pos[3 * j] - xi; const double dy = pos[3 * j + 1] - yi;
            const double dz = pos[3 * j + 2] - zi;
            const double r2 = dx * dx + dy * dy + dz * dz + r_eps * r_eps;
            const double h_ij = hi >= hj ? hi : hj;
            const double factor = mj * force_kernel_f64(r2, h_ij, kernel_id);
            fx += factor * dx; fy += factor * dy; fz += factor * dz;)
        out[3 * (i - i0)] = fx;
        out[3 * (i - i0) + 1] = fy;
        out[3 * (i - i0) + 2] = fz;
    }
}

NBO_CLONES
void NBO_NAME(nbo_forces_f32)(const float* pos, const float* mass, const float* h, int kernel_id, float r_eps,
                              long n, long i0, long i1, float* out) {
#pragma omp parallel for schedule(static)
    for (long i = i0; i < i1; ++i) {
        double fx = 0.0, fy = 0.0, fz = 0.0;
        const float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        const float hi = h[i];
        NBO_JLOOP(
            const float mj = mass[j]; const float hj = h[j];
            const float dx = pos[3 * j] - xi; const float dy = pos[3 * j + 1] - yi;
            const float dz = pos[3 * j + 2] - zi;
            const float r2 = dx * dx + dy * dy + dz * dz + r_eps * r_eps;
            const float h_ij = hi >= hj ? hi : hj;
            const double factor = (double)mj * force_kernel_f32(r2, h_ij, kernel_id);
            fx += factor * (double)dx; fy += factor * (double)dy; fz += factor * (double)dz;)
        out[3 * (i - i0)] = (float)fx;
        out[3 * (i - i0) + 1] = (float)fy;
        out[3 * (i - i0) + 2] = (float)fz;
    }
}

/* k2_gpu != 0: use the reference GPU formula for dehnen_k2 (see above) */
NBO_CLONES
void NBO_NAME(nbo_potential_f64)(const double* pos, const double* mass, const double* h, int kernel_id,
                                 double r_eps, long n, long i0, long i1, int k2_gpu, double* out) {
    const double eps2 = r_eps * r_eps;
#pragma omp parallel for schedule(static)
    for (long i = i0; i < i1; ++i) {
        const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        const double hi = h[i];
        double pot_i = 0.0;
        NBO_JLOOP(
            const double dx = pos[3 * j] - xi; const double dy = pos[3 * j + 1] - yi;
            const double dz = pos[3 * j + 2] - zi;
            const double r2 = dx * dx + dy * dy + dz * dz + eps2;
            const double hj = h[j];
            const double h_eff = hi >= hj ? hi : hj;
            const double kv = (kernel_id == 3 && k2_gpu) ? potential_kernel_k2_gpu(r2, h_eff)
                                                         : potential_kernel_f64(r2, h_eff, kernel_id);
            pot_i += mass[j] * kv;)
        out[i - i0] = pot_i;
    }
}

NBO_CLONES
void NBO_NAME(nbo_potential_f32)(const float* pos, const float* mass, const float* h, int kernel_id, float r_eps,
                                 long n, long i0, long i1, int k2_gpu, float* out) {
    const float eps2 = r_eps * r_eps;
#pragma omp parallel for schedule(static)
    for (long i = i0; i < i1; ++i) {
        const float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        const float hi = h[i];
        double pot_i = 0.0;
        NBO_JLOOP(
            const float dx = pos[3 * j] - xi; const float dy = pos[3 * j + 1] - yi;
            const float dz = pos[3 * j + 2] - zi;
            const float r2 = dx * dx + dy * dy + dz * dz + eps2;
            const float hj = h[j];
            const float h_eff = hi >= hj ? hi : hj;
            const double kv = (kernel_id == 3 && k2_gpu) ? potential_kernel_k2_gpu((double)r2, (double)h_eff)
                                                         : potential_kernel_f32(r2, h_eff, kernel_id);
            pot_i += (double)mass[j] * kv;)
        out[i - i0] = (float)pot_i;
    }
}

/* torchrun exports OMP_NUM_THREADS=1; the timing legs ask for all host cores explicitly */
void NBO_NAME(nbo_set_num_threads)(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int NBO_NAME(nbo_num_threads)(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
