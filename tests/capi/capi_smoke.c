/* capi_smoke.c — calls libnbody_sm100.so through its C ABI only (no Python, no torch): what a cgo / JNI /
 * ctypes binding on the reference side would do.  Built and run by tests/test_capi_c.py on the GPU box.
 *
 *   gcc capi_smoke.c -I../../include -I/usr/local/cuda/include -L<libdir> -lnbody_sm100 -L/usr/local/cuda/lib64 -lcudart -lm
 *
 * Computes Plummer-softened accelerations + potential for N particles in float32_kahan and float64 with
 * nb_direct (one-sided sweep) and nb_direct_sym (pair-symmetric sweep) and checks both against a plain double
 * loop (the reference's CPU algorithm, fields.py:220-256).
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "nbody_sm100.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %d at %s:%d\n", (int)e_, __FILE__, __LINE__); return 2; } } while (0)
#define CHECK_NB(x) do { int r_ = (x); if (r_ != 0) { printf("nb error %d (%s) at %s:%d\n", r_, nb_error_string(r_), __FILE__, __LINE__); return 3; } } while (0)

static double frand(unsigned long long* s) {
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)((*s >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53);
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 3001;
    const double G = 4.300917270069976e-06;
    unsigned long long seed = 12345;
    double *pos = malloc(sizeof(double) * 3 * n), *mass = malloc(sizeof(double) * n), *h = malloc(sizeof(double) * n);
    for (long i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) pos[3 * i + d] = 2.0 * frand(&seed) - 1.0;
        mass[i] = 1.0 + frand(&seed);
        h[i] = 0.01 + 0.04 * frand(&seed);
    }
    /* reference loop in double: a_i = G sum m_j (r^2 + max(h_i,h_j)^2 + eps^2)^-3/2 (r_j - r_i) */
    double *aref = calloc(3 * n, sizeof(double)), *pref = calloc(n, sizeof(double));
    for (long i = 0; i < n; ++i)
        for (long j = 0; j < n; ++j) {
            if (i == j) continue;
            const double dx = pos[3 * j] - pos[3 * i], dy = pos[3 * j + 1] - pos[3 * i + 1], dz = pos[3 * j + 2] - pos[3 * i + 2];
            const double hh = h[i] >= h[j] ? h[i] : h[j];
            const double d = dx * dx + dy * dy + dz * dz + 1e-30 + hh * hh;
            const double w = 1.0 / (d * sqrt(d));
            aref[3 * i] += G * mass[j] * w * dx; aref[3 * i + 1] += G * mass[j] * w * dy; aref[3 * i + 2] += G * mass[j] * w * dz;
            pref[i] -= G * mass[j] / sqrt(d);
        }

    int sm = 0, major = 0, minor = 0, clk = 0;
    size_t tot = 0, fre = 0;
    char name[128];
    CHECK_NB(nb_device_info(&sm, &major, &minor, &tot, &fre, &clk, name, sizeof name));
    printf("device %s sm_%d%d, %d SMs, lib version %d, tile %d\n", name, major, minor, sm, nb_version(), nb_tile_j());

    double *d_pos, *d_mass, *d_h;
    CHECK_CUDA(cudaMalloc((void**)&d_pos, sizeof(double) * 3 * n));
    CHECK_CUDA(cudaMalloc((void**)&d_mass, sizeof(double) * n));
    CHECK_CUDA(cudaMalloc((void**)&d_h, sizeof(double) * n));
    CHECK_CUDA(cudaMemcpy(d_pos, pos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_mass, mass, sizeof(double) * n, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_h, h, sizeof(double) * n, cudaMemcpyHostToDevice));
    cudaStream_t st;
    CHECK_CUDA(cudaStreamCreate(&st));

    int fail = 0;
    const int precs[2] = {NB_PREC_F32_COMP, NB_PREC_F64};
    for (int pp = 0; pp < 4; ++pp) {
        const int prec = precs[pp & 1];
        const int sym = pp >> 1;                 /* second pass: the pair-symmetric entry points */
        const size_t es = prec == NB_PREC_F64 ? 8 : 4;
        int grid = 0;
        CHECK_NB(sym ? nb_sym_grid(prec, NB_KERNEL_PLUMMER, NB_MODE_BOTH, &grid)
                     : nb_sweep_grid(prec, NB_KERNEL_PLUMMER, NB_MODE_BOTH, &grid));
        const size_t ws_bytes = sym ? nb_sym_workspace_bytes(nb_padded_count(n)) : nb_workspace_bytes(prec, n, grid, 1);
        void *packed, *ws, *d_acc, *d_phi;
        CHECK_CUDA(cudaMalloc(&packed, nb_packed_bytes(prec, n)));
        CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
        CHECK_CUDA(cudaMalloc(&d_acc, es * 3 * n));
        CHECK_CUDA(cudaMalloc(&d_phi, es * n));
        /* float64 inputs are cast to the kernel dtype inside nb_pack */
        CHECK_NB((sym ? nb_direct_sym : nb_direct)(prec, NB_KERNEL_PLUMMER, NB_MODE_BOTH, d_pos, NB_DT_F64, d_mass, NB_DT_F64, 0.0,
                                                   d_h, NB_DT_F64, 0.0, n, G, prec == NB_PREC_F64 ? 1e-30 : 1e-15, packed, ws,
                                                   ws_bytes, grid, d_acc, d_phi, st));
        CHECK_CUDA(cudaStreamSynchronize(st));
        void* acc = malloc(es * 3 * n); void* phi = malloc(es * n);
        CHECK_CUDA(cudaMemcpy(acc, d_acc, es * 3 * n, cudaMemcpyDeviceToHost));
        CHECK_CUDA(cudaMemcpy(phi, d_phi, es * n, cudaMemcpyDeviceToHost));
        double num = 0, den = 0, pmax = 0;
        for (long i = 0; i < n; ++i) {
            for (int d = 0; d < 3; ++d) {
                const double a = prec == NB_PREC_F64 ? ((double*)acc)[3 * i + d] : (double)((float*)acc)[3 * i + d];
                num += (a - aref[3 * i + d]) * (a - aref[3 * i + d]); den += aref[3 * i + d] * aref[3 * i + d];
            }
            const double ph = prec == NB_PREC_F64 ? ((double*)phi)[i] : (double)((float*)phi)[i];
            const double e = fabs(ph - pref[i]) / fabs(pref[i]);
            if (e > pmax) pmax = e;
        }
        const double rms = sqrt(num / den), tol = prec == NB_PREC_F64 ? 1e-12 : 1e-5;
        printf("prec %d %s: grid %d, acc rel-RMS %.3e, phi max rel %.3e -> %s\n", prec, sym ? "symmetric" : "one-sided", grid, rms, pmax, (rms < tol && pmax < 10 * tol) ? "ok" : "FAIL");
        if (!(rms < tol && pmax < 10 * tol)) fail = 1;
        /* invalid arguments are reported, not crashed on */
        if (nb_direct(prec, 9, NB_MODE_BOTH, d_pos, NB_DT_F64, d_mass, NB_DT_F64, 0.0, d_h, NB_DT_F64, 0.0, n, G, 1e-15, packed, ws, ws_bytes, grid, d_acc, d_phi, st) != NB_E_INVALID_ARG) fail = 1;
        if (nb_sweep(prec, NB_KERNEL_PLUMMER, NB_MODE_BOTH, packed, n, 0, 1, 0, 0, 0, n, 1e-15, ws, 16, 0, grid, st) != NB_E_WORKSPACE) fail = 1;
        cudaFree(packed); cudaFree(ws); cudaFree(d_acc); cudaFree(d_phi); free(acc); free(phi);
    }
    printf(fail ? "CAPI SMOKE FAILED\n" : "CAPI SMOKE OK\n");
    return fail;
}
