// nb_ubench.cu — CUDA-core peak microbenchmarks: the roofline denominators for
// the pair sweep (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
// Each kernel keeps 8 independent dependency chains per thread in registers,
// 256 threads x (8 x SMs) CTAs, no memory traffic in the timed loop.
#include "nb_common.cuh"

namespace nb {

constexpr int UB_CHAINS = 8;

template <int WHICH>
__global__ void __launch_bounds__(256) ubench_kernel(float* out, int iters, float seed) {
    const float t = seed + threadIdx.x * 1e-7f;
    if (WHICH == 0) {   // scalar FFMA
        float a[UB_CHAINS];
        const float b = 1.0f + t * 1e-3f, c = t * 1e-3f;
#pragma unroll
        for (int k = 0; k < UB_CHAINS; ++k) a[k] = t + k;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 32; ++r)
#pragma unroll
                for (int k = 0; k < UB_CHAINS; ++k) a[k] = fmaf(a[k], b, c);
        }
        float s = 0;
#pragma unroll
        for (int k = 0; k < UB_CHAINS; ++k) s += a[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (WHICH >= 1 && WHICH <= 4) {
        f2 a[UB_CHAINS], f[UB_CHAINS], d[UB_CHAINS];
        const f2 b = pk(1.0f + t * 1e-3f, 1.0f - t * 1e-3f), c = pk(t * 1e-3f, -t * 1e-3f);
#pragma unroll
        for (int k = 0; k < UB_CHAINS; ++k) {
            a[k] = pk(t + k, t - k);
            f[k] = pk(1.0f + 1e-4f * k + t * 1e-3f, 1.0f - 1e-4f * k);
            d[k] = pk(1e-3f * k + t, -1e-3f * k);
        }
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int k = 0; k < UB_CHAINS; ++k) {
                    if (WHICH == 1) a[k] = fma2(a[k], b, c);            // two loop-invariant operands
                    if (WHICH == 2) a[k] = fma2(f[k], d[k], a[k]);      // three distinct register pairs
                    if (WHICH == 3) a[k] = add2(a[k], d[k]);
                    if (WHICH == 4) a[k] = mul2(a[k], f[k]);
                }
        }
        float s = 0;
#pragma unroll
        for (int k = 0; k < UB_CHAINS; ++k) s += lo_of(a[k]) + hi_of(a[k]);
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (WHICH == 5) {   // MUFU.RSQ
        float a[UB_CHAINS];
#pragma unroll
        for (int k = 0; k < UB_CHAINS; ++k) a[k] = 1.0f + t + k;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int k = 0; k < UB_CHAINS; ++k) a[k] = rsqrt_approx(a[k]);
        }
        float s = 0;
#pragma unroll
        for (int k = 0; k < UB_CHAINS; ++k) s += a[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (WHICH == 6 || WHICH == 8) {
        // the Plummer force+potential instruction mix on register-resident data:
        // packed (6) : 3 FADD2 + 2 FMNMX + 3 FFMA2 + 2 MUFU + 3 FMUL2 + 3 FFMA2 + 1 FADD2 per 2 interactions
        // scalar (8) : the same maths with scalar FADD/FFMA/FMUL
        f2 xi[4], yi[4], zi[4], ax[4], ay[4], az[4], sp[4];
        float c0[4], c1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            xi[q] = pk(t + q, t - q); yi[q] = pk(2 * t + q, t - 2 * q); zi[q] = pk(t * q, -t * q);
            c0[q] = 1e-4f * (q + 1); c1[q] = 2e-4f * (q + 1);
            ax[q] = ay[q] = az[q] = sp[q] = 0ull;
        }
        float xj = 0.3f + t, yj = 0.2f - t, zj = 0.1f + t, mj = 1.0f, cj = 1.5e-4f;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (WHICH == 6) {
                        const f2 dx = sub2(splat(xj), xi[q]), dy = sub2(splat(yj), yi[q]), dz = sub2(splat(zj), zi[q]);
                        const f2 c  = pk(fmaxf(cj, c0[q]), fmaxf(cj, c1[q]));
                        const f2 d  = fma2(dx, dx, fma2(dy, dy, fma2(dz, dz, c)));
                        float d0, d1;
                        upk(d, d0, d1);
                        const f2 rinv = pk(rsqrt_approx(d0), rsqrt_approx(d1));
                        const f2 mr = mul2(splat(mj), rinv);
                        sp[q] = add2(sp[q], mr);
                        const f2 f = mul2(mr, mul2(rinv, rinv));
                        ax[q] = fma2(f, dx, ax[q]); ay[q] = fma2(f, dy, ay[q]); az[q] = fma2(f, dz, az[q]);
                    } else {
                        float x0, x1, y0, y1, z0, z1, a0, a1, b0, b1, e0, e1, s0, s1;
                        upk(xi[q], x0, x1); upk(yi[q], y0, y1); upk(zi[q], z0, z1);
                        upk(ax[q], a0, a1); upk(ay[q], b0, b1); upk(az[q], e0, e1); upk(sp[q], s0, s1);
                        {
                            const float dx = xj - x0, dy = yj - y0, dz = zj - z0;
                            const float d = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaxf(cj, c0[q]))));
                            const float ri = rsqrt_approx(d), mr = mj * ri, f = mr * (ri * ri);
                            s0 += mr; a0 = fmaf(f, dx, a0); b0 = fmaf(f, dy, b0); e0 = fmaf(f, dz, e0);
                        }
                        {
                            const float dx = xj - x1, dy = yj - y1, dz = zj - z1;
                            const float d = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaxf(cj, c1[q]))));
                            const float ri = rsqrt_approx(d), mr = mj * ri, f = mr * (ri * ri);
                            s1 += mr; a1 = fmaf(f, dx, a1); b1 = fmaf(f, dy, b1); e1 = fmaf(f, dz, e1);
                        }
                        ax[q] = pk(a0, a1); ay[q] = pk(b0, b1); az[q] = pk(e0, e1); sp[q] = pk(s0, s1);
                    }
                }
                xj += 1e-3f; yj -= 1e-3f; zj += 2e-3f; cj += 1e-7f;
            }
        }
        float s = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            s += lo_of(ax[q]) + hi_of(ax[q]) + lo_of(ay[q]) + hi_of(ay[q]) + lo_of(az[q]) + hi_of(az[q]) + lo_of(sp[q]) +
                 hi_of(sp[q]);
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (WHICH >= 9 && WHICH <= 12) {
        // co-issue probes: 12 FMA-pipe ops per group + k ops of another pipe
        //  9: 12 FFMA2 + 2 MUFU    10: 12 FFMA2 + 4 FMNMX   11: 12 FFMA + 2 MUFU   12: 12 FFMA2 + 2 MUFU + 2 FMNMX
        f2 a[12];
        float s[4];
        const f2 b = pk(1.0f + t * 1e-3f, 1.0f - t * 1e-3f), c = pk(t * 1e-3f, -t * 1e-3f);
        const float bs = 1.0f + t * 1e-3f, cs = t * 1e-3f;
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] = pk(t + k, t - k);
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = 1.0f + t + k;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    if (WHICH == 11) {
                        float lo, hi;
                        upk(a[k], lo, hi);
                        a[k] = pk(fmaf(lo, bs, cs), hi);
                    } else {
                        a[k] = fma2(a[k], b, c);
                    }
                }
                if (WHICH == 9 || WHICH == 11 || WHICH == 12) {
                    s[0] = rsqrt_approx(s[0]);
                    s[1] = rsqrt_approx(s[1]);
                }
                if (WHICH == 10) {
                    s[0] = fmaxf(-s[0], cs); s[1] = fmaxf(-s[1], bs); s[2] = fmaxf(-s[2], cs); s[3] = fmaxf(-s[3], bs);
                }
                if (WHICH == 12) {
                    s[2] = fmaxf(-s[2], cs); s[3] = fmaxf(-s[3], bs);
                }
            }
        }
        float r = s[0] + s[1] + s[2] + s[3];
#pragma unroll
        for (int k = 0; k < 12; ++k) r += lo_of(a[k]) + hi_of(a[k]);
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (WHICH >= 13 && WHICH <= 15) {
        // register-resident copy of the uniform-tile Plummer inner loop (what the sweep issues per pair of
        // interactions): 13 = force+potential, 14 = force + potential accumulated on the FP64 pipe
        // (F2F + DADD instead of FADD2), 15 = force only
        f2 xi[2], yi[2], zi[2], ax[2], ay[2], az[2], sp[2];
        double spd[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            xi[q] = pk(t + q, t - q); yi[q] = pk(2 * t + q, t - 2 * q); zi[q] = pk(t * q, -t * q);
            ax[q] = ay[q] = az[q] = sp[q] = 0ull;
        }
        const f2 c = pk(1e-4f + t * 1e-6f, 1e-4f);
        float xj = 0.3f + t, yj = 0.2f - t, zj = 0.1f + t;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const f2 dx = sub2(splat(xj), xi[q]), dy = sub2(splat(yj), yi[q]), dz = sub2(splat(zj), zi[q]);
                    const f2 d  = fma2(dx, dx, fma2(dy, dy, fma2(dz, dz, c)));
                    float d0, d1;
                    upk(d, d0, d1);
                    const float r0 = rsqrt_approx(d0), r1 = rsqrt_approx(d1);
                    const f2 rinv = pk(r0, r1);
                    if (WHICH == 13) sp[q] = add2(sp[q], rinv);
                    if (WHICH == 14) { spd[2 * q] += static_cast<double>(r0); spd[2 * q + 1] += static_cast<double>(r1); }
                    const f2 f = mul2(rinv, mul2(rinv, rinv));
                    ax[q] = fma2(f, dx, ax[q]); ay[q] = fma2(f, dy, ay[q]); az[q] = fma2(f, dz, az[q]);
                }
                xj += 1e-3f; yj -= 1e-3f; zj += 2e-3f;
            }
        }
        float s = static_cast<float>(spd[0] + spd[1] + spd[2] + spd[3]);
#pragma unroll
        for (int q = 0; q < 2; ++q)
            s += lo_of(ax[q]) + hi_of(ax[q]) + lo_of(ay[q]) + hi_of(ay[q]) + lo_of(az[q]) + hi_of(az[q]) + lo_of(sp[q]) +
                 hi_of(sp[q]);
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (WHICH == 16) {
        // accumulate trios: acc_c = f * d_c + acc_c, c = x,y,z — three FFMA2 sharing the first operand
        f2 f[4], d[4][3], acc[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[k] = pk(1e-3f * k + t, 1e-3f);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) { d[k][cc] = pk(t + k + cc, t - k); acc[k][cc] = 0ull; }
        }
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) acc[k][cc] = fma2(f[k], d[k][cc], acc[k][cc]);
        }
        float s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) s += lo_of(acc[k][cc]) + hi_of(acc[k][cc]);
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (WHICH >= 17 && WHICH <= 20) {
        // FP64 operand-count probes (12 independent chains each):
        //   17: acc = fma(f, d, acc)   three distinct 64-bit registers, f shared by a trio (the sweep's accumulates)
        //   18: acc = fma(d, d, acc)   two distinct registers (the r^2 chain)
        //   19: acc = fma(f, d, acc)   three distinct registers, nothing shared
        //   20: acc = d - acc          DADD, two distinct registers
        double f[12], d[12], acc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            f[k]   = 1.0 + 1e-9 * (t + k);
            d[k]   = 1e-3 * (t - k);
            acc[k] = 0.0;
            asm volatile("" : "+d"(f[k]), "+d"(d[k]));   // opaque: keep them in their own registers
        }
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    if (WHICH == 17) acc[k] = fma(f[k / 3], d[k], acc[k]);
                    if (WHICH == 18) acc[k] = fma(d[k], d[k], acc[k]);
                    if (WHICH == 19) acc[k] = fma(f[k], d[k], acc[k]);
                    if (WHICH == 20) acc[k] = d[k] - acc[k];
                }
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < 12; ++k) s += acc[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<float>(s);
    } else if (WHICH == 7) {   // DFMA
        double a[UB_CHAINS];
        const double b = 1.0 + t * 1e-3, c = t * 1e-3;
#pragma unroll
        for (int k = 0; k < UB_CHAINS; ++k) a[k] = t + k;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int k = 0; k < UB_CHAINS; ++k) a[k] = fma(a[k], b, c);
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < UB_CHAINS; ++k) s += a[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<float>(s);
    }
}

template <int WHICH>
static int run_ubench(int iters, int ctas, float* buf, float* ms, cudaStream_t st) {
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return (int)e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return (int)e;
    ubench_kernel<WHICH><<<ctas, 256, 0, st>>>(buf, iters / 8 + 1, 0.5f);   // warm-up
    cudaEventRecord(e0, st);
    ubench_kernel<WHICH><<<ctas, 256, 0, st>>>(buf, iters, 0.5f);
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return (int)e;
}

}  // namespace nb

using namespace nb;

int nb_ubench_impl(int which, int iters, double* lane_ops_per_s, double* ms_out) {
    if (iters <= 0 || !lane_ops_per_s) return NB_E_INVALID_ARG;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    const int ctas = sms * 8;
    float* buf = nullptr;
    e = cudaMalloc(&buf, sizeof(float) * 256 * ctas);   // benchmark-only scratch, freed below
    if (e != cudaSuccess) return (int)e;
    float ms = 0.f;
    int rc = 0;
    double ops_per_thread_iter = 0;   // lane-operations per thread per loop iteration
    switch (which) {
        case 0: rc = run_ubench<0>(iters / 4 + 1, ctas, buf, &ms, 0); ops_per_thread_iter = 32.0 * UB_CHAINS; iters = iters / 4 + 1; break;
        case 1: rc = run_ubench<1>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 2 * 8.0 * UB_CHAINS; break;
        case 2: rc = run_ubench<2>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 2 * 8.0 * UB_CHAINS; break;
        case 3: rc = run_ubench<3>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 2 * 8.0 * UB_CHAINS; break;
        case 4: rc = run_ubench<4>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 2 * 8.0 * UB_CHAINS; break;
        case 5: rc = run_ubench<5>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 8.0 * UB_CHAINS; break;
        // mixes: report INTERACTIONS per second in lane_ops_per_s (2 rounds x 4 pairs x 2 lanes = 16 / iter)
        case 6: rc = run_ubench<6>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 16.0; break;
        case 7: rc = run_ubench<7>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 8.0 * UB_CHAINS; break;
        case 8: rc = run_ubench<8>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 16.0; break;
        // co-issue probes: report FMA-pipe lane-ops only (4 rounds x 12 ops; packed count 2)
        case 9: rc = run_ubench<9>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 2 * 48.0; break;
        case 10: rc = run_ubench<10>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 2 * 48.0; break;
        case 11: rc = run_ubench<11>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 48.0; break;
        case 12: rc = run_ubench<12>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 2 * 48.0; break;
        // uniform-tile inner-loop copies: INTERACTIONS per second (4 rounds x 2 pairs x 2 lanes = 16 / iter)
        case 13: rc = run_ubench<13>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 16.0; break;
        case 14: rc = run_ubench<14>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 16.0; break;
        case 15: rc = run_ubench<15>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 16.0; break;
        case 16: rc = run_ubench<16>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 2 * 48.0; break;
        // FP64 operand-count probes: FP64 lane-ops (4 rounds x 12 chains)
        case 17: rc = run_ubench<17>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 48.0; break;
        case 18: rc = run_ubench<18>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 48.0; break;
        case 19: rc = run_ubench<19>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 48.0; break;
        case 20: rc = run_ubench<20>(iters, ctas, buf, &ms, 0); ops_per_thread_iter = 48.0; break;
        default: rc = NB_E_INVALID_ARG;
    }
    cudaFree(buf);
    if (rc) return rc;
    const double total = ops_per_thread_iter * (double)iters * 256.0 * (double)ctas;
    *lane_ops_per_s = total / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return 0;
}
