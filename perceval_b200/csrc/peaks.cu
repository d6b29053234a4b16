// peaks.cu -- roofline denominators measured on the device the bench runs on (bench.py only, not on the product path).
//   kind 0: FP64 FMA peak, TFLOP/s (2 flops per DFMA); MEASURED_PEAKS.json has no FP64 figure
//   kind 1: HBM stream copy (read + write bytes), GB/s -- same method as MEASURED_PEAKS.json:hbm_gbs
//   kind 2: L2-resident read bandwidth (32 MiB working set), GB/s
//   kind 3: HBM read-only stream, GB/s
#include "common.cuh"

__global__ void __launch_bounds__(256) fp64_fma_kernel(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) st_stream(dst + i, ld_stream(src + i));
}

__global__ void __launch_bounds__(256) read_kernel(const double2 *__restrict__ src, uint64_t n, int reps, double *out) {
    double s = 0.0;
    for (int r = 0; r < reps; ++r)
        for (uint64_t i = blockIdx.x * (uint64_t)256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
            const double2 v = __ldcg(src + i);
            s += v.x + v.y;
        }
    if (s == 123.456) out[0] = s;
}

static int time_ms(cudaEvent_t e0, cudaEvent_t e1, float *ms) {
    FOCK_CUDA(cudaEventSynchronize(e1));
    FOCK_CUDA(cudaEventElapsedTime(ms, e0, e1));
    return 0;
}

extern "C" int fock_measure_peak(fock_ctx *c, int kind, double *out_value) {
    FOCK_REQUIRE(c && out_value, FOCK_ERR_ARG, "fock_measure_peak: bad argument");
    ScopedDevice sd(c->device);
    cudaEvent_t e0, e1;
    FOCK_CUDA(cudaEventCreate(&e0));
    FOCK_CUDA(cudaEventCreate(&e1));
    float best = 1e30f, ms = 0;
    int rc = 0;
    if (kind == 0) {
        const int iters = 4096, grid = c->sm_count * 8;
        for (int rep = 0; rep < 6 && !rc; ++rep) {
            cudaEventRecord(e0);
            fp64_fma_kernel<<<grid, 256>>>(c->d_scratch, iters, 1.0);
            cudaEventRecord(e1);
            rc = time_ms(e0, e1, &ms);
            if (rep > 0 && ms < best) best = ms;
        }
        const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
        *out_value = flops / (best * 1e-3) / 1e12;
    } else if (kind == 1 || kind == 3) {
        const uint64_t n = (uint64_t)1 << 27;  // 2 GiB of double2
        double2 *a = nullptr, *b = nullptr;
        FOCK_CUDA(cudaMalloc(&a, n * 16));
        if (kind == 1) FOCK_CUDA(cudaMalloc(&b, n * 16));
        cudaMemset(a, 0, n * 16);
        for (int rep = 0; rep < 8 && !rc; ++rep) {
            cudaEventRecord(e0);
            if (kind == 1) copy_kernel<<<c->sm_count * 8, 256>>>(a, b, n);
            else read_kernel<<<c->sm_count * 8, 256>>>(a, n, 1, c->d_scratch);
            cudaEventRecord(e1);
            rc = time_ms(e0, e1, &ms);
            if (rep > 1 && ms < best) best = ms;
        }
        *out_value = (kind == 1 ? 2.0 : 1.0) * n * 16.0 / (best * 1e-3) / 1e9;
        cudaFree(a);
        cudaFree(b);
    } else if (kind == 2) {
        const uint64_t n = (uint64_t)1 << 21;  // 32 MiB
        const int reps = 64;
        double2 *a = nullptr;
        FOCK_CUDA(cudaMalloc(&a, n * 16));
        cudaMemset(a, 0, n * 16);
        for (int rep = 0; rep < 6 && !rc; ++rep) {
            cudaEventRecord(e0);
            read_kernel<<<c->sm_count * 8, 256>>>(a, n, reps, c->d_scratch);
            cudaEventRecord(e1);
            rc = time_ms(e0, e1, &ms);
            if (rep > 0 && ms < best) best = ms;
        }
        *out_value = (double)reps * n * 16.0 / (best * 1e-3) / 1e9;
        cudaFree(a);
    } else {
        fock_set_error("fock_measure_peak: unknown kind %d", kind);
        rc = FOCK_ERR_ARG;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (!rc) rc = fock_check_cuda(cudaGetLastError(), "fock_measure_peak");
    return rc;
}
