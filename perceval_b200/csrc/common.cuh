// common.cuh -- shared helpers for libfock_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fock_b200.h"

#define FOCK_QMAX 64   // max modes for rank / SLOS kernels
#define FOCK_TMAX 33   // photons 0..32
#define FOCK_NMAX 32

struct fock_ctx {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    size_t total_mem;
    uint64_t *d_bt;   // [FOCK_QMAX][FOCK_TMAX]  Bt[q][T] = C(T-1+q, q), 0 for T==0, saturating
    uint64_t *d_dt;   // [FOCK_QMAX][FOCK_TMAX]  Dt[q][T] = Bt[q][T]-Bt[q][T-1]
    int *d_status;    // device-side error flag
    double *d_scratch; // small scratch (peaks.cu sinks)
    double *d_vacuum;  // constant complex 1 + 0i: SLOS layer 0 (the vacuum coefficient), written once at creation
    uint64_t launches;
    void *mu_state;    // owned by slos_mu.cu (cached tail occupation tables)
    cudaStream_t side[2];           // high-priority side streams for the small launches of a layer (slos.cu: SideLaunch)
    cudaEvent_t ev_begin, ev_end;   // optional: recorded around every probability-layer launch (fock_profile_events)
};

void fock_set_error(const char *fmt, ...);
int fock_check_cuda(cudaError_t e, const char *what);
const uint64_t *fock_host_bt();  // host copy of Bt
const uint64_t *fock_host_dt();

#define FOCK_CUDA(x)                                              \
    do {                                                          \
        int _rc = fock_check_cuda((x), #x);                       \
        if (_rc) return _rc;                                      \
    } while (0)

#define FOCK_REQUIRE(cond, code, ...)                             \
    do {                                                          \
        if (!(cond)) {                                            \
            fock_set_error(__VA_ARGS__);                          \
            return (code);                                        \
        }                                                         \
    } while (0)

struct ScopedDevice {
    int prev;
    explicit ScopedDevice(int dev) : prev(-1) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~ScopedDevice() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Stream-ordered scratch memory: allocated and released on the caller's stream (cudaMallocAsync / cudaFreeAsync from the
// device's default pool, whose release threshold fock_create raises so that the memory is re-used, not returned).  Calls
// on different streams or host threads therefore never share a scratch buffer, and no call synchronises the device.
struct StreamScratch {
    void *ptr = nullptr;
    cudaStream_t st = nullptr;
    int alloc(size_t bytes, cudaStream_t s) {
        st = s;
        return fock_check_cuda(cudaMallocAsync(&ptr, bytes ? bytes : 16, s), "cudaMallocAsync");
    }
    ~StreamScratch() {
        if (ptr) cudaFreeAsync(ptr, st);
    }
};

// ---------------------------------------------------------------- complex helpers (double2 = re, im)
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {  // a*b + c
    double re = fma(a.x, b.x, c.x);
    double im = fma(a.x, b.y, c.y);
    re = fma(-a.y, b.y, re);
    im = fma(a.y, b.x, im);
    return make_double2(re, im);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// streaming 16-byte accesses
__device__ __forceinline__ double2 ld_stream(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(double2 *p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// contiguous range -> L2 in one instruction (bytes: multiple of 16, 16-byte aligned address)
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- Philox4x32-10 (same keying as oracle/fock_oracle.c)
__host__ __device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t ctr_hi, uint64_t ctr_lo, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// d-th uniform double in [0,1) of sample idx
__host__ __device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t idx, uint32_t d) {
    uint32_t w[4];
    philox4x32_10(seed, idx, (uint64_t)(d >> 1), w);
    uint32_t a = w[(d & 1) * 2], b = w[(d & 1) * 2 + 1];
    uint64_t bits = (((uint64_t)a << 32) | b) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
