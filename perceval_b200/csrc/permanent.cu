// permanent.cu -- batched Glynn Gray-code permanents (complex128, sm_100a) and the Naive backend's amplitude path.
//
// Replaces xq.permanent_cx(M) (reference perceval/backends/_naive.py:70-71) and the Python triple loop that builds
// the sub-matrix (_naive.py:51-68).
//
//     perm(M) = 2^{1-n} * sum_{g=0}^{2^{n-1}-1} (-1)^{popc(gray(g))} * prod_j v_j(g),
//     v_j(g)  = sum_i delta_i(g) M[i,j],  delta_0 = +1,  delta_{b+1} = -1 iff bit b of gray(g) is set.
//
// Work split: the Gray range is cut into aligned chunks of L = 2^c codes; one thread walks one chunk, re-seeding its
// n column sums from the chunk's first code and then applying one +-2*M[row] rank-1 update per step.  Because chunk
// starts are multiples of L, ctz(g+1) -- the row that flips -- is identical for all lanes of a warp, so the row is
// read from shared memory as a broadcast.  The n column sums live in registers (kernel templated on n, fully
// unrolled); the per-thread partial sums are folded with warp shuffles, then per block, then by a second tiny
// kernel in a fixed order (deterministic result).  The kernel is bound by the FP64 pipe: per step 2n DFMA for the
// update and 4(n-1) DMUL/DFMA for the product chain.
#include "common.cuh"

#define GLYNN_BLOCK 128

#include <stdlib.h>
static int glynn_env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int N, bool SMEM>
__device__ __forceinline__ double2 glynn_chunk(const double2 *__restrict__ M, uint64_t g0, uint64_t g1) {
    double2 v[N];
    const uint64_t gray0 = g0 ^ (g0 >> 1);
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = M[j];
#pragma unroll 1
    for (int i = 1; i < N; ++i) {
        const double d = ((gray0 >> (i - 1)) & 1) ? -1.0 : 1.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double2 e = M[i * N + j];
            v[j].x = fma(d, e.x, v[j].x);
            v[j].y = fma(d, e.y, v[j].y);
        }
    }
    double sgn = (__popcll(gray0) & 1) ? -1.0 : 1.0;
    double2 total = make_double2(0.0, 0.0), inner = make_double2(0.0, 0.0);
#pragma unroll 1
    for (uint64_t g = g0; g < g1; ++g) {
        // product of the n column sums: four independent chains for ILP, then combined
        double2 p0 = v[0], p1, p2, p3;
        if (N >= 4) {
            p1 = v[1]; p2 = v[2]; p3 = v[3];
#pragma unroll
            for (int j = 4; j + 3 < N; j += 4) {
                p0 = cmul(p0, v[j]);
                p1 = cmul(p1, v[j + 1]);
                p2 = cmul(p2, v[j + 2]);
                p3 = cmul(p3, v[j + 3]);
            }
#pragma unroll
            for (int j = N - (N % 4); j < N; ++j) p0 = cmul(p0, v[j]);
            p0 = cmul(cmul(p0, p1), cmul(p2, p3));
        } else {
#pragma unroll
            for (int j = 1; j < N; ++j) p0 = cmul(p0, v[j]);
        }
        inner.x = fma(sgn, p0.x, inner.x);
        inner.y = fma(sgn, p0.y, inner.y);
        sgn = -sgn;
        const uint64_t gn = g + 1;
        if ((gn & 63) == 0) {  // two-level accumulation keeps the round-off of long chunks small
            total.x += inner.x; total.y += inner.y;
            inner = make_double2(0.0, 0.0);
        }
        if (gn < g1) {
            const int b = __ffsll((long long)gn) - 1;
            const uint64_t ngray = gn ^ (gn >> 1);
            const double d = ((ngray >> b) & 1) ? -2.0 : 2.0;
            const double2 *row = M + (b + 1) * N;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 e = row[j];
                v[j].x = fma(d, e.x, v[j].x);
                v[j].y = fma(d, e.y, v[j].y);
            }
        }
    }
    total.x += inner.x; total.y += inner.y;
    return total;
}

// one matrix per blockIdx.y (looped), Gray chunks over blockIdx.x * GLYNN_BLOCK threads
#ifndef GLYNN_MINB
#define GLYNN_MINB 1   // tuning knob (tools/build_variant.py): minimum CTAs per SM the register allocation must allow
#endif
template <int N>
__global__ void __launch_bounds__(GLYNN_BLOCK, GLYNN_MINB) glynn_big_kernel(const double2 *__restrict__ mats, uint64_t B, double2 *__restrict__ partials,
                                                                uint64_t gbegin, uint64_t gend, int chunk_log2) {
    __shared__ double2 sM[N * N];
    __shared__ double2 s_red[GLYNN_BLOCK / 32];
    for (uint64_t b = blockIdx.y; b < B; b += gridDim.y) {
        __syncthreads();
        for (int i = threadIdx.x; i < N * N; i += GLYNN_BLOCK) sM[i] = mats[b * (uint64_t)(N * N) + i];
        __syncthreads();
        const uint64_t chunk = (uint64_t)blockIdx.x * GLYNN_BLOCK + threadIdx.x;
        // chunks are aligned to multiples of L in absolute Gray index
        const uint64_t first_aligned = (gbegin >> chunk_log2) << chunk_log2;
        uint64_t g0 = first_aligned + (chunk << chunk_log2), g1 = g0 + ((uint64_t)1 << chunk_log2);
        if (g0 < gbegin) g0 = gbegin;
        if (g1 > gend) g1 = gend;
        double2 t = make_double2(0.0, 0.0);
        if (g0 < g1) t = glynn_chunk<N, true>(sM, g0, g1);
        t.x = warp_sum(t.x);
        t.y = warp_sum(t.y);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            double2 s = s_red[0];
#pragma unroll
            for (int w = 1; w < GLYNN_BLOCK / 32; ++w) { s.x += s_red[w].x; s.y += s_red[w].y; }
            partials[b * (uint64_t)gridDim.x + blockIdx.x] = s;
        }
    }
}

__global__ void glynn_reduce_kernel(const double2 *__restrict__ partials, uint64_t B, uint32_t per_mat, double scale,
                                    double2 *__restrict__ out) {
    const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double2 s = make_double2(0.0, 0.0);
    for (uint32_t i = 0; i < per_mat; ++i) {
        const double2 p = partials[b * per_mat + i];
        s.x += p.x; s.y += p.y;
    }
    out[b] = make_double2(s.x * scale, s.y * scale);
}

// many small permanents: one thread per matrix, whole Gray range, matrix read through L1
template <int N>
__global__ void __launch_bounds__(GLYNN_BLOCK) glynn_small_kernel(const double2 *__restrict__ mats, uint64_t B, double2 *__restrict__ out,
                                                                  uint64_t gbegin, uint64_t gend, double scale) {
    for (uint64_t b = blockIdx.x * (uint64_t)GLYNN_BLOCK + threadIdx.x; b < B; b += (uint64_t)gridDim.x * GLYNN_BLOCK) {
        const double2 t = glynn_chunk<N, false>(mats + b * (uint64_t)(N * N), gbegin, gend);
        out[b] = make_double2(t.x * scale, t.y * scale);
    }
}

__global__ void perm_trivial_kernel(int n, const double2 *__restrict__ mats, uint64_t B, double2 *__restrict__ out) {
    const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    out[b] = (n == 0) ? make_double2(1.0, 0.0) : mats[b];
}

// ---------------------------------------------------------------- dispatch
template <int N>
static int launch_glynn(fock_ctx *c, const double2 *mats, uint64_t B, double2 *out, uint64_t g0, uint64_t g1, cudaStream_t st) {
    const uint64_t G = g1 - g0;
    const double scale = ldexp(1.0, 1 - N);
    // >= ~24 waves of 3 CTAs / SM: a 4-5 wave grid loses ~12 % to the partial last wave (ncu: 4.61 waves at n=30, B=8)
    const uint64_t target_threads = (uint64_t)c->sm_count * 3 * GLYNN_BLOCK * (uint64_t)glynn_env_int("FOCK_GLYNN_WAVES", 24);
    if (G * B <= 4096 * B && (G <= 4096) && B >= 1024) {
        // many small permanents
        uint64_t g = (B + GLYNN_BLOCK - 1) / GLYNN_BLOCK;
        if (g > (uint64_t)c->sm_count * 32) g = (uint64_t)c->sm_count * 32;
        glynn_small_kernel<N><<<(unsigned)g, GLYNN_BLOCK, 0, st>>>(mats, B, out, g0, g1, scale);
        c->launches++;
        return fock_check_cuda(cudaGetLastError(), "glynn_small_kernel");
    }
    // chunk length: at least 256 codes (amortises the n^2 re-seed), at most what keeps ~target_threads busy
    int chunk_log2 = 8;
    while (chunk_log2 < 20 && (G >> (chunk_log2 + 1)) * B >= target_threads) ++chunk_log2;
    while (chunk_log2 > 0 && ((uint64_t)1 << chunk_log2) > G) --chunk_log2;
    const uint64_t first_aligned = (g0 >> chunk_log2) << chunk_log2;
    const uint64_t nchunks = ((g1 - first_aligned) + (((uint64_t)1 << chunk_log2) - 1)) >> chunk_log2;
    const uint64_t blocks_x = (nchunks + GLYNN_BLOCK - 1) / GLYNN_BLOCK;
    FOCK_REQUIRE(blocks_x < (1u << 31), FOCK_ERR_LIMIT, "glynn: too many chunks");
    const unsigned gy = (unsigned)(B < 32768 ? B : 32768);
    StreamScratch scratch;   // per-call partial sums, stream-ordered
    if (int rc = scratch.alloc((size_t)B * blocks_x * sizeof(double2), st)) return rc;
    double2 *partials = (double2 *)scratch.ptr;
    dim3 grid((unsigned)blocks_x, gy);
    glynn_big_kernel<N><<<grid, GLYNN_BLOCK, 0, st>>>(mats, B, partials, g0, g1, chunk_log2);
    c->launches++;
    if (int rc = fock_check_cuda(cudaGetLastError(), "glynn_big_kernel")) return rc;
    glynn_reduce_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(partials, B, (uint32_t)blocks_x, scale, out);
    c->launches++;
    return fock_check_cuda(cudaGetLastError(), "glynn_reduce_kernel");
}

typedef int (*glynn_fn)(fock_ctx *, const double2 *, uint64_t, double2 *, uint64_t, uint64_t, cudaStream_t);
#define G1(n) launch_glynn<n>
static const glynn_fn g_glynn_table[FOCK_NMAX + 1] = {
    nullptr, nullptr, G1(2),  G1(3),  G1(4),  G1(5),  G1(6),  G1(7),  G1(8),  G1(9),  G1(10),
    G1(11),  G1(12),  G1(13), G1(14), G1(15), G1(16), G1(17), G1(18), G1(19), G1(20), G1(21),
    G1(22),  G1(23),  G1(24), G1(25), G1(26), G1(27), G1(28), G1(29), G1(30), G1(31), G1(32)};

extern "C" int glynn_permanent_batch(fock_ctx *c, int n, const double *d_mats, uint64_t B, double *d_out, uint64_t gray_begin,
                                     uint64_t gray_end, void *stream) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "glynn_permanent_batch: ctx is NULL");
    FOCK_REQUIRE(n >= 0 && n <= FOCK_NMAX, FOCK_ERR_LIMIT, "glynn_permanent_batch: n=%d outside [0,%d]", n, FOCK_NMAX);
    FOCK_REQUIRE(d_out && (d_mats || n == 0), FOCK_ERR_ARG, "glynn_permanent_batch: NULL device pointer");
    if (B == 0) return FOCK_OK;
    ScopedDevice sd(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 1) {
        perm_trivial_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(n, (const double2 *)d_mats, B, (double2 *)d_out);
        c->launches++;
        FOCK_CUDA(cudaGetLastError());
        return FOCK_OK;
    }
    const uint64_t G = (uint64_t)1 << (n - 1);
    if (gray_begin == 0 && gray_end == 0) gray_end = G;
    FOCK_REQUIRE(gray_begin < gray_end && gray_end <= G, FOCK_ERR_ARG, "glynn_permanent_batch: bad Gray range [%llu,%llu) of %llu",
                 (unsigned long long)gray_begin, (unsigned long long)gray_end, (unsigned long long)G);
    return g_glynn_table[n](c, (const double2 *)d_mats, B, (double2 *)d_out, gray_begin, gray_end, st);
}

extern "C" int glynn_permanent_batch_host(fock_ctx *c, int n, const double *h_mats, uint64_t B, double *h_out) {
    FOCK_REQUIRE(c && h_out && (h_mats || n == 0), FOCK_ERR_ARG, "glynn_permanent_batch_host: bad argument");
    if (B == 0) return FOCK_OK;
    ScopedDevice sd(c->device);
    double *dm = nullptr, *dout = nullptr;
    int rc = FOCK_OK;
    size_t mb = 16 * (size_t)B * (size_t)(n > 0 ? n * n : 1);
    if ((rc = fock_check_cuda(cudaMalloc(&dm, mb), "cudaMalloc")) == 0 && (rc = fock_check_cuda(cudaMalloc(&dout, 16 * B), "cudaMalloc")) == 0) {
        if (n > 0) rc = fock_check_cuda(cudaMemcpy(dm, h_mats, mb, cudaMemcpyHostToDevice), "H2D");
        if (!rc) rc = glynn_permanent_batch(c, n, dm, B, dout, 0, 0, nullptr);
        if (!rc) rc = fock_check_cuda(cudaMemcpy(h_out, dout, 16 * B, cudaMemcpyDeviceToHost), "D2H");
    }
    cudaFree(dm);
    cudaFree(dout);
    return rc;
}

// ---------------------------------------------------------------- Naive backend: sub-matrix gather + permanent + normalisation
// reference _naive.py:51-68: M[r,c] = U[out_mode(r), in_mode(c)], modes ascending, repeated per occupancy.
// One warp per output state; in_cols (n ints) lists the input photons' modes.
__global__ void __launch_bounds__(128) naive_submatrix_kernel(int m, int n, const double2 *__restrict__ U, const int *__restrict__ in_cols,
                                                              const uint64_t *__restrict__ bt, const uint64_t *__restrict__ ranks,
                                                              const uint8_t *__restrict__ states, uint64_t B, double2 *__restrict__ mats,
                                                              double *__restrict__ out_fact) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    __shared__ uint8_t s_rows[4][FOCK_NMAX];
    uint8_t *rows = s_rows[threadIdx.x >> 5];
    for (uint64_t b = warp; b < B; b += nwarps) {
        __syncwarp();
        if (lane == 0) {
            double fact = 1.0;
            int r = 0;
            if (states) {
                const uint8_t *st = states + b * (uint64_t)m;
                for (int j = 0; j < m; ++j) {
                    const int s = st[j];
                    for (int t = 0; t < s && r < n; ++t) { rows[r++] = (uint8_t)j; fact *= (double)(t + 1); }
                }
            } else {
                uint64_t rem = ranks[b];
                int Tprev = n;
                for (int j = 0; j < m - 1 && Tprev > 0; ++j) {
                    const uint64_t *row = bt + (m - 1 - j) * FOCK_TMAX;
                    int T = Tprev;
                    while (row[T] > rem) --T;
                    rem -= row[T];
                    for (int t = 0; t < Tprev - T; ++t) { rows[r++] = (uint8_t)j; fact *= (double)(t + 1); }
                    Tprev = T;
                }
                for (int t = 0; t < Tprev; ++t) { rows[r++] = (uint8_t)(m - 1); fact *= (double)(t + 1); }
            }
            out_fact[b] = fact;
        }
        __syncwarp();
        double2 *M = mats + b * (uint64_t)(n * n);
        for (int e = lane; e < n * n; e += 32) {
            const int r = e / n, cidx = e - r * n;
            M[e] = U[(size_t)rows[r] * m + in_cols[cidx]];
        }
    }
}

__global__ void naive_scale_kernel(int n, uint64_t B, const double2 *__restrict__ perms, const double *__restrict__ out_fact,
                                   double in_fact, double2 *__restrict__ amps) {
    const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double2 p = perms[b];
    if (n >= 2) {  // _naive.py:49: n == 1 returns M[0,0] un-normalised (the factor is 1 anyway)
        const double f = 1.0 / sqrt(in_fact * out_fact[b]);
        p.x *= f; p.y *= f;
    }
    amps[b] = p;
}

static int naive_impl(fock_ctx *c, int m, int n, const double *d_U, const uint8_t *in_state, const uint64_t *d_ranks,
                      const uint8_t *d_states, uint64_t B, double *d_amps, void *stream) {
    FOCK_REQUIRE(c && d_U && in_state && d_amps && (d_ranks || d_states), FOCK_ERR_ARG, "naive_amplitudes: bad argument");
    FOCK_REQUIRE(m >= 1 && m <= 255, FOCK_ERR_LIMIT, "naive_amplitudes: m=%d outside [1,255]", m);
    FOCK_REQUIRE(d_states || m <= FOCK_QMAX, FOCK_ERR_LIMIT, "naive_amplitudes: rank input needs m <= %d", FOCK_QMAX);
    FOCK_REQUIRE(n >= 0 && n <= FOCK_NMAX, FOCK_ERR_LIMIT, "naive_amplitudes: n=%d outside [0,%d]", n, FOCK_NMAX);
    int cols[FOCK_NMAX], nn = 0;
    double in_fact = 1.0;
    for (int i = 0; i < m; ++i)
        for (int t = 0; t < in_state[i]; ++t) {
            FOCK_REQUIRE(nn < n, FOCK_ERR_ARG, "naive_amplitudes: input state holds more than n=%d photons", n);
            cols[nn++] = i;
            in_fact *= (double)(t + 1);
        }
    FOCK_REQUIRE(nn == n, FOCK_ERR_ARG, "naive_amplitudes: input state holds %d photons, expected %d", nn, n);
    if (B == 0) return FOCK_OK;
    ScopedDevice sd(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int nsq = n > 0 ? n * n : 1;
    // scratch: mats (B*n*n complex) | perms (B complex) | out_fact (B double) | in_cols (32 int)
    const size_t bytes = 16 * (size_t)B * nsq + 16 * B + 8 * B + 256;
    StreamScratch scratch;
    if (int rc = scratch.alloc(bytes, st)) return rc;
    char *base = (char *)scratch.ptr;
    double2 *mats = (double2 *)base;
    double2 *perms = (double2 *)(base + 16 * (size_t)B * nsq);
    double *ofact = (double *)(base + 16 * (size_t)B * nsq + 16 * B);
    int *d_cols = (int *)(base + 16 * (size_t)B * nsq + 16 * B + 8 * B);
    if (n > 0) FOCK_CUDA(cudaMemcpyAsync(d_cols, cols, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    uint64_t g = (B + 3) / 4;
    if (g > (uint64_t)c->sm_count * 16) g = (uint64_t)c->sm_count * 16;
    naive_submatrix_kernel<<<(unsigned)g, 128, 0, st>>>(m, n, (const double2 *)d_U, d_cols, c->d_bt, d_ranks, d_states, B, mats, ofact);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    if (int rc = glynn_permanent_batch(c, n, (const double *)mats, B, (double *)perms, 0, 0, stream)) return rc;
    naive_scale_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(n, B, perms, ofact, in_fact, (double2 *)d_amps);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    return FOCK_OK;
}

extern "C" int naive_amplitudes(fock_ctx *c, int m, int n, const double *d_U, const uint8_t *in_state, const uint64_t *d_out_ranks,
                                uint64_t B, double *d_amps, void *stream) {
    return naive_impl(c, m, n, d_U, in_state, d_out_ranks, nullptr, B, d_amps, stream);
}

extern "C" int naive_amplitudes_states(fock_ctx *c, int m, int n, const double *d_U, const uint8_t *in_state,
                                       const uint8_t *d_out_states, uint64_t B, double *d_amps, void *stream) {
    return naive_impl(c, m, n, d_U, in_state, nullptr, d_out_states, B, d_amps, stream);
}
