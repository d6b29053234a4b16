// cc2017.cu -- Clifford & Clifford (2018) "Algorithm A" exact boson sampler, batched over samples (complex128, sm_100a).
//
// Replaces xq.Clifford2017.set_unitary / .set_input_state / .sample / .samples (reference call sites
// perceval/backends/_clifford2017.py:39-57; algorithm: docs/source/backends.rst:108-115).
//
//   A = columns of U for the input photons, randomly permuted;  r_1 ~ |A[i,1]|^2;
//   for k = 2..n:  w_i = |Per(A[(r_1..r_{k-1}, i), 1..k])|^2  for every row i, by Laplace expansion along the new row:
//                  w_i = |sum_l A[i,l] * Per_l|^2,  Per_l = permanent of B = A[(r_1..r_{k-1}), 1..k] without column l;
//                  r_k ~ w.   Output = occupation numbers of the multiset {r_1..r_n}.
//
// The k leave-one-column-out permanents come from ONE Glynn Gray-code sweep over the (k-1) x k matrix B: for each
// sign vector the k column sums v_c are updated by one +-2*B[row] step and the k products prod_{c != l} v_c are
// formed from prefix/suffix products.
//
// Mapping: one warp per sample.  The warp is 8 groups of 4 lanes; the Gray range is cut into 8 aligned chunks (one
// per group) and inside a group the k columns are dealt to the 4 lanes in blocks of H = ceil(n/4) columns, so the
// column sums, suffix products and accumulators (3*H complex) all live in registers; the 4 lanes exchange their
// block products with two xor-shuffles per code.  Group partial sums are folded with shuffles.  RNG: Philox4x32-10
// keyed by (seed, sample index) -- identical draws to oracle/fock_oracle.c, independent of batch split / GPU count.
#include "common.cuh"

#define CC_WARPS 4
#define CC_BLOCK (CC_WARPS * 32)

__device__ __forceinline__ double2 shfl_xor_c(double2 v, int mask) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}

__global__ void cc_gather_columns_kernel(int m, int n, const double2 *__restrict__ U, const int *__restrict__ cols,
                                         double2 *__restrict__ At) {
    // At[c][i] = U[i][cols[c]]   (n x m, so that rows i are contiguous for a fixed input photon c)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * m) return;
    const int c = idx / m, i = idx - c * m;
    At[idx] = U[(size_t)i * m + cols[c]];
}

struct CcArgs {
    int m, n;
    const double2 *At;  // n x m
    uint64_t count, seed, offset;
    uint8_t *out;       // count x m
};

#ifndef CC_MINB
#define CC_MINB 1   // tuning knob (tools/build_variant.py): minimum CTAs per SM the register allocation must allow
#endif
template <int H>
__global__ void __launch_bounds__(CC_BLOCK, CC_MINB) cc2017_kernel(const CcArgs a) {
    constexpr int NP = 4 * H;  // padded column count
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.m, n = a.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 2, q = lane & 3;
    // per-warp shared layout
    const size_t bytes_B = (size_t)(n > 1 ? n - 1 : 1) * NP * sizeof(double2);
    const size_t bytes_sp = NP * sizeof(double2);
    const size_t bytes_w = (((size_t)m * sizeof(double)) + 15) & ~(size_t)15;
    const size_t bytes_i = 2 * FOCK_NMAX * sizeof(int);
    const size_t per_warp = bytes_B + bytes_sp + bytes_w + bytes_i;
    unsigned char *base = smem_raw + warp * per_warp;
    double2 *sB = (double2 *)base;
    double2 *sSp = (double2 *)(base + bytes_B);
    double *sW = (double *)(base + bytes_B + bytes_sp);
    int *sPerm = (int *)(base + bytes_B + bytes_sp + bytes_w);
    int *sRows = sPerm + FOCK_NMAX;

    const uint64_t warps_total = (uint64_t)gridDim.x * CC_WARPS;
    for (uint64_t smp = (uint64_t)blockIdx.x * CC_WARPS + warp; smp < a.count; smp += warps_total) {
        const uint64_t idx = a.offset + smp;
        uint8_t *out = a.out + smp * (uint64_t)m;
        for (int i = lane; i < m; i += 32) out[i] = 0;
        // random column permutation (Fisher-Yates, draws 0..n-2)
        if (lane == 0) {
            for (int i = 0; i < n; ++i) sPerm[i] = i;
            for (int i = 0; i < n - 1; ++i) {
                int j = i + (int)(philox_uniform(a.seed, idx, (uint32_t)i) * (double)(n - i));
                if (j > n - 1) j = n - 1;
                const int t = sPerm[i]; sPerm[i] = sPerm[j]; sPerm[j] = t;
            }
        }
        __syncwarp();

        for (int k = 1; k <= n; ++k) {
            const int r = k - 1;
            if (r == 0) {
                for (int c = lane; c < NP; c += 32) sSp[c] = make_double2(c == 0 ? 1.0 : 0.0, 0.0);
            } else {
                // stage B (r x NP): real columns c < k, pad columns are the multiplicative identity (row 0 = 1, rest 0)
                for (int e = lane; e < r * NP; e += 32) {
                    const int i = e / NP, c = e - i * NP;
                    double2 val;
                    if (c < k) val = a.At[(size_t)sPerm[c] * m + sRows[i]];
                    else val = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
                    sB[e] = val;
                }
                __syncwarp();
                const uint64_t C = (uint64_t)1 << (r - 1);
                const uint64_t per = C >= 8 ? (C >> 3) : 1;
                const uint64_t g0 = (uint64_t)grp * per;
                const bool active = g0 < C;
                double2 v[H], acc[H], suf[H];
#pragma unroll
                for (int h = 0; h < H; ++h) acc[h] = make_double2(0.0, 0.0);
                if (active) {
                    const uint64_t gray0 = g0 ^ (g0 >> 1);
#pragma unroll
                    for (int h = 0; h < H; ++h) v[h] = sB[q * H + h];
                    for (int i = 1; i < r; ++i) {
                        const double d = ((gray0 >> (i - 1)) & 1) ? -1.0 : 1.0;
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            const double2 e = sB[i * NP + q * H + h];
                            v[h].x = fma(d, e.x, v[h].x);
                            v[h].y = fma(d, e.y, v[h].y);
                        }
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < H; ++h) v[h] = make_double2(1.0, 0.0);
                }
                double sgn = (active && (__popcll(g0 ^ (g0 >> 1)) & 1)) ? -1.0 : 1.0;
                if (!active) sgn = 0.0;
                // all 32 lanes run the same trip count (shuffles inside); inactive groups contribute 0
#pragma unroll 1
                for (uint64_t t = 0; t < per; ++t) {
                    const uint64_t g = g0 + t;
                    // suffix products inside the lane's block
                    suf[H - 1] = v[H - 1];
#pragma unroll
                    for (int h = H - 2; h >= 0; --h) suf[h] = cmul(v[h], suf[h + 1]);
                    // product of the other three lanes' blocks
                    const double2 P = suf[0];
                    const double2 qd = shfl_xor_c(P, 1);
                    const double2 pp = cmul(P, qd);
                    const double2 rr = shfl_xor_c(pp, 2);
                    const double2 other = cmul(qd, rr);
                    double2 pre = make_double2(sgn * other.x, sgn * other.y);
#pragma unroll
                    for (int h = 0; h < H - 1; ++h) {
                        acc[h] = cfma(pre, suf[h + 1], acc[h]);
                        pre = cmul(pre, v[h]);
                    }
                    acc[H - 1].x += pre.x;
                    acc[H - 1].y += pre.y;
                    sgn = -sgn;
                    const uint64_t gn = g + 1;
                    if (t + 1 < per) {
                        const int b = __ffsll((long long)gn) - 1;
                        const uint64_t ngray = gn ^ (gn >> 1);
                        const double d = ((ngray >> b) & 1) ? -2.0 : 2.0;
                        const double2 *row = sB + (b + 1) * NP + q * H;
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            const double2 e = row[h];
                            v[h].x = fma(d, e.x, v[h].x);
                            v[h].y = fma(d, e.y, v[h].y);
                        }
                    }
                }
                // fold the 8 groups (lanes with equal q)
                const double scale = ldexp(1.0, 1 - r);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    double2 s = acc[h];
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        const double2 t2 = shfl_xor_c(s, o);
                        s.x += t2.x; s.y += t2.y;
                    }
                    if (grp == 0) sSp[q * H + h] = make_double2(s.x * scale, s.y * scale);
                }
            }
            __syncwarp();
            // weights of the m candidate rows
            double part = 0.0;
            for (int i = lane; i < m; i += 32) {
                double2 s = make_double2(0.0, 0.0);
                for (int c = 0; c < k; ++c) s = cfma(a.At[(size_t)sPerm[c] * m + i], sSp[c], s);
                const double w = s.x * s.x + s.y * s.y;
                sW[i] = w;
                part += w;
            }
            __syncwarp();
            const double tot = warp_sum(part);
            const double x = philox_uniform(a.seed, idx, (uint32_t)(n - 1 + k - 1)) * tot;
            // segment scan: lane owns rows [lo, hi)
            const int seg = (m + 31) / 32;
            const int lo = min(lane * seg, m), hi = min(lo + seg, m);
            double segsum = 0.0;
            for (int i = lo; i < hi; ++i) segsum += sW[i];
            double incl = segsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t2 = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t2;
            }
            const double excl = incl - segsum;
            const unsigned ball = __ballot_sync(0xffffffffu, (x < incl) && (hi > lo));
            int pick;
            if (ball == 0) {
                pick = m - 1;
            } else {
                const int owner = __ffs(ball) - 1;
                int p = hi - 1;
                if (lane == owner) {
                    double cum = excl;
                    for (int i = lo; i < hi; ++i) {
                        cum += sW[i];
                        if (x < cum) { p = i; break; }
                    }
                }
                pick = __shfl_sync(0xffffffffu, p, owner);
            }
            if (lane == 0) {
                while (pick > 0 && sW[pick] == 0.0) --pick;
                sRows[k - 1] = pick;
            }
            __syncwarp();
        }
        if (lane == 0)
            for (int k = 0; k < n; ++k) out[sRows[k]] += 1;
        __syncwarp();
    }
}

template <int H>
static int launch_cc(fock_ctx *c, const CcArgs &a, cudaStream_t st) {
    const int NP = 4 * H;
    const size_t per_warp = (size_t)(a.n > 1 ? a.n - 1 : 1) * NP * 16 + (size_t)NP * 16 + ((((size_t)a.m * 8) + 15) & ~(size_t)15) + 2 * FOCK_NMAX * sizeof(int);
    const size_t smem = per_warp * CC_WARPS;
    FOCK_REQUIRE(smem <= 220 * 1024, FOCK_ERR_LIMIT, "cc2017_samples: m=%d, n=%d needs %zu B of shared memory per CTA", a.m, a.n, smem);
    FOCK_CUDA(cudaFuncSetAttribute(cc2017_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FOCK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cc2017_kernel<H>, CC_BLOCK, smem));
    if (occ < 1) occ = 1;
    uint64_t blocks = (a.count + CC_WARPS - 1) / CC_WARPS;
    const uint64_t cap = (uint64_t)c->sm_count * occ;
    if (blocks > cap) blocks = cap;
    cc2017_kernel<H><<<(unsigned)blocks, CC_BLOCK, smem, st>>>(a);
    c->launches++;
    return fock_check_cuda(cudaGetLastError(), "cc2017_kernel");
}

extern "C" int cc2017_samples(fock_ctx *c, int m, int n, const double *d_U, const uint8_t *in_state, uint64_t count, uint64_t seed,
                              uint64_t offset, uint8_t *d_out, void *stream) {
    FOCK_REQUIRE(c && d_U && in_state && (d_out || count == 0), FOCK_ERR_ARG, "cc2017_samples: bad argument");
    FOCK_REQUIRE(m >= 1 && m <= 16384, FOCK_ERR_LIMIT, "cc2017_samples: m=%d outside [1,16384]", m);
    FOCK_REQUIRE(n >= 0 && n <= FOCK_NMAX, FOCK_ERR_LIMIT, "cc2017_samples: n=%d outside [0,%d]", n, FOCK_NMAX);
    int cols[FOCK_NMAX], nn = 0;
    for (int i = 0; i < m; ++i)
        for (int t = 0; t < in_state[i]; ++t) {
            FOCK_REQUIRE(nn < n, FOCK_ERR_ARG, "cc2017_samples: input state holds more than n=%d photons", n);
            cols[nn++] = i;
        }
    FOCK_REQUIRE(nn == n, FOCK_ERR_ARG, "cc2017_samples: input state holds %d photons, expected %d", nn, n);
    if (count == 0) return FOCK_OK;
    ScopedDevice sd(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        FOCK_CUDA(cudaMemsetAsync(d_out, 0, count * (uint64_t)m, st));
        return FOCK_OK;
    }
    const size_t bytes = 16 * (size_t)n * m + 256;
    StreamScratch scratch;   // gathered input columns, stream-ordered
    if (int rc = scratch.alloc(bytes, st)) return rc;
    double2 *At = (double2 *)scratch.ptr;
    int *d_cols = (int *)((char *)scratch.ptr + 16 * (size_t)n * m);
    FOCK_CUDA(cudaMemcpyAsync(d_cols, cols, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    cc_gather_columns_kernel<<<(n * m + 255) / 256, 256, 0, st>>>(m, n, (const double2 *)d_U, d_cols, At);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    CcArgs a;
    a.m = m; a.n = n; a.At = At; a.count = count; a.seed = seed; a.offset = offset; a.out = d_out;
    const int H = (n + 3) / 4;
    switch (H) {
        case 1: return launch_cc<1>(c, a, st);
        case 2: return launch_cc<2>(c, a, st);
        case 3: return launch_cc<3>(c, a, st);
        case 4: return launch_cc<4>(c, a, st);
        case 5: return launch_cc<5>(c, a, st);
        case 6: return launch_cc<6>(c, a, st);
        case 7: return launch_cc<7>(c, a, st);
        default: return launch_cc<8>(c, a, st);
    }
}

extern "C" int cc2017_samples_host(fock_ctx *c, int m, int n, const double *h_U, const uint8_t *in_state, uint64_t count, uint64_t seed,
                                   uint64_t offset, uint8_t *h_out) {
    FOCK_REQUIRE(c && h_U && in_state && (h_out || count == 0), FOCK_ERR_ARG, "cc2017_samples_host: bad argument");
    if (count == 0) return FOCK_OK;
    ScopedDevice sd(c->device);
    double *dU = nullptr;
    uint8_t *dout = nullptr;
    int rc = FOCK_OK;
    if ((rc = fock_check_cuda(cudaMalloc(&dU, 16 * (size_t)m * m), "cudaMalloc")) == 0 &&
        (rc = fock_check_cuda(cudaMalloc(&dout, count * (size_t)m), "cudaMalloc")) == 0) {
        rc = fock_check_cuda(cudaMemcpy(dU, h_U, 16 * (size_t)m * m, cudaMemcpyHostToDevice), "H2D");
        if (!rc) rc = cc2017_samples(c, m, n, dU, in_state, count, seed, offset, dout, nullptr);
        if (!rc) rc = fock_check_cuda(cudaMemcpy(h_out, dout, count * (size_t)m, cudaMemcpyDeviceToHost), "D2H");
    }
    cudaFree(dU);
    cudaFree(dout);
    return rc;
}
