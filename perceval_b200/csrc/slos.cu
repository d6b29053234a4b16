// slos.cu -- SLOS layer propagation over combinatorially ranked Fock states (complex128, sm_100a).
//
// Replaces exqalibur's FSMap.compute_slos_layer(u, m, mk, coefs, parent_coefs) (reference call site
// perceval/backends/_slos.py:99; the reference's own Python twin of the loop is _slos.py:91-97) and
// xq.all_prob_normalize_output (_slos.py:199,213).
//
// Formulation (gather, owner-computes on the child layer):
//     c_k[s] = sum_{j : s_j > 0} U[j, mk] * c_{k-1}[s - e_j]
// No FSMap is materialised (it would be 8*m*N(k-1) bytes, 68 GB at 12 photons / 24 modes).  With
// T_i = photons strictly right of mode i and q_i = m-1-i,
//     rank_k(s)           = sum_{i<m-1} Bt[q_i][T_i]
//     rank_{k-1}(s - e_j) = rank_k(s) - E_j(s),   E_j = sum_{i<j} Dt[q_i][T_i],   Dt[q][T] = Bt[q][T]-Bt[q][T-1]
// so one walk over the modes un-ranks the child AND yields every parent rank with one shared-memory look-up per
// mode.  Bt/Dt (binomials), the U column and the factorial table are staged in shared memory.
#include "common.cuh"

#define SLOS_BLOCK 256

struct SlosArgs {
    int m, k, mk;
    const uint64_t *bt, *dt;
    const double2 *U;        // m*m row-major
    const double2 *parent;   // parent ranks [pbegin, pend)
    uint64_t pbegin, pend;
    double2 *child;          // child ranks [cbegin, cend) (may be null when probs only)
    double *probs;           // may be null
    double *sum;             // may be null
    double inv_in_fact;      // 1 / prod(in!)
    uint64_t cbegin, cend;
    int *status;
};

// MODE bit0: write coefficients, bit1: write probabilities (+ optional sum)
template <int MODE>
__global__ void __launch_bounds__(SLOS_BLOCK) slos_layer_gather_kernel(const SlosArgs a) {
    __shared__ uint64_t s_bt[FOCK_QMAX * FOCK_TMAX];
    __shared__ uint64_t s_dt[FOCK_QMAX * FOCK_TMAX];
    __shared__ double2 s_u[FOCK_QMAX];
    __shared__ double s_fact[FOCK_TMAX];
    __shared__ double s_red[SLOS_BLOCK / 32];

    const int m = a.m, k = a.k;
    for (int i = threadIdx.x; i < m * FOCK_TMAX; i += SLOS_BLOCK) {
        s_bt[i] = a.bt[i];
        s_dt[i] = a.dt[i];
    }
    for (int i = threadIdx.x; i < m; i += SLOS_BLOCK) s_u[i] = a.U[(size_t)i * m + a.mk];
    if (threadIdx.x == 0) {
        double f = 1.0;
        s_fact[0] = 1.0;
        for (int i = 1; i < FOCK_TMAX; ++i) {
            f *= (double)i;
            s_fact[i] = f;
        }
    }
    __syncthreads();

    const double2 *__restrict__ parent = a.parent - a.pbegin;  // index by absolute parent rank
    double local_sum = 0.0;
    bool oob = false;

    for (uint64_t r = a.cbegin + (uint64_t)blockIdx.x * SLOS_BLOCK + threadIdx.x; r < a.cend;
         r += (uint64_t)gridDim.x * SLOS_BLOCK) {
        uint64_t rem = r, E = 0;
        int Tprev = k;
        double2 acc = make_double2(0.0, 0.0);
        double fact = 1.0;
        int i = 0;
        for (; i < m - 1; ++i) {
            const int q = m - 1 - i;
            const uint64_t *row = s_bt + q * FOCK_TMAX;
            int T = Tprev;
            while (row[T] > rem) --T;
            rem -= row[T];
            const int si = Tprev - T;
            if (si > 0) {
                const uint64_t pr = r - E;
                if (pr < a.pbegin || pr >= a.pend) oob = true;
                else acc = cfma(s_u[i], parent[pr], acc);
                if (MODE & 2) fact *= s_fact[si];
            }
            if (T == 0) break;  // every remaining mode is empty
            E += s_dt[q * FOCK_TMAX + T];
            Tprev = T;
        }
        if (i == m - 1 && Tprev > 0) {  // last mode holds the remaining photons
            const uint64_t pr = r - E;
            if (pr < a.pbegin || pr >= a.pend) oob = true;
            else acc = cfma(s_u[m - 1], parent[pr], acc);
            if (MODE & 2) fact *= s_fact[Tprev];
        }
        if (MODE & 1) a.child[r - a.cbegin] = acc;
        if (MODE & 2) {
            const double p = (acc.x * acc.x + acc.y * acc.y) * a.inv_in_fact * fact;
            a.probs[r - a.cbegin] = p;
            local_sum += p;
        }
    }
    if (oob && a.status) atomicExch(a.status, 1);
    if ((MODE & 2) && a.sum) {
        local_sum = warp_sum(local_sum);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local_sum;
        __syncthreads();
        if (threadIdx.x < 32) {
            double v = threadIdx.x < SLOS_BLOCK / 32 ? s_red[threadIdx.x] : 0.0;
            v = warp_sum(v);
            if (threadIdx.x == 0) atomicAdd(a.sum, v);
        }
    }
}

// stand-alone epilogues: MODE 0 -> probabilities (+sum), 1 -> amplitudes
template <int MODE>
__global__ void __launch_bounds__(SLOS_BLOCK) slos_epilogue_kernel(int m, int n, const uint64_t *__restrict__ bt,
                                                                   const double2 *__restrict__ coefs, double inv_in_fact,
                                                                   double *__restrict__ probs, double2 *__restrict__ amps,
                                                                   double *sum, uint64_t begin, uint64_t end) {
    __shared__ uint64_t s_bt[FOCK_QMAX * FOCK_TMAX];
    __shared__ double s_fact[FOCK_TMAX];
    __shared__ double s_red[SLOS_BLOCK / 32];
    for (int i = threadIdx.x; i < m * FOCK_TMAX; i += SLOS_BLOCK) s_bt[i] = bt[i];
    if (threadIdx.x == 0) {
        double f = 1.0;
        s_fact[0] = 1.0;
        for (int i = 1; i < FOCK_TMAX; ++i) {
            f *= (double)i;
            s_fact[i] = f;
        }
    }
    __syncthreads();
    double local_sum = 0.0;
    for (uint64_t r = begin + (uint64_t)blockIdx.x * SLOS_BLOCK + threadIdx.x; r < end; r += (uint64_t)gridDim.x * SLOS_BLOCK) {
        uint64_t rem = r;
        int Tprev = n;
        double fact = 1.0;
        for (int i = 0; i < m - 1 && Tprev > 0; ++i) {
            const uint64_t *row = s_bt + (m - 1 - i) * FOCK_TMAX;
            int T = Tprev;
            while (row[T] > rem) --T;
            rem -= row[T];
            fact *= s_fact[Tprev - T];
            Tprev = T;
        }
        fact *= s_fact[Tprev];
        const double2 c = coefs[r - begin];
        if (MODE == 0) {
            const double p = (c.x * c.x + c.y * c.y) * inv_in_fact * fact;
            probs[r - begin] = p;
            local_sum += p;
        } else {
            const double f = sqrt(fact * inv_in_fact);
            amps[r - begin] = make_double2(c.x * f, c.y * f);
        }
    }
    if (MODE == 0 && sum) {
        local_sum = warp_sum(local_sum);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local_sum;
        __syncthreads();
        if (threadIdx.x < 32) {
            double v = threadIdx.x < SLOS_BLOCK / 32 ? s_red[threadIdx.x] : 0.0;
            v = warp_sum(v);
            if (threadIdx.x == 0) atomicAdd(sum, v);
        }
    }
}

// ---------------------------------------------------------------- host side
static int slos_check(const char *who, fock_ctx *c, int m, int k) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "%s: ctx is NULL", who);
    FOCK_REQUIRE(m >= 1 && m <= FOCK_QMAX, FOCK_ERR_LIMIT, "%s: m=%d outside [1,%d]", who, m, FOCK_QMAX);
    FOCK_REQUIRE(k >= 0 && k <= FOCK_NMAX, FOCK_ERR_LIMIT, "%s: photon count %d outside [0,%d]", who, k, FOCK_NMAX);
    FOCK_REQUIRE(fock_count(m, k) != UINT64_MAX, FOCK_ERR_LIMIT, "%s: C(%d+%d-1,%d) overflows 64 bits", who, k, m, k);
    return 0;
}

static unsigned slos_grid(fock_ctx *c, uint64_t cnt) {
    uint64_t g = (cnt + SLOS_BLOCK - 1) / SLOS_BLOCK;
    uint64_t cap = (uint64_t)c->sm_count * 8;  // 8 CTAs of 256 threads = 2048 threads / SM, persistent grid-stride
    if (g > cap) g = cap;
    return (unsigned)(g ? g : 1);
}

static int slos_layer_impl(fock_ctx *c, int m, int k, const double *d_U, int mk, const double *d_parent, uint64_t pb,
                           uint64_t pe, double *d_child, double *d_probs, double *d_sum, double in_prodnfact, uint64_t cb,
                           uint64_t ce, void *stream, const char *who) {
    if (int rc = slos_check(who, c, m, k)) return rc;
    FOCK_REQUIRE(k >= 1, FOCK_ERR_ARG, "%s: child layer must hold >= 1 photon", who);
    FOCK_REQUIRE(mk >= 0 && mk < m, FOCK_ERR_ARG, "%s: input mode %d outside [0,%d)", who, mk, m);
    FOCK_REQUIRE(cb <= ce && ce <= fock_count(m, k), FOCK_ERR_ARG, "%s: bad child range", who);
    FOCK_REQUIRE(pb <= pe && pe <= fock_count(m, k - 1), FOCK_ERR_ARG, "%s: bad parent range", who);
    FOCK_REQUIRE(d_U && d_parent, FOCK_ERR_ARG, "%s: NULL device pointer", who);
    FOCK_REQUIRE(d_child || d_probs, FOCK_ERR_ARG, "%s: no output buffer", who);
    if (cb == ce) return FOCK_OK;
    ScopedDevice sd(c->device);
    SlosArgs a;
    a.m = m; a.k = k; a.mk = mk;
    a.bt = c->d_bt; a.dt = c->d_dt;
    a.U = (const double2 *)d_U;
    a.parent = (const double2 *)d_parent;
    a.pbegin = pb; a.pend = pe;
    a.child = (double2 *)d_child;
    a.probs = d_probs;
    a.sum = d_sum;
    a.inv_in_fact = 1.0 / in_prodnfact;
    a.cbegin = cb; a.cend = ce;
    a.status = c->d_status;
    unsigned grid = slos_grid(c, ce - cb);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_probs && d_child) slos_layer_gather_kernel<3><<<grid, SLOS_BLOCK, 0, st>>>(a);
    else if (d_probs) slos_layer_gather_kernel<2><<<grid, SLOS_BLOCK, 0, st>>>(a);
    else slos_layer_gather_kernel<1><<<grid, SLOS_BLOCK, 0, st>>>(a);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    return FOCK_OK;
}

extern "C" int slos_layer(fock_ctx *c, int m, int k, const double *d_U, int mk, const double *d_parent, uint64_t pb,
                          uint64_t pe, double *d_child, uint64_t cb, uint64_t ce, void *stream) {
    FOCK_REQUIRE(d_child != nullptr, FOCK_ERR_ARG, "slos_layer: d_child is NULL");
    return slos_layer_impl(c, m, k, d_U, mk, d_parent, pb, pe, d_child, nullptr, nullptr, 1.0, cb, ce, stream, "slos_layer");
}

extern "C" int slos_layer_probs(fock_ctx *c, int m, int k, const double *d_U, int mk, const double *d_parent, uint64_t pb,
                                uint64_t pe, double *d_child, double *d_probs, double *d_sum, double in_prodnfact,
                                uint64_t cb, uint64_t ce, void *stream) {
    FOCK_REQUIRE(d_probs != nullptr, FOCK_ERR_ARG, "slos_layer_probs: d_probs is NULL");
    FOCK_REQUIRE(in_prodnfact > 0, FOCK_ERR_ARG, "slos_layer_probs: in_prodnfact must be > 0");
    return slos_layer_impl(c, m, k, d_U, mk, d_parent, pb, pe, d_child, d_probs, d_sum, in_prodnfact, cb, ce, stream,
                           "slos_layer_probs");
}

extern "C" int slos_probs_epilogue(fock_ctx *c, int m, int n, const double *d_coefs, double in_prodnfact, double *d_probs,
                                   double *d_sum, uint64_t begin, uint64_t end, void *stream) {
    if (int rc = slos_check("slos_probs_epilogue", c, m, n)) return rc;
    FOCK_REQUIRE(begin <= end && end <= fock_count(m, n), FOCK_ERR_ARG, "slos_probs_epilogue: bad range");
    FOCK_REQUIRE(d_coefs && d_probs && in_prodnfact > 0, FOCK_ERR_ARG, "slos_probs_epilogue: bad argument");
    if (begin == end) return FOCK_OK;
    ScopedDevice sd(c->device);
    slos_epilogue_kernel<0><<<slos_grid(c, end - begin), SLOS_BLOCK, 0, (cudaStream_t)stream>>>(
        m, n, c->d_bt, (const double2 *)d_coefs, 1.0 / in_prodnfact, d_probs, nullptr, d_sum, begin, end);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    return FOCK_OK;
}

extern "C" int slos_amplitudes_epilogue(fock_ctx *c, int m, int n, const double *d_coefs, double in_prodnfact, double *d_amps,
                                        uint64_t begin, uint64_t end, void *stream) {
    if (int rc = slos_check("slos_amplitudes_epilogue", c, m, n)) return rc;
    FOCK_REQUIRE(begin <= end && end <= fock_count(m, n), FOCK_ERR_ARG, "slos_amplitudes_epilogue: bad range");
    FOCK_REQUIRE(d_coefs && d_amps && in_prodnfact > 0, FOCK_ERR_ARG, "slos_amplitudes_epilogue: bad argument");
    if (begin == end) return FOCK_OK;
    ScopedDevice sd(c->device);
    slos_epilogue_kernel<1><<<slos_grid(c, end - begin), SLOS_BLOCK, 0, (cudaStream_t)stream>>>(
        m, n, c->d_bt, (const double2 *)d_coefs, 1.0 / in_prodnfact, nullptr, (double2 *)d_amps, nullptr, begin, end);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    return FOCK_OK;
}

// perceval/backends/_slos.py:61-86 with a single target: take the mode with most remaining photons, first index on ties
extern "C" int slos_order(int m, const uint8_t *in_state, int *order) {
    FOCK_REQUIRE(m >= 1 && in_state && order, FOCK_ERR_ARG, "slos_order: bad argument");
    int t[256];
    FOCK_REQUIRE(m <= 256, FOCK_ERR_LIMIT, "slos_order: m > 256");
    int n = 0;
    for (int i = 0; i < m; ++i) { t[i] = in_state[i]; n += t[i]; }
    for (int k = 0; k < n; ++k) {
        int best = 0;
        for (int i = 1; i < m; ++i)
            if (t[i] > t[best]) best = i;
        order[k] = best;
        t[best]--;
    }
    return FOCK_OK;
}

static double host_prodnfact(int m, const uint8_t *s) {
    double p = 1.0;
    for (int i = 0; i < m; ++i)
        for (int v = 2; v <= s[i]; ++v) p *= v;
    return p;
}

extern "C" int slos_prob_distribution(fock_ctx *c, int m, const double *d_U, const uint8_t *in_state, double *d_work_a,
                                      double *d_work_b, double *d_coefs, double *d_probs, double *d_sum, void *stream) {
    FOCK_REQUIRE(c && d_U && in_state && d_probs, FOCK_ERR_ARG, "slos_prob_distribution: bad argument");
    int n = 0;
    for (int i = 0; i < m; ++i) n += in_state[i];
    if (int rc = slos_check("slos_prob_distribution", c, m, n)) return rc;
    ScopedDevice sd(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_sum) FOCK_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), st));
    const double one[2] = {1.0, 0.0};
    if (n == 0) {
        const double p1 = 1.0;
        FOCK_CUDA(cudaMemcpyAsync(d_probs, &p1, sizeof(double), cudaMemcpyHostToDevice, st));
        if (d_coefs) FOCK_CUDA(cudaMemcpyAsync(d_coefs, one, 16, cudaMemcpyHostToDevice, st));
        if (d_sum) FOCK_CUDA(cudaMemcpyAsync(d_sum, &p1, sizeof(double), cudaMemcpyHostToDevice, st));
        FOCK_CUDA(cudaStreamSynchronize(st));
        return FOCK_OK;
    }
    int order[FOCK_NMAX];
    slos_order(m, in_state, order);
    const double inf = host_prodnfact(m, in_state);
    // layers alternate between the two workspaces so that layer n-1 lands in work_a
    FOCK_REQUIRE(n == 1 || d_work_a, FOCK_ERR_ARG, "slos_prob_distribution: d_work_a is NULL");
    FOCK_REQUIRE(n <= 2 || d_work_b, FOCK_ERR_ARG, "slos_prob_distribution: d_work_b is NULL");
    // layer 0 = [1] lives in the scratch area of the context
    double *layer0 = c->d_scratch + 8;
    FOCK_CUDA(cudaMemcpyAsync(layer0, one, 16, cudaMemcpyHostToDevice, st));
    const double *prev = layer0;
    for (int k = 1; k <= n; ++k) {
        const uint64_t np = fock_count(m, k - 1), nc = fock_count(m, k);
        if (k == n) {
            int rc = slos_layer_impl(c, m, k, d_U, order[k - 1], prev, 0, np, d_coefs, d_probs, d_sum, inf, 0, nc, stream,
                                     "slos_prob_distribution");
            if (rc) return rc;
        } else {
            double *cur = ((n - 1 - k) % 2 == 0) ? d_work_a : d_work_b;
            int rc = slos_layer_impl(c, m, k, d_U, order[k - 1], prev, 0, np, cur, nullptr, nullptr, 1.0, 0, nc, stream,
                                     "slos_prob_distribution");
            if (rc) return rc;
            prev = cur;
        }
    }
    // the 16-byte host source of layer0 must stay valid until the copy ran
    FOCK_CUDA(cudaStreamSynchronize(st));
    return FOCK_OK;
}

extern "C" int slos_prob_distribution_host(fock_ctx *c, int m, const double *h_U, const uint8_t *in_state, double *h_probs,
                                           double *h_sum) {
    FOCK_REQUIRE(c && h_U && in_state && h_probs, FOCK_ERR_ARG, "slos_prob_distribution_host: bad argument");
    int n = 0;
    for (int i = 0; i < m; ++i) n += in_state[i];
    if (int rc = slos_check("slos_prob_distribution_host", c, m, n)) return rc;
    ScopedDevice sd(c->device);
    const uint64_t N = fock_count(m, n), Na = n >= 1 ? fock_count(m, n - 1) : 0, Nb = n >= 2 ? fock_count(m, n - 2) : 0;
    double *dU = nullptr, *da = nullptr, *db = nullptr, *dp = nullptr, *ds = nullptr;
    int rc = FOCK_OK;
    cudaError_t e;
#define HOSTCALL(x) do { e = (x); if (e != cudaSuccess) { rc = fock_check_cuda(e, #x); goto done; } } while (0)
    HOSTCALL(cudaMalloc(&dU, 16 * (size_t)m * m));
    if (Na) HOSTCALL(cudaMalloc(&da, 16 * Na));
    if (Nb) HOSTCALL(cudaMalloc(&db, 16 * Nb));
    HOSTCALL(cudaMalloc(&dp, 8 * N));
    HOSTCALL(cudaMalloc(&ds, 8));
    HOSTCALL(cudaMemcpy(dU, h_U, 16 * (size_t)m * m, cudaMemcpyHostToDevice));
    rc = slos_prob_distribution(c, m, dU, in_state, da, db, nullptr, dp, ds, nullptr);
    if (rc) goto done;
    HOSTCALL(cudaMemcpy(h_probs, dp, 8 * N, cudaMemcpyDeviceToHost));
    if (h_sum) HOSTCALL(cudaMemcpy(h_sum, ds, 8, cudaMemcpyDeviceToHost));
    {
        int status = 0;
        HOSTCALL(cudaMemcpy(&status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost));
        if (status) { fock_set_error("slos: parent rank outside the resident window"); rc = FOCK_ERR_ARG; }
    }
done:
#undef HOSTCALL
    cudaFree(dU); cudaFree(da); cudaFree(db); cudaFree(dp); cudaFree(ds);
    return rc;
}
