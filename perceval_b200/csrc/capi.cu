// capi.cu -- context lifecycle, error reporting, binomial tables, FSArray count / rank / unrank.
//
// Replaces xq.FSArray(m,n).count()/.find()/iteration (reference perceval/backends/_slos.py:156-168,190;
// perceval/utils/states.py:255-298).  The order (descending lexicographic) is the one pinned by reference
// tests/utils/test_statevector.py:430-438.  No state list is ever materialised: rank and unrank are closed-form
// walks over a table of binomials:
//     rank(s) = sum_{i=0}^{m-2} Bt[q_i][T_i],  T_i = photons strictly right of mode i,  q_i = m-1-i,
//     Bt[q][T] = C(T-1+q, q)  (0 for T = 0)                                   (hockey-stick form)
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

static thread_local char g_err[512] = "";

void fock_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int fock_check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    fock_set_error("CUDA error %s (%s) in %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
    return FOCK_ERR_CUDA;
}

extern "C" const char *fock_last_error(void) { return g_err; }
extern "C" const char *fock_version(void) { return "fock_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------- host binomial tables
static uint64_t h_bt[FOCK_QMAX * FOCK_TMAX];
static uint64_t h_dt[FOCK_QMAX * FOCK_TMAX];
static std::once_flag g_tables_once;

static uint64_t binom_sat(int n, int k) {
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) {
        r = r * (unsigned)(n - k + i) / (unsigned)i;
        if (r > (unsigned __int128)UINT64_MAX) return UINT64_MAX;
    }
    return (uint64_t)r;
}

static void init_tables() {
    for (int q = 0; q < FOCK_QMAX; ++q)
        for (int T = 0; T < FOCK_TMAX; ++T) {
            uint64_t b = (T == 0) ? 0 : binom_sat(T - 1 + q, q);
            h_bt[q * FOCK_TMAX + T] = b;
        }
    for (int q = 0; q < FOCK_QMAX; ++q)
        for (int T = 0; T < FOCK_TMAX; ++T) {
            uint64_t b = h_bt[q * FOCK_TMAX + T], a = T ? h_bt[q * FOCK_TMAX + T - 1] : 0;
            h_dt[q * FOCK_TMAX + T] = (b == UINT64_MAX) ? UINT64_MAX : b - a;
        }
}
const uint64_t *fock_host_bt() {
    std::call_once(g_tables_once, init_tables);
    return h_bt;
}
const uint64_t *fock_host_dt() {
    std::call_once(g_tables_once, init_tables);
    return h_dt;
}

extern "C" uint64_t fock_count(int m, int n) {
    if (n < 0 || m < 0) return 0;
    if (m == 0) return n == 0 ? 1 : 0;
    return binom_sat(n + m - 1, n);
}

// ---------------------------------------------------------------- context
void slos_mu_init(fock_ctx *c);      // slos_mu.cu
void slos_mu_destroy(fock_ctx *c);
void slos_thin_destroy(fock_ctx *c); // slos_thin.cu
extern "C" int fock_create(int device, fock_ctx **out) {
    FOCK_REQUIRE(out != nullptr, FOCK_ERR_ARG, "fock_create: out is NULL");
    int ndev = 0;
    FOCK_CUDA(cudaGetDeviceCount(&ndev));
    FOCK_REQUIRE(device >= 0 && device < ndev, FOCK_ERR_ARG, "fock_create: device %d out of range (%d devices)", device, ndev);
    ScopedDevice sd(device);
    cudaDeviceProp prop;
    FOCK_CUDA(cudaGetDeviceProperties(&prop, device));
    FOCK_REQUIRE(prop.major >= 10, FOCK_ERR_LIMIT,
                 "fock_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    fock_ctx *c = new fock_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->total_mem = prop.totalGlobalMem;
    c->launches = 0;
    c->mu_state = nullptr;
    c->ev_begin = c->ev_end = nullptr;
    c->side[0] = c->side[1] = nullptr;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (int i = 0; i < 2; ++i) FOCK_CUDA(cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, hi));
    }
    slos_mu_init(c);
    {   // keep stream-ordered scratch (StreamScratch) in the pool between calls
        cudaMemPool_t pool;
        uint64_t keep = UINT64_MAX;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    size_t tb = sizeof(uint64_t) * FOCK_QMAX * FOCK_TMAX;
    FOCK_CUDA(cudaMalloc(&c->d_bt, tb));
    FOCK_CUDA(cudaMalloc(&c->d_dt, tb));
    FOCK_CUDA(cudaMalloc(&c->d_status, sizeof(int)));
    FOCK_CUDA(cudaMalloc(&c->d_scratch, 64 * sizeof(double)));
    FOCK_CUDA(cudaMalloc(&c->d_vacuum, 2 * sizeof(double)));
    {
        const double one[2] = {1.0, 0.0};
        FOCK_CUDA(cudaMemcpy(c->d_vacuum, one, sizeof one, cudaMemcpyHostToDevice));
    }
    FOCK_CUDA(cudaMemcpy(c->d_bt, fock_host_bt(), tb, cudaMemcpyHostToDevice));
    FOCK_CUDA(cudaMemcpy(c->d_dt, fock_host_dt(), tb, cudaMemcpyHostToDevice));
    FOCK_CUDA(cudaMemset(c->d_status, 0, sizeof(int)));
    *out = c;
    return FOCK_OK;
}


extern "C" int fock_destroy(fock_ctx *c) {
    if (!c) return FOCK_OK;
    ScopedDevice sd(c->device);
    slos_mu_destroy(c);
    slos_thin_destroy(c);
    for (int i = 0; i < 2; ++i)
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
    cudaFree(c->d_bt);
    cudaFree(c->d_dt);
    cudaFree(c->d_status);
    cudaFree(c->d_scratch);
    cudaFree(c->d_vacuum);
    delete c;
    return FOCK_OK;
}

extern "C" int fock_device_info(fock_ctx *c, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "fock_device_info: ctx is NULL");
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (total_mem) *total_mem = c->total_mem;
    return FOCK_OK;
}

extern "C" int fock_check_status(fock_ctx *c, void *stream) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "fock_check_status: ctx is NULL");
    ScopedDevice sd(c->device);
    int status = 0;
    FOCK_CUDA(cudaMemcpyAsync(&status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    FOCK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (status) {
        FOCK_CUDA(cudaMemsetAsync(c->d_status, 0, sizeof(int), (cudaStream_t)stream));
        fock_set_error("device kernel flagged an error (SLOS parent rank outside the resident window)");
        return FOCK_ERR_ARG;
    }
    return FOCK_OK;
}

extern "C" uint64_t fock_launch_count(fock_ctx *c) { return c ? c->launches : 0; }

extern "C" int fock_profile_events(fock_ctx *c, void *ev_begin, void *ev_end) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "fock_profile_events: ctx is NULL");
    FOCK_REQUIRE((ev_begin == nullptr) == (ev_end == nullptr), FOCK_ERR_ARG, "fock_profile_events: pass two events or two NULLs");
    c->ev_begin = (cudaEvent_t)ev_begin;
    c->ev_end = (cudaEvent_t)ev_end;
    return FOCK_OK;
}

// ---------------------------------------------------------------- host rank / unrank (any m <= 64, n <= 32)
static int check_mn(const char *who, int m, int n) {
    FOCK_REQUIRE(m >= 1 && m <= FOCK_QMAX, FOCK_ERR_LIMIT, "%s: m=%d outside [1,%d]", who, m, FOCK_QMAX);
    FOCK_REQUIRE(n >= 0 && n <= FOCK_NMAX, FOCK_ERR_LIMIT, "%s: n=%d outside [0,%d]", who, n, FOCK_NMAX);
    FOCK_REQUIRE(fock_count(m, n) != UINT64_MAX, FOCK_ERR_LIMIT, "%s: C(%d+%d-1,%d) overflows 64 bits", who, n, m, n);
    return 0;
}

extern "C" int fock_rank_host(int m, int n, const uint8_t *s, uint64_t cnt, uint64_t *out) {
    if (int rc = check_mn("fock_rank_host", m, n)) return rc;
    const uint64_t *bt = fock_host_bt();
    for (uint64_t i = 0; i < cnt; ++i) {
        const uint8_t *st = s + i * (uint64_t)m;
        int tot = 0;
        for (int j = 0; j < m; ++j) tot += st[j];
        if (tot != n) {  // xq.FSArray.find -> npos
            out[i] = UINT64_MAX;
            continue;
        }
        uint64_t r = 0;
        int T = n;
        for (int j = 0; j < m - 1; ++j) {
            T -= st[j];
            r += bt[(m - 1 - j) * FOCK_TMAX + T];
        }
        out[i] = r;
    }
    return FOCK_OK;
}

extern "C" int fock_unrank_host(int m, int n, const uint64_t *ranks, uint64_t cnt, uint8_t *out) {
    if (int rc = check_mn("fock_unrank_host", m, n)) return rc;
    const uint64_t *bt = fock_host_bt();
    uint64_t N = fock_count(m, n);
    for (uint64_t i = 0; i < cnt; ++i) {
        uint64_t rem = ranks[i];
        FOCK_REQUIRE(rem < N, FOCK_ERR_ARG, "fock_unrank_host: rank %llu >= count %llu", (unsigned long long)rem, (unsigned long long)N);
        uint8_t *st = out + i * (uint64_t)m;
        int Tprev = n;
        for (int j = 0; j < m - 1; ++j) {
            const uint64_t *row = bt + (m - 1 - j) * FOCK_TMAX;
            int T = Tprev;
            while (row[T] > rem) --T;
            rem -= row[T];
            st[j] = (uint8_t)(Tprev - T);
            Tprev = T;
        }
        st[m - 1] = (uint8_t)Tprev;
    }
    return FOCK_OK;
}

// ---------------------------------------------------------------- device rank / unrank
__global__ void __launch_bounds__(256) rank_kernel(int m, int n, const uint64_t *__restrict__ bt, const uint8_t *__restrict__ states,
                                                   uint64_t cnt, uint64_t *__restrict__ ranks) {
    __shared__ uint64_t s_bt[FOCK_QMAX * FOCK_TMAX];
    for (int i = threadIdx.x; i < m * FOCK_TMAX; i += blockDim.x) s_bt[i] = bt[i];
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t *st = states + i * (uint64_t)m;
        int tot = 0;
        for (int j = 0; j < m; ++j) tot += st[j];
        uint64_t r = 0;
        int T = n;
        for (int j = 0; j < m - 1; ++j) {
            T -= st[j];
            r += s_bt[(m - 1 - j) * FOCK_TMAX + max(T, 0)];
        }
        ranks[i] = (tot == n) ? r : UINT64_MAX;
    }
}

__global__ void __launch_bounds__(256) unrank_kernel(int m, int n, const uint64_t *__restrict__ bt, const uint64_t *__restrict__ ranks,
                                                     uint64_t first, uint64_t cnt, uint8_t *__restrict__ states) {
    __shared__ uint64_t s_bt[FOCK_QMAX * FOCK_TMAX];
    for (int i = threadIdx.x; i < m * FOCK_TMAX; i += blockDim.x) s_bt[i] = bt[i];
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t rem = ranks ? ranks[i] : first + i;
        uint8_t *st = states + i * (uint64_t)m;
        int Tprev = n;
        for (int j = 0; j < m - 1; ++j) {
            const uint64_t *row = s_bt + (m - 1 - j) * FOCK_TMAX;
            int T = Tprev;
            while (row[T] > rem) --T;
            rem -= row[T];
            st[j] = (uint8_t)(Tprev - T);
            Tprev = T;
        }
        st[m - 1] = (uint8_t)Tprev;
    }
}

static unsigned grid_for(fock_ctx *c, uint64_t cnt, int block) {
    uint64_t g = (cnt + block - 1) / block;
    uint64_t cap = (uint64_t)c->sm_count * 16;
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    return (unsigned)g;
}

extern "C" int fock_rank(fock_ctx *c, int m, int n, const uint8_t *d_states, uint64_t cnt, uint64_t *d_ranks, void *stream) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "fock_rank: ctx is NULL");
    if (int rc = check_mn("fock_rank", m, n)) return rc;
    if (cnt == 0) return FOCK_OK;
    ScopedDevice sd(c->device);
    rank_kernel<<<grid_for(c, cnt, 256), 256, 0, (cudaStream_t)stream>>>(m, n, c->d_bt, d_states, cnt, d_ranks);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    return FOCK_OK;
}

extern "C" int fock_unrank(fock_ctx *c, int m, int n, const uint64_t *d_ranks, uint64_t cnt, uint8_t *d_states, void *stream) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "fock_unrank: ctx is NULL");
    FOCK_REQUIRE(d_ranks != nullptr, FOCK_ERR_ARG, "fock_unrank: d_ranks is NULL");
    if (int rc = check_mn("fock_unrank", m, n)) return rc;
    if (cnt == 0) return FOCK_OK;
    ScopedDevice sd(c->device);
    unrank_kernel<<<grid_for(c, cnt, 256), 256, 0, (cudaStream_t)stream>>>(m, n, c->d_bt, d_ranks, 0, cnt, d_states);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    return FOCK_OK;
}

extern "C" int fock_enumerate(fock_ctx *c, int m, int n, uint64_t begin, uint64_t end, uint8_t *d_states, void *stream) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "fock_enumerate: ctx is NULL");
    if (int rc = check_mn("fock_enumerate", m, n)) return rc;
    FOCK_REQUIRE(begin <= end && end <= fock_count(m, n), FOCK_ERR_ARG, "fock_enumerate: bad range");
    if (begin == end) return FOCK_OK;
    ScopedDevice sd(c->device);
    unrank_kernel<<<grid_for(c, end - begin, 256), 256, 0, (cudaStream_t)stream>>>(m, n, c->d_bt, nullptr, begin, end - begin, d_states);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    return FOCK_OK;
}

// ---- FSMask on the device (replaces xq.FSMask(m, n, masks[, at_least_modes]).match(state) over a whole FSArray,
// reference perceval/backends/_abstract_backends.py:130-137, tests/utils/test_mask.py:32-45): flags[i] = 1 iff state
// #(begin+i) of FSArray(m, n) matches ANY mask.  conds holds nmask*m int8: -1 accepts anything, v >= 0 fixes the count
// (">= v" on the modes whose bit is set in at_least_bits).  allow_missing is the partial match used on intermediate
// layers: a mode passes if its count can still grow into the condition.
#define FOCK_MASK_MAXCOND 2048
__constant__ signed char c_mask_cond[FOCK_MASK_MAXCOND];

// allow_missing: 0 exact match, 1 the reference's partial match, 2 + b partial match that can still be completed with b more
// photons (the conditioned modes lack sum max(0, c_i - v_i) photons; pruned SLOS layers, slos_masked.cu)
__host__ __device__ inline bool mask_match_one(const uint8_t *st, int m, const signed char *cond, int nmask, uint64_t at_least_bits, int allow_missing) {
    for (int k = 0; k < nmask; ++k) {
        const signed char *cd = cond + k * m;
        bool ok = true;
        int deficit = 0;
        for (int i = 0; i < m && ok; ++i) {
            const int c = cd[i];
            if (c < 0) continue;
            const int v = st[i];
            const bool ge = i < 64 && ((at_least_bits >> i) & 1ull);
            if (allow_missing) {
                ok = ge || v <= c;
                if (v < c) deficit += c - v;
            } else ok = ge ? (v >= c) : (v == c);
        }
        if (ok && allow_missing >= 2 && deficit > allow_missing - 2) ok = false;
        if (ok) return true;
    }
    return false;
}

__global__ void __launch_bounds__(256) mask_match_kernel(int m, int n, const uint64_t *__restrict__ bt, int nmask, uint64_t at_least_bits,
                                                         int allow_missing, uint64_t first, uint64_t cnt, uint8_t *__restrict__ flags) {
    __shared__ uint64_t s_bt[FOCK_QMAX * FOCK_TMAX];
    for (int i = threadIdx.x; i < m * FOCK_TMAX; i += blockDim.x) s_bt[i] = bt[i];
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t rem = first + i;
        uint8_t st[FOCK_QMAX];
        int Tprev = n;
        for (int j = 0; j < m - 1; ++j) {
            const uint64_t *row = s_bt + (m - 1 - j) * FOCK_TMAX;
            int T = Tprev;
            while (row[T] > rem) --T;
            rem -= row[T];
            st[j] = (uint8_t)(Tprev - T);
            Tprev = T;
        }
        st[m - 1] = (uint8_t)Tprev;
        flags[i] = mask_match_one(st, m, c_mask_cond, nmask, at_least_bits, allow_missing) ? 1 : 0;
    }
}

static int mask_check(const char *who, int m, int n, const int8_t *h_conds, int nmask) {
    if (int rc = check_mn(who, m, n)) return rc;
    FOCK_REQUIRE(h_conds != nullptr && nmask >= 1, FOCK_ERR_ARG, "%s: no mask conditions", who);
    FOCK_REQUIRE((size_t)nmask * m <= FOCK_MASK_MAXCOND, FOCK_ERR_LIMIT, "%s: more than %d mask conditions", who, FOCK_MASK_MAXCOND);
    return FOCK_OK;
}

extern "C" int fock_mask_match(fock_ctx *c, int m, int n, const int8_t *h_conds, int nmask, uint64_t at_least_bits, int allow_missing,
                               uint64_t begin, uint64_t end, uint8_t *d_flags, void *stream) {
    FOCK_REQUIRE(c != nullptr && d_flags != nullptr, FOCK_ERR_ARG, "fock_mask_match: NULL argument");
    if (int rc = mask_check("fock_mask_match", m, n, h_conds, nmask)) return rc;
    FOCK_REQUIRE(begin <= end && end <= fock_count(m, n), FOCK_ERR_ARG, "fock_mask_match: bad range");
    if (begin == end) return FOCK_OK;
    ScopedDevice sd(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    static std::mutex mask_lock;   // one constant-bank copy of the conditions per device: calls are serialised (each ends in a sync)
    std::lock_guard<std::mutex> guard(mask_lock);
    FOCK_CUDA(cudaMemcpyToSymbolAsync(c_mask_cond, h_conds, (size_t)nmask * m, 0, cudaMemcpyHostToDevice, st));
    mask_match_kernel<<<grid_for(c, end - begin, 256), 256, 0, st>>>(m, n, c->d_bt, nmask, at_least_bits, allow_missing, begin, end - begin, d_flags);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    FOCK_CUDA(cudaStreamSynchronize(st));   // h_conds may be pageable host memory: it must stay valid until the copy ran
    return FOCK_OK;
}

extern "C" int fock_mask_match_host(int m, int n, const int8_t *h_conds, int nmask, uint64_t at_least_bits, int allow_missing,
                                    const uint8_t *h_states, uint64_t cnt, uint8_t *h_flags) {
    FOCK_REQUIRE(h_states != nullptr && h_flags != nullptr, FOCK_ERR_ARG, "fock_mask_match_host: NULL argument");
    if (int rc = mask_check("fock_mask_match_host", m, n, h_conds, nmask)) return rc;
    for (uint64_t i = 0; i < cnt; ++i)
        h_flags[i] = mask_match_one(h_states + i * (uint64_t)m, m, (const signed char *)h_conds, nmask, at_least_bits, allow_missing) ? 1 : 0;
    return FOCK_OK;
}
