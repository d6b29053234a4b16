// slos_masked.cu -- SLOS layers over a PRUNED rank space (masks / heralds).
//
// Reference: with a mask, SLOSBackend builds every layer on xq.FSArray(m, k, mask) -- only the states that can still grow
// into an output the mask accepts (perceval/backends/_slos.py:156-166; mask semantics _abstract_backends.py:103-137,
// tests/utils/test_mask.py:32-45) -- so masked layers shrink.  Here a pruned layer k is a sorted list of kept ranks of
// FSArray(m, k) plus a packed complex128 vector in the same order.  The kept sets are closed under photon removal
// (removing a photon never breaks a partial match), so every parent of a kept child is kept: its packed position is found
// by binary search in the parent's rank list (log2 |kept| L2-resident probes per edge; the lists are 8 B / kept state).
//   child_c[i] = sum_{j: s_j>0} U[j,mk] * parent_c[pos(rank_i - E_j)]
// Optional fused epilogue on the last layer: probabilities |c|^2 prod(s!)/prod(in!) (+ sum) and/or amplitudes
// c sqrt(prod(s!)/prod(in!)) (reference _slos.py:187-199).  Accumulation order = modes ascending, as every other kernel.
#include "common.cuh"

#define MSK_BLOCK 256

struct MaskedArgs {
    int m, k, mk;
    const uint64_t *bt, *dt;
    const double2 *U;
    const uint64_t *cranks;   // kept child ranks, ascending
    uint64_t nc;
    const uint64_t *pranks;   // kept parent ranks, ascending
    uint64_t np;
    const double2 *parent;    // packed parent coefficients
    double2 *child;           // packed child coefficients (MODE & 1)
    double *probs;            // (MODE & 2)
    double2 *amps;            // (MODE & 4)
    double *sum;
    double inv_in_fact;
    int *status;
};

__device__ __forceinline__ uint64_t msk_find(const uint64_t *__restrict__ v, uint64_t n, uint64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(v + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n && __ldg(v + lo) == key) ? lo : UINT64_MAX;
}

template <int MODE>
__global__ void __launch_bounds__(MSK_BLOCK) slos_masked_layer_kernel(const MaskedArgs a) {
    __shared__ uint64_t s_bt[FOCK_QMAX * FOCK_TMAX];
    __shared__ uint64_t s_dt[FOCK_QMAX * FOCK_TMAX];
    __shared__ double2 s_u[FOCK_QMAX];
    __shared__ double s_fact[FOCK_TMAX];
    __shared__ double s_red[MSK_BLOCK / 32];
    const int m = a.m, k = a.k;
    for (int i = threadIdx.x; i < m * FOCK_TMAX; i += MSK_BLOCK) {
        s_bt[i] = a.bt[i];
        s_dt[i] = a.dt[i];
    }
    for (int i = threadIdx.x; i < m; i += MSK_BLOCK) s_u[i] = a.U[(size_t)i * m + a.mk];
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
    bool missing = false;
    for (uint64_t idx = (uint64_t)blockIdx.x * MSK_BLOCK + threadIdx.x; idx < a.nc; idx += (uint64_t)gridDim.x * MSK_BLOCK) {
        const uint64_t r = a.cranks[idx];
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
                const uint64_t pos = msk_find(a.pranks, a.np, r - E);
                if (pos == UINT64_MAX) missing = true;
                else acc = cfma(s_u[i], a.parent[pos], acc);
                if (MODE & 6) fact *= s_fact[si];
            }
            if (T == 0) break;
            E += s_dt[q * FOCK_TMAX + T];
            Tprev = T;
        }
        if (i == m - 1 && Tprev > 0) {
            const uint64_t pos = msk_find(a.pranks, a.np, r - E);
            if (pos == UINT64_MAX) missing = true;
            else acc = cfma(s_u[m - 1], a.parent[pos], acc);
            if (MODE & 6) fact *= s_fact[Tprev];
        }
        if (MODE & 1) a.child[idx] = acc;
        if (MODE & 2) {
            const double p = (acc.x * acc.x + acc.y * acc.y) * a.inv_in_fact * fact;
            a.probs[idx] = p;
            local_sum += p;
        }
        if (MODE & 4) {
            const double f = sqrt(fact * a.inv_in_fact);
            a.amps[idx] = make_double2(acc.x * f, acc.y * f);
        }
    }
    if (missing && a.status) atomicExch(a.status, 1);
    if ((MODE & 2) && a.sum) {
        local_sum = warp_sum(local_sum);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local_sum;
        __syncthreads();
        if (threadIdx.x < 32) {
            double v = threadIdx.x < MSK_BLOCK / 32 ? s_red[threadIdx.x] : 0.0;
            v = warp_sum(v);
            if (threadIdx.x == 0) atomicAdd(a.sum, v);
        }
    }
}

extern "C" int slos_layer_masked(fock_ctx *c, int m, int k, const double *d_U, int mk, const uint64_t *d_parent_ranks, uint64_t n_parent,
                                 const double *d_parent, const uint64_t *d_child_ranks, uint64_t n_child, double *d_child, double *d_probs,
                                 double *d_amps, double *d_sum, double in_prodnfact, void *stream) {
    FOCK_REQUIRE(c != nullptr, FOCK_ERR_ARG, "slos_layer_masked: ctx is NULL");
    FOCK_REQUIRE(m >= 1 && m <= FOCK_QMAX && k >= 1 && k <= FOCK_NMAX, FOCK_ERR_LIMIT, "slos_layer_masked: m=%d, k=%d outside the limits", m, k);
    FOCK_REQUIRE(fock_count(m, k) != UINT64_MAX, FOCK_ERR_LIMIT, "slos_layer_masked: C(%d+%d-1,%d) overflows 64 bits", k, m, k);
    FOCK_REQUIRE(mk >= 0 && mk < m, FOCK_ERR_ARG, "slos_layer_masked: input mode %d outside [0,%d)", mk, m);
    FOCK_REQUIRE(d_child || d_probs || d_amps, FOCK_ERR_ARG, "slos_layer_masked: no output buffer");
    FOCK_REQUIRE((d_probs == nullptr && d_amps == nullptr) || in_prodnfact > 0, FOCK_ERR_ARG, "slos_layer_masked: in_prodnfact must be > 0");
    if (n_child == 0) return FOCK_OK;
    FOCK_REQUIRE(d_U && d_child_ranks && (n_parent == 0 || (d_parent && d_parent_ranks)), FOCK_ERR_ARG, "slos_layer_masked: NULL device pointer");
    ScopedDevice sd(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    MaskedArgs a;
    a.m = m; a.k = k; a.mk = mk;
    a.bt = c->d_bt; a.dt = c->d_dt;
    a.U = (const double2 *)d_U;
    a.cranks = d_child_ranks; a.nc = n_child;
    a.pranks = d_parent_ranks; a.np = n_parent;
    a.parent = (const double2 *)d_parent;
    a.child = (double2 *)d_child;
    a.probs = d_probs;
    a.amps = (double2 *)d_amps;
    a.sum = d_sum;
    a.inv_in_fact = in_prodnfact > 0 ? 1.0 / in_prodnfact : 1.0;
    a.status = c->d_status;
    uint64_t g = (n_child + MSK_BLOCK - 1) / MSK_BLOCK;
    const uint64_t cap = (uint64_t)c->sm_count * 8;
    if (g > cap) g = cap;
    const int mode = (d_child ? 1 : 0) | (d_probs ? 2 : 0) | (d_amps ? 4 : 0);
    switch (mode) {
        case 1: slos_masked_layer_kernel<1><<<(unsigned)g, MSK_BLOCK, 0, st>>>(a); break;
        case 2: slos_masked_layer_kernel<2><<<(unsigned)g, MSK_BLOCK, 0, st>>>(a); break;
        case 3: slos_masked_layer_kernel<3><<<(unsigned)g, MSK_BLOCK, 0, st>>>(a); break;
        case 4: slos_masked_layer_kernel<4><<<(unsigned)g, MSK_BLOCK, 0, st>>>(a); break;
        case 5: slos_masked_layer_kernel<5><<<(unsigned)g, MSK_BLOCK, 0, st>>>(a); break;
        case 6: slos_masked_layer_kernel<6><<<(unsigned)g, MSK_BLOCK, 0, st>>>(a); break;
        default: slos_masked_layer_kernel<7><<<(unsigned)g, MSK_BLOCK, 0, st>>>(a); break;
    }
    c->launches++;
    return fock_check_cuda(cudaGetLastError(), "slos_masked_layer_kernel");
}
