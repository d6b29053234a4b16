// slos_mu.cu -- cached tail occupation tables shared by the SLOS tile kernels (slos.cu, slos_thin.cu).
//
// For a tail width D and a tail photon count u, entry t of the table is the occupation of the D tail modes of tail rank t
// in FS(D, u), 4 bits per mode (mode i in bits [4i, 4i+4)).  The tables depend on (D, u) only -- not on the layer, the
// unitary or the prefix -- so they are built once per context and serve every layer, unitary and call: a thread of a tile
// kernel reads 8 bytes instead of un-ranking its tail by search (8 B per tail rank; 243 MB for D = 16, u <= 12).
// FSArray order: reference perceval/utils/states.py:255-298 (xq.FSArray iteration), tests/utils/test_statevector.py:430-438.
#include <mutex>

#include "slos_tile.cuh"

struct MuState {
    std::mutex lock;
    uint64_t *tab[24][FOCK_TMAX];
    size_t bytes;
};

__global__ void __launch_bounds__(256) mu_tuple_kernel(int D, int u, uint32_t S, const uint64_t *__restrict__ bt, uint64_t *__restrict__ out) {
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    if (t >= S) return;
    uint64_t rem = t, tup = 0;
    int Tprev = u;
    for (int i = 0; i < D; ++i) {
        int T = 0;
        if (i < D - 1) {
            const uint64_t *row = bt + (D - 1 - i) * FOCK_TMAX;
            T = Tprev;
            while (__ldg(row + T) > rem) --T;
            rem -= __ldg(row + T);
        }
        tup |= (uint64_t)(Tprev - T) << (4 * i);
        Tprev = T;
    }
    out[t] = tup;
}

void slos_mu_init(fock_ctx *c) {
    MuState *ms = new MuState();
    memset(ms->tab, 0, sizeof ms->tab);
    ms->bytes = 0;
    c->mu_state = ms;
}

// Thread-safe; a table is complete (stream synchronised once, at build time) before its pointer is handed out, so a
// launch on any other stream may read it.
int slos_mu_tuples(fock_ctx *c, int D, int u, uint32_t S, cudaStream_t st, const uint64_t **out) {
    MuState *ms = (MuState *)c->mu_state;
    FOCK_REQUIRE(ms != nullptr, FOCK_ERR_ARG, "slos_mu: context not initialised");
    FOCK_REQUIRE(D <= 16 && u < 16, FOCK_ERR_LIMIT, "slos_mu: occupation tuples hold 16 modes of <= 15 photons");
    std::lock_guard<std::mutex> g(ms->lock);
    if (!ms->tab[D][u]) {
        uint64_t *d = nullptr;
        FOCK_CUDA(cudaMalloc(&d, (size_t)S * 8));
        mu_tuple_kernel<<<(S + 255) / 256, 256, 0, st>>>(D, u, S, c->d_bt, d);
        FOCK_CUDA(cudaGetLastError());
        FOCK_CUDA(cudaStreamSynchronize(st));
        c->launches++;
        ms->tab[D][u] = d;
        ms->bytes += (size_t)S * 8;
    }
    *out = ms->tab[D][u];
    return FOCK_OK;
}

size_t slos_mu_bytes(fock_ctx *c) {
    MuState *ms = (MuState *)c->mu_state;
    return ms ? ms->bytes : 0;
}

void slos_mu_destroy(fock_ctx *c) {
    MuState *ms = (MuState *)c->mu_state;
    if (!ms) return;
    for (int d = 0; d < 24; ++d)
        for (int u = 0; u < FOCK_TMAX; ++u)
            if (ms->tab[d][u]) cudaFree(ms->tab[d][u]);
    delete ms;
    c->mu_state = nullptr;
}
