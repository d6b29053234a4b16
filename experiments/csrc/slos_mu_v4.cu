// slos_mu.cu -- (1) cached tail occupation tables shared by every SLOS tile kernel, (2) the v4 warp-specialised tile kernel.
//
// (1) mu_tuple_kernel / slos_mu_tuples: for a tail width D and a tail photon count u, entry t of the table is the
//     occupation of the D tail modes of tail rank t in FS(D, u), 4 bits per mode.  The tables depend on (D, u) only, so
//     they are built once per context and serve every layer, unitary and call; the tile kernels (slos.cu v2 -- the
//     default --, v4 here, v5 in slos_thin.cu) read 8 bytes per thread instead of un-ranking their tail by search.
// (2) slos_mu_kernel (FOCK_SLOS_KERNEL=v4, parity-tested, not the default): same tiling, prefix sweep and rounding
//     sequence as slos_tile_kernel (replaces FSMap.compute_slos_layer, reference call site perceval/backends/_slos.py:99,
//     python twin :91-97, and xq.all_prob_normalize_output, _slos.py:199,213); the sweep is specialised per warp on the
//     number of tail slots it executes, the unitary entry of a slot is one LDS at a pre-computed shared address, and
//     unused slots are neutralised by a zero coefficient instead of predicates (28 % fewer instructions, same speed:
//     profiles/README.md).  FOCK_MU_DEBUG switches off parts of its memory traffic for timing experiments only.
#include <stdlib.h>

#include "slos_tile.cuh"

#define MU_DB 128   // prefix descriptors per batch

struct __align__(16) MuDesc {
    uint64_t cbase;
    const char *tptr;   // byte address of the tail-parent block of this prefix
    double pfact;
    int nz, pad;
};

__device__ __forceinline__ double mu_factorial(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f *= (double)i;
    return f;
}

__device__ __forceinline__ double2 mu_ld_row(const char *base, uint32_t off) {
    return __ldg((const double2 *)(base + off));
}
__device__ __forceinline__ double2 mu_ld_tail(const char *base, uint32_t off) { return __ldg((const double2 *)(base + off)); }

__device__ __forceinline__ double2 mu_lds16(uint32_t saddr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void mu_bar() { asm volatile("bar.sync 0;" ::: "memory"); }

struct MuShared {
    double2 *s_u;
    MuDesc *s_desc;
    double2 *e_u;
    uint64_t *e_ptr;
};

// The prefix sweep of one thread, specialised on the number W of tail slots its WARP executes (W >= the largest number of
// occupied tail modes among the 32 lanes, rounded up to the next instantiated width).  Slots a lane does not need carry a
// zero coefficient and offset 0, so the loop body has no per-lane predicates: per slot it is 2 integer adds, one LDG.128,
// one LDS.128 (the lane's unitary entry) and 4 DFMAs.  Every load of a child -- up to 4 prefix rows and the first FIRST tail
// slots -- is issued before the first FMA, so a warp pays one memory round trip per child.
template <int D, int W, int MODE, bool RANGECHK>
__device__ __forceinline__ void mu_sweep(const TileArgs &a, const MuShared &sh, const uint32_t (&toffb)[D], const uint32_t (&uadr)[(D + 1) / 2],
                                         const uint32_t t, const int u, const int w, const uint64_t rho_a, const uint64_t rho_b,
                                         const double tfact, const bool active, double &local_sum) {
    constexpr int FIRST = W < 8 ? W : 8;
    constexpr int URES = 0;   // unitary entries of the first URES slots stay in registers for the whole sweep
    double2 ures[URES > 0 ? URES : 1];
#pragma unroll
    for (int c = 0; c < URES; ++c) ures[c] = mu_lds16((c & 1) ? (uadr[c / 2] >> 16) : (uadr[c / 2] & 0xFFFFu));
    const int m = a.m, p = a.p, maxnz = a.maxnz;
    const int tid = threadIdx.x;
    const uint64_t *__restrict__ bt = a.bt;
    const uint64_t *__restrict__ dt = a.dt;
    const char *__restrict__ parent_b = (const char *)a.parent;
    const uint32_t t16 = t << 4;
    for (uint64_t rho0 = rho_a; rho0 < rho_b; rho0 += MU_DB) {
        const int nb = (int)((rho_b - rho0) < (uint64_t)MU_DB ? (rho_b - rho0) : (uint64_t)MU_DB);
        mu_bar();
        if (tid < nb) {   // cooperative prefix descriptors: thread i un-ranks prefix rho0 + i of FS(p, w)
            uint64_t rem = rho0 + tid;
            int Tprev = w;
            uint64_t base = 0, E = 0;
            int nz = 0;
            double pf = 1.0;
            for (int i = 0; i < p; ++i) {
                int T = 0;
                if (i < p - 1) {
                    const uint64_t *row = bt + (p - 1 - i) * FOCK_TMAX;
                    T = Tprev;
                    while (__ldg(row + T) > rem) --T;
                    rem -= __ldg(row + T);
                }
                const int si = Tprev - T;
                const int Tfull = T + u;
                if (si > 0) {
                    sh.e_ptr[tid * maxnz + nz] = E;
                    sh.e_u[tid * maxnz + nz] = sh.s_u[i];
                    ++nz;
                    pf *= mu_factorial(si);
                }
                base += __ldg(bt + (m - 1 - i) * FOCK_TMAX + Tfull);
                if (Tfull > 0) E += __ldg(dt + (m - 1 - i) * FOCK_TMAX + Tfull);
                Tprev = T;
            }
            for (int e = 0; e < nz; ++e) sh.e_ptr[tid * maxnz + e] = (uint64_t)(parent_b + ((base - sh.e_ptr[tid * maxnz + e]) << 4));
            MuDesc td;
            td.cbase = base;
            td.tptr = parent_b + ((base - E) << 4);
            td.pfact = pf;
            td.nz = nz;
            td.pad = 0;
            sh.s_desc[tid] = td;
        }
        mu_bar();
        if (!active) continue;
        for (int i = 0; i < nb; ++i) {
            const MuDesc td = sh.s_desc[i];
            const uint64_t r = td.cbase + t;
            if (RANGECHK && (r < a.cbegin || r >= a.cend)) continue;
            const uint64_t *ep = sh.e_ptr + i * maxnz;
            const double2 *pu = sh.e_u + i * maxnz;
            const int nz = (a.nslots & 1) ? 0 : td.nz;
            double2 pv0, pv1, pv2, pv3, tv[FIRST > 0 ? FIRST : 1];
            if (nz > 0) pv0 = mu_ld_row((const char *)ep[0], t16);
            if (nz > 1) pv1 = mu_ld_row((const char *)ep[1], t16);
            if (nz > 2) pv2 = mu_ld_row((const char *)ep[2], t16);
            if (nz > 3) pv3 = mu_ld_row((const char *)ep[3], t16);
            const char *tptr = (a.nslots & 2) ? (const char *)a.parent : td.tptr;
#pragma unroll
            for (int c = 0; c < FIRST; ++c) tv[c] = mu_ld_tail(tptr, (a.nslots & 2) ? 0u : toffb[c]);
            // two independent accumulators (even / odd edges): a lone chain of dependent DFMAs stalls the warp between issues
            double2 acc = make_double2(0.0, 0.0), acb = make_double2(0.0, 0.0);
            if (nz > 0) acc = cfma(pu[0], pv0, acc);
            if (nz > 1) acb = cfma(pu[1], pv1, acb);
            if (nz > 2) acc = cfma(pu[2], pv2, acc);
            if (nz > 3) acb = cfma(pu[3], pv3, acb);
            for (int e = 4; e < nz; ++e) acc = cfma(pu[e], mu_ld_row((const char *)ep[e], t16), acc);
#pragma unroll
            for (int c = 0; c < FIRST; ++c) {
                const double2 uc = c < URES ? ures[c] : ((a.nslots & 4) ? make_double2(0.5, 0.25) : mu_lds16((c & 1) ? (uadr[c / 2] >> 16) : (uadr[c / 2] & 0xFFFFu)));
                if (c & 1) acb = cfma(uc, tv[c], acb);
                else acc = cfma(uc, tv[c], acc);
            }
            if (W > FIRST) {
                double2 tw[W > FIRST ? W - FIRST : 1];
#pragma unroll
                for (int c = FIRST; c < W; ++c) tw[c - FIRST] = mu_ld_tail(tptr, (a.nslots & 2) ? 0u : toffb[c]);
#pragma unroll
                for (int c = FIRST; c < W; ++c) {
                    const double2 uc = mu_lds16((c & 1) ? (uadr[c / 2] >> 16) : (uadr[c / 2] & 0xFFFFu));
                    if (c & 1) acb = cfma(uc, tw[c - FIRST], acb);
                    else acc = cfma(uc, tw[c - FIRST], acc);
                }
            }
            acc.x += acb.x;
            acc.y += acb.y;
            if (MODE & 1) a.child[r - a.cbegin] = acc;
            if (MODE & 2) {
                const double pr = (acc.x * acc.x + acc.y * acc.y) * a.inv_in_fact * (td.pfact * tfact);
                __stcs(a.probs + (r - a.cbegin), pr);
                local_sum += pr;
            }
        }
    }
}

// v4 tile kernel.  Tiling, prefix sweep and rounding sequence (one accumulator, modes in ascending order) are those of
// slos_tile_kernel (v2); only classes whose tail block holds >= 256 states (G == 1) are handled here, the few small ones
// go to v2.  What changed: the sweep is specialised per warp on the tail-slot count (mu_sweep), the unitary entry of a
// slot is one LDS at a pre-computed shared address, and unused slots are neutralised by a zero coefficient instead of
// predicates -- 5.2e9 -> see profiles/README.md for the instruction counts.
template <int D, int MODE, bool RANGECHK, int MINB>
__global__ void __launch_bounds__(TILE_BLOCK, MINB) slos_mu_kernel(const __grid_constant__ TileArgs a) {
    extern __shared__ __align__(16) unsigned char mu_smem[];
    const int m = a.m, p = a.p, maxnz = a.maxnz;
    const int tid = threadIdx.x;
    MuShared sh;
    sh.s_u = (double2 *)mu_smem;             // [m] column mk of U, then one zero entry
    sh.s_desc = (MuDesc *)(sh.s_u + m + 1);
    sh.e_u = (double2 *)(sh.s_desc + MU_DB);
    sh.e_ptr = (uint64_t *)(sh.e_u + MU_DB * maxnz);
    __shared__ double s_red[TILE_BLOCK / 32];
    const uint64_t *__restrict__ bt = a.bt;
    const uint64_t *__restrict__ dt = a.dt;

    for (int i = tid; i < m; i += TILE_BLOCK) sh.s_u[i] = a.U[(size_t)i * m + a.mk];
    if (tid == 0) sh.s_u[m] = make_double2(0.0, 0.0);
    int ci = 0;
    for (int c = 1; c < a.ncls; ++c)
        if ((uint64_t)blockIdx.x >= a.cls[c].item_begin) ci = c;
    const int w = a.cls[ci].w, u = a.cls[ci].u;
    const uint32_t S = a.cls[ci].S, nchunks = a.cls[ci].nchunks;
    const uint64_t local = (uint64_t)blockIdx.x - a.cls[ci].item_begin;
    const uint32_t chunk = (uint32_t)(local % nchunks);
    const uint64_t range = local / nchunks;
    const uint64_t rho_a = a.cls[ci].rho_lo + range * a.cls[ci].per_item;
    uint64_t rho_b = rho_a + a.cls[ci].per_item;
    if (rho_b > a.cls[ci].rho_lo + a.cls[ci].np) rho_b = a.cls[ci].rho_lo + a.cls[ci].np;
    __syncthreads();

    // ---- per-thread tail set-up.  The occupation tuple of tail rank t comes from a table cached per (D, u) (built once by
    // mu_tuple_kernel), so there is no un-ranking search here: T_i and E_i are running sums, Dt look-ups are independent
    // L1 hits, and the (byte offset, shared address of the U entry) pairs of the occupied modes are compacted through a
    // private shared-memory column instead of D*D predicated moves.
    const uint32_t t = chunk * TILE_BLOCK + tid;
    const bool active = t < S;
    const uint32_t s_zero = (uint32_t)__cvta_generic_to_shared(sh.s_u + m);
    const uint32_t s_tail0 = (uint32_t)__cvta_generic_to_shared(sh.s_u + p);
    uint2 *s_col = (uint2 *)(sh.e_ptr + MU_DB * maxnz) + tid;   // [D][TILE_BLOCK] : entry c of this thread at s_col[c * TILE_BLOCK]
    uint32_t toffb[D];            // 16 * local rank of (tau - e_mode) in FS(D, u-1); 0 for unused entries
    uint32_t uadr[(D + 1) / 2];   // shared-window address of the U entry of slot c (16 bits each); unused -> zero entry
    int cnt = 0;
    double tfact = 1.0;
    if (active) {
        const uint64_t tup = __ldg(a.tup[ci] + t);
        uint32_t E = 0;
        int T = u;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const int si = (int)((tup >> (4 * i)) & 15u);
            T -= si;
            if (si > 0) {
                s_col[cnt * TILE_BLOCK] = make_uint2((t - E) << 4, s_tail0 + 16u * (uint32_t)i);
                ++cnt;
                if (si > 1) tfact *= mu_factorial(si);
            }
            if (i < D - 1 && T > 0) E += (uint32_t)__ldg(dt + (D - 1 - i) * FOCK_TMAX + T);
        }
    }
#pragma unroll
    for (int c = 0; c < (D + 1) / 2; ++c) uadr[c] = 0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        uint2 v = make_uint2(0u, s_zero);
        if (c < cnt) v = s_col[c * TILE_BLOCK];
        toffb[c] = v.x;
        uadr[c / 2] |= (c & 1) ? (v.y << 16) : v.y;
    }
    const int wcnt = __reduce_max_sync(0xffffffffu, cnt);   // warp-uniform number of tail slots
    double local_sum = 0.0;
#define MU_SWEEP(WW) mu_sweep<D, (WW) < D ? (WW) : D, MODE, RANGECHK>(a, sh, toffb, uadr, t, u, w, rho_a, rho_b, tfact, active, local_sum)
    if (wcnt <= 4) MU_SWEEP(4);
    else if (wcnt <= 6) MU_SWEEP(6);
    else if (wcnt <= 8) MU_SWEEP(8);
    else if (wcnt <= 10) MU_SWEEP(10);
    else if (wcnt <= 12) MU_SWEEP(12);
    else MU_SWEEP(D);
#undef MU_SWEEP
    if ((MODE & 2) && a.sum) {
        local_sum = warp_sum(local_sum);
        if ((tid & 31) == 0) s_red[tid >> 5] = local_sum;
        __syncthreads();
        if (tid < 32) {
            double v = tid < TILE_BLOCK / 32 ? s_red[tid] : 0.0;
            v = warp_sum(v);
            if (tid == 0) atomicAdd(a.sum, v);
        }
    }
}

// ---------------------------------------------------------------- host side
static int mu_env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int D, int MODE, bool CHK, int MINB>
static int mu_launch1(const TileArgs &a, unsigned grid, size_t smem, cudaStream_t st) {
    FOCK_CUDA(cudaFuncSetAttribute(slos_mu_kernel<D, MODE, CHK, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    slos_mu_kernel<D, MODE, CHK, MINB><<<grid, TILE_BLOCK, smem, st>>>(a);
    return fock_check_cuda(cudaGetLastError(), "slos_mu_kernel");
}

template <int D, int MINB>
static int mu_launch(const TileArgs &a, bool wc, bool wp, bool chk, unsigned grid, size_t smem, cudaStream_t st) {
    if (wp && wc) return chk ? mu_launch1<D, 3, true, MINB>(a, grid, smem, st) : mu_launch1<D, 3, false, MINB>(a, grid, smem, st);
    if (wp) return chk ? mu_launch1<D, 2, true, MINB>(a, grid, smem, st) : mu_launch1<D, 2, false, MINB>(a, grid, smem, st);
    return chk ? mu_launch1<D, 1, true, MINB>(a, grid, smem, st) : mu_launch1<D, 1, false, MINB>(a, grid, smem, st);
}


// ---- cached tail occupation tables: for (D, u), entry t holds the occupation of the D tail modes of rank t in FS(D, u),
// 4 bits per mode (mode i in bits [4i, 4i+4)).  They depend on (D, u) only -- not on the layer, the unitary or the
// prefix -- so a chain of layers, and every later call, re-uses them (8 B per tail rank; 243 MB for D = 16, u <= 12).
struct MuState {
    uint64_t *tab[24][FOCK_TMAX];
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

int slos_mu_tuples(fock_ctx *c, int D, int u, uint32_t S, cudaStream_t st, const uint64_t **out) {
    MuState *ms = (MuState *)c->mu_state;
    if (!ms) {
        ms = new MuState();
        memset(ms, 0, sizeof *ms);
        c->mu_state = ms;
    }
    FOCK_REQUIRE(D <= 16 && u < 16, FOCK_ERR_LIMIT, "slos_mu: occupation tuples hold 16 modes of <= 15 photons");
    if (!ms->tab[D][u]) {
        uint64_t *d = nullptr;
        FOCK_CUDA(cudaMalloc(&d, (size_t)S * 8));
        mu_tuple_kernel<<<(S + 255) / 256, 256, 0, st>>>(D, u, S, c->d_bt, d);
        FOCK_CUDA(cudaGetLastError());
        c->launches++;
        ms->tab[D][u] = d;
    }
    *out = ms->tab[D][u];
    return FOCK_OK;
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

// tail widths the v4 kernel is instantiated for
bool slos_mu_supports(int D, int k) { return (D == 8 || D == 12 || D == 16) && k <= 15; }   // tuples: 4 bits / mode

// `a` is a finished work plan of slos_layer_tiles (slos.cu) holding only classes with G == 1; the whole parent layer must
// be resident and 16-byte aligned.
int slos_mu_launch(fock_ctx *c, int D, TileArgs &a, bool want_child, bool want_probs, bool rangechk, unsigned grid, cudaStream_t st) {
    const size_t smem = (size_t)(a.m + 1) * 16 + (size_t)MU_DB * sizeof(MuDesc) + (size_t)MU_DB * a.maxnz * 24 + (size_t)D * TILE_BLOCK * 8 + 16;
    for (int i = 0; i < a.ncls; ++i)
        if (int rc = slos_mu_tuples(c, D, a.cls[i].u, a.cls[i].S, st, &a.tup[i])) return rc;
    static int dbg = -1;
    if (dbg < 0) dbg = mu_env_int("FOCK_MU_DEBUG", 0);   // timing experiments only: 1 = no prefix rows, 2 = tail loads hit one line
    a.nslots = dbg;
    int rc;
    switch (D) {
        case 8: rc = mu_launch<8, 2>(a, want_child, want_probs, rangechk, grid, smem, st); break;
        case 12: rc = mu_launch<12, 2>(a, want_child, want_probs, rangechk, grid, smem, st); break;
        case 16: rc = mu_launch<16, 2>(a, want_child, want_probs, rangechk, grid, smem, st); break;
        default:
            fock_set_error("slos_mu: tail width %d not instantiated", D);
            return FOCK_ERR_LIMIT;
    }
    if (rc) return rc;
    c->launches++;
    return FOCK_OK;
}
