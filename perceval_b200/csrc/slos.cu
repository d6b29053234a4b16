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
#include <stdlib.h>

#include "common.cuh"
#include "slos_tile.cuh"

#define SLOS_BLOCK 256
#define SLOS_SUB0_MIN (1ull << 18)   // smallest weight-0 tile that goes to a sub-layer call (slos_layer_impl)

struct SlosArgs {
    int m, k, mk;
    const uint64_t *bt, *dt;
    const double2 *U;        // m*m row-major
    const double2 *parent;   // parent ranks [pbegin, pend) minus the hole [gap_b, gap_e), stored packed
    uint64_t pbegin, pend, gap_b, gap_e;
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
                if (pr < a.pbegin || pr >= a.pend || (pr >= a.gap_b && pr < a.gap_e)) oob = true;
                else acc = cfma(s_u[i], parent[pr >= a.gap_e ? pr - (a.gap_e - a.gap_b) : pr], acc);
                if (MODE & 2) fact *= s_fact[si];
            }
            if (T == 0) break;  // every remaining mode is empty
            E += s_dt[q * FOCK_TMAX + T];
            Tprev = T;
        }
        if (i == m - 1 && Tprev > 0) {  // last mode holds the remaining photons
            const uint64_t pr = r - E;
            if (pr < a.pbegin || pr >= a.pend || (pr >= a.gap_b && pr < a.gap_e)) oob = true;
            else acc = cfma(s_u[m - 1], parent[pr >= a.gap_e ? pr - (a.gap_e - a.gap_b) : pr], acc);
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


// ================================================================================================================
// Tile kernel (v2): prefix-uniform formulation.
//
// Split the m modes into a PREFIX (first p = m - D modes) and a TAIL (last D modes).  All children that share the prefix
// pi (weight w) form one contiguous rank range ("tile") [base(pi), base(pi) + S), S = C(u + D - 1, u), u = k - w, in
// which the local index t is the rank of the tail inside FS(D, u).  For every child of the tile
//     parent rank for a prefix mode j (pi_j > 0)  =  (base - E_j) + t          -> aligned, perfectly coalesced row
//     parent rank for a tail mode   j (tau_j > 0) =  (base - E_p) + (t - tailE_j(tau))   -> gather inside one small block
// with E_j = sum_{i<j} Dt[m-1-i][T_i] depending on the prefix only, and tailE_j depending on (u, t) only.
// A CTA therefore fixes (w, a chunk of 256 tail indices) once -- each thread un-ranks ITS tail a single time and keeps
// the <= D tail parent offsets and U entries in registers -- and then sweeps hundreds of prefixes: per prefix the only
// per-thread work is the loads, the complex FMAs and the store.  Prefix descriptors (bases, non-zero modes, prod pi_i!)
// are un-ranked cooperatively, one prefix per thread, into shared memory and read back as broadcasts.
// Classes with S < 256 pack G = 256 / S prefixes per sweep step so the lanes stay busy.
// ================================================================================================================

__device__ __forceinline__ double c_factorial(int n) {   // exact for n <= 22, correctly rounded products beyond
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f *= (double)i;
    return f;
}
#define TILE_DB 128   // descriptors per batch (256 -> one CTA per SM: 18.8 ms vs 15.0 ms for the last 12/24 layer)

// window-checked parent load (CHECK == 2): false if rank r is not resident
__device__ __forceinline__ bool tile_win_load(const TileArgs &a, const double2 *__restrict__ parent, uint64_t r, double2 &v) {
    if (r < a.pbegin || r >= a.pend || (r >= a.gap_b && r < a.gap_e)) {
        v = make_double2(0.0, 0.0);
        return false;
    }
    v = parent[r >= a.gap_e ? r - (a.gap_e - a.gap_b) : r];
    return true;
}

#ifndef TILE_MINB
#define TILE_MINB 2
#endif
#define TILE_VP 4   // prefix edges whose loads are issued together with the tail loads
#define TILE_TB 8   // tail parents loaded per batch

// CHECK: 0 = whole layers, 1 = child range only (sharded child, resident parent layer), 2 = child range + parent window
template <int D, int MODE, int CHECK>
__global__ void __launch_bounds__(TILE_BLOCK, TILE_MINB) slos_tile_kernel(const __grid_constant__ TileArgs a) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    const int m = a.m, p = a.p, maxnz = a.maxnz;
    const int tid = threadIdx.x;
    uint64_t *s_bt = (uint64_t *)tile_smem;
    uint64_t *s_dt = s_bt + m * FOCK_TMAX;
    double2 *s_u = (double2 *)(s_dt + m * FOCK_TMAX);
    double *s_fact = (double *)(s_u + m);
    TileDesc *s_desc = (TileDesc *)(s_fact + 34);
    double2 *e_u = (double2 *)(s_desc + TILE_DB);
    uint64_t *e_pb = (uint64_t *)(e_u + TILE_DB * maxnz);
    __shared__ double s_red[TILE_BLOCK / 32];

    for (int i = tid; i < m * FOCK_TMAX; i += TILE_BLOCK) {
        s_bt[i] = a.bt[i];
        s_dt[i] = a.dt[i];
    }
    for (int i = tid; i < m; i += TILE_BLOCK) s_u[i] = a.U[(size_t)(a.urow0 + i) * a.ustride + a.mk];
    if (tid == 0) {
        double f = 1.0;
        s_fact[0] = 1.0;
        for (int i = 1; i < FOCK_TMAX; ++i) { f *= (double)i; s_fact[i] = f; }
    }
    // ---- which class / work item
    int ci = 0;
    for (int c = 1; c < a.ncls; ++c)
        if ((uint64_t)blockIdx.x >= a.cls[c].item_begin) ci = c;
    const int w = a.cls[ci].w, u = a.cls[ci].u;
    const uint32_t S = a.cls[ci].S, G = a.cls[ci].G, nchunks = a.cls[ci].nchunks;
    const uint64_t local = (uint64_t)blockIdx.x - a.cls[ci].item_begin;
    const uint32_t chunk = (uint32_t)(local % nchunks);
    const uint64_t range = local / nchunks;
    const uint64_t rho_a = a.cls[ci].rho_lo + range * a.cls[ci].per_item;
    uint64_t rho_b = rho_a + a.cls[ci].per_item;
    if (rho_b > a.cls[ci].rho_lo + a.cls[ci].np) rho_b = a.cls[ci].rho_lo + a.cls[ci].np;
    __syncthreads();

    // ---- per-thread tail: un-rank t in FS(D, u) once; keep (mode, parent offset) of the non-zero tail modes, compacted
    uint32_t g, t;
    if (G > 1) { g = tid / S; t = tid - g * S; } else { g = 0; t = chunk * TILE_BLOCK + tid; }
    const bool active = (g < G) && (t < S);
    uint32_t toff[D];            // local rank of (tau - e_mode) in FS(D, u-1) for the c-th occupied tail mode
    uint32_t tmode[(D + 5) / 6]; // the tail mode of entry c, 5 bits each, 6 per word
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < (D + 5) / 6; ++c) tmode[c] = 0;
    double tfact = 1.0;
#pragma unroll
    for (int c = 0; c < D; ++c) toff[c] = 0;
    const uint64_t *__restrict__ tuples = a.tup[ci];
    if (active && tuples != nullptr) {
        // fast set-up: the occupation tuple of tail rank t comes from the table cached per (D, u) (slos_mu.cu), so T_i / E_i
        // are running sums with independent table look-ups, and the (offset, mode) pairs of the occupied modes are
        // compacted through a private shared-memory column instead of D*D predicated moves.  The set-up is not
        // amortised for the classes with few prefixes (w <= 3 holds 40 % of the children at 12 photons / 24 modes): the
        // searching un-rank below cost 27 % of the kernel there (profiles/README.md).
        uint2 *s_col = (uint2 *)(e_pb + TILE_DB * maxnz) + tid;   // entry c of this thread at s_col[c * TILE_BLOCK]
        const uint64_t tup = __ldg(tuples + t);
        uint32_t E = 0;
        int T = u;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const int si = (int)((tup >> (4 * (i & 15))) & 15u);
            T -= si;
            if (si > 0) {
                s_col[cnt * TILE_BLOCK] = make_uint2(t - E, (uint32_t)i);
                ++cnt;
                tfact *= s_fact[si];
            }
            if (i < D - 1 && T > 0) E += (uint32_t)s_dt[(D - 1 - i) * FOCK_TMAX + T];
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
            if (c < cnt) {
                const uint2 v = s_col[c * TILE_BLOCK];
                toff[c] = v.x;
                tmode[c / 6] |= v.y << (5 * (c % 6));
            }
        }
    } else if (active) {
        uint64_t rem = t;
        uint32_t E = 0;
        int Tprev = u;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            int T = 0;
            if (i < D - 1) {
                const uint64_t *row = s_bt + (D - 1 - i) * FOCK_TMAX;
                T = Tprev;
                while (row[T] > rem) --T;
                rem -= row[T];
            }
            const int si = Tprev - T;
            if (si > 0) {
                const uint32_t off = t - E;
#pragma unroll
                for (int c = 0; c < D; ++c)
                    if (c == cnt) { toff[c] = off; tmode[c / 6] |= (uint32_t)i << (5 * (c % 6)); }
                ++cnt;
                tfact *= s_fact[si];
            }
            if (i < D - 1 && T > 0) E += (uint32_t)s_dt[(D - 1 - i) * FOCK_TMAX + T];
            Tprev = T;
        }
    }
    const int wcnt = __reduce_max_sync(0xffffffffu, cnt);   // warp-uniform trip count of the tail phase

    const double2 *__restrict__ parent = a.parent - a.pbegin;
    const double2 *__restrict__ parent_t = parent + t;
    const double2 *s_ut = s_u + p;
    double local_sum = 0.0;
    bool oob = false;

    for (uint64_t rho0 = rho_a; rho0 < rho_b; rho0 += TILE_DB) {
        const int nb = (int)((rho_b - rho0) < (uint64_t)TILE_DB ? (rho_b - rho0) : (uint64_t)TILE_DB);
        __syncthreads();
        // ---- cooperative prefix descriptors: thread i un-ranks prefix rho0 + i of FS(p, w)
        if (tid < nb) {
            uint64_t rem = rho0 + tid;
            int Tprev = w;            // prefix photons still to place
            uint64_t base = 0, E = 0;
            int nz = 0;
            double pf = 1.0;
            for (int i = 0; i < p; ++i) {
                int T = 0;
                if (i < p - 1) {
                    const uint64_t *row = s_bt + (p - 1 - i) * FOCK_TMAX;
                    T = Tprev;
                    while (row[T] > rem) --T;
                    rem -= row[T];
                }
                const int si = Tprev - T;
                const int Tfull = T + u;   // photons right of mode i in the full state
                if (si > 0) {
                    e_pb[tid * maxnz + nz] = E;
                    e_u[tid * maxnz + nz] = s_u[i];
                    ++nz;
                    pf *= s_fact[si];
                }
                if (a.slab) {   // slab-major: E is the PREFIX-local rank difference (same identity on FS(p, w))
                    if (T > 0) E += s_dt[(p - 1 - i) * FOCK_TMAX + T];
                } else {
                    base += s_bt[(m - 1 - i) * FOCK_TMAX + Tfull];
                    if (Tfull > 0) E += s_dt[(m - 1 - i) * FOCK_TMAX + Tfull];
                }
                Tprev = T;
            }
            uint64_t tb;
            if (a.slab) {
                const uint64_t rho = rho0 + tid;
                base = a.cls[ci].coff + rho * S;
                for (int e = 0; e < nz; ++e) e_pb[tid * maxnz + e] = a.cls[ci].roff + (rho - e_pb[tid * maxnz + e]) * S;
                tb = a.cls[ci].toff + rho * a.cls[ci].Sp;
            } else {
                for (int e = 0; e < nz; ++e) e_pb[tid * maxnz + e] = base - e_pb[tid * maxnz + e];
                tb = base - E;   // only meaningful (and only used) when u >= 1
            }
            TileDesc td;
            td.cbase = base;
            td.tbase = tb;
            td.pfact = pf;
            td.nz = nz;
            td.pad = 0;
            s_desc[tid] = td;
        }
        __syncthreads();
        if (!active) continue;
        for (int i = (int)g; i < nb; i += (int)G) {
            const TileDesc td = s_desc[i];
            const uint64_t r = td.cbase + t;
            if (CHECK >= 1 && (r < a.cbegin || r >= a.cend)) continue;
            const uint64_t *pb = e_pb + i * maxnz;
            const double2 *pu = e_u + i * maxnz;
            const int nz = td.nz;
            // ---- phase 1: prefix rows + first batch of tail parents: every load is issued before any arithmetic.
            // Values of entries >= nz / >= wcnt stay undefined and are never consumed (the FMAs carry the same guards).
            const double2 *__restrict__ tbp = parent + td.tbase;   // tail-parent block of this tile
            double2 pv[TILE_VP], tv[TILE_TB];
#pragma unroll
            for (int e = 0; e < TILE_VP; ++e) {
                if (e < nz) {
                    if (CHECK == 2) { if (!tile_win_load(a, parent, pb[e] + t, pv[e])) oob = true; }
                    else pv[e] = parent_t[pb[e]];
                }
            }
#pragma unroll
            for (int c = 0; c < TILE_TB && c < D; ++c) {
                if (c < wcnt) {
                    if (CHECK == 2) { if (!tile_win_load(a, parent, td.tbase + toff[c], tv[c]) && c < cnt) oob = true; }
                    else tv[c] = tbp[toff[c]];
                }
            }
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int e = 0; e < TILE_VP; ++e)
                if (e < nz) acc = cfma(pu[e], pv[e], acc);
            for (int e = TILE_VP; e < nz; ++e) {
                if (CHECK == 2) {
                    double2 v;
                    if (!tile_win_load(a, parent, pb[e] + t, v)) { oob = true; continue; }
                    acc = cfma(pu[e], v, acc);
                    continue;
                }
                acc = cfma(pu[e], parent_t[pb[e]], acc);
            }
#pragma unroll
            for (int c = 0; c < TILE_TB && c < D; ++c)
                if (c < cnt) acc = cfma(s_ut[(tmode[c / 6] >> (5 * (c % 6))) & 31u], tv[c], acc);
            // ---- further tail batches (only when some lane of the warp has more than TILE_TB occupied tail modes)
#pragma unroll
            for (int c0 = TILE_TB; c0 < D; c0 += TILE_TB) {
                if (c0 < wcnt) {
#pragma unroll
                    for (int c = c0; c < c0 + TILE_TB && c < D; ++c) {
                        if (c < wcnt) {
                            if (CHECK == 2) { if (!tile_win_load(a, parent, td.tbase + toff[c], tv[c - c0]) && c < cnt) oob = true; }
                            else tv[c - c0] = tbp[toff[c]];
                        }
                    }
#pragma unroll
                    for (int c = c0; c < c0 + TILE_TB && c < D; ++c)
                        if (c < cnt) acc = cfma(s_ut[(tmode[c / 6] >> (5 * (c % 6))) & 31u], tv[c - c0], acc);
                }
            }
            if (MODE & 1) a.child[r - a.cbegin] = acc;
            if (MODE & 2) {
                const double pr = (acc.x * acc.x + acc.y * acc.y) * a.inv_in_fact * (td.pfact * tfact);
                __stcs(a.probs + (r - a.cbegin), pr);   // never re-read on the device: keep it out of L2's way
                local_sum += pr;
            }
        }
    }
    if (CHECK == 2 && oob && a.status) atomicExch(a.status, 1);
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
// The small launches of a layer (the weight-0 sub-layer, the classes with tiles < 256 states: 222 CTAs that run 0.3 ms) go to
// high-priority side streams of the context, forked from and joined to the caller's stream with per-call events, so that they
// run under the main launch instead of before / after it with most SMs idle.
struct SideLaunch {
    fock_ctx *c;
    cudaStream_t st;
    cudaEvent_t fork = nullptr;
    bool used[2] = {false, false};
    SideLaunch(fock_ctx *c_, cudaStream_t st_) : c(c_), st(st_) {}
    cudaStream_t stream(int i) {   // side stream i, ordered after everything enqueued on the caller's stream so far
        if (!c->side[i]) return st;   // created with the context (fock_create)
        if (!fork) {
            if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess) return st;
            cudaEventRecord(fork, st);
        }
        if (!used[i]) {
            cudaStreamWaitEvent(c->side[i], fork, 0);
            used[i] = true;
        }
        return c->side[i];
    }
    void join() {                  // the caller's stream continues after the side launches
        for (int i = 0; i < 2; ++i)
            if (used[i]) {
                cudaEvent_t done;
                if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) == cudaSuccess) {
                    cudaEventRecord(done, c->side[i]);
                    cudaStreamWaitEvent(st, done, 0);
                    cudaEventDestroy(done);   // released when it has completed
                } else {
                    cudaStreamSynchronize(c->side[i]);
                }
                used[i] = false;
            }
        if (fork) { cudaEventDestroy(fork); fork = nullptr; }
    }
    ~SideLaunch() { join(); }
};

bool slos_thin_supports(int D, int k);           // slos_thin.cu
int slos_thin_launch(fock_ctx *c, int D, TileArgs &a, bool want_child, bool want_probs, bool rangechk, unsigned grid, cudaStream_t st);
int slos_mu_tuples(fock_ctx *c, int D, int u, uint32_t S, cudaStream_t st, const uint64_t **out);                   // slos_mu.cu

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

// ---- host helpers for the tile kernel
uint64_t slos_host_prefix_base(int m, int p, int w, int u, uint64_t rho);
static uint64_t host_prefix_base(int m, int p, int w, int u, uint64_t rho) { return slos_host_prefix_base(m, p, w, u, rho); }
uint64_t slos_host_prefix_base(int m, int p, int w, int u, uint64_t rho) {
    // child rank of (prefix #rho of FS(p, w), tail |u,0,..,0>) in FS(m, w + u)
    const uint64_t *bt = fock_host_bt();
    uint64_t rem = rho, base = 0;
    int Tprev = w;
    for (int i = 0; i < p; ++i) {
        int T = 0;
        if (i < p - 1) {
            const uint64_t *row = bt + (p - 1 - i) * FOCK_TMAX;
            T = Tprev;
            while (row[T] > rem) --T;
            rem -= row[T];
        }
        base += bt[(m - 1 - i) * FOCK_TMAX + (T + u)];
        Tprev = T;
    }
    return base;
}

static int slos_tail_modes(int m) {
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("FOCK_SLOS_TAIL");
        forced = e ? atoi(e) : -1;
    }
    // tail width: wide enough that almost every child lives in a tile of >= 256 states
    int D = forced > 0 ? forced : (m >= 20 ? 16 : (m >= 16 ? 12 : (m >= 12 ? 8 : (m >= 8 ? 6 : (m >= 6 ? 4 : 0)))));   // D = 20 (m >= 26) measured 45 % slower than 16 at 13 photons / 26 modes
    if (D != 4 && D != 6 && D != 8 && D != 10 && D != 12 && D != 16 && D != 20) D = 0;
    if (D > m - 1) D = 0;
    return D;
}

template <int D>
static int launch_tile(fock_ctx *c, TileArgs &a, bool want_child, bool want_probs, int check, unsigned grid, size_t smem, cudaStream_t st) {
#define TILE_LAUNCH(MODE, CHK)                                                                                              \
    do {                                                                                                                  \
        FOCK_CUDA(cudaFuncSetAttribute(slos_tile_kernel<D, MODE, CHK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        slos_tile_kernel<D, MODE, CHK><<<grid, TILE_BLOCK, smem, st>>>(a);                                                \
    } while (0)
#define TILE_LAUNCH_CHK(MODE)                                                                                             \
    do {                                                                                                                  \
        if (check == 2) TILE_LAUNCH(MODE, 2);                                                                             \
        else if (check == 1) TILE_LAUNCH(MODE, 1);                                                                        \
        else TILE_LAUNCH(MODE, 0);                                                                                        \
    } while (0)
    if (want_probs && want_child) TILE_LAUNCH_CHK(3);
    else if (want_probs) TILE_LAUNCH_CHK(2);
    else TILE_LAUNCH_CHK(1);
#undef TILE_LAUNCH_CHK
#undef TILE_LAUNCH
    c->launches++;
    return fock_check_cuda(cudaGetLastError(), "slos_tile_kernel");
}

// slab-major layouts (slos_layer_slab): per prefix weight w the prefixes [rho[2w], rho[2w+1]) to compute, the element offset of
// parent slab w' (w' = 0..k-1) and of child slab w (w = 0..k); index(w, rho, t) = off[w] + rho * |FS(D, photons - w)| + t
struct SlabSpec {
    const uint64_t *rho, *parent_off, *child_off;
};

static int slos_layer_tiles(fock_ctx *c, int D, int m, int k, const double *d_U, int mk, const double *d_parent, uint64_t pb,
                            uint64_t pe, double *d_child, double *d_probs, double *d_sum, double in_prodnfact, uint64_t cb,
                            uint64_t ce, cudaStream_t st, int gfilter = 0, uint64_t gap_b = UINT64_MAX,
                            uint64_t gap_e = UINT64_MAX, const SlabSpec *slab = nullptr, bool skip_w0 = false, int ustride = 0,
                            int urow0 = 0) {
    // gfilter: 0 = every class (tile kernel), 1 = only classes whose tail block fills a CTA (S >= 256) in the hybrid thin
    //          kernel (slos_thin.cu), 2 = only the small classes (S < 256) in the tile kernel
    const int p = m - D;
    TileArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.k = k; a.mk = mk; a.p = p;
    a.maxnz = p < k ? p : k;
    if (a.maxnz < 1) a.maxnz = 1;
    a.bt = c->d_bt; a.dt = c->d_dt;
    a.U = (const double2 *)d_U;
    a.parent = (const double2 *)d_parent;
    a.pbegin = pb; a.pend = pe;
    const bool gapped = gap_b < gap_e && gap_b < pe;
    a.gap_b = gapped ? gap_b : pe;
    a.gap_e = gapped ? gap_e : pe;
    a.child = (double2 *)d_child;
    a.probs = d_probs;
    a.sum = d_sum;
    a.inv_in_fact = 1.0 / in_prodnfact;
    a.cbegin = cb; a.cend = ce;
    a.status = c->d_status;
    a.slab = slab ? 1 : 0;
    a.ustride = ustride > 0 ? ustride : m;
    a.urow0 = urow0;
    const bool full = slab != nullptr || (cb == 0 && ce == fock_count(m, k));
    uint64_t items = 0;
    int ncls = 0;
    // classes with small tail blocks first: their CTAs walk many prefixes with little work each and would otherwise run
    // alone at the end of the grid
    for (int w = k; w >= (skip_w0 ? 1 : 0); --w) {
        const int u = k - w;
        const uint64_t np_total = fock_count(p, w), S64 = fock_count(D, u);
        FOCK_REQUIRE(S64 < (1ull << 32), FOCK_ERR_LIMIT, "slos: tail block too large for the tile kernel");
        if ((gfilter == 1 && S64 < TILE_BLOCK) || (gfilter == 2 && S64 >= TILE_BLOCK)) continue;
        uint64_t lo = 0, hi = np_total;
        if (slab) {
            lo = slab->rho[2 * w];
            hi = slab->rho[2 * w + 1];
            FOCK_REQUIRE(lo <= hi && hi <= np_total, FOCK_ERR_ARG, "slos_layer_slab: bad prefix range for weight %d", w);
        } else if (!full) {
            // prefixes whose tile [base, base+S) intersects [cb, ce); base is increasing in rho
            uint64_t l = 0, h = np_total;
            while (l < h) { uint64_t mid = (l + h) / 2; if (host_prefix_base(m, p, w, u, mid) + S64 > cb) h = mid; else l = mid + 1; }
            lo = l;
            l = lo; h = np_total;
            while (l < h) { uint64_t mid = (l + h) / 2; if (host_prefix_base(m, p, w, u, mid) >= ce) h = mid; else l = mid + 1; }
            hi = l;
        }
        if (hi <= lo) continue;
        TileClass &tc = a.cls[ncls++];
        tc.w = w; tc.u = u;
        tc.S = (uint32_t)S64;
        tc.G = S64 >= TILE_BLOCK ? 1u : (uint32_t)(TILE_BLOCK / S64);
        tc.nchunks = S64 >= TILE_BLOCK ? (uint32_t)((S64 + TILE_BLOCK - 1) / TILE_BLOCK) : 1u;
        tc.rho_lo = lo; tc.np = hi - lo;
        if (slab) {
            tc.coff = slab->child_off[w];
            tc.roff = w >= 1 ? slab->parent_off[w - 1] : 0;
            tc.toff = u >= 1 ? slab->parent_off[w] : 0;
            tc.Sp = u >= 1 ? fock_count(D, u - 1) : 0;
        }
        // prefixes per CTA: 128 sweep steps for full tiles (512: 21.0 ms, 128: 20.5 ms for the 12/24 chain -- mid-size layers need the CTAs); packed small tiles (G prefixes per step) get short items so that
        // they spread over many CTAs instead of one CTA walking tens of thousands of prefixes
        tc.per_item = (uint64_t)(tc.G > 1 ? 32 : 128) * tc.G;
        tc.item_begin = items;
        items += ((tc.np + tc.per_item - 1) / tc.per_item) * tc.nchunks;
    }
    a.ncls = ncls;
    if (items == 0) return FOCK_OK;
    FOCK_REQUIRE(items < (1ull << 31), FOCK_ERR_LIMIT, "slos: too many work items");
    size_t smem = (size_t)2 * m * FOCK_TMAX * 8 + (size_t)m * 16 + 34 * 8 + (size_t)TILE_DB * sizeof(TileDesc) +
                  (size_t)TILE_DB * a.maxnz * 24 + 16;
    if (D <= 16 && k <= 15) {   // cached occupation tuples (4 bits / tail mode): see slos_mu.cu
        for (int i = 0; i < ncls; ++i)
            if (int rc = slos_mu_tuples(c, D, a.cls[i].u, a.cls[i].S, st, &a.tup[i])) return rc;
        smem += (size_t)D * TILE_BLOCK * 8;
    }
    const bool parent_whole = pb == 0 && pe == fock_count(m, k - 1) && !gapped;
    const int check = !parent_whole ? 2 : (full ? 0 : 1);
    const bool wc = d_child != nullptr, wp = d_probs != nullptr;
    if (gfilter == 1) return slos_thin_launch(c, D, a, wc, wp, !full, (unsigned)items, st);
    switch (D) {
        case 4: return launch_tile<4>(c, a, wc, wp, check, (unsigned)items, smem, st);
        case 6: return launch_tile<6>(c, a, wc, wp, check, (unsigned)items, smem, st);
        case 8: return launch_tile<8>(c, a, wc, wp, check, (unsigned)items, smem, st);
        case 10: return launch_tile<10>(c, a, wc, wp, check, (unsigned)items, smem, st);
        case 12: return launch_tile<12>(c, a, wc, wp, check, (unsigned)items, smem, st);
        case 16: return launch_tile<16>(c, a, wc, wp, check, (unsigned)items, smem, st);
        default: return launch_tile<20>(c, a, wc, wp, check, (unsigned)items, smem, st);
    }
}

static int slos_layer_impl(fock_ctx *c, int m, int k, const double *d_U, int mk, const double *d_parent, uint64_t pb,
                           uint64_t pe, double *d_child, double *d_probs, double *d_sum, double in_prodnfact, uint64_t cb,
                           uint64_t ce, void *stream, const char *who, uint64_t gap_b = UINT64_MAX, uint64_t gap_e = UINT64_MAX) {
    if (int rc = slos_check(who, c, m, k)) return rc;
    FOCK_REQUIRE(k >= 1, FOCK_ERR_ARG, "%s: child layer must hold >= 1 photon", who);
    FOCK_REQUIRE(mk >= 0 && mk < m, FOCK_ERR_ARG, "%s: input mode %d outside [0,%d)", who, mk, m);
    FOCK_REQUIRE(cb <= ce && ce <= fock_count(m, k), FOCK_ERR_ARG, "%s: bad child range", who);
    FOCK_REQUIRE(pb <= pe && pe <= fock_count(m, k - 1), FOCK_ERR_ARG, "%s: bad parent range", who);
    FOCK_REQUIRE(d_U && d_parent, FOCK_ERR_ARG, "%s: NULL device pointer", who);
    FOCK_REQUIRE(d_child || d_probs, FOCK_ERR_ARG, "%s: no output buffer", who);
    if (cb == ce) return FOCK_OK;
    ScopedDevice sd(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    // measurement hook (bench.py roofline): CUDA events of the caller around the probability-layer launches, on their stream
    struct EventBracket {
        cudaEvent_t e;
        cudaStream_t s;
        EventBracket(cudaEvent_t b, cudaEvent_t e_, cudaStream_t s_) : e(e_), s(s_) { if (b) cudaEventRecord(b, s_); }
        ~EventBracket() { if (e) cudaEventRecord(e, s); }
    } bracket(d_probs ? c->ev_begin : nullptr, d_probs ? c->ev_end : nullptr, st);
    // Policy.  Small layers / few modes: per-child gather kernel.  Otherwise the prefix/tail tile kernel; large probability
    // layers with at most 8 prefix modes and the whole parent resident run their full tiles in the hybrid thin kernel
    // (slos_thin.cu: last 12/24 layer 12.2 ms vs 13.5 ms) and only their small classes in the tile kernel.
    // FOCK_SLOS_KERNEL=gather|tile pins one kernel (parity tests run every kernel on the same inputs).
    static int force = -1;   // 0 auto, 1 gather, 2 tile
    if (force < 0) {
        const char *e = getenv("FOCK_SLOS_KERNEL");
        force = (e && !strcmp(e, "gather")) ? 1 : ((e && !strcmp(e, "tile")) ? 2 : 0);
    }
    if (force != 1 && (ce - cb) >= 32768) {
        const bool gapped = gap_b < gap_e && gap_b < pe;
        const bool parent_full = (pb == 0 && pe == fock_count(m, k - 1)) && !gapped;
        const int D = slos_tail_modes(m);
        // The slab of prefix weight 0 (every photon in the tail: ONE prefix, 2 % of the states at 12/24) is an independent
        // layer on the D tail modes.  A sweep over one prefix does not amortise the per-thread tail set-up (73 ps per state
        // against 12.7 ps elsewhere), so a whole layer hands that block to a sub-layer call with its own prefix/tail split
        // (same modes in the same order: results are bit-identical).
        const bool full = cb == 0 && ce == fock_count(m, k);
        const int Dsub = D > 0 ? slos_tail_modes(D) : 0;
        const bool sub0 = force == 0 && full && parent_full && Dsub > 0 && fock_count(D, k) >= SLOS_SUB0_MIN && ((uintptr_t)d_parent & 15) == 0;
        SideLaunch side(c, st);
        if (sub0) {
            const uint64_t S = fock_count(D, k), Sp = fock_count(D, k - 1), cbase = fock_count(m, k) - S, pbase = fock_count(m, k - 1) - Sp;
            if (int rc = slos_layer_tiles(c, Dsub, D, k, d_U, mk, d_parent + 2 * pbase, 0, Sp, d_child ? d_child + 2 * cbase : nullptr,
                                          d_probs ? d_probs + cbase : nullptr, d_sum, in_prodnfact, 0, S, side.stream(0), 0, UINT64_MAX,
                                          UINT64_MAX, nullptr, false, m, m - D)) return rc;
        }
        const bool thin = force == 0 && d_probs != nullptr && D == 16 && m - D <= 8 && (ce - cb) >= (1ull << 25);
        if (D > 0 && thin && parent_full && slos_thin_supports(D, k) && ((uintptr_t)d_parent & 15) == 0) {
            if (int rc = slos_layer_tiles(c, D, m, k, d_U, mk, d_parent, pb, pe, d_child, d_probs, d_sum, in_prodnfact, cb, ce, side.stream(1), 2,
                                          UINT64_MAX, UINT64_MAX, nullptr, sub0)) return rc;
            return slos_layer_tiles(c, D, m, k, d_U, mk, d_parent, pb, pe, d_child, d_probs, d_sum, in_prodnfact, cb, ce, st, 1,
                                    UINT64_MAX, UINT64_MAX, nullptr, sub0);
        }
        if (D > 0) return slos_layer_tiles(c, D, m, k, d_U, mk, d_parent, pb, pe, d_child, d_probs, d_sum, in_prodnfact, cb, ce, st, 0, gap_b, gap_e,
                                           nullptr, sub0);
    }
    SlosArgs a;
    a.m = m; a.k = k; a.mk = mk;
    a.bt = c->d_bt; a.dt = c->d_dt;
    a.U = (const double2 *)d_U;
    a.parent = (const double2 *)d_parent;
    a.pbegin = pb; a.pend = pe;
    a.gap_b = (gap_b < gap_e && gap_b < pe) ? gap_b : pe;
    a.gap_e = (gap_b < gap_e && gap_b < pe) ? gap_e : pe;
    a.child = (double2 *)d_child;
    a.probs = d_probs;
    a.sum = d_sum;
    a.inv_in_fact = 1.0 / in_prodnfact;
    a.cbegin = cb; a.cend = ce;
    a.status = c->d_status;
    unsigned grid = slos_grid(c, ce - cb);
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

// Segmented parent: the resident parent ranks are seg = {b0, e0, b1, e1} with e0 <= b1, stored packed (segment 0 then
// segment 1; b1 == e1 means one segment).  Used by the recompute-window partition of a layer chain (dist.py): the parents
// a contiguous child range needs through mode j are exactly one contiguous range, and their union over the modes is one
// or two ranges.
static int seg_check(const char *who, int m, int k, const uint64_t *seg) {
    FOCK_REQUIRE(seg != nullptr, FOCK_ERR_ARG, "%s: parent_seg is NULL", who);
    FOCK_REQUIRE(seg[0] <= seg[1] && seg[1] <= seg[2] && seg[2] <= seg[3] && seg[3] <= fock_count(m, k - 1), FOCK_ERR_ARG,
                 "%s: parent segments must be ordered and inside the parent layer", who);
    return FOCK_OK;
}

extern "C" int slos_layer_seg(fock_ctx *c, int m, int k, const double *d_U, int mk, const double *d_parent, const uint64_t *seg,
                              double *d_child, uint64_t cb, uint64_t ce, void *stream) {
    FOCK_REQUIRE(d_child != nullptr, FOCK_ERR_ARG, "slos_layer_seg: d_child is NULL");
    FOCK_REQUIRE(k >= 1, FOCK_ERR_ARG, "slos_layer_seg: child layer must hold >= 1 photon");
    if (int rc = seg_check("slos_layer_seg", m, k, seg)) return rc;
    const bool two = seg[2] < seg[3];
    return slos_layer_impl(c, m, k, d_U, mk, d_parent, seg[0], two ? seg[3] : seg[1], d_child, nullptr, nullptr, 1.0, cb, ce, stream,
                           "slos_layer_seg", two ? seg[1] : UINT64_MAX, two ? seg[2] : UINT64_MAX);
}

extern "C" int slos_layer_probs_seg(fock_ctx *c, int m, int k, const double *d_U, int mk, const double *d_parent, const uint64_t *seg,
                                    double *d_child, double *d_probs, double *d_sum, double in_prodnfact, uint64_t cb, uint64_t ce,
                                    void *stream) {
    FOCK_REQUIRE(d_probs != nullptr, FOCK_ERR_ARG, "slos_layer_probs_seg: d_probs is NULL");
    FOCK_REQUIRE(in_prodnfact > 0, FOCK_ERR_ARG, "slos_layer_probs_seg: in_prodnfact must be > 0");
    FOCK_REQUIRE(k >= 1, FOCK_ERR_ARG, "slos_layer_probs_seg: child layer must hold >= 1 photon");
    if (int rc = seg_check("slos_layer_probs_seg", m, k, seg)) return rc;
    const bool two = seg[2] < seg[3];
    return slos_layer_impl(c, m, k, d_U, mk, d_parent, seg[0], two ? seg[3] : seg[1], d_child, d_probs, d_sum, in_prodnfact, cb, ce,
                           stream, "slos_layer_probs_seg", two ? seg[1] : UINT64_MAX, two ? seg[2] : UINT64_MAX);
}

// One layer in SLAB-MAJOR layout (multi-GPU slab partition, perceval_b200/slab.py): parent and child are stored by prefix
// weight ("slab"), then prefix rank, then tail rank; only the prefixes [rho_lo, rho_hi) of every slab are computed.  The same
// gather as slos_layer (reference _slos.py:91-97), with the addresses of the aligned prefix rows and of the tail-parent block
// taken from the slab offsets instead of the FSArray rank.
extern "C" int slos_layer_slab(fock_ctx *c, int m, int k, int p, const double *d_U, int mk, const double *d_parent, double *d_child,
                               double *d_probs, double *d_sum, double in_prodnfact, const uint64_t *h_rho_ranges,
                               const uint64_t *h_parent_slab_off, const uint64_t *h_child_slab_off, void *stream) {
    if (int rc = slos_check("slos_layer_slab", c, m, k)) return rc;
    FOCK_REQUIRE(k >= 1, FOCK_ERR_ARG, "slos_layer_slab: child layer must hold >= 1 photon");
    FOCK_REQUIRE(mk >= 0 && mk < m, FOCK_ERR_ARG, "slos_layer_slab: input mode %d outside [0,%d)", mk, m);
    FOCK_REQUIRE(d_U && d_parent && (d_child || d_probs), FOCK_ERR_ARG, "slos_layer_slab: NULL device pointer");
    FOCK_REQUIRE(h_rho_ranges && h_parent_slab_off && h_child_slab_off, FOCK_ERR_ARG, "slos_layer_slab: NULL layout table");
    FOCK_REQUIRE(d_probs == nullptr || in_prodnfact > 0, FOCK_ERR_ARG, "slos_layer_slab: in_prodnfact must be > 0");
    const int D = slos_tail_modes(m);
    FOCK_REQUIRE(D > 0 && p == m - D, FOCK_ERR_ARG, "slos_layer_slab: the prefix of m = %d modes holds %d modes (got %d)", m, m - D, p);
    FOCK_REQUIRE(((uintptr_t)d_parent & 15) == 0, FOCK_ERR_ARG, "slos_layer_slab: parent buffer must be 16-byte aligned");
    ScopedDevice sd(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    struct EventBracket {
        cudaEvent_t e;
        cudaStream_t s;
        EventBracket(cudaEvent_t b, cudaEvent_t e_, cudaStream_t s_) : e(e_), s(s_) { if (b) cudaEventRecord(b, s_); }
        ~EventBracket() { if (e) cudaEventRecord(e, s); }
    } bracket(d_probs ? c->ev_begin : nullptr, d_probs ? c->ev_end : nullptr, st);
    SlabSpec spec{h_rho_ranges, h_parent_slab_off, h_child_slab_off};
    uint64_t children = 0;
    for (int w = 0; w <= k; ++w) children += (h_rho_ranges[2 * w + 1] - h_rho_ranges[2 * w]) * fock_count(D, k - w);
    if (children == 0) return FOCK_OK;
    const uint64_t np = fock_count(m, k - 1), nc = fock_count(m, k);
    // the weight-0 slab as a sub-layer on the tail modes (see slos_layer_impl)
    const int Dsub = slos_tail_modes(D);
    const bool sub0 = h_rho_ranges[1] > h_rho_ranges[0] && Dsub > 0 && fock_count(D, k) >= SLOS_SUB0_MIN;
    SideLaunch side(c, st);
    if (sub0) {
        const uint64_t S = fock_count(D, k), Sp = fock_count(D, k - 1);
        const bool only = children == S;
        if (int rc = slos_layer_tiles(c, Dsub, D, k, d_U, mk, d_parent + 2 * h_parent_slab_off[0], 0, Sp,
                                      d_child ? d_child + 2 * h_child_slab_off[0] : nullptr, d_probs ? d_probs + h_child_slab_off[0] : nullptr,
                                      d_sum, in_prodnfact, 0, S, only ? st : side.stream(0), 0, UINT64_MAX, UINT64_MAX, nullptr, false, m,
                                      m - D)) return rc;
        if (only) return FOCK_OK;
    }
    const bool thin = d_probs != nullptr && D == 16 && m - D <= 8 && children >= (1ull << 24) && slos_thin_supports(D, k);
    if (thin) {
        if (int rc = slos_layer_tiles(c, D, m, k, d_U, mk, d_parent, 0, np, d_child, d_probs, d_sum, in_prodnfact, 0, nc, side.stream(1), 2,
                                      UINT64_MAX, UINT64_MAX, &spec, sub0)) return rc;
        return slos_layer_tiles(c, D, m, k, d_U, mk, d_parent, 0, np, d_child, d_probs, d_sum, in_prodnfact, 0, nc, st, 1, UINT64_MAX,
                                UINT64_MAX, &spec, sub0);
    }
    return slos_layer_tiles(c, D, m, k, d_U, mk, d_parent, 0, np, d_child, d_probs, d_sum, in_prodnfact, 0, nc, st, 0, UINT64_MAX,
                            UINT64_MAX, &spec, sub0);
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
    if (n == 0) {   // one state (the vacuum), probability 1: device-to-device from the context's constant 1 + 0i
        FOCK_CUDA(cudaMemcpyAsync(d_probs, c->d_vacuum, sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (d_coefs) FOCK_CUDA(cudaMemcpyAsync(d_coefs, c->d_vacuum, 16, cudaMemcpyDeviceToDevice, st));
        if (d_sum) FOCK_CUDA(cudaMemcpyAsync(d_sum, c->d_vacuum, sizeof(double), cudaMemcpyDeviceToDevice, st));
        return FOCK_OK;
    }
    int order[FOCK_NMAX];
    slos_order(m, in_state, order);
    const double inf = host_prodnfact(m, in_state);
    // layers alternate between the two workspaces so that layer n-1 lands in work_a
    FOCK_REQUIRE(n == 1 || d_work_a, FOCK_ERR_ARG, "slos_prob_distribution: d_work_a is NULL");
    FOCK_REQUIRE(n <= 2 || d_work_b, FOCK_ERR_ARG, "slos_prob_distribution: d_work_b is NULL");
    const double *prev = c->d_vacuum;   // layer 0 = [1]: read-only constant of the context (safe on any stream)
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
    return FOCK_OK;   // asynchronous on `stream`
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
