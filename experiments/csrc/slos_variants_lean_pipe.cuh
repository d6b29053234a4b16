// Round-1 SLOS tile-kernel variants that lost their benchmarks (profiles/README.md): v2d lean addressing, v2c cp.async
// double buffering.  Kept for the record; NOT part of libfock_b200.so (cut out of perceval_b200/csrc/slos.cu in round 2).
// ================================================================================================================
// Lean tile kernel (v2d): the tiling and sweep of slos_tile_kernel with the per-edge instruction count cut down --
// descriptors carry ready-made byte addresses (row base / tail-block base), tail offsets are pre-scaled to bytes, the U
// entry of an occupied tail mode is one byte-extract + one LDS, unused tail lanes load a harmless valid address instead
// of being predicated, and prefix loads branch on the CTA-uniform edge count.  Same rounding sequence as v1 / v2.
// Used when the whole parent layer is resident; otherwise slos_tile_kernel (with its parent-window checks) runs.
// ================================================================================================================
#define LEAN_DB 128

struct __align__(16) LeanDesc {
    uint64_t cbase;
    const char *tptr;   // byte address of the tail-parent block of this prefix
    double pfact;
    int nz, pad;
};

__device__ __forceinline__ double2 ldg16(const char *base, uint32_t off) { return __ldg((const double2 *)(base + off)); }

template <int D, int MODE, bool RANGECHK>
__global__ void __launch_bounds__(TILE_BLOCK, 2) slos_lean_kernel(const __grid_constant__ TileArgs a) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    const int m = a.m, p = a.p, maxnz = a.maxnz;
    const int tid = threadIdx.x;
    double2 *s_u = (double2 *)tile_smem;
    LeanDesc *s_desc = (LeanDesc *)(s_u + m);
    double2 *e_u = (double2 *)(s_desc + LEAN_DB);
    uint64_t *e_ptr = (uint64_t *)(e_u + LEAN_DB * maxnz);
    __shared__ double s_red[TILE_BLOCK / 32];
    const uint64_t *__restrict__ bt = a.bt;
    const uint64_t *__restrict__ dt = a.dt;

    for (int i = tid; i < m; i += TILE_BLOCK) s_u[i] = a.U[(size_t)i * m + a.mk];
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

    // ---- per-thread tail: un-rank t in FS(D, u) once; byte offsets and U byte offsets of the occupied tail modes
    uint32_t g, t;
    if (G > 1) { g = tid / S; t = tid - g * S; } else { g = 0; t = chunk * TILE_BLOCK + tid; }
    const bool active = (g < G) && (t < S);
    uint32_t toffb[D];            // 16 * local rank of (tau - e_mode) in FS(D, u-1); 0 for unused entries
    uint32_t uoff[(D + 3) / 4];   // 16 * tail mode of entry c, one byte each
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < (D + 3) / 4; ++c) uoff[c] = 0;
    double tfact = 1.0;
#pragma unroll
    for (int c = 0; c < D; ++c) toffb[c] = 0;
    if (active) {
        uint64_t rem = t;
        uint32_t E = 0;
        int Tprev = u;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            int T = 0;
            if (i < D - 1) {
                const uint64_t *row = bt + (D - 1 - i) * FOCK_TMAX;
                T = Tprev;
                while (__ldg(row + T) > rem) --T;
                rem -= __ldg(row + T);
            }
            const int si = Tprev - T;
            if (si > 0) {
                const uint32_t off = (t - E) << 4;
#pragma unroll
                for (int c = 0; c < D; ++c)
                    if (c == cnt) { toffb[c] = off; uoff[c / 4] |= (uint32_t)(i * 16) << (8 * (c % 4)); }
                ++cnt;
                tfact *= c_factorial(si);
            }
            if (i < D - 1 && T > 0) E += (uint32_t)__ldg(dt + (D - 1 - i) * FOCK_TMAX + T);
            Tprev = T;
        }
    }
    const int wcnt = __reduce_max_sync(0xffffffffu, cnt);
    const char *__restrict__ parent_b = (const char *)a.parent;
    const char *s_utb = (const char *)(s_u + p);
    const uint32_t t16 = t << 4;
    double local_sum = 0.0;

    for (uint64_t rho0 = rho_a; rho0 < rho_b; rho0 += LEAN_DB) {
        const int nb = (int)((rho_b - rho0) < (uint64_t)LEAN_DB ? (rho_b - rho0) : (uint64_t)LEAN_DB);
        __syncthreads();
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
                    e_ptr[tid * maxnz + nz] = E;
                    e_u[tid * maxnz + nz] = s_u[i];
                    ++nz;
                    pf *= c_factorial(si);
                }
                base += __ldg(bt + (m - 1 - i) * FOCK_TMAX + Tfull);
                if (Tfull > 0) E += __ldg(dt + (m - 1 - i) * FOCK_TMAX + Tfull);
                Tprev = T;
            }
            for (int e = 0; e < nz; ++e) e_ptr[tid * maxnz + e] = (uint64_t)(parent_b + ((base - e_ptr[tid * maxnz + e]) << 4));
            LeanDesc td;
            td.cbase = base;
            td.tptr = parent_b + ((base - E) << 4);
            td.pfact = pf;
            td.nz = nz;
            td.pad = 0;
            s_desc[tid] = td;
        }
        __syncthreads();
        if (!active) continue;
        for (int i = (int)g; i < nb; i += (int)G) {
            const LeanDesc td = s_desc[i];
            const uint64_t r = td.cbase + t;
            if (RANGECHK && (r < a.cbegin || r >= a.cend)) continue;
            const uint64_t *ep = e_ptr + i * maxnz;
            const double2 *pu = e_u + i * maxnz;
            const int nz = td.nz;
            // ---- every load of this child is issued before any arithmetic
            double2 pv0, pv1, pv2, pv3, tv[8];
            if (nz > 0) pv0 = ldg16((const char *)ep[0], t16);
            if (nz > 1) pv1 = ldg16((const char *)ep[1], t16);
            if (nz > 2) pv2 = ldg16((const char *)ep[2], t16);
            if (nz > 3) pv3 = ldg16((const char *)ep[3], t16);
#pragma unroll
            for (int c = 0; c < 8 && c < D; ++c)
                if (c < wcnt) tv[c] = ldg16(td.tptr, toffb[c]);
            double2 acc = make_double2(0.0, 0.0);
            if (nz > 0) acc = cfma(pu[0], pv0, acc);
            if (nz > 1) acc = cfma(pu[1], pv1, acc);
            if (nz > 2) acc = cfma(pu[2], pv2, acc);
            if (nz > 3) acc = cfma(pu[3], pv3, acc);
            for (int e = 4; e < nz; ++e) acc = cfma(pu[e], ldg16((const char *)ep[e], t16), acc);
#pragma unroll
            for (int c = 0; c < 8 && c < D; ++c)
                if (c < cnt) acc = cfma(*(const double2 *)(s_utb + ((uoff[c / 4] >> (8 * (c % 4))) & 0xFFu)), tv[c], acc);
            if (D > 8 && wcnt > 8) {
#pragma unroll
                for (int c = 8; c < D; ++c)
                    if (c < wcnt) tv[c - 8] = ldg16(td.tptr, toffb[c]);
#pragma unroll
                for (int c = 8; c < D; ++c)
                    if (c < cnt) acc = cfma(*(const double2 *)(s_utb + ((uoff[c / 4] >> (8 * (c % 4))) & 0xFFu)), tv[c - 8], acc);
            }
            if (MODE & 1) a.child[r - a.cbegin] = acc;
            if (MODE & 2) {
                const double pr = (acc.x * acc.x + acc.y * acc.y) * a.inv_in_fact * (td.pfact * tfact);
                __stcs(a.probs + (r - a.cbegin), pr);
                local_sum += pr;
            }
        }
    }
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

// ================================================================================================================
// Pipelined tile kernel (v2c): the same tiling and sweep as slos_tile_kernel, but every parent a thread needs for prefix
// i+1 is copied global -> shared with cp.async (LDGSTS, 16 B per lane, no registers) while the thread multiplies the
// parents of prefix i out of its own shared-memory slots.  A thread only ever reads slots it filled itself, so the
// only synchronisation is cp.async.wait_group.  Slot budget: an occupied mode holds >= 1 photon, so a child never has
// more than k parents: nslots = max over classes of min(p,w) + min(D,u) <= k.
// ================================================================================================================
#define PIPE_DB 32

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D, int MODE, bool RANGECHK>
__global__ void __launch_bounds__(TILE_BLOCK, 2) slos_pipe_kernel(const __grid_constant__ TileArgs a) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    const int m = a.m, p = a.p, maxnz = a.maxnz, nslots = a.nslots;
    const int tid = threadIdx.x;
    double2 *s_u = (double2 *)tile_smem;
    TileDesc *s_desc = (TileDesc *)(s_u + m);
    double2 *e_u = (double2 *)(s_desc + PIPE_DB);
    uint64_t *e_pb = (uint64_t *)(e_u + PIPE_DB * maxnz);
    double2 *slots = (double2 *)(e_pb + PIPE_DB * maxnz);   // [2][nslots][TILE_BLOCK]
    __shared__ double s_red[TILE_BLOCK / 32];
    const uint64_t *__restrict__ bt = a.bt;
    const uint64_t *__restrict__ dt = a.dt;

    for (int i = tid; i < m; i += TILE_BLOCK) s_u[i] = a.U[(size_t)i * m + a.mk];
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
    const int nzs = p < w ? p : w;   // slots [0, nzs) hold prefix parents, [nzs, nzs + cnt) tail parents
    __syncthreads();

    // ---- per-thread tail: un-rank t in FS(D, u) once; keep (mode, parent offset) of the occupied tail modes, compacted
    uint32_t g, t;
    if (G > 1) { g = tid / S; t = tid - g * S; } else { g = 0; t = chunk * TILE_BLOCK + tid; }
    const bool active = (g < G) && (t < S);
    uint32_t toff[D];
    uint32_t tmode[(D + 5) / 6];
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < (D + 5) / 6; ++c) tmode[c] = 0;
    double tfact = 1.0;
#pragma unroll
    for (int c = 0; c < D; ++c) toff[c] = 0;
    if (active) {
        uint64_t rem = t;
        uint32_t E = 0;
        int Tprev = u;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            int T = 0;
            if (i < D - 1) {
                const uint64_t *row = bt + (D - 1 - i) * FOCK_TMAX;
                T = Tprev;
                while (__ldg(row + T) > rem) --T;
                rem -= __ldg(row + T);
            }
            const int si = Tprev - T;
            if (si > 0) {
                const uint32_t off = t - E;
#pragma unroll
                for (int c = 0; c < D; ++c)
                    if (c == cnt) { toff[c] = off; tmode[c / 6] |= (uint32_t)i << (5 * (c % 6)); }
                ++cnt;
                tfact *= c_factorial(si);
            }
            if (i < D - 1 && T > 0) E += (uint32_t)__ldg(dt + (D - 1 - i) * FOCK_TMAX + T);
            Tprev = T;
        }
    }
    const int wcnt = __reduce_max_sync(0xffffffffu, cnt);
    const double2 *__restrict__ parent = a.parent - a.pbegin;
    const double2 *__restrict__ parent_t = parent + t;
    const double2 *s_ut = s_u + p;
    double2 *myslot = slots + tid;
    double local_sum = 0.0;

    for (uint64_t rho0 = rho_a; rho0 < rho_b; rho0 += PIPE_DB) {
        const int nb = (int)((rho_b - rho0) < (uint64_t)PIPE_DB ? (rho_b - rho0) : (uint64_t)PIPE_DB);
        __syncthreads();
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
                    e_pb[tid * maxnz + nz] = E;
                    e_u[tid * maxnz + nz] = s_u[i];
                    ++nz;
                    pf *= c_factorial(si);
                }
                base += __ldg(bt + (m - 1 - i) * FOCK_TMAX + Tfull);
                if (Tfull > 0) E += __ldg(dt + (m - 1 - i) * FOCK_TMAX + Tfull);
                Tprev = T;
            }
            for (int e = 0; e < nz; ++e) e_pb[tid * maxnz + e] = base - e_pb[tid * maxnz + e];
            TileDesc td;
            td.cbase = base;
            td.tbase = base - E;
            td.pfact = pf;
            td.nz = nz;
            td.pad = 0;
            s_desc[tid] = td;
        }
        __syncthreads();
        if (!active || (int)g >= nb) continue;

        // issue the copies of prefix i into buffer b
        auto issue = [&](int i, int b) {
            const TileDesc td = s_desc[i];
            if (RANGECHK) {
                const uint64_t r = td.cbase + t;
                if (r < a.cbegin || r >= a.cend) return;
            }
            double2 *dst = myslot + (size_t)b * nslots * TILE_BLOCK;
            const uint64_t *pb = e_pb + i * maxnz;
            for (int e = 0; e < td.nz; ++e) cp_async16(dst + e * TILE_BLOCK, parent_t + pb[e]);
            const double2 *__restrict__ tbp = parent + td.tbase;
            double2 *dt2 = dst + nzs * TILE_BLOCK;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                if (c >= wcnt) break;
                if (c < cnt) cp_async16(dt2 + c * TILE_BLOCK, tbp + toff[c]);
            }
        };
        int i = (int)g, b = 0;
        issue(i, 0);
        cp_async_commit();
        for (; i < nb; i += (int)G) {
            const int in = i + (int)G;
            if (in < nb) issue(in, b ^ 1);
            cp_async_commit();
            if (in < nb) cp_async_wait<1>();
            else cp_async_wait<0>();
            const TileDesc td = s_desc[i];
            const uint64_t r = td.cbase + t;
            if (!RANGECHK || (r >= a.cbegin && r < a.cend)) {
                const double2 *src = myslot + (size_t)b * nslots * TILE_BLOCK;
                const double2 *pu = e_u + i * maxnz;
                // one accumulator, modes in ascending order: the same rounding sequence as the v1 / v2 kernels, so results
                // are bit-identical whichever kernel (and whichever sharding) produced them
                double2 acc = make_double2(0.0, 0.0);
                for (int e = 0; e < td.nz; ++e) acc = cfma(pu[e], src[e * TILE_BLOCK], acc);
                const double2 *st2 = src + nzs * TILE_BLOCK;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    if (c >= wcnt) break;
                    if (c < cnt) acc = cfma(s_ut[(tmode[c / 6] >> (5 * (c % 6))) & 31u], st2[c * TILE_BLOCK], acc);
                }
                if (MODE & 1) a.child[r - a.cbegin] = acc;
                if (MODE & 2) {
                    const double pr = (acc.x * acc.x + acc.y * acc.y) * a.inv_in_fact * (td.pfact * tfact);
                    __stcs(a.probs + (r - a.cbegin), pr);
                    local_sum += pr;
                }
            }
            b ^= 1;
        }
    }
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

