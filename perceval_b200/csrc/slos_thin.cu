// slos_thin.cu -- hybrid "thin-thread" SLOS tile kernel for large probability layers (complex128, sm_100a).
//
// Same tiling and prefix sweep as slos_tile_kernel (slos.cu; replaces FSMap.compute_slos_layer, reference call site
// perceval/backends/_slos.py:99, python twin :91-97, and xq.all_prob_normalize_output, _slos.py:199,213), built for <= 64
// registers (8 warps per scheduler instead of 4):
//   * the per-thread tail offsets live in a shared-memory column (one LDS.32 per slot) instead of 16 registers;
//   * the 8 LEADING tail modes are walked as WARP-uniform slots: slot s of a warp is the s-th leading tail mode that any of
//     its 32 lanes occupies, its unitary entry comes from the constant bank (column mk of U, one of TH_USLOTS per-launch
//     copies; LDC: no L1TEX wavefront), a lane that does not occupy the mode keeps a zero in its landing registers;
//   * the 8 TRAILING tail modes are per-lane compacted entries (offset | mode in one shared-memory word, unitary by LDS);
//   * the sweep is specialised on the warp's slot counts.
// Accumulation order per child: prefix modes ascending, then tail modes ascending -- the order of the gather and tile
// kernels (a skipped mode adds +0), so results are bit-identical.  Slot statistics and measurements: profiles/README.md.
// (The all-uniform-slot predecessor, v5, lives in experiments/csrc/slos_thin_v5.cu.)
#include <stdlib.h>

#include <mutex>

#include "slos_tile.cuh"

#define TH_DB 128     // prefix descriptors per batch
#define TH_ROWS 2     // prefix rows loaded up front

#ifndef TH_MINB
#define TH_MINB 4
#endif

int slos_mu_tuples(fock_ctx *c, int D, int u, uint32_t S, cudaStream_t st, const uint64_t **out);   // slos_mu.cu

// column mk of U in the constant bank (stream-ordered D2D copy before each launch): the unitary entry of a tail slot is
// warp-uniform, so it is fetched through the constant cache (LDC: no L1TEX wavefront, no LSU write-back) with the slot's
// mode read from a per-warp packed list.  (Doing the same for the prefix rows, or specialising the sweep on more slot
// counts, measured slower: 13.0 ms / 18.9 ms vs 12.45 ms for the last 12/24 layer.)
#define TH_USLOTS 8
__constant__ double2 c_th_u[TH_USLOTS][FOCK_QMAX];

struct __align__(16) ThDesc {
    uint64_t cbase;
    const char *tptr;   // byte address of the tail-parent block of this prefix
    double pfact;
    int nz, pad;
};

__device__ __forceinline__ double th_factorial(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f *= (double)i;
    return f;
}
__device__ __forceinline__ void th_bar() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ double2 th_ldg(const char *p) { return __ldg((const double2 *)p); }

struct ThShared {
    double2 *s_u;       // [m] column mk of U
    ThDesc *s_desc;     // [TH_DB]
    double2 *e_u;       // [TH_DB][maxnz]
    uint64_t *e_ptr;    // [TH_DB][maxnz]
    uint32_t *s_col;    // [D][TILE_BLOCK] byte offset of the tail parent of slot s for each thread
};

// ================================================================================================================
// v6: hybrid tail.  Slot statistics (profiles/README.md): a warp of 32 consecutive tail ranks needs 10.5 warp-uniform slots
// (v5) but only 6.3 compacted per-lane slots (v2), whose per-lane unitary look-ups cost 4 L1TEX wavefronts each.  The 8
// LEADING tail modes are almost warp-uniform (3.3 slots on average), the 8 TRAILING ones are not (3.5 compacted
// slots): v6 walks the leading modes as warp-uniform slots (unitary entry from the constant bank, as v5) and the
// trailing modes as per-lane compacted entries (offset and mode packed in one shared-memory word, unitary entry by LDS).
// ================================================================================================================
#define TH6_LEAD 8

template <int D, int WL, int WT, int MODE, bool RANGECHK>
__device__ __forceinline__ void th6_sweep(const TileArgs &a, const TileClass &cl, const ThShared &sh, const uint32_t pm, const uint32_t wm_lo, const uint32_t t,
                                          const int u, const int w, const uint64_t rho_a, const uint64_t rho_b, const bool active,
                                          double &local_sum) {
    const int m = a.m, p = a.p, maxnz = a.maxnz;
    const int tid = threadIdx.x;
    const uint64_t *__restrict__ bt = a.bt;
    const uint64_t *__restrict__ dt = a.dt;
    const char *__restrict__ parent_b = (const char *)a.parent;
    const uint32_t t16 = t << 4;
    const uint32_t *col = sh.s_col + tid;
    const uint32_t s_utrail = (uint32_t)__cvta_generic_to_shared(sh.s_u + p + TH6_LEAD);
#ifdef TH6_REGCOL
    // tuning variant (tools/build_variant.py): the slot offsets and the tail factorial in registers instead of one LDS per slot
    // and sweep step -- needs a larger register budget (TH_MINB = 3)
    uint32_t rcl[WL], rct[WT];
#pragma unroll
    for (int s = 0; s < WL; ++s) rcl[s] = col[s * TILE_BLOCK];
#pragma unroll
    for (int c = 0; c < WT; ++c) rct[c] = col[(TH6_LEAD + c) * TILE_BLOCK];
    const double tfreg = __hiloint2double((int)col[(D + 1) * TILE_BLOCK], (int)col[D * TILE_BLOCK]);
#define TH6_COLL(s) rcl[s]
#define TH6_COLT(c) rct[c]
#define TH6_TF tfreg
#else
#define TH6_COLL(s) col[(s) * TILE_BLOCK]
#define TH6_COLT(c) col[(TH6_LEAD + (c)) * TILE_BLOCK]
#define TH6_TF __hiloint2double((int)col[(D + 1) * TILE_BLOCK], (int)col[D * TILE_BLOCK])
#endif
    for (uint64_t rho0 = rho_a; rho0 < rho_b; rho0 += TH_DB) {
        const int nb = (int)((rho_b - rho0) < (uint64_t)TH_DB ? (rho_b - rho0) : (uint64_t)TH_DB);
        th_bar();
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
                    pf *= th_factorial(si);
                }
                if (a.slab) {   // slab-major: E is the PREFIX-local rank difference (same identity on FS(p, w))
                    if (T > 0) E += __ldg(dt + (p - 1 - i) * FOCK_TMAX + T);
                } else {
                    base += __ldg(bt + (m - 1 - i) * FOCK_TMAX + Tfull);
                    if (Tfull > 0) E += __ldg(dt + (m - 1 - i) * FOCK_TMAX + Tfull);
                }
                Tprev = T;
            }
            uint64_t tb;
            if (a.slab) {
                const uint64_t rho = rho0 + tid;
                base = cl.coff + rho * cl.S;
                for (int e = 0; e < nz; ++e)
                    sh.e_ptr[tid * maxnz + e] = (uint64_t)(parent_b + ((cl.roff + (rho - sh.e_ptr[tid * maxnz + e]) * cl.S) << 4));
                tb = cl.toff + rho * cl.Sp;
            } else {
                for (int e = 0; e < nz; ++e) sh.e_ptr[tid * maxnz + e] = (uint64_t)(parent_b + ((base - sh.e_ptr[tid * maxnz + e]) << 4));
                tb = base - E;
            }
            ThDesc td;
            td.cbase = base;
            td.tptr = parent_b + (tb << 4);
            td.pfact = pf;
            td.nz = nz;
            td.pad = 0;
            sh.s_desc[tid] = td;
        }
        th_bar();
        if (!active) continue;
#pragma unroll 1
        for (int i = 0; i < nb; ++i) {
            const ThDesc td = sh.s_desc[i];
            const uint64_t r = td.cbase + t;
            if (RANGECHK && (r < a.cbegin || r >= a.cend)) continue;
            const uint64_t *ep = sh.e_ptr + i * maxnz;
            const double2 *pu = sh.e_u + i * maxnz;
            const int nz = td.nz;
            // ---- group A: two prefix rows + the warp-uniform slots of the leading tail modes
            double2 pv[TH_ROWS];
#pragma unroll
            for (int e = 0; e < TH_ROWS; ++e)
                if (e < nz) pv[e] = th_ldg((const char *)ep[e] + t16);
            double2 tv[4];
#pragma unroll
            for (int s = 0; s < 4 && s < WL; ++s) {
                tv[s] = make_double2(0.0, 0.0);
                if (pm & (1u << s)) tv[s] = th_ldg(td.tptr + TH6_COLL(s));
            }
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int e = 0; e < TH_ROWS; ++e)
                if (e < nz) acc = cfma(pu[e], pv[e], acc);
            for (int e = TH_ROWS; e < nz; e += 2) {   // further rows two at a time (one round trip per pair)
                pv[0] = th_ldg((const char *)ep[e] + t16);
                if (e + 1 < nz) pv[1] = th_ldg((const char *)ep[e + 1] + t16);
                acc = cfma(pu[e], pv[0], acc);
                if (e + 1 < nz) acc = cfma(pu[e + 1], pv[1], acc);
            }
#pragma unroll
            for (int s = 0; s < 4 && s < WL; ++s) acc = cfma(c_th_u[a.uslot][p + ((wm_lo >> (4 * s)) & 15u)], tv[s], acc);
            if (WL > 4) {
#pragma unroll
                for (int s = 4; s < WL; ++s) {
                    tv[s - 4] = make_double2(0.0, 0.0);
                    if (pm & (1u << s)) tv[s - 4] = th_ldg(td.tptr + TH6_COLL(s));
                }
#pragma unroll
                for (int s = 4; s < WL; ++s) acc = cfma(c_th_u[a.uslot][p + ((wm_lo >> (4 * s)) & 15u)], tv[s - 4], acc);
            }
            // ---- group B: per-lane compacted entries of the trailing tail modes (offset | mode in one word)
#pragma unroll
            for (int c0 = 0; c0 < WT; c0 += 4) {
                uint32_t x[4];
#pragma unroll
                for (int c = c0; c < c0 + 4 && c < WT; ++c) {
                    x[c - c0] = 0;
                    tv[c - c0] = make_double2(0.0, 0.0);
                    if (pm & (1u << (TH6_LEAD + c))) {
                        x[c - c0] = TH6_COLT(c);
                        tv[c - c0] = th_ldg(td.tptr + (x[c - c0] & ~15u));
                    }
                }
#pragma unroll
                for (int c = c0; c < c0 + 4 && c < WT; ++c) {
                    double2 uc;
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(uc.x), "=d"(uc.y) : "r"(s_utrail + ((x[c - c0] & 15u) << 4)));
                    acc = cfma(uc, tv[c - c0], acc);
                }
            }
            if (MODE & 1) a.child[r - a.cbegin] = acc;
            if (MODE & 2) {
                const double pr = (acc.x * acc.x + acc.y * acc.y) * a.inv_in_fact * (td.pfact * TH6_TF);
                __stcs(a.probs + (r - a.cbegin), pr);
                local_sum += pr;
            }
        }
    }
}

template <int D, int MODE, bool RANGECHK>
__global__ void __launch_bounds__(TILE_BLOCK, TH_MINB) slos_thin6_kernel(const __grid_constant__ TileArgs a) {
    static_assert(D == 16, "v6 splits a 16-mode tail in 8 leading + 8 trailing modes");
    extern __shared__ __align__(16) unsigned char th_smem[];
    const int m = a.m, maxnz = a.maxnz;
    const int tid = threadIdx.x;
    ThShared sh;
    sh.s_u = (double2 *)th_smem;
    sh.s_desc = (ThDesc *)(sh.s_u + m);
    sh.e_u = (double2 *)(sh.s_desc + TH_DB);
    sh.e_ptr = (uint64_t *)(sh.e_u + TH_DB * maxnz);
    sh.s_col = (uint32_t *)(sh.e_ptr + TH_DB * maxnz);
    __shared__ double s_red[TILE_BLOCK / 32];
    const uint64_t *__restrict__ dt = a.dt;

    for (int i = tid; i < m; i += TILE_BLOCK) sh.s_u[i] = a.U[(size_t)(a.urow0 + i) * a.ustride + a.mk];
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

    const uint32_t t = chunk * TILE_BLOCK + tid;
    const bool active = t < S;
    uint32_t occ = 0;
    double tfact = 1.0;
    uint64_t tup = 0;
    if (active) tup = __ldg(a.tup[ci] + t);
#pragma unroll
    for (int i = 0; i < TH6_LEAD; ++i)
        if ((tup >> (4 * i)) & 15u) occ |= 1u << i;
    const uint32_t wbits = __reduce_or_sync(0xffffffffu, occ);   // leading tail modes any lane of the warp occupies
    uint32_t pm = 0;        // bits 0..7: lane occupies the mode of leading slot s; bits 8..15: trailing entry c present
    uint32_t wm_lo = 0;     // leading tail mode of slot s, 4 bits each (warp-uniform)
    int cnt_t = 0;
    {
        uint32_t E = 0;
        int T = u;
        int slot = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const int si = (int)((tup >> (4 * i)) & 15u);
            T -= si;
            if (i < TH6_LEAD) {
                if (wbits & (1u << i)) {        // warp-uniform
                    if (si > 0) {
                        sh.s_col[slot * TILE_BLOCK + tid] = (t - E) << 4;
                        pm |= 1u << slot;
                    }
                    wm_lo |= (uint32_t)i << (4 * slot);
                    ++slot;
                }
            } else if (si > 0) {
                sh.s_col[(TH6_LEAD + cnt_t) * TILE_BLOCK + tid] = ((t - E) << 4) | (uint32_t)(i - TH6_LEAD);
                pm |= 1u << (TH6_LEAD + cnt_t);
                ++cnt_t;
            }
            if (si > 1) tfact *= th_factorial(si);
            if (active && i < D - 1 && T > 0) E += (uint32_t)__ldg(dt + (D - 1 - i) * FOCK_TMAX + T);
        }
    }
    sh.s_col[D * TILE_BLOCK + tid] = (uint32_t)__double2loint(tfact);
    sh.s_col[(D + 1) * TILE_BLOCK + tid] = (uint32_t)__double2hiint(tfact);
    __syncwarp();
    const int nlead = __popc(wbits);
    const int ntrail = __reduce_max_sync(0xffffffffu, cnt_t);
    double local_sum = 0.0;
#define TH6_SWEEP(WL, WT) th6_sweep<D, WL, WT, MODE, RANGECHK>(a, a.cls[ci], sh, pm, wm_lo, t, u, w, rho_a, rho_b, active, local_sum)
    if (nlead <= 4) {
        if (ntrail <= 4) TH6_SWEEP(4, 4);
        else TH6_SWEEP(4, 8);
    } else {
        if (ntrail <= 4) TH6_SWEEP(8, 4);
        else TH6_SWEEP(8, 8);
    }
#undef TH6_SWEEP
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
template <int MODE, bool CHK>
static int th_launch1(const TileArgs &a, unsigned grid, size_t smem, cudaStream_t st) {
    FOCK_CUDA(cudaFuncSetAttribute(slos_thin6_kernel<16, MODE, CHK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    slos_thin6_kernel<16, MODE, CHK><<<grid, TILE_BLOCK, smem, st>>>(a);
    return fock_check_cuda(cudaGetLastError(), "slos_thin6_kernel");
}

static int th_launch(const TileArgs &a, bool wc, bool wp, bool chk, unsigned grid, size_t smem, cudaStream_t st) {
    if (wp && wc) return chk ? th_launch1<3, true>(a, grid, smem, st) : th_launch1<3, false>(a, grid, smem, st);
    if (wp) return chk ? th_launch1<2, true>(a, grid, smem, st) : th_launch1<2, false>(a, grid, smem, st);
    return chk ? th_launch1<1, true>(a, grid, smem, st) : th_launch1<1, false>(a, grid, smem, st);
}

bool slos_thin_supports(int D, int k) { return D == 16 && k <= 15; }

// The unitary column of a launch lives in one of TH_USLOTS constant-bank copies (the symbol is per device).  A launch takes
// the next copy of its device round-robin, makes its stream wait for the launch that used that copy TH_USLOTS launches ago,
// and records an event after the kernel.  Launches on different streams / from different host threads therefore use
// different copies and run concurrently; the mutex only covers the enqueue.
struct ThinDeviceState {
    std::mutex lock;
    cudaEvent_t done[TH_USLOTS] = {};
    unsigned next = 0;
};
static ThinDeviceState g_thin[64];

void slos_thin_destroy(fock_ctx *c) {
    if (!c || c->device < 0 || c->device >= 64) return;
    ThinDeviceState &ds = g_thin[c->device];
    std::lock_guard<std::mutex> g(ds.lock);
    for (int i = 0; i < TH_USLOTS; ++i)
        if (ds.done[i]) {
            cudaEventSynchronize(ds.done[i]);   // another context of this device may still use the ring: keep the events
        }
}

// `a` is a finished work plan of slos_layer_tiles (slos.cu) holding only classes with G == 1 and D == 16; the whole parent
// layer must be resident and 16-byte aligned.
int slos_thin_launch(fock_ctx *c, int D, TileArgs &a, bool want_child, bool want_probs, bool rangechk, unsigned grid, cudaStream_t st) {
    FOCK_REQUIRE(D == 16, FOCK_ERR_LIMIT, "slos_thin: tail width %d not instantiated", D);
    FOCK_REQUIRE(c->device >= 0 && c->device < 64, FOCK_ERR_LIMIT, "slos_thin: device index %d", c->device);
    for (int i = 0; i < a.ncls; ++i)
        if (int rc = slos_mu_tuples(c, D, a.cls[i].u, a.cls[i].S, st, &a.tup[i])) return rc;
    void *sym = nullptr;
    FOCK_CUDA(cudaGetSymbolAddress(&sym, c_th_u));
    const size_t smem = (size_t)a.m * 16 + (size_t)TH_DB * sizeof(ThDesc) + (size_t)TH_DB * a.maxnz * 24 +
                        (size_t)(D + 2) * TILE_BLOCK * 4 + 16;
    ThinDeviceState &ds = g_thin[c->device];
    std::lock_guard<std::mutex> g(ds.lock);
    const unsigned slot = ds.next++ % TH_USLOTS;
    if (!ds.done[slot]) FOCK_CUDA(cudaEventCreateWithFlags(&ds.done[slot], cudaEventDisableTiming));
    else FOCK_CUDA(cudaStreamWaitEvent(st, ds.done[slot], 0));
    a.uslot = (int)slot;
    FOCK_CUDA(cudaMemcpy2DAsync((char *)sym + (size_t)slot * FOCK_QMAX * 16, 16, a.U + (size_t)a.urow0 * a.ustride + a.mk, (size_t)a.ustride * 16, 16, (size_t)a.m,
                                cudaMemcpyDeviceToDevice, st));
    if (int rc = th_launch(a, want_child, want_probs, rangechk, grid, smem, st)) return rc;
    FOCK_CUDA(cudaEventRecord(ds.done[slot], st));
    c->launches++;
    return FOCK_OK;
}
