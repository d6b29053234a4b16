// slos_blk.cu -- SLOS layer kernel v3: tail blocks staged in shared memory with bulk async copies (sm_100a).
//
// Same operator as slos.cu (reference perceval/backends/_slos.py:91-99: child[s] = sum_j U[j,mk] * parent[s - e_j]),
// different execution plan.  Modes are split  LEAD (pl) | MID (pg) | TAIL (D <= 8):
//   * TAIL: the tail-parent block of a prefix -- |FS(D,u-1)| contiguous complex128, a few KB -- is brought into shared
//     memory by cp.async.bulk (mbarrier-tracked ring of stages, issued several steps ahead by one thread: no registers,
//     no scoreboard stall) and every tail edge becomes a shared-memory gather.  What a tail column needs (<= D packed
//     (mode, stage offset) pairs and prod tau_i!) is tabulated ONCE per context in global memory, so a thread is not tied
//     to a column: a step flattens G prefixes x cw columns over the 256 threads and stays dense for every tile size.
//     Large tail blocks are split by the occupation a of the first tail mode: sub-block a needs two contiguous runs
//     (the aligned run of the mode-0 edge and its own sub-block), and the table holds offsets into that stage layout.
//   * MID + LEAD: prefix edges are aligned rows parent[rowbase + t]: streaming 16-byte loads, four in flight per thread.
//     Work items are ordered LEAD-major: all CTAs in flight work inside the tail blocks of a few lead prefixes, whose
//     parent block (|FS(m-pl, u)| values) stays L2-resident, so MID rows and tail blocks hit L2 and only the pl LEAD
//     rows stream from HBM.
//   * prefix descriptors (bases, packed (mode, row base) edge lists, prod pi_i!) are un-ranked cooperatively, one
//     prefix per thread, 128 at a time.
// Classes whose stage would exceed the 13-bit offset field / the stage budget are left to the v2 tile kernel.
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

#define BLK_THREADS 256
#define BLK_DB 128            // prefix descriptors per batch
#define BLK_GMAX 128          // prefixes per step
#define BLK_MAXGRP 34
#define BLK_OFFMASK 0x1FFFu
#define BLK_EDGEMASK 0x00FFFFFFFFFFFFFFull
#define BLK_PLAIN_MAX 1024u   // largest tail-parent block (elements) staged whole
#define BLK_STAGE_MAX 3072u   // largest stage (elements) = 48 KB
#define BLK_CWMAX 512u

__constant__ double c_fact[FOCK_TMAX] = {1.0, 1.0, 2.0, 6.0, 24.0, 120.0, 720.0, 5040.0, 40320.0, 362880.0, 3628800.0, 39916800.0,
    479001600.0, 6227020800.0, 87178291200.0, 1307674368000.0, 20922789888000.0, 355687428096000.0, 6402373705728000.0,
    121645100408832000.0, 2432902008176640000.0, 51090942171709440000.0, 1124000727777607680000.0,
    25852016738884976640000.0, 620448401733239439360000.0, 15511210043330985984000000.0, 403291461126605635584000000.0,
    10888869450418352160768000000.0, 304888344611713860501504000000.0, 8841761993739701954543616000000.0,
    265252859812191058636308480000000.0, 8222838654177922817725562880000000.0, 263130836933693530167218012160000000.0};

struct BlkCls {   // one tail class: (u) staged whole, or (u, a) = sub-block a of a large class
    int u, a;                       // a = -1: whole block
    uint32_t col0, S;               // columns [col0, col0 + S) of FS(D,u)
    uint32_t run0_src, run0_len;    // stage[0, run0_len)                 <- parent[tbase + run0_src ...]
    uint32_t run1_src, run1_len;    // stage[run0_len, run0_len+run1_len) <- parent[tbase + run1_src ...]
    uint32_t G, CW, nchunks, NS;
    uint32_t staged, per_item;
    uint32_t tab_off, pad;          // table index of column col0
};

struct BlkSeg {   // units of one (mid weight v, class) inside a lead prefix; one unit = per_item mid configurations x S columns
    int v, cid;
    uint32_t begin_local, nranges;   // unit indices [begin_local, begin_local + nranges) inside the lead prefix
};

struct BlkGroup {   // all lead prefixes of one weight: identical unit structure
    int wl, u_rest, seg_begin, nseg;
    uint64_t lam_lo, lam_hi, item_begin;
    uint32_t upl;        // units per batch of Lb lead prefixes
    uint32_t ncut;       // every batch is cut into ncut items (unit boundaries cuts[cut_off .. cut_off+ncut])
    uint32_t Lb;         // lead prefixes per batch (1 when the tail block of one lead prefix is large)
    uint32_t cut_off;
};

struct BlkArgs {
    int m, k, mk, p, pl, D, maxnz, ngrp;
    const uint64_t *bt, *dt;
    const double2 *U;
    const double2 *parent;
    double2 *child;
    double *probs;
    double *sum;
    double inv_in_fact;
    uint64_t cbegin, cend;
    const BlkCls *cls;
    const BlkSeg *segs;
    const uint32_t *cuts;
    const uint4 *tab;
    const double *tfact;
    BlkGroup grp[BLK_MAXGRP];
};

struct __align__(16) BlkDesc {
    uint64_t cbase, tbase;
    double pfact;
    int nz, pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// dynamic shared-memory layout (bytes), shared by host and device
struct BlkSmem {
    uint32_t off_u, off_bar, off_ctx, off_desc, off_edge, off_ring, fixed;
};
__host__ __device__ inline BlkSmem blk_smem_layout(int m, int maxnz) {
    BlkSmem L;
    uint32_t o = 0;
    L.off_u = o; o += (uint32_t)m * 16;
    L.off_bar = o; o += 8 * 8;
    L.off_ctx = o; o += 128;
    L.off_desc = o; o += BLK_DB * (uint32_t)sizeof(BlkDesc);
    L.off_edge = o; o += BLK_DB * (uint32_t)maxnz * 8;
    o = (o + 127) & ~127u;
    L.off_ring = o;
    L.fixed = o;
    return L;
}

// Everything one step needs, written to shared memory by thread 0 so that the hot loop keeps almost nothing live.
struct __align__(16) BlkStepCtx {
    const double2 *parent;
    double2 *child;              // already offset by -cbegin
    double *probs;               // already offset by -cbegin
    const uint4 *tab;
    const double *tfact;
    const double2 *stage;        // staged tail parents of this step (shared memory), or nullptr
    uint64_t cbegin, cend;
    double inv_in_fact;
    uint32_t g0, gc, cw, q256, r256, tcol0, stage_elems;
    int u, p, maxnz;
};

template <int MODE, bool STAGED, bool RANGECHK>
__device__ __noinline__ double blk_step(const BlkStepCtx *__restrict__ sc, const BlkDesc *__restrict__ s_desc,
                                        const uint64_t *__restrict__ s_edge, const double2 *__restrict__ s_u) {
    const uint32_t cw = sc->cw, slots = sc->gc * cw, q256 = sc->q256, r256 = sc->r256;
    const int maxnz = sc->maxnz, u = sc->u;
    const char *__restrict__ parent_b = (const char *)sc->parent;
    const double2 *__restrict__ s_ut = s_u + sc->p;
    uint32_t g = threadIdx.x / cw, tl = threadIdx.x - g * cw;
    g += sc->g0;
    double local_sum = 0.0;
    for (uint32_t idx = threadIdx.x; idx < slots; idx += BLK_THREADS) {
        const BlkDesc d = s_desc[g];
        const uint64_t *__restrict__ ed = s_edge + g * maxnz;
        const uint32_t t = sc->tcol0 + tl;
        const uint64_t r = d.cbase + t;
        if (!RANGECHK || (r >= sc->cbegin && r < sc->cend)) {
            double2 acc = make_double2(0.0, 0.0);
            const char *__restrict__ prow = parent_b + (size_t)t * 16;
            int e0 = 0;
            for (; e0 + 4 <= d.nz; e0 += 4) {
                const uint64_t ea = ed[e0], eb = ed[e0 + 1], ec = ed[e0 + 2], edd = ed[e0 + 3];
                const double2 va = ld_stream((const double2 *)(prow + (ea & BLK_EDGEMASK)));
                const double2 vb = ld_stream((const double2 *)(prow + (eb & BLK_EDGEMASK)));
                const double2 vc = ld_stream((const double2 *)(prow + (ec & BLK_EDGEMASK)));
                const double2 vd = ld_stream((const double2 *)(prow + (edd & BLK_EDGEMASK)));
                acc = cfma(s_u[ea >> 56], va, acc);
                acc = cfma(s_u[eb >> 56], vb, acc);
                acc = cfma(s_u[ec >> 56], vc, acc);
                acc = cfma(s_u[edd >> 56], vd, acc);
            }
            const int rest = d.nz - e0;
            if (rest > 0) {
                const uint64_t ea = ed[e0];
                const uint64_t eb = rest > 1 ? ed[e0 + 1] : ea;
                const uint64_t ec = rest > 2 ? ed[e0 + 2] : ea;
                const double2 va = ld_stream((const double2 *)(prow + (ea & BLK_EDGEMASK)));
                double2 vb = va, vc = va;
                if (rest > 1) vb = ld_stream((const double2 *)(prow + (eb & BLK_EDGEMASK)));
                if (rest > 2) vc = ld_stream((const double2 *)(prow + (ec & BLK_EDGEMASK)));
                acc = cfma(s_u[ea >> 56], va, acc);
                if (rest > 1) acc = cfma(s_u[eb >> 56], vb, acc);
                if (rest > 2) acc = cfma(s_u[ec >> 56], vc, acc);
            }
            double tf = 1.0;
            if (u > 0) {
                const uint4 ent = __ldg(sc->tab + tl);
                if (MODE & 2) tf = __ldg(sc->tfact + tl);
                const double2 *__restrict__ PC;
                if (STAGED) PC = sc->stage + (size_t)(g - sc->g0) * sc->stage_elems;
                else PC = sc->parent + d.tbase;
                uint32_t w2 = ent.x;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c == 2) w2 = ent.y;
                    if (c == 4) w2 = ent.z;
                    if (c == 6) w2 = ent.w;
                    const uint32_t e = (c & 1) ? (w2 >> 16) : (w2 & 0xFFFFu);
                    if (e == 0xFFFFu) break;
                    acc = cfma(s_ut[e >> 13], PC[e & BLK_OFFMASK], acc);
                }
            }
            if (MODE & 1) sc->child[r] = acc;
            if (MODE & 2) {
                const double pr = (acc.x * acc.x + acc.y * acc.y) * sc->inv_in_fact * (d.pfact * tf);
                __stcs(sc->probs + r, pr);
                local_sum += pr;
            }
        }
        g += q256;
        tl += r256;
        if (tl >= cw) { tl -= cw; ++g; }
    }
    return local_sum;
}

__device__ __forceinline__ void blk_issue(const BlkArgs &a, const BlkCls &cl, const BlkDesc *s_desc, double2 *stage, uint64_t *bar,
                                          uint32_t g0, uint32_t gc) {
    const uint32_t se = cl.run0_len + cl.run1_len;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, gc * se * 16);
    for (uint32_t g = 0; g < gc; ++g) {
        const double2 *src = a.parent + s_desc[g0 + g].tbase;
        double2 *dst = stage + (size_t)g * se;
        if (cl.run0_len) bulk_g2s(dst, src + cl.run0_src, cl.run0_len * 16, bar);
        if (cl.run1_len) bulk_g2s(dst + cl.run0_len, src + cl.run1_src, cl.run1_len * 16, bar);
    }
}

template <int MODE, bool RANGECHK>
__global__ void __launch_bounds__(BLK_THREADS, 3) slos_blk_kernel(const __grid_constant__ BlkArgs a) {
    extern __shared__ __align__(128) unsigned char blk_smem[];
    __shared__ double s_red[BLK_THREADS / 32];
    const int m = a.m, p = a.p, pl = a.pl, maxnz = a.maxnz;
    const int tid = threadIdx.x;
    const BlkSmem L = blk_smem_layout(m, maxnz);
    double2 *s_u = (double2 *)(blk_smem + L.off_u);
    uint64_t *s_bar = (uint64_t *)(blk_smem + L.off_bar);
    BlkStepCtx *s_ctx = (BlkStepCtx *)(blk_smem + L.off_ctx);
    BlkDesc *s_desc = (BlkDesc *)(blk_smem + L.off_desc);
    uint64_t *s_edge = (uint64_t *)(blk_smem + L.off_edge);
    double2 *s_ring = (double2 *)(blk_smem + L.off_ring);
    const uint64_t *__restrict__ bt = a.bt;
    const uint64_t *__restrict__ dt = a.dt;

    for (int i = tid; i < m; i += BLK_THREADS) s_u[i] = a.U[(size_t)i * m + a.mk];
    if (tid == 0) {
        for (uint32_t s = 0; s < 8; ++s) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- decode the work item (uniform, every thread): batch of lead prefixes [lam0, lam0 + nlam) x units [sub_a, sub_b)
    int gi = 0;
    for (int g = 1; g < a.ngrp; ++g)
        if ((uint64_t)blockIdx.x >= a.grp[g].item_begin) gi = g;
    const BlkGroup &GR = a.grp[gi];
    const uint64_t local = (uint64_t)blockIdx.x - GR.item_begin;
    const uint64_t lam0 = GR.lam_lo + (local / GR.ncut) * GR.Lb;
    const uint64_t nlam = (GR.lam_hi - lam0) < (uint64_t)GR.Lb ? (GR.lam_hi - lam0) : (uint64_t)GR.Lb;
    const uint32_t cj = (uint32_t)(local % GR.ncut);
    const uint32_t sub_a = __ldg(a.cuts + GR.cut_off + cj), sub_b = __ldg(a.cuts + GR.cut_off + cj + 1);
    const int pg = p - pl, wl = GR.wl;
    double local_sum = 0.0;
    uint32_t phase_bits = 0;   // bit s = parity the next wait on barrier s must use
    __syncthreads();
    {
        int sidx = 0;
        BlkSeg sg = a.segs[GR.seg_begin];
        for (uint32_t sub = sub_a; sub < sub_b; ++sub) {
            while (sub >= sg.begin_local + sg.nranges) sg = a.segs[GR.seg_begin + (++sidx)];
            const BlkCls cl = a.cls[sg.cid];
            const int v = sg.v, u = cl.u, w = wl + v;
            const uint64_t Rv = pg == 0 ? 1 : __ldg(bt + (pg - 1) * FOCK_TMAX + v + 1);   // |FS(pg, v)|
            // this unit = (lead prefix, mid configuration) pairs [q_a, q_b) of the batch, lead-major
            const uint64_t q_a = (uint64_t)(sub - sg.begin_local) * cl.per_item;
            uint64_t q_b = q_a + cl.per_item;
            if (q_b > nlam * Rv) q_b = nlam * Rv;
            const uint32_t G = cl.G, NS = cl.NS;
            const bool staged = cl.staged != 0;
            const uint32_t stage_elems = cl.run0_len + cl.run1_len;
            const uint32_t cw = cl.S;
            __syncthreads();   // nobody still reads the previous unit's step context
            if (tid == 0) {
                BlkStepCtx sc;
                sc.parent = a.parent;
                sc.child = a.child ? a.child - a.cbegin : nullptr;
                sc.probs = a.probs ? a.probs - a.cbegin : nullptr;
                sc.tab = a.tab + cl.tab_off;
                sc.tfact = a.tfact + cl.tab_off;
                sc.stage = nullptr;
                sc.cbegin = a.cbegin; sc.cend = a.cend;
                sc.inv_in_fact = a.inv_in_fact;
                sc.g0 = 0; sc.gc = 0;
                sc.cw = cw; sc.q256 = BLK_THREADS / cw; sc.r256 = BLK_THREADS % cw;
                sc.tcol0 = cl.col0; sc.stage_elems = stage_elems;
                sc.u = u; sc.p = p; sc.maxnz = maxnz;
                *s_ctx = sc;
            }

            for (uint64_t q0 = q_a; q0 < q_b; q0 += BLK_DB) {
                const uint32_t nb = (uint32_t)((q_b - q0) < (uint64_t)BLK_DB ? (q_b - q0) : (uint64_t)BLK_DB);
                __syncthreads();
                // ---- cooperative prefix descriptors: thread i un-ranks pair q0 + i = (lead prefix of FS(pl, wl), mid of FS(pg, v))
                if ((uint32_t)tid < nb) {
                    const uint64_t q = q0 + tid;
                    const uint64_t lq = q / Rv;
                    uint64_t rem = lam0 + lq;
                    const uint64_t grank = q - lq * Rv;
                    int Tprev = wl;
                    uint64_t base = 0, E = 0;
                    int nz = 0;
                    double pf = 1.0;
                    uint64_t *ed = s_edge + (size_t)tid * maxnz;
                    for (int i = 0; i < p; ++i) {
                        // modes [0, pl) are un-ranked in FS(pl, wl), modes [pl, p) in FS(pg, v)
                        const bool lead = i < pl;
                        if (i == pl) { rem = grank; Tprev = v; }
                        const int nloc = lead ? pl : pg, iloc = lead ? i : i - pl;
                        int T = 0;
                        if (iloc < nloc - 1) {
                            const uint64_t *row = bt + (nloc - 1 - iloc) * FOCK_TMAX;
                            T = Tprev;
                            while (__ldg(row + T) > rem) --T;
                            rem -= __ldg(row + T);
                        }
                        const int si = Tprev - T;
                        const int Tfull = T + u + (lead ? v : 0);   // photons right of mode i in the full state
                        if (si > 0) {
                            ed[nz++] = ((uint64_t)i << 56) | E;   // E_j for now; turned into the row base below
                            pf *= c_fact[si];
                        }
                        base += __ldg(bt + (m - 1 - i) * FOCK_TMAX + Tfull);
                        if (Tfull > 0) E += __ldg(dt + (m - 1 - i) * FOCK_TMAX + Tfull);
                        Tprev = T;
                    }
                    for (int e = 0; e < nz; ++e) {
                        const uint64_t ev = ed[e];
                        ed[e] = (ev & ~BLK_EDGEMASK) | ((base - (ev & BLK_EDGEMASK)) << 4);
                    }
                    BlkDesc d;
                    d.cbase = base;
                    d.tbase = base - E;
                    d.pfact = pf;
                    d.nz = nz;
                    d.pad = 0;
                    s_desc[tid] = d;
                }
                __syncthreads();
                const uint32_t nsteps = (nb + G - 1) / G;
                if (staged) {
                    if (tid == 0) {
                        const uint32_t pre = (NS - 1) < nsteps ? (NS - 1) : nsteps;
                        for (uint32_t s = 0; s < pre; ++s) {
                            const uint32_t gg0 = s * G, gcc = (nb - gg0) < G ? (nb - gg0) : G;
                            blk_issue(a, cl, s_desc, s_ring + (size_t)(s % NS) * G * stage_elems, &s_bar[s % NS], gg0, gcc);
                        }
                    }
                    for (uint32_t st = 0; st < nsteps; ++st) {
                        const uint32_t g0 = st * G, gc = (nb - g0) < G ? (nb - g0) : G;
                        const uint32_t sn = st + NS - 1;
                        if (tid == 0 && sn < nsteps) {
                            const uint32_t gg0 = sn * G, gcc = (nb - gg0) < G ? (nb - gg0) : G;
                            blk_issue(a, cl, s_desc, s_ring + (size_t)(sn % NS) * G * stage_elems, &s_bar[sn % NS], gg0, gcc);
                        }
                        const uint32_t stg = st % NS;
                        if (tid == 0) { s_ctx->g0 = g0; s_ctx->gc = gc; s_ctx->stage = s_ring + (size_t)stg * G * stage_elems; }
                        mbar_wait(&s_bar[stg], (phase_bits >> stg) & 1u);
                        phase_bits ^= 1u << stg;
                        __syncthreads();
                        local_sum += blk_step<MODE, true, RANGECHK>(s_ctx, s_desc, s_edge, s_u);
                        __syncthreads();   // every thread is done with this stage before it is refilled
                    }
                } else {
                    for (uint32_t st = 0; st < nsteps; ++st) {
                        const uint32_t g0 = st * G, gc = (nb - g0) < G ? (nb - g0) : G;
                        if (st) __syncthreads();
                        if (tid == 0) { s_ctx->g0 = g0; s_ctx->gc = gc; }
                        __syncthreads();
                        local_sum += blk_step<MODE, false, RANGECHK>(s_ctx, s_desc, s_edge, s_u);
                    }
                }
            }
        }
    }
    if ((MODE & 2) && a.sum) {
        local_sum = warp_sum(local_sum);
        if ((tid & 31) == 0) s_red[tid >> 5] = local_sum;
        __syncthreads();
        if (tid < 32) {
            double v = tid < BLK_THREADS / 32 ? s_red[tid] : 0.0;
            v = warp_sum(v);
            if (tid == 0) atomicAdd(a.sum, v);
        }
    }
}

// ---- column tables: one thread per (class, column)
__global__ void __launch_bounds__(256) blk_table_kernel(int D, int ncls, const BlkCls *__restrict__ cls, const uint64_t *__restrict__ bt,
                                                        const uint64_t *__restrict__ dt, uint32_t total, uint4 *__restrict__ tab,
                                                        double *__restrict__ tfact) {
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= total) return;
    int ci = 0;
    for (int c = 1; c < ncls; ++c)
        if (idx >= cls[c].tab_off) ci = c;
    const BlkCls cl = cls[ci];
    const int u = cl.u;
    const uint32_t t = cl.col0 + (idx - cl.tab_off);
    uint64_t rem = t;
    uint32_t E = 0;
    int Tprev = u, cnt = 0;
    double tf = 1.0;
    uint32_t ent[8];
    for (int c = 0; c < 8; ++c) ent[c] = 0xFFFFu;
    for (int i = 0; i < D; ++i) {
        int T = 0;
        if (i < D - 1) {
            const uint64_t *row = bt + (D - 1 - i) * FOCK_TMAX;
            T = Tprev;
            while (row[T] > rem) --T;
            rem -= row[T];
        }
        const int si = Tprev - T;
        if (si > 0) {
            uint32_t off = t - E;   // offset inside the whole tail-parent block FS(D, u-1)
            if (cl.a >= 0) {        // sub-block a: mode 0 reads run 0 (aligned), deeper modes read run 1 (own sub-block)
                if (i == 0) off = off - cl.run0_src;
                else off = off - cl.run1_src + cl.run0_len;
            }
            ent[cnt++] = ((uint32_t)i << 13) | off;
            tf *= c_fact[si];
        }
        if (i < D - 1 && T > 0) E += (uint32_t)dt[(D - 1 - i) * FOCK_TMAX + T];
        Tprev = T;
    }
    tab[idx] = make_uint4(ent[0] | (ent[1] << 16), ent[2] | (ent[3] << 16), ent[4] | (ent[5] << 16), ent[6] | (ent[7] << 16));
    tfact[idx] = tf;
}

// ---------------------------------------------------------------- host planner
uint64_t slos_host_prefix_base(int m, int p, int w, int u, uint64_t rho);   // slos.cu

struct BlkPlan {
    std::vector<BlkGroup> groups;
    BlkSeg *d_segs = nullptr;
    uint32_t *d_cuts = nullptr;
    uint64_t items = 0;
};

struct BlkTailSet {   // everything that depends on D only
    int D = 0, u_limit = 0;
    std::vector<BlkCls> cls;
    std::vector<std::vector<int>> cids;   // per u
    BlkCls *d_cls = nullptr;
    uint4 *d_tab = nullptr;
    double *d_tfact = nullptr;
    uint32_t max_ring_bytes = 0;
};

struct BlkState {
    std::mutex mu;
    std::map<int, BlkTailSet> tails;
    std::map<std::tuple<int, int, int, int, uint64_t, uint64_t>, BlkPlan> plans;
};

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

int slos_blk_tail_modes(int m) { return m >= 14 ? 8 : (m >= 10 ? 6 : (m >= 6 ? 4 : 0)); }

static int blk_tailset(fock_ctx *c, BlkState *S, int D, BlkTailSet **out) {
    auto it = S->tails.find(D);
    if (it != S->tails.end()) { *out = &it->second; return FOCK_OK; }
    BlkTailSet ts;
    ts.D = D;
    const uint64_t *bt = fock_host_bt();
    const uint32_t unit_children = (uint32_t)env_int("FOCK_BLK_UNIT", 4096);
    const int ns_want = env_int("FOCK_BLK_NS", 4);
    uint32_t tab_total = 0;
    int u_limit = FOCK_NMAX + 1;
    ts.cids.resize(FOCK_NMAX + 1);
    auto finish = [&](BlkCls &cl) {
        cl.CW = cl.S;
        cl.nchunks = 1;
        {
            uint32_t g = (1024 + cl.S - 1) / cl.S;
            cl.G = g < 1 ? 1 : (g > BLK_GMAX ? BLK_GMAX : g);
        }
        const uint32_t se = cl.run0_len + cl.run1_len;
        cl.staged = (cl.a >= 0 || se >= 16) ? 1 : 0;
        if (se == 0) cl.staged = 0;
        uint32_t ns = 1;
        if (cl.staged) {
            ns = (uint32_t)ns_want;
            while (ns > 1 && (size_t)ns * cl.G * se * 16 > 32 * 1024) --ns;
            if (ns > 8) ns = 8;
            const uint32_t ring = ns * cl.G * se * 16;
            if (ring > ts.max_ring_bytes) ts.max_ring_bytes = ring;
        }
        cl.NS = ns;
        uint32_t per = (unit_children + cl.S - 1) / cl.S;   // mid configurations per unit
        if (per < cl.G) per = cl.G;
        cl.per_item = per;
        cl.tab_off = tab_total;
        tab_total += cl.S;
    };
    for (int u = 0; u <= FOCK_NMAX && u < u_limit; ++u) {
        const uint64_t S64 = fock_count(D, u), Sp64 = u ? fock_count(D, u - 1) : 0;
        if (Sp64 <= BLK_PLAIN_MAX) {
            BlkCls cl;
            memset(&cl, 0, sizeof cl);
            cl.u = u; cl.a = -1;
            cl.col0 = 0; cl.S = (uint32_t)S64;
            cl.run1_src = 0; cl.run1_len = (uint32_t)Sp64;
            finish(cl);
            ts.cids[u].push_back((int)ts.cls.size());
            ts.cls.push_back(cl);
            continue;
        }
        // split by the occupation a of the first tail mode; sub-block a = columns [Bt[D-1][u-a], +|FS(D-1,u-a)|)
        std::vector<BlkCls> sub;
        bool ok = D >= 2;
        for (int a = 0; a <= u && ok; ++a) {
            BlkCls cl;
            memset(&cl, 0, sizeof cl);
            cl.u = u; cl.a = a;
            cl.col0 = (uint32_t)bt[(D - 1) * FOCK_TMAX + (u - a)];
            cl.S = (uint32_t)fock_count(D - 1, u - a);
            if (a >= 1) { cl.run0_src = cl.col0; cl.run0_len = cl.S; }
            if (u - a >= 1) { cl.run1_src = (uint32_t)bt[(D - 1) * FOCK_TMAX + (u - 1 - a)]; cl.run1_len = (uint32_t)fock_count(D - 1, u - a - 1); }
            if (cl.run0_len + cl.run1_len > BLK_STAGE_MAX) ok = false;
            sub.push_back(cl);
        }
        if (!ok) { u_limit = u; break; }
        for (auto &cl : sub) {
            finish(cl);
            ts.cids[u].push_back((int)ts.cls.size());
            ts.cls.push_back(cl);
        }
    }
    const int env_ul = env_int("FOCK_BLK_ULIM", 0);
    if (env_ul > 0 && env_ul < u_limit) u_limit = env_ul;
    ts.u_limit = u_limit;
    FOCK_CUDA(cudaMalloc(&ts.d_cls, ts.cls.size() * sizeof(BlkCls)));
    FOCK_CUDA(cudaMalloc(&ts.d_tab, (size_t)tab_total * sizeof(uint4)));
    FOCK_CUDA(cudaMalloc(&ts.d_tfact, (size_t)tab_total * sizeof(double)));
    FOCK_CUDA(cudaMemcpy(ts.d_cls, ts.cls.data(), ts.cls.size() * sizeof(BlkCls), cudaMemcpyHostToDevice));
    blk_table_kernel<<<(tab_total + 255) / 256, 256>>>(D, (int)ts.cls.size(), ts.d_cls, c->d_bt, c->d_dt, tab_total, ts.d_tab, ts.d_tfact);
    c->launches++;
    FOCK_CUDA(cudaGetLastError());
    FOCK_CUDA(cudaDeviceSynchronize());
    auto res = S->tails.emplace(D, std::move(ts));
    *out = &res.first->second;
    return FOCK_OK;
}

// smallest u that the v3 kernel does NOT take; classes u >= this go to the v2 tile kernel
int slos_blk_u_limit(fock_ctx *c, int D, int k) {
    BlkState *S = (BlkState *)c->blk_state;
    if (!S) { S = new BlkState(); c->blk_state = S; }
    std::lock_guard<std::mutex> lk(S->mu);
    BlkTailSet *ts = nullptr;
    if (blk_tailset(c, S, D, &ts)) return 0;
    return ts->u_limit <= k ? ts->u_limit : k + 1;
}

void slos_blk_destroy(fock_ctx *c) {
    BlkState *S = (BlkState *)c->blk_state;
    if (!S) return;
    for (auto &kv : S->tails) { cudaFree(kv.second.d_cls); cudaFree(kv.second.d_tab); cudaFree(kv.second.d_tfact); }
    for (auto &kv : S->plans) { cudaFree(kv.second.d_segs); cudaFree(kv.second.d_cuts); }
    delete S;
    c->blk_state = nullptr;
}

template <int MODE>
static int blk_launch(fock_ctx *c, const BlkArgs &a, bool rangechk, unsigned grid, size_t smem, cudaStream_t st) {
    if (rangechk) {
        FOCK_CUDA(cudaFuncSetAttribute(slos_blk_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        slos_blk_kernel<MODE, true><<<grid, BLK_THREADS, smem, st>>>(a);
    } else {
        FOCK_CUDA(cudaFuncSetAttribute(slos_blk_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        slos_blk_kernel<MODE, false><<<grid, BLK_THREADS, smem, st>>>(a);
    }
    c->launches++;
    return fock_check_cuda(cudaGetLastError(), "slos_blk_kernel");
}

// Runs every class with u < u_limit over child ranks [cb, ce); the parent layer must be fully resident.
int slos_layer_blocks(fock_ctx *c, int D, int m, int k, const double *d_U, int mk, const double *d_parent, double *d_child,
                      double *d_probs, double *d_sum, double in_prodnfact, uint64_t cb, uint64_t ce, cudaStream_t st) {
    BlkState *S = (BlkState *)c->blk_state;
    if (!S) { S = new BlkState(); c->blk_state = S; }
    std::lock_guard<std::mutex> lk(S->mu);
    BlkTailSet *ts = nullptr;
    if (int rc = blk_tailset(c, S, D, &ts)) return rc;
    const int p = m - D;
    int pl = env_int("FOCK_BLK_PL", 8);
    if (pl > 10) pl = 10;
    if (pl > p) pl = p;
    if (pl < 1) pl = 1;
    const int pg = p - pl;
    const bool full = (cb == 0 && ce == fock_count(m, k));
    // ---- plan: lead-weight groups and their segments (cached per layer shape)
    const auto key = std::make_tuple(m, k, D, pl, cb, ce);
    auto pit = S->plans.find(key);
    if (pit == S->plans.end()) {
        BlkPlan plan;
        std::vector<BlkSeg> segs;
        std::vector<uint32_t> cuts;
        uint64_t items = 0;
        const uint64_t item_children = (uint64_t)env_int("FOCK_BLK_ITEM", 16384);
        for (int wl = 0; wl <= k; ++wl) {
            const int ur = k - wl;   // photons right of the lead modes
            const uint64_t nlam = fock_count(pl, wl), blk = fock_count(m - pl, ur);
            uint64_t lo = 0, hi = nlam;
            if (!full) {
                uint64_t l = 0, h = nlam;
                while (l < h) { uint64_t mid = (l + h) / 2; if (slos_host_prefix_base(m, pl, wl, ur, mid) + blk > cb) h = mid; else l = mid + 1; }
                lo = l;
                l = lo; h = nlam;
                while (l < h) { uint64_t mid = (l + h) / 2; if (slos_host_prefix_base(m, pl, wl, ur, mid) >= ce) h = mid; else l = mid + 1; }
                hi = l;
            }
            if (hi <= lo) continue;
            BlkGroup g;
            memset(&g, 0, sizeof g);
            g.wl = wl; g.u_rest = ur;
            g.seg_begin = (int)segs.size();
            g.cut_off = (uint32_t)cuts.size();
            // children of one lead prefix that this kernel takes -> lead prefixes per batch
            uint64_t per_lam_children = 0;
            for (int v = 0; v <= ur; ++v) {
                const int u = ur - v;
                if (u >= ts->u_limit) continue;
                per_lam_children += fock_count(pg, v) * fock_count(D, u);
            }
            if (per_lam_children == 0) continue;
            uint64_t Lb = per_lam_children >= item_children ? 1 : (item_children + per_lam_children - 1) / per_lam_children;
            if (Lb > hi - lo) Lb = hi - lo;
            if (Lb > 65536) Lb = 65536;
            uint64_t upl = 0, acc = 0;
            std::vector<uint32_t> mycuts;
            mycuts.push_back(0);
            for (int v = 0; v <= ur; ++v) {
                const int u = ur - v;
                if (u >= ts->u_limit) continue;
                const uint64_t Rv = fock_count(pg, v);
                if (Rv == 0) continue;
                const uint64_t pairs = Lb * Rv;
                for (int cid : ts->cids[u]) {
                    const BlkCls &cl = ts->cls[cid];
                    BlkSeg sg;
                    sg.v = v; sg.cid = cid;
                    sg.begin_local = (uint32_t)upl;
                    sg.nranges = (uint32_t)((pairs + cl.per_item - 1) / cl.per_item);
                    for (uint32_t r = 0; r < sg.nranges; ++r) {
                        const uint64_t np = (pairs - (uint64_t)r * cl.per_item) < cl.per_item ? (pairs - (uint64_t)r * cl.per_item) : cl.per_item;
                        acc += np * cl.S + 512;   // + fixed cost of a unit, in child-equivalents
                        ++upl;
                        if (acc >= item_children) { mycuts.push_back((uint32_t)upl); acc = 0; }
                    }
                    segs.push_back(sg);
                }
            }
            g.nseg = (int)segs.size() - g.seg_begin;
            if (g.nseg == 0) continue;
            FOCK_REQUIRE(upl < (1ull << 31), FOCK_ERR_LIMIT, "slos: too many units per batch of lead prefixes");
            if (mycuts.back() != (uint32_t)upl) mycuts.push_back((uint32_t)upl);
            g.upl = (uint32_t)upl;
            g.lam_lo = lo; g.lam_hi = hi;
            g.item_begin = items;
            g.ncut = (uint32_t)mycuts.size() - 1;
            g.Lb = (uint32_t)Lb;
            cuts.insert(cuts.end(), mycuts.begin(), mycuts.end());
            items += (((hi - lo) + Lb - 1) / Lb) * g.ncut;
            plan.groups.push_back(g);
        }
        FOCK_REQUIRE(plan.groups.size() <= BLK_MAXGRP, FOCK_ERR_LIMIT, "slos: too many lead-weight groups");
        FOCK_REQUIRE(items < (1ull << 31), FOCK_ERR_LIMIT, "slos: too many work items");
        plan.items = items;
        if (!segs.empty()) {
            FOCK_CUDA(cudaMalloc(&plan.d_segs, segs.size() * sizeof(BlkSeg)));
            FOCK_CUDA(cudaMemcpy(plan.d_segs, segs.data(), segs.size() * sizeof(BlkSeg), cudaMemcpyHostToDevice));
        }
        if (cuts.empty()) cuts.push_back(0);
        FOCK_CUDA(cudaMalloc(&plan.d_cuts, cuts.size() * sizeof(uint32_t)));
        FOCK_CUDA(cudaMemcpy(plan.d_cuts, cuts.data(), cuts.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        if (S->plans.size() > 256) {   // bounded cache
            for (auto &kv : S->plans) { cudaFree(kv.second.d_segs); cudaFree(kv.second.d_cuts); }
            S->plans.clear();
        }
        pit = S->plans.emplace(key, std::move(plan)).first;
    }
    const BlkPlan &plan = pit->second;
    if (plan.items == 0) return FOCK_OK;
    BlkArgs a;
    memset(&a, 0, sizeof a);
    a.m = m; a.k = k; a.mk = mk; a.p = p; a.pl = pl; a.D = D;
    a.maxnz = p < k ? p : k;
    if (a.maxnz < 1) a.maxnz = 1;
    a.bt = c->d_bt; a.dt = c->d_dt;
    a.U = (const double2 *)d_U;
    a.parent = (const double2 *)d_parent;
    a.child = (double2 *)d_child;
    a.probs = d_probs;
    a.sum = d_sum;
    a.inv_in_fact = 1.0 / in_prodnfact;
    a.cbegin = cb; a.cend = ce;
    a.cls = ts->d_cls;
    a.segs = plan.d_segs;
    a.cuts = plan.d_cuts;
    a.tab = ts->d_tab;
    a.tfact = ts->d_tfact;
    a.ngrp = (int)plan.groups.size();
    for (int i = 0; i < a.ngrp; ++i) a.grp[i] = plan.groups[i];
    const size_t smem = blk_smem_layout(m, a.maxnz).fixed + ts->max_ring_bytes;
    const bool wc = d_child != nullptr, wp = d_probs != nullptr;
    if (wc && wp) return blk_launch<3>(c, a, !full, (unsigned)plan.items, smem, st);
    if (wp) return blk_launch<2>(c, a, !full, (unsigned)plan.items, smem, st);
    return blk_launch<1>(c, a, !full, (unsigned)plan.items, smem, st);
}
