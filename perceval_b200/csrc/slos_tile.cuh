// slos_tile.cuh -- work-plan structures shared by the SLOS tile kernels (slos.cu: tile kernel, slos_thin.cu: hybrid thin kernel).
#pragma once
#include "common.cuh"

#define TILE_BLOCK 256

struct TileClass {
    int w, u;
    uint32_t S, G, nchunks, pad;
    uint64_t rho_lo, np;      // prefix ranks [rho_lo, rho_lo + np) of FS(p, w) intersect the child range
    uint64_t item_begin;      // first work item (CTA index) of this class
    uint64_t per_item;        // prefixes per work item
    // slab-major layouts (TileArgs::slab != 0): element offsets with  index(w, rho, t) = off + rho * S + t
    uint64_t coff;            // child slab w          (block size S)
    uint64_t roff;            // parent slab w - 1     (block size S: the aligned rows of the prefix modes)
    uint64_t toff;            // parent slab w         (block size Sp: the tail-parent block of the same prefix)
    uint64_t Sp;              // |FS(D, u - 1)|
};

struct TileArgs {
    int m, k, mk, p, ncls, maxnz;
    const uint64_t *bt, *dt;
    const double2 *U;
    const double2 *parent;
    uint64_t pbegin, pend;     // resident parent ranks [pbegin, pend) ...
    uint64_t gap_b, gap_e;     // ... except the hole [gap_b, gap_e): ranks >= gap_e are stored (gap_e - gap_b) elements earlier
    double2 *child;
    double *probs;
    double *sum;
    double inv_in_fact;
    uint64_t cbegin, cend;
    int *status;
    int ustride, urow0;        // column mk of the unitary = U[(urow0 + i) * ustride + mk], i < m (a sub-layer on the tail modes reads rows p..)
    int slab;                  // 0: parent and child in FSArray rank order; 1: slab-major (offsets in TileClass), see slos_layer_slab
    int uslot;                 // thin kernel: which constant-bank copy of the unitary column this launch reads
    TileClass cls[FOCK_TMAX];
    const uint64_t *tup[FOCK_TMAX];   // per class, the occupation tuple (4 bits / tail mode) of every tail rank
};

struct __align__(16) TileDesc {
    uint64_t cbase, tbase;
    double pfact;
    int nz, pad;
};

