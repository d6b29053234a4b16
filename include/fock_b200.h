/*
 * fock_b200.h -- C ABI of the B200-native Fock-amplitude engine (libfock_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of Quandela/Perceval that this repository accelerates: the
 * arithmetic that Perceval's strong / weak simulation backends delegate to the closed `exqalibur` C++ wheel.
 * The reference has no C ABI of its own (its boundary is exqalibur's private pybind API); every entry point
 * below names the reference call site (path relative to the Perceval tree, file:line) that it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / pybind types.  `double2`-style complex = interleaved (re, im)
 *     doubles, passed as `const double*`.  Fock states are uint8 occupation arrays of length m.
 *   - pointers named d_* are DEVICE pointers (caller-owned, e.g. torch tensors); h_* are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous on
 *     that stream unless stated otherwise.
 *   - every function returns 0 on success, <0 on error; fock_last_error() returns the message (thread local).
 *   - FSArray order = descending lexicographic on the occupation tuple (|n,0,..,0> has rank 0).
 *   - limits: SLOS / rank kernels m <= 64, n <= 32; permanents n <= 32; sampler n <= 32, any m.
 */
#ifndef FOCK_B200_H
#define FOCK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fock_ctx fock_ctx;

#define FOCK_OK 0
#define FOCK_ERR_ARG (-1)
#define FOCK_ERR_CUDA (-2)
#define FOCK_ERR_LIMIT (-3)

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
int fock_create(int device, fock_ctx **out);
int fock_destroy(fock_ctx *ctx);
const char *fock_last_error(void);
const char *fock_version(void);
int fock_device_info(fock_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);
/* synchronises `stream`, returns <0 if a kernel flagged an error since the last call (e.g. a SLOS parent rank
 * outside the resident window), and clears the flag */
int fock_check_status(fock_ctx *ctx, void *stream);

/* ---- FSArray: count / rank / unrank  (replaces xq.FSArray(m,n).count()/.find()/iteration,
 *      perceval/backends/_slos.py:156-168,190 ; perceval/utils/states.py:255-298) -------------------------- */
uint64_t fock_count(int m, int n); /* C(n+m-1, n); UINT64_MAX on overflow */
int fock_rank_host(int m, int n, const uint8_t *h_states, uint64_t cnt, uint64_t *h_ranks);
int fock_unrank_host(int m, int n, const uint64_t *h_ranks, uint64_t cnt, uint8_t *h_states);
int fock_rank(fock_ctx *ctx, int m, int n, const uint8_t *d_states, uint64_t cnt, uint64_t *d_ranks, void *stream);
int fock_unrank(fock_ctx *ctx, int m, int n, const uint64_t *d_ranks, uint64_t cnt, uint8_t *d_states, void *stream);
/* states [begin,end) of FSArray(m,n) in order, written as (end-begin) x m uint8 */
int fock_enumerate(fock_ctx *ctx, int m, int n, uint64_t begin, uint64_t end, uint8_t *d_states, void *stream);

/* ---- FSMask  (replaces xq.FSMask(m, n, masks[, at_least_modes]).match(state[, allow_missing]),
 *      perceval/backends/_abstract_backends.py:130-137 ; perceval/simulators/simulator.py:650-662 ;
 *      tests/utils/test_mask.py:32-45) -------------------------------------------------------------------- */
/* h_conds: nmask x m int8, -1 = any count, v >= 0 = exactly v photons (">= v" on modes whose bit is set in
 * at_least_bits); a state matches if it matches any mask.  allow_missing: 0 = exact match; 1 = the partial match of
 * intermediate layers (a mode passes if its count can still grow into the condition); 2 + b = partial match that can still
 * be completed with b more photons (what xq.FSArray(m, k, mask) keeps on layer k = n - b, _slos.py:156-166).  d_flags[i] = 1 iff state #(begin+i) of FSArray(m,n) matches.  Synchronises `stream`. */
int fock_mask_match(fock_ctx *ctx, int m, int n, const int8_t *h_conds, int nmask, uint64_t at_least_bits, int allow_missing,
                    uint64_t begin, uint64_t end, uint8_t *d_flags, void *stream);
int fock_mask_match_host(int m, int n, const int8_t *h_conds, int nmask, uint64_t at_least_bits, int allow_missing,
                         const uint8_t *h_states, uint64_t cnt, uint8_t *h_flags);

/* ---- SLOS  (replaces xq.FSMap(fsa_k, fsa_km1, True).compute_slos_layer(u, m, mk, coefs, parent_coefs),
 *      perceval/backends/_slos.py:99 ; python twin :91-97) ------------------------------------------------- */
/* One layer: child[s] = sum_{j: s_j>0} U[j,mk] * parent[s - e_j] for child ranks [child_begin, child_end) of
 * FSArray(m,k).  d_U: m*m row-major complex.  d_parent holds parent ranks [parent_begin, parent_end) of
 * FSArray(m,k-1) (pass 0, count(m,k-1) for a full layer); d_child receives child_end-child_begin values.
 * A parent outside the resident window is an error reported through *d_status (device int, may be NULL). */
int slos_layer(fock_ctx *ctx, int m, int k, const double *d_U, int mk, const double *d_parent, uint64_t parent_begin,
               uint64_t parent_end, double *d_child, uint64_t child_begin, uint64_t child_end, void *stream);

/* Last layer fused with the probability epilogue (replaces abs(c)**2/prodnfact + xq.all_prob_normalize_output,
 * perceval/backends/_slos.py:197-199,211-213): writes p[s] = |c_n[s]|^2 * prod(s_i!) / in_prodnfact for the child
 * range, optionally the coefficients too (d_child may be NULL), and atomically adds sum(p) to *d_sum (may be NULL). */
int slos_layer_probs(fock_ctx *ctx, int m, int k, const double *d_U, int mk, const double *d_parent,
                     uint64_t parent_begin, uint64_t parent_end, double *d_child, double *d_probs, double *d_sum,
                     double in_prodnfact, uint64_t child_begin, uint64_t child_end, void *stream);

/* Same two calls with a SEGMENTED resident parent: h_parent_seg = {b0, e0, b1, e1}, e0 <= b1, ranks [b0,e0) then
 * [b1,e1) stored packed at d_parent (b1 == e1: one segment).  The parents a contiguous child range needs through one
 * mode are one contiguous range (perceval/backends/_slos.py:91-97 read as a gather), their union over the modes one or
 * two ranges -- the recompute-window partition of perceval_b200/dist.py keeps only those resident. */
int slos_layer_seg(fock_ctx *ctx, int m, int k, const double *d_U, int mk, const double *d_parent, const uint64_t *h_parent_seg,
                   double *d_child, uint64_t child_begin, uint64_t child_end, void *stream);
int slos_layer_probs_seg(fock_ctx *ctx, int m, int k, const double *d_U, int mk, const double *d_parent,
                         const uint64_t *h_parent_seg, double *d_child, double *d_probs, double *d_sum, double in_prodnfact,
                         uint64_t child_begin, uint64_t child_end, void *stream);

/* One layer in SLAB-MAJOR layout -- the layout of the multi-GPU slab partition (no reference counterpart: the reference keeps
 * every layer whole in host memory, _slos.py:44).  The modes are split as in the tile kernels into a prefix of p modes and a
 * tail of m - p; a layer is stored by prefix weight w ("slab"), then prefix rank rho in FSArray(p, w), then tail rank t in
 * FSArray(m - p, photons - w):  index = slab_off[w] + rho * |FSArray(m - p, photons - w)| + t.  h_rho_ranges holds, per
 * w = 0..k, the prefixes [lo, hi) of the child layer to compute (2 (k+1) entries); h_parent_slab_off the element offset of
 * parent slab w' (k entries, relative to d_parent); h_child_slab_off that of child slab w (k+1 entries, relative to d_child /
 * d_probs; offsets are taken modulo 2^64, so compact per-rank layouts work too).  Same gather as slos_layer; d_child and/or
 * d_probs (+ d_sum) as in slos_layer_probs.  p must be the prefix width the library uses for m modes (m - 16 for m >= 20). */
int slos_layer_slab(fock_ctx *ctx, int m, int k, int p, const double *d_U, int mk, const double *d_parent, double *d_child,
                    double *d_probs, double *d_sum, double in_prodnfact, const uint64_t *h_rho_ranges,
                    const uint64_t *h_parent_slab_off, const uint64_t *h_child_slab_off, void *stream);

/* One layer over a PRUNED rank space (masks / heralds; replaces the layers the reference builds on xq.FSArray(m, k, mask),
 * perceval/backends/_slos.py:156-166): d_child_ranks / d_parent_ranks are the ascending kept ranks of FSArray(m, k) /
 * FSArray(m, k-1), d_parent the packed parent coefficients in that order.  Writes, per kept child, any of: the packed
 * coefficient (d_child), the probability |c|^2 prod(s!)/in_prodnfact (d_probs, + atomic sum into d_sum), the amplitude
 * c sqrt(prod(s!)/in_prodnfact) (d_amps) -- NULL skips an output.  A kept child whose parent is not in the parent list is an
 * error reported through fock_check_status. */
int slos_layer_masked(fock_ctx *ctx, int m, int k, const double *d_U, int mk, const uint64_t *d_parent_ranks, uint64_t n_parent,
                      const double *d_parent, const uint64_t *d_child_ranks, uint64_t n_child, double *d_child, double *d_probs,
                      double *d_amps, double *d_sum, double in_prodnfact, void *stream);

/* Stand-alone epilogues on an existing coefficient range [begin,end) of FSArray(m,n). */
int slos_probs_epilogue(fock_ctx *ctx, int m, int n, const double *d_coefs, double in_prodnfact, double *d_probs,
                        double *d_sum, uint64_t begin, uint64_t end, void *stream);
/* amplitude = coef * sqrt(prod(s!)/in_prodnfact)   (perceval/backends/_slos.py:187-193, :216-223) */
int slos_amplitudes_epilogue(fock_ctx *ctx, int m, int n, const double *d_coefs, double in_prodnfact, double *d_amps,
                             uint64_t begin, uint64_t end, void *stream);

/* Photon insertion order of a single input state (perceval/backends/_slos.py:61-86). h_order gets n ints. */
int slos_order(int m, const uint8_t *h_in_state, int *h_order);

/* Whole chain for one input state with caller-provided ping-pong device workspaces:
 * d_work_a must hold count(m,n-1) complex (0 if n==0), d_work_b count(m,n-2) complex; writes count(m,n) probs
 * (and coefficients if d_coefs != NULL).  Replaces SLOSBackend.set_input_state + prob_distribution/all_prob
 * (perceval/backends/_slos.py:143-145,195-214) for the single-input chain.  Asynchronous on `stream`. */
int slos_prob_distribution(fock_ctx *ctx, int m, const double *d_U, const uint8_t *h_in_state, double *d_work_a,
                           double *d_work_b, double *d_coefs, double *d_probs, double *d_sum, void *stream);
/* Same call with HOST buffers (U in, probabilities out); allocates device workspaces internally, synchronous. */
int slos_prob_distribution_host(fock_ctx *ctx, int m, const double *h_U, const uint8_t *h_in_state, double *h_probs,
                                double *h_sum);

/* ---- permanents / Naive backend (replaces xq.permanent_cx(M), perceval/backends/_naive.py:70-71) ---------- */
/* B matrices n x n (row-major complex, contiguous), Glynn Gray-code range [gray_begin, gray_end) of the
 * 2^(n-1) codes (0, 0 = whole range).  d_out[b] receives the (partial) permanent, already scaled by 2^(1-n). */
int glynn_permanent_batch(fock_ctx *ctx, int n, const double *d_mats, uint64_t B, double *d_out, uint64_t gray_begin,
                          uint64_t gray_end, void *stream);
int glynn_permanent_batch_host(fock_ctx *ctx, int n, const double *h_mats, uint64_t B, double *h_out);
/* NaiveBackend.prob_amplitude for a batch of output states given by rank (perceval/backends/_naive.py:46-68):
 * builds the sub-matrices on device and evaluates perm / sqrt(prod(in!) prod(out!)); n == 1 returns M[0,0]. */
int naive_amplitudes(fock_ctx *ctx, int m, int n, const double *d_U, const uint8_t *h_in_state,
                     const uint64_t *d_out_ranks, uint64_t B, double *d_amps, void *stream);
int naive_amplitudes_states(fock_ctx *ctx, int m, int n, const double *d_U, const uint8_t *h_in_state,
                            const uint8_t *d_out_states, uint64_t B, double *d_amps, void *stream);

/* ---- Clifford & Clifford 2017 sampler (replaces xq.Clifford2017.set_unitary/.set_input_state/.sample/.samples,
 *      perceval/backends/_clifford2017.py:39-57) ------------------------------------------------------------- */
/* Samples [offset, offset+count) of the stream keyed by `seed` (Philox4x32-10, one key per sample index, so the
 * result does not depend on how a batch is split across calls or GPUs).  d_out_states: count x m uint8. */
int cc2017_samples(fock_ctx *ctx, int m, int n, const double *d_U, const uint8_t *h_in_state, uint64_t count,
                   uint64_t seed, uint64_t offset, uint8_t *d_out_states, void *stream);
int cc2017_samples_host(fock_ctx *ctx, int m, int n, const double *h_U, const uint8_t *h_in_state, uint64_t count,
                        uint64_t seed, uint64_t offset, uint8_t *h_out_states);

/* ---- measurement helpers (bench.py only) ------------------------------------------------------------------- */
/* kind 0: FP64 FMA peak (returns TFLOP/s), 1: HBM stream read+write copy GB/s, 2: L2-resident read GB/s */
int fock_measure_peak(fock_ctx *ctx, int kind, double *out_value);
/* cudaEvent_t pair (as void*) recorded on the launching stream immediately before / after every probability-layer launch
 * (slos_layer_probs, slos_layer_probs_seg, the last layer of slos_prob_distribution) until reset with (NULL, NULL): lets
 * a caller time the dominant kernel inside a whole-chain call with CUDA events, without a profiler */
int fock_profile_events(fock_ctx *ctx, void *ev_begin, void *ev_end);
/* number of kernel launches issued through this context since creation */
uint64_t fock_launch_count(fock_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FOCK_B200_H */
