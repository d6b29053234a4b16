/*
 * fock_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the Fock-amplitude hot path that Perceval delegates to the closed
 * `exqalibur~=1.1.0` wheel (reference setup.py:77; source not in /root/reference, wheel not installable
 * offline).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (perceval_b200/) never links or calls it.
 *
 * Parity status: the SLOS / Naive halves are pinned against every golden vector the reference tests hold
 * for this path (tests/test_oracle_golden.py replays reference tests/backends/test_backends.py:39-289,
 * docs naive.rst / slos.rst, the Boson_Bunching notebook known answers).  Clifford&Clifford sample
 * *sequences* are "parity unpinned" (exqalibur's RNG stream is unknowable); only the sampled distribution
 * is pinned (exact pmf by enumeration vs SLOS, see tests/test_cc2017_oracle.py).  No reference test stores an
 * amplitude beyond n = 8 or a permanent of a Haar sub-matrix: at the BASELINE sizes (n = 24 .. 32) the Glynn walk is
 * "parity unpinned" by the reference and is instead arbitrated by the same walk in x87 extended precision
 * (orc_glynn_range_ld; tests/test_oracle_golden.py, tests/test_gpu_parity.py).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 *
 * Conventions
 *   - complex numbers are interleaved (re, im) doubles ("double2"), row-major U (m x m).
 *   - Fock states are uint8 occupation arrays of length m.
 *   - FSArray order = descending lexicographic on the occupation tuple, |n,0,..> first
 *     (pinned by reference tests/utils/test_statevector.py:430-438, tests/utils/test_density_matrix.py:54-59).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double _Complex cplx;

#define ORC_MAXM 512

/* ------------------------------------------------------------------ combinatorics */

/* exact C(n,k) in uint64 (multiplicative, exact at every step) */
uint64_t orc_binom(int n, int k)
{
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (unsigned)(n - k + i) / (unsigned)i;
    return (uint64_t)r;
}

/* xq.FSArray(m,n).count()  (reference perceval/backends/_slos.py:165,197) */
uint64_t orc_count(int m, int n)
{
    if (m <= 0) return n == 0 ? 1 : 0;
    return orc_binom(n + m - 1, n);
}

/* xq.FSArray.find(state) (reference _slos.py:190): index of `s` in descending-lex order.
 * Direct definition: number of states that precede s = for every mode i, every larger value v at
 * that mode with the same prefix, the number of completions on the remaining modes. */
uint64_t orc_rank(int m, int n, const uint8_t *s)
{
    uint64_t r = 0;
    int rem = n; /* photons available at mode i */
    for (int i = 0; i < m - 1; ++i) {
        for (int v = s[i] + 1; v <= rem; ++v) r += orc_count(m - i - 1, rem - v);
        rem -= s[i];
    }
    return r;
}

/* iteration of xq.FSArray (reference perceval/utils/states.py:284-285): state at index r */
void orc_unrank(int m, int n, uint64_t r, uint8_t *s)
{
    int rem = n;
    for (int i = 0; i < m - 1; ++i) {
        int v = rem;
        for (;; --v) {
            uint64_t c = orc_count(m - i - 1, rem - v);
            if (r < c) break;
            r -= c;
        }
        s[i] = (uint8_t)v;
        rem -= v;
    }
    s[m - 1] = (uint8_t)rem;
}

/* successor in descending-lex order; returns 0 when s was the last state |0,..,0,n> */
int orc_next(int m, uint8_t *s)
{
    int i = m - 2;
    while (i >= 0 && s[i] == 0) --i;
    if (i < 0) return 0;
    int tail = s[m - 1];
    s[m - 1] = 0;
    s[i] -= 1;
    s[i + 1] = (uint8_t)(tail + 1);
    return 1;
}

void orc_rank_batch(int m, int n, const uint8_t *states, uint64_t cnt, uint64_t *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)cnt; ++i) out[i] = orc_rank(m, n, states + (size_t)i * m);
}

void orc_unrank_batch(int m, int n, const uint64_t *ranks, uint64_t cnt, uint8_t *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)cnt; ++i) orc_unrank(m, n, ranks[i], out + (size_t)i * m);
}

/* all states of FSArray(m,n) in order, cnt*m bytes */
void orc_enumerate(int m, int n, uint8_t *out)
{
    uint8_t s[ORC_MAXM];
    memset(s, 0, sizeof s);
    s[0] = (uint8_t)n;
    uint64_t i = 0;
    do {
        memcpy(out + i * m, s, (size_t)m);
        ++i;
    } while (orc_next(m, s));
}

static double prodnfact(int m, const uint8_t *s)
{
    double p = 1.0;
    for (int i = 0; i < m; ++i)
        for (int v = 2; v <= s[i]; ++v) p *= v;
    return p;
}

double orc_prodnfact(int m, const uint8_t *s) { return prodnfact(m, s); }

/* ------------------------------------------------------------------ SLOS */

/* One SLOS layer, LITERAL scatter form of reference _slos.py:91-97
 *   coefs.fill(0); for parent_idx, coef_parent: for j in range(m):
 *       idx = fsm.get(parent_idx, j); coefs[idx] += coef_parent * u[j, mk]
 * k = photon count of the child layer.  Serial. */
void orc_slos_layer_scatter(int m, int k, const double *u, int mk, const double *parent, double *child)
{
    const cplx *U = (const cplx *)u;
    const cplx *P = (const cplx *)parent;
    cplx *C = (cplx *)child;
    uint64_t nc = orc_count(m, k), np = orc_count(m, k - 1);
    for (uint64_t i = 0; i < nc; ++i) C[i] = 0;
    uint8_t s[ORC_MAXM];
    memset(s, 0, sizeof s);
    s[0] = (uint8_t)(k - 1);
    for (uint64_t p = 0; p < np; ++p) {
        for (int j = 0; j < m; ++j) {
            s[j] += 1;
            uint64_t idx = orc_rank(m, k, s);
            C[idx] += P[p] * U[(size_t)j * m + mk];
            s[j] -= 1;
        }
        orc_next(m, s);
    }
}

/* Same layer in gather form (mathematically identical; used as the threaded CPU baseline):
 *   c_k[s] = sum_{j: s_j>0} U[j,mk] * c_{k-1}[s - e_j]            
 * computes children [begin,end).  Ranks use the collapsed (hockey-stick) form of SURVEY.md 8a row a1,
 *   rank(s) = sum_{i<m-1} C(T_i-1+q_i, q_i),  T_i = photons right of mode i, q_i = m-1-i,
 * through a per-call binomial table, so that the baseline is a fair multithreaded CPU implementation and not a
 * strawman; it is checked against the literal scatter loop above in tests/test_oracle_golden.py. */
void orc_slos_layer_gather(int m, int k, const double *u, int mk, const double *parent, double *child,
                           uint64_t begin, uint64_t end)
{
    const cplx *U = (const cplx *)u;
    const cplx *P = (const cplx *)parent;
    cplx *C = (cplx *)child;
    /* bt[q][T] = C(T-1+q, q) for T >= 1, 0 for T = 0 */
    int W = k + 1;
    uint64_t *bt = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m * (size_t)W);
    for (int q = 0; q < m; ++q)
        for (int T = 0; T <= k; ++T) bt[q * W + T] = T ? orc_binom(T - 1 + q, q) : 0;
    cplx ucol[ORC_MAXM];
    for (int j = 0; j < m; ++j) ucol[j] = U[(size_t)j * m + mk];
#pragma omp parallel
    {
        int nt = 1, tid = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        tid = omp_get_thread_num();
#endif
        uint64_t tot = end - begin;
        uint64_t lo = begin + tot * (uint64_t)tid / (uint64_t)nt;
        uint64_t hi = begin + tot * (uint64_t)(tid + 1) / (uint64_t)nt;
        if (lo < hi) {
            uint8_t s[ORC_MAXM];
            orc_unrank(m, k, lo, s);
            for (uint64_t r = lo; r < hi; ++r) {
                cplx acc = 0;
                /* parent rank for mode j = r - sum_{i<j} (bt[q_i][T_i] - bt[q_i][T_i-1]) */
                uint64_t E = 0;
                int T = k;
                for (int j = 0; j < m; ++j) {
                    if (s[j]) acc += ucol[j] * P[r - E];
                    T -= s[j];
                    if (T == 0) break;
                    if (j < m - 1) E += bt[(m - 1 - j) * W + T] - bt[(m - 1 - j) * W + T - 1];
                }
                C[r - begin] = acc;
                orc_next(m, s);
            }
        }
    }
    free(bt);
}

/* Photon insertion order of a single input state: reference _slos.py:61-86 (_Path._decompose with one
 * target: repeatedly take the mode with the most remaining photons, first index on ties). out has n entries. */
void orc_slos_order(int m, const uint8_t *in_state, int *out)
{
    uint8_t t[ORC_MAXM];
    memcpy(t, in_state, (size_t)m);
    int n = 0;
    for (int i = 0; i < m; ++i) n += t[i];
    for (int k = 0; k < n; ++k) {
        int best = 0;
        for (int i = 1; i < m; ++i)
            if (t[i] > t[best]) best = i;
        out[k] = best;
        t[best] -= 1;
    }
}

/* Full SLOS run for one input state: un-normalised coefficients of the last layer
 * (what _Path.coefs holds, reference _slos.py:44,88-102).  coefs must hold count(m,n) complex.
 * use_scatter=1 follows the literal reference loop; 0 uses the threaded gather.  */
int orc_slos_coefs(int m, const double *u, const uint8_t *in_state, double *coefs, int use_scatter)
{
    int n = 0;
    for (int i = 0; i < m; ++i) n += in_state[i];
    int *order = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    orc_slos_order(m, in_state, order);
    double *prev = (double *)malloc(16);
    prev[0] = 1.0; /* _slos.py:45-46: layer 0 is filled with 1 */
    prev[1] = 0.0;
    for (int k = 1; k <= n; ++k) {
        uint64_t nc = orc_count(m, k);
        double *cur = (k == n) ? coefs : (double *)malloc(16 * (size_t)nc);
        if (!cur) return -1;
        if (use_scatter)
            orc_slos_layer_scatter(m, k, u, order[k - 1], prev, cur);
        else
            orc_slos_layer_gather(m, k, u, order[k - 1], prev, cur, 0, nc);
        free(prev);
        prev = cur;
    }
    if (n == 0) {
        coefs[0] = 1.0;
        coefs[1] = 0.0;
        free(prev);
    }
    free(order);
    return 0;
}

/* reference _slos.py:195-199 / :205-214: p[i] = |c[i]|^2 / prod(in!) ; then
 * xq.all_prob_normalize_output(c, fsa): p[i] *= prod(s_i!)  (definition inferred from _slos.py:192) */
void orc_slos_probs(int m, int n, const double *coefs, double in_prodnfact, double *probs)
{
    uint64_t N = orc_count(m, n);
#pragma omp parallel
    {
        int nt = 1, tid = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        tid = omp_get_thread_num();
#endif
        uint64_t lo = N * (uint64_t)tid / (uint64_t)nt, hi = N * (uint64_t)(tid + 1) / (uint64_t)nt;
        if (lo < hi) {
            uint8_t s[ORC_MAXM];
            orc_unrank(m, n, lo, s);
            for (uint64_t r = lo; r < hi; ++r) {
                double re = coefs[2 * r], im = coefs[2 * r + 1];
                /* numpy abs(c)**2 = hypot squared; re*re+im*im differs by <= 1 ulp, tests use 1e-10 rel */
                probs[r] = (re * re + im * im) / in_prodnfact * prodnfact(m, s);
                orc_next(m, s);
            }
        }
    }
}

/* reference _slos.py:187-193: amplitude = coefs[idx] * sqrt(prod(out!)/prod(in!)) for every state */
void orc_slos_amplitudes(int m, int n, const double *coefs, double in_prodnfact, double *amps)
{
    uint64_t N = orc_count(m, n);
    uint8_t s[ORC_MAXM];
    memset(s, 0, sizeof s);
    s[0] = (uint8_t)n;
    for (uint64_t r = 0; r < N; ++r) {
        double f = sqrt(prodnfact(m, s) / in_prodnfact);
        amps[2 * r] = coefs[2 * r] * f;
        amps[2 * r + 1] = coefs[2 * r + 1] * f;
        orc_next(m, s);
    }
}

/* ------------------------------------------------------------------ permanent (Naive backend) */

/* xq.permanent_cx(M) (reference _naive.py:70-71).  Glynn formula (references.bib:905-914) with a Gray-code
 * walk over delta in {+-1}^n, delta_0 = +1:
 *   perm(M) = 2^{1-n} * sum_delta (prod_k delta_k) * prod_j ( sum_i delta_i M[i,j] )
 * Gray range [g0, g1) of the 2^{n-1} codes is exposed so the range split of SURVEY 8(e) can be checked. */
/* The n column sums drift by one rounding per Gray step; over the millions of steps of a CPU chunk that drift (not the
 * Kahan-compensated outer sum) would dominate the error of a permanent that is 1e7 times smaller than its terms (Haar
 * sub-matrices at n = 30), so the sums are re-seeded exactly from the Gray code every ORC_RESEED steps. */
#define ORC_RESEED 4096
static void glynn_seed(int n, const cplx *M, uint64_t gray, cplx *v)
{
    for (int j = 0; j < n; ++j) {
        cplx acc = M[j]; /* row 0, delta_0 = +1 */
        for (int i = 1; i < n; ++i) acc += ((gray >> (i - 1)) & 1) ? -M[(size_t)i * n + j] : M[(size_t)i * n + j];
        v[j] = acc;
    }
}

static cplx glynn_range(int n, const cplx *M, uint64_t g0, uint64_t g1)
{
    cplx v[64];
    /* delta for code g: bit b of gray(g) set => delta_{b+1} = -1 */
    uint64_t gray = g0 ^ (g0 >> 1);
    glynn_seed(n, M, gray, v);
    int sign = (__builtin_popcountll(gray) & 1) ? -1 : 1;
    cplx total = 0, comp = 0; /* Kahan on the outer sum */
    for (uint64_t g = g0; g < g1; ++g) {
        cplx prod = v[0];
        for (int j = 1; j < n; ++j) prod *= v[j];
        cplx term = (sign > 0 ? prod : -prod) - comp;
        cplx t = total + term;
        comp = (t - total) - term;
        total = t;
        /* next code: bit that flips between gray(g) and gray(g+1) = ctz(g+1) */
        uint64_t gn = g + 1;
        if (gn < g1) {
            uint64_t ngray = gn ^ (gn >> 1);
            if ((gn & (ORC_RESEED - 1)) == 0) {
                glynn_seed(n, M, ngray, v);
            } else {
                int b = __builtin_ctzll(gn);
                int now_minus = (int)((ngray >> b) & 1);
                const cplx *row = M + (size_t)(b + 1) * n;
                if (now_minus)
                    for (int j = 0; j < n; ++j) v[j] -= 2.0 * row[j];
                else
                    for (int j = 0; j < n; ++j) v[j] += 2.0 * row[j];
            }
            sign = -sign;
        }
    }
    return total;
}

/* The same walk in x87 extended precision (64-bit mantissa): the accuracy arbiter between this oracle and the device
 * kernels at the sizes where a double-precision Glynn sum is within a few units of the 1e-10 tolerance (n >= 28). */
typedef long double _Complex lcplx;
static lcplx glynn_range_ld(int n, const cplx *M, uint64_t g0, uint64_t g1)
{
    lcplx v[64];
    uint64_t gray = g0 ^ (g0 >> 1);
    int sign = (__builtin_popcountll(gray) & 1) ? -1 : 1;
    lcplx total = 0;
    for (uint64_t g = g0; g < g1; ++g) {
        if (g == g0 || (g & (ORC_RESEED - 1)) == 0) {
            uint64_t gr = g ^ (g >> 1);
            for (int j = 0; j < n; ++j) {
                lcplx acc = (lcplx)M[j];
                for (int i = 1; i < n; ++i) acc += ((gr >> (i - 1)) & 1) ? -(lcplx)M[(size_t)i * n + j] : (lcplx)M[(size_t)i * n + j];
                v[j] = acc;
            }
        }
        lcplx prod = v[0];
        for (int j = 1; j < n; ++j) prod *= v[j];
        total += (sign > 0 ? prod : -prod);
        uint64_t gn = g + 1;
        if (gn < g1 && (gn & (ORC_RESEED - 1)) != 0) {
            int b = __builtin_ctzll(gn);
            uint64_t ngray = gn ^ (gn >> 1);
            const cplx *row = M + (size_t)(b + 1) * n;
            if ((ngray >> b) & 1)
                for (int j = 0; j < n; ++j) v[j] -= 2.0L * (lcplx)row[j];
            else
                for (int j = 0; j < n; ++j) v[j] += 2.0L * (lcplx)row[j];
        }
        sign = -sign;
    }
    return total;
}

void orc_glynn_range_ld(int n, const double *mat, uint64_t g0, uint64_t g1, double *out)
{
    const cplx *M = (const cplx *)mat;
    if (n == 0) { out[0] = 1; out[1] = 0; return; }
    long double sr = 0, si = 0;
    uint64_t tot = g1 - g0;
    int chunks = 1;
#ifdef _OPENMP
    chunks = omp_get_max_threads() * 8;
#endif
    if ((uint64_t)chunks > tot) chunks = (int)(tot ? tot : 1);
#pragma omp parallel for schedule(dynamic) reduction(+ : sr, si)
    for (int c = 0; c < chunks; ++c) {
        uint64_t lo = g0 + tot * (uint64_t)c / (uint64_t)chunks, hi = g0 + tot * (uint64_t)(c + 1) / (uint64_t)chunks;
        lcplx r = glynn_range_ld(n, M, lo, hi);
        sr += creall(r);
        si += cimagl(r);
    }
    long double scale = ldexpl(1.0L, 1 - n);
    out[0] = (double)(sr * scale);
    out[1] = (double)(si * scale);
}

/* out = [re, im].  Threaded over Gray chunks. */
void orc_glynn_range(int n, const double *mat, uint64_t g0, uint64_t g1, double *out)
{
    const cplx *M = (const cplx *)mat;
    if (n == 0) { out[0] = 1; out[1] = 0; return; }
    double sr = 0, si = 0;
    uint64_t tot = g1 - g0;
    int chunks = 1;
#ifdef _OPENMP
    chunks = omp_get_max_threads() * 8;
#endif
    if ((uint64_t)chunks > tot) chunks = (int)(tot ? tot : 1);
#pragma omp parallel for schedule(dynamic) reduction(+ : sr, si)
    for (int c = 0; c < chunks; ++c) {
        uint64_t lo = g0 + tot * (uint64_t)c / (uint64_t)chunks, hi = g0 + tot * (uint64_t)(c + 1) / (uint64_t)chunks;
        cplx r = glynn_range(n, M, lo, hi);
        sr += creal(r);
        si += cimag(r);
    }
    double scale = ldexp(1.0, 1 - n);
    out[0] = sr * scale;
    out[1] = si * scale;
}

void orc_permanent(int n, const double *mat, double *out)
{
    if (n == 0) { out[0] = 1; out[1] = 0; return; }
    orc_glynn_range(n, mat, 0, (uint64_t)1 << (n - 1), out);
}

/* Ryser formula, independent cross-check of Glynn (n <= 20), serial */
void orc_permanent_ryser(int n, const double *mat, double *out)
{
    const cplx *M = (const cplx *)mat;
    cplx total = 0;
    for (uint64_t S = 1; S < ((uint64_t)1 << n); ++S) {
        cplx prod = 1;
        for (int i = 0; i < n; ++i) {
            cplx rs = 0;
            for (int j = 0; j < n; ++j)
                if ((S >> j) & 1) rs += M[(size_t)i * n + j];
            prod *= rs;
        }
        int bits = __builtin_popcountll(S);
        total += ((n - bits) & 1) ? -prod : prod;
    }
    out[0] = creal(total);
    out[1] = cimag(total);
}

/* reference _naive.py:51-68 (_compute_submatrix): rows = output photons, cols = input photons,
 * M[r,c] = U[out_mode(r), in_mode(c)], modes ascending, repeated per occupancy. */
void orc_naive_submatrix(int m, int n, const double *u, const uint8_t *in_state, const uint8_t *out_state, double *mat)
{
    const cplx *U = (const cplx *)u;
    cplx *M = (cplx *)mat;
    int col = 0;
    for (int ik = 0; ik < m; ++ik)
        for (int a = 0; a < in_state[ik]; ++a) {
            int row = 0;
            for (int ok = 0; ok < m; ++ok)
                for (int b = 0; b < out_state[ok]; ++b) {
                    M[(size_t)row * n + col] = U[(size_t)ok * m + ik];
                    ++row;
                }
            ++col;
        }
}

/* reference _naive.py:46-49: amplitude = perm(M)/sqrt(prod(in!) prod(out!)); n=1 returns M[0,0];
 * photon-number mismatch -> 0 ; n=0 -> 1 */
void orc_naive_amplitude(int m, const double *u, const uint8_t *in_state, const uint8_t *out_state, double *out)
{
    int n = 0, no = 0;
    for (int i = 0; i < m; ++i) { n += in_state[i]; no += out_state[i]; }
    if (n != no) { out[0] = out[1] = 0; return; }
    if (n == 0) { out[0] = 1; out[1] = 0; return; }
    double *mat = (double *)malloc(16 * (size_t)n * n);
    orc_naive_submatrix(m, n, u, in_state, out_state, mat);
    if (n == 1) { out[0] = mat[0]; out[1] = mat[1]; free(mat); return; }
    orc_permanent(n, mat, out);
    double p = sqrt(prodnfact(m, in_state) * prodnfact(m, out_state));
    out[0] /= p;
    out[1] /= p;
    free(mat);
}

/* ------------------------------------------------------------------ Clifford & Clifford 2017 */

/* Philox4x32-10 counter RNG (Salmon et al. 2011) -- the engine's device kernels use the same generator
 * with the same (seed, sample index) keying, so oracle and device draw identical uniforms. */
static inline void philox_round(uint32_t c[4], uint32_t k[2])
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
static void philox4x32_10(uint64_t seed, uint64_t ctr_hi, uint64_t ctr_lo, uint32_t out[4])
{
    uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, 16);
}
/* d-th uniform double in [0,1) of sample `idx`: 53 bits from two 32-bit words of block d/2 */
double orc_uniform(uint64_t seed, uint64_t idx, uint32_t d)
{
    uint32_t w[4];
    philox4x32_10(seed, idx, (uint64_t)(d >> 1), w);
    uint32_t a = w[(d & 1) * 2], b = w[(d & 1) * 2 + 1];
    uint64_t bits = (((uint64_t)a << 32) | b) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

/* all k leave-one-column-out permanents of the (k-1) x k matrix B (rows r_1..r_{k-1}, columns alpha_1..alpha_k)
 * in one Glynn sweep ("adapted Glynn ... n simultaneous sub-permanents", reference docs/source/backends.rst:114-115).
 * B is row-major (k-1) x k.  out[l] = Per(B without column l). */
static void subperms(int k, const cplx *B, cplx *out)
{
    int r = k - 1;
    for (int l = 0; l < k; ++l) out[l] = 0;
    if (r == 0) { out[0] = 1; return; }
    cplx v[64], pre[65], suf[65];
    uint64_t ncodes = (uint64_t)1 << (r - 1);
    for (uint64_t g = 0; g < ncodes; ++g) {
        uint64_t gray = g ^ (g >> 1);
        for (int c = 0; c < k; ++c) {
            cplx acc = B[c];
            for (int i = 1; i < r; ++i) acc += ((gray >> (i - 1)) & 1) ? -B[(size_t)i * k + c] : B[(size_t)i * k + c];
            v[c] = acc;
        }
        double sign = (__builtin_popcountll(gray) & 1) ? -1.0 : 1.0;
        pre[0] = 1;
        for (int c = 0; c < k; ++c) pre[c + 1] = pre[c] * v[c];
        suf[k] = 1;
        for (int c = k - 1; c >= 0; --c) suf[c] = suf[c + 1] * v[c];
        for (int l = 0; l < k; ++l) out[l] += sign * pre[l] * suf[l + 1];
    }
    double scale = ldexp(1.0, 1 - r);
    for (int l = 0; l < k; ++l) out[l] *= scale;
}

/* One sample of Clifford & Clifford (2018) Algorithm A  (reference docs/source/backends.rst:108-115;
 * call site perceval/backends/_clifford2017.py:39-57):
 *   A = columns of U for the input photons, randomly permuted (Fisher-Yates, uniforms 0..n-2);
 *   r_1 ~ |A[i,1]|^2 ; for k=2..n: w_i = |Per(A[(r_1..r_{k-1}, i), 1..k])|^2 by Laplace expansion on the
 *   new row with the k sub-permanents computed together; r_k ~ w (uniform n-1+k-1 ...).
 * Output = occupation array of the sorted multiset.
 * Uniform draw d of sample idx: d in [0,n-1) permutation, d = n-1+k-1 for the k-th row choice. */
void orc_cc2017_sample(int m, int n, const double *u, const uint8_t *in_state, uint64_t seed, uint64_t idx,
                       uint8_t *out_state)
{
    const cplx *U = (const cplx *)u;
    memset(out_state, 0, (size_t)m);
    if (n == 0) return;
    int cols[64], rows[64];
    int c = 0;
    for (int ik = 0; ik < m; ++ik)
        for (int a = 0; a < in_state[ik]; ++a) cols[c++] = ik;
    for (int i = 0; i < n - 1; ++i) { /* Fisher-Yates: swap i with i + floor(u*(n-i)) */
        int j = i + (int)(orc_uniform(seed, idx, (uint32_t)i) * (double)(n - i));
        if (j > n - 1) j = n - 1;
        int t = cols[i]; cols[i] = cols[j]; cols[j] = t;
    }
    double *w = (double *)malloc(sizeof(double) * (size_t)m);
    cplx B[64 * 64], sp[64];
    for (int k = 1; k <= n; ++k) {
        /* B = A[(r_1..r_{k-1}), alpha_1..alpha_k] */
        for (int i = 0; i < k - 1; ++i)
            for (int l = 0; l < k; ++l) B[(size_t)i * k + l] = U[(size_t)rows[i] * m + cols[l]];
        subperms(k, B, sp);
        double tot = 0;
        for (int i = 0; i < m; ++i) {
            cplx acc = 0;
            for (int l = 0; l < k; ++l) acc += U[(size_t)i * m + cols[l]] * sp[l];
            w[i] = creal(acc) * creal(acc) + cimag(acc) * cimag(acc);
            tot += w[i];
        }
        double x = orc_uniform(seed, idx, (uint32_t)(n - 1 + k - 1)) * tot;
        int pick = m - 1;
        double cum = 0;
        for (int i = 0; i < m; ++i) {
            cum += w[i];
            if (x < cum) { pick = i; break; }
        }
        while (pick > 0 && w[pick] == 0.0) --pick; /* never pick a zero-weight row through round-off */
        rows[k - 1] = pick;
    }
    for (int k = 0; k < n; ++k) out_state[rows[k]] += 1;
    free(w);
}

void orc_cc2017_samples(int m, int n, const double *u, const uint8_t *in_state, uint64_t count, uint64_t seed,
                        uint64_t offset, uint8_t *out_states)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)count; ++i)
        orc_cc2017_sample(m, n, u, in_state, seed, offset + (uint64_t)i, out_states + (size_t)i * m);
}

/* Exact output pmf of Algorithm A by enumerating every column permutation and every row sequence
 * (tiny sizes only: n <= 4, m <= 6).  pmf has count(m,n) entries in FSArray order.  Used to show the
 * sampler's law equals the SLOS distribution. */
static void cc_enum_rows(int m, int n, const cplx *U, const int *cols, int k, int *rows, double p, double *pmf)
{
    if (k > n) {
        uint8_t s[ORC_MAXM];
        memset(s, 0, (size_t)m);
        for (int i = 0; i < n; ++i) s[rows[i]] += 1;
        pmf[orc_rank(m, n, s)] += p;
        return;
    }
    cplx B[64], sp[8];
    for (int i = 0; i < k - 1; ++i)
        for (int l = 0; l < k; ++l) B[i * k + l] = U[(size_t)rows[i] * m + cols[l]];
    subperms(k, B, sp);
    double w[64], tot = 0;
    for (int i = 0; i < m; ++i) {
        cplx acc = 0;
        for (int l = 0; l < k; ++l) acc += U[(size_t)i * m + cols[l]] * sp[l];
        w[i] = creal(acc) * creal(acc) + cimag(acc) * cimag(acc);
        tot += w[i];
    }
    if (tot <= 0) return;
    for (int i = 0; i < m; ++i) {
        if (w[i] == 0) continue;
        rows[k - 1] = i;
        cc_enum_rows(m, n, U, cols, k + 1, rows, p * w[i] / tot, pmf);
    }
}
static void cc_enum_perm(int m, int n, const cplx *U, int *cols, int pos, double p, double *pmf)
{
    if (pos == n) {
        int rows[8];
        cc_enum_rows(m, n, U, cols, 1, rows, p, pmf);
        return;
    }
    for (int j = pos; j < n; ++j) {
        int t = cols[pos]; cols[pos] = cols[j]; cols[j] = t;
        cc_enum_perm(m, n, U, cols, pos + 1, p, pmf);
        t = cols[pos]; cols[pos] = cols[j]; cols[j] = t;
    }
}
void orc_cc2017_exact_pmf(int m, int n, const double *u, const uint8_t *in_state, double *pmf)
{
    uint64_t N = orc_count(m, n);
    for (uint64_t i = 0; i < N; ++i) pmf[i] = 0;
    int cols[8], c = 0;
    for (int ik = 0; ik < m; ++ik)
        for (int a = 0; a < in_state[ik]; ++a) cols[c++] = ik;
    double nperm = 1;
    for (int i = 2; i <= n; ++i) nperm *= i;
    cc_enum_perm(m, n, (const cplx *)u, cols, 0, 1.0 / nperm, pmf);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
