/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See plonky2_oracle.h for the status header.
 *
 * Plain C11 (+ OpenMP where the reference uses rayon) restatement of the reference's CPU
 * commit path.  Citations are reference-relative file:line.
 */
#include "plonky2_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "poseidon_constants.h"

typedef unsigned __int128 u128;
#define EPS 0xFFFFFFFFULL

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Goldilocks field, p = 2^64 - 2^32 + 1.  Values are raw u64, possibly >= p; compare canonical.
 * field/src/goldilocks_field.rs
 * ---------------------------------------------------------------------------------------- */

/* goldilocks_field.rs:221-228 */
uint64_t orc_gl_canon(uint64_t a) { return a >= ORC_P ? a - ORC_P : a; }

/* goldilocks_field.rs:249-265 */
static inline uint64_t gl_add(uint64_t a, uint64_t b) {
    /* branch-free on the common carry (the reference uses add/sbb inline asm for this,
     * goldilocks_field.rs:345-368); the double overflow is the rare, hinted branch. */
    uint64_t s, s2;
    uint64_t over = __builtin_add_overflow(a, b, &s);
    uint64_t over2 = __builtin_add_overflow(s, (0 - over) & EPS, &s2);
    if (__builtin_expect(over2, 0)) s2 += EPS;
    return s2;
}

/* goldilocks_field.rs:280-294 */
static inline uint64_t gl_sub(uint64_t a, uint64_t b) {
    uint64_t d, d2;
    uint64_t under = __builtin_sub_overflow(a, b, &d);
    uint64_t under2 = __builtin_sub_overflow(d, (0 - under) & EPS, &d2);
    if (__builtin_expect(under2, 0)) d2 -= EPS;
    return d2;
}
uint64_t orc_gl_add(uint64_t a, uint64_t b) { return gl_add(a, b); }
uint64_t orc_gl_sub(uint64_t a, uint64_t b) { return gl_sub(a, b); }

/* goldilocks_field.rs:390-403 (reduce128) */
static inline uint64_t reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
    uint64_t t0 = lo - hi_hi;
    if (__builtin_expect(lo < hi_hi, 0)) t0 -= EPS; /* branch_hint(): a borrow is exceedingly rare */
    uint64_t t1 = hi_lo * EPS;
    uint64_t t2;
    uint64_t c = __builtin_add_overflow(t0, t1, &t2); /* add_no_canonicalize_trashing_input */
    return t2 + ((0 - c) & EPS);
}

/* goldilocks_field.rs:381-385 (reduce96) */
static inline uint64_t reduce96(uint64_t lo, uint32_t hi) {
    uint64_t t1 = (uint64_t)hi * EPS;
    uint64_t t2;
    uint64_t c = __builtin_add_overflow(lo, t1, &t2);
    return t2 + ((0 - c) & EPS);
}

/* goldilocks_field.rs:303-310 */
uint64_t orc_gl_mul(uint64_t a, uint64_t b) { return reduce128((u128)a * b); }

static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return reduce128((u128)a * b); }

uint64_t orc_gl_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1, b = a;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_mul(b, b);
        e >>= 1;
    }
    return r;
}

/* Fermat inverse a^(p-2) (goldilocks_field.rs:112-151 computes the same power with an
 * addition chain). */
uint64_t orc_gl_inv(uint64_t a) { return orc_gl_pow(a, ORC_P - 2); }

/* field/src/types.rs:239-278; exp <= 32 here */
uint64_t orc_gl_inverse_2exp(unsigned k) { return ORC_P - ((ORC_P - 1) >> k); }

/* field/src/types.rs:280-284, goldilocks_field.rs:91 */
uint64_t orc_gl_primitive_root(unsigned k) {
    uint64_t b = 7277203076849721926ULL;
    for (unsigned i = k; i < 32; i++) b = gl_mul(b, b);
    return b;
}

/* field/src/types.rs:453-455, goldilocks_field.rs:84 */
uint64_t orc_gl_coset_shift(void) { return 14293326489335486720ULL; }

/* field/src/extension/quadratic.rs:186-199, goldilocks_extensions.rs:13-26 (W = 7) */
void orc_ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
    uint64_t c0 = gl_add(gl_mul(a[0], b[0]), gl_mul(7, gl_mul(a[1], b[1])));
    uint64_t c1 = gl_add(gl_mul(a[0], b[1]), gl_mul(a[1], b[0]));
    out[0] = c0;
    out[1] = c1;
}

/* ------------------------------------------------------------------------------------------
 * Bit reversal: result[i] = arr[bitrev(i)]   (util/src/lib.rs:49-97, 181-230)
 * elem_words = u64 words per element (1 for base field, 2 for F_p^2, ...).
 * ---------------------------------------------------------------------------------------- */
static inline size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

static unsigned log2_strict(size_t n) {
    unsigned k = 0;
    while (((size_t)1 << k) < n) k++;
    return k;
}

void orc_reverse_index_bits(uint64_t *arr, size_t n, size_t w) {
    unsigned bits = log2_strict(n);
    uint64_t tmp[16];
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, bits);
        if (i < j) {
            memcpy(tmp, arr + i * w, w * 8);
            memcpy(arr + i * w, arr + j * w, w * 8);
            memcpy(arr + j * w, tmp, w * 8);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * FFT (field/src/fft.rs)
 * ---------------------------------------------------------------------------------------- */

/* fft.rs:14-33.  Row j (= lg_half_m) holds omega_{2^(j+1)}^i for i < 2^j; omega_m does not
 * depend on n, so one table built for the largest size serves all.  Row j lives at
 * table[2^j .. 2^(j+1)). */
static uint64_t *g_roots = NULL;
static unsigned g_roots_lg = 0;

static const uint64_t *root_table(unsigned lg_n) {
    const uint64_t *ret;
#pragma omp critical(orc_roots)
    {
        if (g_roots_lg < lg_n || !g_roots) {
            free(g_roots);
            size_t n = (size_t)1 << lg_n;
            g_roots = (uint64_t *)malloc((n < 2 ? 2 : n) * 8);
            g_roots[0] = 1;
            for (unsigned j = 0; j < lg_n; j++) {
                uint64_t w = orc_gl_primitive_root(j + 1), cur = 1;
                size_t half = (size_t)1 << j;
                for (size_t i = 0; i < half; i++) {
                    g_roots[half + i] = cur;
                    cur = gl_mul(cur, w);
                }
            }
            g_roots_lg = lg_n;
        }
        ret = g_roots;
    }
    return ret;
}

/* fft.rs:165-202 (fft_classic) with the scalar loop of fft.rs:138-156 */
static void fft_classic(uint64_t *v, unsigned lg_n, unsigned r, const uint64_t *roots) {
    size_t n = (size_t)1 << lg_n;
    orc_reverse_index_bits(v, n, 1);
    if (r > 0) {
        size_t mask = ~(((size_t)1 << r) - 1);
        for (size_t i = 0; i < n; i++) v[i] = v[i & mask];
    }
    for (unsigned lg_half_m = r; lg_half_m < lg_n; lg_half_m++) {
        size_t half_m = (size_t)1 << lg_half_m, m = half_m * 2;
        const uint64_t *row = roots + half_m;
        for (size_t k = 0; k < n; k += m) {
            for (size_t j = 0; j < half_m; j++) {
                uint64_t t = gl_mul(row[j], v[k + half_m + j]);
                uint64_t u = v[k + j];
                v[k + j] = gl_add(u, t);
                v[k + half_m + j] = gl_sub(u, t);
            }
        }
    }
}

/* fft.rs:53-61 */
void orc_fft(uint64_t *v, unsigned lg_n, unsigned zero_factor) {
    fft_classic(v, lg_n, zero_factor, root_table(lg_n));
}

/* fft.rs:68-91 */
void orc_ifft(uint64_t *v, unsigned lg_n) {
    size_t n = (size_t)1 << lg_n;
    uint64_t n_inv = orc_gl_inverse_2exp(lg_n);
    fft_classic(v, lg_n, 0, root_table(lg_n));
    v[0] = gl_mul(v[0], n_inv);
    if (n > 1) v[n / 2] = gl_mul(v[n / 2], n_inv);
    for (size_t i = 1; i < n / 2; i++) {
        size_t j = n - i;
        uint64_t ci = gl_mul(v[j], n_inv), cj = gl_mul(v[i], n_inv);
        v[i] = ci;
        v[j] = cj;
    }
}

/* field/src/polynomial/mod.rs:280-293 */
void orc_coset_fft(uint64_t *v, unsigned lg_n, uint64_t shift, unsigned zero_factor) {
    size_t n = (size_t)1 << lg_n;
    uint64_t pw = 1;
    for (size_t i = 0; i < n; i++) {
        v[i] = gl_mul(pw, v[i]);
        pw = gl_mul(pw, shift);
    }
    orc_fft(v, lg_n, zero_factor);
}

/* Direct O(n^2) evaluation on shift*H, the statement fft.rs:230-249 and
 * polynomial/mod.rs:477-495 test against: out[k] = sum_i in[i] (shift w^k)^i. */
void orc_fft_naive(const uint64_t *in, uint64_t *out, unsigned lg_n, uint64_t shift) {
    size_t n = (size_t)1 << lg_n;
    uint64_t w = orc_gl_primitive_root(lg_n);
    for (size_t k = 0; k < n; k++) {
        uint64_t x = gl_mul(shift, orc_gl_pow(w, k));
        uint64_t acc = 0;
        for (size_t i = n; i-- > 0;) acc = gl_add(gl_mul(acc, x), in[i]);
        out[k] = acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * Poseidon, width 12 (core/src/poseidon.rs, core/src/poseidon_goldilocks.rs)
 * ---------------------------------------------------------------------------------------- */

/* poseidon.rs:546-552 */
static inline uint64_t sbox(uint64_t x) {
    uint64_t x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

/* poseidon.rs:504-513 */
static inline void constant_layer(uint64_t s[12], unsigned round) {
#pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
        /* add_canonical_u64 (goldilocks_field.rs:206-211) */
        uint64_t c = POSEIDON_ALL_ROUND_CONSTANTS[i + 12 * round];
        uint64_t t;
        uint64_t ov = __builtin_add_overflow(s[i], c, &t);
        s[i] = t + ((0 - ov) & EPS);
    }
}

/* poseidon.rs:178-198, 245-264 (mds_row_shf + mds_layer).  The coefficients are < 2^6, so the
 * 32-bit halves of the state can be accumulated in u64 without overflow and recombined, as
 * poseidon_goldilocks.rs:217-248 does. */
static inline __attribute__((always_inline)) void mds_layer(uint64_t s[12]) {
    uint64_t lo[12], hi[12], out[12];
#pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
        lo[i] = s[i] & EPS;
        hi[i] = s[i] >> 32;
    }
#pragma GCC unroll 12
    for (int r = 0; r < 12; r++) {
        uint64_t al = 0, ah = 0;
#pragma GCC unroll 12
        for (int i = 0; i < 12; i++) {
            const int k = (i + r) % 12;
            al += lo[k] * POSEIDON_MDS_CIRC[i];
            ah += hi[k] * POSEIDON_MDS_CIRC[i];
        }
        al += lo[r] * POSEIDON_MDS_DIAG[r];
        ah += hi[r] * POSEIDON_MDS_DIAG[r];
        u128 sum = (u128)al + ((u128)ah << 32);
        out[r] = reduce96((uint64_t)sum, (uint32_t)(sum >> 64));
    }
    memcpy(s, out, sizeof out);
}

/* poseidon.rs:574-581 */
static void full_rounds(uint64_t s[12], unsigned *round) {
    for (int k = 0; k < 4; k++) {
        constant_layer(s, *round);
#pragma GCC unroll 12
        for (int i = 0; i < 12; i++) s[i] = sbox(s[i]);
        mds_layer(s);
        (*round)++;
    }
}

/* poseidon.rs:584-596 with :302-313 (first constant layer), :315-342 (initial matrix),
 * :378-408 (fast partial MDS) */
static void partial_rounds_fast(uint64_t s[12], unsigned *round) {
    for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]);
    uint64_t t[12];
    t[0] = s[0];
    for (int c = 1; c < 12; c++) t[c] = 0;
#pragma GCC unroll 12
    for (int r = 1; r < 12; r++)
#pragma GCC unroll 12
        for (int c = 1; c < 12; c++)
            t[c] = gl_add(t[c], gl_mul(s[r], POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (c - 1)]));
    memcpy(s, t, sizeof t);

    for (int i = 0; i < 22; i++) {
        s[0] = sbox(s[0]);
        s[0] = gl_add(s[0], POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS[i]);
        /* d = [M_00 | w_hat] . state, accumulated in 160 bits then reduced (reduce_u160,
         * poseidon.rs:46-52); any exact mod-p evaluation gives the same canonical value. */
        u128 acc_lo = 0;
        uint32_t acc_hi = 0;
#pragma GCC unroll 12
        for (int j = 1; j < 12; j++) {
            u128 prod = (u128)s[j] * POSEIDON_FAST_PARTIAL_ROUND_W_HATS[i * 11 + j - 1];
            u128 n = acc_lo + prod;
            acc_hi += n < prod;
            acc_lo = n;
        }
        {
            u128 prod = (u128)s[0] * (POSEIDON_MDS_CIRC[0] + POSEIDON_MDS_DIAG[0]);
            u128 n = acc_lo + prod;
            acc_hi += n < prod;
            acc_lo = n;
        }
        uint64_t red_hi = reduce96((uint64_t)(acc_lo >> 64), acc_hi);
        uint64_t d = reduce128(((u128)red_hi << 64) + (uint64_t)acc_lo);
#pragma GCC unroll 12
        for (int j = 1; j < 12; j++)
            s[j] = gl_add(s[j], gl_mul(s[0], POSEIDON_FAST_PARTIAL_ROUND_VS[i * 11 + j - 1]));
        s[0] = d;
    }
    *round += 22;
}

/* poseidon.rs:599-609 */
void orc_poseidon(uint64_t s[12]) {
    unsigned round = 0;
    full_rounds(s, &round);
    partial_rounds_fast(s, &round);
    full_rounds(s, &round);
}

/* poseidon.rs:613-633 (poseidon_naive, the KAT-defining form) */
void orc_poseidon_naive(uint64_t s[12]) {
    unsigned round = 0;
    full_rounds(s, &round);
    for (int k = 0; k < 22; k++) {
        constant_layer(s, round);
        s[0] = sbox(s[0]);
        mds_layer(s);
        round++;
    }
    full_rounds(s, &round);
}

/* core/src/hashing.rs:68-95 (hash_n_to_m_no_pad, 4 outputs) */
void orc_hash_no_pad(const uint64_t *in, size_t len, uint64_t out[4]) {
    uint64_t s[12] = {0};
    for (size_t off = 0; off < len; off += 8) {
        size_t c = len - off < 8 ? len - off : 8;
        memcpy(s, in + off, c * 8);
        orc_poseidon(s);
    }
    memcpy(out, s, 32);
}

/* core/src/hashing.rs:150-168 (fork-specific domain-separated leaf hash) */
void orc_hash_leaf(const uint64_t *in, size_t len, uint64_t out[4]) {
    uint64_t s[12] = {0};
    s[8] = (uint64_t)len + 1;
    for (size_t off = 0; off < len; off += 8) {
        size_t c = len - off < 8 ? len - off : 8;
        memcpy(s, in + off, c * 8);
        orc_poseidon(s);
    }
    memcpy(out, s, 32);
}

/* core/src/hashing.rs:47-64 (compress) */
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint64_t s[12] = {0};
    memcpy(s, l, 32);
    memcpy(s + 4, r, 32);
    orc_poseidon(s);
    memcpy(out, s, 32);
}

static void canon4(uint64_t d[4]) {
    for (int i = 0; i < 4; i++) d[i] = orc_gl_canon(d[i]);
}

/* ------------------------------------------------------------------------------------------
 * Merkle tree (plonky2/src/hash/merkle_tree.rs)
 * ---------------------------------------------------------------------------------------- */

/* merkle_tree.rs:56-83 (fill_subtree).  digests_buf holds 2*(n_leaves-1) digests. */
static void fill_subtree(uint64_t *digests_buf, size_t n_digests, const uint64_t *leaves,
                         size_t n_leaves, size_t leaf_len, uint64_t out[4]) {
    if (n_digests == 0) {
        orc_hash_leaf(leaves, leaf_len, out);
        canon4(out);
        return;
    }
    size_t half = n_digests / 2;
    uint64_t *left_mem = digests_buf + (half - 1) * 4;
    uint64_t *right_mem = digests_buf + half * 4;
    uint64_t l[4], r[4];
    size_t hl = n_leaves / 2;
    if (n_leaves >= 2048) {
#pragma omp task shared(l)
        fill_subtree(digests_buf, half - 1, leaves, hl, leaf_len, l);
#pragma omp task shared(r)
        fill_subtree(right_mem + 4, half - 1, leaves + hl * leaf_len, hl, leaf_len, r);
#pragma omp taskwait
    } else {
        fill_subtree(digests_buf, half - 1, leaves, hl, leaf_len, l);
        fill_subtree(right_mem + 4, half - 1, leaves + hl * leaf_len, hl, leaf_len, r);
    }
    memcpy(left_mem, l, 32);
    memcpy(right_mem, r, 32);
    orc_two_to_one(l, r, out);
    canon4(out);
}

/* merkle_tree.rs:163-194 (new) + :85-119 (fill_digests_buf) */
int orc_merkle_tree_new(const uint64_t *leaves, size_t n_leaves, size_t leaf_len,
                        unsigned cap_height, uint64_t *digests, uint64_t *cap) {
    if (n_leaves == 0 || (n_leaves & (n_leaves - 1))) return 1;
    unsigned lg = log2_strict(n_leaves);
    if (cap_height > lg) return 1;
    size_t n_cap = (size_t)1 << cap_height;
    size_t n_digests = 2 * (n_leaves - n_cap);
    if (n_digests == 0) {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n_leaves; i++) {
            orc_hash_leaf(leaves + i * leaf_len, leaf_len, cap + i * 4);
            canon4(cap + i * 4);
        }
        return 0;
    }
    size_t sub_d = n_digests >> cap_height, sub_l = n_leaves >> cap_height;
#pragma omp parallel
#pragma omp single
    for (size_t t = 0; t < n_cap; t++) {
#pragma omp task firstprivate(t)
        fill_subtree(digests + t * sub_d * 4, sub_d, leaves + t * sub_l * leaf_len, sub_l, leaf_len,
                     cap + t * 4);
    }
    return 0;
}

/* merkle_tree.rs:121-160 (merkle_tree_prove) */
void orc_merkle_prove(size_t leaf_index, size_t n_leaves, unsigned cap_height,
                      const uint64_t *digests, uint64_t *siblings) {
    unsigned num_layers = log2_strict(n_leaves) - cap_height;
    size_t digest_len = 2 * (n_leaves - ((size_t)1 << cap_height));
    size_t tree_index = leaf_index >> num_layers;
    size_t tree_len = digest_len >> cap_height;
    const uint64_t *tree = digests + tree_len * tree_index * 4;
    size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
    for (unsigned i = 0; i < num_layers; i++) {
        size_t parity = pair_index & 1;
        pair_index >>= 1;
        size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
        size_t sibling_index = 2 * siblings_index + (1 - parity);
        memcpy(siblings + i * 4, tree + sibling_index * 4, 32);
    }
}

/* core/src/merkle_proofs.rs:59-97 with a single leaf (verify_merkle_proof_to_cap) */
int orc_merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index,
                      const uint64_t *cap, unsigned cap_height, const uint64_t *siblings,
                      unsigned n_siblings) {
    (void)cap_height;
    uint64_t cur[4];
    orc_hash_leaf(leaf, leaf_len, cur);
    for (unsigned i = 0; i < n_siblings; i++) {
        uint64_t nxt[4];
        if (leaf_index & 1)
            orc_two_to_one(siblings + i * 4, cur, nxt);
        else
            orc_two_to_one(cur, siblings + i * 4, nxt);
        leaf_index >>= 1;
        memcpy(cur, nxt, 32);
    }
    for (int k = 0; k < 4; k++)
        if (orc_gl_canon(cur[k]) != orc_gl_canon(cap[leaf_index * 4 + k])) return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * PolynomialBatch (plonky2/src/fri/oracle.rs:168-283)
 * ---------------------------------------------------------------------------------------- */

int orc_batch_from_coeffs(const uint64_t *coeffs, size_t n_cols, unsigned lg_n, unsigned rate_bits,
                          unsigned cap_height, const uint64_t *salt, uint64_t *leaves_out,
                          uint64_t *digests_out, uint64_t *cap_out, double scope_ms[4]) {
    size_t n = (size_t)1 << lg_n, N = n << rate_bits;
    unsigned lg_N = lg_n + rate_bits;
    if (cap_height > lg_N) return 1;
    size_t leaf_len = n_cols + (salt ? 4 : 0);
    double t0 = now_ms();
    /* "FFT + blinding": oracle.rs:202-206, 267-283 -- par over columns */
    uint64_t *lde = (uint64_t *)malloc(leaf_len * N * 8);
    if (!lde) return 2;
    root_table(lg_N);
    uint64_t g = orc_gl_coset_shift();
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < n_cols; c++) {
        uint64_t *col = lde + c * N;
        memcpy(col, coeffs + c * n, n * 8);          /* lde(): zero-pad, polynomial/mod.rs:199-201 */
        memset(col + n, 0, (N - n) * 8);
        orc_coset_fft(col, lg_N, g, rate_bits);      /* coset_fft_with_options(g, Some(r), table) */
    }
    if (salt) memcpy(lde + n_cols * N, salt, 4 * N * 8); /* oracle.rs:259-263, salt injected */
    double t1 = now_ms();
    /* "transpose LDEs" + reverse_index_bits_in_place(leaves): oracle.rs:208-209 */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) {
        size_t src = bitrev(i, lg_N);
        uint64_t *leaf = leaves_out + i * leaf_len;
        for (size_t c = 0; c < leaf_len; c++) leaf[c] = orc_gl_canon(lde[c * N + src]);
    }
    free(lde);
    double t2 = now_ms();
    /* "build Merkle tree": oracle.rs:210-214 */
    int rc = orc_merkle_tree_new(leaves_out, N, leaf_len, cap_height, digests_out, cap_out);
    double t3 = now_ms();
    if (scope_ms) {
        scope_ms[1] = t1 - t0;
        scope_ms[2] = t2 - t1;
        scope_ms[3] = t3 - t2;
    }
    return rc;
}

int orc_batch_from_values(const uint64_t *values, size_t n_cols, unsigned lg_n, unsigned rate_bits,
                          unsigned cap_height, const uint64_t *salt, uint64_t *coeffs_out,
                          uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out,
                          double scope_ms[4]) {
    size_t n = (size_t)1 << lg_n;
    double t0 = now_ms();
    root_table(lg_n + rate_bits);
    /* "IFFT": oracle.rs:176-180 -- par over columns */
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < n_cols; c++) {
        uint64_t *col = coeffs_out + c * n;
        if (col != values + c * n) memcpy(col, values + c * n, n * 8);
        orc_ifft(col, lg_n);
        for (size_t i = 0; i < n; i++) col[i] = orc_gl_canon(col[i]);
    }
    double t1 = now_ms();
    if (scope_ms) scope_ms[0] = t1 - t0;
    return orc_batch_from_coeffs(coeffs_out, n_cols, lg_n, rate_bits, cap_height, salt, leaves_out,
                                 digests_out, cap_out, scope_ms);
}

/* ------------------------------------------------------------------------------------------
 * Challenger (core/src/challenger.rs)
 * ---------------------------------------------------------------------------------------- */
void orc_challenger_init(orc_challenger *c) { memset(c, 0, sizeof *c); }

/* challenger.rs:125-140 */
static void duplexing(orc_challenger *c) {
    for (uint32_t i = 0; i < c->n_in; i++) c->sponge_state[i] = c->input_buffer[i];
    c->n_in = 0;
    orc_poseidon(c->sponge_state);
    for (int i = 0; i < 8; i++) c->output_buffer[i] = c->sponge_state[i];
    c->n_out = 8;
}

/* challenger.rs:35-46 */
void orc_challenger_observe(orc_challenger *c, const uint64_t *elems, size_t n) {
    for (size_t i = 0; i < n; i++) {
        c->n_out = 0;
        c->input_buffer[c->n_in++] = elems[i];
        if (c->n_in == 8) duplexing(c);
    }
}

/* challenger.rs:78-89: pop from the END of the output buffer */
uint64_t orc_challenger_get(orc_challenger *c) {
    if (c->n_in != 0 || c->n_out == 0) duplexing(c);
    return orc_gl_canon(c->output_buffer[--c->n_out]);
}

/* ------------------------------------------------------------------------------------------
 * FRI commit phase
 * ---------------------------------------------------------------------------------------- */

/* core/src/fri.rs:50-61 */
unsigned orc_fri_reduction_arity_bits(unsigned degree_bits, unsigned rate_bits, unsigned cap_height,
                                      unsigned arity_bits, unsigned final_poly_bits, unsigned *out) {
    unsigned k = 0;
    while (degree_bits > final_poly_bits && degree_bits + rate_bits - arity_bits >= cap_height) {
        out[k++] = arity_bits;
        degree_bits -= arity_bits;
    }
    return k;
}

/* coset_fft over F_p^2 = two base-field lanes (extension/mod.rs:73-76: the ext 2-adic
 * generator is the embedded base one, so twiddles are base-field). */
static void ext_coset_fft(uint64_t *v /* n x 2 */, unsigned lg_n, uint64_t shift) {
    size_t n = (size_t)1 << lg_n;
    uint64_t *lane = (uint64_t *)malloc(n * 8);
    for (int k = 0; k < 2; k++) {
        for (size_t i = 0; i < n; i++) lane[i] = v[2 * i + k];
        orc_coset_fft(lane, lg_n, shift, 0);
        for (size_t i = 0; i < n; i++) v[2 * i + k] = orc_gl_canon(lane[i]);
    }
    free(lane);
}

/* plonky2/src/fri/prover.rs:85-143 */
int orc_fri_committed_trees(uint64_t *coeffs, uint64_t *values, unsigned lg_n, unsigned rate_bits,
                            unsigned cap_height, const unsigned *arity_bits, unsigned n_rounds,
                            orc_challenger *challenger, uint64_t *caps_out, uint64_t **leaves_out,
                            uint64_t **digests_out, uint64_t *betas_out, uint64_t *final_poly_out) {
    size_t n = (size_t)1 << lg_n;
    size_t cap_words = ((size_t)1 << cap_height) * 4;
    uint64_t shift = orc_gl_coset_shift();
    for (unsigned step = 0; step < n_rounds; step++) {
        unsigned ab = arity_bits[step];
        size_t arity = (size_t)1 << ab;
        /* reverse_index_bits_in_place(values); leaves = chunks of `arity`, flattened */
        orc_reverse_index_bits(values, n, 2);
        for (size_t i = 0; i < 2 * n; i++) values[i] = orc_gl_canon(values[i]);
        size_t n_leaves = n >> ab, leaf_len = 2 * arity;
        if (cap_height > log2_strict(n_leaves)) return 1;
        size_t n_dig = 2 * (n_leaves - ((size_t)1 << cap_height));
        uint64_t *dig = (uint64_t *)malloc((n_dig ? n_dig : 1) * 32);
        uint64_t *cap = caps_out + step * cap_words;
        int rc = orc_merkle_tree_new(values, n_leaves, leaf_len, cap_height, dig, cap);
        if (rc) {
            free(dig);
            return rc;
        }
        if (leaves_out && leaves_out[step]) memcpy(leaves_out[step], values, n * 16);
        if (digests_out && digests_out[step]) memcpy(digests_out[step], dig, n_dig * 32);
        free(dig);
        orc_challenger_observe(challenger, cap, cap_words);
        uint64_t beta[2];
        beta[0] = orc_challenger_get(challenger);
        beta[1] = orc_challenger_get(challenger);
        if (betas_out) {
            betas_out[2 * step] = beta[0];
            betas_out[2 * step + 1] = beta[1];
        }
        /* reduce_with_powers over chunks (core/src/plonk_common.rs:87-98) */
        for (size_t i = 0; i < n_leaves; i++) {
            uint64_t sum[2] = {0, 0};
            for (size_t j = arity; j-- > 0;) {
                uint64_t t[2];
                orc_ext_mul(sum, beta, t);
                sum[0] = gl_add(t[0], coeffs[2 * (i * arity + j)]);
                sum[1] = gl_add(t[1], coeffs[2 * (i * arity + j) + 1]);
            }
            coeffs[2 * i] = orc_gl_canon(sum[0]);
            coeffs[2 * i + 1] = orc_gl_canon(sum[1]);
        }
        n = n_leaves;
        lg_n -= ab;
        if (step + 1 == n_rounds) continue;
        shift = orc_gl_pow(shift, arity);
        memcpy(values, coeffs, n * 16);
        ext_coset_fft(values, lg_n, shift);
    }
    size_t fin = n >> rate_bits;
    memcpy(final_poly_out, coeffs, fin * 16);
    return 0;
}

/* plonky2/src/fri/prover.rs:159-208, smallest-witness rule */
uint64_t orc_fri_proof_of_work(orc_challenger *ch, unsigned pow_bits) {
    uint64_t base[12];
    memcpy(base, ch->sponge_state, sizeof base);
    for (uint32_t i = 0; i < ch->n_in; i++) base[i] = ch->input_buffer[i];
    uint32_t pos = ch->n_in;
    uint64_t w;
    for (w = 0;; w++) {
        uint64_t s[12];
        memcpy(s, base, sizeof s);
        s[pos] = w;
        orc_poseidon(s);
        uint64_t resp = orc_gl_canon(s[7]);
        unsigned lz = resp ? (unsigned)__builtin_clzll(resp) : 64;
        if (lz >= pow_bits) break;
    }
    orc_challenger_observe(ch, &w, 1);
    (void)orc_challenger_get(ch);
    return w;
}

/* ------------------------------------------------------------------------------------------
 * Opening side of prove_openings
 * ---------------------------------------------------------------------------------------- */

/* field/src/polynomial/mod.rs:155-160 (eval: Horner from the leading coefficient), with the base
 * coefficients embedded in F_p^2 (to_extension, proof.rs:299-304). */
void orc_eval_poly_ext(const uint64_t *coeffs, size_t n, const uint64_t point[2], uint64_t out[2]) {
    uint64_t acc[2] = {0, 0};
    for (size_t i = n; i-- > 0;) {
        uint64_t t[2];
        orc_ext_mul(acc, point, t);
        acc[0] = gl_add(t[0], coeffs[i]);
        acc[1] = t[1];
    }
    out[0] = orc_gl_canon(acc[0]);
    out[1] = orc_gl_canon(acc[1]);
}

/* plonky2/src/fri/oracle.rs:129-165.  composition = sum_t weight_t * poly_t (the alpha-powers and
 * FriCoefficient values are folded into the weights by the caller, reducing.rs:63-72);
 * quotient = divide_by_linear(point) (division.rs:77-90) padded with one zero;
 * final = final * shift + quotient (reducing.rs:94-97). */
void orc_reduce_openings(size_t n_batches, const size_t *n_terms, const uint64_t *const *term_polys,
                         const uint64_t *weights, const uint64_t *points, const uint64_t *shifts,
                         unsigned degree_log, uint64_t *final_out) {
    size_t n = (size_t)1 << degree_log;
    uint64_t *comp = (uint64_t *)malloc(n * 16);
    uint64_t *bs = (uint64_t *)malloc(n * 16);
    size_t base = 0;
    for (size_t i = 0; i < 2 * n; i++) final_out[i] = 0;
    for (size_t b = 0; b < n_batches; b++) {
        for (size_t j = 0; j < n; j++) {
            uint64_t a0 = 0, a1 = 0;
            for (size_t t = 0; t < n_terms[b]; t++) {
                uint64_t c = term_polys[base + t][j];
                a0 = gl_add(a0, gl_mul(weights[2 * (base + t)], c));
                a1 = gl_add(a1, gl_mul(weights[2 * (base + t) + 1], c));
            }
            comp[2 * j] = a0;
            comp[2 * j + 1] = a1;
        }
        /* divide_by_linear: bs = scan from the top of acc = acc*z + c; drop the last (= p(z)); reverse */
        const uint64_t *z = points + 2 * b;
        uint64_t acc[2] = {0, 0};
        for (size_t i = n; i-- > 0;) {
            uint64_t t[2];
            orc_ext_mul(acc, z, t);
            acc[0] = gl_add(t[0], comp[2 * i]);
            acc[1] = gl_add(t[1], comp[2 * i + 1]);
            bs[2 * i] = acc[0];
            bs[2 * i + 1] = acc[1];
        }
        /* quotient[k] = bs[k+1] for k < n-1, quotient[n-1] = 0 (the pushed zero) */
        for (size_t k = 0; k < n; k++) {
            uint64_t q[2] = {0, 0}, f[2];
            if (k + 1 < n) {
                q[0] = bs[2 * (k + 1)];
                q[1] = bs[2 * (k + 1) + 1];
            }
            orc_ext_mul(final_out + 2 * k, shifts + 2 * b, f);
            final_out[2 * k] = orc_gl_canon(gl_add(f[0], q[0]));
            final_out[2 * k + 1] = orc_gl_canon(gl_add(f[1], q[1]));
        }
        base += n_terms[b];
    }
    free(comp);
    free(bs);
}

/* ------------------------------------------------------------------------------------------
 * Plonk permutation argument and quotient polynomials
 * plonky2/src/plonk/prover.rs:402-480,640-866, plonky2/src/plonk/vanishing_poly.rs:166-330,
 * plonky2/src/util/partial_products.rs, field/src/zero_poly_coset.rs
 * ---------------------------------------------------------------------------------------- */

/* compute_filter, plonky2/src/gates/gate.rs:326-333 */
static uint64_t gate_filter(const orc_gate *g, uint64_t s, int many_selectors) {
    uint64_t f = 1;
    for (unsigned i = g->group_start; i < g->group_end; i++)
        if (i != g->index) f = gl_mul(f, gl_sub(i, s));
    if (many_selectors) f = gl_mul(f, gl_sub(0xFFFFFFFFULL /* UNUSED_SELECTOR */, s));
    return f;
}

static void poseidon_gate_eval(const uint64_t *wires, uint64_t *c);

/* F_p^2 helpers on pairs for the extension-field gates */
static inline void ext_sub2(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
    out[0] = gl_sub(a[0], b[0]);
    out[1] = gl_sub(a[1], b[1]);
}
static inline void ext_add2(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
    out[0] = gl_add(a[0], b[0]);
    out[1] = gl_add(a[1], b[1]);
}

/* two_adic_subgroup (field/src/types.rs:292-295) and barycentric_weights (field/src/interpolation.rs:53-65)
 * of CosetInterpolationGate, per subgroup_bits */
static void coset_interpolation_tables(unsigned bits, uint64_t *domain, uint64_t *weights) {
    const unsigned n = 1u << bits;
    const uint64_t g = orc_gl_primitive_root(bits);
    uint64_t v = 1;
    for (unsigned i = 0; i < n; i++, v = gl_mul(v, g)) domain[i] = orc_gl_canon(v);
    for (unsigned i = 0; i < n; i++) {
        uint64_t d = 1;
        for (unsigned j = 0; j < n; j++)
            if (j != i) d = gl_mul(d, gl_sub(domain[i], domain[j]));
        weights[i] = orc_gl_inv(d);
    }
}

/* partial_interpolate, plonky2/src/gates/coset_interpolation.rs:572-599 */
static void partial_interpolate(const uint64_t *domain, const uint64_t *weights, const uint64_t *values /* [..][2] */,
                                unsigned lo, unsigned hi, const uint64_t x[2], uint64_t ev[2], uint64_t prod[2]) {
    for (unsigned j = lo; j < hi; j++) {
        uint64_t val[2] = {gl_mul(values[2 * j], weights[j]), gl_mul(values[2 * j + 1], weights[j])};
        uint64_t term[2] = {gl_sub(x[0], domain[j]), x[1]};
        uint64_t a[2], b[2];
        orc_ext_mul(ev, term, a);
        orc_ext_mul(val, prod, b);
        ext_add2(a, b, ev);
        orc_ext_mul(prod, term, a);
        prod[0] = a[0];
        prod[1] = a[1];
    }
}

/* eval_unfiltered of the supported gates; `consts` already has the selector prefix removed
 * (gate.rs:179).  Adds filter * constraint_k into acc[k] (vanishing_poly.rs:700-726). */
static void gate_eval_add(const orc_gate *g, const uint64_t *consts, const uint64_t *wires,
                          const uint64_t pih[4], uint64_t filter, uint64_t *acc) {
    switch (g->kind) {
    case ORC_GATE_NOOP: /* gates/noop.rs: no constraints */
    case ORC_GATE_LOOKUP: /* gates/lookup.rs:190-195, gates/lookup_table.rs:177-181: no gate constraints */
    case ORC_GATE_LOOKUP_TABLE:
        break;
    case ORC_GATE_CONSTANT: /* gates/constant.rs:121-129 */
        for (unsigned i = 0; i < g->param; i++)
            acc[i] = gl_add(acc[i], gl_mul(filter, gl_sub(consts[i], wires[i])));
        break;
    case ORC_GATE_PUBLIC_INPUT: /* gates/public_input.rs:103-113 */
        for (unsigned i = 0; i < 4; i++)
            acc[i] = gl_add(acc[i], gl_mul(filter, gl_sub(wires[i], pih[i])));
        break;
    case ORC_GATE_ARITHMETIC: /* gates/arithmetic_base.rs:168-185 */
        for (unsigned i = 0; i < g->param; i++) {
            uint64_t m0 = wires[4 * i], m1 = wires[4 * i + 1], ad = wires[4 * i + 2], out = wires[4 * i + 3];
            uint64_t computed = gl_add(gl_mul(gl_mul(m0, m1), consts[0]), gl_mul(ad, consts[1]));
            acc[i] = gl_add(acc[i], gl_mul(filter, gl_sub(out, computed)));
        }
        break;
    case ORC_GATE_ARITHMETIC_EXT: /* gates/arithmetic_extension.rs:92-110 */
        for (unsigned i = 0; i < g->param; i++) {
            const uint64_t *w = wires + 8 * i;
            uint64_t pr[2];
            orc_ext_mul(w, w + 2, pr);
            for (int k = 0; k < 2; k++) {
                uint64_t computed = gl_add(gl_mul(pr[k], consts[0]), gl_mul(w[4 + k], consts[1]));
                acc[2 * i + k] = gl_add(acc[2 * i + k], gl_mul(filter, gl_sub(w[6 + k], computed)));
            }
        }
        break;
    case ORC_GATE_MUL_EXT: /* gates/multiplication_extension.rs:86-101 */
        for (unsigned i = 0; i < g->param; i++) {
            const uint64_t *w = wires + 6 * i;
            uint64_t pr[2];
            orc_ext_mul(w, w + 2, pr);
            for (int k = 0; k < 2; k++)
                acc[2 * i + k] = gl_add(acc[2 * i + k], gl_mul(filter, gl_sub(w[4 + k], gl_mul(pr[k], consts[0]))));
        }
        break;
    case ORC_GATE_BASE_SUM_2: { /* gates/base_sum.rs:153-170, B = 2 */
        uint64_t sum = 0;
        for (unsigned i = g->param; i >= 1; i--) sum = gl_add(gl_mul(sum, 2), wires[i]); /* reduce_with_powers */
        acc[0] = gl_add(acc[0], gl_mul(filter, gl_sub(sum, wires[0])));
        for (unsigned i = 1; i <= g->param; i++)
            acc[i] = gl_add(acc[i], gl_mul(filter, gl_mul(wires[i], gl_sub(wires[i], 1))));
        break;
    }
    case ORC_GATE_POSEIDON: {
        uint64_t c[123];
        poseidon_gate_eval(wires, c);
        for (unsigned i = 0; i < 123; i++) acc[i] = gl_add(acc[i], gl_mul(filter, c[i]));
        break;
    }
    case ORC_GATE_RANDOM_ACCESS: { /* gates/random_access.rs:144-189; wire layout :78-127 */
        const unsigned bits = g->param & 0xFF, copies = (g->param >> 8) & 0xFF, extra = g->param >> 16;
        const unsigned vec = 1u << bits, routed = (2 + vec) * copies + extra;
        unsigned k = 0;
        for (unsigned copy = 0; copy < copies; copy++) {
            const uint64_t *w = wires + (2 + vec) * copy; /* [index, claimed, items...] */
            const uint64_t *b = wires + routed + copy * bits;
            uint64_t items[64];
            for (unsigned i = 0; i < vec; i++) items[i] = w[2 + i];
            for (unsigned i = 0; i < bits; i++, k++)
                acc[k] = gl_add(acc[k], gl_mul(filter, gl_mul(b[i], gl_sub(b[i], 1))));
            uint64_t rec = 0;
            for (unsigned i = bits; i-- > 0;) rec = gl_add(gl_add(rec, rec), b[i]);
            acc[k] = gl_add(acc[k], gl_mul(filter, gl_sub(rec, w[0])));
            k++;
            for (unsigned i = 0, len = vec; i < bits; i++, len >>= 1)
                for (unsigned j = 0; j < len / 2; j++)
                    items[j] = gl_add(items[2 * j], gl_mul(b[i], gl_sub(items[2 * j + 1], items[2 * j])));
            acc[k] = gl_add(acc[k], gl_mul(filter, gl_sub(items[0], w[1])));
            k++;
        }
        for (unsigned i = 0; i < extra; i++, k++)
            acc[k] = gl_add(acc[k], gl_mul(filter, gl_sub(consts[i], wires[(2 + vec) * copies + i])));
        break;
    }
    case ORC_GATE_REDUCING:       /* gates/reducing.rs:109-133: output 0..2, alpha 2..4, old_acc 4..6, coeffs 6.. */
    case ORC_GATE_REDUCING_EXT: { /* gates/reducing_extension.rs:113-132: coefficients are F_p^2 pairs */
        const int ext = g->kind == ORC_GATE_REDUCING_EXT;
        const unsigned n = g->param, start_accs = 6 + (ext ? 2 * n : n);
        const uint64_t *alpha = wires + 2, *cur = wires + 4;
        for (unsigned i = 0; i < n; i++) {
            const uint64_t *nxt = (i == n - 1) ? wires : wires + start_accs + 2 * i;
            uint64_t t[2];
            orc_ext_mul(cur, alpha, t);
            if (ext) {
                t[0] = gl_add(t[0], wires[6 + 2 * i]);
                t[1] = gl_add(t[1], wires[7 + 2 * i]);
            } else {
                t[0] = gl_add(t[0], wires[6 + i]);
            }
            acc[2 * i] = gl_add(acc[2 * i], gl_mul(filter, gl_sub(t[0], nxt[0])));
            acc[2 * i + 1] = gl_add(acc[2 * i + 1], gl_mul(filter, gl_sub(t[1], nxt[1])));
            cur = nxt;
        }
        break;
    }
    case ORC_GATE_POSEIDON_MDS: { /* gates/poseidon_mds.rs:150-169: the MDS layer on 12 F_p^2 lanes */
        for (unsigned r = 0; r < 12; r++)
            for (unsigned k = 0; k < 2; k++) {
                uint64_t sum = gl_mul(wires[2 * r + k], POSEIDON_MDS_DIAG[r]);
                for (unsigned i = 0; i < 12; i++)
                    sum = gl_add(sum, gl_mul(wires[2 * ((i + r) % 12) + k], POSEIDON_MDS_CIRC[i]));
                acc[2 * r + k] = gl_add(acc[2 * r + k], gl_mul(filter, gl_sub(wires[24 + 2 * r + k], sum)));
            }
        break;
    }
    case ORC_GATE_EXPONENTIATION: { /* gates/exponentiation.rs:210-245: base 0, bits 1..1+n (LE), output 1+n */
        const unsigned n = g->param;
        const uint64_t base = wires[0], *bit = wires + 1, *inter = wires + 2 + n;
        for (unsigned i = 0; i < n; i++) {
            const uint64_t prev = i == 0 ? 1 : gl_mul(inter[i - 1], inter[i - 1]);
            const uint64_t cur_bit = bit[n - i - 1];
            const uint64_t computed = gl_mul(prev, gl_add(gl_mul(cur_bit, base), gl_sub(1, cur_bit)));
            acc[i] = gl_add(acc[i], gl_mul(filter, gl_sub(computed, inter[i])));
        }
        acc[n] = gl_add(acc[n], gl_mul(filter, gl_sub(wires[1 + n], inter[n - 1])));
        break;
    }
    case ORC_GATE_COSET_INTERPOLATION: { /* gates/coset_interpolation.rs:260-307; wires :77-155 */
        const unsigned bits = g->param & 0xFF, degree = g->param >> 8, points = 1u << bits;
        const unsigned inter = (points - 2) / (degree - 1), start = 1 + 2 * points + 4;
        uint64_t domain[64], weights[64];
        coset_interpolation_tables(bits, domain, weights);
        const uint64_t shift = wires[0], shift_inv = wires[start + 2 * (2 * inter + 1)];
        const uint64_t *point = wires + 1 + 2 * points, *x = wires + start + 4 * inter, *values = wires + 1;
        unsigned k = 0;
        acc[k] = gl_add(acc[k], gl_mul(filter, gl_sub(gl_mul(shift, shift_inv), 1)));
        k++;
        for (unsigned j = 0; j < 2; j++, k++)
            acc[k] = gl_add(acc[k], gl_mul(filter, gl_sub(point[j], gl_mul(x[j], shift))));
        uint64_t ev[2] = {0, 0}, prod[2] = {1, 0};
        partial_interpolate(domain, weights, values, 0, degree, x, ev, prod);
        for (unsigned i = 0; i < inter; i++) {
            const uint64_t *iev = wires + start + 2 * i, *iprod = wires + start + 2 * (inter + i);
            for (unsigned j = 0; j < 2; j++, k++) acc[k] = gl_add(acc[k], gl_mul(filter, gl_sub(iev[j], ev[j])));
            for (unsigned j = 0; j < 2; j++, k++) acc[k] = gl_add(acc[k], gl_mul(filter, gl_sub(iprod[j], prod[j])));
            const unsigned lo = 1 + (degree - 1) * (i + 1);
            const unsigned hi = lo + degree - 1 < points ? lo + degree - 1 : points;
            ev[0] = iev[0], ev[1] = iev[1], prod[0] = iprod[0], prod[1] = iprod[1];
            partial_interpolate(domain, weights, values, lo, hi, x, ev, prod);
        }
        const uint64_t *out = wires + 3 + 2 * points;
        for (unsigned j = 0; j < 2; j++, k++) acc[k] = gl_add(acc[k], gl_mul(filter, gl_sub(out[j], ev[j])));
        break;
    }
    }
}

/* PoseidonGate::eval_unfiltered_base_one, plonky2/src/gates/poseidon.rs:204-283.  Wire layout
 * poseidon.rs:43-100: inputs 0..11, outputs 12..23, swap 24, deltas 25..28, first-half S-box
 * inputs 29.., partial 65.., second-half 87... */
static void poseidon_gate_eval(const uint64_t *wires, uint64_t *c /* 123 constraints */) {
    unsigned k = 0;
    const uint64_t swap = wires[24];
    c[k++] = gl_mul(swap, gl_sub(swap, 1));
    for (int i = 0; i < 4; i++)
        c[k++] = gl_sub(gl_mul(swap, gl_sub(wires[i + 4], wires[i])), wires[25 + i]);
    uint64_t st[12];
    for (int i = 0; i < 4; i++) {
        st[i] = gl_add(wires[i], wires[25 + i]);
        st[i + 4] = gl_sub(wires[i + 4], wires[25 + i]);
    }
    for (int i = 8; i < 12; i++) st[i] = wires[i];
    unsigned round = 0;
    for (int r = 0; r < 4; r++) {
        constant_layer(st, round);
        if (r != 0)
            for (int i = 0; i < 12; i++) {
                const uint64_t in = wires[29 + 12 * (r - 1) + i];
                c[k++] = gl_sub(st[i], in);
                st[i] = in;
            }
        for (int i = 0; i < 12; i++) st[i] = sbox(st[i]);
        mds_layer(st);
        round++;
    }
    /* partial_first_constant_layer + mds_partial_layer_init, core/src/poseidon.rs:302-342 */
    for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]);
    uint64_t t[12];
    t[0] = st[0];
    for (int cc = 1; cc < 12; cc++) t[cc] = 0;
    for (int r = 1; r < 12; r++)
        for (int cc = 1; cc < 12; cc++)
            t[cc] = gl_add(t[cc], gl_mul(st[r], POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (cc - 1)]));
    memcpy(st, t, sizeof t);
    for (int r = 0; r < 22; r++) {
        const uint64_t in = wires[65 + r];
        c[k++] = gl_sub(st[0], in);
        st[0] = sbox(in);
        if (r != 21) st[0] = gl_add(st[0], POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS[r]);
        /* mds_partial_layer_fast, core/src/poseidon.rs:378-408 */
        uint64_t d = gl_mul(st[0], POSEIDON_MDS_CIRC[0] + POSEIDON_MDS_DIAG[0]);
        for (int j = 1; j < 12; j++) d = gl_add(d, gl_mul(st[j], POSEIDON_FAST_PARTIAL_ROUND_W_HATS[r * 11 + j - 1]));
        for (int j = 1; j < 12; j++) st[j] = gl_add(st[j], gl_mul(st[0], POSEIDON_FAST_PARTIAL_ROUND_VS[r * 11 + j - 1]));
        st[0] = d;
    }
    round += 22;
    for (int r = 0; r < 4; r++) {
        constant_layer(st, round);
        for (int i = 0; i < 12; i++) {
            const uint64_t in = wires[87 + 12 * r + i];
            c[k++] = gl_sub(st[i], in);
            st[i] = in;
        }
        for (int i = 0; i < 12; i++) st[i] = sbox(st[i]);
        mds_layer(st);
        round++;
    }
    for (int i = 0; i < 12; i++) c[k++] = gl_sub(st[i], wires[12 + i]);
}

unsigned orc_gate_num_constraints(const orc_gate *g) {
    switch (g->kind) {
    case ORC_GATE_POSEIDON: return 123; /* poseidon.rs:416-422 */
    case ORC_GATE_ARITHMETIC_EXT: return 2 * g->param;
    case ORC_GATE_MUL_EXT: return 2 * g->param;
    case ORC_GATE_BASE_SUM_2: return 1 + g->param;
    case ORC_GATE_CONSTANT: return g->param;
    case ORC_GATE_PUBLIC_INPUT: return 4;
    case ORC_GATE_ARITHMETIC: return g->param;
    case ORC_GATE_RANDOM_ACCESS: /* random_access.rs:279-282 */
        return ((g->param >> 8) & 0xFF) * ((g->param & 0xFF) + 2) + (g->param >> 16);
    case ORC_GATE_REDUCING:
    case ORC_GATE_REDUCING_EXT: return 2 * g->param;
    case ORC_GATE_POSEIDON_MDS: return 24;
    case ORC_GATE_EXPONENTIATION: return g->param + 1;
    case ORC_GATE_COSET_INTERPOLATION: { /* coset_interpolation.rs:406-410 */
        const unsigned points = 1u << (g->param & 0xFF), degree = g->param >> 8;
        return 5 + 4 * ((points - 2) / (degree - 1));
    }
    default: return 0;
    }
}

static unsigned circuit_num_gate_constraints(const orc_circuit *c) {
    unsigned m = 0;
    for (unsigned i = 0; i < c->num_gates; i++) {
        unsigned k = orc_gate_num_constraints(&c->gates[i]);
        if (k > m) m = k;
    }
    return m;
}

/* eval_vanishing_poly_base_batch for one point (vanishing_poly.rs:166-330, without lookups).
 * x is the evaluation point itself (the reference passes the coset-shifted x), z_h_x = Z_H(x)
 * = x^n - 1.  res[num_challenges]. */
void orc_eval_vanishing_poly_base(const orc_circuit *c, uint64_t x, uint64_t z_h_x, const uint64_t *constants,
                                  const uint64_t *wires, const uint64_t *local_zs, const uint64_t *next_zs,
                                  const uint64_t *partial_products, const uint64_t *s_sigmas,
                                  const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas,
                                  const uint64_t pih[4], uint64_t *res) {
    orc_eval_vanishing_poly_base_lookup(c, x, z_h_x, constants, wires, local_zs, next_zs, partial_products, s_sigmas,
                                        NULL, NULL, betas, gammas, NULL, alphas, pih, res);
}

/* get_lut_poly(...).eval(delta), vanishing_poly.rs:29-52: the table's combos input + b * output, padded with
 * the first entry to whole LookupTableGate rows, as the coefficients of a polynomial in reverse order -- i.e.
 * a Horner accumulation over the entries in table order. */
uint64_t orc_lut_poly_eval(const orc_circuit *c, unsigned lut_index, const uint64_t deltas4[4]) {
    const orc_lookups *lk = c->lookups;
    const unsigned n = lk->lut_lens[lut_index], slots = c->num_routed_wires / 3;
    const unsigned padded = (slots - n % slots) % slots;
    const uint16_t *t = lk->luts[lut_index];
    const uint64_t b = deltas4[1], delta = deltas4[3];
    uint64_t acc = 0;
    for (unsigned k = 0; k < n + padded; k++) {
        const unsigned e = k < n ? k : 0;
        acc = gl_add(gl_mul(acc, delta), gl_add(t[2 * e], gl_mul(b, t[2 * e + 1])));
    }
    return orc_gl_canon(acc);
}

/* check_lookup_constraints_batch, vanishing_poly.rs:521-680, for one challenge: 4 + num_luts + 2 num_sldc
 * constraints into out[].  lookup_selectors = constants[num_selectors ..]; d = this challenge's four deltas. */
static unsigned lookup_constraints(const orc_circuit *c, const uint64_t *wires, const uint64_t *local_zs,
                                   const uint64_t *next_zs, const uint64_t *sel, const uint64_t d[4], uint64_t *out) {
    const orc_lookups *lk = c->lookups;
    const unsigned num_lu_slots = c->num_routed_wires / 2, num_lut_slots = c->num_routed_wires / 3;
    const unsigned lu_degree = lk->lookup_degree, num_sldc = lk->num_lookup_polys - 1;
    const unsigned lut_degree = (num_lut_slots + num_sldc - 1) / num_sldc;
    enum { TRANS_SRE = 0, TRANS_LDC = 1, INIT_SRE = 2, LAST_LDC = 3, START_END = 4 };
    const uint64_t z_re = local_zs[0], next_z_re = next_zs[0];
    const uint64_t *z_x = local_zs + 1, *z_gx = next_zs + 1;
    uint64_t looked[num_lut_slots], looking[num_lu_slots], lookup[num_lut_slots];
    for (unsigned s = 0; s < num_lut_slots; s++) {
        looked[s] = gl_add(wires[3 * s], gl_mul(d[0], wires[3 * s + 1]));
        lookup[s] = gl_add(wires[3 * s], gl_mul(d[1], wires[3 * s + 1]));
    }
    for (unsigned s = 0; s < num_lu_slots; s++) looking[s] = gl_add(wires[2 * s], gl_mul(d[0], wires[2 * s + 1]));
    unsigned k = 0;
    out[k++] = gl_mul(sel[LAST_LDC], z_x[num_sldc - 1]);
    out[k++] = gl_mul(sel[INIT_SRE], z_x[0]);
    out[k++] = gl_mul(sel[INIT_SRE], z_re);
    for (unsigned r = 0; r < lk->num_luts; r++)
        out[k++] = gl_mul(sel[START_END + r], gl_sub(z_re, orc_lut_poly_eval(c, r, d)));
    uint64_t cur = next_z_re;
    for (unsigned s = 0; s < num_lut_slots; s++) cur = gl_add(gl_mul(cur, d[3]), lookup[s]);
    out[k++] = gl_mul(sel[TRANS_SRE], gl_sub(z_re, cur));
    for (unsigned poly = 0; poly < num_sldc; poly++) {
        const unsigned t0 = poly * lut_degree, t1 = (poly + 1) * lut_degree < num_lut_slots ? (poly + 1) * lut_degree : num_lut_slots;
        const unsigned u0 = poly * lu_degree, u1 = (poly + 1) * lu_degree < num_lu_slots ? (poly + 1) * lu_degree : num_lu_slots;
        uint64_t lut_prod = 1, lu_prod = 1, lu_sum_prods = 0, lut_sum_prods_mul = 0;
        for (unsigned i = t0; i < t1; i++) lut_prod = gl_mul(lut_prod, gl_sub(d[2], looked[i]));
        for (unsigned i = u0; i < u1; i++) lu_prod = gl_mul(lu_prod, gl_sub(d[2], looking[i]));
        for (unsigned i = u0; i < u1; i++) {
            uint64_t pr = 1;
            for (unsigned j = u0; j < u1; j++)
                if (j != i) pr = gl_mul(pr, gl_sub(d[2], looking[j]));
            lu_sum_prods = gl_add(lu_sum_prods, pr);
        }
        for (unsigned i = t0; i < t1; i++) {
            uint64_t pr = 1;
            for (unsigned j = t0; j < t1; j++)
                if (j != i) pr = gl_mul(pr, gl_sub(d[2], looked[j]));
            lut_sum_prods_mul = gl_add(lut_sum_prods_mul, gl_mul(wires[3 * i + 2], pr));
        }
        const uint64_t prev = poly == 0 ? z_gx[num_sldc - 1] : z_x[poly - 1];
        const uint64_t diff = gl_sub(z_x[poly], prev);
        out[k++] = gl_mul(sel[TRANS_SRE], gl_sub(gl_mul(lut_prod, diff), lut_sum_prods_mul));
        out[k++] = gl_mul(sel[TRANS_LDC], gl_add(gl_mul(lu_prod, diff), lu_sum_prods));
    }
    return k;
}

/* compute_lookup_polys for every challenge (prover.rs:489-636): out[nc * num_lookup_polys][n] value columns,
 * per challenge RE first, then the partial SLDC polynomials.  wires[num_wires][n]. */
void orc_lookup_polys(const orc_circuit *c, const uint64_t *wires, const uint64_t *deltas, uint64_t *out) {
    const orc_lookups *lk = c->lookups;
    const size_t n = (size_t)1 << c->degree_bits;
    const unsigned num_lu_slots = c->num_routed_wires / 2, num_lut_slots = c->num_routed_wires / 3;
    const unsigned max_lu_deg = lk->lookup_degree, npl = (num_lu_slots + max_lu_deg - 1) / max_lu_deg;
    const unsigned max_lut_deg = (num_lut_slots + npl - 1) / npl;
    const unsigned np1 = npl + 1;
    memset(out, 0, (size_t)c->num_challenges * np1 * n * 8);
#define WIRE(col, row) wires[(size_t)(col) * n + (row)]
    for (unsigned ch = 0; ch < c->num_challenges; ch++) {
        const uint64_t *d = deltas + 4 * ch;
        uint64_t *polys = out + (size_t)ch * np1 * n; /* polys[p * n + row] */
        for (unsigned t = 0; t < lk->num_luts; t++) {
            const size_t last_lu = lk->lookup_rows[3 * t], last_lut = lk->lookup_rows[3 * t + 1],
                         first_lut = lk->lookup_rows[3 * t + 2];
            for (size_t row = first_lut + 1; row-- > last_lut;) {
                uint64_t inv[num_lut_slots];
                uint64_t new_re = polys[row + 1];
                for (unsigned s = 0; s < num_lut_slots; s++) {
                    const uint64_t inp = WIRE(3 * s, row), outp = WIRE(3 * s + 1, row);
                    inv[s] = orc_gl_inv(gl_sub(d[2], gl_add(inp, gl_mul(d[0], outp))));
                    new_re = gl_add(gl_mul(new_re, d[3]), gl_add(inp, gl_mul(d[1], outp)));
                }
                polys[row] = orc_gl_canon(new_re);
                for (unsigned slot = 0; slot < npl; slot++) {
                    uint64_t sum = slot != 0 ? polys[(size_t)slot * n + row] : polys[(size_t)npl * n + row + 1];
                    const unsigned hi = (slot + 1) * max_lut_deg < num_lut_slots ? (slot + 1) * max_lut_deg : num_lut_slots;
                    for (unsigned s = slot * max_lut_deg; s < hi; s++)
                        sum = gl_add(sum, gl_mul(WIRE(3 * s + 2, row), inv[s]));
                    polys[(size_t)(slot + 1) * n + row] = orc_gl_canon(sum);
                }
            }
            for (size_t row = last_lut; row-- > last_lu;) {
                uint64_t inv[num_lu_slots];
                for (unsigned s = 0; s < num_lu_slots; s++)
                    inv[s] = orc_gl_inv(gl_sub(d[2], gl_add(WIRE(2 * s, row), gl_mul(d[0], WIRE(2 * s + 1, row)))));
                for (unsigned slot = 0; slot < npl; slot++) {
                    const uint64_t prev = slot == 0 ? polys[(size_t)npl * n + row + 1] : polys[(size_t)slot * n + row];
                    uint64_t sum = 0;
                    const unsigned hi = (slot + 1) * max_lu_deg < num_lu_slots ? (slot + 1) * max_lu_deg : num_lu_slots;
                    for (unsigned s = slot * max_lu_deg; s < hi; s++) sum = gl_add(sum, inv[s]);
                    polys[(size_t)(slot + 1) * n + row] = orc_gl_canon(gl_sub(prev, sum));
                }
            }
        }
    }
#undef WIRE
}

void orc_eval_vanishing_poly_base_lookup(const orc_circuit *c, uint64_t x, uint64_t z_h_x, const uint64_t *constants,
                                         const uint64_t *wires, const uint64_t *local_zs, const uint64_t *next_zs,
                                         const uint64_t *partial_products, const uint64_t *s_sigmas,
                                         const uint64_t *local_lookup_zs, const uint64_t *next_lookup_zs,
                                         const uint64_t *betas, const uint64_t *gammas, const uint64_t *deltas,
                                         const uint64_t *alphas, const uint64_t pih[4], uint64_t *res) {
    const unsigned nc = c->num_challenges, nr = c->num_routed_wires, np = c->num_partial_products;
    const unsigned ngc = circuit_num_gate_constraints(c);
    const int has_lookup = c->lookups != NULL && local_lookup_zs != NULL;
    const unsigned nlp = has_lookup ? c->lookups->num_lookup_polys : 0;
    const unsigned n_lookup_terms = has_lookup ? nc * (4 + c->lookups->num_luts + 2 * (nlp - 1)) : 0;
    const unsigned n_terms = nc + nc * (np + 1) + n_lookup_terms + ngc;
    uint64_t *terms = calloc(n_terms, 8);
    uint64_t *num = malloc(nr * 8), *den = malloc(nr * 8);
    /* L_0(x) = Z_H(x) / (n (x - 1)), zero_poly_coset.rs:93-96 */
    const uint64_t n_f = (uint64_t)1 << c->degree_bits;
    const uint64_t l_0_x = gl_mul(z_h_x, orc_gl_inv(gl_mul(n_f, gl_sub(x, 1))));
    uint64_t *pp_terms = terms + nc;
    for (unsigned i = 0; i < nc; i++) {
        const uint64_t z_x = local_zs[i], z_gx = next_zs[i];
        terms[i] = gl_mul(l_0_x, gl_sub(z_x, 1));
        for (unsigned j = 0; j < nr; j++) {
            const uint64_t s_id = gl_mul(c->k_is[j], x);
            num[j] = gl_add(gl_add(wires[j], gl_mul(betas[i], s_id)), gammas[i]);
            den[j] = gl_add(gl_add(wires[j], gl_mul(betas[i], s_sigmas[j])), gammas[i]);
        }
        /* check_partial_products, util/partial_products.rs:52-95 */
        const uint64_t *pp = partial_products + (size_t)i * np;
        for (unsigned w = 0; w <= np; w++) {
            const uint64_t prev = w == 0 ? z_x : pp[w - 1];
            const uint64_t next = w == np ? z_gx : pp[w];
            uint64_t pn = 1, pd = 1;
            for (unsigned j = w * c->max_degree; j < (w + 1) * c->max_degree && j < nr; j++) {
                pn = gl_mul(pn, num[j]);
                pd = gl_mul(pd, den[j]);
            }
            pp_terms[(size_t)i * (np + 1) + w] = gl_sub(gl_mul(prev, pn), gl_mul(next, pd));
        }
    }
    /* lookup constraints, vanishing_poly.rs:273-292: after the partial-product terms, challenge by challenge */
    if (has_lookup) {
        uint64_t *lt = terms + nc + nc * (np + 1);
        for (unsigned i = 0; i < nc; i++)
            lt += lookup_constraints(c, wires, local_lookup_zs + (size_t)i * nlp, next_lookup_zs + (size_t)i * nlp,
                                     constants + c->num_selectors, deltas + 4 * i, lt);
    }
    /* gate constraints, vanishing_poly.rs:700-726 */
    uint64_t *gate_terms = terms + nc + nc * (np + 1) + n_lookup_terms;
    const unsigned prefix = c->num_selectors + c->num_lookup_selectors;
    for (unsigned g = 0; g < c->num_gates; g++) {
        const orc_gate *gt = &c->gates[g];
        const uint64_t f = gate_filter(gt, constants[gt->selector_index], c->num_selectors > 1);
        gate_eval_add(gt, constants + prefix, wires, pih, f, gate_terms);
    }
    /* reduce_with_powers_multi, core/src/plonk_common.rs:68-85 */
    for (unsigned a = 0; a < nc; a++) {
        uint64_t cum = 0;
        for (unsigned t = n_terms; t-- > 0;) cum = gl_add(gl_mul(cum, alphas[a]), terms[t]);
        res[a] = orc_gl_canon(cum);
    }
    free(terms);
    free(num);
    free(den);
}

/* compute_quotient_polys (prover.rs:640-866).  The three oracles are given as their leaf-major
 * LDE rows (merkle_tree.leaves, [N][leaf_len], N = n << rate_bits; leaf i = point
 * g w_N^bitrev(i)); out = num_challenges coefficient vectors of length n << quotient_degree_bits. */
int orc_compute_quotient_polys(const orc_circuit *c, unsigned rate_bits, const uint64_t *cs_leaves,
                               size_t cs_len, const uint64_t *wires_leaves, size_t wires_len,
                               const uint64_t *zs_leaves, size_t zs_len, const uint64_t *betas,
                               const uint64_t *gammas, const uint64_t *alphas, const uint64_t pih[4],
                               uint64_t *out) {
    return orc_compute_quotient_polys_lookup(c, rate_bits, cs_leaves, cs_len, wires_leaves, wires_len, zs_leaves, zs_len,
                                             betas, gammas, NULL, alphas, pih, out);
}

/* With lookups the third oracle's rows are Z's, partial products, then the lookup polynomials
 * (lookup_range, circuit_data.rs:582-584). */
int orc_compute_quotient_polys_lookup(const orc_circuit *c, unsigned rate_bits, const uint64_t *cs_leaves,
                                      size_t cs_len, const uint64_t *wires_leaves, size_t wires_len,
                                      const uint64_t *zs_leaves, size_t zs_len, const uint64_t *betas,
                                      const uint64_t *gammas, const uint64_t *deltas, const uint64_t *alphas,
                                      const uint64_t pih[4], uint64_t *out) {
    const unsigned qdb = c->quotient_degree_bits, nc = c->num_challenges;
    if (qdb > rate_bits) return 1; /* prover.rs:662-666 */
    const unsigned lg_lde = c->degree_bits + qdb, lg_N = c->degree_bits + rate_bits;
    const size_t lde_size = (size_t)1 << lg_lde;
    const size_t step = (size_t)1 << (rate_bits - qdb), next_step = (size_t)1 << qdb;
    /* ZeroPolyOnCoset, zero_poly_coset.rs:24-66 */
    const size_t rate = (size_t)1 << qdb;
    uint64_t zh_eval[rate], zh_inv[rate];
    const uint64_t g_pow_n = orc_gl_pow(orc_gl_coset_shift(), (uint64_t)1 << c->degree_bits);
    const uint64_t v = orc_gl_primitive_root(qdb);
    for (size_t k = 0; k < rate; k++) {
        zh_eval[k] = gl_sub(gl_mul(g_pow_n, orc_gl_pow(v, k)), 1);
        zh_inv[k] = orc_gl_inv(zh_eval[k]);
    }
    const uint64_t w = orc_gl_primitive_root(lg_lde);
    const unsigned np = c->num_partial_products;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < lde_size; i++) {
        const uint64_t x = gl_mul(orc_gl_coset_shift(), orc_gl_pow(w, i));
        const size_t i_next = (i + next_step) % lde_size;
        /* get_lde_values(i, step), oracle.rs:286-291 */
        const size_t row = bitrev(i * step, lg_N), row_next = bitrev(i_next * step, lg_N);
        const uint64_t *cs = cs_leaves + row * cs_len;
        const uint64_t *zl = zs_leaves + row * zs_len, *zn = zs_leaves + row_next * zs_len;
        uint64_t res[16];
        const size_t lk_off = (size_t)nc * (1 + np);
        const int lk = c->lookups != NULL && deltas != NULL;
        orc_eval_vanishing_poly_base_lookup(c, x, zh_eval[i % rate], cs, wires_leaves + row * wires_len, zl, zn,
                                            zl + nc, cs + c->num_constants, lk ? zl + lk_off : NULL,
                                            lk ? zn + lk_off : NULL, betas, gammas, deltas, alphas, pih, res);
        for (unsigned a = 0; a < nc; a++) out[(size_t)a * lde_size + i] = gl_mul(res[a], zh_inv[i % rate]);
    }
    (void)np;
    /* values.coset_ifft(F::coset_shift()), field/src/polynomial/mod.rs:58-88 */
    const uint64_t g_inv = orc_gl_inv(orc_gl_coset_shift());
    for (unsigned a = 0; a < nc; a++) {
        uint64_t *p = out + (size_t)a * lde_size;
        orc_ifft(p, lg_lde);
        uint64_t s = 1;
        for (size_t i = 0; i < lde_size; i++) {
            p[i] = orc_gl_canon(gl_mul(p[i], s));
            s = gl_mul(s, g_inv);
        }
    }
    return 0;
}

/* wires_permutation_partial_products_and_zs for every challenge (prover.rs:402-480), laid out as
 * the prover commits them (prover.rs:255-261): out[(nc + nc*np)][n] = Z_0..Z_{nc-1}, then the
 * partial products of challenge 0, of challenge 1, ...  wires[num_wires][n], sigmas[nr][n]. */
void orc_partial_products_and_zs(const orc_circuit *c, const uint64_t *wires, const uint64_t *sigmas,
                                 const uint64_t *betas, const uint64_t *gammas, uint64_t *out) {
    const unsigned nc = c->num_challenges, nr = c->num_routed_wires, np = c->num_partial_products;
    const size_t n = (size_t)1 << c->degree_bits;
    const uint64_t w = orc_gl_primitive_root(c->degree_bits);
    uint64_t *chunk = malloc(n * (np + 1) * 8);
    for (unsigned ch = 0; ch < nc; ch++) {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) {
            const uint64_t x = orc_gl_pow(w, i);
            uint64_t q[nr];
            for (unsigned j = 0; j < nr; j++) {
                const uint64_t wv = wires[(size_t)j * n + i];
                const uint64_t num = gl_add(gl_add(wv, gl_mul(betas[ch], gl_mul(c->k_is[j], x))), gammas[ch]);
                const uint64_t den = gl_add(gl_add(wv, gl_mul(betas[ch], sigmas[(size_t)j * n + i])), gammas[ch]);
                q[j] = gl_mul(num, orc_gl_inv(den));
            }
            /* quotient_chunk_products, util/partial_products.rs:13-24 */
            for (unsigned k = 0; k <= np; k++) {
                uint64_t p = 1;
                for (unsigned j = k * c->max_degree; j < (k + 1) * c->max_degree && j < nr; j++) p = gl_mul(p, q[j]);
                chunk[i * (np + 1) + k] = p;
            }
        }
        /* partial_products_and_z_gx + the swap (prover.rs:466-473) */
        uint64_t z_x = 1;
        for (size_t i = 0; i < n; i++) {
            uint64_t acc = z_x;
            out[(size_t)ch * n + i] = orc_gl_canon(z_x);
            for (unsigned k = 0; k <= np; k++) {
                acc = gl_mul(acc, chunk[i * (np + 1) + k]);
                if (k < np) out[((size_t)nc + (size_t)ch * np + k) * n + i] = orc_gl_canon(acc);
            }
            z_x = acc;
        }
    }
    free(chunk);
}
