/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU oracle: a plain-C restatement of the reference's (Quantus-Network/qp-plonky2)
 * polynomial-commitment path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (qp-plonky2_b200/) never links, imports or calls it.
 *
 * Parity status: the Poseidon permutation is PINNED by the reference's four known-answer
 * vectors (core/src/poseidon_goldilocks.rs:455-490) and the bit-reversal by the reference's
 * table (plonky2/src/util/mod.rs:56-123).  Everything above the permutation (leaf hashes,
 * digests, caps, LDE values, FRI commitments) is "parity unpinned" by golden data: the
 * reference holds no vectors for it and no Rust toolchain exists in this image, so those
 * layers are pinned only structurally (FFT == naive evaluation, every-leaf Merkle round trip,
 * FRI fold consistency, and agreement with the independent big-integer restatement
 * oracle/pyref.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference
 * root).
 */
#ifndef PLONKY2_ORACLE_H
#define PLONKY2_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_P 0xFFFFFFFF00000001ULL
#define ORC_SPONGE_WIDTH 12
#define ORC_SPONGE_RATE 8

/* ---- field (field/src/goldilocks_field.rs) ---- */
uint64_t orc_gl_add(uint64_t a, uint64_t b);
uint64_t orc_gl_sub(uint64_t a, uint64_t b);
uint64_t orc_gl_mul(uint64_t a, uint64_t b);
uint64_t orc_gl_canon(uint64_t a);
uint64_t orc_gl_pow(uint64_t a, uint64_t e);
uint64_t orc_gl_inv(uint64_t a);
uint64_t orc_gl_inverse_2exp(unsigned k);
uint64_t orc_gl_primitive_root(unsigned k);
uint64_t orc_gl_coset_shift(void);

/* ---- quadratic extension F_p[X]/(X^2-7) ---- */
void orc_ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);

/* ---- bit reversal / FFT ---- */
void orc_reverse_index_bits(uint64_t *arr, size_t n, size_t elem_words);
void orc_fft(uint64_t *v, unsigned lg_n, unsigned zero_factor);
void orc_ifft(uint64_t *v, unsigned lg_n);
void orc_coset_fft(uint64_t *v, unsigned lg_n, uint64_t shift, unsigned zero_factor);
void orc_fft_naive(const uint64_t *in, uint64_t *out, unsigned lg_n, uint64_t shift);

/* ---- Poseidon ---- */
void orc_poseidon(uint64_t state[12]);
void orc_poseidon_naive(uint64_t state[12]);
void orc_hash_no_pad(const uint64_t *in, size_t len, uint64_t out[4]);
void orc_hash_leaf(const uint64_t *in, size_t len, uint64_t out[4]);
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);

/* ---- Merkle tree ---- */
/* leaves: leaf-major [n_leaves][leaf_len]; digests: 2*(n_leaves - 2^cap_height) x 4;
 * cap: 2^cap_height x 4.  Returns 0, or 1 if cap_height > log2(n_leaves) or n_leaves is not
 * a power of two (the reference's panics). */
int orc_merkle_tree_new(const uint64_t *leaves, size_t n_leaves, size_t leaf_len,
                        unsigned cap_height, uint64_t *digests, uint64_t *cap);
void orc_merkle_prove(size_t leaf_index, size_t n_leaves, unsigned cap_height,
                      const uint64_t *digests, uint64_t *siblings /* (lgN-h) x 4 */);
int orc_merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index,
                      const uint64_t *cap, unsigned cap_height, const uint64_t *siblings,
                      unsigned n_siblings);

/* ---- PolynomialBatch ---- */
/* values/coeffs are column-major [n_cols][n]; leaves_out leaf-major [N][n_cols + (salt?4:0)];
 * salt (optional, may be NULL) is column-major [4][N] in natural (pre bit-reversal) point order,
 * standing in for the reference's F::rand_vec salt columns.  scope_ms[4] receives the wall time
 * of "IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree". */
int orc_batch_from_values(const uint64_t *values, size_t n_cols, unsigned lg_n, unsigned rate_bits,
                          unsigned cap_height, const uint64_t *salt, uint64_t *coeffs_out,
                          uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out,
                          double scope_ms[4]);
int orc_batch_from_coeffs(const uint64_t *coeffs, size_t n_cols, unsigned lg_n, unsigned rate_bits,
                          unsigned cap_height, const uint64_t *salt, uint64_t *leaves_out,
                          uint64_t *digests_out, uint64_t *cap_out, double scope_ms[4]);

/* ---- Challenger (core/src/challenger.rs) ---- */
typedef struct {
    uint64_t sponge_state[12];
    uint64_t input_buffer[8];
    uint64_t output_buffer[8];
    uint32_t n_in, n_out;
} orc_challenger;
void orc_challenger_init(orc_challenger *c);
void orc_challenger_observe(orc_challenger *c, const uint64_t *elems, size_t n);
uint64_t orc_challenger_get(orc_challenger *c);

/* ---- FRI commit phase (plonky2/src/fri/prover.rs:85-143) ---- */
/* Computes reduction_arity_bits for ConstantArityBits(arity_bits, final_poly_bits)
 * (core/src/fri.rs:50-61).  Returns the count, writes into out[] (capacity 64). */
unsigned orc_fri_reduction_arity_bits(unsigned degree_bits, unsigned rate_bits, unsigned cap_height,
                                      unsigned arity_bits, unsigned final_poly_bits, unsigned *out);

/* coeffs/values: n ext elements (2 u64 each, interleaved).  Consumed (overwritten).
 * For round i: caps_out + i*(2^cap_height*4); if leaves_out[i]/digests_out[i] non-NULL they
 * receive that round's tree.  final_poly_out gets (n >> total_arity >> rate_bits) ext elems.
 * betas_out gets one ext element per round. */
int orc_fri_committed_trees(uint64_t *coeffs, uint64_t *values, unsigned lg_n, unsigned rate_bits,
                            unsigned cap_height, const unsigned *arity_bits, unsigned n_rounds,
                            orc_challenger *challenger, uint64_t *caps_out, uint64_t **leaves_out,
                            uint64_t **digests_out, uint64_t *betas_out, uint64_t *final_poly_out);

/* PoW grinding, deterministic rule = smallest witness (serial `find`,
 * maybe_rayon/src/lib.rs:254-259; plonky2/src/fri/prover.rs:159-208). */
uint64_t orc_fri_proof_of_work(orc_challenger *challenger, unsigned pow_bits);

/* ---- opening side (plonky2/src/fri/oracle.rs:129-165, plonky2/src/plonk/proof.rs:289-327) ---- */
/* PolynomialCoeffs::eval of a base-field polynomial at an F_p^2 point (Horner). */
void orc_eval_poly_ext(const uint64_t *coeffs, size_t n, const uint64_t point[2], uint64_t out[2]);
/* reduce_openings_to_unmasked_final_poly with the instance flattened the same way as the C ABI:
 * batch b has n_terms[b] (polynomial pointer, weight) terms, a point and a shift. */
void orc_reduce_openings(size_t n_batches, const size_t *n_terms, const uint64_t *const *term_polys,
                         const uint64_t *weights, const uint64_t *points, const uint64_t *shifts,
                         unsigned degree_log, uint64_t *final_out /* [n][2] */);

/* ---- plonk permutation argument and quotient (plonky2/src/plonk/prover.rs:402-480,640-866,
 *      plonky2/src/plonk/vanishing_poly.rs:166-330) -------------------------------------------- */
enum { ORC_GATE_NOOP = 0, ORC_GATE_CONSTANT = 1, ORC_GATE_PUBLIC_INPUT = 2, ORC_GATE_ARITHMETIC = 3,
       ORC_GATE_POSEIDON = 4, ORC_GATE_ARITHMETIC_EXT = 5, ORC_GATE_MUL_EXT = 6, ORC_GATE_BASE_SUM_2 = 7,
       ORC_GATE_RANDOM_ACCESS = 8, ORC_GATE_REDUCING = 9, ORC_GATE_REDUCING_EXT = 10, ORC_GATE_POSEIDON_MDS = 11,
       ORC_GATE_EXPONENTIATION = 12, ORC_GATE_COSET_INTERPOLATION = 13,
       ORC_GATE_LOOKUP = 14, ORC_GATE_LOOKUP_TABLE = 15 /* param: index of the table */ };
typedef struct {
    uint32_t kind;           /* ORC_GATE_* */
    uint32_t param;          /* num_consts (ConstantGate) / num_ops (Arithmetic*) / num_limbs / num_coeffs /
                              * num_power_bits; RandomAccess: bits | copies << 8 | extra constants << 16;
                              * CosetInterpolation: subgroup_bits | degree << 8 */
    uint32_t index;          /* position in the sorted gate list = selector value */
    uint32_t selector_index; /* SelectorsInfo.selector_indices[index] */
    uint32_t group_start, group_end; /* SelectorsInfo.groups[selector_index] */
} orc_gate;
/* The lookup argument's share of CommonCircuitData / ProverOnlyCircuitData (circuit_data.rs: luts,
 * num_lookup_polys; circuit_builder.rs:78-90 LookupWire): one LookupGate run + one LookupTableGate run per
 * table, rows "upside down" (gadgets/lookup.rs:80-160). */
typedef struct orc_lookups {
    uint32_t num_luts;
    const uint32_t *lut_lens;      /* [num_luts] */
    const uint16_t *const *luts;   /* luts[k] = [lut_lens[k]][2] (input, output) pairs */
    const uint32_t *lookup_rows;   /* [num_luts][3]: last_lu_row, last_lut_row, first_lut_row */
    uint32_t num_lookup_polys;     /* per challenge: RE + the partial SLDC polynomials */
    uint32_t lookup_degree;        /* lookup_accumulator_degree() = quotient_degree_factor - 1 */
} orc_lookups;
typedef struct {
    uint32_t degree_bits, quotient_degree_bits;
    uint32_t num_challenges, num_routed_wires, num_wires;
    uint32_t num_constants;  /* constant columns (selectors included) of the constants_sigmas oracle */
    uint32_t num_partial_products, max_degree; /* per challenge; chunk size = quotient_degree_factor */
    uint32_t num_selectors, num_lookup_selectors;
    uint32_t num_gates;
    const orc_gate *gates;
    const uint64_t *k_is;    /* [num_routed_wires] */
    const struct orc_lookups *lookups; /* NULL: the circuit has no lookup tables */
} orc_circuit;
unsigned orc_gate_num_constraints(const orc_gate *g);
void orc_eval_vanishing_poly_base(const orc_circuit *c, uint64_t x, uint64_t z_h_x, const uint64_t *constants,
                                  const uint64_t *wires, const uint64_t *local_zs, const uint64_t *next_zs,
                                  const uint64_t *partial_products, const uint64_t *s_sigmas,
                                  const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas,
                                  const uint64_t pih[4], uint64_t *res);
int orc_compute_quotient_polys(const orc_circuit *c, unsigned rate_bits, const uint64_t *cs_leaves,
                               size_t cs_len, const uint64_t *wires_leaves, size_t wires_len,
                               const uint64_t *zs_leaves, size_t zs_len, const uint64_t *betas,
                               const uint64_t *gammas, const uint64_t *alphas, const uint64_t pih[4],
                               uint64_t *out);
void orc_partial_products_and_zs(const orc_circuit *c, const uint64_t *wires, const uint64_t *sigmas,
                                 const uint64_t *betas, const uint64_t *gammas, uint64_t *out);
/* ---- lookup argument (plonky2/src/plonk/prover.rs:489-636, vanishing_poly.rs:29-52,330-520) ----
 * deltas: [num_challenges][4] = (ChallengeA, ChallengeB, ChallengeAlpha, ChallengeDelta) per challenge, i.e. the
 * reference's flat `deltas` vector (prover.rs:236-248).  The *_lookup forms take the circuit's lookups into
 * account (c->lookups != NULL); the plain forms above are the same functions with no lookup terms. */
void orc_lookup_polys(const orc_circuit *c, const uint64_t *wires, const uint64_t *deltas, uint64_t *out);
uint64_t orc_lut_poly_eval(const orc_circuit *c, unsigned lut_index, const uint64_t deltas4[4]);
void orc_eval_vanishing_poly_base_lookup(const orc_circuit *c, uint64_t x, uint64_t z_h_x, const uint64_t *constants,
                                         const uint64_t *wires, const uint64_t *local_zs, const uint64_t *next_zs,
                                         const uint64_t *partial_products, const uint64_t *s_sigmas,
                                         const uint64_t *local_lookup_zs, const uint64_t *next_lookup_zs,
                                         const uint64_t *betas, const uint64_t *gammas, const uint64_t *deltas,
                                         const uint64_t *alphas, const uint64_t pih[4], uint64_t *res);
int orc_compute_quotient_polys_lookup(const orc_circuit *c, unsigned rate_bits, const uint64_t *cs_leaves,
                                      size_t cs_len, const uint64_t *wires_leaves, size_t wires_len,
                                      const uint64_t *zs_leaves, size_t zs_len, const uint64_t *betas,
                                      const uint64_t *gammas, const uint64_t *deltas, const uint64_t *alphas,
                                      const uint64_t pih[4], uint64_t *out);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
