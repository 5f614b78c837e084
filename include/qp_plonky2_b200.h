/* qp_plonky2_b200 -- C ABI of the B200-native polynomial-commitment path.
 *
 * Drop-in boundary for the hot path of Quantus-Network/qp-plonky2:
 * the reference has no FFI, so each entry point below replaces one whole Rust function behind a
 * cargo feature; the Rust shim a maintainer would add is shown in INTEGRATION.md.  Reference
 * citations are relative to the reference root.
 *
 * Conventions
 *   - every function returns an int: QP_OK, or the code of the reference's panic / error
 *     condition it mirrors; nothing throws across the boundary.  qp_last_error() gives text.
 *   - field elements are Goldilocks u64 (p = 2^64 - 2^32 + 1); inputs may be non-canonical,
 *     every output is canonical (what the reference serialises, core/src/config.rs:104-109).
 *   - `space` says where a caller buffer lives: QP_HOST (pageable or pinned) or QP_DEVICE.
 *   - column-major inputs are [n_cols][n] (the reference's Vec<PolynomialValues>), rows/leaves
 *     are leaf-major [count][leaf_len] (its Vec<Vec<F>>).
 *   - handles own device memory (LDE values stay on the device, column-major in leaf order);
 *     they are immutable after creation.
 *   - there is NO CPU fallback: without a CUDA device every call fails with QP_ERR_CUDA.
 *   - threading: a context owns ONE stream plus per-call scratch (events, staging buffers, error
 *     text) and is NOT internally locked: every call that takes the context, or a handle created
 *     on it (readers such as qp_batch_prove / qp_batch_open_many included: they launch on that
 *     stream), must be serialised by the caller.  One context per prover thread; a batch that
 *     several provers share (the reference shares the constants/sigmas batch across concurrent
 *     prove() calls, circuit_data.rs:337-349) is shared by giving those provers the same
 *     context under a caller-side mutex, or one copy per context (qp_batch_deserialize).
 *   - stream ordering: the library orders its own work on the context's stream and synchronises
 *     that stream before a call returns host-visible results.  QP_DEVICE inputs must be complete
 *     (or produced on that same stream) before the call; QP_DEVICE outputs are ready once the
 *     call has returned only if the caller waits on that stream (qp_ctx_synchronize) or works on
 *     it.  Pass the producer's / consumer's cudaStream_t to qp_ctx_create to get that for free.
 */
#ifndef QP_PLONKY2_B200_H
#define QP_PLONKY2_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    QP_OK = 0,
    QP_ERR_CUDA = 1,           /* CUDA runtime failure (incl. no device) */
    QP_ERR_CAP_HEIGHT = 2,     /* assert cap_height <= log2(leaves.len()), plonky2/src/hash/merkle_tree.rs:164-170 */
    QP_ERR_NOT_POW2 = 3,       /* log2_strict panic, util/src/lib.rs:25-29 */
    QP_ERR_DEGREE_MISMATCH = 4,/* "Polynomial degrees inconsistent", plonky2/src/fri/oracle.rs:277 */
    QP_ERR_BAD_ARG = 5,        /* null pointer / out-of-range index (slice index panic) */
    QP_ERR_TOO_LARGE = 6,      /* exceeds the field's two-adicity (types.rs:281 assert) or device memory */
    QP_ERR_BLINDING_NO_SALT = 7,/* blinding requested without injected salt ("Cannot set blinding without rand feature", oracle.rs:238) */
    QP_ERR_UNSUPPORTED = 8     /* a feature of the reference outside this library's contract: refused, never approximated */
};

enum { QP_HOST = 0, QP_DEVICE = 1 };

#define QP_SALT_SIZE 4 /* plonky2/src/fri/oracle.rs:29 */

typedef struct qp_ctx qp_ctx;
typedef struct qp_batch qp_batch;
typedef struct qp_tree qp_tree;
typedef struct qp_fri qp_fri;

/* ---- context ------------------------------------------------------------------------------ */
/* One context per (process, GPU).  `stream` is a cudaStream_t (or NULL for a private stream);
 * max_lde_log bounds the twiddle table (log2 of the largest transform, <= 30: the transform
 * kernels index with 32 bits). */
int qp_ctx_create(int device, void* stream, unsigned max_lde_log, qp_ctx** out);
void qp_ctx_destroy(qp_ctx* ctx);
const char* qp_last_error(const qp_ctx* ctx);
int qp_ctx_synchronize(qp_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t qp_ctx_launch_count(const qp_ctx* ctx);

/* ---- PolynomialBatch (plonky2/src/fri/oracle.rs:33-40) -------------------------------------- */
/* PolynomialBatch::from_values (oracle.rs:168-190): values[n_cols][2^degree_log] -> iNTT ->
 * from_coeffs.  `salt` stands in for the reference's F::rand_vec blinding columns
 * (oracle.rs:259-263): [QP_SALT_SIZE][N] in natural point order, required iff blinding != 0.
 * Sharding (multi-GPU, one process per GPU): the batch holds LDE blocks
 * [block_first, block_first + block_count) of the 2^rate_bits coset blocks (= leaves
 * [block_first * n, ...) ), and the cap entries / digests of exactly those leaves.  Pass
 * block_first = 0, block_count = 2^rate_bits for the whole batch. */
int qp_batch_from_values(qp_ctx* ctx, const uint64_t* values, int space, size_t n_cols,
                         unsigned degree_log, unsigned rate_bits, int blinding, unsigned cap_height,
                         const uint64_t* salt, unsigned block_first, unsigned block_count,
                         qp_batch** out);
/* The same call on what the reference passes (`values: Vec<PolynomialValues<F>>`, oracle.rs:168-175):
 * cols[c] points at column c, 2^degree_log words of ordinary (pageable) host memory -- no flattening
 * copy on the caller's side.  Large inputs are staged through the context's pinned ring in groups of
 * 16 columns by a few host threads while the device transforms and hashes the previous group; `salt`
 * is host memory. */
int qp_batch_from_values_cols(qp_ctx* ctx, const uint64_t* const* cols, size_t n_cols,
                              unsigned degree_log, unsigned rate_bits, int blinding, unsigned cap_height,
                              const uint64_t* salt, unsigned block_first, unsigned block_count,
                              qp_batch** out);
/* ---- multi-device context: one process, a list of GPUs ----------------------------------------- */
/* The reference's from_values is one call in one process (oracle.rs:168-175); this is that call over
 * every GPU of the box.  The devices (a power of two of them, <= 2^rate_bits and <= 2^cap_height at commit
 * time) shard the commitment by coset = by cap subtree: columns are sharded for the
 * upload + inverse transform, coefficient pieces move device to device as peer copies over NVLink in
 * column order, device i extends all columns to coset blocks [i 2^r / D, (i + 1) 2^r / D) and hashes exactly
 * those leaves.  Shard i is an ordinary qp_batch living on devices[i] (its leaf indices are local: global
 * leaf = i * (N / D) + local), to be read with the qp_batch_* getters through qp_mctx_ctx(m, i)'s stream. */
typedef struct qp_mctx qp_mctx;
typedef struct qp_mbatch qp_mbatch;
int qp_mctx_create(const int* devices, unsigned n_devices, unsigned max_lde_log, qp_mctx** out);
void qp_mctx_destroy(qp_mctx* m);
unsigned qp_mctx_num_devices(const qp_mctx* m);
qp_ctx* qp_mctx_ctx(qp_mctx* m, unsigned i);
const char* qp_mctx_last_error(const qp_mctx* m);
/* cols[c]: column c, 2^degree_log words of ordinary host memory; salt: host, [QP_SALT_SIZE][N], iff blinding */
int qp_mbatch_from_values_cols(qp_mctx* m, const uint64_t* const* cols, size_t n_cols, unsigned degree_log,
                               unsigned rate_bits, int blinding, unsigned cap_height, const uint64_t* salt,
                               qp_mbatch** out);
/* from_values (is_coeffs = 0) / from_coeffs (1) on a matrix [n_cols][2^degree_log] resident on devices[0] and
 * complete before the call (the Z / partial-product columns and the quotient chunks of a proof are computed
 * there): device 0 owns every piece, the other devices receive them as peer copies.  `salt`: host. */
int qp_mbatch_from_device(qp_mctx* m, const uint64_t* dev0_data, int is_coeffs, size_t n_cols, unsigned degree_log,
                          unsigned rate_bits, int blinding, unsigned cap_height, const uint64_t* salt,
                          qp_mbatch** out);
/* the whole cap, [2^cap_height][4], host */
int qp_mbatch_cap(const qp_mbatch* b, uint64_t* out);
unsigned qp_mbatch_num_shards(const qp_mbatch* b);
qp_batch* qp_mbatch_shard(qp_mbatch* b, unsigned i);   /* owned by the mbatch */
void qp_mbatch_free(qp_mbatch* b);

/* PolynomialBatch::from_coeffs (oracle.rs:193-223). */
int qp_batch_from_coeffs(qp_ctx* ctx, const uint64_t* coeffs, int space, size_t n_cols,
                         unsigned degree_log, unsigned rate_bits, int blinding, unsigned cap_height,
                         const uint64_t* salt, unsigned block_first, unsigned block_count,
                         qp_batch** out);
void qp_batch_free(qp_batch* b);
/* from_coeffs in pieces, for callers whose coefficient columns arrive over time (the multi-GPU
 * path: the LDE of the columns already gathered overlaps the all-gather of the rest):
 *   qp_batch_begin -> qp_batch_put_coeffs(columns [c0, c0 + count)) ... -> qp_batch_end (salt iff
 *   blinding; builds the tree).  Every column must be put exactly once before qp_batch_end; the
 *   result is the batch qp_batch_from_coeffs would have produced.  Getters are valid after end. */
int qp_batch_begin(qp_ctx* ctx, size_t n_cols, unsigned degree_log, unsigned rate_bits, int blinding,
                   unsigned cap_height, unsigned block_first, unsigned block_count, qp_batch** out);
int qp_batch_put_coeffs(qp_batch* b, const uint64_t* coeffs, int space, size_t c0, size_t count);
int qp_batch_end(qp_batch* b, const uint64_t* salt, int space);
/* The same without the copy: qp_batch_coeffs_slot is the device address of column c0's coefficients
 * inside the batch under construction (2^degree_log words per column, columns contiguous), for a
 * producer that writes them in place -- an inverse transform (qp_ifft_columns with a QP_DEVICE output), a
 * collective's receive buffer; qp_batch_extend_columns says "columns [c0, c0 + count) are in place": it
 * runs their LDE and, if `absorb`, advances the leaf sponges over every complete 8-column chunk of the
 * column prefix extended so far, so that leaf hashing overlaps the arrival of later columns (the sponge
 * of hash_leaf, core/src/hashing.rs:150-168, absorbs the columns strictly in order).  The writes into the
 * slot must be ordered before the call on the context's stream. */
uint64_t* qp_batch_coeffs_slot(qp_batch* b, size_t c0);
int qp_batch_extend_columns(qp_batch* b, size_t c0, size_t count, int absorb);

/* iNTT of columns only (the "IFFT" scope, oracle.rs:176-180): values[n_cols][n] -> coeffs.
 * Used by the multi-GPU path, which shards columns for the iNTT and cosets for the rest. */
int qp_ifft_columns(qp_ctx* ctx, const uint64_t* values, int space, size_t n_cols,
                    unsigned degree_log, uint64_t* coeffs_out, int out_space);

/* merkle_tree.cap: (local) cap entries, [n_local_cap][4]; n_local_cap = 2^cap_height *
 * block_count / 2^rate_bits (whole batch: 2^cap_height). */
int qp_batch_cap(const qp_batch* b, uint64_t* out, int space);
size_t qp_batch_cap_len(const qp_batch* b);
/* .polynomials: coefficients [n_cols][n], natural order. */
int qp_batch_coeffs(const qp_batch* b, uint64_t* out, int space);
/* merkle_tree.digests in the reference layout (local slice), [n_digests][4]. */
int qp_batch_digests(const qp_batch* b, uint64_t* out, int space);
size_t qp_batch_digests_len(const qp_batch* b);
/* merkle_tree.leaves[first .. first+count): leaf-major rows [count][leaf_len], salt included. */
int qp_batch_leaves(const qp_batch* b, size_t first, size_t count, uint64_t* out, int space);
size_t qp_batch_leaf_len(const qp_batch* b);
/* write_polynomial_batch / read_polynomial_batch (plonky2/src/util/serialization/mod.rs:1803-1822,
 * 758-784): the byte form in which CircuitData stores its constants/sigmas commitment
 * (circuit_data.rs:371-391) -- polynomials, Merkle tree (leaves, digests, cap), degree_log, rate_bits,
 * blinding; u64 little-endian, canonical field elements.  Whole (unsharded) batches only.
 * Deserialising rebuilds the device batch from the polynomials (and the salt found in the leaves)
 * and fails with QP_ERR_BAD_ARG if the bytes are truncated, inconsistent, or their cap is not the
 * cap of their polynomials. */
/* out = {polynomials.len(), degree_log, rate_bits, cap height, blinding} */
int qp_batch_describe(const qp_batch* b, uint64_t out[5]);
size_t qp_batch_serialized_len(const qp_batch* b);
int qp_batch_serialize(const qp_batch* b, uint8_t* out, size_t capacity);
int qp_batch_deserialize(qp_ctx* ctx, const uint8_t* data, size_t len, qp_batch** out, size_t* consumed);
/* get_lde_values(index, step) (oracle.rs:286-291): row at bit-reversed index*step, salt stripped;
 * out holds n_cols elements (host). */
int qp_batch_get_lde_values(const qp_batch* b, size_t index, size_t step, uint64_t* out);
/* Rows for many leaf indices at once (query openings, fri/prover.rs:246-249). */
int qp_batch_get_leaves(const qp_batch* b, const uint64_t* leaf_indices, unsigned n, uint64_t* out);
/* MerkleTree::prove (merkle_tree.rs:201-207): siblings bottom-up, [lg N - cap_height][4]. */
int qp_batch_prove(const qp_batch* b, size_t leaf_index, uint64_t* siblings_out);
/* The same for n leaves in one call: siblings_out[n][lg N - cap_height][4]. */
int qp_batch_prove_many(const qp_batch* b, const uint64_t* leaf_indices, unsigned n, uint64_t* siblings_out);
/* qp_batch_get_leaves + qp_batch_prove_many in one round trip (fri_prover_query_rounds, fri/prover.rs:246-253). */
int qp_batch_open_many(const qp_batch* b, const uint64_t* leaf_indices, unsigned n, uint64_t* rows_out,
                       uint64_t* siblings_out);
/* TimingTree scopes (oracle.rs:176-214), milliseconds of device time:
 * [0] "IFFT", [1] "FFT + blinding", [2] "transpose LDEs" (always 0: fused away), [3] "build Merkle tree". */
int qp_batch_timing(const qp_batch* b, double ms[4]);
/* Device time (CUDA events on the launching stream) per kernel group, milliseconds:
 * [0] iNTT passes, [1] LDE passes, [2] leaf_hash_kernel, [3] tree_level_kernel launches. */
int qp_batch_kernel_timing(const qp_batch* b, double ms[4]);
/* Raw device pointers for device-resident consumers (quotient / openings kernels). */
const uint64_t* qp_batch_device_lde(const qp_batch* b);    /* [leaf_len][N_local], leaf order */
const uint64_t* qp_batch_device_coeffs(const qp_batch* b); /* [n_cols][n] */

/* ---- MerkleTree (plonky2/src/hash/merkle_tree.rs:163-207) ---------------------------------- */
/* MerkleTree::new(leaves, cap_height) on caller-provided leaf-major rows. */
int qp_merkle_tree_new(qp_ctx* ctx, const uint64_t* leaves, int space, size_t n_leaves,
                       size_t leaf_len, unsigned cap_height, qp_tree** out);
/* One shard of MerkleTree::new for the multi-GPU form (the reference parallelises over cap subtrees,
 * merkle_tree.rs:85-119): shard `shard` of `n_shards` (a power of two <= 2^cap_height) passes the leaves of
 * ITS subtrees -- rows [shard * n_leaves_total / n_shards, ...) -- and gets their block of `digests` and
 * their 2^cap_height / n_shards cap entries (leaf indices of the getters are local to the shard). */
int qp_merkle_tree_new_shard(qp_ctx* ctx, const uint64_t* shard_leaves, int space, size_t n_leaves_total,
                             size_t leaf_len, unsigned cap_height, unsigned shard, unsigned n_shards,
                             qp_tree** out);
void qp_tree_free(qp_tree* t);
int qp_tree_cap(const qp_tree* t, uint64_t* out, int space);
int qp_tree_digests(const qp_tree* t, uint64_t* out, int space);
size_t qp_tree_digests_len(const qp_tree* t);
int qp_tree_prove(const qp_tree* t, size_t leaf_index, uint64_t* siblings_out);
int qp_tree_get(const qp_tree* t, size_t leaf_index, uint64_t* out); /* MerkleTree::get */

/* ---- BatchFriOracle::from_values / from_coeffs (plonky2/src/batch_fri/oracle.rs:78-160) ----------
 * The commitment of batch FRI: polynomials of non-increasing power-of-two length; every run of equal
 * length is one PolynomialBatch-style LDE (leaf order), and the runs are the matrices of one
 * BatchMerkleTree (below), tallest first.  polys[i]: 2^degree_bits[i] values / coefficients.
 * Errors: QP_ERR_BAD_ARG (empty, lengths increasing, blinding != 0: salt injection is not implemented
 * for batch oracles), QP_ERR_CAP_HEIGHT, QP_ERR_TOO_LARGE. */
typedef struct qp_batch_fri qp_batch_fri;
int qp_batch_fri_from_values(qp_ctx* ctx, const uint64_t* const* polys, const unsigned* degree_bits, size_t n_polys,
                             int space, unsigned rate_bits, int blinding, unsigned cap_height, qp_batch_fri** out);
int qp_batch_fri_from_coeffs(qp_ctx* ctx, const uint64_t* const* polys, const unsigned* degree_bits, size_t n_polys,
                             int space, unsigned rate_bits, int blinding, unsigned cap_height, qp_batch_fri** out);
void qp_batch_fri_free(qp_batch_fri* o);
size_t qp_batch_fri_num_groups(const qp_batch_fri* o);                 /* = degree_bits.len() after dedup */
/* group g (tallest first): its degree_bits and number of polynomials (= leaf width of its matrix) */
int qp_batch_fri_group(const qp_batch_fri* o, size_t g, unsigned* degree_bits, size_t* n_polys);
int qp_batch_fri_coeffs(const qp_batch_fri* o, size_t g, uint64_t* out, int space);   /* [n_polys][2^degree_bits] */
int qp_batch_fri_cap(const qp_batch_fri* o, uint64_t* out, int space);                /* [2^cap_height][4] */
size_t qp_batch_fri_digests_len(const qp_batch_fri* o);
int qp_batch_fri_digests(const qp_batch_fri* o, uint64_t* out, int space);
/* group g as a PolynomialBatch view (its coefficients for qp_opening_term / qp_batch_eval_polys, its LDE rows);
 * owned by the oracle, valid until qp_batch_fri_free */
const qp_batch* qp_batch_fri_group_batch(const qp_batch_fri* o, size_t g);
/* batch_merkle_tree.open_batch(leaf_index): log2(N_0) - cap_height siblings */
int qp_batch_fri_open(const qp_batch_fri* o, size_t leaf_index, uint64_t* siblings_out);
/* batch_merkle_tree.values(leaf_index): every group's row, concatenated (host) */
int qp_batch_fri_values(const qp_batch_fri* o, size_t leaf_index, uint64_t* out);

/* ---- BatchMerkleTree::new (plonky2/src/hash/batch_merkle_tree.rs:40-130) ---------------------
 * One tree over several matrices of strictly decreasing power-of-two heights (the oracle of
 * batch FRI): the tree over the tallest matrix is capped at the height of the next one, whose rows
 * are hashed together with those cap entries (hash_leaf(digest || row)), and so on down to
 * `cap_height`.  matrices[i]: row-major [heights[i]][widths[i]].  Errors are the reference's
 * asserts: QP_ERR_BAD_ARG (empty / heights not strictly decreasing), QP_ERR_NOT_POW2,
 * QP_ERR_CAP_HEIGHT (cap_height > log2 of the last height). */
typedef struct qp_batch_tree qp_batch_tree;
int qp_batch_merkle_tree_new(qp_ctx* ctx, const uint64_t* const* matrices, int space, const size_t* heights,
                             const size_t* widths, size_t n_matrices, unsigned cap_height, qp_batch_tree** out);
void qp_batch_tree_free(qp_batch_tree* t);
int qp_batch_tree_cap(const qp_batch_tree* t, uint64_t* out, int space);        /* [2^cap_height][4] */
size_t qp_batch_tree_digests_len(const qp_batch_tree* t);                       /* 2 (heights[0] - 2^cap_height) */
int qp_batch_tree_digests(const qp_batch_tree* t, uint64_t* out, int space);    /* the stages' digests, concatenated */
/* open_batch (batch_merkle_tree.rs:133-153): log2(heights[0]) - cap_height siblings, 4 words each */
int qp_batch_tree_open(const qp_batch_tree* t, size_t leaf_index, uint64_t* siblings_out);
/* values (batch_merkle_tree.rs:155-164): row leaf_index >> (log2 heights[0] - log2 heights[i]) of every
 * matrix, concatenated (sum of widths words) */
int qp_batch_tree_values(const qp_batch_tree* t, size_t leaf_index, uint64_t* out);

/* ---- Poseidon primitives (core/src/poseidon.rs:599-609, hashing.rs) ------------------------- */
/* Batch of width-12 permutations, states [count][12] (in place). */
int qp_poseidon_permute(qp_ctx* ctx, uint64_t* states, int space, size_t count);

/* ---- transforms (field/src/fft.rs, polynomial/mod.rs) ---------------------------------------- */
/* coset_fft_with_options(shift) of `n_vec` polynomials [n_vec][2^lg_n] -> values in NATURAL
 * order (bit_reversed = 0) or bit-reversed order (bit_reversed = 1).  shift = 1 gives fft. */
int qp_coset_fft(qp_ctx* ctx, const uint64_t* coeffs, int space, size_t n_vec, unsigned lg_n,
                 uint64_t shift, int bit_reversed, uint64_t* out, int out_space);

/* ---- FRI commit phase (plonky2/src/fri/prover.rs:85-143) ------------------------------------ */
/* The transcript (Challenger, core/src/challenger.rs) stays with the caller: each round is
 *   qp_fri_commit_round -> cap   (observe_cap, get_extension_challenge on the caller's side)
 *   qp_fri_fold_round(beta)      (fold by beta; unless last: coset FFT with shift^arity)
 * Inputs are F_p^2 arrays [n][2] (interleaved c0, c1): the LDE polynomial coefficients and
 * its values in natural order (fri_proof's two arguments, prover.rs:24-31). */
int qp_fri_begin(qp_ctx* ctx, const uint64_t* coeffs_ext, const uint64_t* values_ext, int space,
                 unsigned lg_n, unsigned rate_bits, unsigned cap_height, qp_fri** out);
int qp_fri_commit_round(qp_fri* f, unsigned arity_bits, uint64_t* cap_out /* [2^cap_height][4] host */);
int qp_fri_fold_round(qp_fri* f, const uint64_t beta[2], int is_last);
/* Batch FRI (plonky2/src/batch_fri/prover.rs:122-141): after a fold round that is not the last, when the next
 * polynomial of the batch has exactly the current length:
 *     final_values = final_values * beta + values[polynomial_index];  final_coeffs = final_values.coset_ifft(shift)
 * `lower` = the FRI state of that polynomial (qp_fri_begin / qp_fri_begin_from_openings), no round committed.
 * Errors: QP_ERR_DEGREE_MISMATCH (lengths differ), QP_ERR_BAD_ARG (called out of order). */
int qp_fri_mix_values(qp_fri* f, const qp_fri* lower, const uint64_t beta[2]);
/* log2 of the current codeword length (lde_bits minus the arities folded so far) */
unsigned qp_fri_domain_bits(const qp_fri* f);
/* Final polynomial: coeffs truncated to len >> rate_bits (prover.rs:138-141); returns its
 * length in *len_out (ext elements) and writes [len][2] to out (host). */
int qp_fri_final_poly(qp_fri* f, uint64_t* out, size_t* len_out);
/* Tree of commit round r: leaf (x_index >> arity_bits) flattened [2*arity], and its path. */
int qp_fri_tree_get(const qp_fri* f, unsigned round, size_t leaf_index, uint64_t* out);
int qp_fri_tree_prove(const qp_fri* f, unsigned round, size_t leaf_index, uint64_t* siblings_out);
/* n openings of one commit-phase tree in one call (fri_prover_query_round, prover.rs:250-258):
 * leaves_out[n][2*arity], siblings_out[n][layers][4]. */
int qp_fri_tree_open_many(const qp_fri* f, unsigned round, const uint64_t* leaf_indices, unsigned n,
                          uint64_t* leaves_out, uint64_t* siblings_out);
int qp_fri_tree_digests(const qp_fri* f, unsigned round, uint64_t* out, int space);
size_t qp_fri_tree_digests_len(const qp_fri* f, unsigned round);
unsigned qp_fri_num_rounds(const qp_fri* f);
void qp_fri_free(qp_fri* f);

/* ---- opening side of prove_openings (plonky2/src/fri/oracle.rs:320-358) --------------------- */
/* PolynomialCoeffs::eval of every polynomial of the batch at an F_p^2 point -- the work of
 * OpeningSet::new (plonky2/src/plonk/proof.rs:289-327, field/src/polynomial/mod.rs:155-160).
 * out: [n_cols][2] (host). */
int qp_batch_eval_polys(const qp_batch* b, const uint64_t point[2], uint64_t* out);

/* reduce_openings_to_unmasked_final_poly (oracle.rs:129-165) followed by the padded coset FFT
 * (oracle.rs:329-343), entirely on the device; the result is a FRI state ready for
 * qp_fri_commit_round.  The caller flattens the FriInstanceInfo into, per batch (= per opening
 * point), a list of (polynomial, F_p^2 weight) terms: an opening expression at position i of the
 * batch contributes alpha^i * coefficient (One -> 1, PointPower(k) -> point^k, Constant(c) -> c)
 * for each of its terms (core/src/fri_structure.rs:59-108, core/src/reducing.rs:63-72).  `shift`
 * is alpha^(number of expressions of this batch), applied to the running sum before this batch's
 * quotient is added (shift_poly, reducing.rs:94-97).  All batches must have the same degree. */
typedef struct {
    const qp_batch* batch;   /* oracle holding the polynomial */
    size_t poly_index;       /* column inside that oracle */
    uint64_t weight[2];
} qp_opening_term;
typedef struct {
    uint64_t point[2];
    const qp_opening_term* terms;
    size_t n_terms;
    uint64_t shift[2];
} qp_opening_batch;
int qp_fri_begin_from_openings(qp_ctx* ctx, const qp_opening_batch* batches, size_t n_batches,
                               unsigned degree_log, unsigned rate_bits, unsigned cap_height, qp_fri** out);
/* The unmasked final polynomial produced by the call above: [2^degree_log][2] (host). */
int qp_fri_initial_coeffs(const qp_fri* f, uint64_t* out);

/* Proof-of-work grinding (plonky2/src/fri/prover.rs:159-208).  `state12` is the duplex state
 * with the pending inputs already written (duplex_intermediate_state), `witness_pos` the lane
 * the candidate goes to.  Returns the SMALLEST witness whose response (lane 7 after the
 * permutation) has >= min_leading_zeros leading zero bits -- the deterministic rule of the
 * serial `find` (maybe_rayon/src/lib.rs:254-259). */
int qp_fri_proof_of_work(qp_ctx* ctx, const uint64_t state12[12], unsigned witness_pos,
                         unsigned min_leading_zeros, uint64_t* witness_out);

/* ---- Plonk permutation argument and quotient polynomials ------------------------------------ */
/* What CommonCircuitData / ProverOnlyCircuitData contribute to these two steps
 * (plonky2/src/plonk/circuit_data.rs:412-470), created once per circuit and kept on the device.
 * The gate set (common_data.gates + selectors_info) arrives as a straight-line constraint program
 * over F_p: 64-bit words  op | dst << 8 | a << 16 | b << 24 | c << 32  (dst, a, b: registers < 256;
 * c: a column, a pool slot or a constraint index) with
 *   1 LDW dst = wire c            2 LDK dst = constants_sigmas polynomial c
 *   3 LDP dst = public_inputs_hash[c]   4 LDI dst = pool[c]
 *   5 ADD  6 SUB  7 MUL   (dst = r[a] op r[b])
 *   10 MULI dst = r[a] * pool[c]     11 ADDI dst = r[a] + pool[c]     13 FMAI dst = r[a] * pool[c] + r[b]
 *   8 EMIT a, c  constraint number c of the current gate has the value r[a]
 *   9 GATE a     end of a gate; r[a] holds its filter (compute_filter, gates/gate.rs:326-333)
 *   12 WAIT      all column loads issued so far have arrived.  LDW / LDK are asynchronous: after
 *                issuing one, the device waits until at most 3 loads are in flight, so a program
 *                must not read a loaded register before 3 further loads or a WAIT have been issued
 *                (a program that puts a WAIT after every load is always valid).
 *   14 NATIVE_POSEIDON  (a segment of its own: this one word) a PoseidonGate evaluated by the device's
 *                native kernel instead of the interpreter: dst = end of the gate's selector group,
 *                a = the gate's index in common_data.gates (= its selector value), b = start of the
 *                group, c = selector polynomial | (num_selectors > 1) << 16.  At most one per program.
 *   0 END        end of a segment.  A program may consist of several self-contained segments
 *                (no register is live across an END); their contributions add up, and for small
 *                circuits different thread blocks evaluate different segments.
 * The program evaluates evaluate_gate_constraints_base_batch (vanishing_poly.rs:700-726); a Rust
 * shim produces it by running Gate::eval_unfiltered_base_one over a recording field type, the
 * Python mirror (qp-plonky2_b200/plonk.py) from its own gate classes. */
/* One lookup table and where its gates sit (LookupTable = Vec<(u16, u16)>, gates/lookup_table.rs:33;
 * LookupWire, plonk/circuit_builder.rs:78-90: rows are "upside down" -- LookupGate rows
 * [last_lu_row, last_lut_row), LookupTableGate rows [last_lut_row, first_lut_row], gadgets/lookup.rs:80-160). */
typedef struct {
    const uint16_t* table;           /* [len][2]: (input, output) pairs (host) */
    size_t len;
    uint32_t last_lu_row, last_lut_row, first_lut_row;
} qp_lookup_table;
typedef struct {
    uint32_t degree_bits;            /* common_data.degree_bits() */
    uint32_t quotient_degree_bits;   /* log2_ceil(quotient_degree_factor) */
    uint32_t num_challenges;         /* config.num_challenges (<= 4) */
    uint32_t num_routed_wires, num_wires;
    uint32_t num_constants;          /* common_data.num_constants: sigmas start at this polynomial */
    uint32_t num_partial_products;   /* per challenge */
    uint32_t max_degree;             /* permutation_partial_product_degree() = quotient_degree_factor */
    const uint64_t* k_is;            /* [num_routed_wires] (host) */
    const uint64_t* sigmas;          /* prover_data.sigmas as columns [num_routed_wires][n]; may be NULL */
    int sigmas_space;
    const uint64_t* program;         /* (host) */
    size_t program_len;
    const uint64_t* pool;            /* field constants of the program (host) */
    size_t pool_len;
    uint32_t program_regs;           /* registers the program uses */
    /* Lookup argument (compute_lookup_polys, plonky2/src/plonk/prover.rs:489-636; the lookup terms of
     * vanishing_poly.rs:273-292,521-680).  A circuit without lookup tables has all of these zero / NULL.
     * num_lookup_polys = common_data.num_lookup_polys (RE + the partial SLDC polynomials, per challenge:
     * 1 + ceil((num_routed_wires / 2) / (max_degree - 1)), circuit_builder.rs:1284-1290);
     * num_lookup_selectors = 4 + n_luts (gates/selectors.rs:27-75), stored right after the num_selectors
     * selector polynomials of the constants (circuit_builder.rs:1183-1194); luts = common_data.luts with
     * prover_data.lookup_rows (one LookupWire per table).  Inconsistent declarations are QP_ERR_BAD_ARG. */
    uint32_t num_lookup_polys;
    uint32_t num_lookup_selectors;
    uint32_t num_selectors;          /* selectors_info.num_selectors(); read only when there are lookups */
    const qp_lookup_table* luts;     /* (host) */
    size_t n_luts;
} qp_circuit_desc;
typedef struct qp_circuit qp_circuit;
int qp_circuit_create(qp_ctx* ctx, const qp_circuit_desc* desc, qp_circuit** out);
void qp_circuit_free(qp_circuit* c);
/* The scalar fields the circuit was created with (pointers are NULL). */
int qp_circuit_describe(const qp_circuit* c, qp_circuit_desc* out);
int qp_circuit_has_sigmas(const qp_circuit* c);
/* Stream-ordered device scratch ([n_words] u64) for host-side drivers that keep intermediate
 * polynomials on the device between calls (qp_prove). */
int qp_dev_alloc(qp_ctx* ctx, size_t n_words, uint64_t** out);
void qp_dev_free(qp_ctx* ctx, uint64_t* p);
/* Copy n_words u64 between host and device buffers on the context's stream (synchronous). */
int qp_memcpy(qp_ctx* ctx, uint64_t* dst, int dst_space, const uint64_t* src, int src_space, size_t n_words);
/* device memory of src_ctx's GPU -> device memory of dst_ctx's GPU (peer copy), complete at return */
int qp_memcpy_peer(qp_ctx* dst_ctx, uint64_t* dst, qp_ctx* src_ctx, const uint64_t* src, size_t n_words);
/* all_wires_permutation_partial_products (plonky2/src/plonk/prover.rs:402-480) followed by the
 * Z-first ordering of prover.rs:255-261.  wires: witness columns [>= num_routed_wires][n] (values on
 * H); betas, gammas: [num_challenges] (host).  out: [(nc + nc * num_partial_products)][n] value
 * columns = Z_0 .. Z_{nc-1}, then the partial products of challenge 0, 1, ... -- the input of the
 * second PolynomialBatch::from_values of prove(). */
int qp_circuit_partial_products_and_zs(qp_circuit* c, const uint64_t* wires, int space, const uint64_t* betas,
                                       const uint64_t* gammas, uint64_t* out, int out_space);
/* compute_all_lookup_polys (plonky2/src/plonk/prover.rs:489-636) for a circuit with lookup tables.  wires: the
 * witness columns [>= num_routed_wires][n]; deltas: [num_challenges][4] = (ChallengeA, ChallengeB, ChallengeAlpha,
 * ChallengeDelta) per challenge -- the reference's flat `deltas` (prover.rs:236-248: betas, gammas, then the
 * additional challenges).  out: [num_challenges * num_lookup_polys][n] value columns, per challenge RE first,
 * then the partial SLDC polynomials: the columns prove() appends to the Z / partial-product batch (prover.rs:265-271). */
int qp_circuit_lookup_polys(qp_circuit* c, const uint64_t* wires, int space, const uint64_t* deltas, uint64_t* out,
                            int out_space);
/* The lookup challenges of the proof in progress: the quotient entry points below evaluate the lookup terms with
 * them (and with get_lut_poly(..).eval(delta) of every table, prover.rs:687-716, computed here).  Must be called
 * before them when the circuit has lookup tables. */
int qp_circuit_set_lookup_challenges(qp_circuit* c, const uint64_t* deltas);
/* compute_quotient_polys (plonky2/src/plonk/prover.rs:640-866; with lookup tables zs_partial_products also holds
 * the lookup polynomials, lookup_range circuit_data.rs:582): evaluates the
 * vanishing polynomial on the coset of size n << quotient_degree_bits from the three oracles' LDEs
 * (which stay on the device), divides by Z_H and interpolates (coset_ifft).  out:
 * [num_challenges][n << quotient_degree_bits] coefficients; read as [num_challenges *
 * 2^quotient_degree_bits][n] they are the chunks prove() commits with from_coeffs (prover.rs:309-333). */
int qp_circuit_compute_quotient_polys(qp_circuit* c, const qp_batch* constants_sigmas, const qp_batch* wires,
                                      const qp_batch* zs_partial_products, const uint64_t* betas,
                                      const uint64_t* gammas, const uint64_t* alphas,
                                      const uint64_t public_inputs_hash[4], uint64_t* out, int out_space);
/* The same in two steps for a coset-sharded (multi-GPU) commitment: every term of the vanishing polynomial at a
 * point needs that point and the next row of the same coset (prover.rs:679,750), so device i evaluates the
 * positions of ITS shard from ITS shards of the three oracles (step 1: out_dev = [nc][*pos_count] on that device,
 * leaf order, local; *pos_first = where the block belongs), the blocks are gathered side by side on one device and
 * step 2 turns the whole domain's values (leaf order, [nc][n << quotient_degree_bits], consumed) into the
 * quotient coefficients there. */
int qp_circuit_quotient_values_shard(qp_circuit* c, const qp_batch* constants_sigmas, const qp_batch* wires,
                                     const qp_batch* zs_partial_products, const uint64_t* betas,
                                     const uint64_t* gammas, const uint64_t* alphas,
                                     const uint64_t public_inputs_hash[4], uint64_t* out_dev, size_t* pos_first,
                                     size_t* pos_count);
int qp_circuit_quotient_finish(qp_circuit* c, uint64_t* vals_leaf_order, uint64_t* out, int out_space);

#ifdef __cplusplus
}
#endif
#endif
