/* qp_plonky2_host -- host-side mirror of the reference's transcript around the commit path.
 *
 * NOT part of the drop-in FFI: in a Rust build these roles are played by the reference's own
 * `Challenger` (core/src/challenger.rs) and `fri_proof` driver (plonky2/src/fri/prover.rs:24-71),
 * which call the device entry points of qp_plonky2_b200.h.  This header exists because the
 * image has no Rust toolchain: it gives the C++/Python harnesses the same orchestration, so
 * the parity tests and bench.py exercise the path exactly the way the shim would.
 *
 * The Fiat-Shamir transcript is inherently serial (one permutation per 8 observed elements)
 * and stays on the host, as SURVEY.md section 8 (a14) specifies; it is never timed as part of
 * a kernel and never substitutes for a device computation.
 */
#ifndef QP_PLONKY2_HOST_H
#define QP_PLONKY2_HOST_H
#include "qp_plonky2_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* core/src/challenger.rs:12-18 */
typedef struct {
    uint64_t sponge_state[12];
    uint64_t input_buffer[8];
    uint64_t output_buffer[8];
    uint32_t n_in, n_out;
} qp_challenger;

void qp_challenger_init(qp_challenger* c);                                   /* Challenger::new */
void qp_challenger_observe(qp_challenger* c, const uint64_t* elems, size_t n);/* observe_elements / observe_cap */
uint64_t qp_challenger_get(qp_challenger* c);                                /* get_challenge */

/* core/src/fri.rs:50-61 (ConstantArityBits).  Returns the number of reductions. */
unsigned qp_fri_reduction_arity_bits(unsigned degree_bits, unsigned rate_bits, unsigned cap_height,
                                     unsigned arity_bits, unsigned final_poly_bits, unsigned out[64]);

/* fri_committed_trees (plonky2/src/fri/prover.rs:85-143) driven over the device step API:
 * per round commit -> observe_cap -> beta = get_extension_challenge -> fold.
 * caps_out: [n_rounds][2^cap_height][4]; final_poly_out: [(n >> sum arity) >> rate_bits][2]. */
int qp_fri_committed_trees(qp_ctx* ctx, const uint64_t* coeffs_ext, const uint64_t* values_ext, int space,
                           unsigned lg_n, unsigned rate_bits, unsigned cap_height,
                           const unsigned* arity_bits, unsigned n_rounds, qp_challenger* challenger,
                           uint64_t* caps_out, uint64_t* final_poly_out, size_t* final_len_out,
                           qp_fri** fri_out);

/* The same loop on an existing FRI state (e.g. one built by qp_fri_begin_from_openings). */
int qp_fri_run_commit_phase(qp_fri* f, unsigned cap_height, const unsigned* arity_bits, unsigned n_rounds,
                            qp_challenger* challenger, uint64_t* caps_out, uint64_t* final_poly_out,
                            size_t* final_len_out);

/* fri_proof_of_work (prover.rs:159-208): grinds on the device, then observes the witness and
 * draws the response on the transcript, as the reference does. */
int qp_fri_grind(qp_ctx* ctx, qp_challenger* challenger, unsigned proof_of_work_bits, uint64_t* witness_out);

/* fri_proof (plonky2/src/fri/prover.rs:24-71) on an existing FRI state: commit phase, final
 * polynomial, proof of work, query rounds -- serialised exactly as Write::write_fri_proof does
 * (plonky2/src/util/serialization/mod.rs:1654-1667): commit-phase caps || per query [per initial
 * tree: leaf values, u8 path length, siblings; per commit round: 2^arity F_p^2 evals, path] ||
 * final polynomial || pow witness, all little-endian canonical u64.
 * initial_oracles are the whole (unsharded) batches whose trees the queries open, in order.
 * The call consumes the FRI state and advances the transcript, so it cannot be repeated: give it
 * a buffer of at least qp_fri_proof_len() bytes.  *len_out receives the bytes written. */
size_t qp_fri_proof_len(const size_t* oracle_leaf_lens, size_t n_oracles, unsigned lde_bits, unsigned rate_bits,
                        unsigned cap_height, const unsigned* arity_bits, unsigned n_rounds,
                        unsigned num_query_rounds);
int qp_fri_proof(qp_ctx* ctx, const qp_batch* const* initial_oracles, size_t n_oracles, qp_fri* f,
                 qp_challenger* challenger, unsigned rate_bits, unsigned cap_height, const unsigned* arity_bits,
                 unsigned n_rounds, unsigned proof_of_work_bits, unsigned num_query_rounds, uint8_t* out,
                 size_t capacity, size_t* len_out);

/* The same with coset-sharded initial oracles (multi-GPU commitments): oracle_shards[t * n_shards + s] is shard s
 * of oracle t; each query index is opened on the shard that owns it.  The bytes are those of qp_fri_proof on the
 * whole batches. */
int qp_fri_proof_sharded(qp_ctx* ctx, const qp_batch* const* oracle_shards, size_t n_oracles, unsigned n_shards,
                         qp_fri* f, qp_challenger* challenger, unsigned rate_bits, unsigned cap_height,
                         const unsigned* arity_bits, unsigned n_rounds, unsigned proof_of_work_bits,
                         unsigned num_query_rounds, uint8_t* out, size_t capacity, size_t* len_out);

/* ---- batch FRI (plonky2/src/batch_fri/prover.rs:25-230): one FRI proof over polynomials of several degrees ----
 * `f` is the FRI state of the largest final polynomial, lower[k] those of the smaller ones in strictly decreasing
 * length (each built by qp_fri_begin on (lde_final_poly, lde_final_values), or by qp_fri_begin_from_openings on the
 * opening batches of that degree with the groups of the BatchFriOracles as polynomial sources,
 * qp_batch_fri_group_batch -- that is BatchFriOracle::prove_openings, batch_fri/oracle.rs:163-229).
 * qp_batch_fri_run_commit_phase = batch_fri_committed_trees; qp_batch_fri_proof = batch_fri_proof followed by
 * write_fri_proof (the byte layout of qp_fri_proof; an oracle's initial opening is values(x) of all its matrices
 * and open_batch(x)).  Errors: QP_ERR_DEGREE_MISMATCH where the reference asserts (lengths not strictly
 * decreasing, reduction_arity_bits not covering every polynomial, oracle height != the FRI domain). */
int qp_batch_fri_run_commit_phase(qp_fri* f, qp_fri* const* lower, size_t n_lower, unsigned cap_height,
                                  const unsigned* arity_bits, unsigned n_rounds, qp_challenger* challenger,
                                  uint64_t* caps_out, uint64_t* final_poly_out, size_t* final_len_out);
int qp_batch_fri_proof(qp_ctx* ctx, const qp_batch_fri* const* initial_oracles, size_t n_oracles, qp_fri* f,
                       qp_fri* const* lower, size_t n_lower, qp_challenger* challenger, unsigned rate_bits,
                       unsigned cap_height, const unsigned* arity_bits, unsigned n_rounds,
                       unsigned proof_of_work_bits, unsigned num_query_rounds, uint8_t* out, size_t capacity,
                       size_t* len_out);

/* ---- circuit data for the quotient (plonky2/src/plonk/circuit_data.rs:412-470) ---------------- */
/* The gates of a circuit, compiled on the host into the constraint program of qp_circuit_desc
 * (qp_plonky2_b200.h): gates are sorted by (degree, id) like CircuitBuilder::build does
 * (circuit_builder.rs:1177-1179), grouped into selector polynomials (gates/selectors.rs:99-166,
 * max_degree = quotient_degree_factor + 1), filtered (gates/gate.rs:326-333) and evaluated
 * symbolically (Gate::eval_unfiltered of gates/{noop,constant,public_input,arithmetic_base,poseidon,
 * arithmetic_extension,multiplication_extension,base_sum,random_access,reducing,reducing_extension,
 * poseidon_mds,exponentiation,coset_interpolation}.rs) -- the gate set of a recursive verifier circuit
 * under standard_recursion_config. */
enum { QP_GATE_NOOP = 0, QP_GATE_CONSTANT = 1, QP_GATE_PUBLIC_INPUT = 2, QP_GATE_ARITHMETIC = 3, QP_GATE_POSEIDON = 4,
       QP_GATE_ARITHMETIC_EXT = 5, QP_GATE_MUL_EXT = 6, QP_GATE_BASE_SUM_2 = 7, QP_GATE_RANDOM_ACCESS = 8,
       QP_GATE_REDUCING = 9, QP_GATE_REDUCING_EXT = 10, QP_GATE_POSEIDON_MDS = 11, QP_GATE_EXPONENTIATION = 12,
       QP_GATE_COSET_INTERPOLATION = 13,
       QP_GATE_LOOKUP = 14, QP_GATE_LOOKUP_TABLE = 15 /* gates/lookup.rs, gates/lookup_table.rs; qp_program_create_lookups */ };
typedef struct {
    uint32_t kind;   /* QP_GATE_* */
    uint32_t param;  /* ConstantGate: num_consts; Arithmetic / ArithmeticExtension / MulExtension: num_ops;
                        BaseSumGate<2>: num_limbs; Reducing / ReducingExtension: num_coeffs;
                        Exponentiation: num_power_bits;
                        RandomAccess: bits | num_copies << 8 | num_extra_constants << 16;
                        CosetInterpolation: subgroup_bits | degree << 8;
                        Lookup / LookupTable: index of the gate's table in `luts`; otherwise 0 */
} qp_gate_desc;
/* Column loads of a compiled program are issued this many loads ahead of their first use (the
 * device keeps that many asynchronous loads in flight; include/qp_plonky2_b200.h, op 12 WAIT). */
#define QP_PROGRAM_LOAD_LEAD 3
typedef struct qp_program qp_program;
int qp_program_create(const qp_gate_desc* gates, size_t n_gates, unsigned max_degree, qp_program** out);
/* The same for a circuit with lookup tables: LookupGate / LookupTableGate have no constraints of their own, but their
 * ids -- "LookupGate {num_slots: .., lut_hash: [..]}", "LookupTableGate {num_slots: .., lut_hash: [..], last_lut_row:
 * ..}" with the Keccak-256 of the table (lookup.rs:44-55,72-77; lookup_table.rs:50-62,86-92) -- take part in the
 * gate ordering, and the gates' constants start after num_selectors + num_lookup_selectors (= 4 + n_luts) columns
 * (gate.rs:179, circuit_builder.rs:1183-1197). */
int qp_program_create_lookups(const qp_gate_desc* gates, size_t n_gates, unsigned max_degree, unsigned num_routed_wires,
                              const qp_lookup_table* luts, size_t n_luts, qp_program** out);
/* keccak_hash::keccak (Keccak-256, the pre-NIST padding) -- the lut_hash of the two lookup gates. */
void qp_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);
/* The same compiler on a recording made elsewhere (a Rust shim's recording field type run over
 * Gate::eval_unfiltered_base_one, or qp-plonky2_b200/plonk.py): nodes in creation order (operands
 * before their consumers), ops numbered like the program's (include/qp_plonky2_b200.h):
 *   LDW / LDK: a = column; LDP: a = index into public_inputs_hash; LDI: a = pool slot;
 *   ADD / SUB / MUL: a, b = nodes; MULI / ADDI: a = node, b = pool slot.
 * actions: every gate's EMITs (node = the constraint's value, k = its index in the gate) followed
 * by one GATE (node = the gate's filter).  The gate accessors below are empty for such a program. */
typedef struct { uint32_t op, a, b; } qp_dag_node;
typedef struct { uint32_t op, node, k; } qp_dag_action;
int qp_program_from_dag(const qp_dag_node* nodes, size_t n_nodes, const uint64_t* pool, size_t pool_len,
                        const qp_dag_action* actions, size_t n_actions, qp_program** out);
void qp_program_free(qp_program* p);
size_t qp_program_code(const qp_program* p, const uint64_t** code);
size_t qp_program_pool(const qp_program* p, const uint64_t** pool);
unsigned qp_program_regs(const qp_program* p);
/* The code is a sequence of self-contained segments separated by OP_END words (a large gate set is
 * cut into pieces of similar cost that different thread blocks evaluate; the quotient is their
 * sum).  -> number of segments; offsets[i] = first word of segment i. */
size_t qp_program_segments(const qp_program* p, const uint32_t** offsets);
unsigned qp_program_num_selectors(const qp_program* p);        /* SelectorsInfo::num_selectors */
unsigned qp_program_num_gate_constants(const qp_program* p);   /* max over gates of num_constants */
unsigned qp_program_num_gate_constraints(const qp_program* p); /* common_data.num_gate_constraints */
/* Gate at position `sorted_index` of common_data.gates: its index in the caller's list, its
 * selector polynomial and that selector's group range (SelectorsInfo). */
int qp_program_gate(const qp_program* p, unsigned sorted_index, unsigned* original_index,
                    unsigned* selector_index, unsigned* group_start, unsigned* group_end);

/* ---- prove (plonky2/src/plonk/prover.rs:176-398) ------------------------------------------------ */
/* hash_no_pad (core/src/hashing.rs:68-95) on the host: public-input hash, circuit digest. */
void qp_hash_no_pad(const uint64_t* elems, size_t n, uint64_t out[4]);
/* circuit_digest (circuit_builder.rs:1289-1303) for an empty domain separator. */
void qp_circuit_digest(const uint64_t* constants_sigmas_cap, size_t cap_len, unsigned degree_bits, uint64_t out[4]);

typedef struct {
    uint32_t rate_bits, cap_height, proof_of_work_bits, num_query_rounds; /* FriConfig, core/src/fri.rs */
    uint32_t arity_bits, final_poly_bits;                                  /* ConstantArityBits(arity_bits, final_poly_bits) */
    uint32_t quotient_degree_factor;
} qp_prover_config;

/* prove_with_partition_witness from the full witness on, serialised like
 * write_proof_with_public_inputs (plonky2/src/util/serialization/mod.rs:2040-2079).  `wires`:
 * witness matrix [num_wires][n] (host or device); the circuit must have been created with sigmas.
 * Lookup tables: the circuit's (qp_circuit_desc.luts; the deltas are drawn and the lookup polynomials committed
 * with the Z's, prover.rs:236-271).  PoW witness = smallest valid one (zero-knowledge blinding: qp_prove_zk below).  Two-call protocol: with out == NULL only
 * *len_out (an upper bound of the proof size) is written.  timing_ms (optional, 7 entries) receives the
 * reference's TimingTree scopes: wires commitment, partial products, their commitment, quotient polys,
 * quotient commitment, opening set, opening proofs. */
int qp_prove(qp_ctx* ctx, qp_circuit* circuit, const qp_batch* constants_sigmas, const uint64_t circuit_digest[4],
             const qp_prover_config* cfg, const uint64_t* wires, int space, const uint64_t* public_inputs,
             size_t n_public_inputs, uint8_t* out, size_t capacity, size_t* len_out, double* timing_ms);

/* The same with config.zero_knowledge (prover.rs:210,280,328; standard_recursion_zk_config,
 * core/src/circuit_config.rs:80-90): the wires, Z / partial-products and quotient oracles carry QP_SALT_SIZE
 * salt columns per leaf, FriParams.leaf_hiding is observed as 1 and the query openings include the salted
 * leaves.  The reference draws the salt from its RNG (fri/oracle.rs:259-263); here it is injected --
 * [QP_SALT_SIZE][N] per oracle, N = n << rate_bits, natural point order, in the same memory space as `wires`
 * -- so that the proof is reproducible.  (The blinding GATES of a zk circuit, circuit_builder.rs:990-1075,
 * are the circuit builder's business: they are rows of the witness the caller brings.) */
int qp_prove_zk(qp_ctx* ctx, qp_circuit* circuit, const qp_batch* constants_sigmas, const uint64_t circuit_digest[4],
                const qp_prover_config* cfg, const uint64_t* wires, int space, const uint64_t* public_inputs,
                size_t n_public_inputs, const uint64_t* wires_salt, const uint64_t* zs_salt,
                const uint64_t* quotient_salt, uint8_t* out, size_t capacity, size_t* len_out, double* timing_ms);

/* The same from the witness as the reference holds it -- MatrixWitness.wire_values, one heap vector per wire
 * (plonky2/src/iop/witness.rs; prover.rs:201-206): wire_cols[w] points at the 2^degree_bits values of wire w in
 * ordinary (pageable) host memory, no flattening copy on the caller's side.  The columns are read once, by the wires
 * commitment (qp_batch_from_values_cols); the salts (all three or none) are host memory. */
int qp_prove_cols(qp_ctx* ctx, qp_circuit* circuit, const qp_batch* constants_sigmas, const uint64_t circuit_digest[4],
                  const qp_prover_config* cfg, const uint64_t* const* wire_cols, const uint64_t* public_inputs,
                  size_t n_public_inputs, const uint64_t* wires_salt, const uint64_t* zs_salt,
                  const uint64_t* quotient_salt, uint8_t* out, size_t capacity, size_t* len_out, double* timing_ms);

/* prove() over every GPU of a multi-device context (qp_mctx, include/qp_plonky2_b200.h): the four commitments and
 * the evaluation of the vanishing polynomial are sharded by coset = by cap subtree, the rest (Z / partial products,
 * the inverse transform of the gathered quotient values, openings, FRI commit phase) runs on devices[0]; query
 * openings come from the shard that owns the leaf.  circuits[i] must have been created on qp_mctx_ctx(m, i)
 * (circuits[0] with sigmas); constants_sigmas is the multi-device commitment of the constants + sigmas columns;
 * `wires` is HOST memory [num_wires][n].  Same two-call protocol, same bytes as qp_prove. */
int qp_mprove(qp_mctx* m, qp_circuit* const* circuits, qp_mbatch* constants_sigmas, const uint64_t circuit_digest[4],
              const qp_prover_config* cfg, const uint64_t* wires, const uint64_t* public_inputs,
              size_t n_public_inputs, uint8_t* out, size_t capacity, size_t* len_out, double* timing_ms);
/* ... from MatrixWitness.wire_values: wire_cols[w] = the host vector of wire w (pageable or pinned). */
int qp_mprove_cols(qp_mctx* m, qp_circuit* const* circuits, qp_mbatch* constants_sigmas, const uint64_t circuit_digest[4],
                   const qp_prover_config* cfg, const uint64_t* const* wire_cols, const uint64_t* public_inputs,
                   size_t n_public_inputs, uint8_t* out, size_t capacity, size_t* len_out, double* timing_ms);

#ifdef __cplusplus
}
#endif
#endif
