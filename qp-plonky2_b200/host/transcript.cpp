// Host-side transcript mirror (include/qp_plonky2_host.h).  Serial Fiat-Shamir logic only; all
// polynomial / hashing work of the commit path is done by the device entry points it calls.
#include "../../include/qp_plonky2_host.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../csrc/poseidon_constants.h"

namespace {

typedef unsigned __int128 u128;
constexpr uint64_t P = 0xFFFFFFFF00000001ULL;

constexpr uint64_t EPS = 0xFFFFFFFFULL;  // 2^64 mod p = 2^32 - 1

// The transcript is ~120 serial permutations per proof -- a quarter of a 2^12-row proof's latency
// with a textbook `% p` permutation -- so the host permutation is written like a CPU prover's:
// Goldilocks reduction without division or data-dependent branches (field/src/goldilocks_field.rs:
// 390-403), the MDS layer on 32-bit halves (small entries never carry), the partial rounds in
// the reference's fast form (core/src/poseidon.rs:584-596).  tools/microbench/host_poseidon_*.cpp.
inline uint64_t reduce128(u128 x) {   // branch-free: the conditions are data-dependent coin flips
    const uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
    uint64_t t0 = lo - hi_hi;
    t0 -= EPS & (0 - (uint64_t)(lo < hi_hi));
    const uint64_t t1 = hi_lo * EPS;
    uint64_t r = t0 + t1;
    r += EPS & (0 - (uint64_t)(r < t1));
    return r;                          // some representative < 2^64; canonicalised where it matters
}
inline uint64_t canon(uint64_t r) { return r - (P & (0 - (uint64_t)(r >= P))); }
inline uint64_t fmul(uint64_t a, uint64_t b) { return reduce128((u128)a * b); }
inline uint64_t fadd(uint64_t a, uint64_t b) {   // a any u64, b < p  ->  some representative
    uint64_t s = a + b;
    s += EPS & (0 - (uint64_t)(s < a));   // wrapped: + 2^64 = + EPS (cannot wrap again: b < p)
    return s;
}
// lo + hi 2^32 with lo, hi < 2^42 -> mod p  (2^64 = EPS)
inline uint64_t fold(uint64_t lo, uint64_t hi) {
    // value = lo + (hi & EPS) 2^32 + (hi >> 32) 2^64
    const u128 v = (u128)lo + ((u128)(hi & EPS) << 32) + (u128)(hi >> 32) * EPS;
    const uint64_t l = (uint64_t)v, h = (uint64_t)(v >> 64);  // h <= 1
    uint64_t r = l + h * EPS;
    r += EPS & (0 - (uint64_t)(r < l));
    return r;
}
// MDS layer (core/src/poseidon.rs:178-198): circulant with entries < 2^6 plus one diagonal entry.
// 32-bit halves times small constants never carry: twelve independent 64-bit accumulators per half,
// no 128-bit carry chains, one fold per row.
static inline void mds(uint64_t s[12]) {
    uint64_t lo[24], hi[24];
    for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = s[i] & EPS; hi[i] = hi[i + 12] = s[i] >> 32; }
#pragma GCC unroll 12
    for (int r = 0; r < 12; r++) {
        uint64_t al = 0, ah = 0;
#pragma GCC unroll 12
        for (int i = 0; i < 12; i++) { al += lo[r + i] * POSEIDON_MDS_CIRC[i]; ah += hi[r + i] * POSEIDON_MDS_CIRC[i]; }
        if (r == 0) { al += lo[0] * POSEIDON_MDS_DIAG[0]; ah += hi[0] * POSEIDON_MDS_DIAG[0]; }
        s[r] = fold(al, ah);
    }
}
void host_permute(uint64_t s[12]) {
    // poseidon(), core/src/poseidon.rs:599-609: 4 full rounds, 22 partial rounds in the fast form
    // (:584-596 -- one S-box, a dot product and a rank-one update per round), 4 full rounds
    int round = 0;
    for (int half = 0; half < 2; half++) {
        for (int r = 0; r < 4; r++, round++) {
            for (int i = 0; i < 12; i++) s[i] = fadd(s[i], POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i]);
            for (int i = 0; i < 12; i++) {
                const uint64_t x = s[i], x2 = fmul(x, x), x4 = fmul(x2, x2);
                s[i] = fmul(fmul(x, x2), x4);
            }
            mds(s);
        }
        if (half) break;
        for (int i = 0; i < 12; i++) s[i] = fadd(s[i], POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]);  // :304-312
        {   // mds_partial_layer_init, :339-365
            u128 acc[12] = {0};
            for (int r = 1; r < 12; r++) {
                const uint64_t x = canon(s[r]);
                for (int c = 1; c < 12; c++)
                    acc[c] += (u128)reduce128((u128)x * POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (c - 1)]);
            }
            for (int c = 1; c < 12; c++) s[c] = reduce128(acc[c]);
        }
        for (int r = 0; r < 22; r++) {
            const uint64_t x = s[0], x2 = fmul(x, x), x4 = fmul(x2, x2);
            const uint64_t s0 = fadd(fmul(fmul(x, x2), x4), POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS[r]);
            // mds_partial_layer_fast, :378-408
            u128 d = (u128)reduce128((u128)s0 * (POSEIDON_MDS_CIRC[0] + POSEIDON_MDS_DIAG[0]));
            for (int j = 1; j < 12; j++) {
                d += (u128)reduce128((u128)s[j] * POSEIDON_FAST_PARTIAL_ROUND_W_HATS[11 * r + j - 1]);
                s[j] = reduce128((u128)s0 * POSEIDON_FAST_PARTIAL_ROUND_VS[11 * r + j - 1] + s[j]);
            }
            s[0] = reduce128(d);
        }
        round += 22;
    }
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
}
void duplexing(qp_challenger* c) {  // challenger.rs:125-140
    for (uint32_t i = 0; i < c->n_in; i++) c->sponge_state[i] = c->input_buffer[i];
    c->n_in = 0;
    host_permute(c->sponge_state);
    std::memcpy(c->output_buffer, c->sponge_state, 8 * sizeof(uint64_t));
    c->n_out = 8;
}

}  // namespace

extern "C" void qp_challenger_init(qp_challenger* c) { std::memset(c, 0, sizeof *c); }

extern "C" void qp_challenger_observe(qp_challenger* c, const uint64_t* elems, size_t n) {
    for (size_t i = 0; i < n; i++) {  // challenger.rs:35-46
        c->n_out = 0;
        c->input_buffer[c->n_in++] = elems[i] % P;
        if (c->n_in == 8) duplexing(c);
    }
}

extern "C" uint64_t qp_challenger_get(qp_challenger* c) {  // challenger.rs:78-89
    if (c->n_in != 0 || c->n_out == 0) duplexing(c);
    return c->output_buffer[--c->n_out];
}

extern "C" unsigned qp_fri_reduction_arity_bits(unsigned degree_bits, unsigned rate_bits, unsigned cap_height,
                                                unsigned arity_bits, unsigned final_poly_bits, unsigned out[64]) {
    unsigned k = 0;
    if (arity_bits == 0) return 0;  // the reference asserts arity_bits > 0 (core/src/fri.rs:50-61)
    while (degree_bits > final_poly_bits && degree_bits + rate_bits >= cap_height + arity_bits && k < 64) {
        if (arity_bits > degree_bits) break;  // would fold below one coefficient
        out[k++] = arity_bits;
        degree_bits -= arity_bits;
    }
    return k;
}

extern "C" int qp_fri_run_commit_phase(qp_fri* f, unsigned cap_height, const unsigned* arity_bits,
                                       unsigned n_rounds, qp_challenger* ch, uint64_t* caps_out,
                                       uint64_t* final_poly_out, size_t* final_len_out) {
    if (!f || !ch || (n_rounds && (!arity_bits || !caps_out))) return QP_ERR_BAD_ARG;
    const size_t cap_words = ((size_t)1 << cap_height) * 4;
    int rc = QP_OK;
    for (unsigned step = 0; step < n_rounds && !rc; step++) {
        uint64_t* cap = caps_out + step * cap_words;
        rc = qp_fri_commit_round(f, arity_bits[step], cap);
        if (rc) break;
        qp_challenger_observe(ch, cap, cap_words);                              // observe_cap, prover.rs:106
        uint64_t beta[2] = {qp_challenger_get(ch), qp_challenger_get(ch)};      // get_extension_challenge
        rc = qp_fri_fold_round(f, beta, step + 1 == n_rounds);
    }
    if (!rc) rc = qp_fri_final_poly(f, final_poly_out, final_len_out);
    return rc;
}

extern "C" int qp_fri_committed_trees(qp_ctx* ctx, const uint64_t* coeffs_ext, const uint64_t* values_ext,
                                      int space, unsigned lg_n, unsigned rate_bits, unsigned cap_height,
                                      const unsigned* arity_bits, unsigned n_rounds, qp_challenger* ch,
                                      uint64_t* caps_out, uint64_t* final_poly_out, size_t* final_len_out,
                                      qp_fri** fri_out) {
    if (!ch || !fri_out || (n_rounds && (!arity_bits || !caps_out))) return QP_ERR_BAD_ARG;
    qp_fri* f = nullptr;
    int rc = qp_fri_begin(ctx, coeffs_ext, values_ext, space, lg_n, rate_bits, cap_height, &f);
    if (rc) return rc;
    rc = qp_fri_run_commit_phase(f, cap_height, arity_bits, n_rounds, ch, caps_out, final_poly_out, final_len_out);
    if (rc) {
        qp_fri_free(f);
        f = nullptr;
    }
    *fri_out = f;
    return rc;
}

extern "C" int qp_fri_grind(qp_ctx* ctx, qp_challenger* ch, unsigned pow_bits, uint64_t* witness_out) {
    if (!ch || !witness_out) return QP_ERR_BAD_ARG;
    // duplex_intermediate_state (prover.rs:180-183): pending inputs overwritten into the state
    uint64_t st[12];
    std::memcpy(st, ch->sponge_state, sizeof st);
    for (uint32_t i = 0; i < ch->n_in; i++) st[i] = ch->input_buffer[i];
    // min_leading_zeros = pow_bits + (64 - order.bits()) = pow_bits  (prover.rs:167)
    int rc = qp_fri_proof_of_work(ctx, st, ch->n_in, pow_bits, witness_out);
    if (rc) return rc;
    qp_challenger_observe(ch, witness_out, 1);
    (void)qp_challenger_get(ch);
    return QP_OK;
}

// ---------------------------------------------------------------------------------------------
// fri_proof: commit phase + final poly + PoW + query rounds, serialised like write_fri_proof
// ---------------------------------------------------------------------------------------------
namespace {
struct ByteSink {
    uint8_t* out;
    size_t cap, len = 0;
    void u8(uint8_t x) {
        if (out && len < cap) out[len] = x;
        len++;
    }
    void u64s(const uint64_t* p, size_t n) {  // write_field_vec: canonical LE u64 (serialization/mod.rs:1313-1330)
        if (out && len + 8 * n <= cap) std::memcpy(out + len, p, 8 * n);
        len += 8 * n;
    }
    void path(const uint64_t* sib, unsigned layers) {  // write_merkle_proof (mod.rs:1529-1543)
        u8((uint8_t)layers);
        u64s(sib, 4 * (size_t)layers);
    }
};
}  // namespace

extern "C" int qp_fri_proof(qp_ctx* ctx, const qp_batch* const* oracles, size_t n_oracles, qp_fri* f,
                            qp_challenger* ch, unsigned rate_bits, unsigned cap_height, const unsigned* arity_bits,
                            unsigned n_rounds, unsigned pow_bits, unsigned num_queries, uint8_t* out,
                            size_t capacity, size_t* len_out) {
    return qp_fri_proof_sharded(ctx, oracles, n_oracles, 1, f, ch, rate_bits, cap_height, arity_bits, n_rounds, pow_bits,
                                num_queries, out, capacity, len_out);
}

// The same with the initial oracles held as coset shards (multi-GPU commitments): oracles[t * n_shards + s] is
// shard s of oracle t, i.e. leaves [s N / S, (s + 1) N / S) with their part of the tree; a query index is
// opened on the shard that owns it (the shard's cap entries ARE the global ones, so the path is the reference's).
extern "C" int qp_fri_proof_sharded(qp_ctx* ctx, const qp_batch* const* oracles, size_t n_oracles, unsigned n_shards,
                                    qp_fri* f, qp_challenger* ch, unsigned rate_bits, unsigned cap_height,
                                    const unsigned* arity_bits, unsigned n_rounds, unsigned pow_bits,
                                    unsigned num_queries, uint8_t* out, size_t capacity, size_t* len_out) {
    if (!ctx || !f || !ch || !len_out || (n_oracles && !oracles) || (n_rounds && !arity_bits)) return QP_ERR_BAD_ARG;
    if (n_shards == 0 || (n_shards & (n_shards - 1)) || n_shards > (1u << cap_height)) return QP_ERR_BAD_ARG;
    const size_t cap_words = ((size_t)1 << cap_height) * 4;
    // QP_TRACE=1: wall-clock of the stages on stderr (every stage ends synchronised)
    static const bool trace = getenv("QP_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[qp_fri_proof] %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(now - t_prev).count());
        t_prev = now;
    };
    // commit phase (prover.rs:38-49)
    std::vector<uint64_t> caps(cap_words * (n_rounds ? n_rounds : 1));
    size_t final_len = 0;
    int rc = qp_fri_run_commit_phase(f, cap_height, arity_bits, n_rounds, ch, caps.data(), nullptr, &final_len);
    if (rc) return rc;
    std::vector<uint64_t> final_poly(2 * (final_len ? final_len : 1));
    lap("commit phase");
    rc = qp_fri_final_poly(f, final_poly.data(), &final_len);
    if (rc) return rc;
    lap("final poly");
    // observe_final_poly (prover.rs:145-157), then grinding (prover.rs:159-208)
    qp_challenger_observe(ch, final_poly.data(), 2 * final_len);
    uint64_t pow_witness = 0;
    rc = qp_fri_grind(ctx, ch, pow_bits, &pow_witness);
    if (rc) return rc;
    lap("proof of work");
    // query indices (prover.rs:221-226): challenge mod lde size
    unsigned lde_bits = 0;
    {
        // lde size = final_len << rate_bits << sum(arity)
        size_t n = (final_len << rate_bits);
        for (unsigned i = 0; i < n_rounds; i++) n <<= arity_bits[i];
        while (((size_t)1 << lde_bits) < n) lde_bits++;
    }
    const size_t lde = (size_t)1 << lde_bits;
    std::vector<uint64_t> x(num_queries);
    for (unsigned q = 0; q < num_queries; q++) x[q] = qp_challenger_get(ch) % lde;

    // gather everything per tree with one call each
    std::vector<std::vector<uint64_t>> o_rows(n_oracles), o_paths(n_oracles);
    std::vector<size_t> o_len(n_oracles);
    const unsigned o_layers = lde_bits - cap_height;
    const size_t per_shard = lde / n_shards;
    for (size_t t = 0; t < n_oracles && !rc; t++) {
        const qp_batch* const* sh = oracles + t * n_shards;
        if (qp_batch_cap_len(sh[0]) * n_shards != ((size_t)1 << cap_height)) return QP_ERR_BAD_ARG;  // not S shards of the tree
        o_len[t] = qp_batch_leaf_len(sh[0]);
        o_rows[t].resize((size_t)num_queries * o_len[t] + 1);
        o_paths[t].resize((size_t)num_queries * o_layers * 4 + 1);
        if (n_shards == 1) {
            rc = qp_batch_open_many(sh[0], x.data(), num_queries, o_rows[t].data(), o_paths[t].data());
            continue;
        }
        for (unsigned s = 0; s < n_shards && !rc; s++) {
            std::vector<uint64_t> local;
            std::vector<unsigned> who;
            for (unsigned q = 0; q < num_queries; q++)
                if (x[q] / per_shard == s) {
                    local.push_back(x[q] % per_shard);
                    who.push_back(q);
                }
            if (local.empty()) continue;
            std::vector<uint64_t> rows(local.size() * o_len[t] + 1), paths(local.size() * o_layers * 4 + 1);
            rc = qp_batch_open_many(sh[s], local.data(), (unsigned)local.size(), rows.data(), paths.data());
            for (size_t k = 0; k < who.size() && !rc; k++) {
                std::memcpy(o_rows[t].data() + (size_t)who[k] * o_len[t], rows.data() + k * o_len[t], o_len[t] * 8);
                std::memcpy(o_paths[t].data() + (size_t)who[k] * o_layers * 4, paths.data() + k * o_layers * 4,
                            (size_t)o_layers * 32);
            }
        }
    }
    lap("oracle openings");
    std::vector<std::vector<uint64_t>> r_rows(n_rounds), r_paths(n_rounds);
    std::vector<unsigned> r_layers(n_rounds);
    {
        std::vector<uint64_t> idx(x);
        unsigned bits = lde_bits;
        for (unsigned i = 0; i < n_rounds && !rc; i++) {
            for (auto& v : idx) v >>= arity_bits[i];
            bits -= arity_bits[i];
            r_layers[i] = bits - cap_height;
            r_rows[i].resize((size_t)num_queries * (2u << arity_bits[i]) + 1);
            r_paths[i].resize((size_t)num_queries * r_layers[i] * 4 + 1);
            rc = qp_fri_tree_open_many(f, i, idx.data(), num_queries, r_rows[i].data(), r_paths[i].data());
        }
    }
    if (rc) return rc;
    lap("commit-phase tree openings");

    ByteSink w{out, capacity};
    for (unsigned i = 0; i < n_rounds; i++) w.u64s(caps.data() + i * cap_words, cap_words);  // write_merkle_cap
    for (unsigned q = 0; q < num_queries; q++) {
        for (size_t t = 0; t < n_oracles; t++) {  // write_fri_initial_proof
            w.u64s(o_rows[t].data() + (size_t)q * o_len[t], o_len[t]);
            w.path(o_paths[t].data() + (size_t)q * o_layers * 4, o_layers);
        }
        for (unsigned i = 0; i < n_rounds; i++) {  // write_fri_query_step
            const size_t row = 2u << arity_bits[i];
            w.u64s(r_rows[i].data() + (size_t)q * row, row);
            w.path(r_paths[i].data() + (size_t)q * r_layers[i] * 4, r_layers[i]);
        }
    }
    w.u64s(final_poly.data(), 2 * final_len);
    w.u64s(&pow_witness, 1);
    *len_out = w.len;
    return (out && w.len > capacity) ? QP_ERR_BAD_ARG : QP_OK;
}

// ---------------------------------------------------------------------------------------------
// batch FRI (plonky2/src/batch_fri/prover.rs): one FRI proof over polynomials of several degrees
// ---------------------------------------------------------------------------------------------
// batch_fri_committed_trees (prover.rs:83-150): the commit phase of `f` (the largest polynomial); whenever the
// folded codeword reaches the length of the next polynomial of the batch it absorbs that polynomial's values.
extern "C" int qp_batch_fri_run_commit_phase(qp_fri* f, qp_fri* const* lower, size_t n_lower, unsigned cap_height,
                                             const unsigned* arity_bits, unsigned n_rounds, qp_challenger* ch,
                                             uint64_t* caps_out, uint64_t* final_poly_out, size_t* final_len_out) {
    if (!f || !ch || (n_rounds && (!arity_bits || !caps_out)) || (n_lower && !lower)) return QP_ERR_BAD_ARG;
    // "The polynomial vectors should be sorted by degree, from largest to smallest, with no duplicate degrees"
    // and "reduction_arity_bits covers all polynomials" (prover.rs:36-52)
    {
        unsigned cur = qp_fri_domain_bits(f);
        size_t idx = 0;
        for (size_t k = 0; k < n_lower; k++) {
            if (!lower[k]) return QP_ERR_BAD_ARG;
            const unsigned prev = k ? qp_fri_domain_bits(lower[k - 1]) : cur;
            if (qp_fri_domain_bits(lower[k]) >= prev) return QP_ERR_DEGREE_MISMATCH;
        }
        for (unsigned i = 0; i < n_rounds; i++) {
            if (arity_bits[i] > cur) return QP_ERR_BAD_ARG;
            cur -= arity_bits[i];
            if (idx < n_lower && cur == qp_fri_domain_bits(lower[idx])) idx++;
        }
        if (idx != n_lower) return QP_ERR_DEGREE_MISMATCH;
    }
    const size_t cap_words = ((size_t)1 << cap_height) * 4;
    size_t polynomial_index = 0;
    int rc = QP_OK;
    for (unsigned step = 0; step < n_rounds && !rc; step++) {
        uint64_t* cap = caps_out + step * cap_words;
        rc = qp_fri_commit_round(f, arity_bits[step], cap);
        if (rc) break;
        qp_challenger_observe(ch, cap, cap_words);
        uint64_t beta[2] = {qp_challenger_get(ch), qp_challenger_get(ch)};
        const bool last = step + 1 == n_rounds;
        rc = qp_fri_fold_round(f, beta, last);
        if (rc || last) continue;
        if (polynomial_index < n_lower && qp_fri_domain_bits(f) == qp_fri_domain_bits(lower[polynomial_index]))
            rc = qp_fri_mix_values(f, lower[polynomial_index++], beta);
    }
    // (prover.rs:142: every polynomial must have been absorbed; one whose length is only reached by the LAST
    // fold never is -- the reference asserts, so do we)
    if (!rc && polynomial_index != n_lower) rc = QP_ERR_DEGREE_MISMATCH;
    if (!rc) rc = qp_fri_final_poly(f, final_poly_out, final_len_out);
    return rc;
}

// batch_fri_proof (prover.rs:25-80) + write_fri_proof: commit phase, final polynomial, PoW, query rounds with
// the initial openings taken from BatchMerkleTrees (values(x) of every matrix + open_batch(x), prover.rs:189-203).
extern "C" int qp_batch_fri_proof(qp_ctx* ctx, const qp_batch_fri* const* oracles, size_t n_oracles, qp_fri* f,
                                  qp_fri* const* lower, size_t n_lower, qp_challenger* ch, unsigned rate_bits,
                                  unsigned cap_height, const unsigned* arity_bits, unsigned n_rounds,
                                  unsigned pow_bits, unsigned num_queries, uint8_t* out, size_t capacity,
                                  size_t* len_out) {
    if (!ctx || !f || !ch || !len_out || (n_oracles && !oracles) || (n_rounds && !arity_bits)) return QP_ERR_BAD_ARG;
    const size_t cap_words = ((size_t)1 << cap_height) * 4;
    const unsigned lde_bits = qp_fri_domain_bits(f);
    const size_t lde = (size_t)1 << lde_bits;
    std::vector<uint64_t> caps(cap_words * (n_rounds ? n_rounds : 1));
    size_t final_len = 0;
    int rc = qp_batch_fri_run_commit_phase(f, lower, n_lower, cap_height, arity_bits, n_rounds, ch, caps.data(), nullptr,
                                           &final_len);
    if (rc) return rc;
    std::vector<uint64_t> final_poly(2 * (final_len ? final_len : 1));
    rc = qp_fri_final_poly(f, final_poly.data(), &final_len);
    if (rc) return rc;
    qp_challenger_observe(ch, final_poly.data(), 2 * final_len);
    uint64_t pow_witness = 0;
    rc = qp_fri_grind(ctx, ch, pow_bits, &pow_witness);
    if (rc) return rc;
    std::vector<uint64_t> x(num_queries);
    for (unsigned q = 0; q < num_queries; q++) x[q] = qp_challenger_get(ch) % lde;
    // initial openings
    const unsigned o_layers = lde_bits - cap_height;
    std::vector<size_t> o_len(n_oracles, 0);
    std::vector<std::vector<uint64_t>> o_rows(n_oracles), o_paths(n_oracles);
    for (size_t t = 0; t < n_oracles && !rc; t++) {
        if (!oracles[t]) return QP_ERR_BAD_ARG;
        for (size_t g = 0; g < qp_batch_fri_num_groups(oracles[t]); g++) {
            unsigned db = 0;
            size_t np = 0;
            rc = qp_batch_fri_group(oracles[t], g, &db, &np);
            if (rc) return rc;
            if (g == 0 && db + rate_bits != lde_bits) return QP_ERR_DEGREE_MISMATCH;
            o_len[t] += np;
        }
        o_rows[t].resize((size_t)num_queries * o_len[t] + 1);
        o_paths[t].resize((size_t)num_queries * o_layers * 4 + 1);
        for (unsigned q = 0; q < num_queries && !rc; q++) {
            rc = qp_batch_fri_values(oracles[t], x[q], o_rows[t].data() + (size_t)q * o_len[t]);
            if (!rc) rc = qp_batch_fri_open(oracles[t], x[q], o_paths[t].data() + (size_t)q * o_layers * 4);
        }
    }
    if (rc) return rc;
    std::vector<std::vector<uint64_t>> r_rows(n_rounds), r_paths(n_rounds);
    std::vector<unsigned> r_layers(n_rounds);
    {
        std::vector<uint64_t> idx(x);
        unsigned bits = lde_bits;
        for (unsigned i = 0; i < n_rounds && !rc; i++) {
            for (auto& v : idx) v >>= arity_bits[i];
            bits -= arity_bits[i];
            r_layers[i] = bits - cap_height;
            r_rows[i].resize((size_t)num_queries * (2u << arity_bits[i]) + 1);
            r_paths[i].resize((size_t)num_queries * r_layers[i] * 4 + 1);
            rc = qp_fri_tree_open_many(f, i, idx.data(), num_queries, r_rows[i].data(), r_paths[i].data());
        }
    }
    if (rc) return rc;
    ByteSink w{out, capacity};
    for (unsigned i = 0; i < n_rounds; i++) w.u64s(caps.data() + i * cap_words, cap_words);
    for (unsigned q = 0; q < num_queries; q++) {
        for (size_t t = 0; t < n_oracles; t++) {
            w.u64s(o_rows[t].data() + (size_t)q * o_len[t], o_len[t]);
            w.path(o_paths[t].data() + (size_t)q * o_layers * 4, o_layers);
        }
        for (unsigned i = 0; i < n_rounds; i++) {
            const size_t row = 2u << arity_bits[i];
            w.u64s(r_rows[i].data() + (size_t)q * row, row);
            w.path(r_paths[i].data() + (size_t)q * r_layers[i] * 4, r_layers[i]);
        }
    }
    w.u64s(final_poly.data(), 2 * final_len);
    w.u64s(&pow_witness, 1);
    *len_out = w.len;
    return (out && w.len > capacity) ? QP_ERR_BAD_ARG : QP_OK;
}

extern "C" size_t qp_fri_proof_len(const size_t* oracle_leaf_lens, size_t n_oracles, unsigned lde_bits,
                                   unsigned rate_bits, unsigned cap_height, const unsigned* arity_bits,
                                   unsigned n_rounds, unsigned num_queries) {
    size_t total = (size_t)n_rounds * ((size_t)1 << cap_height) * 32;
    size_t per_query = 0;
    for (size_t t = 0; t < n_oracles; t++) per_query += oracle_leaf_lens[t] * 8 + 1 + (size_t)(lde_bits - cap_height) * 32;
    unsigned bits = lde_bits;
    for (unsigned i = 0; i < n_rounds; i++) {
        bits -= arity_bits[i];
        per_query += ((size_t)2 << arity_bits[i]) * 8 + 1 + (size_t)(bits - cap_height) * 32;
    }
    total += per_query * num_queries;
    total += (((size_t)1 << bits) >> rate_bits) * 16 + 8;  // final polynomial + pow witness
    return total;
}
