// prove_with_partition_witness from the full witness on (plonky2/src/plonk/prover.rs:176-398), as a
// host driver over the device entry points of qp_plonky2_b200.h and the host transcript.  Every
// polynomial-sized step runs on the device; this file only sequences them, feeds the Fiat-Shamir
// transcript (serial, host: core/src/challenger.rs) and lays out the proof bytes
// (write_proof_with_public_inputs, plonky2/src/util/serialization/mod.rs:2040-2079).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/qp_plonky2_host.h"

namespace {

typedef unsigned __int128 u128;
constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t W = 7;  // F_p^2 = F_p[X]/(X^2 - 7), field/src/goldilocks_extensions.rs:13-26

struct Ext {
    uint64_t a, b;
};
inline uint64_t fmul(uint64_t x, uint64_t y) { return (uint64_t)(((u128)x * y) % P); }
inline uint64_t fadd(uint64_t x, uint64_t y) { return (uint64_t)(((u128)x + y) % P); }
inline Ext emul(Ext x, Ext y) {  // field/src/extension/quadratic.rs:186-199
    return Ext{fadd(fmul(x.a, y.a), fmul(W, fmul(x.b, y.b))), fadd(fmul(x.a, y.b), fmul(x.b, y.a))};
}
inline uint64_t fpow(uint64_t x, uint64_t e) {
    uint64_t r = 1;
    for (; e; e >>= 1, x = fmul(x, x))
        if (e & 1) r = fmul(r, x);
    return r;
}

void put_u64s(std::vector<uint8_t>& out, const uint64_t* v, size_t n) {  // little-endian canonical u64
    const size_t at = out.size();
    out.resize(at + 8 * n);
    for (size_t i = 0; i < n; i++) {
        uint64_t x = v[i] % P;
        for (int k = 0; k < 8; k++) out[at + 8 * i + k] = (uint8_t)(x >> (8 * k));
    }
}

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double lap(qp_ctx* ctx) {
        qp_ctx_synchronize(ctx);
        auto t1 = std::chrono::steady_clock::now();
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
        return ms;
    }
};

}  // namespace

extern "C" void qp_hash_no_pad(const uint64_t* elems, size_t n, uint64_t out[4]) {
    // hash_no_pad = overwrite-mode sponge, rate 8 (core/src/hashing.rs:68-95): exactly what the
    // challenger's duplex does with its input buffer, so reuse it.
    qp_challenger c;
    qp_challenger_init(&c);
    uint64_t state[12] = {0};
    for (size_t i = 0; i < n; i += 8) {
        const size_t k = n - i < 8 ? n - i : 8;
        std::memcpy(c.sponge_state, state, sizeof state);
        c.n_in = 0;
        qp_challenger_observe(&c, elems + i, k);  // a full chunk permutes here ...
        if (k < 8) (void)qp_challenger_get(&c);    // ... a short one when a challenge is drawn
        std::memcpy(state, c.sponge_state, sizeof state);
        c.n_out = 0;
    }
    for (int i = 0; i < 4; i++) out[i] = state[i] % P;
}

extern "C" void qp_circuit_digest(const uint64_t* cap, size_t cap_len, unsigned degree_bits, uint64_t out[4]) {
    // circuit_builder.rs:1289-1303 with domain_separator = [] ; hash_pad([]) pads to [1,0,0,0,0,0,0,1]
    const uint64_t pad[8] = {1, 0, 0, 0, 0, 0, 0, 1};
    uint64_t ds[4];
    qp_hash_no_pad(pad, 8, ds);
    std::vector<uint64_t> v(cap, cap + 4 * cap_len);
    v.insert(v.end(), ds, ds + 4);
    v.push_back(degree_bits);
    qp_hash_no_pad(v.data(), v.size(), out);
}

extern "C" int qp_prove(qp_ctx* ctx, qp_circuit* circuit, const qp_batch* constants_sigmas,
                        const uint64_t circuit_digest[4], const qp_prover_config* cfg, const uint64_t* wires,
                        int space, const uint64_t* public_inputs, size_t n_public_inputs, uint8_t* out,
                        size_t capacity, size_t* len_out, double* timing_ms) {
    return qp_prove_zk(ctx, circuit, constants_sigmas, circuit_digest, cfg, wires, space, public_inputs, n_public_inputs,
                       nullptr, nullptr, nullptr, out, capacity, len_out, timing_ms);
}

// config.zero_knowledge (plonky2/src/plonk/prover.rs:210,280,328): the wires, Z / partial-products and
// quotient oracles are committed with four salt columns per leaf (fri/oracle.rs:259-263; the reference
// draws them from its RNG, here they are injected so that the proof is reproducible), FriParams.leaf_hiding
// is observed as 1 (core/src/fri.rs:311) and the query openings carry the salted leaves
// (fri_verifier.rs:222 strips them again).  All three salts or none.
static int prove_impl(qp_ctx* ctx, qp_circuit* circuit, const qp_batch* constants_sigmas,
                      const uint64_t circuit_digest[4], const qp_prover_config* cfg, const uint64_t* wires, int space,
                      const uint64_t* const* wire_cols, const uint64_t* public_inputs, size_t n_public_inputs,
                      const uint64_t* wires_salt, const uint64_t* zs_salt, const uint64_t* quotient_salt, uint8_t* out,
                      size_t capacity, size_t* len_out, double* timing_ms);

extern "C" int qp_prove_zk(qp_ctx* ctx, qp_circuit* circuit, const qp_batch* constants_sigmas,
                           const uint64_t circuit_digest[4], const qp_prover_config* cfg, const uint64_t* wires,
                           int space, const uint64_t* public_inputs, size_t n_public_inputs,
                           const uint64_t* wires_salt, const uint64_t* zs_salt, const uint64_t* quotient_salt,
                           uint8_t* out, size_t capacity, size_t* len_out, double* timing_ms) {
    return prove_impl(ctx, circuit, constants_sigmas, circuit_digest, cfg, wires, space, nullptr, public_inputs,
                      n_public_inputs, wires_salt, zs_salt, quotient_salt, out, capacity, len_out, timing_ms);
}

// The witness as the reference holds it: MatrixWitness.wire_values, one heap vector per wire
// (plonky2/src/iop/witness.rs; prover.rs:201-206 wraps each in a PolynomialValues).  The columns are only needed by
// the wires commitment (staged through the pinned ring, qp_batch_from_values_cols); the permutation and lookup
// arguments read the routed wires' values back out of the committed coefficients on the device.
extern "C" int qp_prove_cols(qp_ctx* ctx, qp_circuit* circuit, const qp_batch* constants_sigmas,
                             const uint64_t circuit_digest[4], const qp_prover_config* cfg,
                             const uint64_t* const* wire_cols, const uint64_t* public_inputs, size_t n_public_inputs,
                             const uint64_t* wires_salt, const uint64_t* zs_salt, const uint64_t* quotient_salt,
                             uint8_t* out, size_t capacity, size_t* len_out, double* timing_ms) {
    if (out && !wire_cols) return QP_ERR_BAD_ARG;
    return prove_impl(ctx, circuit, constants_sigmas, circuit_digest, cfg, nullptr, QP_HOST, wire_cols, public_inputs,
                      n_public_inputs, wires_salt, zs_salt, quotient_salt, out, capacity, len_out, timing_ms);
}

static int prove_impl(qp_ctx* ctx, qp_circuit* circuit, const qp_batch* constants_sigmas,
                      const uint64_t circuit_digest[4], const qp_prover_config* cfg, const uint64_t* wires, int space,
                      const uint64_t* const* wire_cols, const uint64_t* public_inputs, size_t n_public_inputs,
                      const uint64_t* wires_salt, const uint64_t* zs_salt, const uint64_t* quotient_salt, uint8_t* out,
                      size_t capacity, size_t* len_out, double* timing_ms) {
    if (!ctx || !circuit || !constants_sigmas || !circuit_digest || !cfg || !len_out) return QP_ERR_BAD_ARG;
    qp_circuit_desc d;
    int rc = qp_circuit_describe(circuit, &d);
    if (rc) return rc;
    const unsigned nc = d.num_challenges, np = d.num_partial_products, qdf = cfg->quotient_degree_factor;
    const size_t n = (size_t)1 << d.degree_bits;
    const size_t n_pre = (size_t)d.num_constants + d.num_routed_wires;
    const size_t n_zs = (size_t)nc * (1 + np), n_q = (size_t)nc * qdf;
    // lookup argument: the RE / partial SLDC polynomials are committed after the Z's and partial products
    // (prover.rs:265-271; lookup_range, circuit_data.rs:582) and opened at zeta and g zeta
    const bool has_lookup = d.num_lookup_polys != 0;
    const size_t n_lk = (size_t)nc * d.num_lookup_polys, n_zs_all = n_zs + n_lk;
    const size_t cap_words = ((size_t)4) << cfg->cap_height;
    unsigned arities[64];
    const unsigned n_rounds = qp_fri_reduction_arity_bits(d.degree_bits, cfg->rate_bits, cfg->cap_height,
                                                          cfg->arity_bits, cfg->final_poly_bits, arities);
    const bool zk = wires_salt || zs_salt || quotient_salt;
    if (zk && !(wires_salt && zs_salt && quotient_salt)) return QP_ERR_BLINDING_NO_SALT;
    const size_t hide = zk ? QP_SALT_SIZE : 0;
    const size_t leaf_lens[4] = {n_pre, d.num_wires + hide, n_zs_all + hide, n_q + hide};
    const size_t fri_len = qp_fri_proof_len(leaf_lens, 4, d.degree_bits + cfg->rate_bits, cfg->rate_bits,
                                            cfg->cap_height, arities, n_rounds, cfg->num_query_rounds);
    const size_t n_open = n_pre + d.num_wires + n_zs_all + nc + n_lk + n_q;
    const size_t total = 8 * (3 * cap_words + 2 * n_open) + fri_len + 8 * (1 + n_public_inputs);
    *len_out = total;
    if (!out) return QP_OK;
    if (capacity < total || (!wires && !wire_cols) || (n_public_inputs && !public_inputs)) return QP_ERR_BAD_ARG;
    if (wire_cols)
        for (unsigned c = 0; c < d.num_wires; c++)
            if (!wire_cols[c]) return QP_ERR_BAD_ARG;
    if (!qp_circuit_has_sigmas(circuit)) return QP_ERR_BAD_ARG;
    if (qdf == 0 || qdf > (1u << d.quotient_degree_bits)) return QP_ERR_BAD_ARG;
    if (qp_batch_leaf_len(constants_sigmas) < n_pre) return QP_ERR_BAD_ARG;

    Timer tm;
    double scopes[7] = {0};
    uint64_t pih[4];
    qp_hash_no_pad(public_inputs, n_public_inputs, pih);  // prover.rs:185-186
    qp_batch *wb = nullptr, *zb = nullptr, *qb = nullptr;
    qp_fri* fri = nullptr;
    uint64_t *d_zs = nullptr, *d_q = nullptr, *d_salt_z = nullptr, *d_salt_q = nullptr, *d_routed = nullptr;
    std::vector<uint8_t> bytes;
    bytes.reserve(total);
    auto cleanup = [&]() {
        if (fri) qp_fri_free(fri);
        qp_batch_free(wb);
        qp_batch_free(zb);
        qp_batch_free(qb);
        qp_dev_free(ctx, d_zs);
        qp_dev_free(ctx, d_q);
        qp_dev_free(ctx, d_salt_z);
        qp_dev_free(ctx, d_salt_q);
        qp_dev_free(ctx, d_routed);
    };
#define QP_STEP(expr)                                                                                   \
    do {                                                                                                \
        rc = (expr);                                                                                    \
        if (rc) {                                                                                       \
            if (getenv("QP_TRACE")) fprintf(stderr, "[prover.cpp:%d] rc=%d: %s\n", __LINE__, rc, #expr); \
            cleanup();                                                                                  \
            return rc;                                                                                  \
        }                                                                                               \
    } while (0)

    // the Z and quotient oracles are committed from device data: their salt has to be there too
    const uint64_t *salt_z = zs_salt, *salt_q = quotient_salt;
    if (zk && space != QP_DEVICE) {
        const size_t words = (size_t)QP_SALT_SIZE << (d.degree_bits + cfg->rate_bits);
        QP_STEP(qp_dev_alloc(ctx, words, &d_salt_z));
        QP_STEP(qp_dev_alloc(ctx, words, &d_salt_q));
        QP_STEP(qp_memcpy(ctx, d_salt_z, QP_DEVICE, zs_salt, QP_HOST, words));
        QP_STEP(qp_memcpy(ctx, d_salt_q, QP_DEVICE, quotient_salt, QP_HOST, words));
        salt_z = d_salt_z;
        salt_q = d_salt_q;
    }
    // wires commitment, prover.rs:201-214
    if (wire_cols)
        QP_STEP(qp_batch_from_values_cols(ctx, wire_cols, d.num_wires, d.degree_bits, cfg->rate_bits, zk ? 1 : 0,
                                          cfg->cap_height, wires_salt, 0, 1u << cfg->rate_bits, &wb));
    else
        QP_STEP(qp_batch_from_values(ctx, wires, space, d.num_wires, d.degree_bits, cfg->rate_bits, zk ? 1 : 0,
                                     cfg->cap_height, wires_salt, 0, 1u << cfg->rate_bits, &wb));
    scopes[0] = tm.lap(ctx);
    // transcript, prover.rs:216-234; FriParams::observe core/src/fri.rs:289-321
    qp_challenger ch;
    qp_challenger_init(&ch);
    {
        std::vector<uint64_t> v = {cfg->rate_bits, cfg->cap_height, cfg->proof_of_work_bits,
                                   1, cfg->arity_bits, cfg->final_poly_bits,  // ConstantArityBits::serialize
                                   cfg->num_query_rounds, zk ? 1u : 0u /* leaf_hiding */, d.degree_bits};
        for (unsigned i = 0; i < n_rounds; i++) v.push_back(arities[i]);
        qp_challenger_observe(&ch, v.data(), v.size());
    }
    qp_challenger_observe(&ch, circuit_digest, 4);
    qp_challenger_observe(&ch, pih, 4);
    std::vector<uint64_t> cap(cap_words);
    QP_STEP(qp_batch_cap(wb, cap.data(), QP_HOST));
    qp_challenger_observe(&ch, cap.data(), cap_words);
    put_u64s(bytes, cap.data(), cap_words);
    std::vector<uint64_t> betas(nc), gammas(nc), alphas(nc);
    for (auto& b : betas) b = qp_challenger_get(&ch);
    for (auto& g : gammas) g = qp_challenger_get(&ch);
    // prover.rs:227-243: four lookup challenges per challenge; betas and gammas are reused for the first 2 nc
    std::vector<uint64_t> deltas;
    if (has_lookup) {
        deltas = betas;
        deltas.insert(deltas.end(), gammas.begin(), gammas.end());
        for (unsigned i = 0; i < 2 * nc; i++) deltas.push_back(qp_challenger_get(&ch));
    }
    // Z and partial products (prover.rs:250-261), kept on the device
    QP_STEP(qp_dev_alloc(ctx, n_zs_all * n, &d_zs));
    // The permutation and lookup arguments read the routed wires' VALUES.  A host witness has just been uploaded and
    // interpolated for the commitment: transforming the routed columns' coefficients back on the device (exact, a
    // few ms at 2^20 rows) replaces a second trip of 80 columns over PCIe (14 ms at 2^20 rows).
    const uint64_t* routed = wires;
    int routed_space = space;
    if (space != QP_DEVICE) {
        QP_STEP(qp_dev_alloc(ctx, (size_t)d.num_routed_wires * n, &d_routed));
        QP_STEP(qp_coset_fft(ctx, qp_batch_device_coeffs(wb), QP_DEVICE, d.num_routed_wires, d.degree_bits, 1, 0, d_routed,
                             QP_DEVICE));
        routed = d_routed;
        routed_space = QP_DEVICE;
    }
    QP_STEP(qp_circuit_partial_products_and_zs(circuit, routed, routed_space, betas.data(), gammas.data(), d_zs, QP_DEVICE));
    if (has_lookup)  // compute_all_lookup_polys, prover.rs:262-263 (also hands the deltas to the quotient evaluation)
        QP_STEP(qp_circuit_lookup_polys(circuit, routed, routed_space, deltas.data(), d_zs + n_zs * n, QP_DEVICE));
    scopes[1] = tm.lap(ctx);
    QP_STEP(qp_batch_from_values(ctx, d_zs, QP_DEVICE, n_zs_all, d.degree_bits, cfg->rate_bits, zk ? 1 : 0, cfg->cap_height,
                                 salt_z, 0, 1u << cfg->rate_bits, &zb));
    scopes[2] = tm.lap(ctx);
    QP_STEP(qp_batch_cap(zb, cap.data(), QP_HOST));
    qp_challenger_observe(&ch, cap.data(), cap_words);
    put_u64s(bytes, cap.data(), cap_words);
    for (auto& a : alphas) a = qp_challenger_get(&ch);
    // quotient polynomials (prover.rs:293-320): coefficients [nc][n << qdb] on the device
    const size_t n_lde = n << d.quotient_degree_bits;
    QP_STEP(qp_dev_alloc(ctx, (size_t)nc * n_lde, &d_q));
    QP_STEP(qp_circuit_compute_quotient_polys(circuit, constants_sigmas, wb, zb, betas.data(), gammas.data(),
                                              alphas.data(), pih, d_q, QP_DEVICE));
    scopes[3] = tm.lap(ctx);
    const uint64_t* d_chunks = d_q;
    uint64_t* d_trim = nullptr;
    if (qdf != (1u << d.quotient_degree_bits)) {
        // trim_to_len(quotient_degree): the tail must be zero ("Quotient has failed, ...")
        std::vector<uint64_t> host((size_t)nc * n_lde);
        QP_STEP(qp_memcpy(ctx, host.data(), QP_HOST, d_q, QP_DEVICE, host.size()));
        std::vector<uint64_t> trimmed((size_t)nc * qdf * n);
        for (unsigned a = 0; a < nc; a++) {
            for (size_t i = (size_t)qdf * n; i < n_lde; i++)
                if (host[a * n_lde + i] % P) {
                    cleanup();
                    return QP_ERR_BAD_ARG;
                }
            std::memcpy(&trimmed[(size_t)a * qdf * n], &host[a * n_lde], (size_t)qdf * n * 8);
        }
        QP_STEP(qp_dev_alloc(ctx, trimmed.size(), &d_trim));
        rc = qp_memcpy(ctx, d_trim, QP_DEVICE, trimmed.data(), QP_HOST, trimmed.size());
        if (rc) {
            qp_dev_free(ctx, d_trim);
            cleanup();
            return rc;
        }
        d_chunks = d_trim;
    }
    rc = qp_batch_from_coeffs(ctx, d_chunks, QP_DEVICE, n_q, d.degree_bits, cfg->rate_bits, zk ? 1 : 0, cfg->cap_height,
                              salt_q, 0, 1u << cfg->rate_bits, &qb);
    qp_dev_free(ctx, d_trim);
    QP_STEP(rc);
    scopes[4] = tm.lap(ctx);
    QP_STEP(qp_batch_cap(qb, cap.data(), QP_HOST));
    qp_challenger_observe(&ch, cap.data(), cap_words);
    put_u64s(bytes, cap.data(), cap_words);
    // zeta, prover.rs:338-347
    const Ext zeta{qp_challenger_get(&ch), qp_challenger_get(&ch)};
    {
        Ext z = zeta;
        for (unsigned i = 0; i < d.degree_bits; i++) z = emul(z, z);
        if (z.a == 1 && z.b == 0) {  // "Opening point is in the subgroup."
            cleanup();
            return QP_ERR_BAD_ARG;
        }
    }
    const uint64_t g = fpow(7277203076849721926ULL, (uint64_t)1 << (32 - d.degree_bits));  // primitive_root_of_unity
    const Ext zeta_next{fmul(g, zeta.a), fmul(g, zeta.b)};
    // OpeningSet::new, proof.rs:289-327
    std::vector<uint64_t> cs_eval(2 * qp_batch_leaf_len(constants_sigmas)), w_eval(2 * (size_t)d.num_wires),
        z_eval(2 * n_zs_all), zn_eval(2 * n_zs_all), q_eval(2 * n_q);
    const uint64_t pz[2] = {zeta.a, zeta.b}, pzn[2] = {zeta_next.a, zeta_next.b};
    QP_STEP(qp_batch_eval_polys(constants_sigmas, pz, cs_eval.data()));
    QP_STEP(qp_batch_eval_polys(wb, pz, w_eval.data()));
    QP_STEP(qp_batch_eval_polys(zb, pz, z_eval.data()));
    QP_STEP(qp_batch_eval_polys(zb, pzn, zn_eval.data()));
    QP_STEP(qp_batch_eval_polys(qb, pz, q_eval.data()));
    scopes[5] = tm.lap(ctx);
    // observe_openings(to_fri_openings()), proof.rs:328-368: zeta batch = constants, sigmas, wires,
    // zs, partial products, quotient polys, lookup_zs; zeta_next batch = zs, lookup_zs
    qp_challenger_observe(&ch, cs_eval.data(), 2 * n_pre);
    qp_challenger_observe(&ch, w_eval.data(), 2 * (size_t)d.num_wires);
    qp_challenger_observe(&ch, z_eval.data(), 2 * (size_t)nc);
    qp_challenger_observe(&ch, z_eval.data() + 2 * nc, 2 * (size_t)nc * np);
    qp_challenger_observe(&ch, q_eval.data(), 2 * n_q);
    qp_challenger_observe(&ch, z_eval.data() + 2 * n_zs, 2 * n_lk);   // lookup_zs close the zeta batch (proof.rs:330-343)
    qp_challenger_observe(&ch, zn_eval.data(), 2 * (size_t)nc);
    qp_challenger_observe(&ch, zn_eval.data() + 2 * n_zs, 2 * n_lk);  // zeta_next batch: zs_next, lookup_zs_next
    // write_opening_set, serialization/mod.rs:1495-1508: constants, sigmas, wires, zs, zs_next,
    // lookup_zs, lookup_zs_next, partial products, quotient polys
    put_u64s(bytes, cs_eval.data(), 2 * n_pre);
    put_u64s(bytes, w_eval.data(), 2 * (size_t)d.num_wires);
    put_u64s(bytes, z_eval.data(), 2 * (size_t)nc);
    put_u64s(bytes, zn_eval.data(), 2 * (size_t)nc);
    put_u64s(bytes, z_eval.data() + 2 * n_zs, 2 * n_lk);   // lookup_zs, lookup_zs_next
    put_u64s(bytes, zn_eval.data() + 2 * n_zs, 2 * n_lk);
    put_u64s(bytes, z_eval.data() + 2 * nc, 2 * (size_t)nc * np);
    put_u64s(bytes, q_eval.data(), 2 * n_q);
    // prove_openings (fri/oracle.rs:320-358) on get_fri_instance(zeta) (circuit_data.rs:592-612)
    const Ext alpha{qp_challenger_get(&ch), qp_challenger_get(&ch)};
    const qp_batch* oracles[4] = {constants_sigmas, wb, zb, qb};
    std::vector<qp_opening_term> t0, t1;
    Ext w{1, 0};
    auto add_terms = [&](std::vector<qp_opening_term>& t, const qp_batch* b, size_t first, size_t count) {
        for (size_t i = first; i < first + count; i++) {  // reduce_polys: sum_i alpha^i p_i, core/src/reducing.rs:63-72
            t.push_back(qp_opening_term{b, i, {w.a, w.b}});
            w = emul(w, alpha);
        }
    };
    // fri_all_openings / fri_next_batch_openings, circuit_data.rs:711-747
    add_terms(t0, oracles[0], 0, n_pre);
    add_terms(t0, oracles[1], 0, d.num_wires);
    add_terms(t0, oracles[2], 0, n_zs);
    add_terms(t0, oracles[3], 0, n_q);
    add_terms(t0, oracles[2], n_zs, n_lk);
    const Ext shift0 = w;  // shift_poly: alpha^count, reducing.rs:94-97
    w = Ext{1, 0};
    add_terms(t1, oracles[2], 0, nc);
    add_terms(t1, oracles[2], n_zs, n_lk);
    const Ext shift1 = w;
    qp_opening_batch ob[2] = {{{zeta.a, zeta.b}, t0.data(), t0.size(), {shift0.a, shift0.b}},
                              {{zeta_next.a, zeta_next.b}, t1.data(), t1.size(), {shift1.a, shift1.b}}};
    const auto t_open = std::chrono::steady_clock::now();
    QP_STEP(qp_fri_begin_from_openings(ctx, ob, 2, d.degree_bits, cfg->rate_bits, cfg->cap_height, &fri));
    if (getenv("QP_TRACE"))
        fprintf(stderr, "[qp_prove] %-30s %8.1f us\n", "fri_begin_from_openings",
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_open).count());
    const size_t at = bytes.size();
    bytes.resize(at + fri_len);
    size_t got = 0;
    rc = qp_fri_proof(ctx, oracles, 4, fri, &ch, cfg->rate_bits, cfg->cap_height, arities, n_rounds,
                      cfg->proof_of_work_bits, cfg->num_query_rounds, bytes.data() + at, fri_len, &got);
    if (!rc && got != fri_len) rc = QP_ERR_BAD_ARG;
    QP_STEP(rc);
    scopes[6] = tm.lap(ctx);
    // public inputs: u64 length, then the elements (serialization/mod.rs:2077-2078)
    const uint64_t npi = n_public_inputs;
    for (int k = 0; k < 8; k++) bytes.push_back((uint8_t)(npi >> (8 * k)));
    put_u64s(bytes, public_inputs, n_public_inputs);
    cleanup();
    if (bytes.size() != total) return QP_ERR_BAD_ARG;
    std::memcpy(out, bytes.data(), total);
    if (timing_ms) std::memcpy(timing_ms, scopes, sizeof scopes);
    return QP_OK;
#undef QP_STEP
}


// ---------------------------------------------------------------------------------------------------------
// prove() over every GPU of a multi-device context (BASELINE.json configs[4]: a large circuit whose LDEs do not
// fit one GPU).  What is sharded by coset = by cap subtree: the four commitments -- LDE +
// Merkle hashing, >= 90 % of the work of a large proof -- and the evaluation of the vanishing polynomial (a point
// needs only its own coset, prover.rs:679,750).  What runs on device 0: the Z / partial-product columns (a
// prefix product over the rows), the inverse transform of the quotient values (gathered there as peer copies:
// num_challenges * 8 n words), the openings and the FRI commit phase of the final polynomial (every shard keeps
// the full coefficient matrices, so device 0 has all it needs; the FRI oracle is 1/16 of one LDE column pair
// per round).  Query openings are served by the shard that owns the leaf.  The proof is byte-identical to
// qp_prove's.
// ---------------------------------------------------------------------------------------------------------
extern "C" int qp_mprove(qp_mctx* m, qp_circuit* const* circuits, qp_mbatch* constants_sigmas,
                         const uint64_t circuit_digest[4], const qp_prover_config* cfg, const uint64_t* wires,
                         const uint64_t* public_inputs, size_t n_public_inputs, uint8_t* out, size_t capacity,
                         size_t* len_out, double* timing_ms) {
    if (!m || !circuits || !circuits[0] || !cfg) return QP_ERR_BAD_ARG;
    qp_circuit_desc d;
    int rc = qp_circuit_describe(circuits[0], &d);
    if (rc) return rc;
    std::vector<const uint64_t*> cols(d.num_wires, nullptr);
    for (unsigned c = 0; wires && c < d.num_wires; c++) cols[c] = wires + ((size_t)c << d.degree_bits);
    return qp_mprove_cols(m, circuits, constants_sigmas, circuit_digest, cfg, wires ? cols.data() : nullptr, public_inputs,
                          n_public_inputs, out, capacity, len_out, timing_ms);
}

// The same from the reference's MatrixWitness.wire_values (one host vector per wire, pageable or pinned).
extern "C" int qp_mprove_cols(qp_mctx* m, qp_circuit* const* circuits, qp_mbatch* constants_sigmas,
                              const uint64_t circuit_digest[4], const qp_prover_config* cfg,
                              const uint64_t* const* wire_cols, const uint64_t* public_inputs, size_t n_public_inputs,
                              uint8_t* out, size_t capacity, size_t* len_out, double* timing_ms) {
    if (!m || !circuits || !constants_sigmas || !circuit_digest || !cfg || !len_out) return QP_ERR_BAD_ARG;
    const unsigned D = qp_mctx_num_devices(m);
    if (D == 0 || qp_mbatch_num_shards(constants_sigmas) != D) return QP_ERR_BAD_ARG;
    for (unsigned e = 0; e < D; e++)
        if (!circuits[e]) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = qp_mctx_ctx(m, 0);
    qp_circuit_desc d;
    int rc = qp_circuit_describe(circuits[0], &d);
    if (rc) return rc;
    const unsigned nc = d.num_challenges, np = d.num_partial_products, qdf = cfg->quotient_degree_factor;
    const size_t n = (size_t)1 << d.degree_bits;
    const size_t n_pre = (size_t)d.num_constants + d.num_routed_wires;
    const size_t n_zs = (size_t)nc * (1 + np), n_q = (size_t)nc * qdf;
    // lookup argument: the RE / partial SLDC polynomials are committed after the Z's and partial products
    // (prover.rs:265-271; lookup_range, circuit_data.rs:582) and opened at zeta and g zeta
    const bool has_lookup = d.num_lookup_polys != 0;
    const size_t n_lk = (size_t)nc * d.num_lookup_polys, n_zs_all = n_zs + n_lk;
    const size_t cap_words = ((size_t)4) << cfg->cap_height;
    unsigned arities[64];
    const unsigned n_rounds = qp_fri_reduction_arity_bits(d.degree_bits, cfg->rate_bits, cfg->cap_height,
                                                          cfg->arity_bits, cfg->final_poly_bits, arities);
    const size_t leaf_lens[4] = {n_pre, d.num_wires, n_zs_all, n_q};
    const size_t fri_len = qp_fri_proof_len(leaf_lens, 4, d.degree_bits + cfg->rate_bits, cfg->rate_bits,
                                            cfg->cap_height, arities, n_rounds, cfg->num_query_rounds);
    const size_t n_open = n_pre + d.num_wires + n_zs_all + nc + n_lk + n_q;
    const size_t total = 8 * (3 * cap_words + 2 * n_open) + fri_len + 8 * (1 + n_public_inputs);
    *len_out = total;
    if (!out) return QP_OK;
    if (capacity < total || !wire_cols || (n_public_inputs && !public_inputs)) return QP_ERR_BAD_ARG;
    for (unsigned c = 0; c < d.num_wires; c++)
        if (!wire_cols[c]) return QP_ERR_BAD_ARG;
    if (!qp_circuit_has_sigmas(circuits[0])) return QP_ERR_BAD_ARG;
    if (qdf == 0 || qdf > (1u << d.quotient_degree_bits)) return QP_ERR_BAD_ARG;
    if (qp_batch_leaf_len(qp_mbatch_shard(constants_sigmas, 0)) < n_pre) return QP_ERR_BAD_ARG;

    Timer tm;
    double scopes[7] = {0};
    uint64_t pih[4];
    qp_hash_no_pad(public_inputs, n_public_inputs, pih);
    qp_mbatch *wb = nullptr, *zb = nullptr, *qb = nullptr;
    qp_fri* fri = nullptr;
    uint64_t *d_zs = nullptr, *d_q = nullptr, *d_vals = nullptr, *d_routed = nullptr;
    std::vector<uint8_t> bytes;
    bytes.reserve(total);
    auto cleanup = [&]() {
        if (fri) qp_fri_free(fri);
        qp_mbatch_free(wb);
        qp_mbatch_free(zb);
        qp_mbatch_free(qb);
        qp_dev_free(ctx, d_zs);
        qp_dev_free(ctx, d_q);
        qp_dev_free(ctx, d_vals);
        qp_dev_free(ctx, d_routed);
    };
#define QP_STEP(expr)                                                                                   \
    do {                                                                                                \
        rc = (expr);                                                                                    \
        if (rc) {                                                                                       \
            if (getenv("QP_TRACE")) fprintf(stderr, "[prover.cpp:%d] rc=%d: %s\n", __LINE__, rc, #expr); \
            cleanup();                                                                                  \
            return rc;                                                                                  \
        }                                                                                               \
    } while (0)

    // wires commitment over all devices, prover.rs:201-214
    QP_STEP(qp_mbatch_from_values_cols(m, wire_cols, d.num_wires, d.degree_bits, cfg->rate_bits, 0, cfg->cap_height, nullptr,
                                       &wb));
    scopes[0] = tm.lap(ctx);
    qp_challenger ch;
    qp_challenger_init(&ch);
    {
        std::vector<uint64_t> v = {cfg->rate_bits, cfg->cap_height, cfg->proof_of_work_bits,
                                   1, cfg->arity_bits, cfg->final_poly_bits,
                                   cfg->num_query_rounds, 0 /* leaf_hiding */, d.degree_bits};
        for (unsigned i = 0; i < n_rounds; i++) v.push_back(arities[i]);
        qp_challenger_observe(&ch, v.data(), v.size());
    }
    qp_challenger_observe(&ch, circuit_digest, 4);
    qp_challenger_observe(&ch, pih, 4);
    std::vector<uint64_t> cap(cap_words);
    QP_STEP(qp_mbatch_cap(wb, cap.data()));
    qp_challenger_observe(&ch, cap.data(), cap_words);
    put_u64s(bytes, cap.data(), cap_words);
    std::vector<uint64_t> betas(nc), gammas(nc), alphas(nc);
    for (auto& b : betas) b = qp_challenger_get(&ch);
    for (auto& g : gammas) g = qp_challenger_get(&ch);
    // prover.rs:227-243: four lookup challenges per challenge; betas and gammas are reused for the first 2 nc
    std::vector<uint64_t> deltas;
    if (has_lookup) {
        deltas = betas;
        deltas.insert(deltas.end(), gammas.begin(), gammas.end());
        for (unsigned i = 0; i < 2 * nc; i++) deltas.push_back(qp_challenger_get(&ch));
    }
    // Z and partial products on device 0 (prover.rs:250-261)
    QP_STEP(qp_dev_alloc(ctx, n_zs_all * n, &d_zs));
    // the routed wires' values from device 0's copy of the coefficient matrix (see qp_prove_zk): no second upload
    QP_STEP(qp_dev_alloc(ctx, (size_t)d.num_routed_wires * n, &d_routed));
    QP_STEP(qp_coset_fft(ctx, qp_batch_device_coeffs(qp_mbatch_shard(wb, 0)), QP_DEVICE, d.num_routed_wires, d.degree_bits, 1, 0,
                         d_routed, QP_DEVICE));
    QP_STEP(qp_circuit_partial_products_and_zs(circuits[0], d_routed, QP_DEVICE, betas.data(), gammas.data(), d_zs, QP_DEVICE));
    if (has_lookup) {  // on device 0 like the Z's; every shard's quotient evaluation needs the challenges
        QP_STEP(qp_circuit_lookup_polys(circuits[0], d_routed, QP_DEVICE, deltas.data(), d_zs + n_zs * n, QP_DEVICE));
        for (unsigned e = 1; e < D; e++) QP_STEP(qp_circuit_set_lookup_challenges(circuits[e], deltas.data()));
    }
    QP_STEP(qp_ctx_synchronize(ctx));
    scopes[1] = tm.lap(ctx);
    QP_STEP(qp_mbatch_from_device(m, d_zs, 0, n_zs_all, d.degree_bits, cfg->rate_bits, 0, cfg->cap_height, nullptr, &zb));
    scopes[2] = tm.lap(ctx);
    QP_STEP(qp_mbatch_cap(zb, cap.data()));
    qp_challenger_observe(&ch, cap.data(), cap_words);
    put_u64s(bytes, cap.data(), cap_words);
    for (auto& a : alphas) a = qp_challenger_get(&ch);
    // quotient values, coset shard by coset shard, gathered on device 0 in leaf order (prover.rs:293-320)
    const size_t n_lde = n << d.quotient_degree_bits;
    QP_STEP(qp_dev_alloc(ctx, (size_t)nc * n_lde, &d_vals));
    QP_STEP(qp_dev_alloc(ctx, (size_t)nc * n_lde, &d_q));
    QP_STEP(qp_ctx_synchronize(ctx));   // the peers write into d_vals
    {
        std::vector<int> rcs(D, QP_OK);
        std::vector<std::thread> th;
        for (unsigned e = 0; e < D; e++)
            th.emplace_back([&, e] {
                qp_ctx* ce = qp_mctx_ctx(m, e);
                const qp_batch* wsh = qp_mbatch_shard(wb, e);
                const size_t n_loc = (size_t)qp_batch_cap_len(wsh) ? ((n << cfg->rate_bits) / D) : 0;
                uint64_t* d_part = nullptr;
                int r = qp_dev_alloc(ce, (size_t)nc * n_loc, &d_part);
                size_t first = 0, count = 0;
                if (!r)
                    r = qp_circuit_quotient_values_shard(circuits[e], qp_mbatch_shard(constants_sigmas, e), wsh,
                                                         qp_mbatch_shard(zb, e), betas.data(), gammas.data(), alphas.data(),
                                                         pih, d_part, &first, &count);
                for (unsigned a = 0; a < nc && !r && count; a++)
                    r = qp_memcpy_peer(ctx, d_vals + (size_t)a * n_lde + first, ce, d_part + (size_t)a * count, count);
                qp_dev_free(ce, d_part);
                rcs[e] = r;
            });
        for (auto& t : th) t.join();
        for (unsigned e = 0; e < D; e++)
            if (rcs[e]) {
                cleanup();
                return rcs[e];
            }
    }
    QP_STEP(qp_circuit_quotient_finish(circuits[0], d_vals, d_q, QP_DEVICE));
    scopes[3] = tm.lap(ctx);
    const uint64_t* d_chunks = d_q;
    uint64_t* d_trim = nullptr;
    if (qdf != (1u << d.quotient_degree_bits)) {
        std::vector<uint64_t> host((size_t)nc * n_lde);
        QP_STEP(qp_memcpy(ctx, host.data(), QP_HOST, d_q, QP_DEVICE, host.size()));
        std::vector<uint64_t> trimmed((size_t)nc * qdf * n);
        for (unsigned a = 0; a < nc; a++) {
            for (size_t i = (size_t)qdf * n; i < n_lde; i++)
                if (host[a * n_lde + i] % P) {
                    cleanup();
                    return QP_ERR_BAD_ARG;
                }
            std::memcpy(&trimmed[(size_t)a * qdf * n], &host[a * n_lde], (size_t)qdf * n * 8);
        }
        QP_STEP(qp_dev_alloc(ctx, trimmed.size(), &d_trim));
        rc = qp_memcpy(ctx, d_trim, QP_DEVICE, trimmed.data(), QP_HOST, trimmed.size());
        if (rc) {
            qp_dev_free(ctx, d_trim);
            cleanup();
            return rc;
        }
        d_chunks = d_trim;
    }
    QP_STEP(qp_ctx_synchronize(ctx));
    rc = qp_mbatch_from_device(m, d_chunks, 1, n_q, d.degree_bits, cfg->rate_bits, 0, cfg->cap_height, nullptr, &qb);
    qp_dev_free(ctx, d_trim);
    QP_STEP(rc);
    scopes[4] = tm.lap(ctx);
    QP_STEP(qp_mbatch_cap(qb, cap.data()));
    qp_challenger_observe(&ch, cap.data(), cap_words);
    put_u64s(bytes, cap.data(), cap_words);
    const Ext zeta{qp_challenger_get(&ch), qp_challenger_get(&ch)};
    {
        Ext z = zeta;
        for (unsigned i = 0; i < d.degree_bits; i++) z = emul(z, z);
        if (z.a == 1 && z.b == 0) {
            cleanup();
            return QP_ERR_BAD_ARG;
        }
    }
    const uint64_t g = fpow(7277203076849721926ULL, (uint64_t)1 << (32 - d.degree_bits));
    const Ext zeta_next{fmul(g, zeta.a), fmul(g, zeta.b)};
    // every shard holds the full coefficient matrices: device 0 evaluates (proof.rs:289-327)
    const qp_batch *cs0 = qp_mbatch_shard(constants_sigmas, 0), *w0 = qp_mbatch_shard(wb, 0), *z0 = qp_mbatch_shard(zb, 0),
                   *q0 = qp_mbatch_shard(qb, 0);
    std::vector<uint64_t> cs_eval(2 * qp_batch_leaf_len(cs0)), w_eval(2 * (size_t)d.num_wires), z_eval(2 * n_zs_all),
        zn_eval(2 * n_zs_all), q_eval(2 * n_q);
    const uint64_t pz[2] = {zeta.a, zeta.b}, pzn[2] = {zeta_next.a, zeta_next.b};
    QP_STEP(qp_batch_eval_polys(cs0, pz, cs_eval.data()));
    QP_STEP(qp_batch_eval_polys(w0, pz, w_eval.data()));
    QP_STEP(qp_batch_eval_polys(z0, pz, z_eval.data()));
    QP_STEP(qp_batch_eval_polys(z0, pzn, zn_eval.data()));
    QP_STEP(qp_batch_eval_polys(q0, pz, q_eval.data()));
    scopes[5] = tm.lap(ctx);
    qp_challenger_observe(&ch, cs_eval.data(), 2 * n_pre);
    qp_challenger_observe(&ch, w_eval.data(), 2 * (size_t)d.num_wires);
    qp_challenger_observe(&ch, z_eval.data(), 2 * (size_t)nc);
    qp_challenger_observe(&ch, z_eval.data() + 2 * nc, 2 * (size_t)nc * np);
    qp_challenger_observe(&ch, q_eval.data(), 2 * n_q);
    qp_challenger_observe(&ch, z_eval.data() + 2 * n_zs, 2 * n_lk);   // lookup_zs close the zeta batch (proof.rs:330-343)
    qp_challenger_observe(&ch, zn_eval.data(), 2 * (size_t)nc);
    qp_challenger_observe(&ch, zn_eval.data() + 2 * n_zs, 2 * n_lk);  // zeta_next batch: zs_next, lookup_zs_next
    put_u64s(bytes, cs_eval.data(), 2 * n_pre);
    put_u64s(bytes, w_eval.data(), 2 * (size_t)d.num_wires);
    put_u64s(bytes, z_eval.data(), 2 * (size_t)nc);
    put_u64s(bytes, zn_eval.data(), 2 * (size_t)nc);
    put_u64s(bytes, z_eval.data() + 2 * n_zs, 2 * n_lk);   // lookup_zs, lookup_zs_next
    put_u64s(bytes, zn_eval.data() + 2 * n_zs, 2 * n_lk);
    put_u64s(bytes, z_eval.data() + 2 * nc, 2 * (size_t)nc * np);
    put_u64s(bytes, q_eval.data(), 2 * n_q);
    const Ext alpha{qp_challenger_get(&ch), qp_challenger_get(&ch)};
    const qp_batch* oracles0[4] = {cs0, w0, z0, q0};
    std::vector<qp_opening_term> t0, t1;
    Ext w{1, 0};
    auto add_terms = [&](std::vector<qp_opening_term>& t, const qp_batch* b, size_t first, size_t count) {
        for (size_t i = first; i < first + count; i++) {
            t.push_back(qp_opening_term{b, i, {w.a, w.b}});
            w = emul(w, alpha);
        }
    };
    add_terms(t0, oracles0[0], 0, n_pre);
    add_terms(t0, oracles0[1], 0, d.num_wires);
    add_terms(t0, oracles0[2], 0, n_zs);
    add_terms(t0, oracles0[3], 0, n_q);
    add_terms(t0, oracles0[2], n_zs, n_lk);
    const Ext shift0 = w;
    w = Ext{1, 0};
    add_terms(t1, oracles0[2], 0, nc);
    add_terms(t1, oracles0[2], n_zs, n_lk);
    const Ext shift1 = w;
    qp_opening_batch ob[2] = {{{zeta.a, zeta.b}, t0.data(), t0.size(), {shift0.a, shift0.b}},
                              {{zeta_next.a, zeta_next.b}, t1.data(), t1.size(), {shift1.a, shift1.b}}};
    QP_STEP(qp_fri_begin_from_openings(ctx, ob, 2, d.degree_bits, cfg->rate_bits, cfg->cap_height, &fri));
    std::vector<const qp_batch*> shards(4 * (size_t)D);
    qp_mbatch* mbs[4] = {constants_sigmas, wb, zb, qb};
    for (unsigned t = 0; t < 4; t++)
        for (unsigned e = 0; e < D; e++) shards[(size_t)t * D + e] = qp_mbatch_shard(mbs[t], e);
    const size_t at = bytes.size();
    bytes.resize(at + fri_len);
    size_t got = 0;
    rc = qp_fri_proof_sharded(ctx, shards.data(), 4, D, fri, &ch, cfg->rate_bits, cfg->cap_height, arities, n_rounds,
                              cfg->proof_of_work_bits, cfg->num_query_rounds, bytes.data() + at, fri_len, &got);
    if (!rc && got != fri_len) rc = QP_ERR_BAD_ARG;
    QP_STEP(rc);
    scopes[6] = tm.lap(ctx);
    const uint64_t npi = n_public_inputs;
    for (int k = 0; k < 8; k++) bytes.push_back((uint8_t)(npi >> (8 * k)));
    put_u64s(bytes, public_inputs, n_public_inputs);
    cleanup();
    if (bytes.size() != total) return QP_ERR_BAD_ARG;
    std::memcpy(out, bytes.data(), total);
    if (timing_ms) std::memcpy(timing_ms, scopes, sizeof scopes);
    return QP_OK;
#undef QP_STEP
}
