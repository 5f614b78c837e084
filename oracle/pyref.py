"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Second, independent restatement of the reference's commit path in pure Python big integers
(every value reduced mod p, no u64 tricks).  It exists because the Rust reference cannot be
compiled in this image: the C oracle (plonky2_oracle.c) is cross-checked against this file on
small cases, so a transcription slip in one restatement shows up as a disagreement.
Deliberately written with DIFFERENT algorithms from the C oracle: recursive radix-2 FFT in
natural order, Poseidon in the naive (KAT-defining) form, level-by-level Merkle construction
addressed through the closed-form digest index.

Citations are reference-relative file:line.  Small inputs only (pure-Python loops).
"""
import os
import re

P = 0xFFFFFFFF00000001
GENERATOR = 14293326489335486720  # field/src/goldilocks_field.rs:84 (coset shift, types.rs:453-455)
POWER_OF_TWO_GENERATOR = 7277203076849721926  # goldilocks_field.rs:91
W = 7  # F_p^2 = F_p[X]/(X^2 - 7), field/src/goldilocks_extensions.rs:13-26

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load_constants():
    text = open(os.path.join(_HERE, "poseidon_constants.h")).read()
    out = {}
    for m in re.finditer(r"static const uint64_t (\w+)\[(\d+)\] = \{(.*?)\};", text, re.S):
        vals = [int(t, 16) for t in re.findall(r"0x([0-9a-fA-F]+)ULL", m.group(3))]
        assert len(vals) == int(m.group(2))
        out[m.group(1)] = vals
    return out


_K = _load_constants()
RC = _K["POSEIDON_ALL_ROUND_CONSTANTS"]
CIRC = _K["POSEIDON_MDS_CIRC"]
DIAG = _K["POSEIDON_MDS_DIAG"]


# ----- field -------------------------------------------------------------------------------
def inv(a):
    return pow(a, P - 2, P)


def primitive_root_of_unity(k):
    """field/src/types.rs:280-284"""
    return pow(POWER_OF_TWO_GENERATOR, 1 << (32 - k), P)


def bitrev(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def reverse_index_bits(a):
    """util/src/lib.rs:49-97: result[i] = arr[bitrev(i)]"""
    bits = (len(a) - 1).bit_length()
    return [a[bitrev(i, bits)] for i in range(len(a))]


# ----- FFT ---------------------------------------------------------------------------------
def _fft_rec(x, w):
    n = len(x)
    if n == 1:
        return list(x)
    e = _fft_rec(x[0::2], w * w % P)
    o = _fft_rec(x[1::2], w * w % P)
    out = [0] * n
    t = 1
    for k in range(n // 2):
        out[k] = (e[k] + t * o[k]) % P
        out[k + n // 2] = (e[k] - t * o[k]) % P
        t = t * w % P
    return out


def fft(coeffs):
    """y_k = sum_i x_i w^(ik), natural order in and out (field/src/fft.rs:159-202)."""
    n = len(coeffs)
    return _fft_rec([c % P for c in coeffs], primitive_root_of_unity(n.bit_length() - 1))


def ifft(values):
    """field/src/fft.rs:68-91: forward transform, reverse indices 1..n-1, scale by 1/n."""
    n = len(values)
    b = fft(values)
    ninv = inv(n)
    return [b[(n - i) % n] * ninv % P for i in range(n)]


def coset_fft(coeffs, shift):
    """field/src/polynomial/mod.rs:280-293"""
    return fft([c * pow(shift, i, P) % P for i, c in enumerate(coeffs)])


def lde_onto_coset(coeffs, rate_bits):
    """polynomial/mod.rs:199-201 + coset_fft_with_options(coset_shift) (oracle.rs:278-280)"""
    n = len(coeffs)
    return coset_fft(list(coeffs) + [0] * (n * ((1 << rate_bits) - 1)), GENERATOR)


# ----- Poseidon (naive form, core/src/poseidon.rs:613-633) ---------------------------------
def _mds(s):
    """poseidon.rs:178-198: row r = sum_i s[(i+r)%12]*CIRC[i] + s[r]*DIAG[r]"""
    return [(sum(s[(i + r) % 12] * CIRC[i] for i in range(12)) + s[r] * DIAG[r]) % P for r in range(12)]


def poseidon(state):
    s = [x % P for x in state]
    for r in range(30):
        s = [(s[i] + RC[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            s = [pow(x, 7, P) for x in s]
        else:
            s[0] = pow(s[0], 7, P)
        s = _mds(s)
    return s


def hash_no_pad(x):
    """core/src/hashing.rs:68-95"""
    s = [0] * 12
    for off in range(0, len(x), 8):
        chunk = x[off : off + 8]
        s[: len(chunk)] = [c % P for c in chunk]
        s = poseidon(s)
    return s[:4]


def hash_leaf(x):
    """core/src/hashing.rs:150-168: capacity lane 8 = len+1, overwrite-mode absorb."""
    s = [0] * 12
    s[8] = len(x) + 1
    for off in range(0, len(x), 8):
        chunk = x[off : off + 8]
        s[: len(chunk)] = [c % P for c in chunk]
        s = poseidon(s)
    return s[:4]


def two_to_one(l, r):
    """core/src/hashing.rs:47-64"""
    return poseidon(list(l) + list(r) + [0] * 4)[:4]


# ----- Merkle tree (plonky2/src/hash/merkle_tree.rs) ---------------------------------------
def digest_index(layer, k):
    """Position, inside a cap-subtree block, of node k of `layer` (0 = leaf digests); derived
    from merkle_tree_prove, merkle_tree.rs:139-158."""
    return 2 * ((k >> 1) << (layer + 1)) + 2 * ((1 << layer) - 1) + (k & 1)


def merkle_tree(leaves, cap_height):
    n = len(leaves)
    lg = n.bit_length() - 1
    assert 1 << lg == n and cap_height <= lg
    ncap = 1 << cap_height
    sub = n >> cap_height
    sub_d = 2 * (sub - 1)
    digests = [None] * (2 * (n - ncap))
    cap = []
    for t in range(ncap):
        level = [hash_leaf(l) for l in leaves[t * sub : (t + 1) * sub]]
        layer = 0
        while len(level) > 1:
            for k, d in enumerate(level):
                digests[t * sub_d + digest_index(layer, k)] = d
            level = [two_to_one(level[2 * k], level[2 * k + 1]) for k in range(len(level) // 2)]
            layer += 1
        cap.append(level[0])
    return digests, cap


def merkle_prove(i, n, cap_height, digests):
    lg = n.bit_length() - 1
    layers = lg - cap_height
    sub_d = 2 * ((n >> cap_height) - 1)
    t = i >> layers
    k = i & ((1 << layers) - 1)
    out = []
    for layer in range(layers):
        out.append(digests[t * sub_d + digest_index(layer, k ^ 1)])
        k >>= 1
    return out


def merkle_verify(leaf, i, cap, siblings):
    """core/src/merkle_proofs.rs:59-97"""
    cur = hash_leaf(leaf)
    for s in siblings:
        cur = two_to_one(s, cur) if i & 1 else two_to_one(cur, s)
        i >>= 1
    return cur == list(cap[i])


# ----- BatchMerkleTree (plonky2/src/hash/batch_merkle_tree.rs) -----------------------------
def batch_merkle_tree(matrices, cap_height):
    """BatchMerkleTree::new (batch_merkle_tree.rs:40-130): matrices of strictly decreasing power-of-two
    heights; the tree over the tallest one is capped at the height of the next, whose rows are then
    hashed together with those cap entries (hash_leaf of digest || row), and so on down to the cap.
    -> (digests, cap, leaf_heights)"""
    assert matrices and all(len(m) & (len(m) - 1) == 0 for m in matrices)
    assert all(len(a) > len(b) for a, b in zip(matrices, matrices[1:]))
    assert cap_height <= len(matrices[-1]).bit_length() - 1
    heights = [len(m) for m in matrices] + [1 << cap_height]
    digests, cap = [], None
    for j, m in enumerate(matrices):
        rows = [list(r) for r in m] if j == 0 else [list(cap[i]) + list(m[i]) for i in range(len(m))]
        d, cap = merkle_tree(rows, heights[j + 1].bit_length() - 1)
        digests += d
    return digests, cap, [h.bit_length() - 1 for h in heights[:-1]]


def batch_merkle_open(i, matrices, cap_height, digests):
    """open_batch (batch_merkle_tree.rs:133-153)"""
    heights = [len(m) for m in matrices] + [1 << cap_height]
    lg0 = heights[0].bit_length() - 1
    out, pos = [], 0
    for j in range(len(matrices)):
        cur, nxt = heights[j], heights[j + 1]
        nd = 2 * (cur - nxt)
        out += merkle_prove(i >> (lg0 - (cur.bit_length() - 1)), cur, nxt.bit_length() - 1, digests[pos:pos + nd])
        pos += nd
    return out


def batch_merkle_verify(leaf_data, leaf_heights, i, cap, siblings):
    """verify_batch_merkle_proof_to_cap (core/src/merkle_proofs.rs:59-97)"""
    cur = hash_leaf(leaf_data[0])
    height, nxt = leaf_heights[0], 1
    for s in siblings:
        cur = two_to_one(s, cur) if i & 1 else two_to_one(cur, s)
        i >>= 1
        height -= 1
        if nxt < len(leaf_heights) and height == leaf_heights[nxt]:
            cur = hash_leaf(list(cur) + list(leaf_data[nxt]))
            nxt += 1
    return nxt == len(leaf_data) and cur == list(cap[i])


# ----- PolynomialBatch (plonky2/src/fri/oracle.rs:168-223) ---------------------------------
def batch_from_values(cols, rate_bits, cap_height, salt=None):
    coeffs = [ifft(c) for c in cols]
    return (coeffs,) + batch_from_coeffs(coeffs, rate_bits, cap_height, salt)


def batch_from_coeffs(coeffs, rate_bits, cap_height, salt=None):
    lde = [lde_onto_coset(c, rate_bits) for c in coeffs]
    if salt is not None:
        lde += [list(s) for s in salt]
    N = len(lde[0])
    bits = N.bit_length() - 1
    leaves = [[col[bitrev(i, bits)] for col in lde] for i in range(N)]
    digests, cap = merkle_tree(leaves, cap_height)
    return leaves, digests, cap


def batch_fri_from_coeffs(polys, rate_bits, cap_height):
    """BatchFriOracle::from_coeffs (plonky2/src/batch_fri/oracle.rs:105-160) without blinding:
    polynomials of non-increasing length; every run of equal length is extended onto the coset,
    transposed and bit-reversed into one leaf matrix; the matrices go to BatchMerkleTree::new.
    -> (leaf matrices, digests, cap, degree_bits tallest first)"""
    lens = [len(p) for p in polys]
    assert all(a >= b for a, b in zip(lens, lens[1:]))
    mats, start = [], 0
    for i in range(len(polys)):
        if i == len(polys) - 1 or lens[i] > lens[i + 1]:
            lde = [lde_onto_coset(c, rate_bits) for c in polys[start:i + 1]]
            bits = len(lde[0]).bit_length() - 1
            mats.append([[col[bitrev(r, bits)] for col in lde] for r in range(len(lde[0]))])
            start = i + 1
    digests, cap, _ = batch_merkle_tree(mats, cap_height)
    return mats, digests, cap, sorted({n.bit_length() - 1 for n in lens}, reverse=True)


# ----- Challenger (core/src/challenger.rs) -------------------------------------------------
class Challenger:
    def __init__(self):
        self.state = [0] * 12
        self.inp = []
        self.out = []

    def _duplex(self):
        self.state[: len(self.inp)] = self.inp
        self.inp = []
        self.state = poseidon(self.state)
        self.out = self.state[:8]

    def observe(self, elems):
        for e in elems:
            self.out = []
            self.inp.append(int(e) % P)
            if len(self.inp) == 8:
                self._duplex()

    def get_challenge(self):
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()

    def get_extension_challenge(self):
        return (self.get_challenge(), self.get_challenge())


# ----- F_p^2 and the FRI commit phase (plonky2/src/fri/prover.rs:85-143) --------------------
def ext_mul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def ext_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fri_committed_trees(coeffs, values, rate_bits, cap_height, arity_bits, challenger):
    coeffs = [tuple(c) for c in coeffs]
    values = [tuple(v) for v in values]
    shift = GENERATOR
    caps, betas, trees = [], [], []
    for step, ab in enumerate(arity_bits):
        arity = 1 << ab
        values = reverse_index_bits(values)
        leaves = [
            [x for e in values[i : i + arity] for x in e] for i in range(0, len(values), arity)
        ]
        digests, cap = merkle_tree(leaves, cap_height)
        trees.append((leaves, digests))
        caps.append(cap)
        challenger.observe([x for d in cap for x in d])
        beta = challenger.get_extension_challenge()
        betas.append(beta)
        folded = []
        for i in range(0, len(coeffs), arity):
            acc = (0, 0)
            for c in reversed(coeffs[i : i + arity]):
                acc = ext_add(ext_mul(acc, beta), c)
            folded.append(acc)
        coeffs = folded
        if step + 1 == len(arity_bits):
            continue
        shift = pow(shift, arity, P)
        lane0 = coset_fft([c[0] for c in coeffs], shift)
        lane1 = coset_fft([c[1] for c in coeffs], shift)
        values = list(zip(lane0, lane1))
    final = coeffs[: len(coeffs) >> rate_bits]
    return caps, betas, final, trees


# ----- opening side (plonky2/src/fri/oracle.rs:129-165) --------------------------------------
def eval_poly_ext(coeffs, point):
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = ext_add(ext_mul(acc, point), (c % P, 0))
    return acc


def reduce_openings(batches, n):
    """batches: list of dict(point, shift, terms=[(poly list, weight)]) -> final poly list of ext."""
    final = [(0, 0)] * n
    for b in batches:
        comp = [(0, 0)] * n
        for poly, w in b["terms"]:
            comp = [ext_add(comp[j], ((w[0] * poly[j]) % P, (w[1] * poly[j]) % P)) for j in range(n)]
        # (comp(X) - comp(z)) / (X - z) by synthetic division, then one zero pad
        z = b["point"]
        q = [(0, 0)] * n
        acc = (0, 0)
        for i in range(n - 1, 0, -1):
            acc = ext_add(ext_mul(acc, z), comp[i])
            q[i - 1] = acc
        final = [ext_add(ext_mul(final[k], b["shift"]), q[k]) for k in range(n)]
    return final


# ----- batch FRI (plonky2/src/batch_fri/prover.rs, batch_fri/oracle.rs:163-229) ---------------
def coset_ifft(values, shift):
    """field/src/polynomial/mod.rs:58-88: ifft, then coefficient i times shift^-i."""
    co = ifft(values)
    s_inv = inv(shift)
    return [c * pow(s_inv, i, P) % P for i, c in enumerate(co)]


def fri_proof_of_work(challenger, pow_bits):
    """plonky2/src/fri/prover.rs:159-208 with the serial smallest-witness rule."""
    st = list(challenger.state)
    st[: len(challenger.inp)] = challenger.inp
    pos = len(challenger.inp)
    w = 0
    while True:
        t = list(st)
        t[pos] = w
        resp = poseidon(t)[7]                       # squeeze().last()
        if resp < (1 << (64 - pow_bits)):           # leading_zeros >= pow_bits
            break
        w += 1
    challenger.observe([w])
    challenger.get_challenge()
    return w


def batch_fri_committed_trees(coeffs, values_list, rate_bits, cap_height, arity_bits, challenger):
    """batch_fri_committed_trees (batch_fri/prover.rs:83-150).  coeffs: the LDE coefficients of the largest
    final polynomial; values_list[k]: the LDE values (natural order) of final polynomial k, lengths strictly
    decreasing.  -> (caps, betas, final_poly, trees)"""
    final_coeffs = [tuple(c) for c in coeffs]
    final_values = [tuple(v) for v in values_list[0]]
    assert all(len(a) > len(b) for a, b in zip(values_list, values_list[1:]))
    shift = GENERATOR
    polynomial_index = 1
    caps, betas, trees = [], [], []
    for step, ab in enumerate(arity_bits):
        arity = 1 << ab
        final_values = reverse_index_bits(final_values)
        leaves = [[x for e in final_values[i : i + arity] for x in e] for i in range(0, len(final_values), arity)]
        digests, cap = merkle_tree(leaves, cap_height)
        trees.append((leaves, digests))
        caps.append(cap)
        challenger.observe([x for d in cap for x in d])
        beta = challenger.get_extension_challenge()
        betas.append(beta)
        folded = []
        for i in range(0, len(final_coeffs), arity):
            acc = (0, 0)
            for c in reversed(final_coeffs[i : i + arity]):
                acc = ext_add(ext_mul(acc, beta), c)
            folded.append(acc)
        final_coeffs = folded
        if step + 1 == len(arity_bits):
            continue
        shift = pow(shift, arity, P)
        final_values = list(zip(coset_fft([c[0] for c in final_coeffs], shift),
                                coset_fft([c[1] for c in final_coeffs], shift)))
        if polynomial_index != len(values_list) and len(final_values) == len(values_list[polynomial_index]):
            final_values = [ext_add(ext_mul(f, beta), tuple(v))
                            for f, v in zip(final_values, values_list[polynomial_index])]
            polynomial_index += 1
        final_coeffs = list(zip(coset_ifft([v[0] for v in final_values], shift),
                                coset_ifft([v[1] for v in final_values], shift)))
    assert polynomial_index == len(values_list)
    final = final_coeffs[: len(final_coeffs) >> rate_bits]
    challenger.observe([x for c in final for x in c])
    return caps, betas, final, trees


def _u64s(xs):
    return b"".join(int(x).to_bytes(8, "little") for x in xs)


def batch_fri_proof_bytes(oracles, coeffs, values_list, challenger, rate_bits, cap_height, arity_bits, pow_bits,
                          num_query_rounds):
    """batch_fri_proof (batch_fri/prover.rs:25-80) + write_fri_proof (serialization/mod.rs:1654-1667).
    oracles: list of (leaf matrices, digests) as batch_fri_from_coeffs returns them."""
    n = len(coeffs)
    caps, betas, final, trees = batch_fri_committed_trees(coeffs, values_list, rate_bits, cap_height, arity_bits,
                                                          challenger)
    w = fri_proof_of_work(challenger, pow_bits)
    xs = [challenger.get_challenge() % n for _ in range(num_query_rounds)]
    out = bytearray()
    for cap in caps:
        out += _u64s(x for d in cap for x in d)
    lg0 = n.bit_length() - 1
    for x in xs:
        for mats, digests in oracles:
            # BatchMerkleTree::values(x) flattened, then open_batch(x) (prover.rs:189-203)
            row = []
            for m in mats:
                row += list(m[x >> (lg0 - (len(m).bit_length() - 1))])
            sib = batch_merkle_open(x, mats, cap_height, digests)
            out += _u64s(row) + bytes([len(sib)]) + _u64s(v for d in sib for v in d)
        idx, m = x, n
        for k, a in enumerate(arity_bits):
            idx >>= a
            m >>= a
            leaves, digests = trees[k]
            sib = merkle_prove(idx, m, cap_height, digests)
            out += _u64s(leaves[idx]) + bytes([len(sib)]) + _u64s(v for d in sib for v in d)
    out += _u64s(x for c in final for x in c)
    out += int(w).to_bytes(8, "little")
    return bytes(out)


def ext_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = ext_mul(r, a)
        a = ext_mul(a, a)
        e >>= 1
    return r


def fri_coefficient(coeff, point):
    """FriCoefficient (core/src/fri_structure.rs:100-108) at the batch point."""
    if coeff == "one" or coeff == ("one",):
        return (1, 0)
    if coeff[0] == "point_power":
        return ext_pow(tuple(point), coeff[1])
    return (coeff[1][0] % P, coeff[1][1] % P)


def batch_prove_openings_bytes(degree_bits, instances, oracle_polys, oracles, challenger, rate_bits, cap_height,
                               arity_bits, pow_bits, num_query_rounds):
    """BatchFriOracle::prove_openings (batch_fri/oracle.rs:163-229).  oracle_polys[o]: the coefficient vectors
    of oracle o (tallest first); oracles[o] = (leaf matrices, digests); instances as in the reference:
    instances[i]["batches"][b] = dict(point, openings=[[(oracle_index, polynomial_index, coefficient), ...], ...])."""
    alpha = challenger.get_extension_challenge()
    final_coeffs, final_values = [], []
    for d, inst in zip(degree_bits, instances):
        n = 1 << d
        batches = []
        for b in inst["batches"]:
            terms, apow = [], (1, 0)
            for expr in b["openings"]:
                for oi, pi, coeff in expr:
                    assert len(oracle_polys[oi][pi]) == n
                    terms.append((oracle_polys[oi][pi], ext_mul(apow, fri_coefficient(coeff, b["point"]))))
                apow = ext_mul(apow, alpha)
            batches.append(dict(point=tuple(b["point"]), shift=apow, terms=terms))
        fin = reduce_openings(batches, n)
        lde = fin + [(0, 0)] * (n * ((1 << rate_bits) - 1))
        final_coeffs.append(lde)
        final_values.append(list(zip(coset_fft([c[0] for c in lde], GENERATOR), coset_fft([c[1] for c in lde], GENERATOR))))
    return batch_fri_proof_bytes(oracles, final_coeffs[0], final_values, challenger, rate_bits, cap_height, arity_bits,
                                 pow_bits, num_query_rounds)
