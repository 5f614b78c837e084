"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

ctypes binding of the CPU oracle (oracle/plonky2_oracle.c), a plain-C restatement of the
reference's commit path.  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by qp-plonky2_b200/.

Parity status: Poseidon permutation pinned by the reference's KATs; layers above it are
"parity unpinned" by golden data (the reference holds none and cannot be compiled here) and are
pinned structurally and against oracle/pyref.py.  See plonky2_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libplonky2_oracle.so")

P = 0xFFFFFFFF00000001


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = [os.path.join(_HERE, f) for f in ("plonky2_oracle.c", "plonky2_oracle.h", "poseidon_constants.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None

u64p = C.POINTER(C.c_uint64)


def _ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


class Challenger(C.Structure):
    """core/src/challenger.rs state, laid out as orc_challenger."""

    _fields_ = [
        ("sponge_state", C.c_uint64 * 12),
        ("input_buffer", C.c_uint64 * 8),
        ("output_buffer", C.c_uint64 * 8),
        ("n_in", C.c_uint32),
        ("n_out", C.c_uint32),
    ]

    def __init__(self):
        super().__init__()
        lib().orc_challenger_init(C.byref(self))

    def observe(self, elems):
        a = np.ascontiguousarray(np.asarray(elems, dtype=np.uint64).reshape(-1))
        lib().orc_challenger_observe(C.byref(self), _ptr(a), a.size)

    def get_challenge(self) -> int:
        return int(lib().orc_challenger_get(C.byref(self)))

    def get_extension_challenge(self):
        return (self.get_challenge(), self.get_challenge())

    def clone(self):
        c = Challenger.__new__(Challenger)
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Challenger))
        return c


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u64, u32, sz, dbl = C.c_uint64, C.c_uint, C.c_size_t, C.c_double
    sig = {
        "orc_gl_add": (u64, [u64, u64]),
        "orc_gl_sub": (u64, [u64, u64]),
        "orc_gl_mul": (u64, [u64, u64]),
        "orc_gl_canon": (u64, [u64]),
        "orc_gl_pow": (u64, [u64, u64]),
        "orc_gl_inv": (u64, [u64]),
        "orc_gl_inverse_2exp": (u64, [u32]),
        "orc_gl_primitive_root": (u64, [u32]),
        "orc_gl_coset_shift": (u64, []),
        "orc_ext_mul": (None, [u64p, u64p, u64p]),
        "orc_reverse_index_bits": (None, [u64p, sz, sz]),
        "orc_fft": (None, [u64p, u32, u32]),
        "orc_ifft": (None, [u64p, u32]),
        "orc_coset_fft": (None, [u64p, u32, u64, u32]),
        "orc_fft_naive": (None, [u64p, u64p, u32, u64]),
        "orc_poseidon": (None, [u64p]),
        "orc_poseidon_naive": (None, [u64p]),
        "orc_hash_no_pad": (None, [u64p, sz, u64p]),
        "orc_hash_leaf": (None, [u64p, sz, u64p]),
        "orc_two_to_one": (None, [u64p, u64p, u64p]),
        "orc_merkle_tree_new": (C.c_int, [u64p, sz, sz, u32, u64p, u64p]),
        "orc_merkle_prove": (None, [sz, sz, u32, u64p, u64p]),
        "orc_merkle_verify": (C.c_int, [u64p, sz, sz, u64p, u32, u64p, u32]),
        "orc_batch_from_values": (C.c_int, [u64p, sz, u32, u32, u32, u64p, u64p, u64p, u64p, u64p, C.POINTER(dbl)]),
        "orc_batch_from_coeffs": (C.c_int, [u64p, sz, u32, u32, u32, u64p, u64p, u64p, u64p, C.POINTER(dbl)]),
        "orc_challenger_init": (None, [C.POINTER(Challenger)]),
        "orc_challenger_observe": (None, [C.POINTER(Challenger), u64p, sz]),
        "orc_challenger_get": (u64, [C.POINTER(Challenger)]),
        "orc_fri_reduction_arity_bits": (u32, [u32, u32, u32, u32, u32, C.POINTER(u32)]),
        "orc_fri_committed_trees": (
            C.c_int,
            [u64p, u64p, u32, u32, u32, C.POINTER(u32), u32, C.POINTER(Challenger), u64p,
             C.POINTER(u64p), C.POINTER(u64p), u64p, u64p],
        ),
        "orc_fri_proof_of_work": (u64, [C.POINTER(Challenger), u32]),
        "orc_eval_poly_ext": (None, [u64p, sz, u64p, u64p]),
        "orc_reduce_openings": (None, [sz, C.POINTER(sz), C.POINTER(u64p), u64p, u64p, u64p, u32, u64p]),
        "orc_num_threads": (C.c_int, []),
        "orc_eval_vanishing_poly_base": (None, [C.c_void_p, u64, u64, u64p, u64p, u64p, u64p, u64p, u64p, u64p, u64p,
                                                u64p, u64p, u64p]),
        "orc_compute_quotient_polys": (C.c_int, [C.c_void_p, u32, u64p, sz, u64p, sz, u64p, sz, u64p, u64p, u64p,
                                                 u64p, u64p]),
        "orc_partial_products_and_zs": (None, [C.c_void_p, u64p, u64p, u64p, u64p, u64p]),
        "orc_lookup_polys": (None, [C.c_void_p, u64p, u64p, u64p]),
        "orc_lut_poly_eval": (u64, [C.c_void_p, u32, u64p]),
        "orc_eval_vanishing_poly_base_lookup": (None, [C.c_void_p, u64, u64] + [u64p] * 14),
        "orc_compute_quotient_polys_lookup": (C.c_int, [C.c_void_p, u32, u64p, sz, u64p, sz, u64p, sz, u64p, u64p, u64p,
                                                        u64p, u64p, u64p]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


# ---------------------------------------------------------------------------------------------
# numpy-level helpers
# ---------------------------------------------------------------------------------------------

def rand_felts(shape, seed):
    """Uniform canonical Goldilocks elements (rejection of values >= p is a 2^-32 event; we
    reduce instead, which is equally uniform for test purposes)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a).astype(np.uint64)


def poseidon(state):
    s = np.array(state, dtype=np.uint64).copy()
    lib().orc_poseidon(_ptr(s))
    return canon(s)


def poseidon_naive(state):
    s = np.array(state, dtype=np.uint64).copy()
    lib().orc_poseidon_naive(_ptr(s))
    return canon(s)


def canon(a):
    a = np.asarray(a, dtype=np.uint64)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a).astype(np.uint64)


def hash_leaf(x):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_leaf(_ptr(a) if a.size else None, a.size, _ptr(out))
    return canon(out)


def hash_no_pad(x):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_no_pad(_ptr(a) if a.size else None, a.size, _ptr(out))
    return canon(out)


def two_to_one(l, r):
    l = np.ascontiguousarray(np.asarray(l, dtype=np.uint64))
    r = np.ascontiguousarray(np.asarray(r, dtype=np.uint64))
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_two_to_one(_ptr(l), _ptr(r), _ptr(out))
    return canon(out)


def fft(v, zero_factor=0):
    a = np.array(v, dtype=np.uint64).copy()
    lib().orc_fft(_ptr(a), int(a.size).bit_length() - 1, zero_factor)
    return canon(a)


def ifft(v):
    a = np.array(v, dtype=np.uint64).copy()
    lib().orc_ifft(_ptr(a), int(a.size).bit_length() - 1)
    return canon(a)


def coset_fft(v, shift, zero_factor=0):
    a = np.array(v, dtype=np.uint64).copy()
    lib().orc_coset_fft(_ptr(a), int(a.size).bit_length() - 1, shift, zero_factor)
    return canon(a)


def fft_naive(v, shift=1):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.uint64))
    out = np.zeros_like(a)
    lib().orc_fft_naive(_ptr(a), _ptr(out), int(a.size).bit_length() - 1, shift)
    return canon(out)


def reverse_index_bits(v):
    a = np.array(v, dtype=np.uint64).copy()
    n = a.shape[0]
    w = a.size // n
    lib().orc_reverse_index_bits(_ptr(a), n, w)
    return a


class MerkleTree:
    """plonky2/src/hash/merkle_tree.rs: leaves [N][L], digests (reference layout), cap."""

    def __init__(self, leaves, cap_height):
        leaves = np.ascontiguousarray(np.asarray(leaves, dtype=np.uint64))
        n, L = leaves.shape
        if n == 0 or n & (n - 1) or cap_height > n.bit_length() - 1:
            raise ValueError("cap_height should be at most log2(leaves.len())")
        self.leaves = leaves
        self.cap_height = cap_height
        self.digests = np.zeros((2 * (n - (1 << cap_height)), 4), dtype=np.uint64)
        self.cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
        rc = lib().orc_merkle_tree_new(_ptr(leaves), n, L, cap_height, _ptr(self.digests), _ptr(self.cap))
        if rc:
            raise ValueError("merkle_tree_new failed")

    def prove(self, i):
        n = self.leaves.shape[0]
        k = n.bit_length() - 1 - self.cap_height
        sib = np.zeros((k, 4), dtype=np.uint64)
        lib().orc_merkle_prove(i, n, self.cap_height, _ptr(self.digests), _ptr(sib))
        return sib


def merkle_verify(leaf, index, cap, siblings):
    leaf = np.ascontiguousarray(np.asarray(leaf, dtype=np.uint64))
    cap = np.ascontiguousarray(np.asarray(cap, dtype=np.uint64))
    sib = np.ascontiguousarray(np.asarray(siblings, dtype=np.uint64)).reshape(-1, 4)
    ch = int(cap.shape[0]).bit_length() - 1
    return bool(lib().orc_merkle_verify(_ptr(leaf), leaf.size, index, _ptr(cap), ch,
                                        _ptr(sib) if sib.size else None, sib.shape[0]))


class PolynomialBatch:
    """plonky2/src/fri/oracle.rs:33-40 -- polynomials (coeffs), merkle tree, scope timings."""

    SCOPES = ("IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree")

    @classmethod
    def from_values(cls, values, rate_bits, cap_height, salt=None):
        return cls._make(values, rate_bits, cap_height, salt, True)

    @classmethod
    def from_coeffs(cls, coeffs, rate_bits, cap_height, salt=None):
        return cls._make(coeffs, rate_bits, cap_height, salt, False)

    @classmethod
    def _make(cls, data, rate_bits, cap_height, salt, is_values):
        data = np.ascontiguousarray(np.asarray(data, dtype=np.uint64))
        ncols, n = data.shape
        lg_n = n.bit_length() - 1
        assert 1 << lg_n == n
        N = n << rate_bits
        L = ncols + (4 if salt is not None else 0)
        if cap_height > lg_n + rate_bits:
            raise ValueError("cap_height should be at most log2(leaves.len())")
        if salt is not None:
            salt = np.ascontiguousarray(np.asarray(salt, dtype=np.uint64))
            assert salt.shape == (4, N)
        self = cls()
        self.degree_log, self.rate_bits, self.blinding = lg_n, rate_bits, salt is not None
        self.leaves = np.zeros((N, L), dtype=np.uint64)
        self.digests = np.zeros((2 * (N - (1 << cap_height)), 4), dtype=np.uint64)
        self.cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
        ms = (C.c_double * 4)()
        if is_values:
            self.polynomials = np.zeros_like(data)
            rc = lib().orc_batch_from_values(_ptr(data), ncols, lg_n, rate_bits, cap_height, _ptr(salt),
                                             _ptr(self.polynomials), _ptr(self.leaves),
                                             _ptr(self.digests), _ptr(self.cap), ms)
        else:
            self.polynomials = canon(data)
            rc = lib().orc_batch_from_coeffs(_ptr(data), ncols, lg_n, rate_bits, cap_height, _ptr(salt),
                                             _ptr(self.leaves), _ptr(self.digests), _ptr(self.cap), ms)
        if rc:
            raise ValueError("orc_batch failed rc=%d" % rc)
        self.scope_ms = dict(zip(cls.SCOPES, list(ms)))
        return self

    def get_lde_values(self, index, step=1):
        """oracle.rs:286-291"""
        bits = self.degree_log + self.rate_bits
        i = int(format(index * step, "0%db" % bits)[::-1], 2)
        row = self.leaves[i]
        return row[: len(row) - (4 if self.blinding else 0)]


def serialize_polynomial_batch(batch, rate_bits, blinding=False):
    """write_polynomial_batch (plonky2/src/util/serialization/mod.rs:1803-1822; write_merkle_tree
    :1476-1491): every integer a u64 LE, field elements canonical, the bool one byte."""
    u64 = lambda x: np.uint64(x).tobytes()
    co = np.ascontiguousarray(batch.polynomials, dtype="<u8")
    out = [u64(co.shape[0])]
    for col in co:
        out += [u64(col.size), col.tobytes()]
    leaves = np.ascontiguousarray(batch.leaves, dtype="<u8")
    out.append(u64(leaves.shape[0]))
    for row in leaves:
        out += [u64(row.size), row.tobytes()]
    dig = np.ascontiguousarray(batch.digests, dtype="<u8")
    out += [u64(dig.shape[0]), dig.tobytes()]
    cap = np.ascontiguousarray(batch.cap, dtype="<u8")
    out += [u64(cap.shape[0].bit_length() - 1), cap.tobytes()]
    out += [u64(co.shape[1].bit_length() - 1), u64(rate_bits), bytes([1 if blinding else 0])]
    return b"".join(out)


def fri_reduction_arity_bits(degree_bits, rate_bits, cap_height, arity_bits=4, final_poly_bits=5):
    out = (C.c_uint * 64)()
    k = lib().orc_fri_reduction_arity_bits(degree_bits, rate_bits, cap_height, arity_bits, final_poly_bits, out)
    return [int(out[i]) for i in range(k)]


def fri_committed_trees(coeffs, values, rate_bits, cap_height, arity_bits, challenger, keep_trees=False):
    """plonky2/src/fri/prover.rs:85-143.  coeffs/values: [n][2] ext elements.
    Returns dict(caps, betas, final_poly, leaves?, digests?)."""
    co = np.array(coeffs, dtype=np.uint64).copy()
    va = np.array(values, dtype=np.uint64).copy()
    n = co.shape[0]
    lg_n = n.bit_length() - 1
    R = len(arity_bits)
    ncap = 1 << cap_height
    caps = np.zeros((R, ncap, 4), dtype=np.uint64)
    betas = np.zeros((R, 2), dtype=np.uint64)
    tot = sum(arity_bits)
    final = np.zeros(((n >> tot) >> rate_bits, 2), dtype=np.uint64)
    ab = (C.c_uint * max(R, 1))(*arity_bits)
    leaves, digests = [], []
    lp = dp = None
    if keep_trees:
        m = n
        for a in arity_bits:
            leaves.append(np.zeros((m >> a, 2 << a), dtype=np.uint64))
            digests.append(np.zeros((2 * ((m >> a) - ncap), 4), dtype=np.uint64))
            m >>= a
        lp = (u64p * R)(*[_ptr(x) for x in leaves])
        dp = (u64p * R)(*[_ptr(x) if x.size else None for x in digests])
    rc = lib().orc_fri_committed_trees(_ptr(co), _ptr(va), lg_n, rate_bits, cap_height, ab, R,
                                       C.byref(challenger), _ptr(caps), lp, dp, _ptr(betas), _ptr(final))
    if rc:
        raise ValueError("fri_committed_trees rc=%d" % rc)
    return dict(caps=caps, betas=betas, final_poly=final, leaves=leaves, digests=digests)


def fri_proof_of_work(challenger, pow_bits):
    return int(lib().orc_fri_proof_of_work(C.byref(challenger), pow_bits))


def eval_poly_ext(coeffs, point):
    """PolynomialCoeffs::eval at an F_p^2 point -> (c0, c1)"""
    a = np.ascontiguousarray(np.asarray(coeffs, dtype=np.uint64))
    pt = np.array(point, dtype=np.uint64)
    out = np.zeros(2, dtype=np.uint64)
    lib().orc_eval_poly_ext(_ptr(a), a.size, _ptr(pt), _ptr(out))
    return out


def reduce_openings(batches, degree_log):
    """reduce_openings_to_unmasked_final_poly.  batches: list of dict(point=(a,b), shift=(a,b),
    terms=[(poly ndarray[n], (w0, w1)), ...]).  Returns final poly [n][2]."""
    nb = len(batches)
    n_terms = (C.c_size_t * nb)(*[len(b["terms"]) for b in batches])
    polys = [np.ascontiguousarray(np.asarray(p, dtype=np.uint64)) for b in batches for p, _ in b["terms"]]
    ptrs = (u64p * max(len(polys), 1))(*[_ptr(p) for p in polys])
    w = np.array([list(wt) for b in batches for _, wt in b["terms"]], dtype=np.uint64).reshape(-1, 2)
    pts = np.array([list(b["point"]) for b in batches], dtype=np.uint64)
    sh = np.array([list(b["shift"]) for b in batches], dtype=np.uint64)
    out = np.zeros((1 << degree_log, 2), dtype=np.uint64)
    lib().orc_reduce_openings(nb, n_terms, ptrs, _ptr(np.ascontiguousarray(w)) if w.size else None, _ptr(pts), _ptr(sh),
                              degree_log, _ptr(out))
    return out


def fri_proof_bytes(oracle_batches, coeffs, values, challenger, rate_bits, cap_height, arity_bits, pow_bits,
                    num_query_rounds) -> bytes:
    """fri_proof (plonky2/src/fri/prover.rs:24-71) + write_fri_proof
    (plonky2/src/util/serialization/mod.rs:1654-1667) from the oracle's pieces.
    oracle_batches: oracle.PolynomialBatch objects (initial trees); coeffs/values: [n][2]."""
    n = coeffs.shape[0]
    r = fri_committed_trees(coeffs, values, rate_bits, cap_height, arity_bits, challenger, keep_trees=True)
    final = r["final_poly"]
    challenger.observe(final.reshape(-1))                      # observe_final_poly
    w = fri_proof_of_work(challenger, pow_bits)
    xs = [challenger.get_challenge() % n for _ in range(num_query_rounds)]
    out = bytearray()
    out += r["caps"].astype("<u8").tobytes()
    lg = n.bit_length() - 1
    for x in xs:
        for b in oracle_batches:
            out += b.leaves[x].astype("<u8").tobytes()
            sib = np.zeros((lg - cap_height, 4), dtype=np.uint64)
            lib().orc_merkle_prove(x, n, cap_height, _ptr(b.digests), _ptr(sib) if sib.size else None)
            out += bytes([sib.shape[0]]) + sib.astype("<u8").tobytes()
        idx, m = x, n
        for k, a in enumerate(arity_bits):
            idx >>= a
            m >>= a
            out += r["leaves"][k][idx].astype("<u8").tobytes()
            layers = (m.bit_length() - 1) - cap_height
            sib = np.zeros((layers, 4), dtype=np.uint64)
            if layers:
                lib().orc_merkle_prove(idx, m, cap_height, _ptr(r["digests"][k]), _ptr(sib))
            out += bytes([layers]) + sib.astype("<u8").tobytes()
    out += final.astype("<u8").tobytes()
    out += int(w).to_bytes(8, "little")
    return bytes(out)


# ---------------------------------------------------------------------------------------------
# plonk permutation argument and quotient (plonky2/src/plonk/prover.rs:402-480,640-866)
# ---------------------------------------------------------------------------------------------
GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC, GATE_POSEIDON = 0, 1, 2, 3, 4
GATE_ARITHMETIC_EXT, GATE_MUL_EXT, GATE_BASE_SUM_2 = 5, 6, 7
(GATE_RANDOM_ACCESS, GATE_REDUCING, GATE_REDUCING_EXT, GATE_POSEIDON_MDS, GATE_EXPONENTIATION,
 GATE_COSET_INTERPOLATION) = range(8, 14)
GATE_LOOKUP, GATE_LOOKUP_TABLE = 14, 15


class _OrcGate(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("param", C.c_uint32), ("index", C.c_uint32),
                ("selector_index", C.c_uint32), ("group_start", C.c_uint32), ("group_end", C.c_uint32)]


class _OrcLookups(C.Structure):
    _fields_ = [("num_luts", C.c_uint32), ("lut_lens", C.POINTER(C.c_uint32)),
                ("luts", C.POINTER(C.POINTER(C.c_uint16))), ("lookup_rows", C.POINTER(C.c_uint32)),
                ("num_lookup_polys", C.c_uint32), ("lookup_degree", C.c_uint32)]


class _OrcCircuit(C.Structure):
    _fields_ = [("degree_bits", C.c_uint32), ("quotient_degree_bits", C.c_uint32),
                ("num_challenges", C.c_uint32), ("num_routed_wires", C.c_uint32), ("num_wires", C.c_uint32),
                ("num_constants", C.c_uint32), ("num_partial_products", C.c_uint32), ("max_degree", C.c_uint32),
                ("num_selectors", C.c_uint32), ("num_lookup_selectors", C.c_uint32), ("num_gates", C.c_uint32),
                ("gates", C.POINTER(_OrcGate)), ("k_is", u64p), ("lookups", C.POINTER(_OrcLookups))]


class Circuit:
    """The circuit description the oracle's plonk functions take.  `gates`: list of
    (kind, param, selector_index, (group_start, group_end)) in SORTED gate order."""

    def __init__(self, degree_bits, quotient_degree_bits, num_challenges, num_routed_wires, num_wires,
                 num_constants, num_partial_products, max_degree, num_selectors, gates, k_is, luts=None,
                 lookup_rows=None):
        """luts: list of [(input, output), ...] tables (common_data.luts); lookup_rows: one
        (last_lu_row, last_lut_row, first_lut_row) per table (prover_data.lookup_rows)."""
        self._gates = (_OrcGate * len(gates))()
        for i, (kind, param, sel, (gs, ge)) in enumerate(gates):
            self._gates[i] = _OrcGate(kind, param, i, sel, gs, ge)
        self._k_is = np.ascontiguousarray(k_is, dtype=np.uint64)
        self.luts = [np.ascontiguousarray(t, dtype=np.uint16).reshape(-1, 2) for t in (luts or [])]
        self.num_lookup_polys = 0
        lk_ptr = C.POINTER(_OrcLookups)()
        if self.luts:
            lookup_degree = max_degree - 1                      # lookup_accumulator_degree, circuit_data.rs:557-559
            self.num_lookup_polys = -(-(num_routed_wires // 2) // lookup_degree) + 1   # circuit_builder.rs:1284-1290
            self._lut_lens = np.array([len(t) for t in self.luts], dtype=np.uint32)
            self._lut_ptrs = (C.POINTER(C.c_uint16) * len(self.luts))(
                *[t.ctypes.data_as(C.POINTER(C.c_uint16)) for t in self.luts])
            self._rows = np.ascontiguousarray(lookup_rows, dtype=np.uint32).reshape(len(self.luts), 3)
            self._lk = _OrcLookups(len(self.luts), self._lut_lens.ctypes.data_as(C.POINTER(C.c_uint32)), self._lut_ptrs,
                                   self._rows.ctypes.data_as(C.POINTER(C.c_uint32)), self.num_lookup_polys, lookup_degree)
            lk_ptr = C.pointer(self._lk)
        self.c = _OrcCircuit(degree_bits, quotient_degree_bits, num_challenges, num_routed_wires, num_wires,
                             num_constants, num_partial_products, max_degree, num_selectors,
                             4 + len(self.luts) if self.luts else 0, len(gates), self._gates, _ptr(self._k_is), lk_ptr)
        self.num_challenges = num_challenges
        self.degree_bits, self.quotient_degree_bits = degree_bits, quotient_degree_bits
        self.num_partial_products = num_partial_products

    def eval_vanishing_poly_base(self, x, constants, wires, local_zs, next_zs, partial_products, s_sigmas,
                                 betas, gammas, alphas, pih):
        """eval_vanishing_poly_base_batch for one point x (vanishing_poly.rs:166-330)."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in
                (constants, wires, local_zs, next_zs, partial_products, s_sigmas, betas, gammas, alphas, pih)]
        res = np.zeros(self.num_challenges, dtype=np.uint64)
        z_h = (pow(int(x), 1 << self.degree_bits, P) - 1) % P
        lib().orc_eval_vanishing_poly_base(C.byref(self.c), int(x), z_h, *[_ptr(a) for a in arrs], _ptr(res))
        return res

    def compute_quotient_polys(self, rate_bits, cs_leaves, wires_leaves, zs_leaves, betas, gammas, alphas, pih,
                               deltas=None):
        """compute_quotient_polys (prover.rs:640-866) from the three oracles' leaf-major LDE rows.  deltas: the
        lookup challenges [nc][4] of a circuit with lookup tables (the lookup polynomials are then the last
        columns of the third oracle)."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (betas, gammas, alphas, pih)]
        L = [np.ascontiguousarray(a, dtype=np.uint64) for a in (cs_leaves, wires_leaves, zs_leaves)]
        out = np.zeros((self.num_challenges, 1 << (self.degree_bits + self.quotient_degree_bits)), dtype=np.uint64)
        if deltas is None:
            rc = lib().orc_compute_quotient_polys(C.byref(self.c), rate_bits, _ptr(L[0]), L[0].shape[1], _ptr(L[1]),
                                                  L[1].shape[1], _ptr(L[2]), L[2].shape[1],
                                                  *[_ptr(a) for a in arrs], _ptr(out))
        else:
            d = np.ascontiguousarray(deltas, dtype=np.uint64)
            rc = lib().orc_compute_quotient_polys_lookup(
                C.byref(self.c), rate_bits, _ptr(L[0]), L[0].shape[1], _ptr(L[1]), L[1].shape[1], _ptr(L[2]),
                L[2].shape[1], _ptr(arrs[0]), _ptr(arrs[1]), _ptr(d), _ptr(arrs[2]), _ptr(arrs[3]), _ptr(out))
        assert rc == 0, "Having constraints of degree higher than the rate is not supported yet."
        return out

    def lookup_polys(self, wires, deltas):
        """compute_all_lookup_polys (prover.rs:489-636): -> [nc * num_lookup_polys][n] value columns."""
        w = np.ascontiguousarray(wires, dtype=np.uint64)
        d = np.ascontiguousarray(deltas, dtype=np.uint64)
        out = np.zeros((self.num_challenges * self.num_lookup_polys, 1 << self.degree_bits), dtype=np.uint64)
        lib().orc_lookup_polys(C.byref(self.c), _ptr(w), _ptr(d), _ptr(out))
        return out

    def eval_vanishing_poly_base_lookup(self, x, constants, wires, local_zs, next_zs, partial_products, s_sigmas,
                                        local_lookup_zs, next_lookup_zs, betas, gammas, deltas, alphas, pih):
        """eval_vanishing_poly_base_batch for one point x, lookup terms included."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in
                (constants, wires, local_zs, next_zs, partial_products, s_sigmas, local_lookup_zs, next_lookup_zs,
                 betas, gammas, deltas, alphas, pih)]
        res = np.zeros(self.num_challenges, dtype=np.uint64)
        z_h = (pow(int(x), 1 << self.degree_bits, P) - 1) % P
        lib().orc_eval_vanishing_poly_base_lookup(C.byref(self.c), int(x), z_h, *[_ptr(a) for a in arrs], _ptr(res))
        return res

    def partial_products_and_zs(self, wires, sigmas, betas, gammas):
        """all_wires_permutation_partial_products, Z columns first (prover.rs:255-261,402-480)."""
        w = np.ascontiguousarray(wires, dtype=np.uint64)
        s = np.ascontiguousarray(sigmas, dtype=np.uint64)
        b = np.ascontiguousarray(betas, dtype=np.uint64)
        g = np.ascontiguousarray(gammas, dtype=np.uint64)
        nc = self.num_challenges
        out = np.zeros((nc * (1 + self.num_partial_products), 1 << self.degree_bits), dtype=np.uint64)
        lib().orc_partial_products_and_zs(C.byref(self.c), _ptr(w), _ptr(s), _ptr(b), _ptr(g), _ptr(out))
        return out
