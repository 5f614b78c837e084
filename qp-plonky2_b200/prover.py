"""prove_with_partition_witness from the full witness on (plonky2/src/plonk/prover.rs:176-398).  The
driver is native (qp-plonky2_b200/host/prover.cpp, `qp_prove` of include/qp_plonky2_host.h); this module
is its ctypes face.  Every polynomial-sized step runs on the device:

    wires commitment            PolynomialBatch.from_values            prover.rs:201-214
    transcript                  Challenger (host, core/src/challenger.rs)   prover.rs:216-234
    Z / partial products        Circuit.partial_products_and_zs        prover.rs:250-261
    lookup polynomials          Circuit.lookup_polys                   prover.rs:262-271,489-636
    their commitment            PolynomialBatch.from_values            prover.rs:275-287
    quotient polynomials        Circuit.compute_quotient_polys         prover.rs:293-307
    quotient commitment         PolynomialBatch.from_coeffs            prover.rs:322-334
    openings at zeta, g zeta    PolynomialBatch.eval_polys             prover.rs:349-361 (OpeningSet::new)
    opening proof               fri_from_openings + fri_proof          prover.rs:364-380 (prove_openings)
    serialization               write_proof_with_public_inputs         util/serialization/mod.rs:2040-2079

Out of scope: circuit building and witness generation -- the caller brings
the constants/sigmas commitment and the witness matrix, like `prove_with_partition_witness` gets
them from ProverOnlyCircuitData and the generators.  Lookup tables: the circuit's (CommonCircuitData(luts=...,
lookup_rows=...)): the deltas are drawn after the gammas and the lookup polynomials are committed with the Z's
(prover.rs:227-271).  Zero-knowledge mode (`salts=`): the three prover oracles are committed with injected salt
columns and leaf_hiding is on.
Deterministic where the reference is not: the PoW witness is the smallest one (serial `find`).
"""
import ctypes as C

import numpy as np

P = 0xFFFFFFFF00000001

class FriConfig:
    """core/src/fri.rs FriConfig with FriReductionStrategy::ConstantArityBits; defaults are
    standard_recursion_config (core/src/circuit_config.rs:51-70)."""

    def __init__(self, rate_bits=3, cap_height=4, proof_of_work_bits=16, arity_bits=4, final_poly_bits=5,
                 num_query_rounds=28):
        self.rate_bits, self.cap_height = rate_bits, cap_height
        self.proof_of_work_bits, self.num_query_rounds = proof_of_work_bits, num_query_rounds
        self.arity_bits, self.final_poly_bits = arity_bits, final_poly_bits

    def observe(self, challenger, degree_bits, reduction_arity_bits):
        """FriConfig::observe + FriParams::observe, core/src/fri.rs:289-321."""
        challenger.observe_elements([self.rate_bits, self.cap_height, self.proof_of_work_bits])
        challenger.observe_elements([1, self.arity_bits, self.final_poly_bits])  # ConstantArityBits.serialize()
        challenger.observe_element(self.num_query_rounds)
        challenger.observe_element(0)  # leaf_hiding
        challenger.observe_element(degree_bits)
        challenger.observe_elements(list(reduction_arity_bits))


def hash_no_pad(elems):
    """hash_no_pad (core/src/hashing.rs:68-95), host side (qp_hash_no_pad)."""
    from . import lib
    a = np.ascontiguousarray(np.asarray([int(e) % P for e in elems], dtype=np.uint64))
    out = np.zeros(4, dtype=np.uint64)
    lib().qp_hash_no_pad(a.ctypes.data if a.size else None, a.size, out.ctypes.data)
    return out


def circuit_digest(constants_sigmas_cap, degree_bits):
    """circuit_builder.rs:1289-1303 with an empty domain separator (qp_circuit_digest)."""
    from . import lib
    cap = np.ascontiguousarray(np.asarray(constants_sigmas_cap, dtype=np.uint64).reshape(-1, 4))
    out = np.zeros(4, dtype=np.uint64)
    lib().qp_circuit_digest(cap.ctypes.data, cap.shape[0], degree_bits, out.ctypes.data)
    return out


class _Config(C.Structure):
    _fields_ = [("rate_bits", C.c_uint32), ("cap_height", C.c_uint32), ("proof_of_work_bits", C.c_uint32),
                ("num_query_rounds", C.c_uint32), ("arity_bits", C.c_uint32), ("final_poly_bits", C.c_uint32),
                ("quotient_degree_factor", C.c_uint32)]


class ProverData:
    """What prove() needs from ProverOnlyCircuitData (circuit_data.rs:497-540): the constants /
    sigmas commitment (circuit_builder.rs:1207-1227) and the circuit digest."""

    def __init__(self, ctx, circuit, constants_sigmas_values, fri_config=None):
        from . import PolynomialBatch, fri_reduction_arity_bits
        c = circuit.common
        self.ctx, self.circuit = ctx, circuit
        self.fri = fri_config or FriConfig(c.rate_bits, c.cap_height)
        self.constants_sigmas_commitment = PolynomialBatch.from_values(
            ctx, constants_sigmas_values, self.fri.rate_bits, False, self.fri.cap_height)
        self.reduction_arity_bits = fri_reduction_arity_bits(
            c.degree_bits, self.fri.rate_bits, self.fri.cap_height, self.fri.arity_bits, self.fri.final_poly_bits)
        self.circuit_digest = circuit_digest(self.constants_sigmas_commitment.merkle_tree.cap, c.degree_bits)


SCOPES = ("compute wires commitment", "compute partial products", "commit to partial products, Z's",
          "compute quotient polys", "commit to quotient polys", "construct the opening set", "compute opening proofs")


def _wire_columns(wires, c):
    """A witness given as a list / tuple of per-wire host vectors (MatrixWitness.wire_values) -> (ctypes pointer array,
    the arrays it points into), or None for a matrix."""
    if not isinstance(wires, (list, tuple)):
        return None
    cols = [np.ascontiguousarray(np.asarray(w, dtype=np.uint64).ravel()) for w in wires]
    if len(cols) != c.num_wires or any(w.size != 1 << c.degree_bits for w in cols):
        raise ValueError("wires must be num_wires vectors of n values")
    return (C.c_void_p * len(cols))(*[w.ctypes.data for w in cols]), cols


def prove(prover_data, wires, public_inputs, timing=None, salts=None):
    """-> bytes of ProofWithPublicInputs.  salts: None, or (wires_salt, zs_salt, quotient_salt), each [4][N]
    (N = n << rate_bits) in the same memory space as `wires`: config.zero_knowledge (prover.rs:210,280,328).
    wires: the witness -- a matrix [num_wires][n] (numpy, or a torch CUDA tensor: then nothing but caps, openings and
    the FRI proof leaves the device), or a LIST of num_wires separately allocated host vectors, the reference's
    MatrixWitness.wire_values (qp_prove_cols: no flattening copy); public_inputs: field elements (already part of
    the witness).  `timing` (dict) receives the reference's TimingTree scopes in milliseconds."""
    from . import QP_HOST, _buf, lib
    pd = prover_data
    ctx, c, f = pd.ctx, pd.circuit.common, pd.fri
    cfg = _Config(f.rate_bits, f.cap_height, f.proof_of_work_bits, f.num_query_rounds, f.arity_bits,
                  f.final_poly_bits, c.quotient_degree_factor)
    as_cols = _wire_columns(wires, c)
    if as_cols is None:
        ptr, space, keep, shape = _buf(wires)
        if tuple(shape) != (c.num_wires, 1 << c.degree_bits):
            raise ValueError("wires must be [num_wires][n]")
    else:
        space = QP_HOST
    pis = np.ascontiguousarray(np.asarray([int(x) % P for x in public_inputs], dtype=np.uint64))
    digest = np.ascontiguousarray(pd.circuit_digest, dtype=np.uint64)
    need = C.c_size_t()
    sp, skeep = [None, None, None], []
    if salts is not None:
        for k, sarr in enumerate(salts):
            q, sspace, keep_s, sshape = _buf(sarr)
            if sspace != space or tuple(sshape) != (4, (1 << c.degree_bits) << f.rate_bits):
                raise ValueError("salts must be [4][N] in the same memory space as the wires")
            sp[k] = q
            skeep.append(keep_s)
    head = (ctx._h, pd.circuit._h, pd.constants_sigmas_commitment._h, digest.ctypes.data, C.byref(cfg))
    tail = (pis.ctypes.data if pis.size else None, pis.size, sp[0], sp[1], sp[2])
    if as_cols is None:
        fn, args = lib().qp_prove_zk, head + (ptr, space) + tail
    else:
        fn, args = lib().qp_prove_cols, head + (as_cols[0],) + tail
    ctx.check(fn(*args, None, 0, C.byref(need), None))
    buf = (C.c_uint8 * need.value)()
    ms = (C.c_double * 7)()
    rc = fn(*args, buf, need.value, C.byref(need), ms)
    if rc:
        from . import QpError
        raise QpError(rc, lib().qp_last_error(ctx._h).decode() or "qp_prove failed")
    if timing is not None:
        timing.update(dict(zip(SCOPES, list(ms))))
    return bytes(buf)


class MultiProverData:
    """ProverData over a MultiContext: one Circuit per device (the gate program, k_is and -- on device 0 -- the
    sigmas), the constants / sigmas commitment as a MultiBatch."""

    def __init__(self, mctx, common, sigmas, constants_sigmas_values, fri_config=None):
        from . import MultiBatch, fri_reduction_arity_bits
        from .plonk import Circuit
        self.mctx, self.common = mctx, common
        self.fri = fri_config or FriConfig(common.rate_bits, common.cap_height)
        self.circuits = [Circuit(ctx, common, sigmas if i == 0 else None) for i, ctx in enumerate(mctx.contexts)]
        self.constants_sigmas_commitment = MultiBatch.from_values_cols(
            mctx, list(constants_sigmas_values), self.fri.rate_bits, False, self.fri.cap_height)
        self.reduction_arity_bits = fri_reduction_arity_bits(
            common.degree_bits, self.fri.rate_bits, self.fri.cap_height, self.fri.arity_bits, self.fri.final_poly_bits)
        self.circuit_digest = circuit_digest(self.constants_sigmas_commitment.cap, common.degree_bits)


def mprove(prover_data, wires, public_inputs, timing=None):
    """prove() over every GPU of the MultiContext (qp_mprove): the four commitments and the quotient evaluation are
    sharded by coset, the rest runs on device 0; the bytes are prove()'s.  wires: host matrix [num_wires][n], or a list
    of num_wires separately allocated host vectors (MatrixWitness.wire_values, qp_mprove_cols)."""
    from . import QpError, lib
    pd = prover_data
    c, f = pd.common, pd.fri
    cfg = _Config(f.rate_bits, f.cap_height, f.proof_of_work_bits, f.num_query_rounds, f.arity_bits,
                  f.final_poly_bits, c.quotient_degree_factor)
    as_cols = _wire_columns(wires, c)
    if as_cols is None:
        w = np.ascontiguousarray(np.asarray(wires, dtype=np.uint64))
        if w.shape != (c.num_wires, 1 << c.degree_bits):
            raise ValueError("wires must be [num_wires][n]")
    pis = np.ascontiguousarray(np.asarray([int(x) % P for x in public_inputs], dtype=np.uint64))
    digest = np.ascontiguousarray(pd.circuit_digest, dtype=np.uint64)
    circs = (C.c_void_p * len(pd.circuits))(*[x._h.value for x in pd.circuits])
    need = C.c_size_t()
    fn = lib().qp_mprove if as_cols is None else lib().qp_mprove_cols
    args = (pd.mctx._h, circs, pd.constants_sigmas_commitment._h, digest.ctypes.data, C.byref(cfg),
            w.ctypes.data if as_cols is None else as_cols[0], pis.ctypes.data if pis.size else None, pis.size)
    rc = fn(*args, None, 0, C.byref(need), None)
    if rc:
        raise QpError(rc, "qp_mprove (sizing) failed")
    buf = (C.c_uint8 * need.value)()
    ms = (C.c_double * 7)()
    rc = fn(*args, buf, need.value, C.byref(need), ms)
    if rc:
        msgs = [lib().qp_last_error(x._h).decode() for x in pd.mctx.contexts] + [lib().qp_mctx_last_error(pd.mctx._h).decode()]
        raise QpError(rc, "qp_mprove failed: " + " | ".join(x for x in msgs if x))
    if timing is not None:
        timing.update(dict(zip(SCOPES, list(ms))))
    return bytes(buf)
