"""qp_plonky2_b200 -- B200-native polynomial-commitment path of Quantus-Network/qp-plonky2.

Python host mirror of the reference interface for the hot path (names, argument meaning and
error behaviour follow the Rust API; citations are reference-relative file:line):

    PolynomialBatch.from_values / from_coeffs      plonky2/src/fri/oracle.rs:168-223
    PolynomialBatch.get_lde_values                 plonky2/src/fri/oracle.rs:286-291
    MerkleTree(leaves, cap_height) / .prove / .get plonky2/src/hash/merkle_tree.rs:163-207
    Challenger                                     core/src/challenger.rs
    fri_committed_trees / fri_proof_of_work        plonky2/src/fri/prover.rs:85-208

Everything goes through the C ABI in include/qp_plonky2_b200.h (ctypes, plain pointers): the
same entry points a Rust shim would bind.  There is no CPU fallback: if the
CUDA library is not built, or no GPU is present, the calls raise.

Host data are numpy uint64 arrays; device data may be passed as torch CUDA tensors (int64 or
uint64 storage), in which case no host<->device copy of the input happens.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# QP_PLONKY2_LIB: load a differently-built library (kernel experiments, tools/exp_variants.py)
LIB_PATH = os.environ.get("QP_PLONKY2_LIB") or os.path.join(_HERE, "libqp_plonky2_b200.so")
_SOURCES = [
    os.path.join(_HERE, "csrc", f)
    for f in ("qp_plonky2.cu", "goldilocks.cuh", "poseidon.cuh", "poseidon_constants.h", "ntt.cuh",
              "merkle.cuh", "fri.cuh", "openings.cuh", "quotient.cuh", "multi_device.inl")
] + [
    os.path.join(_HERE, "host", "transcript.cpp"),
    os.path.join(_HERE, "host", "plonk_host.cpp"),
    os.path.join(_HERE, "host", "prover.cpp"),
    os.path.join(_ROOT, "include", "qp_plonky2_b200.h"),
    os.path.join(_ROOT, "include", "qp_plonky2_host.h"),
]

P = 0xFFFFFFFF00000001
SALT_SIZE = 4
QP_HOST, QP_DEVICE = 0, 1

ERRORS = {
    1: "CUDA failure",
    2: "cap_height should be at most log2(leaves.len())",
    3: "Not a power of two",
    4: "Polynomial degrees inconsistent",
    5: "bad argument",
    6: "too large",
    7: "Cannot set blinding without salt",
    8: "not supported by this library (refused, not approximated)",
}


class QpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("qp_plonky2_b200 error %d (%s): %s" % (code, ERRORS.get(code, "?"), msg))
        self.code = code


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in _SOURCES
    )
    if not (force or stale):
        return LIB_PATH
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH,
        os.path.join(_HERE, "csrc", "qp_plonky2.cu"), os.path.join(_HERE, "host", "transcript.cpp"),
        os.path.join(_HERE, "host", "plonk_host.cpp"), os.path.join(_HERE, "host", "prover.cpp"),
    ]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


u64p = C.POINTER(C.c_uint64)


class _ChallengerState(C.Structure):
    _fields_ = [
        ("sponge_state", C.c_uint64 * 12),
        ("input_buffer", C.c_uint64 * 8),
        ("output_buffer", C.c_uint64 * 8),
        ("n_in", C.c_uint32),
        ("n_out", C.c_uint32),
    ]


_lib = None


def lib():
    """Load the C-ABI library.  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "qp_plonky2_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback" % LIB_PATH
        )
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, sz, i32 = C.c_void_p, C.c_uint64, C.c_uint, C.c_size_t, C.c_int
    pp = C.POINTER(vp)
    sig = {
        "qp_ctx_create": (i32, [i32, vp, u32, pp]),
        "qp_ctx_destroy": (None, [vp]),
        "qp_last_error": (C.c_char_p, [vp]),
        "qp_ctx_synchronize": (i32, [vp]),
        "qp_ctx_launch_count": (u64, [vp]),
        "qp_batch_from_values": (i32, [vp, vp, i32, sz, u32, u32, i32, u32, vp, u32, u32, pp]),
        "qp_batch_from_coeffs": (i32, [vp, vp, i32, sz, u32, u32, i32, u32, vp, u32, u32, pp]),
        "qp_batch_from_values_cols": (i32, [vp, vp, sz, u32, u32, i32, u32, vp, u32, u32, pp]),
        "qp_merkle_tree_new_shard": (i32, [vp, vp, i32, sz, sz, u32, u32, u32, pp]),
        "qp_mctx_create": (i32, [vp, u32, u32, pp]),
        "qp_mctx_destroy": (None, [vp]),
        "qp_mctx_num_devices": (u32, [vp]),
        "qp_mctx_ctx": (vp, [vp, u32]),
        "qp_mctx_last_error": (C.c_char_p, [vp]),
        "qp_mbatch_from_values_cols": (i32, [vp, vp, sz, u32, u32, i32, u32, vp, pp]),
        "qp_mbatch_cap": (i32, [vp, vp]),
        "qp_mbatch_num_shards": (u32, [vp]),
        "qp_mbatch_shard": (vp, [vp, u32]),
        "qp_mbatch_free": (None, [vp]),
        "qp_batch_free": (None, [vp]),
        "qp_batch_begin": (i32, [vp, sz, u32, u32, i32, u32, u32, u32, pp]),
        "qp_batch_put_coeffs": (i32, [vp, vp, i32, sz, sz]),
        "qp_batch_coeffs_slot": (vp, [vp, sz]),
        "qp_batch_extend_columns": (i32, [vp, sz, sz, i32]),
        "qp_batch_end": (i32, [vp, vp, i32]),
        "qp_ifft_columns": (i32, [vp, vp, i32, sz, u32, vp, i32]),
        "qp_batch_cap": (i32, [vp, vp, i32]),
        "qp_batch_cap_len": (sz, [vp]),
        "qp_batch_coeffs": (i32, [vp, vp, i32]),
        "qp_batch_digests": (i32, [vp, vp, i32]),
        "qp_batch_digests_len": (sz, [vp]),
        "qp_batch_leaves": (i32, [vp, sz, sz, vp, i32]),
        "qp_batch_leaf_len": (sz, [vp]),
        "qp_batch_get_lde_values": (i32, [vp, sz, sz, vp]),
        "qp_batch_get_leaves": (i32, [vp, vp, u32, vp]),
        "qp_batch_prove": (i32, [vp, sz, vp]),
        "qp_batch_timing": (i32, [vp, C.POINTER(C.c_double)]),
        "qp_batch_kernel_timing": (i32, [vp, C.POINTER(C.c_double)]),
        "qp_batch_device_lde": (vp, [vp]),
        "qp_batch_device_coeffs": (vp, [vp]),
        "qp_merkle_tree_new": (i32, [vp, vp, i32, sz, sz, u32, pp]),
        "qp_tree_free": (None, [vp]),
        "qp_tree_cap": (i32, [vp, vp, i32]),
        "qp_tree_digests": (i32, [vp, vp, i32]),
        "qp_tree_digests_len": (sz, [vp]),
        "qp_tree_prove": (i32, [vp, sz, vp]),
        "qp_tree_get": (i32, [vp, sz, vp]),
        "qp_poseidon_permute": (i32, [vp, vp, i32, sz]),
        "qp_coset_fft": (i32, [vp, vp, i32, sz, u32, u64, i32, vp, i32]),
        "qp_fri_begin": (i32, [vp, vp, vp, i32, u32, u32, u32, pp]),
        "qp_fri_commit_round": (i32, [vp, u32, vp]),
        "qp_fri_fold_round": (i32, [vp, vp, i32]),
        "qp_fri_final_poly": (i32, [vp, vp, C.POINTER(sz)]),
        "qp_fri_tree_get": (i32, [vp, u32, sz, vp]),
        "qp_fri_tree_prove": (i32, [vp, u32, sz, vp]),
        "qp_fri_tree_digests": (i32, [vp, u32, vp, i32]),
        "qp_fri_tree_digests_len": (sz, [vp, u32]),
        "qp_fri_num_rounds": (u32, [vp]),
        "qp_fri_free": (None, [vp]),
        "qp_fri_proof_of_work": (i32, [vp, vp, u32, u32, u64p]),
        "qp_batch_eval_polys": (i32, [vp, vp, vp]),
        "qp_batch_prove_many": (i32, [vp, vp, u32, vp]),
        "qp_fri_tree_open_many": (i32, [vp, u32, vp, u32, vp, vp]),
        "qp_fri_proof": (i32, [vp, pp, sz, vp, C.POINTER(_ChallengerState), u32, u32, C.POINTER(u32), u32, u32, u32,
                               vp, sz, C.POINTER(sz)]),
        "qp_fri_proof_len": (sz, [C.POINTER(sz), sz, u32, u32, u32, C.POINTER(u32), u32, u32]),
        "qp_fri_begin_from_openings": (i32, [vp, vp, sz, u32, u32, u32, pp]),
        "qp_fri_initial_coeffs": (i32, [vp, vp]),
        "qp_fri_run_commit_phase": (i32, [vp, u32, C.POINTER(u32), u32, C.POINTER(_ChallengerState), vp, vp,
                                          C.POINTER(sz)]),
        # host transcript mirror (include/qp_plonky2_host.h)
        "qp_challenger_init": (None, [C.POINTER(_ChallengerState)]),
        "qp_challenger_observe": (None, [C.POINTER(_ChallengerState), vp, sz]),
        "qp_challenger_get": (u64, [C.POINTER(_ChallengerState)]),
        "qp_fri_reduction_arity_bits": (u32, [u32, u32, u32, u32, u32, C.POINTER(u32)]),
        "qp_fri_committed_trees": (i32, [vp, vp, vp, i32, u32, u32, u32, C.POINTER(u32), u32,
                                         C.POINTER(_ChallengerState), vp, vp, C.POINTER(sz), pp]),
        "qp_fri_grind": (i32, [vp, C.POINTER(_ChallengerState), u32, u64p]),
        # plonk permutation argument and quotient (plonk.py)
        "qp_circuit_create": (i32, [vp, vp, pp]),
        "qp_circuit_free": (None, [vp]),
        "qp_circuit_partial_products_and_zs": (i32, [vp, vp, i32, vp, vp, vp, i32]),
        "qp_circuit_compute_quotient_polys": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
        "qp_circuit_describe": (i32, [vp, vp]),
        "qp_circuit_has_sigmas": (i32, [vp]),
        "qp_dev_alloc": (i32, [vp, sz, pp]),
        "qp_batch_open_many": (i32, [vp, vp, u32, vp, vp]),
        "qp_batch_describe": (i32, [vp, vp]),
        "qp_batch_serialized_len": (sz, [vp]),
        "qp_batch_serialize": (i32, [vp, vp, sz]),
        "qp_batch_deserialize": (i32, [vp, vp, sz, pp, C.POINTER(sz)]),
        "qp_batch_fri_from_values": (i32, [vp, vp, vp, sz, i32, u32, i32, u32, pp]),
        "qp_batch_fri_from_coeffs": (i32, [vp, vp, vp, sz, i32, u32, i32, u32, pp]),
        "qp_batch_fri_free": (None, [vp]),
        "qp_batch_fri_num_groups": (sz, [vp]),
        "qp_batch_fri_group": (i32, [vp, sz, C.POINTER(u32), C.POINTER(sz)]),
        "qp_batch_fri_coeffs": (i32, [vp, sz, vp, i32]),
        "qp_batch_fri_cap": (i32, [vp, vp, i32]),
        "qp_batch_fri_digests_len": (sz, [vp]),
        "qp_batch_fri_digests": (i32, [vp, vp, i32]),
        "qp_batch_fri_open": (i32, [vp, sz, vp]),
        "qp_batch_fri_values": (i32, [vp, sz, vp]),
        "qp_batch_fri_group_batch": (vp, [vp, sz]),
        "qp_fri_mix_values": (i32, [vp, vp, vp]),
        "qp_fri_domain_bits": (u32, [vp]),
        "qp_batch_fri_run_commit_phase": (i32, [vp, pp, sz, u32, C.POINTER(u32), u32, C.POINTER(_ChallengerState), vp, vp,
                                                C.POINTER(sz)]),
        "qp_batch_fri_proof": (i32, [vp, pp, sz, vp, pp, sz, C.POINTER(_ChallengerState), u32, u32, C.POINTER(u32), u32,
                                     u32, u32, vp, sz, C.POINTER(sz)]),
        "qp_batch_merkle_tree_new": (i32, [vp, vp, i32, vp, vp, sz, u32, pp]),
        "qp_batch_tree_free": (None, [vp]),
        "qp_batch_tree_cap": (i32, [vp, vp, i32]),
        "qp_batch_tree_digests_len": (sz, [vp]),
        "qp_batch_tree_digests": (i32, [vp, vp, i32]),
        "qp_batch_tree_open": (i32, [vp, sz, vp]),
        "qp_batch_tree_values": (i32, [vp, sz, vp]),
        "qp_dev_free": (None, [vp, vp]),
        "qp_memcpy": (i32, [vp, vp, i32, vp, i32, sz]),
        "qp_memcpy_peer": (i32, [vp, vp, vp, vp, sz]),
        "qp_mbatch_from_device": (i32, [vp, vp, i32, sz, u32, u32, i32, u32, vp, pp]),
        "qp_circuit_quotient_values_shard": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(sz), C.POINTER(sz)]),
        "qp_circuit_quotient_finish": (i32, [vp, vp, vp, i32]),
        "qp_fri_proof_sharded": (i32, [vp, pp, sz, u32, vp, C.POINTER(_ChallengerState), u32, u32, C.POINTER(u32), u32,
                                       u32, u32, vp, sz, C.POINTER(sz)]),
        # host side of the quotient / prove (include/qp_plonky2_host.h)
        "qp_program_create": (i32, [vp, sz, u32, pp]),
        "qp_program_create_lookups": (i32, [vp, sz, u32, u32, vp, sz, pp]),
        "qp_keccak256": (None, [vp, sz, vp]),
        "qp_circuit_lookup_polys": (i32, [vp, vp, i32, vp, vp, i32]),
        "qp_circuit_set_lookup_challenges": (i32, [vp, vp]),
        "qp_program_from_dag": (i32, [vp, sz, vp, sz, vp, sz, pp]),
        "qp_program_free": (None, [vp]),
        "qp_program_code": (sz, [vp, pp]),
        "qp_program_pool": (sz, [vp, pp]),
        "qp_program_regs": (u32, [vp]),
        "qp_program_segments": (sz, [vp, pp]),
        "qp_program_num_selectors": (u32, [vp]),
        "qp_program_num_gate_constants": (u32, [vp]),
        "qp_program_num_gate_constraints": (u32, [vp]),
        "qp_program_gate": (i32, [vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "qp_hash_no_pad": (None, [vp, sz, vp]),
        "qp_circuit_digest": (None, [vp, sz, u32, vp]),
        "qp_prove": (i32, [vp, vp, vp, vp, vp, vp, i32, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(C.c_double)]),
        "qp_mprove": (i32, [vp, vp, vp, vp, vp, vp, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(C.c_double)]),
        "qp_prove_zk": (i32, [vp, vp, vp, vp, vp, vp, i32, vp, sz, vp, vp, vp, vp, sz, C.POINTER(sz), C.POINTER(C.c_double)]),
        "qp_prove_cols": (i32, [vp, vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp, sz, C.POINTER(sz), C.POINTER(C.c_double)]),
        "qp_mprove_cols": (i32, [vp, vp, vp, vp, vp, vp, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    L._exported = sorted(sig)
    _lib = L
    return L


# ---------------------------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _buf(x):
    """-> (pointer, space, keepalive, shape) for numpy (host) or torch CUDA (device) arrays."""
    if _is_torch(x):
        assert x.is_contiguous() and x.element_size() == 8
        space = QP_DEVICE if x.is_cuda else QP_HOST
        return C.c_void_p(x.data_ptr()), space, x, tuple(x.shape)
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    return C.c_void_p(a.ctypes.data), QP_HOST, a, a.shape


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


class Context:
    """One per (process, GPU): stream, twiddle table, scratch pool."""

    def __init__(self, device: int = 0, max_lde_log: int = 24, stream=None, private_stream: bool = False):
        """stream: a cudaStream_t handle the library launches on.  Default (None): torch's current
        stream for `device` when torch is loaded with CUDA -- device tensors handed to the library
        and device outputs it produces are then ordered with the caller's torch work without any
        extra synchronisation (the legacy default stream is passed as cudaStreamLegacy).
        private_stream=True asks for the library's own non-blocking stream instead; the caller then
        owns the ordering (see "stream ordering" in include/qp_plonky2_b200.h)."""
        self._h = C.c_void_p()
        if stream is None and not private_stream:
            import sys

            torch = sys.modules.get("torch")
            if torch is not None and torch.cuda.is_available():
                stream = torch.cuda.current_stream(device).cuda_stream or 1   # 1 = cudaStreamLegacy
        rc = lib().qp_ctx_create(device, C.c_void_p(stream) if stream else None, max_lde_log, C.byref(self._h))
        if rc:
            raise QpError(rc, "qp_ctx_create failed (no usable CUDA device? there is no CPU fallback)")
        self.device = device

    def check(self, rc):
        if rc:
            raise QpError(rc, lib().qp_last_error(self._h).decode())

    def synchronize(self):
        self.check(lib().qp_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().qp_ctx_launch_count(self._h))

    def close(self):
        if self._h:
            lib().qp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- primitives ----
    def poseidon(self, states):
        """Batch of width-12 permutations (core/src/poseidon.rs:599-609)."""
        a = np.ascontiguousarray(np.asarray(states, dtype=np.uint64)).reshape(-1, 12).copy()
        self.check(lib().qp_poseidon_permute(self._h, _np_ptr(a), QP_HOST, a.shape[0]))
        return a

    def coset_fft(self, coeffs, shift=1, bit_reversed=False):
        """coset_fft_with_options (field/src/polynomial/mod.rs:280-293) of every row."""
        p, space, keep, shape = _buf(coeffs)
        n_vec, n = (1, shape[0]) if len(shape) == 1 else shape
        lg = int(n).bit_length() - 1
        if (1 << lg) != n:
            raise QpError(3, "Not a power of two: %d" % n)
        out = np.zeros((n_vec, n), dtype=np.uint64)
        self.check(lib().qp_coset_fft(self._h, p, space, n_vec, lg, shift, int(bit_reversed), _np_ptr(out), QP_HOST))
        return out.reshape(shape)

    def ifft_columns(self, values, out_device=None, sync=True):
        """The "IFFT" scope alone (oracle.rs:176-180); out_device: optional torch CUDA tensor (sync=False
        leaves the result stream-ordered on the context's stream, no host synchronisation)."""
        p, space, keep, shape = _buf(values)
        n_cols, n = shape
        lg = int(n).bit_length() - 1
        if (1 << lg) != n:
            raise QpError(3, "Not a power of two: %d" % n)
        if out_device is not None:
            self.check(lib().qp_ifft_columns(self._h, p, space, n_cols, lg, C.c_void_p(out_device.data_ptr()), QP_DEVICE))
            if sync:
                self.synchronize()
            return out_device
        out = np.zeros((n_cols, n), dtype=np.uint64)
        self.check(lib().qp_ifft_columns(self._h, p, space, n_cols, lg, _np_ptr(out), QP_HOST))
        return out


class _BorrowedContext(Context):
    """A context owned by someone else (a MultiContext's per-device context): same methods, never destroyed here."""

    def __init__(self, handle, device):
        self._h = C.c_void_p(handle)
        self.device = device

    def close(self):
        self._h = C.c_void_p()


class MultiContext:
    """One process, a list of GPUs (qp_mctx): the coset-sharded commit driven from inside the library, peer
    copies over NVLink between the devices, no torch.distributed."""

    def __init__(self, devices, max_lde_log: int = 24):
        self.devices = list(devices)
        self._h = C.c_void_p()
        arr = (C.c_int * len(self.devices))(*self.devices)
        rc = lib().qp_mctx_create(arr, len(self.devices), max_lde_log, C.byref(self._h))
        if rc:
            raise QpError(rc, "qp_mctx_create failed (devices must be distinct, a power of two of them)")
        self.contexts = [_BorrowedContext(lib().qp_mctx_ctx(self._h, i), d) for i, d in enumerate(self.devices)]

    def check(self, rc):
        if rc:
            raise QpError(rc, lib().qp_mctx_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().qp_mctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiBatch:
    """PolynomialBatch::from_values over every device of a MultiContext; `shards[i]` is the PolynomialBatch of
    device i (coset blocks [i 2^r / D, (i + 1) 2^r / D), leaf indices local to the shard)."""

    def __init__(self):
        self._h = C.c_void_p()

    @classmethod
    def from_values_cols(cls, mctx, columns, rate_bits, blinding, cap_height, salt=None):
        cols = [np.ascontiguousarray(np.asarray(c, dtype=np.uint64).ravel()) for c in columns]
        if len(cols) == 0:
            raise QpError(5, "polynomials[0]: index out of bounds (empty batch)")
        if len({c.size for c in cols}) != 1:
            raise QpError(4, "Polynomial degrees inconsistent")
        n = cols[0].size
        lg = int(n).bit_length() - 1
        if n == 0 or (1 << lg) != n:
            raise QpError(3, "Not a power of two: %d" % n)
        sp = skeep = None
        if blinding:
            if salt is None:
                raise QpError(7, "blinding=True needs salt[4][N] (the reference draws it from its RNG)")
            skeep = np.ascontiguousarray(np.asarray(salt, dtype=np.uint64))
            sp = _np_ptr(skeep)
        ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        self = cls()
        self.mctx, self.cap_height = mctx, cap_height
        mctx.check(lib().qp_mbatch_from_values_cols(mctx._h, ptrs, len(cols), lg, rate_bits, int(bool(blinding)),
                                                    cap_height, sp, C.byref(self._h)))
        D = int(lib().qp_mbatch_num_shards(self._h))
        blocks = (1 << rate_bits) // D
        self.shards = []
        for i in range(D):
            b = PolynomialBatch()
            b._h = C.c_void_p(lib().qp_mbatch_shard(self._h, i))
            b.ctx = mctx.contexts[i]
            b._borrowed = True
            b._describe(len(cols), lg, rate_bits, blinding, cap_height, i * blocks, blocks)
            self.shards.append(b)
        return self

    @property
    def cap(self):
        out = np.zeros((1 << self.cap_height, 4), dtype=np.uint64)
        self.mctx.check(lib().qp_mbatch_cap(self._h, _np_ptr(out)))
        return out

    def free(self):
        if self._h:
            for b in self.shards:
                b._h = C.c_void_p()
            lib().qp_mbatch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _Cap:
    pass


class BatchMerkleView:
    """The `merkle_tree` field of a PolynomialBatch: cap / digests / leaves served from the
    device on demand (materialising 9 GB of leaves per commit on the host would erase the win,
    SURVEY.md section 7)."""

    def __init__(self, batch):
        self._b = batch

    @property
    def cap(self):
        b = self._b
        out = np.zeros((lib().qp_batch_cap_len(b._h), 4), dtype=np.uint64)
        b.ctx.check(lib().qp_batch_cap(b._h, _np_ptr(out), QP_HOST))
        return out

    @property
    def digests(self):
        b = self._b
        out = np.zeros((lib().qp_batch_digests_len(b._h), 4), dtype=np.uint64)
        if out.size:
            b.ctx.check(lib().qp_batch_digests(b._h, _np_ptr(out), QP_HOST))
        return out

    def leaves(self, first=0, count=None):
        b = self._b
        if count is None:
            count = b.n_local_leaves - first
        out = np.zeros((count, b.leaf_len), dtype=np.uint64)
        if out.size:
            b.ctx.check(lib().qp_batch_leaves(b._h, first, count, _np_ptr(out), QP_HOST))
        return out

    def get(self, i):
        return self.leaves(i, 1)[0]

    def open_many(self, indices):
        """rows and Merkle paths of many leaves in one round trip -> (rows [n][leaf_len], paths [n][layers][4])"""
        b = self._b
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
        layers = b.local_lg_leaves - b.local_cap_height
        rows = np.zeros((idx.size, b.leaf_len), dtype=np.uint64)
        paths = np.zeros((idx.size, layers, 4), dtype=np.uint64)
        b.ctx.check(lib().qp_batch_open_many(b._h, _np_ptr(idx), idx.size, _np_ptr(rows), _np_ptr(paths) if layers else None))
        return rows, paths

    def get_many(self, indices):
        b = self._b
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
        out = np.zeros((idx.size, b.leaf_len), dtype=np.uint64)
        b.ctx.check(lib().qp_batch_get_leaves(b._h, _np_ptr(idx), idx.size, _np_ptr(out)))
        return out

    def prove(self, leaf_index):
        b = self._b
        k = b.local_lg_leaves - b.local_cap_height
        out = np.zeros((k, 4), dtype=np.uint64)
        b.ctx.check(lib().qp_batch_prove(b._h, leaf_index, _np_ptr(out) if k else None))
        return out


class PolynomialBatch:
    """plonky2/src/fri/oracle.rs:33-40.  Fields: polynomials (coefficients), merkle_tree,
    degree_log, rate_bits, blinding.  `timing` holds the reference's four TimingTree scopes."""

    SCOPES = ("IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree")

    def __init__(self):
        self._h = C.c_void_p()

    @classmethod
    def from_values(cls, ctx, values, rate_bits, blinding, cap_height, salt=None, block_first=0,
                    block_count=None):
        return cls._make(ctx, values, rate_bits, blinding, cap_height, salt, block_first, block_count, True)

    @classmethod
    def from_coeffs(cls, ctx, polynomials, rate_bits, blinding, cap_height, salt=None, block_first=0,
                    block_count=None):
        return cls._make(ctx, polynomials, rate_bits, blinding, cap_height, salt, block_first, block_count, False)

    @classmethod
    def from_values_cols(cls, ctx, columns, rate_bits, blinding, cap_height, salt=None, block_first=0,
                         block_count=None):
        """from_values on a list of separately allocated host columns (numpy uint64 vectors) -- the
        reference's `Vec<PolynomialValues<F>>` (oracle.rs:168-175): no flattening copy, the library stages
        the pageable columns itself (qp_batch_from_values_cols)."""
        cols = [np.ascontiguousarray(np.asarray(c, dtype=np.uint64).ravel()) for c in columns]
        if len(cols) == 0:
            raise QpError(5, "polynomials[0]: index out of bounds (empty batch)")
        if len({c.size for c in cols}) != 1:
            raise QpError(4, "Polynomial degrees inconsistent")
        n = cols[0].size
        lg = int(n).bit_length() - 1
        if n == 0 or (1 << lg) != n:
            raise QpError(3, "Not a power of two: %d" % n)
        if block_count is None:
            block_count = 1 << rate_bits
        sp = skeep = None
        if blinding:
            if salt is None:
                raise QpError(7, "blinding=True needs salt[4][N] (the reference draws it from its RNG)")
            skeep = np.ascontiguousarray(np.asarray(salt, dtype=np.uint64))
            if skeep.shape != (SALT_SIZE, n << rate_bits):
                raise QpError(5, "salt must be [4][N]")
            sp = _np_ptr(skeep)
        ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        self = cls()
        self.ctx = ctx
        rc = lib().qp_batch_from_values_cols(ctx._h, ptrs, len(cols), lg, rate_bits, int(bool(blinding)), cap_height,
                                             sp, block_first, block_count, C.byref(self._h))
        ctx.check(rc)
        self._describe(len(cols), lg, rate_bits, blinding, cap_height, block_first, block_count)
        return self

    def _describe(self, n_cols, lg, rate_bits, blinding, cap_height, block_first, block_count):
        self.n_cols, self.degree_log, self.rate_bits = n_cols, lg, rate_bits
        self.blinding, self.cap_height = bool(blinding), cap_height
        self.block_first, self.block_count = block_first, block_count
        self.leaf_len = int(lib().qp_batch_leaf_len(self._h))
        self.n_local_leaves = block_count << lg
        self.local_lg_leaves = self.n_local_leaves.bit_length() - 1
        self.local_cap_height = int(lib().qp_batch_cap_len(self._h)).bit_length() - 1
        self.merkle_tree = BatchMerkleView(self)
        ms = (C.c_double * 4)()
        lib().qp_batch_timing(self._h, ms)
        self.timing = dict(zip(self.SCOPES, list(ms)))
        lib().qp_batch_kernel_timing(self._h, ms)
        self.kernel_ms = dict(zip(("intt", "lde", "leaf_hash", "tree_levels"), list(ms)))

    @classmethod
    def _make(cls, ctx, data, rate_bits, blinding, cap_height, salt, block_first, block_count, is_values):
        if not _is_torch(data):
            # Vec<PolynomialValues>: every column must have the same length (oracle.rs:277)
            rows = list(data) if not isinstance(data, np.ndarray) else data
            if len(rows) == 0:
                raise QpError(5, "polynomials[0]: index out of bounds (empty batch)")
            if not isinstance(rows, np.ndarray) and len({len(r) for r in rows}) != 1:
                raise QpError(4, "Polynomial degrees inconsistent")
        p, space, keep, shape = _buf(data)
        if len(shape) != 2:
            raise QpError(5, "expected [n_cols][n]")
        n_cols, n = shape
        lg = int(n).bit_length() - 1
        if n == 0 or (1 << lg) != n:
            raise QpError(3, "Not a power of two: %d" % n)
        if block_count is None:
            block_count = 1 << rate_bits
        sp = None
        skeep = None
        if blinding:
            if salt is None:
                raise QpError(7, "blinding=True needs salt[4][N] (the reference draws it from its RNG)")
            sp, sspace, skeep, sshape = _buf(salt)
            if sspace != space or tuple(sshape) != (SALT_SIZE, n << rate_bits):
                raise QpError(5, "salt must be [4][N] in the same memory space as the data")
        self = cls()
        self.ctx = ctx
        fn = lib().qp_batch_from_values if is_values else lib().qp_batch_from_coeffs
        rc = fn(ctx._h, p, space, n_cols, lg, rate_bits, int(bool(blinding)), cap_height, sp, block_first,
                block_count, C.byref(self._h))
        ctx.check(rc)
        self._describe(n_cols, lg, rate_bits, blinding, cap_height, block_first, block_count)
        return self

    # ---- write_polynomial_batch / read_polynomial_batch (serialization/mod.rs:1803-1822, 758-784) ----
    def to_bytes(self):
        n = int(lib().qp_batch_serialized_len(self._h))
        out = np.zeros(n, dtype=np.uint8)
        self.ctx.check(lib().qp_batch_serialize(self._h, out.ctypes.data, n))
        return out.tobytes()

    @classmethod
    def from_bytes(cls, ctx, data):
        """-> (batch, bytes consumed).  The device batch is rebuilt from the stored polynomials; the
        stored cap must be theirs."""
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        self = cls()
        self.ctx = ctx
        used = C.c_size_t()
        ctx.check(lib().qp_batch_deserialize(ctx._h, buf.ctypes.data, buf.size, C.byref(self._h), C.byref(used)))
        d = np.zeros(5, dtype=np.uint64)
        lib().qp_batch_describe(self._h, d.ctypes.data)
        self.n_cols, self.degree_log, self.rate_bits, self.cap_height = (int(x) for x in d[:4])
        self.blinding = bool(d[4])
        self.block_first, self.block_count = 0, 1 << self.rate_bits
        self.leaf_len = int(lib().qp_batch_leaf_len(self._h))
        self.n_local_leaves = self.block_count << self.degree_log
        self.local_lg_leaves = self.n_local_leaves.bit_length() - 1
        self.local_cap_height = self.cap_height
        self.merkle_tree = BatchMerkleView(self)
        self.timing, self.kernel_ms = {}, {}
        return self, int(used.value)

    # ---- from_coeffs in pieces (columns arriving over time, e.g. from an all-gather) ----
    @classmethod
    def begin(cls, ctx, n_cols, degree_log, rate_bits, blinding, cap_height, block_first=0, block_count=None):
        if block_count is None:
            block_count = 1 << rate_bits
        self = cls()
        self.ctx = ctx
        ctx.check(lib().qp_batch_begin(ctx._h, n_cols, degree_log, rate_bits, int(bool(blinding)), cap_height,
                                       block_first, block_count, C.byref(self._h)))
        self.n_cols, self.degree_log, self.rate_bits = n_cols, degree_log, rate_bits
        self.blinding, self.cap_height = bool(blinding), cap_height
        self.block_first, self.block_count = block_first, block_count
        return self

    def put_coeffs(self, coeffs, c0):
        """Coefficient columns [c0, c0 + len(coeffs)): copied in and extended (LDE) right away."""
        p, space, keep, shape = _buf(coeffs)
        if len(shape) != 2 or shape[1] != 1 << self.degree_log:
            raise QpError(4, "Polynomial degrees inconsistent")
        self.ctx.check(lib().qp_batch_put_coeffs(self._h, p, space, c0, shape[0]))

    def coeffs_slot(self, c0, count):
        """Device view (torch int64 tensor [count][n], zero-copy) of coefficient columns [c0, c0 + count) of
        a batch under construction: a producer (inverse transform, collective) writes them in place, then
        extend_columns(c0, count) runs their LDE."""
        import torch

        n = 1 << self.degree_log
        ptr = lib().qp_batch_coeffs_slot(self._h, c0)
        if not ptr or c0 + count > self.n_cols:
            raise QpError(5, "column range out of bounds")

        class _Raw:
            __cuda_array_interface__ = {"shape": (count, n), "typestr": "<i8", "data": (int(ptr), False), "version": 3}

        return torch.as_tensor(_Raw(), device=torch.device("cuda", self.ctx.device))

    def extend_columns(self, c0, count, absorb=True):
        """Columns [c0, c0 + count) are in place (coeffs_slot): LDE now; with `absorb` the leaf sponges also
        advance over every complete 8-column chunk of the extended column prefix."""
        self.ctx.check(lib().qp_batch_extend_columns(self._h, c0, count, int(bool(absorb))))

    def end(self, salt=None):
        sp, sspace, skeep = None, QP_HOST, None
        if salt is not None:
            sp, sspace, skeep, _ = _buf(salt)
        self.ctx.check(lib().qp_batch_end(self._h, sp, sspace))
        self.leaf_len = int(lib().qp_batch_leaf_len(self._h))
        self.n_local_leaves = self.block_count << self.degree_log
        self.local_lg_leaves = self.n_local_leaves.bit_length() - 1
        self.local_cap_height = int(lib().qp_batch_cap_len(self._h)).bit_length() - 1
        self.merkle_tree = BatchMerkleView(self)
        ms = (C.c_double * 4)()
        lib().qp_batch_timing(self._h, ms)
        self.timing = dict(zip(self.SCOPES, list(ms)))
        lib().qp_batch_kernel_timing(self._h, ms)
        self.kernel_ms = dict(zip(("intt", "lde", "leaf_hash", "tree_levels"), list(ms)))
        return self

    @property
    def polynomials(self):
        out = np.zeros((self.n_cols, 1 << self.degree_log), dtype=np.uint64)
        self.ctx.check(lib().qp_batch_coeffs(self._h, _np_ptr(out), QP_HOST))
        return out

    def get_lde_values(self, index, step=1):
        out = np.zeros(self.n_cols, dtype=np.uint64)
        self.ctx.check(lib().qp_batch_get_lde_values(self._h, index, step, _np_ptr(out)))
        return out

    def eval_polys(self, point):
        """Every polynomial of the batch evaluated at an F_p^2 point (OpeningSet::new,
        plonky2/src/plonk/proof.rs:289-327) -> [n_cols][2]."""
        pt = np.array([int(point[0]), int(point[1])], dtype=np.uint64)
        out = np.zeros((self.n_cols, 2), dtype=np.uint64)
        self.ctx.check(lib().qp_batch_eval_polys(self._h, _np_ptr(pt), _np_ptr(out)))
        return out

    @property
    def device_lde_ptr(self):
        return lib().qp_batch_device_lde(self._h)

    @property
    def device_coeffs_ptr(self):
        return lib().qp_batch_device_coeffs(self._h)

    def free(self):
        if self._h and self.ctx._h and not getattr(self, "_borrowed", False):
            lib().qp_batch_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MerkleTree:
    """plonky2/src/hash/merkle_tree.rs:163-207 on caller-provided leaf-major rows."""

    def __init__(self, ctx, leaves, cap_height, shard=0, n_shards=1):
        """shard / n_shards: the multi-GPU form -- `leaves` are the rows of cap subtrees
        [shard * 2^h / n_shards, ...) only (qp_merkle_tree_new_shard); `cap` / `digests` are then this
        shard's slices of the reference's arrays and leaf indices are local."""
        self._h = C.c_void_p()
        self.ctx = ctx
        if not _is_torch(leaves) and not isinstance(leaves, np.ndarray):
            leaves = list(leaves)
            if len({len(r) for r in leaves}) > 1:
                raise QpError(5, "ragged leaves are not supported by the device path")
            leaves = np.asarray(leaves, dtype=np.uint64).reshape(len(leaves), -1)
        p, space, keep, shape = _buf(leaves)
        n, L = shape
        if n_shards != 1:
            ctx.check(lib().qp_merkle_tree_new_shard(ctx._h, p, space, n * n_shards, L, cap_height, shard, n_shards,
                                                     C.byref(self._h)))
            cap_height -= n_shards.bit_length() - 1
        else:
            ctx.check(lib().qp_merkle_tree_new(ctx._h, p, space, n, L, cap_height, C.byref(self._h)))
        self.n_leaves, self.leaf_len, self.cap_height = n, L, cap_height

    @property
    def cap(self):
        out = np.zeros((1 << self.cap_height, 4), dtype=np.uint64)
        self.ctx.check(lib().qp_tree_cap(self._h, _np_ptr(out), QP_HOST))
        return out

    @property
    def digests(self):
        out = np.zeros((lib().qp_tree_digests_len(self._h), 4), dtype=np.uint64)
        if out.size:
            self.ctx.check(lib().qp_tree_digests(self._h, _np_ptr(out), QP_HOST))
        return out

    def get(self, i):
        out = np.zeros(self.leaf_len, dtype=np.uint64)
        self.ctx.check(lib().qp_tree_get(self._h, i, _np_ptr(out) if out.size else None))
        return out

    def prove(self, i):
        k = self.n_leaves.bit_length() - 1 - self.cap_height
        out = np.zeros((k, 4), dtype=np.uint64)
        self.ctx.check(lib().qp_tree_prove(self._h, i, _np_ptr(out) if k else None))
        return out

    def free(self):
        if self._h and self.ctx._h:
            lib().qp_tree_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BatchMerkleTree:
    """plonky2/src/hash/batch_merkle_tree.rs: `leaves` = matrices (2-D arrays, row-major) of strictly
    decreasing power-of-two heights; same fields and methods as the reference (`cap`, `digests`,
    `leaf_heights`, `open_batch`, `values`)."""

    def __init__(self, ctx, leaves, cap_height):
        self._h = C.c_void_p()
        self.ctx, self.cap_height = ctx, cap_height
        mats = [np.ascontiguousarray(np.asarray(m, dtype=np.uint64).reshape(len(m), -1)) for m in leaves]
        if not mats:
            raise QpError(1, "no leaves")
        self.heights = [m.shape[0] for m in mats]
        self.widths = [m.shape[1] for m in mats]
        self.leaf_heights = [h.bit_length() - 1 for h in self.heights]
        n = len(mats)
        ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in mats])
        hs = (C.c_size_t * n)(*self.heights)
        ws = (C.c_size_t * n)(*self.widths)
        ctx.check(lib().qp_batch_merkle_tree_new(ctx._h, ptrs, QP_HOST, hs, ws, n, cap_height, C.byref(self._h)))

    @property
    def cap(self):
        out = np.zeros((1 << self.cap_height, 4), dtype=np.uint64)
        self.ctx.check(lib().qp_batch_tree_cap(self._h, _np_ptr(out), QP_HOST))
        return out

    @property
    def digests(self):
        out = np.zeros((lib().qp_batch_tree_digests_len(self._h), 4), dtype=np.uint64)
        if out.size:
            self.ctx.check(lib().qp_batch_tree_digests(self._h, _np_ptr(out), QP_HOST))
        return out

    def open_batch(self, leaf_index):
        k = self.leaf_heights[0] - self.cap_height
        out = np.zeros((k, 4), dtype=np.uint64)
        self.ctx.check(lib().qp_batch_tree_open(self._h, leaf_index, _np_ptr(out) if k else None))
        return out

    def values(self, leaf_index):
        out = np.zeros(sum(self.widths), dtype=np.uint64)
        self.ctx.check(lib().qp_batch_tree_values(self._h, leaf_index, _np_ptr(out) if out.size else None))
        return np.split(out, np.cumsum(self.widths)[:-1])

    def free(self):
        if self._h and self.ctx._h:
            lib().qp_batch_tree_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BatchFriOracle:
    """plonky2/src/batch_fri/oracle.rs:30-160 (the commitment; blinding = False): polynomials of
    non-increasing power-of-two length -> `polynomials` (coefficients), `degree_bits` (deduplicated,
    tallest first), and the BatchMerkleTree over the per-degree LDE matrices (`cap`, `digests`,
    `open_batch`, `values`)."""

    def __init__(self):
        self._h = C.c_void_p()

    @classmethod
    def from_values(cls, ctx, values, rate_bits, blinding, cap_height):
        return cls._make(ctx, values, rate_bits, blinding, cap_height, True)

    @classmethod
    def from_coeffs(cls, ctx, polynomials, rate_bits, blinding, cap_height):
        return cls._make(ctx, polynomials, rate_bits, blinding, cap_height, False)

    @classmethod
    def _make(cls, ctx, polys, rate_bits, blinding, cap_height, is_values):
        arrs = [np.ascontiguousarray(np.asarray(p, dtype=np.uint64).ravel()) for p in polys]
        for a in arrs:
            if a.size == 0 or a.size & (a.size - 1):
                raise QpError(3, "Not a power of two: %d" % a.size)
        self = cls()
        self.ctx, self.rate_bits, self.cap_height, self.blinding = ctx, rate_bits, cap_height, bool(blinding)
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        bits = (C.c_uint32 * max(n, 1))(*[a.size.bit_length() - 1 for a in arrs])
        fn = lib().qp_batch_fri_from_values if is_values else lib().qp_batch_fri_from_coeffs
        ctx.check(fn(ctx._h, ptrs, bits, n, QP_HOST, rate_bits, int(bool(blinding)), cap_height, C.byref(self._h)))
        self.degree_bits, self.group_sizes = [], []
        for g in range(int(lib().qp_batch_fri_num_groups(self._h))):
            d, k = C.c_uint32(), C.c_size_t()
            lib().qp_batch_fri_group(self._h, g, C.byref(d), C.byref(k))
            self.degree_bits.append(int(d.value))
            self.group_sizes.append(int(k.value))
        return self

    @property
    def polynomials(self):
        out = []
        for g, (d, k) in enumerate(zip(self.degree_bits, self.group_sizes)):
            co = np.zeros((k, 1 << d), dtype=np.uint64)
            self.ctx.check(lib().qp_batch_fri_coeffs(self._h, g, _np_ptr(co), QP_HOST))
            out += list(co)
        return out

    @property
    def cap(self):
        out = np.zeros((1 << self.cap_height, 4), dtype=np.uint64)
        self.ctx.check(lib().qp_batch_fri_cap(self._h, _np_ptr(out), QP_HOST))
        return out

    @property
    def digests(self):
        out = np.zeros((lib().qp_batch_fri_digests_len(self._h), 4), dtype=np.uint64)
        if out.size:
            self.ctx.check(lib().qp_batch_fri_digests(self._h, _np_ptr(out), QP_HOST))
        return out

    def open_batch(self, leaf_index):
        k = self.degree_bits[0] + self.rate_bits - self.cap_height
        out = np.zeros((k, 4), dtype=np.uint64)
        self.ctx.check(lib().qp_batch_fri_open(self._h, leaf_index, _np_ptr(out) if k else None))
        return out

    def values(self, leaf_index):
        out = np.zeros(sum(self.group_sizes), dtype=np.uint64)
        self.ctx.check(lib().qp_batch_fri_values(self._h, leaf_index, _np_ptr(out)))
        return np.split(out, np.cumsum(self.group_sizes)[:-1])

    @property
    def leaf_len(self):
        return sum(self.group_sizes)

    def locate(self, polynomial_index):
        """(group, column) of `polynomials[polynomial_index]` (the polynomials are sorted tallest first)."""
        for g, k in enumerate(self.group_sizes):
            if polynomial_index < k:
                return g, polynomial_index
            polynomial_index -= k
        raise IndexError("polynomial index out of range")

    def group_batch(self, g):
        """Group g as a PolynomialBatch handle (what an opening term refers to); owned by this oracle."""
        return _BorrowedBatch(lib().qp_batch_fri_group_batch(self._h, g))

    @staticmethod
    def prove_openings(ctx, degree_bits, instances, oracles, challenger, rate_bits, cap_height, arity_bits,
                       proof_of_work_bits, num_query_rounds) -> bytes:
        """BatchFriOracle::prove_openings (plonky2/src/batch_fri/oracle.rs:163-229) -> the FriProof bytes.
        instances[i] (one per entry of degree_bits, tallest first): dict(batches=[dict(point=(a, b),
        openings=[expression, ...])]); an expression is a list of terms (oracle_index, polynomial_index,
        coefficient) with coefficient "one" | ("point_power", k) | ("constant", (c0, c1))
        (core/src/fri_structure.rs:59-108)."""
        if len(degree_bits) != len(instances):
            raise QpError(5, "degree_bits and instances differ in length")
        alpha = challenger.get_extension_challenge()
        fris = []
        try:
            for d, inst in zip(degree_bits, instances):
                flat = []
                for b in inst["batches"]:
                    terms, apow = [], (1, 0)
                    for expr in b["openings"]:                     # reduce_polys: sum_i alpha^i expr_i
                        for oi, pi, coeff in expr:
                            g, col = oracles[oi].locate(pi)
                            if oracles[oi].degree_bits[g] != d:
                                raise QpError(6, "opening term of another degree")
                            terms.append((oracles[oi].group_batch(g), col, _ext_mul(apow, _coefficient(coeff, b["point"]))))
                        apow = _ext_mul(apow, alpha)
                    flat.append(dict(point=b["point"], shift=apow, terms=terms))   # shift_poly: alpha^len(openings)
                fris.append(fri_from_openings(ctx, flat, d, rate_bits, cap_height))
            return batch_fri_proof(ctx, oracles, fris, challenger, rate_bits, cap_height, arity_bits, proof_of_work_bits,
                                   num_query_rounds)
        finally:
            for f in fris:
                f.free()

    def free(self):
        if self._h and self.ctx._h:
            lib().qp_batch_fri_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _BorrowedBatch:
    """A qp_batch handle owned by something else (a group of a BatchFriOracle)."""

    def __init__(self, h):
        self._h = C.c_void_p(h)


_GL_P = 0xFFFFFFFF00000001


def _ext_mul(a, b):      # F_p^2 = F_p[X] / (X^2 - 7)
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % _GL_P, (a[0] * b[1] + a[1] * b[0]) % _GL_P)


def _coefficient(coeff, point):
    """FriCoefficient (core/src/fri_structure.rs:100-108) evaluated at the batch point."""
    if coeff == "one" or coeff == ("one",):
        return (1, 0)
    if coeff[0] == "point_power":
        r, base, e = (1, 0), (int(point[0]) % _GL_P, int(point[1]) % _GL_P), int(coeff[1])
        while e:
            if e & 1:
                r = _ext_mul(r, base)
            base = _ext_mul(base, base)
            e >>= 1
        return r
    if coeff[0] == "constant":
        return (int(coeff[1][0]) % _GL_P, int(coeff[1][1]) % _GL_P)
    raise ValueError("unknown coefficient %r" % (coeff,))


class Challenger:
    """core/src/challenger.rs -- host-side duplex sponge (see include/qp_plonky2_host.h)."""

    def __init__(self):
        self._s = _ChallengerState()
        lib().qp_challenger_init(C.byref(self._s))

    def observe_elements(self, elems):
        a = np.ascontiguousarray(np.asarray(elems, dtype=np.uint64).reshape(-1))
        if a.size:
            lib().qp_challenger_observe(C.byref(self._s), _np_ptr(a), a.size)

    observe_cap = observe_elements

    def observe_element(self, e):
        self.observe_elements([e])

    def get_challenge(self) -> int:
        return int(lib().qp_challenger_get(C.byref(self._s)))

    def get_n_challenges(self, n):
        return [self.get_challenge() for _ in range(n)]

    def get_extension_challenge(self):
        return (self.get_challenge(), self.get_challenge())


def fri_reduction_arity_bits(degree_bits, rate_bits, cap_height, arity_bits=4, final_poly_bits=5):
    """FriReductionStrategy::ConstantArityBits (core/src/fri.rs:50-61)."""
    out = (C.c_uint * 64)()
    k = lib().qp_fri_reduction_arity_bits(degree_bits, rate_bits, cap_height, arity_bits, final_poly_bits, out)
    return [int(out[i]) for i in range(k)]


class FriCommitment:
    """Result of fri_committed_trees: per-round trees (device-resident) + final polynomial."""

    def __init__(self, ctx, handle, caps, final_poly, arity_bits, lg_n, cap_height):
        self.ctx, self._h = ctx, handle
        self.caps, self.final_poly = caps, final_poly
        self.arity_bits, self.lg_n, self.cap_height = list(arity_bits), lg_n, cap_height

    def tree_get(self, rnd, leaf_index):
        out = np.zeros(2 << self.arity_bits[rnd], dtype=np.uint64)
        self.ctx.check(lib().qp_fri_tree_get(self._h, rnd, leaf_index, _np_ptr(out)))
        return out

    def tree_prove(self, rnd, leaf_index):
        lg = self.lg_n - sum(self.arity_bits[: rnd + 1])
        k = lg - self.cap_height
        out = np.zeros((k, 4), dtype=np.uint64)
        self.ctx.check(lib().qp_fri_tree_prove(self._h, rnd, leaf_index, _np_ptr(out) if k else None))
        return out

    def tree_digests(self, rnd):
        out = np.zeros((lib().qp_fri_tree_digests_len(self._h, rnd), 4), dtype=np.uint64)
        if out.size:
            self.ctx.check(lib().qp_fri_tree_digests(self._h, rnd, _np_ptr(out), QP_HOST))
        return out

    def free(self):
        if self._h and self.ctx._h:
            lib().qp_fri_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fri_committed_trees(ctx, coeffs, values, challenger, rate_bits, cap_height, arity_bits):
    """plonky2/src/fri/prover.rs:85-143.  coeffs/values: F_p^2 arrays [n][2] (natural order)."""
    pc, space, k1, shape = _buf(coeffs)
    pv, space2, k2, shape2 = _buf(values)
    if space != space2 or tuple(shape) != tuple(shape2) or len(shape) != 2 or shape[1] != 2:
        raise QpError(5, "coeffs/values must both be [n][2] in the same memory space")
    n = shape[0]
    lg = int(n).bit_length() - 1
    if n == 0 or (1 << lg) != n:
        raise QpError(3, "Not a power of two: %d" % n)
    R = len(arity_bits)
    caps = np.zeros((R, 1 << cap_height, 4), dtype=np.uint64)
    tot = sum(arity_bits)
    final = np.zeros(((n >> tot) >> rate_bits, 2), dtype=np.uint64)
    ab = (C.c_uint * max(R, 1))(*arity_bits)
    h = C.c_void_p()
    flen = C.c_size_t()
    rc = lib().qp_fri_committed_trees(ctx._h, pc, pv, space, lg, rate_bits, cap_height, ab, R,
                                      C.byref(challenger._s), _np_ptr(caps), _np_ptr(final) if final.size else None,
                                      C.byref(flen), C.byref(h))
    ctx.check(rc)
    return FriCommitment(ctx, h, caps, final, arity_bits, lg, cap_height)


class _OpeningTerm(C.Structure):
    _fields_ = [("batch", C.c_void_p), ("poly_index", C.c_size_t), ("weight", C.c_uint64 * 2)]


class _OpeningBatch(C.Structure):
    _fields_ = [("point", C.c_uint64 * 2), ("terms", C.POINTER(_OpeningTerm)), ("n_terms", C.c_size_t),
                ("shift", C.c_uint64 * 2)]


def fri_from_openings(ctx, batches, degree_log, rate_bits, cap_height):
    """reduce_openings_to_unmasked_final_poly + padded coset FFT on the device
    (plonky2/src/fri/oracle.rs:129-165, 329-343).  batches: list of dict(point=(a, b), shift=(a, b),
    terms=[(PolynomialBatch, poly_index, (w0, w1)), ...]).  Returns a FriCommitment whose commit
    phase has not run yet (use fri_commit_phase)."""
    keep = []
    arr = (_OpeningBatch * len(batches))()
    for i, b in enumerate(batches):
        terms = (_OpeningTerm * max(len(b["terms"]), 1))()
        for k, (pb, idx, w) in enumerate(b["terms"]):
            terms[k].batch = pb._h
            terms[k].poly_index = idx
            terms[k].weight[0], terms[k].weight[1] = int(w[0]), int(w[1])
        keep.append(terms)
        arr[i].point[0], arr[i].point[1] = int(b["point"][0]), int(b["point"][1])
        arr[i].terms = terms
        arr[i].n_terms = len(b["terms"])
        arr[i].shift[0], arr[i].shift[1] = int(b["shift"][0]), int(b["shift"][1])
    h = C.c_void_p()
    ctx.check(lib().qp_fri_begin_from_openings(ctx._h, arr, len(batches), degree_log, rate_bits, cap_height, C.byref(h)))
    return FriCommitment(ctx, h, None, None, [], degree_log + rate_bits, cap_height)


def fri_commit_phase(fri, challenger, rate_bits, arity_bits):
    """fri_committed_trees' loop (prover.rs:93-126) on an existing device FRI state."""
    R = len(arity_bits)
    n = 1 << fri.lg_n
    caps = np.zeros((R, 1 << fri.cap_height, 4), dtype=np.uint64)
    final = np.zeros(((n >> sum(arity_bits)) >> rate_bits, 2), dtype=np.uint64)
    ab = (C.c_uint * max(R, 1))(*arity_bits)
    flen = C.c_size_t()
    fri.ctx.check(lib().qp_fri_run_commit_phase(fri._h, fri.cap_height, ab, R, C.byref(challenger._s), _np_ptr(caps),
                                                _np_ptr(final) if final.size else None, C.byref(flen)))
    fri.caps, fri.final_poly, fri.arity_bits = caps, final, list(arity_bits)
    return fri


def fri_proof(ctx, oracles, fri, challenger, rate_bits, arity_bits, proof_of_work_bits, num_query_rounds) -> bytes:
    """fri_proof (plonky2/src/fri/prover.rs:24-71) on a device FRI state, serialised like
    write_fri_proof (plonky2/src/util/serialization/mod.rs:1654-1667).  `oracles`: the initial
    PolynomialBatch trees the queries open, in order."""
    R = len(arity_bits)
    ab = (C.c_uint * max(R, 1))(*arity_bits)
    lens = (C.c_size_t * max(len(oracles), 1))(*[o.leaf_len for o in oracles])
    need = lib().qp_fri_proof_len(lens, len(oracles), fri.lg_n, rate_bits, fri.cap_height, ab, R, num_query_rounds)
    buf = (C.c_uint8 * need)()
    hs = (C.c_void_p * max(len(oracles), 1))(*[o._h for o in oracles])
    got = C.c_size_t()
    ctx.check(lib().qp_fri_proof(ctx._h, hs, len(oracles), fri._h, C.byref(challenger._s), rate_bits, fri.cap_height,
                                 ab, R, proof_of_work_bits, num_query_rounds, buf, need, C.byref(got)))
    assert got.value == need, (got.value, need)
    return bytes(buf)


def batch_fri_proof(ctx, oracles, fris, challenger, rate_bits, cap_height, arity_bits, proof_of_work_bits,
                    num_query_rounds) -> bytes:
    """batch_fri_proof (plonky2/src/batch_fri/prover.rs:25-80) serialised like write_fri_proof.  oracles:
    BatchFriOracle objects (the initial BatchMerkleTrees); fris[0]: the FRI state of the largest final polynomial,
    fris[1:]: those of the smaller ones, strictly decreasing (fri_from_openings / fri_begin)."""
    R = len(arity_bits)
    ab = (C.c_uint * max(R, 1))(*arity_bits)
    lens = (C.c_size_t * max(len(oracles), 1))(*[o.leaf_len for o in oracles])
    need = lib().qp_fri_proof_len(lens, len(oracles), fris[0].lg_n, rate_bits, cap_height, ab, R, num_query_rounds)
    buf = (C.c_uint8 * need)()
    hs = (C.c_void_p * max(len(oracles), 1))(*[o._h for o in oracles])
    lower = (C.c_void_p * max(len(fris) - 1, 1))(*[f._h for f in fris[1:]])
    got = C.c_size_t()
    ctx.check(lib().qp_batch_fri_proof(ctx._h, hs, len(oracles), fris[0]._h, lower, len(fris) - 1, C.byref(challenger._s),
                                       rate_bits, cap_height, ab, R, proof_of_work_bits, num_query_rounds, buf, need,
                                       C.byref(got)))
    assert got.value == need, (got.value, need)
    return bytes(buf)


def fri_begin(ctx, coeffs, values, rate_bits, cap_height):
    """A device FRI state from (lde_polynomial_coeffs, lde_polynomial_values), F_p^2 arrays [n][2] in natural
    order -- fri_proof's / batch_fri_proof's two polynomial arguments."""
    co = np.ascontiguousarray(np.asarray(coeffs, dtype=np.uint64))
    va = np.ascontiguousarray(np.asarray(values, dtype=np.uint64))
    n = co.shape[0]
    lg = int(n).bit_length() - 1
    if n == 0 or (1 << lg) != n or co.shape != va.shape:
        raise QpError(3, "Not a power of two / shapes differ")
    h = C.c_void_p()
    ctx.check(lib().qp_fri_begin(ctx._h, _np_ptr(co), _np_ptr(va), QP_HOST, lg, rate_bits, cap_height, C.byref(h)))
    return FriCommitment(ctx, h, None, None, [], lg, cap_height)


def fri_initial_coeffs(fri, degree_log):
    out = np.zeros((1 << degree_log, 2), dtype=np.uint64)
    fri.ctx.check(lib().qp_fri_initial_coeffs(fri._h, _np_ptr(out)))
    return out


def fri_proof_of_work(ctx, challenger, proof_of_work_bits) -> int:
    """plonky2/src/fri/prover.rs:159-208, smallest-witness rule."""
    w = C.c_uint64()
    ctx.check(lib().qp_fri_grind(ctx._h, C.byref(challenger._s), proof_of_work_bits, C.byref(w)))
    return int(w.value)
