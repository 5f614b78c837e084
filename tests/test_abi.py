"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/*.h declares; with no GPU the product path fails loudly instead of
falling back to anything."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def qp():
    import qp_plonky2_b200 as m

    m.build()
    return m


def declared_symbols():
    names = set()
    for h in ("qp_plonky2_b200.h", "qp_plonky2_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(qp_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol(qp):
    raw = ctypes.CDLL(qp.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(raw, s), "missing export: " + s
    # and the Python binding binds exactly the declared set
    assert sorted(qp.lib()._exported) == syms


def test_library_has_sm100a_code(qp):
    """the shipped .so carries sm_100a SASS (cuobjdump lists the cubin)"""
    import shutil
    import subprocess

    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cu, "-lelf", qp.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(qp):
    """Without a CUDA device the product path raises (QP_ERR_CUDA); it never computes on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(qp.QpError) as e:
        qp.Context(0)
    assert e.value.code == 1


def test_host_transcript_mirror(qp):
    """Challenger (host boundary consumer) against the oracle; FRI arity schedule."""
    import oracle

    a, b = qp.Challenger(), oracle.Challenger()
    for i in range(1, 10):
        xs = oracle.rand_felts(3 * i, i)
        a.observe_elements(xs)
        b.observe(xs)
        for _ in range(i):
            assert a.get_challenge() == b.get_challenge()
    for d in (12, 13, 14, 20, 23):
        assert qp.fri_reduction_arity_bits(d, 3, 4) == oracle.fri_reduction_arity_bits(d, 3, 4)


def test_product_does_not_import_oracle():
    """only tests/, smoke() and bench.py may touch oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qp-plonky2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "plonky2_oracle" not in text and "oracle/" not in text, f
