#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the hot path into small JSON fixtures.

Runs in the build container only (needs /root/reference); the fixtures it writes are committed,
so nothing at test time reads the reference.  Sources:

  poseidon_kat.json       core/src/poseidon_goldilocks.rs:455-490  -- the four width-12
                          permutation known-answer vectors ("calculated with (modified)
                          hadeshash reference implementation")
  reverse_index_bits.json plonky2/src/util/mod.rs:56-123 -- 256-entry bit-reversal table and
                          the [10,20,30,40] case
  field_inputs.json       field/src/prime_field_testing.rs:8-17 -- boundary inputs the
                          reference's field tests run add/sub/mul/neg/square over

    python tests/golden/make_golden.py
"""
import json
import os
import re

REF = os.environ.get("QP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
P = 0xFFFFFFFF00000001


def ints(text):
    return [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\b\d+\b", text)]


def poseidon_kat():
    src = open(os.path.join(REF, "core/src/poseidon_goldilocks.rs")).read()
    m = re.search(r"let test_vectors12[^=]*=\s*vec!\[(.*?)\n\s*\];", src, re.S)
    body = m.group(1).replace("neg_one", str(P - 1))
    body = re.sub(r"//[^\n]*", "", body)
    groups = re.findall(r"\[([^\[\]]*)\]", body)
    assert len(groups) == 8, len(groups)
    vecs = []
    for k in range(0, 8, 2):
        i, o = ints(groups[k]), ints(groups[k + 1])
        assert len(i) == 12 and len(o) == 12
        vecs.append({"input": [str(x) for x in i], "output": [str(x) for x in o]})
    return vecs


def reverse_table():
    src = open(os.path.join(REF, "plonky2/src/util/mod.rs")).read()
    m = re.search(r"let output256: Vec<u64> = vec!\[(.*?)\];", src, re.S)
    tab = ints(m.group(1))
    assert len(tab) == 256
    m2 = re.search(r"reverse_index_bits\(&\[([^\]]*)\]\), vec!\[([^\]]*)\]", src)
    return {"output256": tab, "small_in": ints(m2.group(1)), "small_out": ints(m2.group(2))}


def field_inputs():
    c = 10
    xs = (
        list(range(0, c))
        + list(range((1 << 31) - c, (1 << 31) + c))
        + list(range((1 << 32) - c, (1 << 32) + c))
        + list(range((1 << 63) - c, (1 << 63) + c))
        + list(range(P - c, P))
    )
    return [str(x) for x in xs if x < P]


def main():
    out = {
        "poseidon_kat.json": poseidon_kat(),
        "reverse_index_bits.json": reverse_table(),
        "field_inputs.json": field_inputs(),
    }
    for name, obj in out.items():
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(obj, f, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
