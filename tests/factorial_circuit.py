"""Test infrastructure: the reference's `factorial` example (plonky2/examples/factorial.rs; BASELINE.json configs[0])
as a REAL circuit -- the statement "initial * 2 * 3 * ... * 100 = result" with public inputs (initial, result) --
laid out by a minimal builder of its own:

    99 multiplications           ArithmeticGate ops (c0 = 1, c1 = 0), chained by copy constraints
    the constants 2 .. 100, 0    ConstantGate(2) rows, copied into the multiplications / the hash inputs
    public inputs                one PoseidonGate row hashes (initial, result, 0, ..) like hash_n_to_hash_no_pad and its
                                 four outputs are copied into the PublicInputGate row (circuit_builder.rs: the
                                 public-input hash is computed IN the circuit and tied to the gate's wires)

The reference's CircuitBuilder places gates, orders wire partitions and fills unused cells in its own way, so the proof
BYTES of its factorial circuit are not reproducible here (no Rust toolchain); the statement, the gate set, the
configuration (standard_recursion_config) and the prover path are the same, and the restated verifier decides.
Returns an object with SynthCircuit's interface (tests/synth_circuit.py)."""
import numpy as np

from qp_plonky2_b200 import plonk

import gate_witness
from synth_circuit import SynthCircuit, _mulmod

P = plonk.P


def factorial_circuit(initial=1, first=2, last=100, num_wires=143, num_routed_wires=80, num_challenges=2,
                      quotient_degree_factor=8, rate_bits=3, cap_height=4):
    nr = num_routed_wires
    ops_per_row = nr // 4
    factors = list(range(first, last + 1))
    n_arith = -(-len(factors) // ops_per_row)
    n_const = -(-(len(factors) + 1) // 2)                     # the factors and one 0, two constants per row
    rows_needed = 1 + n_const + n_arith + 1
    degree_bits = max(5, (rows_needed - 1).bit_length())
    n = 1 << degree_bits
    gates = [plonk.NoopGate(), plonk.ConstantGate(2), plonk.PublicInputGate(), plonk.ArithmeticGate.new_from_config(nr),
             plonk.PoseidonGate()]
    c = plonk.CommonCircuitData(degree_bits, gates, num_wires, nr, num_challenges, quotient_degree_factor, rate_bits,
                                cap_height)
    idx = {g.id().split(" ")[0].split("(")[0]: i for i, g in enumerate(c.gates)}
    gc0 = c.num_selectors + c.num_lookup_selectors
    row_gate = np.full(n, idx["NoopGate"], dtype=np.int64)
    PI_ROW = 0
    const_rows = list(range(1, 1 + n_const))
    arith_rows = list(range(1 + n_const, 1 + n_const + n_arith))
    POS_ROW = 1 + n_const + n_arith
    row_gate[PI_ROW] = idx["PublicInputGate"]
    row_gate[const_rows] = idx["ConstantGate"]
    row_gate[arith_rows] = idx["ArithmeticGate"]
    row_gate[POS_ROW] = idx["PoseidonGate"]
    consts = np.zeros((c.num_constants, n), dtype=np.uint64)
    for s, (a, b) in enumerate(c.groups):
        in_group = (row_gate >= a) & (row_gate < b)
        consts[s] = np.where(in_group, row_gate, plonk.UNUSED_SELECTOR if c.num_selectors > 1 else row_gate)
    wires = np.zeros((num_wires, n), dtype=np.uint64)
    classes = []                                              # copy constraints: lists of (column, row) that are equal

    # constants: value k lives in ConstantGate row const_rows[k // 2], wire k % 2 (constant.rs: wire_i = const_i)
    values = factors + [0]
    cell_of_const = {}
    for k, v in enumerate(values):
        r, w = const_rows[k // 2], k % 2
        consts[gc0 + w, r] = v
        wires[w, r] = v
        cell_of_const[v] = (w, r)
    # the chain cur_j = cur_(j-1) * factor_j (arithmetic_base.rs: out = c0 m0 m1 + c1 addend)
    cur, cur_cell = int(initial) % P, None
    initial_cells = []
    for j, f in enumerate(factors):
        r, op = arith_rows[j // ops_per_row], j % ops_per_row
        consts[gc0, r], consts[gc0 + 1, r] = 1, 0
        wires[4 * op, r], wires[4 * op + 1, r] = cur, f
        if cur_cell is None:
            initial_cells.append((4 * op, r))
        else:
            classes.append([cur_cell, (4 * op, r)])
        classes.append([cell_of_const[f], (4 * op + 1, r)])
        cur = cur * f % P
        wires[4 * op + 3, r] = cur
        cur_cell = (4 * op + 3, r)
    for r in arith_rows:                                      # unused operations of the last row: 0 * 0 = 0
        consts[gc0, r], consts[gc0 + 1, r] = 1, 0
    # public inputs: hash_n_to_hash_no_pad([initial, result]) in one PoseidonGate row (swap = 0)
    public_inputs = [int(initial) % P, cur]
    row = gate_witness.generate(plonk.PoseidonGate(), public_inputs + [0] * 10, 0)
    for w, v in row.items():
        wires[w, POS_ROW] = v
    classes.append(initial_cells + [(0, POS_ROW)])
    classes.append([cur_cell, (1, POS_ROW)])
    classes.append([cell_of_const[0]] + [(k, POS_ROW) for k in range(2, 12)])
    from qp_plonky2_b200 import prover as _prover
    pih = _prover.hash_no_pad(public_inputs)
    for k in range(4):
        assert int(wires[12 + k, POS_ROW]) == int(pih[k]), "the in-circuit hash is hash_no_pad of the public inputs"
        wires[k, PI_ROW] = pih[k]
        classes.append([(12 + k, POS_ROW), (k, PI_ROW)])
    # sigma: one cycle per class (circuit_builder.rs sigma_vecs does the same with its own cell order)
    sigma_row = np.tile(np.arange(n, dtype=np.int64), (nr, 1))
    sigma_col = np.tile(np.arange(nr, dtype=np.int64)[:, None], (1, n))
    merged = {}
    for cls in classes:                                       # union of classes that share a cell
        cells = set(cls)
        for cell in list(cells):
            if cell in merged:
                cells |= merged[cell]
        for cell in cells:
            merged[cell] = cells
    done = set()
    for cells in merged.values():
        key = id(cells)
        if key in done:
            continue
        done.add(key)
        cyc = sorted(cells)
        assert all(col < nr for col, _ in cyc), "copy constraints need routed wires"
        assert len({int(wires[col, r]) for col, r in cyc}) == 1, "a copy class holds one value"
        for (col, r), (ncol, nrow) in zip(cyc, cyc[1:] + cyc[:1]):
            sigma_col[col, r], sigma_row[col, r] = ncol, nrow
    w = plonk.primitive_root_of_unity(degree_bits)
    sub = np.ones(n, dtype=np.uint64)
    m, wm = 1, w
    while m < n:
        sub[m:2 * m] = _mulmod(sub[:m], np.uint64(wm))
        m, wm = 2 * m, wm * wm % P
    k_arr = np.array([int(k) for k in c.k_is], dtype=np.uint64)
    sc = object.__new__(SynthCircuit)
    sc.n, sc.common, sc.row_gate, sc.constants, sc.wires = n, c, row_gate, consts, wires
    sc.sigmas, sc.subgroup = _mulmod(k_arr[sigma_col], sub[sigma_row]), sub
    sc.public_inputs, sc.public_inputs_hash, sc._oracle_circuit = public_inputs, pih, None
    return sc
