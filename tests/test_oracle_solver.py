"""The witness-solver oracle: circuit_synth builds a BatchCreateUser-shaped system in the flat form the cgo shim
would export from gnark; oracle/py/solver.py restates gnark's level walk (run-time search for the unsolved wire, hints, mid-solve
commitment); oracle/c/orc_solver.c is the fast port.  Pinned here: the in-circuit Poseidon gadget equals the native hash (the
reference's own circuit test checks exactly that, circuit/batch_create_user_circuit_test.go:62-76), every constraint of the solved
system holds, C == Python, and the proof made from the solution passes the toxic-waste check."""
import numpy as np
import pytest

import groth16 as g16
import orc
import poseidon as ps
import solver
from bn254 import R
from helpers import circuit_instance, circuit_synth, oracle_poseidon_constants

SMALL = dict(users=3, assets_per_user=2, cex_assets=3, tiers=2, merkle_depth=2, chain_perms=3, limb_bits=8)


def test_poseidon_gadget_equals_native_hash():
    cs_mod = circuit_synth()
    cb = cs_mod.CircuitBuilder(1, oracle_poseidon_constants, limb_bits=8)
    cb.section("body", 1)
    xs = [cb.secret("field") for _ in range(14)]
    h2 = cb.materialise(cb.poseidon(xs[:2]))
    h5 = cb.materialise(cb.poseidon(xs[:5]))
    h14 = cb.materialise(cb.poseidon(xs))                     # 12 + 2: one width-13 permutation chained into a width-3 one
    flat = cb.flatten()
    inputs = [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in cs_mod.draw_inputs(flat, 3)]
    w, a, b, c, _ = solver.solve_program(flat, inputs)
    vals = inputs[1:]
    wire = lambda le: w[cb.flatten()["n_public"] + flat["n_secret"] + cs_mod.unref(next(iter(le.t)))[4]]
    assert wire(h2) == ps.poseidon(vals[:2]) and wire(h5) == ps.poseidon(vals[:5]) and wire(h14) == ps.poseidon(vals)


def test_python_solver_and_c_solver_agree():
    inst = circuit_instance(seed=5, with_key=False, **SMALL)
    flat = inst["flat"]
    challenge = lambda vals: 0x1234567 + len(vals)
    w, a, b, c, info = solver.solve_program(flat, inst["inputs"], challenge)
    assert all(x * y % R == z for x, y, z in zip(a, b, c))
    assert len(info["committed"]) == len(flat["private_committed"]) > 0
    # the gadget set is all there
    fns = set(int(x) for x in flat["hint_fn"])
    assert fns == {1, 2, 3, 4, 5, 6, 7, 8}
    cw, ca, cb_, cc = orc.solve(flat, inst["inputs_mont"], lambda v: challenge(v))
    assert orc.fr_unmont(cw) == w and orc.fr_unmont(ca) == a and orc.fr_unmont(cb_) == b and orc.fr_unmont(cc) == c
    # multiplicities really count: the range table's counters sum to the number of queries
    h = [i for i, f in enumerate(flat["hint_fn"]) if f == 7][0]
    first, n_out = int(flat["hint_out_first"][h]), int(flat["hint_n_out"][h])
    assert sum(w[first:first + n_out]) == int(flat["hint_in_end"][h]) - int(flat["hint_in_ptr"][h])


def test_unsatisfied_input_is_rejected():
    inst = circuit_instance(seed=6, with_key=False, **SMALL)
    flat = inst["flat"]
    # a 64-bit range-checked input set to 2^70: the limb decomposition no longer recomposes
    first, n_s, count, specs = [x for x in flat["secret_layout"] if any(k == "uint" for k, _ in x[3])][0]
    j = [k for k, _ in specs].index("uint")
    bad = list(inst["inputs"]); bad[first - 1 + j] = 1 << 70
    with pytest.raises(solver.Unsatisfied):
        solver.solve_program(flat, bad, lambda v: 1 << 200)
    with pytest.raises(RuntimeError, match="not satisfied"):          # the C port checks all constraints after the walk, like the GPU
        orc.solve(flat, orc.fr_mont(bad), lambda v: 1 << 200)


def test_c_prove_program_passes_the_toxic_waste_check():
    inst = circuit_instance(seed=7, **SMALL)
    flat, arr, sc = inst["flat"], inst["arr"], inst["sc"]
    r, s = 0x1111111111111111222222222222222 % R, 0x3333333333333333444444444444444 % R
    proof_bytes, secs = orc.groth16_prove_program(arr, flat, sc["infinity_a"], sc["infinity_b"], inst["inputs_mont"], r, s)
    proof = g16.proof_from_raw_bytes(proof_bytes)
    # the same solution from the Python solver, with the real commitment challenge
    box = {}
    def commit_fn(vals):
        box["pt"] = orc.g1_unpack(orc.g1_msm(arr["ck_basis"], orc.fr_mont(vals)))[0]
        return g16.commitment_challenge(box["pt"])
    w, a, b, c, info = solver.solve_program(flat, inst["inputs"], commit_fn)
    assert proof["Commitments"][0] == box["pt"]
    assert g16.check_in_exponent(inst["cs"], inst["tox"], proof, dict(w=w, a=a, b=b, c=c), r, s)
    # C hash_to_field == Python hash_to_field
    import bn254 as bn
    assert orc.commitment_challenge(bn.g1_raw_bytes(box["pt"])) == g16.commitment_challenge(box["pt"])
