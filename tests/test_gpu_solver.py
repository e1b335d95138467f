"""GPU parity of the witness solver: zkpor_program_upload / zkpor_r1cs_solve / zkpor_groth16_prove_solve through
the C-ABI against the oracle's solver on BatchCreateUser-shaped circuits -- wires, a, b, c bit-exact, the mid-solve commitment, and
the proof bytes of the whole groth16.Prove (src/prover/prover/prover.go:269, hint registration :68)."""
import numpy as np
import pytest

import groth16 as g16
import orc
import zkpor_b200 as zk
from bn254 import R
from helpers import circuit_instance, make_pk

pytestmark = pytest.mark.gpu

SMALL = dict(users=3, assets_per_user=2, cex_assets=3, tiers=2, merkle_depth=2, chain_perms=3, limb_bits=8)
MEDIUM = dict(users=150, assets_per_user=2, cex_assets=5, tiers=3, merkle_depth=8, chain_perms=6, limb_bits=8)


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def small():
    return circuit_instance(seed=11, **SMALL)


def oracle_solution(inst):
    arr = inst["arr"]
    box = {}
    def commit_fn(vals_mont):
        box["pt"] = orc.g1_unpack(orc.g1_msm(arr["ck_basis"], vals_mont))[0]
        return g16.commitment_challenge(box["pt"])
    w, a, b, c = orc.solve(inst["flat"], inst["inputs_mont"], commit_fn)
    return w, a, b, c, box["pt"]


def test_solve_small_vs_oracle(ctx, small):
    prog = zk.Program(ctx, small["flat"])
    st = prog.stats()
    assert st["count_hints"] >= 2 and st["narrow_levels"] > 0          # both schedule regimes and the special hints are exercised
    pk = make_pk(zk, ctx, small)
    w, a, b, c, cm = prog.solve(small["inputs_mont"], pk)
    ow, oa, ob, oc, opt = oracle_solution(small)
    assert np.array_equal(w, ow)
    assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(c, oc)
    assert orc.g1_unpack(cm)[0] == opt
    prog.close(); pk.close()


def test_prove_solve_bytes_vs_oracle(ctx, small):
    prog = zk.Program(ctx, small["flat"])
    pk = make_pk(zk, ctx, small)
    r, s = 0xABCDEF0123456789ABCDEF % R, 0x123456789ABCDEF0123 % R
    got = pk.prove_solve(prog, small["inputs_mont"], r, s)
    want, _ = orc.groth16_prove_program(small["arr"], small["flat"], small["sc"]["infinity_a"], small["sc"]["infinity_b"], small["inputs_mont"], r, s)
    assert got == want and len(got) == 388
    # twice: the program and the key are reusable, the commitment is not carried over
    assert pk.prove_solve(prog, small["inputs_mont"], r, s) == want
    # the classic entry point on the oracle's solution gives the same bytes (commitment computed once vs. mid-solve)
    w, a, b, c, _ = oracle_solution(small)
    assert pk.prove(w, a, b, c, small["flat"]["n_constraints"], r, s) == want
    prog.close(); pk.close()


def test_unsatisfied_and_out_of_table_inputs_fail_loudly(ctx, small):
    flat = small["flat"]
    prog = zk.Program(ctx, flat)
    pk = make_pk(zk, ctx, small)
    first, n_s, count, specs = [x for x in flat["secret_layout"] if any(k == "uint" for k, _ in x[3])][0]
    j = [k for k, _ in specs].index("uint")
    bad = list(small["inputs"]); bad[first - 1 + j] = 1 << 70
    with pytest.raises(zk.ZkporError, match="not satisfied|outside|division"):
        prog.solve(orc.fr_mont(bad), pk)
    # a lookup index beyond its table
    first, n_s, count, specs = [x for x in flat["secret_layout"] if any(k == "below" for k, _ in x[3])][-1]
    j = [k for k, _ in specs].index("below")
    bad = list(small["inputs"]); bad[first - 1 + j] = 10 ** 6
    with pytest.raises(zk.ZkporError, match="outside|not satisfied"):
        prog.solve(orc.fr_mont(bad), pk)
    # and the context is still usable afterwards
    w, *_ = prog.solve(small["inputs_mont"], pk)
    assert np.array_equal(w, oracle_solution(small)[0])
    # a commitment hint without a key
    with pytest.raises(zk.ZkporError, match="commitment"):
        prog.solve(small["inputs_mont"], None)
    prog.close(); pk.close()


def test_invalid_programs_are_rejected(ctx, small):
    flat = dict(small["flat"])
    lv = np.array(flat["level_instr"]).copy()
    # swap the first and the last level's first instructions: the schedule no longer respects the dependencies
    lp = flat["level_ptr"]
    lv[int(lp[0])], lv[int(lp[-2])] = lv[int(lp[-2])], lv[int(lp[0])]
    flat["level_instr"] = lv
    with pytest.raises(zk.ZkporError, match="unsolved|never solved"):
        zk.Program(ctx, flat)
    flat = dict(small["flat"]); fn = np.array(flat["hint_fn"]).copy(); fn[0] = 99; flat["hint_fn"] = fn
    with pytest.raises(zk.ZkporError, match="unknown hint"):
        zk.Program(ctx, flat)


def test_solve_medium_wide_levels_vs_oracle(ctx):
    inst = circuit_instance(seed=21, **MEDIUM)
    flat = inst["flat"]
    assert flat["n_constraints"] > 200_000
    prog = zk.Program(ctx, flat)
    st = prog.stats()
    assert st["wide_levels"] > 100 and st["narrow_runs"] >= 1
    pk = make_pk(zk, ctx, inst)
    w, a, b, c, cm = prog.solve(inst["inputs_mont"], pk)
    ow, oa, ob, oc, opt = oracle_solution(inst)
    assert np.array_equal(w, ow) and np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(c, oc)
    r, s = 77, 99
    want, _ = orc.groth16_prove_program(inst["arr"], flat, inst["sc"]["infinity_a"], inst["sc"]["infinity_b"], inst["inputs_mont"], r, s)
    assert pk.prove_solve(prog, inst["inputs_mont"], r, s) == want
    prog.close(); pk.close()


def test_deferred_tail_gives_the_same_proof(ctx, monkeypatch):
    """The schedule's last run of narrow levels (the serial sponge of the CEX commitment) is started early on a side stream and its
    wires are added to the A / B / K multiplications afterwards; the proof bytes must not change.  ZKPOR_TAIL_MIN forces the
    deferral on a circuit whose tail is far shorter than the default threshold."""
    inst = circuit_instance(seed=31, **dict(MEDIUM, users=40, chain_perms=12))
    flat = inst["flat"]
    r, s = 1234567, 7654321
    want, _ = orc.groth16_prove_program(inst["arr"], flat, inst["sc"]["infinity_a"], inst["sc"]["infinity_b"], inst["inputs_mont"], r, s)
    monkeypatch.setenv("ZKPOR_TAIL_MIN", "0")
    plain = zk.Program(ctx, flat)
    assert plain.stats()["deferred_tail"]["levels"] == 0
    monkeypatch.setenv("ZKPOR_TAIL_MIN", "1")
    prog = zk.Program(ctx, flat)
    tail = prog.stats()["deferred_tail"]
    assert tail["levels"] > 100 and tail["wires"] > 100 and tail["starts_before_step"] > 0
    pk = make_pk(zk, ctx, inst)
    assert pk.prove_solve(plain, inst["inputs_mont"], r, s) == want
    for _ in range(3):                                   # the key's tail points are built on the first proof and reused
        assert pk.prove_solve(prog, inst["inputs_mont"], r, s) == want
    # the plain solve entry point runs the same program in place
    w, a, b, c, _ = prog.solve(inst["inputs_mont"], pk)
    ow, oa, ob, oc, _ = oracle_solution(inst)
    assert np.array_equal(w, ow) and np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(c, oc)
    # an unsatisfiable input still fails loudly (and leaves nothing running)
    first, n_s, count, specs = [x for x in flat["secret_layout"] if any(k == "uint" for k, _ in x[3])][0]
    j = [k for k, _ in specs].index("uint")
    bad = list(inst["inputs"]); bad[first - 1 + j] = 1 << 70
    with pytest.raises(zk.ZkporError, match="not satisfied|outside|division"):
        pk.prove_solve(prog, orc.fr_mont(bad), r, s)
    assert pk.prove_solve(prog, inst["inputs_mont"], r, s) == want
    prog.close(); plain.close(); pk.close()


@pytest.mark.parametrize("pipe", ["0", "1"])
def test_narrow_levels_pipelined_and_plain_agree_with_oracle(ctx, monkeypatch, pipe):
    """The serial sponge (narrow levels) can run software-pipelined (ZKPOR_NARROW_PIPE=1; opt-in): at upload the terms of each row that read a wire solved one
    level earlier are moved to the end of their lists, and the rest of a level is summed while the previous level finishes
    (k_solve_narrow_pipe).  Both forms must give the oracle's wires -- on a chain long enough to have full rounds (13 instructions per
    level, not pipelined), partial rounds (one instruction) and the chain's hand-over between permutations."""
    monkeypatch.setenv("ZKPOR_NARROW_PIPE", pipe)
    inst = circuit_instance(seed=41, **dict(MEDIUM, users=20, chain_perms=9))
    prog = zk.Program(ctx, inst["flat"])
    assert prog.stats()["narrow_levels"] > 1000
    pk = make_pk(zk, ctx, inst)
    ow, oa, ob, oc, _ = oracle_solution(inst)
    for _ in range(2):
        w, a, b, c, _ = prog.solve(inst["inputs_mont"], pk)
        assert np.array_equal(w, ow) and np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(c, oc)
    r, s = 5, 6
    want, _ = orc.groth16_prove_program(inst["arr"], inst["flat"], inst["sc"]["infinity_a"], inst["sc"]["infinity_b"], inst["inputs_mont"], r, s)
    assert pk.prove_solve(prog, inst["inputs_mont"], r, s) == want
    prog.close(); pk.close()
