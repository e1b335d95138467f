"""Validate the C oracle (oracle/c -> oracle/_build/liborc.so) against the Python big-int oracle through the
committed golden vectors (tests/golden/oracle_vectors.json, made by tests/golden/make_golden.py) and live."""
import json
import os

import numpy as np
import pytest

import bn254 as bn
import orc
import poseidon as ps
from bn254 import FP2, G1_GEN, G2_GEN, R, SplitMix64

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def vec():
    return json.load(open(os.path.join(GOLDEN, "oracle_vectors.json")))


def H(xs):
    return [int(x, 16) for x in xs]


def test_c_field_and_points():
    rng = SplitMix64(1)
    ks = [rng.field(R) for _ in range(20)] + [0, 1, 2, R - 1]
    g1 = orc.g1_unpack(orc.g1_fixed_base(orc.ints_to_limbs(ks)))
    g2 = orc.g2_unpack(orc.g2_fixed_base(orc.ints_to_limbs(ks)))
    for k, p1, p2 in zip(ks, g1, g2):
        assert p1 == bn.pt_mul(G1_GEN, k)
        assert p2 == bn.pt_mul(G2_GEN, k, FP2)


def test_c_poseidon_constants_and_hash(vec):
    for t in (2, 3, 5, 6, 10, 11, 13):
        rc, mds, rp = orc.poseidon_constants(t)
        prc, pmds = ps.constants(t)
        assert rp == ps.ROUNDS_P[t - 2]
        assert orc.fr_unmont(rc) == prc
        assert orc.fr_unmont(mds) == [x for row in pmds for x in row]
    for lane in (0, 1):
        orc.poseidon_set_out_lane(lane)
        for case in vec["poseidon"]:
            got = orc.fr_unmont(orc.poseidon_hash(orc.fr_mont(H(case["in"]))))[0]
            assert got == int(case[f"lane{lane}"], 16)
    orc.poseidon_set_out_lane(1)


def test_c_account_leaves(vec):
    for lane in (0, 1):
        orc.poseidon_set_out_lane(lane)
        for lv in vec["leaves"]:
            ids = np.frombuffer(bytes.fromhex(lv["id"]), dtype=np.uint8)
            tot = orc.be32_array([lv["equity"], lv["debt"], lv["collateral"]]).reshape(1, 96)
            got = orc.account_leaves(ids, tot, np.array(lv["flat"], dtype=np.uint64), lv["tier"])
            assert got.tobytes().hex() == lv[f"leaf_lane{lane}"]
    orc.poseidon_set_out_lane(1)


def test_c_merkle(vec):
    for tv in vec["merkle"]:
        orc.poseidon_set_out_lane(tv["lane"])
        cap, depth = tv["capacity"], tv["depth"]
        nil = bytes.fromhex(vec["nil_account_hash"][f"lane{tv['lane']}"])
        leaves = np.zeros((cap, 32), dtype=np.uint8)
        dirty = np.zeros((cap + 63) // 64, dtype=np.uint64)
        for k, v in tv["leaves"].items():
            leaves[int(k)] = np.frombuffer(bytes.fromhex(v), dtype=np.uint8)
            dirty[int(k) >> 6] |= np.uint64(1 << (int(k) & 63))
        nodes, root = orc.merkle_build(leaves, cap, depth, nil, dirty)
        assert root.hex() == tv["root"]
        pr = orc.merkle_proofs(leaves, nodes, cap, depth, nil, [9, 20], dirty)
        assert [pr[0, l].tobytes().hex() for l in range(depth)] == tv["proof_9"]
        assert [pr[1, l].tobytes().hex() for l in range(depth)] == tv["proof_20"]
    orc.poseidon_set_out_lane(1)


def test_c_msm(vec):
    m = vec["msm"]
    pk = orc.ints_to_limbs(H(m["point_scalars"]))
    sc = orc.fr_mont(H(m["scalars"]))
    g1 = orc.g1_fixed_base(pk); g2 = orc.g2_fixed_base(pk)
    assert orc.fp_unmont(orc.g1_msm(g1, sc)) == H(m["g1"])
    assert orc.fp_unmont(orc.g2_msm(g2, sc)) == H(m["g2"])
    # larger, threaded, against the discrete-log identity  sum s_i (k_i G) = (sum s_i k_i) G
    rng = SplitMix64(77)
    n = 5000
    ks = [1 + rng.field(R - 1) for _ in range(n)]
    ss = [rng.field(R) if i % 3 else rng.next() & 0xFFFF for i in range(n)]
    pts = orc.g1_fixed_base(orc.ints_to_limbs(ks))
    dot = sum(k * s for k, s in zip(ks, ss)) % R
    assert orc.g1_unpack(orc.g1_msm(pts, orc.fr_mont(ss)))[0] == bn.pt_mul(G1_GEN, dot)
    pts2 = orc.g2_fixed_base(orc.ints_to_limbs(ks[:600]))
    dot2 = sum(k * s for k, s in zip(ks[:600], ss[:600])) % R
    assert orc.g2_unpack(orc.g2_msm(pts2, orc.fr_mont(ss[:600])))[0] == bn.pt_mul(G2_GEN, dot2, FP2)


def test_c_ntt_and_compute_h(vec):
    t = vec["ntt"]
    v = orc.fr_mont(H(t["v"]))
    assert orc.fr_unmont(orc.ntt(v, 5, False, False, False)) == H(t["fft_dif"])
    assert orc.fr_unmont(orc.ntt(v, 5, False, True, True)) == H(t["fft_dit_coset"])
    assert orc.fr_unmont(orc.ntt(v, 5, True, False, False)) == H(t["ifft_dif"])
    assert orc.fr_unmont(orc.ntt(v, 5, True, False, True)) == H(t["ifft_dif_coset"])
    h = orc.compute_h(orc.fr_mont(H(t["a"])), orc.fr_mont(H(t["b"])), orc.fr_mont(H(t["c"])), 5)
    assert orc.fr_unmont(h) == H(t["h_bitrev"])


def pk_arrays_from_golden(g):
    s = g["pk_scalars"]
    lim = lambda xs: orc.ints_to_limbs(H(xs))
    arr = dict(A=orc.g1_fixed_base(lim(s["A"])), B1=orc.g1_fixed_base(lim(s["B"])), B2=orc.g2_fixed_base(lim(s["B"])),
               K=orc.g1_fixed_base(lim(s["K"])), Z=orc.g1_fixed_base(lim(s["Z"])),
               ck_basis=orc.g1_fixed_base(lim(s["ck"])), ck_basis_exp_sigma=orc.g1_fixed_base(lim(s["ck_sigma"])),
               alpha1=orc.g1_fixed_base(lim([s["alpha"]])), beta1=orc.g1_fixed_base(lim([s["beta"]])),
               delta1=orc.g1_fixed_base(lim([s["delta"]])), beta2=orc.g2_fixed_base(lim([s["beta"]])),
               delta2=orc.g2_fixed_base(lim([s["delta"]])), log_n=g["log_n"])
    return arr


def filtered_wires(g):
    w = H(g["wires"])
    wa = [w[i] for i in range(len(w)) if not g["infinity_a"][i]]
    wb = [w[i] for i in range(len(w)) if not g["infinity_b"][i]]
    drop = set(g["private_committed"]) | {g["commitment_index"]}
    wk = [w[i] for i in range(g["nb_public"], len(w)) if i not in drop]
    cm = [w[i] for i in g["private_committed"]]
    return wa, wb, wk, cm


def test_c_groth16_prove_bytes(vec):
    g = vec["groth16"]
    arr = pk_arrays_from_golden(g)
    wa, wb, wk, cm = filtered_wires(g)
    proof = orc.groth16_prove(arr, orc.fr_mont(wa), orc.fr_mont(wb), orc.fr_mont(wk), orc.fr_mont(cm),
                              orc.fr_mont(H(g["a"])), orc.fr_mont(H(g["b"])), orc.fr_mont(H(g["c"])),
                              int(g["r"], 16), int(g["s"], 16))
    assert proof.hex() == g["proof_raw"]


def test_c_reference_fixture_wide_sponge():
    """C oracle against the reference fixture: nil leaf of the fixture's circuit generation -> proof[15] (see test_oracle_kat.py)."""
    import base64
    fx = json.load(open(os.path.join(GOLDEN, "user_config_proof.json")))
    pr = [int.from_bytes(base64.b64decode(x), "big") for x in fx["Proof"]]
    orc.poseidon_set_out_lane(1)
    hv = lambda xs: orc.fr_unmont(orc.poseidon_hash(orc.fr_mont(xs)))[0]
    v = hv([0, 0, 0, 0, hv([0] * 584)])
    for _ in range(15):
        v = hv([v, v])
    assert v == pr[15]
