"""Regenerates tests/golden/*.  Run in the build container (needs /root/reference for the fixture copy):
    python tests/golden/make_golden.py
* user_config_proof.json : data copied from /root/reference/src/verifier/config/user_config.json (the reference's
  only byte-level fixture for the hot path; Root + 28 siblings).
* oracle_vectors.json    : outputs of oracle/py (Python big-int ground truth) on seeded inputs; the C oracle and the
  CUDA product are compared with these so that the GPU box needs neither /root/reference nor long Python runs."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle", "py"))

import bn254 as bn  # noqa: E402
import groth16 as g16  # noqa: E402
import merkle  # noqa: E402
import ntt  # noqa: E402
import poseidon as ps  # noqa: E402
from bn254 import FP2, G1_GEN, G2_GEN, R, SplitMix64  # noqa: E402


def main():
    ref = "/root/reference/src/verifier/config/user_config.json"
    if os.path.exists(ref):
        fx = json.load(open(ref))
        json.dump({"source": "src/verifier/config/user_config.json", "AccountIndex": fx["AccountIndex"],
                   "Root": fx["Root"], "Proof": fx["Proof"]}, open(os.path.join(HERE, "user_config_proof.json"), "w"), indent=1)
    out = {}
    rng = SplitMix64(0xB200)
    # Poseidon: every width the circuit uses, both output lanes
    pv = []
    for n_in in (1, 2, 4, 5, 9, 10, 12, 13, 24, 29, 100):
        ins = [rng.field(R) for _ in range(n_in)]
        pv.append({"in": [hex(x) for x in ins], "lane0": hex(ps.poseidon(ins, 0)), "lane1": hex(ps.poseidon(ins, 1))})
    out["poseidon"] = pv
    # account leaves
    lv = []
    for n_assets, tier in ((0, 50), (3, 50), (50, 50), (51, 500), (120, 500)):
        idxs = sorted({rng.next() % 500 for _ in range(n_assets * 3)})[:n_assets] if n_assets else []
        if tier == 50 and n_assets == 50:
            idxs = list(range(0, 100, 2))
        assets = [(i, rng.next() >> 20, rng.next() >> 24, rng.next() >> 30, rng.next() >> 30, rng.next() >> 30) for i in idxs]
        acc_id = rng.field(R).to_bytes(32, "big")
        eq, debt, col = rng.next() << 10, rng.next(), 0 if n_assets == 0 else rng.next()
        lv.append({"id": acc_id.hex(), "equity": eq, "debt": debt, "collateral": col, "assets": assets, "tier": merkle.assets_count_tier(len(assets)),
                   "flat": merkle.padding_account_assets(assets),
                   "leaf_lane1": merkle.account_leaf(acc_id, eq, debt, col, assets, 1).hex(),
                   "leaf_lane0": merkle.account_leaf(acc_id, eq, debt, col, assets, 0).hex()})
    out["leaves"] = lv
    out["nil_account_hash"] = {"lane0": merkle.nil_account_hash(0).hex(), "lane1": merkle.nil_account_hash(1).hex()}
    # Merkle tree, depth 10, ragged capacity, sparse sets
    tv = []
    for lane in (0, 1):
        nil = merkle.nil_account_hash(lane)
        t = merkle.FixedDepthMerkleTree(10, nil, 37, lane)
        leaves = {}
        for k in (0, 1, 2, 3, 4, 9, 17, 36):
            leaves[k] = ps.poseidon_bytes([bytes([k + 1]), b"\x05"], lane)
            t.set(k, leaves[k])
        t.build()
        tv.append({"lane": lane, "depth": 10, "capacity": 37, "leaves": {str(k): v.hex() for k, v in leaves.items()},
                   "root": t.root.hex(), "proof_9": [x.hex() for x in t.get_proof(9)], "proof_20": [x.hex() for x in t.get_proof(20)]})
    out["merkle"] = tv
    # MSM G1/G2 (mathematical definition), n = 33
    pts_k = [1 + rng.field(R - 1) for _ in range(33)]
    sc = [rng.field(R) for _ in range(33)]
    sc[3] = 0; sc[4] = 1; sc[5] = R - 1; sc[6] = 2 ** 16; sc[7] = sc[8]
    g1pts = [bn.pt_mul(G1_GEN, k) for k in pts_k]
    g2pts = [bn.pt_mul(G2_GEN, k, FP2) for k in pts_k]
    dot = sum(k * s for k, s in zip(pts_k, sc)) % R
    r1 = bn.msm_naive(g1pts, sc); r2 = bn.msm_naive(g2pts, sc, FP2)
    assert r1 == bn.pt_mul(G1_GEN, dot) and r2 == bn.pt_mul(G2_GEN, dot, FP2)
    out["msm"] = {"point_scalars": [hex(k) for k in pts_k], "scalars": [hex(s) for s in sc],
                  "g1": [hex(r1[0]), hex(r1[1])], "g2": [hex(r2[0][0]), hex(r2[0][1]), hex(r2[1][0]), hex(r2[1][1])]}
    # NTT / computeH, n = 32 with 27 constraints
    d = ntt.Domain(27)
    a = [rng.field(R) for _ in range(27)]; b = [rng.field(R) for _ in range(27)]
    c = [x * y % R for x, y in zip(a, b)]
    v = [rng.field(R) for _ in range(32)]
    out["ntt"] = {"logn": 5, "v": [hex(x) for x in v],
                  "fft_dif": [hex(x) for x in ntt.fft(d, v, "DIF")],
                  "fft_dit_coset": [hex(x) for x in ntt.fft(d, v, "DIT", coset=True)],
                  "ifft_dif": [hex(x) for x in ntt.fft_inverse(d, v, "DIF")],
                  "ifft_dif_coset": [hex(x) for x in ntt.fft_inverse(d, v, "DIF", coset=True)],
                  "a": [hex(x) for x in a], "b": [hex(x) for x in b], "c": [hex(x) for x in c],
                  "h_bitrev": [hex(x) for x in ntt.compute_h(a, b, c, d)]}
    # Groth16: synthetic R1CS, 61 constraints -> n = 64
    cs = g16.synth_r1cs(61, 9, 0xB200)
    tox = g16.toxic_from_seed(3)
    pk, vk = g16.setup(cs, tox)
    pub, sec = g16.synth_inputs(cs, 5)
    r, s = rng.field(R), rng.field(R)
    proof, aux = g16.prove(cs, pk, pub, sec, r, s)
    assert g16.check_in_exponent(cs, tox, proof, aux, r, s)
    hx = lambda xs: [hex(x) for x in xs]
    out["groth16"] = {
        "n_constraints": 61, "nb_secret": 9, "seed": 0xB200, "toxic_seed": 3, "input_seed": 5, "r": hex(r), "s": hex(s),
        "log_n": 6, "pk_scalars": {"A": hx(pk["A_s"]), "B": hx(pk["B_s"]), "K": hx(pk["K_s"]), "Z": hx(pk["Z_s"]),
                                   "ck": hx(pk["ck_basis_s"]), "ck_sigma": hx([x * tox["sigma"] % R for x in pk["ck_basis_s"]]),
                                   "alpha": hex(tox["alpha"]), "beta": hex(tox["beta"]), "delta": hex(tox["delta"])},
        "infinity_a": [int(x) for x in pk["infinity_a"]], "infinity_b": [int(x) for x in pk["infinity_b"]],
        "nb_public": cs.nb_public, "commitment_index": cs.commitment_index, "private_committed": cs.private_committed,
        "wires": hx(aux["w"]), "a": hx(aux["a"]), "b": hx(aux["b"]), "c": hx(aux["c"]), "h": hx(aux["h"]),
        "proof_raw": g16.proof_raw_bytes(proof).hex()}
    json.dump(out, open(os.path.join(HERE, "oracle_vectors.json"), "w"))
    print("wrote golden vectors")


if __name__ == "__main__":
    main()
