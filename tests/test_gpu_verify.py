"""GPU parity: pairing product and groth16.Verify through the C-ABI against the oracle.
Replaces gnark-crypto bn254.Pair / PairingCheck and gnark groth16.Verify (src/prover/prover/prover.go:276,
src/verifier/main.go:284).  GT values are compared bit for bit with oracle/py/pairing.py (exact exponent (q^12-1)/r);
Verify is compared as accept / reject on GPU-made proofs, tampered proofs and wrong public inputs."""
import numpy as np
import pytest

import bn254 as bn
import groth16 as g16
import orc
import pairing as pr
import zkpor_b200 as zk
from bn254 import FP2, G1_GEN, G2_GEN, R, SplitMix64
from helpers import make_pk, synthetic_instance, vk_arrays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def gt_from_arr(a):
    v = orc.fp_unmont(np.ascontiguousarray(a).reshape(12, 4))
    return pr.from_tower(tuple(tuple((v[6 * i + 2 * j], v[6 * i + 2 * j + 1]) for j in range(3)) for i in range(2)))


def test_pairing_product_vs_oracle(ctx):
    rng = SplitMix64(11)
    ks = [(1 + rng.field(R - 1), 1 + rng.field(R - 1)) for _ in range(3)]
    ps = [bn.pt_mul(G1_GEN, a) for a, _ in ks]; qs = [bn.pt_mul(G2_GEN, b, FP2) for _, b in ks]
    P, Q = orc.g1_pack(ps), orc.g2_pack(qs)
    # single pairings, bit-exact GT
    for i in range(3):
        assert gt_from_arr(zk.pairing_product(ctx, P[i:i + 1].copy(), Q[i:i + 1].copy(), 1)) == pr.pairing(ps[i], qs[i])
    # product of three with one shared final exponentiation
    assert gt_from_arr(zk.pairing_product(ctx, P, Q, 3)) == pr.pairing_product(list(zip(ps, qs)))
    # bilinearity against the generator pairing: e(aP, bQ) = e(P, Q)^(ab)
    e = gt_from_arr(zk.pairing_product(ctx, orc.g1_pack([G1_GEN]), orc.g2_pack([G2_GEN]), 1))
    assert gt_from_arr(zk.pairing_product(ctx, P[:1].copy(), Q[:1].copy(), 1)) == pr.f12_pow(e, ks[0][0] * ks[0][1] % R)
    # infinity on either side contributes 1; empty product is 1
    one = pr.F12_ONE
    assert gt_from_arr(zk.pairing_product(ctx, np.zeros((1, 8), dtype=np.uint64), Q[:1].copy(), 1)) == one
    assert gt_from_arr(zk.pairing_product(ctx, P[:1].copy(), np.zeros((1, 16), dtype=np.uint64), 1)) == one
    assert gt_from_arr(zk.pairing_product(ctx, P, Q, 0)) == one


def test_pairing_check(ctx):
    a = 0xC0FFEE
    P = orc.g1_pack([bn.pt_mul(G1_GEN, a), bn.pt_neg(G1_GEN)])
    Q = orc.g2_pack([G2_GEN, bn.pt_mul(G2_GEN, a, FP2)])
    assert zk.pairing_check(ctx, P, Q, 2)                      # e(aP, Q) e(-P, aQ) = 1
    Q2 = orc.g2_pack([G2_GEN, bn.pt_mul(G2_GEN, a + 1, FP2)])
    assert not zk.pairing_check(ctx, P, Q2, 2)
    import torch
    tp = torch.from_numpy(P.view(np.int64)).cuda(); tq = torch.from_numpy(Q.view(np.int64)).cuda()
    assert zk.pairing_check(ctx, tp, tq, 2)                    # device-resident inputs


@pytest.fixture(scope="module")
def proved(ctx):
    """one synthetic circuit, three GPU-made proofs with different witnesses and blinding"""
    inst = synthetic_instance(300, 24, seed=77)
    vk = zk.VerifyingKey(**vk_arrays(inst))
    pk = make_pk(zk, ctx, inst)
    m = orc.fr_mont
    proofs, pubs = [], []
    rng = SplitMix64(78)
    for k in range(3):
        cur = inst if k == 0 else synthetic_instance(300, 24, seed=77, input_seed=500 + k)
        proofs.append(pk.prove(m(cur["w"]), m(cur["a"]), m(cur["b"]), m(cur["c"]), 300, rng.field(R), rng.field(R)))
        pubs.append(cur["w"][1:cur["cs"].nb_public])
    pk.close()
    return inst, vk, proofs, pubs


def test_verify_accepts_gpu_proofs_and_matches_oracle(ctx, proved):
    inst, vk, proofs, pubs = proved
    ovk = inst["vk"]
    for pf, pub in zip(proofs, pubs):
        assert vk.verify(ctx, pf, orc.fr_mont(pub))
    assert g16.verify(ovk, g16.proof_from_raw_bytes(proofs[0]), pubs[0])          # the oracle agrees (CPU pairing)


def test_verify_rejects(ctx, proved):
    inst, vk, proofs, pubs = proved
    pf, pub = proofs[0], pubs[0]
    good = g16.proof_from_raw_bytes(pf)
    def enc(**kw):
        d = dict(good); d.update(kw); return g16.proof_raw_bytes(d)
    assert not vk.verify(ctx, enc(Krs=bn.pt_add(good["Krs"], G1_GEN)), orc.fr_mont(pub))
    assert not vk.verify(ctx, enc(Ar=bn.pt_mul(good["Ar"], 2)), orc.fr_mont(pub))
    assert not vk.verify(ctx, enc(Bs=bn.pt_add(good["Bs"], G2_GEN, FP2)), orc.fr_mont(pub))
    assert not vk.verify(ctx, enc(CommitmentPok=bn.pt_add(good["CommitmentPok"], G1_GEN)), orc.fr_mont(pub))     # Pedersen check
    assert not vk.verify(ctx, enc(Commitments=[bn.pt_add(good["Commitments"][0], G1_GEN)]), orc.fr_mont(pub))    # changes the challenge
    assert not vk.verify(ctx, pf, orc.fr_mont([(pub[0] + 1) % R] + list(pub[1:])))                                # wrong public input
    assert not vk.verify(ctx, proofs[1], orc.fr_mont(pub))                                                        # proof of another statement
    # malformed encodings are errors, not verdicts
    bad = bytearray(pf); bad[5] ^= 0xFF
    with pytest.raises(zk.ZkporError):
        vk.verify(ctx, bytes(bad), orc.fr_mont(pub))                 # Ar no longer on the curve
    with pytest.raises(zk.ZkporError):
        vk.verify(ctx, pf[:-1], orc.fr_mont(pub))
    with pytest.raises(zk.ZkporError):
        vk.verify(ctx, pf, orc.fr_mont(list(pub) + [1]))              # witness length != len(vk.K) - 2


def test_verify_batch(ctx, proved):
    inst, vk, proofs, pubs = proved
    pw = np.stack([orc.fr_mont(p) for p in pubs])
    assert vk.verify_batch(ctx, proofs, pw, seed=bytes(range(32)))
    assert vk.verify_batch(ctx, proofs[:1], pw[:1])
    good = g16.proof_from_raw_bytes(proofs[1])
    bad = dict(good); bad["Krs"] = bn.pt_add(good["Krs"], G1_GEN)
    assert not vk.verify_batch(ctx, [proofs[0], g16.proof_raw_bytes(bad), proofs[2]], pw)
    bad = dict(good); bad["CommitmentPok"] = bn.pt_add(good["CommitmentPok"], G1_GEN)
    assert not vk.verify_batch(ctx, [proofs[0], g16.proof_raw_bytes(bad), proofs[2]], pw)
    assert not vk.verify_batch(ctx, proofs, pw[::-1].copy())          # witnesses attached to the wrong proofs
