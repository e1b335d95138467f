"""GPU parity: batch decoding of gnark-crypto point encodings (what pk.UnsafeReadFrom does, prover.go:342-346)
against the oracle's restatement of ecc/bn254/marshal.go."""
import numpy as np
import pytest

import bn254 as bn
import orc
import zkpor_b200 as zk
from bn254 import FP2, G1_GEN, G2_GEN, P, R, SplitMix64

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def points(n, seed):
    rng = SplitMix64(seed)
    ks = orc.ints_to_limbs([1 + rng.field(R - 1) for _ in range(n)])
    g1 = orc.g1_unpack(orc.g1_fixed_base(ks)); g2 = orc.g2_unpack(orc.g2_fixed_base(ks))
    g1[3] = None; g2[5] = None                    # infinity
    g1[7] = bn.pt_neg(g1[6]); g2[9] = bn.pt_neg(g2[8], FP2)   # both signs of the same x
    return g1, g2


def test_g1_decode_compressed_and_raw(ctx):
    g1, _ = points(300, 1)
    comp = np.frombuffer(b"".join(bn.g1_compressed_bytes(p) for p in g1), dtype=np.uint8).copy()
    raw = np.frombuffer(b"".join(bn.g1_raw_bytes(p) for p in g1), dtype=np.uint8).copy()
    assert orc.g1_unpack(ctx.g1_decode_batch(comp, len(g1), True)) == g1
    assert orc.g1_unpack(ctx.g1_decode_batch(raw, len(g1), False)) == g1


def test_g2_decode_compressed_and_raw(ctx):
    _, g2 = points(200, 2)
    comp = np.frombuffer(b"".join(bn.g2_compressed_bytes(p) for p in g2), dtype=np.uint8).copy()
    raw = np.frombuffer(b"".join(bn.g2_raw_bytes(p) for p in g2), dtype=np.uint8).copy()
    assert orc.g2_unpack(ctx.g2_decode_batch(comp, len(g2), True)) == g2
    assert orc.g2_unpack(ctx.g2_decode_batch(raw, len(g2), False)) == g2


def test_decode_rejects_invalid_points(ctx):
    g1, g2 = points(20, 3)
    comp = bytearray(b"".join(bn.g1_compressed_bytes(p) for p in g1))
    # find an x with no square root for x^3+3 and plant it at index 11
    x = 5
    while bn.fp_sqrt((x ** 3 + 3) % P) is not None:
        x += 1
    b = bytearray(x.to_bytes(32, "big")); b[0] |= 0x80
    comp[11 * 32:12 * 32] = b
    with pytest.raises(zk.ZkporError, match="index 11"):
        ctx.g1_decode_batch(np.frombuffer(bytes(comp), dtype=np.uint8).copy(), 20, True)
    # coordinate >= q
    bad = bytearray(b"".join(bn.g1_raw_bytes(p) for p in g1)); bad[64:96] = (P + 1).to_bytes(32, "big")
    with pytest.raises(zk.ZkporError, match="index 1"):
        ctx.g1_decode_batch(np.frombuffer(bytes(bad), dtype=np.uint8).copy(), 20, False)


def test_decoded_key_feeds_the_msm(ctx):
    """decode -> device-resident points -> MSM, without a host round trip of the points"""
    import torch
    g1, _ = points(500, 4)
    comp = np.frombuffer(b"".join(bn.g1_compressed_bytes(p) for p in g1), dtype=np.uint8).copy()
    dev = torch.empty(500 * 8, dtype=torch.int64, device="cuda")
    ctx.g1_decode_batch(comp, 500, True, out=dev)
    rng = SplitMix64(5)
    sc = orc.fr_mont([rng.field(R) for _ in range(500)])
    assert np.array_equal(ctx.msm_g1(dev, sc, 500), orc.g1_msm(orc.g1_pack(g1), sc))


# ------------------------------------------------------------------------------------------------ containers (a11 / f1)
def _small_instance():
    import groth16 as g16
    from helpers import synthetic_instance
    inst = synthetic_instance(200, 80, seed=41)
    pk_py, vk_py = g16.setup(inst["cs"], inst["tox"])
    return inst, pk_py, vk_py


def test_proof_containers_round_trip_vs_oracle(ctx):
    """Proof.ReadFrom / WriteTo / WriteRawTo (verifier/main.go:208-216, prover.go:201) against the oracle's restatement"""
    import containers as ct
    import groth16 as g16
    from helpers import make_pk, oracle_proof
    inst, pk_py, vk_py = _small_instance()
    raw = oracle_proof(inst, 11, 13)
    proof = g16.proof_from_raw_bytes(raw)
    comp = ct.proof_bytes(proof)
    assert len(raw) == 388 and len(comp) == 196
    assert zk.proof_decode(ctx, comp) == raw and zk.proof_decode(ctx, raw) == raw
    assert zk.proof_encode(ctx, raw, compressed=True) == comp and zk.proof_encode(ctx, comp, compressed=False) == raw
    # a proof made on the GPU goes through the same path
    pk = make_pk(zk, ctx, inst)
    m = orc.fr_mont
    got = pk.prove(m(inst["w"]), m(inst["a"]), m(inst["b"]), m(inst["c"]), len(inst["a"]), 11, 13)
    assert got == raw and zk.proof_decode(ctx, zk.proof_encode(ctx, got)) == raw
    pk.close()
    # damaged inputs fail loudly: truncated, and an x that is not on the curve
    with pytest.raises(zk.ZkporError, match="truncated"):
        zk.proof_decode(ctx, comp[:100])
    bad = bytearray(comp); bad[31] ^= 1
    try:
        out = zk.proof_decode(ctx, bytes(bad))
        assert out != raw                      # another valid x: decodes to a different point
    except zk.ZkporError as e:
        assert "invalid point" in str(e)


def test_vk_containers_vs_oracle(ctx):
    """vk.ReadFrom / WriteTo (prover.go:358-362, verifier/main.go:33-34, keygen/main.go:46-62): 524 bytes, and the decoded key verifies"""
    import containers as ct
    from helpers import oracle_proof
    inst, pk_py, vk_py = _small_instance()
    for raw in (False, True):
        b = ct.vk_bytes(vk_py, raw=raw)
        vk = zk.vk_decode(ctx, b)
        assert vk["bytes_consumed"] == len(b) and vk["n_commitments"] == 1 and len(vk["public_committed"]) == 0
        assert np.array_equal(vk["g1_k"], orc.g1_pack(vk_py["K"])) and np.array_equal(vk["g2_gamma"], orc.g2_pack([vk_py["gamma2"]]).ravel())
        assert np.array_equal(vk["g1_beta"], orc.g1_pack([vk_py["beta1"]]).ravel()) and np.array_equal(vk["g2_ped_g_root_sigma_neg"], orc.g2_pack([vk_py["ped_g_root_sigma_neg"]]).ravel())
        assert zk.vk_encode(ctx, vk, raw=raw) == b
    assert len(ct.vk_bytes(vk_py)) == 524
    # the decoded key is the one groth16.Verify needs
    vk = zk.vk_decode(ctx, ct.vk_bytes(vk_py))
    v = zk.VerifyingKey(alpha1=vk["g1_alpha"], beta2=vk["g2_beta"], gamma2=vk["g2_gamma"], delta2=vk["g2_delta"], K=vk["g1_k"], n_commitments=1,
                        public_committed=(), ped_g=vk["g2_ped_g"], ped_g_root_sigma_neg=vk["g2_ped_g_root_sigma_neg"])
    raw_proof = oracle_proof(inst, 5, 6)
    pub = orc.fr_mont(inst["w"][1:inst["cs"].nb_public])
    assert v.verify(ctx, raw_proof, pub)
    with pytest.raises(zk.ZkporError, match="truncated"):
        zk.vk_decode(ctx, ct.vk_bytes(vk_py)[:300])


def test_pk_containers_vs_oracle(ctx):
    """pk.ReadFrom / UnsafeReadFrom straight into HBM (prover.go:342-346) and pk.WriteTo / WriteRawTo back (keygen/main.go:46-62)"""
    import containers as ct
    from helpers import make_pk, oracle_proof
    inst, pk_py, vk_py = _small_instance()
    cs = inst["cs"]
    want = oracle_proof(inst, 21, 22)
    m = orc.fr_mont
    for raw in (False, True):
        b = ct.pk_bytes(pk_py, raw=raw)
        pk = zk.ProvingKey.read(ctx, b, cs.nb_public, cs.private_committed, cs.commitment_index)
        assert pk.bytes_consumed == len(b)
        assert pk.prove(m(inst["w"]), m(inst["a"]), m(inst["b"]), m(inst["c"]), len(inst["a"]), 21, 22) == want
        assert pk.write(raw=raw) == b and pk.write(raw=not raw) == ct.pk_bytes(pk_py, raw=not raw)
        pk.close()
    # a key uploaded from arrays writes the same file
    pk = make_pk(zk, ctx, inst)
    assert pk.write() == ct.pk_bytes(pk_py)
    pk.close()
    b = ct.pk_bytes(pk_py)
    with pytest.raises(zk.ZkporError, match="truncated|inconsistent"):
        zk.ProvingKey.read(ctx, b[:len(b) // 2], cs.nb_public, cs.private_committed, cs.commitment_index)
    with pytest.raises(zk.ZkporError, match="not a proving key"):
        zk.ProvingKey.read(ctx, b"\x00" * 400, cs.nb_public, cs.private_committed, cs.commitment_index)


def test_point_encode_batch_is_the_inverse_of_decode(ctx):
    g1, g2 = points(300, 7)
    p1, p2 = orc.g1_pack(g1), orc.g2_pack(g2)
    for comp in (1, 0):
        enc = np.zeros((len(g1), 32 if comp else 64), dtype=np.uint8)
        zk._check(zk.lib().zkpor_g1_encode_batch(ctx._h, zk._ptr(p1), len(g1), comp, zk._ptr(enc)))
        assert enc.tobytes() == b"".join((bn.g1_compressed_bytes if comp else bn.g1_raw_bytes)(p) for p in g1)
        assert np.array_equal(ctx.g1_decode_batch(enc, len(g1), bool(comp)), p1)
        enc = np.zeros((len(g2), 64 if comp else 128), dtype=np.uint8)
        zk._check(zk.lib().zkpor_g2_encode_batch(ctx._h, zk._ptr(p2), len(g2), comp, zk._ptr(enc)))
        assert enc.tobytes() == b"".join((bn.g2_compressed_bytes if comp else bn.g2_raw_bytes)(p) for p in g2)
        assert np.array_equal(ctx.g2_decode_batch(enc, len(g2), bool(comp)), p2)


def test_containers_against_the_committed_fixture(ctx):
    """tests/golden/containers.json: committed bytes of a seeded instance (made by tests/golden/make_containers.py from the oracle)"""
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "containers.json")))
    comp, raw = bytes.fromhex(fx["proof_compressed"]), bytes.fromhex(fx["proof_raw"])
    assert zk.proof_decode(ctx, comp) == raw and zk.proof_encode(ctx, raw, compressed=True) == comp
    vkb = bytes.fromhex(fx["vk_compressed"])
    vk = zk.vk_decode(ctx, vkb)
    assert vk["bytes_consumed"] == 524 and zk.vk_encode(ctx, vk) == vkb
