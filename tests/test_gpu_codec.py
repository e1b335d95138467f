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
