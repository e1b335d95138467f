"""GPU parity: zkpor_msm_g1 / zkpor_msm_g2 through the C-ABI against the oracle (bit-exact affine output).
Replaces gnark-crypto MultiExp inside groth16.Prove (src/prover/prover/prover.go:269)."""
import numpy as np
import pytest

import bn254 as bn
import orc
import zkpor_b200 as zk
from bn254 import FP2, G1_GEN, G2_GEN, R, SplitMix64
from helpers import H, g1_points, g2_points, golden, rand_scalars, rand_scalars_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def test_golden_msm_vectors(ctx):
    m = golden()["msm"]
    pk = orc.ints_to_limbs(H(m["point_scalars"]))
    sc = orc.fr_mont(H(m["scalars"]))
    assert orc.fp_unmont(ctx.msm_g1(orc.g1_fixed_base(pk), sc, len(sc))) == H(m["g1"])
    assert orc.fp_unmont(ctx.msm_g2(orc.g2_fixed_base(pk), sc, len(sc))) == H(m["g2"])


@pytest.mark.parametrize("n", [1, 2, 31, 257, 4096, 20000])
@pytest.mark.parametrize("kind", ["uniform", "witness"])
def test_g1_vs_oracle(ctx, n, kind):
    pts = g1_points(n, 100 + n)
    sc = orc.fr_mont(rand_scalars(n, 200 + n, kind))
    assert np.array_equal(ctx.msm_g1(pts, sc, n), orc.g1_msm(pts, sc))


@pytest.mark.parametrize("n", [1, 3, 500, 6000])
@pytest.mark.parametrize("kind", ["uniform", "witness"])
def test_g2_vs_oracle(ctx, n, kind):
    pts = g2_points(n, 300 + n)
    sc = orc.fr_mont(rand_scalars(n, 400 + n, kind))
    assert np.array_equal(ctx.msm_g2(pts, sc, n), orc.g2_msm(pts, sc))


def test_edge_cases(ctx):
    n = 64
    pts = g1_points(n, 7)
    # all-zero scalars -> infinity; empty input -> infinity
    assert not ctx.msm_g1(pts, orc.fr_mont([0] * n), n).any()
    assert not ctx.msm_g1(pts, orc.fr_mont([0] * n), 0).any()
    # repeated points in one bucket (doubling inside the accumulator), P and -P (passes through infinity), infinity inputs
    pts[1] = pts[0]; pts[2] = pts[0]
    negp = orc.g1_unpack(pts[3:4])[0]; pts[4] = orc.g1_pack([bn.pt_neg(negp)])[0]
    pts[5] = 0
    ss = [5, 5, 5, 9, 9, 1234] + rand_scalars(n - 6, 8)
    ss[10] = R - 1; ss[11] = 1; ss[12] = (1 << 253) + 12345; ss[13] = (1 << 16); ss[14] = (1 << 15)
    sc = orc.fr_mont(ss)
    assert np.array_equal(ctx.msm_g1(pts, sc, n), orc.g1_msm(pts, sc))
    # canonical (non-Montgomery) scalars
    assert np.array_equal(ctx.msm_g1(pts, orc.ints_to_limbs(ss), n, zk.ZKPOR_SCALARS_PLAIN), orc.g1_msm(pts, sc))
    # same scalar everywhere: one bucket per window gets every point
    sc2 = orc.fr_mont([0xDEADBEEFCAFEF00D] * n)
    assert np.array_equal(ctx.msm_g1(pts, sc2, n), orc.g1_msm(pts, sc2))


@pytest.mark.parametrize("rounds", [0, 1, 2, 3, -1])
def test_affine_tree_special_cases(ctx, rounds):
    """Bucket lists full of repeated points, opposite points and infinities: every pair class of the batched-affine
    tree (doubling, cancellation to infinity, infinity operands, odd element carried over) at every tree depth, and
    the extended-Jacobian-only path (rounds = 0), all bit-exact against the oracle."""
    n = 6000
    rng = SplitMix64(77)
    base = g1_points(6, 70)
    neg = orc.g1_pack([bn.pt_neg(q) for q in orc.g1_unpack(base)])
    pool = np.concatenate([base, neg, np.zeros((1, 8), dtype=np.uint64)])          # 6 points, their negatives, infinity
    pts = np.ascontiguousarray(pool[[rng.next() % len(pool) for _ in range(n)]])
    svals = [3, 3, 3, R - 3, 0x10001, (1 << 200) + 17, R - 1, 1]                       # few distinct scalars -> few, long bucket lists
    ss = [svals[rng.next() % len(svals)] for _ in range(n)]
    sc = orc.fr_mont(ss)
    want = orc.g1_msm(pts, sc)
    ctx.set_affine_rounds(rounds)
    try:
        assert np.array_equal(ctx.msm_g1(pts, sc, n), want)
        # mixed with uniform scalars (lists of ordinary length next to the degenerate ones)
        ss2 = [ss[i] if i % 3 else rng.field(R) for i in range(n)]
        sc2 = orc.fr_mont(ss2)
        assert np.array_equal(ctx.msm_g1(pts, sc2, n), orc.g1_msm(pts, sc2))
        b2 = g2_points(5, 71)
        neg2 = orc.g2_pack([bn.pt_neg(q, FP2) for q in orc.g2_unpack(b2)])
        pool2 = np.concatenate([b2, neg2, np.zeros((1, 16), dtype=np.uint64)])
        n2 = 3000
        p2 = np.ascontiguousarray(pool2[[rng.next() % len(pool2) for _ in range(n2)]])
        s2 = orc.fr_mont([svals[rng.next() % len(svals)] if i % 4 else rng.field(R) for i in range(n2)])
        assert np.array_equal(ctx.msm_g2(p2, s2, n2), orc.g2_msm(p2, s2))
    finally:
        ctx.set_affine_rounds(0)


@pytest.mark.parametrize("rounds", [0, 2, -1])
def test_affine_rounds_agree_at_2pow18(ctx, rounds):
    n = 1 << 18
    pts = g1_points(2048, 81); pts = np.ascontiguousarray(np.tile(pts, (n // 2048, 1)))
    sc = rand_scalars_np(n, 82)
    ctx.set_affine_rounds(rounds)
    try:
        assert np.array_equal(ctx.msm_g1(pts, sc, n), orc.g1_msm(pts, sc))
    finally:
        ctx.set_affine_rounds(0)


def test_heavy_buckets_skewed_witness(ctx):
    """witness-like scalars at a size where one bucket (the value 1) holds far more than the heavy threshold"""
    n = 150000
    pts = g1_points(4096, 41); pts = np.ascontiguousarray(np.tile(pts, (n // 4096 + 1, 1))[:n])
    sc = orc.fr_mont(rand_scalars(n, 42, "witness"))
    assert np.array_equal(ctx.msm_g1(pts, sc, n), orc.g1_msm(pts, sc))
    n2 = 60000
    p2 = g2_points(1024, 43); p2 = np.ascontiguousarray(np.tile(p2, (n2 // 1024 + 1, 1))[:n2])
    s2 = orc.fr_mont([1] * 30000 + rand_scalars(n2 - 30000, 44, "witness"))
    assert np.array_equal(ctx.msm_g2(p2, s2, n2), orc.g2_msm(p2, s2))


def test_partitioned_sort_skewed_and_plain(ctx):
    """n >= 2^18 takes the two-level (partitioned) counting sort: witness-like scalars (one bucket with ~15% of all terms, most
    partitions nearly empty), canonical-integer input, and the largest canonical scalars (top window fully populated)."""
    n = 300000
    pts = g1_points(4096, 51); pts = np.ascontiguousarray(np.tile(pts, (n // 4096 + 1, 1))[:n])
    ss = rand_scalars(n, 52, "witness")
    ss[7] = R - 1; ss[8] = R - 2; ss[9] = (1 << 253) + 1; ss[10] = (1 << 240) - 1; ss[11] = 1 << 240
    sc = orc.fr_mont(ss)
    want = orc.g1_msm(pts, sc)
    assert np.array_equal(ctx.msm_g1(pts, sc, n), want)
    assert np.array_equal(ctx.msm_g1(pts, orc.ints_to_limbs(ss), n, zk.ZKPOR_SCALARS_PLAIN), want)
    n2 = 1 << 18
    p2 = g2_points(512, 53); p2 = np.ascontiguousarray(np.tile(p2, (n2 // 512, 1)))
    s2 = rand_scalars_np(n2, 54)
    assert np.array_equal(ctx.msm_g2(p2, s2, n2), orc.g2_msm(p2, s2))


def test_device_pointer_inputs(ctx):
    import torch
    n = 5000
    pts = g1_points(n, 21); sc = orc.fr_mont(rand_scalars(n, 22))
    tp = torch.from_numpy(pts.view(np.int64)).cuda(); ts = torch.from_numpy(sc.view(np.int64)).cuda()
    torch.cuda.synchronize()
    assert np.array_equal(ctx.msm_g1(tp, ts, n), orc.g1_msm(pts, sc))
    t = ctx.last_timings()
    assert t["h2d"] < 5.0 and t["accumulate"] > 0


def test_partials_combine_like_multi_gpu(ctx):
    """point-chunk sharding: sum of per-chunk partial results == whole MSM (what 8 ranks + one all-gather do)."""
    n = 9000
    pts = g1_points(n, 31); sc = orc.fr_mont(rand_scalars(n, 32))
    cuts = [0, 1000, 1001, 5000, n]
    parts = np.stack([ctx.msm_g1_partial(pts[a:b].copy(), sc[a:b].copy(), b - a) for a, b in zip(cuts, cuts[1:])])
    assert np.array_equal(zk.g1_sum_partials(parts), orc.g1_msm(pts, sc))
    p2 = g2_points(700, 33); s2 = orc.fr_mont(rand_scalars(700, 34))
    parts2 = np.stack([ctx.msm_g2_partial(p2[a:b].copy(), s2[a:b].copy(), b - a) for a, b in ((0, 300), (300, 700))])
    assert np.array_equal(zk.g2_sum_partials(parts2), orc.g2_msm(p2, s2))


def test_g1_2pow20_vs_oracle_and_linearity(ctx):
    """BASELINE config 2 smallest size, bit-exact against the CPU oracle; then the size-independent property
    MSM(s) + MSM(t) = MSM(s + t) with resident points."""
    import torch
    n = 1 << 20
    rng = SplitMix64(99)
    base = orc.g1_fixed_base(orc.ints_to_limbs([1 + rng.field(R - 1) for _ in range(4096)]))
    pts = np.ascontiguousarray(np.tile(base, (n // 4096, 1)))          # 2^20 points (repeats are legal inputs)
    s = rand_scalars_np(n, 1); t = rand_scalars_np(n, 2)
    got = ctx.msm_g1(pts, s, n)
    assert np.array_equal(got, orc.g1_msm(pts, s))
    dp = torch.from_numpy(pts.view(np.int64)).cuda()
    # s + t mod r with Python ints in object arrays
    S = np.array(orc.limbs_to_ints(s), dtype=object); T = np.array(orc.limbs_to_ints(t), dtype=object)
    st = orc.ints_to_limbs(list((S + T) % R))
    a = orc.g1_unpack(got)[0]; b = orc.g1_unpack(ctx.msm_g1(dp, t, n))[0]; c = orc.g1_unpack(ctx.msm_g1(dp, st, n))[0]
    assert bn.pt_add(a, b) == c
