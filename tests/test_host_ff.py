"""CPU check of the product's field / curve headers (csrc/ff.cuh, csrc/ec.cuh) through their __host__ paths:
the limb schedule of Fe::mul (run against an emulated carry flag) and the XYZZ formulas are compared with the
oracle before any GPU time is spent.  The GPU tests re-check the same functions through the PTX path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import bn254 as bn
import orc
from bn254 import FP2, G1_GEN, G2_GEN, P, R, SplitMix64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "zkmerkle-proof-of-solvency_b200")


@pytest.fixture(scope="module")
def ht():
    so = os.path.join(PKG, "_build", "libzkpor_hosttest.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(PKG, "csrc", "hosttest.cpp"), "-o", so])
    return C.CDLL(so)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def edge_values(mod):
    return [0, 1, 2, mod - 1, mod - 2, (1 << 253), (1 << 32) - 1, (1 << 64) - 1, (1 << 224) + 5, mod >> 1]


@pytest.mark.parametrize("which", ["fp", "fr"])
def test_mont_mul_schedule(ht, which):
    mod = P if which == "fp" else R
    fn = ht.ht_fp_mul if which == "fp" else ht.ht_fr_mul
    rng = SplitMix64(5)
    vals = edge_values(mod) + [rng.field(mod) for _ in range(300)]
    rinv = pow(bn.MONT_R, -1, mod)
    for i in range(len(vals) - 1):
        a, b = vals[i], vals[(i * 7 + 3) % len(vals)]
        A = orc.ints_to_limbs([a]); B = orc.ints_to_limbs([b])
        o = np.zeros(4, dtype=np.uint64); oref = np.zeros(4, dtype=np.uint64)
        fn(p(A), p(B), p(o), p(oref))
        want = a * b * rinv % mod
        assert orc.limbs_to_ints(o)[0] == want == orc.limbs_to_ints(oref)[0]
        ok = np.zeros(4, dtype=np.uint64)
        (ht.ht_fp_mul_kara if which == "fp" else ht.ht_fr_mul_kara)(p(A), p(B), p(ok))
        assert orc.limbs_to_ints(ok)[0] == want, (hex(a), hex(b))


def test_fr_add_sub_neg_inv(ht):
    rng = SplitMix64(6)
    vals = edge_values(R) + [rng.field(R) for _ in range(100)]
    for i in range(len(vals)):
        a, b = vals[i], vals[(i * 5 + 1) % len(vals)]
        A = orc.ints_to_limbs([a]); B = orc.ints_to_limbs([b])
        oa, os_, on = (np.zeros(4, dtype=np.uint64) for _ in range(3))
        ht.ht_fr_addsub(p(A), p(B), p(oa), p(os_), p(on))
        assert orc.limbs_to_ints(oa)[0] == (a + b) % R
        assert orc.limbs_to_ints(os_)[0] == (a - b) % R
        assert orc.limbs_to_ints(on)[0] == (-a) % R
    for a in vals[:20]:
        o = np.zeros(4, dtype=np.uint64)
        ht.ht_fr_inv(p(orc.fr_mont([a])), p(o))
        assert orc.fr_unmont(o)[0] == (pow(a, -1, R) if a else 0)


@pytest.mark.parametrize("which", [0, 1])
def test_divstep_inverse_matches_fermat_and_python(ht, which):
    """Fe::inv (Bernstein-Yang divsteps on 30-bit limbs, ff.cuh) against Fermat in the same header and Python's pow."""
    mod = P if which == 0 else R
    rng = SplitMix64(17 + which)
    vals = edge_values(mod) + [3, (1 << 30) - 1, 1 << 30, 1 << 60, (mod >> 1) + 1] + [rng.field(mod) for _ in range(1500)]
    vals += [rng.field(1 << k) for k in range(1, 254)]
    arr = orc.ints_to_limbs(vals)
    assert ht.ht_inv_crosscheck(p(arr), len(vals), which) == 0
    fn = ht.ht_fp_inv if which == 0 else ht.ht_fr_inv
    rinv = pow(bn.MONT_R, -1, mod)
    for v in vals[:60]:
        o = np.zeros(4, dtype=np.uint64)
        fn(p(orc.ints_to_limbs([v])), p(o))
        plain = v * rinv % mod
        assert orc.limbs_to_ints(o)[0] == (pow(plain, -1, mod) * bn.MONT_R % mod if plain else 0)


def test_xyzz_accumulate_g1_g2(ht):
    rng = SplitMix64(9)
    ks = [1 + rng.field(R - 1) for _ in range(12)]
    # special cases: repeated point (doubling), opposite points (-> infinity), infinity input
    ks[3] = ks[2]; ks[6] = ks[5]
    neg = [0] * 12; neg[6] = 1; neg[8] = 1
    g1 = [bn.pt_mul(G1_GEN, k) for k in ks]; g2 = [bn.pt_mul(G2_GEN, k, FP2) for k in ks]
    g1[10] = None; g2[10] = None
    want = sum((-k if n else k) for i, (k, n) in enumerate(zip(ks, neg)) if i != 10) % R
    negs = np.array(neg, dtype=np.uint8)
    o1 = np.zeros(8, dtype=np.uint64); o2 = np.zeros(16, dtype=np.uint64)
    ht.ht_g1_accumulate(p(orc.g1_pack(g1)), p(negs), 12, p(o1))
    ht.ht_g2_accumulate(p(orc.g2_pack(g2)), p(negs), 12, p(o2))
    assert orc.g1_unpack(o1)[0] == bn.pt_mul(G1_GEN, want)
    assert orc.g2_unpack(o2)[0] == bn.pt_mul(G2_GEN, want, FP2)
    # P + (-P) first, then more: accumulator passes through infinity
    o1[:] = 0
    ht.ht_g1_accumulate(p(orc.g1_pack([g1[0], g1[0], g1[1]])), p(np.array([0, 1, 0], dtype=np.uint8)), 3, p(o1))
    assert orc.g1_unpack(o1)[0] == g1[1]


def test_xyzz_add_mul(ht):
    rng = SplitMix64(10)
    for _ in range(3):
        k1, k2, k3 = (1 + rng.field(R - 1) for _ in range(3))
        for pack, unpack, gen, F, fn, w in ((orc.g1_pack, orc.g1_unpack, G1_GEN, bn.FP, ht.ht_g1_add_mul, 8),
                                            (orc.g2_pack, orc.g2_unpack, G2_GEN, FP2, ht.ht_g2_add_mul, 16)):
            a, b = bn.pt_mul(gen, k1, F), bn.pt_mul(gen, k2, F)
            oa, om, od = (np.zeros(w, dtype=np.uint64) for _ in range(3))
            fn(p(pack([a])), p(pack([b])), p(orc.ints_to_limbs([k3])), p(oa), p(om), p(od))
            assert unpack(oa)[0] == bn.pt_mul(gen, k1 + k2, F)
            assert unpack(om)[0] == bn.pt_mul(gen, k1 * k3, F)
            assert unpack(od)[0] == bn.pt_mul(gen, 2 * k1, F)


def test_pairing_header_matches_oracle(ht):
    """csrc/pairing.cuh through its __host__ path: Fp12 tower arithmetic, the Miller loop value and e(P, Q), bit for bit
    against oracle/py/pairing.py (dense-polynomial Fp12, generic line functions)."""
    import pairing as pr

    def to_arr(t):
        return orc.fp_mont([v for c in t for b in c for v in b])

    def from_arr(a):
        v = orc.fp_unmont(a.reshape(12, 4))
        return pr.from_tower(tuple(tuple((v[6 * i + 2 * j], v[6 * i + 2 * j + 1]) for j in range(3)) for i in range(2)))

    rng = SplitMix64(4)
    x = tuple(rng.field(P) for _ in range(12)); y = tuple(rng.field(P) for _ in range(12))
    om, osq, oi = (np.zeros(48, dtype=np.uint64) for _ in range(3))
    ht.ht_fp12_ops(p(to_arr(pr.to_tower(x))), p(to_arr(pr.to_tower(y))), p(om), p(osq), p(oi))
    assert from_arr(om) == pr.f12_mul(x, y) and from_arr(osq) == pr.f12_sqr(x) and from_arr(oi) == pr.f12_inv(x)
    a, b = rng.field(R), rng.field(R)
    pa = bn.pt_mul(G1_GEN, a); qb = bn.pt_mul(G2_GEN, b, FP2)
    mil, gt = np.zeros(48, dtype=np.uint64), np.zeros(48, dtype=np.uint64)
    ht.ht_pairing(p(orc.g1_pack([pa])), p(orc.g2_pack([qb])), p(mil), p(gt))
    want_m = pr.miller_loop(pr.untwist(qb), pr.embed_g1(pa))
    assert from_arr(mil) == want_m
    assert from_arr(gt) == pr.final_exponentiation(want_m)
