"""Full-size parity through size-independent properties (BASELINE.json configs 2 and 3): the key arrays are
(k0 + i*d)*G, so every MSM result -- and therefore every proof element -- has a discrete log the oracle can compute
with O(n) field additions; computeH is checked by the quotient identity at a random point."""
import os
import sys

import numpy as np
import pytest

import bn254 as bn
import groth16 as g16
import orc
import zkpor_b200 as zk
from bn254 import FP2, G1_GEN, G2_GEN, R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload shape + synthetic key builders shared with the benchmark)

pytestmark = pytest.mark.gpu
RINV = pow(1 << 256, -1, R)


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def host(t):
    return t.cpu().numpy().view(np.uint64).reshape(-1, 4)


def dlog_dot(scalars_mont_np, k0, d):
    """sum_i s_i * (k0 + i*d) for Montgomery-form scalars given as limb patterns"""
    s, t = orc.fr_index_sums(scalars_mont_np)
    return (k0 * s + d * t) * RINV % R


@pytest.mark.parametrize("log_n,kind", [(26, 0), (24, 2)])
def test_g1_msm_full_size_exact(ctx, log_n, kind):
    import torch
    n = 1 << log_n
    pts = bench.dev_buf(torch, n * 64); sc = bench.dev_buf(torch, n * 32)
    zk.synth_points_g1(ctx, 12345, 67891, n, pts)
    zk.synth_scalars(ctx, 7, n, kind, sc)
    got = orc.g1_unpack(ctx.msm_g1(pts, sc, n))[0]
    assert got == bn.pt_mul(G1_GEN, dlog_dot(host(sc), 12345, 67891))


def test_g2_msm_2pow24_exact(ctx):
    import torch
    n = 1 << 24
    pts = bench.dev_buf(torch, n * 128); sc = bench.dev_buf(torch, n * 32)
    zk.synth_points_g2(ctx, 222, 333, n, pts)
    zk.synth_scalars(ctx, 9, n, 0, sc)
    got = orc.g2_unpack(ctx.msm_g2(pts, sc, n))[0]
    assert got == bn.pt_mul(G2_GEN, dlog_dot(host(sc), 222, 333), FP2)


def test_compute_h_quotient_identity_2pow24(ctx):
    import torch
    log_n = 24; n = 1 << log_n; m = n - 12345
    a, b, c, h = (bench.dev_buf(torch, n * 32) for _ in range(4))
    zk.synth_scalars(ctx, 1, m, 0, a); zk.synth_scalars(ctx, 2, m, 0, b)
    ctx.fr_mul(a, b, c, m)
    ctx.compute_h(a, b, c, m, log_n, out=h)
    x0 = 0x1234567890ABCDEF0FEDCBA987654321 % R
    A = orc.eval_barycentric(host(a)[:m], log_n, x0); B = orc.eval_barycentric(host(b)[:m], log_n, x0); Cc = orc.eval_barycentric(host(c)[:m], log_n, x0)
    hx = orc.poly_eval_bitrev(host(h), log_n, x0)
    assert hx * (pow(x0, n, R) - 1) % R == (A * B - Cc) % R
    assert orc.fr_unmont(host(h)[n - 1])[0] == 0      # deg h <= n-2 (coefficient n-1 sits at bitrev(n-1) = n-1)


@pytest.mark.parametrize("log_n,scalars", [(22, "uniform"), (22, "witness")])
def test_prove_bench_shape_exact_by_discrete_logs(ctx, log_n, scalars):
    """The benchmark workload itself (smaller domain): the 388 proof bytes equal the bytes derived from the key's
    discrete logs, the wire vector and h -- Ar, Bs, Krs, Commitment, Pok all exact."""
    import torch
    sh = bench.shape_for(log_n)
    pk, arrays, _ = bench.build_key(torch, zk, ctx, sh)
    wires, a, b, c = bench.build_inputs(torch, zk, ctx, sh, scalars)
    r, s = 0xABCDEF123456789 % R, 0x987654321FEDCBA % R
    proof = pk.prove(wires, a, b, c, sh["n_constraints"], r, s)
    h = ctx.compute_h(a, b, c, sh["n_constraints"], log_n)
    w = host(wires)
    keep_k = np.ones(sh["W"], dtype=bool); keep_k[:sh["n_public"]] = False
    keep_k[sh["committed"].astype(np.int64)] = False; keep_k[sh["commitment_index"]] = False
    S, T = bench.SEEDS, bench.TOXIC
    dA = dlog_dot(np.ascontiguousarray(w[sh["inf_a"] == 0]), *S["A"]); dB = dlog_dot(np.ascontiguousarray(w[sh["inf_b"] == 0]), *S["B"])
    dK = dlog_dot(np.ascontiguousarray(w[keep_k]), *S["K"]); dZ = dlog_dot(np.ascontiguousarray(h[:sh["n_z"]]), *S["Z"])
    dC = dlog_dot(np.ascontiguousarray(w[sh["committed"].astype(np.int64)]), *S["CK"])
    ar = (dA + T["alpha"] + r * T["delta"]) % R
    bs = (dB + T["beta"] + s * T["delta"]) % R
    krs = (dK + dZ - r * s * T["delta"] + s * ar + r * bs) % R
    want = dict(Ar=bn.pt_mul(G1_GEN, ar), Bs=bn.pt_mul(G2_GEN, bs, FP2), Krs=bn.pt_mul(G1_GEN, krs),
                Commitments=[bn.pt_mul(G1_GEN, dC)], CommitmentPok=bn.pt_mul(G1_GEN, dC * T["sigma"] % R))
    assert proof == g16.proof_raw_bytes(want)
    pk.close()
