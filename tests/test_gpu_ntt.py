"""GPU parity: zkpor_ntt / zkpor_compute_h against the oracle (gnark-crypto fft.Domain conventions, gnark computeH)."""
import numpy as np
import pytest

import ntt as pyntt
import orc
import zkpor_b200 as zk
from bn254 import R, SplitMix64
from helpers import H, golden, rand_scalars, rand_scalars_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def test_golden_ntt_vectors(ctx):
    t = golden()["ntt"]
    v = orc.fr_mont(H(t["v"]))
    assert orc.fr_unmont(ctx.ntt(v.copy(), 5, False, False, False)) == H(t["fft_dif"])
    assert orc.fr_unmont(ctx.ntt(v.copy(), 5, False, True, True)) == H(t["fft_dit_coset"])
    assert orc.fr_unmont(ctx.ntt(v.copy(), 5, True, False, False)) == H(t["ifft_dif"])
    assert orc.fr_unmont(ctx.ntt(v.copy(), 5, True, False, True)) == H(t["ifft_dif_coset"])
    h = ctx.compute_h(orc.fr_mont(H(t["a"])), orc.fr_mont(H(t["b"])), orc.fr_mont(H(t["c"])), len(t["a"]), 5)
    assert orc.fr_unmont(h) == H(t["h_bitrev"])


@pytest.mark.parametrize("logn", [1, 2, 3, 4, 7, 10, 13, 14])
def test_all_variants_vs_oracle(ctx, logn):
    v = rand_scalars_np(1 << logn, 50 + logn)
    for inverse in (False, True):
        for dit in (False, True):
            for coset in (False, True):
                got = ctx.ntt(v.copy(), logn, inverse, dit, coset)
                assert np.array_equal(got, orc.ntt(v, logn, inverse, dit, coset)), (logn, inverse, dit, coset)


@pytest.mark.parametrize("logn,m", [(1, 2), (4, 9), (12, 4096), (12, 2500), (16, 60000)])
def test_compute_h_vs_oracle(ctx, logn, m):
    a = rand_scalars_np(m, 60 + logn); b = rand_scalars_np(m, 61 + logn)
    c = np.zeros_like(a)
    orc.lib().orc_fr_mul_batch(a.ctypes.data_as(orc.C.c_void_p), b.ctypes.data_as(orc.C.c_void_p), c.ctypes.data_as(orc.C.c_void_p), orc.C.c_size_t(m))
    assert np.array_equal(ctx.compute_h(a, b, c, m, logn), orc.compute_h(a, b, c, logn))


def test_roundtrip_and_quotient_identity_2pow20(ctx):
    """size-independent properties at a size the oracle is not asked to match: iNTT(NTT(x)) = x, and
    h(x0) * (x0^n - 1) = A(x0)*B(x0) - C(x0) at a random point for the h returned by compute_h."""
    import torch
    logn, n = 20, 1 << 20
    v = rand_scalars_np(n, 5)
    d = torch.from_numpy(v.view(np.int64)).cuda()
    ctx.ntt(d, logn, False, False, True); ctx.ntt(d, logn, True, True, True)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint64), v)
    m = n - 1000
    a = rand_scalars_np(m, 6); b = rand_scalars_np(m, 7); c = np.zeros_like(a)
    orc.lib().orc_fr_mul_batch(a.ctypes.data_as(orc.C.c_void_p), b.ctypes.data_as(orc.C.c_void_p), c.ctypes.data_as(orc.C.c_void_p), orc.C.c_size_t(m))
    h = ctx.compute_h(a, b, c, m, logn)
    # evaluate everything at x0 with the oracle's C NTT: coefficients of A, B, C via iNTT; Horner in Python on 2^20 terms is slow,
    # so evaluate through a second domain instead: coset evaluations at g'*w^k for another shift are not available -> use
    # barycentric evaluation: P(x0) = (x0^n - 1)/n * sum_k P(w^k) w^k / (x0 - w^k), computed with numpy object arrays on a
    # strided SAMPLE is not exact; so check the exact identity on the coefficient side instead: h natural has degree <= n-2.
    hn = orc.fr_unmont(h[[int(format(n - 1, "020b")[::-1], 2)]])   # coefficient n-1 sits at bitrev(n-1) = n-1
    assert hn == [0]
    # and spot-check 3 coefficients of h*Z = A*B - C through the oracle on the same inputs at 2^20 is the oracle test above
    # at smaller sizes; here compare a strided sample of h with the oracle's compute_h (costs ~10 s of CPU)
    want = orc.compute_h(a, b, c, logn)
    assert np.array_equal(h, want)
