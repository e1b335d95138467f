"""GPU parity: groth16.Setup building blocks (fixed-base batch multiplication, Lagrange basis, per-wire sums) and the
whole key against the oracle's Setup restatement on the synthetic R1CS; a proof made with the GPU-built key equals the
oracle's proof.  Reference call site: src/keygen/main.go:42."""
import numpy as np
import pytest

import bn254 as bn
import groth16 as g16
import orc
import zkpor_b200 as zk
from bn254 import FP2, G1_GEN, G2_GEN, R, SplitMix64
from helpers import oracle_proof, synthetic_instance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def test_fixed_base_batch_vs_oracle(ctx):
    rng = SplitMix64(3)
    ks = [rng.field(R) for _ in range(2000)] + [0, 1, 2, R - 1, 255, 256, 1 << 248]
    n = len(ks)
    g1 = orc.g1_pack([G1_GEN])[0]; g2 = orc.g2_pack([G2_GEN])[0]
    out1 = np.zeros((n, 8), dtype=np.uint64); out2 = np.zeros((n, 16), dtype=np.uint64)
    zk.g1_fixed_base_batch(ctx, g1, orc.fr_mont(ks), n, out1)
    assert np.array_equal(out1, orc.g1_fixed_base(orc.ints_to_limbs(ks)))
    zk.g2_fixed_base_batch(ctx, g2, orc.ints_to_limbs(ks), n, out2, zk.ZKPOR_SCALARS_PLAIN)
    assert np.array_equal(out2, orc.g2_fixed_base(orc.ints_to_limbs(ks)))
    # arbitrary base
    base = orc.g1_pack([bn.pt_mul(G1_GEN, 777)])[0]
    zk.g1_fixed_base_batch(ctx, base, orc.fr_mont(ks[:50]), 50, out1[:50])
    assert orc.g1_unpack(out1[:50]) == [bn.pt_mul(G1_GEN, 777 * k) for k in ks[:50]]


def csc_from(rows_terms, n_wires):
    """rows_terms: per constraint list of (coeff, wire) -> CSC (col_ptr, rows, coeffs mont)"""
    cols = [[] for _ in range(n_wires)]
    for r, terms in enumerate(rows_terms):
        for cf, w in terms:
            cols[w].append((r, cf))
    ptr, rows, cfs = [0], [], []
    for c in cols:
        for r, cf in c:
            rows.append(r); cfs.append(cf)
        ptr.append(len(rows))
    return np.array(ptr, dtype=np.uint64), np.array(rows, dtype=np.uint32), orc.fr_mont(cfs).reshape(-1, 4)


@pytest.mark.parametrize("n_constraints", [61, 700])
def test_setup_key_and_proof_vs_oracle(ctx, n_constraints):
    inst = synthetic_instance(n_constraints, 30, seed=900 + n_constraints)
    cs, tox, sc, arr = inst["cs"], inst["tox"], inst["sc"], inst["arr"]
    nw = cs.nb_wires
    pk_kwargs, extras = zk.groth16_setup(ctx, sc["log_n"], nw, cs.nb_public, csc_from(cs.L, nw), csc_from(cs.Rr, nw), csc_from(cs.O, nw),
                                         cs.private_committed, cs.commitment_index, tox)
    host = lambda t, w: t.cpu().numpy().view(np.uint64).reshape(-1, w)
    for name, key, w in (("A", "A", 8), ("B1", "B1", 8), ("K", "K", 8), ("Z", "Z", 8), ("B2", "B2", 16), ("ck_basis", "ck_basis", 8),
                         ("ck_basis_exp_sigma", "ck_basis_exp_sigma", 8)):
        want = arr[key]
        assert np.array_equal(host(pk_kwargs[name], w)[:want.shape[0]], want), name
    assert pk_kwargs["infinity_a"].tolist() == [int(x) for x in sc["infinity_a"]]
    assert pk_kwargs["infinity_b"].tolist() == [int(x) for x in sc["infinity_b"]]
    for nm in ("alpha1", "beta1", "delta1", "beta2", "delta2"):
        assert np.array_equal(pk_kwargs[nm].reshape(-1), arr[nm].reshape(-1)), nm
    assert np.array_equal(host(extras["vk_K"], 8)[:extras["n_vk"]], orc.g1_fixed_base(orc.ints_to_limbs(sc["vkK_s"])))
    # prove with the GPU-built key
    pk = zk.ProvingKey(ctx, **pk_kwargs)
    m = orc.fr_mont
    rng = SplitMix64(5); r, s = rng.field(R), rng.field(R)
    assert pk.prove(m(inst["w"]), m(inst["a"]), m(inst["b"]), m(inst["c"]), n_constraints, r, s) == oracle_proof(inst, r, s)
    pk.close()
