"""GPU parity: zkpor_groth16_prove -- proof.WriteRawTo bytes bit-exact against the oracle on the same key, witness
and (r, s).  Replaces groth16.Prove at src/prover/prover/prover.go:269."""
import numpy as np
import pytest

import groth16 as g16
import orc
import zkpor_b200 as zk
from bn254 import R, SplitMix64
from helpers import H, golden, make_pk, oracle_proof, r1cs_csr, synthetic_instance
from test_oracle_c import pk_arrays_from_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def test_golden_proof_bytes(ctx):
    g = golden()["groth16"]
    arr = pk_arrays_from_golden(g)
    pk = zk.ProvingKey(ctx, log_n=g["log_n"], A=arr["A"], B1=arr["B1"], K=arr["K"], Z=arr["Z"], B2=arr["B2"],
                       alpha1=arr["alpha1"], beta1=arr["beta1"], delta1=arr["delta1"], beta2=arr["beta2"], delta2=arr["delta2"],
                       n_a=len(g["pk_scalars"]["A"]), n_b=len(g["pk_scalars"]["B"]), n_k=len(g["pk_scalars"]["K"]), n_z=len(g["pk_scalars"]["Z"]),
                       infinity_a=g["infinity_a"], infinity_b=g["infinity_b"], n_public=g["nb_public"],
                       ck_basis=arr["ck_basis"], ck_basis_exp_sigma=arr["ck_basis_exp_sigma"],
                       private_committed=g["private_committed"], commitment_index=g["commitment_index"])
    m = orc.fr_mont
    proof = pk.prove(m(H(g["wires"])), m(H(g["a"])), m(H(g["b"])), m(H(g["c"])), g["n_constraints"], int(g["r"], 16), int(g["s"], 16))
    assert len(proof) == 388 and proof.hex() == g["proof_raw"]
    # the mid-solve Pedersen commitment (BSB22 hint) through its own entry point
    w = H(g["wires"])
    cm = pk.commit(m([w[i] for i in g["private_committed"]]))
    assert orc.g1_unpack(cm)[0] == g16.proof_from_raw_bytes(proof)["Commitments"][0]


@pytest.mark.parametrize("n_constraints,nb_secret", [(700, 40), (4000, 300)])
def test_synthetic_prove_vs_oracle(ctx, n_constraints, nb_secret):
    inst = synthetic_instance(n_constraints, nb_secret, seed=n_constraints)
    rng = SplitMix64(n_constraints + 5)
    r, s = rng.field(R), rng.field(R)
    want = oracle_proof(inst, r, s)
    pk = make_pk(zk, ctx, inst)
    m = orc.fr_mont
    got = pk.prove(m(inst["w"]), m(inst["a"]), m(inst["b"]), m(inst["c"]), n_constraints, r, s)
    assert got == want
    # the proof is a valid Groth16 proof for the synthetic statement (toxic-waste check, small case only)
    if n_constraints <= 1000:
        proof = g16.proof_from_raw_bytes(got)
        aux = dict(w=inst["w"], a=inst["a"], b=inst["b"], c=inst["c"])
        assert g16.check_in_exponent(inst["cs"], inst["tox"], proof, aux, r, s)
    # multi-GPU code path on one GPU: two "ranks" each holding a chunk of every key array, partials combined
    arr, sc = inst["arr"], inst["sc"]
    h = ctx.compute_h(m(inst["a"]), m(inst["b"]), m(inst["c"]), n_constraints, arr["log_n"])
    wa, wb, wk, cm = m(inst["wa"]), m(inst["wb"]), m(inst["wk"]), m(inst["committed"])
    nz = arr["Z"].shape[0]
    parts = []
    for rank in range(2):
        sl = lambda x: slice(rank * (len(x) // 2), len(x) // 2 if rank == 0 else len(x))
        key = zk.ProvingKey(ctx, log_n=arr["log_n"], A=arr["A"][sl(arr["A"])].copy(), B1=arr["B1"][sl(arr["B1"])].copy(),
                            K=arr["K"][sl(arr["K"])].copy(), Z=arr["Z"][sl(arr["Z"])].copy(), B2=arr["B2"][sl(arr["B2"])].copy(),
                            alpha1=arr["alpha1"], beta1=arr["beta1"], delta1=arr["delta1"], beta2=arr["beta2"], delta2=arr["delta2"],
                            n_a=len(range(*sl(arr["A"]).indices(len(arr["A"])))), n_b=len(range(*sl(arr["B1"]).indices(len(arr["B1"])))),
                            n_k=len(range(*sl(arr["K"]).indices(len(arr["K"])))), n_z=len(range(*sl(arr["Z"]).indices(nz))),
                            ck_basis=arr["ck_basis"][sl(arr["ck_basis"])].copy(), ck_basis_exp_sigma=arr["ck_basis_exp_sigma"][sl(arr["ck_basis"])].copy())
        hz = h[:nz]
        parts.append(key.prove_partial(wa[sl(wa)].copy(), wb[sl(wb)].copy(), wk[sl(wk)].copy(), cm[sl(cm)].copy(), hz[sl(hz)].copy(),
                                       len(range(*sl(hz).indices(nz)))))
        key.close()
    assert pk.finish(np.stack(parts), r, s) == want
    pk.close()


def test_constraint_evaluation_and_prove_from_wires(ctx):
    """a = L w, b = R w, c = O w on the device (the linear-algebra half of r1cs.Solve) bit-exact against the oracle's solver, and
    groth16.Prove from the wire vector alone gives the same proof bytes as with a, b, c handed over."""
    n_constraints = 900
    inst = synthetic_instance(n_constraints, 60, seed=31)
    cs = inst["cs"]
    mats, table = r1cs_csr(cs)
    r1 = zk.R1CS(ctx, n_constraints, cs.nb_wires, mats, table)
    m = orc.fr_mont
    a, b, c = r1.eval(m(inst["w"]))
    assert orc.fr_unmont(a) == inst["a"] and orc.fr_unmont(b) == inst["b"]
    assert orc.fr_unmont(c) == [x * y % R for x, y in zip(inst["a"], inst["b"])] == inst["c"]
    import torch
    dw = torch.from_numpy(m(inst["w"]).view(np.int64)).cuda()
    a2, _, _ = r1.eval(dw)                                   # device-resident wires
    assert np.array_equal(a, a2)
    rng = SplitMix64(32)
    r, s = rng.field(R), rng.field(R)
    pk = make_pk(zk, ctx, inst)
    want = oracle_proof(inst, r, s)
    assert pk.prove_wires(r1, m(inst["w"]), r, s) == want
    assert pk.prove(m(inst["w"]), m(inst["a"]), m(inst["b"]), m(inst["c"]), n_constraints, r, s) == want
    # malformed matrices are rejected on upload
    bad = [(mats[0][0], mats[0][1].copy(), mats[0][2]), mats[1], mats[2]]
    bad[0][1][0] = cs.nb_wires
    with pytest.raises(zk.ZkporError):
        zk.R1CS(ctx, n_constraints, cs.nb_wires, bad, table)
    pk.close(); r1.close()


def test_prove_and_verify_without_commitment(ctx):
    """circuit without a BSB22 commitment: 324-byte proof (zero CommitmentPok), vk without a Pedersen key"""
    import groth16 as g
    from helpers import pk_arrays
    cs = g.synth_r1cs(300, 20, seed=41, with_commitment=False)
    tox = g.toxic_from_seed(42)
    sc = g.setup_scalars(cs, tox)
    arr = pk_arrays(sc, tox)
    pub, sec = g.synth_inputs(cs, 43)
    w, a, b, c, _, _ = g.solve(cs, None, pub, sec)
    pk = zk.ProvingKey(ctx, log_n=arr["log_n"], A=arr["A"], B1=arr["B1"], K=arr["K"], Z=arr["Z"], B2=arr["B2"],
                       alpha1=arr["alpha1"], beta1=arr["beta1"], delta1=arr["delta1"], beta2=arr["beta2"], delta2=arr["delta2"],
                       n_a=len(sc["A_s"]), n_b=len(sc["B_s"]), n_k=len(sc["K_s"]), n_z=len(sc["Z_s"]),
                       infinity_a=sc["infinity_a"], infinity_b=sc["infinity_b"], n_public=cs.nb_public)
    m = orc.fr_mont
    proof = pk.prove(m(w), m(a), m(b), m(c), 300, 0x1234, 0x5678)
    assert len(proof) == 324
    from helpers import oracle_vk, vk_arrays
    vk = zk.VerifyingKey(**vk_arrays(dict(vk=oracle_vk(cs, sc, tox))))
    assert vk.verify(ctx, proof, m(pub))
    assert not vk.verify(ctx, proof, m([(pub[0] + 1) % R]))
    pk.close()


def test_prove_large_skewed_witness(ctx):
    """~2.9e5 wires, a fifth of them equal to 1 and a fifth equal to 0: the shared wire sort takes the partitioned path (n >= 2^18),
    the bucket of the value 1 is a heavy bucket in every view (skip markers in the CTA-per-chunk path), most K points are at infinity
    (wires no constraint reads).  Proof bytes bit-exact against the CPU oracle; the proof verifies."""
    import time
    t0 = time.time()
    inst = synthetic_instance(20000, 270000, seed=91)
    assert len(inst["w"]) >= (1 << 18) and sum(1 for v in inst["w"] if v == 1) > 40000
    r, s = 0x1111_2222_3333, 0x4444_5555_6666
    want = oracle_proof(inst, r, s)
    pk = make_pk(zk, ctx, inst)
    m = orc.fr_mont
    got = pk.prove(m(inst["w"]), m(inst["a"]), m(inst["b"]), m(inst["c"]), 20000, r, s)
    assert got == want
    from helpers import vk_arrays
    vk = zk.VerifyingKey(**vk_arrays(inst))
    assert vk.verify(ctx, got, m(inst["w"][1:inst["cs"].nb_public]))
    pk.close()
    print("large skewed instance: %.1f s" % (time.time() - t0))
