"""Shared test helpers: seeded inputs and a synthetic Groth16 instance built with the oracle (test infrastructure)."""
import json
import os

import numpy as np

import groth16 as g16
import orc
from bn254 import R, SplitMix64

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def golden():
    return json.load(open(os.path.join(GOLDEN, "oracle_vectors.json")))


def H(xs):
    return [int(x, 16) for x in xs]


def rand_scalars(n, seed, kind="uniform"):
    """kind: uniform | witness (60% zero, 15% one, 20% < 2^16, 5% uniform -- SURVEY.md 8(d) config 2)"""
    rng = SplitMix64(seed)
    out = []
    for _ in range(n):
        if kind == "uniform":
            out.append(rng.field(R))
        else:
            t = rng.next() % 100
            out.append(0 if t < 60 else 1 if t < 75 else rng.next() & 0xFFFF if t < 95 else rng.field(R))
    return out


def rand_scalars_np(n, seed):
    """n uniform Fr elements as (n,4) uint64 Montgomery-looking limbs, vectorised (values are uniform < r by
    rejection on the top limb; they are *used as* Montgomery representations, which is a bijection of Fr)."""
    rs = np.random.RandomState(seed)
    out = rs.randint(0, 1 << 63, size=(n, 4), dtype=np.int64).astype(np.uint64) * np.uint64(2) + rs.randint(0, 2, size=(n, 4)).astype(np.uint64)
    out[:, 3] &= np.uint64((1 << 60) - 1)   # < 2^252 < r
    return np.ascontiguousarray(out)


def g1_points(n, seed):
    rng = SplitMix64(seed)
    return orc.g1_fixed_base(orc.ints_to_limbs([1 + rng.field(R - 1) for _ in range(n)]))


def g2_points(n, seed):
    rng = SplitMix64(seed)
    return orc.g2_fixed_base(orc.ints_to_limbs([1 + rng.field(R - 1) for _ in range(n)]))


def pk_arrays(sc, toxic):
    lim = orc.ints_to_limbs
    return dict(A=orc.g1_fixed_base(lim(sc["A_s"])), B1=orc.g1_fixed_base(lim(sc["B_s"])), B2=orc.g2_fixed_base(lim(sc["B_s"])),
                K=orc.g1_fixed_base(lim(sc["K_s"])), Z=orc.g1_fixed_base(lim(sc["Z_s"])),
                ck_basis=orc.g1_fixed_base(lim(sc["ck_basis_s"])), ck_basis_exp_sigma=orc.g1_fixed_base(lim(sc["ck_sigma_s"])),
                alpha1=orc.g1_fixed_base(lim([toxic["alpha"]])), beta1=orc.g1_fixed_base(lim([toxic["beta"]])),
                delta1=orc.g1_fixed_base(lim([toxic["delta"]])), beta2=orc.g2_fixed_base(lim([toxic["beta"]])),
                delta2=orc.g2_fixed_base(lim([toxic["delta"]])), log_n=sc["log_n"])


def synthetic_instance(n_constraints, nb_secret, seed, input_seed=None):
    """R1CS + key + solved witness, all via the oracle.  Returns dict with arrays in gnark memory layout.
    input_seed: another assignment of the same circuit (same key)."""
    cs = g16.synth_r1cs(n_constraints, nb_secret, seed)
    tox = g16.toxic_from_seed(seed + 1)
    sc = g16.setup_scalars(cs, tox)
    arr = pk_arrays(sc, tox)
    pub, sec = g16.synth_inputs(cs, seed + 2 if input_seed is None else input_seed)
    commit_fn = lambda vals: orc.g1_unpack(orc.g1_msm(arr["ck_basis"], orc.fr_mont(vals)))[0]
    w, a, b, c, cpt, cvals = g16.solve(cs, None, pub, sec, commit_fn)
    wa = [w[i] for i in range(len(w)) if not sc["infinity_a"][i]]
    wb = [w[i] for i in range(len(w)) if not sc["infinity_b"][i]]
    drop = set(cs.private_committed) | {cs.commitment_index}
    wk = [w[i] for i in range(cs.nb_public, len(w)) if i not in drop]
    return dict(cs=cs, tox=tox, sc=sc, arr=arr, w=w, a=a, b=b, c=c, wa=wa, wb=wb, wk=wk, committed=cvals, commitment=cpt, vk=oracle_vk(cs, sc, tox))


def oracle_vk(cs, sc, tox):
    """groth16.VerifyingKey as Python points (the dict oracle/py/groth16.py verify() takes), from the setup scalars"""
    from bn254 import FP2, G1_GEN, G2_GEN, pt_mul
    g1 = lambda s: pt_mul(G1_GEN, s)
    g2 = lambda s: pt_mul(G2_GEN, s, FP2)
    ped_g = g2(tox["ped_g2"])
    return dict(alpha1=g1(tox["alpha"]), beta2=g2(tox["beta"]), gamma2=g2(tox["gamma"]), delta2=g2(tox["delta"]),
                K=[g1(s) for s in sc["vkK_s"]], ped_g=ped_g, ped_g_root_sigma_neg=pt_mul(ped_g, (-pow(tox["sigma"], -1, R)) % R, FP2),
                public_and_commitment_committed=[[]] if cs.commitment_index >= 0 else [])


def vk_arrays(inst):
    """keyword arguments of zkpor_b200.VerifyingKey (numpy, gnark memory layout) for a synthetic instance"""
    vk = inst["vk"]
    return dict(alpha1=orc.g1_pack([vk["alpha1"]]), beta2=orc.g2_pack([vk["beta2"]]), gamma2=orc.g2_pack([vk["gamma2"]]),
                delta2=orc.g2_pack([vk["delta2"]]), K=orc.g1_pack(vk["K"]), n_commitments=len(vk["public_and_commitment_committed"]),
                public_committed=(), ped_g=orc.g2_pack([vk["ped_g"]]), ped_g_root_sigma_neg=orc.g2_pack([vk["ped_g_root_sigma_neg"]]))


def oracle_proof(inst, r, s):
    m = orc.fr_mont
    return orc.groth16_prove(inst["arr"], m(inst["wa"]), m(inst["wb"]), m(inst["wk"]), m(inst["committed"]),
                             m(inst["a"]), m(inst["b"]), m(inst["c"]), r, s)


def make_pk(zk, ctx, inst):
    arr, sc, cs = inst["arr"], inst["sc"], inst["cs"]
    return zk.ProvingKey(ctx, log_n=arr["log_n"], A=arr["A"], B1=arr["B1"], K=arr["K"], Z=arr["Z"], B2=arr["B2"],
                         alpha1=arr["alpha1"], beta1=arr["beta1"], delta1=arr["delta1"], beta2=arr["beta2"], delta2=arr["delta2"],
                         n_a=len(sc["A_s"]), n_b=len(sc["B_s"]), n_k=len(sc["K_s"]), n_z=len(sc["Z_s"]),
                         infinity_a=sc["infinity_a"], infinity_b=sc["infinity_b"], n_public=cs.nb_public,
                         ck_basis=arr["ck_basis"], ck_basis_exp_sigma=arr["ck_basis_exp_sigma"],
                         private_committed=cs.private_committed, commitment_index=cs.commitment_index)


def r1cs_csr(cs):
    """CSR matrices + coefficient table of an oracle R1CS in the layout zkpor_b200.R1CS takes (ids 0..4 = 0, 1, 2, -1, -2 as in gnark's table)"""
    table = [0, 1, 2, R - 1, R - 2]
    ids = {v: i for i, v in enumerate(table)}
    mats = []
    for rows in (cs.L, cs.Rr, cs.O):
        rp, wi, ci = [0], [], []
        for terms in rows:
            for cf, w in terms:
                if cf not in ids:
                    ids[cf] = len(table); table.append(cf)
                wi.append(w); ci.append(ids[cf])
            rp.append(len(wi))
        mats.append((np.array(rp, dtype=np.uint64), np.array(wi, dtype=np.uint32), np.array(ci, dtype=np.uint32)))
    return mats, orc.fr_mont(table)


# ----------------------------------------------------------------------------------------------- circuit-shaped instances (solver tests)
def oracle_poseidon_constants(t):
    import poseidon as ps
    rc, mds = ps.constants(t)
    return rc, mds, ps.ROUNDS_P[t - 2]


def circuit_synth():
    import importlib
    return importlib.import_module("zkmerkle-proof-of-solvency_b200.circuit_synth")


def circuit_instance(seed=1, with_key=True, **shape):
    """A BatchCreateUser-shaped circuit (circuit_synth.batch_create_user_like), its flat program, a satisfying input assignment
    and -- with_key -- a Groth16 key made by the oracle's Setup from explicit toxic waste."""
    cs_mod = circuit_synth()
    cb = cs_mod.batch_create_user_like(poseidon_constants=oracle_poseidon_constants, **shape)
    flat = cb.flatten()
    inputs = cs_mod.draw_inputs(flat, seed)                       # canonical limbs
    ints = [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in inputs]
    inst = dict(flat=flat, inputs=ints, inputs_mont=orc.fr_mont(ints))
    if with_key:
        import solver
        cs = solver.to_r1cs(flat)
        tox = g16.toxic_from_seed(seed + 1)
        sc = g16.setup_scalars(cs, tox)
        inst.update(cs=cs, tox=tox, sc=sc, arr=pk_arrays(sc, tox))
    return inst
