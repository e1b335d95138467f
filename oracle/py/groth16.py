"""ORACLE (test infrastructure, NOT product code) -- Groth16 Setup / Prove / Verify as gnark v0.10 does them.

The reference reaches this code at src/prover/prover/prover.go:269 (groth16.Prove), :276 (groth16.Verify),
src/keygen/main.go:42 (groth16.Setup) and src/prover/prover/prover.go:201 (proof.WriteRawTo).  The
implementation is out of tree (bnb-chain/gnark v0.10.1-0.20240910145009-4b5261061f04, backend/groth16/bn254
{setup,prove,verify,marshal}.go + gnark-crypto fr/pedersen, fr/hash_to_field); this file restates it
 over a *synthetic* R1CS because the gnark frontend cannot be run here.
"Parity unpinned" against gnark bytes: the maths (QAP identity) is pinned by `check_in_exponent`, which
re-derives every proof element from the toxic waste; the byte layout follows App. B.3.

Shape kept identical to the BatchCreateUser circuit: wire 0 = ONE, 1 public input, exactly one BSB22
commitment with no public committed wires (vk is then 524 B -- README.md:54).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

import bn254 as bn
from bn254 import FP, FP2, G1_GEN, G2_GEN, R, SplitMix64, pt_add, pt_mul
from ntt import Domain, bitrev, compute_h

COMMITMENT_DST = b"bsb22-commitment"


# ----------------------------------------------------------------------------- hash_to_field (RFC 9380 XMD/SHA-256)
def expand_msg_xmd(msg: bytes, dst: bytes, length: int) -> bytes:
    ell = (length + 31) // 32
    assert ell <= 255 and len(dst) <= 255
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(64) + msg + length.to_bytes(2, "big") + b"\x00" + dst_prime).digest()
    bi = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = bi
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return out[:length]


def fr_hash(msg: bytes, dst: bytes, count: int):
    """gnark-crypto fr.Hash: L = 48 bytes per element, big-endian, reduced mod r."""
    L = 48
    rnd = expand_msg_xmd(msg, dst, count * L)
    return [int.from_bytes(rnd[i * L:(i + 1) * L], "big") % R for i in range(count)]


def commitment_challenge(commitment_pt, public_committed_vals=()) -> int:
    """Prove's BSB22 hint override / Verify: hash_to_field("bsb22-commitment")(Marshal(commitment) || publics)."""
    msg = bn.g1_raw_bytes(commitment_pt) + b"".join(bn.fe_bytes(v) for v in public_committed_vals)
    return fr_hash(msg, COMMITMENT_DST, 1)[0]


# ----------------------------------------------------------------------------- synthetic R1CS
@dataclass
class R1CS:
    nb_public: int            # includes the ONE wire (index 0)
    nb_secret: int
    nb_internal: int
    L: list                   # per constraint: list of (coeff, wire)
    Rr: list
    O: list
    commitment_index: int = -1            # wire holding the commitment challenge (-1: none)
    private_committed: list = field(default_factory=list)   # sorted wire indices
    commit_after: int = 0                 # number of constraints solved before the commitment hint fires

    @property
    def nb_wires(self): return self.nb_public + self.nb_secret + self.nb_internal
    @property
    def nb_constraints(self): return len(self.L)


def synth_r1cs(n_constraints: int, nb_secret: int, seed: int, with_commitment: bool = True, fan: int = 3) -> R1CS:
    """Each constraint k defines one fresh internal wire:  (sum l_i w_i) * (sum r_j w_j) = w_out(k).
    The commitment wire (when present) sits among the internal wires after `commit_after` constraints; every
    later constraint may read it.  A quarter of the constraints use a single-term left side and a constant
    right side so that some wires never appear in A or in B (=> InfinityA / InfinityB are exercised)."""
    rng = SplitMix64(seed)
    nb_public = 2
    base = nb_public + nb_secret
    commit_after = n_constraints // 2 if with_commitment else n_constraints + 1
    c_idx = base + commit_after if with_commitment else -1
    L, Rr, O = [], [], []

    def out_wire(k):
        return base + k + (1 if with_commitment and k >= commit_after else 0)

    for k in range(n_constraints):
        avail = out_wire(k)  # wires with index < avail are already defined (commitment wire included when k >= commit_after)
        def lin(nterms):
            return [(rng.field(R) if rng.next() & 3 else 1 + (rng.next() & 0xFFFF), rng.next() % avail) for _ in range(nterms)]
        if k % 4 == 3:
            l, r = lin(1), [(1 + rng.next() % 1000, 0)]
        else:
            l, r = lin(1 + rng.next() % fan), lin(1 + rng.next() % fan)
        if with_commitment and k == commit_after:
            l.append((1, c_idx))  # make sure the challenge is actually used
        L.append(l); Rr.append(r); O.append([(1, out_wire(k))])
    cs = R1CS(nb_public, nb_secret, n_constraints + (1 if with_commitment else 0), L, Rr, O,
              commitment_index=c_idx, commit_after=commit_after)
    if with_commitment:
        cand = list(range(nb_public, c_idx))
        pick = sorted({cand[rng.next() % len(cand)] for _ in range(max(2, len(cand) // 3))})
        cs.private_committed = pick
    return cs


def synth_inputs(cs: R1CS, seed: int):
    rng = SplitMix64(seed ^ 0xABCDEF)
    pub = [rng.field(R) for _ in range(cs.nb_public - 1)]
    sec = []
    for _ in range(cs.nb_secret):  # witness-like mix: zeros, ones, small, uniform
        t = rng.next() % 10
        sec.append(0 if t < 2 else 1 if t < 4 else rng.next() & 0xFFFF if t < 7 else rng.field(R))
    return pub, sec


def _dot(terms, w):
    return sum(cf * w[i] for cf, i in terms) % R


def solve(cs: R1CS, pk, public_inputs, secret_inputs, commit_fn=None):
    """r1cs.Solve restated for the synthetic system.  Returns (wires, a, b, c, commitment_pt, committed_vals).
    commit_fn(values) -> affine point overrides the Pedersen commitment MSM (the BSB22 hint override of Prove);
    the default uses the Python group law on pk["ck_basis"]."""
    w = [None] * cs.nb_wires
    w[0] = 1
    for i, v in enumerate(public_inputs):
        w[1 + i] = v % R
    for i, v in enumerate(secret_inputs):
        w[cs.nb_public + i] = v % R
    a, b, c = [], [], []
    commitment_pt, committed_vals = None, []
    for k in range(cs.nb_constraints):
        if cs.commitment_index >= 0 and k == cs.commit_after:
            committed_vals = [w[i] for i in cs.private_committed]
            commitment_pt = commit_fn(committed_vals) if commit_fn else bn.msm_naive(pk["ck_basis"], committed_vals)
            w[cs.commitment_index] = commitment_challenge(commitment_pt)
        av, bv = _dot(cs.L[k], w), _dot(cs.Rr[k], w)
        (cf, out), = cs.O[k]
        w[out] = av * bv % R * pow(cf, -1, R) % R
        a.append(av); b.append(bv); c.append(av * bv % R)
    assert all(v is not None for v in w)
    return w, a, b, c, commitment_pt, committed_vals


# ----------------------------------------------------------------------------- Setup
def toxic_from_seed(seed: int):
    rng = SplitMix64(seed ^ 0x70C1C)
    names = ("alpha", "beta", "gamma", "delta", "tau", "sigma", "ped_g2")
    return {nm: 1 + rng.field(R - 1) for nm in names}


def setup_scalars(cs: R1CS, toxic: dict):
    """The scalar side of groth16.Setup: the discrete logs (w.r.t. the generators) of every key element.
    Cheap in Python; the points are k*G for these k (setup() below, or a fixed-base pass of the C oracle)."""
    d = Domain(cs.nb_constraints)
    n = d.n
    al, be, ga, de, tau, sigma = (toxic[k] for k in ("alpha", "beta", "gamma", "delta", "tau", "sigma"))
    ga_inv, de_inv = pow(ga, -1, R), pow(de, -1, R)
    # Lagrange basis at tau:  L_k(tau) = (tau^n - 1)/n * w^k / (tau - w^k)
    zt = (pow(tau, n, R) - 1) % R
    lag, wk = [], 1
    for k in range(n):
        lag.append(zt * d.card_inv % R * wk % R * pow((tau - wk) % R, -1, R) % R)
        wk = wk * d.gen % R
    nw = cs.nb_wires
    A, B, C = [0] * nw, [0] * nw, [0] * nw
    for k in range(cs.nb_constraints):
        for cf, i in cs.L[k]: A[i] = (A[i] + cf * lag[k]) % R
        for cf, i in cs.Rr[k]: B[i] = (B[i] + cf * lag[k]) % R
        for cf, i in cs.O[k]: C[i] = (C[i] + cf * lag[k]) % R
    committed = set(cs.private_committed)
    vkK, pkK, ckK = [], [], []
    for i in range(nw):
        t = (A[i] * be + B[i] * al + C[i]) % R
        if i < cs.nb_public or i == cs.commitment_index:
            vkK.append(t * ga_inv % R)
        elif i in committed:
            ckK.append(t * ga_inv % R)
        else:
            pkK.append(t * de_inv % R)
    zdt = zt * de_inv % R
    Z_nat = []
    for _ in range(n):
        Z_nat.append(zdt); zdt = zdt * tau % R
    Z = [Z_nat[bitrev(i, d.logn)] for i in range(n)][:n - 1]  # gnark >= 0.9: bit-reversed, n-1 kept
    return dict(domain=d, log_n=d.logn, A_s=[x for x in A if x], B_s=[x for x in B if x], K_s=pkK, Z_s=Z, vkK_s=vkK,
                infinity_a=[x == 0 for x in A], infinity_b=[x == 0 for x in B], ck_basis_s=ckK,
                ck_sigma_s=[x * sigma % R for x in ckK])


def setup(cs: R1CS, toxic: dict):
    """groth16.Setup with explicit toxic waste.  Returns (pk, vk); pk also carries the *scalars* behind each
    point array (suffix _s) so the C oracle / CUDA tests can rebuild the same points with a fixed-base pass."""
    sc = setup_scalars(cs, toxic)
    al, be, ga, de, sigma = (toxic[k] for k in ("alpha", "beta", "gamma", "delta", "sigma"))
    g1 = lambda s: pt_mul(G1_GEN, s)
    g2 = lambda s: pt_mul(G2_GEN, s, FP2)
    ped_g = g2(toxic["ped_g2"])
    pk = dict(sc)
    pk.update(alpha1=g1(al), beta1=g1(be), delta1=g1(de), beta2=g2(be), delta2=g2(de),
              A=[g1(s) for s in sc["A_s"]], B1=[g1(s) for s in sc["B_s"]], B2=[g2(s) for s in sc["B_s"]],
              K=[g1(s) for s in sc["K_s"]], Z=[g1(s) for s in sc["Z_s"]],
              ck_basis=[g1(s) for s in sc["ck_basis_s"]], ck_basis_exp_sigma=[g1(s) for s in sc["ck_sigma_s"]])
    vk = dict(alpha1=pk["alpha1"], beta1=pk["beta1"], delta1=pk["delta1"], beta2=pk["beta2"], delta2=pk["delta2"],
              gamma2=g2(ga), K=[g1(s) for s in sc["vkK_s"]], K_s=sc["vkK_s"],
              ped_g=ped_g, ped_g_root_sigma_neg=pt_mul(ped_g, (-pow(sigma, -1, R)) % R, FP2),
              public_and_commitment_committed=[[]] if cs.commitment_index >= 0 else [])
    return pk, vk


# ----------------------------------------------------------------------------- Prove
def filter_wires(cs: R1CS, pk, w):
    wa = [w[i] for i in range(len(w)) if not pk["infinity_a"][i]]
    wb = [w[i] for i in range(len(w)) if not pk["infinity_b"][i]]
    drop = set(cs.private_committed) | ({cs.commitment_index} if cs.commitment_index >= 0 else set())
    wk = [w[i] for i in range(cs.nb_public, len(w)) if i not in drop]
    return wa, wb, wk


def prove(cs: R1CS, pk, public_inputs, secret_inputs, r: int, s: int):
    """groth16.Prove with injected (r, s) (gnark samples them with crypto/rand).  Returns the proof dict and
    the intermediate vectors the CUDA parity tests feed to the C-ABI."""
    w, a, b, c, commitment_pt, committed_vals = solve(cs, pk, public_inputs, secret_inputs)
    d = pk["domain"]
    pok = bn.msm_naive(pk["ck_basis_exp_sigma"], committed_vals) if cs.commitment_index >= 0 else None
    # single commitment => Fold(poks, challenge) = poks[0]
    h = compute_h(a, b, c, d)
    wa, wb, wk = filter_wires(cs, pk, w)
    kr = (-(r * s)) % R
    d_r, d_s, d_kr = (pt_mul(pk["delta1"], x) for x in (r, s, kr))
    ar = pt_add(pt_add(bn.msm_naive(pk["A"], wa), pk["alpha1"]), d_r)
    bs1 = pt_add(pt_add(bn.msm_naive(pk["B1"], wb), pk["beta1"]), d_s)
    bs2 = pt_add(pt_add(bn.msm_naive(pk["B2"], wb, FP2), pt_mul(pk["delta2"], s, FP2), FP2), pk["beta2"], FP2)
    krs = pt_add(bn.msm_naive(pk["K"], wk), d_kr)
    krs = pt_add(krs, bn.msm_naive(pk["Z"], h[:d.n - 1]))
    krs = pt_add(krs, pt_mul(ar, s))
    krs = pt_add(krs, pt_mul(bs1, r))
    proof = dict(Ar=ar, Bs=bs2, Krs=krs, Commitments=[commitment_pt] if commitment_pt is not None or cs.commitment_index >= 0 else [],
                 CommitmentPok=pok)
    aux = dict(w=w, a=a, b=b, c=c, h=h, wa=wa, wb=wb, wk=wk, committed=committed_vals)
    return proof, aux


def proof_raw_bytes(proof) -> bytes:
    """proof.WriteRawTo: Ar(64) Bs(128) Krs(64) u32be(len) Commitments(64 each) CommitmentPok(64) = 388 B for 1."""
    out = bn.g1_raw_bytes(proof["Ar"]) + bn.g2_raw_bytes(proof["Bs"]) + bn.g1_raw_bytes(proof["Krs"])
    out += len(proof["Commitments"]).to_bytes(4, "big")
    for cpt in proof["Commitments"]:
        out += bn.g1_raw_bytes(cpt)
    out += bn.g1_raw_bytes(proof["CommitmentPok"])
    return out


def proof_from_raw_bytes(b: bytes):
    ar = bn.g1_from_bytes(b[0:64]); bs = bn.g2_from_bytes(b[64:192]); krs = bn.g1_from_bytes(b[192:256])
    k = int.from_bytes(b[256:260], "big")
    cm = [bn.g1_from_bytes(b[260 + 64 * i:324 + 64 * i]) for i in range(k)]
    pok = bn.g1_from_bytes(b[260 + 64 * k:324 + 64 * k])
    return dict(Ar=ar, Bs=bs, Krs=krs, Commitments=cm, CommitmentPok=pok)


# ----------------------------------------------------------------------------- Verify (toxic-waste check; pairing in pairing.py)
def check_in_exponent(cs: R1CS, toxic, proof, aux, r: int, s: int) -> bool:
    """Re-derive the discrete logs of Ar, Bs, Krs from the toxic waste and the solved wires, check that the
    proof points are exactly those multiples of the generators, and check the Groth16 verification equation
        ar*bs = alpha*beta + ksum*gamma + krs*delta            (all in Fr)
    plus the Pedersen relation pok = sigma * commitment."""
    d = Domain(cs.nb_constraints)
    n = d.n
    al, be, ga, de, tau, sigma = (toxic[k] for k in ("alpha", "beta", "gamma", "delta", "tau", "sigma"))
    zt = (pow(tau, n, R) - 1) % R
    lag, wk_ = [], 1
    for k in range(n):
        lag.append(zt * d.card_inv % R * wk_ % R * pow((tau - wk_) % R, -1, R) % R)
        wk_ = wk_ * d.gen % R
    At = sum(aux["a"][k] * lag[k] for k in range(cs.nb_constraints)) % R
    Bt = sum(aux["b"][k] * lag[k] for k in range(cs.nb_constraints)) % R
    Ct = sum(aux["c"][k] * lag[k] for k in range(cs.nb_constraints)) % R
    ar = (At + al + r * de) % R
    bs = (Bt + be + s * de) % R
    ht = (At * Bt - Ct) * pow(zt, -1, R) % R
    w = aux["w"]
    # per-wire (beta*A_i + alpha*B_i + C_i)(tau)
    nw = cs.nb_wires
    Ai, Bi, Ci = [0] * nw, [0] * nw, [0] * nw
    for k in range(cs.nb_constraints):
        for cf, i in cs.L[k]: Ai[i] = (Ai[i] + cf * lag[k]) % R
        for cf, i in cs.Rr[k]: Bi[i] = (Bi[i] + cf * lag[k]) % R
        for cf, i in cs.O[k]: Ci[i] = (Ci[i] + cf * lag[k]) % R
    kk = [(be * Ai[i] + al * Bi[i] + Ci[i]) % R for i in range(nw)]
    committed = set(cs.private_committed)
    pub_idx = [i for i in range(nw) if i < cs.nb_public or i == cs.commitment_index]
    ksum_pub = sum(w[i] * kk[i] for i in pub_idx) % R          # gamma * (vk.K . publicWitness)
    kcommit = sum(w[i] * kk[i] for i in committed) % R          # gamma * log(commitment)
    kpriv = sum(w[i] * kk[i] for i in range(nw) if i not in committed and i not in pub_idx) % R
    de_inv, ga_inv = pow(de, -1, R), pow(ga, -1, R)
    krs = (kpriv * de_inv + ht * zt % R * de_inv + s * ar + r * bs - r * s * de) % R
    ok = proof["Ar"] == pt_mul(G1_GEN, ar)
    ok &= proof["Bs"] == pt_mul(G2_GEN, bs, FP2)
    ok &= proof["Krs"] == pt_mul(G1_GEN, krs)
    if cs.commitment_index >= 0:
        ok &= proof["Commitments"][0] == pt_mul(G1_GEN, kcommit * ga_inv % R)
        ok &= proof["CommitmentPok"] == pt_mul(G1_GEN, kcommit * ga_inv % R * sigma % R)
        ok &= w[cs.commitment_index] == commitment_challenge(proof["Commitments"][0])
    lhs = ar * bs % R
    rhs = (al * be + (ksum_pub + kcommit) % R * ga_inv % R * ga + krs * de) % R
    ok &= lhs == rhs
    return bool(ok)


# ----------------------------------------------------------------------------- Verify (gnark backend/groth16/bn254/verify.go restated)
def verify(vk, proof, public_witness) -> bool:
    """groth16.Verify (src/prover/prover/prover.go:276, src/verifier/main.go:284): recompute the commitment challenge,
    check the Pedersen proof of knowledge e(C, G) e(pok, GRootSigmaNeg) = 1 (pok = sigma C, GRootSigmaNeg = G^(-1/sigma)) and
    e(Krs, -delta) e(Ar, Bs) e(K0 + sum pw_i K_i + C, -gamma) = e(alpha, beta)."""
    import pairing as pr
    pw = list(public_witness)
    ksum = vk["K"][0]
    if proof["Commitments"]:
        cpt = proof["Commitments"][0]
        pw = pw + [commitment_challenge(cpt, [pw[i - 1] for i in vk["public_and_commitment_committed"][0]])]
        if not pr.pairing_check([(cpt, vk["ped_g"]), (proof["CommitmentPok"], vk["ped_g_root_sigma_neg"])]):
            return False
        ksum = pt_add(ksum, cpt)
    assert len(pw) == len(vk["K"]) - 1
    for k, v in zip(vk["K"][1:], pw):
        ksum = pt_add(ksum, pt_mul(k, v))
    return pr.pairing_check([(proof["Krs"], bn.pt_neg(vk["delta2"], FP2)), (proof["Ar"], proof["Bs"]),
                             (ksum, bn.pt_neg(vk["gamma2"], FP2)), (bn.pt_neg(vk["alpha1"]), vk["beta2"])])
