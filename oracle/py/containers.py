"""ORACLE (test infrastructure, NOT product code) -- the byte containers of gnark's Groth16 objects, restated in Python.

Restates gnark v0.10 backend/groth16/bn254/marshal.go (Proof / VerifyingKey / ProvingKey writeTo and ReadFrom) and gnark-crypto
v0.14 ecc/bn254/marshal.go (Encoder / Decoder) + fr/fft/domain.go (Domain.WriteTo) -- all out of tree (/root/reference/go.mod:57-60);
call sites: src/prover/prover/prover.go:201,317-362, src/verifier/main.go:33-34,208-216, src/keygen/main.go:46-62.  Layouts as in
SURVEY.md App. B.3 [memory]; pinned only by the reference's own file sizes: a verifying key of one public input and one commitment is
524 bytes (README.md:54,57), which this module reproduces (tests/test_oracle_kat.py).  PARITY UNPINNED beyond that: the Domain header's
trailing flag byte and the []bool packing are from memory of the pinned gnark-crypto version.
Points are affine tuples of Python ints (None = infinity), as everywhere in oracle/py."""
import bn254 as bn
from bn254 import FP2, R
from ntt import FR_GEN


def _g1(pt, raw):
    return bn.g1_raw_bytes(pt) if raw else bn.g1_compressed_bytes(pt)


def _g2(pt, raw):
    return bn.g2_raw_bytes(pt) if raw else bn.g2_compressed_bytes(pt)


def _u32(v):
    return int(v).to_bytes(4, "big")


def _u64(v):
    return int(v).to_bytes(8, "big")


def _fr(v):
    return (int(v) % R).to_bytes(32, "big")


def _slice_g1(pts, raw):
    return _u32(len(pts)) + b"".join(_g1(p, raw) for p in pts)


def _slice_g2(pts, raw):
    return _u32(len(pts)) + b"".join(_g2(p, raw) for p in pts)


def _bools(bits):
    out = bytearray((len(bits) + 7) // 8)
    for i, b in enumerate(bits):
        if b:
            out[i >> 3] |= 1 << (i & 7)
    return _u32(len(bits)) + bytes(out)


class _Reader:
    def __init__(self, b):
        self.b, self.o = bytes(b), 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("truncated")
        v = self.b[self.o:self.o + n]; self.o += n
        return v

    def u32(self):
        return int.from_bytes(self.take(4), "big")

    def u64(self):
        return int.from_bytes(self.take(8), "big")

    def g1(self, raw_hint=False):
        f = self.b[self.o] >> 6
        raw = f == 0 or (f == 1 and raw_hint)
        return bn.g1_from_bytes(self.take(64 if raw else 32)), raw

    def g2(self, raw_hint=False):
        f = self.b[self.o] >> 6
        raw = f == 0 or (f == 1 and raw_hint)
        return bn.g2_from_bytes(self.take(128 if raw else 64)), raw

    def g1_slice(self, raw_hint):
        return [self.g1(raw_hint)[0] for _ in range(self.u32())]

    def g2_slice(self, raw_hint):
        return [self.g2(raw_hint)[0] for _ in range(self.u32())]

    def bools(self):
        n = self.u32()
        by = self.take((n + 7) // 8)
        return [bool((by[i >> 3] >> (i & 7)) & 1) for i in range(n)]


# ------------------------------------------------------------------------------------------------ proof
def proof_bytes(proof, raw=False) -> bytes:
    """Proof.WriteTo (compressed: 196 B with one commitment) / WriteRawTo (388 B)"""
    out = _g1(proof["Ar"], raw) + _g2(proof["Bs"], raw) + _g1(proof["Krs"], raw)
    out += _slice_g1(proof["Commitments"], raw)
    return out + _g1(proof["CommitmentPok"], raw)


def proof_from_bytes(b):
    r = _Reader(b)
    ar, raw = r.g1()
    bs, _ = r.g2(raw); krs, _ = r.g1(raw)
    cm = r.g1_slice(raw)
    pok, _ = r.g1(raw)
    return dict(Ar=ar, Bs=bs, Krs=krs, Commitments=cm, CommitmentPok=pok), r.o


# ------------------------------------------------------------------------------------------------ verifying key
def vk_bytes(vk, raw=False) -> bytes:
    """VerifyingKey.WriteTo: alpha1 beta1 beta2 gamma2 delta1 delta2 | K | PublicAndCommitmentCommitted | Pedersen vk"""
    out = _g1(vk["alpha1"], raw) + _g1(vk["beta1"], raw) + _g2(vk["beta2"], raw) + _g2(vk["gamma2"], raw) + _g1(vk["delta1"], raw) + _g2(vk["delta2"], raw)
    out += _slice_g1(vk["K"], raw)
    pcc = vk["public_and_commitment_committed"]
    out += _u32(len(pcc))
    for inner in pcc:
        out += _u32(len(inner)) + b"".join(_u64(x) for x in inner)
    if pcc:
        out += _g2(vk["ped_g"], raw) + _g2(vk["ped_g_root_sigma_neg"], raw)
    return out


def vk_from_bytes(b):
    r = _Reader(b)
    a1, raw = r.g1()
    vk = dict(alpha1=a1)
    vk["beta1"], _ = r.g1(raw); vk["beta2"], _ = r.g2(raw); vk["gamma2"], _ = r.g2(raw); vk["delta1"], _ = r.g1(raw); vk["delta2"], _ = r.g2(raw)
    vk["K"] = r.g1_slice(raw)
    pcc = []
    for _ in range(r.u32()):
        pcc.append([r.u64() for _ in range(r.u32())])
    vk["public_and_commitment_committed"] = pcc
    if pcc:
        vk["ped_g"], _ = r.g2(raw); vk["ped_g_root_sigma_neg"], _ = r.g2(raw)
    return vk, r.o


# ------------------------------------------------------------------------------------------------ proving key
def pk_bytes(pk, raw=False) -> bytes:
    """ProvingKey.WriteTo: fft.Domain | alpha1 beta1 delta1 | A B1 Z K | beta2 delta2 | B2 | nbWires NbInfinityA NbInfinityB InfinityA
    InfinityB | u32 nbCommitmentKeys | Basis BasisExpSigma.  `pk` = the dict oracle/py/groth16.py setup() returns."""
    d = pk["domain"]
    out = _u64(d.n) + _fr(d.card_inv) + _fr(d.gen) + _fr(d.gen_inv) + _fr(FR_GEN) + _fr(pow(FR_GEN, -1, R)) + b"\x01"
    out += _g1(pk["alpha1"], raw) + _g1(pk["beta1"], raw) + _g1(pk["delta1"], raw)
    out += _slice_g1(pk["A"], raw) + _slice_g1(pk["B1"], raw) + _slice_g1(pk["Z"], raw) + _slice_g1(pk["K"], raw)
    out += _g2(pk["beta2"], raw) + _g2(pk["delta2"], raw) + _slice_g2(pk["B2"], raw)
    ia, ib = list(pk["infinity_a"]), list(pk["infinity_b"])
    out += _u64(len(ia)) + _u64(sum(ia)) + _u64(sum(ib)) + _bools(ia) + _bools(ib)
    has = len(pk["ck_basis"]) > 0
    out += _u32(1 if has else 0)
    if has:
        out += _slice_g1(pk["ck_basis"], raw) + _slice_g1(pk["ck_basis_exp_sigma"], raw)
    return out
