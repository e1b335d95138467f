"""ORACLE (test infrastructure, NOT product code) -- BN254 fields and groups in Python big ints.

Ground truth for every other layer: the C oracle (oracle/c) and the CUDA product are both
checked against this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import it.

The reference keeps this arithmetic out of tree, in
  github.com/bnb-chain/gnark-crypto v0.14.1-0.20240910145340-609ab3a7eb9b  (ecc/bn254, ecc/bn254/fp, ecc/bn254/fr)
pinned at /root/reference/go.mod:57-60.  This is a restatement of the published curve
(alt_bn128 / BN254, EIP-196/197 parameters) and of gnark-crypto's encodings:
  * Fp, Fr elements are 4 x u64 little-endian limbs in Montgomery form, R = 2^256;
  * Bytes()/Marshal() are 32-byte big-endian canonical;
  * G1 raw = X||Y (64 B), compressed = X with flag bits (32 B); G2 = X.A1||X.A0[||Y.A1||Y.A0].
Pinned by known answers: 2*G1 (EIP-196 vector), r*G = O, on-curve checks of both generators.
"""
from __future__ import annotations

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fp modulus q
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr modulus r
MONT_R = 1 << 256
FR_GEN = 5  # Fr multiplicative generator (gnark-crypto fft.Domain.FrMultiplicativeGen)
FR_TWO_ADICITY = 28
FR_ROOT_2_28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904
B1 = 3  # y^2 = x^3 + 3

G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


def inv(a: int, m: int) -> int:
    return pow(a, -1, m)


# ----------------------------------------------------------------------------- Fp2 = Fp[u]/(u^2+1)
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return f2_mul(a, a)
def f2_scalar(a, k): return (a[0] * k % P, a[1] * k % P)


def f2_inv(a):
    d = inv((a[0] * a[0] + a[1] * a[1]) % P, P)
    return (a[0] * d % P, (-a[1] * d) % P)


F2_ZERO = (0, 0)
F2_ONE = (1, 0)
B2 = f2_mul((3, 0), f2_inv((9, 1)))  # twist coefficient b' = 3/(9+u)


class Field:
    """Minimal field-ops vtable so the group law is written once for G1 (Fp) and G2 (Fp2)."""

    def __init__(self, add, sub, mul, neg, invf, zero, one, b):
        self.add, self.sub, self.mul, self.neg, self.inv = add, sub, mul, neg, invf
        self.zero, self.one, self.b = zero, one, b


FP = Field(lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a, b: a * b % P,
           lambda a: (-a) % P, lambda a: inv(a, P), 0, 1, B1)
FP2 = Field(f2_add, f2_sub, f2_mul, f2_neg, f2_inv, F2_ZERO, F2_ONE, B2)

# ----------------------------------------------------------------------------- group law (affine, None = infinity)


def is_on_curve(pt, F=FP):
    if pt is None:
        return True
    x, y = pt
    return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), F.b)


def pt_neg(pt, F=FP):
    return None if pt is None else (pt[0], F.neg(pt[1]))


def pt_add(p1, p2, F=FP):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 != y2 or y1 == F.zero:
            return None
        xx = F.mul(x1, x1)
        lam = F.mul(F.add(F.add(xx, xx), xx), F.inv(F.add(y1, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


# Jacobian (X, Y, Z): x = X/Z^2, y = Y/Z^3; Z == 0 is infinity.  Used for speed in scalar mul / MSM.
def jac_from_affine(pt, F=FP):
    return (F.one, F.one, F.zero) if pt is None else (pt[0], pt[1], F.one)


def jac_to_affine(j, F=FP):
    X, Y, Z = j
    if Z == F.zero:
        return None
    zi = F.inv(Z)
    zi2 = F.mul(zi, zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


def jac_double(j, F=FP):
    X, Y, Z = j
    if Z == F.zero:
        return j
    A = F.mul(X, X)
    B = F.mul(Y, Y)
    C = F.mul(B, B)
    t = F.add(X, B)
    D = F.sub(F.sub(F.mul(t, t), A), C)
    D = F.add(D, D)
    E = F.add(F.add(A, A), A)
    Fq = F.mul(E, E)
    X3 = F.sub(Fq, F.add(D, D))
    C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
    Z3 = F.mul(Y, Z); Z3 = F.add(Z3, Z3)
    return (X3, Y3, Z3)


def jac_add(j1, j2, F=FP):
    if j1[2] == F.zero:
        return j2
    if j2[2] == F.zero:
        return j1
    X1, Y1, Z1 = j1
    X2, Y2, Z2 = j2
    Z1Z1 = F.mul(Z1, Z1)
    Z2Z2 = F.mul(Z2, Z2)
    U1 = F.mul(X1, Z2Z2)
    U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
    S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        if S1 != S2:
            return (F.one, F.one, F.zero)
        return jac_double(j1, F)
    H = F.sub(U2, U1)
    Rr = F.sub(S2, S1)
    HH = F.mul(H, H)
    HHH = F.mul(H, HH)
    V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.mul(Rr, Rr), HHH), F.add(V, V))
    Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def pt_mul(pt, k: int, F=FP):
    """k*pt for an affine point; k is reduced mod r (both groups have prime order r)."""
    k %= R
    acc = (F.one, F.one, F.zero)
    base = jac_from_affine(pt, F)
    while k:
        if k & 1:
            acc = jac_add(acc, base, F)
        base = jac_double(base, F)
        k >>= 1
    return jac_to_affine(acc, F)


def msm_naive(points, scalars, F=FP):
    """sum_i scalars[i]*points[i] -- the mathematical definition gnark-crypto MultiExp computes."""
    acc = (F.one, F.one, F.zero)
    for pt, s in zip(points, scalars):
        if pt is None or s % R == 0:
            continue
        acc = jac_add(acc, jac_from_affine(pt_mul(pt, s, F), F), F)
    return jac_to_affine(acc, F)


# ----------------------------------------------------------------------------- limb / byte codecs
def to_mont_limbs(a: int, mod: int):
    """4 x u64 little-endian limbs of a*R mod m: gnark-crypto's in-memory fp/fr.Element."""
    v = a * MONT_R % mod
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_mont_limbs(limbs, mod: int) -> int:
    v = sum(int(l) << (64 * i) for i, l in enumerate(limbs))
    return v * inv(MONT_R, mod) % mod


def fe_bytes(a: int) -> bytes:
    """fr/fp.Element.Bytes(): 32-byte big-endian canonical."""
    return int(a).to_bytes(32, "big")


# gnark-crypto ecc/bn254/marshal.go flag bits (top 2 bits of byte 0)
M_UNCOMPRESSED = 0b00 << 6
M_COMPRESSED_SMALLEST = 0b10 << 6
M_COMPRESSED_LARGEST = 0b11 << 6
M_COMPRESSED_INFINITY = 0b01 << 6


def _lex_largest_fp(y: int) -> bool:
    return y > (P - 1) // 2


def g1_raw_bytes(pt) -> bytes:
    """G1Affine.RawBytes(): X||Y big-endian; infinity = all-zero with flag 01 (compressed infinity mask is
    also what gnark writes for raw infinity: byte0 = 0b01<<6 per marshal.go mUncompressedInfinity=0b01<<6)."""
    if pt is None:
        return bytes([M_COMPRESSED_INFINITY]) + bytes(63)
    return fe_bytes(pt[0]) + fe_bytes(pt[1])


def g1_compressed_bytes(pt) -> bytes:
    if pt is None:
        return bytes([M_COMPRESSED_INFINITY]) + bytes(31)
    b = bytearray(fe_bytes(pt[0]))
    b[0] |= M_COMPRESSED_LARGEST if _lex_largest_fp(pt[1]) else M_COMPRESSED_SMALLEST
    return bytes(b)


def fp_sqrt(a: int):
    """q = 3 mod 4 => sqrt = a^((q+1)/4) when a is a square."""
    y = pow(a, (P + 1) // 4, P)
    return y if y * y % P == a % P else None


def g1_from_bytes(b: bytes):
    flag = b[0] & 0xC0
    if flag == M_COMPRESSED_INFINITY:
        return None
    if flag == M_UNCOMPRESSED:
        return (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:64], "big"))
    x = int.from_bytes(bytes([b[0] & 0x3F]) + b[1:32], "big")
    y = fp_sqrt((x * x * x + 3) % P)
    if y is None:
        raise ValueError("not on curve")
    if _lex_largest_fp(y) != (flag == M_COMPRESSED_LARGEST):
        y = P - y
    return (x, y)


def _lex_largest_fp2(y) -> bool:
    # gnark-crypto E2.LexicographicallyLargest: compare A1 first, then A0
    if y[1] == 0:
        return _lex_largest_fp(y[0])
    return _lex_largest_fp(y[1])


def g2_raw_bytes(pt) -> bytes:
    if pt is None:
        return bytes([M_COMPRESSED_INFINITY]) + bytes(127)
    (x, y) = pt
    return fe_bytes(x[1]) + fe_bytes(x[0]) + fe_bytes(y[1]) + fe_bytes(y[0])


def g2_compressed_bytes(pt) -> bytes:
    if pt is None:
        return bytes([M_COMPRESSED_INFINITY]) + bytes(63)
    (x, y) = pt
    b = bytearray(fe_bytes(x[1]) + fe_bytes(x[0]))
    b[0] |= M_COMPRESSED_LARGEST if _lex_largest_fp2(y) else M_COMPRESSED_SMALLEST
    return bytes(b)


def fp2_sqrt(a):
    """Square root in Fp2 (complex method, q = 3 mod 4); returns None if a is a non-residue."""
    if a == F2_ZERO:
        return F2_ZERO
    a0, a1 = a
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt((-a0) % P)
        return (0, s)
    norm = (a0 * a0 + a1 * a1) % P
    n = fp_sqrt(norm)
    if n is None:
        return None
    half = inv(2, P)
    for cand in ((a0 + n) * half % P, (a0 - n) * half % P):
        x0 = fp_sqrt(cand)
        if x0 is not None and x0 != 0:
            x1 = a1 * inv(2 * x0 % P, P) % P
            if f2_mul((x0, x1), (x0, x1)) == (a0 % P, a1 % P):
                return (x0, x1)
    return None


def g2_from_bytes(b: bytes):
    flag = b[0] & 0xC0
    if flag == M_COMPRESSED_INFINITY:
        return None
    x1 = int.from_bytes(bytes([b[0] & 0x3F]) + b[1:32], "big")
    x0 = int.from_bytes(b[32:64], "big")
    if flag == M_UNCOMPRESSED:
        return ((x0, x1), (int.from_bytes(b[96:128], "big"), int.from_bytes(b[64:96], "big")))
    x = (x0, x1)
    y = fp2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B2))
    if y is None:
        raise ValueError("not on curve")
    if _lex_largest_fp2(y) != (flag == M_COMPRESSED_LARGEST):
        y = f2_neg(y)
    return (x, y)


# ----------------------------------------------------------------------------- deterministic test data
class SplitMix64:
    """Seeded generator shared (by restatement) with the C oracle and the CUDA tests."""

    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def field(self, mod: int) -> int:
        v = 0
        for _ in range(4):
            v = (v << 64) | self.next()
        return v % mod
