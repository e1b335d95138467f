"""ORACLE (test infrastructure, not product code): BN254 optimal-ate pairing in Python big ints.

Restates what gnark-crypto's `bn254.Pair` / `bn254.PairingCheck` compute for `groth16.Verify`
(src/prover/prover/prover.go:276, src/verifier/main.go:284; implementation out of tree:
github.com/bnb-chain/gnark-crypto v0.14.1-0.20240910145340-609ab3a7eb9b, ecc/bn254/pairing.go) from the published
definition: e(P, Q) = f_{6x+2,Q}(P) * l_{[6x+2]Q, pi(Q)}(P) * l_{[6x+2]Q + pi(Q), -pi^2(Q)}(P), raised to
(q^12 - 1)/r, x = 4965661367192848881.  Deliberately written the slow, obviously-correct way: Fp12 = Fp[w]/(w^12 - 18 w^6 + 82)
as dense polynomials, G2 points untwisted into E(Fp12), generic chord/tangent lines, Frobenius by exponentiation.

Pinned by: bilinearity e(aP, bQ) = e(P, Q)^(ab), non-degeneracy, e(P, Q)^r = 1 (tests/test_oracle_kat.py).  gnark's
GT *representation* differs by a fixed power (its final exponentiation uses a multiple of (q^12-1)/r), so GT bytes are
not comparable with gnark's; pairing-product EQUALITIES -- all that Verify uses -- are.

Tower basis used by the product (csrc/pairing.cuh), Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - (9+u)), Fp12 = Fp6[w]/(w^2 - v):
an element sum_k a_k w^k (a_k in Fp2, k = 0..5) has tower coordinates C0 = (a_0, a_2, a_4), C1 = (a_1, a_3, a_5); see
`to_tower` / `from_tower`.
"""
from bn254 import P, R

ATE_X = 4965661367192848881
ATE_LOOP = 6 * ATE_X + 2
FINAL_EXP = (P ** 12 - 1) // R


# ------------------------------------------------------------------ Fp12 as polynomials mod w^12 - 18 w^6 + 82
def f12(coeffs):
    return tuple(c % P for c in coeffs)


F12_ZERO = f12([0] * 12)
F12_ONE = f12([1] + [0] * 11)


def f12_add(a, b): return tuple((x + y) % P for x, y in zip(a, b))
def f12_sub(a, b): return tuple((x - y) % P for x, y in zip(a, b))
def f12_neg(a): return tuple((-x) % P for x in a)
def f12_scalar(a, k): return tuple(x * k % P for x in a)


def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):   # w^12 = 18 w^6 - 82
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return tuple(x % P for x in t[:12])


def f12_sqr(a): return f12_mul(a, a)


def f12_pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


def _poly_divmod_inv(a):
    """inverse in Fp[w]/(m) by the extended Euclid on polynomials"""
    m = [82] + [0] * 5 + [-18 % P] + [0] * 5 + [1]
    def deg(p):
        d = len(p) - 1
        while d >= 0 and p[d] % P == 0:
            d -= 1
        return d
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], m[:]
    while deg(low) > 0:
        # r = high divided by low
        dl, dh = deg(low), deg(high)
        r = [0] * 13
        tmp = high[:]
        inv_lead = pow(low[dl], -1, P)
        for i in range(dh - dl, -1, -1):
            q = tmp[dl + i] * inv_lead % P
            r[i] = q
            for j in range(dl + 1):
                tmp[i + j] = (tmp[i + j] - q * low[j]) % P
        nm, new = hm[:], high[:]
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % P
                new[i + j] = (new[i + j] - low[i] * r[j]) % P
        lm, low, hm, high = nm, new, lm, low
    c = pow(low[0], -1, P)
    return tuple(x * c % P for x in lm[:12])


def f12_inv(a):
    r = _poly_divmod_inv(a)
    assert f12_mul(r, a) == F12_ONE
    return r


def from_fp2_coeffs(aks):
    """sum_k a_k w^k with a_k = (x, y) in Fp2, u = w^6 - 9  ->  dense Fp polynomial"""
    out = [0] * 12
    for k, (x, y) in enumerate(aks):
        out[k] = (out[k] + x - 9 * y) % P
        out[k + 6] = (out[k + 6] + y) % P
    return tuple(out)


def to_fp2_coeffs(a):
    return [((a[k] + 9 * a[k + 6]) % P, a[k + 6]) for k in range(6)]


def to_tower(a):
    """dense polynomial -> ((c00, c01, c02), (c10, c11, c12)), each an Fp2 pair: gnark-crypto E12{C0, C1 E6{B0, B1, B2}}"""
    ak = to_fp2_coeffs(a)
    return ((ak[0], ak[2], ak[4]), (ak[1], ak[3], ak[5]))


def from_tower(t):
    (c00, c01, c02), (c10, c11, c12) = t
    return from_fp2_coeffs([c00, c10, c01, c11, c02, c12])


# ------------------------------------------------------------------ points in E(Fp12): y^2 = x^3 + 3
def untwist(q):
    """G2 point on the twist y^2 = x^3 + 3/(9+u) -> E(Fp12): (x w^2, y w^3)"""
    if q is None:
        return None
    x, y = q
    return (from_fp2_coeffs([(0, 0), (0, 0), x, (0, 0), (0, 0), (0, 0)]), from_fp2_coeffs([(0, 0), (0, 0), (0, 0), y, (0, 0), (0, 0)]))


def embed_g1(p):
    if p is None:
        return None
    return (f12([p[0]] + [0] * 11), f12([p[1]] + [0] * 11))


def e12_on_curve(pt):
    x, y = pt
    return f12_sub(f12_sqr(y), f12_mul(f12_sqr(x), x)) == f12([3] + [0] * 11)


def e12_double(pt):
    x, y = pt
    lam = f12_mul(f12_scalar(f12_sqr(x), 3), f12_inv(f12_scalar(y, 2)))
    nx = f12_sub(f12_sqr(lam), f12_scalar(x, 2))
    return (nx, f12_sub(f12_mul(lam, f12_sub(x, nx)), y))


def e12_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        return e12_double(p1) if y1 == y2 else None
    lam = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
    nx = f12_sub(f12_sub(f12_sqr(lam), x1), x2)
    return (nx, f12_sub(f12_mul(lam, f12_sub(x1, nx)), y1))


def linefunc(p1, p2, t):
    """line through p1, p2 (tangent when equal) evaluated at t"""
    (x1, y1), (x2, y2), (xt, yt) = p1, p2, t
    if x1 != x2:
        lam = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
    elif y1 == y2:
        lam = f12_mul(f12_scalar(f12_sqr(x1), 3), f12_inv(f12_scalar(y1, 2)))
    else:
        return f12_sub(xt, x1)
    return f12_sub(f12_mul(lam, f12_sub(xt, x1)), f12_sub(yt, y1))


def miller_loop(q12, p12):
    if q12 is None or p12 is None:
        return F12_ONE
    rpt, f = q12, F12_ONE
    for bit in bin(ATE_LOOP)[3:]:
        f = f12_mul(f12_sqr(f), linefunc(rpt, rpt, p12))
        rpt = e12_double(rpt)
        if bit == "1":
            f = f12_mul(f, linefunc(rpt, q12, p12))
            rpt = e12_add(rpt, q12)
    q1 = (f12_pow(q12[0], P), f12_pow(q12[1], P))
    nq2 = (f12_pow(q1[0], P), f12_neg(f12_pow(q1[1], P)))
    f = f12_mul(f, linefunc(rpt, q1, p12))
    rpt = e12_add(rpt, q1)
    f = f12_mul(f, linefunc(rpt, nq2, p12))
    return f


def final_exponentiation(f):
    return f12_pow(f, FINAL_EXP)


def pairing(g1_pt, g2_pt):
    """e(P, Q) for P in G1 (affine Fp pair or None), Q in G2 (affine Fp2 pair or None)"""
    return final_exponentiation(miller_loop(untwist(g2_pt), embed_g1(g1_pt)))


def pairing_product(pairs):
    """prod e(P_i, Q_i): one shared final exponentiation (what PairingCheck / Verify evaluate)"""
    f = F12_ONE
    for p, q in pairs:
        f = f12_mul(f, miller_loop(untwist(q), embed_g1(p)))
    return final_exponentiation(f)


def pairing_check(pairs) -> bool:
    return pairing_product(pairs) == F12_ONE
