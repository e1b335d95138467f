"""ORACLE (test infrastructure, NOT product code) -- Fr NTT on gnark-crypto's fft.Domain conventions and computeH.

Out-of-tree code restated: gnark-crypto v0.14 ecc/bn254/fr/fft (Domain, FFT,
FFTInverse, OnCoset, DIF/DIT) and gnark v0.10 backend/groth16/bn254/prove.go computeH, reached from
src/prover/prover/prover.go:269.

  * Domain(m): cardinality n = next pow2 >= m, generator = ROOT_2_28^(2^(28-log n)), coset shift = 5.
  * DIF: natural order in  -> bit-reversed out;  DIT: bit-reversed in -> natural out.
  * computeH: h = iNTT_coset( (NTT_coset(iNTT a) * NTT_coset(iNTT b) - NTT_coset(iNTT c)) * 1/(g^n - 1) ),
    returned in BIT-REVERSED order (DIF output, no final permutation) because gnark >= 0.9 stores pk.G1.Z
    bit-reversed at Setup; `h_natural` gives the coefficient vector for checks.
"""
from __future__ import annotations

from bn254 import FR_GEN, FR_ROOT_2_28, FR_TWO_ADICITY, R


def bitrev(i: int, logn: int) -> int:
    return int(format(i, f"0{logn}b")[::-1], 2) if logn else 0


def bit_reverse_permute(a):
    n = len(a)
    logn = n.bit_length() - 1
    return [a[bitrev(i, logn)] for i in range(n)]


class Domain:
    def __init__(self, m: int):
        n = 1
        while n < m:
            n <<= 1
        self.n = n
        self.logn = n.bit_length() - 1
        assert self.logn <= FR_TWO_ADICITY
        self.gen = pow(FR_ROOT_2_28, 1 << (FR_TWO_ADICITY - self.logn), R)
        self.gen_inv = pow(self.gen, -1, R)
        self.card_inv = pow(n, -1, R)
        self.coset = FR_GEN
        self.coset_inv = pow(FR_GEN, -1, R)


def _dif(a, w):
    """Gentleman-Sande butterflies, natural in -> bit-reversed out."""
    a = list(a)
    n = len(a)
    m = n
    while m > 1:
        half = m >> 1
        wm = pow(w, n // m, R)
        for start in range(0, n, m):
            tw = 1
            for j in range(half):
                u, v = a[start + j], a[start + j + half]
                a[start + j] = (u + v) % R
                a[start + j + half] = (u - v) * tw % R
                tw = tw * wm % R
        m = half
    return a


def _dit(a, w):
    """Cooley-Tukey butterflies, bit-reversed in -> natural out."""
    a = list(a)
    n = len(a)
    m = 2
    while m <= n:
        half = m >> 1
        wm = pow(w, n // m, R)
        for start in range(0, n, m):
            tw = 1
            for j in range(half):
                u, v = a[start + j], a[start + j + half] * tw % R
                a[start + j] = (u + v) % R
                a[start + j + half] = (u - v) % R
                tw = tw * wm % R
        m <<= 1
    return a


def fft(d: Domain, a, decimation: str, coset: bool = False):
    """domain.FFT(a, DIF|DIT[, OnCoset()])"""
    n, logn = d.n, d.logn
    a = list(a)
    if coset:
        if decimation == "DIF":  # natural input: a[i] *= g^i
            a = [a[i] * pow(d.coset, i, R) % R for i in range(n)]
        else:  # bit-reversed input: position i holds coefficient bitrev(i)
            a = [a[i] * pow(d.coset, bitrev(i, logn), R) % R for i in range(n)]
    return _dif(a, d.gen) if decimation == "DIF" else _dit(a, d.gen)


def fft_inverse(d: Domain, a, decimation: str, coset: bool = False):
    """domain.FFTInverse(a, DIF|DIT[, OnCoset()]) -- includes the 1/n scaling."""
    n, logn = d.n, d.logn
    a = _dif(a, d.gen_inv) if decimation == "DIF" else _dit(a, d.gen_inv)
    if not coset:
        return [x * d.card_inv % R for x in a]
    if decimation == "DIF":  # output bit-reversed
        return [a[i] * d.card_inv % R * pow(d.coset_inv, bitrev(i, logn), R) % R for i in range(n)]
    return [a[i] * d.card_inv % R * pow(d.coset_inv, i, R) % R for i in range(n)]


def naive_dft(a, w):
    n = len(a)
    return [sum(a[j] * pow(w, i * j, R) for j in range(n)) % R for i in range(n)]


def compute_h(a, b, c, d: Domain):
    """gnark computeH; inputs are the R1CS evaluation vectors (len = #constraints), zero-padded to n.
    Output: n elements, bit-reversed coefficient order (entry bitrev(i) = coefficient i)."""
    n = d.n
    pad = lambda v: list(v) + [0] * (n - len(v))
    a, b, c = pad(a), pad(b), pad(c)
    a = fft_inverse(d, a, "DIF"); b = fft_inverse(d, b, "DIF"); c = fft_inverse(d, c, "DIF")
    a = fft(d, a, "DIT", coset=True); b = fft(d, b, "DIT", coset=True); c = fft(d, c, "DIT", coset=True)
    den = pow((pow(d.coset, n, R) - 1) % R, -1, R)
    a = [(a[i] * b[i] - c[i]) * den % R for i in range(n)]
    return fft_inverse(d, a, "DIF", coset=True)


def h_natural(h_bitrev):
    return bit_reverse_permute(h_bitrev)


def poly_eval(coeffs, x):
    acc = 0
    for cf in reversed(coeffs):
        acc = (acc * x + cf) % R
    return acc
