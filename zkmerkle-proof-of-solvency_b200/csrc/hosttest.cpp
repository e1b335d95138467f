// Host-side self-test shim (NOT part of the product library): exposes the __host__ paths of ff.cuh / ec.cuh to
// tests/test_host_ff.py through ctypes so that the limb schedule of Fe::mul and the XYZZ formulas are checked on the
// CPU against the oracle before any GPU time is spent.
#define FF_HOST_EMULATE_PTX 1
#include "pairing.cuh"
using namespace ff;
using namespace ec;
extern "C" {
void ht_fp_mul(const uint32_t *a, const uint32_t *b, uint32_t *o, uint32_t *oref) {
    Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32);
    Fp r = Fp::mul(x, y), rr = Fp::mul_ref(x, y); memcpy(o, &r, 32); memcpy(oref, &rr, 32);
}
void ht_fp_mul_kara(const uint32_t *a, const uint32_t *b, uint32_t *o) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp r = Fp::mul_kara(x, y); memcpy(o, &r, 32); }
void ht_fr_mul_kara(const uint32_t *a, const uint32_t *b, uint32_t *o) { Fr x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fr r = Fr::mul_kara(x, y); memcpy(o, &r, 32); }
void ht_fr_mul(const uint32_t *a, const uint32_t *b, uint32_t *o, uint32_t *oref) {
    Fr x, y; memcpy(&x, a, 32); memcpy(&y, b, 32);
    Fr r = Fr::mul(x, y), rr = Fr::mul_ref(x, y); memcpy(o, &r, 32); memcpy(oref, &rr, 32);
}
void ht_fr_addsub(const uint32_t *a, const uint32_t *b, uint32_t *oadd, uint32_t *osub, uint32_t *oneg) {
    Fr x, y; memcpy(&x, a, 32); memcpy(&y, b, 32);
    Fr s = Fr::add(x, y), d = Fr::sub(x, y), n = Fr::neg(x); memcpy(oadd, &s, 32); memcpy(osub, &d, 32); memcpy(oneg, &n, 32);
}
void ht_fr_inv(const uint32_t *a, uint32_t *o) { Fr x; memcpy(&x, a, 32); Fr r = Fr::inv(x); memcpy(o, &r, 32); }
void ht_fp_inv(const uint32_t *a, uint32_t *o) { Fp x; memcpy(&x, a, 32); Fp r = Fp::inv(x); memcpy(o, &r, 32); }
// divstep inversion against Fermat on n values (32 bytes each, already reduced); returns the number of mismatches
int ht_inv_crosscheck(const uint32_t *vals, int n, int which) {
    int bad = 0;
    for (int i = 0; i < n; i++) {
        if (which == 0) { Fp x; memcpy(&x, vals + 8 * i, 32); bad += !(Fp::inv(x) == Fp::inv_fermat(x)); }
        else { Fr x; memcpy(&x, vals + 8 * i, 32); bad += !(Fr::inv(x) == Fr::inv_fermat(x)); }
    }
    return bad;
}
// sum_i (+/-) p_i accumulated with add_affine, then a general add of the accumulator with itself-shifted copy
void ht_g1_accumulate(const uint32_t *pts, const uint8_t *neg, int n, uint32_t *out_aff) {
    G1XYZZ acc = G1XYZZ::inf();
    for (int i = 0; i < n; i++) { G1Affine p; memcpy(&p, pts + 16 * i, 64); acc.add_affine(p, neg[i]); }
    G1Affine r = acc.to_affine(); memcpy(out_aff, &r, 64);
}
void ht_g2_accumulate(const uint32_t *pts, const uint8_t *neg, int n, uint32_t *out_aff) {
    G2XYZZ acc = G2XYZZ::inf();
    for (int i = 0; i < n; i++) { G2Affine p; memcpy(&p, pts + 32 * i, 128); acc.add_affine(p, neg[i]); }
    G2Affine r = acc.to_affine(); memcpy(out_aff, &r, 128);
}
void ht_g1_add_mul(const uint32_t *p, const uint32_t *q, const uint32_t *k, uint32_t *out_add, uint32_t *out_mul, uint32_t *out_dbl) {
    G1Affine a, b; memcpy(&a, p, 64); memcpy(&b, q, 64);
    G1XYZZ x = G1XYZZ::from_affine(a), y = G1XYZZ::from_affine(b);
    G1XYZZ s = x; s.add(y); G1Affine r = s.to_affine(); memcpy(out_add, &r, 64);
    r = x.mul_256(k).to_affine(); memcpy(out_mul, &r, 64);
    G1XYZZ d = x; d.add(x); r = d.to_affine(); memcpy(out_dbl, &r, 64);
}
void ht_g2_add_mul(const uint32_t *p, const uint32_t *q, const uint32_t *k, uint32_t *out_add, uint32_t *out_mul, uint32_t *out_dbl) {
    G2Affine a, b; memcpy(&a, p, 128); memcpy(&b, q, 128);
    G2XYZZ x = G2XYZZ::from_affine(a), y = G2XYZZ::from_affine(b);
    G2XYZZ s = x; s.add(y); G2Affine r = s.to_affine(); memcpy(out_add, &r, 128);
    r = x.mul_256(k).to_affine(); memcpy(out_mul, &r, 128);
    G2XYZZ d = x; d.add(x); r = d.to_affine(); memcpy(out_dbl, &r, 128);
}
// pairing: Miller loop value and e(P, Q) = miller^((q^12-1)/r), both as 12 Fp Montgomery elements in tower order
void ht_pairing(const uint32_t *p, const uint32_t *q, uint32_t *out_miller, uint32_t *out_gt) {
    G1Affine a; G2Affine b; memcpy(&a, p, 64); memcpy(&b, q, 128);
    pairing::Fp12 f = pairing::miller_loop(a, b);
    memcpy(out_miller, &f, 384);
    pairing::Fp12 e = pairing::final_exponentiation(f);
    memcpy(out_gt, &e, 384);
}
void ht_fp12_ops(const uint32_t *a, const uint32_t *b, uint32_t *out_mul, uint32_t *out_sqr, uint32_t *out_inv) {
    pairing::Fp12 x, y; memcpy(&x, a, 384); memcpy(&y, b, 384);
    pairing::Fp12 m = pairing::Fp12::mul(x, y), s = pairing::Fp12::sqr(x), i = pairing::Fp12::inv(x);
    memcpy(out_mul, &m, 384); memcpy(out_sqr, &s, 384); memcpy(out_inv, &i, 384);
}
}
