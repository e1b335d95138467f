/* ORACLE (test infrastructure, NOT product code) -- BN254 Fp / Fr / Fp2 arithmetic on the CPU.
 *
 * Restates gnark-crypto's fp.Element / fr.Element (bnb-chain/gnark-crypto v0.14.1-0.20240910145340-609ab3a7eb9b,
 * ecc/bn254/{fp,fr}, pinned at /root/reference/go.mod:57-60; source NOT in /root/reference): 4 x u64
 * little-endian limbs, Montgomery form with R = 2^256.  Validated against oracle/py/bn254.py (Python big ints),
 * which is itself pinned by the EIP-196 vectors (tests/test_oracle_kat.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 */
#ifndef ORC_FIELD_H
#define ORC_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;         /* one field element, Montgomery form */
typedef struct { fe a0, a1; } fe2;            /* Fp2 = Fp[u]/(u^2+1), a0 + a1*u     */

typedef struct {
    uint64_t m[4];    /* modulus                          */
    uint64_t inv;     /* -m^-1 mod 2^64                   */
    uint64_t r2[4];   /* R^2 mod m (to-Montgomery factor) */
    uint64_t one[4];  /* R mod m                          */
} fparams;

extern const fparams ORC_FP, ORC_FR;

static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) { return memcmp(a, b, sizeof(fe)) == 0; }

static inline int fe_geq_m(const uint64_t *t, const uint64_t *m) {
    for (int i = 3; i >= 0; i--) { if (t[i] > m[i]) return 1; if (t[i] < m[i]) return 0; }
    return 1;
}
static inline void fe_sub_m(uint64_t *t, const uint64_t *m) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - m[i] - (uint64_t)b; t[i] = (uint64_t)d; b = (d >> 64) & 1; }
}

static inline void fe_add(fe *z, const fe *x, const fe *y, const fparams *P) {
    /* both moduli are < 2^254: the sum fits four words; one branch-free conditional subtraction */
    u128 c = (u128)x->l[0] + y->l[0]; const uint64_t t0 = (uint64_t)c;
    c = (u128)x->l[1] + y->l[1] + (uint64_t)(c >> 64); const uint64_t t1 = (uint64_t)c;
    c = (u128)x->l[2] + y->l[2] + (uint64_t)(c >> 64); const uint64_t t2 = (uint64_t)c;
    c = (u128)x->l[3] + y->l[3] + (uint64_t)(c >> 64); const uint64_t t3 = (uint64_t)c;
    u128 b = (u128)t0 - P->m[0]; const uint64_t r0 = (uint64_t)b;
    b = (u128)t1 - P->m[1] - (uint64_t)((b >> 64) & 1); const uint64_t r1 = (uint64_t)b;
    b = (u128)t2 - P->m[2] - (uint64_t)((b >> 64) & 1); const uint64_t r2 = (uint64_t)b;
    b = (u128)t3 - P->m[3] - (uint64_t)((b >> 64) & 1); const uint64_t r3 = (uint64_t)b;
    const int keep = (int)((b >> 64) & 1);
    z->l[0] = keep ? t0 : r0; z->l[1] = keep ? t1 : r1; z->l[2] = keep ? t2 : r2; z->l[3] = keep ? t3 : r3;
}
static inline void fe_sub(fe *z, const fe *x, const fe *y, const fparams *P) {
    u128 b = (u128)x->l[0] - y->l[0]; const uint64_t t0 = (uint64_t)b;
    b = (u128)x->l[1] - y->l[1] - (uint64_t)((b >> 64) & 1); const uint64_t t1 = (uint64_t)b;
    b = (u128)x->l[2] - y->l[2] - (uint64_t)((b >> 64) & 1); const uint64_t t2 = (uint64_t)b;
    b = (u128)x->l[3] - y->l[3] - (uint64_t)((b >> 64) & 1); const uint64_t t3 = (uint64_t)b;
    const uint64_t mask = (uint64_t)0 - (uint64_t)((b >> 64) & 1);   /* borrow: add the modulus back */
    u128 c = (u128)t0 + (P->m[0] & mask); z->l[0] = (uint64_t)c;
    c = (u128)t1 + (P->m[1] & mask) + (uint64_t)(c >> 64); z->l[1] = (uint64_t)c;
    c = (u128)t2 + (P->m[2] & mask) + (uint64_t)(c >> 64); z->l[2] = (uint64_t)c;
    c = (u128)t3 + (P->m[3] & mask) + (uint64_t)(c >> 64); z->l[3] = (uint64_t)c;
}
static inline void fe_neg(fe *z, const fe *x, const fparams *P) {
    if (fe_is_zero(x)) { *z = *x; return; }
    fe m; memcpy(m.l, P->m, 32); fe_sub(z, &m, x, P);
}
static inline void fe_dbl(fe *z, const fe *x, const fparams *P) { fe_add(z, x, x, P); }

/* CIOS Montgomery product z = x*y/R mod m, fully unrolled, with gnark-crypto's "no-carry" shortcut (field/generator: valid when the top
 * word of the modulus is < 2^63 - 1, as for both BN254 moduli): the two carry words of textbook CIOS collapse into t3 = C + A. */
static inline void fe_mul(fe *z, const fe *x, const fe *y, const fparams *P) {
    const uint64_t x0 = x->l[0], x1 = x->l[1], x2 = x->l[2], x3 = x->l[3];
    const uint64_t q0 = P->m[0], q1 = P->m[1], q2 = P->m[2], q3 = P->m[3], inv = P->inv;
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#pragma GCC unroll 4
    for (int i = 0; i < 4; i++) {
        const uint64_t yi = y->l[i];
        u128 a = (u128)x0 * yi + t0;
        uint64_t A = (uint64_t)(a >> 64);
        const uint64_t m = (uint64_t)a * inv;
        u128 c = (u128)m * q0 + (uint64_t)a;
        uint64_t C = (uint64_t)(c >> 64);
        a = (u128)x1 * yi + t1 + A; A = (uint64_t)(a >> 64); c = (u128)m * q1 + (uint64_t)a + C; t0 = (uint64_t)c; C = (uint64_t)(c >> 64);
        a = (u128)x2 * yi + t2 + A; A = (uint64_t)(a >> 64); c = (u128)m * q2 + (uint64_t)a + C; t1 = (uint64_t)c; C = (uint64_t)(c >> 64);
        a = (u128)x3 * yi + t3 + A; A = (uint64_t)(a >> 64); c = (u128)m * q3 + (uint64_t)a + C; t2 = (uint64_t)c; C = (uint64_t)(c >> 64);
        t3 = C + A;
    }
    /* one conditional subtraction, branch-free */
    u128 b = (u128)t0 - q0; const uint64_t r0 = (uint64_t)b;
    b = (u128)t1 - q1 - (uint64_t)((b >> 64) & 1); const uint64_t r1 = (uint64_t)b;
    b = (u128)t2 - q2 - (uint64_t)((b >> 64) & 1); const uint64_t r2 = (uint64_t)b;
    b = (u128)t3 - q3 - (uint64_t)((b >> 64) & 1); const uint64_t r3 = (uint64_t)b;
    const int keep = (int)((b >> 64) & 1);           /* borrow: t < q */
    z->l[0] = keep ? t0 : r0; z->l[1] = keep ? t1 : r1; z->l[2] = keep ? t2 : r2; z->l[3] = keep ? t3 : r3;
}
static inline void fe_sqr(fe *z, const fe *x, const fparams *P) { fe_mul(z, x, x, P); }

static inline void fe_to_mont(fe *z, const fe *x, const fparams *P) { fe r2; memcpy(r2.l, P->r2, 32); fe_mul(z, x, &r2, P); }
static inline void fe_from_mont(fe *z, const fe *x, const fparams *P) { fe one = {{1, 0, 0, 0}}; fe_mul(z, x, &one, P); }
static inline void fe_one(fe *z, const fparams *P) { memcpy(z->l, P->one, 32); }

/* z = x^e, e given as 4 little-endian u64 (plain integer) */
static inline void fe_pow(fe *z, const fe *x, const uint64_t e[4], const fparams *P) {
    fe acc; fe_one(&acc, P);
    for (int i = 255; i >= 0; i--) {
        fe_sqr(&acc, &acc, P);
        if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, x, P);
    }
    *z = acc;
}
static inline void fe_inv(fe *z, const fe *x, const fparams *P) {   /* Fermat: x^(m-2); 0 -> 0 */
    uint64_t e[4]; memcpy(e, P->m, 32); e[0] -= 2;                  /* both moduli end in ...47 / ...01: no borrow */
    fe_pow(z, x, e, P);
}

/* ---- named instances ---- */
#define FPF(name) static inline void fp_##name
static inline void fp_add(fe *z, const fe *x, const fe *y) { fe_add(z, x, y, &ORC_FP); }
static inline void fp_sub(fe *z, const fe *x, const fe *y) { fe_sub(z, x, y, &ORC_FP); }
static inline void fp_mul(fe *z, const fe *x, const fe *y) { fe_mul(z, x, y, &ORC_FP); }
static inline void fp_sqr(fe *z, const fe *x) { fe_mul(z, x, x, &ORC_FP); }
static inline void fp_neg(fe *z, const fe *x) { fe_neg(z, x, &ORC_FP); }
static inline void fp_inv(fe *z, const fe *x) { fe_inv(z, x, &ORC_FP); }
static inline void fp_one(fe *z) { fe_one(z, &ORC_FP); }
static inline int fp_is_zero(const fe *x) { return fe_is_zero(x); }
static inline int fp_eq(const fe *x, const fe *y) { return fe_eq(x, y); }

static inline void fr_add(fe *z, const fe *x, const fe *y) { fe_add(z, x, y, &ORC_FR); }
static inline void fr_sub(fe *z, const fe *x, const fe *y) { fe_sub(z, x, y, &ORC_FR); }
static inline void fr_mul(fe *z, const fe *x, const fe *y) { fe_mul(z, x, y, &ORC_FR); }
static inline void fr_sqr(fe *z, const fe *x) { fe_mul(z, x, x, &ORC_FR); }
static inline void fr_inv(fe *z, const fe *x) { fe_inv(z, x, &ORC_FR); }
static inline void fr_one(fe *z) { fe_one(z, &ORC_FR); }

/* ---- Fp2 ---- */
static inline void fp2_add(fe2 *z, const fe2 *x, const fe2 *y) { fp_add(&z->a0, &x->a0, &y->a0); fp_add(&z->a1, &x->a1, &y->a1); }
static inline void fp2_sub(fe2 *z, const fe2 *x, const fe2 *y) { fp_sub(&z->a0, &x->a0, &y->a0); fp_sub(&z->a1, &x->a1, &y->a1); }
static inline void fp2_neg(fe2 *z, const fe2 *x) { fp_neg(&z->a0, &x->a0); fp_neg(&z->a1, &x->a1); }
static inline void fp2_mul(fe2 *z, const fe2 *x, const fe2 *y) {
    fe t0, t1, s0, s1, m;                      /* Karatsuba: (a0b0 - a1b1) + ((a0+a1)(b0+b1) - a0b0 - a1b1) u */
    fp_mul(&t0, &x->a0, &y->a0); fp_mul(&t1, &x->a1, &y->a1);
    fp_add(&s0, &x->a0, &x->a1); fp_add(&s1, &y->a0, &y->a1);
    fp_mul(&m, &s0, &s1);
    fp_sub(&z->a0, &t0, &t1);
    fp_sub(&m, &m, &t0); fp_sub(&z->a1, &m, &t1);
}
static inline void fp2_sqr(fe2 *z, const fe2 *x) { fe2 t = *x; fp2_mul(z, &t, &t); }
static inline void fp2_inv(fe2 *z, const fe2 *x) {
    fe n0, n1, d; fp_sqr(&n0, &x->a0); fp_sqr(&n1, &x->a1); fp_add(&d, &n0, &n1); fp_inv(&d, &d);
    fp_mul(&z->a0, &x->a0, &d); fp_mul(&n1, &x->a1, &d); fp_neg(&z->a1, &n1);
}
static inline void fp2_one(fe2 *z) { fp_one(&z->a0); memset(&z->a1, 0, sizeof(fe)); }
static inline int fp2_is_zero(const fe2 *x) { return fe_is_zero(&x->a0) && fe_is_zero(&x->a1); }
static inline int fp2_eq(const fe2 *x, const fe2 *y) { return fe_eq(&x->a0, &y->a0) && fe_eq(&x->a1, &y->a1); }

#endif
