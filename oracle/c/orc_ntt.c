/* ORACLE (test infrastructure, NOT product code) -- Fr NTT and computeH on the CPU.
 * Restates gnark-crypto v0.14 ecc/bn254/fr/fft (Domain, FFT, FFTInverse, OnCoset; out of tree) and gnark v0.10
 * backend/groth16/bn254/prove.go computeH, both reached from src/prover/prover/prover.go:269.  Conventions in
 * oracle/py/ntt.py, against which this file is validated (tests/test_oracle_c.py). */
#include <stdlib.h>
#include <omp.h>
#include "orc.h"
#include "orc_field.h"

static const uint64_t ROOT_2_28[4] = {   /* 5^((r-1)/2^28), plain */
    0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL, 0x00e0a7eb8ef62abcULL, 0x2a3c09f0a58a7e85ULL};

static void fr_from_u64(fe *z, uint64_t v) { fe t = {{v, 0, 0, 0}}; fe_to_mont(z, &t, &ORC_FR); }

static void domain_gen(fe *g, int logn) {
    fe t; memcpy(t.l, ROOT_2_28, 32); fe_to_mont(&t, &t, &ORC_FR);
    for (int i = logn; i < 28; i++) fr_sqr(&t, &t);
    *g = t;
}

static size_t bitrev(size_t i, int logn) {
    size_t r = 0;
    for (int b = 0; b < logn; b++) { r = (r << 1) | (i & 1); i >>= 1; }
    return r;
}

/* tw[j] = w^j for j < n/2 */
static fe *twiddles(const fe *w, size_t half) {
    fe *tw = (fe *)malloc(sizeof(fe) * (half ? half : 1));
    fr_one(&tw[0]);
    for (size_t j = 1; j < half; j++) fr_mul(&tw[j], &tw[j - 1], w);
    return tw;
}

static void dif(fe *a, int logn, const fe *tw, int threads) {
    size_t n = (size_t)1 << logn;
    for (size_t m = n; m > 1; m >>= 1) {
        size_t half = m >> 1, step = n / m;
        #pragma omp parallel for schedule(static) num_threads(threads)
        for (size_t idx = 0; idx < n / 2; idx++) {
            size_t start = (idx / half) * m, j = idx % half;
            fe u = a[start + j], v = a[start + j + half], d;
            fr_add(&a[start + j], &u, &v);
            fr_sub(&d, &u, &v);
            fr_mul(&a[start + j + half], &d, &tw[j * step]);
        }
    }
}
static void dit(fe *a, int logn, const fe *tw, int threads) {
    size_t n = (size_t)1 << logn;
    for (size_t m = 2; m <= n; m <<= 1) {
        size_t half = m >> 1, step = n / m;
        #pragma omp parallel for schedule(static) num_threads(threads)
        for (size_t idx = 0; idx < n / 2; idx++) {
            size_t start = (idx / half) * m, j = idx % half;
            fe u = a[start + j], v;
            fr_mul(&v, &a[start + j + half], &tw[j * step]);
            fr_add(&a[start + j], &u, &v);
            fr_sub(&a[start + j + half], &u, &v);
        }
    }
}

/* a[pos] *= base^(exponent(pos)) * extra, exponent = pos (natural) or bitrev(pos) */
static void scale_pow(fe *a, int logn, const fe *base, const fe *extra, int bitrev_idx, int threads) {
    size_t n = (size_t)1 << logn;
    fe *pw = (fe *)malloc(sizeof(fe) * n);
    pw[0] = *extra;
    for (size_t i = 1; i < n; i++) fr_mul(&pw[i], &pw[i - 1], base);
    #pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &pw[bitrev_idx ? bitrev(i, logn) : i]);
    free(pw);
}

void orc_ntt(uint64_t *data, int logn, int inverse, int is_dit, int coset, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    fe *a = (fe *)data;
    size_t n = (size_t)1 << logn;
    fe g, w, one, five, five_inv, ninv;
    domain_gen(&g, logn); fr_one(&one); fr_from_u64(&five, 5); fr_inv(&five_inv, &five);
    fr_from_u64(&ninv, (uint64_t)n); fr_inv(&ninv, &ninv);
    if (inverse) fr_inv(&w, &g); else w = g;
    fe *tw = twiddles(&w, n / 2);
    if (!inverse && coset) scale_pow(a, logn, &five, &one, is_dit, threads);      /* DIF: natural in; DIT: bit-reversed in */
    if (is_dit) dit(a, logn, tw, threads); else dif(a, logn, tw, threads);
    if (inverse) {
        if (coset) scale_pow(a, logn, &five_inv, &ninv, !is_dit, threads);         /* DIF: bit-reversed out */
        else scale_pow(a, logn, &one, &ninv, 0, threads);
    }
    free(tw);
}

void orc_compute_h(const uint64_t *a_in, const uint64_t *b_in, const uint64_t *c_in, size_t m, int logn, uint64_t *out_h, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    size_t n = (size_t)1 << logn;
    fe *a = (fe *)out_h, *b = (fe *)calloc(n, sizeof(fe)), *c = (fe *)calloc(n, sizeof(fe));
    memset(a, 0, n * sizeof(fe));
    memcpy(a, a_in, m * sizeof(fe)); memcpy(b, b_in, m * sizeof(fe)); memcpy(c, c_in, m * sizeof(fe));
    fe *v[3] = {a, b, c};
    for (int k = 0; k < 3; k++) { orc_ntt((uint64_t *)v[k], logn, 1, 0, 0, threads); orc_ntt((uint64_t *)v[k], logn, 0, 1, 1, threads); }
    fe den, five, one; fr_from_u64(&five, 5); fr_one(&one);
    den = five; for (int i = 0; i < logn; i++) fr_sqr(&den, &den);     /* 5^n */
    fr_sub(&den, &den, &one); fr_inv(&den, &den);
    #pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < n; i++) { fe t; fr_mul(&t, &a[i], &b[i]); fr_sub(&t, &t, &c[i]); fr_mul(&a[i], &t, &den); }
    orc_ntt((uint64_t *)a, logn, 1, 0, 1, threads);
    free(b); free(c);
}
