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

/* ---- full-size check helpers (tests of the 2^26 configurations): O(n) field work, threaded ---- */
/* out_sum = sum_i v_i, out_isum = sum_i i*v_i, the limb patterns v_i taken as integers mod r */
void orc_fr_index_sums(const uint64_t *v, size_t n, uint64_t *out_sum, uint64_t *out_isum, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    fe S, I; memset(&S, 0, sizeof S); memset(&I, 0, sizeof I);
    #pragma omp parallel num_threads(threads)
    {
        fe s, t; memset(&s, 0, sizeof s); memset(&t, 0, sizeof t);
        #pragma omp for schedule(static)
        for (size_t i = 0; i < n; i++) {
            const fe *x = (const fe *)(v + 4 * i);
            fe im, p; fr_from_u64(&im, (uint64_t)i);
            fr_add(&s, &s, x); fr_mul(&p, x, &im); fr_add(&t, &t, &p);
        }
        #pragma omp critical
        { fr_add(&S, &S, &s); fr_add(&I, &I, &t); }
    }
    memcpy(out_sum, &S, 32); memcpy(out_isum, &I, 32);
}
/* P(x0) for P given by its evaluations on the size-2^logn subgroup (Montgomery in/out):
 * P(x0) = (x0^n - 1)/n * sum_k e_k w^k / (x0 - w^k); x0 must not be in the subgroup.  Batch inversion per thread. */
void orc_eval_barycentric(const uint64_t *evals, size_t m, int logn, const uint64_t *x0, uint64_t *out, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    size_t n = (size_t)1 << logn;
    fe g, x, acc_total; domain_gen(&g, logn); memcpy(&x, x0, 32); memset(&acc_total, 0, sizeof acc_total);
    #pragma omp parallel num_threads(threads)
    {
        int T = omp_get_num_threads(), t = omp_get_thread_num();
        size_t per = (m + T - 1) / T, lo = (size_t)t * per, hi = lo + per > m ? m : lo + per;
        fe acc; memset(&acc, 0, sizeof acc);
        if (lo < hi) {
            enum { B = 1024 };
            fe wk, e; fe_one(&wk, &ORC_FR);
            uint64_t ex[4] = {lo, 0, 0, 0}; fe_pow(&wk, &g, ex, &ORC_FR);
            fe den[B], pre[B], num[B];
            for (size_t s0 = lo; s0 < hi; s0 += B) {
                size_t cnt = hi - s0 < B ? hi - s0 : B;
                fe run; fr_one(&run);
                for (size_t j = 0; j < cnt; j++) {
                    fr_sub(&den[j], &x, &wk); pre[j] = run; fr_mul(&run, &run, &den[j]);
                    fr_mul(&num[j], (const fe *)(evals + 4 * (s0 + j)), &wk); fr_mul(&wk, &wk, &g);
                }
                fe inv; fr_inv(&inv, &run);
                for (size_t j = cnt; j-- > 0;) { fr_mul(&e, &inv, &pre[j]); fr_mul(&inv, &inv, &den[j]); fr_mul(&e, &e, &num[j]); fr_add(&acc, &acc, &e); }
            }
        }
        #pragma omp critical
        fr_add(&acc_total, &acc_total, &acc);
    }
    fe xn = x, one, ninv; fr_one(&one);
    for (int i = 0; i < logn; i++) fr_sqr(&xn, &xn);
    fr_sub(&xn, &xn, &one); fr_from_u64(&ninv, (uint64_t)n); fr_inv(&ninv, &ninv);
    fr_mul(&acc_total, &acc_total, &xn); fr_mul(&acc_total, &acc_total, &ninv);
    memcpy(out, &acc_total, 32);
}
/* sum_i c_{bitrev(i)}... : evaluates the polynomial whose coefficient bitrev(pos) is stored at pos (computeH output) */
void orc_poly_eval_bitrev(const uint64_t *coef, int logn, const uint64_t *x0, uint64_t *out, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    size_t n = (size_t)1 << logn;
    fe x, total; memcpy(&x, x0, 32); memset(&total, 0, sizeof total);
    /* thread t takes coefficient range [lo, hi): Horner on the range, times x^lo */
    #pragma omp parallel num_threads(threads)
    {
        int T = omp_get_num_threads(), t = omp_get_thread_num();
        size_t per = (n + T - 1) / T, lo = (size_t)t * per, hi = lo + per > n ? n : lo + per;
        if (lo < hi) {
            fe acc; memset(&acc, 0, sizeof acc);
            for (size_t i = hi; i-- > lo;) { fr_mul(&acc, &acc, &x); fr_add(&acc, &acc, (const fe *)(coef + 4 * bitrev(i, logn))); }
            fe xp; uint64_t ex[4] = {lo, 0, 0, 0}; fe_pow(&xp, &x, ex, &ORC_FR);
            fr_mul(&acc, &acc, &xp);
            #pragma omp critical
            fr_add(&total, &total, &acc);
        }
    }
    memcpy(out, &total, 32);
}
