/* ORACLE (test infrastructure) -- short-Weierstrass group law y^2 = x^3 + b over a base field, written once and
 * instantiated for G1 (Fp) and G2 (Fp2) by orc_curve.c.  Restates the published formulas (EFD "dbl-2009-l",
 * "add-2007-bl", "madd-2007-bl") that gnark-crypto's G1Jac/G2Jac also implement (ecc/bn254/g1.go, g2.go, out of
 * tree); results are compared in affine form, so the choice of projective formulas does not affect parity.
 *
 * Before including define:  CV(name)  FT  F_add F_sub F_mul F_sqr F_neg F_inv F_one F_is_zero F_eq
 */

typedef struct { FT x, y; } CV(aff);          /* (0,0) = infinity, as gnark-crypto stores it */
typedef struct { FT x, y, z; } CV(jac);       /* z = 0 = infinity */

static inline int CV(aff_is_inf)(const CV(aff) *p) { return F_is_zero(&p->x) && F_is_zero(&p->y); }
static inline void CV(jac_set_inf)(CV(jac) *p) { F_one(&p->x); F_one(&p->y); memset(&p->z, 0, sizeof(FT)); }
static inline int CV(jac_is_inf)(const CV(jac) *p) { return F_is_zero(&p->z); }

static inline void CV(jac_from_aff)(CV(jac) *r, const CV(aff) *p) {
    if (CV(aff_is_inf)(p)) { CV(jac_set_inf)(r); return; }
    r->x = p->x; r->y = p->y; F_one(&r->z);
}

static void CV(jac_double)(CV(jac) *r, const CV(jac) *p) {
    if (CV(jac_is_inf)(p)) { *r = *p; return; }
    FT A, B, C, D, E, Fq, t, X3, Y3, Z3;
    F_sqr(&A, &p->x); F_sqr(&B, &p->y); F_sqr(&C, &B);
    F_add(&t, &p->x, &B); F_sqr(&t, &t); F_sub(&t, &t, &A); F_sub(&t, &t, &C); F_add(&D, &t, &t);
    F_add(&E, &A, &A); F_add(&E, &E, &A);
    F_sqr(&Fq, &E);
    F_add(&t, &D, &D); F_sub(&X3, &Fq, &t);
    F_add(&C, &C, &C); F_add(&C, &C, &C); F_add(&C, &C, &C);
    F_sub(&t, &D, &X3); F_mul(&t, &E, &t); F_sub(&Y3, &t, &C);
    F_mul(&Z3, &p->y, &p->z); F_add(&Z3, &Z3, &Z3);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static void CV(jac_add)(CV(jac) *r, const CV(jac) *p, const CV(jac) *q) {
    if (CV(jac_is_inf)(p)) { *r = *q; return; }
    if (CV(jac_is_inf)(q)) { *r = *p; return; }
    FT Z1Z1, Z2Z2, U1, U2, S1, S2, H, Rr, HH, HHH, V, t, X3, Y3, Z3;
    F_sqr(&Z1Z1, &p->z); F_sqr(&Z2Z2, &q->z);
    F_mul(&U1, &p->x, &Z2Z2); F_mul(&U2, &q->x, &Z1Z1);
    F_mul(&S1, &p->y, &q->z); F_mul(&S1, &S1, &Z2Z2);
    F_mul(&S2, &q->y, &p->z); F_mul(&S2, &S2, &Z1Z1);
    if (F_eq(&U1, &U2)) {
        if (F_eq(&S1, &S2)) { CV(jac_double)(r, p); } else { CV(jac_set_inf)(r); }
        return;
    }
    F_sub(&H, &U2, &U1); F_sub(&Rr, &S2, &S1);
    F_sqr(&HH, &H); F_mul(&HHH, &H, &HH); F_mul(&V, &U1, &HH);
    F_sqr(&X3, &Rr); F_sub(&X3, &X3, &HHH); F_add(&t, &V, &V); F_sub(&X3, &X3, &t);
    F_sub(&t, &V, &X3); F_mul(&Y3, &Rr, &t); F_mul(&t, &S1, &HHH); F_sub(&Y3, &Y3, &t);
    F_mul(&Z3, &p->z, &q->z); F_mul(&Z3, &Z3, &H);
    r->x = X3; r->y = Y3; r->z = Z3;
}

/* r = p + q with q affine (the bucket-accumulation step of Pippenger) */
static void CV(jac_add_mixed)(CV(jac) *r, const CV(jac) *p, const CV(aff) *q) {
    if (CV(aff_is_inf)(q)) { *r = *p; return; }
    if (CV(jac_is_inf)(p)) { CV(jac_from_aff)(r, q); return; }
    FT Z1Z1, U2, S2, H, Rr, HH, HHH, V, t, X3, Y3, Z3;
    F_sqr(&Z1Z1, &p->z);
    F_mul(&U2, &q->x, &Z1Z1);
    F_mul(&S2, &q->y, &p->z); F_mul(&S2, &S2, &Z1Z1);
    if (F_eq(&p->x, &U2)) {
        if (F_eq(&p->y, &S2)) { CV(jac_double)(r, p); } else { CV(jac_set_inf)(r); }
        return;
    }
    F_sub(&H, &U2, &p->x); F_sub(&Rr, &S2, &p->y);
    F_sqr(&HH, &H); F_mul(&HHH, &H, &HH); F_mul(&V, &p->x, &HH);
    F_sqr(&X3, &Rr); F_sub(&X3, &X3, &HHH); F_add(&t, &V, &V); F_sub(&X3, &X3, &t);
    F_sub(&t, &V, &X3); F_mul(&Y3, &Rr, &t); F_mul(&t, &p->y, &HHH); F_sub(&Y3, &Y3, &t);
    F_mul(&Z3, &p->z, &H);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static void CV(jac_to_aff)(CV(aff) *r, const CV(jac) *p) {
    if (CV(jac_is_inf)(p)) { memset(r, 0, sizeof(*r)); return; }
    FT zi, zi2, zi3;
    F_inv(&zi, &p->z); F_sqr(&zi2, &zi); F_mul(&zi3, &zi2, &zi);
    F_mul(&r->x, &p->x, &zi2); F_mul(&r->y, &p->y, &zi3);
}

/* r = k * p, k a plain 256-bit integer (4 LE u64) */
static void CV(jac_mul)(CV(jac) *r, const CV(jac) *p, const uint64_t k[4]) {
    CV(jac) acc; CV(jac_set_inf)(&acc);
    for (int i = 255; i >= 0; i--) {
        CV(jac_double)(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) CV(jac_add)(&acc, &acc, p);
    }
    *r = acc;
}

/* Montgomery batch normalisation of n Jacobian points */
static void CV(batch_to_aff)(CV(aff) *out, const CV(jac) *in, size_t n) {
    FT *pre = (FT *)malloc(sizeof(FT) * (n + 1));
    FT acc; F_one(&acc);
    for (size_t i = 0; i < n; i++) { pre[i] = acc; if (!CV(jac_is_inf)(&in[i])) F_mul(&acc, &acc, &in[i].z); }
    FT inv; F_inv(&inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (CV(jac_is_inf)(&in[i])) { memset(&out[i], 0, sizeof(out[i])); continue; }
        FT zi, zi2, zi3; F_mul(&zi, &inv, &pre[i]); F_mul(&inv, &inv, &in[i].z);
        F_sqr(&zi2, &zi); F_mul(&zi3, &zi2, &zi);
        F_mul(&out[i].x, &in[i].x, &zi2); F_mul(&out[i].y, &in[i].y, &zi3);
    }
    free(pre);
}

/* Extended Jacobian coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ = 0: infinity) -- what gnark-crypto's MultiExp keeps its
 * buckets in (g1JacExtended, ecc/bn254/g1.go, out of tree): a mixed addition is 8M + 2S (EFD "madd-2008-s"). */
typedef struct { FT x, y, zz, zzz; } CV(xyzz);
static inline void CV(xyzz_set_inf)(CV(xyzz) *p) { memset(p, 0, sizeof(*p)); }
static inline int CV(xyzz_is_inf)(const CV(xyzz) *p) { return F_is_zero(&p->zz); }

/* p += (neg ? -q : q), q affine */
static inline void CV(xyzz_add_mixed)(CV(xyzz) *p, const CV(aff) *q, int neg) {
    if (CV(aff_is_inf)(q)) return;
    FT qy = q->y;
    if (neg) F_neg(&qy, &qy);
    if (CV(xyzz_is_inf)(p)) { p->x = q->x; p->y = qy; F_one(&p->zz); F_one(&p->zzz); return; }
    FT U2, S2, Pp, Rr, PP, PPP, Q, t, X3;
    F_mul(&U2, &q->x, &p->zz); F_mul(&S2, &qy, &p->zzz);
    F_sub(&Pp, &U2, &p->x); F_sub(&Rr, &S2, &p->y);
    if (F_is_zero(&Pp)) {
        if (!F_is_zero(&Rr)) { CV(xyzz_set_inf)(p); return; }
        /* doubling of the affine point ("mdbl-2008-s") */
        FT U, V, W, S, M;
        F_add(&U, &qy, &qy); F_sqr(&V, &U); F_mul(&W, &U, &V); F_mul(&S, &q->x, &V);
        F_sqr(&M, &q->x); F_add(&t, &M, &M); F_add(&M, &M, &t);
        F_sqr(&X3, &M); F_sub(&X3, &X3, &S); F_sub(&X3, &X3, &S);
        F_sub(&t, &S, &X3); F_mul(&t, &M, &t); F_mul(&S, &W, &qy); F_sub(&p->y, &t, &S);
        p->x = X3; p->zz = V; p->zzz = W;
        return;
    }
    F_sqr(&PP, &Pp); F_mul(&PPP, &Pp, &PP); F_mul(&Q, &p->x, &PP);
    F_sqr(&X3, &Rr); F_sub(&X3, &X3, &PPP); F_sub(&X3, &X3, &Q); F_sub(&X3, &X3, &Q);
    F_sub(&t, &Q, &X3); F_mul(&t, &Rr, &t); F_mul(&Q, &p->y, &PPP); F_sub(&p->y, &t, &Q);
    p->x = X3;
    F_mul(&p->zz, &p->zz, &PP); F_mul(&p->zzz, &p->zzz, &PPP);
}
/* the same point in Jacobian coordinates: Z = ZZZ, X' = X*ZZ^2, Y' = Y*ZZZ^2 */
static inline void CV(xyzz_to_jac)(CV(jac) *r, const CV(xyzz) *p) {
    if (CV(xyzz_is_inf)(p)) { CV(jac_set_inf)(r); return; }
    FT t;
    F_sqr(&t, &p->zz); F_mul(&r->x, &p->x, &t);
    F_sqr(&t, &p->zzz); F_mul(&r->y, &p->y, &t);
    r->z = p->zzz;
}

/* Pippenger bucket method with signed digits, as gnark-crypto's MultiExp does it (ecc/bn254/multiexp.go, out of tree: c-bit windows,
 * digits in [-2^(c-1), 2^(c-1)), 2^(c-1) extended-Jacobian buckets per window, one running-sum reduction per window).  The digits come
 * without a carry chain from s' = s + K, K = sum over all windows but the top one of 2^(c-1) * 2^(wc): digit_w = window_w(s') - 2^(c-1)
 * (the top window is taken as it is; its value is small because s < 2^254).  One (window, chunk of points) task at a time so that the
 * caller can spread tasks over OpenMP threads. */
static inline uint64_t CV(window_of)(const uint64_t *s, int bit, int c) {
    const int limb = bit >> 6, off = bit & 63;
    if (limb >= 4) return 0;
    uint64_t d = s[limb] >> off;
    if (off + c > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - off);
    return d & (((uint64_t)1 << c) - 1);
}
static void CV(msm_window_chunk)(CV(jac) *out, const CV(aff) *pts, const uint64_t *sc /* n x 4, already offset by K */, size_t lo, size_t hi,
                                 int w, int c, int top) {
    const size_t half = (size_t)1 << (c - 1), nb = half + 1;
    CV(xyzz) *bk = (CV(xyzz) *)calloc(nb, sizeof(CV(xyzz)));
    const int bit = w * c;
    for (size_t i = lo; i < hi; i++) {
        const int64_t raw = (int64_t)CV(window_of)(sc + 4 * i, bit, c);
        const int64_t d = top ? raw : raw - (int64_t)half;
        if (i + 8 < hi) __builtin_prefetch(&pts[i + 8]);
        if (d > 0) CV(xyzz_add_mixed)(&bk[d - 1], &pts[i], 0);
        else if (d < 0) CV(xyzz_add_mixed)(&bk[-d - 1], &pts[i], 1);
    }
    CV(jac) run, sum, b; CV(jac_set_inf)(&run); CV(jac_set_inf)(&sum);
    for (size_t i = nb; i-- > 0;) {
        if (!CV(xyzz_is_inf)(&bk[i])) { CV(xyzz_to_jac)(&b, &bk[i]); CV(jac_add)(&run, &run, &b); }
        CV(jac_add)(&sum, &sum, &run);
    }
    free(bk);
    *out = sum;
}

static void CV(msm)(CV(jac) *out, const CV(aff) *pts, const uint64_t *sc /* plain */, size_t n, int threads) {
    if (n == 0) { CV(jac_set_inf)(out); return; }
    int c = 4;
    { size_t t = n; int lg = 0; while (t >>= 1) lg++; c = lg <= 6 ? 3 : lg <= 10 ? 7 : lg <= 14 ? 11 : lg <= 18 ? 13 : lg <= 21 ? 15 : 16; }
    const int nw = (255 + c - 1) / c;
    /* s' = s + K */
    uint64_t K[4] = {0, 0, 0, 0};
    for (int w = 0; w < nw - 1; w++) { const int b = w * c + c - 1; K[b >> 6] |= (uint64_t)1 << (b & 63); }
    uint64_t *sk = (uint64_t *)malloc(32 * n);
    #pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < n; i++) {
        u128 cy = 0;
        for (int j = 0; j < 4; j++) { cy += (u128)sc[4 * i + j] + K[j]; sk[4 * i + j] = (uint64_t)cy; cy >>= 64; }
    }
    /* tasks = windows x point chunks: about three tasks per thread keep a dynamic schedule balanced (16 windows on 16 threads
       with one chunk each would leave most threads idle whenever one window is slower) */
    int chunks = 1;
    if (threads > 1 && n >= (1u << 14)) chunks = (3 * threads + nw - 1) / nw;
    size_t per = (n + chunks - 1) / chunks;
    int ntask = nw * chunks;
    CV(jac) *part = (CV(jac) *)malloc(sizeof(CV(jac)) * ntask);
    #pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int t = 0; t < ntask; t++) {
        int w = t / chunks, ch = t % chunks;
        size_t lo = (size_t)ch * per, hi = lo + per; if (hi > n) hi = n; if (lo > hi) lo = hi;
        CV(msm_window_chunk)(&part[t], pts, sk, lo, hi, w, c, w == nw - 1);
    }
    free(sk);
    CV(jac) acc; CV(jac_set_inf)(&acc);
    for (int w = nw - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) CV(jac_double)(&acc, &acc);
        for (int ch = 0; ch < chunks; ch++) CV(jac_add)(&acc, &acc, &part[w * chunks + ch]);
    }
    free(part);
    *out = acc;
}

/* out[i] = k_i * base for plain scalars: 8-bit fixed-base table, mixed adds, one batch normalisation */
static void CV(fixed_base_batch)(CV(aff) *out, const CV(aff) *base, const uint64_t *sc, size_t n, int threads) {
    enum { WB = 8, NWIN = 32, TSZ = 255 };
    CV(aff) *tab = (CV(aff) *)malloc(sizeof(CV(aff)) * NWIN * TSZ);
    CV(jac) *tj = (CV(jac) *)malloc(sizeof(CV(jac)) * NWIN * TSZ);
    CV(jac) wbase; CV(jac_from_aff)(&wbase, base);
    for (int w = 0; w < NWIN; w++) {
        CV(jac) acc = wbase;
        for (int k = 0; k < TSZ; k++) { tj[w * TSZ + k] = acc; CV(jac_add)(&acc, &acc, &wbase); }
        wbase = acc;   /* 256 * previous base */
    }
    CV(batch_to_aff)(tab, tj, (size_t)NWIN * TSZ);
    free(tj);
    CV(jac) *res = (CV(jac) *)malloc(sizeof(CV(jac)) * n);
    #pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < n; i++) {
        CV(jac) acc; CV(jac_set_inf)(&acc);
        const uint8_t *b = (const uint8_t *)(sc + 4 * i);
        for (int w = 0; w < NWIN; w++) if (b[w]) CV(jac_add_mixed)(&acc, &acc, &tab[w * TSZ + b[w] - 1]);
        res[i] = acc;
    }
    CV(batch_to_aff)(out, res, n);
    free(res); free(tab);
}
