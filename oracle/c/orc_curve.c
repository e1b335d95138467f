/* ORACLE (test infrastructure, NOT product code) -- G1/G2 group operations and MultiExp on the CPU.
 * Restates what the reference reaches through groth16.Prove (src/prover/prover/prover.go:269): gnark-crypto
 * G1Jac.MultiExp / G2Jac.MultiExp and BatchScalarMultiplicationG1/G2 (ecc/bn254/multiexp.go, g1.go, g2.go, out of
 * tree).  The result of a multi-exponentiation is a group element; parity is on its affine coordinates. */
#include <stdlib.h>
#include <omp.h>
#include "orc.h"
#include "orc_field.h"

const fparams ORC_FP = {
  {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
  0x87d20782e4866389ULL,
  {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL},
  {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}
};
const fparams ORC_FR = {
  {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
  0xc2e1f593efffffffULL,
  {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL},
  {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}
};

/* ---- G1 over Fp ---- */
#define CV(n) g1_##n
#define FT fe
#define F_add fp_add
#define F_sub fp_sub
#define F_mul fp_mul
#define F_sqr fp_sqr
#define F_neg fp_neg
#define F_inv fp_inv
#define F_one fp_one
#define F_is_zero fp_is_zero
#define F_eq fp_eq
#include "orc_curve_impl.h"
#undef CV
#undef FT
#undef F_add
#undef F_sub
#undef F_mul
#undef F_sqr
#undef F_neg
#undef F_inv
#undef F_one
#undef F_is_zero
#undef F_eq

/* ---- G2 over Fp2 ---- */
#define CV(n) g2_##n
#define FT fe2
#define F_add fp2_add
#define F_sub fp2_sub
#define F_mul fp2_mul
#define F_sqr fp2_sqr
#define F_neg fp2_neg
#define F_inv fp2_inv
#define F_one fp2_one
#define F_is_zero fp2_is_zero
#define F_eq fp2_eq
#include "orc_curve_impl.h"

int orc_num_threads(void) { return omp_get_max_threads(); }

void orc_to_mont(uint64_t *io, size_t n, int which) {
    const fparams *P = which ? &ORC_FR : &ORC_FP;
    for (size_t i = 0; i < n; i++) fe_to_mont((fe *)(io + 4 * i), (fe *)(io + 4 * i), P);
}
void orc_from_mont(uint64_t *io, size_t n, int which) {
    const fparams *P = which ? &ORC_FR : &ORC_FP;
    for (size_t i = 0; i < n; i++) fe_from_mont((fe *)(io + 4 * i), (fe *)(io + 4 * i), P);
}
void orc_fr_mul_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) fr_mul((fe *)(out + 4 * i), (const fe *)(a + 4 * i), (const fe *)(b + 4 * i));
}

static void g1_gen(g1_aff *g) {
    fe one = {{1, 0, 0, 0}}, two = {{2, 0, 0, 0}};
    fe_to_mont(&g->x, &one, &ORC_FP); fe_to_mont(&g->y, &two, &ORC_FP);
}
static void g2_gen(g2_aff *g) {   /* EIP-197 generator, plain limbs, converted below */
    static const uint64_t c[4][4] = {
        {0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL},   /* x.a0 */
        {0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL},   /* x.a1 */
        {0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL},   /* y.a0 */
        {0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL},   /* y.a1 */
    };
    fe t;
    memcpy(t.l, c[0], 32); fe_to_mont(&g->x.a0, &t, &ORC_FP);
    memcpy(t.l, c[1], 32); fe_to_mont(&g->x.a1, &t, &ORC_FP);
    memcpy(t.l, c[2], 32); fe_to_mont(&g->y.a0, &t, &ORC_FP);
    memcpy(t.l, c[3], 32); fe_to_mont(&g->y.a1, &t, &ORC_FP);
}

void orc_g1_fixed_base(const uint64_t *sc, size_t n, uint64_t *out, int threads) {
    g1_aff g; g1_gen(&g); g1_fixed_base_batch((g1_aff *)out, &g, sc, n, threads > 0 ? threads : omp_get_max_threads());
}
void orc_g2_fixed_base(const uint64_t *sc, size_t n, uint64_t *out, int threads) {
    g2_aff g; g2_gen(&g); g2_fixed_base_batch((g2_aff *)out, &g, sc, n, threads > 0 ? threads : omp_get_max_threads());
}

static uint64_t *scalars_to_plain(const uint64_t *mont, size_t n) {
    uint64_t *p = (uint64_t *)malloc(32 * (n ? n : 1));
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fe_from_mont((fe *)(p + 4 * i), (const fe *)(mont + 4 * i), &ORC_FR);
    return p;
}

void orc_g1_msm_jac(g1_jac *out, const uint64_t *pts, const uint64_t *sc_mont, size_t n, int threads) {
    uint64_t *plain = scalars_to_plain(sc_mont, n);
    g1_msm(out, (const g1_aff *)pts, plain, n, threads > 0 ? threads : omp_get_max_threads());
    free(plain);
}
void orc_g2_msm_jac(g2_jac *out, const uint64_t *pts, const uint64_t *sc_mont, size_t n, int threads) {
    uint64_t *plain = scalars_to_plain(sc_mont, n);
    g2_msm(out, (const g2_aff *)pts, plain, n, threads > 0 ? threads : omp_get_max_threads());
    free(plain);
}
void orc_g1_msm(const uint64_t *pts, const uint64_t *sc, size_t n, uint64_t *out, int threads) {
    g1_jac j; orc_g1_msm_jac(&j, pts, sc, n, threads); g1_jac_to_aff((g1_aff *)out, &j);
}
void orc_g2_msm(const uint64_t *pts, const uint64_t *sc, size_t n, uint64_t *out, int threads) {
    g2_jac j; orc_g2_msm_jac(&j, pts, sc, n, threads); g2_jac_to_aff((g2_aff *)out, &j);
}

void orc_g1_add(const uint64_t *p, const uint64_t *q, uint64_t *out) {
    g1_jac a, b; g1_jac_from_aff(&a, (const g1_aff *)p); g1_jac_from_aff(&b, (const g1_aff *)q);
    g1_jac_add(&a, &a, &b); g1_jac_to_aff((g1_aff *)out, &a);
}
void orc_g1_scalar_mul(const uint64_t *p, const uint64_t *k, uint64_t *out) {
    g1_jac a; g1_jac_from_aff(&a, (const g1_aff *)p); g1_jac_mul(&a, &a, k); g1_jac_to_aff((g1_aff *)out, &a);
}
void orc_g2_add(const uint64_t *p, const uint64_t *q, uint64_t *out) {
    g2_jac a, b; g2_jac_from_aff(&a, (const g2_aff *)p); g2_jac_from_aff(&b, (const g2_aff *)q);
    g2_jac_add(&a, &a, &b); g2_jac_to_aff((g2_aff *)out, &a);
}
void orc_g2_scalar_mul(const uint64_t *p, const uint64_t *k, uint64_t *out) {
    g2_jac a; g2_jac_from_aff(&a, (const g2_aff *)p); g2_jac_mul(&a, &a, k); g2_jac_to_aff((g2_aff *)out, &a);
}
int orc_g1_on_curve(const uint64_t *p) {
    const g1_aff *a = (const g1_aff *)p;
    if (g1_aff_is_inf(a)) return 1;
    fe l, r, three = {{3, 0, 0, 0}}; fe_to_mont(&three, &three, &ORC_FP);
    fp_sqr(&l, &a->y); fp_sqr(&r, &a->x); fp_mul(&r, &r, &a->x); fp_add(&r, &r, &three);
    return fp_eq(&l, &r);
}
int orc_g2_on_curve(const uint64_t *p) {
    const g2_aff *a = (const g2_aff *)p;
    if (g2_aff_is_inf(a)) return 1;
    fe2 l, r, b, nine_u; fe t = {{9, 0, 0, 0}}, three = {{3, 0, 0, 0}};
    fe_to_mont(&nine_u.a0, &t, &ORC_FP); fp_one(&nine_u.a1);
    fp2_inv(&b, &nine_u); fe_to_mont(&three, &three, &ORC_FP);
    fp_mul(&b.a0, &b.a0, &three); fp_mul(&b.a1, &b.a1, &three);   /* b' = 3/(9+u) */
    fp2_sqr(&l, &a->y); fp2_sqr(&r, &a->x); fp2_mul(&r, &r, &a->x); fp2_add(&r, &r, &b);
    return fp2_eq(&l, &r);
}

/* ---- helpers shared with orc_groth16.c ---- */
void orc__g1_finish(uint8_t *out64, const g1_jac *p);
void orc__g2_finish(uint8_t *out128, const g2_jac *p);
static void be32(uint8_t *o, const fe *m) {
    fe p; fe_from_mont(&p, m, &ORC_FP);
    for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) o[31 - (8 * i + b)] = (uint8_t)(p.l[i] >> (8 * b));
}
/* gnark-crypto RawBytes: X||Y big-endian (G2: X.A1||X.A0||Y.A1||Y.A0); infinity = 0x40 then zeros */
void orc__g1_finish(uint8_t *o, const g1_jac *p) {
    g1_aff a; g1_jac_to_aff(&a, p);
    if (g1_aff_is_inf(&a)) { memset(o, 0, 64); o[0] = 0x40; return; }
    be32(o, &a.x); be32(o + 32, &a.y);
}
void orc__g2_finish(uint8_t *o, const g2_jac *p) {
    g2_aff a; g2_jac_to_aff(&a, p);
    if (g2_aff_is_inf(&a)) { memset(o, 0, 128); o[0] = 0x40; return; }
    be32(o, &a.x.a1); be32(o + 32, &a.x.a0); be32(o + 64, &a.y.a1); be32(o + 96, &a.y.a0);
}

/* ---- Groth16 prove: the group part (orc_groth16_prove lives here to reuse the static group law) ---- */
void orc_compute_h(const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t m, int logn, uint64_t *out_h, int threads);

static int prove_core(const orc_pk *pk, const uint64_t *wa, const uint64_t *wb, const uint64_t *wk, const uint64_t *committed,
                      const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n_constraints,
                      const uint64_t *r_plain, const uint64_t *s_plain, uint8_t *out, int threads, const g1_jac *commit_done);

int orc_groth16_prove(const orc_pk *pk, const uint64_t *wa, const uint64_t *wb, const uint64_t *wk, const uint64_t *committed,
                      const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n_constraints,
                      const uint64_t *r_plain, const uint64_t *s_plain, uint8_t *out, int threads) {
    return prove_core(pk, wa, wb, wk, committed, a, b, c, n_constraints, r_plain, s_plain, out, threads, NULL);
}

static int prove_core(const orc_pk *pk, const uint64_t *wa, const uint64_t *wb, const uint64_t *wk, const uint64_t *committed,
                      const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n_constraints,
                      const uint64_t *r_plain, const uint64_t *s_plain, uint8_t *out, int threads, const g1_jac *commit_done) {
    /* Restates gnark v0.10 backend/groth16/bn254/prove.go after the solver has run:
     *   pok  = MSM(BasisExpSigma, committed)            commitment = MSM(Basis, committed)
     *   h    = computeH(a, b, c)                        (bit-reversed, paired with pk.G1.Z as stored)
     *   Ar   = MSM(A, wa) + alpha1 + r*delta1           Bs1 = MSM(B1, wb) + beta1 + s*delta1
     *   Bs   = MSM(B2, wb) + s*delta2 + beta2
     *   Krs  = MSM(K, wk) + MSM(Z, h[:n-1]) + (-r*s)*delta1 + s*Ar + r*Bs1
     * proof.WriteRawTo layout: Ar 64 | Bs 128 | Krs 64 | u32be 1 | Commitment 64 | Pok 64 = 388 bytes. */
    if (threads <= 0) threads = omp_get_max_threads();
    size_t n = (size_t)1 << pk->log_n;
    if (pk->n_z != n - 1) return -1;
    uint64_t *h = (uint64_t *)malloc(32 * n);
    orc_compute_h(a, b, c, n_constraints, pk->log_n, h, threads);

    g1_jac ar, bs1, krs, kz, t, commit, pok; g2_jac bs2, t2;
    if (commit_done) commit = *commit_done; else orc_g1_msm_jac(&commit, pk->ck_basis, committed, pk->n_ck, threads);
    orc_g1_msm_jac(&pok, pk->ck_basis_exp_sigma, committed, pk->n_ck, threads);
    orc_g1_msm_jac(&ar, pk->A, wa, pk->n_a, threads);
    orc_g1_msm_jac(&bs1, pk->B1, wb, pk->n_b, threads);
    orc_g2_msm_jac(&bs2, pk->B2, wb, pk->n_b, threads);
    orc_g1_msm_jac(&krs, pk->K, wk, pk->n_k, threads);
    orc_g1_msm_jac(&kz, pk->Z, h, pk->n_z, threads);
    free(h);

    fe rm, sm, kr; fe rp, sp;
    memcpy(rp.l, r_plain, 32); memcpy(sp.l, s_plain, 32);
    fe_to_mont(&rm, &rp, &ORC_FR); fe_to_mont(&sm, &sp, &ORC_FR);
    fr_mul(&kr, &rm, &sm); fe_neg(&kr, &kr, &ORC_FR); fe_from_mont(&kr, &kr, &ORC_FR);

    g1_jac d1, dr, ds, dkr; g1_jac_from_aff(&d1, (const g1_aff *)pk->delta1);
    g1_jac_mul(&dr, &d1, rp.l); g1_jac_mul(&ds, &d1, sp.l); g1_jac_mul(&dkr, &d1, kr.l);
    g1_jac_add_mixed(&ar, &ar, (const g1_aff *)pk->alpha1); g1_jac_add(&ar, &ar, &dr);
    g1_jac_add_mixed(&bs1, &bs1, (const g1_aff *)pk->beta1); g1_jac_add(&bs1, &bs1, &ds);
    g2_jac_from_aff(&t2, (const g2_aff *)pk->delta2); g2_jac_mul(&t2, &t2, sp.l);
    g2_jac_add(&bs2, &bs2, &t2); g2_jac_add_mixed(&bs2, &bs2, (const g2_aff *)pk->beta2);
    g1_jac_add(&krs, &krs, &dkr); g1_jac_add(&krs, &krs, &kz);
    g1_jac_mul(&t, &ar, sp.l); g1_jac_add(&krs, &krs, &t);
    g1_jac_mul(&t, &bs1, rp.l); g1_jac_add(&krs, &krs, &t);

    orc__g1_finish(out, &ar); orc__g2_finish(out + 64, &bs2); orc__g1_finish(out + 192, &krs);
    out[256] = 0; out[257] = 0; out[258] = 0; out[259] = 1;
    orc__g1_finish(out + 260, &commit); orc__g1_finish(out + 324, &pok);
    return 0;
}

/* ---- groth16.Prove from the circuit inputs: solver (orc_solver.c) + the proof above ---------------------------------------- */
typedef struct { const orc_pk *pk; g1_jac commit; uint64_t *committed; int threads; } commit_ctx;
static void commit_cb(const uint64_t *vals, size_t n, uint64_t *challenge, void *user) {
    commit_ctx *cc = (commit_ctx *)user;
    uint8_t raw[64];
    orc_g1_msm_jac(&cc->commit, cc->pk->ck_basis, vals, n, cc->threads);
    orc__g1_finish(raw, &cc->commit);
    cc->committed = (uint64_t *)malloc(32 * (n ? n : 1));
    memcpy(cc->committed, vals, 32 * n);
    orc_commitment_challenge(raw, 64, challenge);
}
/* keep[i] != 0 -> element i goes to the output, order preserved; parallel two-pass compaction */
static size_t compact32(const uint64_t *src, size_t n, const uint8_t *keep, uint64_t *dst, int threads) {
    const size_t nb = (size_t)threads * 8, per = (n + nb - 1) / nb;
    size_t *cnt = (size_t *)calloc(nb + 1, sizeof(size_t));
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t b = 0; b < nb; b++) { size_t c = 0; for (size_t i = b * per; i < n && i < (b + 1) * per; i++) c += keep[i] != 0; cnt[b + 1] = c; }
    for (size_t b = 0; b < nb; b++) cnt[b + 1] += cnt[b];
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t b = 0; b < nb; b++) { size_t o = cnt[b]; for (size_t i = b * per; i < n && i < (b + 1) * per; i++) if (keep[i]) { memcpy(dst + 4 * o, src + 4 * i, 32); o++; } }
    const size_t total = cnt[nb];
    free(cnt);
    return total;
}

int orc_groth16_prove_program(const orc_pk *pk, const orc_program *prog, const uint8_t *infinity_a, const uint8_t *infinity_b,
                              uint64_t commitment_index, const uint64_t *inputs_mont, const uint64_t *r_plain, const uint64_t *s_plain,
                              uint8_t *out, double *seconds, uint64_t *err_at, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    const size_t nw = prog->n_wires, m = prog->n_constraints;
    uint64_t *w = (uint64_t *)malloc(32 * nw), *a = (uint64_t *)malloc(32 * m), *b = (uint64_t *)malloc(32 * m), *c = (uint64_t *)malloc(32 * m);
    commit_ctx cc; memset(&cc, 0, sizeof(cc)); cc.pk = pk; cc.threads = threads;
    double t0 = omp_get_wtime();
    int rc = orc_solve(prog, inputs_mont, w, a, b, c, commit_cb, &cc, err_at, threads);
    double t1 = omp_get_wtime();
    if (rc == 0) {
        uint8_t *keep = (uint8_t *)malloc(nw);
        uint64_t *wa = (uint64_t *)malloc(32 * nw), *wb = (uint64_t *)malloc(32 * nw), *wk = (uint64_t *)malloc(32 * nw);
        for (size_t i = 0; i < nw; i++) keep[i] = !infinity_a[i];
        const size_t na = compact32(w, nw, keep, wa, threads);
        for (size_t i = 0; i < nw; i++) keep[i] = !infinity_b[i];
        const size_t nb = compact32(w, nw, keep, wb, threads);
        for (size_t i = 0; i < nw; i++) keep[i] = i >= prog->n_public;
        for (size_t i = 0; i < prog->n_committed; i++) keep[prog->private_committed[i]] = 0;
        if (cc.committed) keep[commitment_index] = 0;
        const size_t nk = compact32(w, nw, keep, wk, threads);
        if (na != pk->n_a || nb != pk->n_b || nk != pk->n_k) rc = -2;
        else rc = prove_core(pk, wa, wb, wk, cc.committed, a, b, c, m, r_plain, s_plain, out, threads, cc.committed ? &cc.commit : NULL);
        free(keep); free(wa); free(wb); free(wk);
    }
    if (seconds) { seconds[0] = t1 - t0; seconds[1] = omp_get_wtime() - t1; }
    free(cc.committed); free(w); free(a); free(b); free(c);
    return rc;
}
