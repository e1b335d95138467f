/* ORACLE (test infrastructure, NOT product code) -- gnark's R1CS solver on the CPU, over the flat program arrays.
 *
 * Restates r1cs.Solve as groth16.Prove runs it (src/prover/prover/prover.go:269 -> gnark constraint/bn254/solver.go + system.go,
 * out of tree, bnb-chain/gnark v0.10.1-0.20240910145009-4b5261061f04): levels in order, the instructions of a level in parallel
 * over the host threads (gnark: one goroutine per chunk of a level), an R1C instruction finds its ONE unsolved wire at run time,
 * hints evaluate their input expressions and write consecutive output wires.  Hints: IntegerDivision (circuit/utils.go:103-110,
 * registered at prover.go:68), bits.NBits, InvZero, rangecheck decomposition, logderivlookup lookup, logderivarg multiplicity count,
 * CmpNOp, and the BSB22 commitment placeholder (Pedersen commitment + hash_to_field).
 * Checked against oracle/py/solver.py (tests/test_oracle_solver.py).  Divisions inside one chunk of a level share one inversion
 * (Montgomery's trick): gnark-crypto's Inverse is ~10x faster than this port's Fermat inversion, the batching keeps the CPU baseline
 * from being handicapped by it.
 */
#include <omp.h>
#include <stdlib.h>
#include "orc.h"
#include "orc_field.h"

#define H_DIVMOD 1
#define H_NBITS 2
#define H_INVZERO 3
#define H_DECOMPOSE 4
#define H_LOOKUP 5
#define H_CMP 6
#define H_COUNT 7
#define H_COMMIT 8
#define CHUNK 128

/* ---- SHA-256, RFC 9380 expand_message_xmd, gnark-crypto fr.Hash ------------------------------------------------------------ */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
    0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
typedef struct { uint32_t h[8]; uint8_t buf[64]; size_t fill; uint64_t bits; } sha256_t;
static uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_block(sha256_t *s, const uint8_t *p) {
    uint32_t w[64], a, b, c, d, e, f, g, h;
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) w[i] = w[i - 16] + (ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] + (ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10));
    a = s->h[0]; b = s->h[1]; c = s->h[2]; d = s->h[3]; e = s->h[4]; f = s->h[5]; g = s->h[6]; h = s->h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (ror(e, 6) ^ ror(e, 11) ^ ror(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
        uint32_t t2 = (ror(a, 2) ^ ror(a, 13) ^ ror(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s->h[0] += a; s->h[1] += b; s->h[2] += c; s->h[3] += d; s->h[4] += e; s->h[5] += f; s->h[6] += g; s->h[7] += h;
}
static void sha_init(sha256_t *s) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(s->h, iv, 32); s->fill = 0; s->bits = 0;
}
static void sha_update(sha256_t *s, const void *data, size_t n) {
    const uint8_t *p = (const uint8_t *)data;
    s->bits += 8 * (uint64_t)n;
    while (n--) { s->buf[s->fill++] = *p++; if (s->fill == 64) { sha_block(s, s->buf); s->fill = 0; } }
}
static void sha_final(sha256_t *s, uint8_t out[32]) {
    uint64_t bits = s->bits; uint8_t x = 0x80, z = 0, lb[8];
    sha_update(s, &x, 1);
    while (s->fill != 56) sha_update(s, &z, 1);
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    sha_update(s, lb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = s->h[i] >> 24; out[4 * i + 1] = s->h[i] >> 16; out[4 * i + 2] = s->h[i] >> 8; out[4 * i + 3] = s->h[i]; }
}

/* hash_to_field("bsb22-commitment")(msg) -> one Fr element, Montgomery (48 uniform bytes, big-endian, reduced) */
void orc_commitment_challenge(const uint8_t *msg, size_t len, uint64_t *out_mont) {
    static const char dst[] = "bsb22-commitment";
    const uint8_t dlen = (uint8_t)(sizeof(dst) - 1);
    uint8_t b0[32], bi[32], u[64], zpad[64] = {0}, lib[3] = {0, 48, 0}, one = 1, two = 2;
    sha256_t s;
    sha_init(&s); sha_update(&s, zpad, 64); sha_update(&s, msg, len); sha_update(&s, lib, 3); sha_update(&s, dst, dlen); sha_update(&s, &dlen, 1); sha_final(&s, b0);
    sha_init(&s); sha_update(&s, b0, 32); sha_update(&s, &one, 1); sha_update(&s, dst, dlen); sha_update(&s, &dlen, 1); sha_final(&s, bi);
    memcpy(u, bi, 32);
    for (int k = 0; k < 32; k++) bi[k] ^= b0[k];
    sha_init(&s); sha_update(&s, bi, 32); sha_update(&s, &two, 1); sha_update(&s, dst, dlen); sha_update(&s, &dlen, 1); sha_final(&s, u + 32);
    fe acc = {{0, 0, 0, 0}}, k256 = {{256, 0, 0, 0}}, byte;
    fe_to_mont(&k256, &k256, &ORC_FR);
    for (int i = 0; i < 48; i++) {
        fe t = {{u[i], 0, 0, 0}};
        fe_to_mont(&byte, &t, &ORC_FR);
        fr_mul(&acc, &acc, &k256); fr_add(&acc, &acc, &byte);
    }
    memcpy(out_mont, acc.l, 32);
}

/* ---- the solver ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const orc_program *p; fe *w; uint8_t *solved; fe *a, *b, *c;
    fe one, minus_one; uint32_t one_id, minus_one_id;
    volatile int err; volatile uint64_t err_at;
} solver_t;

static inline void term_acc(const solver_t *s, fe *acc, uint32_t cid, const fe *x) {
    if (cid == s->one_id) { fr_add(acc, acc, x); return; }
    if (cid == s->minus_one_id) { fr_sub(acc, acc, x); return; }
    fe t; fr_mul(&t, (const fe *)s->p->coeffs + cid, x); fr_add(acc, acc, &t);
}
/* sum of the solved terms of a row; *unk = position of the last unsolved term, returns how many are unsolved */
static inline int eval_row(const solver_t *s, const uint64_t *ptr, const uint32_t *wire, const uint32_t *coef, uint64_t row, fe *acc, uint64_t *unk) {
    int n = 0;
    memset(acc, 0, sizeof(fe));
    for (uint64_t e = ptr[row]; e < ptr[row + 1]; e++) {
        if (!s->solved[wire[e]]) { *unk = e; n++; continue; }
        term_acc(s, acc, coef[e], s->w + wire[e]);
    }
    return n;
}
static inline void aux_eval(const solver_t *s, uint64_t row, fe *acc) {
    uint64_t unk; const orc_program *p = s->p;
    if (eval_row(s, p->aux_row_ptr, p->aux_wire, p->aux_coeff, row, acc, &unk)) ((solver_t *)s)->err = 5;
}
static void fail(solver_t *s, int code, uint64_t at) { if (!s->err) { s->err = code; s->err_at = at; } }

static void divmod256(const uint64_t *x, const uint64_t *d, uint64_t *q, uint64_t *r) {
    memset(q, 0, 32); memset(r, 0, 32);
    for (int bit = 255; bit >= 0; bit--) {
        for (int i = 3; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
        r[0] = (r[0] << 1) | ((x[bit >> 6] >> (bit & 63)) & 1);
        if (fe_geq_m(r, d)) { fe_sub_m(r, d); q[bit >> 6] |= 1ull << (bit & 63); }
    }
}
static inline void fr_from_u64(fe *z, uint64_t v) { fe t = {{v, 0, 0, 0}}; fe_to_mont(z, &t, &ORC_FR); }

/* a pending division: w[wire] = num / den (den == 0 with zero_ok -> 0) */
typedef struct { uint32_t wire; fe num, den; int zero_ok; } pending_t;

static void flush_pending(solver_t *s, pending_t *pd, int n, uint64_t at) {
    fe prefix[CHUNK], acc, inv, t;
    fr_one(&acc);
    for (int i = 0; i < n; i++) {
        prefix[i] = acc;
        if (fe_is_zero(&pd[i].den)) { if (!pd[i].zero_ok) fail(s, 2, at); continue; }
        fr_mul(&acc, &acc, &pd[i].den);
    }
    fr_inv(&inv, &acc);
    for (int i = n - 1; i >= 0; i--) {
        if (fe_is_zero(&pd[i].den)) { memset(&s->w[pd[i].wire], 0, sizeof(fe)); continue; }
        fr_mul(&t, &inv, &prefix[i]);                 /* 1 / den_i */
        fr_mul(&inv, &inv, &pd[i].den);
        fr_mul(&s->w[pd[i].wire], &t, &pd[i].num);
    }
}

/* instructions [q0, q1) of one level; wires written here are flagged solved by the caller after the level */
static void run_chunk(solver_t *s, uint64_t q0, uint64_t q1, uint32_t *done, size_t *n_done) {
    const orc_program *p = s->p;
    pending_t pd[CHUNK]; int npd = 0;
    for (uint64_t q = q0; q < q1; q++) {
        const uint32_t ins = p->level_instr[q], arg = p->instr_arg[ins];
        if (p->instr_kind[ins] == 0) {
            fe av, bv, cv, t; uint64_t ua = 0, ub = 0, uc = 0;
            const int na = eval_row(s, p->l_row_ptr, p->l_wire, p->l_coeff, arg, &av, &ua);
            const int nb = eval_row(s, p->r_row_ptr, p->r_wire, p->r_coeff, arg, &bv, &ub);
            const int nc = eval_row(s, p->o_row_ptr, p->o_wire, p->o_coeff, arg, &cv, &uc);
            if (na + nb + nc > 1) { fail(s, 1, arg); continue; }
            if (na + nb + nc == 0) continue;           /* assertion: checked over a, b, c after the solve */
            pending_t *e = &pd[npd++];
            e->zero_ok = 0;
            if (nc) {
                e->wire = p->o_wire[uc]; fr_mul(&t, &av, &bv); fr_sub(&e->num, &t, &cv); e->den = ((const fe *)p->coeffs)[p->o_coeff[uc]];
            } else if (na) {
                e->wire = p->l_wire[ua]; fr_mul(&t, &av, &bv); fr_sub(&e->num, &cv, &t); fr_mul(&e->den, (const fe *)p->coeffs + p->l_coeff[ua], &bv);
            } else {
                e->wire = p->r_wire[ub]; fr_mul(&t, &av, &bv); fr_sub(&e->num, &cv, &t); fr_mul(&e->den, (const fe *)p->coeffs + p->r_coeff[ub], &av);
            }
            done[(*n_done)++] = e->wire;
        } else {
            const uint32_t fn = p->hint_fn[arg], param = p->hint_param[arg], out = p->hint_out_first[arg], n_out = p->hint_n_out[arg];
            const uint64_t r0 = p->hint_in_ptr[arg], r1 = p->hint_in_end[arg];
            fe x, y, xp, yp;
            switch (fn) {
            case H_DIVMOD: {
                aux_eval(s, r0, &x); aux_eval(s, r0 + 1, &y); fe_from_mont(&xp, &x, &ORC_FR); fe_from_mont(&yp, &y, &ORC_FR);
                if (fe_is_zero(&yp)) { fail(s, 2, arg); break; }
                fe qq, rr; divmod256(xp.l, yp.l, qq.l, rr.l);
                fe_to_mont(&s->w[out], &qq, &ORC_FR); fe_to_mont(&s->w[out + 1], &rr, &ORC_FR);
                break;
            }
            case H_NBITS:
                aux_eval(s, r0, &x); fe_from_mont(&xp, &x, &ORC_FR);
                for (uint32_t k = 0; k < n_out; k++) { if (k < 256 && ((xp.l[k >> 6] >> (k & 63)) & 1)) s->w[out + k] = s->one; else memset(&s->w[out + k], 0, sizeof(fe)); }
                break;
            case H_INVZERO: {
                pending_t *e = &pd[npd++];
                aux_eval(s, r0, &e->den); e->num = s->one; e->wire = out; e->zero_ok = 1;
                break;
            }
            case H_DECOMPOSE:
                aux_eval(s, r0, &x); fe_from_mont(&xp, &x, &ORC_FR);
                for (uint32_t k = 0; k < n_out; k++) {
                    const uint32_t lo = k * param; uint64_t limb = 0;
                    if (lo < 256) {
                        limb = xp.l[lo >> 6] >> (lo & 63);
                        if ((lo & 63) + param > 64 && (lo >> 6) + 1 < 4) limb |= xp.l[(lo >> 6) + 1] << (64 - (lo & 63));
                        limb &= (1ull << param) - 1;
                    }
                    fr_from_u64(&s->w[out + k], limb);
                }
                break;
            case H_LOOKUP: {
                const uint64_t t0 = p->table_ptr[param], t1 = p->table_ptr[param + 1];
                for (uint64_t r = r0; r < r1; r++) {
                    aux_eval(s, r, &x); fe_from_mont(&xp, &x, &ORC_FR);
                    if (xp.l[1] | xp.l[2] | xp.l[3] || xp.l[0] >= t1 - t0) { fail(s, 3, arg); break; }
                    aux_eval(s, t0 + xp.l[0], &s->w[out + (uint32_t)(r - r0)]);
                }
                break;
            }
            case H_CMP: {
                aux_eval(s, r0, &x); aux_eval(s, r0 + 1, &y); fe_from_mont(&xp, &x, &ORC_FR); fe_from_mont(&yp, &y, &ORC_FR);
                const int ge = fe_geq_m(xp.l, yp.l), le = fe_geq_m(yp.l, xp.l);
                if (ge && le) memset(&s->w[out], 0, sizeof(fe)); else s->w[out] = ge ? s->one : s->minus_one;
                break;
            }
            default: fail(s, 4, arg);
            }
            for (uint32_t k = 0; k < n_out; k++) done[(*n_done)++] = out + k;
        }
        if (npd == CHUNK) { flush_pending(s, pd, npd, q0); npd = 0; }
    }
    if (npd) flush_pending(s, pd, npd, q0);
}

int orc_solve(const orc_program *p, const uint64_t *inputs_mont, uint64_t *wires, uint64_t *oa, uint64_t *ob, uint64_t *oc,
              orc_commit_fn commit, void *user, uint64_t *err_at, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    solver_t s; memset(&s, 0, sizeof(s));
    s.p = p; s.w = (fe *)wires;
    s.solved = (uint8_t *)calloc(p->n_wires, 1);
    fr_one(&s.one); fe_neg(&s.minus_one, &s.one, &ORC_FR);
    s.one_id = s.minus_one_id = 0xFFFFFFFFu;
    for (uint64_t i = 0; i < p->n_coeffs; i++) {
        if (s.one_id == 0xFFFFFFFFu && fe_eq((const fe *)p->coeffs + i, &s.one)) s.one_id = (uint32_t)i;
        if (s.minus_one_id == 0xFFFFFFFFu && fe_eq((const fe *)p->coeffs + i, &s.minus_one)) s.minus_one_id = (uint32_t)i;
    }
    const uint64_t n_in = p->n_public - 1 + p->n_secret;
    s.w[0] = s.one;
    memcpy(s.w + 1, inputs_mont, n_in * 32);
    memset(s.solved, 1, n_in + 1);
    for (uint64_t l = 0; l < p->n_levels && !s.err; l++) {
        const uint64_t q0 = p->level_ptr[l], q1 = p->level_ptr[l + 1];
        /* special hints of this level first (they read earlier levels only) */
        for (uint64_t q = q0; q < q1 && !s.err; q++) {
            const uint32_t ins = p->level_instr[q], arg = p->instr_arg[ins];
            if (p->instr_kind[ins] != 1) continue;
            const uint32_t fn = p->hint_fn[arg], out = p->hint_out_first[arg], n_out = p->hint_n_out[arg];
            if (fn == H_COUNT) {
                uint32_t *cnt = (uint32_t *)calloc(n_out, 4);
                const uint64_t r0 = p->hint_in_ptr[arg], r1 = p->hint_in_end[arg];
#pragma omp parallel for num_threads(threads) schedule(static)
                for (uint64_t r = r0; r < r1; r++) {
                    fe x, xp; aux_eval(&s, r, &x); fe_from_mont(&xp, &x, &ORC_FR);
                    if (xp.l[1] | xp.l[2] | xp.l[3] || xp.l[0] >= n_out) { fail(&s, 3, arg); continue; }
                    __atomic_fetch_add(&cnt[xp.l[0]], 1u, __ATOMIC_RELAXED);
                }
                for (uint32_t k = 0; k < n_out; k++) { fr_from_u64(&s.w[out + k], cnt[k]); s.solved[out + k] = 1; }
                free(cnt);
            } else if (fn == H_COMMIT) {
                if (!commit) { fail(&s, 6, arg); break; }
                fe *vals = (fe *)malloc(32 * (p->n_committed ? p->n_committed : 1));
                for (uint64_t i = 0; i < p->n_committed; i++) {
                    if (!s.solved[p->private_committed[i]]) fail(&s, 7, p->private_committed[i]);
                    vals[i] = s.w[p->private_committed[i]];
                }
                commit((const uint64_t *)vals, p->n_committed, (uint64_t *)&s.w[out], user);
                s.solved[out] = 1;
                free(vals);
            }
        }
        /* the rest of the level: chunks over the threads; solved flags are raised after the level */
        const uint64_t n_l = q1 - q0, n_chunks = (n_l + CHUNK - 1) / CHUNK;
#pragma omp parallel num_threads(threads) if (n_l > 4 * CHUNK && threads > 1)
        {
            size_t cap = 4 * CHUNK, nd = 0; uint32_t *dn = (uint32_t *)malloc(4 * cap);
#pragma omp for schedule(dynamic, 4)
            for (uint64_t ch = 0; ch < n_chunks; ch++) {
                const uint64_t a0 = q0 + ch * CHUNK, a1 = a0 + CHUNK < q1 ? a0 + CHUNK : q1;
                size_t need = 0;   /* outputs of this chunk: one per R1C, n_out per hint */
                for (uint64_t q = a0; q < a1; q++) { const uint32_t ins = p->level_instr[q]; need += p->instr_kind[ins] == 0 ? 1 : p->hint_n_out[p->instr_arg[ins]]; }
                if (nd + need > cap) { cap = (nd + need) * 2; dn = (uint32_t *)realloc(dn, 4 * cap); }
                uint64_t b0 = a0;  /* runs between the special hints handled above */
                for (uint64_t q = a0; q <= a1; q++) {
                    int special = 0;
                    if (q < a1) { const uint32_t ins = p->level_instr[q]; if (p->instr_kind[ins] == 1) { const uint32_t fn = p->hint_fn[p->instr_arg[ins]]; special = fn == H_COUNT || fn == H_COMMIT; } }
                    if (q == a1 || special) { if (q > b0) run_chunk(&s, b0, q, dn, &nd); b0 = q + 1; }
                }
            }
            /* the implicit barrier of the loop: every chunk of the level is done; raise the flags */
            for (size_t i = 0; i < nd; i++) s.solved[dn[i]] = 1;
            free(dn);
        }
    }
    int rc = s.err;
    if (!rc) for (uint64_t i = 0; i < p->n_wires; i++) if (!s.solved[i]) { rc = 8; s.err_at = i; break; }
    free(s.solved);
    if (!rc && oa && ob && oc) {
        /* a = L w, b = R w, c = O w; gnark's check: every constraint satisfied */
        volatile int bad = 0; volatile uint64_t bad_at = 0;
        solver_t full = s; full.solved = NULL;
#pragma omp parallel for num_threads(threads) schedule(static)
        for (uint64_t k = 0; k < p->n_constraints; k++) {
            fe av = {{0, 0, 0, 0}}, bv = av, cv = av, t;
            for (uint64_t e = p->l_row_ptr[k]; e < p->l_row_ptr[k + 1]; e++) term_acc(&full, &av, p->l_coeff[e], full.w + p->l_wire[e]);
            for (uint64_t e = p->r_row_ptr[k]; e < p->r_row_ptr[k + 1]; e++) term_acc(&full, &bv, p->r_coeff[e], full.w + p->r_wire[e]);
            for (uint64_t e = p->o_row_ptr[k]; e < p->o_row_ptr[k + 1]; e++) term_acc(&full, &cv, p->o_coeff[e], full.w + p->o_wire[e]);
            fr_mul(&t, &av, &bv);
            if (!fe_eq(&t, &cv) && !bad) { bad = 1; bad_at = k; }
            ((fe *)oa)[k] = av; ((fe *)ob)[k] = bv; ((fe *)oc)[k] = cv;
        }
        if (bad) { rc = 9; s.err_at = bad_at; }
    }
    if (err_at) *err_at = s.err_at;
    return rc;
}
