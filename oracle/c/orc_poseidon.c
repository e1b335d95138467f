/* ORACLE (test infrastructure, NOT product code) -- Poseidon (x^5, R_F=8, iden3 R_P table, Grain constants),
 * FixedDepthMerkleTree and account-leaf hashing on the CPU.
 *
 * Restates: bnb-chain gnark-crypto ecc/bn254/fr/poseidon (out of tree; call sites src/utils/account_tree.go:19,
 * src/utils/utils.go:748, src/witness/main.go:181), src/utils/merkletree/merkletree.go:137-308 and
 * src/utils/utils.go:188-221,744-750.  Conventions and parity status: oracle/py/poseidon.py, oracle/py/merkle.py.
 * This file implements its own Grain LFSR so that the CUDA product's constants are checked against an
 * independent generator. */
#include <stdlib.h>
#include <omp.h>
#include "orc.h"
#include "orc_field.h"

#define MAX_T 13
static const int ROUNDS_P[16] = {56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68};
static int g_out_lane = 1;   /* see oracle/py/poseidon.py PARITY STATUS */

typedef struct { int ready, rp; fe *rc; fe *mds; } pconst;
static pconst g_pc[MAX_T + 1];

void orc_poseidon_set_out_lane(int lane) { g_out_lane = lane; }
int orc_poseidon_get_out_lane(void) { return g_out_lane; }

/* ---- Grain LFSR (Hades reference parameter script, self-shrinking mode) ---- */
typedef struct { uint8_t s[80]; } grain;
static int grain_clock(grain *g) {
    int nb = g->s[62] ^ g->s[51] ^ g->s[38] ^ g->s[23] ^ g->s[13] ^ g->s[0];
    memmove(g->s, g->s + 1, 79); g->s[79] = (uint8_t)nb;
    return nb;
}
static void grain_init(grain *g, int t, int rf, int rp) {
    int pos = 0;
    const int vals[6] = {1, 0, 254, t, rf, rp}, widths[6] = {2, 4, 12, 12, 10, 10};
    for (int k = 0; k < 6; k++) for (int i = widths[k] - 1; i >= 0; i--) g->s[pos++] = (uint8_t)((vals[k] >> i) & 1);
    while (pos < 80) g->s[pos++] = 1;
    for (int i = 0; i < 160; i++) grain_clock(g);
}
static int grain_bit(grain *g) { for (;;) { int b1 = grain_clock(g), b2 = grain_clock(g); if (b1) return b2; } }
static void grain_draw(grain *g, uint64_t v[4]) {   /* 254 bits, most significant first */
    v[0] = v[1] = v[2] = v[3] = 0;
    for (int i = 0; i < 254; i++) {
        v[3] = (v[3] << 1) | (v[2] >> 63); v[2] = (v[2] << 1) | (v[1] >> 63); v[1] = (v[1] << 1) | (v[0] >> 63);
        v[0] = (v[0] << 1) | (uint64_t)grain_bit(g);
    }
}

static void pc_build(int t) {
    pconst *pc = &g_pc[t];
    int rp = ROUNDS_P[t - 2], nrc = (8 + rp) * t;
    grain g; grain_init(&g, t, 8, rp);
    pc->rc = (fe *)malloc(sizeof(fe) * nrc); pc->mds = (fe *)malloc(sizeof(fe) * t * t); pc->rp = rp;
    for (int i = 0; i < nrc;) {
        uint64_t v[4]; grain_draw(&g, v);
        if (fe_geq_m(v, ORC_FR.m)) continue;                   /* rejection sampling */
        fe p; memcpy(p.l, v, 32); fe_to_mont(&pc->rc[i++], &p, &ORC_FR);
    }
    for (;;) {
        fe xy[2 * MAX_T]; int ok = 1;
        for (int i = 0; i < 2 * t; i++) {
            uint64_t v[4]; grain_draw(&g, v);
            if (fe_geq_m(v, ORC_FR.m)) fe_sub_m(v, ORC_FR.m);  /* F(v): reduce (v < 2^254 < 2r) */
            fe p; memcpy(p.l, v, 32); fe_to_mont(&xy[i], &p, &ORC_FR);
        }
        for (int i = 0; i < 2 * t && ok; i++) for (int j = 0; j < i; j++) if (fe_eq(&xy[i], &xy[j])) ok = 0;
        for (int i = 0; i < t && ok; i++) for (int j = 0; j < t; j++) {
            fe s; fr_add(&s, &xy[i], &xy[t + j]);
            if (fe_is_zero(&s)) { ok = 0; break; }
            fr_inv(&pc->mds[i * t + j], &s);
        }
        if (ok) break;
    }
    pc->ready = 1;
}
static const pconst *pc_get(int t) {
    if (!g_pc[t].ready) {
        #pragma omp critical(orc_pc)
        { if (!g_pc[t].ready) pc_build(t); }
    }
    return &g_pc[t];
}

void orc_poseidon_constants(int t, uint64_t *rc, uint64_t *mds, int *rounds_p) {
    const pconst *pc = pc_get(t);
    if (rc) memcpy(rc, pc->rc, sizeof(fe) * (8 + pc->rp) * t);
    if (mds) memcpy(mds, pc->mds, sizeof(fe) * t * t);
    if (rounds_p) *rounds_p = pc->rp;
}

static void sbox(fe *x) { fe x2, x4; fr_sqr(&x2, x); fr_sqr(&x4, &x2); fr_mul(x, &x4, x); }

static void permute(fe *s, int t) {
    const pconst *pc = pc_get(t);
    int rounds = 8 + pc->rp;
    fe n[MAX_T];
    for (int r = 0; r < rounds; r++) {
        for (int i = 0; i < t; i++) fr_add(&s[i], &s[i], &pc->rc[r * t + i]);
        if (r < 4 || r >= 4 + pc->rp) { for (int i = 0; i < t; i++) sbox(&s[i]); } else sbox(&s[0]);
        for (int i = 0; i < t; i++) {
            fe acc, m; memset(&acc, 0, sizeof acc);
            for (int j = 0; j < t; j++) { fr_mul(&m, &pc->mds[i * t + j], &s[j]); fr_add(&acc, &acc, &m); }
            n[i] = acc;
        }
        memcpy(s, n, sizeof(fe) * t);
    }
}
void orc_poseidon_permute(uint64_t *state, int t) { permute((fe *)state, t); }

/* poseidon.Poseidon(input...): 12 per permutation, lane 0 carried, last chunk at width rem+1 */
static void hash_fe(const fe *in, size_t n, fe *out) {
    fe st[MAX_T]; memset(st, 0, sizeof st);
    size_t start = 0; int width = MAX_T;
    if (n > 12) for (size_t i = 0; i < n / 12; i++) { memcpy(&st[1], &in[start], 12 * sizeof(fe)); permute(st, 13); start += 12; }
    if (start < n) { size_t rem = n - start; memcpy(&st[1], &in[start], rem * sizeof(fe)); permute(st, (int)rem + 1); width = (int)rem + 1; }
    *out = st[g_out_lane < width ? g_out_lane : 0];
}
void orc_poseidon_hash(const uint64_t *in, size_t n, uint64_t *out) { hash_fe((const fe *)in, n, (fe *)out); }

static void fe_from_be(fe *z, const uint8_t *b) {   /* canonical (< r) big-endian -> Montgomery */
    fe p;
    for (int i = 0; i < 4; i++) { uint64_t v = 0; for (int k = 0; k < 8; k++) v = (v << 8) | b[8 * (3 - i) + k]; p.l[i] = v; }
    fe_to_mont(z, &p, &ORC_FR);
}
static void fe_to_be(uint8_t *b, const fe *m) {
    fe p; fe_from_mont(&p, m, &ORC_FR);
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) b[8 * (3 - i) + k] = (uint8_t)(p.l[i] >> (8 * (7 - k)));
}
void orc_poseidon_hash_be(const uint8_t *in, size_t n, uint8_t *out) {
    fe *e = (fe *)malloc(sizeof(fe) * n), o;
    for (size_t i = 0; i < n; i++) fe_from_be(&e[i], in + 32 * i);
    hash_fe(e, n, &o); fe_to_be(out, &o); free(e);
}
static void node_hash(const uint8_t *l, const uint8_t *r, uint8_t *out) {
    fe e[2], o; fe_from_be(&e[0], l); fe_from_be(&e[1], r); hash_fe(e, 2, &o); fe_to_be(out, &o);
}
void orc_poseidon_node_batch(const uint8_t *pairs, size_t count, uint8_t *out, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    pc_get(3);
    #pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < count; i++) node_hash(pairs + 64 * i, pairs + 64 * i + 32, out + 32 * i);
}

/* ---- FixedDepthMerkleTree.  Output representation: for level l = 1..depth a dense array of
 * level_len(l) = max(1, ceil(capacity / 2^l)) nodes; a position the reference leaves "not dirty" holds
 * nilHashes[l] (which is what getNodeAt returns for it, merkletree.go:315-331). ---- */
size_t orc_merkle_level_len(size_t capacity, int level) {
    size_t len = (capacity + (((size_t)1 << level) - 1)) >> level;
    return len ? len : 1;
}
size_t orc_merkle_nodes_total(size_t capacity, int depth) {
    size_t t = 0; for (int l = 1; l <= depth; l++) t += orc_merkle_level_len(capacity, l); return t;
}
static int bit_get(const uint64_t *bs, size_t i) { return bs ? (int)((bs[i >> 6] >> (i & 63)) & 1) : 1; }

void orc_merkle_build(const uint8_t *leaves, const uint64_t *dirty, size_t capacity, int depth, const uint8_t *nil_leaf,
                      uint8_t *out_nodes, uint8_t *out_root, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    pc_get(3);
    uint8_t (*nil)[32] = (uint8_t (*)[32])malloc(32 * (depth + 1));
    memcpy(nil[0], nil_leaf, 32);
    for (int l = 1; l <= depth; l++) node_hash(nil[l - 1], nil[l - 1], nil[l]);
    /* child-level dirty flags, one byte per position */
    size_t prev_len = capacity;
    uint8_t *prev_dirty = (uint8_t *)malloc(prev_len ? prev_len : 1);
    for (size_t i = 0; i < capacity; i++) prev_dirty[i] = (uint8_t)bit_get(dirty, i);
    const uint8_t *prev = leaves;
    uint8_t *cur = out_nodes;
    for (int l = 1; l <= depth; l++) {
        size_t len = orc_merkle_level_len(capacity, l);
        uint8_t *cur_dirty = (uint8_t *)calloc(len, 1);
        #pragma omp parallel for schedule(static) num_threads(threads)
        for (size_t p = 0; p < len; p++) {
            size_t lc = 2 * p, rc = 2 * p + 1;
            int dl = lc < prev_len && prev_dirty[lc], dr = rc < prev_len && prev_dirty[rc];
            if (!dl && !dr) { memcpy(cur + 32 * p, nil[l], 32); continue; }
            const uint8_t *L = dl ? prev + 32 * lc : nil[l - 1], *Rr = dr ? prev + 32 * rc : nil[l - 1];
            node_hash(L, Rr, cur + 32 * p);
            cur_dirty[p] = 1;
        }
        free(prev_dirty); prev_dirty = cur_dirty; prev_len = len; prev = cur; cur += 32 * len;
    }
    memcpy(out_root, prev, 32);   /* level `depth` has exactly one node when capacity <= 2^depth */
    free(prev_dirty); free(nil);
}

void orc_merkle_proofs(const uint8_t *leaves, const uint64_t *dirty, const uint8_t *nodes, size_t capacity, int depth,
                       const uint8_t *nil_leaf, const uint32_t *keys, size_t nkeys, uint8_t *out) {
    uint8_t (*nil)[32] = (uint8_t (*)[32])malloc(32 * (depth + 1));
    memcpy(nil[0], nil_leaf, 32);
    for (int l = 1; l <= depth; l++) node_hash(nil[l - 1], nil[l - 1], nil[l]);
    for (size_t k = 0; k < nkeys; k++) {
        size_t pos = keys[k];
        const uint8_t *lvl = nodes;
        for (int l = 0; l < depth; l++) {
            size_t sib = pos ^ 1;
            const uint8_t *src;
            if (l == 0) src = (sib < capacity && bit_get(dirty, sib)) ? leaves + 32 * sib : nil[0];
            else { size_t len = orc_merkle_level_len(capacity, l); src = sib < len ? lvl + 32 * sib : nil[l]; lvl += 32 * len; }
            memcpy(out + 32 * (k * depth + l), src, 32);
            pos >>= 1;
        }
    }
    free(nil);
}

/* AccountInfoToHash over already padded flat assets (PaddingAccountAssets is host logic, oracle/py/merkle.py):
 * leaf = Poseidon5(id, equity, debt, collateral, Poseidon(pack_triples(flat))) */
void orc_account_leaves(const uint8_t *ids, const uint8_t *totals, const uint64_t *flat, size_t n, int tier, uint8_t *out, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    for (int t = 2; t <= MAX_T; t++) pc_get(t);
    size_t nflat = (size_t)tier * 6, nel = (nflat + 2) / 3;
    #pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t a = 0; a < n; a++) {
        fe *e = (fe *)malloc(sizeof(fe) * nel), in5[5];
        const uint64_t *f = flat + a * nflat;
        for (size_t i = 0; i < nel; i++) {   /* a*2^128 + b*2^64 + c  (< 2^192 < r) */
            fe p = {{3 * i + 2 < nflat ? f[3 * i + 2] : 0, 3 * i + 1 < nflat ? f[3 * i + 1] : 0, f[3 * i], 0}};
            fe_to_mont(&e[i], &p, &ORC_FR);
        }
        hash_fe(e, nel, &in5[4]);
        fe_from_be(&in5[0], ids + 32 * a);
        for (int k = 0; k < 3; k++) fe_from_be(&in5[1 + k], totals + 96 * a + 32 * k);
        fe o; hash_fe(in5, 5, &o); fe_to_be(out + 32 * a, &o);
        free(e);
    }
}
