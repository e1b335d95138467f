// Groth16 Prove on the GPU with the proving key resident in HBM.
//
// Replaces the body of gnark's groth16.Prove (backend/groth16/bn254/prove.go, out of tree) after the constraint
// solver has run -- call site src/prover/prover/prover.go:269 -- and gnark-crypto's pedersen Commit/ProveKnowledge:
//   commitment = MSM(Basis, committed)                 pok = MSM(BasisExpSigma, committed)
//   h   = computeH(a, b, c)                            (bit-reversed, pairs with pk.G1.Z as gnark stores it)
//   Ar  = MSM(A, wA) + alpha1 + r*delta1               Bs1 = MSM(B1, wB) + beta1 + s*delta1
//   Bs  = MSM(B2, wB) + s*delta2 + beta2
//   Krs = MSM(K, wK) + MSM(Z, h[:n-1]) + (-r*s)*delta1 + s*Ar + r*Bs1
// wA / wB drop the wires whose A / B query is the point at infinity (pk.InfinityA/B); wK drops the public wires, the
// committed wires and the commitment wire.  Output = proof.WriteRawTo bytes.
// A, B1, B2 and K multiply (subsets of) the SAME wire vector, so the digits and the counting sort by bucket are computed
// once over all wires; each multiplication then takes its own view of the shared lists (msm_view: skip bitmap + rank map).
#include "internal.h"
#include <algorithm>
#include <cstdlib>
#include <type_traits>

using namespace ff;
using namespace ec;


namespace zk {

__global__ void k_gather_fr(const Fr *__restrict__ src, const uint32_t *__restrict__ idx, uint64_t n, Fr *__restrict__ dst) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *s = reinterpret_cast<const uint4 *>(src + idx[i]);
    uint4 *d = reinterpret_cast<uint4 *>(dst + i);
    d[0] = __ldg(s); d[1] = __ldg(s + 1);
}

template <class P>
__global__ void k_gather_points(const P *__restrict__ src, const uint32_t *__restrict__ idx, uint64_t n, P *__restrict__ dst) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr uint32_t Q = sizeof(P) / 16;
    if (t >= n * Q) return;
    const uint64_t i = t / Q; const uint32_t q = (uint32_t)(t % Q);
    reinterpret_cast<uint4 *>(dst + i)[q] = __ldg(reinterpret_cast<const uint4 *>(src + idx[i]) + q);
}

static int32_t upload(void **dst, const void *src, size_t bytes) {
    *dst = nullptr;
    if (bytes == 0) return ZKPOR_OK;
    ZK_CUDA(cudaMalloc(dst, bytes));
    ZK_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyDefault));
    return ZKPOR_OK;
}

// ZKPOR_OVERLAP_NTT=1 runs computeH on the copy stream concurrently with the wire-side sorts and multiplications.  Measured on
// B200: 1 114 vs 1 120 ms per proof -- the sort kernels fill every thread slot of the SMs, so the NTT blocks only trickle in and the
// accumulations slow down by what the NTT gains.  Off by default.
static bool overlap_ntt() {
    static const bool on = [] { const char *v = getenv("ZKPOR_OVERLAP_NTT"); return v != nullptr && atoi(v) != 0; }();
    return on;
}

struct ProofParts { G1XYZZ ar, bs1, krs_k, krs_z, commit, pok; G2XYZZ bs2; };

// gnark prove.go tail: blinding, Krs assembly, Jacobian -> affine, WriteRawTo layout
static void assemble_proof(const ProofParts &pp, const G1Affine &alpha1, const G1Affine &beta1, const G1Affine &delta1, const G2Affine &beta2,
                           const G2Affine &delta2, const uint8_t r_be[32], const uint8_t s_be[32], bool has_commitment, uint8_t *out,
                           uint32_t *out_len) {
    Fr r_plain, s_plain;
    fe_from_be32(&r_plain, r_be); fe_from_be32(&s_plain, s_be);
    // gnark draws r, s with fr.Element.SetRandom: canonical by construction; bytes >= the modulus are reduced (fr.SetBytes semantics)
    r_plain = Fr::from_mont(Fr::to_mont(r_plain)); s_plain = Fr::from_mont(Fr::to_mont(s_plain));
    Fr kr = Fr::from_mont(Fr::neg(Fr::mul(Fr::to_mont(r_plain), Fr::to_mont(s_plain))));   // -(r*s), plain
    G1XYZZ d1 = G1XYZZ::from_affine(delta1);
    G1XYZZ ar = pp.ar; ar.add_affine(alpha1, false); ar.add(d1.mul_256(r_plain.l));
    G1XYZZ bs1 = pp.bs1; bs1.add_affine(beta1, false); bs1.add(d1.mul_256(s_plain.l));
    G2XYZZ bs2 = pp.bs2; bs2.add(G2XYZZ::from_affine(delta2).mul_256(s_plain.l)); bs2.add_affine(beta2, false);
    G1XYZZ krs = pp.krs_k; krs.add(d1.mul_256(kr.l)); krs.add(pp.krs_z);
    krs.add(ar.mul_256(s_plain.l)); krs.add(bs1.mul_256(r_plain.l));
    g1_to_raw_bytes(out, ar.to_affine());
    g2_to_raw_bytes(out + 64, bs2.to_affine());
    g1_to_raw_bytes(out + 192, krs.to_affine());
    out[256] = 0; out[257] = 0; out[258] = 0; out[259] = has_commitment ? 1 : 0;
    uint32_t len = 260;
    if (has_commitment) {
        g1_to_raw_bytes(out + 260, pp.commit.to_affine()); len += 64;
        g1_to_raw_bytes(out + len, pp.pok.to_affine()); len += 64;
    } else {
        // gnark writes CommitmentPok unconditionally: the zero point
        g1_to_raw_bytes(out + len, G1Affine::inf()); len += 64;
    }
    *out_len = len;
}

static int32_t combine_commit(zkpor_ctx *ctx, zkpor_pk *pk, G1XYZZ *commit, G1XYZZ *pok);

int32_t pk_commit_and_pok(zkpor_ctx *ctx, zkpor_pk *pk, const Fr *d_wires, G1XYZZ *commit, G1XYZZ *pok) {
    *commit = G1XYZZ::inf(); *pok = G1XYZZ::inf();
    if (!pk->has_commitment) return ZKPOR_OK;
    if (pk->n_ck == 0) return combine_commit(ctx, pk, commit, pok);
    ZK_TRY(pk->sub.reserve(pk->n_ck * 32));
    Fr *sub = pk->sub.as<Fr>();
    MsmSorted srt;
    ZK_LAUNCH(ctx, k_gather_fr, grid_for(pk->n_ck, 256), 256, 0, d_wires, (const uint32_t *)pk->idx_c, pk->n_ck, sub);
    ZK_TRY(msm_sort(ctx, sub, pk->n_ck, ZKPOR_SCALARS_MONT, &srt));
    ZK_TRY(msm_accumulate_g1(ctx, pk->ck, srt, commit));
    ZK_TRY(msm_accumulate_g1(ctx, pk->ck_sigma, srt, pok));
    return combine_commit(ctx, pk, commit, pok);
}

// sharded key: every rank has summed its share of the committed wires; all ranks need the whole commitment (the challenge wire)
static int32_t combine_commit(zkpor_ctx *ctx, zkpor_pk *pk, G1XYZZ *commit, G1XYZZ *pok) {
    if (pk->shard_world <= 1) return ZKPOR_OK;
    G1XYZZ mine[2] = {*commit, *pok}, all[2 * 8];
    ZK_TRY(comm_all_gather_host(ctx, mine, all, sizeof mine));
    *commit = G1XYZZ::inf(); *pok = G1XYZZ::inf();
    for (int j = 0; j < pk->shard_world; j++) { commit->add(all[2 * j]); pok->add(all[2 * j + 1]); }
    return ZKPOR_OK;
}

static void tail_key_free(zkpor_pk *pk) {
    zkpor_pk::TailKey &t = pk->tail;
    for (void *q : {(void *)t.w_a, (void *)t.w_b, (void *)t.w_k, (void *)t.A, (void *)t.B1, (void *)t.K, (void *)t.B2}) if (q) cudaFree(q);
    t = zkpor_pk::TailKey();
}

// The A / B / K key points of the wires a program's deferred tail solves (within this key's wire range), as compact arrays: the tail's
// contribution to the three wire multiplications is a small multiplication of its own once the tail has finished (Sum s_i P_i is
// linear in the wire vector).  Built on the first proof of a (key, program) pair.
static int32_t tail_key_build(zkpor_ctx *ctx, zkpor_pk *pk, const SolverTail &tail) {
    if (pk->tail.prog_uid == tail.uid) return ZKPOR_OK;
    tail_key_free(pk);
    const uint64_t words = (pk->n_wires + 31) / 32;
    std::vector<uint2> maps[3];
    const uint2 *dmap[3] = {pk->map_a, pk->map_b, pk->map_k};
    for (int k = 0; k < 3; k++) { maps[k].resize(words); if (words) ZK_CUDA(cudaMemcpy(maps[k].data(), dmap[k], words * sizeof(uint2), cudaMemcpyDeviceToHost)); }
    std::vector<uint32_t> wid[3], pid[3];
    for (uint32_t wire : *tail.wires) {
        if (wire < pk->wire_first || wire >= pk->wire_first + pk->n_wires) continue;
        const uint64_t j = wire - pk->wire_first; const uint32_t bit = (uint32_t)(j & 31);
        for (int k = 0; k < 3; k++) {
            const uint2 m = maps[k][j >> 5];
            if ((m.x >> bit) & 1u) continue;
            wid[k].push_back(wire); pid[k].push_back(m.y + (uint32_t)__builtin_popcount(~m.x & ((1u << bit) - 1u)));
        }
    }
    zkpor_pk::TailKey &t = pk->tail;
    t.n_a = wid[0].size(); t.n_b = wid[1].size(); t.n_k = wid[2].size();
    uint32_t **wdst[3] = {&t.w_a, &t.w_b, &t.w_k};
    for (int k = 0; k < 3; k++) {
        const uint64_t n = wid[k].size();
        if (n == 0) continue;
        ZK_TRY(upload((void **)wdst[k], wid[k].data(), n * 4));
        uint32_t *d_pid = nullptr;
        ZK_TRY(upload((void **)&d_pid, pid[k].data(), n * 4));
        int32_t rc = ZKPOR_OK;
        auto gather = [&](auto **dst, const auto *src) {
            using P = std::remove_pointer_t<std::remove_pointer_t<decltype(dst)>>;
            if (cudaMalloc((void **)dst, n * sizeof(P)) != cudaSuccess) { set_error("prove: out of device memory for the tail's key points"); rc = ZKPOR_ERR_OOM; return; }
            k_gather_points<P><<<grid_for(n * (sizeof(P) / 16), 256), 256, 0, ctx->stream>>>(src, d_pid, n, *dst);
        };
        if (k == 0) gather(&t.A, pk->A);
        if (k == 1) { gather(&t.B1, pk->B1); if (rc == ZKPOR_OK) gather(&t.B2, pk->B2); }
        if (k == 2) gather(&t.K, pk->K);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_pid);
        ZK_TRY(rc);
    }
    t.prog_uid = tail.uid;
    return ZKPOR_OK;
}

// adds the tail wires' terms to Ar, Bs1, Bs2 and Krs-K (the shared sort saw those wires as zero)
static int32_t tail_delta(zkpor_ctx *ctx, zkpor_pk *pk, const Fr *dw, ProofParts &pp) {
    const zkpor_pk::TailKey &t = pk->tail;
    ZK_TRY(pk->sub.reserve(std::max({t.n_a, t.n_b, t.n_k, pk->n_ck, (uint64_t)1}) * 32));
    Fr *sub = pk->sub.as<Fr>();
    MsmSorted srt;
    if (t.n_a) {
        G1XYZZ r;
        ZK_LAUNCH(ctx, k_gather_fr, grid_for(t.n_a, 256), 256, 0, dw, (const uint32_t *)t.w_a, t.n_a, sub);
        ZK_TRY(msm_g1_dev(ctx, t.A, sub, t.n_a, ZKPOR_SCALARS_MONT, &r));
        pp.ar.add(r);
    }
    if (t.n_b) {
        G1XYZZ r; G2XYZZ r2;
        ZK_LAUNCH(ctx, k_gather_fr, grid_for(t.n_b, 256), 256, 0, dw, (const uint32_t *)t.w_b, t.n_b, sub);
        ZK_TRY(msm_sort(ctx, sub, t.n_b, ZKPOR_SCALARS_MONT, &srt));
        ZK_TRY(msm_accumulate_g1(ctx, t.B1, srt, &r));
        ZK_TRY(msm_accumulate_g2(ctx, t.B2, srt, &r2));
        pp.bs1.add(r); pp.bs2.add(r2);
    }
    if (t.n_k) {
        G1XYZZ r;
        ZK_LAUNCH(ctx, k_gather_fr, grid_for(t.n_k, 256), 256, 0, dw, (const uint32_t *)t.w_k, t.n_k, sub);
        ZK_TRY(msm_g1_dev(ctx, t.K, sub, t.n_k, ZKPOR_SCALARS_MONT, &r));
        pp.krs_k.add(r);
    }
    return ZKPOR_OK;
}

// ZKPOR_DEFER_TAIL=0: the solver's tail runs in place (measurement knob; the proof is the same)
static bool defer_tail_enabled() {
    static const bool on = [] { const char *v = getenv("ZKPOR_DEFER_TAIL"); return v == nullptr || atoi(v) != 0; }();
    return on;
}

}  // namespace zk

using namespace zk;

extern "C" {

int32_t zkpor_pk_free(zkpor_ctx *ctx, zkpor_pk *pk) {
    (void)ctx;
    if (!pk) return ZKPOR_OK;
    void *ptrs[] = {pk->A, pk->B1, pk->K, pk->Z, pk->ck, pk->ck_sigma, pk->B2, pk->idx_c, pk->map_a, pk->map_b, pk->map_k};
    for (void *p : ptrs) if (p) cudaFree(p);
    tail_key_free(pk);
    pk->wires.release(); pk->sub.release();
    delete pk;
    return ZKPOR_OK;
}

// Uploads the part of the key that multiplies wires [w0, w1) (w0 a multiple of 32), the committed wires number [c0, c1) and
// Z[z0, z1); the whole key is the range (0, n_wires, 0, n_committed, 0, n_z).  `d` always describes the WHOLE key.
static int32_t pk_upload_range(zkpor_ctx *ctx, const zkpor_pk_desc *d, uint64_t w0, uint64_t w1, uint64_t c0, uint64_t c1, uint64_t z0, uint64_t z1,
                               int rank, int world, zkpor_pk **out) {
    ZK_REQUIRE(ctx && d && out, "pk_upload: null argument");
    ZK_REQUIRE(d->log_n >= 1 && d->log_n <= 28, "pk_upload: log_n out of range");
    ZK_REQUIRE(d->g1_alpha && d->g1_beta && d->g1_delta && d->g2_beta && d->g2_delta, "pk_upload: missing alpha/beta/delta");
    ZK_REQUIRE(d->n_wires < (1ull << 32), "pk_upload: too many wires");
    ZK_REQUIRE(d->n_committed == 0 || (d->ck_basis && d->ck_basis_exp_sigma && (d->n_wires == 0 || d->private_committed)),
               "pk_upload: n_committed > 0 needs ck_basis, ck_basis_exp_sigma and (with wire maps) private_committed");
    ZK_REQUIRE(world == 1 || d->n_wires > 0, "pk_upload_shard: the key needs its infinity maps");
    ZK_CUDA(cudaSetDevice(ctx->device));
    zkpor_pk *pk = new zkpor_pk();
    *out = nullptr;
    pk->log_n = d->log_n; pk->n_wires = w1 - w0; pk->n_wires_total = d->n_wires; pk->wire_first = w0; pk->z_first = z0;
    pk->shard_rank = rank; pk->shard_world = world;
    pk->n_z = z1 - z0; pk->n_ck = c1 - c0;
    pk->has_commitment = d->n_committed > 0 || d->ck_basis != nullptr;
    memcpy(&pk->alpha1, d->g1_alpha, 64); memcpy(&pk->beta1, d->g1_beta, 64); memcpy(&pk->delta1, d->g1_delta, 64);
    memcpy(&pk->beta2, d->g2_beta, 128); memcpy(&pk->delta2, d->g2_delta, 128);
    int32_t rc = ZKPOR_OK;
    auto up = [&](void **dst, const void *src, size_t first, size_t count, size_t elem) {
        if (rc == ZKPOR_OK) rc = upload(dst, (const uint8_t *)src + first * elem, count * elem);
    };
    uint64_t ra0 = 0, rb0 = 0, rk0 = 0, ra1 = d->n_a, rb1 = d->n_b, rk1 = d->n_k;   // this range's slice of the compact arrays
    if (d->n_wires > 0) {
        if (!d->infinity_a || !d->infinity_b) { set_error("pk_upload: infinity maps missing"); rc = ZKPOR_ERR_INVALID_ARG; }
        else {
            std::vector<uint32_t> ic;
            std::vector<uint8_t> drop(d->n_wires, 0);
            for (uint64_t i = 0; i < d->n_committed; i++) {
                if (d->private_committed[i] >= d->n_wires) { set_error("pk_upload: committed wire out of range"); rc = ZKPOR_ERR_INVALID_ARG; break; }
                drop[d->private_committed[i]] = 1;
                if (i >= c0 && i < c1) ic.push_back((uint32_t)d->private_committed[i]);
            }
            if (rc == ZKPOR_OK && pk->has_commitment) {
                if (d->commitment_index >= d->n_wires) { set_error("pk_upload: commitment wire out of range"); rc = ZKPOR_ERR_INVALID_ARG; }
                else drop[d->commitment_index] = 1;
            }
            if (rc == ZKPOR_OK) {
                // per 32 wires: skip bits and the rank (index in this range's compact key array) of the first wire of the word
                const uint64_t words_all = (d->n_wires + 31) / 32, word0 = w0 / 32, words = (w1 - w0 + 31) / 32;
                std::vector<uint2> ma(words), mb(words), mk(words);
                uint64_t ra = 0, rb = 0, rk = 0;
                for (uint64_t w = 0; w < words_all; w++) {
                    uint32_t ba = 0, bb = 0, bk = 0;
                    for (uint32_t j = 0; j < 32; j++) {
                        const uint64_t i = w * 32 + j;
                        const bool in = i < d->n_wires;
                        if (!in || d->infinity_a[i]) ba |= 1u << j;
                        if (!in || d->infinity_b[i]) bb |= 1u << j;
                        if (!in || i < d->n_public || drop[i]) bk |= 1u << j;
                    }
                    if (w == word0) { ra0 = ra; rb0 = rb; rk0 = rk; }
                    if (w == word0 + words) { ra1 = ra; rb1 = rb; rk1 = rk; }
                    if (w >= word0 && w < word0 + words) {
                        const uint64_t k = w - word0;
                        ma[k] = make_uint2(ba, (uint32_t)(ra - ra0)); mb[k] = make_uint2(bb, (uint32_t)(rb - rb0)); mk[k] = make_uint2(bk, (uint32_t)(rk - rk0));
                    }
                    ra += 32 - __builtin_popcount(ba); rb += 32 - __builtin_popcount(bb); rk += 32 - __builtin_popcount(bk);
                }
                if (word0 >= words_all) { ra0 = ra; rb0 = rb; rk0 = rk; }
                if (word0 + words >= words_all) { ra1 = ra; rb1 = rb; rk1 = rk; }
                if (ra != d->n_a || rb != d->n_b || rk != d->n_k) {
                    set_error("pk_upload: key sizes inconsistent with infinity/commitment maps (A %llu/%llu, B %llu/%llu, K %llu/%llu)", (unsigned long long)ra,
                              (unsigned long long)d->n_a, (unsigned long long)rb, (unsigned long long)d->n_b, (unsigned long long)rk, (unsigned long long)d->n_k);
                    rc = ZKPOR_ERR_INVALID_ARG;
                }
                up((void **)&pk->map_a, ma.data(), 0, words, sizeof(uint2)); up((void **)&pk->map_b, mb.data(), 0, words, sizeof(uint2));
                up((void **)&pk->map_k, mk.data(), 0, words, sizeof(uint2));
            }
            up((void **)&pk->idx_c, ic.data(), 0, ic.size(), 4);
        }
    }
    pk->n_a = ra1 - ra0; pk->n_b = rb1 - rb0; pk->n_k = rk1 - rk0;
    up((void **)&pk->A, d->g1_a, ra0, pk->n_a, 64); up((void **)&pk->B1, d->g1_b, rb0, pk->n_b, 64); up((void **)&pk->K, d->g1_k, rk0, pk->n_k, 64);
    up((void **)&pk->Z, d->g1_z, z0, pk->n_z, 64); up((void **)&pk->B2, d->g2_b, rb0, pk->n_b, 128);
    up((void **)&pk->ck, d->ck_basis, c0, pk->n_ck, 64); up((void **)&pk->ck_sigma, d->ck_basis_exp_sigma, c0, pk->n_ck, 64);
    if (rc == ZKPOR_OK && d->n_z != (1ull << d->log_n) - 1 && d->n_wires > 0) { set_error("pk_upload: len(Z) must be 2^log_n - 1"); rc = ZKPOR_ERR_INVALID_ARG; }
    if (rc != ZKPOR_OK) { zkpor_pk_free(ctx, pk); return rc; }
    *out = pk;
    return ZKPOR_OK;
}

int32_t zkpor_pk_upload(zkpor_ctx *ctx, const zkpor_pk_desc *d, zkpor_pk **out) {
    ZK_REQUIRE(ctx && d && out, "pk_upload: null argument");
    return pk_upload_range(ctx, d, 0, d->n_wires, 0, d->n_committed, 0, d->n_z, 0, 1, out);
}

// point-chunk split of the key over the ranks of the context's group: wires in `world` contiguous ranges (multiples of 32), the committed
// wires by count, Z in the n/world chunks that computeH's sharded transform leaves on each rank
int32_t zkpor_pk_upload_shard(zkpor_ctx *ctx, const zkpor_pk_desc *d, zkpor_pk **out) {
    ZK_REQUIRE(ctx && d && out, "pk_upload_shard: null argument");
    int rank, world; comm_info(ctx, &rank, &world);
    ZK_REQUIRE(d->log_n >= 1 && d->log_n <= 28 && ((uint64_t)1 << d->log_n) >= (uint64_t)world * world * 2, "pk_upload_shard: domain too small for this many ranks");
    const uint64_t W = d->n_wires, per = (((W + world - 1) / world) + 31) & ~31ull;
    const uint64_t w0 = std::min(W, per * rank), w1 = std::min(W, w0 + per);
    const uint64_t c0 = d->n_committed * rank / world, c1 = d->n_committed * (rank + 1) / world;
    const uint64_t m = ((uint64_t)1 << d->log_n) / world, z0 = std::min(d->n_z, m * rank), z1 = std::min(d->n_z, m * (rank + 1));
    return pk_upload_range(ctx, d, w0, w1, c0, c1, z0, z1, rank, world, out);
}

int32_t zkpor_pk_shard_info(zkpor_pk *pk, uint64_t out8[8]) {
    ZK_REQUIRE(pk && out8, "pk_shard_info: null argument");
    out8[0] = pk->shard_rank; out8[1] = pk->shard_world; out8[2] = pk->wire_first; out8[3] = pk->n_wires; out8[4] = pk->n_a; out8[5] = pk->n_b;
    out8[6] = pk->n_k; out8[7] = pk->n_z;
    return ZKPOR_OK;
}

int32_t zkpor_pk_commit(zkpor_ctx *ctx, zkpor_pk *pk, const void *committed_values, void *out_affine64) {
    ZK_REQUIRE(ctx && pk && committed_values && out_affine64, "pk_commit: null argument");
    ZK_REQUIRE(pk->has_commitment, "pk_commit: key has no commitment");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    const void *ds;
    ZK_TRY(to_device(ctx, committed_values, pk->n_ck * 32, ctx->in_scalars, &ds));
    G1XYZZ r; ZK_TRY(msm_g1_dev(ctx, pk->ck, ds, pk->n_ck, ZKPOR_SCALARS_MONT, &r));
    G1Affine a = r.to_affine(); memcpy(out_affine64, &a, 64);
    stages_collect(ctx);
    return ZKPOR_OK;
}

// shared body of zkpor_groth16_prove (a, b, c from the caller) and zkpor_groth16_prove_wires (a, b, c = L w, R w, O w on the device)
// and zkpor_groth16_prove_solve (the wires themselves come from the device solver: `prog` set, `wires` = the circuit's inputs)
static int32_t prove_body(zkpor_ctx *ctx, zkpor_pk *pk, zkpor_r1cs *cs, zkpor_program *prog, const void *wires, const void *a, const void *b,
                          const void *c, uint64_t n_constraints, const uint8_t r_be[32], const uint8_t s_be[32], uint8_t *out_proof, uint32_t *out_len) {
    ZK_REQUIRE(pk->n_wires_total > 0, "prove: key was uploaded without wire maps (a key for zkpor_groth16_prove_partial?)");
    int rank, world; comm_info(ctx, &rank, &world);
    ZK_REQUIRE(pk->shard_world == world && pk->shard_rank == rank, "prove: the key's shard does not match the context's rank in its group");
    ZK_REQUIRE(world == 1 || cs != nullptr, "prove: one proof across several GPUs needs the constraint system on the device (prove_wires / prove_solve)");
    const size_t n = (size_t)1 << pk->log_n, m = n / (size_t)world;   // m: this rank's share of the domain
    ZK_REQUIRE(n_constraints > 0 && n_constraints <= n, "prove: n_constraints exceeds the domain");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    // wires -> HBM on the compute stream (every wire-only MSM needs them); a, b, c -> padded device vectors on the
    // copy stream, so that their H2D transfer overlaps the commitment / A / B / K multi-scalar multiplications
    const void *dw;
    ProofParts pp;
    pp.commit = G1XYZZ::inf(); pp.pok = G1XYZZ::inf();
    bool commit_done = false, deferred = false;
    SolverTail tail{};
    if (prog != nullptr) {
        // r1cs.Solve on the device: inputs -> wires[1 ..], every other wire by the level schedule; the commitment hint leaves the
        // commitment and its proof of knowledge behind.  Sharded: every rank solves the whole system (the schedule is a latency
        // chain, not throughput work) and contributes its share of the commitment.
        ZK_TRY(pk->wires.reserve(pk->n_wires_total * 32));
        Fr *w = pk->wires.as<Fr>();
        const Fr one = Fr::one();
        stage_begin(ctx, ST_H2D);
        ZK_CUDA(cudaMemcpyAsync(w, &one, 32, cudaMemcpyHostToDevice, ctx->stream));
        ZK_CUDA(cudaMemcpyAsync(w + 1, wires, program_inputs(prog) * 32, cudaMemcpyDefault, ctx->stream));
        stage_end(ctx, ST_H2D);
        stage_begin(ctx, ST_SOLVE);
        deferred = defer_tail_enabled() && solver_tail_info(prog, &tail);
        if (deferred) ZK_TRY(tail_key_build(ctx, pk, tail));
        ZK_TRY(solver_run(ctx, prog, pk, w, &pp.commit, &pp.pok, &commit_done, deferred));
        if (deferred) stage_end(ctx, ST_SOLVE);   // the head of the schedule; the tail runs beside the multiplications below
        dw = w;
    } else {
        stage_begin(ctx, ST_H2D);
        ZK_TRY(to_device(ctx, wires, pk->n_wires_total * 32, pk->wires, &dw));
        stage_end(ctx, ST_H2D);
    }
    const size_t bytes = m * sizeof(Fr), in_bytes = n_constraints * sizeof(Fr);
    ZK_TRY(ctx->ntt_a.reserve(bytes)); ZK_TRY(ctx->ntt_b.reserve(bytes)); ZK_TRY(ctx->ntt_c.reserve(bytes));
    const void *src[3] = {a, b, c};
    Fr *dst[3] = {ctx->ntt_a.as<Fr>(), ctx->ntt_b.as<Fr>(), ctx->ntt_c.as<Fr>()};
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));   // the previous call's use of the NTT buffers is over
    // a = L w, b = R w, c = O w on the device (and, after a solve, gnark's satisfaction check): needs every wire, so with a deferred
    // tail it runs after the wire multiplications instead of ahead of them
    auto eval_abc = [&]() -> int32_t {
        if (world > 1) {
            // this rank's rows rank, rank + world, ... of a, b, c: the cyclic split the sharded transform starts from
            ZK_TRY(r1cs_eval_strided_dev(ctx, cs, (const Fr *)dw, dst[0], dst[1], dst[2], (uint64_t)rank, (uint64_t)world, m));
            int32_t ok_all[8], ok_mine = prog != nullptr ? r1cs_check_dev(ctx, dst[0], dst[1], dst[2], m) : ZKPOR_OK;
            if (prog != nullptr) {
                ZK_TRY(comm_all_gather_host(ctx, &ok_mine, ok_all, sizeof ok_mine));
                for (int j = 0; j < world; j++)
                    if (ok_all[j] != ZKPOR_OK) { if (ok_mine == ZKPOR_OK) set_error("solve: a constraint is not satisfied (found by rank %d)", j); return ok_all[j]; }
                if (!deferred) stage_end(ctx, ST_SOLVE);
            }
            return ZKPOR_OK;
        }
        for (int k = 0; k < 3; k++) if (bytes > in_bytes) ZK_CUDA(cudaMemsetAsync((uint8_t *)dst[k] + in_bytes, 0, bytes - in_bytes, ctx->stream));
        ZK_TRY(r1cs_eval_dev(ctx, cs, (const Fr *)dw, dst[0], dst[1], dst[2]));
        if (prog != nullptr) {   // gnark's Solve fails on the first unsatisfied constraint; here one pass over a, b, c
            ZK_TRY(r1cs_check_dev(ctx, dst[0], dst[1], dst[2], n_constraints));
            if (!deferred) stage_end(ctx, ST_SOLVE);
        }
        return ZKPOR_OK;
    };
    if (world > 1) ZK_TRY(ctx->dist_tmp.reserve(bytes));
    if (cs != nullptr) {
        if (!deferred) ZK_TRY(eval_abc());
    } else {
        for (int k = 0; k < 3; k++) {
            ZK_CUDA(cudaMemcpyAsync(dst[k], src[k], in_bytes, cudaMemcpyDefault, ctx->copy_stream));
            if (bytes > in_bytes) ZK_CUDA(cudaMemsetAsync((uint8_t *)dst[k] + in_bytes, 0, bytes - in_bytes, ctx->copy_stream));
        }
    }
    // computeH runs on the copy stream, right behind the transfers of a, b, c, concurrently with the wire-side sorts and
    // multiplications of the compute stream: the counting sorts are bound by L2 atomics and scattered stores and leave the
    // integer pipe idle, which the NTT butterflies fill.  The Z multiplication waits for h (copy_done).
    const bool overlap = overlap_ntt() && world == 1 && !deferred;
    if (overlap) {
        if (cs != nullptr) {   // a, b, c were produced on the compute stream
            ZK_CUDA(cudaEventRecord(ctx->copy_done, ctx->stream));
            ZK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done, 0));
        }
        cudaStream_t compute_stream = ctx->stream;
        ctx->stream = ctx->copy_stream;
        const int32_t rc_h = compute_h_dev(ctx, dst[0], dst[1], dst[2], pk->log_n);
        ctx->stream = compute_stream;
        ZK_TRY(rc_h);
    }
    ZK_CUDA(cudaEventRecord(ctx->copy_done, ctx->copy_stream));

    MsmSorted srt;
    if (!commit_done) ZK_TRY(pk_commit_and_pok(ctx, pk, (const Fr *)dw, &pp.commit, &pp.pok));
    pp.ar = G1XYZZ::inf(); pp.bs1 = G1XYZZ::inf(); pp.bs2 = G2XYZZ::inf(); pp.krs_k = G1XYZZ::inf(); pp.krs_z = G1XYZZ::inf();
    // one digit extraction + counting sort over (this rank's range of) the wire vector, four accumulations through their wire maps
    const Fr *dw_mine = (const Fr *)dw + pk->wire_first;
    if (pk->n_wires) {
        if (deferred) ctx->scalar_mask = tail.mask + pk->wire_first / 32;   // wire ranges start at multiples of 32
        ZK_TRY(msm_sort(ctx, dw_mine, pk->n_wires, ZKPOR_SCALARS_MONT, &srt));
        MsmSorted view;
        if (pk->n_a) {
            ZK_TRY(msm_view(ctx, srt, pk->map_a, &view));
            ZK_TRY(msm_accumulate_g1(ctx, pk->A, view, &pp.ar, pk->n_a));
        }
        if (pk->n_b) {
            ZK_TRY(msm_view(ctx, srt, pk->map_b, &view));      // one view, two accumulations (G1 and G2)
            ZK_TRY(msm_accumulate_g1(ctx, pk->B1, view, &pp.bs1, pk->n_b));
            ZK_TRY(msm_accumulate_g2(ctx, pk->B2, view, &pp.bs2, pk->n_b));
        }
        if (pk->n_k) {
            ZK_TRY(msm_view(ctx, srt, pk->map_k, &view));
            ZK_TRY(msm_accumulate_g1(ctx, pk->K, view, &pp.krs_k, pk->n_k));
        }
    }
    if (deferred) {
        ZK_TRY(solver_tail_join(ctx, prog));
        ZK_TRY(tail_delta(ctx, pk, (const Fr *)dw, pp));
        ZK_TRY(eval_abc());
    }
    ZK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
    if (world > 1) ZK_TRY(compute_h_dist(ctx, dst[0], dst[1], dst[2], ctx->dist_tmp.as<Fr>(), pk->log_n));
    else if (!overlap) ZK_TRY(compute_h_dev(ctx, dst[0], dst[1], dst[2], pk->log_n));
    if (pk->n_z) ZK_TRY(msm_g1_dev(ctx, pk->Z, dst[0], pk->n_z, ZKPOR_SCALARS_MONT, &pp.krs_z));
    if (world > 1) {
        // the one collective of the multiplications: 4 G1 + 1 G2 partial sums per rank, every rank finishes the proof
        struct Part { G1XYZZ ar, bs1, krs_k, krs_z; G2XYZZ bs2; } mine = {pp.ar, pp.bs1, pp.krs_k, pp.krs_z, pp.bs2}, all[8];
        ZK_TRY(comm_all_gather_host(ctx, &mine, all, sizeof mine));
        pp.ar = G1XYZZ::inf(); pp.bs1 = G1XYZZ::inf(); pp.bs2 = G2XYZZ::inf(); pp.krs_k = G1XYZZ::inf(); pp.krs_z = G1XYZZ::inf();
        for (int j = 0; j < world; j++) { pp.ar.add(all[j].ar); pp.bs1.add(all[j].bs1); pp.krs_k.add(all[j].krs_k); pp.krs_z.add(all[j].krs_z); pp.bs2.add(all[j].bs2); }
    }
    assemble_proof(pp, pk->alpha1, pk->beta1, pk->delta1, pk->beta2, pk->delta2, r_be, s_be, pk->has_commitment, out_proof, out_len);
    stages_collect(ctx);
    return ZKPOR_OK;
}

// a failing rank of a sharded proof must not leave its peers waiting at a host barrier
static int32_t prove_impl(zkpor_ctx *ctx, zkpor_pk *pk, zkpor_r1cs *cs, zkpor_program *prog, const void *wires, const void *a, const void *b,
                          const void *c, uint64_t n_constraints, const uint8_t r_be[32], const uint8_t s_be[32], uint8_t *out_proof, uint32_t *out_len) {
    const int32_t rc = prove_body(ctx, pk, cs, prog, wires, a, b, c, n_constraints, r_be, s_be, out_proof, out_len);
    if (rc != ZKPOR_OK) { comm_abort(ctx); ctx->scalar_mask = nullptr; if (prog) solver_tail_abandon(ctx, prog); }
    return rc;
}

int32_t zkpor_groth16_prove(zkpor_ctx *ctx, zkpor_pk *pk, const void *wires, const void *a, const void *b, const void *c,
                            uint64_t n_constraints, const uint8_t r_be[32], const uint8_t s_be[32], uint8_t *out_proof, uint32_t *out_len) {
    ZK_REQUIRE(ctx && pk && wires && a && b && c && r_be && s_be && out_proof && out_len, "prove: null argument");
    return prove_impl(ctx, pk, nullptr, nullptr, wires, a, b, c, n_constraints, r_be, s_be, out_proof, out_len);
}

int32_t zkpor_groth16_prove_wires(zkpor_ctx *ctx, zkpor_pk *pk, zkpor_r1cs *cs, const void *wires, const uint8_t r_be[32], const uint8_t s_be[32],
                                  uint8_t *out_proof, uint32_t *out_len) {
    ZK_REQUIRE(ctx && pk && cs && wires && r_be && s_be && out_proof && out_len, "prove_wires: null argument");
    ZK_REQUIRE(r1cs_wires(cs) == pk->n_wires_total, "prove_wires: the constraint system and the key disagree on the number of wires");
    return prove_impl(ctx, pk, cs, nullptr, wires, nullptr, nullptr, nullptr, r1cs_rows(cs), r_be, s_be, out_proof, out_len);
}

int32_t zkpor_groth16_prove_solve(zkpor_ctx *ctx, zkpor_pk *pk, zkpor_program *prog, const void *inputs, const uint8_t r_be[32],
                                  const uint8_t s_be[32], uint8_t *out_proof, uint32_t *out_len) {
    ZK_REQUIRE(ctx && pk && prog && inputs && r_be && s_be && out_proof && out_len, "prove_solve: null argument");
    zkpor_r1cs *cs = program_matrices(prog);
    ZK_REQUIRE(r1cs_wires(cs) == pk->n_wires_total, "prove_solve: the program and the key disagree on the number of wires");
    return prove_impl(ctx, pk, cs, prog, inputs, nullptr, nullptr, nullptr, r1cs_rows(cs), r_be, s_be, out_proof, out_len);
}

int32_t zkpor_groth16_prove_partial(zkpor_ctx *ctx, zkpor_pk *pk, const void *wires_a, const void *wires_b, const void *wires_k,
                                    const void *committed, const void *h_chunk, uint64_t n_h, void *out_partials) {
    ZK_REQUIRE(ctx && pk && out_partials, "prove_partial: null argument");
    ZK_REQUIRE(n_h <= pk->n_z, "prove_partial: h chunk longer than the Z chunk");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    ProofParts pp;
    pp.ar = G1XYZZ::inf(); pp.bs1 = G1XYZZ::inf(); pp.bs2 = G2XYZZ::inf(); pp.krs_k = G1XYZZ::inf(); pp.krs_z = G1XYZZ::inf();
    pp.commit = G1XYZZ::inf(); pp.pok = G1XYZZ::inf();
    const void *ds; MsmSorted srt;
    if (pk->n_ck && committed) {
        ZK_TRY(to_device(ctx, committed, pk->n_ck * 32, ctx->in_scalars, &ds));
        ZK_TRY(msm_sort(ctx, ds, pk->n_ck, ZKPOR_SCALARS_MONT, &srt));
        ZK_TRY(msm_accumulate_g1(ctx, pk->ck, srt, &pp.commit));
        ZK_TRY(msm_accumulate_g1(ctx, pk->ck_sigma, srt, &pp.pok));
    }
    if (pk->n_a && wires_a) {
        ZK_TRY(to_device(ctx, wires_a, pk->n_a * 32, ctx->in_scalars, &ds));
        ZK_TRY(msm_g1_dev(ctx, pk->A, ds, pk->n_a, ZKPOR_SCALARS_MONT, &pp.ar));
    }
    if (pk->n_b && wires_b) {
        ZK_TRY(to_device(ctx, wires_b, pk->n_b * 32, ctx->in_scalars, &ds));
        ZK_TRY(msm_sort(ctx, ds, pk->n_b, ZKPOR_SCALARS_MONT, &srt));
        ZK_TRY(msm_accumulate_g1(ctx, pk->B1, srt, &pp.bs1));
        ZK_TRY(msm_accumulate_g2(ctx, pk->B2, srt, &pp.bs2));
    }
    if (pk->n_k && wires_k) {
        ZK_TRY(to_device(ctx, wires_k, pk->n_k * 32, ctx->in_scalars, &ds));
        ZK_TRY(msm_g1_dev(ctx, pk->K, ds, pk->n_k, ZKPOR_SCALARS_MONT, &pp.krs_k));
    }
    if (n_h && h_chunk) {
        ZK_TRY(to_device(ctx, h_chunk, n_h * 32, ctx->in_scalars, &ds));
        ZK_TRY(msm_g1_dev(ctx, pk->Z, ds, n_h, ZKPOR_SCALARS_MONT, &pp.krs_z));
    }
    uint8_t *o = (uint8_t *)out_partials;
    memcpy(o, &pp.ar, 128); memcpy(o + 128, &pp.bs1, 128); memcpy(o + 256, &pp.krs_k, 128); memcpy(o + 384, &pp.krs_z, 128);
    memcpy(o + 512, &pp.commit, 128); memcpy(o + 640, &pp.pok, 128); memcpy(o + 768, &pp.bs2, 256);
    stages_collect(ctx);
    return ZKPOR_OK;
}

int32_t zkpor_groth16_finish(const void *partials, uint32_t k, const void *g1_alpha, const void *g1_beta, const void *g1_delta,
                             const void *g2_beta, const void *g2_delta, const uint8_t r_be[32], const uint8_t s_be[32],
                             int32_t has_commitment, uint8_t *out_proof, uint32_t *out_len) {
    ZK_REQUIRE(partials && g1_alpha && g1_beta && g1_delta && g2_beta && g2_delta && r_be && s_be && out_proof && out_len, "finish: null argument");
    ProofParts pp;
    pp.ar = G1XYZZ::inf(); pp.bs1 = G1XYZZ::inf(); pp.bs2 = G2XYZZ::inf(); pp.krs_k = G1XYZZ::inf(); pp.krs_z = G1XYZZ::inf();
    pp.commit = G1XYZZ::inf(); pp.pok = G1XYZZ::inf();
    for (uint32_t i = 0; i < k; i++) {
        const uint8_t *p = (const uint8_t *)partials + (size_t)i * ZKPOR_PROVE_PARTIAL_BYTES;
        G1XYZZ g; G2XYZZ g2;
        memcpy(&g, p, 128); pp.ar.add(g); memcpy(&g, p + 128, 128); pp.bs1.add(g); memcpy(&g, p + 256, 128); pp.krs_k.add(g);
        memcpy(&g, p + 384, 128); pp.krs_z.add(g); memcpy(&g, p + 512, 128); pp.commit.add(g); memcpy(&g, p + 640, 128); pp.pok.add(g);
        memcpy(&g2, p + 768, 256); pp.bs2.add(g2);
    }
    G1Affine al, be, de; G2Affine be2, de2;
    memcpy(&al, g1_alpha, 64); memcpy(&be, g1_beta, 64); memcpy(&de, g1_delta, 64); memcpy(&be2, g2_beta, 128); memcpy(&de2, g2_delta, 128);
    assemble_proof(pp, al, be, de, be2, de2, r_be, s_be, has_commitment != 0, out_proof, out_len);
    return ZKPOR_OK;
}

}  // extern "C"
