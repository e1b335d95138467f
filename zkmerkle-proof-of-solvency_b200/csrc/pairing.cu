// groth16.Verify and the pairing product behind it.
// Replaces gnark backend/groth16/bn254/verify.go (out of tree) as called from src/prover/prover/prover.go:276 and
// src/verifier/main.go:284, and gnark-crypto bn254.PairingCheck / pedersen.VerifyingKey.Verify underneath it.
//   GPU: one Miller loop per thread (k_miller_loops), subgroup checks and the rho-scalings of the batch verifier.
//   Host: the product of the loop values and ONE final exponentiation per pairing product (O(1) work), SHA-256 for the
//   BSB22 commitment challenge (hash_to_field, RFC 9380 expand_message_xmd) and the 388-byte proof parsing.
#include "internal.h"
#include "pairing.cuh"

using namespace ff;
using namespace ec;
using pairing::Fp12;

namespace zk {

__global__ void __launch_bounds__(64) k_miller_loops(const G1Affine *__restrict__ P, const G2Affine *__restrict__ Q, uint64_t n, Fp12 *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pairing::miller_loop(P[i], Q[i]);
}

// out[i] = k_i * P_i for 128-bit plain scalars (4 x u32 each): the rho-scaling of the batch verifier
__global__ void __launch_bounds__(128) k_scale_g1_128(const G1Affine *__restrict__ P, const uint32_t *__restrict__ k, uint64_t n, G1Affine *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1XYZZ acc = G1XYZZ::inf();
    const G1Affine p = P[i];
    for (int b = 127; b >= 0; b--) {
        acc = acc.dbl();
        if ((k[4 * i + (b >> 5)] >> (b & 31)) & 1) acc.add_affine(p, false);
    }
    out[i] = acc.to_affine();
}

// ok[i] = 1 iff Q_i is on the twist and r * Q_i = infinity (G2Affine.IsInSubGroup)
__global__ void __launch_bounds__(64) k_g2_in_subgroup(const G2Affine *__restrict__ Q, uint64_t n, Fp2 twist_b, uint8_t *__restrict__ ok) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const G2Affine q = Q[i];
    if (q.is_inf()) { ok[i] = 1; return; }
    bool on = Fp2::sqr(q.y) == Fp2::add(Fp2::mul(Fp2::sqr(q.x), q.x), twist_b);
    uint32_t r[8];
    for (int k = 0; k < 8; k++) r[k] = FrParams::M(k);
    ok[i] = on && G2XYZZ::from_affine(q).mul_256(r).is_inf();
}

// ------------------------------------------------------------------------------------------------ host helpers
struct Sha256 {
    uint32_t h[8]; uint8_t buf[64]; uint64_t len = 0; uint32_t fill = 0;
    Sha256() { static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u}; memcpy(h, iv, 32); }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t *p) {
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u,
            0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
            0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u,
            0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
            0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
            0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const void *data, size_t n) {
        const uint8_t *p = (const uint8_t *)data;
        len += n;
        while (n) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take); fill += (uint32_t)take; p += take; n -= take;
            if (fill == 64) { block(buf); fill = 0; }
        }
    }
    void final(uint8_t out[32]) {
        uint64_t bits = len * 8;
        uint8_t pad = 0x80; update(&pad, 1);
        uint8_t z = 0; while (fill != 56) update(&z, 1);
        uint8_t lb[8]; for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
    }
};

// RFC 9380 expand_message_xmd with SHA-256, len_in_bytes <= 255*32
static void expand_msg_xmd(const uint8_t *msg, size_t msg_len, const char *dst, size_t dst_len, uint8_t *out, size_t out_len) {
    const size_t ell = (out_len + 31) / 32;
    uint8_t dst_prime_len = (uint8_t)dst_len, b0[32], bi[32], zpad[64] = {0};
    uint8_t lib[3] = {(uint8_t)(out_len >> 8), (uint8_t)out_len, 0};
    Sha256 s0; s0.update(zpad, 64); s0.update(msg, msg_len); s0.update(lib, 3); s0.update(dst, dst_len); s0.update(&dst_prime_len, 1); s0.final(b0);
    uint8_t one = 1;
    Sha256 s1; s1.update(b0, 32); s1.update(&one, 1); s1.update(dst, dst_len); s1.update(&dst_prime_len, 1); s1.final(bi);
    size_t done = 0;
    for (size_t i = 1; i <= ell; i++) {
        size_t take = out_len - done < 32 ? out_len - done : 32;
        memcpy(out + done, bi, take); done += take;
        if (i == ell) break;
        uint8_t x[32], idx = (uint8_t)(i + 1);
        for (int k = 0; k < 32; k++) x[k] = b0[k] ^ bi[k];
        Sha256 s; s.update(x, 32); s.update(&idx, 1); s.update(dst, dst_len); s.update(&dst_prime_len, 1); s.final(bi);
    }
}

// big-endian bytes (any length) -> Montgomery Fr, reduced mod r
static Fr fr_from_be_bytes(const uint8_t *p, size_t n) {
    const Fr k256 = Fr::from_u64(256);
    Fr acc = Fr::zero();
    for (size_t i = 0; i < n; i++) acc = Fr::add(Fr::mul(acc, k256), Fr::from_u64(p[i]));
    return acc;
}
static void fr_to_be32(uint8_t out[32], const Fr &mont) {
    Fr p = Fr::from_mont(mont);
    for (int i = 0; i < 8; i++) { uint32_t v = p.l[7 - i]; out[4 * i] = v >> 24; out[4 * i + 1] = v >> 16; out[4 * i + 2] = v >> 8; out[4 * i + 3] = v; }
}
// gnark-crypto fr.Hash(msg, dst, 1): 48 uniform bytes, big-endian, reduced mod r
static Fr hash_to_fr(const uint8_t *msg, size_t len, const char *dst) {
    uint8_t u[48];
    expand_msg_xmd(msg, len, dst, strlen(dst), u, 48);
    return fr_from_be_bytes(u, 48);
}

static bool fp_from_be32(const uint8_t *p, Fp *out) {   // canonical big-endian -> Montgomery; false when >= q
    Fp v;
    for (int i = 0; i < 8; i++) { const uint8_t *b = p + 4 * (7 - i); v.l[i] = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }
    for (int i = 7; i >= 0; i--) {
        if (v.l[i] < FpParams::M(i)) break;
        if (v.l[i] > FpParams::M(i) || i == 0) return false;
    }
    *out = Fp::to_mont(v);
    return true;
}
static Fp2 twist_b() { return Fp2::mul(Fp2{Fp::from_u64(3), Fp::zero()}, Fp2::inv(Fp2{Fp::from_u64(9), Fp::one()})); }   // 3/(9+u)
// G1Affine.SetBytes for the uncompressed encoding (flags 00 = point, 01 = infinity), on-curve check included
static bool g1_from_raw(const uint8_t *p, G1Affine *out) {
    if ((p[0] >> 6) == 1) { for (int i = 1; i < 64; i++) if (p[i]) return false; if (p[0] != 0x40) return false; *out = G1Affine::inf(); return true; }
    if ((p[0] >> 6) != 0) return false;
    if (!fp_from_be32(p, &out->x) || !fp_from_be32(p + 32, &out->y)) return false;
    if (out->is_inf()) return true;
    return Fp::sqr(out->y) == Fp::add(Fp::mul(Fp::sqr(out->x), out->x), Fp::from_u64(3));
}
static bool g2_from_raw(const uint8_t *p, G2Affine *out) {
    if ((p[0] >> 6) == 1) { for (int i = 1; i < 128; i++) if (p[i]) return false; if (p[0] != 0x40) return false; *out = G2Affine::inf(); return true; }
    if ((p[0] >> 6) != 0) return false;
    if (!fp_from_be32(p, &out->x.a1) || !fp_from_be32(p + 32, &out->x.a0) || !fp_from_be32(p + 64, &out->y.a1) || !fp_from_be32(p + 96, &out->y.a0)) return false;
    if (out->is_inf()) return true;
    return Fp2::sqr(out->y) == Fp2::add(Fp2::mul(Fp2::sqr(out->x), out->x), twist_b());
}

struct ParsedProof { G1Affine ar, krs, commitment, pok; G2Affine bs; uint32_t n_commit; };
static int32_t parse_proof(const uint8_t *b, uint32_t len, uint64_t want_commitments, ParsedProof *pp) {
    ZK_REQUIRE(len >= 260, "verify: proof shorter than Ar|Bs|Krs|count");
    ZK_REQUIRE(g1_from_raw(b, &pp->ar) && g2_from_raw(b + 64, &pp->bs) && g1_from_raw(b + 192, &pp->krs), "verify: Ar/Bs/Krs is not a valid uncompressed point");
    pp->n_commit = ((uint32_t)b[256] << 24) | ((uint32_t)b[257] << 16) | ((uint32_t)b[258] << 8) | b[259];
    ZK_REQUIRE(pp->n_commit == want_commitments && pp->n_commit <= 1, "verify: commitment count differs from the verifying key's");
    ZK_REQUIRE(len == 260 + 64 * pp->n_commit + 64, "verify: proof length does not match its commitment count (raw encoding expected)");
    pp->commitment = G1Affine::inf();
    if (pp->n_commit) ZK_REQUIRE(g1_from_raw(b + 260, &pp->commitment), "verify: commitment is not a valid point");
    ZK_REQUIRE(g1_from_raw(b + 260 + 64 * pp->n_commit, &pp->pok), "verify: commitment proof of knowledge is not a valid point");
    return ZKPOR_OK;
}

static G1Affine g1_neg(const G1Affine &p) { return G1Affine{p.x, Fp::neg(p.y)}; }
static G2Affine g2_neg(const G2Affine &p) { return G2Affine{p.x, Fp2::neg(p.y)}; }
static G1XYZZ g1_mul_fr(const G1Affine &p, const Fr &k_mont) { Fr k = Fr::from_mont(k_mont); return G1XYZZ::from_affine(p).mul_256(k.l); }

// Miller loops on the GPU, product on the host; *out = prod (not yet exponentiated)
static int32_t miller_product(zkpor_ctx *ctx, const G1Affine *P, const G2Affine *Q, uint64_t n, bool host_inputs, Fp12 *out) {
    *out = Fp12::one();
    if (n == 0) return ZKPOR_OK;
    const void *dp = P, *dq = Q;
    if (host_inputs) {
        ZK_TRY(ctx->in_points.reserve(n * (sizeof(G1Affine) + sizeof(G2Affine))));
        uint8_t *base = ctx->in_points.as<uint8_t>();
        ZK_CUDA(cudaMemcpyAsync(base, P, n * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
        ZK_CUDA(cudaMemcpyAsync(base + n * sizeof(G1Affine), Q, n * sizeof(G2Affine), cudaMemcpyHostToDevice, ctx->stream));
        dp = base; dq = base + n * sizeof(G1Affine);
    }
    ZK_TRY(ctx->misc.reserve(n * sizeof(Fp12)));
    ZK_LAUNCH(ctx, k_miller_loops, grid_for(n, 64), 64, 0, (const G1Affine *)dp, (const G2Affine *)dq, n, ctx->misc.as<Fp12>());
    std::vector<Fp12> f(n);
    ZK_CUDA(cudaMemcpyAsync(f.data(), ctx->misc.p, n * sizeof(Fp12), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    Fp12 acc = f[0];
    for (uint64_t i = 1; i < n; i++) acc = Fp12::mul(acc, f[i]);
    *out = acc;
    return ZKPOR_OK;
}

// commitment challenge of Verify / Prove's hint: hash_to_field("bsb22-commitment")(Marshal(commitment) || public committed)
static Fr commitment_challenge(const G1Affine &commitment, const zkpor_vk_desc *vk, const Fr *pw) {
    std::vector<uint8_t> msg(64 + 32 * vk->n_public_committed);
    g1_to_raw_bytes(msg.data(), commitment);
    for (uint64_t j = 0; j < vk->n_public_committed; j++) fr_to_be32(msg.data() + 64 + 32 * j, pw[vk->public_committed[j] - 1]);
    return hash_to_fr(msg.data(), msg.size(), "bsb22-commitment");
}

Fr commitment_challenge_g1(const G1Affine &commitment) {
    uint8_t msg[64];
    g1_to_raw_bytes(msg, commitment);
    return hash_to_fr(msg, 64, "bsb22-commitment");
}

static int32_t check_vk(const zkpor_vk_desc *vk, uint64_t n_public) {
    ZK_REQUIRE(vk && vk->g1_alpha && vk->g2_beta && vk->g2_gamma && vk->g2_delta && vk->g1_k, "verify: null verifying-key field");
    ZK_REQUIRE(vk->n_commitments <= 1, "verify: at most one BSB22 commitment is supported (the reference circuits have exactly one)");
    ZK_REQUIRE(vk->n_k == 1 + n_public + vk->n_commitments, "verify: len(vk.G1.K) != 1 + public inputs + commitments");
    if (vk->n_commitments) ZK_REQUIRE(vk->g2_ped_g && vk->g2_ped_g_root_sigma_neg, "verify: null Pedersen verifying key");
    for (uint64_t j = 0; j < vk->n_public_committed; j++) ZK_REQUIRE(vk->public_committed[j] >= 1 && vk->public_committed[j] <= n_public, "verify: public committed index out of range");
    return ZKPOR_OK;
}

}  // namespace zk

using namespace zk;

extern "C" {

int32_t zkpor_pairing_product(zkpor_ctx *ctx, const void *g1_points, const void *g2_points, uint64_t n, void *out_gt384) {
    ZK_REQUIRE(ctx != nullptr && out_gt384 != nullptr && (n == 0 || (g1_points && g2_points)), "pairing_product: null argument");
    ZK_CUDA(cudaSetDevice(ctx->device));
    const bool host = n && !is_device_ptr(g1_points);
    ZK_REQUIRE(n == 0 || host == !is_device_ptr(g2_points), "pairing_product: G1 and G2 arrays must both be host or both be device memory");
    Fp12 f;
    ZK_TRY(miller_product(ctx, (const G1Affine *)g1_points, (const G2Affine *)g2_points, n, host, &f));
    Fp12 e = pairing::final_exponentiation(f);
    memcpy(out_gt384, &e, sizeof e);
    return ZKPOR_OK;
}

int32_t zkpor_pairing_check(zkpor_ctx *ctx, const void *g1_points, const void *g2_points, uint64_t n, int32_t *out_ok) {
    ZK_REQUIRE(out_ok != nullptr, "pairing_check: null output");
    Fp12 e;
    ZK_TRY(zkpor_pairing_product(ctx, g1_points, g2_points, n, &e));
    *out_ok = e == Fp12::one();
    return ZKPOR_OK;
}

int32_t zkpor_groth16_verify(zkpor_ctx *ctx, const zkpor_vk_desc *vk, const uint8_t *proof_raw, uint32_t proof_len,
                             const void *public_witness, uint64_t n_public, int32_t *out_ok) {
    ZK_REQUIRE(ctx != nullptr && proof_raw != nullptr && out_ok != nullptr && (n_public == 0 || public_witness != nullptr), "verify: null argument");
    ZK_TRY(check_vk(vk, n_public));
    ZK_CUDA(cudaSetDevice(ctx->device));
    *out_ok = 0;
    ParsedProof pp;
    ZK_TRY(parse_proof(proof_raw, proof_len, vk->n_commitments, &pp));
    const Fr *pw = (const Fr *)public_witness;
    const G1Affine *K = (const G1Affine *)vk->g1_k;
    // sum_i pw_i K[1+i] + K[0] (+ challenge K[last] + commitment)
    G1XYZZ ksum = G1XYZZ::from_affine(K[0]);
    for (uint64_t i = 0; i < n_public; i++) ksum.add(g1_mul_fr(K[1 + i], pw[i]));
    if (pp.n_commit) {
        ksum.add(g1_mul_fr(K[1 + n_public], commitment_challenge(pp.commitment, vk, pw)));
        ksum.add_affine(pp.commitment, false);
    }
    // one launch: Bs subgroup check + all Miller loops.  pairs 0-3: Groth16 equation, pairs 4-5: Pedersen proof of knowledge
    G1Affine P[6] = {pp.krs, pp.ar, ksum.to_affine(), g1_neg(*(const G1Affine *)vk->g1_alpha), pp.commitment, pp.pok};
    G2Affine Q[6] = {g2_neg(*(const G2Affine *)vk->g2_delta), pp.bs, g2_neg(*(const G2Affine *)vk->g2_gamma), *(const G2Affine *)vk->g2_beta, G2Affine::inf(), G2Affine::inf()};
    if (pp.n_commit) { Q[4] = *(const G2Affine *)vk->g2_ped_g; Q[5] = *(const G2Affine *)vk->g2_ped_g_root_sigma_neg; }   // e(C, G) e(pok, G^(-1/sigma)) = 1, pok = sigma C
    ZK_TRY(ctx->io.reserve(sizeof(G2Affine) + 64));
    ZK_CUDA(cudaMemcpyAsync(ctx->io.p, &pp.bs, sizeof(G2Affine), cudaMemcpyHostToDevice, ctx->stream));
    uint8_t *d_ok = ctx->io.as<uint8_t>() + sizeof(G2Affine);
    ZK_LAUNCH(ctx, k_g2_in_subgroup, 1, 64, 0, ctx->io.as<G2Affine>(), (uint64_t)1, twist_b(), d_ok);
    uint8_t in_sub = 0;
    ZK_CUDA(cudaMemcpyAsync(&in_sub, d_ok, 1, cudaMemcpyDeviceToHost, ctx->stream));
    Fp12 groth, ped;
    ZK_TRY(miller_product(ctx, P, Q, 4, true, &groth));
    ZK_TRY(miller_product(ctx, P + 4, Q + 4, pp.n_commit ? 2 : 0, true, &ped));
    if (!in_sub) return ZKPOR_OK;   // proof.isValid() fails: Bs outside G2
    const bool ok_groth = pairing::final_exponentiation(groth) == Fp12::one();
    const bool ok_ped = !pp.n_commit || pairing::final_exponentiation(ped) == Fp12::one();
    *out_ok = ok_groth && ok_ped;
    return ZKPOR_OK;
}

int32_t zkpor_groth16_verify_batch(zkpor_ctx *ctx, const zkpor_vk_desc *vk, const uint8_t *proofs_raw, uint32_t proof_len,
                                   uint64_t proof_stride, const void *public_witnesses, uint64_t n_public, uint64_t count,
                                   const uint8_t seed32[32], int32_t *out_ok) {
    ZK_REQUIRE(ctx != nullptr && proofs_raw != nullptr && out_ok != nullptr && seed32 != nullptr && (n_public == 0 || public_witnesses != nullptr), "verify_batch: null argument");
    ZK_REQUIRE(count > 0 && proof_stride >= proof_len, "verify_batch: empty batch or stride shorter than a proof");
    ZK_TRY(check_vk(vk, n_public));
    ZK_CUDA(cudaSetDevice(ctx->device));
    *out_ok = 0;
    const G1Affine *K = (const G1Affine *)vk->g1_k;
    const bool has_c = vk->n_commitments == 1;
    // host pass: parse, challenges, 128-bit coefficients rho_i (Groth16) and rho'_i (Pedersen) bound to the whole batch
    uint8_t batch_digest[32];
    { Sha256 s; s.update(seed32, 32); for (uint64_t i = 0; i < count; i++) s.update(proofs_raw + i * proof_stride, proof_len); s.update(public_witnesses, count * n_public * 32); s.final(batch_digest); }
    std::vector<ParsedProof> pr(count);
    std::vector<G1Affine> ar(count), krs(count), cm(count), pok(count);
    std::vector<G2Affine> bs(count);
    std::vector<uint32_t> rho(4 * count), rho2(4 * count);
    std::vector<Fr> kscal(vk->n_k, Fr::zero());     // coefficient of every K[j] in sum_i rho_i * ksum_i
    for (uint64_t i = 0; i < count; i++) {
        ZK_TRY(parse_proof(proofs_raw + i * proof_stride, proof_len, vk->n_commitments, &pr[i]));
        ar[i] = pr[i].ar; krs[i] = pr[i].krs; cm[i] = pr[i].commitment; pok[i] = pr[i].pok; bs[i] = pr[i].bs;
        uint8_t d[32], ib[8];
        for (int k = 0; k < 8; k++) ib[k] = (uint8_t)(i >> (56 - 8 * k));
        { Sha256 s; s.update(batch_digest, 32); s.update(ib, 8); s.final(d); }
        for (int k = 0; k < 4; k++) { memcpy(&rho[4 * i + k], d + 4 * k, 4); memcpy(&rho2[4 * i + k], d + 16 + 4 * k, 4); }
        Fr r = Fr::zero();
        for (int k = 0; k < 4; k++) r.l[k] = rho[4 * i + k];
        r = Fr::to_mont(r);
        const Fr *pw = (const Fr *)public_witnesses + i * n_public;
        kscal[0] = Fr::add(kscal[0], r);
        for (uint64_t j = 0; j < n_public; j++) kscal[1 + j] = Fr::add(kscal[1 + j], Fr::mul(r, pw[j]));
        if (has_c) kscal[1 + n_public] = Fr::add(kscal[1 + n_public], Fr::mul(r, commitment_challenge(pr[i].commitment, vk, pw)));
    }
    // device: subgroup checks of every Bs, rho_i * Ar_i, and the folded sums (MSMs over the batch)
    ZK_TRY(ctx->io.reserve(count * (sizeof(G2Affine) + 2 * sizeof(G1Affine) + 16 + 1) + 256));
    uint8_t *base = ctx->io.as<uint8_t>();
    G2Affine *d_bs = (G2Affine *)base;
    G1Affine *d_ar = (G1Affine *)(base + count * sizeof(G2Affine)), *d_sar = d_ar + count;
    uint32_t *d_rho = (uint32_t *)(d_sar + count);
    uint8_t *d_ok = (uint8_t *)(d_rho + 4 * count);
    ZK_CUDA(cudaMemcpyAsync(d_bs, bs.data(), count * sizeof(G2Affine), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(d_ar, ar.data(), count * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(d_rho, rho.data(), count * 16, cudaMemcpyHostToDevice, ctx->stream));
    ZK_LAUNCH(ctx, k_g2_in_subgroup, grid_for(count, 64), 64, 0, d_bs, count, twist_b(), d_ok);
    ZK_LAUNCH(ctx, k_scale_g1_128, grid_for(count, 128), 128, 0, d_ar, d_rho, count, d_sar);
    std::vector<uint8_t> okv(count);
    std::vector<G1Affine> sar(count);
    ZK_CUDA(cudaMemcpyAsync(okv.data(), d_ok, count, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(sar.data(), d_sar, count * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint64_t i = 0; i < count; i++) if (!okv[i]) return ZKPOR_OK;
    // 128-bit coefficients as 256-bit plain scalars for the MSMs
    std::vector<uint32_t> s256(8 * count, 0), s256b(8 * count, 0);
    for (uint64_t i = 0; i < count; i++) for (int k = 0; k < 4; k++) { s256[8 * i + k] = rho[4 * i + k]; s256b[8 * i + k] = rho2[4 * i + k]; }
    G1XYZZ s_krs, s_cm, s_cm2, s_pok2;
    ZK_TRY(zkpor_msm_g1_partial(ctx, krs.data(), s256.data(), count, ZKPOR_SCALARS_PLAIN, &s_krs));
    G1XYZZ ksum = G1XYZZ::inf();
    for (uint64_t j = 0; j < vk->n_k; j++) ksum.add(g1_mul_fr(K[j], kscal[j]));
    Fr rho_sum = kscal[0];
    std::vector<G1Affine> P(sar);
    std::vector<G2Affine> Q(bs);
    if (has_c) {
        ZK_TRY(zkpor_msm_g1_partial(ctx, cm.data(), s256.data(), count, ZKPOR_SCALARS_PLAIN, &s_cm));
        ZK_TRY(zkpor_msm_g1_partial(ctx, cm.data(), s256b.data(), count, ZKPOR_SCALARS_PLAIN, &s_cm2));
        ZK_TRY(zkpor_msm_g1_partial(ctx, pok.data(), s256b.data(), count, ZKPOR_SCALARS_PLAIN, &s_pok2));
        ksum.add(s_cm);
        P.push_back(s_cm2.to_affine()); Q.push_back(*(const G2Affine *)vk->g2_ped_g);
        P.push_back(s_pok2.to_affine()); Q.push_back(*(const G2Affine *)vk->g2_ped_g_root_sigma_neg);
    }
    P.push_back(s_krs.to_affine()); Q.push_back(g2_neg(*(const G2Affine *)vk->g2_delta));
    P.push_back(ksum.to_affine()); Q.push_back(g2_neg(*(const G2Affine *)vk->g2_gamma));
    P.push_back(g1_neg(g1_mul_fr(*(const G1Affine *)vk->g1_alpha, rho_sum).to_affine())); Q.push_back(*(const G2Affine *)vk->g2_beta);
    Fp12 f;
    ZK_TRY(miller_product(ctx, P.data(), Q.data(), P.size(), true, &f));
    *out_ok = pairing::final_exponentiation(f) == Fp12::one();
    return ZKPOR_OK;
}

}  // extern "C"
