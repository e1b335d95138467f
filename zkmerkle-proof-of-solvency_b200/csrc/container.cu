// Byte containers of the Groth16 objects (formats per App. B.3):
//   proof.ReadFrom / WriteTo / WriteRawTo     src/verifier/main.go:208-216, src/prover/prover/prover.go:201
//   vk.ReadFrom / WriteTo                     src/prover/prover/prover.go:358-362, src/verifier/main.go:33-34, src/keygen/main.go:46-62
//   pk.ReadFrom / UnsafeReadFrom / WriteTo    src/prover/prover/prover.go:342-346, src/keygen/main.go:46-62
// gnark's marshal.go and gnark-crypto's Encoder/Decoder are out of tree (go.mod:57-60); what is restated here: everything big-endian;
// a point is written compressed (G1 32 B, G2 64 B, X.A1 first) by WriteTo and raw (64 / 128 B) by WriteRawTo, the top two bits of its
// first byte say which (00 raw, 10 / 11 compressed with the smaller / larger y, 01 infinity) and the decoder follows them point by
// point; a slice is a u32 length followed by its elements; []bool is a u32 length followed by ceil(len/8) bytes, bit i%8 of byte i/8.
// The vk layout reproduces the reference's 524-byte files (README.md:54,57).  The r1cs container (CBOR + intcomp) stays with gnark.
//
// The per-point work of a 2^26 key -- one square root in Fp or Fp2 per compressed point on the way in, one parity test on the way out --
// runs on the GPU (codec.cu's decode kernels, the encode kernels below) on the byte ranges of each array, straight from / into the
// caller's file buffer.
#include "internal.h"
#include <algorithm>

using namespace ff;
using namespace ec;

namespace zk {

__device__ __forceinline__ void store_be_fp(uint8_t *p, const Fp &plain, uint32_t flag_bits) {
    uint32_t *w = reinterpret_cast<uint32_t *>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t v = plain.l[7 - i];
        if (i == 0) v |= flag_bits << 30;
        w[i] = __byte_perm(v, 0, 0x0123);
    }
}
// canonical v > (q-1)/2
__device__ __forceinline__ bool plain_lex_largest(const Fp &v) {
    const uint32_t H[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
    for (int i = 7; i >= 0; i--) { if (v.l[i] > H[i]) return true; if (v.l[i] < H[i]) return false; }
    return false;
}

__global__ void __launch_bounds__(128) k_g1_encode(const G1Affine *__restrict__ in, uint64_t n, int compressed, uint8_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const G1Affine p = in[i];
    uint8_t *o = out + i * (compressed ? 32 : 64);
    if (p.is_inf()) {
        uint4 z = make_uint4(0, 0, 0, 0), f = make_uint4(0x40u, 0, 0, 0);    // byte 0 = 0b01 << 6
        uint4 *w = reinterpret_cast<uint4 *>(o);
        w[0] = f; w[1] = z;
        if (!compressed) { w[2] = z; w[3] = z; }
        return;
    }
    const Fp x = Fp::from_mont(p.x), y = Fp::from_mont(p.y);
    if (compressed) store_be_fp(o, x, plain_lex_largest(y) ? 3u : 2u);
    else { store_be_fp(o, x, 0u); store_be_fp(o + 32, y, 0u); }
}
__global__ void __launch_bounds__(128) k_g2_encode(const G2Affine *__restrict__ in, uint64_t n, int compressed, uint8_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const G2Affine p = in[i];
    uint8_t *o = out + i * (compressed ? 64 : 128);
    if (p.is_inf()) {
        uint4 z = make_uint4(0, 0, 0, 0), f = make_uint4(0x40u, 0, 0, 0);
        uint4 *w = reinterpret_cast<uint4 *>(o);
        w[0] = f;
        for (int k = 1; k < (compressed ? 4 : 8); k++) w[k] = z;
        return;
    }
    const Fp x0 = Fp::from_mont(p.x.a0), x1 = Fp::from_mont(p.x.a1), y0 = Fp::from_mont(p.y.a0), y1 = Fp::from_mont(p.y.a1);
    if (compressed) {
        const bool largest = y1.is_zero() ? plain_lex_largest(y0) : plain_lex_largest(y1);     // E2.LexicographicallyLargest: A1 first
        store_be_fp(o, x1, largest ? 3u : 2u); store_be_fp(o + 32, x0, 0u);
    } else {
        store_be_fp(o, x1, 0u); store_be_fp(o + 32, x0, 0u); store_be_fp(o + 64, y1, 0u); store_be_fp(o + 96, y0, 0u);
    }
}

template <bool G2>
static int32_t encode_batch(zkpor_ctx *ctx, const void *points, uint64_t n, int32_t compressed, void *out_bytes) {
    ZK_REQUIRE(ctx && ((points && out_bytes) || n == 0), "encode: null argument");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return ZKPOR_OK;
    const size_t in_sz = G2 ? 128 : 64, out_sz = (G2 ? 64 : 32) * (compressed ? 1 : 2);
    const void *din;
    ZK_TRY(to_device(ctx, points, n * in_sz, ctx->in_points, &din));
    const bool out_dev = is_device_ptr(out_bytes);
    void *dout = out_bytes;
    if (!out_dev) { ZK_TRY(ctx->io.reserve(n * out_sz)); dout = ctx->io.p; }
    if (G2) ZK_LAUNCH(ctx, k_g2_encode, grid_for(n, 128), 128, 0, (const G2Affine *)din, n, (int)compressed, (uint8_t *)dout);
    else ZK_LAUNCH(ctx, k_g1_encode, grid_for(n, 128), 128, 0, (const G1Affine *)din, n, (int)compressed, (uint8_t *)dout);
    if (!out_dev) ZK_CUDA(cudaMemcpyAsync(out_bytes, dout, n * out_sz, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

// ------------------------------------------------------------------------------------------------ host byte streams
struct Reader {
    const uint8_t *p, *end; bool ok = true;
    Reader(const uint8_t *b, uint64_t n) : p(b), end(b + n) {}
    const uint8_t *take(uint64_t n) { if (!ok || (uint64_t)(end - p) < n) { ok = false; return nullptr; } const uint8_t *q = p; p += n; return q; }
    uint32_t u32() { const uint8_t *q = take(4); return q ? ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3] : 0; }
    uint64_t u64() { const uint64_t hi = u32(); return (hi << 32) | u32(); }
};
struct Writer {
    uint8_t *p; uint64_t cap, len = 0;     // p == nullptr: count only
    Writer(uint8_t *b, uint64_t c) : p(b), cap(c) {}
    uint8_t *room(uint64_t n) { uint8_t *q = (p && len + n <= cap) ? p + len : nullptr; len += n; return q; }
    void bytes(const void *src, uint64_t n) { if (uint8_t *q = room(n)) memcpy(q, src, n); }
    void u32(uint32_t v) { const uint8_t b[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v}; bytes(b, 4); }
    void u64(uint64_t v) { u32((uint32_t)(v >> 32)); u32((uint32_t)v); }
    bool fits() const { return p == nullptr || len <= cap; }
};

// size of the point encoding that starts with byte b0 (the decoder reads the compressed size first and more when the flag says raw);
// 01 (infinity) has the size of its neighbours: `raw_hint`
static inline bool flag_is_raw(uint8_t b0, bool raw_hint) { const uint32_t f = b0 >> 6; return f == 0 ? true : f == 1 ? raw_hint : false; }

// `count` points of one group from the stream, each in the encoding its flag bits name, into `out` (host, affine Montgomery)
template <bool G2>
static int32_t read_points_host(zkpor_ctx *ctx, Reader &r, uint64_t count, bool raw_hint, void *out) {
    const uint64_t csz = G2 ? 64 : 32, psz = G2 ? 128 : 64;
    for (uint64_t i = 0; i < count; i++) {
        if (r.p >= r.end) { r.ok = false; break; }
        const bool raw = flag_is_raw(r.p[0], raw_hint);
        raw_hint = raw;
        const uint8_t *q = r.take(raw ? 2 * csz : csz);
        if (!q) break;
        ZK_TRY((G2 ? zkpor_g2_decode_batch : zkpor_g1_decode_batch)(ctx, q, 1, raw ? 0 : 1, (uint8_t *)out + i * psz));
    }
    if (!r.ok) { set_error("container: truncated input"); return ZKPOR_ERR_INVALID_ARG; }
    return ZKPOR_OK;
}
// a slice whose elements share one encoding (a key array): decoded in one launch from the file's byte range into DEVICE memory
template <bool G2>
static int32_t read_point_slice_dev(zkpor_ctx *ctx, Reader &r, uint64_t *count, void **d_out) {
    const uint64_t csz = G2 ? 64 : 32, psz = G2 ? 128 : 64;
    *d_out = nullptr;
    const uint64_t n = r.u32();
    *count = n;
    if (!r.ok) { set_error("container: truncated input"); return ZKPOR_ERR_INVALID_ARG; }
    if (n == 0) return ZKPOR_OK;
    // the encoding of the slice: the first element that is not the point at infinity tells (compressed elements are csz apart)
    bool raw = false, known = false;
    for (uint64_t i = 0; i < n && !known; i++) {
        const uint8_t *q = r.p + i * csz;                      // valid while every element so far was an infinity of the compressed size
        if (q >= r.end) break;
        const uint32_t f = q[0] >> 6;
        if (f == 1) { bool zeros = true; for (uint64_t k = 1; k < csz && q + k < r.end; k++) zeros &= q[k] == 0; if (zeros) continue; }
        raw = f == 0; known = true;
    }
    const uint8_t *q = r.take(n * (raw ? 2 * csz : csz));
    if (!q) { set_error("container: truncated point slice"); return ZKPOR_ERR_INVALID_ARG; }
    ZK_CUDA(cudaMalloc(d_out, n * psz));
    const int32_t rc = (G2 ? zkpor_g2_decode_batch : zkpor_g1_decode_batch)(ctx, q, n, raw ? 0 : 1, *d_out);
    if (rc != ZKPOR_OK) { cudaFree(*d_out); *d_out = nullptr; }
    return rc;
}
template <bool G2>
static int32_t write_points(zkpor_ctx *ctx, Writer &w, const void *pts /* host or device */, uint64_t n, bool raw) {
    const uint64_t sz = (G2 ? 64 : 32) * (raw ? 2 : 1);
    uint8_t *q = w.room(n * sz);
    if (!q || n == 0) return ZKPOR_OK;
    return encode_batch<G2>(ctx, pts, n, raw ? 0 : 1, q);
}
static void write_fr(Writer &w, const Fr &mont) {
    const Fr p = Fr::from_mont(mont);
    uint8_t b[32];
    for (int i = 0; i < 8; i++) { const uint32_t v = p.l[7 - i]; b[4 * i] = v >> 24; b[4 * i + 1] = v >> 16; b[4 * i + 2] = v >> 8; b[4 * i + 3] = v; }
    w.bytes(b, 32);
}

static const uint32_t ROOT_2_28_PLAIN[8] = {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu, 0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u};

}  // namespace zk

using namespace zk;

extern "C" {

int32_t zkpor_g1_encode_batch(zkpor_ctx *ctx, const void *points, uint64_t n, int32_t compressed, void *out_bytes) {
    return encode_batch<false>(ctx, points, n, compressed, out_bytes);
}
int32_t zkpor_g2_encode_batch(zkpor_ctx *ctx, const void *points, uint64_t n, int32_t compressed, void *out_bytes) {
    return encode_batch<true>(ctx, points, n, compressed, out_bytes);
}

// ---- proof ------------------------------------------------------------------------------------------------------------------
int32_t zkpor_proof_decode(zkpor_ctx *ctx, const uint8_t *in, uint64_t in_len, uint8_t *out_raw, uint32_t *out_len, uint64_t *consumed) {
    ZK_REQUIRE(ctx && in && out_raw && out_len, "proof_decode: null argument");
    Reader r(in, in_len);
    G1Affine ar, krs, pok; G2Affine bs;
    std::vector<G1Affine> cm;
    ZK_TRY(read_points_host<false>(ctx, r, 1, false, &ar));
    const bool raw = (in[0] >> 6) == 0;
    ZK_TRY(read_points_host<true>(ctx, r, 1, raw, &bs));
    ZK_TRY(read_points_host<false>(ctx, r, 1, raw, &krs));
    const uint32_t nc = r.u32();
    if (!r.ok || nc > 16) { set_error("proof_decode: truncated input or implausible commitment count"); return ZKPOR_ERR_INVALID_ARG; }
    cm.resize(nc);
    if (nc) ZK_TRY(read_points_host<false>(ctx, r, nc, raw, cm.data()));
    ZK_TRY(read_points_host<false>(ctx, r, 1, raw, &pok));
    const uint32_t need = 64 + 128 + 64 + 4 + 64 * nc + 64;
    ZK_REQUIRE(*out_len >= need, "proof_decode: output buffer too small (260 + 64 per commitment + 64 bytes)");
    g1_to_raw_bytes(out_raw, ar); g2_to_raw_bytes(out_raw + 64, bs); g1_to_raw_bytes(out_raw + 192, krs);
    out_raw[256] = 0; out_raw[257] = 0; out_raw[258] = (uint8_t)(nc >> 8); out_raw[259] = (uint8_t)nc;
    for (uint32_t i = 0; i < nc; i++) g1_to_raw_bytes(out_raw + 260 + 64 * i, cm[i]);
    g1_to_raw_bytes(out_raw + 260 + 64 * nc, pok);
    *out_len = need;
    if (consumed) *consumed = (uint64_t)(r.p - in);
    return ZKPOR_OK;
}

int32_t zkpor_proof_encode(zkpor_ctx *ctx, const uint8_t *raw, uint32_t raw_len, int32_t compressed, uint8_t *out, uint32_t *out_len) {
    ZK_REQUIRE(ctx && raw && out && out_len, "proof_encode: null argument");
    // through the decoder: validates the points and accepts either input encoding
    uint8_t canon[260 + 64 * 17]; uint32_t clen = sizeof canon; uint64_t used = 0;
    ZK_TRY(zkpor_proof_decode(ctx, raw, raw_len, canon, &clen, &used));
    if (!compressed) { ZK_REQUIRE(*out_len >= clen, "proof_encode: output buffer too small"); memcpy(out, canon, clen); *out_len = clen; return ZKPOR_OK; }
    const uint32_t nc = ((uint32_t)canon[258] << 8) | canon[259];
    const uint32_t need = 32 + 64 + 32 + 4 + 32 * nc + 32;
    ZK_REQUIRE(*out_len >= need, "proof_encode: output buffer too small");
    G1Affine g1[19]; G2Affine g2;
    ZK_TRY(zkpor_g1_decode_batch(ctx, canon, 1, 0, &g1[0]));
    ZK_TRY(zkpor_g2_decode_batch(ctx, canon + 64, 1, 0, &g2));
    ZK_TRY(zkpor_g1_decode_batch(ctx, canon + 192, 1, 0, &g1[1]));
    ZK_TRY(zkpor_g1_decode_batch(ctx, canon + 260, nc + 1, 0, &g1[2]));
    Writer w(out, *out_len);
    ZK_TRY(write_points<false>(ctx, w, &g1[0], 1, false));
    ZK_TRY(write_points<true>(ctx, w, &g2, 1, false));
    ZK_TRY(write_points<false>(ctx, w, &g1[1], 1, false));
    w.u32(nc);
    ZK_TRY(write_points<false>(ctx, w, &g1[2], nc + 1, false));
    *out_len = (uint32_t)w.len;
    return ZKPOR_OK;
}

// ---- verifying key ----------------------------------------------------------------------------------------------------------
int32_t zkpor_vk_decode(zkpor_ctx *ctx, const uint8_t *in, uint64_t in_len, zkpor_vk_host *out, void *k_points, uint64_t k_cap,
                        uint64_t *public_committed, uint64_t pc_cap, uint64_t *consumed) {
    ZK_REQUIRE(ctx && in && out, "vk_decode: null argument");
    memset(out, 0, sizeof *out);
    Reader r(in, in_len);
    const bool raw = in_len > 0 && (in[0] >> 6) == 0;
    ZK_TRY(read_points_host<false>(ctx, r, 1, raw, out->g1_alpha));
    ZK_TRY(read_points_host<false>(ctx, r, 1, raw, out->g1_beta));
    ZK_TRY(read_points_host<true>(ctx, r, 1, raw, out->g2_beta));
    ZK_TRY(read_points_host<true>(ctx, r, 1, raw, out->g2_gamma));
    ZK_TRY(read_points_host<false>(ctx, r, 1, raw, out->g1_delta));
    ZK_TRY(read_points_host<true>(ctx, r, 1, raw, out->g2_delta));
    const uint64_t nk = r.u32();
    if (!r.ok) { set_error("vk_decode: truncated input"); return ZKPOR_ERR_INVALID_ARG; }
    out->n_k = nk;
    ZK_REQUIRE(nk == 0 || (k_points && k_cap >= nk), "vk_decode: k_points too small for vk.G1.K");
    ZK_TRY(read_points_host<false>(ctx, r, nk, raw, k_points));
    // PublicAndCommitmentCommitted: [][]uint64 -- u32 outer length, per entry u32 length + u64 values
    const uint32_t outer = r.u32();
    if (!r.ok || outer > 1) { set_error("vk_decode: truncated input, or more than one commitment (the reference circuits have exactly one)"); return ZKPOR_ERR_INVALID_ARG; }
    out->n_commitments = outer;
    if (outer == 1) {
        const uint32_t inner = r.u32();
        ZK_REQUIRE(inner == 0 || (public_committed && pc_cap >= inner), "vk_decode: public_committed too small");
        for (uint32_t i = 0; i < inner; i++) public_committed[i] = r.u64();
        out->n_public_committed = inner;
        ZK_TRY(read_points_host<true>(ctx, r, 1, raw, out->g2_ped_g));
        ZK_TRY(read_points_host<true>(ctx, r, 1, raw, out->g2_ped_g_root_sigma_neg));
    }
    if (!r.ok) { set_error("vk_decode: truncated input"); return ZKPOR_ERR_INVALID_ARG; }
    if (consumed) *consumed = (uint64_t)(r.p - in);
    return ZKPOR_OK;
}

int32_t zkpor_vk_encode(zkpor_ctx *ctx, const zkpor_vk_host *vk, const void *k_points, const uint64_t *public_committed, int32_t raw,
                        uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
    ZK_REQUIRE(ctx && vk && out_len && (vk->n_k == 0 || k_points), "vk_encode: null argument");
    ZK_REQUIRE(vk->n_commitments <= 1 && (vk->n_public_committed == 0 || public_committed), "vk_encode: at most one commitment");
    Writer w(out, out_cap);
    const bool rw = raw != 0;
    ZK_TRY(write_points<false>(ctx, w, vk->g1_alpha, 1, rw)); ZK_TRY(write_points<false>(ctx, w, vk->g1_beta, 1, rw));
    ZK_TRY(write_points<true>(ctx, w, vk->g2_beta, 1, rw)); ZK_TRY(write_points<true>(ctx, w, vk->g2_gamma, 1, rw));
    ZK_TRY(write_points<false>(ctx, w, vk->g1_delta, 1, rw)); ZK_TRY(write_points<true>(ctx, w, vk->g2_delta, 1, rw));
    w.u32((uint32_t)vk->n_k);
    ZK_TRY(write_points<false>(ctx, w, k_points, vk->n_k, rw));
    w.u32((uint32_t)vk->n_commitments);
    if (vk->n_commitments == 1) {
        w.u32((uint32_t)vk->n_public_committed);
        for (uint64_t i = 0; i < vk->n_public_committed; i++) w.u64(public_committed[i]);
        ZK_TRY(write_points<true>(ctx, w, vk->g2_ped_g, 1, rw)); ZK_TRY(write_points<true>(ctx, w, vk->g2_ped_g_root_sigma_neg, 1, rw));
    }
    *out_len = w.len;
    ZK_REQUIRE(w.fits(), "vk_encode: output buffer too small (out_len holds the size needed)");
    return ZKPOR_OK;
}

// ---- proving key ------------------------------------------------------------------------------------------------------------
int32_t zkpor_pk_read(zkpor_ctx *ctx, const uint8_t *in, uint64_t in_len, const zkpor_pk_cs_info *info, zkpor_pk **out, uint64_t *consumed) {
    ZK_REQUIRE(ctx && in && info && out, "pk_read: null argument");
    ZK_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    Reader r(in, in_len);
    // fft.Domain: Cardinality u64, then CardinalityInv, Generator, GeneratorInv, FrMultiplicativeGen, FrMultiplicativeGenInv (32 B each,
    // all recomputable from the cardinality) and the withPrecompute flag
    const uint64_t card = r.u64();
    r.take(5 * 32); r.take(1);
    if (!r.ok || card == 0 || (card & (card - 1)) != 0 || card > (1ull << 28)) { set_error("pk_read: not a proving key (domain cardinality %llu)", (unsigned long long)card); return ZKPOR_ERR_INVALID_ARG; }
    zkpor_pk_desc d; memset(&d, 0, sizeof d);
    d.log_n = (uint32_t)__builtin_ctzll(card);
    G1Affine alpha1, beta1, delta1; G2Affine beta2, delta2;
    void *dev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};      // A, B1, Z, K, B2, ck, ck_sigma
    std::vector<uint8_t> inf_a, inf_b;
    int32_t rc = ZKPOR_OK;
    auto step = [&](int32_t v) { if (rc == ZKPOR_OK) rc = v; };
    const bool raw = r.p < r.end && (r.p[0] >> 6) == 0;
    step(read_points_host<false>(ctx, r, 1, raw, &alpha1)); step(read_points_host<false>(ctx, r, 1, raw, &beta1)); step(read_points_host<false>(ctx, r, 1, raw, &delta1));
    uint64_t n_ck = 0, n_ck2 = 0;
    if (rc == ZKPOR_OK) step(read_point_slice_dev<false>(ctx, r, &d.n_a, &dev[0]));
    if (rc == ZKPOR_OK) step(read_point_slice_dev<false>(ctx, r, &d.n_b, &dev[1]));
    if (rc == ZKPOR_OK) step(read_point_slice_dev<false>(ctx, r, &d.n_z, &dev[2]));
    if (rc == ZKPOR_OK) step(read_point_slice_dev<false>(ctx, r, &d.n_k, &dev[3]));
    step(read_points_host<true>(ctx, r, 1, raw, &beta2)); step(read_points_host<true>(ctx, r, 1, raw, &delta2));
    uint64_t n_b2 = 0;
    if (rc == ZKPOR_OK) step(read_point_slice_dev<true>(ctx, r, &n_b2, &dev[4]));
    if (rc == ZKPOR_OK) {
        d.n_wires = r.u64();
        const uint64_t nb_inf_a = r.u64(), nb_inf_b = r.u64();
        auto bools = [&](std::vector<uint8_t> &dst) {
            const uint64_t n = r.u32();
            const uint8_t *q = r.take((n + 7) / 8);
            if (!q || n != d.n_wires) { r.ok = false; return; }
            dst.resize(n);
            for (uint64_t i = 0; i < n; i++) dst[i] = (q[i >> 3] >> (i & 7)) & 1u;
        };
        bools(inf_a); bools(inf_b);
        const uint32_t n_keys = r.u32();
        if (!r.ok || n_keys > 1 || n_b2 != d.n_b) { set_error("pk_read: truncated or inconsistent key (wires %llu, commitment keys %u)", (unsigned long long)d.n_wires, n_keys); rc = ZKPOR_ERR_INVALID_ARG; }
        if (rc == ZKPOR_OK) {
            uint64_t ca = 0, cb = 0;
            for (uint8_t v : inf_a) ca += v;
            for (uint8_t v : inf_b) cb += v;
            if (ca != nb_inf_a || cb != nb_inf_b) { set_error("pk_read: NbInfinityA/B disagree with the infinity maps"); rc = ZKPOR_ERR_INVALID_ARG; }
        }
        if (rc == ZKPOR_OK && n_keys == 1) {
            step(read_point_slice_dev<false>(ctx, r, &n_ck, &dev[5]));
            if (rc == ZKPOR_OK) step(read_point_slice_dev<false>(ctx, r, &n_ck2, &dev[6]));
            if (rc == ZKPOR_OK && (n_ck != n_ck2 || n_ck != info->n_committed)) { set_error("pk_read: commitment key of %llu points, the constraint system commits %llu wires", (unsigned long long)n_ck, (unsigned long long)info->n_committed); rc = ZKPOR_ERR_INVALID_ARG; }
        }
    }
    if (rc == ZKPOR_OK) {
        d.n_public = info->n_public;
        d.g1_a = dev[0]; d.g1_b = dev[1]; d.g1_z = dev[2]; d.g1_k = dev[3]; d.g2_b = dev[4];
        d.g1_alpha = &alpha1; d.g1_beta = &beta1; d.g1_delta = &delta1; d.g2_beta = &beta2; d.g2_delta = &delta2;
        d.infinity_a = inf_a.data(); d.infinity_b = inf_b.data();
        d.n_committed = n_ck; d.ck_basis = dev[5]; d.ck_basis_exp_sigma = dev[6];
        d.private_committed = info->private_committed; d.commitment_index = info->commitment_index;
        rc = zkpor_pk_upload(ctx, &d, out);
    }
    for (void *q : dev) if (q) cudaFree(q);
    if (rc == ZKPOR_OK && consumed) *consumed = (uint64_t)(r.p - in);
    return rc;
}

int32_t zkpor_pk_write(zkpor_ctx *ctx, zkpor_pk *pk, int32_t raw, uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
    ZK_REQUIRE(ctx && pk && out_len, "pk_write: null argument");
    ZK_REQUIRE(pk->shard_world == 1 && pk->n_wires_total > 0, "pk_write: needs a whole key uploaded with its infinity maps");
    ZK_CUDA(cudaSetDevice(ctx->device));
    Writer w(out, out_cap);
    const bool rw = raw != 0;
    // fft.Domain
    const uint64_t card = 1ull << pk->log_n;
    Fr gen; memcpy(gen.l, ROOT_2_28_PLAIN, 32); gen = Fr::to_mont(gen);
    for (uint32_t i = pk->log_n; i < 28; i++) gen = Fr::sqr(gen);
    const Fr five = Fr::from_u64(5);
    w.u64(card);
    write_fr(w, Fr::inv(Fr::from_u64(card))); write_fr(w, gen); write_fr(w, Fr::inv(gen)); write_fr(w, five); write_fr(w, Fr::inv(five));
    const uint8_t with_precompute = 1; w.bytes(&with_precompute, 1);
    ZK_TRY(write_points<false>(ctx, w, &pk->alpha1, 1, rw)); ZK_TRY(write_points<false>(ctx, w, &pk->beta1, 1, rw)); ZK_TRY(write_points<false>(ctx, w, &pk->delta1, 1, rw));
    w.u32((uint32_t)pk->n_a); ZK_TRY(write_points<false>(ctx, w, pk->A, pk->n_a, rw));
    w.u32((uint32_t)pk->n_b); ZK_TRY(write_points<false>(ctx, w, pk->B1, pk->n_b, rw));
    w.u32((uint32_t)pk->n_z); ZK_TRY(write_points<false>(ctx, w, pk->Z, pk->n_z, rw));
    w.u32((uint32_t)pk->n_k); ZK_TRY(write_points<false>(ctx, w, pk->K, pk->n_k, rw));
    ZK_TRY(write_points<true>(ctx, w, &pk->beta2, 1, rw)); ZK_TRY(write_points<true>(ctx, w, &pk->delta2, 1, rw));
    w.u32((uint32_t)pk->n_b); ZK_TRY(write_points<true>(ctx, w, pk->B2, pk->n_b, rw));
    // InfinityA / InfinityB back from the skip bits of the wire maps
    const uint64_t W = pk->n_wires_total, words = (W + 31) / 32;
    std::vector<uint2> ma(words), mb(words);
    if (words) { ZK_CUDA(cudaMemcpy(ma.data(), pk->map_a, words * sizeof(uint2), cudaMemcpyDeviceToHost)); ZK_CUDA(cudaMemcpy(mb.data(), pk->map_b, words * sizeof(uint2), cudaMemcpyDeviceToHost)); }
    w.u64(W); w.u64(W - pk->n_a); w.u64(W - pk->n_b);
    for (const std::vector<uint2> *m : {&ma, &mb}) {
        w.u32((uint32_t)W);
        uint8_t *q = w.room((W + 7) / 8);
        if (q) for (uint64_t i = 0; i < (W + 7) / 8; i++) {
            uint8_t b = (uint8_t)((*m)[i >> 2].x >> (8 * (i & 3)));
            if (i == (W + 7) / 8 - 1 && (W & 7)) b &= (uint8_t)((1u << (W & 7)) - 1u);     // the map marks the padding wires as skipped
            q[i] = b;
        }
    }
    w.u32(pk->has_commitment ? 1u : 0u);
    if (pk->has_commitment) {
        w.u32((uint32_t)pk->n_ck); ZK_TRY(write_points<false>(ctx, w, pk->ck, pk->n_ck, rw));
        w.u32((uint32_t)pk->n_ck); ZK_TRY(write_points<false>(ctx, w, pk->ck_sigma, pk->n_ck, rw));
    }
    *out_len = w.len;
    ZK_REQUIRE(w.fits(), "pk_write: output buffer too small (out_len holds the size needed)");
    return ZKPOR_OK;
}

}  // extern "C"
