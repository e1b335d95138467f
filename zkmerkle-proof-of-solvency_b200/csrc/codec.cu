// Batch decoding of gnark-crypto BN254 point encodings on the GPU.
//
// pk.UnsafeReadFrom (src/prover/prover/prover.go:342-346) decodes ~4 x 2^26 compressed G1 points and 2^26 compressed
// G2 points per tier: one square root in Fp (Fp2) per point.  The reference only prints how long that load takes
// (prover.go:301-365).  Encoding (gnark-crypto ecc/bn254/marshal.go, out of tree): big-endian
// coordinates, top two bits of byte 0 = 00 uncompressed, 10 / 11 compressed with the lexicographically smaller / larger
// y, 01 infinity; G2 writes X.A1 before X.A0.  q = 3 mod 4, so sqrt(a) = a^((q+1)/4) when it exists.
#include "internal.h"

using namespace ff;
using namespace ec;

namespace zk {

__device__ __forceinline__ Fp load_be_fp(const uint8_t *p, uint32_t mask_top) {   // plain (canonical) limbs
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
    Fp v;
#pragma unroll
    for (int i = 0; i < 8; i++) v.l[i] = __byte_perm(w[7 - i], 0, 0x0123);
    v.l[7] &= mask_top;
    return v;
}
__device__ __forceinline__ bool lt_q(const Fp &v) {
    for (int i = 7; i >= 0; i--) { if (v.l[i] < FpParams::M(i)) return true; if (v.l[i] > FpParams::M(i)) return false; }
    return false;
}
// canonical y > (q-1)/2
__device__ __forceinline__ bool lex_largest_fp(const Fp &mont) {
    Fp v = Fp::from_mont(mont);
    // (q-1)/2 limbs
    const uint32_t H[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
    for (int i = 7; i >= 0; i--) { if (v.l[i] > H[i]) return true; if (v.l[i] < H[i]) return false; }
    return false;
}
__device__ __forceinline__ Fp fp_sqrt_candidate(const Fp &a) {   // a^((q+1)/4)
    const uint32_t E[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
    return Fp::pow(a, E);
}
__device__ __forceinline__ Fp fp_three() { return Fp::from_u64(3); }

__global__ void __launch_bounds__(128) k_g1_decode(const uint8_t *__restrict__ in, uint64_t n, int compressed, G1Affine *__restrict__ out,
                                                   unsigned long long *__restrict__ first_bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = in + i * (compressed ? 32 : 64);
    const uint32_t flag = p[0] >> 6;
    G1Affine r = G1Affine::inf();
    if (flag != 1) {   // 01 = infinity
        Fp x = load_be_fp(p, 0x3fffffffu);
        bool ok = lt_q(x);
        x = Fp::to_mont(x);
        if (flag == 0) {
            Fp y = load_be_fp(p + 32, 0xffffffffu);
            ok = ok && !compressed && lt_q(y);
            r = G1Affine{x, Fp::to_mont(y)};
        } else {
            Fp rhs = Fp::add(Fp::mul(Fp::sqr(x), x), fp_three());
            Fp y = fp_sqrt_candidate(rhs);
            ok = ok && compressed && Fp::sqr(y) == rhs;
            if (lex_largest_fp(y) != (flag == 3)) y = Fp::neg(y);
            r = G1Affine{x, y};
        }
        if (!ok) { atomicMin(first_bad, (unsigned long long)i); r = G1Affine::inf(); }
    }
    out[i] = r;
}

// square root in Fp2 (complex method); ok = false when a is not a square
__device__ __forceinline__ Fp2 fp2_sqrt(const Fp2 &a, bool &ok) {
    ok = true;
    if (a.is_zero()) return a;
    const Fp half = Fp::inv(Fp::from_u64(2));
    if (a.a1.is_zero()) {
        Fp s = fp_sqrt_candidate(a.a0);
        if (Fp::sqr(s) == a.a0) return Fp2{s, Fp::zero()};
        Fp na = Fp::neg(a.a0);
        s = fp_sqrt_candidate(na);
        ok = Fp::sqr(s) == na;
        return Fp2{Fp::zero(), s};
    }
    Fp norm = Fp::add(Fp::sqr(a.a0), Fp::sqr(a.a1));
    Fp nrt = fp_sqrt_candidate(norm);
    if (Fp::sqr(nrt) != norm) { ok = false; return a; }
    Fp cand = Fp::mul(Fp::add(a.a0, nrt), half);
    Fp x0 = fp_sqrt_candidate(cand);
    if (Fp::sqr(x0) != cand || x0.is_zero()) {
        cand = Fp::mul(Fp::sub(a.a0, nrt), half);
        x0 = fp_sqrt_candidate(cand);
        if (Fp::sqr(x0) != cand || x0.is_zero()) { ok = false; return a; }
    }
    Fp x1 = Fp::mul(a.a1, Fp::inv(Fp::dbl(x0)));
    Fp2 r{x0, x1};
    ok = Fp2::sqr(r) == a;
    return r;
}
__global__ void __launch_bounds__(128) k_g2_decode(const uint8_t *__restrict__ in, uint64_t n, int compressed, G2Affine *__restrict__ out,
                                                   Fp2 b_twist, unsigned long long *__restrict__ first_bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = in + i * (compressed ? 64 : 128);
    const uint32_t flag = p[0] >> 6;
    G2Affine r = G2Affine::inf();
    if (flag != 1) {
        Fp x1 = load_be_fp(p, 0x3fffffffu), x0 = load_be_fp(p + 32, 0xffffffffu);   // X.A1 first
        bool ok = lt_q(x0) && lt_q(x1);
        Fp2 x{Fp::to_mont(x0), Fp::to_mont(x1)};
        if (flag == 0) {
            Fp y1 = load_be_fp(p + 64, 0xffffffffu), y0 = load_be_fp(p + 96, 0xffffffffu);
            ok = ok && !compressed && lt_q(y0) && lt_q(y1);
            r = G2Affine{x, Fp2{Fp::to_mont(y0), Fp::to_mont(y1)}};
        } else {
            Fp2 rhs = Fp2::add(Fp2::mul(Fp2::sqr(x), x), b_twist);
            bool sq;
            Fp2 y = fp2_sqrt(rhs, sq);
            ok = ok && compressed && sq;
            bool largest = y.a1.is_zero() ? lex_largest_fp(y.a0) : lex_largest_fp(y.a1);   // E2.LexicographicallyLargest
            if (largest != (flag == 3)) y = Fp2::neg(y);
            r = G2Affine{x, y};
        }
        if (!ok) { atomicMin(first_bad, (unsigned long long)i); r = G2Affine::inf(); }
    }
    out[i] = r;
}

template <bool G2>
static int32_t decode_batch(zkpor_ctx *ctx, const void *in_bytes, uint64_t n, int32_t compressed, void *out_points) {
    ZK_REQUIRE(ctx && ((in_bytes && out_points) || n == 0), "decode: null argument");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return ZKPOR_OK;
    stages_reset(ctx);
    const size_t in_sz = (G2 ? 64 : 32) * (compressed ? 1 : 2), out_sz = G2 ? 128 : 64;
    const void *din;
    stage_begin(ctx, ST_H2D);
    ZK_TRY(to_device(ctx, in_bytes, n * in_sz, ctx->io, &din));
    stage_end(ctx, ST_H2D);
    const bool out_dev = is_device_ptr(out_points);
    void *dout = out_points;
    if (!out_dev) { ZK_TRY(ctx->in_points.reserve(n * out_sz)); dout = ctx->in_points.p; }
    ZK_TRY(ctx->misc.reserve(256));
    unsigned long long *bad = ctx->misc.as<unsigned long long>();
    ZK_CUDA(cudaMemsetAsync(bad, 0xff, 8, ctx->stream));
    if (G2) {
        // b' = 3/(9+u) on the host
        Fp2 nine_u{Fp::from_u64(9), Fp::one()};
        Fp2 inv = Fp2::inv(nine_u);
        Fp three = Fp::from_u64(3);
        Fp2 b{Fp::mul(inv.a0, three), Fp::mul(inv.a1, three)};
        ZK_LAUNCH(ctx, k_g2_decode, grid_for(n, 128), 128, 0, (const uint8_t *)din, n, (int)compressed, (G2Affine *)dout, b, bad);
    } else {
        ZK_LAUNCH(ctx, k_g1_decode, grid_for(n, 128), 128, 0, (const uint8_t *)din, n, (int)compressed, (G1Affine *)dout, bad);
    }
    unsigned long long first_bad = ~0ull;
    ZK_CUDA(cudaMemcpyAsync(&first_bad, bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (!out_dev) {
        stage_begin(ctx, ST_D2H);
        ZK_CUDA(cudaMemcpyAsync(out_points, dout, n * out_sz, cudaMemcpyDeviceToHost, ctx->stream));
        stage_end(ctx, ST_D2H);
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    if (first_bad != ~0ull) { set_error("invalid point encoding at index %llu (not on the curve, coordinate >= q, or wrong flag)", first_bad); return ZKPOR_ERR_INVALID_ARG; }
    return ZKPOR_OK;
}

}  // namespace zk

using namespace zk;

extern "C" {
int32_t zkpor_g1_decode_batch(zkpor_ctx *ctx, const void *in_bytes, uint64_t n, int32_t compressed, void *out_points) {
    return decode_batch<false>(ctx, in_bytes, n, compressed, out_points);
}
int32_t zkpor_g2_decode_batch(zkpor_ctx *ctx, const void *in_bytes, uint64_t n, int32_t compressed, void *out_points) {
    return decode_batch<true>(ctx, in_bytes, n, compressed, out_points);
}
}
