// Synthetic workload generators (bench / full-size parity tooling, not on the proving path).
//
// The reference's 12 GB proving keys cannot be produced here (groth16.Setup needs the gnark frontend; SURVEY.md 8(d)
// config 3), so bench.py and the full-size tests build a key of the same SHAPE directly in HBM:
//   points[i] = (k0 + i*d) * G        -- distinct curve points with KNOWN discrete logs, so the result of a 2^26-term
//                                        MSM can be checked exactly by the oracle as (sum_i s_i (k0 + i d)) * G;
//   scalars   = counter-based uniform Fr elements (or the witness-like mix of SURVEY.md 8(d) config 2).
// Generation: one scalar multiplication per thread, then a chain of mixed additions, then Montgomery batch
// normalisation to affine -- ~25 field multiplications per point instead of ~3500.
#include "internal.h"

using namespace ff;
using namespace ec;

namespace zk {

static const int CHAIN = 64;

template <class F>
__global__ void __launch_bounds__(128) k_chain(Affine<F> gen, Affine<F> step, Fr k0, Fr d, uint64_t n, XYZZ<F> *__restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = t * CHAIN;
    if (lo >= n) return;
    uint64_t hi = lo + CHAIN < n ? lo + CHAIN : n;
    Fr k = Fr::from_mont(Fr::add(k0, Fr::mul(d, Fr::from_u64(lo))));
    XYZZ<F> acc = XYZZ<F>::from_affine(gen).mul_256(k.l);
    for (uint64_t i = lo; i < hi; i++) { out[i] = acc; acc.add_affine(step, false); }
}

// Montgomery batch inversion of ZZZ over runs of CHAIN points, then x = X/ZZ, y = Y/ZZZ
template <class F>
__global__ void __launch_bounds__(128) k_normalise(const XYZZ<F> *__restrict__ in, uint64_t n, F *__restrict__ prefix, Affine<F> *__restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = t * CHAIN;
    if (lo >= n) return;
    uint64_t hi = lo + CHAIN < n ? lo + CHAIN : n;
    F acc = F::one();
    for (uint64_t i = lo; i < hi; i++) { prefix[i] = acc; acc = F::mul(acc, in[i].ZZZ); }
    F inv = F::inv(acc);
    for (uint64_t i = hi; i-- > lo;) {
        XYZZ<F> p = in[i];
        F zi = F::mul(inv, prefix[i]);          // 1/ZZZ_i
        inv = F::mul(inv, p.ZZZ);
        F zz_inv = F::sqr(F::mul(zi, p.ZZ));    // 1/ZZ_i
        out[i] = Affine<F>{F::mul(p.X, zz_inv), F::mul(p.Y, zi)};
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ bool lt_modulus(const Fr &v) {
    for (int i = 7; i >= 0; i--) { if (v.l[i] < FrParams::M(i)) return true; if (v.l[i] > FrParams::M(i)) return false; }
    return false;
}
// kind 0: uniform in [0, r).  kind 1: 60% zero, 15% one, 20% < 2^16, 5% uniform (witness-like), canonical integers.
// kind 2: the kind-1 values in Montgomery form -- what a gnark wire vector holds.  For kind 0 the limb pattern can be
// read as Montgomery or canonical: uniform either way.
__global__ void k_synth_scalars(uint64_t seed, uint64_t n, int kind, Fr *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = Fr::zero();
    uint64_t base = mix64(seed ^ (i * 0xD1342543DE82EF95ull));
    bool uniform = true;
    if (kind == 1 || kind == 2) {
        uint32_t t = (uint32_t)(base % 100);
        if (t < 60) uniform = false;
        else if (t < 75) { v.l[0] = 1; uniform = false; }
        else if (t < 95) { v.l[0] = (uint32_t)(base >> 32) & 0xFFFF; uniform = false; }
    }
    if (uniform) {
        for (uint64_t ctr = 0;; ctr++) {
            for (int k = 0; k < 4; k++) { uint64_t w = mix64(base + 4 * ctr + k + 1); v.l[2 * k] = (uint32_t)w; v.l[2 * k + 1] = (uint32_t)(w >> 32); }
            v.l[7] &= 0x3FFFFFFFu;
            if (lt_modulus(v)) break;
        }
    }
    if (kind == 2) v = Fr::to_mont(v);
    uint4 *d = reinterpret_cast<uint4 *>(out + i); const uint4 *s = reinterpret_cast<const uint4 *>(&v);
    d[0] = s[0]; d[1] = s[1];
}

template <class F>
static int32_t synth_points(zkpor_ctx *ctx, const Affine<F> &gen, const uint8_t k0_be[32], const uint8_t d_be[32], uint64_t n, void *out_dev) {
    ZK_REQUIRE(ctx && k0_be && d_be && out_dev, "synth_points: null argument");
    ZK_REQUIRE(is_device_ptr(out_dev), "synth_points: output must be device memory");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return ZKPOR_OK;
    Fr k0p, dp; fe_from_be32(&k0p, k0_be); fe_from_be32(&dp, d_be);
    Affine<F> step = XYZZ<F>::from_affine(gen).mul_256(dp.l).to_affine();
    ZK_TRY(ctx->buckets.reserve(n * sizeof(XYZZ<F>)));
    ZK_TRY(ctx->partials.reserve(n * sizeof(F)));
    const uint64_t threads = (n + CHAIN - 1) / CHAIN;
    ZK_LAUNCH(ctx, (k_chain<F>), grid_for(threads, 128), 128, 0, gen, step, Fr::to_mont(k0p), Fr::to_mont(dp), n, ctx->buckets.as<XYZZ<F>>());
    ZK_LAUNCH(ctx, (k_normalise<F>), grid_for(threads, 128), 128, 0, (const XYZZ<F> *)ctx->buckets.p, n, ctx->partials.as<F>(), (Affine<F> *)out_dev);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

}  // namespace zk

using namespace zk;

extern "C" {

int32_t zkpor_synth_points_g1(zkpor_ctx *ctx, const uint8_t k0_be[32], const uint8_t d_be[32], uint64_t n, void *out_dev) {
    G1Affine g{Fp::one(), Fp::from_u64(2)};
    return synth_points<Fp>(ctx, g, k0_be, d_be, n, out_dev);
}
int32_t zkpor_synth_points_g2(zkpor_ctx *ctx, const uint8_t k0_be[32], const uint8_t d_be[32], uint64_t n, void *out_dev) {
    static const uint32_t G2X0[8] = {0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu};
    static const uint32_t G2X1[8] = {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u};
    static const uint32_t G2Y0[8] = {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u};
    static const uint32_t G2Y1[8] = {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};
    G2Affine g;
    memcpy(g.x.a0.l, G2X0, 32); memcpy(g.x.a1.l, G2X1, 32); memcpy(g.y.a0.l, G2Y0, 32); memcpy(g.y.a1.l, G2Y1, 32);
    g.x.a0 = Fp::to_mont(g.x.a0); g.x.a1 = Fp::to_mont(g.x.a1); g.y.a0 = Fp::to_mont(g.y.a0); g.y.a1 = Fp::to_mont(g.y.a1);
    return synth_points<Fp2>(ctx, g, k0_be, d_be, n, out_dev);
}
int32_t zkpor_synth_scalars(zkpor_ctx *ctx, uint64_t seed, uint64_t n, int32_t kind, void *out_dev) {
    ZK_REQUIRE(ctx && out_dev, "synth_scalars: null argument");
    ZK_REQUIRE(is_device_ptr(out_dev), "synth_scalars: output must be device memory");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return ZKPOR_OK;
    ZK_LAUNCH(ctx, k_synth_scalars, grid_for(n, 256), 256, 0, seed, n, (int)kind, (Fr *)out_dev);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

}  // extern "C"
