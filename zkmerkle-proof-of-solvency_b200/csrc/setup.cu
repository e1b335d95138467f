// groth16.Setup building blocks on the GPU (reference call site src/keygen/main.go:42).
//
// gnark's Setup (backend/groth16/bn254/setup.go, out of tree) is: Lagrange basis at tau -> per-wire A_i(tau), B_i(tau),
// C_i(tau) (sparse column sums) -> K_i = (beta*A_i + alpha*B_i + C_i)/{gamma,delta}, Z_i = tau^i (tau^n - 1)/delta ->
// curve.BatchScalarMultiplicationG1/G2 of ~2^26 scalars each against the generators.  The last step dominates
// (hours on a CPU for the two 2^26 tiers); here it is one fixed-base pass: 8-bit windows, 32 mixed additions per
// scalar from a 32 x 255 table resident in L2, then Montgomery batch normalisation.
#include "internal.h"

using namespace ff;
using namespace ec;

namespace zk {

static const int FB_WINDOWS = 32, FB_ENTRIES = 255, FB_CHAIN = 64;

// table[w*255 + d-1] = d * 2^(8w) * base  (one block per window; thread d builds its entry by double-and-add)
template <class F>
__global__ void k_fb_table(Affine<F> base, XYZZ<F> *__restrict__ out) {
    const int w = blockIdx.x, d = threadIdx.x + 1;
    if (d > FB_ENTRIES) return;
    XYZZ<F> b = XYZZ<F>::from_affine(base);
    for (int k = 0; k < 8 * w; k++) b = b.dbl();
    out[w * FB_ENTRIES + d - 1] = b.mul_u32((uint32_t)d);
}

template <class F>
__global__ void __launch_bounds__(128) k_fixed_base(const Affine<F> *__restrict__ table, const Fr *__restrict__ scalars, uint64_t n, int mont,
                                                    XYZZ<F> *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s; { const uint4 *p = reinterpret_cast<const uint4 *>(scalars + i); uint4 *d = reinterpret_cast<uint4 *>(&s); d[0] = __ldg(p); d[1] = __ldg(p + 1); }
    if (mont) s = Fr::from_mont(s);
    XYZZ<F> acc = XYZZ<F>::inf();
#pragma unroll 1
    for (int w = 0; w < FB_WINDOWS; w++) {
        uint32_t d = (s.l[w >> 2] >> (8 * (w & 3))) & 0xffu;
        if (d) {
            Affine<F> p; const uint4 *src = reinterpret_cast<const uint4 *>(table + w * FB_ENTRIES + d - 1); uint4 *dst = reinterpret_cast<uint4 *>(&p);
#pragma unroll
            for (int k = 0; k < (int)(sizeof(Affine<F>) / 16); k++) dst[k] = __ldg(src + k);
            acc.add_affine(p, false);
        }
    }
    out[i] = acc;
}

// XYZZ -> affine with per-thread Montgomery batch inversion (runs of FB_CHAIN); infinity -> (0,0)
template <class F>
__global__ void __launch_bounds__(128) k_fb_normalise(const XYZZ<F> *__restrict__ in, uint64_t n, F *__restrict__ prefix, Affine<F> *__restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = t * FB_CHAIN;
    if (lo >= n) return;
    uint64_t hi = lo + FB_CHAIN < n ? lo + FB_CHAIN : n;
    F acc = F::one();
    for (uint64_t i = lo; i < hi; i++) { prefix[i] = acc; F z = in[i].ZZZ; if (!z.is_zero()) acc = F::mul(acc, z); }
    F inv = F::inv(acc);
    for (uint64_t i = hi; i-- > lo;) {
        XYZZ<F> p = in[i];
        if (p.ZZZ.is_zero()) { out[i] = Affine<F>::inf(); continue; }
        F zi = F::mul(inv, prefix[i]);
        inv = F::mul(inv, p.ZZZ);
        F zz_inv = F::sqr(F::mul(zi, p.ZZ));
        out[i] = Affine<F>{F::mul(p.X, zz_inv), F::mul(p.Y, zi)};
    }
}

template <class F>
static int32_t fixed_base_batch(zkpor_ctx *ctx, const void *base_affine, const void *scalars, uint64_t n, uint32_t flags, void *out_points) {
    ZK_REQUIRE(ctx && base_affine && ((scalars && out_points) || n == 0), "fixed_base_batch: null argument");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return ZKPOR_OK;
    stages_reset(ctx);
    Affine<F> base;
    if (is_device_ptr(base_affine)) ZK_CUDA(cudaMemcpy(&base, base_affine, sizeof base, cudaMemcpyDeviceToHost));
    else memcpy(&base, base_affine, sizeof base);
    const void *ds;
    stage_begin(ctx, ST_H2D);
    ZK_TRY(to_device(ctx, scalars, n * 32, ctx->in_scalars, &ds));
    stage_end(ctx, ST_H2D);
    const bool out_dev = is_device_ptr(out_points);
    void *dout = out_points;
    if (!out_dev) { ZK_TRY(ctx->in_points.reserve(n * sizeof(Affine<F>))); dout = ctx->in_points.p; }
    const size_t tab_n = (size_t)FB_WINDOWS * FB_ENTRIES;
    ZK_TRY(ctx->windows.reserve(tab_n * (sizeof(XYZZ<F>) + sizeof(Affine<F>) + sizeof(F))));
    XYZZ<F> *tab_x = ctx->windows.as<XYZZ<F>>();
    Affine<F> *tab_a = (Affine<F> *)(tab_x + tab_n);
    F *tab_p = (F *)(tab_a + tab_n);
    ZK_LAUNCH(ctx, (k_fb_table<F>), FB_WINDOWS, 256, 0, base, tab_x);
    ZK_LAUNCH(ctx, (k_fb_normalise<F>), grid_for((tab_n + FB_CHAIN - 1) / FB_CHAIN, 128), 128, 0, (const XYZZ<F> *)tab_x, (uint64_t)tab_n, tab_p, tab_a);
    ZK_TRY(ctx->buckets.reserve(n * sizeof(XYZZ<F>)));
    ZK_TRY(ctx->partials.reserve(n * sizeof(F)));
    ZK_LAUNCH(ctx, (k_fixed_base<F>), grid_for(n, 128), 128, 0, (const Affine<F> *)tab_a, (const Fr *)ds, n, (flags & ZKPOR_SCALARS_PLAIN) ? 0 : 1,
              ctx->buckets.as<XYZZ<F>>());
    ZK_LAUNCH(ctx, (k_fb_normalise<F>), grid_for((n + FB_CHAIN - 1) / FB_CHAIN, 128), 128, 0, (const XYZZ<F> *)ctx->buckets.p, n, ctx->partials.as<F>(),
              (Affine<F> *)dout);
    if (!out_dev) {
        stage_begin(ctx, ST_D2H);
        ZK_CUDA(cudaMemcpyAsync(out_points, dout, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
        stage_end(ctx, ST_D2H);
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    return ZKPOR_OK;
}

// ---- scalar side of Setup --------------------------------------------------------------------------------------------
// out[k] = (tau^n - 1)/n * w^k / (tau - w^k), k < n  -- the Lagrange basis of the size-n domain evaluated at tau
__global__ void __launch_bounds__(128) k_lagrange(Fr tau, Fr scale /* (tau^n-1)/n */, Fr gen, Fr ginv, uint32_t log_n, Fr *__restrict__ out) {
    const uint64_t n = 1ull << log_n;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = t * FB_CHAIN;
    if (lo >= n) return;
    uint64_t hi = lo + FB_CHAIN < n ? lo + FB_CHAIN : n;
    // w^lo by square-and-multiply, then a running product
    Fr wk = Fr::one(), b = gen;
    for (uint64_t e = lo; e; e >>= 1) { if (e & 1) wk = Fr::mul(wk, b); b = Fr::sqr(b); }
    Fr acc = Fr::one(), w = wk;
    for (uint64_t i = lo; i < hi; i++) { out[i] = acc; acc = Fr::mul(acc, Fr::sub(tau, w)); w = Fr::mul(w, gen); }   // prefix products in out[]
    Fr inv = Fr::inv(acc);
    // backwards: w currently = w^hi; step back with gen^-1 would cost an inversion, recompute from the stored prefix instead
    for (uint64_t i = hi; i-- > lo;) {
        w = Fr::mul(w, ginv);                          // w^i
        Fr den_inv = Fr::mul(inv, out[i]);             // 1/(tau - w^i)
        inv = Fr::mul(inv, Fr::sub(tau, w));
        out[i] = Fr::mul(Fr::mul(scale, w), den_inv);
    }
}

// out[i] = sum over the column-i entries (CSC) of coeff * lagrange[row]
__global__ void k_wire_sums(const uint64_t *__restrict__ col_ptr, const uint32_t *__restrict__ rows, const Fr *__restrict__ coeffs,
                            const Fr *__restrict__ lagrange, uint64_t n_wires, Fr *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_wires) return;
    Fr acc = Fr::zero();
    for (uint64_t e = col_ptr[i]; e < col_ptr[i + 1]; e++) acc = Fr::add(acc, Fr::mul(coeffs[e], lagrange[rows[e]]));
    out[i] = acc;
}
// out = (ka*a + kb*b + kc*c) * k
__global__ void k_lincomb3(const Fr *__restrict__ a, const Fr *__restrict__ b, const Fr *__restrict__ c, Fr ka, Fr kb, Fr kc, Fr k, uint64_t n,
                           Fr *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = Fr::add(Fr::add(Fr::mul(ka, a[i]), Fr::mul(kb, b[i])), Fr::mul(kc, c[i]));
    out[i] = Fr::mul(v, k);
}
// out[i] = first * ratio^e(i), e(i) = i or bitrev(i, log_n)
__global__ void k_powers(Fr first, Fr ratio, uint64_t n, uint32_t log_n, int bitrev, Fr *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t e = bitrev ? (uint64_t)(__brevll(i) >> (64 - log_n)) : i;
    Fr acc = first, b = ratio;
    for (; e; e >>= 1) { if (e & 1) acc = Fr::mul(acc, b); b = Fr::sqr(b); }
    out[i] = acc;
}

static Fr fr_from_be_mont(const uint8_t be[32]) { Fr p; fe_from_be32(&p, be); return Fr::to_mont(p); }

}  // namespace zk

using namespace zk;

extern "C" {

int32_t zkpor_g1_fixed_base_batch(zkpor_ctx *ctx, const void *base_affine64, const void *scalars, uint64_t n, uint32_t flags, void *out_points) {
    return fixed_base_batch<Fp>(ctx, base_affine64, scalars, n, flags, out_points);
}
int32_t zkpor_g2_fixed_base_batch(zkpor_ctx *ctx, const void *base_affine128, const void *scalars, uint64_t n, uint32_t flags, void *out_points) {
    return fixed_base_batch<Fp2>(ctx, base_affine128, scalars, n, flags, out_points);
}

int32_t zkpor_setup_lagrange(zkpor_ctx *ctx, const uint8_t tau_be[32], uint32_t log_n, void *out_dev) {
    ZK_REQUIRE(ctx && tau_be && out_dev, "setup_lagrange: null argument");
    ZK_REQUIRE(log_n >= 1 && log_n <= 28, "setup_lagrange: log_n out of range");
    ZK_REQUIRE(is_device_ptr(out_dev), "setup_lagrange: output must be device memory");
    ZK_CUDA(cudaSetDevice(ctx->device));
    static const uint32_t ROOT[8] = {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu, 0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u};
    Fr gen; memcpy(gen.l, ROOT, 32); gen = Fr::to_mont(gen);
    for (uint32_t i = log_n; i < 28; i++) gen = Fr::sqr(gen);
    Fr tau = fr_from_be_mont(tau_be);
    Fr tn = tau; for (uint32_t i = 0; i < log_n; i++) tn = Fr::sqr(tn);
    Fr scale = Fr::mul(Fr::sub(tn, Fr::one()), Fr::inv(Fr::from_u64(1ull << log_n)));
    const uint64_t n = 1ull << log_n;
    ZK_LAUNCH(ctx, k_lagrange, grid_for((n + FB_CHAIN - 1) / FB_CHAIN, 128), 128, 0, tau, scale, gen, Fr::inv(gen), log_n, (Fr *)out_dev);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

int32_t zkpor_setup_wire_sums(zkpor_ctx *ctx, const uint64_t *col_ptr, const uint32_t *rows, const void *coeffs, uint64_t nnz, const void *lagrange_dev,
                              uint64_t n_wires, void *out_dev) {
    ZK_REQUIRE(ctx && col_ptr && lagrange_dev && out_dev && (nnz == 0 || (rows && coeffs)), "setup_wire_sums: null argument");
    ZK_REQUIRE(is_device_ptr(lagrange_dev) && is_device_ptr(out_dev), "setup_wire_sums: lagrange / output must be device memory");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n_wires == 0) return ZKPOR_OK;
    const void *d_ptr, *d_rows = nullptr, *d_coef = nullptr;
    ZK_TRY(to_device(ctx, col_ptr, (n_wires + 1) * 8, ctx->io, &d_ptr));
    if (nnz) { ZK_TRY(to_device(ctx, rows, nnz * 4, ctx->in_scalars, &d_rows)); ZK_TRY(to_device(ctx, coeffs, nnz * 32, ctx->in_points, &d_coef)); }
    ZK_LAUNCH(ctx, k_wire_sums, grid_for(n_wires, 256), 256, 0, (const uint64_t *)d_ptr, (const uint32_t *)d_rows, (const Fr *)d_coef,
              (const Fr *)lagrange_dev, n_wires, (Fr *)out_dev);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

int32_t zkpor_fr_lincomb3(zkpor_ctx *ctx, const void *a, const void *b, const void *c, const uint8_t ka_be[32], const uint8_t kb_be[32],
                          const uint8_t kc_be[32], const uint8_t k_be[32], uint64_t n, void *out) {
    ZK_REQUIRE(ctx && a && b && c && ka_be && kb_be && kc_be && k_be && out, "fr_lincomb3: null argument");
    ZK_REQUIRE(is_device_ptr(a) && is_device_ptr(b) && is_device_ptr(c) && is_device_ptr(out), "fr_lincomb3: operands must be device memory");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n) ZK_LAUNCH(ctx, k_lincomb3, grid_for(n, 256), 256, 0, (const Fr *)a, (const Fr *)b, (const Fr *)c, fr_from_be_mont(ka_be), fr_from_be_mont(kb_be),
                     fr_from_be_mont(kc_be), fr_from_be_mont(k_be), n, (Fr *)out);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

int32_t zkpor_fr_powers(zkpor_ctx *ctx, const uint8_t first_be[32], const uint8_t ratio_be[32], uint64_t n, uint32_t log_n, int32_t bitrev, void *out_dev) {
    ZK_REQUIRE(ctx && first_be && ratio_be && out_dev, "fr_powers: null argument");
    ZK_REQUIRE(is_device_ptr(out_dev), "fr_powers: output must be device memory");
    ZK_REQUIRE(!bitrev || (log_n >= 1 && log_n <= 28 && n <= (1ull << log_n)), "fr_powers: bit-reversed order needs n <= 2^log_n");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n) ZK_LAUNCH(ctx, k_powers, grid_for(n, 256), 256, 0, fr_from_be_mont(first_be), fr_from_be_mont(ratio_be), n, log_n ? log_n : 1u, (int)bitrev, (Fr *)out_dev);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

}  // extern "C"
