// Scalar-field NTT on sm_100a and the Groth16 quotient computation.
//
// Replaces gnark-crypto fft.Domain.FFT / FFTInverse (ecc/bn254/fr/fft, out of tree) and gnark's computeH
// (backend/groth16/bn254/prove.go, out of tree), both reached from src/prover/prover/prover.go:269.  Conventions are
// gnark's: DIF = natural in / bit-reversed out, DIT = bit-reversed in / natural out, coset shift 5, inverse scales by
// 1/n.  Each pass keeps 2^R elements of one butterfly group in registers and runs R levels on them (R = 3: radix-8),
// so a 2^26 transform is 9 read+write sweeps of the vector; twiddles come from one cached table w^j, j < n/2.
#include <map>
#include "internal.h"

using namespace ff;

namespace zk {

static const uint32_t ROOT_2_28_PLAIN[8] = {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu, 0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u};

struct NttDomain {
    uint32_t log_n = 0;
    Fr *tw_fwd = nullptr, *tw_inv = nullptr;          // w^j, w^-j for j < n/2
    Fr *cf_lo = nullptr, *cf_hi = nullptr;            // coset 5^e, two-level: lo[e & LO_MASK] * hi[e >> LO_BITS]
    Fr *cs_lo = nullptr, *cs_hi = nullptr;            // 5^e / n (inverse-then-coset fused scaling of computeH)
    Fr *ci_lo = nullptr, *ci_hi = nullptr;            // 5^-e * den / n
    Fr *ni_lo = nullptr, *ni_hi = nullptr;            // 5^-e / n (plain coset inverse)
    Fr n_inv;                                         // 1/n
};
struct NttCache { std::map<uint32_t, NttDomain> doms; };

static const uint32_t LO_BITS = 12;

// out[k] = extra * base^k
__global__ void k_pow_table(Fr *out, Fr base, Fr extra, uint32_t count) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    Fr acc = extra, b = base;
    for (uint32_t e = k; e; e >>= 1) { if (e & 1) acc = Fr::mul(acc, b); b = Fr::sqr(b); }
    out[k] = acc;
}
// out[j] = lo[j & mask] * hi[j >> LO_BITS]
__global__ void k_table_product(Fr *out, const Fr *lo, const Fr *hi, size_t count) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    out[j] = Fr::mul(lo[j & ((1u << LO_BITS) - 1)], hi[j >> LO_BITS]);
}

__device__ __forceinline__ Fr ld_fr(const Fr *p) {
    Fr r; const uint4 *s = reinterpret_cast<const uint4 *>(p); uint4 *d = reinterpret_cast<uint4 *>(&r);
    d[0] = s[0]; d[1] = s[1]; return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr *p) {
    Fr r; const uint4 *s = reinterpret_cast<const uint4 *>(p); uint4 *d = reinterpret_cast<uint4 *>(&r);
    d[0] = __ldg(s); d[1] = __ldg(s + 1); return r;
}
__device__ __forceinline__ void st_fr(Fr *p, const Fr &v) {
    uint4 *d = reinterpret_cast<uint4 *>(p); const uint4 *s = reinterpret_cast<const uint4 *>(&v);
    d[0] = s[0]; d[1] = s[1];
}

// R levels of decimation-in-frequency starting at level `level` (level 0 has half-size n/2)
template <int R>
__global__ void __launch_bounds__(128, 4) k_ntt_dif(Fr *__restrict__ data, const Fr *__restrict__ tw, uint32_t log_n, uint32_t level) {
    const uint32_t log_h = log_n - 1 - level, log_q = log_h - (R - 1);
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ((size_t)1 << (log_n - R))) return;
    size_t j = t & (((size_t)1 << log_q) - 1), blk = t >> log_q;
    Fr *p = data + (blk << (log_h + 1)) + j;
    Fr x[1 << R];
#pragma unroll
    for (int k = 0; k < (1 << R); k++) x[k] = ld_fr(p + ((size_t)k << log_q));
#pragma unroll
    for (int l = 0; l < R; l++) {
        const int dist = 1 << (R - 1 - l);
        const uint32_t sh = log_n - 1 - (log_h - l);
#pragma unroll
        for (int k = 0; k < (1 << R); k++) {
            if (k & dist) continue;
            size_t o = j + ((size_t)(k & (dist - 1)) << log_q);
            Fr w = ldg_fr(tw + (o << sh));
            Fr u = x[k], v = x[k + dist];
            x[k] = Fr::add(u, v);
            x[k + dist] = Fr::mul(Fr::sub(u, v), w);
        }
    }
#pragma unroll
    for (int k = 0; k < (1 << R); k++) st_fr(p + ((size_t)k << log_q), x[k]);
}

// R levels of decimation-in-time starting at level `level` (level 0 has half-size 1)
template <int R>
__global__ void __launch_bounds__(128, 4) k_ntt_dit(Fr *__restrict__ data, const Fr *__restrict__ tw, uint32_t log_n, uint32_t level) {
    const uint32_t log_q = level;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ((size_t)1 << (log_n - R))) return;
    size_t j = t & (((size_t)1 << log_q) - 1), blk = t >> log_q;
    Fr *p = data + (blk << (log_q + R)) + j;
    Fr x[1 << R];
#pragma unroll
    for (int k = 0; k < (1 << R); k++) x[k] = ld_fr(p + ((size_t)k << log_q));
#pragma unroll
    for (int l = 0; l < R; l++) {
        const int dist = 1 << l;
        const uint32_t sh = log_n - 1 - (log_q + l);
#pragma unroll
        for (int k = 0; k < (1 << R); k++) {
            if (k & dist) continue;
            size_t o = j + ((size_t)(k & (dist - 1)) << log_q);
            Fr w = ldg_fr(tw + (o << sh));
            Fr u = x[k], v = Fr::mul(x[k + dist], w);
            x[k] = Fr::add(u, v);
            x[k + dist] = Fr::sub(u, v);
        }
    }
#pragma unroll
    for (int k = 0; k < (1 << R); k++) st_fr(p + ((size_t)k << log_q), x[k]);
}

// data[i] *= lo[e & mask] * hi[e >> LO_BITS], e = i or bitrev(i)
__global__ void k_scale_pow(Fr *__restrict__ data, const Fr *__restrict__ lo, const Fr *__restrict__ hi, uint32_t log_n, int bitrev) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_n)) return;
    uint32_t e = bitrev ? (__brev((uint32_t)i) >> (32 - log_n)) : (uint32_t)i;
    Fr f = Fr::mul(ldg_fr(lo + (e & ((1u << LO_BITS) - 1))), ldg_fr(hi + (e >> LO_BITS)));
    st_fr(data + i, Fr::mul(ld_fr(data + i), f));
}
__global__ void k_scale_const(Fr *__restrict__ data, Fr f, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(data + i, Fr::mul(ld_fr(data + i), f));
}
__global__ void k_mul_vec(const Fr *__restrict__ a, const Fr *__restrict__ b, Fr *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, Fr::mul(ld_fr(a + i), ld_fr(b + i)));
}
// a = a*b - c
__global__ void k_ab_minus_c(Fr *__restrict__ a, const Fr *__restrict__ b, const Fr *__restrict__ c, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(a + i, Fr::sub(Fr::mul(ld_fr(a + i), ld_fr(b + i)), ld_fr(c + i)));
}

static Fr host_pow_u64(Fr base, uint64_t e) {
    Fr acc = Fr::one();
    for (; e; e >>= 1) { if (e & 1) acc = Fr::mul(acc, base); base = Fr::sqr(base); }
    return acc;
}

static int32_t build_two_level(zkpor_ctx *ctx, Fr base, Fr extra, uint32_t log_count, Fr **lo, Fr **hi) {
    const uint32_t lo_n = log_count < LO_BITS ? (1u << log_count) : (1u << LO_BITS);
    const uint32_t hi_n = log_count < LO_BITS ? 1u : (1u << (log_count - LO_BITS));
    ZK_CUDA(cudaMalloc((void **)lo, sizeof(Fr) * lo_n));
    ZK_CUDA(cudaMalloc((void **)hi, sizeof(Fr) * hi_n));
    ZK_LAUNCH(ctx, k_pow_table, grid_for(lo_n, 128), 128, 0, *lo, base, Fr::one(), lo_n);
    ZK_LAUNCH(ctx, k_pow_table, grid_for(hi_n, 128), 128, 0, *hi, host_pow_u64(base, 1ull << LO_BITS), extra, hi_n);
    return ZKPOR_OK;
}

static int32_t get_domain(zkpor_ctx *ctx, uint32_t log_n, NttDomain **out) {
    if (!ctx->ntt_tables) ctx->ntt_tables = new NttCache();
    NttCache *cache = (NttCache *)ctx->ntt_tables;
    auto it = cache->doms.find(log_n);
    if (it != cache->doms.end()) { *out = &it->second; return ZKPOR_OK; }
    NttDomain d; d.log_n = log_n;
    Fr root; memcpy(root.l, ROOT_2_28_PLAIN, 32); root = Fr::to_mont(root);
    for (uint32_t i = log_n; i < 28; i++) root = Fr::sqr(root);
    const Fr gen = root, gen_inv = Fr::inv(root);
    const Fr five = Fr::from_u64(5), five_inv = Fr::inv(five);
    d.n_inv = Fr::inv(Fr::from_u64(1ull << log_n));
    Fr den = five; for (uint32_t i = 0; i < log_n; i++) den = Fr::sqr(den);     // 5^n
    den = Fr::inv(Fr::sub(den, Fr::one()));
    const size_t half = log_n ? ((size_t)1 << (log_n - 1)) : 1;
    // twiddle tables via a temporary two-level table
    for (int dir = 0; dir < 2; dir++) {
        Fr *lo, *hi, *full;
        ZK_TRY(build_two_level(ctx, dir ? gen_inv : gen, Fr::one(), log_n ? log_n - 1 : 0, &lo, &hi));
        ZK_CUDA(cudaMalloc((void **)&full, sizeof(Fr) * half));
        ZK_LAUNCH(ctx, k_table_product, grid_for(half, 256), 256, 0, full, lo, hi, half);
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(lo); cudaFree(hi);
        (dir ? d.tw_inv : d.tw_fwd) = full;
    }
    ZK_TRY(build_two_level(ctx, five, Fr::one(), log_n, &d.cf_lo, &d.cf_hi));
    ZK_TRY(build_two_level(ctx, five, d.n_inv, log_n, &d.cs_lo, &d.cs_hi));
    ZK_TRY(build_two_level(ctx, five_inv, Fr::mul(d.n_inv, den), log_n, &d.ci_lo, &d.ci_hi));
    ZK_TRY(build_two_level(ctx, five_inv, d.n_inv, log_n, &d.ni_lo, &d.ni_hi));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    cache->doms[log_n] = d;
    *out = &cache->doms[log_n];
    return ZKPOR_OK;
}

static int32_t run_dif(zkpor_ctx *ctx, Fr *d, const Fr *tw, uint32_t log_n) {
    uint32_t level = 0;
    while (level < log_n) {
        uint32_t R = log_n - level >= 3 ? 3 : log_n - level;
        size_t threads = (size_t)1 << (log_n - R);
        KTimed kt(ctx, KC_NTT_PASS, (uint64_t)1 << log_n);
        if (R == 3) ZK_LAUNCH(ctx, k_ntt_dif<3>, grid_for(threads, 128), 128, 0, d, tw, log_n, level);
        else if (R == 2) ZK_LAUNCH(ctx, k_ntt_dif<2>, grid_for(threads, 128), 128, 0, d, tw, log_n, level);
        else ZK_LAUNCH(ctx, k_ntt_dif<1>, grid_for(threads, 128), 128, 0, d, tw, log_n, level);
        kt.stop();
        level += R;
    }
    return ZKPOR_OK;
}
static int32_t run_dit(zkpor_ctx *ctx, Fr *d, const Fr *tw, uint32_t log_n) {
    uint32_t level = 0;
    while (level < log_n) {
        uint32_t R = log_n - level >= 3 ? 3 : log_n - level;
        size_t threads = (size_t)1 << (log_n - R);
        KTimed kt(ctx, KC_NTT_PASS, (uint64_t)1 << log_n);
        if (R == 3) ZK_LAUNCH(ctx, k_ntt_dit<3>, grid_for(threads, 128), 128, 0, d, tw, log_n, level);
        else if (R == 2) ZK_LAUNCH(ctx, k_ntt_dit<2>, grid_for(threads, 128), 128, 0, d, tw, log_n, level);
        else ZK_LAUNCH(ctx, k_ntt_dit<1>, grid_for(threads, 128), 128, 0, d, tw, log_n, level);
        kt.stop();
        level += R;
    }
    return ZKPOR_OK;
}

int32_t ntt_dev(zkpor_ctx *ctx, Fr *d, uint32_t log_n, bool inverse, bool dit, bool coset) {
    ZK_REQUIRE(log_n <= 28, "ntt: log_n exceeds the 2-adicity of Fr (28)");
    if (log_n == 0) return ZKPOR_OK;
    NttDomain *dom; ZK_TRY(get_domain(ctx, log_n, &dom));
    const size_t n = (size_t)1 << log_n;
    if (!inverse) {
        if (coset) ZK_LAUNCH(ctx, k_scale_pow, grid_for(n, 256), 256, 0, d, dom->cf_lo, dom->cf_hi, log_n, dit ? 1 : 0);
        ZK_TRY(dit ? run_dit(ctx, d, dom->tw_fwd, log_n) : run_dif(ctx, d, dom->tw_fwd, log_n));
    } else {
        ZK_TRY(dit ? run_dit(ctx, d, dom->tw_inv, log_n) : run_dif(ctx, d, dom->tw_inv, log_n));
        if (coset) ZK_LAUNCH(ctx, k_scale_pow, grid_for(n, 256), 256, 0, d, dom->ni_lo, dom->ni_hi, log_n, dit ? 0 : 1);
        else ZK_LAUNCH(ctx, k_scale_const, grid_for(n, 256), 256, 0, d, dom->n_inv, n);
    }
    return ZKPOR_OK;
}

// a, b, c: device vectors of n = 2^log_n elements (zero padded).  Result h in a, bit-reversed coefficient order.
//   x <- DIF_inv(x); x[pos] *= 5^bitrev(pos)/n; x <- DIT_fwd(x)        for x in a, b, c
//   a <- a*b - c ; a <- DIF_inv(a) ; a[pos] *= 5^-bitrev(pos) * den/n
int32_t compute_h_dev(zkpor_ctx *ctx, Fr *a, Fr *b, Fr *c, uint32_t log_n) {
    ZK_REQUIRE(log_n >= 1 && log_n <= 28, "compute_h: log_n out of range");
    NttDomain *dom; ZK_TRY(get_domain(ctx, log_n, &dom));
    const size_t n = (size_t)1 << log_n;
    stage_begin(ctx, ST_NTT);
    Fr *v[3] = {a, b, c};
    for (int k = 0; k < 3; k++) {
        ZK_TRY(run_dif(ctx, v[k], dom->tw_inv, log_n));
        ZK_LAUNCH(ctx, k_scale_pow, grid_for(n, 256), 256, 0, v[k], dom->cs_lo, dom->cs_hi, log_n, 1);
        ZK_TRY(run_dit(ctx, v[k], dom->tw_fwd, log_n));
    }
    ZK_LAUNCH(ctx, k_ab_minus_c, grid_for(n, 256), 256, 0, a, (const Fr *)b, (const Fr *)c, n);
    ZK_TRY(run_dif(ctx, a, dom->tw_inv, log_n));
    ZK_LAUNCH(ctx, k_scale_pow, grid_for(n, 256), 256, 0, a, dom->ci_lo, dom->ci_hi, log_n, 1);
    stage_end(ctx, ST_NTT);
    return ZKPOR_OK;
}

// ------------------------------------------------------------------------------------------------ computeH across N GPUs
// One size-n transform split over N = 2^k ranks with ONE all-to-all: rank g holds the cyclic subsequence
// x_g[j] = x[g + N j] (m = n/N elements).
//   inverse (evaluations -> coefficients):  X[k' + m t] = sum_g w^(-g(k' + m t)) Y_g[k'],  Y_g = the size-m inverse transform of x_g:
//     local DIF (bit-reversed positions p, k' = bitrev(p)), scale by w^(-g k')/n, all-to-all (rank r takes positions [r m/N, (r+1) m/N)
//     of every g), then a radix-N butterfly over g.  Rank r ends up with X[k' + m t] for its positions and every t, stored at
//     [q N + bitrev_k(t)]: with that order the ranks' pieces are exactly the contiguous chunks [r m, (r+1) m) of the bit-reversed
//     coefficient vector -- the order gnark stores pk.G1.Z in, so the Z multiplication shards by plain point chunks.
//   forward (coefficients in that order -> evaluations, cyclic again): radix-N butterfly over t, all-to-all back, scale by w^(g k'),
//     local DIT.
// computeH = inverse, coset scale 5^k, forward (for a, b, c), a*b - c, inverse, scale 5^(-k) den: three exchange phases (3 + 3 + 1 vectors).
struct DistDomain {
    uint32_t log_n = 0, log_w = 0; int rank = -1;
    Fr *inv_lo = nullptr, *inv_hi = nullptr;     // (w^-g)^e / n
    Fr *invd_lo = nullptr, *invd_hi = nullptr;   // (w^-g)^e * den / n
    Fr *fwd_lo = nullptr, *fwd_hi = nullptr;     // (w^g)^e
    Fr *c5_lo = nullptr, *c5_hi = nullptr, *c5i_lo = nullptr, *c5i_hi = nullptr;   // 5^e, 5^-e for e < m
    Fr *consts = nullptr;                        // om[g*N + t] = w^(m g t) (N*N), then om_inv (N*N), then 5^(m t), 5^(-m t) (N each)
};
struct DistCache { std::map<uint64_t, DistDomain> doms; };

// in[g*chunk + q] (one value per source rank g) -> out[q*N + bitrev_k(t)] = (sum_g in[g][q] * om_inv[g][t]) * c(k') * cm[t],
// k' = bitrev(first + q): the radix-N butterfly that completes an inverse transform, with the coset factor 5^(+-(k' + m t)) folded in
__global__ void k_dist_bfly_inv(const Fr *__restrict__ in, Fr *__restrict__ out, const Fr *__restrict__ om_inv, const Fr *__restrict__ cm,
                                const Fr *__restrict__ c_lo, const Fr *__restrict__ c_hi, uint32_t log_m, uint32_t log_w, size_t first, size_t chunk) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= chunk) return;
    const uint32_t N = 1u << log_w;
    const uint32_t kp = log_m ? (__brev((uint32_t)(first + q)) >> (32 - log_m)) : 0;
    const Fr base = Fr::mul(ldg_fr(c_lo + (kp & ((1u << LO_BITS) - 1))), ldg_fr(c_hi + (kp >> LO_BITS)));
    Fr x[8];
    for (uint32_t g = 0; g < N; g++) x[g] = ld_fr(in + (size_t)g * chunk + q);
    for (uint32_t t = 0; t < N; t++) {
        Fr acc = x[0];
        for (uint32_t g = 1; g < N; g++) acc = Fr::add(acc, Fr::mul(x[g], ldg_fr(om_inv + g * N + t)));
        const uint32_t tr = log_w ? (__brev(t) >> (32 - log_w)) : 0;
        st_fr(out + q * N + tr, Fr::mul(acc, Fr::mul(base, ldg_fr(cm + t))));
    }
}
// in[q*N + bitrev_k(t)] -> out[g*chunk + q] = sum_t in[q][t] * om[g][t]: the radix-N butterfly that opens a forward transform
__global__ void k_dist_bfly_fwd(const Fr *__restrict__ in, Fr *__restrict__ out, const Fr *__restrict__ om, uint32_t log_w, size_t chunk) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= chunk) return;
    const uint32_t N = 1u << log_w;
    Fr x[8];
    for (uint32_t t = 0; t < N; t++) { const uint32_t tr = log_w ? (__brev(t) >> (32 - log_w)) : 0; x[t] = ld_fr(in + q * N + tr); }
    for (uint32_t g = 0; g < N; g++) {
        Fr acc = x[0];
        for (uint32_t t = 1; t < N; t++) acc = Fr::add(acc, Fr::mul(x[t], ldg_fr(om + g * N + t)));
        st_fr(out + (size_t)g * chunk + q, acc);
    }
}

static int32_t get_dist_domain(zkpor_ctx *ctx, uint32_t log_n, uint32_t log_w, int rank, DistDomain **out) {
    if (!ctx->dist_tables) ctx->dist_tables = new DistCache();
    DistCache *cache = (DistCache *)ctx->dist_tables;
    const uint64_t key = ((uint64_t)log_n << 32) | ((uint64_t)log_w << 16) | (uint64_t)rank;
    auto it = cache->doms.find(key);
    if (it != cache->doms.end()) { *out = &it->second; return ZKPOR_OK; }
    DistDomain d; d.log_n = log_n; d.log_w = log_w; d.rank = rank;
    const uint32_t log_m = log_n - log_w, N = 1u << log_w;
    Fr root; memcpy(root.l, ROOT_2_28_PLAIN, 32); root = Fr::to_mont(root);
    for (uint32_t i = log_n; i < 28; i++) root = Fr::sqr(root);
    const Fr w = root, w_inv = Fr::inv(root);
    const Fr five = Fr::from_u64(5), five_inv = Fr::inv(five);
    const Fr n_inv = Fr::inv(Fr::from_u64(1ull << log_n));
    Fr den = five; for (uint32_t i = 0; i < log_n; i++) den = Fr::sqr(den);
    den = Fr::inv(Fr::sub(den, Fr::one()));
    ZK_TRY(build_two_level(ctx, host_pow_u64(w_inv, (uint64_t)rank), n_inv, log_m, &d.inv_lo, &d.inv_hi));
    ZK_TRY(build_two_level(ctx, host_pow_u64(w_inv, (uint64_t)rank), Fr::mul(n_inv, den), log_m, &d.invd_lo, &d.invd_hi));
    ZK_TRY(build_two_level(ctx, host_pow_u64(w, (uint64_t)rank), Fr::one(), log_m, &d.fwd_lo, &d.fwd_hi));
    ZK_TRY(build_two_level(ctx, five, Fr::one(), log_m, &d.c5_lo, &d.c5_hi));
    ZK_TRY(build_two_level(ctx, five_inv, Fr::one(), log_m, &d.c5i_lo, &d.c5i_hi));
    std::vector<Fr> hc(2 * (size_t)N * N + 2 * N);
    const Fr om = host_pow_u64(w, 1ull << log_m), om_inv = host_pow_u64(w_inv, 1ull << log_m);   // primitive N-th roots
    const Fr c5m = host_pow_u64(five, 1ull << log_m), c5mi = host_pow_u64(five_inv, 1ull << log_m);
    for (uint32_t g = 0; g < N; g++) for (uint32_t t = 0; t < N; t++) {
        hc[(size_t)g * N + t] = host_pow_u64(om, (uint64_t)g * t);
        hc[(size_t)N * N + (size_t)g * N + t] = host_pow_u64(om_inv, (uint64_t)g * t);
    }
    for (uint32_t t = 0; t < N; t++) { hc[2 * (size_t)N * N + t] = host_pow_u64(c5m, t); hc[2 * (size_t)N * N + N + t] = host_pow_u64(c5mi, t); }
    ZK_CUDA(cudaMalloc((void **)&d.consts, hc.size() * sizeof(Fr)));
    ZK_CUDA(cudaMemcpy(d.consts, hc.data(), hc.size() * sizeof(Fr), cudaMemcpyHostToDevice));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    cache->doms[key] = d;
    *out = &cache->doms[key];
    return ZKPOR_OK;
}

// x: this rank's m elements in cyclic order; result: this rank's chunk of the bit-reversed coefficient vector, times 5^(+k) (coset > 0)
// or 5^(-k) den (coset < 0).  tmp: m elements of scratch.
static int32_t dist_inverse(zkpor_ctx *ctx, DistDomain *dd, NttDomain *loc, Fr *x, Fr *tmp, int coset) {
    const uint32_t log_m = dd->log_n - dd->log_w, N = 1u << dd->log_w;
    const size_t m = (size_t)1 << log_m, chunk = m >> dd->log_w;
    ZK_TRY(run_dif(ctx, x, loc->tw_inv, log_m));
    if (coset < 0) ZK_LAUNCH(ctx, k_scale_pow, grid_for(m, 256), 256, 0, x, dd->invd_lo, dd->invd_hi, log_m, 1);
    else ZK_LAUNCH(ctx, k_scale_pow, grid_for(m, 256), 256, 0, x, dd->inv_lo, dd->inv_hi, log_m, 1);
    ZK_TRY(comm_all_to_all(ctx, x, tmp, chunk * sizeof(Fr)));
    const Fr *lo = coset > 0 ? dd->c5_lo : dd->c5i_lo, *hi = coset > 0 ? dd->c5_hi : dd->c5i_hi;
    const Fr *om_inv = dd->consts + (size_t)N * N, *cm = dd->consts + 2 * (size_t)N * N + (coset > 0 ? 0 : N);
    ZK_LAUNCH(ctx, k_dist_bfly_inv, grid_for(chunk, 128), 128, 0, (const Fr *)tmp, x, om_inv, cm, lo, hi, log_m, dd->log_w, (size_t)dd->rank * chunk, chunk);
    return ZKPOR_OK;
}

// x: this rank's chunk of the bit-reversed coefficient vector; result: this rank's m evaluations in cyclic order
static int32_t dist_forward(zkpor_ctx *ctx, DistDomain *dd, NttDomain *loc, Fr *x, Fr *tmp) {
    const uint32_t log_m = dd->log_n - dd->log_w;
    const size_t m = (size_t)1 << log_m, chunk = m >> dd->log_w;
    ZK_LAUNCH(ctx, k_dist_bfly_fwd, grid_for(chunk, 128), 128, 0, (const Fr *)x, tmp, (const Fr *)dd->consts, dd->log_w, chunk);
    ZK_TRY(comm_all_to_all(ctx, tmp, x, chunk * sizeof(Fr)));
    ZK_LAUNCH(ctx, k_scale_pow, grid_for(m, 256), 256, 0, x, dd->fwd_lo, dd->fwd_hi, log_m, 1);
    ZK_TRY(run_dit(ctx, x, loc->tw_fwd, log_m));
    return ZKPOR_OK;
}

int32_t compute_h_dist(zkpor_ctx *ctx, Fr *a, Fr *b, Fr *c, Fr *tmp, uint32_t log_n) {
    int rank = 0, world = 1;
    comm_info(ctx, &rank, &world);
    uint32_t log_w = 0;
    while ((1 << log_w) < world) log_w++;
    ZK_REQUIRE((1 << log_w) == world && log_w <= 3, "compute_h: the number of ranks must be 1, 2, 4 or 8");
    ZK_REQUIRE(log_n >= 2 * log_w + 1 && log_n <= 28, "compute_h: domain too small for this many ranks");
    DistDomain *dd; ZK_TRY(get_dist_domain(ctx, log_n, log_w, rank, &dd));
    NttDomain *loc; ZK_TRY(get_domain(ctx, log_n - log_w, &loc));
    const size_t m = (size_t)1 << (log_n - log_w);
    stage_begin(ctx, ST_NTT);
    Fr *v[3] = {a, b, c};
    for (int k = 0; k < 3; k++) ZK_TRY(dist_inverse(ctx, dd, loc, v[k], tmp, +1));
    for (int k = 0; k < 3; k++) ZK_TRY(dist_forward(ctx, dd, loc, v[k], tmp));
    ZK_LAUNCH(ctx, k_ab_minus_c, grid_for(m, 256), 256, 0, a, (const Fr *)b, (const Fr *)c, m);
    ZK_TRY(dist_inverse(ctx, dd, loc, a, tmp, -1));
    stage_end(ctx, ST_NTT);
    return ZKPOR_OK;
}

}  // namespace zk

using namespace zk;

extern "C" {

void zk_free_ntt(zkpor_ctx *ctx) {
    if (!ctx->ntt_tables) return;
    NttCache *cache = (NttCache *)ctx->ntt_tables;
    for (auto &kv : cache->doms) {
        NttDomain &d = kv.second;
        Fr *ptrs[] = {d.tw_fwd, d.tw_inv, d.cf_lo, d.cf_hi, d.cs_lo, d.cs_hi, d.ci_lo, d.ci_hi, d.ni_lo, d.ni_hi};
        for (Fr *p : ptrs) if (p) cudaFree(p);
    }
    delete cache;
    ctx->ntt_tables = nullptr;
    if (ctx->dist_tables) {
        DistCache *dc = (DistCache *)ctx->dist_tables;
        for (auto &kv : dc->doms) {
            DistDomain &d = kv.second;
            Fr *ptrs[] = {d.inv_lo, d.inv_hi, d.invd_lo, d.invd_hi, d.fwd_lo, d.fwd_hi, d.c5_lo, d.c5_hi, d.c5i_lo, d.c5i_hi, d.consts};
            for (Fr *p : ptrs) if (p) cudaFree(p);
        }
        delete dc;
        ctx->dist_tables = nullptr;
    }
}

int32_t zkpor_ntt(zkpor_ctx *ctx, void *data, uint32_t log_n, int32_t inverse, int32_t decimation, int32_t coset) {
    ZK_REQUIRE(ctx != nullptr && data != nullptr, "ntt: null argument");
    ZK_REQUIRE(log_n <= 28, "ntt: log_n exceeds the 2-adicity of Fr (28)");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    const size_t bytes = sizeof(Fr) << log_n;
    const bool on_dev = is_device_ptr(data);
    Fr *d = (Fr *)data;
    if (!on_dev) {
        ZK_TRY(ctx->ntt_a.reserve(bytes));
        d = ctx->ntt_a.as<Fr>();
        stage_begin(ctx, ST_H2D);
        ZK_CUDA(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        stage_end(ctx, ST_H2D);
    }
    stage_begin(ctx, ST_NTT);
    ZK_TRY(ntt_dev(ctx, d, log_n, inverse != 0, decimation != 0, coset != 0));
    stage_end(ctx, ST_NTT);
    if (!on_dev) {
        stage_begin(ctx, ST_D2H);
        ZK_CUDA(cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        stage_end(ctx, ST_D2H);
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    return ZKPOR_OK;
}

int32_t zkpor_fr_mul(zkpor_ctx *ctx, const void *a, const void *b, void *out, uint64_t n) {
    ZK_REQUIRE(ctx && a && b && out, "fr_mul: null argument");
    ZK_REQUIRE(is_device_ptr(a) && is_device_ptr(b) && is_device_ptr(out), "fr_mul: operands must be device memory");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n) ZK_LAUNCH(ctx, k_mul_vec, grid_for(n, 256), 256, 0, (const Fr *)a, (const Fr *)b, (Fr *)out, (size_t)n);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

int32_t zkpor_compute_h(zkpor_ctx *ctx, const void *a, const void *b, const void *c, uint64_t m, uint32_t log_n, void *out_h) {
    ZK_REQUIRE(ctx != nullptr && a && b && c && out_h, "compute_h: null argument");
    ZK_REQUIRE(log_n >= 1 && log_n <= 28 && m <= (1ull << log_n) && m > 0, "compute_h: size out of range");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    const size_t n = (size_t)1 << log_n, bytes = n * sizeof(Fr), in_bytes = m * sizeof(Fr);
    ZK_TRY(ctx->ntt_a.reserve(bytes)); ZK_TRY(ctx->ntt_b.reserve(bytes)); ZK_TRY(ctx->ntt_c.reserve(bytes));
    const void *src[3] = {a, b, c};
    Fr *dst[3] = {ctx->ntt_a.as<Fr>(), ctx->ntt_b.as<Fr>(), ctx->ntt_c.as<Fr>()};
    stage_begin(ctx, ST_H2D);
    for (int k = 0; k < 3; k++) {
        ZK_CUDA(cudaMemcpyAsync(dst[k], src[k], in_bytes, cudaMemcpyDefault, ctx->stream));
        if (bytes > in_bytes) ZK_CUDA(cudaMemsetAsync((uint8_t *)dst[k] + in_bytes, 0, bytes - in_bytes, ctx->stream));
    }
    stage_end(ctx, ST_H2D);
    ZK_TRY(compute_h_dev(ctx, dst[0], dst[1], dst[2], log_n));
    stage_begin(ctx, ST_D2H);
    ZK_CUDA(cudaMemcpyAsync(out_h, dst[0], bytes, cudaMemcpyDefault, ctx->stream));
    stage_end(ctx, ST_D2H);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    return ZKPOR_OK;
}

}  // extern "C"
