// BN254 multi-scalar multiplication on sm_100a: signed-digit Pippenger with a counting sort by bucket.
//
// Replaces gnark-crypto G1Jac.MultiExp / G2Jac.MultiExp (ecc/bn254/multiexp.go, out of tree) called from
// groth16.Prove -- src/prover/prover/prover.go:269.  Pipeline (all on ctx->stream, points/scalars resident in HBM):
// (steps 1-4, the scalar side, live in msm_sort.cu)
//   1. k_from_mont      scalars Montgomery -> canonical                              (streaming, 64 B/term)
//   2-4. counting sort of the signed point references (index << 1 | sign) by (window, bucket):
//        n <  2^18: k_digits<HIST> (RED histogram), k_scan, k_digits<SCATTER> (one returning L2 atomic per term and window)
//        n >= 2^18: k_part_count / k_part_scan / k_part_scatter / k_part_sort -- 64 partitions per window, counters and
//                   cursors in shared memory
//        then the heavy-bucket plan and the population schedule (slots by decreasing reference count); for a sort shared
//        by several multiplications (groth16.cu) msm_view derives each multiplication's lists from it
//   5. k_accumulate     one thread per (window, bucket): XYZZ += affine point, gathered 64/128 B loads
//   6. k_reduce_level   sum_b b*B_b per window by a tree of short running sums (8 buckets per thread and level)
//   7. host             Horner over the nwin window sums (nwin*c doublings) -- O(1) work, 2 KB copied back
// The result of an MSM is a group element, so any bucket order gives bit-identical affine output.
#include "internal.h"
#include <cstdlib>

using namespace ff;
using namespace ec;

namespace zk {

// ------------------------------------------------------------------------------------------------ point side
template <class F> __device__ __forceinline__ Affine<F> load_affine(const Affine<F> *p) {
    Affine<F> r;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
    uint4 *dst = reinterpret_cast<uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Affine<F>) / 16); k++) dst[k] = __ldg(src + k);
    return r;
}

#ifndef ZK_G2_ACC_MINB
#define ZK_G2_ACC_MINB 1
#endif
template <class F>
__global__ void __launch_bounds__(128, sizeof(F) == sizeof(Fp) ? 1 : ZK_G2_ACC_MINB) k_accumulate(const Affine<F> *__restrict__ points, const uint32_t *__restrict__ sorted,
                                                    const uint32_t *__restrict__ off, const uint32_t *__restrict__ cnt,
                                                    uint64_t n, MsmPlan plan, uint32_t heavy_t, const uint32_t *__restrict__ order,
                                                    XYZZ<F> *__restrict__ buckets) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)plan.nwin * plan.nb) return;
    const size_t t = order[tid];
    uint32_t w = (uint32_t)(t / plan.nb);
    const uint32_t *idx = sorted + (size_t)w * n + off[t];
    uint32_t m = cnt[t];
    if (m > heavy_t) return;   // reduced by k_accumulate_heavy / k_heavy_combine
    XYZZ<F> acc = XYZZ<F>::inf();
    if (sizeof(F) != sizeof(Fp)) {
        // G2: the accumulator alone is 64 registers; a prefetched point would spill, so only the reference is ahead
        uint32_t e = m ? __ldg(idx) : 0;
        for (uint32_t k = 0; k < m; k++) {
            uint32_t e1 = k + 1 < m ? __ldg(idx + k + 1) : 0;
            Affine<F> p = load_affine(points + (e >> 1));
            acc.add_affine(p, e & 1);
            e = e1;
        }
    } else if (m) {
        // two-deep software pipeline: the reference for k+2 and the point for k+1 are in flight while k is added
        uint32_t e = __ldg(idx), e1 = m > 1 ? __ldg(idx + 1) : 0;
        Affine<F> p = load_affine(points + (e >> 1));
        for (uint32_t k = 0; k < m; k++) {
            Affine<F> pn = p; uint32_t e2 = 0;
            if (k + 1 < m) pn = load_affine(points + (e1 >> 1));
            if (k + 2 < m) e2 = __ldg(idx + k + 2);
            acc.add_affine(p, e & 1);
            p = pn; e = e1; e1 = e2;
        }
    }
    buckets[t] = acc;
}

// one CTA per block descriptor: 128 strided partial sums, then a shared-memory tree
template <class F>
__global__ void __launch_bounds__(128) k_accumulate_heavy(const Affine<F> *__restrict__ points, const uint32_t *__restrict__ sorted,
                                                          const HeavyBlk *__restrict__ blks, const uint32_t *__restrict__ counters,
                                                          uint64_t n, MsmPlan plan, XYZZ<F> *__restrict__ parts) {
    __shared__ XYZZ<F> sh[128];
    const uint32_t nblk = counters[0];
    // persistent CTAs striding over the block descriptors: with no heavy bucket (uniform scalars) the launch is a few
    // hundred CTAs that exit at once (a grid of max_blks ~ 4*10^5 empty CTAs cost 11 ms per MSM in the first version)
    for (uint32_t b = blockIdx.x; b < nblk; b += gridDim.x) {
        const HeavyBlk blk = blks[b];
        const uint32_t *idx = sorted + (size_t)(blk.slot / plan.nb) * n + blk.start;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t k = threadIdx.x; k < blk.count; k += 128) {
            uint32_t e = __ldg(idx + k);
            if (e == REF_SKIP) continue;   // shared sort: entry not in this multiplication (never produced otherwise: index < 2^31)
            Affine<F> p = load_affine(points + (e >> 1));
            acc.add_affine(p, e & 1);
        }
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (uint32_t s = 64; s > 0; s >>= 1) {
            if (threadIdx.x < s) { XYZZ<F> a = sh[threadIdx.x]; a.add(sh[threadIdx.x + s]); sh[threadIdx.x] = a; }
            __syncthreads();
        }
        if (threadIdx.x == 0) parts[b] = sh[0];
        __syncthreads();
    }
}

// one CTA per heavy bucket: strided sums of its block partials, then a shared-memory tree
template <class F>
__global__ void __launch_bounds__(128) k_heavy_combine(const XYZZ<F> *__restrict__ parts, const HeavyBkt *__restrict__ bkts,
                                                       const uint32_t *__restrict__ counters, XYZZ<F> *__restrict__ buckets) {
    __shared__ XYZZ<F> sh[128];
    const uint32_t nbkt = counters[1];
    for (uint32_t i = blockIdx.x; i < nbkt; i += gridDim.x) {
        const HeavyBkt bk = bkts[i];
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t k = threadIdx.x; k < bk.nblk; k += 128) acc.add(parts[bk.first_blk + k]);
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (uint32_t s = 64; s > 0; s >>= 1) {
            if (threadIdx.x < s) { XYZZ<F> a = sh[threadIdx.x]; a.add(sh[threadIdx.x + s]); sh[threadIdx.x] = a; }
            __syncthreads();
        }
        if (threadIdx.x == 0) buckets[bk.slot] = sh[0];
        __syncthreads();
    }
}

// Bucket reduction W = sum_j (j+1) X_j of every window, as a tree of short running sums (no scalar multiplications):
// split X into segments of L: (i+1) = L g + (t+1), so W(X) = sum_g S_g + L (W(R) - T) with S_g = sum_t (t+1) X_{Lg+t},
// R_g = sum_t X_{Lg+t}, T = the total.  Recursing on R until one element (= T) is left gives
//     W = sum_l L^l U_l  -  T (L + L^2 + ... + L^(K-1)),      U_l = sum_g S^l_g,   K levels.
// Level l: thread (w, g) walks its <= L elements once from the top -- run += X, sum += run (2 additions per element) -- and
// also sums the carried array A^l (the descendants' sum_{j<l} L^j S^j), writing R_g and A^{l+1}_g = sum A^l + L^l S^l_g.  After
// the last level A = sum_l L^l U_l and R = T; the host subtracts the multiple of T.  Chains are L additions long (the
// previous scheme: 64 + a 19-bit scalar multiplication per thread, a quarter as many threads).
template <class F>
__global__ void __launch_bounds__(128) k_reduce_level(const XYZZ<F> *__restrict__ X, const XYZZ<F> *__restrict__ A, uint32_t m_in, uint32_t m_out,
                                                      uint32_t L, uint32_t dbl, uint32_t nwin, XYZZ<F> *__restrict__ R, XYZZ<F> *__restrict__ An) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nwin * m_out) return;
    const uint32_t w = t / m_out, g = t - w * m_out;
    const uint32_t lo = g * L, hi = min(lo + L, m_in);
    const XYZZ<F> *x = X + (size_t)w * m_in;
    XYZZ<F> run = XYZZ<F>::inf(), sum = XYZZ<F>::inf();
    for (uint32_t j = hi; j-- > lo;) { run.add(x[j]); sum.add(run); }
    for (uint32_t k = 0; k < dbl; k++) sum = sum.dbl();       // L^l S
    if (A != nullptr) {
        const XYZZ<F> *a = A + (size_t)w * m_in;
        for (uint32_t j = lo; j < hi; j++) sum.add(a[j]);
    }
    R[t] = run; An[t] = sum;
}

template <class F>
static int32_t msm_accumulate(zkpor_ctx *ctx, const void *d_points, const MsmSorted &s, XYZZ<F> *host_out, uint64_t terms = 0) {
    const MsmPlan plan = s.plan;
    const size_t slots = (size_t)plan.nwin * plan.nb;
    ZK_TRY(ctx->buckets.reserve(slots * sizeof(XYZZ<F>)));
    stage_begin(ctx, ST_ACCUM);
    {
        KTimed kt(ctx, sizeof(F) == sizeof(Fp) ? KC_ACCUM_G1 : KC_ACCUM_G2, terms ? terms : s.n);   // units = points actually added
        // light buckets: batched-affine tree rounds + XYZZ tail (msm_affine.cu) when the lists are long enough, else XYZZ only
        bool done = false;
        if (!s.is_view) ZK_TRY(msm_tree_sums(ctx, (const Affine<F> *)d_points, s, ctx->buckets.as<XYZZ<F>>(), &done));
        if (!done) {
            ZK_LAUNCH(ctx, (k_accumulate<F>), grid_for(slots, 128), 128, 0, (const Affine<F> *)d_points, s.idx, s.off, s.cnt, s.n, plan, s.heavy_t,
                      s.order, ctx->buckets.as<XYZZ<F>>());
        }
        ZK_TRY(ctx->heavy_part.reserve((size_t)s.max_blks * sizeof(XYZZ<F>)));
        ZK_LAUNCH(ctx, (k_accumulate_heavy<F>), (s.max_blks < 8u * ctx->sm_count ? s.max_blks : 8u * ctx->sm_count), 128, 0, (const Affine<F> *)d_points, s.idx, s.blks, s.counters, s.n, plan,
                  ctx->heavy_part.as<XYZZ<F>>());
        ZK_LAUNCH(ctx, (k_heavy_combine<F>), (s.max_bkts < 4u * ctx->sm_count ? s.max_bkts : 4u * ctx->sm_count), 128, 0, (const XYZZ<F> *)ctx->heavy_part.p, s.bkts, s.counters,
                  ctx->buckets.as<XYZZ<F>>());
        kt.stop();
    }
    stage_end(ctx, ST_ACCUM);
    stage_begin(ctx, ST_REDUCE);
    const uint32_t LOG_L = 3, L = 1u << LOG_L;
    // level sizes nb -> ceil(nb/L) -> ... -> 1; R^l and A^{l+1} of every level live side by side in `partials`
    uint32_t sizes[40], K = 0;
    sizes[0] = plan.nb;
    while (sizes[K] > 1) { sizes[K + 1] = (sizes[K] + L - 1) / L; K++; }
    size_t total_out = 0;
    for (uint32_t l = 1; l <= K; l++) total_out += sizes[l];
    ZK_TRY(ctx->partials.reserve((total_out ? total_out : 1) * plan.nwin * sizeof(XYZZ<F>) * 2));
    const XYZZ<F> *X = ctx->buckets.as<XYZZ<F>>(), *A = nullptr;
    XYZZ<F> *out = ctx->partials.as<XYZZ<F>>();
    for (uint32_t l = 0; l < K; l++) {
        const size_t cnt_out = (size_t)plan.nwin * sizes[l + 1];
        XYZZ<F> *R = out, *An = out + cnt_out;
        ZK_LAUNCH(ctx, (k_reduce_level<F>), grid_for(cnt_out, 128), 128, 0, X, A, sizes[l], sizes[l + 1], L, l * LOG_L, plan.nwin, R, An);
        X = R; A = An; out += 2 * cnt_out;
    }
    stage_end(ctx, ST_REDUCE);
    // per window (T, sum_l L^l U_l) -> host; W = that sum - T (L + ... + L^(K-1)); Horner over the windows
    std::vector<XYZZ<F>> ws(2 * (size_t)plan.nwin);
    stage_begin(ctx, ST_D2H);
    if (K == 0) {   // one bucket per window: W = X
        ZK_CUDA(cudaMemcpyAsync(ws.data(), X, plan.nwin * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        ZK_CUDA(cudaMemcpyAsync(ws.data(), X, plan.nwin * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, ctx->stream));                 // T per window
        ZK_CUDA(cudaMemcpyAsync(ws.data() + plan.nwin, A, plan.nwin * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, ctx->stream));    // sum_l L^l U_l
    }
    stage_end(ctx, ST_D2H);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t cmul = 0;
    for (uint32_t l = 1; l < K; l++) cmul += 1u << (l * LOG_L);
    auto window_sum = [&](uint32_t w) {
        if (K == 0) return ws[w];
        XYZZ<F> r = ws[plan.nwin + w];
        if (cmul) r.add(ws[w].mul_u32(cmul).negated());
        return r;
    };
    XYZZ<F> acc = window_sum(plan.nwin - 1);
    for (int w = (int)plan.nwin - 2; w >= 0; w--) {
        for (uint32_t k = 0; k < plan.c; k++) acc = acc.dbl();
        acc.add(window_sum((uint32_t)w));
    }
    *host_out = acc;
    return ZKPOR_OK;
}

int32_t msm_accumulate_g1(zkpor_ctx *ctx, const void *d_points, const MsmSorted &s, G1XYZZ *o, uint64_t terms) { return msm_accumulate<Fp>(ctx, d_points, s, o, terms); }
int32_t msm_accumulate_g2(zkpor_ctx *ctx, const void *d_points, const MsmSorted &s, G2XYZZ *o, uint64_t terms) { return msm_accumulate<Fp2>(ctx, d_points, s, o, terms); }

int32_t msm_g1_dev(zkpor_ctx *ctx, const void *d_points, const void *d_scalars, uint64_t n, uint32_t flags, G1XYZZ *o) {
    if (n == 0) { *o = G1XYZZ::inf(); return ZKPOR_OK; }
    MsmSorted s; ZK_TRY(msm_sort(ctx, d_scalars, n, flags, &s));
    return msm_accumulate<Fp>(ctx, d_points, s, o);
}
int32_t msm_g2_dev(zkpor_ctx *ctx, const void *d_points, const void *d_scalars, uint64_t n, uint32_t flags, G2XYZZ *o) {
    if (n == 0) { *o = G2XYZZ::inf(); return ZKPOR_OK; }
    MsmSorted s; ZK_TRY(msm_sort(ctx, d_scalars, n, flags, &s));
    return msm_accumulate<Fp2>(ctx, d_points, s, o);
}

}  // namespace zk

// ------------------------------------------------------------------------------------------------ C-ABI
using namespace zk;

template <class F>
static int32_t msm_entry(zkpor_ctx *ctx, const void *points, const void *scalars, uint64_t n, uint32_t flags, XYZZ<F> *out) {
    ZK_REQUIRE(ctx != nullptr, "msm: null context");
    ZK_REQUIRE(n == 0 || (points != nullptr && scalars != nullptr), "msm: null input");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    if (n == 0) { *out = XYZZ<F>::inf(); return ZKPOR_OK; }
    const void *dp, *ds;
    stage_begin(ctx, ST_H2D);
    ZK_TRY(to_device(ctx, points, n * sizeof(Affine<F>), ctx->in_points, &dp));
    ZK_TRY(to_device(ctx, scalars, n * 32, ctx->in_scalars, &ds));
    stage_end(ctx, ST_H2D);
    MsmSorted s;
    ZK_TRY(msm_sort(ctx, ds, n, flags, &s));
    ZK_TRY(msm_accumulate<F>(ctx, dp, s, out));
    stages_collect(ctx);
    return ZKPOR_OK;
}

extern "C" {

int32_t zkpor_msm_g1(zkpor_ctx *ctx, const void *points, const void *scalars, uint64_t n, uint32_t flags, void *out_affine64) {
    ZK_REQUIRE(out_affine64 != nullptr, "msm: null output");
    G1XYZZ r; ZK_TRY(msm_entry<Fp>(ctx, points, scalars, n, flags, &r));
    G1Affine a = r.to_affine(); memcpy(out_affine64, &a, sizeof a);
    return ZKPOR_OK;
}
int32_t zkpor_msm_g2(zkpor_ctx *ctx, const void *points, const void *scalars, uint64_t n, uint32_t flags, void *out_affine128) {
    ZK_REQUIRE(out_affine128 != nullptr, "msm: null output");
    G2XYZZ r; ZK_TRY(msm_entry<Fp2>(ctx, points, scalars, n, flags, &r));
    G2Affine a = r.to_affine(); memcpy(out_affine128, &a, sizeof a);
    return ZKPOR_OK;
}
int32_t zkpor_msm_g1_partial(zkpor_ctx *ctx, const void *points, const void *scalars, uint64_t n, uint32_t flags, void *out) {
    ZK_REQUIRE(out != nullptr, "msm: null output");
    G1XYZZ r; ZK_TRY(msm_entry<Fp>(ctx, points, scalars, n, flags, &r)); memcpy(out, &r, sizeof r);
    return ZKPOR_OK;
}
int32_t zkpor_msm_g2_partial(zkpor_ctx *ctx, const void *points, const void *scalars, uint64_t n, uint32_t flags, void *out) {
    ZK_REQUIRE(out != nullptr, "msm: null output");
    G2XYZZ r; ZK_TRY(msm_entry<Fp2>(ctx, points, scalars, n, flags, &r)); memcpy(out, &r, sizeof r);
    return ZKPOR_OK;
}
int32_t zkpor_g1_sum_partials(const void *partials, uint32_t k, void *out_affine64) {
    ZK_REQUIRE(partials != nullptr && out_affine64 != nullptr, "sum_partials: null argument");
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t i = 0; i < k; i++) { G1XYZZ p; memcpy(&p, (const uint8_t *)partials + (size_t)i * sizeof p, sizeof p); acc.add(p); }
    G1Affine a = acc.to_affine(); memcpy(out_affine64, &a, sizeof a);
    return ZKPOR_OK;
}
int32_t zkpor_g2_sum_partials(const void *partials, uint32_t k, void *out_affine128) {
    ZK_REQUIRE(partials != nullptr && out_affine128 != nullptr, "sum_partials: null argument");
    G2XYZZ acc = G2XYZZ::inf();
    for (uint32_t i = 0; i < k; i++) { G2XYZZ p; memcpy(&p, (const uint8_t *)partials + (size_t)i * sizeof p, sizeof p); acc.add(p); }
    G2Affine a = acc.to_affine(); memcpy(out_affine128, &a, sizeof a);
    return ZKPOR_OK;
}

}  // extern "C"
