// BN254 multi-scalar multiplication on sm_100a: signed-digit Pippenger with a counting sort by bucket.
//
// Replaces gnark-crypto G1Jac.MultiExp / G2Jac.MultiExp (ecc/bn254/multiexp.go, out of tree) called from
// groth16.Prove -- src/prover/prover/prover.go:269.  Pipeline (all on ctx->stream, points/scalars resident in HBM):
//   1. k_from_mont      scalars Montgomery -> canonical                              (streaming, 64 B/term)
//   2. k_digits<HIST>   c-bit signed digits of every scalar, histogram per (window, bucket)   (32 B/term read)
//   3. k_scan           exclusive scan of the histogram per window
//   4. k_digits<SCATTER> signed point references scattered to their bucket segment   (32 B read + 4*nwin B write)
//   5. k_accumulate     one thread per (window, bucket): XYZZ += affine point, gathered 64/128 B loads
//   6. k_reduce_seg / k_sum_groups   sum_b b*B_b per window by segmented running sums
//   7. host             Horner over the nwin window sums (nwin*c doublings) -- O(1) work, 2 KB copied back
// The result of an MSM is a group element, so any bucket order gives bit-identical affine output.
#include "internal.h"

using namespace ff;
using namespace ec;

namespace zk {

MsmPlan msm_plan(uint64_t n) {
    // cost model in mixed-add units: every window adds n points and reduces nb buckets with 2 general adds (~1.4x)
    double best = 1e300; uint32_t best_c = 4;
    for (uint32_t c = 8; c <= 20; c++) {   // c >= 8 keeps nwin <= 32 (k_digits holds one key per window in registers)
        uint32_t nwin = (255 + c - 1) / c;
        double nb = (double)(1u << (c - 1));
        double cost = nwin * ((double)n + 2.8 * nb + 2000.0);
        if (cost < best) { best = cost; best_c = c; }
    }
    MsmPlan p; p.c = best_c; p.nwin = (255 + best_c - 1) / best_c; p.nb = 1u << (best_c - 1);
    return p;
}

// ------------------------------------------------------------------------------------------------ scalar side
__global__ void k_from_mont(const Fr *__restrict__ in, Fr *__restrict__ out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = Fr::from_mont(in[i]);
}

__device__ __forceinline__ uint32_t window_bits(const uint32_t *s, uint32_t off, uint32_t c) {
    uint32_t limb = off >> 5, sh = off & 31;
    uint64_t v = s[limb];
    if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

// MODE 0: histogram.  MODE 1: scatter (cursor[] starts as the exclusive scan and is advanced atomically).
template <int MODE>
__global__ void k_digits(const uint32_t *__restrict__ scalars /* plain, 8 x u32 each */, uint64_t n, MsmPlan plan,
                         uint32_t *__restrict__ counter, uint32_t *__restrict__ sorted) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + 8 * i);
    uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
    // All windows' atomics are issued before any result is consumed: a returning atomic costs microseconds under load
    // (ncu: the scatter kernel sat at 4.6 % issue utilisation waiting on them one at a time).
    const uint32_t MAXW = 32;
    uint32_t key[MAXW], pos[MAXW];
    uint32_t carry = 0;
#pragma unroll
    for (uint32_t w = 0; w < MAXW; w++) {
        key[w] = 0;
        if (w < plan.nwin) {
            uint32_t d = window_bits(s, w * plan.c, plan.c) + carry;
            uint32_t neg = 0;
            if (d > plan.nb) { d = (1u << plan.c) - d; neg = 1; carry = 1; } else carry = 0;
            key[w] = (d << 1) | neg;
        }
    }
#pragma unroll
    for (uint32_t w = 0; w < MAXW; w++) {
        pos[w] = 0;
        if (w < plan.nwin && (key[w] >> 1)) {
            size_t slot = (size_t)w * plan.nb + ((key[w] >> 1) - 1);
            if (MODE == 0) atomicAdd(&counter[slot], 1u);
            else pos[w] = atomicAdd(&counter[slot], 1u);
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (uint32_t w = 0; w < MAXW; w++)
            if (w < plan.nwin && (key[w] >> 1)) sorted[(size_t)w * n + pos[w]] = ((uint32_t)i << 1) | (key[w] & 1u);
    }
}

// one block per window: exclusive scan of cnt -> off, and cur = off
__global__ void k_scan(const uint32_t *__restrict__ cnt, uint32_t *__restrict__ off, uint32_t *__restrict__ cur, uint32_t nb) {
    __shared__ uint32_t part[1024];
    const uint32_t w = blockIdx.x, t = threadIdx.x, T = blockDim.x;
    const uint32_t per = (nb + T - 1) / T, lo = t * per, hi = min(lo + per, nb);
    const uint32_t *c = cnt + (size_t)w * nb;
    uint32_t sum = 0;
    for (uint32_t k = lo; k < hi; k++) sum += c[k];
    part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < T; d <<= 1) {   // Hillis-Steele inclusive scan over the per-thread sums
        uint32_t v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - sum;
    for (uint32_t k = lo; k < hi; k++) { off[(size_t)w * nb + k] = run; cur[(size_t)w * nb + k] = run; run += c[k]; }
}

// ------------------------------------------------------------------------------------------------ point side
template <class F> __device__ __forceinline__ Affine<F> load_affine(const Affine<F> *p) {
    Affine<F> r;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
    uint4 *dst = reinterpret_cast<uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Affine<F>) / 16); k++) dst[k] = __ldg(src + k);
    return r;
}

template <class F>
__global__ void __launch_bounds__(128) k_accumulate(const Affine<F> *__restrict__ points, const uint32_t *__restrict__ sorted,
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

// ---- bucket schedule ----------------------------------------------------------------------------------------------
// Threads of a warp run until the fullest of their 32 buckets is done (ncu, first version: 22 of 32 lanes active on
// average).  Buckets are therefore handed to threads in order of decreasing population: a counting sort of the
// (window, bucket) slots by their reference count, so that the 32 buckets of a warp have (almost) equal length.

__global__ void k_size_hist(const uint32_t *__restrict__ cnt, size_t slots, uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[SIZE_BINS];
    for (uint32_t i = threadIdx.x; i < SIZE_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < slots; t += (size_t)gridDim.x * blockDim.x) {
        uint32_t c = cnt[t];
        atomicAdd(&sh[c < SIZE_BINS ? c : SIZE_BINS - 1], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < SIZE_BINS; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// cursor[b] = number of slots in bins above b (descending order); one block of SIZE_BINS/2 threads, trivial size
__global__ void k_size_scan(const uint32_t *__restrict__ hist, uint32_t *__restrict__ cursor, uint32_t *__restrict__ bin_start) {
    __shared__ uint32_t sh[SIZE_BINS];
    for (uint32_t i = threadIdx.x; i < SIZE_BINS; i += blockDim.x) sh[i] = hist[SIZE_BINS - 1 - i];   // reversed
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = 0; for (uint32_t i = 0; i < SIZE_BINS; i++) { uint32_t v = sh[i]; sh[i] = run; run += v; } }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < SIZE_BINS; i += blockDim.x) { cursor[SIZE_BINS - 1 - i] = sh[i]; bin_start[SIZE_BINS - 1 - i] = sh[i]; }
}
__global__ void k_size_scatter(const uint32_t *__restrict__ cnt, size_t slots, uint32_t *__restrict__ cursor, uint32_t *__restrict__ order) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slots) return;
    uint32_t c = cnt[t];
    order[atomicAdd(&cursor[c < SIZE_BINS ? c : SIZE_BINS - 1], 1u)] = (uint32_t)t;
}

// ---- heavy buckets ----------------------------------------------------------------------------------------------
static const uint32_t HEAVY_CHUNK = 4096;

// one thread per (window, bucket): buckets above the threshold reserve ceil(cnt / HEAVY_CHUNK) block descriptors
__global__ void k_heavy_plan(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off, size_t slots, uint32_t heavy_t,
                             HeavyBlk *__restrict__ blks, HeavyBkt *__restrict__ bkts, uint32_t *__restrict__ counters) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slots) return;
    uint32_t c = cnt[t];
    if (c <= heavy_t) return;
    uint32_t nblk = (c + HEAVY_CHUNK - 1) / HEAVY_CHUNK;
    uint32_t first = atomicAdd(&counters[0], nblk), bi = atomicAdd(&counters[1], 1u);
    bkts[bi] = HeavyBkt{(uint32_t)t, first, nblk};
    for (uint32_t k = 0; k < nblk; k++) {
        uint32_t rem = c - k * HEAVY_CHUNK;
        blks[first + k] = HeavyBlk{(uint32_t)t, off[t] + k * HEAVY_CHUNK, rem < HEAVY_CHUNK ? rem : HEAVY_CHUNK};
    }
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

// thread (w, seg): sum_{j in seg} (j+1) * B[w][j]  via running sums, segment length L
template <class F>
__global__ void __launch_bounds__(128) k_reduce_seg(const XYZZ<F> *__restrict__ buckets, MsmPlan plan, uint32_t L, uint32_t segs,
                                                    XYZZ<F> *__restrict__ partials) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= plan.nwin * segs) return;
    uint32_t w = t / segs, seg = t % segs;
    uint32_t lo = seg * L, hi = min(lo + L, plan.nb);
    const XYZZ<F> *B = buckets + (size_t)w * plan.nb;
    XYZZ<F> run = XYZZ<F>::inf(), sum = XYZZ<F>::inf();
    for (uint32_t j = hi; j-- > lo;) { run.add(B[j]); sum.add(run); }
    // sum = sum_j (j - lo + 1) B_j ; add lo * run
    if (lo) sum.add(run.mul_u32(lo));
    partials[t] = sum;
}

// out[w*groups_out + g] = sum_{k < group} in[w*count_in + g*group + k]
template <class F>
__global__ void __launch_bounds__(128) k_sum_groups(const XYZZ<F> *__restrict__ in, uint32_t count_in, uint32_t group, uint32_t groups_out,
                                                    uint32_t nwin, XYZZ<F> *__restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nwin * groups_out) return;
    uint32_t w = t / groups_out, g = t % groups_out;
    uint32_t lo = g * group, hi = min(lo + group, count_in);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = lo; k < hi; k++) acc.add(in[(size_t)w * count_in + k]);
    out[t] = acc;
}

int32_t msm_sort(zkpor_ctx *ctx, const void *d_scalars, uint64_t n, uint32_t flags, MsmSorted *out) {
    ZK_REQUIRE(n > 0 && n < (1ull << 31), "msm: n must be in [1, 2^31)");
    MsmPlan plan = msm_plan(n);
    const size_t slots = (size_t)plan.nwin * plan.nb;
    ZK_TRY(ctx->bucket_cnt.reserve(slots * 4));
    ZK_TRY(ctx->bucket_off.reserve(slots * 4));
    ZK_TRY(ctx->bucket_cur.reserve(slots * 4));
    ZK_TRY(ctx->sort_idx.reserve((size_t)plan.nwin * n * 4));
    stage_begin(ctx, ST_DIGITS);
    const uint32_t *plain = (const uint32_t *)d_scalars;
    if (!(flags & ZKPOR_SCALARS_PLAIN)) {
        ZK_TRY(ctx->misc.reserve(n * 32));
        ZK_LAUNCH(ctx, k_from_mont, grid_for(n, 256), 256, 0, (const Fr *)d_scalars, ctx->misc.as<Fr>(), n);
        plain = ctx->misc.as<uint32_t>();
    }
    ZK_CUDA(cudaMemsetAsync(ctx->bucket_cnt.p, 0, slots * 4, ctx->stream));
    { KTimed kt(ctx, KC_SORT, n);
      ZK_LAUNCH(ctx, k_digits<0>, grid_for(n, 256), 256, 0, plain, n, plan, ctx->bucket_cnt.as<uint32_t>(), (uint32_t *)nullptr);
      kt.stop(); }
    stage_end(ctx, ST_DIGITS);
    stage_begin(ctx, ST_SORT);
    ZK_LAUNCH(ctx, k_scan, plan.nwin, 1024, 0, ctx->bucket_cnt.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(),
              ctx->bucket_cur.as<uint32_t>(), plan.nb);
    { KTimed kt(ctx, KC_SORT, n);
      ZK_LAUNCH(ctx, k_digits<1>, grid_for(n, 256), 256, 0, plain, n, plan, ctx->bucket_cur.as<uint32_t>(), ctx->sort_idx.as<uint32_t>());
      kt.stop(); }
    // heavy-bucket plan (device side, no host round trip): thresholds well above the uniform-case bucket size
    {
        const uint64_t avg = n / plan.nb + 1, total = (uint64_t)plan.nwin * n;
        // The top window of a 254-bit scalar has only 254 - c*(nwin-1) bits: its buckets are 2^(c-1) / 2^topbits times fuller than
        // the others (32x at c = 20: ~4096 references at n = 2^26) and still far too many for one CTA each, so the threshold sits
        // above them (2 * HEAVY_CHUNK - 2 = SIZE_BINS - 2, which also keeps every light bucket in an exact population bin).
        out->heavy_t = (uint32_t)(16 * avg > SIZE_BINS - 2 ? (16 * avg < 0xFFFFFFFFull ? 16 * avg : 0xFFFFFFFFull) : SIZE_BINS - 2);
        out->max_bkts = (uint32_t)(total / out->heavy_t + 1);
        out->max_blks = (uint32_t)(total / HEAVY_CHUNK + out->max_bkts);
        const size_t b_blk = (size_t)out->max_blks * sizeof(HeavyBlk), b_bkt = (size_t)out->max_bkts * sizeof(HeavyBkt);
        ZK_TRY(ctx->heavy.reserve(256 + b_blk + b_bkt));
        uint8_t *base = ctx->heavy.as<uint8_t>();
        ZK_CUDA(cudaMemsetAsync(base, 0, 8, ctx->stream));
        out->counters = (const uint32_t *)base; out->blks = (const HeavyBlk *)(base + 256); out->bkts = (const HeavyBkt *)(base + 256 + b_blk);
        ZK_LAUNCH(ctx, k_heavy_plan, grid_for(slots, 256), 256, 0, ctx->bucket_cnt.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(), slots, out->heavy_t,
                  (HeavyBlk *)out->blks, (HeavyBkt *)out->bkts, (uint32_t *)base);
    }
    // bucket schedule: slots in order of decreasing population
    {
        ZK_TRY(ctx->order.reserve(slots * 4 + 3 * SIZE_BINS * 4));
        uint32_t *order = ctx->order.as<uint32_t>(), *hist = order + slots, *cursor = hist + SIZE_BINS, *bin_start = cursor + SIZE_BINS;
        ZK_CUDA(cudaMemsetAsync(hist, 0, SIZE_BINS * 4, ctx->stream));
        ZK_LAUNCH(ctx, k_size_hist, 4 * ctx->sm_count, 256, 0, ctx->bucket_cnt.as<uint32_t>(), slots, hist);
        ZK_LAUNCH(ctx, k_size_scan, 1, 256, 0, (const uint32_t *)hist, cursor, bin_start);
        ZK_LAUNCH(ctx, k_size_scatter, grid_for(slots, 256), 256, 0, ctx->bucket_cnt.as<uint32_t>(), slots, cursor, order);
        out->order = order; out->hist = hist; out->bin_start = bin_start;
    }
    stage_end(ctx, ST_SORT);
    out->plan = plan; out->n = n;
    out->idx = ctx->sort_idx.as<uint32_t>(); out->off = ctx->bucket_off.as<uint32_t>(); out->cnt = ctx->bucket_cnt.as<uint32_t>();
    return ZKPOR_OK;
}

template <class F>
static int32_t msm_accumulate(zkpor_ctx *ctx, const void *d_points, const MsmSorted &s, XYZZ<F> *host_out) {
    const MsmPlan plan = s.plan;
    const size_t slots = (size_t)plan.nwin * plan.nb;
    ZK_TRY(ctx->buckets.reserve(slots * sizeof(XYZZ<F>)));
    stage_begin(ctx, ST_ACCUM);
    {
        KTimed kt(ctx, sizeof(F) == sizeof(Fp) ? KC_ACCUM_G1 : KC_ACCUM_G2, s.n);
        // light buckets: batched-affine tree rounds + XYZZ tail (msm_affine.cu) when the lists are long enough, else XYZZ only
        bool done = false;
        ZK_TRY(msm_tree_sums(ctx, (const Affine<F> *)d_points, s, ctx->buckets.as<XYZZ<F>>(), &done));
        if (!done)
            ZK_LAUNCH(ctx, (k_accumulate<F>), grid_for(slots, 128), 128, 0, (const Affine<F> *)d_points, s.idx, s.off, s.cnt, s.n, plan, s.heavy_t,
                      s.order, ctx->buckets.as<XYZZ<F>>());
        ZK_TRY(ctx->heavy_part.reserve((size_t)s.max_blks * sizeof(XYZZ<F>)));
        ZK_LAUNCH(ctx, (k_accumulate_heavy<F>), (s.max_blks < 8u * ctx->sm_count ? s.max_blks : 8u * ctx->sm_count), 128, 0, (const Affine<F> *)d_points, s.idx, s.blks, s.counters, s.n, plan,
                  ctx->heavy_part.as<XYZZ<F>>());
        ZK_LAUNCH(ctx, (k_heavy_combine<F>), (s.max_bkts < 4u * ctx->sm_count ? s.max_bkts : 4u * ctx->sm_count), 128, 0, (const XYZZ<F> *)ctx->heavy_part.p, s.bkts, s.counters,
                  ctx->buckets.as<XYZZ<F>>());
        kt.stop();
    }
    stage_end(ctx, ST_ACCUM);
    stage_begin(ctx, ST_REDUCE);
    const uint32_t L = plan.nb >= 32 ? 32 : plan.nb;
    uint32_t count = (plan.nb + L - 1) / L;
    ZK_TRY(ctx->partials.reserve((size_t)plan.nwin * count * sizeof(XYZZ<F>) * 2));
    XYZZ<F> *cur = ctx->partials.as<XYZZ<F>>(), *nxt = cur + (size_t)plan.nwin * count;
    ZK_LAUNCH(ctx, (k_reduce_seg<F>), grid_for((size_t)plan.nwin * count, 128), 128, 0, ctx->buckets.as<XYZZ<F>>(), plan, L, count, cur);
    while (count > 1) {
        uint32_t groups = (count + 31) / 32;
        ZK_LAUNCH(ctx, (k_sum_groups<F>), grid_for((size_t)plan.nwin * groups, 128), 128, 0, cur, count, 32u, groups, plan.nwin, nxt);
        XYZZ<F> *t = cur; cur = nxt; nxt = t; count = groups;
    }
    stage_end(ctx, ST_REDUCE);
    // window sums -> host, Horner
    std::vector<XYZZ<F>> ws(plan.nwin);
    stage_begin(ctx, ST_D2H);
    ZK_CUDA(cudaMemcpyAsync(ws.data(), cur, plan.nwin * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, ctx->stream));
    stage_end(ctx, ST_D2H);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    XYZZ<F> acc = ws[plan.nwin - 1];
    for (int w = (int)plan.nwin - 2; w >= 0; w--) {
        for (uint32_t k = 0; k < plan.c; k++) acc = acc.dbl();
        acc.add(ws[w]);
    }
    *host_out = acc;
    return ZKPOR_OK;
}

int32_t msm_accumulate_g1(zkpor_ctx *ctx, const void *d_points, const MsmSorted &s, G1XYZZ *o) { return msm_accumulate<Fp>(ctx, d_points, s, o); }
int32_t msm_accumulate_g2(zkpor_ctx *ctx, const void *d_points, const MsmSorted &s, G2XYZZ *o) { return msm_accumulate<Fp2>(ctx, d_points, s, o); }

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
