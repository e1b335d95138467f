// Batched-affine bucket accumulation for the Pippenger MSM of msm.cu (same reference seam: gnark-crypto G1Jac.MultiExp /
// G2Jac.MultiExp inside groth16.Prove, src/prover/prover/prover.go:269).
#include "internal.h"

using namespace ff;
using namespace ec;

namespace zk {

template <class F> __device__ __forceinline__ Affine<F> load_affine(const Affine<F> *p) {
    Affine<F> r;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
    uint4 *dst = reinterpret_cast<uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Affine<F>) / 16); k++) dst[k] = __ldg(src + k);
    return r;
}

// ---- batched-affine bucket accumulation ---------------------------------------------------------------------------
// An XYZZ mixed addition costs 10 field products (G2: 28 base-field products); an affine addition costs one inversion
// + 3.  With the inversion shared by a batch of K independent additions (Montgomery's trick: 3 products per element)
// and computed by divsteps (ff.cuh Fe::inv: the pipe time of ~15 products instead of ~320 for Fermat) an addition costs
// ~6.5 products (G2: ~18).  Additions of a batch must be independent, so a bucket is summed as a binary tree: round r
// adds neighbours 2j, 2j+1 of every bucket's list -- round 0 gathers key points through the sorted references, later
// rounds stream the previous round's sums -- and writes ceil(m/2) points.  After R rounds the few points left per bucket
// are summed in XYZZ form (k_accumulate_pts) and the bucket reduction continues unchanged.
// Layout of round r >= 1: window-major with stride cap_r = (cap_{r-1} + nb)/2 + 1, bucket b of a window starting at
// off_r = (off_{r-1} + b) >> 1: the recurrence keeps the lists disjoint (gap >= ceil(m/2)) without another prefix sum.
// A thread walks G consecutive entries of the population-ordered slot list as one stream of pairs, cut into batches
// of K; prefix products and the pair positions of a batch live in local memory (L1/L2 resident).
template <class F> __device__ __forceinline__ void store_affine(Affine<F> *p, const Affine<F> &v) {
    uint4 *dst = reinterpret_cast<uint4 *>(p);
    const uint4 *src = reinterpret_cast<const uint4 *>(&v);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Affine<F>) / 16); k++) dst[k] = src[k];
}
template <class F> __device__ __forceinline__ F load_field(const F *p) {
    F r;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
    uint4 *dst = reinterpret_cast<uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) dst[k] = __ldg(src + k);
    return r;
}
// element `pos` of a round's input: round 0 = key point through its signed reference, later = previous round's sum
template <class F, bool FIRST>
__device__ __forceinline__ Affine<F> ba_load(const Affine<F> *__restrict__ points, const uint32_t *__restrict__ sorted,
                                             const Affine<F> *__restrict__ qin, uint64_t pos) {
    if (FIRST) {
        uint32_t e = __ldg(sorted + pos);
        Affine<F> p = load_affine(points + (e >> 1));
        if (e & 1) p.y = F::neg(p.y);
        return p;
    }
    return load_affine(qin + pos);
}
template <class F, bool FIRST>
__device__ __forceinline__ F ba_load_x(const Affine<F> *__restrict__ points, const uint32_t *__restrict__ sorted,
                                       const Affine<F> *__restrict__ qin, uint64_t pos) {
    if (FIRST) return load_field(&points[__ldg(sorted + pos) >> 1].x);
    return load_field(&qin[pos].x);
}
// the value inverted for a + b: x_b - x_a, 2 y_a for a doubling, 1 when no division is needed (infinity in or out)
template <class F> __device__ __forceinline__ F ba_denominator(const Affine<F> &a, const Affine<F> &b) {
    if (a.is_inf() || b.is_inf()) return F::one();
    F d = F::sub(b.x, a.x);
    if (!d.is_zero()) return d;
    if (a.y == b.y && !a.y.is_zero()) return F::dbl(a.y);
    return F::one();
}
template <class F> __device__ __forceinline__ Affine<F> ba_sum(const Affine<F> &a, const Affine<F> &b, const F &dinv) {
    if (a.is_inf()) return b;
    if (b.is_inf()) return a;
    F lam;
    if (a.x == b.x) {
        if (!(a.y == b.y) || a.y.is_zero()) return Affine<F>::inf();
        F xx = F::sqr(a.x);
        lam = F::mul(F::add(F::dbl(xx), xx), dinv);
    } else {
        lam = F::mul(F::sub(b.y, a.y), dinv);
    }
    Affine<F> r;
    r.x = F::sub(F::sub(F::sqr(lam), a.x), b.x);
    r.y = F::sub(F::mul(lam, F::sub(a.x, r.x)), a.y);
    return r;
}

// pairs of round r in a bucket list that starts with c references
__host__ __device__ __forceinline__ uint32_t pairs_in_round(uint32_t c, uint32_t r) { return (uint32_t)((((uint64_t)c + (1ull << r) - 1) >> r) >> 1); }

// Balanced schedule.  Light slots sit in `order` by decreasing population and every population bin c is exact, so the
// number of pairs in front of a bin is a 8192-entry prefix sum per round: pp[r][d], d = SIZE_BINS-1-c ascending,
// pp[r][SIZE_BINS] = all pairs of the round.  Thread g of a round takes pairs [g*W, (g+1)*W) of that flat index space
// (binary search for the bin, a division for the slot) -- the same work for every thread whatever the populations are
// (the top window of a 254-bit scalar fills 32x fewer, 32x longer lists than the others).
__global__ void __launch_bounds__(1024) k_pair_prefix(const uint32_t *__restrict__ hist, uint32_t heavy_t, uint64_t *__restrict__ pp) {
    __shared__ uint64_t part[1024];
    const uint32_t r = blockIdx.x, t = threadIdx.x;
    constexpr uint32_t PER = SIZE_BINS / 1024;
    uint64_t loc[PER], sum = 0;
#pragma unroll
    for (uint32_t i = 0; i < PER; i++) {
        const uint32_t c = SIZE_BINS - 1 - (t * PER + i);
        loc[i] = sum;
        if (c <= heavy_t) sum += (uint64_t)hist[c] * pairs_in_round(c, r);
    }
    part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint64_t v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const uint64_t base = part[t] - sum;
    uint64_t *out = pp + (size_t)r * (SIZE_BINS + 1);
#pragma unroll
    for (uint32_t i = 0; i < PER; i++) out[t * PER + i] = base + loc[i];
    if (t == 1023) out[SIZE_BINS] = base + sum;
}

// Position of a thread in the flat pair space of a round: population bin -> slot of the bin -> pair of the slot.
struct PairWalker {
    const uint32_t *order, *off, *hist, *bin_start;
    MsmPlan plan; uint32_t round; uint64_t cap_in, cap_out;
    uint32_t c, p, hc, bs, rank, j;     // bin population, pairs per slot, slots in bin, bin start in `order`, slot, pair
    uint32_t ibase, obase; bool need_slot;
    // input position of the current pair's first element and output position of its sum; then step to the next pair
    __device__ __forceinline__ void next(uint32_t &src, uint32_t &dst) {
        if (need_slot) {
            const uint32_t t = order[bs + rank];
            const uint32_t w = t / plan.nb, b = t - w * plan.nb;
            uint32_t o = off[t];
            for (uint32_t r = 0; r < round; r++) o = (o + b) >> 1;
            ibase = (uint32_t)((uint64_t)w * cap_in + o);
            obase = (uint32_t)((uint64_t)w * cap_out + ((o + b) >> 1));
            need_slot = false;
        }
        src = ibase + 2 * j; dst = obase + j;
        if (++j == p) {
            j = 0; need_slot = true;
            if (++rank == hc) {
                rank = 0;
                do { c--; p = pairs_in_round(c, round); hc = hist[c]; } while (c > 1 && (hc == 0 || p == 0));
                bs = bin_start[c];
            }
        }
    }
};

// One round of the tree.  Per batch of K pairs: forward pass (x-coordinates only: denominators and their running product),
// one inversion, backward pass (full points: slopes and sums).  Loads run one pair ahead of the arithmetic in both passes.
template <class F, int K, bool FIRST, int MINB>
__global__ void __launch_bounds__(128, MINB) k_affine_round(const Affine<F> *__restrict__ points, const uint32_t *__restrict__ sorted,
                                                            const Affine<F> *__restrict__ qin, Affine<F> *__restrict__ qout,
                                                            const uint32_t *__restrict__ off, const uint32_t *__restrict__ order,
                                                            const uint32_t *__restrict__ hist, const uint32_t *__restrict__ bin_start,
                                                            const uint64_t *__restrict__ pp, MsmPlan plan, uint32_t round,
                                                            uint64_t cap_in, uint64_t cap_out, uint32_t W) {
    const uint64_t *P = pp + (size_t)round * (SIZE_BINS + 1);
    const uint64_t total = P[SIZE_BINS];
    const uint64_t first = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * W;
    if (first >= total) return;
    uint32_t left = (uint32_t)(total - first < W ? total - first : W);   // pairs of this thread
    // bin of pair `first`: P[d] <= first < P[d+1]
    uint32_t lo = 0, hi = SIZE_BINS;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (P[mid] <= first) lo = mid; else hi = mid; }
    PairWalker wk;
    wk.order = order; wk.off = off; wk.hist = hist; wk.bin_start = bin_start; wk.plan = plan; wk.round = round; wk.cap_in = cap_in; wk.cap_out = cap_out;
    wk.c = SIZE_BINS - 1 - lo;
    wk.p = pairs_in_round(wk.c, round);
    wk.hc = hist[wk.c]; wk.bs = bin_start[wk.c];
    wk.rank = (uint32_t)((first - P[lo]) / wk.p); wk.j = (uint32_t)((first - P[lo]) % wk.p);
    wk.ibase = wk.obase = 0; wk.need_slot = true;

    F pre[K];
    uint32_t src[K], dst[K];
    while (left) {
        const int kb = left < (uint32_t)K ? (int)left : K;
        // ---- forward: denominators and prefix products
        F acc = F::one();
        uint32_t ps, pd;
        wk.next(ps, pd);
        F nx1 = ba_load_x<F, FIRST>(points, sorted, qin, ps), nx2 = ba_load_x<F, FIRST>(points, sorted, qin, (uint64_t)ps + 1);
        for (int k = 0; k < kb; k++) {
            const uint32_t pos = ps;
            const F x1 = nx1, x2 = nx2;
            src[k] = ps; dst[k] = pd;
            if (k + 1 < kb) {
                wk.next(ps, pd);
                nx1 = ba_load_x<F, FIRST>(points, sorted, qin, ps); nx2 = ba_load_x<F, FIRST>(points, sorted, qin, (uint64_t)ps + 1);
            }
            F d = F::sub(x2, x1);
            if (d.is_zero() || x1.is_zero() || x2.is_zero())   // doubling, cancellation or (possibly) infinity: classify on the full points
                d = ba_denominator(ba_load<F, FIRST>(points, sorted, qin, pos), ba_load<F, FIRST>(points, sorted, qin, (uint64_t)pos + 1));
            pre[k] = acc;
            acc = k ? F::mul(acc, d) : d;
        }
        left -= (uint32_t)kb;
        F inv = F::inv(acc);
        // ---- backward: inverse of every denominator, slope, sum
        Affine<F> na = ba_load<F, FIRST>(points, sorted, qin, src[kb - 1]), nb = ba_load<F, FIRST>(points, sorted, qin, (uint64_t)src[kb - 1] + 1);
        for (int kk = kb - 1; kk >= 0; kk--) {
            const Affine<F> a = na, b = nb;
            if (kk) { na = ba_load<F, FIRST>(points, sorted, qin, src[kk - 1]); nb = ba_load<F, FIRST>(points, sorted, qin, (uint64_t)src[kk - 1] + 1); }
            F dinv = inv;
            if (kk) { dinv = F::mul(inv, pre[kk]); inv = F::mul(inv, ba_denominator(a, b)); }
            store_affine(qout + dst[kk], ba_sum(a, b, dinv));
        }
    }
}

// the odd element of a list has no partner in this round: it is carried to the next round as it is
template <class F, bool FIRST>
__global__ void __launch_bounds__(256) k_affine_carry(const Affine<F> *__restrict__ points, const uint32_t *__restrict__ sorted,
                                                      const Affine<F> *__restrict__ qin, Affine<F> *__restrict__ qout,
                                                      const uint32_t *__restrict__ off, const uint32_t *__restrict__ cnt,
                                                      const uint32_t *__restrict__ order, MsmPlan plan, uint32_t heavy_t, uint32_t round,
                                                      uint64_t cap_in, uint64_t cap_out, size_t slots) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const uint32_t t = order[i];
    const uint32_t c0 = cnt[t];
    if (c0 == 0 || c0 > heavy_t) return;
    const uint32_t m = (uint32_t)(((uint64_t)c0 + (1ull << round) - 1) >> round);
    if (!(m & 1)) return;
    const uint32_t w = t / plan.nb, b = t - w * plan.nb;
    uint32_t o = off[t];
    for (uint32_t r = 0; r < round; r++) o = (o + b) >> 1;
    store_affine(qout + (uint64_t)w * cap_out + ((o + b) >> 1) + (m >> 1), ba_load<F, FIRST>(points, sorted, qin, (uint64_t)w * cap_in + o + (m - 1)));
}

// after the affine rounds: thread (window, bucket) sums the ceil(cnt / 2^rounds) points left of its list in XYZZ form
template <class F>
__global__ void __launch_bounds__(128) k_accumulate_pts(const Affine<F> *__restrict__ q, const uint32_t *__restrict__ off,
                                                        const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ order, MsmPlan plan,
                                                        uint32_t heavy_t, uint32_t rounds, uint64_t cap, XYZZ<F> *__restrict__ buckets) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)plan.nwin * plan.nb) return;
    const uint32_t t = order[tid];
    const uint32_t c0 = cnt[t];
    if (c0 > heavy_t) return;
    const uint32_t w = t / plan.nb, b = t - w * plan.nb;
    uint32_t o = off[t];
    for (uint32_t r = 0; r < rounds; r++) o = (o + b) >> 1;
    const uint32_t m = (uint32_t)(((uint64_t)c0 + (1ull << rounds) - 1) >> rounds);
    const Affine<F> *src = q + (uint64_t)w * cap + o;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = 0; k < m; k++) acc.add_affine(load_affine(src + k), false);
    buckets[t] = acc;
}

template <class F>
static int32_t tree_sums(zkpor_ctx *ctx, const Affine<F> *d_points, const MsmSorted &s, XYZZ<F> *buckets, bool *done) {
    const MsmPlan plan = s.plan;
    const size_t slots = (size_t)plan.nwin * plan.nb;
    *done = false;
    // R leaves ~2-4 points per bucket for the XYZZ tail; 0 = XYZZ only
    constexpr int K = sizeof(F) == sizeof(Fp) ? 64 : 32;   // pairs per inversion (prefix products: 2 KB of local memory per thread)
    const uint64_t avg = s.n / plan.nb;
    uint32_t R = 0;
    while (R < 16 && (avg >> (R + 1)) >= 2) R++;
    // Measured on B200 (profiles/r01_SUMMARY.md): the affine rounds run at 0.9-1.0x the time of the XYZZ accumulation -- the
    // divstep inversions are long dependent chains that leave the IMAD pipe half idle -- so they are opt-in (affine_rounds > 0
    // caps the depth, < 0 selects the automatic depth) and the default is the XYZZ path.
    if (ctx->affine_rounds == 0) R = 0;
    else if (ctx->affine_rounds > 0 && (uint32_t)ctx->affine_rounds < R) R = (uint32_t)ctx->affine_rounds;
    if (s.heavy_t > SIZE_BINS - 2) R = 0;   // the balanced schedule needs an exact population bin for every light bucket
    if ((uint64_t)plan.nwin * s.n >= 0xFFFFFFFFull) R = 0;   // batch positions are kept as 32-bit offsets
    if (R == 0) return ZKPOR_OK;
    uint64_t cap[18];
    cap[0] = s.n;
    for (uint32_t r = 0; r < R; r++) cap[r + 1] = (cap[r] + plan.nb) / 2 + 1;
    if (ctx->tree_a.reserve((size_t)plan.nwin * cap[1] * sizeof(Affine<F>)) != ZKPOR_OK ||
        (R >= 2 && ctx->tree_b.reserve((size_t)plan.nwin * cap[2] * sizeof(Affine<F>)) != ZKPOR_OK)) {
        ctx->tree_a.release(); ctx->tree_b.release();   // not enough HBM for the intermediate sums: plain XYZZ accumulation
        return ZKPOR_OK;
    }
    // pair prefix per round over the population bins (balanced schedule)
    ZK_TRY(ctx->tree_meta.reserve((size_t)R * (SIZE_BINS + 1) * sizeof(uint64_t)));
    uint64_t *pp = ctx->tree_meta.as<uint64_t>();
    ZK_LAUNCH(ctx, k_pair_prefix, R, 1024, 0, s.hist, s.heavy_t, pp);
    // G2: 170 registers (8 warps/SM) without a bound; capped at 128 (16 warps/SM) it spills ~0.4 KB to L1
    const bool tight = sizeof(F) == sizeof(Fp) || ctx->g2_tight_regs;   // G1 fits 128 registers without spilling
    const uint32_t W = 2 * K;   // pairs per thread: whole batches
    for (uint32_t r = 0; r < R; r++) {
        const Affine<F> *qin = r == 0 ? nullptr : ((r & 1) ? ctx->tree_a.as<Affine<F>>() : ctx->tree_b.as<Affine<F>>());
        Affine<F> *qout = (r & 1) ? ctx->tree_b.as<Affine<F>>() : ctx->tree_a.as<Affine<F>>();
        const uint64_t max_pairs = ((uint64_t)plan.nwin * s.n) >> (r + 1);
        const size_t threads = (size_t)(max_pairs / W + 1);
#define ZK_ROUND(FIRST, MINB)                                                                                                              \
        ZK_LAUNCH(ctx, (k_affine_round<F, K, FIRST, MINB>), grid_for(threads, 128), 128, 0, d_points, s.idx, qin, qout, s.off, s.order, s.hist, \
                  s.bin_start, pp, plan, r, cap[r], cap[r + 1], W)
        if (r == 0) { if (tight) ZK_ROUND(true, 4); else ZK_ROUND(true, 1); }
        else { if (tight) ZK_ROUND(false, 4); else ZK_ROUND(false, 1); }
#undef ZK_ROUND
        if (r == 0)
            ZK_LAUNCH(ctx, (k_affine_carry<F, true>), grid_for(slots, 256), 256, 0, d_points, s.idx, qin, qout, s.off, s.cnt, s.order, plan, s.heavy_t, r,
                      cap[r], cap[r + 1], slots);
        else
            ZK_LAUNCH(ctx, (k_affine_carry<F, false>), grid_for(slots, 256), 256, 0, d_points, s.idx, qin, qout, s.off, s.cnt, s.order, plan, s.heavy_t, r,
                      cap[r], cap[r + 1], slots);
    }
    const Affine<F> *q = (R & 1) ? ctx->tree_a.as<Affine<F>>() : ctx->tree_b.as<Affine<F>>();
    ZK_LAUNCH(ctx, (k_accumulate_pts<F>), grid_for(slots, 128), 128, 0, q, s.off, s.cnt, s.order, plan, s.heavy_t, R, cap[R], buckets);
    *done = true;
    return ZKPOR_OK;
}

int32_t msm_tree_sums(zkpor_ctx *ctx, const G1Affine *d_points, const MsmSorted &s, G1XYZZ *buckets, bool *done) { return tree_sums<Fp>(ctx, d_points, s, buckets, done); }
int32_t msm_tree_sums(zkpor_ctx *ctx, const G2Affine *d_points, const MsmSorted &s, G2XYZZ *buckets, bool *done) { return tree_sums<Fp2>(ctx, d_points, s, buckets, done); }

}  // namespace zk
