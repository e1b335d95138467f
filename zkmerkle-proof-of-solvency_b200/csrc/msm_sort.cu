// Scalar side of the Pippenger MSM of msm.cu: digit extraction, counting sort of the point references by bucket, the
// population schedule, the heavy-bucket plan and the per-multiplication views of a shared sort.  No curve arithmetic here.
// Same reference seam as msm.cu: gnark-crypto G1Jac.MultiExp / G2Jac.MultiExp inside groth16.Prove (src/prover/prover/prover.go:269).
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
// mask (optional): scalars whose bit is set are not read and count as zero -- wires that a concurrent kernel is still solving
// (the solver's deferred tail); their terms are added by a multiplication of their own afterwards
__global__ void k_from_mont(const Fr *__restrict__ in, Fr *__restrict__ out, uint64_t n, const uint32_t *__restrict__ mask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask != nullptr && ((mask[i >> 5] >> (i & 31)) & 1u)) { out[i] = Fr::zero(); return; }
    out[i] = Fr::from_mont(in[i]);
}

// plain scalars may be any 256-bit integers: bring them below r (at most five subtractions), so that the signed-digit recoding's
// top carry always fits the windows sized for 254-bit values
__global__ void k_canon_plain(const Fr *__restrict__ in, Fr *__restrict__ out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = in[i];
    const Fr m = Fr::modulus();
    for (int k = 0; k < 6; k++) {
        bool ge = true;
        for (int j = 7; j >= 0; j--) { if (x.l[j] > m.l[j]) break; if (x.l[j] < m.l[j]) { ge = false; break; } }
        if (!ge) break;
        uint64_t borrow = 0;
        for (int j = 0; j < 8; j++) { uint64_t t = (uint64_t)x.l[j] - m.l[j] - borrow; x.l[j] = (uint32_t)t; borrow = (t >> 63) & 1; }
    }
    out[i] = x;
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

// Shared sort (groth16.cu): A, B1, B2 and K multiply subsets of ONE wire vector, so the digits and the counting sort are
// computed once over all wires.  Each multiplication then gets a VIEW of the sorted lists: its map (word i>>5 = skip bits of
// 32 wires -- query point at infinity / wire not in this multiplication -- and the rank of the first of them in the compact
// key array) turns wire references into key-point references.  Light lists are compacted (k_view_lists: one warp
// per list; the kept count becomes the list length, so the population schedule balances REAL additions); the lists of heavy
// buckets keep their length and get REF_SKIP markers (k_view_heavy: one CTA per chunk), which k_accumulate_heavy steps over.
__device__ __forceinline__ uint32_t view_ref(const uint2 *__restrict__ map, uint32_t e) {
    const uint32_t wire = e >> 1, bit = wire & 31;
    const uint2 mw = __ldg(map + (wire >> 5));
    if ((mw.x >> bit) & 1u) return REF_SKIP;
    return ((mw.y + __popc(~mw.x & ((1u << bit) - 1u))) << 1) | (e & 1u);
}
// one warp per list: 32 consecutive references per step (coalesced), ballot + popcount compaction, coalesced stores
__global__ void __launch_bounds__(256) k_view_lists(const uint32_t *__restrict__ sorted, uint32_t *__restrict__ out, const uint32_t *__restrict__ off,
                                                    const uint32_t *__restrict__ cnt, uint64_t n, MsmPlan plan, uint32_t heavy_t,
                                                    const uint2 *__restrict__ map, uint32_t *__restrict__ cnt_v) {
    const size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (t >= (size_t)plan.nwin * plan.nb) return;
    const uint32_t m = cnt[t];
    if (m > heavy_t) { if (lane == 0) cnt_v[t] = m; return; }
    const size_t base = (size_t)(t / plan.nb) * n + off[t];
    uint32_t kept = 0;
    for (uint32_t k0 = 0; k0 < m; k0 += 32) {
        const uint32_t k = k0 + lane;
        const uint32_t r = k < m ? view_ref(map, __ldg(sorted + base + k)) : REF_SKIP;
        const uint32_t keep = __ballot_sync(0xFFFFFFFFu, r != REF_SKIP);
        if (r != REF_SKIP) out[base + kept + __popc(keep & ((1u << lane) - 1u))] = r;
        kept += __popc(keep);
    }
    if (lane == 0) cnt_v[t] = kept;
}
__global__ void __launch_bounds__(256) k_view_heavy(const uint32_t *__restrict__ sorted, uint32_t *__restrict__ out, const HeavyBlk *__restrict__ blks,
                                                    const uint32_t *__restrict__ counters, uint64_t n, MsmPlan plan, const uint2 *__restrict__ map) {
    const uint32_t nblk = counters[0];
    for (uint32_t b = blockIdx.x; b < nblk; b += gridDim.x) {
        const HeavyBlk blk = blks[b];
        const size_t base = (size_t)(blk.slot / plan.nb) * n + blk.start;
        for (uint32_t k = threadIdx.x; k < blk.count; k += blockDim.x) out[base + k] = view_ref(map, __ldg(sorted + base + k));
    }
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

// ---- partitioned counting sort ------------------------------------------------------------------------------------
// The direct scatter above needs one RETURNING L2 atomic per (term, window) -- ~25 G/s on B200, 35 ms for 2^26 terms.  For
// large inputs the sort runs in two levels instead: (A) every (term, window) entry goes to one of PARTS partitions of its
// window by the top bits of its bucket -- counting and cursors in shared memory, one global reservation per (CTA,
// partition); (B) one CTA per (window, partition) owns <= 8192 buckets: histogram, scan and cursors all in shared memory.
// The only global atomics left are the reservations; cnt[] / off[] fall out of (B), so the RED histogram pass goes too.
static const uint32_t PART_BITS = 6, PARTS = 1u << PART_BITS, PART_TILE = 4096;
struct PartPlan { uint32_t shift_norm, shift_top; };   // fine bits (bucket-1 & mask) of an ordinary window / of the top window

__device__ __forceinline__ void load_scalar(const uint32_t *__restrict__ scalars, uint64_t i, uint32_t (&s)[8]) {
    const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + 8 * i);
    uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
}
// signed digit of window w given the carry of window w-1: bucket d in [0, nb], neg, carry out
__device__ __forceinline__ uint32_t signed_digit(const uint32_t (&s)[8], uint32_t w, const MsmPlan &plan, uint32_t &carry, uint32_t &neg) {
    uint32_t d = window_bits(s, w * plan.c, plan.c) + carry;
    neg = 0;
    if (d > plan.nb) { d = (1u << plan.c) - d; neg = 1; carry = 1; } else carry = 0;
    return d;
}

// (A0) entries per (window, partition)
__global__ void __launch_bounds__(256) k_part_count(const uint32_t *__restrict__ scalars, uint64_t n, MsmPlan plan, PartPlan pp, uint32_t *__restrict__ part_cnt) {
    __shared__ uint32_t sh[32 * PARTS];
    const uint32_t nbins = plan.nwin * PARTS;
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint64_t tile = (uint64_t)blockIdx.x * PART_TILE;
    for (uint32_t it = 0; it < PART_TILE / 256; it++) {
        const uint64_t i = tile + it * 256 + threadIdx.x;
        if (i >= n) break;
        uint32_t s[8]; load_scalar(scalars, i, s);
        uint32_t carry = 0, neg;
        for (uint32_t w = 0; w < plan.nwin; w++) {
            const uint32_t d = signed_digit(s, w, plan, carry, neg);
            if (d) atomicAdd(&sh[w * PARTS + ((d - 1) >> (w + 1 == plan.nwin ? pp.shift_top : pp.shift_norm))], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) if (sh[i]) atomicAdd(&part_cnt[i], sh[i]);
}
// exclusive scan of the partition counts (<= 2048 values, one CTA): part_base[0..nbins], cursors = copy
__global__ void __launch_bounds__(1024) k_part_scan(const uint32_t *__restrict__ part_cnt, uint32_t nbins, uint64_t *__restrict__ part_base, unsigned long long *__restrict__ part_cur) {
    __shared__ uint64_t sh[2048];
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = i < nbins ? part_cnt[i] : 0;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t run = 0; for (uint32_t i = 0; i < nbins; i++) { uint64_t v = sh[i]; sh[i] = run; run += v; } part_base[nbins] = run; }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) { part_base[i] = sh[i]; part_cur[i] = sh[i]; }
}
// (A1) entries (fine bucket << 32 | signed reference) written to their partition.  A CTA takes 1024 terms, four per thread
// with their scalars in registers, and walks the windows in the OUTER loop (the carry chain of the signed digits runs
// along it): the 1024 entries of one window leave together as ~64 runs of 128 bytes, so the L2 merges them into whole
// sectors (terms-outer order spread every run over the CTA's lifetime: 13 GB of DRAM writes for 7 GB of entries).
static const uint32_t SCATTER_TILE = 1024;
__global__ void __launch_bounds__(256) k_part_scatter(const uint32_t *__restrict__ scalars, uint64_t n, MsmPlan plan, PartPlan pp,
                                                      unsigned long long *__restrict__ part_cur, uint64_t *__restrict__ entries) {
    __shared__ uint32_t sh[PARTS];
    __shared__ unsigned long long base[PARTS];
    const uint64_t tile = (uint64_t)blockIdx.x * SCATTER_TILE;
    uint32_t s[4][8], carry[4];
    bool live[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint64_t i = tile + k * 256 + threadIdx.x;
        live[k] = i < n; carry[k] = 0;
        if (live[k]) load_scalar(scalars, i, s[k]);
        else { for (int j = 0; j < 8; j++) s[k][j] = 0; }
    }
    for (uint32_t w = 0; w < plan.nwin; w++) {
        const uint32_t sft = w + 1 == plan.nwin ? pp.shift_top : pp.shift_norm;
        if (threadIdx.x < PARTS) sh[threadIdx.x] = 0;
        __syncthreads();
        uint32_t d[4], neg[4], pos[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            d[k] = signed_digit(s[k], w, plan, carry[k], neg[k]);
            pos[k] = d[k] ? atomicAdd(&sh[(d[k] - 1) >> sft], 1u) : 0;   // position inside the CTA's run of that partition
        }
        __syncthreads();
        if (threadIdx.x < PARTS) {
            const uint32_t c = sh[threadIdx.x];
            base[threadIdx.x] = c ? atomicAdd(&part_cur[w * PARTS + threadIdx.x], (unsigned long long)c) : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!d[k]) continue;
            const uint32_t part = (d[k] - 1) >> sft, fine = (d[k] - 1) & ((1u << sft) - 1u);
            const uint64_t i = tile + k * 256 + threadIdx.x;
            entries[base[part] + pos[k]] = ((uint64_t)fine << 32) | (((uint32_t)i << 1) | neg[k]);
        }
    }
}
// (B) one CTA per (window, partition): cnt / off of its buckets and the sorted references
__global__ void __launch_bounds__(1024) k_part_sort(const uint64_t *__restrict__ entries, const uint64_t *__restrict__ part_base, uint64_t n, MsmPlan plan,
                                                    PartPlan pp, uint32_t *__restrict__ cnt, uint32_t *__restrict__ off, uint32_t *__restrict__ sorted) {
    __shared__ uint32_t sh[8192];
    __shared__ uint32_t part[1024];
    const uint32_t bin = blockIdx.x, w = bin / PARTS, pidx = bin % PARTS, t = threadIdx.x;
    const uint32_t sft = w + 1 == plan.nwin ? pp.shift_top : pp.shift_norm, nfine = 1u << sft;
    const uint64_t lo = part_base[bin], hi = part_base[bin + 1];
    const uint32_t first_bucket = pidx << sft;                      // bucket-1 of fine index 0
    if (first_bucket >= plan.nb) return;                            // partition beyond the window's buckets (never populated)
    for (uint32_t i = t; i < nfine; i += 1024) sh[i] = 0;
    __syncthreads();
    // four loads in flight per thread: the pass is bound by load latency, not by the shared-memory atomics
    for (uint64_t e = lo + t; e < hi; e += 4096) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = e + u * 1024 < hi ? __ldg(entries + e + u * 1024) : ~0ull;
#pragma unroll
        for (int u = 0; u < 4; u++) if (e + u * 1024 < hi) atomicAdd(&sh[(uint32_t)(v[u] >> 32)], 1u);
    }
    __syncthreads();
    // exclusive scan of sh[0..nfine): 8 (or fewer) consecutive counters per thread
    const uint32_t per = (nfine + 1023) / 1024, a = t * per, b = a + per < nfine ? a + per : nfine;
    uint32_t sum = 0;
    for (uint32_t i = a; i < b; i++) sum += sh[i];
    part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const uint32_t wrel = (uint32_t)(lo - part_base[w * PARTS]);    // first position of this partition inside the window
    uint32_t run = part[t] - sum;
    for (uint32_t i = a; i < b; i++) {
        const uint32_t c = sh[i];
        const uint32_t bucket = first_bucket + i;
        if (bucket < plan.nb) { cnt[(size_t)w * plan.nb + bucket] = c; off[(size_t)w * plan.nb + bucket] = wrel + run; }
        sh[i] = run;                                                // cursor
        run += c;
    }
    __syncthreads();
    uint32_t *dst = sorted + (size_t)w * n + wrel;
    for (uint64_t e = lo + t; e < hi; e += 4096) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = e + u * 1024 < hi ? __ldg(entries + e + u * 1024) : ~0ull;
#pragma unroll
        for (int u = 0; u < 4; u++) if (e + u * 1024 < hi) dst[atomicAdd(&sh[(uint32_t)(v[u] >> 32)], 1u)] = (uint32_t)v[u];
    }
}

// counting sort of the slots by their reference count (descending): order[], and hist / bin_start behind it in the same buffer
static int32_t population_order(zkpor_ctx *ctx, const uint32_t *cnt, size_t slots, uint32_t *order, MsmSorted *out) {
    uint32_t *hist = order + slots, *cursor = hist + SIZE_BINS, *bin_start = cursor + SIZE_BINS;
    ZK_CUDA(cudaMemsetAsync(hist, 0, SIZE_BINS * 4, ctx->stream));
    ZK_LAUNCH(ctx, k_size_hist, 4 * ctx->sm_count, 256, 0, cnt, slots, hist);
    ZK_LAUNCH(ctx, k_size_scan, 1, 256, 0, (const uint32_t *)hist, cursor, bin_start);
    ZK_LAUNCH(ctx, k_size_scatter, grid_for(slots, 256), 256, 0, cnt, slots, cursor, order);
    out->order = order; out->hist = hist; out->bin_start = bin_start;
    return ZKPOR_OK;
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
    ZK_TRY(ctx->misc.reserve(n * 32));
    const uint32_t *mask = ctx->scalar_mask;
    ctx->scalar_mask = nullptr;
    ZK_REQUIRE(mask == nullptr || !(flags & ZKPOR_SCALARS_PLAIN), "msm_sort: a scalar mask needs Montgomery scalars");
    if (!(flags & ZKPOR_SCALARS_PLAIN)) ZK_LAUNCH(ctx, k_from_mont, grid_for(n, 256), 256, 0, (const Fr *)d_scalars, ctx->misc.as<Fr>(), n, mask);
    else ZK_LAUNCH(ctx, k_canon_plain, grid_for(n, 256), 256, 0, (const Fr *)d_scalars, ctx->misc.as<Fr>(), n);
    plain = ctx->misc.as<uint32_t>();
    ZK_CUDA(cudaMemsetAsync(ctx->bucket_cnt.p, 0, slots * 4, ctx->stream));
    const bool partitioned = n >= (1u << 18) && plan.c >= PART_BITS + 2 && plan.nwin <= 32 && !ctx->direct_scatter;
    if (partitioned) {
        // fine bits per window: an ordinary window has c-1 bucket bits; the top window only 256 - c*(nwin-1) (any 256-bit
        // integer is accepted as a scalar; canonical ones use two bits fewer, i.e. a quarter of its partitions)
        const uint32_t top_raw = 256 > plan.c * (plan.nwin - 1) ? 256 - plan.c * (plan.nwin - 1) : 0;
        const uint32_t top_bits = top_raw < plan.c - 1 ? top_raw : plan.c - 1;   // bucket-1 < 2^top_bits
        PartPlan pp;
        pp.shift_norm = plan.c - 1 - PART_BITS;
        pp.shift_top = top_bits > PART_BITS ? top_bits - PART_BITS : 0;
        const uint32_t nbins = plan.nwin * PARTS;
        ZK_TRY(ctx->part_buf.reserve((size_t)plan.nwin * n * 8));
        ZK_TRY(ctx->part_meta.reserve((size_t)nbins * 4 + (size_t)(nbins + 1) * 8 * 2 + 64));
        uint32_t *part_cnt = ctx->part_meta.as<uint32_t>();
        uint64_t *part_base = (uint64_t *)(ctx->part_meta.as<uint8_t>() + (((size_t)nbins * 4 + 15) & ~(size_t)15));
        unsigned long long *part_cur = (unsigned long long *)(part_base + nbins + 1);
        ZK_CUDA(cudaMemsetAsync(part_cnt, 0, (size_t)nbins * 4, ctx->stream));
        const int tiles = grid_for(n, PART_TILE);
        { KTimed kt(ctx, KC_SORT, n);
          ZK_LAUNCH(ctx, k_part_count, tiles, 256, 0, plain, n, plan, pp, part_cnt);
          kt.stop(); }
        stage_end(ctx, ST_DIGITS);
        stage_begin(ctx, ST_SORT);
        ZK_LAUNCH(ctx, k_part_scan, 1, 1024, 0, (const uint32_t *)part_cnt, nbins, part_base, part_cur);
        { KTimed kt(ctx, KC_SORT, n);
          ZK_LAUNCH(ctx, k_part_scatter, grid_for(n, SCATTER_TILE), 256, 0, plain, n, plan, pp, part_cur, ctx->part_buf.as<uint64_t>());
          ZK_LAUNCH(ctx, k_part_sort, nbins, 1024, 0, (const uint64_t *)ctx->part_buf.p, (const uint64_t *)part_base, n, plan, pp, ctx->bucket_cnt.as<uint32_t>(),
                    ctx->bucket_off.as<uint32_t>(), ctx->sort_idx.as<uint32_t>());
          kt.stop(); }
    } else {
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
    }
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
    ZK_TRY(ctx->order.reserve(slots * 4 + 3 * SIZE_BINS * 4));
    ZK_TRY(population_order(ctx, ctx->bucket_cnt.as<uint32_t>(), slots, ctx->order.as<uint32_t>(), out));
    stage_end(ctx, ST_SORT);
    out->plan = plan; out->n = n;
    out->idx = ctx->sort_idx.as<uint32_t>(); out->off = ctx->bucket_off.as<uint32_t>(); out->cnt = ctx->bucket_cnt.as<uint32_t>();
    return ZKPOR_OK;
}

int32_t msm_view(zkpor_ctx *ctx, const MsmSorted &s, const uint2 *map, MsmSorted *view) {
    const MsmPlan plan = s.plan;
    const size_t slots = (size_t)plan.nwin * plan.nb, total = (size_t)plan.nwin * s.n;
    ZK_TRY(ctx->sort_idx2.reserve(total * 4));
    ZK_TRY(ctx->view_cnt.reserve(slots * 4));
    ZK_TRY(ctx->view_order.reserve(slots * 4 + 3 * SIZE_BINS * 4));
    *view = s;
    stage_begin(ctx, ST_SORT);
    {
        KTimed kt(ctx, KC_SORT, s.n);
        ZK_LAUNCH(ctx, k_view_lists, grid_for(slots * 32, 256), 256, 0, s.idx, ctx->sort_idx2.as<uint32_t>(), s.off, s.cnt, s.n, plan, s.heavy_t, map,
                  ctx->view_cnt.as<uint32_t>());
        ZK_LAUNCH(ctx, k_view_heavy, (s.max_blks < 8u * ctx->sm_count ? s.max_blks : 8u * ctx->sm_count), 256, 0, s.idx, ctx->sort_idx2.as<uint32_t>(), s.blks,
                  s.counters, s.n, plan, map);
        kt.stop();
    }
    view->idx = ctx->sort_idx2.as<uint32_t>(); view->cnt = ctx->view_cnt.as<uint32_t>(); view->is_view = true;
    ZK_TRY(population_order(ctx, view->cnt, slots, ctx->view_order.as<uint32_t>(), view));
    stage_end(ctx, ST_SORT);
    return ZKPOR_OK;
}

}  // namespace zk
