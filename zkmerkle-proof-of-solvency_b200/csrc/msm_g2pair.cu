// G2 bucket accumulation with TWO LANES PER BUCKET: lane 2k holds the a0 components, lane 2k+1 the a1 components of every
// Fp2 value of bucket k's accumulator (X, Y, ZZ, ZZZ) and of the point being added.
// Why: the one-thread-per-bucket kernel (msm.cu k_accumulate<Fp2>) needs 255 registers -- 8 warps per SM, two per scheduler --
// and the carry chains of the field products then keep the IMAD.WIDE pipe only 73 % busy (G1, at 128 registers and 16 warps:
// 93 %).  Split across a lane pair every value is 8 registers instead of 16 and the kernel fits 128.
// Arithmetic on split values (all branch-free across the two roles -- even and odd lanes of a warp never diverge):
//   add / sub / neg   component-wise, no communication
//   sqr(a)            lane0: (a0+a1)(a0-a1), lane1: 2 a0 a1           -- one product per lane, one 8-word exchange
//   mul2(x,y,u,v)     TWO Fp2 products at once, Karatsuba spread evenly: lane0: x0y0, u0v0, (x0+x1)(y0+y1); lane1: x1y1, u1v1,
//                     (u0+u1)(v0+v1) -- three products per lane, four 8-word exchanges (operands, then M1-A / C and B / D)
// The mixed addition madd-2008-s is exactly four independent product pairs and two squarings: 14 products per lane = the 28
// base-field products of the one-thread form, no extra multiplications.
// STATUS: opt-in (ZKPOR_G2_PAIR; see g2_pair_lanes in msm.cu): gives the oracle's result, but measured 27 % slower than one thread
// per bucket (2^22 terms: 44 ms vs 35 ms) -- the unconditional special-case arithmetic and the spills cost more than the occupancy gains.
// Same reference seam as msm.cu: gnark-crypto G2Jac.MultiExp inside groth16.Prove (src/prover/prover/prover.go:269).
#include "internal.h"
#include "fp2_split.cuh"
#include <cstdlib>

using namespace ff;
using namespace ec;

namespace zk {

namespace {

// device lane pair: role = lane & 1, exchange = 8 shuffles with the partner lane; `mask` names the lanes that execute the call
struct LanePair {
    uint32_t mask; bool role;
    __device__ __forceinline__ Fp xchg(const Fp &v) const {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(mask, v.l[i], 1);
        return r;
    }
    __device__ __forceinline__ Fp sel_role(const Fp &a, const Fp &b) const {   // role ? a : b without a branch
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = role ? a.l[i] : b.l[i];
        return r;
    }
    __device__ __forceinline__ bool pair_zero(const Fp &v) const {             // both components zero (every lane of `mask` must call this)
        const int z = v.is_zero();
        return z & __shfl_xor_sync(mask, z, 1);
    }
};
using SplitAcc = fp2split::Acc<Fp>;

__device__ __forceinline__ Fp load_fp(const Fp *p) {
    Fp r;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
    uint4 lo = __ldg(src), hi = __ldg(src + 1);
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w; r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}

}  // namespace

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_accumulate_g2_pair(const G2Affine *__restrict__ points, const uint32_t *__restrict__ sorted,
                                                               const uint32_t *__restrict__ off, const uint32_t *__restrict__ cnt, uint64_t n,
                                                               MsmPlan plan, uint32_t heavy_t, const uint32_t *__restrict__ order,
                                                               G2XYZZ *__restrict__ buckets) {
    const uint32_t FULL = 0xFFFFFFFFu;
    const size_t slots = (size_t)plan.nwin * plan.nb;
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;   // bucket (pair) index in the population order
    const bool role = threadIdx.x & 1;
    const bool live = q < slots;                                             // whole pairs are live or not
    uint32_t t = 0, m = 0;
    const uint32_t *idx = sorted;
    if (live) {
        t = order[q];
        m = cnt[t];
        idx = sorted + (size_t)(t / plan.nb) * n + off[t];
        if (m > heavy_t) m = 0;                                              // summed by k_accumulate_heavy
    }
    // every lane of the warp runs the same number of steps (the shuffles need all of them); lists of a warp are equally long
    uint32_t steps = m;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) steps = max(steps, __shfl_xor_sync(FULL, steps, d));

    const Fp one_split = role ? Fp::zero() : Fp::one();                      // Fp2 one = (1, 0)
    SplitAcc acc;
    acc.X = one_split; acc.Y = one_split; acc.ZZ = Fp::zero(); acc.ZZZ = Fp::zero();
    bool acc_inf = true;
    // component `role` of point e (x then y: 4 Fp per G2Affine)
    auto load_pt = [&](uint32_t e, Fp &x, Fp &y) {
        const Fp *base = reinterpret_cast<const Fp *>(points + (e >> 1));
        x = load_fp(base + role); y = load_fp(base + 2 + role);
    };
    uint32_t e = m ? __ldg(idx) : 0;
    Fp px = Fp::zero(), py = Fp::zero();
    if (m) load_pt(e, px, py);
    const LanePair ln{FULL, role};
    for (uint32_t k = 0; k < steps; k++) {
        const bool active = k < m;
        // next reference and point in flight while this one is added
        uint32_t en = 0; Fp nx = px, ny = py;
        if (k + 1 < m) { en = __ldg(idx + k + 1); load_pt(en, nx, ny); }
        // NOTE: everything that shuffles (pair_zero, madd, ...) is evaluated by EVERY lane before it is combined with lane-dependent
        // conditions: a shuffle behind a short-circuited `&&` deadlocked the first version of this kernel
        const bool px_zero = ln.pair_zero(px), py_zero = ln.pair_zero(py);
        const bool p_inf = px_zero & py_zero;
        const Fp y2 = (e & 1) ? Fp::neg(py) : py;
        // madd-2008-s on split values (fp2_split.cuh; computed unconditionally, special cases are selected afterwards)
        Fp P, R;
        SplitAcc nxt = fp2split::madd(ln, acc, px, y2, P, R);
        const bool p_zero = ln.pair_zero(P), r_zero = ln.pair_zero(R);
        bool nxt_inf = false;
        // same x: doubling (same y) or cancellation -- rare, pair-uniform; the lanes concerned are named by a ballot
        const bool same_x = active & !p_inf & !acc_inf & p_zero;
        const bool need_dbl = same_x & r_zero;
        const uint32_t dmask = __ballot_sync(FULL, need_dbl);
        if (need_dbl) nxt = fp2split::dbl_affine(LanePair{dmask, role}, px, y2);
        if (same_x & !r_zero) nxt_inf = true;
        // first point of the list: the accumulator becomes the point
        if (acc_inf) { nxt.X = px; nxt.Y = y2; nxt.ZZ = one_split; nxt.ZZZ = one_split; nxt_inf = false; }
        if (active & !p_inf) {
            acc = nxt; acc_inf = nxt_inf;
            if (nxt_inf) { acc.X = one_split; acc.Y = one_split; acc.ZZ = Fp::zero(); acc.ZZZ = Fp::zero(); }
        }
        e = en; px = nx; py = ny;
    }
    if (live && cnt[t] <= heavy_t) {
        Fp *out = reinterpret_cast<Fp *>(buckets + t);   // X.a0 X.a1 Y.a0 Y.a1 ZZ.a0 ZZ.a1 ZZZ.a0 ZZZ.a1
        if (acc_inf) { acc.X = one_split; acc.Y = one_split; acc.ZZ = Fp::zero(); acc.ZZZ = Fp::zero(); }
        out[0 + role] = acc.X; out[2 + role] = acc.Y; out[4 + role] = acc.ZZ; out[6 + role] = acc.ZZZ;
    }
}

int32_t msm_g2_pair_accumulate(zkpor_ctx *ctx, const G2Affine *d_points, const MsmSorted &s, G2XYZZ *buckets) {
    const size_t slots = (size_t)s.plan.nwin * s.plan.nb;
    // registers per thread: 128 / 168 / 244 at 4 / 3 / 2 resident CTAs per SM (the two lower ones spill 0.9 / 0.4 KB)
    static const int minb = [] { const char *v = getenv("ZKPOR_G2_PAIR"); const int k = v ? atoi(v) : 4; return k == 2 || k == 3 ? k : 4; }();
    if (minb == 2)
        ZK_LAUNCH(ctx, k_accumulate_g2_pair<2>, grid_for(slots * 2, 128), 128, 0, d_points, s.idx, s.off, s.cnt, s.n, s.plan, s.heavy_t, s.order, buckets);
    else if (minb == 3)
        ZK_LAUNCH(ctx, k_accumulate_g2_pair<3>, grid_for(slots * 2, 128), 128, 0, d_points, s.idx, s.off, s.cnt, s.n, s.plan, s.heavy_t, s.order, buckets);
    else
        ZK_LAUNCH(ctx, k_accumulate_g2_pair<4>, grid_for(slots * 2, 128), 128, 0, d_points, s.idx, s.off, s.cnt, s.n, s.plan, s.heavy_t, s.order, buckets);
    return ZKPOR_OK;
}

}  // namespace zk
