// Witness solver on the GPU: gnark's r1cs.Solve (constraint/bn254/solver.go + system.go, out of tree) -- the first step of
// groth16.Prove, src/prover/prover/prover.go:269 -- with the hint functions the reference circuit uses (IntegerDivision,
// circuit/utils.go:103-110, registered at prover.go:68; gnark's std hints behind api.ToBinary / api.IsZero / rangecheck.Check /
// logderivlookup; the BSB22 commitment placeholder that Prove overrides).
//
// gnark walks cs.Levels: the instructions of a level are independent, levels are barriers.  An R1C instruction has exactly one
// unsolved wire, found at run time; here it is found ONCE, at upload, by a dry run of the same schedule on solved-flags, so the
// solving kernels only evaluate and store.
//
// Schedule.  The reference circuit has two very different regimes: thousands of levels that are ~10^3..10^6
// instructions wide (one block per user: range checks, lookups, Merkle paths) and a tail of ~10^5 levels that are 1..13 wide (the
// two 10 000-element CEX commitments are serial sponge chains).  So:
//   * a WIDE level is one grid launch, 8 lanes per instruction: the lanes split the terms of the three linear expressions
//     (Poseidon rows carry up to ~80 terms), reduce with shuffles, lane 0 solves and stores;
//   * a run of consecutive NARROW levels is ONE launch of a single CTA: a warp per instruction, __syncthreads between levels --
//     a level then costs a dependent L2 round trip instead of a kernel launch;
//   * COUNT (multiplicities) is a histogram over all queries; COMMIT gathers the committed wires, runs the Pedersen commitment and
//     its proof of knowledge (one sort, two accumulations), hashes the point to the field on the host, stores the challenge.
// The solver writes wire values only; a = Lw, b = Rw, c = Ow and the satisfaction check are one wide pass afterwards (r1cs.cu).
#include "internal.h"
#include <algorithm>
#include <cstdlib>

using namespace ff;
using namespace ec;

namespace zk {

static const uint32_t NARROW_MAX = 96;          // levels up to this many instructions are fused into single-CTA runs (env ZKPOR_NARROW_MAX)
static const int NARROW_THREADS = 512;          // upper bound of the fused-run CTA (env ZKPOR_NARROW_THREADS picks fewer warps)
static const uint64_t SOLVE_NONE = ~0ull;
static const uint32_t HINT_BIT = 0x80000000u;
// solve_e[row]: which term of the constraint is the unknown -- side in bits 62..63, position in the side's term list in bits 0..35 --
// and, for rows of narrow levels, how many of each side's terms read a wire solved in the level just before (bits 36..59, eight per
// side; those terms are moved to the end of the side's list at upload): everything else of a row can be summed one level ahead
// (k_solve_narrow_pipe).  SE_ALLFRESH: a side has more than 255 such terms, nothing of the row is summed ahead.
static const uint64_t SE_POS_MASK = (1ull << 36) - 1, SE_ALLFRESH = 1ull << 60;
__host__ __device__ __forceinline__ int se_side(uint64_t se) { return (int)(se >> 62); }
__host__ __device__ __forceinline__ uint64_t se_pos(uint64_t se) { return se & SE_POS_MASK; }
__host__ __device__ __forceinline__ uint32_t se_fresh(uint64_t se, int side) { return (uint32_t)(se >> (36 + 8 * side)) & 255u; }

enum StepKind { STEP_WIDE = 0, STEP_NARROW, STEP_COUNT, STEP_COMMIT };
struct Step { int kind; uint64_t a, b; bool has_div; uint64_t n_long; };   // WIDE: sched range [a, b), the first n_long rows long; NARROW: levels [a, b); COUNT / COMMIT: hint id a
enum SolveErr { SE_OK = 0, SE_UNSOLVED = 1, SE_DIV0 = 2, SE_INDEX = 3, SE_HINT = 4 };

struct Pending;
struct ProgView {
    const uint64_t *ptr[3]; const uint32_t *wire[3], *coef[3];
    const uint64_t *aux_ptr; const uint32_t *aux_wire, *aux_coef;
    const Fr *coeffs; uint32_t one_id, minus_one_id;
    const uint32_t *sched; const uint64_t *lvl_start;
    const uint32_t *hint_fn, *hint_param, *hint_out, *hint_nout; const uint64_t *hint_in0, *hint_in1;
    const uint64_t *table_ptr;
    uint64_t *solve_e;
    Fr *w; uint8_t *solved; unsigned long long *err;      // err: (code << 56) | row or hint id, first writer wins
    struct Pending *pend;                                 // wide levels: divisions deferred to k_solve_div, slot = position in the level
    uint8_t *step_div;                                    // dry run: step s has at least one division
    unsigned long long *prof;                             // ZKPOR_NARROW_PROF: cycle counters of k_solve_narrow_pipe's phases (development)
    uint32_t *wstep; uint32_t cur_step;                   // dry run: wstep[wire] = 1 + the schedule step that solves it (0: an input)
};
// w[wire] = num / den, written by the evaluating group, consumed (and cleared) by k_solve_div
struct alignas(16) Pending { Fr num, den; uint32_t wire, state, pad0, pad1; };   // state: 0 empty, 1 division, 2 division where den = 0 gives 0 (InvZero)
static const uint64_t NO_SLOT = ~0ull;

}  // namespace zk

struct zkpor_program {
    uint64_t n_wires = 0, n_public = 0, n_secret = 0, n_rows = 0, n_instr = 0, n_levels = 0, n_hints = 0, n_aux_rows = 0, n_tables = 0;
    zkpor_r1cs *cs = nullptr;
    uint64_t *aux_ptr = nullptr; uint32_t *aux_wire = nullptr, *aux_coef = nullptr;
    uint32_t *sched = nullptr; uint64_t *lvl_start = nullptr;
    uint32_t *hint_fn = nullptr, *hint_param = nullptr, *hint_out = nullptr, *hint_nout = nullptr; uint64_t *hint_in0 = nullptr, *hint_in1 = nullptr;
    uint64_t *table_ptr = nullptr;
    uint64_t *solve_e = nullptr;
    zk::Pending *pend = nullptr; uint64_t pend_cap = 0;
    uint8_t *step_div = nullptr;
    unsigned long long *err = nullptr;
    uint32_t *counters = nullptr; uint64_t counters_cap = 0;
    uint32_t minus_one_id = 0xFFFFFFFFu;
    std::vector<zk::Step> steps;
    std::vector<uint32_t> h_hint_fn, h_hint_nout, h_hint_out; std::vector<uint64_t> h_hint_in0, h_hint_in1;
    uint64_t stats[4] = {0, 0, 0, 0}, stats_long = 0;
    bool has_commit = false;
    uint32_t narrow_max = zk::NARROW_MAX; int narrow_threads = zk::NARROW_THREADS; uint32_t long_row = 0;
    zk::DevBuf wires, abc;
    // The deferred tail (see run_schedule): the last step, when it is a long run of narrow levels, starts on a side stream as soon as
    // the steps it reads from have been enqueued (tail_after) and runs beside the rest of the proof; its wires are listed here.
    int64_t tail_step = -1; uint64_t tail_after = 0, n_tail_wires = 0, uid = 0;
    uint32_t *tail_wires = nullptr, *tail_mask = nullptr; uint32_t *wstep = nullptr;
    std::vector<uint32_t> h_tail_wires;
    cudaEvent_t tail_go = nullptr, tail_done = nullptr; bool tail_running = false;
    unsigned long long *prof = nullptr;
    bool narrow_pipe = false;                      // narrow levels by k_solve_narrow_pipe (rows partitioned at upload)
    const char *trace_path = nullptr;              // env ZKPOR_SOLVE_TRACE: per-step device times of the next solve, written as CSV
};

namespace zk {

__device__ __forceinline__ void solve_fail(const ProgView &v, int code, uint64_t where) {
    atomicCAS(v.err, 0ull, ((unsigned long long)code << 56) | where);
}

// one term of a linear expression: coefficient ids 1 and -1 skip the product
__device__ __forceinline__ Fr term_acc(const ProgView &v, const Fr &acc, uint32_t cid, const Fr &x) {
    if (cid == v.one_id) return Fr::add(acc, x);
    if (cid == v.minus_one_id) return Fr::sub(acc, x);
    return Fr::add(acc, Fr::mul(v.coeffs[cid], x));
}

// ---- unreduced sums.  A linear expression of a Poseidon row has ~80 terms, spread over the lanes of a group: reducing modulo r
// after every addition of the tree sum (an addition, a comparison and a conditional subtraction of eight limbs, per shuffle step) cost
// more than the products.  The lanes therefore add their residues as plain integers in nine limbs (room for 2^32 terms), the tree
// adds nine limbs per step, and ONE lane reduces the total: S = hi * 2^256 + lo, hi * 2^256 mod r = the Montgomery form of hi (a
// 64-entry table in constant memory; a product beyond it), then lo + table[hi] < 2^256 + r comes below r by conditional
// subtractions of 4r, 2r and r.
struct Wide9 { uint32_t l[9]; };
__constant__ Fr c_hi_mont[64];                    // c_hi_mont[k] = k * 2^256 mod r
__device__ __forceinline__ Wide9 w9_zero() { Wide9 a; for (int i = 0; i < 9; i++) a.l[i] = 0; return a; }
__device__ __forceinline__ void w9_add(Wide9 &a, const Fr &x) {
    a.l[0] = ptx::add_cc(a.l[0], x.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) a.l[i] = ptx::addc_cc(a.l[i], x.l[i]);
    a.l[8] = ptx::addc(a.l[8], 0u);
}
__device__ __forceinline__ void w9_addw(Wide9 &a, const Wide9 &b) {
    a.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) a.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    a.l[8] = ptx::addc(a.l[8], b.l[8]);
}
// v -= k*r when v >= k*r (v: nine limbs, k*r < 2^256)
template <int K>
__device__ __forceinline__ void w9_csub(Wide9 &v) {
    Wide9 d;
    uint32_t m[8]; uint32_t carry = 0;                    // K * r, K = 1, 2, 4
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)FrParams::M(i) * K + carry; m[i] = (uint32_t)t; carry = (uint32_t)(t >> 32); }
    d.l[0] = ptx::sub_cc(v.l[0], m[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) d.l[i] = ptx::subc_cc(v.l[i], m[i]);
    d.l[8] = ptx::subc_cc(v.l[8], 0u);
    const uint32_t borrow = ptx::subc(0u, 0u);            // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < 9; i++) v.l[i] = borrow ? v.l[i] : d.l[i];
}
__device__ __forceinline__ Fr w9_reduce(const Wide9 &s) {
    const uint32_t hi = s.l[8];
    const Fr t = hi < 64 ? c_hi_mont[hi] : Fr::from_u64(hi);
    Wide9 v;
    v.l[0] = ptx::add_cc(s.l[0], t.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) v.l[i] = ptx::addc_cc(s.l[i], t.l[i]);
    v.l[8] = ptx::addc(0u, 0u);
    w9_csub<4>(v); w9_csub<2>(v); w9_csub<1>(v);   // 2^256 < 5.3 r, so v < 6.3 r: minus 4r if possible -> < 4r; minus 2r -> < 2r; minus r -> < r
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = v.l[i];
    return r;
}
// one term: the coefficients 1 and -1 skip the product
__device__ __forceinline__ void term_w9(const ProgView &v, Wide9 &acc, uint32_t cid, const Fr &x) {
    if (cid == v.one_id) w9_add(acc, x);
    else if (cid == v.minus_one_id) w9_add(acc, Fr::neg(x));
    else w9_add(acc, Fr::mul(v.coeffs[cid], x));
}
// tree sum over the G lanes of a group; valid on lane 0
template <int G>
__device__ __forceinline__ Wide9 group_sum_w9(Wide9 acc, unsigned mask) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        Wide9 o;
#pragma unroll
        for (int i = 0; i < 9; i++) o.l[i] = __shfl_down_sync(mask, acc.l[i], off, G);
        w9_addw(acc, o);
    }
    return acc;
}
__device__ __forceinline__ Wide9 lane_dot_w9(const ProgView &v, const uint64_t *ptr, const uint32_t *wire, const uint32_t *coef, uint64_t row, uint64_t skip,
                                             int lane, int stride) {
    Wide9 acc = w9_zero();
    const uint64_t e1 = ptr[row + 1];
    for (uint64_t e = ptr[row] + lane; e < e1; e += stride) {
        if (e == skip) continue;
        term_w9(v, acc, coef[e], v.w[wire[e]]);
    }
    return acc;
}

// tree sum over the G lanes of a group; valid on lane 0
template <int G>
__device__ __forceinline__ Fr group_sum(Fr acc, unsigned mask) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.l[i] = __shfl_down_sync(mask, acc.l[i], off, G);
        acc = Fr::add(acc, o);
    }
    return acc;
}

// sum over the terms of row `row` except position `skip`, lane `lane` of `stride` lanes taking every stride-th term (partial sum of the lane)
__device__ __forceinline__ Fr lane_dot(const ProgView &v, const uint64_t *ptr, const uint32_t *wire, const uint32_t *coef, uint64_t row, uint64_t skip,
                                       int lane, int stride) {
    Fr acc = Fr::zero();
    const uint64_t e1 = ptr[row + 1];
    for (uint64_t e = ptr[row] + lane; e < e1; e += stride) {
        if (e == skip) continue;
        acc = term_acc(v, acc, coef[e], v.w[wire[e]]);
    }
    return acc;
}

// the same sum by the G lanes of a group; valid on lane 0
template <int G>
__device__ __forceinline__ Fr group_dot(const ProgView &v, const uint64_t *ptr, const uint32_t *wire, const uint32_t *coef, uint64_t row,
                                        uint64_t skip, int lane, unsigned mask) {
    if (G >= 8) {   // long rows: unreduced sums, one reduction (lane 0 of the group)
        const Wide9 s = group_sum_w9<G>(lane_dot_w9(v, ptr, wire, coef, row, skip, lane, G), mask);
        return lane == 0 ? w9_reduce(s) : Fr::zero();
    }
    return group_sum<G>(lane_dot(v, ptr, wire, coef, row, skip, lane, G), mask);
}

__device__ __forceinline__ Fr aux_eval(const ProgView &v, uint64_t row) {
    Fr acc = Fr::zero();
    for (uint64_t e = v.aux_ptr[row], e1 = v.aux_ptr[row + 1]; e < e1; e++) acc = term_acc(v, acc, v.aux_coef[e], v.w[v.aux_wire[e]]);
    return acc;
}

__device__ __forceinline__ bool geq256(const uint32_t *a, const uint32_t *b) {
    for (int i = 7; i >= 0; i--) { if (a[i] > b[i]) return true; if (a[i] < b[i]) return false; }
    return true;
}

// big.Int DivMod on canonical 256-bit values: word division for a one-word divisor, shift-subtract otherwise
__device__ void divmod256(const uint32_t *x, const uint32_t *d, uint32_t *q, uint32_t *r) {
    bool small = true;
    for (int i = 1; i < 8; i++) small &= d[i] == 0;
    if (small) {
        uint64_t rem = 0;
        for (int i = 7; i >= 0; i--) { uint64_t cur = (rem << 32) | x[i]; q[i] = (uint32_t)(cur / d[0]); rem = cur % d[0]; }
        for (int i = 1; i < 8; i++) r[i] = 0;
        r[0] = (uint32_t)rem;
        return;
    }
    for (int i = 0; i < 8; i++) { q[i] = 0; r[i] = 0; }
    for (int bit = 255; bit >= 0; bit--) {
        for (int i = 7; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
        r[0] = (r[0] << 1) | ((x[bit >> 5] >> (bit & 31)) & 1u);
        if (geq256(r, d)) {
            uint64_t borrow = 0;
            for (int i = 0; i < 8; i++) { uint64_t t = (uint64_t)r[i] - d[i] - borrow; r[i] = (uint32_t)t; borrow = (t >> 63) & 1; }
            q[bit >> 5] |= 1u << (bit & 31);
        }
    }
}

__device__ __forceinline__ Fr fr_from_words(const uint32_t *p) { Fr t; for (int i = 0; i < 8; i++) t.l[i] = p[i]; return Fr::to_mont(t); }

// hint functions except COUNT and COMMIT; lane 0 of the group runs them
template <bool DRY>
__device__ void exec_hint(const ProgView &v, uint32_t h, uint64_t slot, uint64_t step) {
    const uint32_t fn = v.hint_fn[h], param = v.hint_param[h], out = v.hint_out[h], n_out = v.hint_nout[h];
    if (DRY) {
        for (uint32_t k = 0; k < n_out; k++) { v.solved[out + k] = 1; v.wstep[out + k] = v.cur_step + 1; }
        if (fn == ZKPOR_HINT_INVZERO && step != NO_SLOT) v.step_div[step] = 1;
        return;
    }
    const uint64_t r0 = v.hint_in0[h], r1 = v.hint_in1[h];
    switch (fn) {
    case ZKPOR_HINT_DIVMOD: {
        const Fr x = Fr::from_mont(aux_eval(v, r0)), d = Fr::from_mont(aux_eval(v, r0 + 1));
        if (d.is_zero()) { solve_fail(v, SE_DIV0, h); return; }
        uint32_t q[8], r[8];
        divmod256(x.l, d.l, q, r);
        v.w[out] = fr_from_words(q); v.w[out + 1] = fr_from_words(r);
        break;
    }
    case ZKPOR_HINT_NBITS: {
        const Fr x = Fr::from_mont(aux_eval(v, r0));
        for (uint32_t k = 0; k < n_out; k++) v.w[out + k] = (k < 256 && ((x.l[k >> 5] >> (k & 31)) & 1u)) ? Fr::one() : Fr::zero();
        break;
    }
    case ZKPOR_HINT_INVZERO:
        if (slot != NO_SLOT) { Pending &pd = v.pend[slot]; pd.num = Fr::one(); pd.den = aux_eval(v, r0); pd.wire = out; pd.state = 2; }
        else v.w[out] = Fr::inv(aux_eval(v, r0));
        break;
    case ZKPOR_HINT_DECOMPOSE: {
        const Fr x = Fr::from_mont(aux_eval(v, r0));
        for (uint32_t k = 0; k < n_out; k++) {
            const uint32_t lo = k * param;
            uint64_t limb = 0;
            if (lo < 256) {
                limb = x.l[lo >> 5] >> (lo & 31);
                if ((lo & 31) + param > 32 && (lo >> 5) + 1 < 8) limb |= (uint64_t)x.l[(lo >> 5) + 1] << (32 - (lo & 31));
                limb &= (param >= 32) ? 0xFFFFFFFFull : ((1ull << param) - 1);
            }
            v.w[out + k] = Fr::from_u64(limb);
        }
        break;
    }
    case ZKPOR_HINT_LOOKUP: {
        const uint64_t t0 = v.table_ptr[param], t1 = v.table_ptr[param + 1];
        for (uint64_t r = r0; r < r1; r++) {
            const Fr q = Fr::from_mont(aux_eval(v, r));
            bool ok = true;
            for (int i = 2; i < 8; i++) ok &= q.l[i] == 0;
            const uint64_t idx = (uint64_t)q.l[0] | ((uint64_t)q.l[1] << 32);
            if (!ok || idx >= t1 - t0) { solve_fail(v, SE_INDEX, h); return; }
            v.w[out + (uint32_t)(r - r0)] = aux_eval(v, t0 + idx);
        }
        break;
    }
    case ZKPOR_HINT_CMP: {
        const Fr x = Fr::from_mont(aux_eval(v, r0)), y = Fr::from_mont(aux_eval(v, r0 + 1));
        const bool ge = geq256(x.l, y.l), le = geq256(y.l, x.l);
        v.w[out] = (ge && le) ? Fr::zero() : (ge ? Fr::one() : Fr::neg(Fr::one()));
        break;
    }
    default: solve_fail(v, SE_HINT, h);
    }
}

// the unknown wire of constraint `row` (term `pos` of side `side`) from the known parts a, b, c of the three sides; one lane
__device__ __forceinline__ void finish_instr(const ProgView &v, uint64_t row, int side, uint64_t pos, const Fr &a, const Fr &b, const Fr &c, uint64_t slot) {
    const uint32_t cid = v.coef[side][pos];
    Fr num, den;
    bool unit = false, neg = false;
    if (side == 2) {                                      // L*R = known + cf*x
        num = Fr::sub(Fr::mul(a, b), c);
        unit = cid == v.one_id; neg = cid == v.minus_one_id;
        den = v.coeffs[cid];
    } else {                                              // (known + cf*x) * other = c
        const Fr &known = side == 0 ? a : b, &other = side == 0 ? b : a;
        num = Fr::sub(c, Fr::mul(known, other));
        den = cid == v.one_id ? other : Fr::mul(v.coeffs[cid], other);
        if (den.is_zero()) { solve_fail(v, SE_DIV0, row); return; }
    }
    const uint32_t wire = v.wire[side][pos];
    if (unit) v.w[wire] = num;
    else if (neg) v.w[wire] = Fr::neg(num);
    else if (slot != NO_SLOT) { Pending &pd = v.pend[slot]; pd.num = num; pd.den = den; pd.wire = wire; pd.state = 1; }
    else v.w[wire] = Fr::mul(num, Fr::inv(den));
}

// one instruction by a group of G lanes (lane = index in the group, mask = the group's lanes)
// slot: where a division is parked for k_solve_div (wide levels), NO_SLOT = divide in place (narrow runs); step: the schedule step (dry run)
template <int G, bool DRY>
__device__ __forceinline__ void exec_instr(const ProgView &v, uint32_t packed, int lane, unsigned mask, uint64_t slot, uint64_t step) {
    if (packed & HINT_BIT) {
        if (lane == 0) exec_hint<DRY>(v, packed & ~HINT_BIT, slot, step);
        return;
    }
    const uint64_t row = packed;
    if (DRY) {
        // the one wire of this constraint that is not solved yet
        uint64_t cand = SOLVE_NONE; uint32_t found = 0;
        for (int side = 0; side < 3; side++)
            for (uint64_t e = v.ptr[side][row] + lane, e1 = v.ptr[side][row + 1]; e < e1; e += G)
                if (!v.solved[v.wire[side][e]]) { cand = ((uint64_t)side << 62) | e; found++; }
        uint32_t total = found;
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) total += __shfl_xor_sync(mask, total, off, G);
        if (total > 1) { if (lane == 0) solve_fail(v, SE_UNSOLVED, row); return; }
        if (total == 0) { if (lane == 0) v.solve_e[row] = SOLVE_NONE; return; }
        if (found) {
            const int side = se_side(cand); const uint64_t pos = se_pos(cand);
            v.solve_e[row] = cand; v.solved[v.wire[side][pos]] = 1; v.wstep[v.wire[side][pos]] = v.cur_step + 1;
            const uint32_t cid = v.coef[side][pos];
            if (step != NO_SLOT && (side != 2 || (cid != v.one_id && cid != v.minus_one_id))) v.step_div[step] = 1;
        }
        return;
    }
    const uint64_t se = v.solve_e[row];
    if (se == SOLVE_NONE) return;                         // an assertion: checked with a, b, c after the solve
    const int side = se_side(se);
    const uint64_t pos = se_pos(se);
    if (G == 32) {
        // both products' sides fit a half-warp (the full rounds of a Poseidon permutation: 13 terms): L on lanes 0..15 and R on lanes
        // 16..31 at once, one term per lane -- a third of the serial work of three whole-warp sums
        const uint64_t a0 = v.ptr[0][row], a1 = v.ptr[0][row + 1], b0 = v.ptr[1][row], b1 = v.ptr[1][row + 1];
        if (a1 - a0 <= 16 && b1 - b0 <= 16) {
            const int half = lane >> 4, hl = lane & 15;
            const uint64_t e = (half ? b0 : a0) + hl, e1 = half ? b1 : a1;
            Wide9 acc = w9_zero();
            if (e < e1 && !(side == half && e == pos)) term_w9(v, acc, (half ? v.coef[1] : v.coef[0])[e], v.w[(half ? v.wire[1] : v.wire[0])[e]]);
            acc = group_sum_w9<16>(acc, mask);
            const Fr mine = hl == 0 ? w9_reduce(acc) : Fr::zero();
            Fr bb;
#pragma unroll
            for (int i = 0; i < 8; i++) bb.l[i] = __shfl_sync(mask, mine.l[i], 16);
            const uint64_t c0 = v.ptr[2][row], c1 = v.ptr[2][row + 1];
            const Fr cc = (side == 2 && c1 - c0 == 1) ? Fr::zero() : group_dot<G>(v, v.ptr[2], v.wire[2], v.coef[2], row, side == 2 ? pos : SOLVE_NONE, lane, mask);
            if (lane == 0) finish_instr(v, row, side, pos, mine, bb, cc, slot);
            return;
        }
    }
    const Fr a = group_dot<G>(v, v.ptr[0], v.wire[0], v.coef[0], row, side == 0 ? pos : SOLVE_NONE, lane, mask);
    const Fr b = group_dot<G>(v, v.ptr[1], v.wire[1], v.coef[1], row, side == 1 ? pos : SOLVE_NONE, lane, mask);
    const Fr c = group_dot<G>(v, v.ptr[2], v.wire[2], v.coef[2], row, side == 2 ? pos : SOLVE_NONE, lane, mask);
    if (lane != 0) return;
    finish_instr(v, row, side, pos, a, b, c, slot);
}

// One wide level.  Its instructions are ordered long rows first (zkpor_program_upload): the first n_long take a whole warp each -- a
// Poseidon row has ~80 terms per side, one product per lane -- the others (one to eight terms per side: most of a compiled circuit)
// WIDE_GS lanes each, so that a level of 10^5 short rows is a few hundred CTAs instead of 10^4.
static const int WIDE_GS = 4;
static const uint32_t WIDE_LONG_ROW = 9;          // terms on the longest side from which a row counts as long
template <bool DRY>
__global__ void __launch_bounds__(256) k_solve_wide(ProgView v, uint64_t pos0, uint64_t n_long, uint64_t count, uint64_t step) {
    const uint64_t long_blocks = (n_long + 7) / 8;
    if (blockIdx.x < long_blocks) {
        const uint64_t g = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
        if (g >= n_long) return;
        exec_instr<32, DRY>(v, v.sched[pos0 + g], threadIdx.x & 31, 0xFFFFFFFFu, g, step);
    } else {
        const uint64_t g = n_long + ((uint64_t)(blockIdx.x - long_blocks) * 256 + threadIdx.x) / WIDE_GS;
        if (g >= count) return;                           // whole groups leave together (WIDE_GS divides the block size)
        const int lane = threadIdx.x & (WIDE_GS - 1);
        const unsigned mask = ((1u << WIDE_GS) - 1u) << ((threadIdx.x & 31) & ~(WIDE_GS - 1));
        exec_instr<WIDE_GS, DRY>(v, v.sched[pos0 + g], lane, mask, g, step);
    }
}

// sched[i] -> the longest side of its constraint row (hints: 0), capped at 255: the upload orders a wide level by it
__global__ void k_row_class(ProgView v, uint64_t n, uint8_t *cls) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t packed = v.sched[i];
    uint64_t m = 0;
    if (!(packed & HINT_BIT))
        for (int side = 0; side < 3; side++) m = max(m, v.ptr[side][(uint64_t)packed + 1] - v.ptr[side][packed]);
    cls[i] = (uint8_t)min(m, (uint64_t)255);
}

// The divisions of a wide level, one thread per DIV_BATCH consecutive slots sharing one inversion (Montgomery's trick): an inversion is
// ~23 K instructions of one lane; done by the evaluating group it would idle the group's other lanes for all of them.
static const int DIV_BATCH = 8;
__global__ void __launch_bounds__(128) k_solve_div(ProgView v, uint64_t count) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, s0 = t * DIV_BATCH;
    if (s0 >= count) return;
    Fr pre[DIV_BATCH];
    uint32_t st[DIV_BATCH];
    Fr acc = Fr::one();
#pragma unroll
    for (int k = 0; k < DIV_BATCH; k++) {
        st[k] = s0 + k < count ? v.pend[s0 + k].state : 0;
        pre[k] = acc;
        if (st[k]) {
            const Fr d = v.pend[s0 + k].den;
            if (d.is_zero()) { if (st[k] == 1) solve_fail(v, SE_DIV0, s0 + k); st[k] |= 4; }
            else acc = Fr::mul(acc, d);
        }
    }
    Fr inv = Fr::inv(acc);
#pragma unroll
    for (int k = DIV_BATCH - 1; k >= 0; k--) {
        if (!st[k]) continue;
        Pending &pd = v.pend[s0 + k];
        if (st[k] & 4) v.w[pd.wire] = Fr::zero();
        else { v.w[pd.wire] = Fr::mul(pd.num, Fr::mul(inv, pre[k])); inv = Fr::mul(inv, pd.den); }
        pd.state = 0;
    }
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// levels [l0, l1), each at most a few dozen instructions: one CTA, a barrier per level.
// A level's critical path would be four dependent global loads (schedule entry -> row pointers -> term lists -> wire values) plus
// the arithmetic; the first three are static data, so every warp pulls them into L1 ahead of time: the schedule entries of level
// l+3, the row pointers of level l+2 and the term lists of level l+1 while level l is being solved.  What remains per level is one
// L2 round trip for the wire values the previous level has just written, and the arithmetic -- which for ONE warp per instruction
// is a latency chain of its own: a field product is ~250 dependent instructions of a lane (~800 cycles with nothing else to issue),
// and an 80-term Poseidon row makes every lane do three of them per side, three sides in sequence.  The serial sponge tail of the
// circuit has one or two instructions per level, so the CTA's other warps are idle: there, the three sides of an instruction go to
// different warps and a side's terms to up to NARROW_SUB warps (one product per lane), the partial sums meet in shared memory, and
// one lane finishes (a*b, the subtraction, the store).  Two barriers per level instead of one, a third of the dependent products.
static const int NARROW_SUB = 4;                  // warps per side at most
static const int NARROW_SPLIT_INSTR = NARROW_THREADS / 32 / 3;   // instructions of a level that can be split by side
// static data of the levels after l into L1, by `count` warps of which this is number `rank`: schedule entries of level l+3, row
// pointers of level l+2, term lists of level l+1
__device__ __forceinline__ void narrow_prefetch(const ProgView &v, uint64_t l, uint64_t l1, int rank, int count, int lane) {
    if (l + 3 < l1) {
        const uint64_t s0 = v.lvl_start[l + 3], s1 = v.lvl_start[l + 4];
        for (uint64_t p = s0 + (uint64_t)(rank * 32 + lane) * 32; p < s1; p += (uint64_t)count * 32 * 32) prefetch_l1(v.sched + p);
    }
    if (l + 2 < l1) {
        const uint64_t s0 = v.lvl_start[l + 2], s1 = v.lvl_start[l + 3];
        for (uint64_t p = s0 + rank; p < s1; p += count) {
            const uint32_t packed = v.sched[p];
            if (packed & HINT_BIT) { if (lane == 0) { prefetch_l1(v.hint_in0 + (packed & ~HINT_BIT)); prefetch_l1(v.hint_out + (packed & ~HINT_BIT)); } }
            else if (lane < 3) prefetch_l1(v.ptr[lane] + packed);
            else if (lane == 3) prefetch_l1(v.solve_e + packed);
        }
    }
    if (l + 1 < l1) {
        const uint64_t s0 = v.lvl_start[l + 1], s1 = v.lvl_start[l + 2];
        for (uint64_t p = s0 + rank; p < s1; p += count) {
            const uint32_t packed = v.sched[p];
            if (packed & HINT_BIT) continue;
            for (int side = 0; side < 3; side++) {
                const uint64_t e0 = v.ptr[side][packed], e1 = v.ptr[side][packed + 1];
                for (uint64_t e = (e0 & ~31ull) + (uint64_t)lane * 32; e < e1; e += 32 * 32) { prefetch_l1(v.wire[side] + e); prefetch_l1(v.coef[side] + e); }
            }
        }
    }
}

template <bool DRY>
__global__ void __launch_bounds__(NARROW_THREADS) k_solve_narrow(ProgView v, uint64_t l0, uint64_t l1) {
    __shared__ uint32_t part[NARROW_SPLIT_INSTR][3][NARROW_SUB][12];   // nine limbs of an unreduced partial sum
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // Everything that maps a warp to its role is worked out once: a level must not pay runtime divisions (they were 15 % of the
    // kernel's instructions).  subs: 4 bits per level width k = 1..7 -> warps per side; role[s-1]: this warp's (instruction, side,
    // part) when a side is split over s warps.
    uint32_t subs = 0, role[NARROW_SUB];
    for (int k = 1; k <= 7; k++) subs |= (uint32_t)(DRY || 3 * k > nwarps ? 0 : min(NARROW_SUB, nwarps / (3 * k))) << (4 * k);
#pragma unroll
    for (int sb = 1; sb <= NARROW_SUB; sb++) {
        const int per = 3 * sb, ins = warp / per, rem = warp - ins * per, side = rem / sb, prt = rem - side * sb;
        role[sb - 1] = (uint32_t)ins | ((uint32_t)side << 8) | ((uint32_t)prt << 16) | ((uint32_t)(rem == 0) << 24);
    }
    for (uint64_t l = l0; l < l1; l++) {
        const uint64_t s0 = v.lvl_start[l], s1 = v.lvl_start[l + 1];
        const int k = (int)(s1 - s0);
        const int sub = k <= 7 ? (int)((subs >> (4 * k)) & 15u) : 0;
        if (sub == 0) {
            narrow_prefetch(v, l, l1, warp, nwarps, lane);
            for (uint64_t p = s0 + warp; p < s1; p += nwarps) exec_instr<32, DRY>(v, v.sched[p], lane, 0xFFFFFFFFu, NO_SLOT, NO_SLOT);
            __syncthreads();
            continue;
        }
        // split mode: warp -> (instruction, side, part); the warps beyond the 3 * k * sub working ones prefetch
        const uint32_t rl = sub == 1 ? role[0] : sub == 2 ? role[1] : sub == 3 ? role[2] : role[3];
        const int ins = (int)(rl & 255u), side = (int)((rl >> 8) & 255u), prt = (int)((rl >> 16) & 255u);
        const bool first = (rl >> 24) != 0;
        const int working = 3 * k * sub;
        uint32_t packed = HINT_BIT;
        uint64_t se = SOLVE_NONE;
        if (warp >= working) narrow_prefetch(v, l, l1, warp - working, nwarps - working, lane);
        else {
            packed = v.sched[s0 + ins];
            if (packed & HINT_BIT) { if (first && lane == 0) exec_hint<false>(v, packed & ~HINT_BIT, NO_SLOT, NO_SLOT); }
            else {
                se = v.solve_e[packed];
                if (se != SOLVE_NONE) {
                    const uint64_t skip = se_side(se) == side ? se_pos(se) : SOLVE_NONE;
                    const Wide9 acc = group_sum_w9<32>(lane_dot_w9(v, v.ptr[side], v.wire[side], v.coef[side], packed, skip, prt * 32 + lane, sub * 32), 0xFFFFFFFFu);
                    if (lane == 0) for (int i = 0; i < 9; i++) part[ins][side][prt][i] = acc.l[i];
                }
            }
        }
        __syncthreads();
        if (warp < working && first && !(packed & HINT_BIT) && se != SOLVE_NONE) {
            // lanes 0, 1, 2 of the instruction's first warp total and reduce one side each; lane 0 finishes
            Fr mine = Fr::zero();
            if (lane < 3) {
                Wide9 t = w9_zero();
                for (int q = 0; q < sub; q++) { Wide9 o; for (int i = 0; i < 9; i++) o.l[i] = part[ins][lane][q][i]; w9_addw(t, o); }
                mine = w9_reduce(t);
            }
            Fr b, c;
#pragma unroll
            for (int i = 0; i < 8; i++) { b.l[i] = __shfl_sync(0xFFFFFFFFu, mine.l[i], 1); c.l[i] = __shfl_sync(0xFFFFFFFFu, mine.l[i], 2); }
            if (lane == 0) finish_instr(v, packed, se_side(se), se_pos(se), mine, b, c, NO_SLOT);
        } else if (working == nwarps && !(warp < working && first)) {
            // no spare warp in this level: the warps that are not finishing prefetch while the finishers work
            narrow_prefetch(v, l, l1, warp, nwarps, lane);
        }
        __syncthreads();
    }
}

// ---- narrow levels, software-pipelined --------------------------------------------------------------------------------------------
// In the serial sponge a level is one S-box constraint whose sides are ~80-term linear expressions, but only the terms that read the
// wire solved ONE level earlier (one or two) have to wait for it.  zkpor_program_upload moves those "fresh" terms to the end of each
// side's list (solve_e carries their count); here the other ("old") terms of level l + 1 are summed by eleven warps WHILE the finishing
// warp of level l reduces, multiplies and stores -- so a level's critical path is: the fresh terms' products, a small sum, the finish.
// Levels of one to three instructions run this way; wider ones go warp-per-instruction as in k_solve_narrow.
static const int PIPE_MAX_K = 3, PIPE_FIN0 = 11;     // finisher of instruction i = warp PIPE_FIN0 + i; warps 0..10 sum ahead

// sum of the old terms of level l (k instructions) into oldp[ins][side][part]; warps 0 .. PIPE_FIN0-1
__device__ __forceinline__ void pipe_old_phase(const ProgView &v, uint64_t s0, int k, int warp, int lane, uint32_t (*oldp)[3][3][12]) {
    const int sub = k == 1 ? 3 : 1;
    if (warp >= 3 * k * sub) return;
    const int ins = k == 1 ? 0 : warp / 3, side = k == 1 ? warp / 3 : warp % 3, prt = k == 1 ? warp % 3 : 0;
    const uint32_t packed = v.sched[s0 + ins];
    Wide9 acc = w9_zero();
    if (!(packed & HINT_BIT)) {
        const uint64_t se = v.solve_e[packed];
        if (se != SOLVE_NONE && !(se & SE_ALLFRESH)) {
            const uint64_t e0 = v.ptr[side][packed], e1 = v.ptr[side][(uint64_t)packed + 1] - se_fresh(se, side);
            const uint64_t skip = se_side(se) == side ? se_pos(se) : SOLVE_NONE;
            for (uint64_t e = e0 + prt * 32 + lane; e < e1; e += sub * 32) {
                if (e == skip) continue;
                term_w9(v, acc, v.coef[side][e], v.w[v.wire[side][e]]);
            }
            acc = group_sum_w9<32>(acc, 0xFFFFFFFFu);
        }
    }
    if (lane == 0) for (int i = 0; i < 9; i++) oldp[ins][side][prt][i] = acc.l[i];
}

__global__ void __launch_bounds__(NARROW_THREADS) k_solve_narrow_pipe(ProgView v, uint64_t l0, uint64_t l1) {
    __shared__ uint32_t oldp[2][PIPE_MAX_K][3][3][12];
    __shared__ uint32_t freshp[PIPE_MAX_K][3][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;      // launched with 16 warps
    bool have_old = false;
    int cur = 0;
    uint64_t s0 = v.lvl_start[l0];
    long long pf[6] = {0, 0, 0, 0, 0, 0}, tq = 0;          // prof: [0] fresh work, [1] wait at barrier 1, [2] finish / old work, [3] wait at barrier 2, [4] pipelined levels, [5] other levels' cycles
    const bool prof = v.prof != nullptr && lane == 0 && (warp == 0 || warp == PIPE_FIN0);
    for (uint64_t l = l0; l < l1; l++) {
        const uint64_t s1 = v.lvl_start[l + 1];
        const int k = (int)(s1 - s0);
        if (prof) tq = clock64();
        if (k > PIPE_MAX_K) {
            narrow_prefetch(v, l, l1, warp, nwarps, lane);
            for (uint64_t p = s0 + warp; p < s1; p += nwarps) exec_instr<32, false>(v, v.sched[p], lane, 0xFFFFFFFFu, NO_SLOT, NO_SLOT);
            __syncthreads();
            if (prof) pf[5] += clock64() - tq;
            have_old = false; s0 = s1;
            continue;
        }
        if (!have_old) {                                    // first level of a pipelined stretch: its old terms were not summed ahead
            if (warp < PIPE_FIN0) pipe_old_phase(v, s0, k, warp, lane, oldp[cur]);
            __syncthreads();
        }
        // fresh terms: one warp per (instruction, side); the other warps prefetch the static data of the levels ahead
        uint32_t packed = HINT_BIT;
        uint64_t se = SOLVE_NONE;
        if (warp < 3 * k) {
            const int ins = warp / 3, side = warp % 3;
            packed = v.sched[s0 + ins];
            if (!(packed & HINT_BIT)) {
                se = v.solve_e[packed];
                if (se != SOLVE_NONE) {
                    const uint64_t e1 = v.ptr[side][(uint64_t)packed + 1];
                    const uint64_t e0 = (se & SE_ALLFRESH) ? v.ptr[side][packed] : e1 - se_fresh(se, side);
                    const uint64_t skip = se_side(se) == side ? se_pos(se) : SOLVE_NONE;
                    Wide9 acc = w9_zero();
                    for (uint64_t e = e0 + lane; e < e1; e += 32) {
                        if (e == skip) continue;
                        term_w9(v, acc, v.coef[side][e], v.w[v.wire[side][e]]);
                    }
                    if (e1 - e0 > 1) acc = group_sum_w9<32>(acc, 0xFFFFFFFFu);
                    if (lane == 0) for (int i = 0; i < 9; i++) freshp[ins][side][i] = acc.l[i];
                }
            }
        } else narrow_prefetch(v, l, l1, warp - 3 * k, nwarps - 3 * k, lane);
        if (prof) { const long long t = clock64(); pf[0] += t - tq; tq = t; }
        __syncthreads();
        if (prof) { const long long t = clock64(); pf[1] += t - tq; tq = t; }
        // finish level l (warps PIPE_FIN0 ..) while warps 0 .. PIPE_FIN0-1 sum the old terms of level l + 1
        const uint64_t s2 = l + 1 < l1 ? v.lvl_start[l + 2] : s1;
        const int k1 = (int)(s2 - s1);
        const bool pipe1 = l + 1 < l1 && k1 <= PIPE_MAX_K && k1 > 0;
        if (warp >= PIPE_FIN0 && warp - PIPE_FIN0 < k) {
            const int ins = warp - PIPE_FIN0;
            const uint32_t pk = v.sched[s0 + ins];
            if (pk & HINT_BIT) { if (lane == 0) exec_hint<false>(v, pk & ~HINT_BIT, NO_SLOT, NO_SLOT); }
            else {
                const uint64_t sq = v.solve_e[pk];
                if (sq != SOLVE_NONE) {
                    Fr mine = Fr::zero();
                    if (lane < 3) {
                        Wide9 t;
                        for (int i = 0; i < 9; i++) t.l[i] = freshp[ins][lane][i];
                        const int sub = k == 1 ? 3 : 1;
                        for (int q = 0; q < sub; q++) { Wide9 o; for (int i = 0; i < 9; i++) o.l[i] = oldp[cur][ins][lane][q][i]; w9_addw(t, o); }
                        mine = w9_reduce(t);
                    }
                    Fr b, c;
#pragma unroll
                    for (int i = 0; i < 8; i++) { b.l[i] = __shfl_sync(0xFFFFFFFFu, mine.l[i], 1); c.l[i] = __shfl_sync(0xFFFFFFFFu, mine.l[i], 2); }
                    if (lane == 0) finish_instr(v, pk, se_side(sq), se_pos(sq), mine, b, c, NO_SLOT);
                }
            }
        } else if (warp < PIPE_FIN0 && pipe1) pipe_old_phase(v, s1, k1, warp, lane, oldp[cur ^ 1]);
        if (prof) { const long long t = clock64(); pf[2] += t - tq; tq = t; }
        __syncthreads();
        if (prof) { pf[3] += clock64() - tq; pf[4]++; }
        have_old = pipe1; cur ^= 1; s0 = s1;
    }
    if (prof) for (int i = 0; i < 6; i++) atomicAdd(v.prof + (warp == 0 ? 0 : 6) + i, (unsigned long long)pf[i]);
}

// upload: wlevel[wire] = 1 + the level that solves it, for the wires solved in narrow levels [la, lb); one thread per level
__global__ void k_narrow_wlevel(ProgView v, uint64_t la, uint64_t lb, uint32_t *wlevel) {
    const uint64_t l = la + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= lb) return;
    for (uint64_t p = v.lvl_start[l], p1 = v.lvl_start[l + 1]; p < p1; p++) {
        const uint32_t packed = v.sched[p];
        if (packed & HINT_BIT) {
            const uint32_t h = packed & ~HINT_BIT;
            for (uint32_t q = 0; q < v.hint_nout[h]; q++) wlevel[v.hint_out[h] + q] = (uint32_t)(l + 1);
        } else {
            const uint64_t se = v.solve_e[packed];
            if (se != SOLVE_NONE) wlevel[v.wire[se_side(se)][se_pos(se)]] = (uint32_t)(l + 1);
        }
    }
}
// upload: in every row of the narrow levels [la, lb) the terms that read a wire solved one level earlier go to the end of their side's
// list (a sum does not depend on the order of its terms); their count and the unknown's new position are recorded in solve_e
__global__ void k_narrow_partition(ProgView v, uint64_t la, uint64_t lb, const uint32_t *__restrict__ wlevel) {
    const uint64_t l = la + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= lb) return;
    for (uint64_t p = v.lvl_start[l], p1 = v.lvl_start[l + 1]; p < p1; p++) {
        const uint32_t packed = v.sched[p];
        if (packed & HINT_BIT) continue;
        uint64_t se = v.solve_e[packed];
        if (se == SOLVE_NONE) continue;
        const int us = se_side(se);
        uint64_t upos = se_pos(se), out = (uint64_t)us << 62;
        for (int side = 0; side < 3; side++) {
            uint32_t *wi = const_cast<uint32_t *>(v.wire[side]), *ci = const_cast<uint32_t *>(v.coef[side]);
            uint64_t i = v.ptr[side][packed], j = v.ptr[side][(uint64_t)packed + 1];
            uint64_t nf = 0;
            while (i < j) {
                const bool fresh = !(side == us && i == upos) && wlevel[wi[i]] == (uint32_t)l;
                if (!fresh) { i++; continue; }
                j--;
                const uint32_t tw = wi[i], tc = ci[i]; wi[i] = wi[j]; ci[i] = ci[j]; wi[j] = tw; ci[j] = tc;
                if (side == us) { if (upos == j) upos = i; }        // the unknown's term sat at j and moved to i (i itself is fresh, never the unknown)
                nf++;
            }
            if (nf > 255) out |= SE_ALLFRESH; else out |= nf << (36 + 8 * side);
        }
        v.solve_e[packed] = out | upos;
    }
}

__global__ void k_count_queries(ProgView v, uint64_t r0, uint64_t r1, uint32_t n_out, uint32_t *cnt, uint32_t hint) {
    const uint64_t r = r0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= r1) return;
    const Fr q = Fr::from_mont(aux_eval(v, r));
    bool ok = true;
    for (int i = 1; i < 8; i++) ok &= q.l[i] == 0;
    if (!ok || q.l[0] >= n_out) { solve_fail(v, SE_INDEX, hint); return; }
    atomicAdd(cnt + q.l[0], 1u);
}
__global__ void k_count_store(ProgView v, const uint32_t *cnt, uint32_t out, uint32_t n_out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_out) v.w[out + k] = Fr::from_u64(cnt[k]);
}
__global__ void k_mark_solved(uint8_t *solved, uint32_t *wstep, uint32_t step, uint64_t first, uint64_t n) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { solved[first + k] = 1; wstep[first + k] = step + 1; }
}

// Upload-time analysis of the deferred tail, one thread per schedule entry of the step: which wires does the step solve (appended to
// out_wires when given; *cursor counts them), and which is the latest earlier step it reads from (dep = max of wstep over its reads).
__global__ void k_tail_scan(ProgView v, uint64_t p0, uint64_t p1, uint32_t tail_mark, uint32_t *out_wires, unsigned long long *cursor, uint32_t *dep) {
    const uint64_t p = p0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    const uint32_t packed = v.sched[p];
    uint32_t d = 0;
    auto reads = [&](uint32_t wire) { const uint32_t ws = v.wstep[wire]; if (ws != tail_mark && ws > d) d = ws; };
    auto aux_rows = [&](uint64_t r0, uint64_t r1) { for (uint64_t e = v.aux_ptr[r0], e1 = v.aux_ptr[r1]; e < e1; e++) reads(v.aux_wire[e]); };
    if (packed & HINT_BIT) {
        const uint32_t h = packed & ~HINT_BIT, n_out = v.hint_nout[h], out = v.hint_out[h];
        aux_rows(v.hint_in0[h], v.hint_in1[h]);
        if (v.hint_fn[h] == ZKPOR_HINT_LOOKUP) aux_rows(v.table_ptr[v.hint_param[h]], v.table_ptr[v.hint_param[h] + 1]);
        const unsigned long long at = atomicAdd(cursor, (unsigned long long)n_out);
        if (out_wires) for (uint32_t k = 0; k < n_out; k++) out_wires[at + k] = out + k;
    } else {
        for (int side = 0; side < 3; side++)
            for (uint64_t e = v.ptr[side][packed], e1 = v.ptr[side][(uint64_t)packed + 1]; e < e1; e++) reads(v.wire[side][e]);
        const uint64_t se = v.solve_e[packed];
        if (se != SOLVE_NONE) {
            const unsigned long long at = atomicAdd(cursor, 1ull);
            if (out_wires) out_wires[at] = v.wire[se_side(se)][se_pos(se)];
        }
    }
    if (d) atomicMax(dep, d);
}
__global__ void k_wire_mask(const uint32_t *__restrict__ wires, uint64_t n, uint32_t *mask) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicOr(mask + (wires[i] >> 5), 1u << (wires[i] & 31));
}

__global__ void k_check_abc(const Fr *__restrict__ a, const Fr *__restrict__ b, const Fr *__restrict__ c, uint64_t n, unsigned long long *first_bad) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (Fr::mul(a[k], b[k]) != c[k]) atomicMin(first_bad, (unsigned long long)k);
}

int32_t r1cs_check_dev(zkpor_ctx *ctx, const Fr *d_a, const Fr *d_b, const Fr *d_c, uint64_t n_rows) {
    ZK_TRY(ctx->misc.reserve(64));
    unsigned long long *bad = ctx->misc.as<unsigned long long>();
    ZK_CUDA(cudaMemsetAsync(bad, 0xFF, 8, ctx->stream));
    ZK_LAUNCH(ctx, k_check_abc, grid_for(n_rows, 256), 256, 0, d_a, d_b, d_c, n_rows, bad);
    unsigned long long h = 0;
    ZK_CUDA(cudaMemcpyAsync(&h, bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h != ~0ull) { set_error("solve: constraint #%llu is not satisfied", h); return ZKPOR_ERR_STATE; }
    return ZKPOR_OK;
}

static ProgView make_view(zkpor_program *p, Fr *w, uint8_t *solved) {
    ProgView v;
    for (int m = 0; m < 3; m++) { v.ptr[m] = p->cs->row_ptr[m]; v.wire[m] = p->cs->wire_ids[m]; v.coef[m] = p->cs->coeff_ids[m]; }
    v.aux_ptr = p->aux_ptr; v.aux_wire = p->aux_wire; v.aux_coef = p->aux_coef;
    v.coeffs = p->cs->coeffs; v.one_id = p->cs->one_id; v.minus_one_id = p->minus_one_id;
    v.sched = p->sched; v.lvl_start = p->lvl_start;
    v.hint_fn = p->hint_fn; v.hint_param = p->hint_param; v.hint_out = p->hint_out; v.hint_nout = p->hint_nout; v.hint_in0 = p->hint_in0; v.hint_in1 = p->hint_in1;
    v.table_ptr = p->table_ptr; v.solve_e = p->solve_e; v.w = w; v.solved = solved; v.err = p->err;
    v.pend = p->pend; v.step_div = p->step_div; v.wstep = p->wstep; v.cur_step = 0; v.prof = p->prof;
    return v;
}

static int32_t solve_error(zkpor_program *p, zkpor_ctx *ctx, const char *what) {
    unsigned long long e = 0;
    ZK_CUDA(cudaMemcpyAsync(&e, p->err, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e == 0) return ZKPOR_OK;
    const int code = (int)(e >> 56); const unsigned long long where = e & ((1ull << 56) - 1);
    const char *msg = code == SE_UNSOLVED ? "constraint has more than one unsolved wire (levels are not a valid schedule)"
                    : code == SE_DIV0 ? "division by zero" : code == SE_INDEX ? "lookup / multiplicity index outside its table" : "unknown hint function";
    set_error("%s: %s (constraint row or hint record #%llu)", what, msg, where);
    return ZKPOR_ERR_STATE;
}

// runs the schedule: DRY = find every R1C instruction's unknown on solved-flags, else solve
// defer_tail (proofs only): the program's tail step -- a long run of narrow levels at the end of the schedule: in the reference circuit
// the serial sponge of the CEX commitment, ~170 000 levels of one to thirteen instructions, 70 % of the solve on one SM -- is launched
// on the context's high-priority side stream as soon as the steps it reads from are enqueued, and the caller goes on with everything
// that is linear in the wire vector (the A, B, K multiplications, with the tail's wires masked to zero) before solver_tail_join.
template <bool DRY>
static int32_t run_schedule(zkpor_ctx *ctx, zkpor_program *p, zkpor_pk *pk, Fr *w, uint8_t *solved, G1XYZZ *commit, G1XYZZ *pok, bool defer_tail = false) {
    ProgView v = make_view(p, w, solved);
    ZK_CUDA(cudaMemsetAsync(p->err, 0, 8, ctx->stream));
    const bool trace = !DRY && p->trace_path != nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace) { tev.resize(p->steps.size() + 1); for (auto &e : tev) cudaEventCreate(&e); }
    defer_tail = defer_tail && !DRY && p->tail_step >= 0;
    for (size_t si = 0; si < p->steps.size(); si++) {
        const Step &s = p->steps[si];
        if (trace) cudaEventRecord(tev[si], ctx->stream);
        v.cur_step = (uint32_t)si;
        if (defer_tail && si == p->tail_after) {
            const Step &t = p->steps[(size_t)p->tail_step];
            ZK_CUDA(cudaEventRecord(p->tail_go, ctx->stream));
            ZK_CUDA(cudaStreamWaitEvent(ctx->tail_stream, p->tail_go, 0));
            cudaStream_t main_stream = ctx->stream;
            ctx->stream = ctx->tail_stream;
            int32_t rc = ZKPOR_OK;
            { KTimed kt(ctx, KC_SOLVE_NARROW, t.b - t.a);
              if (p->narrow_pipe) k_solve_narrow_pipe<<<1, NARROW_THREADS, 0, ctx->stream>>>(v, t.a, t.b);
              else k_solve_narrow<false><<<1, p->narrow_threads, 0, ctx->stream>>>(v, t.a, t.b);
              ctx->launches++;
              if (cudaGetLastError() != cudaSuccess) { set_error("solve: launch of the deferred tail failed"); rc = ZKPOR_ERR_CUDA; }
              kt.stop(); }
            cudaEventRecord(p->tail_done, ctx->stream);
            ctx->stream = main_stream;
            ZK_TRY(rc);
            p->tail_running = true;
        }
        if (defer_tail && (int64_t)si == p->tail_step) continue;
        switch (s.kind) {
        case STEP_WIDE: {
            const uint64_t count = s.b - s.a, blocks = (s.n_long + 7) / 8 + (((count - s.n_long) * WIDE_GS + 255) / 256);
            KTimed kt(ctx, KC_SOLVE_WIDE, DRY ? 0 : count);
            ZK_LAUNCH(ctx, k_solve_wide<DRY>, (int)blocks, 256, 0, v, s.a, s.n_long, count, (uint64_t)si);
            if (!DRY && s.has_div) ZK_LAUNCH(ctx, k_solve_div, grid_for((count + DIV_BATCH - 1) / DIV_BATCH, 128), 128, 0, v, count);
            kt.stop();
            break;
        }
        case STEP_NARROW: {
            KTimed kt(ctx, KC_SOLVE_NARROW, DRY ? 0 : s.b - s.a);
            if (!DRY && p->narrow_pipe) ZK_LAUNCH(ctx, k_solve_narrow_pipe, 1, NARROW_THREADS, 0, v, s.a, s.b);
            else ZK_LAUNCH(ctx, k_solve_narrow<DRY>, 1, p->narrow_threads, 0, v, s.a, s.b);
            kt.stop();
            break;
        }
        case STEP_COUNT: {
            const uint32_t h = (uint32_t)s.a, n_out = p->h_hint_nout[h], out = p->h_hint_out[h];
            if (DRY) { ZK_LAUNCH(ctx, k_mark_solved, grid_for(n_out, 256), 256, 0, solved, v.wstep, (uint32_t)si, (uint64_t)out, (uint64_t)n_out); break; }
            const uint64_t r0 = p->h_hint_in0[h], r1 = p->h_hint_in1[h];
            ZK_CUDA(cudaMemsetAsync(p->counters, 0, (size_t)n_out * 4, ctx->stream));
            if (r1 > r0) ZK_LAUNCH(ctx, k_count_queries, grid_for(r1 - r0, 256), 256, 0, v, r0, r1, n_out, p->counters, h);
            ZK_LAUNCH(ctx, k_count_store, grid_for(n_out, 256), 256, 0, v, (const uint32_t *)p->counters, out, n_out);
            break;
        }
        case STEP_COMMIT: {
            const uint32_t h = (uint32_t)s.a, out = p->h_hint_out[h];
            if (DRY) { ZK_LAUNCH(ctx, k_mark_solved, grid_for(1, 32), 32, 0, solved, v.wstep, (uint32_t)si, (uint64_t)out, (uint64_t)1); break; }
            // Prove's override of the BSB22 placeholder: Pedersen-commit the private committed wires, hash the
            // commitment to the field; the proof of knowledge shares the sort of the committed values, so it is taken here as well
            if (pk == nullptr || !pk->has_commitment) { set_error("solve: the program has a commitment hint but no proving key with a commitment key was given"); return ZKPOR_ERR_INVALID_ARG; }
            ZK_TRY(pk_commit_and_pok(ctx, pk, w, commit, pok));
            const Fr ch = commitment_challenge_g1(commit->to_affine());
            ZK_CUDA(cudaMemcpyAsync(w + out, &ch, 32, cudaMemcpyHostToDevice, ctx->stream));
            ZK_CUDA(cudaStreamSynchronize(ctx->stream));   // `ch` lives on this frame
            break;
        }
        }
    }
    if (trace) {
        cudaEventRecord(tev[p->steps.size()], ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        if (FILE *f = fopen(p->trace_path, "w")) {
            fprintf(f, "step,kind,count,n_long,has_div,ms\n");
            for (size_t si = 0; si < p->steps.size(); si++) {
                float ms = 0.f; cudaEventElapsedTime(&ms, tev[si], tev[si + 1]);
                const Step &s = p->steps[si];
                fprintf(f, "%zu,%d,%llu,%llu,%d,%.4f\n", si, s.kind, (unsigned long long)(s.kind <= STEP_NARROW ? s.b - s.a : 1), (unsigned long long)s.n_long, (int)s.has_div, ms);
            }
            fclose(f);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    return solve_error(p, ctx, DRY ? "program_upload" : "solve");
}

int32_t solver_run(zkpor_ctx *ctx, zkpor_program *prog, zkpor_pk *pk, Fr *d_wires, G1XYZZ *commit, G1XYZZ *pok, bool *has_commit, bool defer_tail) {
    *has_commit = prog->has_commit;
    return run_schedule<false>(ctx, prog, pk, d_wires, nullptr, commit, pok, defer_tail);
}
bool solver_tail_info(const zkpor_program *prog, SolverTail *out) {
    if (prog->tail_step < 0) return false;
    out->uid = prog->uid; out->mask = prog->tail_mask; out->wires = &prog->h_tail_wires; out->levels = prog->steps[(size_t)prog->tail_step].b - prog->steps[(size_t)prog->tail_step].a;
    return true;
}
// the compute stream waits for the tail; its errors surface here.  A no-op when no tail is in flight.
int32_t solver_tail_join(zkpor_ctx *ctx, zkpor_program *prog) {
    if (!prog->tail_running) return ZKPOR_OK;
    prog->tail_running = false;
    ZK_CUDA(cudaStreamWaitEvent(ctx->stream, prog->tail_done, 0));
    return solve_error(prog, ctx, "solve");
}
// error paths: nothing of the program may still be running when the caller returns
void solver_tail_abandon(zkpor_ctx *ctx, zkpor_program *prog) {
    if (prog && prog->tail_running) { cudaStreamSynchronize(ctx->tail_stream); prog->tail_running = false; }
}
zkpor_r1cs *program_matrices(zkpor_program *prog) { return prog->cs; }
uint64_t program_inputs(const zkpor_program *prog) { return prog->n_public - 1 + prog->n_secret; }

// host copy of an array that may live on either side
template <class T>
static int32_t fetch(std::vector<T> &dst, const T *src, size_t n) {
    dst.resize(n);
    if (n) ZK_CUDA(cudaMemcpy(dst.data(), src, n * sizeof(T), cudaMemcpyDefault));
    return ZKPOR_OK;
}
template <class T>
static int32_t put(T **dst, const T *src, size_t n) {
    *dst = nullptr;
    ZK_CUDA(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) ZK_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyDefault));
    return ZKPOR_OK;
}

}  // namespace zk

using namespace zk;

extern "C" {

int32_t zkpor_program_free(zkpor_ctx *ctx, zkpor_program *p) {
    if (!p) return ZKPOR_OK;
    if (p->cs) zkpor_r1cs_free(ctx, p->cs);
    void *ptrs[] = {p->aux_ptr, p->aux_wire, p->aux_coef, p->sched, p->lvl_start, p->hint_fn, p->hint_param, p->hint_out, p->hint_nout,
                    p->hint_in0, p->hint_in1, p->table_ptr, p->solve_e, p->err, p->counters, p->pend, p->step_div};
    for (void *q : ptrs) if (q) cudaFree(q);
    if (p->tail_running) cudaStreamSynchronize(ctx->tail_stream);
    if (p->prof) {
        unsigned long long h[12];
        cudaDeviceSynchronize();
        if (cudaMemcpy(h, p->prof, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int wv = 0; wv < 2; wv++) {
                const double n = (double)std::max<unsigned long long>(h[6 * wv + 4], 1);
                fprintf(stderr, "narrow prof, %s: per pipelined level: work A %.0f | wait 1 %.0f | work B %.0f | wait 2 %.0f cycles (%llu levels); other levels %.0f cycles in all\n",
                        wv == 0 ? "warp 0 (fresh terms, old terms)" : "finisher warp", h[6 * wv] / n, h[6 * wv + 1] / n, h[6 * wv + 2] / n, h[6 * wv + 3] / n, h[6 * wv + 4], (double)h[6 * wv + 5]);
            }
        cudaFree(p->prof);
    }
    for (void *q : {(void *)p->tail_wires, (void *)p->tail_mask, (void *)p->wstep}) if (q) cudaFree(q);
    if (p->tail_go) cudaEventDestroy(p->tail_go);
    if (p->tail_done) cudaEventDestroy(p->tail_done);
    p->wires.release(); p->abc.release();
    delete p;
    return ZKPOR_OK;
}

int32_t zkpor_program_stats(zkpor_program *prog, uint64_t out4[4]) {
    ZK_REQUIRE(prog && out4, "program_stats: null argument");
    for (int i = 0; i < 4; i++) out4[i] = prog->stats[i];
    return ZKPOR_OK;
}

int32_t zkpor_program_tail_info(zkpor_program *prog, uint64_t out3[3]) {
    ZK_REQUIRE(prog && out3, "program_tail_info: null argument");
    out3[0] = out3[1] = out3[2] = 0;
    if (prog->tail_step >= 0) { const zk::Step &t = prog->steps[(size_t)prog->tail_step]; out3[0] = t.b - t.a; out3[1] = prog->n_tail_wires; out3[2] = prog->tail_after; }
    return ZKPOR_OK;
}

int32_t zkpor_program_r1cs(zkpor_program *prog, zkpor_r1cs **out) {
    ZK_REQUIRE(prog && out, "program_r1cs: null argument");
    *out = prog->cs;
    return ZKPOR_OK;
}

int32_t zkpor_program_tail_wires(zkpor_program *prog, uint32_t *out_wires, uint64_t cap) {
    ZK_REQUIRE(prog && (out_wires || cap == 0), "program_tail_wires: null argument");
    ZK_REQUIRE(cap >= prog->h_tail_wires.size(), "program_tail_wires: buffer too small (zkpor_program_tail_info gives the count)");
    if (!prog->h_tail_wires.empty()) memcpy(out_wires, prog->h_tail_wires.data(), prog->h_tail_wires.size() * 4);
    return ZKPOR_OK;
}

int32_t zkpor_program_upload(zkpor_ctx *ctx, const zkpor_program_desc *d, zkpor_program **out) {
    ZK_REQUIRE(ctx && d && out, "program_upload: null argument");
    ZK_REQUIRE(d->n_public >= 1 && d->n_public + d->n_secret <= d->n_wires && d->n_wires < (1ull << 32), "program_upload: wire counts out of range");
    ZK_REQUIRE(d->n_instr > 0 && d->n_instr < (1ull << 32) && d->n_constraints < (1ull << 31) && d->n_hints < (1ull << 31), "program_upload: sizes out of range");
    ZK_REQUIRE(d->instr_kind && d->instr_arg && d->level_ptr && d->level_instr, "program_upload: null instruction arrays");
    ZK_REQUIRE(d->n_hints == 0 || (d->hint_fn && d->hint_param && d->hint_out_first && d->hint_n_out && d->hint_in_ptr && d->hint_in_end), "program_upload: null hint arrays");
    ZK_REQUIRE(d->n_aux_rows == 0 || d->aux.row_ptr, "program_upload: null auxiliary matrix");
    ZK_REQUIRE(d->n_tables == 0 || d->table_ptr, "program_upload: null table_ptr");
    ZK_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    zkpor_program *p = new zkpor_program();
    p->n_wires = d->n_wires; p->n_public = d->n_public; p->n_secret = d->n_secret; p->n_rows = d->n_constraints; p->n_instr = d->n_instr;
    p->n_levels = d->n_levels; p->n_hints = d->n_hints; p->n_aux_rows = d->n_aux_rows; p->n_tables = d->n_tables;
    int32_t rc = zkpor_r1cs_upload(ctx, d->n_constraints, d->n_wires, &d->l, &d->r, &d->o, d->coeff_table, d->n_coeffs, &p->cs);
    auto fail = [&](int32_t code) { zkpor_program_free(ctx, p); return code; };
    if (rc != ZKPOR_OK) return fail(rc);
    // the coefficient -1, like 1, skips the product
    {
        std::vector<Fr> tab;
        if ((rc = fetch(tab, (const Fr *)d->coeff_table, d->n_coeffs)) != ZKPOR_OK) return fail(rc);
        const Fr one = Fr::one(), m1 = Fr::neg(Fr::one());
        for (uint64_t i = 0; i < d->n_coeffs; i++) {
            if (tab[i] == m1 && p->minus_one_id == 0xFFFFFFFFu) p->minus_one_id = (uint32_t)i;
            if (tab[i] == one && p->cs->one_id == 0xFFFFFFFFu) p->cs->one_id = (uint32_t)i;
        }
    }
    // instruction-level arrays come to the host (O(instructions)); the matrices (O(terms)) never do
    std::vector<uint8_t> kind; std::vector<uint32_t> arg, lvl_instr, hparam; std::vector<uint64_t> lvl_ptr, tptr;
    if ((rc = fetch(kind, d->instr_kind, d->n_instr)) != ZKPOR_OK || (rc = fetch(arg, d->instr_arg, d->n_instr)) != ZKPOR_OK ||
        (rc = fetch(lvl_instr, d->level_instr, d->n_instr)) != ZKPOR_OK || (rc = fetch(lvl_ptr, d->level_ptr, d->n_levels + 1)) != ZKPOR_OK ||
        (rc = fetch(p->h_hint_fn, d->hint_fn, d->n_hints)) != ZKPOR_OK || (rc = fetch(p->h_hint_nout, d->hint_n_out, d->n_hints)) != ZKPOR_OK ||
        (rc = fetch(p->h_hint_out, d->hint_out_first, d->n_hints)) != ZKPOR_OK || (rc = fetch(p->h_hint_in0, d->hint_in_ptr, d->n_hints)) != ZKPOR_OK ||
        (rc = fetch(p->h_hint_in1, d->hint_in_end, d->n_hints)) != ZKPOR_OK || (rc = fetch(hparam, d->hint_param, d->n_hints)) != ZKPOR_OK ||
        (rc = fetch(tptr, d->table_ptr, d->n_tables ? d->n_tables + 1 : 0)) != ZKPOR_OK)
        return fail(rc);
    auto bad = [&](const char *m) { set_error("program_upload: %s", m); return fail(ZKPOR_ERR_INVALID_ARG); };
    if (lvl_ptr[0] != 0 || lvl_ptr[d->n_levels] != d->n_instr) return bad("level_ptr does not span the instructions");
    uint64_t max_out = 1;
    for (uint64_t h = 0; h < d->n_hints; h++) {
        if ((uint64_t)p->h_hint_out[h] + p->h_hint_nout[h] > d->n_wires) return bad("hint outputs out of range");
        if (p->h_hint_in0[h] > p->h_hint_in1[h] || p->h_hint_in1[h] > d->n_aux_rows) return bad("hint inputs out of range");
        const uint32_t fn = p->h_hint_fn[h];
        if (fn < ZKPOR_HINT_DIVMOD || fn > ZKPOR_HINT_COMMIT) return bad("unknown hint function id");
        if ((fn == ZKPOR_HINT_DIVMOD || fn == ZKPOR_HINT_CMP) && (p->h_hint_in1[h] - p->h_hint_in0[h] != 2 || p->h_hint_nout[h] != (fn == ZKPOR_HINT_DIVMOD ? 2u : 1u))) return bad("DIVMOD / CMP hint arity");
        if ((fn == ZKPOR_HINT_NBITS || fn == ZKPOR_HINT_DECOMPOSE || fn == ZKPOR_HINT_INVZERO) && p->h_hint_in1[h] - p->h_hint_in0[h] != 1) return bad("single-input hint arity");
        if (fn == ZKPOR_HINT_DECOMPOSE && (hparam[h] == 0 || hparam[h] > 32)) return bad("DECOMPOSE limb width must be 1..32 bits");
        if (fn == ZKPOR_HINT_LOOKUP && (hparam[h] >= d->n_tables || p->h_hint_in1[h] - p->h_hint_in0[h] != p->h_hint_nout[h])) return bad("LOOKUP table id / arity");
        if (fn == ZKPOR_HINT_COUNT) max_out = std::max<uint64_t>(max_out, p->h_hint_nout[h]);
        if (fn == ZKPOR_HINT_COMMIT) { if (p->h_hint_nout[h] != 1) return bad("COMMIT hint has one output"); p->has_commit = true; }
    }
    for (uint64_t t = 0; t < d->n_tables; t++) if (tptr[t] > tptr[t + 1] || tptr[t + 1] > d->n_aux_rows) return bad("table_ptr out of range");
    if (const char *e = getenv("ZKPOR_NARROW_MAX")) p->narrow_max = (uint32_t)std::max(1, atoi(e));
    if (const char *e = getenv("ZKPOR_NARROW_THREADS")) p->narrow_threads = std::min(NARROW_THREADS, std::max(32, atoi(e) & ~31));
    p->trace_path = getenv("ZKPOR_SOLVE_TRACE");
    if (getenv("ZKPOR_NARROW_PROF")) { if (cudaMalloc((void **)&p->prof, 12 * 8) == cudaSuccess) cudaMemset(p->prof, 0, 12 * 8); else p->prof = nullptr; }
    p->long_row = WIDE_LONG_ROW;
    if (const char *e = getenv("ZKPOR_WIDE_LONG_ROW")) p->long_row = (uint32_t)std::max(0, atoi(e));   // 0: every row takes a warp
    // schedule: instructions in level order, special hints lifted out as steps of their own
    std::vector<uint32_t> sched; sched.reserve(d->n_instr);
    std::vector<uint64_t> lvl_start; lvl_start.reserve(d->n_levels + 1);
    int64_t run_first = -1;
    auto close_run = [&](uint64_t end_level) {
        if (run_first >= 0) { p->steps.push_back({STEP_NARROW, (uint64_t)run_first, end_level, false, 0}); p->stats[1]++; p->stats[2] += end_level - run_first; run_first = -1; }
    };
    for (uint64_t l = 0; l < d->n_levels; l++) {
        lvl_start.push_back(sched.size());
        bool special = false;
        for (uint64_t q = lvl_ptr[l]; q < lvl_ptr[l + 1]; q++) {
            const uint32_t ins = lvl_instr[q];
            if (ins >= d->n_instr) return bad("level_instr out of range");
            if (kind[ins] == ZKPOR_INS_R1C) { if (arg[ins] >= d->n_constraints) return bad("instruction row out of range"); sched.push_back(arg[ins]); continue; }
            if (kind[ins] != ZKPOR_INS_HINT || arg[ins] >= d->n_hints) return bad("instruction kind / hint id out of range");
            const uint32_t fn = p->h_hint_fn[arg[ins]];
            if (fn == ZKPOR_HINT_COUNT || fn == ZKPOR_HINT_COMMIT) {
                if (!special) close_run(l);
                special = true;
                p->steps.push_back({fn == ZKPOR_HINT_COUNT ? STEP_COUNT : STEP_COMMIT, arg[ins], 0, false, 0});
                if (fn == ZKPOR_HINT_COUNT) p->stats[3]++;
            } else sched.push_back(arg[ins] | HINT_BIT);
        }
        const uint64_t n_l = sched.size() - lvl_start.back();
        if (n_l == 0) continue;
        if (n_l <= p->narrow_max) { if (run_first < 0) run_first = (int64_t)l; }
        else { close_run(l); p->steps.push_back({STEP_WIDE, lvl_start.back(), (uint64_t)sched.size(), false, 0}); p->stats[0]++; p->pend_cap = std::max<uint64_t>(p->pend_cap, n_l); }
    }
    lvl_start.push_back(sched.size());
    close_run(d->n_levels);
    // device copies
    if ((rc = put(&p->sched, sched.data(), sched.size())) != ZKPOR_OK || (rc = put(&p->lvl_start, lvl_start.data(), lvl_start.size())) != ZKPOR_OK ||
        (rc = put(&p->hint_fn, d->hint_fn, d->n_hints)) != ZKPOR_OK || (rc = put(&p->hint_param, d->hint_param, d->n_hints)) != ZKPOR_OK ||
        (rc = put(&p->hint_out, d->hint_out_first, d->n_hints)) != ZKPOR_OK || (rc = put(&p->hint_nout, d->hint_n_out, d->n_hints)) != ZKPOR_OK ||
        (rc = put(&p->hint_in0, d->hint_in_ptr, d->n_hints)) != ZKPOR_OK || (rc = put(&p->hint_in1, d->hint_in_end, d->n_hints)) != ZKPOR_OK ||
        (rc = put(&p->table_ptr, tptr.data(), tptr.size())) != ZKPOR_OK ||
        (rc = put(&p->aux_ptr, d->aux.row_ptr, d->n_aux_rows + 1)) != ZKPOR_OK || (rc = put(&p->aux_wire, d->aux.wire_ids, d->aux.nnz)) != ZKPOR_OK ||
        (rc = put(&p->aux_coef, d->aux.coeff_ids, d->aux.nnz)) != ZKPOR_OK)
        return fail(rc);
    p->counters_cap = max_out;
    if (cudaMalloc((void **)&p->solve_e, std::max<uint64_t>(d->n_constraints, 1) * 8) != cudaSuccess || cudaMalloc((void **)&p->err, 8) != cudaSuccess ||
        cudaMalloc((void **)&p->counters, max_out * 4) != cudaSuccess) { set_error("program_upload: out of device memory"); return fail(ZKPOR_ERR_OOM); }
    if (cudaMalloc((void **)&p->pend, std::max<uint64_t>(p->pend_cap, 1) * sizeof(Pending)) != cudaSuccess ||
        cudaMalloc((void **)&p->step_div, p->steps.size() + 1) != cudaSuccess) { set_error("program_upload: out of device memory"); return fail(ZKPOR_ERR_OOM); }
    cudaMemsetAsync(p->pend, 0, std::max<uint64_t>(p->pend_cap, 1) * sizeof(Pending), ctx->stream);
    cudaMemsetAsync(p->step_div, 0, p->steps.size() + 1, ctx->stream);
    // wide levels: long rows first (k_solve_wide gives them a warp each, the short ones WIDE_GS lanes)
    {
        uint8_t *d_cls = nullptr;
        std::vector<uint8_t> cls;
        if (cudaMalloc((void **)&d_cls, std::max<size_t>(sched.size(), 1)) != cudaSuccess) { set_error("program_upload: out of device memory"); return fail(ZKPOR_ERR_OOM); }
        ProgView v0 = make_view(p, nullptr, nullptr);
        k_row_class<<<grid_for(sched.size(), 256), 256, 0, ctx->stream>>>(v0, (uint64_t)sched.size(), d_cls);
        cudaStreamSynchronize(ctx->stream);
        rc = fetch(cls, (const uint8_t *)d_cls, sched.size());
        cudaFree(d_cls);
        if (rc != ZKPOR_OK) return fail(rc);
        std::vector<uint32_t> tmp;
        for (Step &st : p->steps) {
            if (st.kind != STEP_WIDE) continue;
            tmp.clear();
            for (uint64_t i = st.a; i < st.b; i++) if (p->long_row == 0 || cls[i] >= p->long_row) tmp.push_back(sched[i]);
            st.n_long = tmp.size();
            for (uint64_t i = st.a; i < st.b; i++) if (!(p->long_row == 0 || cls[i] >= p->long_row)) tmp.push_back(sched[i]);
            std::copy(tmp.begin(), tmp.end(), sched.begin() + st.a);
            p->stats_long += st.n_long;
        }
        if (cudaMemcpy(p->sched, sched.data(), sched.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("program_upload: schedule copy failed"); return fail(ZKPOR_ERR_CUDA); }
    }
    {   // table of the unreduced sums' reduction (w9_reduce)
        Fr tab[64];
        for (uint64_t k = 0; k < 64; k++) tab[k] = Fr::from_u64(k);
        if (cudaMemcpyToSymbol(c_hi_mont, tab, sizeof tab) != cudaSuccess) { set_error("program_upload: constant table upload failed"); return fail(ZKPOR_ERR_CUDA); }
    }
    // dry run: which wire does every R1C instruction solve for
    uint8_t *solved = nullptr;
    if (cudaMalloc((void **)&solved, d->n_wires) != cudaSuccess) { set_error("program_upload: out of device memory"); return fail(ZKPOR_ERR_OOM); }
    cudaMemsetAsync(solved, 0, d->n_wires, ctx->stream);
    cudaMemsetAsync(solved, 1, d->n_public + d->n_secret, ctx->stream);
    cudaMemsetAsync(p->solve_e, 0xFF, std::max<uint64_t>(d->n_constraints, 1) * 8, ctx->stream);
    if (cudaMalloc((void **)&p->wstep, d->n_wires * 4) != cudaSuccess) { cudaFree(solved); set_error("program_upload: out of device memory"); return fail(ZKPOR_ERR_OOM); }
    cudaMemsetAsync(p->wstep, 0, d->n_wires * 4, ctx->stream);
    rc = run_schedule<true>(ctx, p, nullptr, nullptr, solved, nullptr, nullptr);
    if (rc == ZKPOR_OK) {
        std::vector<uint8_t> sd;
        rc = fetch(sd, (const uint8_t *)p->step_div, p->steps.size());
        if (rc == ZKPOR_OK) for (size_t i = 0; i < p->steps.size(); i++) p->steps[i].has_div = sd[i] != 0;
    }
    if (rc == ZKPOR_OK) {
        // every wire must have been reached
        std::vector<uint8_t> hs;
        rc = fetch(hs, (const uint8_t *)solved, d->n_wires);
        if (rc == ZKPOR_OK)
            for (uint64_t i = 0; i < d->n_wires; i++)
                if (!hs[i]) { set_error("program_upload: wire %llu is never solved by the schedule", (unsigned long long)i); rc = ZKPOR_ERR_INVALID_ARG; break; }
    }
    cudaFree(solved);
    if (rc != ZKPOR_OK) return fail(rc);
    // narrow levels: move each row's fresh terms (wires solved one level earlier) to the end of their lists (k_solve_narrow_pipe)
    {
        const char *e = getenv("ZKPOR_NARROW_PIPE");
        p->narrow_pipe = e != nullptr && atoi(e) != 0 && p->narrow_threads == NARROW_THREADS && p->stats[2] > 0;   // opt-in: measured 4 % slower
    }
    if (p->narrow_pipe) {
        uint32_t *wlevel = nullptr;
        if (cudaMalloc((void **)&wlevel, d->n_wires * 4) != cudaSuccess) { set_error("program_upload: out of device memory"); return fail(ZKPOR_ERR_OOM); }
        cudaMemsetAsync(wlevel, 0, d->n_wires * 4, ctx->stream);
        ProgView v = make_view(p, nullptr, nullptr);
        for (const Step &st : p->steps) if (st.kind == STEP_NARROW) k_narrow_wlevel<<<grid_for(st.b - st.a, 128), 128, 0, ctx->stream>>>(v, st.a, st.b, wlevel);
        for (const Step &st : p->steps) if (st.kind == STEP_NARROW) k_narrow_partition<<<grid_for(st.b - st.a, 128), 128, 0, ctx->stream>>>(v, st.a, st.b, wlevel);
        const cudaError_t ce = cudaStreamSynchronize(ctx->stream);
        cudaFree(wlevel);
        if (ce != cudaSuccess) { set_error("program_upload: narrow-level analysis failed: %s", cudaGetErrorString(ce)); return fail(ZKPOR_ERR_CUDA); }
    }
    // the deferred tail: the last step, if it is a long run of narrow levels
    uint64_t tail_min = 2048;
    if (const char *e = getenv("ZKPOR_TAIL_MIN")) tail_min = (uint64_t)std::max(0, atoi(e));   // 0: never defer
    if (tail_min > 0 && !p->steps.empty() && p->steps.back().kind == STEP_NARROW && p->steps.back().b - p->steps.back().a >= tail_min) {
        const Step &t = p->steps.back();
        const uint64_t p0 = lvl_start[t.a], p1 = lvl_start[t.b];
        const uint32_t mark = (uint32_t)p->steps.size();       // wstep of the wires the last step solves
        unsigned long long *d_cur = nullptr; uint32_t *d_dep = nullptr;
        if (cudaMalloc((void **)&d_cur, 16) != cudaSuccess) { set_error("program_upload: out of device memory"); return fail(ZKPOR_ERR_OOM); }
        d_dep = (uint32_t *)(d_cur + 1);
        cudaMemsetAsync(d_cur, 0, 16, ctx->stream);
        ProgView v = make_view(p, nullptr, nullptr);
        k_tail_scan<<<grid_for(p1 - p0, 128), 128, 0, ctx->stream>>>(v, p0, p1, mark, nullptr, d_cur, d_dep);
        unsigned long long hn[2] = {0, 0};
        cudaMemcpyAsync(hn, d_cur, 16, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        p->n_tail_wires = hn[0];
        p->tail_after = std::min<uint64_t>((uint32_t)hn[1], p->steps.size() - 1);
        const uint64_t mask_words = (d->n_wires + 31) / 32;
        if (cudaMalloc((void **)&p->tail_wires, std::max<uint64_t>(p->n_tail_wires, 1) * 4) != cudaSuccess ||
            cudaMalloc((void **)&p->tail_mask, mask_words * 4) != cudaSuccess) { cudaFree(d_cur); set_error("program_upload: out of device memory"); return fail(ZKPOR_ERR_OOM); }
        cudaMemsetAsync(d_cur, 0, 8, ctx->stream);
        cudaMemsetAsync(p->tail_mask, 0, mask_words * 4, ctx->stream);
        k_tail_scan<<<grid_for(p1 - p0, 128), 128, 0, ctx->stream>>>(v, p0, p1, mark, p->tail_wires, d_cur, d_dep);
        cudaStreamSynchronize(ctx->stream);   // the context's stream does not synchronise with the blocking copy below
        rc = fetch(p->h_tail_wires, (const uint32_t *)p->tail_wires, p->n_tail_wires);
        cudaFree(d_cur);
        if (rc != ZKPOR_OK) return fail(rc);
        std::sort(p->h_tail_wires.begin(), p->h_tail_wires.end());
        cudaMemcpy(p->tail_wires, p->h_tail_wires.data(), p->n_tail_wires * 4, cudaMemcpyHostToDevice);
        if (p->n_tail_wires) k_wire_mask<<<grid_for(p->n_tail_wires, 256), 256, 0, ctx->stream>>>(p->tail_wires, p->n_tail_wires, p->tail_mask);
        if (cudaEventCreateWithFlags(&p->tail_go, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&p->tail_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error("program_upload: tail setup failed"); return fail(ZKPOR_ERR_CUDA); }
        p->tail_step = (int64_t)p->steps.size() - 1;
    }
    cudaFree(p->wstep); p->wstep = nullptr;
    { static uint64_t next_uid = 1; p->uid = __atomic_fetch_add(&next_uid, 1, __ATOMIC_RELAXED); }
    *out = p;
    return ZKPOR_OK;
}

int32_t zkpor_r1cs_solve(zkpor_ctx *ctx, zkpor_program *prog, zkpor_pk *pk, const void *inputs, void *out_wires, void *out_a, void *out_b,
                         void *out_c, void *out_commitment64) {
    ZK_REQUIRE(ctx && prog && inputs, "r1cs_solve: null argument");
    ZK_REQUIRE(!pk || pk->n_wires_total == prog->n_wires, "r1cs_solve: the program and the key disagree on the number of wires");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    Fr *w;
    if (out_wires && is_device_ptr(out_wires)) w = (Fr *)out_wires;
    else { ZK_TRY(prog->wires.reserve(prog->n_wires * 32)); w = prog->wires.as<Fr>(); }
    const Fr one = Fr::one();
    stage_begin(ctx, ST_H2D);
    ZK_CUDA(cudaMemcpyAsync(w, &one, 32, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(w + 1, inputs, program_inputs(prog) * 32, cudaMemcpyDefault, ctx->stream));
    stage_end(ctx, ST_H2D);
    stage_begin(ctx, ST_SOLVE);
    G1XYZZ commit = G1XYZZ::inf(), pok = G1XYZZ::inf();
    bool has_commit = false;
    ZK_TRY(solver_run(ctx, prog, pk, w, &commit, &pok, &has_commit, false));
    // a = L w, b = R w, c = O w and gnark's satisfaction check
    const size_t bytes = prog->n_rows * sizeof(Fr);
    void *outs[3] = {out_a, out_b, out_c};
    Fr *d[3];
    bool need_tmp = false;
    for (int m = 0; m < 3; m++) need_tmp |= !(outs[m] && is_device_ptr(outs[m]));
    if (need_tmp) ZK_TRY(prog->abc.reserve(3 * bytes));
    for (int m = 0; m < 3; m++) d[m] = (outs[m] && is_device_ptr(outs[m])) ? (Fr *)outs[m] : prog->abc.as<Fr>() + (size_t)m * prog->n_rows;
    ZK_TRY(r1cs_eval_dev(ctx, prog->cs, w, d[0], d[1], d[2]));
    ZK_TRY(r1cs_check_dev(ctx, d[0], d[1], d[2], prog->n_rows));
    stage_end(ctx, ST_SOLVE);
    stage_begin(ctx, ST_D2H);
    for (int m = 0; m < 3; m++)
        if (outs[m] && !is_device_ptr(outs[m])) ZK_CUDA(cudaMemcpyAsync(outs[m], d[m], bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_wires && !is_device_ptr(out_wires)) ZK_CUDA(cudaMemcpyAsync(out_wires, w, prog->n_wires * 32, cudaMemcpyDeviceToHost, ctx->stream));
    stage_end(ctx, ST_D2H);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_commitment64) {
        const G1Affine c = has_commit ? commit.to_affine() : G1Affine::inf();
        memcpy(out_commitment64, &c, 64);
    }
    stages_collect(ctx);
    return ZKPOR_OK;
}

}  // extern "C"
