// Batched Poseidon over BN254 Fr and the fixed-depth Merkle tree on sm_100a.
//
// Replaces (a) the bnb-chain gnark-crypto fr/poseidon hashers -- poseidon.Poseidon / PoseidonBytes / NewPoseidon, call
// sites src/utils/account_tree.go:19,27, src/utils/utils.go:748, src/witness/main.go:181 -- and (b)
// FixedDepthMerkleTree.Build / GetProof, src/utils/merkletree/merkletree.go:192-279,297-308, plus the leaf hashing of
// utils.AccountInfoToHash / ComputeUserAssetsCommitment, src/utils/utils.go:744-750,188-221.
//
// Permutation: x^5 S-box, R_F = 8, R_P = iden3 table, round constants and Cauchy MDS from the Grain LFSR (generated
// here on the host at first use).  Sponge: state = [0, in...], 12 inputs per permutation, lane 0 chains, the last
// partial chunk uses the width-(rem+1) permutation (see DESIGN.md "Poseidon parity" for what is pinned).
// Kernels: 2-to-1 node hash = one thread per node (t = 3 state in registers); wide hashes = 16 lanes per hash, one
// state element per lane, MDS row-times-vector with warp shuffles.
#include "internal.h"
#include <algorithm>

using namespace ff;

namespace zk {

static const int MAX_T = 13;
static const int ROUNDS_P_TABLE[16] = {56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68};

// Per width t: the 8 full-round constant vectors (`full`, 8*t), the dense MDS (`mds`), and the sparse-partial-round form
// (Grassi et al., appendix B; derivation restated and checked in oracle/py/poseidon.py `sparse_constants`):
//   kp[p]      lane-0 constant of partial round p (the other lanes' constants are pushed forward into full[4])
//   pre        dense matrix P used instead of the MDS in the last full round of the first half
//   sv[p*t+j]  row 0 of the sparse matrix of round p: sv[0] = M[0][0], sv[j] = v_j ;  sw[p*t+i] = w_i (column 0), sw[0] unused
struct PoseidonTables {
    const Fr *full[MAX_T + 1]; const Fr *mds[MAX_T + 1]; const Fr *pre[MAX_T + 1]; const Fr *kp[MAX_T + 1];
    const Fr *sv[MAX_T + 1]; const Fr *sw[MAX_T + 1]; int rp[MAX_T + 1];
};
struct PoseidonState { PoseidonTables tab; Fr *blob = nullptr; };

// ------------------------------------------------------------------------------------------------ constants (host)
namespace {
struct Grain {
    uint8_t s[80];
    int clock() { int nb = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]; memmove(s, s + 1, 79); s[79] = (uint8_t)nb; return nb; }
    Grain(int t, int rf, int rp) {
        int pos = 0;
        auto put = [&](int v, int w) { for (int i = w - 1; i >= 0; i--) s[pos++] = (v >> i) & 1; };
        put(1, 2); put(0, 4); put(254, 12); put(t, 12); put(rf, 10); put(rp, 10);
        while (pos < 80) s[pos++] = 1;
        for (int i = 0; i < 160; i++) clock();
    }
    int bit() { for (;;) { int a = clock(), b = clock(); if (a) return b; } }
    Fr draw() {   // 254 bits, MSB first, as a plain integer
        Fr v = Fr::zero();
        for (int i = 0; i < 254; i++) {
            for (int k = 7; k > 0; k--) v.l[k] = (v.l[k] << 1) | (v.l[k - 1] >> 31);
            v.l[0] = (v.l[0] << 1) | (uint32_t)bit();
        }
        return v;
    }
};
bool geq_modulus(const Fr &v) {
    for (int i = 7; i >= 0; i--) { if (v.l[i] > FrParams::M(i)) return true; if (v.l[i] < FrParams::M(i)) return false; }
    return true;
}
}  // namespace

static void build_constants(int t, std::vector<Fr> &rc, std::vector<Fr> &mds, int &rp) {
    rp = ROUNDS_P_TABLE[t - 2];
    Grain g(t, 8, rp);
    rc.clear(); mds.assign((size_t)t * t, Fr::zero());
    while ((int)rc.size() < (8 + rp) * t) { Fr v = g.draw(); if (!geq_modulus(v)) rc.push_back(Fr::to_mont(v)); }
    for (;;) {
        std::vector<Fr> xy(2 * t);
        for (auto &e : xy) { Fr v = g.draw(); if (geq_modulus(v)) v = Fr::reduce_once(v); e = Fr::to_mont(v); }
        bool ok = true;
        for (int i = 0; i < 2 * t && ok; i++) for (int j = 0; j < i; j++) if (xy[i] == xy[j]) { ok = false; break; }
        for (int i = 0; i < t && ok; i++) for (int j = 0; j < t; j++) {
            Fr s = Fr::add(xy[i], xy[t + j]);
            if (s.is_zero()) { ok = false; break; }
            mds[(size_t)i * t + j] = Fr::inv(s);
        }
        if (ok) return;
    }
}

// (t-1)x(t-1) inverse over Fr by Gauss-Jordan (host, one-off)
static std::vector<Fr> mat_inv(std::vector<Fr> a, int n) {
    std::vector<Fr> inv((size_t)n * n, Fr::zero());
    for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = Fr::one();
    for (int c = 0; c < n; c++) {
        int piv = c;
        while (piv < n && a[(size_t)piv * n + c].is_zero()) piv++;
        for (int j = 0; j < n; j++) { std::swap(a[(size_t)c * n + j], a[(size_t)piv * n + j]); std::swap(inv[(size_t)c * n + j], inv[(size_t)piv * n + j]); }
        Fr f = Fr::inv(a[(size_t)c * n + c]);
        for (int j = 0; j < n; j++) { a[(size_t)c * n + j] = Fr::mul(a[(size_t)c * n + j], f); inv[(size_t)c * n + j] = Fr::mul(inv[(size_t)c * n + j], f); }
        for (int r = 0; r < n; r++) {
            if (r == c || a[(size_t)r * n + c].is_zero()) continue;
            Fr g = a[(size_t)r * n + c];
            for (int j = 0; j < n; j++) {
                a[(size_t)r * n + j] = Fr::sub(a[(size_t)r * n + j], Fr::mul(g, a[(size_t)c * n + j]));
                inv[(size_t)r * n + j] = Fr::sub(inv[(size_t)r * n + j], Fr::mul(g, inv[(size_t)c * n + j]));
            }
        }
    }
    return inv;
}

struct SparseConsts { std::vector<Fr> full, kp, pre, sv, sw; };

static SparseConsts build_sparse(int t, int rp, const std::vector<Fr> &rc, const std::vector<Fr> &M) {
    SparseConsts o;
    auto at = [&](const std::vector<Fr> &m, int i, int j) -> const Fr & { return m[(size_t)i * t + j]; };
    // 1. push the linear-lane constants of the partial rounds forward
    std::vector<Fr> carry(t, Fr::zero());
    for (int p = 0; p < rp; p++) {
        std::vector<Fr> cp(t);
        for (int i = 0; i < t; i++) cp[i] = Fr::add(rc[(size_t)(4 + p) * t + i], carry[i]);
        o.kp.push_back(cp[0]);
        for (int i = 0; i < t; i++) { Fr acc = Fr::zero(); for (int j = 1; j < t; j++) acc = Fr::add(acc, Fr::mul(at(M, i, j), cp[j])); carry[i] = acc; }
    }
    for (int r = 0; r < 4; r++) for (int i = 0; i < t; i++) o.full.push_back(rc[(size_t)r * t + i]);
    for (int i = 0; i < t; i++) o.full.push_back(Fr::add(rc[(size_t)(4 + rp) * t + i], carry[i]));
    for (int r = 5 + rp; r < 8 + rp; r++) for (int i = 0; i < t; i++) o.full.push_back(rc[(size_t)r * t + i]);
    // 2. T = S_p * diag(1, T^), from the last partial round backwards; T <- diag(1, T^) * M
    std::vector<Fr> T = M;
    o.sv.assign((size_t)rp * t, Fr::zero()); o.sw.assign((size_t)rp * t, Fr::zero());
    const int n = t - 1;
    for (int p = rp - 1; p >= 0; p--) {
        std::vector<Fr> That((size_t)n * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) That[(size_t)i * n + j] = at(T, i + 1, j + 1);
        std::vector<Fr> inv = mat_inv(That, n);
        o.sv[(size_t)p * t] = at(M, 0, 0);
        for (int j = 0; j < n; j++) { Fr acc = Fr::zero(); for (int x = 0; x < n; x++) acc = Fr::add(acc, Fr::mul(at(T, 0, 1 + x), inv[(size_t)x * n + j])); o.sv[(size_t)p * t + 1 + j] = acc; }
        for (int i = 0; i < n; i++) o.sw[(size_t)p * t + 1 + i] = at(T, i + 1, 0);
        std::vector<Fr> nt((size_t)t * t);
        for (int j = 0; j < t; j++) nt[j] = at(M, 0, j);                                   // row 0 of D*M = row 0 of M
        for (int i = 1; i < t; i++) for (int j = 0; j < t; j++) { Fr acc = Fr::zero(); for (int x = 1; x < t; x++) acc = Fr::add(acc, Fr::mul(That[(size_t)(i - 1) * n + (x - 1)], at(M, x, j))); nt[(size_t)i * t + j] = acc; }
        T = nt;
    }
    o.pre = T;
    return o;
}

static int32_t get_tables(zkpor_ctx *ctx, PoseidonTables *out) {
    if (!ctx->pos_consts) {
        PoseidonState *st = new PoseidonState();
        std::vector<Fr> all;
        size_t off[MAX_T + 1][6] = {{0}};
        for (int t = 2; t <= MAX_T; t++) {
            std::vector<Fr> rc, mds; int rp;
            build_constants(t, rc, mds, rp);
            SparseConsts sc = build_sparse(t, rp, rc, mds);
            st->tab.rp[t] = rp;
            const std::vector<Fr> *parts[6] = {&sc.full, &mds, &sc.pre, &sc.kp, &sc.sv, &sc.sw};
            for (int k = 0; k < 6; k++) { off[t][k] = all.size(); all.insert(all.end(), parts[k]->begin(), parts[k]->end()); }
        }
        cudaError_t e = cudaMalloc((void **)&st->blob, all.size() * sizeof(Fr));
        if (e != cudaSuccess) { delete st; set_error("cudaMalloc poseidon tables: %s", cudaGetErrorString(e)); return ZKPOR_ERR_OOM; }
        e = cudaMemcpy(st->blob, all.data(), all.size() * sizeof(Fr), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(st->blob); delete st; set_error("poseidon tables upload: %s", cudaGetErrorString(e)); return ZKPOR_ERR_CUDA; }
        for (int t = 0; t <= MAX_T; t++) {
            const bool on = t >= 2;
            st->tab.full[t] = on ? st->blob + off[t][0] : nullptr; st->tab.mds[t] = on ? st->blob + off[t][1] : nullptr;
            st->tab.pre[t] = on ? st->blob + off[t][2] : nullptr; st->tab.kp[t] = on ? st->blob + off[t][3] : nullptr;
            st->tab.sv[t] = on ? st->blob + off[t][4] : nullptr; st->tab.sw[t] = on ? st->blob + off[t][5] : nullptr;
            if (!on) st->tab.rp[t] = 0;
        }
        ctx->pos_consts = st;
    }
    *out = ((PoseidonState *)ctx->pos_consts)->tab;
    return ZKPOR_OK;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ Fr sbox5(const Fr &x) { Fr x2 = Fr::sqr(x); return Fr::mul(Fr::sqr(x2), x); }

// 32-byte big-endian canonical -> Montgomery
__device__ __forceinline__ Fr load_be_mont(const uint8_t *p) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
    Fr v;
#pragma unroll
    for (int i = 0; i < 8; i++) v.l[i] = __byte_perm(w[7 - i], 0, 0x0123);
    return Fr::to_mont(v);
}
__device__ __forceinline__ void store_be_plain(uint8_t *p, const Fr &mont) {
    Fr v = Fr::from_mont(mont);
    uint32_t *w = reinterpret_cast<uint32_t *>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) w[7 - i] = __byte_perm(v.l[i], 0, 0x0123);
}

// t = 3 permutation, whole state in registers; partial rounds in the sparse form (8 products instead of 12)
__device__ __forceinline__ void permute3(Fr &s0, Fr &s1, Fr &s2, const PoseidonTables &tab) {
    const Fr *__restrict__ full = tab.full[3];
    const Fr *__restrict__ kp = tab.kp[3];
    const Fr *__restrict__ sv = tab.sv[3];
    const Fr *__restrict__ sw = tab.sw[3];
    const int rp = tab.rp[3];
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        if (r == 4) {
#pragma unroll 1
            for (int p = 0; p < rp; p++) {
                s0 = sbox5(Fr::add(s0, kp[p]));
                Fr n0 = Fr::add(Fr::add(Fr::mul(sv[3 * p], s0), Fr::mul(sv[3 * p + 1], s1)), Fr::mul(sv[3 * p + 2], s2));
                s1 = Fr::add(s1, Fr::mul(sw[3 * p + 1], s0));
                s2 = Fr::add(s2, Fr::mul(sw[3 * p + 2], s0));
                s0 = n0;
            }
        }
        const Fr *__restrict__ m = (r == 3) ? tab.pre[3] : tab.mds[3];
        s0 = sbox5(Fr::add(s0, full[3 * r])); s1 = sbox5(Fr::add(s1, full[3 * r + 1])); s2 = sbox5(Fr::add(s2, full[3 * r + 2]));
        Fr n0 = Fr::add(Fr::add(Fr::mul(m[0], s0), Fr::mul(m[1], s1)), Fr::mul(m[2], s2));
        Fr n1 = Fr::add(Fr::add(Fr::mul(m[3], s0), Fr::mul(m[4], s1)), Fr::mul(m[5], s2));
        Fr n2 = Fr::add(Fr::add(Fr::mul(m[6], s0), Fr::mul(m[7], s1)), Fr::mul(m[8], s2));
        s0 = n0; s1 = n1; s2 = n2;
    }
}

__device__ __forceinline__ Fr shfl_fr(const Fr &v, int src) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src, 16);
    return r;
}

__device__ __forceinline__ Fr shfl_down_fr(const Fr &v, int delta) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], delta, 16);
    return r;
}

// Width-t permutation spread over a 16-lane group: lane i owns state[i] (lanes >= t carry garbage, never read).
// Full rounds: ARK, x^5, dense row-times-vector through shuffles.  Partial rounds (sparse form): lane 0 does the
// S-box, every lane one product for the row-0 dot product (4-step shuffle tree) and one for its own update.
__device__ __forceinline__ Fr group_permute(Fr s, int t, int lane, const PoseidonTables &tab) {
    const int li = lane < t ? lane : 0, rp = tab.rp[t];
    const Fr *__restrict__ full = tab.full[t];
    const Fr *__restrict__ kp = tab.kp[t];
    const Fr *__restrict__ sv = tab.sv[t];
    const Fr *__restrict__ sw = tab.sw[t];
    for (int r = 0; r < 8; r++) {
        if (r == 4) {
            for (int p = 0; p < rp; p++) {
                if (lane == 0) s = sbox5(Fr::add(s, kp[p]));
                Fr s0 = shfl_fr(s, 0);
                Fr prod = lane < t ? Fr::mul(sv[p * t + li], s) : Fr::zero();
#pragma unroll
                for (int d = 8; d > 0; d >>= 1) prod = Fr::add(prod, shfl_down_fr(prod, d));
                if (lane == 0) s = prod;
                else if (lane < t) s = Fr::add(s, Fr::mul(sw[p * t + li], s0));
            }
        }
        const Fr *__restrict__ row = ((r == 3) ? tab.pre[t] : tab.mds[t]) + li * t;
        s = sbox5(Fr::add(s, full[r * t + li]));
        Fr acc = Fr::zero();
        for (int j = 0; j < t; j++) acc = Fr::add(acc, Fr::mul(row[j], shfl_fr(s, j)));
        s = acc;
    }
    return s;
}

// poseidon.Poseidon(inputs...) by a 16-lane group; `input(k)` yields the k-th input in Montgomery form.
template <class In>
__device__ __forceinline__ Fr group_hash(In input, uint32_t n_in, int lane, const PoseidonTables &tab, int out_lane) {
    Fr s = Fr::zero();
    uint32_t start = 0;
    int width = MAX_T;
    for (uint32_t i = 0; i < n_in / 12 && n_in > 12; i++) {
        if (lane >= 1 && lane <= 12) s = input(start + lane - 1);
        s = group_permute(s, 13, lane, tab);
        start += 12;
    }
    if (start < n_in) {
        int rem = (int)(n_in - start);
        if (lane >= 1 && lane <= rem) s = input(start + lane - 1);
        s = group_permute(s, rem + 1, lane, tab);
        width = rem + 1;
    }
    return shfl_fr(s, out_lane < width ? out_lane : 0);
}

// ------------------------------------------------------------------------------------------------ kernels
// generic batch: `count` hashes of n_in big-endian elements
__global__ void __launch_bounds__(128) k_hash_batch(const uint8_t *__restrict__ in, uint32_t n_in, uint64_t count, uint8_t *__restrict__ out,
                                                    PoseidonTables tab, int out_lane) {
    uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    int lane = threadIdx.x & 15;
    bool live = g < count;
    uint64_t gi = live ? g : count - 1;
    const uint8_t *base = in + gi * n_in * 32;
    Fr h = group_hash([&](uint32_t k) { return load_be_mont(base + 32 * (size_t)k); }, n_in, lane, tab, out_lane);
    if (live && lane == 0) store_be_plain(out + g * 32, h);
}

// utils.AccountInfoToHash for accounts of one tier; flat = n x tier*6 u64 (PaddingAccountAssets layout)
__global__ void __launch_bounds__(128) k_account_leaves(const uint8_t *__restrict__ ids, const uint8_t *__restrict__ totals,
                                                        const uint64_t *__restrict__ flat, uint64_t n, uint32_t tier, uint8_t *__restrict__ out,
                                                        PoseidonTables tab, int out_lane) {
    uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    int lane = threadIdx.x & 15;
    bool live = g < n;
    uint64_t gi = live ? g : n - 1;
    const uint32_t nflat = tier * 6, nel = (nflat + 2) / 3;
    const uint64_t *f = flat + gi * nflat;
    // packed triple a*2^128 + b*2^64 + c (src/utils/utils.go:196-218); < 2^192 < r
    Fr commit = group_hash([&](uint32_t k) {
        Fr v = Fr::zero();
        uint64_t a = f[3 * k], b = 3 * k + 1 < nflat ? f[3 * k + 1] : 0, c = 3 * k + 2 < nflat ? f[3 * k + 2] : 0;
        v.l[0] = (uint32_t)c; v.l[1] = (uint32_t)(c >> 32); v.l[2] = (uint32_t)b; v.l[3] = (uint32_t)(b >> 32);
        v.l[4] = (uint32_t)a; v.l[5] = (uint32_t)(a >> 32);
        return Fr::to_mont(v);
    }, nel, lane, tab, out_lane);
    const uint8_t *id = ids + gi * 32, *tot = totals + gi * 96;
    Fr h = group_hash([&](uint32_t k) { return k == 0 ? load_be_mont(id) : k < 4 ? load_be_mont(tot + 32 * (k - 1)) : commit; },
                      5, lane, tab, out_lane);
    if (live && lane == 0) store_be_plain(out + g * 32, h);
}

// ---- thread-per-account leaf hashing ---------------------------------------------------------------------------------
// One thread owns one account; its width-13 sponge state lives in shared memory (element-major, so consecutive threads
// touch consecutive 16-byte words: conflict-free), lane 0 stays in registers through the partial rounds.  Compared
// with the 16-lane layout every lane is busy in the partial rounds (one S-box per state instead of one per 16 lanes):
// ~3.5 K field products per width-13 permutation instead of ~7.2 K lane-products.
static const int TPA_THREADS = 64;

struct TpaState {
    uint4 (*p)[TPA_THREADS];
    int tid;
    __device__ __forceinline__ Fr ld(int e) const { Fr r; uint4 *d = reinterpret_cast<uint4 *>(&r); d[0] = p[2 * e][tid]; d[1] = p[2 * e + 1][tid]; return r; }
    __device__ __forceinline__ void st(int e, const Fr &v) const { const uint4 *s = reinterpret_cast<const uint4 *>(&v); p[2 * e][tid] = s[0]; p[2 * e + 1][tid] = s[1]; }
};

template <int T>
__device__ __forceinline__ void tpa_permute(const TpaState &S, const PoseidonTables &tab) {
    const Fr *__restrict__ full = tab.full[T];
    const Fr *__restrict__ kp = tab.kp[T];
    const Fr *__restrict__ sv = tab.sv[T];
    const Fr *__restrict__ sw = tab.sw[T];
    const int rp = tab.rp[T];
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        if (r == 4) {
            Fr s0 = S.ld(0);
#pragma unroll 1
            for (int p = 0; p < rp; p++) {
                s0 = sbox5(Fr::add(s0, kp[p]));
                Fr acc = Fr::mul(sv[p * T], s0);
#pragma unroll 1
                for (int j = 1; j < T; j++) {
                    Fr sj = S.ld(j);
                    acc = Fr::add(acc, Fr::mul(sv[p * T + j], sj));
                    S.st(j, Fr::add(sj, Fr::mul(sw[p * T + j], s0)));
                }
                s0 = acc;
            }
            S.st(0, s0);
        }
        const Fr *__restrict__ m = (r == 3) ? tab.pre[T] : tab.mds[T];
        Fr acc[T];
#pragma unroll
        for (int i = 0; i < T; i++) acc[i] = Fr::zero();
#pragma unroll 1
        for (int j = 0; j < T; j++) {
            Fr sj = sbox5(Fr::add(S.ld(j), full[r * T + j]));
#pragma unroll
            for (int i = 0; i < T; i++) acc[i] = Fr::add(acc[i], Fr::mul(m[i * T + j], sj));
        }
#pragma unroll
        for (int i = 0; i < T; i++) S.st(i, acc[i]);
    }
}

__global__ void __launch_bounds__(TPA_THREADS) k_account_leaves_tpa(const uint8_t *__restrict__ ids, const uint8_t *__restrict__ totals,
                                                                    const uint64_t *__restrict__ flat, uint64_t n, uint32_t tier,
                                                                    uint8_t *__restrict__ out, PoseidonTables tab, int out_lane) {
    __shared__ uint4 smem[2 * MAX_T][TPA_THREADS];
    const uint64_t g = (uint64_t)blockIdx.x * TPA_THREADS + threadIdx.x;
    if (g >= n) return;                      // no block-wide synchronisation below: every thread is independent
    TpaState S{smem, (int)threadIdx.x};
    const uint32_t nflat = tier * 6, nel = (nflat + 2) / 3;
    const uint64_t *f = flat + g * nflat;
    auto packed = [&](uint32_t k) {          // a*2^128 + b*2^64 + c (src/utils/utils.go:196-218)
        Fr v = Fr::zero();
        uint64_t a = f[3 * k], b = 3 * k + 1 < nflat ? f[3 * k + 1] : 0, c = 3 * k + 2 < nflat ? f[3 * k + 2] : 0;
        v.l[0] = (uint32_t)c; v.l[1] = (uint32_t)(c >> 32); v.l[2] = (uint32_t)b; v.l[3] = (uint32_t)(b >> 32);
        v.l[4] = (uint32_t)a; v.l[5] = (uint32_t)(a >> 32);
        return Fr::to_mont(v);
    };
    // assets commitment: 12 elements per width-13 permutation, lane 0 chains
    S.st(0, Fr::zero());
    uint32_t start = 0;
    int width = MAX_T;
    if (nel > 12) {
        for (uint32_t c = 0; c < nel / 12; c++) {
            for (int j = 1; j <= 12; j++) S.st(j, packed(start + j - 1));
            tpa_permute<13>(S, tab);
            start += 12;
        }
    }
    if (start < nel) {
        const int rem = (int)(nel - start);
        for (int j = 1; j <= rem; j++) S.st(j, packed(start + j - 1));
        width = rem + 1;
        switch (width) {     // widths the reference's tiers produce: 50 assets -> 100 = 8*12 + 4, 500 assets -> 1000 = 83*12 + 4
            case 5: tpa_permute<5>(S, tab); break;
            case 13: tpa_permute<13>(S, tab); break;
            default: {       // any other tier size: sequential fallback through the generic widths
                switch (width) {
                    case 2: tpa_permute<2>(S, tab); break; case 3: tpa_permute<3>(S, tab); break; case 4: tpa_permute<4>(S, tab); break;
                    case 6: tpa_permute<6>(S, tab); break; case 7: tpa_permute<7>(S, tab); break; case 8: tpa_permute<8>(S, tab); break;
                    case 9: tpa_permute<9>(S, tab); break; case 10: tpa_permute<10>(S, tab); break; case 11: tpa_permute<11>(S, tab); break;
                    default: tpa_permute<12>(S, tab); break;
                }
            }
        }
    }
    const Fr commit = S.ld(out_lane < width ? out_lane : 0);
    // leaf = Poseidon5(id, equity, debt, collateral, commitment): width 6
    S.st(0, Fr::zero());
    S.st(1, load_be_mont(ids + g * 32));
    for (int k = 0; k < 3; k++) S.st(2 + k, load_be_mont(totals + g * 96 + 32 * k));
    S.st(5, commit);
    tpa_permute<6>(S, tab);
    store_be_plain(out + g * 32, S.ld(out_lane < 6 ? out_lane : 0));
}

// one Merkle level: node p of `cur` from children 2p, 2p+1 of `prev`; non-dirty children read as nil_prev,
// a node with no dirty child becomes nil_cur (what getNodeAt returns for it) and stays non-dirty.
__global__ void __launch_bounds__(128) k_merkle_level(const uint8_t *__restrict__ prev, const uint8_t *__restrict__ prev_dirty, uint64_t prev_len,
                                                      uint8_t *__restrict__ cur, uint8_t *__restrict__ cur_dirty, uint64_t cur_len,
                                                      const uint8_t *__restrict__ nil_prev, const uint8_t *__restrict__ nil_cur,
                                                      PoseidonTables tab, int out_lane) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cur_len) return;
    uint64_t lc = 2 * p, rc_ = 2 * p + 1;
    bool dl = lc < prev_len && prev_dirty[lc], dr = rc_ < prev_len && prev_dirty[rc_];
    uint4 *dst = reinterpret_cast<uint4 *>(cur + 32 * p);
    if (!dl && !dr) {
        const uint4 *s = reinterpret_cast<const uint4 *>(nil_cur);
        dst[0] = s[0]; dst[1] = s[1]; cur_dirty[p] = 0;
        return;
    }
    Fr s0 = Fr::zero(), s1 = load_be_mont(dl ? prev + 32 * lc : nil_prev), s2 = load_be_mont(dr ? prev + 32 * rc_ : nil_prev);
    permute3(s0, s1, s2, tab);
    store_be_plain(cur + 32 * p, out_lane == 0 ? s0 : out_lane == 1 ? s1 : s2);
    cur_dirty[p] = 1;
}

// 2-to-1 hashes of arbitrary pairs (used for the nil chain and by tests): out[i] = H(in[2i], in[2i+1])
__global__ void k_node_pairs(const uint8_t *__restrict__ in, uint64_t count, uint8_t *__restrict__ out, PoseidonTables tab, int out_lane) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr s0 = Fr::zero(), s1 = load_be_mont(in + 64 * i), s2 = load_be_mont(in + 64 * i + 32);
    permute3(s0, s1, s2, tab);
    store_be_plain(out + 32 * i, out_lane == 0 ? s0 : out_lane == 1 ? s1 : s2);
}

// nil[l] = H(nil[l-1], nil[l-1]), sequential (depth <= 32): one thread
__global__ void k_nil_chain(uint8_t *nil, uint32_t depth, PoseidonTables tab, int out_lane) {
    if (threadIdx.x || blockIdx.x) return;
    for (uint32_t l = 1; l <= depth; l++) {
        Fr s0 = Fr::zero(), s1 = load_be_mont(nil + 32 * (l - 1)), s2 = s1;
        permute3(s0, s1, s2, tab);
        store_be_plain(nil + 32 * l, out_lane == 0 ? s0 : out_lane == 1 ? s1 : s2);
    }
}

__global__ void k_set_keys(uint8_t *__restrict__ leaves, uint8_t *__restrict__ dirty, const uint32_t *__restrict__ keys, uint64_t count,
                           const uint8_t *__restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t k = keys[i];
    const uint4 *s = reinterpret_cast<const uint4 *>(vals + 32 * i);
    uint4 *d = reinterpret_cast<uint4 *>(leaves + 32 * (size_t)k);
    d[0] = s[0]; d[1] = s[1]; dirty[k] = 1;
}

struct TreeDev { const uint8_t *node[33]; const uint8_t *dirty[33]; uint64_t len[33]; const uint8_t *nil; uint32_t depth; };

// out[(i*depth + l)] = sibling of key i at level l (GetProof, merkletree.go:297-308)
__global__ void k_get_proofs(TreeDev t, const uint32_t *__restrict__ keys, uint64_t count, uint8_t *__restrict__ out) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * t.depth) return;
    uint32_t l = (uint32_t)(idx % t.depth);
    uint64_t i = idx / t.depth;
    uint64_t sib = ((uint64_t)keys[i] >> l) ^ 1;
    const uint8_t *src = (sib < t.len[l] && t.dirty[l][sib]) ? t.node[l] + 32 * sib : t.nil + 32 * l;
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(out + 32 * idx);
    d[0] = s[0]; d[1] = s[1];
}
__global__ void k_get_leaves(TreeDev t, const uint32_t *__restrict__ keys, uint64_t count, uint8_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t k = keys[i];
    const uint8_t *src = (k < t.len[0] && t.dirty[0][k]) ? t.node[0] + 32 * k : t.nil;
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(out + 32 * i);
    d[0] = s[0]; d[1] = s[1];
}

}  // namespace zk

using namespace zk;

// ---- witness batches: the serial main loop of the witness service, src/witness/witness/witness.go:144-206 ----
// The reference walks the batches one after the other: it adds every account's assets to the running CEX totals (fillCreateUserOp,
// witness.go:319-340) and hashes the 10 000 packed elements of the CEX state before and after each batch (:159-166, :176-183) -- two
// serial 834-permutation sponges per batch on one core.  A prefix sum is not serial: per-batch deltas (one CTA per batch), one scan
// over the batches, then ALL states' commitments at once (16 lanes per state), then the batch commitments (:193-198).
static const uint32_t CEX_ELEMS_PER_ASSET = 20;   // 2 packed totals + 3 x 6 packed tier-ratio pairs (src/utils/utils.go:53-88)
static const uint32_t CEX_FIELDS = 5;             // TotalEquity, TotalDebt, LoanCollateral, MarginCollateral, PortfolioMarginCollateral

// delta[b][asset][field] = sum over the accounts of batch b; flat = PaddingAccountAssets layout (index, equity, debt, loan, margin, pm per slot)
__global__ void __launch_bounds__(256) k_batch_deltas(const uint64_t *__restrict__ flat, uint32_t tier, uint32_t ops_per_batch, uint32_t n_assets,
                                                      uint64_t *__restrict__ delta, unsigned long long *__restrict__ err) {
    extern __shared__ unsigned long long acc[];
    const uint32_t nacc = n_assets * CEX_FIELDS;
    for (uint32_t i = threadIdx.x; i < nacc; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    const uint64_t b = blockIdx.x, slots = (uint64_t)ops_per_batch * tier;
    const uint64_t *f0 = flat + b * slots * 6;
    for (uint64_t sl = threadIdx.x; sl < slots; sl += blockDim.x) {
        const uint64_t *f = f0 + sl * 6;
        const uint64_t idx = f[0];
        if (idx >= n_assets) { atomicCAS(err, 0ull, (2ull << 56) | (b * ops_per_batch + sl / tier)); continue; }
#pragma unroll
        for (uint32_t k = 0; k < CEX_FIELDS; k++) {
            const unsigned long long v = f[1 + k];
            if (v == 0) continue;
            const unsigned long long old = atomicAdd(&acc[idx * CEX_FIELDS + k], v);
            if (old + v < old) atomicCAS(err, 0ull, (1ull << 56) | b);            // utils.SafeAdd panics on overflow
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nacc; i += blockDim.x) delta[b * nacc + i] = acc[i];
}
// totals[0] = initial, totals[b + 1] = totals[b] + delta[b]   (in place: totals[1..] holds the deltas on entry)
__global__ void k_totals_scan(uint64_t *__restrict__ totals, uint32_t nacc, uint64_t n_batches, unsigned long long *__restrict__ err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nacc) return;
    uint64_t run = totals[i];
    for (uint64_t b = 0; b < n_batches; b++) {
        const uint64_t d = totals[(b + 1) * nacc + i], next = run + d;
        if (next < run) atomicCAS(err, 0ull, (1ull << 56) | b);
        totals[(b + 1) * nacc + i] = run = next;
    }
}
__device__ __forceinline__ Fr pack3_mont(uint64_t a, uint64_t b, uint64_t c) {   // a*2^128 + b*2^64 + c
    Fr v = Fr::zero();
    v.l[0] = (uint32_t)c; v.l[1] = (uint32_t)(c >> 32); v.l[2] = (uint32_t)b; v.l[3] = (uint32_t)(b >> 32); v.l[4] = (uint32_t)a; v.l[5] = (uint32_t)(a >> 32);
    return Fr::to_mont(v);
}
// commitment of CEX state g (g = 0 .. n_batches): Poseidon over n_assets x 20 elements, 16 lanes per state
__global__ void __launch_bounds__(128) k_cex_commitments(const uint64_t *__restrict__ totals, const uint64_t *__restrict__ base_price,
                                                         const uint8_t *__restrict__ tier_elems, uint32_t n_assets, uint64_t count,
                                                         uint8_t *__restrict__ out, PoseidonTables tab, int out_lane) {
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool live = g < count;
    const uint64_t *t = totals + (live ? g : count - 1) * n_assets * CEX_FIELDS;
    const Fr h = group_hash([&](uint32_t k) {
        const uint32_t a = k / CEX_ELEMS_PER_ASSET, e = k - a * CEX_ELEMS_PER_ASSET;
        if (e == 0) return pack3_mont(t[a * CEX_FIELDS], t[a * CEX_FIELDS + 1], base_price[a]);
        if (e == 1) return pack3_mont(t[a * CEX_FIELDS + 2], t[a * CEX_FIELDS + 3], t[a * CEX_FIELDS + 4]);
        return load_be_mont(tier_elems + ((size_t)a * (CEX_ELEMS_PER_ASSET - 2) + (e - 2)) * 32);
    }, n_assets * CEX_ELEMS_PER_ASSET, lane, tab, out_lane);
    if (live && lane == 0) store_be_plain(out + g * 32, h);
}
// BatchCommitment = Poseidon(AccountTreeRoot, Before, After, MinAccountIndex, MaxAccountIndex)   (witness.go:193-198)
__global__ void __launch_bounds__(128) k_batch_commitments(const uint8_t *__restrict__ root, const uint8_t *__restrict__ cex_cm,
                                                           const uint32_t *__restrict__ account_index, uint32_t ops_per_batch, uint64_t n_batches,
                                                           uint8_t *__restrict__ out, PoseidonTables tab, int out_lane) {
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool live = g < n_batches;
    const uint64_t b = live ? g : n_batches - 1;
    const Fr h = group_hash([&](uint32_t k) {
        if (k == 0) return load_be_mont(root);
        if (k == 1) return load_be_mont(cex_cm + b * 32);
        if (k == 2) return load_be_mont(cex_cm + (b + 1) * 32);
        return Fr::from_u64(account_index[b * ops_per_batch + (k == 3 ? 0 : ops_per_batch - 1)]);
    }, 5, lane, tab, out_lane);
    if (live && lane == 0) store_be_plain(out + g * 32, h);
}

struct zkpor_tree {
    uint32_t depth = 0;
    uint64_t capacity = 0;
    uint8_t *node[33] = {nullptr};    // node[0] = leaves
    uint8_t *dirty[33] = {nullptr};
    uint64_t len[33] = {0};
    uint8_t *nil = nullptr;           // (depth+1) x 32 B on device
    uint8_t nil_host[33][32];
    uint8_t root[32];
    bool any_set = false, built = false;
    int out_lane = 1;
};

static TreeDev tree_dev(const zkpor_tree *t) {
    TreeDev d;
    for (int l = 0; l < 33; l++) { d.node[l] = t->node[l]; d.dirty[l] = t->dirty[l]; d.len[l] = t->len[l]; }
    d.nil = t->nil; d.depth = t->depth;
    return d;
}

extern "C" {

void zk_free_poseidon(zkpor_ctx *ctx) {
    if (!ctx->pos_consts) return;
    PoseidonState *st = (PoseidonState *)ctx->pos_consts;
    if (st->blob) cudaFree(st->blob);
    delete st;
    ctx->pos_consts = nullptr;
}

int32_t zkpor_poseidon_set_out_lane(zkpor_ctx *ctx, int32_t lane) {
    ZK_REQUIRE(ctx != nullptr, "poseidon: null context");
    ZK_REQUIRE(lane == 0 || lane == 1, "poseidon: output lane must be 0 or 1");
    ctx->poseidon_out_lane = lane;
    return ZKPOR_OK;
}

int32_t zkpor_poseidon_constants(uint32_t t, void *out_round_constants, void *out_mds, uint32_t *out_rounds_p) {
    ZK_REQUIRE(t >= 2 && t <= (uint32_t)MAX_T && out_round_constants && out_mds && out_rounds_p, "poseidon_constants: width must be 2..13, outputs non-null");
    std::vector<Fr> rc, mds; int rp = 0;
    build_constants((int)t, rc, mds, rp);
    memcpy(out_round_constants, rc.data(), rc.size() * sizeof(Fr));
    memcpy(out_mds, mds.data(), mds.size() * sizeof(Fr));
    *out_rounds_p = (uint32_t)rp;
    return ZKPOR_OK;
}

int32_t zkpor_poseidon_hash_batch(zkpor_ctx *ctx, const void *in_be, uint32_t n_in, uint64_t count, void *out_be) {
    ZK_REQUIRE(ctx != nullptr && in_be != nullptr && out_be != nullptr, "poseidon: null argument");
    ZK_REQUIRE(n_in >= 1, "poseidon: empty input");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return ZKPOR_OK;
    stages_reset(ctx);
    PoseidonTables tab; ZK_TRY(get_tables(ctx, &tab));
    const void *din;
    stage_begin(ctx, ST_H2D);
    ZK_TRY(to_device(ctx, in_be, (size_t)count * n_in * 32, ctx->io, &din));
    stage_end(ctx, ST_H2D);
    const bool out_dev = is_device_ptr(out_be);
    uint8_t *dout = (uint8_t *)out_be;
    if (!out_dev) { ZK_TRY(ctx->misc.reserve(count * 32)); dout = ctx->misc.as<uint8_t>(); }
    stage_begin(ctx, ST_POSEIDON);
    ZK_LAUNCH(ctx, k_hash_batch, grid_for(count * 16, 128), 128, 0, (const uint8_t *)din, n_in, count, dout, tab, ctx->poseidon_out_lane);
    stage_end(ctx, ST_POSEIDON);
    if (!out_dev) {
        stage_begin(ctx, ST_D2H);
        ZK_CUDA(cudaMemcpyAsync(out_be, dout, count * 32, cudaMemcpyDeviceToHost, ctx->stream));
        stage_end(ctx, ST_D2H);
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    return ZKPOR_OK;
}

int32_t zkpor_account_leaves(zkpor_ctx *ctx, const void *ids_be, const void *totals_be, const void *flat_assets, uint64_t n,
                             uint32_t tier, void *out_be) {
    ZK_REQUIRE(ctx != nullptr && ids_be && totals_be && flat_assets && out_be, "account_leaves: null argument");
    ZK_REQUIRE(tier >= 1 && tier <= 4096, "account_leaves: tier out of range");
    ZK_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return ZKPOR_OK;
    stages_reset(ctx);
    PoseidonTables tab; ZK_TRY(get_tables(ctx, &tab));
    const void *d_ids, *d_tot, *d_flat;
    stage_begin(ctx, ST_H2D);
    ZK_TRY(to_device(ctx, ids_be, n * 32, ctx->in_scalars, &d_ids));
    ZK_TRY(to_device(ctx, totals_be, n * 96, ctx->io, &d_tot));
    ZK_TRY(to_device(ctx, flat_assets, n * (size_t)tier * 48, ctx->in_points, &d_flat));
    stage_end(ctx, ST_H2D);
    const bool out_dev = is_device_ptr(out_be);
    uint8_t *dout = (uint8_t *)out_be;
    if (!out_dev) { ZK_TRY(ctx->misc.reserve(n * 32)); dout = ctx->misc.as<uint8_t>(); }
    stage_begin(ctx, ST_POSEIDON);
    ZK_LAUNCH(ctx, k_account_leaves_tpa, grid_for(n, TPA_THREADS), TPA_THREADS, 0, (const uint8_t *)d_ids, (const uint8_t *)d_tot,
              (const uint64_t *)d_flat, n, tier, dout, tab, ctx->poseidon_out_lane);
    stage_end(ctx, ST_POSEIDON);
    if (!out_dev) {
        stage_begin(ctx, ST_D2H);
        ZK_CUDA(cudaMemcpyAsync(out_be, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        stage_end(ctx, ST_D2H);
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    return ZKPOR_OK;
}

// ---- tree ------------------------------------------------------------------------------------------------------
int32_t zkpor_tree_free(zkpor_ctx *ctx, zkpor_tree *t) {
    (void)ctx;
    if (!t) return ZKPOR_OK;
    for (int l = 0; l < 33; l++) { if (t->node[l]) cudaFree(t->node[l]); if (t->dirty[l]) cudaFree(t->dirty[l]); }
    if (t->nil) cudaFree(t->nil);
    delete t;
    return ZKPOR_OK;
}

int32_t zkpor_tree_create(zkpor_ctx *ctx, uint32_t depth, const uint8_t nil_leaf[32], uint64_t capacity, zkpor_tree **out) {
    ZK_REQUIRE(ctx != nullptr && nil_leaf != nullptr && out != nullptr, "tree_create: null argument");
    // merkletree.go:138-146 panics; the C-ABI reports
    ZK_REQUIRE(depth <= 32, "depth too large");
    ZK_REQUIRE(depth > 0, "depth must be positive");
    ZK_REQUIRE(capacity <= (1ull << depth), "capacity exceeds maximum for given depth");
    ZK_CUDA(cudaSetDevice(ctx->device));
    PoseidonTables tab; ZK_TRY(get_tables(ctx, &tab));
    zkpor_tree *t = new zkpor_tree();
    t->depth = depth; t->capacity = capacity; t->out_lane = ctx->poseidon_out_lane;
    auto fail = [&](int32_t rc) { zkpor_tree_free(ctx, t); return rc; };
    for (uint32_t l = 0; l <= depth; l++) {
        uint64_t len = l == 0 ? capacity : (capacity + ((1ull << l) - 1)) >> l;
        if (l > 0 && len == 0) len = 1;
        t->len[l] = len;
        size_t nbytes = (size_t)(len ? len : 1);
        if (cudaMalloc((void **)&t->node[l], nbytes * 32) != cudaSuccess || cudaMalloc((void **)&t->dirty[l], nbytes) != cudaSuccess) {
            set_error("tree_create: out of device memory at level %u", l); return fail(ZKPOR_ERR_OOM);
        }
        if (cudaMemsetAsync(t->dirty[l], 0, nbytes, ctx->stream) != cudaSuccess) { set_error("tree_create: memset failed"); return fail(ZKPOR_ERR_CUDA); }
    }
    if (cudaMalloc((void **)&t->nil, 33 * 32) != cudaSuccess) { set_error("tree_create: out of device memory"); return fail(ZKPOR_ERR_OOM); }
    if (cudaMemcpyAsync(t->nil, nil_leaf, 32, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { set_error("tree_create: copy failed"); return fail(ZKPOR_ERR_CUDA); }
    k_nil_chain<<<1, 32, 0, ctx->stream>>>(t->nil, depth, tab, t->out_lane);
    ctx->launches++;
    if (cudaMemcpyAsync(t->nil_host, t->nil, 32 * (depth + 1), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error("tree_create: nil chain failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(ZKPOR_ERR_CUDA); }
    memcpy(t->root, t->nil_host[depth], 32);
    *out = t;
    return ZKPOR_OK;
}

int32_t zkpor_tree_set_range(zkpor_ctx *ctx, zkpor_tree *t, uint64_t first_key, uint64_t count, const void *leaves_be) {
    ZK_REQUIRE(ctx && t && (leaves_be || count == 0), "tree_set: null argument");
    if (first_key + count > t->capacity) { set_error("key %llu out of range for capacity %llu", (unsigned long long)(first_key + count - 1), (unsigned long long)t->capacity); return ZKPOR_ERR_INVALID_ARG; }
    if (count == 0) return ZKPOR_OK;
    ZK_CUDA(cudaSetDevice(ctx->device));
    ZK_CUDA(cudaMemcpyAsync(t->node[0] + 32 * first_key, leaves_be, count * 32, cudaMemcpyDefault, ctx->stream));
    ZK_CUDA(cudaMemsetAsync(t->dirty[0] + first_key, 1, count, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    t->any_set = true; t->built = false;
    return ZKPOR_OK;
}

int32_t zkpor_tree_set_keys(zkpor_ctx *ctx, zkpor_tree *t, const uint32_t *keys, uint64_t count, const void *leaves_be) {
    ZK_REQUIRE(ctx && t && ((keys && leaves_be) || count == 0), "tree_set: null argument");
    if (count == 0) return ZKPOR_OK;
    ZK_REQUIRE(!is_device_ptr(keys), "tree_set_keys: keys must be a host array");
    for (uint64_t i = 0; i < count; i++)
        if (keys[i] >= t->capacity) { set_error("key %u out of range for capacity %llu", keys[i], (unsigned long long)t->capacity); return ZKPOR_ERR_INVALID_ARG; }
    ZK_CUDA(cudaSetDevice(ctx->device));
    const void *dk, *dv;
    ZK_TRY(to_device(ctx, keys, count * 4, ctx->in_scalars, &dk));
    ZK_TRY(to_device(ctx, leaves_be, count * 32, ctx->io, &dv));
    ZK_LAUNCH(ctx, k_set_keys, grid_for(count, 256), 256, 0, t->node[0], t->dirty[0], (const uint32_t *)dk, count, (const uint8_t *)dv);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    t->any_set = true; t->built = false;
    return ZKPOR_OK;
}

int32_t zkpor_tree_build(zkpor_ctx *ctx, zkpor_tree *t) {
    ZK_REQUIRE(ctx && t, "tree_build: null argument");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    PoseidonTables tab; ZK_TRY(get_tables(ctx, &tab));
    if (t->any_set) {
        stage_begin(ctx, ST_POSEIDON);
        for (uint32_t l = 1; l <= t->depth; l++) {
            ZK_LAUNCH(ctx, k_merkle_level, grid_for(t->len[l], 128), 128, 0, (const uint8_t *)t->node[l - 1], (const uint8_t *)t->dirty[l - 1],
                      t->len[l - 1], t->node[l], t->dirty[l], t->len[l], (const uint8_t *)(t->nil + 32 * (l - 1)),
                      (const uint8_t *)(t->nil + 32 * l), tab, t->out_lane);
        }
        stage_end(ctx, ST_POSEIDON);
        ZK_CUDA(cudaMemcpyAsync(t->root, t->node[t->depth], 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    t->built = true;
    return ZKPOR_OK;
}

int32_t zkpor_tree_root(zkpor_ctx *ctx, zkpor_tree *t, uint8_t out_root[32]) {
    ZK_REQUIRE(ctx && t && out_root, "tree_root: null argument");
    memcpy(out_root, t->root, 32);
    return ZKPOR_OK;
}

static int32_t tree_gather(zkpor_ctx *ctx, zkpor_tree *t, const uint32_t *keys, uint64_t count, void *out_be, bool proofs) {
    ZK_REQUIRE(ctx && t && ((keys && out_be) || count == 0), "tree_get: null argument");
    if (count == 0) return ZKPOR_OK;
    ZK_REQUIRE(!is_device_ptr(keys), "tree_get: keys must be a host array");
    for (uint64_t i = 0; i < count; i++)
        if ((uint64_t)keys[i] >= (1ull << t->depth)) { set_error("key %u out of range for tree depth %u", keys[i], t->depth); return ZKPOR_ERR_INVALID_ARG; }
    ZK_CUDA(cudaSetDevice(ctx->device));
    const void *dk;
    ZK_TRY(to_device(ctx, keys, count * 4, ctx->in_scalars, &dk));
    const size_t per = proofs ? (size_t)t->depth * 32 : 32, bytes = per * count;
    const bool out_dev = is_device_ptr(out_be);
    uint8_t *dout = (uint8_t *)out_be;
    if (!out_dev) { ZK_TRY(ctx->misc.reserve(bytes)); dout = ctx->misc.as<uint8_t>(); }
    TreeDev td = tree_dev(t);
    if (proofs) ZK_LAUNCH(ctx, k_get_proofs, grid_for(count * t->depth, 256), 256, 0, td, (const uint32_t *)dk, count, dout);
    else ZK_LAUNCH(ctx, k_get_leaves, grid_for(count, 256), 256, 0, td, (const uint32_t *)dk, count, dout);
    if (!out_dev) ZK_CUDA(cudaMemcpyAsync(out_be, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}
int32_t zkpor_tree_get_proofs(zkpor_ctx *ctx, zkpor_tree *t, const uint32_t *keys, uint64_t count, void *out_be) { return tree_gather(ctx, t, keys, count, out_be, true); }
int32_t zkpor_tree_get_leaves(zkpor_ctx *ctx, zkpor_tree *t, const uint32_t *keys, uint64_t count, void *out_be) { return tree_gather(ctx, t, keys, count, out_be, false); }

int32_t zkpor_tree_level(zkpor_ctx *ctx, zkpor_tree *t, uint32_t level, void **out_dev_ptr, uint64_t *out_len) {
    ZK_REQUIRE(ctx && t && out_dev_ptr && out_len, "tree_level: null argument");
    ZK_REQUIRE(level <= t->depth, "tree_level: level out of range");
    *out_dev_ptr = t->node[level]; *out_len = t->len[level];
    return ZKPOR_OK;
}


// ---- one tree over the GPUs of a group: rank g owns the leaves [g << k, (g + 1) << k), builds the subtree above them,
// the N subtree roots are all-gathered (32 B each) and every rank finishes the top levels.  Every rank creates the tree with the same
// depth / nil leaf / capacity and sets only leaves of its own range; proofs of a key are served by the rank that owns it (the siblings
// below level k live there; those at and above level k are on every rank).
int32_t zkpor_tree_shard_range(zkpor_ctx *ctx, zkpor_tree *t, uint64_t *out_first_key, uint64_t *out_count, uint32_t *out_subtree_level) {
    ZK_REQUIRE(ctx && t && out_first_key && out_count, "tree_shard_range: null argument");
    int rank, world; comm_info(ctx, &rank, &world);
    uint32_t k = 0;
    while (k < t->depth && ((uint64_t)world << k) < t->capacity) k++;
    const uint64_t first = (uint64_t)rank << k;
    *out_first_key = first < t->capacity ? first : t->capacity;
    *out_count = first < t->capacity ? std::min<uint64_t>((uint64_t)1 << k, t->capacity - first) : 0;
    if (out_subtree_level) *out_subtree_level = k;
    return ZKPOR_OK;
}

int32_t zkpor_tree_build_sharded(zkpor_ctx *ctx, zkpor_tree *t) {
    ZK_REQUIRE(ctx && t, "tree_build_sharded: null argument");
    int rank, world; comm_info(ctx, &rank, &world);
    if (world == 1) return zkpor_tree_build(ctx, t);
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    PoseidonTables tab; ZK_TRY(get_tables(ctx, &tab));
    uint64_t first, count; uint32_t k;
    ZK_TRY(zkpor_tree_shard_range(ctx, t, &first, &count, &k));
    stage_begin(ctx, ST_POSEIDON);
    auto level = [&](uint32_t l) -> int32_t {
        ZK_LAUNCH(ctx, k_merkle_level, grid_for(t->len[l], 128), 128, 0, (const uint8_t *)t->node[l - 1], (const uint8_t *)t->dirty[l - 1],
                  t->len[l - 1], t->node[l], t->dirty[l], t->len[l], (const uint8_t *)(t->nil + 32 * (l - 1)), (const uint8_t *)(t->nil + 32 * l), tab, t->out_lane);
        return ZKPOR_OK;
    };
    for (uint32_t l = 1; l <= k; l++) ZK_TRY(level(l));
    // the subtree roots meet: node[k][g] of rank g (the empty-subtree value where a rank has no leaves)
    uint8_t mine[32], all[8 * 32];
    ZK_REQUIRE(world <= 8, "tree_build_sharded: at most 8 ranks");
    if ((uint64_t)rank < t->len[k]) ZK_CUDA(cudaMemcpyAsync(mine, t->node[k] + (size_t)rank * 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
    else memcpy(mine, t->nil_host[k], 32);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    ZK_TRY(comm_all_gather_host(ctx, mine, all, 32));
    const uint64_t have = std::min<uint64_t>((uint64_t)world, t->len[k]);
    ZK_CUDA(cudaMemcpyAsync(t->node[k], all, have * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemsetAsync(t->dirty[k], 1, have, ctx->stream));
    for (uint32_t l = k + 1; l <= t->depth; l++) ZK_TRY(level(l));
    stage_end(ctx, ST_POSEIDON);
    ZK_CUDA(cudaMemcpyAsync(t->root, t->node[t->depth], 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    t->built = true;
    return ZKPOR_OK;
}

int32_t zkpor_witness_batches(zkpor_ctx *ctx, const zkpor_cex_desc *cex, const uint8_t account_tree_root[32], const void *flat_assets,
                              const uint32_t *account_indices, uint64_t n_accounts, uint32_t tier, uint32_t ops_per_batch,
                              uint64_t *out_totals, void *out_cex_commitments, void *out_batch_commitments) {
    ZK_REQUIRE(ctx && cex && account_tree_root && flat_assets && account_indices, "witness_batches: null argument");
    ZK_REQUIRE(cex->n_assets >= 1 && cex->n_assets <= 1200 && cex->base_prices && cex->tier_ratio_elems && cex->initial_totals, "witness_batches: bad CEX description (at most 1200 assets)");
    ZK_REQUIRE(tier >= 1 && ops_per_batch >= 1 && n_accounts > 0 && n_accounts % ops_per_batch == 0,
               "witness_batches: the accounts must fill whole batches (the witness service pads the last one, src/witness/main.go:71-83)");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    PoseidonTables tab; ZK_TRY(get_tables(ctx, &tab));
    const uint64_t nb = n_accounts / ops_per_batch;
    const uint32_t nacc = cex->n_assets * CEX_FIELDS;
    const void *d_flat, *d_idx;
    stage_begin(ctx, ST_H2D);
    ZK_TRY(to_device(ctx, flat_assets, n_accounts * (size_t)tier * 48, ctx->in_points, &d_flat));
    ZK_TRY(to_device(ctx, account_indices, n_accounts * 4, ctx->in_scalars, &d_idx));
    // scratch: totals (nb+1) x nacc u64 | base prices | tier elements | root | commitments (nb+1) x 32 | batch commitments nb x 32 | err
    const size_t o_tot = 0, o_price = o_tot + (nb + 1) * (size_t)nacc * 8, o_tier = o_price + (size_t)cex->n_assets * 8,
                 o_root = o_tier + (size_t)cex->n_assets * (CEX_ELEMS_PER_ASSET - 2) * 32, o_cm = o_root + 32, o_bc = o_cm + (nb + 1) * 32,
                 o_err = o_bc + nb * 32, total = o_err + 16;
    ZK_TRY(ctx->io.reserve(total));
    uint8_t *base = ctx->io.as<uint8_t>();
    uint64_t *d_tot = (uint64_t *)(base + o_tot);
    unsigned long long *d_err = (unsigned long long *)(base + o_err);
    ZK_CUDA(cudaMemcpyAsync(d_tot, cex->initial_totals, (size_t)nacc * 8, cudaMemcpyDefault, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(base + o_price, cex->base_prices, (size_t)cex->n_assets * 8, cudaMemcpyDefault, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(base + o_tier, cex->tier_ratio_elems, (size_t)cex->n_assets * (CEX_ELEMS_PER_ASSET - 2) * 32, cudaMemcpyDefault, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(base + o_root, account_tree_root, 32, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemsetAsync(d_err, 0, 16, ctx->stream));
    stage_end(ctx, ST_H2D);
    stage_begin(ctx, ST_POSEIDON);
    ZK_LAUNCH(ctx, k_batch_deltas, (int)nb, 256, (size_t)nacc * 8, (const uint64_t *)d_flat, tier, ops_per_batch, cex->n_assets, d_tot + nacc, d_err);
    ZK_LAUNCH(ctx, k_totals_scan, grid_for(nacc, 128), 128, 0, d_tot, nacc, nb, d_err);
    ZK_LAUNCH(ctx, k_cex_commitments, grid_for((nb + 1) * 16, 128), 128, 0, (const uint64_t *)d_tot, (const uint64_t *)(base + o_price), (const uint8_t *)(base + o_tier),
              cex->n_assets, nb + 1, base + o_cm, tab, ctx->poseidon_out_lane);
    ZK_LAUNCH(ctx, k_batch_commitments, grid_for(nb * 16, 128), 128, 0, (const uint8_t *)(base + o_root), (const uint8_t *)(base + o_cm), (const uint32_t *)d_idx,
              ops_per_batch, nb, base + o_bc, tab, ctx->poseidon_out_lane);
    stage_end(ctx, ST_POSEIDON);
    unsigned long long err = 0;
    stage_begin(ctx, ST_D2H);
    ZK_CUDA(cudaMemcpyAsync(&err, d_err, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_totals) ZK_CUDA(cudaMemcpyAsync(out_totals, d_tot, (nb + 1) * (size_t)nacc * 8, cudaMemcpyDefault, ctx->stream));
    if (out_cex_commitments) ZK_CUDA(cudaMemcpyAsync(out_cex_commitments, base + o_cm, (nb + 1) * 32, cudaMemcpyDefault, ctx->stream));
    if (out_batch_commitments) ZK_CUDA(cudaMemcpyAsync(out_batch_commitments, base + o_bc, nb * 32, cudaMemcpyDefault, ctx->stream));
    stage_end(ctx, ST_D2H);
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    stages_collect(ctx);
    if (err) {
        const unsigned long long where = err & ((1ull << 56) - 1);
        if ((err >> 56) == 2) set_error("witness_batches: account #%llu holds an asset index beyond the %u CEX assets", where, cex->n_assets);
        else set_error("witness_batches: a CEX total overflows 64 bits in batch #%llu (utils.SafeAdd)", where);
        return ZKPOR_ERR_STATE;
    }
    return ZKPOR_OK;
}

}  // extern "C"
