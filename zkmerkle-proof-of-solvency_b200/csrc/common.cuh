// Shared plumbing of libzkpor_b200: context, error reporting, device buffers, pointer classification, stage timers.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/zkpor_b200.h"

namespace zk {

void set_error(const char *fmt, ...);
const char *get_error();

#define ZK_CUDA(call)                                                                                              \
    do {                                                                                                           \
        cudaError_t e__ = (call);                                                                                  \
        if (e__ != cudaSuccess) {                                                                                  \
            zk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);            \
            return e__ == cudaErrorMemoryAllocation ? ZKPOR_ERR_OOM : ZKPOR_ERR_CUDA;                              \
        }                                                                                                          \
    } while (0)

#define ZK_TRY(expr)                 \
    do {                             \
        int32_t rc__ = (expr);       \
        if (rc__ != ZKPOR_OK) return rc__; \
    } while (0)

#define ZK_REQUIRE(cond, msg)                                   \
    do {                                                        \
        if (!(cond)) { zk::set_error("%s", msg); return ZKPOR_ERR_INVALID_ARG; } \
    } while (0)

// stages whose device time is recorded per call (zkpor_ctx_last_timings)
enum Stage { ST_H2D = 0, ST_DIGITS, ST_SORT, ST_ACCUM, ST_REDUCE, ST_NTT, ST_POSEIDON, ST_D2H, ST_SOLVE, ST_COUNT };

// kernel classes whose individual launches are timed with CUDA events (zkpor_ctx_kernel_stats; bench.py's roofline)
enum KClass { KC_ACCUM_G1 = 0, KC_ACCUM_G2, KC_NTT_PASS, KC_SORT, KC_POSEIDON, KC_SOLVE_WIDE, KC_SOLVE_NARROW, KC_COUNT };
struct KRec { int klass; uint64_t units; cudaEvent_t e0, e1; };

// A grow-only device allocation reused across calls (cudaMalloc is far too slow for the hot path).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int32_t reserve(size_t bytes) {
        if (bytes <= cap) return ZKPOR_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes < ((size_t)256 << 20) ? bytes >> 3 : 0);   // slack only where regrowth is likely: a 2^26 proof's buffers are GBs each
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); p = nullptr; return ZKPOR_ERR_OOM; }
        cap = want;
        return ZKPOR_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

}  // namespace zk

struct zkpor_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D of a/b/c overlaps the wire-only MSMs of a proof
    cudaEvent_t copy_done = nullptr;
    cudaStream_t tail_stream = nullptr;   // highest priority: the solver's deferred tail runs beside the multiplications (solver.cu)
    // scalars whose bit is set read as zero in the NEXT msm_sort (bit i of the mask = scalar i; consumed by that call)
    const uint32_t *scalar_mask = nullptr;
    int sm_count = 0;
    uint64_t launches = 0;
    int poseidon_out_lane = 1;
    // scratch
    zk::DevBuf in_points, in_scalars, sort_idx, sort_idx2, view_cnt, view_order, part_buf, part_meta, bucket_cnt, bucket_off, bucket_cur, buckets, partials, windows, misc, ntt_a, ntt_b, ntt_c, io, heavy, heavy_part, order, tree_a, tree_b, tree_meta, dist_tmp;
    bool g2_tight_regs = true;   // env ZKPOR_G2_TIGHT=0: G2 affine rounds at 170 registers / 8 warps per SM instead of 128 / 16
    bool direct_scatter = false;   // env ZKPOR_DIRECT_SCATTER=1: one-level counting sort (returning L2 atomics) at every size
    int affine_rounds = 0;    // affine tree rounds of an MSM: 0 = XYZZ only (default), -1 = automatic depth, k > 0 = at most k; env ZKPOR_AFFINE_ROUNDS
    void *pinned = nullptr; size_t pinned_cap = 0;
    // poseidon constants on device (built lazily)
    void *pos_consts = nullptr;
    // ntt twiddles cache
    void *ntt_tables = nullptr;
    void *dist_tables = nullptr;   // per-rank tables of the distributed transform (ntt.cu)
    void *comm = nullptr;          // communicator of the one-proof-across-N-GPUs mode (dist.cu): NCCL or in-process
    int stream_priority = 0;
    // stage timers
    cudaEvent_t ev[zk::ST_COUNT][2];
    bool ev_used[zk::ST_COUNT];
    float last_ms[zk::ST_COUNT];
    // per-launch kernel timers
    bool ktime_on = false;
    uint32_t ktime_mask = 0xFFFFFFFFu;   // kernel classes that are timed while ktime_on
    std::vector<zk::KRec> klog;
    std::vector<cudaEvent_t> ev_free;
};

namespace zk {

// true if p is device-accessible memory of the current device (device or managed); host otherwise
inline bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Returns a device pointer holding `bytes` of `src`: src itself when it already is device memory, else a copy in
// `stage` (async on the ctx stream).
inline int32_t to_device(zkpor_ctx *ctx, const void *src, size_t bytes, DevBuf &stage, const void **out) {
    if (bytes == 0) { *out = nullptr; return ZKPOR_OK; }
    if (is_device_ptr(src)) { *out = src; return ZKPOR_OK; }
    ZK_TRY(stage.reserve(bytes));
    ZK_CUDA(cudaMemcpyAsync(stage.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *out = stage.p;
    return ZKPOR_OK;
}

inline void stage_begin(zkpor_ctx *ctx, Stage s) { if (!ctx->ev_used[s]) { cudaEventRecord(ctx->ev[s][0], ctx->stream); } }
inline void stage_end(zkpor_ctx *ctx, Stage s) { cudaEventRecord(ctx->ev[s][1], ctx->stream); ctx->ev_used[s] = true; }
inline void stages_reset(zkpor_ctx *ctx) { for (int i = 0; i < ST_COUNT; i++) { ctx->ev_used[i] = false; ctx->last_ms[i] = 0.f; } }
inline void stages_collect(zkpor_ctx *ctx) {
    for (int i = 0; i < ST_COUNT; i++) {
        if (ctx->ev_used[i]) { float ms = 0.f; if (cudaEventElapsedTime(&ms, ctx->ev[i][0], ctx->ev[i][1]) == cudaSuccess) ctx->last_ms[i] = ms; }
    }
}

inline cudaEvent_t kev_get(zkpor_ctx *ctx) {
    if (!ctx->ev_free.empty()) { cudaEvent_t e = ctx->ev_free.back(); ctx->ev_free.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
// brackets one launch of a timed kernel class: KTimed kt(ctx, KC_x, units); launch; kt.stop();
struct KTimed {
    zkpor_ctx *ctx; KRec rec; bool on;
    KTimed(zkpor_ctx *c, int klass, uint64_t units) : ctx(c), on(c->ktime_on && ((c->ktime_mask >> klass) & 1u)) {
        if (!on) return;
        rec.klass = klass; rec.units = units; rec.e0 = kev_get(c); rec.e1 = kev_get(c);
        cudaEventRecord(rec.e0, c->stream);
    }
    void stop() { if (!on) return; cudaEventRecord(rec.e1, ctx->stream); ctx->klog.push_back(rec); on = false; }
};

inline int grid_for(size_t n, int block) { return (int)((n + block - 1) / block); }

}  // namespace zk

#define ZK_LAUNCH(ctx, kernel, grid, block, smem, ...)                       \
    do {                                                                     \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);     \
        (ctx)->launches++;                                                   \
        ZK_CUDA(cudaGetLastError());                                         \
    } while (0)
