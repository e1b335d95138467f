// Context lifecycle, error reporting and small host helpers of libzkpor_b200.
#include "internal.h"
#include <cstdlib>

namespace zk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
const char *get_error() { return g_err; }

void fe_from_be32(ff::Fr *out, const uint8_t be[32]) {
    for (int i = 0; i < 8; i++) {
        const uint8_t *p = be + 4 * (7 - i);
        out->l[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
}
static void fp_to_be32(uint8_t out[32], const ff::Fp &mont) {
    ff::Fp p = ff::Fp::from_mont(mont);
    for (int i = 0; i < 8; i++) { uint32_t v = p.l[7 - i]; out[4 * i] = v >> 24; out[4 * i + 1] = v >> 16; out[4 * i + 2] = v >> 8; out[4 * i + 3] = v; }
}
// gnark-crypto G1Affine.RawBytes / G2Affine.RawBytes (ecc/bn254/marshal.go): X||Y big-endian, G2 coordinates A1 first;
// infinity = mUncompressedInfinity (0b01<<6) then zeros
void g1_to_raw_bytes(uint8_t out[64], const ec::G1Affine &p) {
    if (p.is_inf()) { memset(out, 0, 64); out[0] = 0x40; return; }
    fp_to_be32(out, p.x); fp_to_be32(out + 32, p.y);
}
void g2_to_raw_bytes(uint8_t out[128], const ec::G2Affine &p) {
    if (p.is_inf()) { memset(out, 0, 128); out[0] = 0x40; return; }
    fp_to_be32(out, p.x.a1); fp_to_be32(out + 32, p.x.a0); fp_to_be32(out + 64, p.y.a1); fp_to_be32(out + 96, p.y.a0);
}

static const char *STAGE_NAMES[ST_COUNT] = {"h2d", "digits", "sort", "accumulate", "reduce", "ntt", "poseidon", "d2h", "solve"};

}  // namespace zk

using namespace zk;

extern "C" {

const char *zkpor_version(void) { return "zkpor_b200 0.1 (sm_100a)"; }
const char *zkpor_last_error(void) { return zk::get_error(); }
const char *zkpor_stage_name(int32_t i) { return (i >= 0 && i < ST_COUNT) ? STAGE_NAMES[i] : ""; }

int32_t zkpor_device_count(int32_t *out_count) {
    ZK_REQUIRE(out_count != nullptr, "device_count: null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *out_count = n;
    return ZKPOR_OK;
}

int32_t zkpor_ctx_create(int32_t device_id, zkpor_ctx **out) {
    ZK_REQUIRE(out != nullptr, "ctx_create: null output");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: libzkpor_b200 has no CPU fallback (%s)", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return ZKPOR_ERR_NO_DEVICE;
    }
    ZK_REQUIRE(device_id >= 0 && device_id < n, "ctx_create: device id out of range");
    ZK_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    ZK_CUDA(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library contains sm_100a code only", device_id, prop.major, prop.minor);
        return ZKPOR_ERR_NO_DEVICE;
    }
    zkpor_ctx *ctx = new zkpor_ctx();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char *v = getenv("ZKPOR_AFFINE_ROUNDS")) ctx->affine_rounds = atoi(v);
    if (const char *v = getenv("ZKPOR_G2_TIGHT")) ctx->g2_tight_regs = atoi(v) != 0;
    if (const char *v = getenv("ZKPOR_DIRECT_SCATTER")) ctx->direct_scatter = atoi(v) != 0;
    ZK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ZK_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    ZK_CUDA(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
    { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); ZK_CUDA(cudaStreamCreateWithPriority(&ctx->tail_stream, cudaStreamNonBlocking, hi)); }
    for (int i = 0; i < ST_COUNT; i++) {
        ZK_CUDA(cudaEventCreate(&ctx->ev[i][0])); ZK_CUDA(cudaEventCreate(&ctx->ev[i][1]));
        ctx->ev_used[i] = false; ctx->last_ms[i] = 0.f;
    }
    *out = ctx;
    return ZKPOR_OK;
}

void zk_free_poseidon(zkpor_ctx *ctx);
void zk_free_ntt(zkpor_ctx *ctx);

int32_t zkpor_ctx_destroy(zkpor_ctx *ctx) {
    if (!ctx) return ZKPOR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->tail_stream) cudaStreamSynchronize(ctx->tail_stream);
    zk::DevBuf *bufs[] = {&ctx->in_points, &ctx->in_scalars, &ctx->sort_idx, &ctx->sort_idx2, &ctx->view_cnt, &ctx->view_order, &ctx->part_buf, &ctx->part_meta, &ctx->bucket_cnt, &ctx->bucket_off, &ctx->bucket_cur,
                          &ctx->buckets, &ctx->partials, &ctx->windows, &ctx->misc, &ctx->ntt_a, &ctx->ntt_b, &ctx->ntt_c, &ctx->io, &ctx->heavy, &ctx->heavy_part, &ctx->order, &ctx->tree_a, &ctx->tree_b, &ctx->tree_meta, &ctx->dist_tmp};
    for (auto *b : bufs) b->release();
    zk_free_poseidon(ctx); zk_free_ntt(ctx);
    zk::comm_free(ctx);
    for (auto &r : ctx->klog) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : ctx->ev_free) cudaEventDestroy(e);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < ST_COUNT; i++) { cudaEventDestroy(ctx->ev[i][0]); cudaEventDestroy(ctx->ev[i][1]); }
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    if (ctx->tail_stream) cudaStreamDestroy(ctx->tail_stream);
    cudaEventDestroy(ctx->copy_done);
    delete ctx;
    return ZKPOR_OK;
}

int32_t zkpor_ctx_sync(zkpor_ctx *ctx) {
    ZK_REQUIRE(ctx != nullptr, "ctx_sync: null context");
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}
int32_t zkpor_msm_set_affine_rounds(zkpor_ctx *ctx, int32_t rounds) {
    ZK_REQUIRE(ctx != nullptr, "msm_set_affine_rounds: null context");
    ctx->affine_rounds = rounds;
    return ZKPOR_OK;
}
int32_t zkpor_ctx_launch_count(zkpor_ctx *ctx, uint64_t *out) {
    ZK_REQUIRE(ctx != nullptr && out != nullptr, "launch_count: null argument");
    *out = ctx->launches;
    return ZKPOR_OK;
}
int32_t zkpor_ctx_stream(zkpor_ctx *ctx, void **out_stream) {
    ZK_REQUIRE(ctx != nullptr && out_stream != nullptr, "ctx_stream: null argument");
    *out_stream = (void *)ctx->stream;
    return ZKPOR_OK;
}
int32_t zkpor_ctx_kernel_timing(zkpor_ctx *ctx, int32_t enable) {
    ZK_REQUIRE(ctx != nullptr, "kernel_timing: null context");
    ZK_CUDA(cudaSetDevice(ctx->device));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->klog) { ctx->ev_free.push_back(r.e0); ctx->ev_free.push_back(r.e1); }
    ctx->klog.clear();
    ctx->ktime_on = (enable & 1) != 0;
    ctx->ktime_mask = (enable >> 8) ? (uint32_t)(enable >> 8) : 0xFFFFFFFFu;
    return ZKPOR_OK;
}
int32_t zkpor_ctx_kernel_stats(zkpor_ctx *ctx, int32_t klass, double *total_ms, uint64_t *launches, uint64_t *units) {
    ZK_REQUIRE(ctx != nullptr && total_ms && launches && units, "kernel_stats: null argument");
    ZK_CUDA(cudaSetDevice(ctx->device));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    double ms = 0; uint64_t n = 0, u = 0;
    for (auto &r : ctx->klog) {
        if (r.klass != klass) continue;
        float t = 0.f; ZK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms += t; n++; u += r.units;
    }
    *total_ms = ms; *launches = n; *units = u;
    return ZKPOR_OK;
}
int32_t zkpor_ctx_last_timings(zkpor_ctx *ctx, float *out_ms, int32_t cap, int32_t *n) {
    ZK_REQUIRE(ctx != nullptr && out_ms != nullptr && n != nullptr, "last_timings: null argument");
    int k = cap < ST_COUNT ? cap : ST_COUNT;
    for (int i = 0; i < k; i++) out_ms[i] = ctx->last_ms[i];
    *n = ST_COUNT;
    return ZKPOR_OK;
}

}  // extern "C"
