// One proof across the N GPUs of a box: the communicator behind the sharded prove path.
//
// The reference scales out with independent prover processes (README.md:126); the north star adds ONE proof split over the GPUs:
// every multi-scalar multiplication shards by point chunk, computeH shards by a four-step transform with all-to-all exchanges
// (ntt.cu compute_h_dist), and the ranks' partial sums (5 G1 + 1 G2 point) meet in one all-gather.  Two communicators:
//   - NCCL (one process per GPU, the torchrun / several-prover-processes deployment): ncclSend/ncclRecv groups for the all-to-all,
//     ncclAllGather for the partial sums.  libnccl.so.2 is resolved at run time (dlopen; an NCCL already loaded in the process -- e.g.
//     torch's -- is reused), so the library has no link-time dependency and loads on boxes without NCCL.
//   - in-process (one process, one host thread per GPU -- what a single Go prover with N goroutines locked to OS threads uses, and what
//     zkpor_ctx_create_multi returns): peers PULL their chunk from the sender's buffer with cudaMemcpyPeerAsync over NVLink, ordered by
//     CUDA events; two host barriers per exchange, no device-side synchronisation.  Contexts may share a device (parity tests on a
//     one-GPU box run the N-rank algorithm on one B200).
#include "internal.h"
#include <nccl.h>
#include <dlfcn.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace zk {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("ZKPOR_NCCL_LIB");
        void *h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already holds (torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.h = h;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.Send = (decltype(api.Send))dlsym(h, "ncclSend");
        api.Recv = (decltype(api.Recv))dlsym(h, "ncclRecv");
        api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Send || !api.Recv || !api.AllGather || !api.GroupStart || !api.GroupEnd ||
            !api.GetErrorString) api.h = nullptr;
    });
    return api.h ? &api : nullptr;
}

#define ZK_NCCL(call)                                                                                              \
    do {                                                                                                           \
        ncclResult_t r__ = (call);                                                                                 \
        if (r__ != ncclSuccess) { zk::set_error("%s failed: %s", #call, nccl_api()->GetErrorString(r__)); return ZKPOR_ERR_CUDA; } \
    } while (0)

// host barrier of the in-process group; abort() releases every waiter for good (a rank that failed must not leave its peers hanging)
struct LocalGroup {
    int n = 0;
    std::mutex mu; std::condition_variable cv;
    int arrived = 0; uint64_t gen = 0; bool aborted = false;
    std::vector<zkpor_ctx *> ctxs;
    std::vector<const void *> send; std::vector<const void *> hsend;
    std::vector<cudaEvent_t> ready, done;
    std::atomic<int> refs{0};
    bool wait() {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return false;
        const uint64_t g = gen;
        if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); return true; }
        cv.wait(lk, [&] { return gen != g || aborted; });
        return gen != g;
    }
    void abort() { std::lock_guard<std::mutex> lk(mu); aborted = true; cv.notify_all(); }
};

struct Comm {
    int rank = 0, world = 1;
    LocalGroup *grp = nullptr;
    ncclComm_t nccl = nullptr;
    void *stage = nullptr; size_t stage_cap = 0;   // device staging of the NCCL all-gather of host data
    uint64_t a2a_calls = 0, a2a_bytes = 0;         // bytes this rank received from other ranks
};

void comm_info(zkpor_ctx *ctx, int *rank, int *world) {
    Comm *c = (Comm *)ctx->comm;
    *rank = c ? c->rank : 0; *world = c ? c->world : 1;
}

void comm_abort(zkpor_ctx *ctx) {
    Comm *c = (Comm *)ctx->comm;
    if (c && c->grp) c->grp->abort();
}

static int32_t aborted() { set_error("sharded call aborted: another rank of the group failed"); return ZKPOR_ERR_STATE; }

int32_t comm_all_to_all(zkpor_ctx *ctx, const void *send, void *recv, size_t bytes) {
    Comm *c = (Comm *)ctx->comm;
    if (!c || c->world == 1) { ZK_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, ctx->stream)); return ZKPOR_OK; }
    c->a2a_calls++; c->a2a_bytes += bytes * (size_t)(c->world - 1);
    if (c->nccl) {
        NcclApi *api = nccl_api();
        ZK_NCCL(api->GroupStart());
        for (int j = 0; j < c->world; j++) {
            ZK_NCCL(api->Send((const uint8_t *)send + (size_t)j * bytes, bytes, ncclUint8, j, c->nccl, ctx->stream));
            ZK_NCCL(api->Recv((uint8_t *)recv + (size_t)j * bytes, bytes, ncclUint8, j, c->nccl, ctx->stream));
        }
        ZK_NCCL(api->GroupEnd());
        return ZKPOR_OK;
    }
    LocalGroup *g = c->grp;
    const int r = c->rank;
    ZK_CUDA(cudaEventRecord(g->ready[r], ctx->stream));   // my chunks are written and my receive buffer is free from here on
    g->send[r] = send;
    if (!g->wait()) return aborted();
    for (int j = 0; j < g->n; j++) if (j != r) ZK_CUDA(cudaStreamWaitEvent(ctx->stream, g->ready[j], 0));
    for (int k = 0; k < g->n; k++) {
        const int j = (r + k) % g->n;                     // staggered, so that the ranks do not all pull from rank 0 first
        const uint8_t *src = (const uint8_t *)g->send[j] + (size_t)r * bytes;
        uint8_t *dst = (uint8_t *)recv + (size_t)j * bytes;
        if (g->ctxs[j]->device == ctx->device) ZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        else ZK_CUDA(cudaMemcpyPeerAsync(dst, ctx->device, src, g->ctxs[j]->device, bytes, ctx->stream));
    }
    ZK_CUDA(cudaEventRecord(g->done[r], ctx->stream));
    if (!g->wait()) return aborted();
    for (int j = 0; j < g->n; j++) if (j != r) ZK_CUDA(cudaStreamWaitEvent(ctx->stream, g->done[j], 0));   // peers have pulled: `send` may be reused
    return ZKPOR_OK;
}

int32_t comm_all_gather_host(zkpor_ctx *ctx, const void *send, void *recv, size_t bytes) {
    Comm *c = (Comm *)ctx->comm;
    if (!c || c->world == 1) { memcpy(recv, send, bytes); return ZKPOR_OK; }
    if (c->nccl) {
        NcclApi *api = nccl_api();
        const size_t need = bytes * (size_t)(c->world + 1);
        if (need > c->stage_cap) {
            if (c->stage) cudaFree(c->stage);
            c->stage = nullptr; c->stage_cap = 0;
            ZK_CUDA(cudaMalloc(&c->stage, need + 4096));
            c->stage_cap = need + 4096;
        }
        uint8_t *st = (uint8_t *)c->stage;
        ZK_CUDA(cudaMemcpyAsync(st, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
        ZK_NCCL(api->AllGather(st, st + bytes, bytes, ncclUint8, c->nccl, ctx->stream));
        ZK_CUDA(cudaMemcpyAsync(recv, st + bytes, bytes * (size_t)c->world, cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        return ZKPOR_OK;
    }
    LocalGroup *g = c->grp;
    g->hsend[c->rank] = send;
    if (!g->wait()) return aborted();
    for (int j = 0; j < g->n; j++) memcpy((uint8_t *)recv + (size_t)j * bytes, g->hsend[j], bytes);
    if (!g->wait()) return aborted();
    return ZKPOR_OK;
}

void comm_free(zkpor_ctx *ctx) {
    Comm *c = (Comm *)ctx->comm;
    if (!c) return;
    if (c->nccl && nccl_api()) nccl_api()->CommDestroy(c->nccl);
    if (c->stage) cudaFree(c->stage);
    if (c->grp) {
        LocalGroup *g = c->grp;
        cudaEventDestroy(g->ready[c->rank]); cudaEventDestroy(g->done[c->rank]);
        if (--g->refs == 0) delete g;
    }
    delete c;
    ctx->comm = nullptr;
}

}  // namespace zk

using namespace zk;

extern "C" {

int32_t zkpor_comm_unique_id(uint8_t out_id128[128]) {
    ZK_REQUIRE(out_id128 != nullptr, "comm_unique_id: null output");
    NcclApi *api = nccl_api();
    if (!api) { set_error("comm_unique_id: libnccl.so.2 could not be loaded (set ZKPOR_NCCL_LIB)"); return ZKPOR_ERR_STATE; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ZK_NCCL(api->GetUniqueId(&id));
    memcpy(out_id128, &id, 128);
    return ZKPOR_OK;
}

int32_t zkpor_ctx_comm_init(zkpor_ctx *ctx, const uint8_t id128[128], int32_t rank, int32_t world) {
    ZK_REQUIRE(ctx && id128, "ctx_comm_init: null argument");
    ZK_REQUIRE(world >= 1 && world <= 8 && (world & (world - 1)) == 0 && rank >= 0 && rank < world, "ctx_comm_init: world must be 1, 2, 4 or 8 and 0 <= rank < world");
    ZK_REQUIRE(ctx->comm == nullptr, "ctx_comm_init: the context already belongs to a group");
    NcclApi *api = nccl_api();
    if (!api) { set_error("ctx_comm_init: libnccl.so.2 could not be loaded (set ZKPOR_NCCL_LIB)"); return ZKPOR_ERR_STATE; }
    ZK_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id; memcpy(&id, id128, 128);
    Comm *c = new Comm();
    c->rank = rank; c->world = world;
    ncclResult_t r = api->CommInitRank(&c->nccl, world, id, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", api->GetErrorString(r)); delete c; return ZKPOR_ERR_CUDA; }
    ctx->comm = c;
    return ZKPOR_OK;
}

int32_t zkpor_ctx_create_multi(const int32_t *device_ids, int32_t n, zkpor_ctx **out_ctxs) {
    ZK_REQUIRE(device_ids && out_ctxs, "ctx_create_multi: null argument");
    ZK_REQUIRE(n >= 1 && n <= 8 && (n & (n - 1)) == 0, "ctx_create_multi: the number of contexts must be 1, 2, 4 or 8");
    for (int i = 0; i < n; i++) out_ctxs[i] = nullptr;
    LocalGroup *g = new LocalGroup();
    g->n = n; g->ctxs.resize(n); g->send.assign(n, nullptr); g->hsend.assign(n, nullptr); g->ready.resize(n); g->done.resize(n);
    int32_t rc = ZKPOR_OK;
    int made = 0;
    for (int i = 0; i < n && rc == ZKPOR_OK; i++) {
        rc = zkpor_ctx_create(device_ids[i], &out_ctxs[i]);
        if (rc != ZKPOR_OK) break;
        made++;
        g->ctxs[i] = out_ctxs[i];
        Comm *c = new Comm();
        c->rank = i; c->world = n; c->grp = g; g->refs++;
        out_ctxs[i]->comm = c;
        if (cudaEventCreateWithFlags(&g->ready[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&g->done[i], cudaEventDisableTiming) != cudaSuccess) {
            set_error("ctx_create_multi: cudaEventCreate failed"); rc = ZKPOR_ERR_CUDA;
        }
    }
    // NVLink peer access between every pair of distinct devices (cudaMemcpyPeerAsync falls back to staging through the host without it)
    for (int i = 0; i < made && rc == ZKPOR_OK; i++)
        for (int j = 0; j < made; j++) {
            if (device_ids[i] == device_ids[j]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[j]);
            if (!can) continue;
            cudaSetDevice(device_ids[i]);
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[j], 0);
            if (e != cudaSuccess) cudaGetLastError();   // already enabled
        }
    if (rc != ZKPOR_OK) {
        for (int i = 0; i < made; i++) { zkpor_ctx_destroy(out_ctxs[i]); out_ctxs[i] = nullptr; }
        if (made == 0) delete g;
        return rc;
    }
    return ZKPOR_OK;
}

int32_t zkpor_ctx_comm_info(zkpor_ctx *ctx, int32_t *out_rank, int32_t *out_world, uint64_t out_stats[2]) {
    ZK_REQUIRE(ctx != nullptr, "ctx_comm_info: null context");
    int r, w; comm_info(ctx, &r, &w);
    if (out_rank) *out_rank = r;
    if (out_world) *out_world = w;
    if (out_stats) { Comm *c = (Comm *)ctx->comm; out_stats[0] = c ? c->a2a_calls : 0; out_stats[1] = c ? c->a2a_bytes : 0; }
    return ZKPOR_OK;
}

int32_t zkpor_multi_prove_solve(zkpor_ctx **ctxs, zkpor_pk **pks, zkpor_program **progs, int32_t n, const void *inputs, const uint8_t r_be[32],
                                const uint8_t s_be[32], uint8_t *out_proof, uint32_t *out_len) {
    ZK_REQUIRE(ctxs && pks && progs && inputs && r_be && s_be && out_proof && out_len, "multi_prove_solve: null argument");
    ZK_REQUIRE(n >= 1 && n <= 8, "multi_prove_solve: 1 to 8 contexts");
    std::vector<int32_t> rcs(n, ZKPOR_OK);
    std::vector<std::string> errs(n);
    std::vector<std::vector<uint8_t>> proofs(n, std::vector<uint8_t>(512));
    std::vector<uint32_t> lens(n, 0);
    std::vector<std::thread> th;
    for (int i = 0; i < n; i++)
        th.emplace_back([&, i] {
            rcs[i] = zkpor_groth16_prove_solve(ctxs[i], pks[i], progs[i], inputs, r_be, s_be, proofs[i].data(), &lens[i]);
            if (rcs[i] != ZKPOR_OK) errs[i] = get_error();   // the message lives in the worker's thread-local slot
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rcs[i] != ZKPOR_OK) { set_error("rank %d: %s", i, errs[i].c_str()); return rcs[i]; }
    for (int i = 1; i < n; i++)
        if (lens[i] != lens[0] || memcmp(proofs[i].data(), proofs[0].data(), lens[0]) != 0) { set_error("multi_prove_solve: rank %d produced a different proof", i); return ZKPOR_ERR_STATE; }
    memcpy(out_proof, proofs[0].data(), lens[0]);
    *out_len = lens[0];
    return ZKPOR_OK;
}

int32_t zkpor_compute_h_sharded(zkpor_ctx *ctx, const void *a, const void *b, const void *c, uint32_t log_n, void *out_h_chunk) {
    ZK_REQUIRE(ctx && a && b && c && out_h_chunk, "compute_h_sharded: null argument");
    int rank, world; comm_info(ctx, &rank, &world);
    ZK_REQUIRE(log_n >= 1 && log_n <= 28 && ((size_t)1 << log_n) >= (size_t)world, "compute_h_sharded: log_n out of range");
    ZK_CUDA(cudaSetDevice(ctx->device));
    stages_reset(ctx);
    const size_t m = ((size_t)1 << log_n) / (size_t)world, bytes = m * sizeof(ff::Fr);
    int32_t rc = ZKPOR_OK;
    auto body = [&]() -> int32_t {
        ZK_TRY(ctx->ntt_a.reserve(bytes)); ZK_TRY(ctx->ntt_b.reserve(bytes)); ZK_TRY(ctx->ntt_c.reserve(bytes)); ZK_TRY(ctx->dist_tmp.reserve(bytes));
        const void *src[3] = {a, b, c};
        ff::Fr *dst[3] = {ctx->ntt_a.as<ff::Fr>(), ctx->ntt_b.as<ff::Fr>(), ctx->ntt_c.as<ff::Fr>()};
        for (int k = 0; k < 3; k++) ZK_CUDA(cudaMemcpyAsync(dst[k], src[k], bytes, cudaMemcpyDefault, ctx->stream));
        ZK_TRY(compute_h_dist(ctx, dst[0], dst[1], dst[2], ctx->dist_tmp.as<ff::Fr>(), log_n));
        ZK_CUDA(cudaMemcpyAsync(out_h_chunk, dst[0], bytes, cudaMemcpyDefault, ctx->stream));
        ZK_CUDA(cudaStreamSynchronize(ctx->stream));
        return ZKPOR_OK;
    };
    rc = body();
    if (rc != ZKPOR_OK) comm_abort(ctx);
    stages_collect(ctx);
    return rc;
}

}  // extern "C"
