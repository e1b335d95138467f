// Constraint evaluation on the GPU: a = L w, b = R w, c = O w for an R1CS whose matrices stay resident in HBM.
// Replaces the part of gnark's r1cs.Solve (constraint/bn254/solver.go, out of tree; reached from groth16.Prove,
// src/prover/prover/prover.go:269) that fills solution.A / .B / .C.  The wire values come from the device solver (solver.cu,
// zkpor_groth16_prove_solve) or from the caller (zkpor_groth16_prove_wires: gnark's solver on the CPU).
#include "internal.h"

using namespace ff;


namespace zk {

// <row, w> by ROW_LANES lanes: lane j takes the terms j, j + ROW_LANES, ... and the partial sums meet by shuffles.  A compiled
// circuit mixes three-term rows with the ~80-term rows of its Poseidon gadgets: a thread per row would walk the long ones serially
// while the other lanes of its warp idle; four lanes per row cut that walk by four and cost the short rows nothing but idle lanes.
// The kernel is bound by the latency of the gathered 32-byte wire values, not by their bytes.
static const int ROW_LANES = 4;
__device__ __forceinline__ Fr row_dot(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ wire_ids, const uint32_t *__restrict__ coeff_ids,
                                      const Fr *__restrict__ coeffs, uint32_t one_id, const Fr *__restrict__ w, uint64_t row, bool live, int lane) {
    Fr acc = Fr::zero();
    if (live) {
        const uint64_t end = row_ptr[row + 1];
        for (uint64_t e = row_ptr[row] + lane; e < end; e += ROW_LANES) {
            const uint32_t id = __ldg(coeff_ids + e);
            const Fr v = w[__ldg(wire_ids + e)];
            acc = Fr::add(acc, id == one_id ? v : Fr::mul(coeffs[id], v));
        }
    }
#pragma unroll
    for (int off = ROW_LANES / 2; off > 0; off >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.l[i] = __shfl_down_sync(0xFFFFFFFFu, acc.l[i], off, ROW_LANES);
        acc = Fr::add(acc, o);
    }
    return acc;                                            // valid on lane 0 of the group
}

__global__ void __launch_bounds__(256) k_r1cs_rows(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ wire_ids,
                                                   const uint32_t *__restrict__ coeff_ids, const Fr *__restrict__ coeffs, uint32_t one_id,
                                                   const Fr *__restrict__ w, uint64_t n_rows, Fr *__restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, row = t / ROW_LANES;
    const int lane = (int)(t % ROW_LANES);
    const Fr acc = row_dot(row_ptr, wire_ids, coeff_ids, coeffs, one_id, w, row, row < n_rows, lane);
    if (row < n_rows && lane == 0) out[row] = acc;
}

// j-th output = row offset + j*stride (the cyclic subsequence a rank of the distributed computeH holds); rows past the end are zero
__global__ void __launch_bounds__(256) k_r1cs_rows_strided(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ wire_ids,
                                                           const uint32_t *__restrict__ coeff_ids, const Fr *__restrict__ coeffs, uint32_t one_id,
                                                           const Fr *__restrict__ w, uint64_t n_rows, uint64_t offset, uint64_t stride, uint64_t count,
                                                           Fr *__restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, j = t / ROW_LANES;
    const int lane = (int)(t % ROW_LANES);
    const uint64_t row = offset + j * stride;
    const Fr acc = row_dot(row_ptr, wire_ids, coeff_ids, coeffs, one_id, w, row, j < count && row < n_rows, lane);
    if (j < count && lane == 0) out[j] = acc;
}

int32_t r1cs_eval_strided_dev(zkpor_ctx *ctx, zkpor_r1cs *cs, const Fr *d_wires, Fr *d_a, Fr *d_b, Fr *d_c, uint64_t offset, uint64_t stride,
                              uint64_t count) {
    Fr *outs[3] = {d_a, d_b, d_c};
    for (int m = 0; m < 3; m++)
        ZK_LAUNCH(ctx, k_r1cs_rows_strided, grid_for(count * ROW_LANES, 256), 256, 0, (const uint64_t *)cs->row_ptr[m], (const uint32_t *)cs->wire_ids[m],
                  (const uint32_t *)cs->coeff_ids[m], (const Fr *)cs->coeffs, cs->one_id, d_wires, cs->n_rows, offset, stride, count, outs[m]);
    return ZKPOR_OK;
}

int32_t r1cs_eval_dev(zkpor_ctx *ctx, zkpor_r1cs *cs, const Fr *d_wires, Fr *d_a, Fr *d_b, Fr *d_c) {
    Fr *outs[3] = {d_a, d_b, d_c};
    for (int m = 0; m < 3; m++)
        ZK_LAUNCH(ctx, k_r1cs_rows, grid_for(cs->n_rows * ROW_LANES, 256), 256, 0, (const uint64_t *)cs->row_ptr[m], (const uint32_t *)cs->wire_ids[m],
                  (const uint32_t *)cs->coeff_ids[m], (const Fr *)cs->coeffs, cs->one_id, d_wires, cs->n_rows, outs[m]);
    return ZKPOR_OK;
}
uint64_t r1cs_rows(const zkpor_r1cs *cs) { return cs->n_rows; }
uint64_t r1cs_wires(const zkpor_r1cs *cs) { return cs->n_wires; }

static int32_t upload(void **dst, const void *src, size_t bytes) {
    *dst = nullptr;
    if (bytes == 0) return ZKPOR_OK;
    ZK_CUDA(cudaMalloc(dst, bytes));
    ZK_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyDefault));
    return ZKPOR_OK;
}

}  // namespace zk

using namespace zk;

extern "C" {

int32_t zkpor_r1cs_free(zkpor_ctx *ctx, zkpor_r1cs *cs) {
    (void)ctx;
    if (!cs) return ZKPOR_OK;
    for (int m = 0; m < 3; m++) { if (cs->row_ptr[m]) cudaFree(cs->row_ptr[m]); if (cs->wire_ids[m]) cudaFree(cs->wire_ids[m]); if (cs->coeff_ids[m]) cudaFree(cs->coeff_ids[m]); }
    if (cs->coeffs) cudaFree(cs->coeffs);
    cs->wires.release(); cs->out.release();
    delete cs;
    return ZKPOR_OK;
}

int32_t zkpor_r1cs_upload(zkpor_ctx *ctx, uint64_t n_constraints, uint64_t n_wires, const zkpor_csr *l, const zkpor_csr *r, const zkpor_csr *o,
                          const void *coeff_table, uint64_t n_coeffs, zkpor_r1cs **out) {
    ZK_REQUIRE(ctx && l && r && o && coeff_table && out, "r1cs_upload: null argument");
    ZK_REQUIRE(n_constraints > 0 && n_wires > 0 && n_wires < (1ull << 32) && n_coeffs > 0 && n_coeffs < (1ull << 32), "r1cs_upload: sizes out of range");
    ZK_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    const zkpor_csr *ms[3] = {l, r, o};
    // validate on the host when the arrays are host memory (ids in range, row pointers monotone and consistent with nnz)
    for (int m = 0; m < 3; m++) {
        ZK_REQUIRE(ms[m]->row_ptr && (ms[m]->nnz == 0 || (ms[m]->wire_ids && ms[m]->coeff_ids)), "r1cs_upload: null matrix array");
        if (is_device_ptr(ms[m]->row_ptr)) continue;
        ZK_REQUIRE(ms[m]->row_ptr[0] == 0 && ms[m]->row_ptr[n_constraints] == ms[m]->nnz, "r1cs_upload: row_ptr does not span nnz");
        for (uint64_t k = 0; k < n_constraints; k++) ZK_REQUIRE(ms[m]->row_ptr[k] <= ms[m]->row_ptr[k + 1], "r1cs_upload: row_ptr not monotone");
        if (!is_device_ptr(ms[m]->wire_ids))
            for (uint64_t e = 0; e < ms[m]->nnz; e++) ZK_REQUIRE(ms[m]->wire_ids[e] < n_wires && ms[m]->coeff_ids[e] < n_coeffs, "r1cs_upload: wire or coefficient id out of range");
    }
    zkpor_r1cs *cs = new zkpor_r1cs();
    cs->n_rows = n_constraints; cs->n_wires = n_wires; cs->n_coeffs = n_coeffs;
    int32_t rc = ZKPOR_OK;
    auto up = [&](void **dst, const void *src, size_t bytes) { if (rc == ZKPOR_OK) rc = upload(dst, src, bytes); };
    for (int m = 0; m < 3; m++) {
        cs->nnz[m] = ms[m]->nnz;
        up((void **)&cs->row_ptr[m], ms[m]->row_ptr, (n_constraints + 1) * 8);
        up((void **)&cs->wire_ids[m], ms[m]->wire_ids, ms[m]->nnz * 4);
        up((void **)&cs->coeff_ids[m], ms[m]->coeff_ids, ms[m]->nnz * 4);
    }
    up((void **)&cs->coeffs, coeff_table, n_coeffs * 32);
    if (rc == ZKPOR_OK && !is_device_ptr(coeff_table)) {
        const Fr one = Fr::one();
        const Fr *t = (const Fr *)coeff_table;
        for (uint64_t i = 0; i < n_coeffs; i++) if (t[i] == one) { cs->one_id = (uint32_t)i; break; }
    }
    if (rc != ZKPOR_OK) { zkpor_r1cs_free(ctx, cs); return rc; }
    *out = cs;
    return ZKPOR_OK;
}

int32_t zkpor_r1cs_eval(zkpor_ctx *ctx, zkpor_r1cs *cs, const void *wires, void *out_a, void *out_b, void *out_c) {
    ZK_REQUIRE(ctx && cs && wires && out_a && out_b && out_c, "r1cs_eval: null argument");
    ZK_CUDA(cudaSetDevice(ctx->device));
    const void *dw;
    ZK_TRY(to_device(ctx, wires, cs->n_wires * 32, cs->wires, &dw));
    void *outs[3] = {out_a, out_b, out_c};
    Fr *d[3];
    const size_t bytes = cs->n_rows * sizeof(Fr);
    bool any_host = false;
    for (int m = 0; m < 3; m++) any_host |= !is_device_ptr(outs[m]);
    if (any_host) ZK_TRY(cs->out.reserve(3 * bytes));
    for (int m = 0; m < 3; m++) d[m] = is_device_ptr(outs[m]) ? (Fr *)outs[m] : cs->out.as<Fr>() + (size_t)m * cs->n_rows;
    ZK_TRY(r1cs_eval_dev(ctx, cs, (const Fr *)dw, d[0], d[1], d[2]));
    for (int m = 0; m < 3; m++)
        if (!is_device_ptr(outs[m])) ZK_CUDA(cudaMemcpyAsync(outs[m], d[m], bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKPOR_OK;
}

}  // extern "C"
