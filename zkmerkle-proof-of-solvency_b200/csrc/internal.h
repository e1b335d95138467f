// Internal (non-exported) interfaces between the translation units of libzkpor_b200.
#pragma once
#include "common.cuh"
#include "ec.cuh"

// R1CS matrices resident in HBM (r1cs.cu)
struct zkpor_r1cs {
    uint64_t n_rows = 0, n_wires = 0, n_coeffs = 0;
    uint64_t *row_ptr[3] = {nullptr, nullptr, nullptr};
    uint32_t *wire_ids[3] = {nullptr, nullptr, nullptr}, *coeff_ids[3] = {nullptr, nullptr, nullptr};
    uint64_t nnz[3] = {0, 0, 0};
    ff::Fr *coeffs = nullptr;
    uint32_t one_id = 0xFFFFFFFFu;   // id of the coefficient 1 (skips the product), if the table has it
    zk::DevBuf wires, out;
};

// Groth16 proving key resident in HBM (groth16.cu)
struct zkpor_pk {
    uint32_t log_n = 0;
    uint64_t n_wires = 0, n_a = 0, n_b = 0, n_k = 0, n_z = 0, n_ck = 0;
    ec::G1Affine *A = nullptr, *B1 = nullptr, *K = nullptr, *Z = nullptr, *ck = nullptr, *ck_sigma = nullptr;
    ec::G2Affine *B2 = nullptr;
    uint32_t *idx_c = nullptr;                                     // gather indices of the committed wires
    uint2 *map_a = nullptr, *map_b = nullptr, *map_k = nullptr;   // wire -> key-point maps of the shared sort (msm.cu map_wire)
    ec::G1Affine alpha1, beta1, delta1;
    ec::G2Affine beta2, delta2;
    bool has_commitment = false;
    // one proof across N GPUs (zkpor_pk_upload_shard): this key holds the points of wires [wire_first, wire_first + n_wires) out of
    // n_wires_total, Z[z_first, z_first + n_z) and the commitment basis of its share of the committed wires
    int shard_rank = 0, shard_world = 1;
    uint64_t wire_first = 0, n_wires_total = 0, z_first = 0;
    zk::DevBuf wires, sub;
    // key points of the wires a program's deferred tail solves (groth16.cu, built on first use per program): compact copies of the
    // A / B / K entries of those wires and the wire ids to gather their values from
    struct TailKey { uint64_t prog_uid = 0, n_a = 0, n_b = 0, n_k = 0; uint32_t *w_a = nullptr, *w_b = nullptr, *w_k = nullptr;
                     ec::G1Affine *A = nullptr, *B1 = nullptr, *K = nullptr; ec::G2Affine *B2 = nullptr; } tail;
};

namespace zk {

// Window plan of one Pippenger run: c-bit signed digits, nwin windows, nb = 2^(c-1) buckets per window.
struct MsmPlan { uint32_t c, nwin, nb; };
MsmPlan msm_plan(uint64_t n);

// Result of the scalar-side half of an MSM (digit extraction + counting sort by bucket), living in ctx scratch:
// for window w, the signed point references of bucket b are sort_idx[w*n + off[w*nb+b] .. +cnt[w*nb+b]).
// Buckets holding more than heavy_t references (skewed witnesses: the value 1 alone is ~15% of a real wire vector) are
// cut into blocks of HEAVY_CHUNK references, each reduced by a whole CTA; see k_heavy_plan in msm.cu.
struct HeavyBlk { uint32_t slot, start, count; };
struct HeavyBkt { uint32_t slot, first_blk, nblk; };
// Population bins of the bucket schedule: slots (window, bucket) are ordered by decreasing reference count with a counting
// sort over SIZE_BINS exact bins (heavy buckets, > heavy_t <= SIZE_BINS - 2 references, share the last bin).
static const uint32_t SIZE_BINS = 8192;
static const uint32_t HEAVY_CHUNK = 4096;        // references per CTA of the heavy-bucket path
static const uint32_t REF_SKIP = 0xFFFFFFFFu;    // entry of a view's heavy list that is not part of the multiplication
struct MsmSorted {
    MsmPlan plan; uint64_t n; const uint32_t *idx, *off, *cnt, *order;   // order: slots by decreasing population
    const uint32_t *hist, *bin_start;   // slots per population bin; first position of a bin in `order`
    bool is_view = false;               // a multiplication's view of a shared sort (msm_view)
    uint32_t heavy_t, max_blks, max_bkts; const HeavyBlk *blks; const HeavyBkt *bkts; const uint32_t *counters;
};

int32_t msm_sort(zkpor_ctx *ctx, const void *d_scalars, uint64_t n, uint32_t flags, MsmSorted *out);
// `terms` = points actually added (length of the compact key array when s is a view), for the per-launch kernel statistics
int32_t msm_accumulate_g1(zkpor_ctx *ctx, const void *d_points, const MsmSorted &s, ec::G1XYZZ *host_out, uint64_t terms = 0);
int32_t msm_accumulate_g2(zkpor_ctx *ctx, const void *d_points, const MsmSorted &s, ec::G2XYZZ *host_out, uint64_t terms = 0);
// One multiplication's view of a sort that ran over a whole wire vector shared by several multiplications: word w of the map =
// {skip bits of wires 32w..32w+31, rank of wire 32w in the compact key array} (zkpor_pk_upload in groth16.cu).  The view lives in
// ctx scratch until the next msm_view / msm_sort.
int32_t msm_view(zkpor_ctx *ctx, const MsmSorted &shared, const uint2 *wire_map, MsmSorted *view);
// sums of the light buckets (cnt <= heavy_t) by batched-affine tree rounds + an XYZZ tail (msm_affine.cu); *done = false
// when the lists are too short for it to pay (or HBM is short) and the caller must run the XYZZ accumulation instead
int32_t msm_tree_sums(zkpor_ctx *ctx, const ec::G1Affine *d_points, const MsmSorted &s, ec::G1XYZZ *buckets, bool *done);
int32_t msm_tree_sums(zkpor_ctx *ctx, const ec::G2Affine *d_points, const MsmSorted &s, ec::G2XYZZ *buckets, bool *done);
// full device MSM on device-resident inputs; result as XYZZ on the host
int32_t msm_g1_dev(zkpor_ctx *ctx, const void *d_points, const void *d_scalars, uint64_t n, uint32_t flags, ec::G1XYZZ *host_out);
int32_t msm_g2_dev(zkpor_ctx *ctx, const void *d_points, const void *d_scalars, uint64_t n, uint32_t flags, ec::G2XYZZ *host_out);

// R1CS matrices resident in HBM (r1cs.cu); out_* are device vectors of n_constraints elements
int32_t r1cs_eval_dev(zkpor_ctx *ctx, zkpor_r1cs *cs, const ff::Fr *d_wires, ff::Fr *d_a, ff::Fr *d_b, ff::Fr *d_c);
uint64_t r1cs_rows(const zkpor_r1cs *cs);
uint64_t r1cs_wires(const zkpor_r1cs *cs);

// Pedersen commitment and proof of knowledge of the key's committed wires, gathered from the device wire vector (groth16.cu):
// one sort, two accumulations.  Called mid-solve by the commitment hint and by the prove entry points that receive solved wires.
int32_t pk_commit_and_pok(zkpor_ctx *ctx, zkpor_pk *pk, const ff::Fr *d_wires, ec::G1XYZZ *commit, ec::G1XYZZ *pok);
// BSB22 challenge: hash_to_field("bsb22-commitment")(commitment.Marshal()) for a commitment without public committed wires (pairing.cu)
ff::Fr commitment_challenge_g1(const ec::G1Affine &commitment);

// witness solver (solver.cu): wires[0] = 1, wires[1 ..] = inputs on entry; every other wire is written.  commit / pok receive the
// commitment hint's by-products when the program has one (has_commit).
// defer_tail: see run_schedule in solver.cu -- the caller must solver_tail_join before it reads the tail's wires
int32_t solver_run(zkpor_ctx *ctx, zkpor_program *prog, zkpor_pk *pk, ff::Fr *d_wires, ec::G1XYZZ *commit, ec::G1XYZZ *pok, bool *has_commit, bool defer_tail);
struct SolverTail { uint64_t uid; const uint32_t *mask; const std::vector<uint32_t> *wires; uint64_t levels; };   // mask: device bitmask over all wires; wires: ascending ids (host)
bool solver_tail_info(const zkpor_program *prog, SolverTail *out);
int32_t solver_tail_join(zkpor_ctx *ctx, zkpor_program *prog);
void solver_tail_abandon(zkpor_ctx *ctx, zkpor_program *prog);
zkpor_r1cs *program_matrices(zkpor_program *prog);
uint64_t program_inputs(const zkpor_program *prog);
// a[k]*b[k] == c[k] for every constraint row; ZKPOR_ERR_STATE naming the first violated row otherwise
int32_t r1cs_check_dev(zkpor_ctx *ctx, const ff::Fr *d_a, const ff::Fr *d_b, const ff::Fr *d_c, uint64_t n_rows);

// NTT (device-resident data)
int32_t ntt_dev(zkpor_ctx *ctx, ff::Fr *d_data, uint32_t log_n, bool inverse, bool dit, bool coset);
int32_t compute_h_dev(zkpor_ctx *ctx, ff::Fr *d_a, ff::Fr *d_b, ff::Fr *d_c, uint32_t log_n);   // result in d_a (bit-reversed)

// communicator of the sharded mode (dist.cu).  all_to_all: chunk i of `send` goes to rank i, chunk j of `recv` comes from rank j
// (device buffers, bytes per peer); all_gather_host: `bytes` of host data from every rank, in rank order, to every rank.
void comm_info(zkpor_ctx *ctx, int *rank, int *world);
int32_t comm_all_to_all(zkpor_ctx *ctx, const void *send, void *recv, size_t bytes_per_peer);
int32_t comm_all_gather_host(zkpor_ctx *ctx, const void *send, void *recv, size_t bytes);
void comm_abort(zkpor_ctx *ctx);   // a failing rank releases its peers' host barriers (in-process group)
void comm_free(zkpor_ctx *ctx);
// computeH across the communicator's ranks (ntt.cu): a, b, c = this rank's n/N evaluations in cyclic order (row g + N j at j); result in
// a = this rank's contiguous chunk of h in gnark's bit-reversed coefficient order; tmp = n/N elements of scratch
int32_t compute_h_dist(zkpor_ctx *ctx, ff::Fr *a, ff::Fr *b, ff::Fr *c, ff::Fr *tmp, uint32_t log_n);
// rows g, g + N, g + 2N, ... of a = L w, b = R w, c = O w (j-th output = row offset + j*stride; rows beyond the system are zero)
int32_t r1cs_eval_strided_dev(zkpor_ctx *ctx, zkpor_r1cs *cs, const ff::Fr *d_wires, ff::Fr *d_a, ff::Fr *d_b, ff::Fr *d_c, uint64_t offset,
                              uint64_t stride, uint64_t count);

// host helpers
void fe_from_be32(ff::Fr *out_plain, const uint8_t be[32]);    // canonical big-endian -> plain limbs (not Montgomery)
void g1_to_raw_bytes(uint8_t out[64], const ec::G1Affine &p);   // gnark RawBytes
void g2_to_raw_bytes(uint8_t out[128], const ec::G2Affine &p);

}  // namespace zk
