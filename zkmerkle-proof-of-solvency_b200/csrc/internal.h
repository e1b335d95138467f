// Internal (non-exported) interfaces between the translation units of libzkpor_b200.
#pragma once
#include "common.cuh"
#include "ec.cuh"

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
// G2 bucket sums of the light buckets with two lanes per bucket (msm_g2pair.cu)
int32_t msm_g2_pair_accumulate(zkpor_ctx *ctx, const ec::G2Affine *d_points, const MsmSorted &s, ec::G2XYZZ *buckets);
// full device MSM on device-resident inputs; result as XYZZ on the host
int32_t msm_g1_dev(zkpor_ctx *ctx, const void *d_points, const void *d_scalars, uint64_t n, uint32_t flags, ec::G1XYZZ *host_out);
int32_t msm_g2_dev(zkpor_ctx *ctx, const void *d_points, const void *d_scalars, uint64_t n, uint32_t flags, ec::G2XYZZ *host_out);

// R1CS matrices resident in HBM (r1cs.cu); out_* are device vectors of n_constraints elements
int32_t r1cs_eval_dev(zkpor_ctx *ctx, zkpor_r1cs *cs, const ff::Fr *d_wires, ff::Fr *d_a, ff::Fr *d_b, ff::Fr *d_c);
uint64_t r1cs_rows(const zkpor_r1cs *cs);
uint64_t r1cs_wires(const zkpor_r1cs *cs);

// NTT (device-resident data)
int32_t ntt_dev(zkpor_ctx *ctx, ff::Fr *d_data, uint32_t log_n, bool inverse, bool dit, bool coset);
int32_t compute_h_dev(zkpor_ctx *ctx, ff::Fr *d_a, ff::Fr *d_b, ff::Fr *d_c, uint32_t log_n);   // result in d_a (bit-reversed)

// host helpers
void fe_from_be32(ff::Fr *out_plain, const uint8_t be[32]);    // canonical big-endian -> plain limbs (not Montgomery)
void g1_to_raw_bytes(uint8_t out[64], const ec::G1Affine &p);   // gnark RawBytes
void g2_to_raw_bytes(uint8_t out[128], const ec::G2Affine &p);

}  // namespace zk
