/* libzkpor_b200 -- C-ABI of the B200-native hot path of binance/zkmerkle-proof-of-solvency.
 *
 * This is the drop-in boundary: exactly the entry points a cgo shim needs to replace the
 * bodies of the reference's hot calls, nothing else.  Each function cites the reference interface it replaces
 * (paths relative to the reference repository).  The Go-side stubs are in INTEGRATION.md.
 *
 * Conventions
 *   - status: every function returns int32_t, 0 = ZKPOR_OK; zkpor_last_error() gives the message of the last
 *     failure on the calling thread.  Nothing throws or aborts across the boundary.
 *   - field elements: 4 x u64 little-endian limbs, Montgomery form, R = 2^256 -- gnark-crypto's fr.Element /
 *     fp.Element in memory, so Go passes unsafe.Pointer(&slice[0]) with zero copies.
 *   - G1 affine = X||Y (64 B), G2 affine = X.A0||X.A1||Y.A0||Y.A1 (128 B), Montgomery limbs, infinity = all zero:
 *     gnark-crypto's bn254.G1Affine / G2Affine in memory.
 *   - hashes / Merkle nodes: 32-byte big-endian canonical (fr.Element.Bytes(); what merkletree.go stores).
 *   - every pointer may be a host pointer or a CUDA device pointer of the context's device; the library detects
 *     which (cudaPointerGetAttributes).  Host buffers are owned by the caller for the duration of the call only.
 *   - a zkpor_ctx is bound to one GPU and is NOT re-entrant (one call in flight per ctx); different contexts may be
 *     driven from different OS threads.
 *   - there is no CPU fallback: without a CUDA device zkpor_ctx_create fails with ZKPOR_ERR_NO_DEVICE.
 */
#ifndef ZKPOR_B200_H
#define ZKPOR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ZKPOR_OK 0
#define ZKPOR_ERR_INVALID_ARG 1
#define ZKPOR_ERR_NO_DEVICE 2
#define ZKPOR_ERR_CUDA 3
#define ZKPOR_ERR_OOM 4
#define ZKPOR_ERR_STATE 5

/* flags */
#define ZKPOR_SCALARS_MONT 0u      /* scalars are Montgomery-form fr.Element (gnark memory layout) -- the default */
#define ZKPOR_SCALARS_PLAIN 1u     /* scalars are canonical integers (4 x u64 LE)                               */

typedef struct zkpor_ctx zkpor_ctx;
typedef struct zkpor_pk zkpor_pk;
typedef struct zkpor_tree zkpor_tree;

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */
const char *zkpor_version(void);
const char *zkpor_last_error(void);
int32_t zkpor_device_count(int32_t *out_count);
int32_t zkpor_ctx_create(int32_t device_id, zkpor_ctx **out);
int32_t zkpor_ctx_destroy(zkpor_ctx *ctx);
int32_t zkpor_ctx_sync(zkpor_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int32_t zkpor_ctx_launch_count(zkpor_ctx *ctx, uint64_t *out);
/* CUDA stream the context launches on, as a cudaStream_t cast to void* (for event timing by the caller) */
int32_t zkpor_ctx_stream(zkpor_ctx *ctx, void **out_stream);
/* device milliseconds spent in the kernels of the last call, by stage (CUDA events on the context's stream).
 * stage names: zkpor_stage_name(i); returns the number of stages via *n. */
int32_t zkpor_ctx_last_timings(zkpor_ctx *ctx, float *out_ms, int32_t cap, int32_t *n);
const char *zkpor_stage_name(int32_t i);
/* Per-launch timing of the dominant kernels (CUDA events on the context's stream around every launch of a class):
 * enable, run, then query.  klass: 0 = G1 bucket accumulation, 1 = G2 bucket accumulation, 2 = NTT butterfly pass,
 * 3 = digit/scatter (sort) kernels, 4 = Poseidon/Merkle kernels, 5 = solver wide-level launches, 6 = solver fused narrow runs.
 * units = terms (MSM), elements (NTT), hashes, instructions (5), levels (6).
 * enable: 0 = off, 1 = every class, 1 | (mask << 8) = the classes whose bit is set in mask (an event pair per launch is not free:
 * a proof has ~8 000 solver launches, so bench.py times only the classes it reports while the clock runs). */
int32_t zkpor_ctx_kernel_timing(zkpor_ctx *ctx, int32_t enable);
int32_t zkpor_ctx_kernel_stats(zkpor_ctx *ctx, int32_t klass, double *total_ms, uint64_t *launches, uint64_t *units);

/* ---- multi-scalar multiplication -------------------------------------------------------------------------------
 * Replaces gnark-crypto G1Jac.MultiExp / G2Jac.MultiExp (ecc/bn254/multiexp.go, out of tree) as called inside
 * groth16.Prove -- src/prover/prover/prover.go:269.  out = sum_i scalars[i] * points[i], affine. */
int32_t zkpor_msm_g1(zkpor_ctx *ctx, const void *points /* n x 64 B */, const void *scalars /* n x 32 B */, uint64_t n,
                     uint32_t flags, void *out_affine64 /* host */);
int32_t zkpor_msm_g2(zkpor_ctx *ctx, const void *points /* n x 128 B */, const void *scalars, uint64_t n, uint32_t flags,
                     void *out_affine128 /* host */);
/* Multi-GPU: the partial result of this rank's point chunk as an extended-Jacobian point (X,Y,ZZ,ZZZ; 128 B for G1,
 * 256 B for G2) so that ranks can exchange partials (one NCCL all-gather) and finish with zkpor_g{1,2}_sum_partials. */
int32_t zkpor_msm_g1_partial(zkpor_ctx *ctx, const void *points, const void *scalars, uint64_t n, uint32_t flags, void *out_xyzz128);
int32_t zkpor_msm_g2_partial(zkpor_ctx *ctx, const void *points, const void *scalars, uint64_t n, uint32_t flags, void *out_xyzz256);
/* Bucket accumulation strategy.  Default (0): every bucket is summed in extended-Jacobian form.  -1 / k > 0: a bucket is
 * first summed as a binary tree of batched affine additions (one shared inversion per batch of 64) for an automatic number
 * of levels / at most k levels.  Results are identical for every setting (an MSM result is a group element); on B200 the
 * affine rounds measured no faster than the default, so the knob exists for measurement and for the parity
 * tests.  Environment override at context creation: ZKPOR_AFFINE_ROUNDS. */
int32_t zkpor_msm_set_affine_rounds(zkpor_ctx *ctx, int32_t rounds);
int32_t zkpor_g1_sum_partials(const void *partials_xyzz /* host, k x 128 B */, uint32_t k, void *out_affine64);
int32_t zkpor_g2_sum_partials(const void *partials_xyzz /* host, k x 256 B */, uint32_t k, void *out_affine128);

/* ---- NTT -------------------------------------------------------------------------------------------------------
 * Replaces gnark-crypto fft.Domain.FFT / FFTInverse (ecc/bn254/fr/fft, out of tree), same conventions:
 * decimation 0 = DIF (natural in, bit-reversed out), 1 = DIT (bit-reversed in, natural out); coset = OnCoset()
 * with shift 5; inverse includes the 1/n scaling.  In place over `data` (n = 2^log_n Montgomery elements). */
int32_t zkpor_ntt(zkpor_ctx *ctx, void *data, uint32_t log_n, int32_t inverse, int32_t decimation, int32_t coset);
/* out[i] = a[i] * b[i] in Fr (Montgomery in/out), device-resident vectors -- fr.Vector element-wise product */
int32_t zkpor_fr_mul(zkpor_ctx *ctx, const void *a, const void *b, void *out, uint64_t n);
/* gnark backend/groth16/bn254/prove.go computeH: a, b, c = R1CS evaluation vectors of length n_constraints,
 * zero-padded to n = 2^log_n; out_h = n elements, bit-reversed coefficient order (pairs with pk.G1.Z as gnark
 * stores it).  out_h may be host or device. */
int32_t zkpor_compute_h(zkpor_ctx *ctx, const void *a, const void *b, const void *c, uint64_t n_constraints, uint32_t log_n,
                        void *out_h);

/* ---- Groth16 proving key resident in HBM -----------------------------------------------------------------------
 * Replaces the in-memory groth16.ProvingKey filled by pk.UnsafeReadFrom (src/prover/prover/prover.go:342-346);
 * the descriptor fields are gnark's bn254 ProvingKey fields (backend/groth16/bn254/setup.go, out of tree). */
typedef struct {
    uint32_t log_n;                  /* pk.Domain.Cardinality = 2^log_n                                           */
    uint64_t n_wires;                /* len(InfinityA) = len(InfinityB) = number of wires                           */
    uint64_t n_public;               /* r1cs.GetNbPublicVariables(), includes the ONE wire                          */
    uint64_t n_a, n_b, n_k, n_z;     /* len(pk.G1.A), len(pk.G1.B) = len(pk.G2.B), len(pk.G1.K), len(pk.G1.Z)        */
    const void *g1_a, *g1_b, *g1_k, *g1_z;     /* affine arrays                                                     */
    const void *g2_b;
    const void *g1_alpha, *g1_beta, *g1_delta; /* single points                                                     */
    const void *g2_beta, *g2_delta;
    const uint8_t *infinity_a, *infinity_b;    /* n_wires bytes each (Go []bool)                                     */
    uint64_t n_committed;            /* len(CommitmentKeys[0].Basis); 0 = circuit without commitment                */
    const void *ck_basis, *ck_basis_exp_sigma;
    const uint64_t *private_committed;         /* n_committed wire indices (ascending)                              */
    uint64_t commitment_index;       /* wire index of the commitment (challenge) wire                               */
} zkpor_pk_desc;
int32_t zkpor_pk_upload(zkpor_ctx *ctx, const zkpor_pk_desc *desc, zkpor_pk **out);
int32_t zkpor_pk_free(zkpor_ctx *ctx, zkpor_pk *pk);

/* Pedersen commitment of the BSB22 hint: commitment = MSM(pk.CommitmentKeys[0].Basis, values)
 * (gnark-crypto fr/pedersen ProvingKey.Commit; called mid-solve by Prove's hint override). */
int32_t zkpor_pk_commit(zkpor_ctx *ctx, zkpor_pk *pk, const void *committed_values, void *out_affine64);

/* Replaces the body of groth16.Prove after the solver has run (src/prover/prover/prover.go:269):
 *   wires   = solution.W            (n_wires Montgomery elements)
 *   a, b, c = solution.A/B/C        (n_constraints each)
 *   r, s    = the blinding scalars gnark draws with crypto/rand, canonical 32-byte big-endian
 * out_proof = proof.WriteRawTo bytes: Ar 64 | Bs 128 | Krs 64 | u32be nbCommitments | Commitment 64 | Pok 64
 * (388 bytes with one commitment, 324 with none). */
int32_t zkpor_groth16_prove(zkpor_ctx *ctx, zkpor_pk *pk, const void *wires, const void *a, const void *b, const void *c,
                            uint64_t n_constraints, const uint8_t r_be[32], const uint8_t s_be[32], uint8_t *out_proof,
                            uint32_t *out_len);
/* Multi-GPU variant: every MSM runs on this rank's chunk of the key (the pk uploaded on this ctx holds only that chunk;
 * see zkpor_pk_desc and INTEGRATION.md); returns the 7 partial sums (Ar, Bs1, Krs-K, Krs-Z, Commit, Pok as G1 XYZZ
 * 128 B each, then Bs as G2 XYZZ 256 B) for one all-gather.  zkpor_groth16_finish combines k ranks' partials. */
#define ZKPOR_PROVE_PARTIAL_BYTES (6 * 128 + 256)
int32_t zkpor_groth16_prove_partial(zkpor_ctx *ctx, zkpor_pk *pk, const void *wires_a, const void *wires_b, const void *wires_k,
                                    const void *committed, const void *h_chunk, uint64_t n_h, void *out_partials);
int32_t zkpor_groth16_finish(const void *partials /* k x ZKPOR_PROVE_PARTIAL_BYTES */, uint32_t k, const void *g1_alpha,
                             const void *g1_beta, const void *g1_delta, const void *g2_beta, const void *g2_delta,
                             const uint8_t r_be[32], const uint8_t s_be[32], int32_t has_commitment, uint8_t *out_proof,
                             uint32_t *out_len);

/* ---- constraint evaluation (the linear-algebra half of r1cs.Solve) ----------------------------------
 * gnark's solver fills solution.A/B/C with <L_k, w>, <R_k, w>, <O_k, w> for every constraint k while it solves
 * (constraint/bn254/solver.go, out of tree; reached from groth16.Prove, src/prover/prover/prover.go:269).  With the three
 * matrices resident in HBM the Go side hands over the wire vector only: 2.1 GB per 2^26 proof instead of 8.4 GB.
 * Matrices are compressed sparse rows as gnark stores linear expressions: per term a wire id and an id into the shared
 * coefficient table (constraint.CoeffTable: Montgomery fr.Elements). */
typedef struct zkpor_r1cs zkpor_r1cs;
typedef struct {
    uint64_t nnz;
    const uint64_t *row_ptr;      /* n_constraints + 1 entries */
    const uint32_t *wire_ids;     /* nnz */
    const uint32_t *coeff_ids;    /* nnz */
} zkpor_csr;
int32_t zkpor_r1cs_upload(zkpor_ctx *ctx, uint64_t n_constraints, uint64_t n_wires, const zkpor_csr *l, const zkpor_csr *r,
                          const zkpor_csr *o, const void *coeff_table /* n_coeffs x 32 B */, uint64_t n_coeffs, zkpor_r1cs **out);
int32_t zkpor_r1cs_free(zkpor_ctx *ctx, zkpor_r1cs *cs);
/* a = L w, b = R w, c = O w (n_constraints Montgomery elements each; host or device outputs) */
int32_t zkpor_r1cs_eval(zkpor_ctx *ctx, zkpor_r1cs *cs, const void *wires, void *out_a, void *out_b, void *out_c);
/* zkpor_groth16_prove with a, b, c evaluated on the device from the wire vector */
int32_t zkpor_groth16_prove_wires(zkpor_ctx *ctx, zkpor_pk *pk, zkpor_r1cs *cs, const void *wires, const uint8_t r_be[32],
                                  const uint8_t s_be[32], uint8_t *out_proof, uint32_t *out_len);

/* ---- witness solver (r1cs.Solve with its hints) --------------------------------------------------------
 * Replaces gnark's constraint solver (constraint/bn254/solver.go + system.go, out of tree), the first step of groth16.Prove --
 * src/prover/prover/prover.go:269 -- including the hint functions: the reference's own IntegerDivision (circuit/utils.go:103-110,
 * registered at src/prover/prover/prover.go:68), gnark's std hints the circuit pulls in (bits.NBits behind api.ToBinary, InvZero
 * behind api.IsZero, the rangecheck limb decomposition, the logderivlookup lookup and the logderivarg multiplicity count, the fork's
 * CmpNOp comparator) and the BSB22 commitment placeholder that Prove overrides (solve -> Pedersen commit -> hash_to_field -> continue).
 *
 * The program is gnark's compiled constraint system flattened by the cgo shim (INTEGRATION.md "solver contract"):
 *   - the three R1CS matrices + coefficient table (as zkpor_r1cs_upload),
 *   - cs.Instructions as (kind, arg): kind R1C -> arg = constraint row; kind HINT -> arg = hint record,
 *   - cs.Levels as level_ptr / level_instr (instruction ids; the instructions of one level are independent),
 *   - hint records: function id, parameter, first output wire and number of outputs (gnark allocates a hint's outputs as
 *     consecutive wires), and the inputs as rows [hint_in_ptr[h], hint_in_end[h]) of the auxiliary matrix `aux` of linear
 *     expressions; lookup table t = rows [table_ptr[t], table_ptr[t+1]) of the same matrix.
 * Which wire an R1C instruction solves for is not part of the contract: gnark finds it at run time (the one unsolved wire); this
 * library finds it once, at upload, by a dry run of the schedule on the GPU. */
#define ZKPOR_INS_R1C 0u
#define ZKPOR_INS_HINT 1u
#define ZKPOR_HINT_DIVMOD 1u      /* out = (in[0] / in[1], in[0] % in[1]) as integers: IntegerDivision, circuit/utils.go:103-110 */
#define ZKPOR_HINT_NBITS 2u       /* out[i] = bit i of in[0]                                   (api.ToBinary)               */
#define ZKPOR_HINT_INVZERO 3u     /* out = 1/in[0], or 0 when in[0] = 0                        (api.IsZero)                 */
#define ZKPOR_HINT_DECOMPOSE 4u   /* out[i] = limb i of in[0], param bits per limb            (rangecheck.Check)           */
#define ZKPOR_HINT_LOOKUP 5u      /* out[i] = table[param][in[i]]                              (logderivlookup.Lookup)      */
#define ZKPOR_HINT_CMP 6u         /* out = -1 / 0 / 1 for in[0] <, =, > in[1] as integers      (api.CmpNOp)                 */
#define ZKPOR_HINT_COUNT 7u       /* out[k] = #{i : in[i] = k}, k < n_out                      (logderivarg multiplicities) */
#define ZKPOR_HINT_COMMIT 8u      /* out = hash_to_field(Pedersen commitment of pk's committed wires)  (BSB22 placeholder)  */
typedef struct zkpor_program zkpor_program;
typedef struct {
    uint64_t n_wires, n_public /* includes the ONE wire */, n_secret, n_constraints;
    zkpor_csr l, r, o;
    const void *coeff_table; uint64_t n_coeffs;
    uint64_t n_instr; const uint8_t *instr_kind; const uint32_t *instr_arg;
    uint64_t n_levels; const uint64_t *level_ptr /* n_levels + 1 */; const uint32_t *level_instr /* n_instr */;
    uint64_t n_hints; const uint32_t *hint_fn, *hint_param, *hint_out_first, *hint_n_out; const uint64_t *hint_in_ptr, *hint_in_end;
    uint64_t n_aux_rows; zkpor_csr aux;
    uint64_t n_tables; const uint64_t *table_ptr /* n_tables + 1 */;
} zkpor_program_desc;
int32_t zkpor_program_upload(zkpor_ctx *ctx, const zkpor_program_desc *desc, zkpor_program **out);
int32_t zkpor_program_free(zkpor_ctx *ctx, zkpor_program *prog);
/* the program's three matrices as a zkpor_r1cs (for zkpor_r1cs_eval / zkpor_groth16_prove_wires on wires solved elsewhere); the handle
 * is borrowed: it lives as long as the program and must not be passed to zkpor_r1cs_free */
int32_t zkpor_program_r1cs(zkpor_program *prog, zkpor_r1cs **out);
/* number of schedule steps by type: wide level launches, fused runs of narrow levels, levels inside those runs, count hints */
int32_t zkpor_program_stats(zkpor_program *prog, uint64_t out4[4]);
/* The deferred tail of the schedule (0, 0, 0 when the program has none): when the schedule ends in a long run of narrow levels -- in
 * the reference circuit the serial Poseidon sponges of the CEX commitments (src/witness/witness/witness.go:159-166 off-circuit,
 * circuit/batch_create_user_circuit.go:129,320 in-circuit), ~170 000 levels of 1..13 instructions -- zkpor_groth16_prove_solve starts
 * that run on a side stream as soon as its inputs are solved and overlaps it with the A, B and K multiplications, which see the
 * run's wires as zero; their terms are added by small multiplications once the run is done (a multi-scalar multiplication is linear
 * in the wire vector).  out3 = levels in the run, wires it solves, schedule step it starts before.  Env ZKPOR_DEFER_TAIL=0 disables
 * the overlap, ZKPOR_TAIL_MIN sets the minimum run length (levels; default 2048). */
int32_t zkpor_program_tail_info(zkpor_program *prog, uint64_t out3[3]);
/* the ids of the wires the deferred run solves, ascending (cap >= the count zkpor_program_tail_info reports; host buffer) */
int32_t zkpor_program_tail_wires(zkpor_program *prog, uint32_t *out_wires, uint64_t cap);
/* r1cs.Solve: inputs = the n_public - 1 public then the n_secret secret values (Montgomery).  out_wires = n_wires elements,
 * out_a / out_b / out_c = n_constraints each (any of the four may be NULL).  pk supplies the commitment key and may be NULL for a
 * program without a commitment hint; out_commitment64 (may be NULL) receives the commitment point.  An unsatisfied constraint,
 * a division by zero or a lookup outside its table is ZKPOR_ERR_STATE with the first offending row in the message. */
int32_t zkpor_r1cs_solve(zkpor_ctx *ctx, zkpor_program *prog, zkpor_pk *pk, const void *inputs, void *out_wires, void *out_a,
                         void *out_b, void *out_c, void *out_commitment64);
/* The whole of groth16.Prove (src/prover/prover/prover.go:269): solve, then the proof as zkpor_groth16_prove makes it.  The
 * commitment and its proof of knowledge are computed once, mid-solve. */
int32_t zkpor_groth16_prove_solve(zkpor_ctx *ctx, zkpor_pk *pk, zkpor_program *prog, const void *inputs, const uint8_t r_be[32],
                                  const uint8_t s_be[32], uint8_t *out_proof, uint32_t *out_len);

/* ---- one proof across the N GPUs of a box -----------------------------------------------------------------
 * The reference scales out with independent prover processes (README.md:126: several provers pull batches from one queue); this
 * adds latency scaling of ONE proof: the key is split by point chunk (wires in N contiguous ranges, Z and the commitment basis in N
 * chunks), computeH runs as a four-step transform with all-to-all exchanges, and the ranks' partial sums meet in one all-gather.
 * Every rank holds the constraint system and solves the whole witness (the solver is a latency chain), so no wire data is exchanged.
 * Two ways to form a group of N = 1, 2, 4 or 8 contexts:
 *   - one process per GPU (torchrun, or N prover processes): rank 0 calls zkpor_comm_unique_id and passes the 128 bytes to the other
 *     ranks by any channel; every rank calls zkpor_ctx_comm_init on its own context.  Exchanges run over NCCL (libnccl.so.2 is
 *     resolved at run time; ZKPOR_NCCL_LIB overrides the path).
 *   - one process (a Go prover with one OS-locked goroutine per GPU): zkpor_ctx_create_multi returns the N contexts already joined;
 *     peers pull over NVLink with cudaMemcpyPeerAsync.  Device ids may repeat (the N-rank algorithm on fewer GPUs, for tests).
 * On a context that belongs to a group, zkpor_pk_upload_shard uploads this rank's part of the key, and zkpor_groth16_prove_solve /
 * zkpor_groth16_prove_wires / zkpor_r1cs_solve with such a key are COLLECTIVE: every rank calls them with the same arguments
 * (each from its own thread or process) and every rank receives the same proof bytes -- identical to the single-GPU proof. */
int32_t zkpor_comm_unique_id(uint8_t out_id128[128]);
int32_t zkpor_ctx_comm_init(zkpor_ctx *ctx, const uint8_t id128[128], int32_t rank, int32_t world);
int32_t zkpor_ctx_create_multi(const int32_t *device_ids, int32_t n, zkpor_ctx **out_ctxs /* n */);
/* rank and size of the context's group (0 and 1 without one); out_stats (may be NULL) = all-to-all exchanges so far and the bytes
 * this rank received from its peers in them */
int32_t zkpor_ctx_comm_info(zkpor_ctx *ctx, int32_t *out_rank, int32_t *out_world, uint64_t out_stats[2]);
/* desc describes the WHOLE key (host or device arrays, as for zkpor_pk_upload); only this rank's chunks are copied to its GPU */
int32_t zkpor_pk_upload_shard(zkpor_ctx *ctx, const zkpor_pk_desc *desc, zkpor_pk **out);
/* shard of a key: rank, world, first wire, wires, then the lengths of its A, B, K and Z chunks */
int32_t zkpor_pk_shard_info(zkpor_pk *pk, uint64_t out8[8]);
/* computeH across the group (collective): a, b, c = this rank's n/N evaluation rows rank, rank + N, rank + 2N, ... (zero beyond
 * the constraint count); out_h_chunk = elements [rank n/N, (rank+1) n/N) of zkpor_compute_h's output */
int32_t zkpor_compute_h_sharded(zkpor_ctx *ctx, const void *a, const void *b, const void *c, uint32_t log_n, void *out_h_chunk);
/* convenience for the one-process form: runs zkpor_groth16_prove_solve on the n contexts from n host threads and returns rank 0's
 * proof (pks[i] / progs[i] live on ctxs[i]; inputs, r, s as for zkpor_groth16_prove_solve) */
int32_t zkpor_multi_prove_solve(zkpor_ctx **ctxs, zkpor_pk **pks, zkpor_program **progs, int32_t n, const void *inputs,
                                const uint8_t r_be[32], const uint8_t s_be[32], uint8_t *out_proof, uint32_t *out_len);

/* ---- pairing / groth16.Verify --------------------------------------------------------------
 * Replaces gnark-crypto bn254.MillerLoop / FinalExponentiation / PairingCheck (ecc/bn254/pairing.go, out of tree) under
 * groth16.Verify -- src/prover/prover/prover.go:276, src/verifier/main.go:284.  Miller loops run one pair per GPU thread;
 * the product and the one final exponentiation are O(1) host work.
 * out_gt384 = prod_i e(P_i, Q_i) as 12 Montgomery Fp elements in gnark-crypto E12 memory order (C0.B0.A0 ... C1.B2.A1),
 * raised to exactly (q^12-1)/r (gnark's GT value is a fixed power of this one; equalities between products agree). */
int32_t zkpor_pairing_product(zkpor_ctx *ctx, const void *g1_points /* n x 64 B */, const void *g2_points /* n x 128 B */, uint64_t n,
                              void *out_gt384 /* host */);
/* *out_ok = 1 iff prod_i e(P_i, Q_i) == 1 (bn254.PairingCheck) */
int32_t zkpor_pairing_check(zkpor_ctx *ctx, const void *g1_points, const void *g2_points, uint64_t n, int32_t *out_ok);

/* groth16.VerifyingKey as gnark holds it in memory (points affine Montgomery; host pointers). */
typedef struct {
    const void *g1_alpha;                         /* vk.G1.Alpha                                                    */
    const void *g2_beta, *g2_gamma, *g2_delta;    /* vk.G2.Beta / Gamma / Delta                                     */
    const void *g1_k; uint64_t n_k;               /* vk.G1.K: ONE wire, public inputs, then one entry per commitment */
    uint64_t n_commitments;                       /* len(vk.PublicAndCommitmentCommitted): 0 or 1 here              */
    const uint64_t *public_committed;             /* vk.PublicAndCommitmentCommitted[0] (1-based public wire ids)   */
    uint64_t n_public_committed;
    const void *g2_ped_g, *g2_ped_g_root_sigma_neg;   /* vk.CommitmentKey.G / .GRootSigmaNeg (pedersen.VerifyingKey)  */
} zkpor_vk_desc;
/* groth16.Verify(proof, vk, publicWitness): proof = the bytes of proof.WriteRawTo (zkpor_groth16_prove's output),
 * public_witness = n_public Montgomery fr.Elements (without the ONE wire).  Recomputes the BSB22 commitment challenge
 * (hash_to_field, DST "bsb22-commitment"), checks the Pedersen proof of knowledge and the Groth16 pairing equation.
 * *out_ok = 1 valid, 0 invalid (a malformed proof is an error return). */
int32_t zkpor_groth16_verify(zkpor_ctx *ctx, const zkpor_vk_desc *vk, const uint8_t *proof_raw, uint32_t proof_len,
                             const void *public_witness, uint64_t n_public, int32_t *out_ok);
/* Batch verifier: `count` proofs of one circuit (proof i at proofs_raw + i*proof_stride, its public witness at
 * public_witnesses + i*n_public*32) checked with ONE pairing product of count + 3 Miller loops: a random linear
 * combination rho_i (derived from `seed32` and the proofs, 128-bit) folds the Groth16 equations; the Pedersen checks
 * are folded the same way.  *out_ok = 1 iff every proof verifies (soundness error 2^-128 per batch). */
int32_t zkpor_groth16_verify_batch(zkpor_ctx *ctx, const zkpor_vk_desc *vk, const uint8_t *proofs_raw, uint32_t proof_len,
                                   uint64_t proof_stride, const void *public_witnesses, uint64_t n_public, uint64_t count,
                                   const uint8_t seed32[32], int32_t *out_ok);

/* ---- Poseidon / Merkle -----------------------------------------------------------------------------------------
 * Replaces the bnb-chain gnark-crypto fr/poseidon hashers (poseidon.Poseidon / PoseidonBytes / NewPoseidon, call
 * sites src/utils/account_tree.go:19,27, src/utils/utils.go:748, src/witness/main.go:181) as batch calls. */
/* output lane of the permutation (see DESIGN.md "Poseidon parity"): default 1 (the in-tree fixture), 0 = iden3 */
int32_t zkpor_poseidon_set_out_lane(zkpor_ctx *ctx, int32_t lane);
/* the permutation's parameters for width t = 2..13 as this library generates them (Grain LFSR, the iden3 / gnark-crypto fork set):
 * (8 + rounds_p) * t round constants, t * t MDS entries (row-major), Montgomery fr.Elements, host buffers -- what the in-circuit
 * gadget (gnark std/hash/poseidon, circuit/utils.go:19,48) must use to agree with the native hash */
int32_t zkpor_poseidon_constants(uint32_t t, void *out_round_constants, void *out_mds, uint32_t *out_rounds_p);
/* count independent hashes of n_in big-endian 32-byte elements each -> count x 32 B (PoseidonBytes semantics) */
int32_t zkpor_poseidon_hash_batch(zkpor_ctx *ctx, const void *in_be, uint32_t n_in, uint64_t count, void *out_be);
/* utils.AccountInfoToHash for a batch of accounts of one asset tier (src/utils/utils.go:744-750,188-221):
 * ids = n x 32 B BE, totals = n x 3 x 32 B BE (equity, debt, collateral), flat_assets = n x tier*6 u64 already laid
 * out by PaddingAccountAssets (host logic).  out = n x 32 B leaf hashes. */
int32_t zkpor_account_leaves(zkpor_ctx *ctx, const void *ids_be, const void *totals_be, const void *flat_assets, uint64_t n,
                             uint32_t tier, void *out_be);

/* merkletree.NewFixedDepthMerkleTree / Set / Build / Root / GetProof (src/utils/merkletree/merkletree.go:137-308).
 * The tree lives in HBM; leaves are uploaded in ranges (Set), Build hashes all levels, proofs are gathered in
 * batches. */
int32_t zkpor_tree_create(zkpor_ctx *ctx, uint32_t depth, const uint8_t nil_leaf[32], uint64_t capacity, zkpor_tree **out);
int32_t zkpor_tree_free(zkpor_ctx *ctx, zkpor_tree *t);
int32_t zkpor_tree_set_range(zkpor_ctx *ctx, zkpor_tree *t, uint64_t first_key, uint64_t count, const void *leaves_be);
int32_t zkpor_tree_set_keys(zkpor_ctx *ctx, zkpor_tree *t, const uint32_t *keys, uint64_t count, const void *leaves_be);
int32_t zkpor_tree_build(zkpor_ctx *ctx, zkpor_tree *t);
int32_t zkpor_tree_root(zkpor_ctx *ctx, zkpor_tree *t, uint8_t out_root[32]);
int32_t zkpor_tree_get_leaves(zkpor_ctx *ctx, zkpor_tree *t, const uint32_t *keys, uint64_t count, void *out_be);
int32_t zkpor_tree_get_proofs(zkpor_ctx *ctx, zkpor_tree *t, const uint32_t *keys, uint64_t count, void *out_be /* count x depth x 32 */);
/* device pointer of level `level` (0 = leaves) and its length in nodes, for multi-GPU subtree exchange */
int32_t zkpor_tree_level(zkpor_ctx *ctx, zkpor_tree *t, uint32_t level, void **out_dev_ptr, uint64_t *out_len);

/* One tree over the GPUs of a group (contexts joined as for the sharded proof): rank g owns the leaves
 * [g << k, (g + 1) << k), k = the smallest level with world << k >= capacity.  Every rank creates the tree with the same depth, nil leaf
 * and capacity and sets only the leaves of its own range (zkpor_tree_shard_range); zkpor_tree_build_sharded (collective) builds the
 * subtrees, all-gathers the N subtree roots (32 B each) and finishes the top levels on every rank, so every rank has the root.  The
 * proofs of a key are served by the rank that owns it. */
int32_t zkpor_tree_shard_range(zkpor_ctx *ctx, zkpor_tree *t, uint64_t *out_first_key, uint64_t *out_count, uint32_t *out_subtree_level);
int32_t zkpor_tree_build_sharded(zkpor_ctx *ctx, zkpor_tree *t);

/* ---- witness batches -------------------------------------------------------------------------------------
 * Replaces the serial main loop of the witness service, src/witness/witness/witness.go:144-206, for all batches of one asset tier at
 * once: the running CEX totals that fillCreateUserOp accumulates account by account (:319-340) become per-batch sums + one scan, the
 * two 10 000-element Poseidon sponges per batch (:159-166, :176-183; the "after" state of a batch is the "before" state of the next)
 * become n_batches + 1 independent hashes, and the batch commitments (:193-198) one more batch of 5-input hashes.  The account proofs of
 * a batch are zkpor_tree_get_proofs on its account indices.  Host logic that stays in Go: gob + s2 serialisation and the DB writer. */
typedef struct {
    uint32_t n_assets;               /* utils.AssetCounts = 500 reserved CEX assets                                                 */
    const uint64_t *base_prices;     /* n_assets: CexAssetInfo.BasePrice                                                            */
    const void *tier_ratio_elems;    /* n_assets x 18 x 32 B big-endian: ConvertTierRatiosToBytes of Loan, Margin, PortfolioMargin
                                        ratios (src/utils/utils.go:26-51) -- static across batches                                   */
    const uint64_t *initial_totals;  /* n_assets x 5: TotalEquity, TotalDebt, LoanCollateral, MarginCollateral, PortfolioMargin-
                                        Collateral before the first batch                                                            */
} zkpor_cex_desc;
/* flat_assets = n_accounts x tier*6 u64 (the PaddingAccountAssets layout zkpor_account_leaves takes), accounts in batch order;
 * account_indices = their tree keys; n_accounts must be a multiple of ops_per_batch (the service pads the last batch,
 * src/witness/main.go:71-83).  With n_batches = n_accounts / ops_per_batch:
 *   out_totals            (n_batches + 1) x n_assets x 5 u64: the CEX totals before batch b (row n_batches: after the last one)
 *   out_cex_commitments   (n_batches + 1) x 32 B: BeforeCEXAssetsCommitment of batch b = row b, AfterCEXAssetsCommitment = row b + 1
 *   out_batch_commitments n_batches x 32 B: BatchCommitment
 * Outputs may be host or device pointers, or NULL.  A total that overflows 64 bits (utils.SafeAdd panics) or an asset index
 * >= n_assets is ZKPOR_ERR_STATE. */
int32_t zkpor_witness_batches(zkpor_ctx *ctx, const zkpor_cex_desc *cex, const uint8_t account_tree_root[32], const void *flat_assets,
                              const uint32_t *account_indices, uint64_t n_accounts, uint32_t tier, uint32_t ops_per_batch,
                              uint64_t *out_totals, void *out_cex_commitments, void *out_batch_commitments);

/* ---- point decoding (pk.UnsafeReadFrom's square roots) -----------------------------------
 * gnark-crypto bn254 encodings (ecc/bn254/marshal.go, out of tree): compressed = 32 B (G1) / 64 B (G2, X.A1 first),
 * raw = 64 B / 128 B, big-endian, flags in the top two bits of byte 0 (00 raw, 10/11 compressed smaller/larger y,
 * 01 infinity).  out = affine Montgomery points (host or device).  A point that is not on the curve, a coordinate
 * >= q or a flag that contradicts `compressed` fails the call with the first offending index in the message. */
int32_t zkpor_g1_decode_batch(zkpor_ctx *ctx, const void *in_bytes, uint64_t n, int32_t compressed, void *out_points);
int32_t zkpor_g2_decode_batch(zkpor_ctx *ctx, const void *in_bytes, uint64_t n, int32_t compressed, void *out_points);

/* ---- containers: the byte formats of proofs and keys (layouts in App. B.3) -------------------------
 * gnark's marshal.go and gnark-crypto's Encoder / Decoder are out of tree (go.mod:57-60).  Everything is big-endian; WriteTo writes
 * points compressed, WriteRawTo raw, and the decoder follows each point's flag bits, so either form is accepted on the way in; a slice
 * is a u32 length + elements; []bool is a u32 length + ceil(len / 8) bytes (bit i % 8 of byte i / 8).  The per-point work of a 2^26 key
 * (a square root per compressed point) runs on the GPU over the byte ranges of the caller's file buffer.  The r1cs container (CBOR +
 * intcomp) is not read here: gnark parses it and the shim hands over the flat program (zkpor_program_upload). */
int32_t zkpor_g1_encode_batch(zkpor_ctx *ctx, const void *points, uint64_t n, int32_t compressed, void *out_bytes);
int32_t zkpor_g2_encode_batch(zkpor_ctx *ctx, const void *points, uint64_t n, int32_t compressed, void *out_bytes);
/* groth16.Proof.ReadFrom (src/verifier/main.go:208-216): compressed or raw bytes -> the raw layout zkpor_groth16_verify takes
 * (Ar 64 | Bs 128 | Krs 64 | u32 nbCommitments | Commitments 64 each | CommitmentPok 64).  *out_len: capacity in, length out. */
int32_t zkpor_proof_decode(zkpor_ctx *ctx, const uint8_t *in, uint64_t in_len, uint8_t *out_raw, uint32_t *out_len, uint64_t *consumed);
/* Proof.WriteTo (compressed != 0: 196 bytes with one commitment) / WriteRawTo (src/prover/prover/prover.go:201) from either form */
int32_t zkpor_proof_encode(zkpor_ctx *ctx, const uint8_t *proof, uint32_t proof_len, int32_t compressed, uint8_t *out, uint32_t *out_len);
/* groth16.VerifyingKey as host data (points = affine Montgomery memory images, as zkpor_vk_desc points at them) */
typedef struct {
    uint8_t g1_alpha[64], g1_beta[64], g1_delta[64];
    uint8_t g2_beta[128], g2_gamma[128], g2_delta[128];
    uint8_t g2_ped_g[128], g2_ped_g_root_sigma_neg[128];   /* vk.CommitmentKey, valid when n_commitments = 1 */
    uint64_t n_k;                                          /* len(vk.G1.K) */
    uint64_t n_commitments;                                /* len(vk.PublicAndCommitmentCommitted): 0 or 1 */
    uint64_t n_public_committed;                           /* len(vk.PublicAndCommitmentCommitted[0]) */
} zkpor_vk_host;
/* vk.ReadFrom (src/prover/prover/prover.go:358-362, src/verifier/main.go:33-34): alpha1 beta1 beta2 gamma2 delta1 delta2 | K slice |
 * PublicAndCommitmentCommitted | Pedersen vk -- 524 bytes for the reference circuits (README.md:54,57).  k_points receives vk.G1.K. */
int32_t zkpor_vk_decode(zkpor_ctx *ctx, const uint8_t *in, uint64_t in_len, zkpor_vk_host *out, void *k_points, uint64_t k_cap,
                        uint64_t *public_committed, uint64_t pc_cap, uint64_t *consumed);
/* vk.WriteTo (raw = 0; src/keygen/main.go:46-62) / WriteRawTo.  out may be NULL to ask for the length. */
int32_t zkpor_vk_encode(zkpor_ctx *ctx, const zkpor_vk_host *vk, const void *k_points, const uint64_t *public_committed, int32_t raw,
                        uint8_t *out, uint64_t out_cap, uint64_t *out_len);
/* what a proving-key file does not hold and the resident key needs, from the constraint system */
typedef struct {
    uint64_t n_public;                   /* r1cs.GetNbPublicVariables(), includes the ONE wire */
    const uint64_t *private_committed;   /* CommitmentInfo[0].PrivateCommitted, ascending */
    uint64_t n_committed;
    uint64_t commitment_index;           /* CommitmentInfo[0].CommitmentIndex as a wire id */
} zkpor_pk_cs_info;
/* pk.ReadFrom / UnsafeReadFrom (src/prover/prover/prover.go:342-346) straight into HBM: fft.Domain header | alpha1 beta1 delta1 |
 * A[] B[] Z[] K[] | beta2 delta2 | B2[] | nbWires NbInfinityA NbInfinityB InfinityA[] InfinityB[] | u32 nbCommitmentKeys | Basis[]
 * BasisExpSigma[].  `in` is the file's bytes in host memory (a 12 GB mmap is fine: each array is staged and decoded in one launch).
 * Points are checked to be on the curve, not in the subgroup (UnsafeReadFrom's contract; G1 has cofactor 1). */
int32_t zkpor_pk_read(zkpor_ctx *ctx, const uint8_t *in, uint64_t in_len, const zkpor_pk_cs_info *info, zkpor_pk **out, uint64_t *consumed);
/* pk.WriteTo (raw = 0; src/keygen/main.go:46-62) / WriteRawTo from the resident key.  out may be NULL to ask for the length. */
int32_t zkpor_pk_write(zkpor_ctx *ctx, zkpor_pk *pk, int32_t raw, uint8_t *out, uint64_t out_cap, uint64_t *out_len);

/* ---- groth16.Setup building blocks (src/keygen/main.go:42) -------------------------------
 * curve.BatchScalarMultiplicationG1/G2: out[i] = scalars[i] * base (affine, host or device). */
int32_t zkpor_g1_fixed_base_batch(zkpor_ctx *ctx, const void *base_affine64, const void *scalars, uint64_t n, uint32_t flags, void *out_points);
int32_t zkpor_g2_fixed_base_batch(zkpor_ctx *ctx, const void *base_affine128, const void *scalars, uint64_t n, uint32_t flags, void *out_points);
/* Lagrange basis of the size-2^log_n domain at tau: out[k] = (tau^n - 1)/n * w^k/(tau - w^k)  (device, Montgomery) */
int32_t zkpor_setup_lagrange(zkpor_ctx *ctx, const uint8_t tau_be[32], uint32_t log_n, void *out_dev);
/* per-wire query evaluation: out[i] = sum_{e in column i} coeffs[e] * lagrange[rows[e]]; the R1CS matrix is passed in
 * compressed-sparse-column form (col_ptr has n_wires+1 entries) -- setupABC of gnark's setup.go */
int32_t zkpor_setup_wire_sums(zkpor_ctx *ctx, const uint64_t *col_ptr, const uint32_t *rows, const void *coeffs, uint64_t nnz,
                              const void *lagrange_dev, uint64_t n_wires, void *out_dev);
/* out = (ka*a + kb*b + kc*c) * k  element-wise (device vectors; scalars canonical big-endian) */
int32_t zkpor_fr_lincomb3(zkpor_ctx *ctx, const void *a, const void *b, const void *c, const uint8_t ka_be[32], const uint8_t kb_be[32],
                          const uint8_t kc_be[32], const uint8_t k_be[32], uint64_t n, void *out);
/* out[i] = first * ratio^i, or first * ratio^bitrev(i, log_n) when bitrev != 0 (pk.G1.Z is stored bit-reversed) */
int32_t zkpor_fr_powers(zkpor_ctx *ctx, const uint8_t first_be[32], const uint8_t ratio_be[32], uint64_t n, uint32_t log_n, int32_t bitrev, void *out_dev);

/* ---- synthetic workloads (bench / full-size parity tooling; not on the proving path) ----------------------------
 * points[i] = (k0 + i*d) * G with known discrete logs, written as affine Montgomery points into DEVICE memory;
 * scalars = counter-based uniform Fr (kind 0) or the witness-like mix of SURVEY.md 8(d) (kind 1). */
int32_t zkpor_synth_points_g1(zkpor_ctx *ctx, const uint8_t k0_be[32], const uint8_t d_be[32], uint64_t n, void *out_dev);
int32_t zkpor_synth_points_g2(zkpor_ctx *ctx, const uint8_t k0_be[32], const uint8_t d_be[32], uint64_t n, void *out_dev);
int32_t zkpor_synth_scalars(zkpor_ctx *ctx, uint64_t seed, uint64_t n, int32_t kind, void *out_dev);

#ifdef __cplusplus
}
#endif
#endif
