/* ORACLE (test infrastructure, NOT product code) -- public entry points of liborc.so (CPU restatement).
 * Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs ONLY.
 * Every function cites the reference interface it restates; see the .c files.
 *
 * Conventions:  field elements = 4 x u64 LE limbs, Montgomery form (gnark-crypto memory
 * layout); hashes / Merkle nodes = 32-byte big-endian canonical; G1 affine = X||Y (8 u64), G2 affine =
 * X.A0||X.A1||Y.A0||Y.A1 (16 u64); infinity = all zero.
 */
#ifndef ORC_H
#define ORC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int  orc_num_threads(void);

/* field helpers (which: 0 = Fp, 1 = Fr) */
void orc_to_mont(uint64_t *inout, size_t n, int which);
void orc_from_mont(uint64_t *inout, size_t n, int which);
void orc_fr_mul_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);

/* curve */
void orc_g1_fixed_base(const uint64_t *scalars_plain, size_t n, uint64_t *out_aff, int threads);   /* k_i * G1 */
void orc_g2_fixed_base(const uint64_t *scalars_plain, size_t n, uint64_t *out_aff, int threads);   /* k_i * G2 */
void orc_g1_msm(const uint64_t *pts, const uint64_t *scalars_mont, size_t n, uint64_t *out_aff, int threads);
void orc_g2_msm(const uint64_t *pts, const uint64_t *scalars_mont, size_t n, uint64_t *out_aff, int threads);
void orc_g1_add(const uint64_t *p, const uint64_t *q, uint64_t *out_aff);
void orc_g1_scalar_mul(const uint64_t *p, const uint64_t *k_plain, uint64_t *out_aff);
void orc_g2_add(const uint64_t *p, const uint64_t *q, uint64_t *out_aff);
void orc_g2_scalar_mul(const uint64_t *p, const uint64_t *k_plain, uint64_t *out_aff);
int  orc_g1_on_curve(const uint64_t *p);
int  orc_g2_on_curve(const uint64_t *p);

/* NTT (gnark-crypto fft.Domain conventions) */
void orc_ntt(uint64_t *data_mont, int logn, int inverse, int dit, int coset, int threads);
void orc_compute_h(const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t m, int logn, uint64_t *out_h, int threads);

/* full-size check helpers */
void orc_fr_index_sums(const uint64_t *v, size_t n, uint64_t *out_sum, uint64_t *out_isum, int threads);
void orc_eval_barycentric(const uint64_t *evals_mont, size_t m, int logn, const uint64_t *x0_mont, uint64_t *out_mont, int threads);
void orc_poly_eval_bitrev(const uint64_t *coef_mont, int logn, const uint64_t *x0_mont, uint64_t *out_mont, int threads);

/* Poseidon / Merkle */
void orc_poseidon_set_out_lane(int lane);
int  orc_poseidon_get_out_lane(void);
void orc_poseidon_constants(int t, uint64_t *rc_mont /* (8+RP)*t*4 */, uint64_t *mds_mont /* t*t*4 */, int *rounds_p);
void orc_poseidon_permute(uint64_t *state_mont, int t);
void orc_poseidon_hash(const uint64_t *in_mont, size_t n_in, uint64_t *out_mont);
void orc_poseidon_hash_be(const uint8_t *in_be, size_t n_in, uint8_t *out_be);
void orc_poseidon_node_batch(const uint8_t *pairs_be, size_t count, uint8_t *out_be, int threads);
size_t orc_merkle_level_len(size_t capacity, int level);
size_t orc_merkle_nodes_total(size_t capacity, int depth);
void orc_merkle_build(const uint8_t *leaves, const uint64_t *dirty, size_t capacity, int depth, const uint8_t *nil_leaf,
                      uint8_t *out_nodes, uint8_t *out_root, int threads);
void orc_merkle_proofs(const uint8_t *leaves, const uint64_t *dirty, const uint8_t *nodes, size_t capacity, int depth,
                       const uint8_t *nil_leaf, const uint32_t *keys, size_t nkeys, uint8_t *out);
void orc_account_leaves(const uint8_t *ids_be, const uint8_t *totals_be, const uint64_t *flat_assets, size_t n_accounts,
                        int tier, uint8_t *out_be, int threads);

/* Groth16 prove (gnark v0.10 backend/groth16/bn254/prove.go restated) */
typedef struct {
    uint64_t n_a, n_b, n_k, n_z, n_ck;        /* point counts: A, B1(=B2), K, Z (= domain-1), commitment basis */
    const uint64_t *A, *B1, *K, *Z, *B2;      /* affine, Montgomery */
    const uint64_t *ck_basis, *ck_basis_exp_sigma;
    const uint64_t *alpha1, *beta1, *delta1, *beta2, *delta2;
    int log_n;                                /* domain size */
} orc_pk;
int orc_groth16_prove(const orc_pk *pk, const uint64_t *wires_a, const uint64_t *wires_b, const uint64_t *wires_k,
                      const uint64_t *committed, const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n_constraints,
                      const uint64_t *r_plain, const uint64_t *s_plain, uint8_t *out_proof388, int threads);

/* r1cs.Solve over the flat program arrays (orc_solver.c; same layout as zkpor_program_desc / circuit_synth.flatten()) */
typedef struct {
    uint64_t n_wires, n_public, n_secret, n_constraints;
    const uint64_t *l_row_ptr; const uint32_t *l_wire, *l_coeff;
    const uint64_t *r_row_ptr; const uint32_t *r_wire, *r_coeff;
    const uint64_t *o_row_ptr; const uint32_t *o_wire, *o_coeff;
    const uint64_t *coeffs; uint64_t n_coeffs;                       /* Montgomery */
    uint64_t n_instr; const uint8_t *instr_kind; const uint32_t *instr_arg;
    uint64_t n_levels; const uint64_t *level_ptr; const uint32_t *level_instr;
    uint64_t n_hints; const uint32_t *hint_fn, *hint_param, *hint_out_first, *hint_n_out; const uint64_t *hint_in_ptr, *hint_in_end;
    const uint64_t *aux_row_ptr; const uint32_t *aux_wire, *aux_coeff;
    const uint64_t *table_ptr;
    const uint64_t *private_committed; uint64_t n_committed;
} orc_program;
/* the overridden BSB22 hint: committed values (Montgomery) -> challenge (Montgomery) */
typedef void (*orc_commit_fn)(const uint64_t *committed_mont, size_t n, uint64_t *challenge_mont, void *user);
/* returns 0, or: 1 more than one unsolved wire, 2 division by zero, 3 index outside a table, 4 unknown hint, 5 hint reads an unsolved
 * wire, 6 commitment hint without callback, 7 committed wire unsolved, 8 wire never solved, 9 constraint not satisfied (*err_at) */
int orc_solve(const orc_program *p, const uint64_t *inputs_mont, uint64_t *wires_mont, uint64_t *a, uint64_t *b, uint64_t *c,
              orc_commit_fn commit, void *user, uint64_t *err_at, int threads);
void orc_commitment_challenge(const uint8_t *msg, size_t len, uint64_t *out_mont);
/* the whole of groth16.Prove: solve (commitment mid-solve), filter the wires by infinity_a / infinity_b / public+committed, prove.
 * seconds[0] = solver, seconds[1] = the rest */
int orc_groth16_prove_program(const orc_pk *pk, const orc_program *prog, const uint8_t *infinity_a, const uint8_t *infinity_b,
                              uint64_t commitment_index, const uint64_t *inputs_mont, const uint64_t *r_plain, const uint64_t *s_plain,
                              uint8_t *out_proof388, double *seconds, uint64_t *err_at, int threads);

#ifdef __cplusplus
}
#endif
#endif
