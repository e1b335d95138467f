// C++ host-side mirror of the reference interfaces that libzkpor_b200 replaces (header-only, above the C-ABI in
// zkpor_b200.h).  The reference is Go and there is no Go toolchain in the build image, so this is the compiled-language
// host layer; names, argument meaning and error behaviour follow the reference:
//
//   merkletree.NewFixedDepthMerkleTree / Set / Build / Root / Get / GetProof / VerifyProof   src/utils/merkletree/merkletree.go:137-355
//   poseidon.PoseidonBytes                                                                   src/utils/utils.go:748
//   utils.PaddingAccountAssets / utils.AccountInfoToHash                                     src/utils/utils.go:147-186,744-750
//   groth16.Prove (after the solver)                                                         src/prover/prover/prover.go:269
//
// Go panics / returned errors become C++ exceptions (zkpor::Error).  Nothing here computes a hash or a group operation
// on the CPU: every call forwards to the GPU library and throws if no CUDA device is present.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "zkpor_b200.h"

namespace zkpor {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int32_t rc) { if (rc != ZKPOR_OK) throw Error(std::string("zkpor: ") + zkpor_last_error()); }

using Hash = std::array<uint8_t, 32>;   // 32-byte big-endian canonical field element

class Context {
  public:
    explicit Context(int device = 0) { check(zkpor_ctx_create(device, &h_)); }
    ~Context() { zkpor_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    zkpor_ctx *get() const { return h_; }
  private:
    zkpor_ctx *h_ = nullptr;
};

namespace utils {
// src/utils/constants.go:103-106
inline int GetAssetsCountOfUser(size_t n_assets) {
    if (n_assets <= 50) return 50;
    if (n_assets <= 500) return 500;
    throw Error("the target counts is less than the length of assets");
}
struct AccountAsset { uint16_t Index; uint64_t Equity, Debt, Loan, Margin, PortfolioMargin; };
// src/utils/utils.go:147-186 -- gaps are filled with the lowest unused asset indices first
inline std::vector<uint64_t> PaddingAccountAssets(const std::vector<AccountAsset> &assets) {
    const size_t target = (size_t)GetAssetsCountOfUser(assets.size()), fields = 6;
    std::vector<uint64_t> flat(target * fields, 0);
    const size_t padding = target - assets.size();
    size_t cur_pad = 0, cur_idx = 0, index = 0;
    for (const auto &a : assets) {
        if (cur_pad < padding) {
            for (size_t j = cur_idx; j < a.Index; j++) {
                cur_pad++; flat[index * fields] = j; index++;
                if (cur_pad >= padding) break;
            }
        }
        uint64_t *f = &flat[index * fields];
        f[0] = a.Index; f[1] = a.Equity; f[2] = a.Debt; f[3] = a.Loan; f[4] = a.Margin; f[5] = a.PortfolioMargin;
        index++; cur_idx = (size_t)a.Index + 1;
    }
    for (size_t i = index; i < target; i++) flat[i * fields] = cur_idx++;
    return flat;
}
// AccountInfoToHash for a batch of accounts of one tier; ids / totals are 32-byte big-endian (big.Int zero = all zero)
inline std::vector<Hash> AccountInfoToHashBatch(Context &ctx, const std::vector<Hash> &ids, const std::vector<std::array<Hash, 3>> &totals,
                                                const std::vector<uint64_t> &flat_assets, uint32_t tier) {
    std::vector<Hash> out(ids.size());
    if (flat_assets.size() != ids.size() * tier * 6 || totals.size() != ids.size()) throw Error("AccountInfoToHashBatch: size mismatch");
    check(zkpor_account_leaves(ctx.get(), ids.data(), totals.data(), flat_assets.data(), ids.size(), tier, out.data()));
    return out;
}
}  // namespace utils

namespace poseidon {
// poseidon.PoseidonBytes(input ...[]byte): every chunk is one big-endian element (shorter chunks are left-padded)
inline Hash PoseidonBytes(Context &ctx, const std::vector<std::vector<uint8_t>> &input) {
    std::vector<uint8_t> buf(input.size() * 32, 0);
    for (size_t i = 0; i < input.size(); i++) {
        if (input[i].size() > 32) throw Error("not support bytes bigger than modulus");
        std::copy(input[i].begin(), input[i].end(), buf.begin() + 32 * i + (32 - input[i].size()));
    }
    Hash out;
    check(zkpor_poseidon_hash_batch(ctx.get(), buf.data(), (uint32_t)input.size(), 1, out.data()));
    return out;
}
}  // namespace poseidon

namespace merkletree {
class FixedDepthMerkleTree {
  public:
    // NewFixedDepthMerkleTree(depth, nilLeafHash, hasherFunc, capacity): the hasher is the GPU Poseidon
    FixedDepthMerkleTree(Context &ctx, int depth, const Hash &nil_leaf, uint64_t capacity) : ctx_(ctx), depth_(depth), capacity_(capacity) {
        if (depth > 32) throw Error("depth too large");
        if (depth <= 0) throw Error("depth must be positive");
        if (capacity > (1ull << depth)) throw Error("capacity exceeds maximum for given depth");
        check(zkpor_tree_create(ctx.get(), (uint32_t)depth, nil_leaf.data(), capacity, &h_));
    }
    ~FixedDepthMerkleTree() { zkpor_tree_free(ctx_.get(), h_); }
    FixedDepthMerkleTree(const FixedDepthMerkleTree &) = delete;
    void Set(uint32_t key, const Hash &value) {
        if (key >= capacity_) throw Error("key " + std::to_string(key) + " out of range for capacity " + std::to_string(capacity_));
        check(zkpor_tree_set_range(ctx_.get(), h_, key, 1, value.data()));
    }
    void SetRange(uint64_t first_key, const std::vector<Hash> &values) { check(zkpor_tree_set_range(ctx_.get(), h_, first_key, values.size(), values.data())); }
    void Build() { check(zkpor_tree_build(ctx_.get(), h_)); }
    Hash Root() const { Hash r; check(zkpor_tree_root(ctx_.get(), h_, r.data())); return r; }
    Hash Get(uint32_t key) const { Hash r; check(zkpor_tree_get_leaves(ctx_.get(), h_, &key, 1, r.data())); return r; }
    std::vector<Hash> GetProof(uint32_t key) const {
        if ((uint64_t)key >= (1ull << depth_)) throw Error("key " + std::to_string(key) + " out of range for tree depth " + std::to_string(depth_));
        std::vector<Hash> p((size_t)depth_);
        check(zkpor_tree_get_proofs(ctx_.get(), h_, &key, 1, p.data()));
        return p;
    }
  private:
    Context &ctx_; zkpor_tree *h_ = nullptr; int depth_; uint64_t capacity_;
};
// VerifyProof(root, key, proof, leaf, depth, hasherFunc)
inline bool VerifyProof(Context &ctx, const Hash &root, uint32_t key, const std::vector<Hash> &proof, const Hash &leaf, int depth) {
    if ((int)proof.size() != depth || (uint64_t)key >= (1ull << depth)) return false;
    Hash node = leaf;
    for (int i = 0; i < depth; i++) {
        uint8_t pair[64];
        const Hash &l = (key & (1u << i)) ? proof[i] : node, &r = (key & (1u << i)) ? node : proof[i];
        std::copy(l.begin(), l.end(), pair); std::copy(r.begin(), r.end(), pair + 32);
        check(zkpor_poseidon_hash_batch(ctx.get(), pair, 2, 1, node.data()));
    }
    return node == root;
}
}  // namespace merkletree

namespace groth16 {
class ProvingKey {   // groth16.ProvingKey resident in HBM
  public:
    ProvingKey(Context &ctx, const zkpor_pk_desc &desc) : ctx_(ctx) { check(zkpor_pk_upload(ctx.get(), &desc, &h_)); }
    ~ProvingKey() { zkpor_pk_free(ctx_.get(), h_); }
    ProvingKey(const ProvingKey &) = delete;
    zkpor_pk *get() const { return h_; }
  private:
    Context &ctx_; zkpor_pk *h_ = nullptr;
};
// groth16.Prove after the solver; returns proof.WriteRawTo bytes.  r, s: 32-byte big-endian canonical.
inline std::vector<uint8_t> Prove(Context &ctx, ProvingKey &pk, const void *wires, const void *a, const void *b, const void *c,
                                  uint64_t n_constraints, const Hash &r, const Hash &s) {
    std::vector<uint8_t> out(388); uint32_t n = 0;
    check(zkpor_groth16_prove(ctx.get(), pk.get(), wires, a, b, c, n_constraints, r.data(), s.data(), out.data(), &n));
    out.resize(n);
    return out;
}
// R1CS matrices resident in HBM + groth16.Prove from the wire vector alone (a, b, c evaluated on the device)
class R1cs {
  public:
    R1cs(Context &ctx, uint64_t n_constraints, uint64_t n_wires, const zkpor_csr &l, const zkpor_csr &r, const zkpor_csr &o, const void *coeff_table,
         uint64_t n_coeffs) : ctx_(ctx) { check(zkpor_r1cs_upload(ctx.get(), n_constraints, n_wires, &l, &r, &o, coeff_table, n_coeffs, &h_)); }
    ~R1cs() { zkpor_r1cs_free(ctx_.get(), h_); }
    R1cs(const R1cs &) = delete;
    zkpor_r1cs *get() const { return h_; }
    void Eval(const void *wires, void *a, void *b, void *c) { check(zkpor_r1cs_eval(ctx_.get(), h_, wires, a, b, c)); }
  private:
    Context &ctx_; zkpor_r1cs *h_ = nullptr;
};
inline std::vector<uint8_t> ProveWires(Context &ctx, ProvingKey &pk, R1cs &cs, const void *wires, const Hash &r, const Hash &s) {
    std::vector<uint8_t> out(388); uint32_t n = 0;
    check(zkpor_groth16_prove_wires(ctx.get(), pk.get(), cs.get(), wires, r.data(), s.data(), out.data(), &n));
    out.resize(n);
    return out;
}
// gnark's compiled constraint system, flattened (zkpor_program_desc): the witness solver's program resident in HBM
class Program {
  public:
    Program(Context &ctx, const zkpor_program_desc &desc) : ctx_(ctx) { check(zkpor_program_upload(ctx.get(), &desc, &h_)); }
    ~Program() { zkpor_program_free(ctx_.get(), h_); }
    Program(const Program &) = delete;
    zkpor_program *get() const { return h_; }
    // r1cs.Solve: inputs = public then secret values; any output may be null
    void Solve(ProvingKey *pk, const void *inputs, void *wires, void *a, void *b, void *c, void *commitment64 = nullptr) {
        check(zkpor_r1cs_solve(ctx_.get(), h_, pk ? pk->get() : nullptr, inputs, wires, a, b, c, commitment64));
    }
  private:
    Context &ctx_; zkpor_program *h_ = nullptr;
};
// the whole of groth16.Prove (src/prover/prover/prover.go:269): witness solver with its hints, commitment mid-solve, proof
inline std::vector<uint8_t> ProveSolve(Context &ctx, ProvingKey &pk, Program &prog, const void *inputs, const Hash &r, const Hash &s) {
    std::vector<uint8_t> out(388); uint32_t n = 0;
    check(zkpor_groth16_prove_solve(ctx.get(), pk.get(), prog.get(), inputs, r.data(), s.data(), out.data(), &n));
    out.resize(n);
    return out;
}
// proof.ReadFrom (src/verifier/main.go:208-216): compressed or raw bytes -> the raw layout Verify takes; WriteTo / WriteRawTo back
inline std::vector<uint8_t> DecodeProof(Context &ctx, const std::vector<uint8_t> &bytes) {
    std::vector<uint8_t> out(260 + 64 * 17); uint32_t n = (uint32_t)out.size();
    check(zkpor_proof_decode(ctx.get(), bytes.data(), bytes.size(), out.data(), &n, nullptr));
    out.resize(n);
    return out;
}
inline std::vector<uint8_t> EncodeProof(Context &ctx, const std::vector<uint8_t> &proof, bool compressed) {
    std::vector<uint8_t> out(260 + 64 * 17); uint32_t n = (uint32_t)out.size();
    check(zkpor_proof_encode(ctx.get(), proof.data(), (uint32_t)proof.size(), compressed ? 1 : 0, out.data(), &n));
    out.resize(n);
    return out;
}
// pk.WriteTo / WriteRawTo from the resident key (src/keygen/main.go:46-62)
inline std::vector<uint8_t> WriteProvingKey(Context &ctx, ProvingKey &pk, bool raw) {
    uint64_t n = 0;
    check(zkpor_pk_write(ctx.get(), pk.get(), raw ? 1 : 0, nullptr, 0, &n));
    std::vector<uint8_t> out(n);
    check(zkpor_pk_write(ctx.get(), pk.get(), raw ? 1 : 0, out.data(), out.size(), &n));
    return out;
}
// groth16.Verify (src/prover/prover/prover.go:276, src/verifier/main.go:284): true = valid; a malformed proof throws.
// public_witness: n_public Montgomery fr.Elements, without the ONE wire.
inline bool Verify(Context &ctx, const zkpor_vk_desc &vk, const std::vector<uint8_t> &proof_raw, const void *public_witness, uint64_t n_public) {
    int32_t ok = 0;
    check(zkpor_groth16_verify(ctx.get(), &vk, proof_raw.data(), (uint32_t)proof_raw.size(), public_witness, n_public, &ok));
    return ok != 0;
}
// every proof of one circuit with a single pairing product (the verifier service's loop, src/verifier/main.go:176-302)
inline bool VerifyBatch(Context &ctx, const zkpor_vk_desc &vk, const std::vector<std::vector<uint8_t>> &proofs, const void *public_witnesses,
                        uint64_t n_public, const Hash &seed) {
    if (proofs.empty()) return true;
    const size_t len = proofs[0].size();
    std::vector<uint8_t> flat(len * proofs.size());
    for (size_t i = 0; i < proofs.size(); i++) {
        if (proofs[i].size() != len) throw Error("zkpor: proofs of one batch must have one length");
        std::copy(proofs[i].begin(), proofs[i].end(), flat.begin() + i * len);
    }
    int32_t ok = 0;
    check(zkpor_groth16_verify_batch(ctx.get(), &vk, flat.data(), (uint32_t)len, len, public_witnesses, n_public, proofs.size(), seed.data(), &ok));
    return ok != 0;
}
}  // namespace groth16

namespace witness {
// the witness service's main loop for all batches of one tier (src/witness/witness/witness.go:144-206)
struct Batches { std::vector<uint64_t> totals; std::vector<Hash> cex_commitments, batch_commitments; };
inline Batches RunBatches(Context &ctx, const zkpor_cex_desc &cex, const Hash &root, const std::vector<uint64_t> &flat_assets,
                          const std::vector<uint32_t> &account_indices, uint32_t tier, uint32_t ops_per_batch) {
    if (ops_per_batch == 0 || account_indices.size() % ops_per_batch != 0) throw Error("zkpor: the accounts must fill whole batches");
    const size_t nb = account_indices.size() / ops_per_batch;
    Batches out;
    out.totals.resize((nb + 1) * (size_t)cex.n_assets * 5); out.cex_commitments.resize(nb + 1); out.batch_commitments.resize(nb);
    check(zkpor_witness_batches(ctx.get(), &cex, root.data(), flat_assets.data(), account_indices.data(), account_indices.size(), tier, ops_per_batch,
                                out.totals.data(), out.cex_commitments.data(), out.batch_commitments.data()));
    return out;
}
}  // namespace witness

// bn254.PairingCheck: prod e(P_i, Q_i) == 1
inline bool PairingCheck(Context &ctx, const void *g1_points, const void *g2_points, uint64_t n) {
    int32_t ok = 0;
    check(zkpor_pairing_check(ctx.get(), g1_points, g2_points, n, &ok));
    return ok != 0;
}

}  // namespace zkpor
