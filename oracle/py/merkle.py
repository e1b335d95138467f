"""ORACLE (test infrastructure, NOT product code) -- FixedDepthMerkleTree and account-leaf hashing.

Restates, function by function:
  src/utils/merkletree/merkletree.go:137-174  NewFixedDepthMerkleTree (nilHashes chain)
  src/utils/merkletree/merkletree.go:179-187  Set
  src/utils/merkletree/merkletree.go:192-279  Build (dirty propagation, nil-subtree shortcut)
  src/utils/merkletree/merkletree.go:297-308  GetProof
  src/utils/merkletree/merkletree.go:334-355  VerifyProof
  src/utils/utils.go:147-186                  PaddingAccountAssets
  src/utils/utils.go:188-221                  ComputeUserAssetsCommitment
  src/utils/utils.go:744-750                  AccountInfoToHash
  src/utils/utils.go:26-88,779-800            ConvertTierRatiosToBytes / ConvertAssetInfoToBytes / ComputeCexAssetsCommitment
  src/utils/constants.go:125-127              NilAccountHash = Poseidon(0,0,0,0,0)
  src/witness/main.go:71-83                   padding-account ids = Fr(sha256(BE32(index)))
"""
from __future__ import annotations

import hashlib

from bn254 import R
from poseidon import PoseidonHasher, poseidon, poseidon_bytes

ACCOUNT_TREE_DEPTH = 28
ASSET_COUNTS = 500
TIER_COUNT = 12
ASSET_TIERS = (50, 500)  # src/utils/constants.go:103-106
U64 = 1 << 64
U128 = 1 << 128


def nil_account_hash(out_lane=None) -> bytes:
    return poseidon([0, 0, 0, 0, 0], out_lane).to_bytes(32, "big")


class FixedDepthMerkleTree:
    def __init__(self, depth: int, nil_leaf: bytes, capacity: int, out_lane=None):
        if depth > 32:
            raise ValueError("depth too large")
        if depth <= 0:
            raise ValueError("depth must be positive")
        if capacity > (1 << depth):
            raise ValueError("capacity exceeds maximum for given depth")
        self.depth, self.capacity, self.lane = depth, capacity, out_lane
        self.leaves = {}
        self.levels = [dict() for _ in range(depth + 1)]  # only dirty (computed) positions
        self.nil = [nil_leaf]
        h = PoseidonHasher(out_lane)
        for _ in range(depth):
            h.reset()
            h.write(self.nil[-1])
            h.write(self.nil[-1])
            self.nil.append(h.sum())
        self.root = self.nil[depth]

    def set(self, key: int, value: bytes):
        if key >= self.capacity:
            raise IndexError(f"key {key} out of range for capacity {self.capacity}")
        self.leaves[key] = bytes(value)

    def _node(self, level: int, pos: int) -> bytes:
        if level == 0:
            return self.leaves.get(pos, self.nil[0])
        return self.levels[level].get(pos, self.nil[level])

    def build(self):
        dirty = sorted({k >> 1 for k in self.leaves})
        h = PoseidonHasher(self.lane)
        for level in range(1, self.depth + 1):
            if not dirty:
                break
            cur = {}
            self.levels[level] = cur
            for pos in dirty:
                h.reset()
                h.write(self._node(level - 1, pos << 1))
                h.write(self._node(level - 1, (pos << 1) | 1))
                cur[pos] = h.sum()
            dirty = sorted({p >> 1 for p in dirty})
        self.root = self.levels[self.depth].get(0, self.nil[self.depth])

    def get(self, key: int) -> bytes:
        return self.leaves.get(key, self.nil[0]) if key < self.capacity else self.nil[0]

    def get_proof(self, key: int):
        if key >= (1 << self.depth):
            raise IndexError("key out of range for tree depth")
        proof, pos = [], key
        for level in range(self.depth):
            proof.append(self._node(level, pos ^ 1))
            pos >>= 1
        return proof


def verify_proof(root: bytes, key: int, proof, leaf: bytes, depth: int, out_lane=None) -> bool:
    if len(proof) != depth or key >= (1 << depth):
        return False
    node = leaf
    h = PoseidonHasher(out_lane)
    for i in range(depth):
        h.reset()
        if key & (1 << i) == 0:
            h.write(node); h.write(proof[i])
        else:
            h.write(proof[i]); h.write(node)
        node = h.sum()
    return node == root


# ----------------------------------------------------------------------------- leaves
def assets_count_tier(n_assets: int) -> int:
    """GetAssetsCountOfUser: the smallest tier that holds the user's assets (src/utils/utils.go:128-145)."""
    for t in ASSET_TIERS:
        if n_assets <= t:
            return t
    raise ValueError("too many assets")


def padding_account_assets(assets):
    """assets: list of (index, equity, debt, loan, margin, pm), index strictly increasing.
    Returns the flat uint64 list of targetCounts*6 entries; gaps are filled with the lowest unused
    indices first (src/utils/utils.go:147-186)."""
    target = assets_count_tier(len(assets))
    flat = [0] * (target * 6)
    padding = target - len(assets)
    cur_pad, cur_idx, index = 0, 0, 0
    for a in assets:
        if cur_pad < padding:
            for j in range(cur_idx, a[0]):
                cur_pad += 1
                flat[index * 6] = j
                index += 1
                if cur_pad >= padding:
                    break
        flat[index * 6:index * 6 + 6] = list(a)
        index += 1
        cur_idx = a[0] + 1
    for i in range(index, target):
        flat[i * 6] = cur_idx
        cur_idx += 1
    return flat


def pack_triples(flat):
    """a*2^128 + b*2^64 + c per 3 uint64 (src/utils/utils.go:196-218); missing tail entries are 0."""
    n = (len(flat) + 2) // 3
    out = []
    for i in range(n):
        a = flat[3 * i] if 3 * i < len(flat) else 0
        b = flat[3 * i + 1] if 3 * i + 1 < len(flat) else 0
        c = flat[3 * i + 2] if 3 * i + 2 < len(flat) else 0
        out.append(a * U128 + b * U64 + c)
    return out


def user_assets_commitment(assets, out_lane=None) -> bytes:
    return poseidon(pack_triples(padding_account_assets(assets)), out_lane).to_bytes(32, "big")


def account_leaf(account_id: bytes, total_equity: int, total_debt: int, total_collateral: int, assets,
                 out_lane=None) -> bytes:
    """AccountInfoToHash: Poseidon5(id, equity, debt, collateral, assetsCommitment); big.Int.Bytes() of 0 is
    the empty slice, which maps to Fr 0."""
    ac = user_assets_commitment(assets, out_lane)
    return poseidon_bytes([account_id,
                           total_equity.to_bytes((total_equity.bit_length() + 7) // 8, "big"),
                           total_debt.to_bytes((total_debt.bit_length() + 7) // 8, "big"),
                           total_collateral.to_bytes((total_collateral.bit_length() + 7) // 8, "big"),
                           ac], out_lane)


def padding_account_id(index: int) -> bytes:
    """src/witness/main.go:75-79: id = Fr(sha256(BE32(index))).Bytes()."""
    return (int.from_bytes(hashlib.sha256(index.to_bytes(4, "big")).digest(), "big") % R).to_bytes(32, "big")


# ----------------------------------------------------------------------------- CEX commitment
def tier_ratios_packed(tiers):
    """tiers: list of (boundary, ratio), even length.  ratio0 + bnd0*2^8 + ratio1*2^126 + bnd1*2^134."""
    out = []
    for i in range(0, len(tiers), 2):
        b0, r0 = tiers[i]
        b1, r1 = tiers[i + 1]
        out.append(r0 + b0 * (1 << 8) + r1 * (1 << 126) + b1 * (1 << 134))
    return out


def cex_asset_packed(a):
    """a: dict(total_equity, total_debt, base_price, loan, margin, pm, loan_ratios, margin_ratios, pm_ratios)
    -> 2 + 3*TIER_COUNT/2 = 20 field elements (src/utils/utils.go:53-88)."""
    out = [a["total_equity"] * U128 + a["total_debt"] * U64 + a["base_price"],
           a["loan"] * U128 + a["margin"] * U64 + a["pm"]]
    out += tier_ratios_packed(a["loan_ratios"]) + tier_ratios_packed(a["margin_ratios"]) + tier_ratios_packed(a["pm_ratios"])
    return out


def cex_assets_commitment(cex_assets, out_lane=None) -> bytes:
    """ComputeCexAssetsCommitment: pad to 500 reserved assets, hash all 10 000 packed elements."""
    empty = dict(total_equity=0, total_debt=0, base_price=0, loan=0, margin=0, pm=0,
                 loan_ratios=[(0, 0)] * TIER_COUNT, margin_ratios=[(0, 0)] * TIER_COUNT, pm_ratios=[(0, 0)] * TIER_COUNT)
    full = list(cex_assets) + [empty] * (ASSET_COUNTS - len(cex_assets))
    elems = []
    for a in full:
        elems += cex_asset_packed(a)
    return poseidon(elems, out_lane).to_bytes(32, "big")


# ----------------------------------------------------------------------------- witness batches
def witness_batches(cex_assets, root: bytes, accounts, ops_per_batch: int, out_lane=None):
    """The witness service's main loop, src/witness/witness/witness.go:144-206, restated as it is written: batch after batch, the CEX
    state is hashed (BeforeCEXAssetsCommitment), every account's assets are added to the running totals (fillCreateUserOp,
    :319-340), the state is hashed again (AfterCEXAssetsCommitment) and BatchCommitment = Poseidon(root, before, after, min, max)
    (:185-198; a zero index is []byte{0}, the element 0).
    cex_assets: list of dicts as cex_asset_packed takes (mutated copies are returned); accounts: list of (account_index, assets) with
    assets = [(index, equity, debt, loan, margin, pm)], a multiple of ops_per_batch of them.
    Returns [(before_totals, before_commitment, after_commitment, batch_commitment)] per batch and the final totals."""
    assert len(accounts) % ops_per_batch == 0
    state = [dict(a) for a in cex_assets]
    out = []
    for b in range(len(accounts) // ops_per_batch):
        before_totals = [(a["total_equity"], a["total_debt"], a["loan"], a["margin"], a["pm"]) for a in state]
        before = cex_assets_commitment(state, out_lane)
        batch = accounts[b * ops_per_batch:(b + 1) * ops_per_batch]
        for _, assets in batch:
            for (idx, eq, debt, loan, margin, pm) in assets:
                a = state[idx]
                for k, v in (("total_equity", eq), ("total_debt", debt), ("loan", loan), ("margin", margin), ("pm", pm)):
                    a[k] += v
                    assert a[k] < (1 << 64), "utils.SafeAdd: overflow"
        after = cex_assets_commitment(state, out_lane)
        lo, hi = batch[0][0], batch[-1][0]
        bc = poseidon_bytes([root, before, after, lo.to_bytes(max(1, (lo.bit_length() + 7) // 8), "big"),
                             hi.to_bytes(max(1, (hi.bit_length() + 7) // 8), "big")], out_lane)
        out.append((before_totals, before, after, bc))
    return out, [(a["total_equity"], a["total_debt"], a["loan"], a["margin"], a["pm"]) for a in state]
