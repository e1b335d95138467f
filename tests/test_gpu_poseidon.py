"""GPU parity: batched Poseidon, account leaves, FixedDepthMerkleTree through the C-ABI against the oracle and the
reference's own fixture (src/verifier/config/user_config.json -> tests/golden/user_config_proof.json)."""
import base64
import json
import os

import numpy as np
import pytest

import merkle
import orc
import poseidon as ps
import zkpor_b200 as zk
from bn254 import R, SplitMix64
from helpers import GOLDEN, H, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def test_reference_fixture_chain_on_gpu(ctx):
    fx = json.load(open(os.path.join(GOLDEN, "user_config_proof.json")))
    pr = [base64.b64decode(x) for x in fx["Proof"]]
    ctx.set_poseidon_out_lane(1)
    for i in range(15, 27):
        assert ctx.poseidon_bytes(pr[i], pr[i]) == pr[i + 1]


def test_reference_fixture_wide_sponge_on_gpu(ctx):
    """The fixture's empty-subtree chain from its nil leaf: Poseidon(0,0,0,0, Poseidon(0 x 584)) hashed up 15 levels = proof[15]
    (test_oracle_kat.py::test_reference_fixture_pins_wide_sponge_and_leaf_hash) -- the wide sponge, the width-6 leaf hash and the
    node hash of the CUDA path against the reference's own bytes, no oracle involved."""
    fx = json.load(open(os.path.join(GOLDEN, "user_config_proof.json")))
    pr = [base64.b64decode(x) for x in fx["Proof"]]
    ctx.set_poseidon_out_lane(1)
    z = bytes(32)
    empty_assets = ctx.poseidon_hash_batch(np.zeros(584 * 32, dtype=np.uint8), 584, 1).tobytes()
    v = ctx.poseidon_bytes(z, z, z, z, empty_assets)
    for _ in range(15):
        v = ctx.poseidon_bytes(v, v)
    assert v == pr[15]
    # and through the tree itself: a depth-28 tree whose nil leaf is that hash has the fixture's empty-subtree siblings
    nil_leaf = ctx.poseidon_bytes(z, z, z, z, empty_assets)
    t = zk.FixedDepthMerkleTree(ctx, 28, nil_leaf, 4)
    t.set_range(0, np.frombuffer(nil_leaf * 4, dtype=np.uint8), 4)
    t.build()
    proof = t.get_proofs(np.array([0], dtype=np.uint32)).reshape(28, 32)
    for lvl in range(15, 28):
        assert proof[lvl].tobytes() == pr[lvl]
    t.close()


def test_published_vectors_lane0(ctx):
    ctx.set_poseidon_out_lane(0)
    assert ctx.poseidon_bytes(b"\x01", b"\x02").hex() == "115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a"
    assert ctx.poseidon_bytes(b"\x01", b"\x02", b"\x03", b"\x04").hex() == "299c867db6c1fdd79dcefa40e4510b9837e60ebb1ce0663dbaa525df65250465"
    assert int.from_bytes(ctx.poseidon_bytes(b"\x01"), "big") == 18586133768512220936620570745912940619677854269274689475585506675881198879027
    ctx.set_poseidon_out_lane(1)


def test_golden_poseidon_all_widths(ctx):
    for lane in (0, 1):
        ctx.set_poseidon_out_lane(lane)
        for case in golden()["poseidon"]:
            ins = H(case["in"])
            got = ctx.poseidon_hash_batch(orc.be32_array(ins), len(ins), 1).tobytes()
            assert got.hex() == "%064x" % int(case[f"lane{lane}"], 16)
    ctx.set_poseidon_out_lane(1)


@pytest.mark.parametrize("n_in", [1, 2, 5, 11, 12, 13, 25, 100])
def test_hash_batch_vs_oracle(ctx, n_in):
    rng = SplitMix64(n_in)
    count = 37
    vals = [rng.field(R) for _ in range(count * n_in)]
    got = ctx.poseidon_hash_batch(orc.be32_array(vals), n_in, count)
    for i in range(count):
        want = orc.fr_unmont(orc.poseidon_hash(orc.fr_mont(vals[i * n_in:(i + 1) * n_in])))[0]
        assert got[i].tobytes() == want.to_bytes(32, "big")


def test_hasher_wrapper(ctx):
    h = zk.PoseidonHasher(ctx)
    h.write(b"\x00"); h.write(b"")
    assert h.sum(b"ab") == b"ab" + ps.poseidon_bytes([b"\x00", b""])
    assert h.data == []
    with pytest.raises(ValueError):
        h.write(R.to_bytes(32, "big"))


def test_golden_account_leaves(ctx):
    for lane in (0, 1):
        ctx.set_poseidon_out_lane(lane)
        for lv in golden()["leaves"]:
            flat = zk.padding_account_assets([tuple(a) for a in lv["assets"]])
            assert flat.tolist() == lv["flat"]
            ids = np.frombuffer(bytes.fromhex(lv["id"]), dtype=np.uint8).copy()
            tot = orc.be32_array([lv["equity"], lv["debt"], lv["collateral"]]).reshape(-1)
            got = ctx.account_leaves(ids, tot, flat, 1, lv["tier"])
            assert got.tobytes().hex() == lv[f"leaf_lane{lane}"]
    ctx.set_poseidon_out_lane(1)


@pytest.mark.parametrize("tier,n", [(50, 3001), (500, 203)])
def test_account_leaves_batch_vs_oracle(ctx, tier, n):
    rng = np.random.RandomState(tier)
    flat = rng.randint(0, 1 << 62, size=(n, tier * 6), dtype=np.int64).astype(np.uint64)
    ids = orc.be32_array([SplitMix64(i).field(R) for i in range(n)])
    tot = orc.be32_array([int(x) for x in rng.randint(0, 1 << 62, size=n * 3)]).reshape(n, 96)
    tot[5] = 0
    got = ctx.account_leaves(ids, tot, flat, n, tier)
    assert np.array_equal(got, orc.account_leaves(ids, tot, flat, tier))


def test_golden_merkle(ctx):
    for tv in golden()["merkle"]:
        ctx.set_poseidon_out_lane(tv["lane"])
        nil = bytes.fromhex(golden()["nil_account_hash"][f"lane{tv['lane']}"])
        t = zk.FixedDepthMerkleTree(ctx, tv["depth"], nil, tv["capacity"])
        assert t.root() == merkle.FixedDepthMerkleTree(tv["depth"], nil, tv["capacity"], tv["lane"]).root   # empty root = nilHashes[depth]
        for k, v in tv["leaves"].items():
            t.set(int(k), bytes.fromhex(v))
        t.build()
        assert t.root().hex() == tv["root"]
        assert [x.hex() for x in t.get_proof(9)] == tv["proof_9"]
        assert [x.hex() for x in t.get_proof(20)] == tv["proof_20"]
        leaf9 = bytes.fromhex(tv["leaves"]["9"])
        assert t.get(9) == leaf9 and t.get(20) == nil
        assert zk.verify_proof(ctx, t.root(), 9, t.get_proof(9), leaf9, tv["depth"])
        assert not zk.verify_proof(ctx, t.root(), 8, t.get_proof(9), leaf9, tv["depth"])
        t.close()
    ctx.set_poseidon_out_lane(1)


@pytest.mark.parametrize("capacity,nset,depth", [(1, 1, 28), (2, 2, 5), (1000, 1000, 28), (100003, 77777, 28), (4096, 4096, 12)])
def test_merkle_build_vs_oracle(ctx, capacity, nset, depth):
    rng = np.random.RandomState(capacity)
    leaves = rng.randint(0, 256, size=(capacity, 32)).astype(np.uint8); leaves[:, 0] &= 0x0F
    nil = merkle.nil_account_hash(1)
    t = zk.FixedDepthMerkleTree(ctx, depth, nil, capacity)
    t.set_range(0, leaves[:nset].copy(), nset)
    t.build()
    dirty = np.zeros((capacity + 63) // 64, dtype=np.uint64)
    for k in range(nset):
        dirty[k >> 6] |= np.uint64(1 << (k & 63))
    nodes, root = orc.merkle_build(leaves, capacity, depth, nil, dirty)
    assert t.root() == root
    keys = sorted(set([0, capacity - 1, nset - 1, min(nset, capacity - 1)] + [int(x) for x in rng.randint(0, capacity, size=50)]))
    assert np.array_equal(t.get_proofs(keys), orc.merkle_proofs(leaves, nodes, capacity, depth, nil, keys, dirty))
    t.close()


def test_merkle_sparse_sets_and_rebuild(ctx):
    """Set on arbitrary keys, Build, more Sets, Build again (merkletree_test.go multiple Set->Build cycles)."""
    capacity, depth = 5000, 20
    nil = merkle.nil_account_hash(1)
    rng = np.random.RandomState(3)
    leaves = np.zeros((capacity, 32), dtype=np.uint8)
    dirty = np.zeros((capacity + 63) // 64, dtype=np.uint64)
    t = zk.FixedDepthMerkleTree(ctx, depth, nil, capacity)
    for rnd in range(2):
        keys = np.unique(rng.randint(0, capacity, size=300)).astype(np.uint32)
        vals = rng.randint(0, 256, size=(keys.size, 32)).astype(np.uint8); vals[:, 0] &= 0x0F
        t.set_keys(keys, vals)
        for k, v in zip(keys, vals):
            leaves[k] = v; dirty[int(k) >> 6] |= np.uint64(1 << (int(k) & 63))
        t.build()
        _, root = orc.merkle_build(leaves, capacity, depth, nil, dirty)
        assert t.root() == root


def test_tree_argument_errors(ctx):
    nil = bytes(32)
    for args in ((33, 1), (0, 1), (3, 9)):
        with pytest.raises(ValueError):
            zk.FixedDepthMerkleTree(ctx, args[0], nil, args[1])
    t = zk.FixedDepthMerkleTree(ctx, 4, nil, 10)
    with pytest.raises(IndexError):
        t.set(10, nil)
    with pytest.raises(zk.ZkporError, match="out of range"):
        t.set_range(8, np.zeros((3, 32), dtype=np.uint8), 3)
    with pytest.raises(IndexError):
        t.get_proof(16)


def test_cex_assets_commitment_10000_elements(ctx):
    """utils.ComputeCexAssetsCommitment (src/utils/utils.go:779-800): 500 assets x 20 packed elements through the
    chained sponge -- 833 full-width permutations + one of width 5, twice per batch in witness.go:159-183."""
    rng = SplitMix64(500)
    assets = []
    for i in range(7):
        tiers = [(int(rng.next() % (1 << 100)), int(rng.next() % 101)) for _ in range(merkle.TIER_COUNT)]
        assets.append(dict(total_equity=rng.next(), total_debt=rng.next(), base_price=rng.next(), loan=rng.next(), margin=rng.next(), pm=rng.next(),
                           loan_ratios=tiers, margin_ratios=tiers[::-1], pm_ratios=tiers))
    empty = dict(total_equity=0, total_debt=0, base_price=0, loan=0, margin=0, pm=0, loan_ratios=[(0, 0)] * 12, margin_ratios=[(0, 0)] * 12, pm_ratios=[(0, 0)] * 12)
    elems = []
    for a in assets + [empty] * (merkle.ASSET_COUNTS - len(assets)):
        elems += merkle.cex_asset_packed(a)
    assert len(elems) == 10000
    for lane in (0, 1):
        ctx.set_poseidon_out_lane(lane); orc.poseidon_set_out_lane(lane)
        got = ctx.poseidon_hash_batch(orc.be32_array(elems), 10000, 1).tobytes()
        want = orc.fr_unmont(orc.poseidon_hash(orc.fr_mont(elems)))[0].to_bytes(32, "big")
        assert got == want
    ctx.set_poseidon_out_lane(1); orc.poseidon_set_out_lane(1)
    # batch commitment = PoseidonBytes(root, before, after, min, max) with the []byte{0} quirk for a zero index (witness.go:185-198)
    assert ctx.poseidon_bytes(got, got, want, b"\x00", (1379).to_bytes(2, "big")) == ps.poseidon_bytes([got, got, want, b"\x00", (1379).to_bytes(2, "big")], 1)


def _fast_cex_commitment(state, out_lane=None):
    """merkle.cex_assets_commitment with the 834 permutations done by the C oracle (the Python loop takes seconds per hash)"""
    empty = dict(total_equity=0, total_debt=0, base_price=0, loan=0, margin=0, pm=0, loan_ratios=[(0, 0)] * 12, margin_ratios=[(0, 0)] * 12, pm_ratios=[(0, 0)] * 12)
    elems = []
    for a in list(state) + [empty] * (merkle.ASSET_COUNTS - len(state)):
        elems += merkle.cex_asset_packed(a)
    return orc.fr_unmont(orc.poseidon_hash(orc.fr_mont(elems)))[0].to_bytes(32, "big")


def test_witness_batches_vs_oracle(ctx, monkeypatch):
    """zkpor_witness_batches against the reference's serial loop as restated in oracle/py/merkle.py (witness.go:144-206): running CEX
    totals, before / after commitments of every batch, batch commitments -- including a first batch that starts at account index 0."""
    rng = SplitMix64(77)
    n_assets, tier, ops, nb = merkle.ASSET_COUNTS, 50, 6, 4
    cex = []
    for i in range(9):
        tiers = [(int(rng.next() % (1 << 100)), int(rng.next() % 101)) for _ in range(merkle.TIER_COUNT)]
        cex.append(dict(total_equity=rng.next() >> 8, total_debt=rng.next() >> 8, base_price=rng.next(), loan=rng.next() >> 8, margin=rng.next() >> 8, pm=rng.next() >> 8,
                        loan_ratios=tiers, margin_ratios=tiers[::-1], pm_ratios=tiers))
    accounts = []
    for j in range(ops * nb):
        k = 1 + rng.next() % 5
        idxs = sorted({int(rng.next() % 9) for _ in range(k)})
        accounts.append((j, [(i, rng.next() >> 12, rng.next() >> 12, rng.next() >> 12, rng.next() >> 12, rng.next() >> 12) for i in idxs]))
    root = (0x1234567890ABCDEF << 64 | 12345).to_bytes(32, "big")
    monkeypatch.setattr(merkle, "cex_assets_commitment", _fast_cex_commitment)
    want, final = merkle.witness_batches(cex, root, accounts, ops)
    # inputs as the Go side holds them
    base_prices = np.zeros(n_assets, dtype=np.uint64); totals0 = np.zeros((n_assets, 5), dtype=np.uint64)
    tier_elems = np.zeros((n_assets, 18, 32), dtype=np.uint8)
    for i, a in enumerate(cex):
        base_prices[i] = a["base_price"]
        totals0[i] = [a["total_equity"], a["total_debt"], a["loan"], a["margin"], a["pm"]]
        packed = merkle.tier_ratios_packed(a["loan_ratios"]) + merkle.tier_ratios_packed(a["margin_ratios"]) + merkle.tier_ratios_packed(a["pm_ratios"])
        tier_elems[i] = orc.be32_array(packed)
    flat = np.array([merkle.padding_account_assets(assets) for _, assets in accounts], dtype=np.uint64)
    assert flat.shape == (ops * nb, tier * 6)
    idx = np.array([j for j, _ in accounts], dtype=np.uint32)
    totals, cm, bc = zk.witness_batches(ctx, base_prices=base_prices, tier_ratio_elems=tier_elems, initial_totals=totals0, root=root, flat_assets=flat,
                                        account_indices=idx, tier=tier, ops_per_batch=ops)
    for b, (before_totals, before, after, batch_cm) in enumerate(want):
        assert [tuple(int(x) for x in row) for row in totals[b][:9]] == before_totals[:9]
        assert cm[b].tobytes() == before and cm[b + 1].tobytes() == after and bc[b].tobytes() == batch_cm
    assert [tuple(int(x) for x in row) for row in totals[nb][:9]] == final[:9]
    # device-resident inputs, outputs not wanted: same commitments
    import torch
    tf = torch.from_numpy(flat.view(np.int64)).cuda()
    _, cm2, _ = zk.witness_batches(ctx, base_prices=base_prices, tier_ratio_elems=tier_elems, initial_totals=totals0, root=root, flat_assets=tf,
                                   account_indices=idx, tier=tier, ops_per_batch=ops)
    assert np.array_equal(cm, cm2)
    # an overflowing total and an asset index beyond the table fail loudly (utils.SafeAdd panics)
    big = totals0.copy(); big[0, 0] = (1 << 64) - 1
    bad_flat = flat.copy(); bad_flat[0, 0] = 0; bad_flat[0, 1] = 5
    with pytest.raises(zk.ZkporError, match="overflows"):
        zk.witness_batches(ctx, base_prices=base_prices, tier_ratio_elems=tier_elems, initial_totals=big, root=root, flat_assets=bad_flat, account_indices=idx, tier=tier, ops_per_batch=ops)
    bad_flat = flat.copy(); bad_flat[3, 0] = 501
    with pytest.raises(zk.ZkporError, match="asset index"):
        zk.witness_batches(ctx, base_prices=base_prices, tier_ratio_elems=tier_elems, initial_totals=totals0, root=root, flat_assets=bad_flat, account_indices=idx, tier=tier, ops_per_batch=ops)
    with pytest.raises(zk.ZkporError, match="whole batches"):
        zk.witness_batches(ctx, base_prices=base_prices, tier_ratio_elems=tier_elems, initial_totals=totals0, root=root, flat_assets=flat[:7], account_indices=idx[:7], tier=tier, ops_per_batch=ops)


@pytest.mark.parametrize("world,capacity", [(2, 1000), (4, 777), (8, 5000), (4, 3)])
def test_tree_built_across_ranks_matches_single(world, capacity):
    """account ranges -> one subtree per GPU -> all-gather of the subtree roots -> top levels on every rank.  The group
    is the in-process one (device ids may repeat), one host thread per rank; the root and every owner-served proof equal the single-GPU
    tree's (FixedDepthMerkleTree.Build / GetProof, merkletree.go:192-308)."""
    depth = 28
    rng = SplitMix64(900 + world + capacity)
    leaves = orc.be32_array([rng.field(R) for _ in range(capacity)])
    nil = merkle.nil_account_hash()
    single = zk.Context(0)
    t1 = zk.FixedDepthMerkleTree(single, depth, nil, capacity)
    t1.set_range(0, leaves, capacity); t1.build()
    _, want_root = orc.merkle_build(leaves, capacity, depth, nil)
    assert t1.root() == want_root
    have = zk.device_count()
    ctxs = zk.create_multi([i % have for i in range(world)])

    def rank_fn(r):
        t = zk.FixedDepthMerkleTree(ctxs[r], depth, nil, capacity)
        first, count, level = t.shard_range()
        if count:
            t.set_range(first, leaves[first:first + count], count)
        t.build_sharded()
        keys = np.arange(first, first + count, dtype=np.uint32)[:: max(1, count // 5)] if count else np.zeros(0, dtype=np.uint32)
        proofs = t.get_proofs(keys) if len(keys) else None
        return t.root(), keys, proofs, (first, count, level)

    res = zk.run_ranks(rank_fn, world)
    covered = 0
    for root, keys, proofs, (first, count, level) in res:
        assert root == want_root
        covered += count
        if len(keys):
            assert np.array_equal(proofs, t1.get_proofs(keys))
    assert covered == capacity
    for cx in ctxs:
        cx.close()
    single.close()
