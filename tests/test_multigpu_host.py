"""N > 1 path on CPU (gloo, world_size 2): point-chunk sharding, the all-gather of the 1 KiB partials and
zkpor_groth16_finish (host arithmetic of the product).  The per-chunk partial sums are produced by the oracle here
(no GPU in this container); on the GPU box test_gpu_groth16.py runs the same combine with zkpor_groth16_prove_partial."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc
import zkpor_b200 as zk
from bn254 import R
from helpers import oracle_proof, synthetic_instance


def affine_to_xyzz(aff, g2=False):
    """affine point limbs -> XYZZ limbs (ZZ = ZZZ = 1; all-zero = infinity stays ZZ = 0)"""
    w = 16 if g2 else 8
    out = np.zeros(2 * w, dtype=np.uint64)
    if not np.asarray(aff).any():
        return out
    out[:w] = aff
    one = orc.fp_mont([1])[0]
    out[w:w + 4] = one
    out[w + w // 2: w + w // 2 + 4] = one
    return out


def _worker(rank, world, port, inst_seed, r, s, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "py"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inst = synthetic_instance(300, 20, inst_seed)
        arr, m = inst["arr"], orc.fr_mont
        h = orc.compute_h(m(inst["a"]), m(inst["b"]), m(inst["c"]), arr["log_n"])[:arr["Z"].shape[0]]
        def part(points, scalars, g2=False):
            lo, hi = zk.chunk_bounds(len(scalars), rank, world)
            f = orc.g2_msm if g2 else orc.g1_msm
            if hi == lo:
                return affine_to_xyzz(np.zeros(16 if g2 else 8, dtype=np.uint64), g2)
            return affine_to_xyzz(f(points[lo:hi].copy(), scalars[lo:hi].copy()), g2)
        wa, wb, wk, cm = m(inst["wa"]), m(inst["wb"]), m(inst["wk"]), m(inst["committed"])
        mine = zk.pack_partial(part(arr["A"], wa), part(arr["B1"], wb), part(arr["K"], wk), part(arr["Z"], h),
                               part(arr["ck_basis"], cm), part(arr["ck_basis_exp_sigma"], cm), part(arr["B2"], wb, True))
        gathered = [torch.zeros(zk.PROVE_PARTIAL_BYTES, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine))
        proof = zk.finish_proof(torch.stack(gathered).numpy(), arr["alpha1"], arr["beta1"], arr["delta1"], arr["beta2"], arr["delta2"], r, s)
        q.put((rank, proof == oracle_proof(inst, r, s)))
    finally:
        dist.destroy_process_group()


def test_sharded_prove_combine_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, 77, 123456789 % R, 987654321 % R, q)) for rk in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True), (1, True)]


def test_chunk_bounds_cover_exactly():
    for L in (0, 1, 7, 1000, (1 << 26) - 1):
        for world in (1, 2, 3, 8):
            cuts = [zk.chunk_bounds(L, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == L
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))


def _tree_worker(rank, world, port, capacity, depth, q):
    """one tree over `world` ranks as zkpor_tree_build_sharded does it: own leaf range -> subtree root (the oracle hashes here: no GPU in
    this container) -> all-gather of the subtree roots -> the top levels on every rank"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "py"))
    import merkle
    from bn254 import SplitMix64
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = SplitMix64(4242)
        leaves = orc.be32_array([rng.field(R) for _ in range(capacity)])
        nil = merkle.nil_account_hash()
        first, count, k = zk.tree_shard_range(rank, world, capacity, depth)
        # the subtree above my leaves is a depth-k tree of its own
        if count:
            _, sub_root = orc.merkle_build(leaves[first:first + count].copy(), count, k, nil)
        else:                                        # a rank without leaves contributes the empty subtree of level k
            sub_root = nil
            for _ in range(k):
                sub_root = orc.poseidon_node_batch(np.frombuffer(sub_root + sub_root, dtype=np.uint8).copy())[0].tobytes()
        gathered = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(np.frombuffer(sub_root, dtype=np.uint8).copy()))
        # top levels: a tree of depth - k over the `world` subtree roots, its empty leaves being the empty subtree of level k
        nil_k = nil
        for _ in range(k):
            nil_k = orc.poseidon_node_batch(np.frombuffer(nil_k + nil_k, dtype=np.uint8).copy())[0].tobytes()
        tops = torch.stack(gathered).numpy()
        have = min(world, -(-capacity // (1 << k)))
        _, root = orc.merkle_build(tops[:have].copy(), have, depth - k, nil_k)
        _, want = orc.merkle_build(leaves, capacity, depth, nil)
        q.put((rank, root == want, first, count))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("capacity", [1000, 3])
def test_sharded_tree_split_gloo_world2(capacity):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + capacity) % 2000
    procs = [ctx.Process(target=_tree_worker, args=(rk, 2, port, capacity, 28, q)) for rk in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert [r[1] for r in res] == [True, True]
    assert sum(r[3] for r in res) == capacity and res[0][2] == 0


def test_tree_shard_ranges_cover_exactly():
    for capacity in (1, 2, 3, 777, 1000, 10_000_000):
        for world in (1, 2, 4, 8):
            parts = [zk.tree_shard_range(r, world, capacity, 28) for r in range(world)]
            assert sum(p[1] for p in parts) == capacity
            assert all(p[0] == r << p[2] or p[1] == 0 for r, p in enumerate(parts))
            assert len({p[2] for p in parts}) == 1 and (world << parts[0][2]) >= capacity
