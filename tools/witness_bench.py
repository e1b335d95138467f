"""Witness-service hot path at scale (BASELINE config 5 shape; `python bench.py --workload witness [--gpus N]` runs this).

N synthetic accounts (95 % in the 50-asset tier, 5 % in the 500-asset tier) go through everything
src/witness does between the CSV parser and the DB writer:
  leaf hashes (utils.AccountInfoToHash, src/witness/main.go:163-195)            zkpor_account_leaves
  FixedDepthMerkleTree.Build at depth 28 (main.go:197)                          zkpor_tree_build / zkpor_tree_build_sharded
  the batch loop (witness/witness.go:144-206): running CEX totals, the two 10 000-element commitments and the batch commitment of
  every batch of 1380 (tier 50) / 200 (tier 500) accounts                       zkpor_witness_batches
  the account proofs of the batches (witness.go:323)                            zkpor_tree_get_proofs (timed on a sample of batches)
N > 1 (torchrun): every rank hashes and owns one account range and its subtree; one all-gather of the subtree roots (NCCL, 32 B per
rank) gives every rank the root; the batch loop is replicated work of seconds and runs on rank 0.
One JSON line: accounts/s end to end, the stage times, and the Poseidon kernels' product rate against the measured integer-pipe ceiling
(a width-13 permutation in the sparse form is ~3.5 K field products; leaf hashing is ALU-bound by two orders of magnitude)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

OPS = {50: 1380, 500: 200}                         # utils.BatchCreateUserOpsCountsTiers
PERM13_PRODUCTS = 3500.0                           # field products of one width-13 permutation, sparse partial rounds
NODE_PRODUCTS = 830.0                              # one width-3 permutation
IMAD_PEAK_GPS = 27.1 * 148 * 1.965 / 128


def main(args=None):
    import torch
    import zkpor_b200 as zk
    n_total = int(getattr(args, "accounts", 0) or (int(sys.argv[1]) if len(sys.argv) > 1 and args is None else 10_000_000))
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = zk.Context(local)
    if world > 1:
        uid = [zk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    depth = 28
    # account population: tier-50 accounts first (whole batches), then tier-500 (the service sorts users by tier, main.go:120-160)
    n500 = (n_total // 20) // OPS[500] * OPS[500]
    n50 = (n_total - n500) // OPS[50] * OPS[50]
    n = n50 + n500
    tree = zk.FixedDepthMerkleTree(ctx, depth, bytes(32), n)
    first, count, level = tree.shard_range() if world > 1 else (0, n, depth)
    lo, hi = first, first + count

    def synth_flat(m, tier, seed):
        """PaddingAccountAssets rows: ascending asset indices below 500, values below 2^40 (sums stay far from 2^64)"""
        g = torch.Generator(device="cuda"); g.manual_seed(seed)
        f = torch.randint(0, 1 << 40, (m, tier, 6), dtype=torch.int64, device="cuda", generator=g)
        step = max(1, 500 // tier)
        f[:, :, 0] = (torch.arange(tier, device="cuda", dtype=torch.int64) * step)[None, :]
        torch.cuda.synchronize()          # the library launches on its own non-blocking stream: torch's kernels must have finished
        return f.reshape(m, tier * 6)

    chunk = 1 << 19
    ids = torch.empty(chunk * 4, dtype=torch.int64, device="cuda"); tot = torch.empty(chunk * 12, dtype=torch.int64, device="cuda")
    for x in (ids, tot):
        zk.synth_scalars(ctx, 6, x.numel() // 4, 0, x); x.view(torch.uint8).view(-1, 32)[:, 0] &= 0x0F
    leaves = torch.empty(max(count, 1) * 4, dtype=torch.int64, device="cuda")
    flat50_all = synth_flat(min(n50, max(0, min(hi, n50) - lo)) if world > 1 else n50, 50, 50 + rank) if n50 else None
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    # ---- leaf hashes of this rank's account range
    done = 0
    t50 = t500 = 0.0
    a50 = max(0, min(hi, n50) - lo)                       # tier-50 accounts of this rank
    tt = time.perf_counter()
    for s in range(0, a50, chunk):
        m = min(chunk, a50 - s)
        ctx.account_leaves(ids, tot, flat50_all[s:s + m], m, 50, out=leaves[done * 4:(done + m) * 4]); done += m
    torch.cuda.synchronize(); t50 = time.perf_counter() - tt
    a500 = count - a50
    tt = time.perf_counter()
    c500 = 1 << 16
    flat500 = synth_flat(min(c500, max(a500, 1)), 500, 500 + rank)
    for s in range(0, a500, c500):
        m = min(c500, a500 - s)
        ctx.account_leaves(ids, tot, flat500[:m], m, 500, out=leaves[done * 4:(done + m) * 4]); done += m
    torch.cuda.synchronize(); t500 = time.perf_counter() - tt
    # ---- tree
    tt = time.perf_counter()
    if count:
        tree.set_range(lo, leaves, count)
    tree.build_sharded() if world > 1 else tree.build()
    torch.cuda.synchronize(); t_build = time.perf_counter() - tt
    root = tree.root()
    # ---- batch loop (rank 0; the other ranks' accounts would arrive as their flat rows -- here rank 0 regenerates what it needs)
    t_batches, nb50, nb500, cm_last = 0.0, n50 // OPS[50], n500 // OPS[500], None
    if rank == 0:
        base_prices = np.arange(1, 501, dtype=np.uint64); tier_elems = np.zeros((500, 18, 32), dtype=np.uint8); tier_elems[:, :, 31] = 7
        totals = np.zeros((500, 5), dtype=np.uint64)
        tt = time.perf_counter()
        if n50:
            f50 = flat50_all if world == 1 else synth_flat(n50, 50, 50)
            tot_b, cm, bc = zk.witness_batches(ctx, base_prices=base_prices, tier_ratio_elems=tier_elems, initial_totals=totals, root=root,
                                               flat_assets=f50, account_indices=np.arange(n50, dtype=np.uint32), tier=50, ops_per_batch=OPS[50])
            totals = tot_b[-1].copy(); cm_last = cm[-1].tobytes()
            del f50
        if n500:
            # the 500-asset tier: 2.4 MB of assets per batch of 200; rows generated per call (the same rows for every chunk of batches)
            per = max(1, c500 // OPS[500]) * OPS[500]
            f500 = synth_flat(per, 500, 500)
            for s in range(0, n500, per):
                m = min(per, n500 - s)
                tot_b, cm, bc = zk.witness_batches(ctx, base_prices=base_prices, tier_ratio_elems=tier_elems, initial_totals=totals, root=root,
                                                   flat_assets=f500[:m], account_indices=np.arange(n50 + s, n50 + s + m, dtype=np.uint32), tier=500, ops_per_batch=OPS[500])
                totals = tot_b[-1].copy(); cm_last = cm[-1].tobytes()
        torch.cuda.synchronize(); t_batches = time.perf_counter() - tt
    # ---- account proofs: a sample of batches of this rank's range, scaled to all of them
    tt = time.perf_counter()
    sample_batches = 8
    got = 0
    for b in range(sample_batches):
        k0 = lo + (b * 9973 * OPS[50]) % max(1, count - OPS[50]) if count > OPS[50] else lo
        keys = np.arange(k0, min(k0 + OPS[50], hi), dtype=np.uint32)
        if len(keys):
            tree.get_proofs(keys); got += len(keys)
    torch.cuda.synchronize(); t_proofs_sample = time.perf_counter() - tt
    t_proofs = t_proofs_sample * (count / max(1, got))
    e1.record(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0 - t_proofs_sample + t_proofs
    if dist is not None:
        t = torch.tensor([wall, t50, t500, t_build, t_proofs], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, t50, t500, t_build, t_proofs = (float(x) for x in t.tolist())
        roots = [None] * world; dist.all_gather_object(roots, root.hex())
        assert len(set(roots)) == 1, "ranks disagree on the root"
    if rank == 0:
        leaf_products = (a50 * (9 * PERM13_PRODUCTS + 1200) + a500 * (84 * PERM13_PRODUCTS + 1200))      # per rank; +1 width-6 leaf hash
        line = {"metric": "accounts/s", "value": n / wall, "unit": "accounts/s", "n_gpus": world, "steps": 1, "warmup": 0, "ms_per_step": wall * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic",
                "config": {"workload": f"witness service hot path: {n} accounts ({n50} tier-50 in {nb50} batches of 1380, {n500} tier-500 in {nb500} batches of 200), depth-{depth} tree",
                           "parallelism": "single GPU" if world == 1 else f"account ranges and subtrees over {world} GPUs, one all-gather of {world} x 32 B (NCCL), subtree level {level}"},
                "stage_s": {"leaf_hash_tier50": t50, "leaf_hash_tier500": t500, "tree_build": t_build, "batch_loop_all_batches": t_batches,
                            "account_proofs_all_accounts_scaled_from_sample": t_proofs},
                "batches": nb50 + nb500, "root": root.hex()[:16], "last_cex_commitment": cm_last.hex()[:16] if cm_last else None,
                "modmul_roofline": {"kernel": "k_account_leaves_tpa (leaf hashing)", "achieved_gps": leaf_products / max(t50 + t500, 1e-9) / 1e9, "peak_gps": IMAD_PEAK_GPS,
                                    "frac": leaf_products / max(t50 + t500, 1e-9) / 1e9 / IMAD_PEAK_GPS, "unit": "1e9 field products/s",
                                    "note": "products counted as 3.5 K per width-13 permutation (sparse partial rounds); per rank"},
                "roofline": {"bound": "hbm", "kernel": "k_account_leaves_tpa", "achieved": (a50 * 2432 + a500 * 24032) / max(t50 + t500, 1e-9) / 1e9, "unit": "GB/s",
                             "note": "2 432 B per tier-50 leaf, 24 032 B per tier-500 leaf; ALU-bound, see modmul_roofline"}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
