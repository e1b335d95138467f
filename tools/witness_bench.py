"""Witness-service hot path at scale (BASELINE config 5 shape): N synthetic accounts -> leaf hashes (utils.AccountInfoToHash)
-> FixedDepthMerkleTree.Build (depth 28) -> GetProof for one batch of 1380 users.  Development tool; prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import zkpor_b200 as zk


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    ctx = zk.Context(0)
    chunk = 1 << 20
    leaves = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    ids = torch.empty(chunk * 4, dtype=torch.int64, device="cuda"); tot = torch.empty(chunk * 12, dtype=torch.int64, device="cuda")
    flat = torch.empty(chunk * 300, dtype=torch.int64, device="cuda")
    for x in (ids, tot):
        zk.synth_scalars(ctx, 6, x.numel() // 4, 0, x); x.view(torch.uint8).view(-1, 32)[:, 0] &= 0x0F
    flat.random_(0, 1 << 62)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for lo in range(0, n, chunk):                      # tier-50 accounts (95% of the reference's population)
        m = min(chunk, n - lo)
        ctx.account_leaves(ids, tot, flat, m, 50, out=leaves[lo * 4:(lo + m) * 4])
    t_leaves = time.perf_counter() - t0
    tree = zk.FixedDepthMerkleTree(ctx, 28, bytes(32), n)
    t0 = time.perf_counter(); tree.set_range(0, leaves, n); tree.build(); t_build = time.perf_counter() - t0
    keys = np.arange(1380, dtype=np.uint32) + 4242
    t0 = time.perf_counter(); pr = tree.get_proofs(keys); t_proofs = time.perf_counter() - t0
    print(json.dumps({"accounts": n, "leaf_hash_s": t_leaves, "accounts_per_s": n / t_leaves, "tree_build_s": t_build,
                      "proofs_1380_ms": t_proofs * 1e3, "root": tree.root().hex()[:16]}))


if __name__ == "__main__":
    main()
