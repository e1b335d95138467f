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
