"""One proof across N ranks (include/zkpor_b200.h "one proof across the N GPUs of a box") through the C-ABI.

The in-process group (zkpor_ctx_create_multi) accepts repeated device ids, so the N-rank algorithm -- key split by point chunk,
four-step computeH with all-to-all exchanges, commitment combined mid-solve, one all-gather of the partial sums -- runs on the
single B200 of the test box with one host thread per rank; on a box with several GPUs the same tests spread over them.
Bar: proof bytes identical to the single-GPU proof and to the oracle's (reference call: src/prover/prover/prover.go:269)."""
import numpy as np
import pytest

import orc
import zkpor_b200 as zk
from bn254 import R, SplitMix64
from helpers import circuit_instance, make_pk

pytestmark = pytest.mark.gpu

MEDIUM = dict(users=40, assets_per_user=2, cex_assets=5, tiers=3, merkle_depth=6, chain_perms=4, limb_bits=8)


def devices(n):
    have = zk.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("world,log_n", [(2, 10), (4, 12), (8, 14), (2, 16)])
def test_compute_h_sharded_matches_single(world, log_n):
    n = 1 << log_n
    rng = SplitMix64(100 + world + log_n)
    m = n - 37                                                        # fewer constraints than the domain: zero padding is part of the contract
    a = orc.fr_mont([rng.field(R) for _ in range(m)]); b = orc.fr_mont([rng.field(R) for _ in range(m)])
    c = np.zeros_like(a)
    orc.lib().orc_fr_mul_batch(a.ctypes.data_as(orc.C.c_void_p), b.ctypes.data_as(orc.C.c_void_p), c.ctypes.data_as(orc.C.c_void_p), orc.C.c_size_t(m))
    single = zk.Context(0)
    want = single.compute_h(a, b, c, m, log_n)
    assert np.array_equal(want, orc.compute_h(a, b, c, log_n))
    single.close()
    pad = lambda v: np.concatenate([v, np.zeros((n - m, 4), dtype=np.uint64)])
    ap, bp, cp = pad(a), pad(b), pad(c)
    ctxs = zk.create_multi(devices(world))

    def rank_fn(r):
        return ctxs[r].compute_h_sharded(np.ascontiguousarray(ap[r::world]), np.ascontiguousarray(bp[r::world]), np.ascontiguousarray(cp[r::world]), log_n)

    for _ in range(2):                                                # twice: the exchange buffers and events are reusable
        chunks = zk.run_ranks(rank_fn, world)
        assert np.array_equal(np.concatenate(chunks), want)
    info = ctxs[1].comm_info()
    assert info["world"] == world and info["rank"] == 1 and info["all_to_all_calls"] == 2 * 7
    for cx in ctxs:
        cx.close()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_prove_solve_is_bit_exact(world):
    inst = circuit_instance(seed=31, **MEDIUM)
    flat = inst["flat"]
    r, s = 0x1234567 % R, 0x7654321 % R
    want, _ = orc.groth16_prove_program(inst["arr"], flat, inst["sc"]["infinity_a"], inst["sc"]["infinity_b"], inst["inputs_mont"], r, s)
    ctxs = zk.create_multi(devices(world))
    progs = [zk.Program(cx, flat) for cx in ctxs]
    import helpers
    orig = zk.ProvingKey

    def shard_pk(cx):
        # make_pk builds the whole-key description; shard=True uploads this rank's chunks only
        zk.ProvingKey = lambda *a, **k: orig(*a, shard=True, **k)
        try:
            return helpers.make_pk(zk, cx, inst)
        finally:
            zk.ProvingKey = orig

    pks = [shard_pk(cx) for cx in ctxs]
    infos = [pk.shard_info() for pk in pks]
    assert [i["rank"] for i in infos] == list(range(world)) and all(i["world"] == world for i in infos)
    assert sum(i["n_wires"] for i in infos) == flat["n_wires"] and sum(i["n_z"] for i in infos) == (1 << inst["arr"]["log_n"]) - 1
    assert sum(i["n_a"] for i in infos) == len(inst["sc"]["A_s"]) and sum(i["n_k"] for i in infos) == len(inst["sc"]["K_s"])
    # every rank from its own thread, every rank gets the proof
    proofs = zk.run_ranks(lambda k: pks[k].prove_solve(progs[k], inst["inputs_mont"], r, s), world)
    assert all(p == want for p in proofs)
    # the library's own thread fan-out
    assert zk.multi_prove_solve(ctxs, pks, progs, inst["inputs_mont"], r, s) == want
    # a failing witness fails on every rank (no rank is left waiting) ...
    flat_l = flat["secret_layout"]
    first, n_s, count, specs = [x for x in flat_l if any(k == "uint" for k, _ in x[3])][0]
    j = [k for k, _ in specs].index("uint")
    bad = list(inst["inputs"]); bad[first - 1 + j] = 1 << 70
    with pytest.raises(zk.ZkporError):
        zk.multi_prove_solve(ctxs, pks, progs, orc.fr_mont(bad), r, s)
    for p in pks + progs:
        p.close()
    for cx in ctxs:
        cx.close()


def test_sharded_key_on_wrong_context_is_rejected():
    inst = circuit_instance(seed=5, users=3, assets_per_user=2, cex_assets=3, tiers=2, merkle_depth=2, chain_perms=3, limb_bits=8)
    ctx = zk.Context(0)
    pk = make_pk(zk, ctx, inst)                                        # whole key, plain context
    prog = zk.Program(ctx, inst["flat"])
    assert pk.shard_info()["world"] == 1
    ctxs = zk.create_multi(devices(2))
    prog2 = zk.Program(ctxs[0], inst["flat"])
    pk2 = make_pk(zk, ctxs[0], inst)
    with pytest.raises(zk.ZkporError, match="shard"):
        # a whole key used on a rank of a 2-group: refused before any collective starts, so nothing hangs
        pk2.prove_solve(prog2, inst["inputs_mont"], 1, 2)
    pk2.close(); prog2.close()
    for cx in ctxs:
        cx.close()
    pk.close(); prog.close(); ctx.close()
