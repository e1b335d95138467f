"""CPU models of the index arithmetic the CUDA kernels rely on (no GPU, no oracle): the invariants are proved here on
integers so that a kernel bug cannot hide behind a wrong derivation.

* bucket reduction tree (csrc/msm.cu k_reduce_level): W = sum_j (j+1) X_j from segment running sums, levels of 8
* batched-affine round layout (csrc/msm_affine.cu): off_r = (off_{r-1} + b) >> 1 keeps the halved lists disjoint
* two-level partition of the counting sort (csrc/msm_sort.cu): (bucket-1) >> shift covers at most 64 partitions per window
"""
import random


def tree_reduce(x, log_l=3):
    L = 1 << log_l
    X, A, K = list(x), None, 0
    while len(X) > 1:
        m = len(X); mo = (m + L - 1) // L
        R, An = [0] * mo, [0] * mo
        for g in range(mo):
            lo, hi = g * L, min(g * L + L, m)
            run = s = 0
            for j in range(hi - 1, lo - 1, -1):
                run += X[j]; s += run
            s <<= K * log_l
            if A is not None:
                s += sum(A[lo:hi])
            R[g], An[g] = run, s
        X, A, K = R, An, K + 1
    if K == 0:
        return X[0]
    return A[0] - sum(1 << (l * log_l) for l in range(1, K)) * X[0]


def test_bucket_reduction_tree_formula():
    rng = random.Random(1)
    for n in [1, 2, 7, 8, 9, 63, 64, 65, 128, 513, 4096, 5000, 1 << 14]:
        x = [rng.randrange(1 << 40) for _ in range(n)]
        assert tree_reduce(x) == sum((j + 1) * v for j, v in enumerate(x))


def test_affine_round_layout_is_disjoint():
    rng = random.Random(3)
    for _ in range(200):
        nb = rng.choice([1, 2, 3, 8, 64, 257])
        cnt = [rng.choice([0, 0, 1, 2, 3, 5, 8, 100, rng.randrange(300)]) for _ in range(nb)]
        off = [sum(cnt[:b]) for b in range(nb)]
        cap, o, m = sum(cnt), off[:], cnt[:]
        for r in range(rng.randrange(1, 9)):
            cap2 = (cap + nb) // 2 + 1
            o2 = [(o[b] + b) >> 1 for b in range(nb)]
            m2 = [(c + 1) // 2 for c in m]
            used = set()
            for b in range(nb):
                assert m2[b] == (cnt[b] + (1 << (r + 1)) - 1) >> (r + 1)       # closed form used by the kernels
                for k in range(m2[b]):
                    pos = o2[b] + k
                    assert pos < cap2 and pos not in used
                    used.add(pos)
            o, m, cap = o2, m2, cap2


def test_partition_plan_covers_every_digit():
    PART_BITS = 6
    for c in range(8, 21):
        nwin = (255 + c - 1) // c
        nb = 1 << (c - 1)
        shift_norm = c - 1 - PART_BITS
        top_raw = max(0, 256 - c * (nwin - 1))
        top_bits = min(top_raw, c - 1)
        shift_top = max(0, top_bits - PART_BITS)
        assert shift_norm >= 0 or c < PART_BITS + 2
        if c >= PART_BITS + 2:
            assert ((nb - 1) >> shift_norm) < (1 << PART_BITS) and (1 << shift_norm) <= 8192
        # top window: any 256-bit integer gives a digit (+ carry) <= 2^top_raw, never negative (2^top_raw <= nb needs top_raw <= c-1)
        max_top_digit = min(nb, (1 << top_raw))
        assert ((max_top_digit - 1) >> shift_top) < (1 << PART_BITS)
        assert (1 << shift_top) <= 8192


def test_witness_batches_restatement_is_a_prefix_sum(monkeypatch):
    """oracle/py/merkle.py witness_batches restates the reference's serial loop (witness.go:144-206).  What the GPU pipeline relies on:
    the after-state of a batch is the before-state of the next (so n + 1 commitments serve n batches), and the totals are the initial
    ones plus the per-batch sums."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "py"))
    import merkle
    import orc
    from bn254 import SplitMix64

    def fast(state, out_lane=None):      # the 834-permutation sponge by the C oracle
        empty = dict(total_equity=0, total_debt=0, base_price=0, loan=0, margin=0, pm=0, loan_ratios=[(0, 0)] * 12, margin_ratios=[(0, 0)] * 12, pm_ratios=[(0, 0)] * 12)
        elems = []
        for a in list(state) + [empty] * (merkle.ASSET_COUNTS - len(state)):
            elems += merkle.cex_asset_packed(a)
        return orc.fr_unmont(orc.poseidon_hash(orc.fr_mont(elems)))[0].to_bytes(32, "big")

    assert fast([]) == merkle.cex_assets_commitment([])          # the C sponge is the Python one
    monkeypatch.setattr(merkle, "cex_assets_commitment", fast)
    rng = SplitMix64(5)
    cex = [dict(total_equity=10 * i, total_debt=i, base_price=7 + i, loan=0, margin=3, pm=1, loan_ratios=[(5, 50)] * 12, margin_ratios=[(6, 60)] * 12, pm_ratios=[(7, 70)] * 12)
           for i in range(4)]
    accounts = [(j, [(int(rng.next() % 4), 1 + j, 2, 3, 4, 5)]) for j in range(6)]
    out, final = merkle.witness_batches(cex, bytes(32), accounts, 2)
    assert len(out) == 3
    for b in range(2):
        assert out[b][2] == out[b + 1][1]                         # after(b) == before(b+1)
    assert sum(t[0] for t in final) == sum(a["total_equity"] for a in cex) + sum(1 + j for j in range(6))
    assert out[0][0][:4] == [(a["total_equity"], a["total_debt"], a["loan"], a["margin"], a["pm"]) for a in cex]
