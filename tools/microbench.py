"""Per-kernel microbenchmarks on one B200 (development tool; bench.py is the contract benchmark).
Usage: python tools/microbench.py [msm] [g2] [ntt] [hash] [check]  -- default: all but check."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import zkpor_b200 as zk

R = zk.R_MOD


def dev_buf(nbytes):
    return torch.empty(nbytes // 8, dtype=torch.int64, device="cuda")


def timed(ctx, fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


def main():
    what = set(sys.argv[1:]) or {"msm", "g2", "ntt", "hash"}
    ctx = zk.Context(0)
    out = {}
    if "msm" in what:
        nmax = 1 << 26
        pts = dev_buf(nmax * 64); sc = dev_buf(nmax * 32)
        t0 = time.perf_counter(); zk.synth_points_g1(ctx, 12345, 67891, nmax, pts); print("synth g1 2^26: %.2fs" % (time.perf_counter() - t0), flush=True)
        for kind in (0, 2):
            zk.synth_scalars(ctx, 7 + kind, nmax, kind, sc)
            for lg in (20, 22, 24, 26):
                n = 1 << lg
                ms = timed(ctx, lambda: ctx.msm_g1(pts, sc, n))
                st = ctx.last_timings()
                print(f"msm_g1 2^{lg} kind={kind}: {ms:.2f} ms  " + " ".join(f"{k}={v:.2f}" for k, v in st.items() if v > 0), flush=True)
                out[f"msm_g1_{lg}_{kind}"] = dict(ms=ms, **st)
        if "check" in what:
            sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
            import orc, bn254 as bn
            zk.synth_scalars(ctx, 7, nmax, 0, sc)
            got = ctx.msm_g1(pts, sc, nmax)
            h = sc.cpu().numpy().view(np.uint64).reshape(-1, 4)
            s, t = orc.fr_index_sums(h)
            rinv = pow(1 << 256, -1, R)
            dot = (12345 * s + 67891 * t) * rinv % R
            print("2^26 exact check:", orc.g1_unpack(got)[0] == bn.pt_mul(bn.G1_GEN, dot), flush=True)
        del pts, sc
    if "g2" in what:
        nmax = 1 << 24
        pts = dev_buf(nmax * 128); sc = dev_buf(nmax * 32)
        t0 = time.perf_counter(); zk.synth_points_g2(ctx, 222, 333, nmax, pts); print("synth g2 2^24: %.2fs" % (time.perf_counter() - t0), flush=True)
        zk.synth_scalars(ctx, 9, nmax, 0, sc)
        for lg in (20, 22, 24):
            n = 1 << lg
            ms = timed(ctx, lambda: ctx.msm_g2(pts, sc, n), reps=2)
            st = ctx.last_timings()
            print(f"msm_g2 2^{lg}: {ms:.2f} ms  " + " ".join(f"{k}={v:.2f}" for k, v in st.items() if v > 0), flush=True)
            out[f"msm_g2_{lg}"] = dict(ms=ms, **st)
        del pts, sc
    if "ntt" in what:
        for lg in (22, 24, 26):
            n = 1 << lg
            a, b, c, h = (dev_buf(n * 32) for _ in range(4))
            for x, sd in ((a, 1), (b, 2), (c, 3)):
                zk.synth_scalars(ctx, sd, n, 0, x)
            ms1 = timed(ctx, lambda: ctx.ntt(a, lg, False, False, False))
            ms = timed(ctx, lambda: ctx.compute_h(a, b, c, n, lg, out=h), reps=2)
            print(f"ntt 2^{lg}: {ms1:.2f} ms   compute_h 2^{lg}: {ms:.2f} ms  ntt-stage={ctx.last_timings()['ntt']:.2f}", flush=True)
            out[f"ntt_{lg}"] = dict(ntt_ms=ms1, compute_h_ms=ms)
            del a, b, c, h
    if "hash" in what:
        n = 1 << 23
        leaves = dev_buf(n * 32)
        zk.synth_scalars(ctx, 5, n, 0, leaves)     # uniform < r: valid canonical big-endian elements as bytes? top byte is LE limb 0 -> mask
        lv = leaves.view(torch.uint8).view(-1, 32); lv[:, 0] &= 0x0F
        t = zk.FixedDepthMerkleTree(ctx, 28, bytes(32), n)
        t.set_range(0, leaves, n)
        ms = timed(ctx, lambda: t.build(), reps=2)
        print(f"merkle build 2^23 leaves depth 28: {ms:.2f} ms  ({n / ms / 1e3:.1f} M node-hashes/s)", flush=True)
        out["merkle_23"] = dict(ms=ms)
        t.close()
        na = 1 << 18
        ids = dev_buf(na * 32); tot = dev_buf(na * 96); flat = dev_buf(na * 50 * 48); o = dev_buf(na * 32)
        for x in (ids, tot):
            zk.synth_scalars(ctx, 6, x.numel() // 4, 0, x); x.view(torch.uint8).view(-1, 32)[:, 0] &= 0x0F
        flat.random_(0, 1 << 62)
        ms = timed(ctx, lambda: ctx.account_leaves(ids, tot, flat, na, 50, out=o), reps=2)
        print(f"account leaves tier 50, 2^18 accounts: {ms:.2f} ms ({na / ms:.1f} accounts/ms)", flush=True)
        out["leaves50_18"] = dict(ms=ms)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
