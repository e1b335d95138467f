"""development probe: one G1 (and optionally G2) MSM per size, meant to run under `ncu --metrics gpu__time_duration.sum`"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zkpor_b200 as zk

ctx = zk.Context(0)
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20, 24]
nmax = 1 << max(sizes)
pts = torch.empty(nmax * 8, dtype=torch.int64, device="cuda"); sc = torch.empty(nmax * 4, dtype=torch.int64, device="cuda")
zk.synth_points_g1(ctx, 12345, 67891, nmax, pts)
zk.synth_scalars(ctx, 7, nmax, 0, sc)
for lg in sizes:
    ctx.msm_g1(pts, sc, 1 << lg)
    ctx.sync()
    print("done", lg, ctx.last_timings(), flush=True)
if len(sys.argv) > 2:
    n2 = 1 << int(sys.argv[2])
    p2 = torch.empty(n2 * 16, dtype=torch.int64, device="cuda")
    zk.synth_points_g2(ctx, 222, 333, n2, p2)
    ctx.msm_g2(p2, sc, n2); ctx.sync()
    print("done g2", ctx.last_timings(), flush=True)
