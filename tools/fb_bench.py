import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, zkpor_b200 as zk
ctx = zk.Context(0)
for lg in (22, 24):
    n = 1 << lg
    sc = torch.empty(n * 4, dtype=torch.int64, device="cuda"); zk.synth_scalars(ctx, 3, n, 0, sc)
    g1 = torch.empty(8, dtype=torch.int64, device="cuda"); zk.synth_points_g1(ctx, 1, 1, 1, g1)
    g2 = torch.empty(16, dtype=torch.int64, device="cuda"); zk.synth_points_g2(ctx, 1, 1, 1, g2)
    o1 = torch.empty(n * 8, dtype=torch.int64, device="cuda"); o2 = torch.empty(n * 16, dtype=torch.int64, device="cuda")
    for name, f, g, o in (("g1", zk.g1_fixed_base_batch, g1, o1), ("g2", zk.g2_fixed_base_batch, g2, o2)):
        f(ctx, g, sc, n, o); torch.cuda.synchronize(); t0 = time.perf_counter(); f(ctx, g, sc, n, o); torch.cuda.synchronize()
        print("fixed-base %s 2^%d: %.1f ms" % (name, lg, (time.perf_counter() - t0) * 1e3), flush=True)
