"""development probe: G2 MSM with the two-lanes-per-bucket kernel (run with ZKPOR_G2_PAIR=2|3|4) -- parity at 2^14 against the oracle,
then the accumulate stage time at 2^22.  Meant to run under a SHORT timeout."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "py")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zkpor_b200 as zk, orc
from helpers import g2_points, rand_scalars_np
ctx = zk.Context(0)
n = 1 << 14
p2 = g2_points(256, 5); p2 = np.ascontiguousarray(np.tile(p2, (n // 256, 1)))
s2 = rand_scalars_np(n, 6)
got = ctx.msm_g2(p2, s2, n)
print("pair kernel parity 2^14:", bool(np.array_equal(got, orc.g2_msm(p2, s2))), flush=True)
nmax = 1 << 22
pts = torch.empty(nmax * 16, dtype=torch.int64, device="cuda"); sc = torch.empty(nmax * 4, dtype=torch.int64, device="cuda")
zk.synth_points_g2(ctx, 222, 333, nmax, pts); zk.synth_scalars(ctx, 9, nmax, 0, sc)
for _ in range(2):
    ctx.msm_g2(pts, sc, nmax); ctx.sync()
print("2^22", os.environ.get("ZKPOR_G2_PAIR"), ctx.last_timings(), flush=True)
