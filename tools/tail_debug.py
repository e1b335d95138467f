"""debug: where does the deferred-tail proof differ (development tool)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "py")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc, zkpor_b200 as zk
from helpers import circuit_instance, make_pk
MEDIUM = dict(users=40, assets_per_user=2, cex_assets=5, tiers=3, merkle_depth=8, chain_perms=12, limb_bits=8)
inst = circuit_instance(seed=31, **MEDIUM)
flat = inst["flat"]
ctx = zk.Context(0)
os.environ["ZKPOR_TAIL_MIN"] = "0"; plain = zk.Program(ctx, flat)
os.environ["ZKPOR_TAIL_MIN"] = "1"; prog = zk.Program(ctx, flat)
print(prog.stats())
pk = make_pk(zk, ctx, inst)
p0 = pk.prove_solve(plain, inst["inputs_mont"], 0, 0)
p1 = pk.prove_solve(prog, inst["inputs_mont"], 0, 0)
for nm, a, b in (("Ar", 0, 64), ("Bs", 64, 192), ("Krs", 192, 256), ("C", 260, 324), ("Pok", 324, 388)):
    print(nm, "same" if p0[a:b] == p1[a:b] else "DIFF")
T = prog.tail_wires()
print("tail wires", len(T), T[:8], T[-4:], "n_wires", flat["n_wires"])
w, *_ = prog.solve(inst["inputs_mont"], pk)
ia = np.asarray(inst["sc"]["infinity_a"]).astype(bool)
A = inst["arr"]["A"].reshape(-1, 8)
keep = ~ia
rank = np.cumsum(keep) - 1
def msm(sel):
    idx = np.nonzero(sel & keep)[0]
    if len(idx) == 0: return None
    return orc.g1_unpack(ctx.msm_g1(np.ascontiguousarray(A[rank[idx]]), np.ascontiguousarray(w[idx]), len(idx)))[0]
allw = np.ones(len(w), dtype=bool); tail = np.zeros(len(w), dtype=bool); tail[T] = True
print("full   ", msm(allw)); print("nontail", msm(~tail)); print("tail   ", msm(tail))
print("in tail & in A:", int((tail & keep).sum()))
import bn254 as bn
alpha = orc.g1_unpack(inst["arr"]["alpha1"])[0]
for nm, p in (("plain", p0), ("deferred", p1)):
    ar = bn.g1_from_bytes(p[0:64])
    msm_part = bn.pt_add(ar, bn.pt_neg(alpha))
    print(nm, "Ar - alpha =", msm_part)
f, nt, tl = msm(allw), msm(~tail), msm(tail)
print("nontail + tail == full:", bn.pt_add(nt, tl) == f)
