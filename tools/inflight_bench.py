"""Two provers on ONE GPU (development tool): each prover has its own context (streams, scratch), key and program; they prove
independent batches from two host threads.  While one proof is in its multiplications (throughput-bound kernels that fill the SMs) the
other is in its solve (a latency chain of ~8 000 small launches), so the pair should finish sooner than two proofs back to back.
Usage: python tools/inflight_bench.py [log_n] [provers] [steps]"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import zkpor_b200 as zk


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    provers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    r, s = bench.RS
    wls = []
    for p in range(provers):
        ctx = zk.Context(0)
        wls.append(bench.Workload(torch, zk, ctx, log_n))
        # the library holds its own copies of the key and the program: drop the generator's
        wls[-1].arrays = None
        for k, v in list(wls[-1].flat.items()):
            if hasattr(v, "device"):
                wls[-1].flat[k] = None
        torch.cuda.empty_cache()
        print("prover", p, "ready: setup %.1f s, HBM in use %.1f GB" % (wls[-1].setup_s, (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9), flush=True)
    ref = wls[0].pk.prove_solve(wls[0].prog, wls[0].inputs, r, s)
    for wl in wls:
        for _ in range(2):
            assert wl.pk.prove_solve(wl.prog, wl.inputs, r, s) == ref

    def one(wl, n, out):
        for _ in range(n):
            out.append(wl.pk.prove_solve(wl.prog, wl.inputs, r, s))

    res = {}
    for k in range(1, provers + 1):
        outs = [[] for _ in range(k)]
        th = [threading.Thread(target=one, args=(wls[i], steps, outs[i])) for i in range(k)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert all(p == ref for o in outs for p in o)
        res[k] = dict(provers=k, proofs=k * steps, seconds=dt, ms_per_proof=dt / (k * steps) * 1e3, proofs_per_hour=k * steps / dt * 3600)
        print(json.dumps(res[k]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(workload=wls[0].describe(), runs=list(res.values())), open(os.path.join(ROOT, "gpurun_out", f"inflight_bench_{log_n}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
