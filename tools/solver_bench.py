"""Witness-solver timing on one B200 (development tool): builds the bench circuit at 2^LOG_N and times zkpor_r1cs_solve alone for a
few schedule settings (env knobs of csrc/solver.cu), with the per-class launch statistics of the library.
Usage: python tools/solver_bench.py [log_n] [settings...]   setting = NARROW_MAX:NARROW_THREADS:WIDE_LONG_ROW (0 = a warp for every row)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import zkpor_b200 as zk


def summarise_trace(path):
    """per-step device times of the last solve (csrc/solver.cu, ZKPOR_SOLVE_TRACE) by step type and width"""
    import csv
    rows = list(csv.DictReader(open(path)))
    kinds = {0: "wide", 1: "narrow", 2: "count", 3: "commit"}
    agg = {}
    for r in rows:
        k, cnt, ms = kinds[int(r["kind"])], int(r["count"]), float(r["ms"])
        if k == "wide":
            b = 1 << max(0, cnt.bit_length() - 1)
            key = (k, b, int(r["has_div"]), int(int(r["n_long"]) > 0))
        else:
            key = (k, 0, 0, 0)
        a = agg.setdefault(key, [0, 0.0, 0])
        a[0] += 1; a[1] += ms; a[2] += cnt
    print("kind  width>=  div long  steps     ms   us/step  instr")
    for key in sorted(agg):
        n, ms, cnt = agg[key]
        print("%-6s %8d %3d %3d %7d %8.2f %8.2f %10d" % (key + (n, ms, 1e3 * ms / n, cnt)))
    print("total ms", sum(a[1] for a in agg.values()), flush=True)


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    settings = sys.argv[2:] or ["96:512:9", "96:512:0", "96:512:17", "96:256:9", "24:512:9"]
    ctx = zk.Context(0)
    wl = bench.Workload(torch, zk, ctx, log_n)
    print(wl.describe(), "setup %.1f s" % wl.setup_s, flush=True)
    wires = bench.dev_buf(torch, wl.sh["W"] * 32)
    out = []
    for st in settings:
        nm, nt, g32 = st.split(":")
        os.environ["ZKPOR_NARROW_MAX"], os.environ["ZKPOR_NARROW_THREADS"], os.environ["ZKPOR_WIDE_LONG_ROW"] = nm, nt, g32
        t0 = time.perf_counter()
        prog = zk.Program(ctx, wl.flat)
        up = time.perf_counter() - t0
        for _ in range(2):
            prog.solve_device(wl.inputs, wires, wl.pk)
        ctx.kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.ExternalStream(ctx.stream())
        torch.cuda.synchronize()
        e0.record(stream)
        n = 3
        for _ in range(n):
            prog.solve_device(wl.inputs, wires, wl.pk)
        e1.record(stream)
        torch.cuda.synchronize()
        wide, narrow = ctx.kernel_stats(5), ctx.kernel_stats(6)
        ctx.kernel_timing(False)
        rec = dict(setting=st, upload_s=up, solve_ms=e0.elapsed_time(e1) / n, stats=prog.stats(), stage=ctx.last_timings().get("solve"),
                   wide_ms=wide["total_ms"] / n, wide_launches=wide["launches"] // n, narrow_ms=narrow["total_ms"] / n, narrow_levels=narrow["units"] // n)
        rec["us_per_wide_level"] = 1e3 * rec["wide_ms"] / max(1, rec["wide_launches"])
        rec["us_per_narrow_level"] = 1e3 * rec["narrow_ms"] / max(1, rec["narrow_levels"])
        print(json.dumps(rec), flush=True)
        out.append(rec)
        prog.close()
        tr = os.environ.get("ZKPOR_SOLVE_TRACE")
        if tr and os.path.exists(tr):
            summarise_trace(tr)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(workload=wl.describe(), runs=out), open(os.path.join(ROOT, "gpurun_out", f"solver_bench_{log_n}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
