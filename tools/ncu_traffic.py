"""DRAM traffic of the bucket-accumulation kernels from an `ncu --set full` capture (development tool).

  capture (GPU box):  ncu --set full --clock-control none -k regex:k_accumulate -c 4 -f -o gpurun_out/r02_accumulate python tools/ncu_traffic.py run 24
  summarise (here):   python tools/ncu_traffic.py summarise gpurun_out/r02_accumulate.ncu-rep 24

`run LOG_N` does one G1 and one G2 multi-scalar multiplication of 2^LOG_N uniform terms (so the number of terms behind each captured launch is
known); `summarise` writes profiles/ncu_traffic.json, which bench.py reads for `roofline.traffic` -- keyed by the hash of csrc/msm.cu, so a
capture from other sources is reported as stale instead of silently reused."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(log_n):
    import torch
    import zkpor_b200 as zk
    ctx = zk.Context(0)
    n = 1 << log_n
    buf = lambda b: torch.empty(b // 8, dtype=torch.int64, device="cuda")
    p1, p2, sc = buf(n * 64), buf((n >> 2) * 128), buf(n * 32)
    zk.synth_points_g1(ctx, 12345, 67891, n, p1); zk.synth_points_g2(ctx, 777, 31, n >> 2, p2); zk.synth_scalars(ctx, 9, n, 0, sc)
    for _ in range(2):
        ctx.msm_g1(p1, sc, n)
        ctx.msm_g2(p2, sc, n >> 2)
    print("done", n)


def summarise(rep, log_n):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    src = hashlib.sha256(open(os.path.join(ROOT, "zkmerkle-proof-of-solvency_b200", "csrc", "msm.cu"), "rb").read()).hexdigest()[:12]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        if "k_accumulate" not in name or "heavy" in name:
            continue
        g2 = "Fp2" in name or "Fe2" in name or "<ff::Fp2" in name
        unit = lambda k: float(d[k]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(rows[1][hdr.index(k)], 1.0)
        rec = dict(kernel="k_accumulate", field="Fp2" if g2 else "Fp", terms=(1 << log_n) >> (2 if g2 else 0),
                   dram_bytes=unit("dram__bytes_read.sum") + unit("dram__bytes_write.sum"), duration_ms=unit("gpu__time_duration.sum") if False else float(d["gpu__time_duration.sum"]),
                   duration_unit=rows[1][hdr.index("gpu__time_duration.sum")], registers=int(float(d["launch__registers_per_thread"])),
                   fmaheavy_pct=float(d.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "nan") or "nan"),
                   issue_active_pct=float(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "nan") or "nan"),
                   warps_active_pct=float(d.get("sm__warps_active.avg.pct_of_peak_sustained_active", "nan") or "nan"),
                   msm_cu_sha=src, source=f"ncu --set full, {os.path.basename(rep)}, 2^{log_n} uniform terms")
        rec["bytes_per_term"] = rec["dram_bytes"] / rec["terms"]
        out.append(rec)
    # one record per field: the last capture of each
    best = {}
    for rec in out:
        best[rec["field"]] = rec
    json.dump(dict(kernels=list(best.values())), open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(list(best.values()), indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(int(sys.argv[2]))
    else:
        summarise(sys.argv[2], int(sys.argv[3]))
