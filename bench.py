#!/usr/bin/env python
"""bench.py -- proofs/hour of the B200-native Groth16 prover on a synthetic zkpor50_1380-shaped batch (2^26 domain).

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one JSON line on rank 0.

One "step" = one pass of the hot path = groth16.Prove after the solver (src/prover/prover/prover.go:269) for one batch:
computeH (7 NTTs of size 2^26 + pointwise) and the proof's multi-scalar multiplications (A, B1, K, Z in G1, B in G2,
the Pedersen commitment and its proof of knowledge), the proving key resident in HBM, through the C-ABI.

  value  : proofs/hour, inputs (wire vector, a, b, c) already resident in HBM when the timed region starts.
  e2e    : same, through the same C-ABI call with PINNED HOST buffers -- the H2D copies of the wire vector and the
           a/b/c vectors and the D2H of the 388 proof bytes are inside the timed region.
  roofline: dominant kernel = G1 bucket accumulation (k_accumulate<Fp>); achieved = 96 B/term (SURVEY.md 8(d):
           64 B affine point + 32 B scalar) x terms per launch / CUDA-event duration per launch, against the measured
           HBM copy bandwidth in MEASURED_PEAKS.json.  The kernel is integer-ALU bound, so the fraction is
           expected to be ~1%: the line also carries achieved field multiplications per second.
  cpu_baseline / --impl reference: the oracle's CPU prover (oracle/c, a restatement of gnark's algorithm -- gnark
           itself cannot be built here: no Go toolchain, modules not vendored) on all host cores, on a bounded sample
           (2^20-domain batch of the same shape), scaled by the domain ratio.

N > 1 (torchrun, one rank per GPU), two modes:
  --mode replicas (default): proofs are independent objects (the reference scales the same way: several prover
      processes pulling batches from one queue, README.md:126) -- every rank holds a full key and proves its own
      batch each step; no data-path collective; value = N*K proofs / max-over-ranks time; scaling = "weak".
  --mode sharded: ONE proof per step across the N GPUs (BASELINE config 4): the key is sharded by point chunk; rank 0
      runs computeH while the others start on the wire-side MSMs (rank 0 holds a correspondingly smaller chunk), h is
      broadcast over NVLink, every rank adds its share of the Z MSM, one NCCL all-gather of the partial sums (2 KiB per
      rank), every rank finishes the proof.  scaling = "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
SHAPE = dict(  # zkpor50_1380 shape parameters, as fractions of the domain size n
    constraints=65_000_000 / (1 << 26),   # README.md:10-21 -> ~65.0 M R1CS on a 2^26 domain
    wires=0.984,                          # nbWires / n  (unknown until keygen; placeholder)
    inf_a=0.25, inf_b=0.25,               # fraction of wires whose A / B query is the point at infinity
    committed=1 / 16,                     # BSB22-committed wires (range-check limbs + lookup entries) / n
)
SEEDS = dict(A=(11, 101), B=(12, 103), K=(13, 107), Z=(14, 109), CK=(15, 113))
TOXIC = dict(alpha=0xA11CE, beta=0xB0B, delta=0xDE17A, sigma=0x51634)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--log-n", type=int, default=26)
    ap.add_argument("--cpu-log-n", type=int, default=20)
    ap.add_argument("--scalars", default="uniform", choices=["uniform", "witness"])
    ap.add_argument("--mode", default="replicas", choices=["replicas", "sharded"],
                    help="N>1: replicas = one independent proof per GPU per step (weak scaling, no data-path collective); "
                         "sharded = ONE proof per step, key sharded by point chunk + NCCL all-gather of partials (strong scaling)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def shape_for(log_n):
    n = 1 << log_n
    W = int(n * SHAPE["wires"]) & ~3
    n_public = 2
    n_ck = max(4, int(n * SHAPE["committed"]))
    c_first = W // 2
    inf_a = np.zeros(W, dtype=np.uint8); inf_a[3::4] = 1
    inf_b = np.zeros(W, dtype=np.uint8); inf_b[1::4] = 1
    inf_a[:n_public] = 0; inf_b[:n_public] = 0
    committed = np.arange(c_first, c_first + n_ck, dtype=np.uint64)
    commitment_index = c_first + n_ck
    n_a = int(W - inf_a.sum()); n_b = int(W - inf_b.sum())
    n_k = W - n_public - n_ck - 1
    return dict(n=n, log_n=log_n, W=W, n_public=n_public, n_ck=n_ck, inf_a=inf_a, inf_b=inf_b, committed=committed,
                commitment_index=commitment_index, n_a=n_a, n_b=n_b, n_k=n_k, n_z=n - 1,
                n_constraints=min(n, int(n * SHAPE["constraints"])))


# Sharded (one proof across N GPUs) schedule: rank 0 runs computeH while the other ranks start on the wire-side MSMs; h is
# then broadcast over NVLink and every rank takes an even share of the Z MSM.  Per-proof kernel time at 2^26 on one B200
# (profiles/r01_SUMMARY.md): computeH 113 ms, wire-side MSMs (A, B1, B2, K, commitment) ~750 ms, Z MSM ~140 ms.
COST_MS = dict(ntt=113.0, msm_wires=750.0, msm_z=140.0)


def shard_weight_rank0(world):
    """fraction of the wire-side MSM work given to rank 0 so that all ranks finish together"""
    if world <= 1:
        return 1.0
    t = (COST_MS["ntt"] + COST_MS["msm_wires"] + COST_MS["msm_z"]) / world
    return min(1.0 / world, max(0.0, (t - COST_MS["ntt"] - COST_MS["msm_z"] / world) / COST_MS["msm_wires"]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def dev_buf(torch, nbytes):
    return torch.empty((nbytes + 7) // 8, dtype=torch.int64, device="cuda")


def single_point(torch, zk, ctx, k, g2=False):
    buf = dev_buf(torch, 128 if g2 else 64)
    (zk.synth_points_g2 if g2 else zk.synth_points_g1)(ctx, k, 1, 1, buf)
    return buf.cpu().numpy().view(np.uint64).copy()


def build_key(torch, zk, ctx, sh, rank=0, world=1):
    """Synthetic proving key of the zkpor50_1380 shape directly in HBM (points with known discrete logs).  With
    world > 1 this rank holds the [rank/world, (rank+1)/world) chunk of every array (point-chunk sharding)."""
    f0 = shard_weight_rank0(world)

    def chunk(L, even=False):
        """rank 0 also runs computeH, so it takes the fraction f0 of every wire-side array (the rest is split evenly over the
        other ranks); the Z array, consumed after h has been broadcast, is split evenly."""
        if world == 1 or even:
            return (L * rank) // world, (L * (rank + 1)) // world
        cut0 = int(L * f0)
        if rank == 0:
            return 0, cut0
        rest = L - cut0
        return cut0 + (rest * (rank - 1)) // (world - 1), cut0 + (rest * rank) // (world - 1)

    arrays, counts = {}, {}
    for name, L, g2 in (("A", sh["n_a"], False), ("B1", sh["n_b"], False), ("K", sh["n_k"], False), ("Z", sh["n_z"], False),
                        ("B2", sh["n_b"], True), ("ck", sh["n_ck"], False), ("ck_sigma", sh["n_ck"], False)):
        lo, hi = chunk(L, even=(name == "Z"))
        key = {"B1": "B", "B2": "B", "ck": "CK", "ck_sigma": "CK"}.get(name, name)
        k0, d = SEEDS[key]
        if name == "ck_sigma":
            k0, d = k0 * TOXIC["sigma"] % R, d * TOXIC["sigma"] % R
        buf = dev_buf(torch, (hi - lo) * (128 if g2 else 64))
        (zk.synth_points_g2 if g2 else zk.synth_points_g1)(ctx, (k0 + lo * d) % R, d, hi - lo, buf)
        arrays[name], counts[name] = buf, (lo, hi)
    pts = dict(alpha1=single_point(torch, zk, ctx, TOXIC["alpha"]), beta1=single_point(torch, zk, ctx, TOXIC["beta"]),
               delta1=single_point(torch, zk, ctx, TOXIC["delta"]), beta2=single_point(torch, zk, ctx, TOXIC["beta"], True),
               delta2=single_point(torch, zk, ctx, TOXIC["delta"], True))
    common = dict(log_n=sh["log_n"], A=arrays["A"], B1=arrays["B1"], K=arrays["K"], Z=arrays["Z"], B2=arrays["B2"],
                  n_a=counts["A"][1] - counts["A"][0], n_b=counts["B1"][1] - counts["B1"][0], n_k=counts["K"][1] - counts["K"][0],
                  n_z=counts["Z"][1] - counts["Z"][0], ck_basis=arrays["ck"], ck_basis_exp_sigma=arrays["ck_sigma"], **pts)
    if world == 1:
        pk = zk.ProvingKey(ctx, infinity_a=sh["inf_a"], infinity_b=sh["inf_b"], n_public=sh["n_public"],
                           private_committed=sh["committed"], commitment_index=sh["commitment_index"], **common)
    else:
        lo, hi = counts["ck"]
        pk = zk.ProvingKey(ctx, private_committed=np.zeros(hi - lo, dtype=np.uint64), **common)
    return pk, arrays, counts


def build_inputs(torch, zk, ctx, sh, kind):
    """wire vector + a, b, c = a o b on the constraint domain (a satisfying assignment's evaluation vectors)."""
    W, m = sh["W"], sh["n_constraints"]
    wires = dev_buf(torch, W * 32); a = dev_buf(torch, m * 32); b = dev_buf(torch, m * 32); c = dev_buf(torch, m * 32)
    zk.synth_scalars(ctx, 0xB200, W, 2 if kind == "witness" else 0, wires)
    zk.synth_scalars(ctx, 0xB201, m, 0, a); zk.synth_scalars(ctx, 0xB202, m, 0, b)
    ctx.fr_mul(a, b, c, m)
    return wires, a, b, c


def cpu_prover_sample(torch, zk, ctx, log_n, kind, threads=0):
    """Times the oracle's CPU Groth16 prover (oracle/c: Pippenger MSM G1/G2, radix-2 NTT, all host threads) on a
    2^log_n-domain batch of the same shape; also returns the GPU proof of the same batch for the parity flag."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
    import orc
    sh = shape_for(log_n)
    pk, arrays, _ = build_key(torch, zk, ctx, sh)
    wires, a, b, c = build_inputs(torch, zk, ctx, sh, kind)
    r, s = 0x1234567890ABCDEF1234567890ABCDEF % R, 0xFEDCBA0987654321FEDCBA0987654321 % R
    gpu_proof = pk.prove(wires, a, b, c, sh["n_constraints"], r, s)
    host = lambda t: t.cpu().numpy().view(np.uint64)
    arr = dict(A=host(arrays["A"]), B1=host(arrays["B1"]), K=host(arrays["K"]), Z=host(arrays["Z"]), B2=host(arrays["B2"]),
               ck_basis=host(arrays["ck"]), ck_basis_exp_sigma=host(arrays["ck_sigma"]), log_n=log_n, **pk.points)
    w = host(wires).reshape(-1, 4)
    keep_k = np.ones(sh["W"], dtype=bool); keep_k[:sh["n_public"]] = False
    keep_k[sh["committed"].astype(np.int64)] = False; keep_k[sh["commitment_index"]] = False
    wa, wb, wk, cm = w[sh["inf_a"] == 0], w[sh["inf_b"] == 0], w[keep_k], w[sh["committed"].astype(np.int64)]
    ha, hb, hc = host(a).reshape(-1, 4), host(b).reshape(-1, 4), host(c).reshape(-1, 4)
    # thread count: the fastest of {OpenMP default, all logical CPUs, half of them} -- the port must not be handicapped by
    # SMT oversubscription or by torchrun's OMP_NUM_THREADS=1
    ncpu = len(os.sched_getaffinity(0))
    cands = [threads] if threads else sorted({max(1, orc.lib().orc_num_threads()), ncpu, max(1, ncpu // 2)})
    if not threads and max(cands) == 1:
        cands = [1]
    best = None
    for nt in cands:
        t0 = time.perf_counter()
        proof_nt = orc.groth16_prove(arr, wa, wb, wk, cm, ha, hb, hc, r, s, threads=nt)
        dt_nt = time.perf_counter() - t0
        if best is None or dt_nt < best[0]:
            best = (dt_nt, nt, proof_nt)
    dt, nthreads, cpu_proof = best
    pk.close()
    return dt, nthreads, cpu_proof == gpu_proof


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    if args.impl == "reference":
        os.environ.pop("OMP_NUM_THREADS", None)   # torchrun pins it to 1; the CPU arm uses every host core
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner on stdout otherwise; stdout carries ONE JSON line
    import torch
    import zkpor_b200 as zk
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1 and args.impl == "native":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = zk.Context(local)
    sh = shape_for(args.log_n)
    ratio = (1 << args.log_n) / (1 << args.cpu_log_n)
    workload = f"zkpor50_1380-shaped synthetic batch: Groth16 Prove after the solver, domain 2^{args.log_n}, " \
               f"{sh['n_constraints']} constraints, {sh['W']} wires, {args.scalars} scalars"

    if args.impl == "reference":
        dt, nthreads, ok = cpu_prover_sample(torch, zk, ctx, args.cpu_log_n, args.scalars)
        times = [dt]
        for _ in range(max(0, min(args.steps, 3) - 1)):      # later steps reuse the thread count the first one selected
            times.append(cpu_prover_sample(torch, zk, ctx, args.cpu_log_n, args.scalars, threads=nthreads)[0])
        t = min(times)
        v = 3600.0 / (t * ratio)
        line = {"impl": "reference", "metric": "proofs/hour", "value": v, "unit": "proofs/hour", "n_gpus": args.gpus, "steps": len(times),
                "warmup": 0, "ms_per_step": t * ratio * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic",
                "config": {"workload": workload, "sampled": f"CPU prover timed on a 2^{args.cpu_log_n}-domain batch of the same shape, "
                                                             f"time x {ratio:.0f} (domain ratio)"},
                "cpu_baseline": {"value": v, "unit": "proofs/hour", "cores": nthreads, "kind": "port",
                                 "sample": f"2^{args.cpu_log_n}-domain batch, {t:.2f} s, scaled x{ratio:.0f}", "parity_with_gpu": ok},
                "e2e": {"value": v, "unit": "proofs/hour", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------- native arm
    t_setup = time.perf_counter()
    sharded = world > 1 and args.mode == "sharded"
    pk, arrays, counts = build_key(torch, zk, ctx, sh, rank, world) if sharded else build_key(torch, zk, ctx, sh)
    wires, a, b, c = build_inputs(torch, zk, ctx, sh, args.scalars)
    r, s = 0x1234567890ABCDEF1234567890ABCDEF % R, 0xFEDCBA0987654321FEDCBA0987654321 % R
    m = sh["n_constraints"]
    stream = torch.cuda.ExternalStream(ctx.stream())

    if not sharded:
        def step(w_, a_, b_, c_):
            return pk.prove(w_, a_, b_, c_, m, r, s)
    else:
        w4 = wires.view(-1, 4)
        def sl(name, idx):
            lo, hi = counts[name]
            return idx[lo:hi]
        ia = torch.from_numpy(np.nonzero(sh["inf_a"] == 0)[0]).cuda(); ib = torch.from_numpy(np.nonzero(sh["inf_b"] == 0)[0]).cuda()
        keep = np.ones(sh["W"], dtype=bool); keep[:sh["n_public"]] = False; keep[sh["committed"].astype(np.int64)] = False; keep[sh["commitment_index"]] = False
        ik = torch.from_numpy(np.nonzero(keep)[0]).cuda(); ic = torch.from_numpy(sh["committed"].astype(np.int64)).cuda()
        ia, ib, ik, ic = sl("A", ia), sl("B1", ib), sl("K", ik), sl("ck", ic)
        h = dev_buf(torch, sh["n"] * 32)
        zlo, zhi = counts["Z"]
        gathered = torch.empty((world, 2, zk.PROVE_PARTIAL_BYTES), dtype=torch.uint8, device="cuda")

        def step(w_, a_, b_, c_):
            if not w_.is_cuda:    # e2e: this rank's H2D copies are part of the step (a, b, c are only needed by the computeH rank)
                w_ = w_.cuda(non_blocking=True)
                if rank == 0:
                    a_, b_, c_ = a_.cuda(non_blocking=True), b_.cuda(non_blocking=True), c_.cuda(non_blocking=True)
            wv = w_.view(-1, 4)
            wa, wb, wk, cm = wv[ia].contiguous(), wv[ib].contiguous(), wv[ik].contiguous(), wv[ic].contiguous()
            torch.cuda.synchronize()
            if rank == 0:
                ctx.compute_h(a_, b_, c_, m, sh["log_n"], out=h)
            part_w = pk.prove_partial(wa, wb, wk, cm, None, 0)   # rank 0's chunk is smaller by the time computeH takes
            dist.broadcast(h, src=0)                              # 2 GiB over NVLink, all ranks arrive together
            torch.cuda.synchronize()
            part_z = pk.prove_partial(None, None, None, None, h.view(-1, 4)[zlo:zhi], zhi - zlo)
            mine = torch.from_numpy(np.stack([part_w, part_z])).cuda()
            dist.all_gather_into_tensor(gathered, mine)
            return pk.finish(gathered.cpu().numpy().reshape(-1, zk.PROVE_PARTIAL_BYTES), r, s)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n_steps, inputs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            proof = step(*inputs)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        ms = max(dev_ms, 0.0)
        if dist is not None:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, wall, proof

    setup_s = time.perf_counter() - t_setup
    for _ in range(args.warmup):
        proof = step(wires, a, b, c)
    sampler = ClockSampler(local); sampler.start()
    ctx.kernel_timing(True)
    l0 = ctx.launch_count()
    ms, wall, proof = run(args.steps, (wires, a, b, c))
    launches = ctx.launch_count() - l0
    kstats = {name: ctx.kernel_stats(k) for k, name in enumerate(["accumulate_g1", "accumulate_g2", "ntt_pass", "digits_scatter"])}
    ctx.kernel_timing(False)
    clocks = sampler.stop()
    proofs_per_step = world if (world > 1 and not sharded) else 1
    value = proofs_per_step * args.steps / (ms / 1e3) * 3600.0

    # ---- e2e: pinned host inputs through the same call
    e2e = None
    if not args.no_e2e:
        hw, ha, hb, hc = (x.cpu().pin_memory() for x in (wires, a, b, c))
        h2d = sum(x.numel() * 8 for x in (hw, ha, hb, hc))
        if not sharded:
            inputs = tuple(x.numpy() for x in (hw, ha, hb, hc))
        else:
            inputs = (hw, ha, hb, hc)
        step(*inputs)
        ems, ewall, eproof = run(args.steps, inputs)
        assert eproof == proof, "e2e proof differs from the device-resident proof"
        e2e = {"value": proofs_per_step * args.steps / (ems / 1e3) * 3600.0, "unit": "proofs/hour", "h2d_bytes_per_step": h2d * proofs_per_step,
               "d2h_bytes_per_step": len(proof) * proofs_per_step,
               "ms_per_step": ems / args.steps}
        del hw, ha, hb, hc

    # ---- roofline of the dominant kernel (G1 bucket accumulation)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    ks = kstats["accumulate_g1"]
    roof = None
    if ks["launches"]:
        per_launch_ms = ks["total_ms"] / ks["launches"]
        terms_per_launch = ks["units"] / ks["launches"]
        achieved = 96.0 * terms_per_launch / (per_launch_ms / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_accumulate<Fp> (G1 bucket accumulation)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel: 86.63 GB for the 49 526 340-term launch of this
                # very workload under `ncu --set full` (profiles/r01_SUMMARY.md) = 1 749 B/term, scaled to the average launch
                "traffic": 1749.2 * terms_per_launch, "traffic_source": "ncu --set full, bench.py at 2^26: 85.75 GB read + 0.88 GB written per 49.5 M-term launch (profiles/r01_SUMMARY.md)",
                "bytes_per_term": 96, "terms_per_launch": terms_per_launch,
                "launch_ms": per_launch_ms, "launches": ks["launches"], "share_of_step": ks["total_ms"] / ms,
                "note": "integer-ALU bound: ~13 windows x 10 field mul per term; see DESIGN.md for the modmul roofline"}
    breakdown = {k: {"ms_per_step": v["total_ms"] / args.steps, "launches_per_step": v["launches"] / args.steps} for k, v in kstats.items()}

    line = {"metric": "proofs/hour", "value": value, "unit": "proofs/hour", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic",
            "config": {"workload": workload, "parallelism": "single GPU" if world == 1 else (f"point-chunk sharded MSM x{world} (rank 0 also runs computeH and broadcasts h over NVLink) + NCCL all-gather of the 2 KiB partials" if sharded
                                                                        else f"{world} independent proofs per step, one per GPU (full key per GPU), no data-path collective"),
                       "proofs_per_step": proofs_per_step,
                       "l2": "inputs (>= 2 GB per vector, 21 GB key) are far larger than the 126 MB L2; no explicit flush needed",
                       "key": "synthetic key in HBM: points (k0 + i*d)*G per array", "timing": "CUDA events on the library stream, max over ranks",
                       "setup_s": setup_s},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernel_breakdown": breakdown, "wall_ms_per_step": wall / args.steps * 1e3,
            "proof_sha": __import__("hashlib").sha256(proof).hexdigest()[:16]}
    if e2e:
        line["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu:
        pk.close(); del arrays
        torch.cuda.empty_cache()
        dt, nthreads, ok = cpu_prover_sample(torch, zk, ctx, args.cpu_log_n, args.scalars)
        line["cpu_baseline"] = {"value": 3600.0 / (dt * ratio), "unit": "proofs/hour", "cores": nthreads, "kind": "port",
                                "sample": f"oracle CPU prover on a 2^{args.cpu_log_n}-domain batch of the same shape: {dt:.2f} s, scaled x{ratio:.0f} (domain ratio)",
                                "parity_with_gpu": ok}
    if dist is not None:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
