#!/usr/bin/env python
"""bench.py -- proofs/hour of the B200-native Groth16 prover on a synthetic zkpor50_1380-shaped batch (2^26 domain).

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one JSON line on rank 0.

One "step" = one pass of the hot path = the WHOLE of groth16.Prove (src/prover/prover/prover.go:269) for one batch PER PROVER, through
the C-ABI call zkpor_groth16_prove_solve, from the circuit's inputs (the 1 public + ~2 M secret values gnark's witness holds):
  witness solver on the device (level schedule, hints, the BSB22 commitment mid-solve; its serial tail runs beside the
  multiplications) -> a = Lw, b = Rw, c = Ow -> computeH (7 NTTs of size 2^26 + pointwise) -> the proof's multi-scalar
  multiplications (A, B1, K, Z in G1, B in G2) -> 388 proof bytes.
The proving key and the compiled constraint system are resident in HBM (as gnark keeps pk and r1cs in memory across proofs).
--provers P (default 2): P provers per GPU, each with its own context, key and program, prove independent batches from P host
threads (the reference runs several prover processes per machine, README.md:126); `one_proof_alone_ms` is one prover's latency.

Workload.  gnark's frontend is Go and cannot compile circuit/batch_create_user_circuit.go here, so the constraint system is a
BatchCreateUser-SHAPED synthetic one (circuit_synth.batch_create_user_like: per-user blocks of 50 assets with range checks, tier
lookups, integer divisions, Poseidon commitments and a 28-level Merkle path; 500 CEX assets with 12 tiers; an 834-permutation
serial sponge chain; one commitment with two log-derivative arguments), sized to ~65.0 M constraints on the 2^26 domain
(README.md:10-21).  The key is synthetic too: points (k0 + i*d)*G per array, so the timed proof can be checked EXACTLY by
discrete-log arithmetic (`parity` in the JSON line; the run fails if it does not hold).

  value  : proofs/hour, inputs already resident in HBM when the timed region starts (device time between two full synchronisations).
  e2e    : same call with PINNED HOST inputs -- the H2D copy of the inputs and the D2H of the 388 proof bytes are inside the
           timed region (the solver runs on the device, so only ~66 MB cross PCIe per proof).
  roofline: dominant kernel = G1 bucket accumulation (k_accumulate<Fp>); achieved = 96 B/term (64 B affine
           point + 32 B scalar) x terms per launch / CUDA-event duration per launch, against the measured HBM copy bandwidth
           in MEASURED_PEAKS.json.  The kernel is integer-ALU bound, so the fraction is ~1%: `modmul_roofline`
           carries achieved field multiplications per second against the measured IMAD.WIDE ceiling.
  cpu_baseline: the oracle's CPU prover (oracle/c: the same solver walk over the host threads, signed-digit Pippenger MSM,
           radix-2 NTT -- a restatement of gnark's algorithm; gnark itself cannot be built here: no Go toolchain, modules not
           vendored) on a bounded sample: the same circuit family on a 2^--cpu-log-n domain, time scaled by the constraint ratio.
  --impl reference: ONE full-size proof (same 2^26 circuit, same key) by that CPU prover on all host threads -- a measurement,
           not an extrapolation; steps/warmup are ignored beyond that (a 2^26 CPU proof takes minutes).
  --workload witness: the witness service's hot path on --accounts synthetic accounts (tools/witness_bench.py).

N > 1 (torchrun, one rank per GPU): proofs are independent objects (the reference scales the same way: several prover
processes pulling batches from one queue, README.md:126) -- every rank holds full keys and proves its own batches each step;
no data-path collective; value = N*P*K proofs / max-over-ranks time; scaling = "weak".  The same line carries `sharded`: ONE
proof across the N GPUs through the library's own NCCL path, whole call and post-solver part.
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
    ap.add_argument("--cpu-log-n", type=int, default=21, help="domain of the bounded CPU sample of the native arm's cpu_baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the discrete-log check of the timed proof (it costs ~1 min of host time at 2^26)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the one-proof-across-N-GPUs measurement")
    ap.add_argument("--provers", type=int, default=2, help="provers per GPU (each with its own context, key and program), proving independent batches concurrently")
    ap.add_argument("--workload", default="prove", choices=["prove", "witness"],
                    help="prove: groth16.Prove (the headline metric); witness: the witness service's hot path on --accounts synthetic accounts (BASELINE config 5)")
    ap.add_argument("--accounts", type=int, default=10_000_000)
    ap.add_argument("--tier", type=int, default=50, choices=[50, 500],
                    help="assets per user of the circuit: 50 = zkpor50_1380 (the headline), 500 = zkpor500_200 (BASELINE config 4: ~62.9 M constraints)")
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


def build_key(torch, zk, ctx, sh, arrays=None, shard=False):
    """Synthetic proving key of the given shape directly in HBM: array X holds (k0_X + i*d_X)*G (known discrete logs).
    shard=True: upload only the chunks of ctx's rank in its group (the arrays describe the whole key)."""
    have = arrays is not None
    arrays = arrays or {}
    for name, L, g2 in (("A", sh["n_a"], False), ("B1", sh["n_b"], False), ("K", sh["n_k"], False), ("Z", sh["n_z"], False),
                        ("B2", sh["n_b"], True), ("ck", sh["n_ck"], False), ("ck_sigma", sh["n_ck"], False)):
        if have:
            break
        key = {"B1": "B", "B2": "B", "ck": "CK", "ck_sigma": "CK"}.get(name, name)
        k0, d = SEEDS[key]
        if name == "ck_sigma":
            k0, d = k0 * TOXIC["sigma"] % R, d * TOXIC["sigma"] % R
        buf = dev_buf(torch, L * (128 if g2 else 64))
        (zk.synth_points_g2 if g2 else zk.synth_points_g1)(ctx, k0 % R, d, L, buf)
        arrays[name] = buf
    pts = dict(alpha1=single_point(torch, zk, ctx, TOXIC["alpha"]), beta1=single_point(torch, zk, ctx, TOXIC["beta"]),
               delta1=single_point(torch, zk, ctx, TOXIC["delta"]), beta2=single_point(torch, zk, ctx, TOXIC["beta"], True),
               delta2=single_point(torch, zk, ctx, TOXIC["delta"], True))
    pk = zk.ProvingKey(ctx, log_n=sh["log_n"], A=arrays["A"], B1=arrays["B1"], K=arrays["K"], Z=arrays["Z"], B2=arrays["B2"],
                       n_a=sh["n_a"], n_b=sh["n_b"], n_k=sh["n_k"], n_z=sh["n_z"], ck_basis=arrays["ck"], ck_basis_exp_sigma=arrays["ck_sigma"],
                       infinity_a=sh["inf_a"], infinity_b=sh["inf_b"], n_public=sh["n_public"], private_committed=sh["committed"],
                       commitment_index=sh["commitment_index"], shard=shard, **pts)
    return pk, arrays, None


def build_inputs(torch, zk, ctx, sh, kind):
    """wire vector + a, b, c = a o b on the constraint domain (a satisfying assignment's evaluation vectors)."""
    W, m = sh["W"], sh["n_constraints"]
    wires = dev_buf(torch, W * 32); a = dev_buf(torch, m * 32); b = dev_buf(torch, m * 32); c = dev_buf(torch, m * 32)
    zk.synth_scalars(ctx, 0xB200, W, 2 if kind == "witness" else 0, wires)
    zk.synth_scalars(ctx, 0xB201, m, 0, a); zk.synth_scalars(ctx, 0xB202, m, 0, b)
    ctx.fr_mul(a, b, c, m)
    return wires, a, b, c


# ---------------------------------------------------------------------------------------------------------------- circuit workload
FULL_CIRCUIT = dict(assets_per_user=50, cex_assets=500, tiers=12, merkle_depth=28, chain_perms=834)   # the tier-50 batch
CONSTRAINT_FILL = 65_000_000 / (1 << 26)         # README.md:10-21: ~65.0 M constraints on the 2^26 domain


TIER = 50                                        # --tier: assets per user (50 -> zkpor50_1380, 500 -> zkpor500_200)


def circuit_params(log_n):
    """BatchCreateUser-shaped circuit for a 2^log_n domain: the per-user block is the tier's one; the global parts (CEX
    assets, the serial commitment chain) scale with the domain below 2^26; the number of users fills the domain to ~97 % (tier 50:
    65.0 M constraints) / ~94 % (tier 500: 62.9 M, README.md:10-21)."""
    f = min(1.0, (1 << log_n) / (1 << 26))
    return dict(FULL_CIRCUIT, assets_per_user=TIER, cex_assets=max(4, int(FULL_CIRCUIT["cex_assets"] * f)), chain_perms=max(2, int(FULL_CIRCUIT["chain_perms"] * f)))


def build_circuit(zk, log_n, xp=np, device=None):
    import importlib
    cs_mod = importlib.import_module("zkmerkle-proof-of-solvency_b200.circuit_synth")
    params = circuit_params(log_n)
    target = int((1 << log_n) * (CONSTRAINT_FILL if TIER == 50 else 62_900_000 / (1 << 26)))
    rows = lambda cb: sum(len(s.rows) * s.count for s in cb.sections)
    r1 = rows(cs_mod.batch_create_user_like(users=1, poseidon_constants=zk.poseidon_constants, **params))
    r2 = rows(cs_mod.batch_create_user_like(users=2, poseidon_constants=zk.poseidon_constants, **params))
    users = max(1, (target - (r1 - (r2 - r1))) // (r2 - r1))
    cb = cs_mod.batch_create_user_like(users=users, poseidon_constants=zk.poseidon_constants, **params)
    flat = cb.flatten(xp=xp, device=device)
    assert flat["n_constraints"] <= (1 << log_n), (flat["n_constraints"], log_n)
    return cs_mod, flat, dict(params, users=int(users))


def circuit_shape(torch, flat, log_n):
    """the key's shape from the constraint system, as gnark's Setup derives it: InfinityA / InfinityB = wires absent from L / R"""
    W = flat["n_wires"]
    dev = flat["l_wire"].device if hasattr(flat["l_wire"], "device") else None
    if dev is not None:
        ia = torch.ones(W, dtype=torch.uint8, device=dev); ia[flat["l_wire"].long()] = 0
        ib = torch.ones(W, dtype=torch.uint8, device=dev); ib[flat["r_wire"].long()] = 0
        inf_a, inf_b = ia.cpu().numpy(), ib.cpu().numpy()
        committed = flat["private_committed"].cpu().numpy().astype(np.uint64)
    else:
        inf_a = np.ones(W, dtype=np.uint8); inf_a[np.asarray(flat["l_wire"], dtype=np.int64)] = 0
        inf_b = np.ones(W, dtype=np.uint8); inf_b[np.asarray(flat["r_wire"], dtype=np.int64)] = 0
        committed = np.asarray(flat["private_committed"], dtype=np.uint64)
    n = 1 << log_n
    n_ck = len(committed)
    return dict(n=n, log_n=log_n, W=W, n_public=flat["n_public"], n_ck=n_ck, inf_a=inf_a, inf_b=inf_b, committed=committed,
                commitment_index=int(flat["commitment_index"]), n_a=int(W - inf_a.sum()), n_b=int(W - inf_b.sum()),
                n_k=W - flat["n_public"] - n_ck - 1, n_z=n - 1, n_constraints=flat["n_constraints"])


def inputs_to_mont(torch, ctx, plain_np):
    """canonical limbs -> Montgomery fr.Elements on the device (one field product by R^2)"""
    x = torch.from_numpy(plain_np.view(np.int64)).cuda()
    r2 = pow(1 << 256, 2, R)
    k = torch.from_numpy(np.array([[(r2 >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64).view(np.int64)).cuda().repeat(x.shape[0], 1)
    out = torch.empty_like(x)
    ctx.fr_mul(x, k, out, x.shape[0])
    return out


def host_limbs(t):
    return t.cpu().numpy().view(np.uint64).reshape(-1, 4)


def check_proof_by_discrete_logs(orc, ctx, sh, wires_dev, a, b, c, proof, r, s):
    """The key arrays are (k0 + i*d)*G, so every proof element has a discrete log computable with O(n) field additions on the host
    (oracle: checker only): the 388 timed bytes must equal the bytes derived from the wire vector, h and the toxic scalars."""
    import bn254 as bn
    import groth16 as g16
    from bn254 import FP2, G1_GEN, G2_GEN
    rinv = pow(1 << 256, -1, R)

    def dlog_dot(limbs, k0, d):
        s_, t_ = orc.fr_index_sums(np.ascontiguousarray(limbs))
        return (k0 * s_ + d * t_) * rinv % R

    h = ctx.compute_h(a, b, c, sh["n_constraints"], sh["log_n"])
    w = host_limbs(wires_dev)
    keep_k = np.ones(sh["W"], dtype=bool); keep_k[:sh["n_public"]] = False
    cidx = sh["committed"].astype(np.int64)
    keep_k[cidx] = False; keep_k[sh["commitment_index"]] = False
    S, T = SEEDS, TOXIC
    dA = dlog_dot(w[sh["inf_a"] == 0], *S["A"]); dB = dlog_dot(w[sh["inf_b"] == 0], *S["B"])
    dK = dlog_dot(w[keep_k], *S["K"]); dZ = dlog_dot(h[:sh["n_z"]], *S["Z"])
    dC = dlog_dot(w[cidx], *S["CK"])
    ar = (dA + T["alpha"] + r * T["delta"]) % R
    bs = (dB + T["beta"] + s * T["delta"]) % R
    krs = (dK + dZ - r * s * T["delta"] + s * ar + r * bs) % R
    cpt = bn.pt_mul(G1_GEN, dC)
    want = dict(Ar=bn.pt_mul(G1_GEN, ar), Bs=bn.pt_mul(G2_GEN, bs, FP2), Krs=bn.pt_mul(G1_GEN, krs),
                Commitments=[cpt], CommitmentPok=bn.pt_mul(G1_GEN, dC * T["sigma"] % R))
    ok = proof == g16.proof_raw_bytes(want)
    # the challenge wire holds hash_to_field of that very commitment
    ok &= orc.fr_unmont(w[sh["commitment_index"]])[0] == g16.commitment_challenge(cpt)
    return bool(ok)


class Workload:
    """circuit + key + inputs of one domain size, resident in HBM"""

    def __init__(self, torch, zk, ctx, log_n, seed=0xB200, keep_generator=True):
        t0 = time.perf_counter()
        self.torch, self.zk, self.ctx, self.log_n = torch, zk, ctx, log_n
        self.cs_mod, self.flat, self.params = build_circuit(zk, log_n, xp=torch, device="cuda")
        self.sh = circuit_shape(torch, self.flat, log_n)
        self.prog = zk.Program(ctx, self.flat)
        self.pk, self.arrays, _ = build_key(torch, zk, ctx, self.sh)
        self.inputs_plain = self.cs_mod.draw_inputs(self.flat, seed)
        self.inputs = inputs_to_mont(torch, ctx, self.inputs_plain)
        self.text = self.describe()
        if not keep_generator:
            # the library holds its own copies of the key and of the constraint system: drop the generator's (~50 GB at 2^26)
            self.arrays = None
            for k, v in list(self.flat.items()):
                if hasattr(v, "device"):
                    self.flat[k] = None
            torch.cuda.empty_cache()
        self.setup_s = time.perf_counter() - t0

    def describe(self):
        f, sh, p = self.flat, self.sh, self.params
        return (f"BatchCreateUser-shaped synthetic batch (circuit_synth: {p['users']} users x {p['assets_per_user']} assets, {p['cex_assets']} CEX assets x "
                f"{p['tiers']} tiers, Merkle depth {p['merkle_depth']}, {p['chain_perms']}-permutation commitment chain): whole groth16.Prove incl. witness solver, "
                f"domain 2^{self.log_n}, {f['n_constraints']} constraints, {f['n_wires']} wires, {f['n_public'] - 1 + f['n_secret']} inputs, "
                f"{f['n_levels']} solver levels, {sh['n_ck']} committed wires")

    def close(self):
        self.pk.close(); self.prog.close()
        self.arrays = None


def cpu_full_prove(torch, zk, ctx, log_n, threads=0):
    """the oracle's CPU prover (solver + proof) on the same circuit family at 2^log_n, all host threads; returns timing, thread count,
    the proof, and the GPU proof of the same batch (parity flag)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
    import orc
    wl = Workload(torch, zk, ctx, log_n)
    r, s = RS
    gpu_proof = wl.pk.prove_solve(wl.prog, wl.inputs, r, s)
    host = lambda t: t.cpu().numpy().view(np.uint64)
    arr = dict(A=host(wl.arrays["A"]), B1=host(wl.arrays["B1"]), K=host(wl.arrays["K"]), Z=host(wl.arrays["Z"]), B2=host(wl.arrays["B2"]),
               ck_basis=host(wl.arrays["ck"]), ck_basis_exp_sigma=host(wl.arrays["ck_sigma"]), log_n=log_n, **wl.pk.points)
    flat_h = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in wl.flat.items()}
    for k in ("l_wire", "l_coeff", "r_wire", "r_coeff", "o_wire", "o_coeff", "aux_wire", "aux_coeff", "instr_arg", "level_instr", "hint_fn", "hint_param",
              "hint_out_first", "hint_n_out"):
        flat_h[k] = flat_h[k].view(np.uint32)
    for k in ("l_row_ptr", "r_row_ptr", "o_row_ptr", "aux_row_ptr", "hint_in_ptr", "hint_in_end", "private_committed"):
        flat_h[k] = flat_h[k].view(np.uint64)
    inputs_h = host_limbs(wl.inputs)
    sh, desc, n_constraints = wl.sh, wl.describe(), wl.flat["n_constraints"]
    wl.close(); del wl
    torch.cuda.empty_cache()
    ncpu = len(os.sched_getaffinity(0))
    nthreads = threads or ncpu
    t0 = time.perf_counter()
    cpu_proof, secs = orc.groth16_prove_program(arr, flat_h, sh["inf_a"], sh["inf_b"], inputs_h, r, s, threads=nthreads)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, solve_s=secs[0], prove_s=secs[1], threads=nthreads, host_cpus=ncpu, parity=cpu_proof == gpu_proof, desc=desc, n_constraints=n_constraints)


RS = (0x1234567890ABCDEF1234567890ABCDEF % R, 0xFEDCBA0987654321FEDCBA0987654321 % R)
DTYPE = "u32 limbs (254-bit modular integers)"


def main():
    args = parse()
    global TIER
    TIER = args.tier
    if args.workload == "witness" and args.impl == "native":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import witness_bench
        return witness_bench.main(args)
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
    r, s = RS

    if args.impl == "reference":
        # ONE full-size proof on the host cores: the same circuit, key and inputs as the native arm (the GPU only generates them)
        res = cpu_full_prove(torch, zk, ctx, args.log_n)
        v = 3600.0 / res["seconds"]
        line = {"impl": "reference", "metric": "proofs/hour", "value": v, "unit": "proofs/hour", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": res["desc"], "note": "one full-size CPU proof, not scaled; --steps/--warmup are not repeated (minutes per proof)"},
                "cpu_baseline": {"value": v, "unit": "proofs/hour", "cores": res["threads"], "host_cpus": res["host_cpus"], "kind": "port",
                                 "sample": f"the full 2^{args.log_n} batch: {res['seconds']:.1f} s (solver {res['solve_s']:.1f} s, proof {res['prove_s']:.1f} s)",
                                 "parity_with_gpu": res["parity"]},
                "published_anchor": "gnark CPU, 32 vCPU m5.8xlarge, earlier circuit: 62 s per proof (docs/updated_proof_of_solvency_to_mitigate_dummy_user_attack.md:201)",
                "e2e": {"value": v, "unit": "proofs/hour", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------- native arm
    # P provers per GPU, each with its own context (streams, scratch), key and program, proving independent batches from P host
    # threads: while one proof is in its multiplications (throughput-bound kernels that fill the SMs) another is in its solve (a
    # latency chain of ~8 000 small launches), as the reference runs several prover processes per machine (README.md:126).
    P = max(1, args.provers)
    wl = Workload(torch, zk, ctx, args.log_n, keep_generator=False)
    sh, pk, prog = wl.sh, wl.pk, wl.prog
    workload = wl.text
    provers = [(ctx, wl)]
    for _ in range(1, P):
        free_gb = torch.cuda.mem_get_info()[0] / 1e9
        used_gb = (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9
        if free_gb < 1.6 * used_gb / len(provers) + 4:    # a prover's scratch grows by about half again during its first proofs
            print(f"bench: {free_gb:.0f} GB free, a prover needs ~{1.6 * used_gb / len(provers):.0f} GB: running {len(provers)} prover(s) per GPU", file=sys.stderr)
            break
        c2 = zk.Context(local)
        provers.append((c2, Workload(torch, zk, c2, args.log_n, keep_generator=False)))
    P = len(provers)
    hbm_gb = (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9

    def step(inputs):
        return pk.prove_solve(prog, inputs, r, s)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n_steps, inputs_of, fn=None):
        """n_steps proofs by every prover (fn: prover 0 only, one thread); device time between two full-device synchronisations"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        who = provers if fn is None else provers[:1]
        outs, stages, errs = [[] for _ in who], [], []

        def work(i):
            c, w = who[i]
            try:
                for _ in range(n_steps):
                    outs[i].append(fn(inputs_of(w)) if fn is not None else w.pk.prove_solve(w.prog, inputs_of(w), r, s))
                    if i == 0:
                        stages.append(c.last_timings())
            except Exception as e:     # surfaces below: a thread must not die silently
                errs.append(e)

        th = [threading.Thread(target=work, args=(i,)) for i in range(1, len(who))]
        barrier()
        e0.record()
        t0 = time.perf_counter()
        for t in th:
            t.start()
        work(0)
        for t in th:
            t.join()
        torch.cuda.synchronize()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        if errs:
            raise errs[0]
        ms = max(e0.elapsed_time(e1), 0.0)
        if dist is not None:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        proofs = [p for o in outs for p in o]
        assert all(p == proofs[0] for p in proofs), "provers disagree on the proof of the same batch"
        return ms, wall, proofs[0], stages

    # warm-up, prover by prover: a prover's scratch reaches its size during its first proofs; if the second one does not fit after all
    # (the estimate above is rough), it is dropped and the run goes on with one prover per GPU instead of failing
    for i, (c, w) in enumerate(list(provers)):
        try:
            for _ in range(args.warmup):
                proof = w.pk.prove_solve(w.prog, w.inputs, r, s)
        except zk.ZkporError as e:
            if i == 0 or "memory" not in str(e).lower():
                raise
            print(f"bench: prover {i} does not fit in HBM ({e}); running {i} prover(s) per GPU", file=sys.stderr)
            for c2, w2 in provers[i:]:
                w2.close(); c2.close()
            provers = provers[:i]
            torch.cuda.empty_cache()
            break
    P = len(provers)
    if dist is not None:                                   # every rank must run the same number of provers: the value counts world * P proofs
        t = torch.tensor([P], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) < P:
            for c2, w2 in provers[int(t.item()):]:
                w2.close(); c2.close()
            provers = provers[:int(t.item())]; P = len(provers)
            torch.cuda.empty_cache()
    sampler = ClockSampler(local); sampler.start()
    ctx.kernel_timing(True, classes=[0, 1, 2, 3, 6])   # not the solver's ~8 000 wide launches per proof: an event pair per launch is not free
    l0 = [c.launch_count() for c, _ in provers]
    torch.cuda.cudart().cudaProfilerStart()          # `ncu --profile-from-start off` captures exactly the timed steps
    ms, wall, proof, stages = run(args.steps, lambda w: w.inputs)
    torch.cuda.cudart().cudaProfilerStop()
    launches = sum(c.launch_count() - l for (c, _), l in zip(provers, l0))
    hbm_peak_gb = (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9
    kstats = {name: ctx.kernel_stats(k) for k, name in ((0, "accumulate_g1"), (1, "accumulate_g2"), (2, "ntt_pass"), (3, "digits_scatter"), (6, "solver_tail"))}
    ctx.kernel_timing(False)
    clocks = sampler.stop()
    proofs_per_step = world * P
    value = proofs_per_step * args.steps / (ms / 1e3) * 3600.0
    # latency and stage breakdown of ONE proof alone on the GPU (the timed region runs P provers per GPU)
    solo_ms, _, _, stages = run(1, lambda w: w.inputs, step)
    solve_ms = float(np.mean([st.get("solve", 0.0) for st in stages]))

    # ---- e2e: pinned host inputs through the same call
    e2e = None
    if not args.no_e2e:
        for c, w in provers:
            w.hin = w.inputs.cpu().pin_memory()
            w.inputs_h = w.hin.numpy()
            w.pk.prove_solve(w.prog, w.inputs_h, r, s)
        ems, ewall, eproof, _ = run(args.steps, lambda w: w.inputs_h)
        assert eproof == proof, "e2e proof differs from the device-resident proof"
        e2e = {"value": proofs_per_step * args.steps / (ems / 1e3) * 3600.0, "unit": "proofs/hour", "h2d_bytes_per_step": wl.hin.numel() * 8 * proofs_per_step,
               "d2h_bytes_per_step": len(proof) * proofs_per_step, "ms_per_step": ems / args.steps}

    # the extra provers have done their part: their ~85 GB each go back before the checks below allocate
    for c, w in provers[1:]:
        w.close(); c.close()
    provers = provers[:1]
    torch.cuda.empty_cache()

    # ---- parity of the TIMED proof: exact, by discrete logs (oracle = checker)
    parity = None
    if not args.no_parity and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
        import orc
        t0 = time.perf_counter()
        m = sh["n_constraints"]
        wires_dev = dev_buf(torch, sh["W"] * 32)
        a, b, c = (dev_buf(torch, m * 32) for _ in range(3))
        prog.solve_device(wl.inputs, wires_dev, pk, abc=(a, b, c))
        ok = check_proof_by_discrete_logs(orc, ctx, sh, wires_dev, a, b, c, proof, r, s)
        parity = {"checked_at": f"2^{args.log_n}", "ok": ok, "method": "all five proof points re-derived from the key's discrete logs, the solved wire vector and h; "
                  "challenge wire = hash_to_field(commitment)", "seconds": time.perf_counter() - t0}
        del wires_dev, a, b, c
        if not ok:
            print(json.dumps({"error": "the timed proof failed the discrete-log parity check", "parity": parity}), flush=True)
            return 1

    # ---- one proof across the N GPUs: the library's sharded path over NCCL
    sharded = None
    if world > 1 and not args.no_sharded:
        # the part after the solver on ONE GPU first (the round-1 scope: wires given), as the base of the post-solver speed-up
        cs_b = prog.r1cs()
        wires_dev = dev_buf(torch, sh["W"] * 32)
        prog.solve_device(wl.inputs, wires_dev, pk)
        wstep = lambda inputs: pk.prove_wires(cs_b, wires_dev, r, s)
        assert wstep(None) == proof
        post1_ms, _, _, _ = run(args.steps, lambda w: None, wstep)
        pk.close()                                        # the whole key leaves HBM; the generator makes it again and this rank keeps its chunks
        torch.cuda.empty_cache()
        uid = [zk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        pk_s, arrays_s, _ = build_key(torch, zk, ctx, sh, shard=True)
        del arrays_s
        torch.cuda.empty_cache()
        sstep = lambda inputs: pk_s.prove_solve(prog, inputs, r, s)
        for _ in range(max(1, args.warmup)):
            sproof = sstep(wl.inputs)
        c0 = ctx.comm_info()
        sms, swall, sproof, sstages = run(args.steps, lambda w: w.inputs, sstep)
        c1 = ctx.comm_info()
        wsstep = lambda inputs: pk_s.prove_wires(cs_b, wires_dev, r, s)
        assert wsstep(None) == proof
        postn_ms, _, _, _ = run(args.steps, lambda w: None, wsstep)
        del wires_dev
        same = torch.tensor([1 if sproof == proof else 0], device="cuda"); dist.all_reduce(same, op=dist.ReduceOp.MIN)
        st = {k: float(np.mean([x.get(k, 0.0) for x in sstages])) for k in sstages[0]}
        one_ms = solo_ms
        sharded = {"ms_per_proof": sms / args.steps, "speedup_vs_1": one_ms / (sms / args.steps), "one_gpu_ms_per_proof": one_ms,
                   "post_solver": {"one_gpu_ms": post1_ms / args.steps, "sharded_ms": postn_ms / args.steps, "speedup": post1_ms / postn_ms,
                                   "what": "zkpor_groth16_prove_wires: the proof from a solved wire vector (constraint evaluation, computeH, the five multiplications); "
                                           "with the solver in front, one proof's latency is bounded below by the serial sponge chain of the circuit (solver_tail)"},
                   "solve_head_ms": st.get("solve", 0.0), "ntt_ms": st.get("ntt", 0.0),
                   "proof_identical_to_one_gpu_on_every_rank": bool(same.item()),
                   "split": "solver replicated on every rank (latency chain); key split by point chunk (wires in N ranges, Z and the commitment basis in N chunks); "
                            "computeH as a four-step transform",
                   "collective": f"NCCL: 7 all-to-all exchanges of n/N^2 elements per peer (computeH), all-gather of 256 B mid-solve (commitment) and of 768 B (partial sums)",
                   "nvlink_bytes_received_per_rank_per_proof": (c1["all_to_all_bytes"] - c0["all_to_all_bytes"]) / args.steps,
                   "shard": pk_s.shard_info(), "wall_ms_per_proof": swall / args.steps * 1e3}
        if not same.item():
            print(json.dumps({"error": "the sharded proof differs from the one-GPU proof", "sharded": sharded}), flush=True)
            return 1
        pk_s.close()

    # ---- roofline of the dominant kernel (G1 bucket accumulation)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    ks = kstats["accumulate_g1"]
    roof, modmul = None, None
    if ks["launches"]:
        per_launch_ms = ks["total_ms"] / ks["launches"]
        terms_per_launch = ks["units"] / ks["launches"]
        achieved = 96.0 * terms_per_launch / (per_launch_ms / 1e3) / 1e9
        traffic = ncu_traffic_per_term("k_accumulate")
        roof = {"bound": "hbm", "kernel": "k_accumulate<Fp> (G1 bucket accumulation)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "traffic": traffic["bytes_per_term"] * terms_per_launch if traffic else None, "traffic_source": traffic["source"] if traffic else "no ncu capture on record",
                "bytes_per_term": 96, "terms_per_launch": terms_per_launch,
                "launch_ms": per_launch_ms, "launches": ks["launches"], "share_of_step": ks["total_ms"] / ms,
                "note": "integer-ALU bound: windows x 10 field products per term; see modmul_roofline"}
        # field products: every term is added into one bucket per window (XYZZ mixed addition = 8M + 2S = 10 products)
        nwin = -(-254 // 20) if args.log_n >= 24 else None
        if nwin:
            gps = terms_per_launch * nwin * 10 / (per_launch_ms / 1e3) / 1e9
            modmul = {"achieved_gps": gps, "peak_gps": IMAD_PEAK_GPS, "frac": gps / IMAD_PEAK_GPS, "unit": "1e9 field products/s",
                      "peak_source": "tools/pipe_probe.cu on B200: 27.1 IMAD.WIDE lane-ops/clk/SM sustained x 148 SMs x 1.965 GHz / 128 per product (profiles/r01_pipe_probe.txt)"}
    breakdown = {k: {"ms_per_step": v["total_ms"] / args.steps, "launches_per_step": v["launches"] / args.steps} for k, v in kstats.items()}

    line = {"metric": "proofs/hour", "value": value, "unit": "proofs/hour", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": workload, "parallelism": f"{P} provers per GPU (own context, key and program each; independent batches from {P} host threads)" + ("" if world == 1 else f" x {world} GPUs, no data-path collective"),
                       "proofs_per_step": proofs_per_step, "provers_per_gpu": P, "hbm_in_use_gb_after_setup": hbm_gb, "hbm_in_use_gb_after_timed_run": hbm_peak_gb, "solver_schedule": prog.stats(),
                       "l2": "inputs (>= 2 GB per vector, 21 GB key, ~6 GB constraint system) are far larger than the 126 MB L2; no explicit flush needed",
                       "key": "synthetic key in HBM: points (k0 + i*d)*G per array", "timing": "CUDA events on the library stream, max over ranks",
                       "setup_s": wl.setup_s},
            "solve_ms": solve_ms, "one_proof_alone_ms": solo_ms, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "modmul_roofline": modmul, "kernel_breakdown": breakdown,
            "stage_ms": {k: float(np.mean([st.get(k, 0.0) for st in stages])) for k in ("h2d", "solve", "ntt") if stages and k in stages[0]},
            "wall_ms_per_step": wall / args.steps * 1e3, "proof_sha": __import__("hashlib").sha256(proof).hexdigest()[:16]}
    if e2e:
        line["e2e"] = e2e
    if parity:
        line["parity"] = parity
    if sharded:
        line["sharded"] = sharded
    if rank == 0 and world == 1 and not args.no_cpu:
        for c, w in provers:
            w.close()
        del wl, pk, prog, provers
        torch.cuda.empty_cache()
        res = cpu_full_prove(torch, zk, ctx, args.cpu_log_n)
        scale = sh["n_constraints"] / res["n_constraints"]
        line["cpu_baseline"] = {"value": 3600.0 / (res["seconds"] * scale), "unit": "proofs/hour", "cores": res["threads"], "host_cpus": res["host_cpus"], "kind": "port",
                                "sample": f"oracle CPU prover (solver + proof) on the same circuit family at 2^{args.cpu_log_n} ({res['n_constraints']} constraints): "
                                          f"{res['seconds']:.2f} s (solver {res['solve_s']:.2f} s), scaled x{scale:.1f} by the constraint ratio; "
                                          f"`bench.py --impl reference` measures the full 2^{args.log_n} batch unscaled",
                                "parity_with_gpu": res["parity"]}
    if dist is not None:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


IMAD_PEAK_GPS = 27.1 * 148 * 1.965 / 128          # measured ceiling of the integer pipe in field products


def ncu_traffic_per_term(kernel_prefix):
    """dram bytes per MSM term of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json: written by
    tools/ncu_traffic.py from an `ncu --set full` run of this bench; keyed by kernel name + the hash of csrc/msm.cu it was built from).
    A capture taken from other sources is reported as stale instead of silently reused."""
    import hashlib
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        k = [x for x in rec["kernels"] if x["kernel"].startswith(kernel_prefix) and x.get("field") == "Fp"][0]
        src = hashlib.sha256(open(os.path.join(ROOT, "zkmerkle-proof-of-solvency_b200", "csrc", "msm.cu"), "rb").read()).hexdigest()[:12]
        stale = "" if k.get("msm_cu_sha") == src else " -- STALE: msm.cu changed since the capture"
        return {"bytes_per_term": k["dram_bytes"] / k["terms"], "source": f"{k['source']}{stale}"}
    except Exception:
        return None


if __name__ == "__main__":
    sys.exit(main())
