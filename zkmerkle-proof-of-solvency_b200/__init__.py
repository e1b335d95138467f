"""zkpor_b200 -- thin ctypes binding of libzkpor_b200.so (CUDA sm_100a kernels behind the C-ABI in
include/zkpor_b200.h) plus a host-side mirror of the reference interfaces this library replaces:

  reference (Go)                                              here
  ----------------------------------------------------------  -------------------------------------------
  groth16.Prove(cs, pk, witness)      prover.go:269           ProvingKey(...).prove(wires, a, b, c, r, s)
  pedersen ProvingKey.Commit          (inside Prove)          ProvingKey.commit(values)
  G1Jac/G2Jac.MultiExp                (inside Prove)          Context.msm_g1 / msm_g2
  fft.Domain.FFT/FFTInverse, computeH (inside Prove)          Context.ntt / compute_h
  poseidon.PoseidonBytes / Poseidon   utils.go:748            Context.poseidon_bytes / poseidon_hash_batch
  utils.AccountInfoToHash             utils.go:744-750        Context.account_leaves (+ padding_account_assets)
  merkletree.FixedDepthMerkleTree     merkletree.go:27-355    FixedDepthMerkleTree(ctx, depth, nil, capacity)

This module never computes on the CPU: there is no fallback.  If the shared library or a CUDA device is missing it
raises -- loudly -- instead of returning anything.  It does not import anything under oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkpor_b200.so")

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ZKPOR_SCALARS_MONT, ZKPOR_SCALARS_PLAIN = 0, 1
PROVE_PARTIAL_BYTES = 6 * 128 + 256
ASSET_TIERS = (50, 500)          # src/utils/constants.go:103-106
ACCOUNT_TREE_DEPTH = 28          # src/utils/constants.go:18


class ZkporError(RuntimeError):
    pass


def build_library(force: bool = False):
    """Compile the CUDA extension for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-j8"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZkporError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(the CUDA extension is mandatory; there is no CPU path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.zkpor_last_error.restype = C.c_char_p
        _lib.zkpor_version.restype = C.c_char_p
        _lib.zkpor_stage_name.restype = C.c_char_p
    return _lib


def _check(rc: int):
    if rc != 0:
        raise ZkporError(f"zkpor error {rc}: {lib().zkpor_last_error().decode()}")


def _ptr(x):
    """numpy array -> host pointer; int -> raw (device) pointer; torch tensor -> data_ptr()"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


def device_count() -> int:
    n = C.c_int32(0)
    _check(lib().zkpor_device_count(C.byref(n)))
    return n.value


# ----------------------------------------------------------------------------------------------- host logic mirrors
def assets_count_tier(n_assets: int) -> int:
    """utils.GetAssetsCountOfUser (src/utils/utils.go:128-145)"""
    for t in ASSET_TIERS:
        if n_assets <= t:
            return t
    raise ValueError("the target counts is less than the length of assets")


def padding_account_assets(assets) -> np.ndarray:
    """utils.PaddingAccountAssets (src/utils/utils.go:147-186).  assets: rows (index, equity, debt, loan, margin,
    portfolio_margin) with strictly increasing index -> flat uint64[tier*6]; gaps take the lowest unused indices."""
    target = assets_count_tier(len(assets))
    flat = np.zeros(target * 6, dtype=np.uint64)
    padding = target - len(assets)
    cur_pad, cur_idx, index = 0, 0, 0
    for a in assets:
        if cur_pad < padding:
            for j in range(cur_idx, int(a[0])):
                cur_pad += 1
                flat[index * 6] = j
                index += 1
                if cur_pad >= padding:
                    break
        flat[index * 6:index * 6 + 6] = a
        index += 1
        cur_idx = int(a[0]) + 1
    for i in range(index, target):
        flat[i * 6] = cur_idx
        cur_idx += 1
    return flat


def poseidon_constants(t: int):
    """(round constants, MDS rows, partial rounds) of the width-t permutation as canonical ints, from the library's own generator"""
    rc = np.zeros(((8 + 70) * t, 4), dtype=np.uint64)
    mds = np.zeros((t * t, 4), dtype=np.uint64)
    rp = C.c_uint32(0)
    _check(lib().zkpor_poseidon_constants(C.c_uint32(t), _ptr(rc), _ptr(mds), C.byref(rp)))
    r_inv = pow(1 << 256, -1, R_MOD)
    unmont = lambda row: (int(row[0]) | int(row[1]) << 64 | int(row[2]) << 128 | int(row[3]) << 192) * r_inv % R_MOD
    rcs = [unmont(rc[i]) for i in range((8 + rp.value) * t)]
    m = [unmont(x) for x in mds]
    return rcs, [m[i * t:(i + 1) * t] for i in range(t)], rp.value


def be32(v: int) -> bytes:
    return int(v).to_bytes(32, "big")


# ----------------------------------------------------------------------------------------------- context
class Context:
    """One GPU, one stream, NOT re-entrant (mirrors: one proof in flight per prover process, prover.go:141-247)."""

    def __init__(self, device: int = 0, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = C.c_void_p(_handle)
        else:
            _check(lib().zkpor_ctx_create(C.c_int32(device), C.byref(self._h)))
        self.device = device

    # --- one proof across N GPUs (include/zkpor_b200.h "one proof across the N GPUs of a box")
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """join an NCCL group (one process per GPU); unique_id = comm_unique_id() of rank 0, passed to every rank"""
        assert len(unique_id) == 128
        idb = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().zkpor_ctx_comm_init(self._h, idb, C.c_int32(rank), C.c_int32(world)))

    def comm_info(self) -> dict:
        r, w = C.c_int32(0), C.c_int32(1)
        st = (C.c_uint64 * 2)()
        _check(lib().zkpor_ctx_comm_info(self._h, C.byref(r), C.byref(w), st))
        return dict(rank=r.value, world=w.value, all_to_all_calls=int(st[0]), all_to_all_bytes=int(st[1]))

    def compute_h_sharded(self, a, b, c, log_n: int, out=None):
        """collective: a, b, c = this rank's cyclic rows; returns this rank's contiguous chunk of h"""
        w = self.comm_info()["world"]
        if out is None:
            out = np.zeros(((1 << log_n) // w, 4), dtype=np.uint64)
        _check(lib().zkpor_compute_h_sharded(self._h, _ptr(a), _ptr(b), _ptr(c), C.c_uint32(log_n), _ptr(out)))
        return out

    def close(self):
        if self._h:
            lib().zkpor_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- bookkeeping
    def sync(self):
        _check(lib().zkpor_ctx_sync(self._h))

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        _check(lib().zkpor_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def stream(self) -> int:
        s = C.c_void_p()
        _check(lib().zkpor_ctx_stream(self._h, C.byref(s)))
        return s.value or 0

    def last_timings(self) -> dict:
        buf = (C.c_float * 16)()
        n = C.c_int32(0)
        _check(lib().zkpor_ctx_last_timings(self._h, buf, 16, C.byref(n)))
        return {lib().zkpor_stage_name(i).decode(): float(buf[i]) for i in range(n.value)}

    def kernel_timing(self, enable: bool, classes=None):
        """classes: kernel class ids to time (None = all); an event pair per launch is not free"""
        mask = 0 if classes is None else sum(1 << k for k in classes)
        _check(lib().zkpor_ctx_kernel_timing(self._h, C.c_int32((1 | (mask << 8)) if enable else 0)))

    def kernel_stats(self, klass: int):
        ms, n, u = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().zkpor_ctx_kernel_stats(self._h, C.c_int32(klass), C.byref(ms), C.byref(n), C.byref(u)))
        return dict(total_ms=ms.value, launches=n.value, units=u.value)

    # --- MSM
    def set_affine_rounds(self, rounds: int):
        """Levels of batched-affine tree summation per bucket (0 = extended-Jacobian only, the default; -1 automatic; k > 0 at most k)."""
        _check(lib().zkpor_msm_set_affine_rounds(self._h, C.c_int32(rounds)))

    def msm_g1(self, points, scalars, n: int, flags: int = ZKPOR_SCALARS_MONT) -> np.ndarray:
        out = np.zeros(8, dtype=np.uint64)
        _check(lib().zkpor_msm_g1(self._h, _ptr(points), _ptr(scalars), C.c_uint64(n), C.c_uint32(flags), _ptr(out)))
        return out

    def msm_g2(self, points, scalars, n: int, flags: int = ZKPOR_SCALARS_MONT) -> np.ndarray:
        out = np.zeros(16, dtype=np.uint64)
        _check(lib().zkpor_msm_g2(self._h, _ptr(points), _ptr(scalars), C.c_uint64(n), C.c_uint32(flags), _ptr(out)))
        return out

    def msm_g1_partial(self, points, scalars, n: int, flags: int = ZKPOR_SCALARS_MONT) -> np.ndarray:
        out = np.zeros(16, dtype=np.uint64)
        _check(lib().zkpor_msm_g1_partial(self._h, _ptr(points), _ptr(scalars), C.c_uint64(n), C.c_uint32(flags), _ptr(out)))
        return out

    def msm_g2_partial(self, points, scalars, n: int, flags: int = ZKPOR_SCALARS_MONT) -> np.ndarray:
        out = np.zeros(32, dtype=np.uint64)
        _check(lib().zkpor_msm_g2_partial(self._h, _ptr(points), _ptr(scalars), C.c_uint64(n), C.c_uint32(flags), _ptr(out)))
        return out

    # --- NTT
    def ntt(self, data, log_n: int, inverse: bool, dit: bool, coset: bool):
        """in place; data = numpy (host) or device pointer / tensor"""
        _check(lib().zkpor_ntt(self._h, _ptr(data), C.c_uint32(log_n), C.c_int32(inverse), C.c_int32(dit), C.c_int32(coset)))
        return data

    def fr_mul(self, a, b, out, n: int):
        _check(lib().zkpor_fr_mul(self._h, _ptr(a), _ptr(b), _ptr(out), C.c_uint64(n)))
        return out

    def compute_h(self, a, b, c, n_constraints: int, log_n: int, out=None):
        if out is None:
            out = np.zeros((1 << log_n, 4), dtype=np.uint64)
        _check(lib().zkpor_compute_h(self._h, _ptr(a), _ptr(b), _ptr(c), C.c_uint64(n_constraints), C.c_uint32(log_n), _ptr(out)))
        return out

    # --- point decoding (pk.UnsafeReadFrom)
    def g1_decode_batch(self, in_bytes, n: int, compressed: bool = True, out=None):
        if out is None:
            out = np.zeros((n, 8), dtype=np.uint64)
        _check(lib().zkpor_g1_decode_batch(self._h, _ptr(in_bytes), C.c_uint64(n), C.c_int32(compressed), _ptr(out)))
        return out

    def g2_decode_batch(self, in_bytes, n: int, compressed: bool = True, out=None):
        if out is None:
            out = np.zeros((n, 16), dtype=np.uint64)
        _check(lib().zkpor_g2_decode_batch(self._h, _ptr(in_bytes), C.c_uint64(n), C.c_int32(compressed), _ptr(out)))
        return out

    # --- Poseidon
    def set_poseidon_out_lane(self, lane: int):
        _check(lib().zkpor_poseidon_set_out_lane(self._h, C.c_int32(lane)))

    def poseidon_hash_batch(self, in_be, n_in: int, count: int, out=None):
        if out is None:
            out = np.zeros((count, 32), dtype=np.uint8)
        _check(lib().zkpor_poseidon_hash_batch(self._h, _ptr(in_be), C.c_uint32(n_in), C.c_uint64(count), _ptr(out)))
        return out

    def poseidon_bytes(self, *chunks: bytes) -> bytes:
        """poseidon.PoseidonBytes(...[]byte): every chunk is one big-endian element (empty = 0), must be < r."""
        vals = [int.from_bytes(c, "big") for c in chunks]
        if any(v >= R_MOD for v in vals):
            raise ValueError("not support bytes bigger than modulus")
        buf = np.frombuffer(b"".join(be32(v) for v in vals), dtype=np.uint8).copy()
        return self.poseidon_hash_batch(buf, len(vals), 1).tobytes()

    def account_leaves(self, ids_be, totals_be, flat_assets, n: int, tier: int, out=None):
        if out is None:
            out = np.zeros((n, 32), dtype=np.uint8)
        _check(lib().zkpor_account_leaves(self._h, _ptr(ids_be), _ptr(totals_be), _ptr(flat_assets), C.c_uint64(n), C.c_uint32(tier), _ptr(out)))
        return out


def _be_arr(v: int):
    return (C.c_uint8 * 32).from_buffer_copy(be32(v % R_MOD))


def synth_points_g1(ctx: Context, k0: int, d: int, n: int, out_dev):
    """out_dev[i] = (k0 + i*d) * G1, affine Montgomery, written to device memory"""
    _check(lib().zkpor_synth_points_g1(ctx._h, _be_arr(k0), _be_arr(d), C.c_uint64(n), _ptr(out_dev)))


def synth_points_g2(ctx: Context, k0: int, d: int, n: int, out_dev):
    _check(lib().zkpor_synth_points_g2(ctx._h, _be_arr(k0), _be_arr(d), C.c_uint64(n), _ptr(out_dev)))


def synth_scalars(ctx: Context, seed: int, n: int, kind: int, out_dev):
    _check(lib().zkpor_synth_scalars(ctx._h, C.c_uint64(seed), C.c_uint64(n), C.c_int32(kind), _ptr(out_dev)))


def g1_sum_partials(partials: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(partials, dtype=np.uint64).reshape(-1, 16)
    out = np.zeros(8, dtype=np.uint64)
    _check(lib().zkpor_g1_sum_partials(_ptr(p), C.c_uint32(p.shape[0]), _ptr(out)))
    return out


def g2_sum_partials(partials: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(partials, dtype=np.uint64).reshape(-1, 32)
    out = np.zeros(16, dtype=np.uint64)
    _check(lib().zkpor_g2_sum_partials(_ptr(p), C.c_uint32(p.shape[0]), _ptr(out)))
    return out


class PoseidonHasher:
    """hash.Hash as returned by poseidon.NewPoseidon(): Write appends ONE element per call, Sum hashes and clears."""

    def __init__(self, ctx: Context):
        self.ctx, self.data = ctx, []

    def reset(self):
        self.data = []

    def write(self, p: bytes) -> int:
        if int.from_bytes(p, "big") >= R_MOD:
            raise ValueError("not support bytes bigger than modulus")
        self.data.append(bytes(p))
        return len(p)

    def sum(self, prefix: bytes = b"") -> bytes:
        out = self.ctx.poseidon_bytes(*self.data)
        self.data = []
        return prefix + out


# ----------------------------------------------------------------------------------------------- Merkle tree
class FixedDepthMerkleTree:
    """merkletree.FixedDepthMerkleTree with the tree resident in HBM (src/utils/merkletree/merkletree.go)."""

    def __init__(self, ctx: Context, depth: int, nil_leaf: bytes, capacity: int):
        if depth > 32:
            raise ValueError("depth too large")
        if depth <= 0:
            raise ValueError("depth must be positive")
        if capacity > (1 << depth):
            raise ValueError("capacity exceeds maximum for given depth")
        self.ctx, self.depth, self.capacity = ctx, depth, capacity
        self._h = C.c_void_p()
        nil = (C.c_uint8 * 32).from_buffer_copy(nil_leaf)
        _check(lib().zkpor_tree_create(ctx._h, C.c_uint32(depth), nil, C.c_uint64(capacity), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().zkpor_tree_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, key: int, value: bytes):
        if key >= self.capacity:
            raise IndexError(f"key {key} out of range for capacity {self.capacity}")
        self.set_range(key, np.frombuffer(value, dtype=np.uint8).copy(), 1)

    def set_range(self, first_key: int, leaves_be, count: int):
        _check(lib().zkpor_tree_set_range(self.ctx._h, self._h, C.c_uint64(first_key), C.c_uint64(count), _ptr(leaves_be)))

    def set_keys(self, keys, leaves_be):
        k = np.ascontiguousarray(keys, dtype=np.uint32)
        _check(lib().zkpor_tree_set_keys(self.ctx._h, self._h, _ptr(k), C.c_uint64(k.size), _ptr(leaves_be)))

    def build(self):
        _check(lib().zkpor_tree_build(self.ctx._h, self._h))

    def root(self) -> bytes:
        out = (C.c_uint8 * 32)()
        _check(lib().zkpor_tree_root(self.ctx._h, self._h, out))
        return bytes(out)

    def get(self, key: int) -> bytes:
        return self.get_leaves([key])[0].tobytes()

    def get_leaves(self, keys) -> np.ndarray:
        k = np.ascontiguousarray(keys, dtype=np.uint32)
        out = np.zeros((k.size, 32), dtype=np.uint8)
        _check(lib().zkpor_tree_get_leaves(self.ctx._h, self._h, _ptr(k), C.c_uint64(k.size), _ptr(out)))
        return out

    def get_proof(self, key: int):
        if key >= (1 << self.depth):
            raise IndexError(f"key {key} out of range for tree depth {self.depth}")
        return [bytes(x) for x in self.get_proofs([key])[0]]

    def get_proofs(self, keys) -> np.ndarray:
        k = np.ascontiguousarray(keys, dtype=np.uint32)
        out = np.zeros((k.size, self.depth, 32), dtype=np.uint8)
        _check(lib().zkpor_tree_get_proofs(self.ctx._h, self._h, _ptr(k), C.c_uint64(k.size), _ptr(out)))
        return out

    def shard_range(self):
        """(first key, number of keys, subtree level) of this context's rank when the tree is built across a group of GPUs"""
        first, count, lvl = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        _check(lib().zkpor_tree_shard_range(self.ctx._h, self._h, C.byref(first), C.byref(count), C.byref(lvl)))
        return first.value, count.value, lvl.value

    def build_sharded(self):
        """collective Build over the context's group: subtrees per rank, all-gather of the subtree roots, top levels everywhere"""
        _check(lib().zkpor_tree_build_sharded(self.ctx._h, self._h))

    def level(self, level: int):
        p, n = C.c_void_p(), C.c_uint64(0)
        _check(lib().zkpor_tree_level(self.ctx._h, self._h, C.c_uint32(level), C.byref(p), C.byref(n)))
        return p.value, n.value


class CexDesc(C.Structure):
    _fields_ = [("n_assets", C.c_uint32), ("base_prices", C.c_void_p), ("tier_ratio_elems", C.c_void_p), ("initial_totals", C.c_void_p)]


def witness_batches(ctx: Context, *, base_prices, tier_ratio_elems, initial_totals, root: bytes, flat_assets, account_indices, tier: int, ops_per_batch: int):
    """The witness service's main loop (witness.go:144-206) for all batches of one tier: returns (totals (nb+1, n_assets, 5) uint64,
    cex_commitments (nb+1, 32) uint8, batch_commitments (nb, 32) uint8).  base_prices (n_assets,) uint64; tier_ratio_elems
    (n_assets, 18, 32) uint8 big-endian; initial_totals (n_assets, 5) uint64; flat_assets (n_accounts, tier*6) uint64."""
    bp = np.ascontiguousarray(base_prices, dtype=np.uint64); te = np.ascontiguousarray(tier_ratio_elems, dtype=np.uint8); it = np.ascontiguousarray(initial_totals, dtype=np.uint64)
    idx = np.ascontiguousarray(account_indices, dtype=np.uint32)
    n_assets, n = bp.size, idx.size
    nb = n // ops_per_batch
    d = CexDesc(n_assets, _ptr(bp), _ptr(te), _ptr(it))
    totals = np.zeros((nb + 1, n_assets, 5), dtype=np.uint64); cm = np.zeros((nb + 1, 32), dtype=np.uint8); bc = np.zeros((nb, 32), dtype=np.uint8)
    rt = (C.c_uint8 * 32).from_buffer_copy(root)
    _check(lib().zkpor_witness_batches(ctx._h, C.byref(d), rt, _ptr(flat_assets), _ptr(idx), C.c_uint64(n), C.c_uint32(tier), C.c_uint32(ops_per_batch),
                                       _ptr(totals), _ptr(cm), _ptr(bc)))
    return totals, cm, bc


def verify_proof(ctx: Context, root: bytes, key: int, proof, leaf: bytes, depth: int) -> bool:
    """merkletree.VerifyProof (merkletree.go:334-355), hashing on the GPU one level at a time."""
    if len(proof) != depth or key >= (1 << depth):
        return False
    node = leaf
    for i in range(depth):
        node = ctx.poseidon_bytes(node, proof[i]) if key & (1 << i) == 0 else ctx.poseidon_bytes(proof[i], node)
    return node == root


# ----------------------------------------------------------------------------------------------- Groth16
class PkDesc(C.Structure):
    _fields_ = [("log_n", C.c_uint32), ("n_wires", C.c_uint64), ("n_public", C.c_uint64),
                ("n_a", C.c_uint64), ("n_b", C.c_uint64), ("n_k", C.c_uint64), ("n_z", C.c_uint64),
                ("g1_a", C.c_void_p), ("g1_b", C.c_void_p), ("g1_k", C.c_void_p), ("g1_z", C.c_void_p), ("g2_b", C.c_void_p),
                ("g1_alpha", C.c_void_p), ("g1_beta", C.c_void_p), ("g1_delta", C.c_void_p), ("g2_beta", C.c_void_p), ("g2_delta", C.c_void_p),
                ("infinity_a", C.c_void_p), ("infinity_b", C.c_void_p),
                ("n_committed", C.c_uint64), ("ck_basis", C.c_void_p), ("ck_basis_exp_sigma", C.c_void_p),
                ("private_committed", C.c_void_p), ("commitment_index", C.c_uint64)]


class PkCsInfo(C.Structure):
    _fields_ = [("n_public", C.c_uint64), ("private_committed", C.c_void_p), ("n_committed", C.c_uint64), ("commitment_index", C.c_uint64)]


class VkHost(C.Structure):
    _fields_ = [("g1_alpha", C.c_uint8 * 64), ("g1_beta", C.c_uint8 * 64), ("g1_delta", C.c_uint8 * 64),
                ("g2_beta", C.c_uint8 * 128), ("g2_gamma", C.c_uint8 * 128), ("g2_delta", C.c_uint8 * 128),
                ("g2_ped_g", C.c_uint8 * 128), ("g2_ped_g_root_sigma_neg", C.c_uint8 * 128),
                ("n_k", C.c_uint64), ("n_commitments", C.c_uint64), ("n_public_committed", C.c_uint64)]


_VK_POINTS = ("g1_alpha", "g1_beta", "g1_delta", "g2_beta", "g2_gamma", "g2_delta", "g2_ped_g", "g2_ped_g_root_sigma_neg")


def proof_decode(ctx: Context, data: bytes) -> bytes:
    """groth16.Proof.ReadFrom (verifier/main.go:208-216): compressed or raw bytes -> the raw layout the verify calls take"""
    src = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(260 + 64 * 17, dtype=np.uint8)
    n, used = C.c_uint32(out.size), C.c_uint64(0)
    _check(lib().zkpor_proof_decode(ctx._h, _ptr(src), C.c_uint64(src.size), _ptr(out), C.byref(n), C.byref(used)))
    return out[:n.value].tobytes()


def proof_encode(ctx: Context, data: bytes, compressed: bool = True) -> bytes:
    """Proof.WriteTo (compressed) / WriteRawTo from either form"""
    src = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(260 + 64 * 17, dtype=np.uint8)
    n = C.c_uint32(out.size)
    _check(lib().zkpor_proof_encode(ctx._h, _ptr(src), C.c_uint32(src.size), C.c_int32(1 if compressed else 0), _ptr(out), C.byref(n)))
    return out[:n.value].tobytes()


def vk_decode(ctx: Context, data: bytes) -> dict:
    """vk.ReadFrom (prover.go:358-362, verifier/main.go:33-34) -> the dict the verify calls take (points as uint64 arrays)"""
    src = np.frombuffer(data, dtype=np.uint8)
    vk = VkHost()
    k = np.zeros((64, 8), dtype=np.uint64)
    pc = np.zeros(64, dtype=np.uint64)
    used = C.c_uint64(0)
    _check(lib().zkpor_vk_decode(ctx._h, _ptr(src), C.c_uint64(src.size), C.byref(vk), _ptr(k), C.c_uint64(k.shape[0]), _ptr(pc), C.c_uint64(pc.size), C.byref(used)))
    out = {nm: np.frombuffer(bytes(getattr(vk, nm)), dtype=np.uint64).copy() for nm in _VK_POINTS}
    out.update(g1_k=k[:vk.n_k].copy(), n_commitments=int(vk.n_commitments), public_committed=pc[:vk.n_public_committed].copy(), bytes_consumed=used.value)
    return out


def vk_encode(ctx: Context, vk: dict, raw: bool = False) -> bytes:
    """vk.WriteTo (keygen/main.go:46-62) / WriteRawTo"""
    h = VkHost()
    for nm in _VK_POINTS:
        if nm in vk and vk[nm] is not None:
            b = np.ascontiguousarray(vk[nm], dtype=np.uint64).tobytes()
            C.memmove(getattr(h, nm), b, len(b))
    k = np.ascontiguousarray(vk["g1_k"], dtype=np.uint64).reshape(-1, 8)
    pc = np.ascontiguousarray(vk.get("public_committed", []), dtype=np.uint64)
    h.n_k, h.n_commitments, h.n_public_committed = k.shape[0], int(vk.get("n_commitments", 0)), pc.size
    n = C.c_uint64(0)
    args = (ctx._h, C.byref(h), _ptr(k), _ptr(pc) if pc.size else None, C.c_int32(1 if raw else 0))
    _check(lib().zkpor_vk_encode(*args, None, C.c_uint64(0), C.byref(n)))
    out = np.zeros(n.value, dtype=np.uint8)
    _check(lib().zkpor_vk_encode(*args, _ptr(out), C.c_uint64(out.size), C.byref(n)))
    return out.tobytes()


class ProvingKey:
    """groth16.ProvingKey resident in HBM (what pk.UnsafeReadFrom fills at prover.go:342-346).  Arrays are numpy
    uint64 in gnark memory layout, or device pointers / tensors for the point arrays."""

    def __init__(self, ctx: Context, *, log_n, A, B1, K, Z, B2, alpha1, beta1, delta1, beta2, delta2, n_a, n_b, n_k, n_z,
                 infinity_a=None, infinity_b=None, n_public=0, ck_basis=None, ck_basis_exp_sigma=None, private_committed=None,
                 commitment_index=0, shard=False):
        """shard=True: the arguments describe the WHOLE key; only the chunks of ctx's rank in its group are uploaded
        (zkpor_pk_upload_shard) and the prove calls become collective."""
        self.ctx = ctx
        d = PkDesc()
        keep = []

        def hp(x, dtype):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                x = np.ascontiguousarray(x, dtype=dtype)
            keep.append(x)
            return _ptr(x)

        d.log_n = log_n
        d.n_wires = 0 if infinity_a is None else len(infinity_a)
        d.n_public = n_public
        d.n_a, d.n_b, d.n_k, d.n_z = n_a, n_b, n_k, n_z
        d.g1_a, d.g1_b, d.g1_k, d.g1_z, d.g2_b = (hp(x, np.uint64) for x in (A, B1, K, Z, B2))
        d.g1_alpha, d.g1_beta, d.g1_delta, d.g2_beta, d.g2_delta = (hp(x, np.uint64) for x in (alpha1, beta1, delta1, beta2, delta2))
        d.infinity_a = hp(None if infinity_a is None else np.asarray(infinity_a, dtype=np.uint8), np.uint8)
        d.infinity_b = hp(None if infinity_b is None else np.asarray(infinity_b, dtype=np.uint8), np.uint8)
        n_ck = 0
        if ck_basis is not None:
            n_ck = len(private_committed) if private_committed is not None else (ck_basis.size // 8)
        d.n_committed = n_ck
        d.ck_basis, d.ck_basis_exp_sigma = hp(ck_basis, np.uint64), hp(ck_basis_exp_sigma, np.uint64)
        d.private_committed = hp(None if private_committed is None else np.asarray(private_committed, dtype=np.uint64), np.uint64)
        d.commitment_index = commitment_index
        self.points = dict(alpha1=np.array(alpha1, dtype=np.uint64), beta1=np.array(beta1, dtype=np.uint64), delta1=np.array(delta1, dtype=np.uint64),
                           beta2=np.array(beta2, dtype=np.uint64), delta2=np.array(delta2, dtype=np.uint64))
        self.has_commitment = ck_basis is not None
        self.log_n, self.n_z = log_n, n_z
        self._h = C.c_void_p()
        _check((lib().zkpor_pk_upload_shard if shard else lib().zkpor_pk_upload)(ctx._h, C.byref(d), C.byref(self._h)))

    @classmethod
    def read(cls, ctx: Context, file_bytes, n_public: int, private_committed=None, commitment_index: int = 0) -> "ProvingKey":
        """pk.ReadFrom / UnsafeReadFrom (prover.go:342-346): the key file's bytes (compressed or raw) decoded straight into HBM.
        n_public, private_committed and commitment_index come from the constraint system (zkpor_pk_cs_info)."""
        buf = np.frombuffer(file_bytes, dtype=np.uint8) if not isinstance(file_bytes, np.ndarray) else file_bytes
        pc = np.ascontiguousarray(private_committed if private_committed is not None else [], dtype=np.uint64)
        info = PkCsInfo(n_public, _ptr(pc) if pc.size else None, pc.size, commitment_index)
        self = cls.__new__(cls)
        self.ctx, self._h, self.points, self.has_commitment = ctx, C.c_void_p(), {}, pc.size > 0
        used = C.c_uint64(0)
        _check(lib().zkpor_pk_read(ctx._h, _ptr(buf), C.c_uint64(buf.size), C.byref(info), C.byref(self._h), C.byref(used)))
        self.bytes_consumed = used.value
        return self

    def write(self, raw: bool = False) -> bytes:
        """pk.WriteTo (compressed) / WriteRawTo from the resident key (keygen/main.go:46-62)"""
        n = C.c_uint64(0)
        _check(lib().zkpor_pk_write(self.ctx._h, self._h, C.c_int32(1 if raw else 0), None, C.c_uint64(0), C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint8)
        _check(lib().zkpor_pk_write(self.ctx._h, self._h, C.c_int32(1 if raw else 0), _ptr(out), C.c_uint64(out.size), C.byref(n)))
        return out.tobytes()

    def shard_info(self) -> dict:
        out = (C.c_uint64 * 8)()
        _check(lib().zkpor_pk_shard_info(self._h, out))
        return dict(zip(("rank", "world", "wire_first", "n_wires", "n_a", "n_b", "n_k", "n_z"), (int(x) for x in out)))

    def close(self):
        if self._h:
            lib().zkpor_pk_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def commit(self, committed_values) -> np.ndarray:
        out = np.zeros(8, dtype=np.uint64)
        _check(lib().zkpor_pk_commit(self.ctx._h, self._h, _ptr(committed_values), _ptr(out)))
        return out

    def prove(self, wires, a, b, c, n_constraints: int, r: int, s: int) -> bytes:
        """groth16.Prove after the solver: returns proof.WriteRawTo bytes."""
        out = np.zeros(388, dtype=np.uint8)
        n = C.c_uint32(0)
        rb = (C.c_uint8 * 32).from_buffer_copy(be32(r % R_MOD))
        sb = (C.c_uint8 * 32).from_buffer_copy(be32(s % R_MOD))
        _check(lib().zkpor_groth16_prove(self.ctx._h, self._h, _ptr(wires), _ptr(a), _ptr(b), _ptr(c), C.c_uint64(n_constraints), rb, sb,
                                         _ptr(out), C.byref(n)))
        return out[:n.value].tobytes()

    def prove_solve(self, prog: "Program", inputs, r: int, s: int) -> bytes:
        """the whole of groth16.Prove (prover.go:269): witness solver (hints, commitment mid-solve) + proof, from the circuit inputs"""
        out = np.zeros(388, dtype=np.uint8)
        n = C.c_uint32(0)
        rb = (C.c_uint8 * 32).from_buffer_copy(be32(r % R_MOD))
        sb = (C.c_uint8 * 32).from_buffer_copy(be32(s % R_MOD))
        _check(lib().zkpor_groth16_prove_solve(self.ctx._h, self._h, prog._h, _ptr(inputs), rb, sb, _ptr(out), C.byref(n)))
        return out[:n.value].tobytes()

    def prove_wires(self, cs: "R1CS", wires, r: int, s: int) -> bytes:
        """groth16.Prove from the wire vector alone: a, b, c = L w, R w, O w are evaluated on the device (R1CS resident in HBM)."""
        out = np.zeros(388, dtype=np.uint8)
        n = C.c_uint32(0)
        rb = (C.c_uint8 * 32).from_buffer_copy(be32(r % R_MOD))
        sb = (C.c_uint8 * 32).from_buffer_copy(be32(s % R_MOD))
        _check(lib().zkpor_groth16_prove_wires(self.ctx._h, self._h, cs._h, _ptr(wires), rb, sb, _ptr(out), C.byref(n)))
        return out[:n.value].tobytes()

    def prove_partial(self, wires_a, wires_b, wires_k, committed, h_chunk, n_h: int) -> np.ndarray:
        out = np.zeros(PROVE_PARTIAL_BYTES, dtype=np.uint8)
        _check(lib().zkpor_groth16_prove_partial(self.ctx._h, self._h, _ptr(wires_a), _ptr(wires_b), _ptr(wires_k), _ptr(committed),
                                                 _ptr(h_chunk), C.c_uint64(n_h), _ptr(out)))
        return out

    def finish(self, partials: np.ndarray, r: int, s: int) -> bytes:
        p = np.ascontiguousarray(partials, dtype=np.uint8).reshape(-1, PROVE_PARTIAL_BYTES)
        out = np.zeros(388, dtype=np.uint8)
        n = C.c_uint32(0)
        rb = (C.c_uint8 * 32).from_buffer_copy(be32(r % R_MOD))
        sb = (C.c_uint8 * 32).from_buffer_copy(be32(s % R_MOD))
        pts = self.points
        _check(lib().zkpor_groth16_finish(_ptr(p), C.c_uint32(p.shape[0]), _ptr(pts["alpha1"]), _ptr(pts["beta1"]), _ptr(pts["delta1"]),
                                          _ptr(pts["beta2"]), _ptr(pts["delta2"]), rb, sb, C.c_int32(self.has_commitment), _ptr(out), C.byref(n)))
        return out[:n.value].tobytes()


# ----------------------------------------------------------------------------------------------- Setup
def g1_fixed_base_batch(ctx: Context, base, scalars, n: int, out, flags: int = ZKPOR_SCALARS_MONT):
    """curve.BatchScalarMultiplicationG1: out[i] = scalars[i] * base"""
    _check(lib().zkpor_g1_fixed_base_batch(ctx._h, _ptr(base), _ptr(scalars), C.c_uint64(n), C.c_uint32(flags), _ptr(out)))
    return out


def g2_fixed_base_batch(ctx: Context, base, scalars, n: int, out, flags: int = ZKPOR_SCALARS_MONT):
    _check(lib().zkpor_g2_fixed_base_batch(ctx._h, _ptr(base), _ptr(scalars), C.c_uint64(n), C.c_uint32(flags), _ptr(out)))
    return out


def groth16_setup(ctx: Context, log_n: int, n_wires: int, nb_public: int, csc_a, csc_b, csc_c, private_committed, commitment_index: int, toxic: dict):
    """groth16.Setup (src/keygen/main.go:42; gnark backend/groth16/bn254/setup.go) with EXPLICIT toxic waste
    {alpha, beta, gamma, delta, tau, sigma} -- gnark draws it with crypto/rand.  csc_x = (col_ptr u64[n_wires+1],
    rows u32[nnz], coeffs Fr-Montgomery (nnz,4) u64) for the A / B / C matrices in column (= wire) order.
    Everything heavy runs on the GPU: Lagrange basis at tau, per-wire sums, the K / Z scalars, and the fixed-base
    batch multiplications.  Returns (pk_kwargs for ProvingKey(...), extras) with point arrays as CUDA tensors."""
    import torch
    n = 1 << log_n
    dev = lambda nb: torch.empty((nb + 7) // 8, dtype=torch.int64, device="cuda")
    be = lambda v: _be_arr(v)
    inv = lambda v: pow(v % R_MOD, -1, R_MOD)
    lag = dev(n * 32)
    _check(lib().zkpor_setup_lagrange(ctx._h, be(toxic["tau"]), C.c_uint32(log_n), _ptr(lag)))
    sums = []
    for col_ptr, rows, coeffs in (csc_a, csc_b, csc_c):
        col_ptr = np.ascontiguousarray(col_ptr, dtype=np.uint64); rows = np.ascontiguousarray(rows, dtype=np.uint32)
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
        out = dev(n_wires * 32)
        _check(lib().zkpor_setup_wire_sums(ctx._h, _ptr(col_ptr), _ptr(rows), _ptr(coeffs), C.c_uint64(rows.size), _ptr(lag), C.c_uint64(n_wires), _ptr(out)))
        sums.append(out)
    A, B, Cc = sums
    k_gamma, k_delta = dev(n_wires * 32), dev(n_wires * 32)
    for out, k in ((k_gamma, inv(toxic["gamma"])), (k_delta, inv(toxic["delta"]))):   # (beta*A + alpha*B + C) / {gamma, delta}
        _check(lib().zkpor_fr_lincomb3(ctx._h, _ptr(A), _ptr(B), _ptr(Cc), be(toxic["beta"]), be(toxic["alpha"]), be(1), be(k), C.c_uint64(n_wires), _ptr(out)))
    v4 = lambda t: t.view(-1, 4)
    committed = np.asarray(private_committed, dtype=np.int64)
    has_commit = commitment_index is not None and commitment_index >= 0
    is_vk = np.zeros(n_wires, dtype=bool); is_vk[:nb_public] = True
    if has_commit:
        is_vk[commitment_index] = True
    is_ck = np.zeros(n_wires, dtype=bool); is_ck[committed] = True
    idx = lambda mask: torch.from_numpy(np.nonzero(mask)[0]).cuda()
    vk_s = v4(k_gamma)[idx(is_vk)].contiguous(); ck_s = v4(k_gamma)[idx(is_ck)].contiguous(); pk_s = v4(k_delta)[idx(~is_vk & ~is_ck)].contiguous()
    ck_sigma_s = torch.empty_like(ck_s)
    if ck_s.shape[0]:
        _check(lib().zkpor_fr_lincomb3(ctx._h, _ptr(ck_s), _ptr(ck_s), _ptr(ck_s), be(1), be(0), be(0), be(toxic["sigma"]), C.c_uint64(ck_s.shape[0]), _ptr(ck_sigma_s)))
    zt = (pow(toxic["tau"], n, R_MOD) - 1) % R_MOD
    Z = dev(n * 32)
    _check(lib().zkpor_fr_powers(ctx._h, be(zt * inv(toxic["delta"]) % R_MOD), be(toxic["tau"]), C.c_uint64(n), C.c_uint32(log_n), C.c_int32(1), _ptr(Z)))
    Z_s = v4(Z)[: n - 1].contiguous()
    inf_a = (v4(A) == 0).all(dim=1); inf_b = (v4(B) == 0).all(dim=1)
    A_s = v4(A)[~inf_a].contiguous(); B_s = v4(B)[~inf_b].contiguous()
    g1 = dev(64); g2 = dev(128)
    synth_points_g1(ctx, 1, 1, 1, g1); synth_points_g2(ctx, 1, 1, 1, g2)

    def fb(scalars, g2_=False):
        cnt = scalars.shape[0]
        out = dev(max(cnt, 1) * (128 if g2_ else 64))
        if cnt:
            (g2_fixed_base_batch if g2_ else g1_fixed_base_batch)(ctx, g2 if g2_ else g1, scalars, cnt, out)
        return out

    def one(v, g2_=False):
        s = torch.from_numpy(np.frombuffer(b"".join(int((v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF).to_bytes(8, "little") for k in range(4)), dtype=np.int64).copy()).cuda()
        o = dev(128 if g2_ else 64)
        (g2_fixed_base_batch if g2_ else g1_fixed_base_batch)(ctx, g2 if g2_ else g1, s, 1, o, ZKPOR_SCALARS_PLAIN)
        return o.cpu().numpy().view(np.uint64).copy()

    pk_kwargs = dict(log_n=log_n, A=fb(A_s), B1=fb(B_s), K=fb(pk_s), Z=fb(Z_s), B2=fb(B_s, True),
                     alpha1=one(toxic["alpha"]), beta1=one(toxic["beta"]), delta1=one(toxic["delta"]), beta2=one(toxic["beta"], True), delta2=one(toxic["delta"], True),
                     n_a=int(A_s.shape[0]), n_b=int(B_s.shape[0]), n_k=int(pk_s.shape[0]), n_z=n - 1,
                     infinity_a=inf_a.cpu().numpy().astype(np.uint8), infinity_b=inf_b.cpu().numpy().astype(np.uint8), n_public=nb_public)
    if has_commit:
        pk_kwargs.update(ck_basis=fb(ck_s), ck_basis_exp_sigma=fb(ck_sigma_s), private_committed=committed.astype(np.uint64), commitment_index=commitment_index)
    extras = dict(vk_K=fb(vk_s), gamma2=one(toxic["gamma"], True), n_vk=int(vk_s.shape[0]))
    return pk_kwargs, extras


# ----------------------------------------------------------------------------------------------- constraint evaluation
class Csr(C.Structure):
    _fields_ = [("nnz", C.c_uint64), ("row_ptr", C.c_void_p), ("wire_ids", C.c_void_p), ("coeff_ids", C.c_void_p)]


class R1CS:
    """The three R1CS matrices resident in HBM (what r1cs.ReadFrom fills at prover.go:317-327, flattened to CSR by the shim):
    matrices = [(row_ptr u64[n+1], wire_ids u32[nnz], coeff_ids u32[nnz])] * 3 for L, R, O; coeff_table = (k, 4) u64 Montgomery."""

    def __init__(self, ctx: Context, n_constraints: int, n_wires: int, matrices, coeff_table):
        self.ctx = ctx
        self.n_constraints = n_constraints
        keep, descs = [], []
        for row_ptr, wire_ids, coeff_ids in matrices:
            rp = np.ascontiguousarray(row_ptr, dtype=np.uint64); wi = np.ascontiguousarray(wire_ids, dtype=np.uint32); ci = np.ascontiguousarray(coeff_ids, dtype=np.uint32)
            keep += [rp, wi, ci]
            descs.append(Csr(wi.size, _ptr(rp), _ptr(wi) if wi.size else None, _ptr(ci) if ci.size else None))
        tab = np.ascontiguousarray(coeff_table, dtype=np.uint64).reshape(-1, 4)
        self._h = C.c_void_p()
        _check(lib().zkpor_r1cs_upload(ctx._h, C.c_uint64(n_constraints), C.c_uint64(n_wires), C.byref(descs[0]), C.byref(descs[1]), C.byref(descs[2]),
                                       _ptr(tab), C.c_uint64(tab.shape[0]), C.byref(self._h)))

    def close(self):
        if self._h and not getattr(self, "_borrowed", False):
            lib().zkpor_r1cs_free(self.ctx._h, self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval(self, wires):
        """(a, b, c) = (L w, R w, O w) as (n_constraints, 4) uint64 Montgomery arrays"""
        outs = [np.zeros((self.n_constraints, 4), dtype=np.uint64) for _ in range(3)]
        _check(lib().zkpor_r1cs_eval(self.ctx._h, self._h, _ptr(wires), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])))
        return outs


# ----------------------------------------------------------------------------------------------- witness solver
class ProgramDesc(C.Structure):
    _fields_ = [("n_wires", C.c_uint64), ("n_public", C.c_uint64), ("n_secret", C.c_uint64), ("n_constraints", C.c_uint64),
                ("l", Csr), ("r", Csr), ("o", Csr), ("coeff_table", C.c_void_p), ("n_coeffs", C.c_uint64),
                ("n_instr", C.c_uint64), ("instr_kind", C.c_void_p), ("instr_arg", C.c_void_p),
                ("n_levels", C.c_uint64), ("level_ptr", C.c_void_p), ("level_instr", C.c_void_p),
                ("n_hints", C.c_uint64), ("hint_fn", C.c_void_p), ("hint_param", C.c_void_p), ("hint_out_first", C.c_void_p),
                ("hint_n_out", C.c_void_p), ("hint_in_ptr", C.c_void_p), ("hint_in_end", C.c_void_p),
                ("n_aux_rows", C.c_uint64), ("aux", Csr), ("n_tables", C.c_uint64), ("table_ptr", C.c_void_p)]


def fr_mont_limbs(values) -> np.ndarray:
    """canonical ints -> (k, 4) uint64 Montgomery limbs (fr.Element memory layout); host-side, for coefficient tables"""
    out = np.zeros((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        m = (int(v) % R_MOD) * (1 << 256) % R_MOD
        out[i] = [(m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]
    return out


class Program:
    """gnark's compiled constraint system, flattened (zkpor_program_desc), resident in HBM: matrices, instructions, levels, hints.
    `flat` = the dict circuit_synth.CircuitBuilder.flatten() returns (numpy arrays, or torch tensors already on the device)."""

    def __init__(self, ctx: Context, flat: dict):
        self.ctx = ctx
        self.n_wires, self.n_constraints = int(flat["n_wires"]), int(flat["n_constraints"])
        self.n_inputs = int(flat["n_public"]) - 1 + int(flat["n_secret"])
        keep = []

        def hp(x, dtype):
            if isinstance(x, np.ndarray):
                x = np.ascontiguousarray(x, dtype=dtype)
            keep.append(x)
            return _ptr(x) if len(x) else None

        def csr(prefix):
            return Csr(len(flat[prefix + "_wire"]), hp(flat[prefix + "_row_ptr"], np.uint64), hp(flat[prefix + "_wire"], np.uint32), hp(flat[prefix + "_coeff"], np.uint32))

        d = ProgramDesc()
        d.n_wires, d.n_public, d.n_secret, d.n_constraints = self.n_wires, int(flat["n_public"]), int(flat["n_secret"]), self.n_constraints
        d.l, d.r, d.o = csr("l"), csr("r"), csr("o")
        tab = fr_mont_limbs(flat["coeffs"])
        keep.append(tab)
        d.coeff_table, d.n_coeffs = _ptr(tab), tab.shape[0]
        d.n_instr, d.instr_kind, d.instr_arg = int(flat["n_instr"]), hp(flat["instr_kind"], np.uint8), hp(flat["instr_arg"], np.uint32)
        d.n_levels, d.level_ptr, d.level_instr = int(flat["n_levels"]), hp(flat["level_ptr"], np.uint64), hp(flat["level_instr"], np.uint32)
        d.n_hints = int(flat["n_hints"])
        d.hint_fn, d.hint_param = hp(flat["hint_fn"], np.uint32), hp(flat["hint_param"], np.uint32)
        d.hint_out_first, d.hint_n_out = hp(flat["hint_out_first"], np.uint32), hp(flat["hint_n_out"], np.uint32)
        d.hint_in_ptr, d.hint_in_end = hp(flat["hint_in_ptr"], np.uint64), hp(flat["hint_in_end"], np.uint64)
        d.n_aux_rows = len(flat["aux_row_ptr"]) - 1
        d.aux = csr("aux")
        d.n_tables, d.table_ptr = int(flat["n_tables"]), hp(np.asarray(flat["table_ptr"], dtype=np.uint64), np.uint64)
        self._h = C.c_void_p()
        _check(lib().zkpor_program_upload(ctx._h, C.byref(d), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().zkpor_program_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self) -> dict:
        out = (C.c_uint64 * 4)()
        _check(lib().zkpor_program_stats(self._h, out))
        t = (C.c_uint64 * 3)()
        _check(lib().zkpor_program_tail_info(self._h, t))
        return dict(wide_levels=out[0], narrow_runs=out[1], narrow_levels=out[2], count_hints=out[3],
                    deferred_tail=dict(levels=t[0], wires=t[1], starts_before_step=t[2]))

    def r1cs(self) -> "R1CS":
        """the program's matrices as an R1CS handle (borrowed: valid while the program lives; closing it is a no-op)"""
        cs = R1CS.__new__(R1CS)
        cs.ctx, cs.n_constraints, cs._h, cs._borrowed = self.ctx, self.n_constraints, C.c_void_p(), True
        _check(lib().zkpor_program_r1cs(self._h, C.byref(cs._h)))
        return cs

    def tail_wires(self) -> np.ndarray:
        n = self.stats()["deferred_tail"]["wires"]
        out = np.zeros(max(n, 1), dtype=np.uint32)
        _check(lib().zkpor_program_tail_wires(self._h, _ptr(out), C.c_uint64(out.size)))
        return out[:n]

    def solve(self, inputs, pk: "ProvingKey" = None, want_abc=True):
        """r1cs.Solve: inputs = (n_public - 1 + n_secret, 4) uint64 Montgomery -> (wires, a, b, c, commitment) as numpy arrays"""
        w = np.zeros((self.n_wires, 4), dtype=np.uint64)
        abc = [np.zeros((self.n_constraints, 4), dtype=np.uint64) for _ in range(3)] if want_abc else [None] * 3
        cm = np.zeros(8, dtype=np.uint64)
        _check(lib().zkpor_r1cs_solve(self.ctx._h, self._h, pk._h if pk is not None else None, _ptr(inputs), _ptr(w), _ptr(abc[0]), _ptr(abc[1]),
                                      _ptr(abc[2]), _ptr(cm)))
        return w, abc[0], abc[1], abc[2], cm

    def solve_device(self, inputs, d_wires, pk: "ProvingKey" = None, abc=(None, None, None)):
        """solve with inputs / wires (and optionally a, b, c) resident in HBM: tensors or raw device pointers"""
        _check(lib().zkpor_r1cs_solve(self.ctx._h, self._h, pk._h if pk is not None else None, _ptr(inputs), _ptr(d_wires), _ptr(abc[0]), _ptr(abc[1]),
                                      _ptr(abc[2]), None))


# ----------------------------------------------------------------------------------------------- pairing / Verify
class VkDesc(C.Structure):
    _fields_ = [("g1_alpha", C.c_void_p), ("g2_beta", C.c_void_p), ("g2_gamma", C.c_void_p), ("g2_delta", C.c_void_p),
                ("g1_k", C.c_void_p), ("n_k", C.c_uint64), ("n_commitments", C.c_uint64),
                ("public_committed", C.c_void_p), ("n_public_committed", C.c_uint64),
                ("g2_ped_g", C.c_void_p), ("g2_ped_g_root_sigma_neg", C.c_void_p)]


def pairing_product(ctx: Context, g1_points, g2_points, n: int) -> np.ndarray:
    """prod_i e(P_i, Q_i) as (12, 4) uint64: Montgomery Fp limbs in gnark-crypto E12 order (bn254.Pair)."""
    out = np.zeros((12, 4), dtype=np.uint64)
    _check(lib().zkpor_pairing_product(ctx._h, _ptr(g1_points), _ptr(g2_points), C.c_uint64(n), _ptr(out)))
    return out


def pairing_check(ctx: Context, g1_points, g2_points, n: int) -> bool:
    ok = C.c_int32(0)
    _check(lib().zkpor_pairing_check(ctx._h, _ptr(g1_points), _ptr(g2_points), C.c_uint64(n), C.byref(ok)))
    return bool(ok.value)


class VerifyingKey:
    """groth16.VerifyingKey (what vk.ReadFrom fills at src/verifier/main.go:33-34): numpy uint64 arrays in gnark memory
    layout.  K = vk.G1.K (ONE wire, public inputs, commitment wire)."""

    def __init__(self, *, alpha1, beta2, gamma2, delta2, K, n_commitments=1, public_committed=(), ped_g=None, ped_g_root_sigma_neg=None):
        h = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.uint64)
        self._keep = dict(alpha1=h(alpha1), beta2=h(beta2), gamma2=h(gamma2), delta2=h(delta2), K=h(K), ped_g=h(ped_g), grsn=h(ped_g_root_sigma_neg),
                          pc=np.ascontiguousarray(public_committed, dtype=np.uint64))
        k = self._keep
        self.desc = VkDesc(_ptr(k["alpha1"]), _ptr(k["beta2"]), _ptr(k["gamma2"]), _ptr(k["delta2"]), _ptr(k["K"]), k["K"].reshape(-1, 8).shape[0],
                           n_commitments, _ptr(k["pc"]) if k["pc"].size else None, k["pc"].size, _ptr(k["ped_g"]), _ptr(k["grsn"]))

    def verify(self, ctx: Context, proof_raw: bytes, public_witness) -> bool:
        """groth16.Verify(proof, vk, publicWitness) (prover.go:276, verifier/main.go:284); public_witness = (n, 4) uint64 Montgomery."""
        pw = np.ascontiguousarray(public_witness, dtype=np.uint64).reshape(-1, 4)
        buf = np.frombuffer(proof_raw, dtype=np.uint8).copy()
        ok = C.c_int32(0)
        _check(lib().zkpor_groth16_verify(ctx._h, C.byref(self.desc), _ptr(buf), C.c_uint32(buf.size), _ptr(pw) if pw.size else None,
                                          C.c_uint64(pw.shape[0]), C.byref(ok)))
        return bool(ok.value)

    def verify_batch(self, ctx: Context, proofs_raw, public_witnesses, seed: bytes = bytes(32)) -> bool:
        """all proofs of one circuit with one pairing product (random linear combination bound to `seed` and the batch)"""
        count = len(proofs_raw)
        plen = len(proofs_raw[0])
        buf = np.frombuffer(b"".join(proofs_raw), dtype=np.uint8).copy()
        pw = np.ascontiguousarray(public_witnesses, dtype=np.uint64).reshape(count, -1, 4)
        sd = np.frombuffer(seed, dtype=np.uint8).copy()
        ok = C.c_int32(0)
        _check(lib().zkpor_groth16_verify_batch(ctx._h, C.byref(self.desc), _ptr(buf), C.c_uint32(plen), C.c_uint64(plen), _ptr(pw) if pw.size else None,
                                                C.c_uint64(pw.shape[1]), C.c_uint64(count), _ptr(sd), C.byref(ok)))
        return bool(ok.value)


# ----------------------------------------------------------------------------------------------- multi-GPU host logic
def comm_unique_id() -> bytes:
    """NCCL unique id for Context.comm_init (rank 0 creates it, every rank receives it by any channel)"""
    out = (C.c_uint8 * 128)()
    _check(lib().zkpor_comm_unique_id(out))
    return bytes(out)


def create_multi(device_ids) -> list:
    """zkpor_ctx_create_multi: len(device_ids) contexts joined in one in-process group (drive each from its own thread)"""
    n = len(device_ids)
    ids = (C.c_int32 * n)(*device_ids)
    hs = (C.c_void_p * n)()
    _check(lib().zkpor_ctx_create_multi(ids, C.c_int32(n), hs))
    return [Context(device_ids[i], _handle=hs[i]) for i in range(n)]


def multi_prove_solve(ctxs, pks, progs, inputs, r: int, s: int) -> bytes:
    """zkpor_multi_prove_solve: one proof across the group's GPUs, one host thread per context inside the library"""
    n = len(ctxs)
    out = np.zeros(388, dtype=np.uint8)
    ln = C.c_uint32(0)
    _check(lib().zkpor_multi_prove_solve((C.c_void_p * n)(*[c._h.value for c in ctxs]), (C.c_void_p * n)(*[p._h.value for p in pks]),
                                         (C.c_void_p * n)(*[p._h.value for p in progs]), C.c_int32(n), _ptr(inputs), _be_arr(r), _be_arr(s),
                                         _ptr(out), C.byref(ln)))
    return out[:ln.value].tobytes()


def tree_shard_range(rank: int, world: int, capacity: int, depth: int):
    """host mirror of zkpor_tree_shard_range: (first key, number of keys, subtree level k) of a rank when one tree is built over
    `world` GPUs -- k is the smallest level with world << k >= capacity; rank g owns the leaves [g << k, (g + 1) << k)"""
    k = 0
    while k < depth and (world << k) < capacity:
        k += 1
    first = rank << k
    if first >= capacity:
        return capacity, 0, k
    return first, min(1 << k, capacity - first), k


def run_ranks(fn, n: int) -> list:
    """fn(rank) on n host threads (ctypes calls release the GIL, so collective library calls proceed concurrently); re-raises the
    first failure"""
    import threading
    res, errs = [None] * n, [None] * n

    def work(i):
        try:
            res[i] = fn(i)
        except BaseException as e:   # noqa: BLE001
            errs[i] = e

    th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return res


def chunk_bounds(length: int, rank: int, world: int):
    """Point-chunk sharding of one key array: rank r owns [r*L/N, (r+1)*L/N)."""
    return (length * rank) // world, (length * (rank + 1)) // world


def finish_proof(partials: np.ndarray, alpha1, beta1, delta1, beta2, delta2, r: int, s: int, has_commitment: bool = True) -> bytes:
    """zkpor_groth16_finish: combine the ranks' partial sums (after the all-gather) into proof.WriteRawTo bytes.
    Pure host arithmetic -- needs no GPU and no context."""
    p = np.ascontiguousarray(partials, dtype=np.uint8).reshape(-1, PROVE_PARTIAL_BYTES)
    out = np.zeros(388, dtype=np.uint8)
    n = C.c_uint32(0)
    pts = [np.ascontiguousarray(x, dtype=np.uint64) for x in (alpha1, beta1, delta1, beta2, delta2)]
    _check(lib().zkpor_groth16_finish(_ptr(p), C.c_uint32(p.shape[0]), *[_ptr(x) for x in pts], _be_arr(r), _be_arr(s),
                                      C.c_int32(1 if has_commitment else 0), _ptr(out), C.byref(n)))
    return out[:n.value].tobytes()


def pack_partial(ar, bs1, krs_k, krs_z, commit, pok, bs2) -> np.ndarray:
    """One rank's 7 partial sums in the layout of zkpor_groth16_prove_partial: six G1 XYZZ (128 B) then one G2 XYZZ
    (256 B).  Inputs are XYZZ limb arrays (16 / 32 u64)."""
    out = np.zeros(PROVE_PARTIAL_BYTES, dtype=np.uint8)
    off = 0
    for x, nbytes in ((ar, 128), (bs1, 128), (krs_k, 128), (krs_z, 128), (commit, 128), (pok, 128), (bs2, 256)):
        b = np.ascontiguousarray(x, dtype=np.uint64).view(np.uint8)
        assert b.size == nbytes
        out[off:off + nbytes] = b
        off += nbytes
    return out
