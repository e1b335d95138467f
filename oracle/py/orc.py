"""ORACLE (test infrastructure, NOT product code) -- ctypes binding of oracle/_build/liborc.so (oracle/c, the fast
CPU restatement) plus int <-> limb helpers shared by the tests.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import bn254 as bn

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "..", "_build", "liborc.so")


def build(force: bool = False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "..")], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_merkle_level_len.restype = C.c_size_t
        _lib.orc_merkle_level_len.argtypes = [C.c_size_t, C.c_int]
        _lib.orc_merkle_nodes_total.restype = C.c_size_t
        _lib.orc_merkle_nodes_total.argtypes = [C.c_size_t, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------- int <-> limbs
MASK64 = (1 << 64) - 1


def ints_to_limbs(vals, nlimbs=4) -> np.ndarray:
    """plain integers -> (n, nlimbs) uint64 little-endian limbs"""
    out = np.empty((len(vals), nlimbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(nlimbs):
            out[i, k] = (v >> (64 * k)) & MASK64
    return out


def limbs_to_ints(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(arr[i, k]) << (64 * k) for k in range(4)) for i in range(arr.shape[0])]


def fr_mont(vals) -> np.ndarray:
    """integers -> Montgomery-form Fr elements (gnark-crypto memory layout)"""
    return ints_to_limbs([v * bn.MONT_R % bn.R for v in vals])


def fr_unmont(arr) -> list:
    rinv = pow(bn.MONT_R, -1, bn.R)
    return [v * rinv % bn.R for v in limbs_to_ints(arr)]


def fp_mont(vals) -> np.ndarray:
    return ints_to_limbs([v * bn.MONT_R % bn.P for v in vals])


def fp_unmont(arr) -> list:
    rinv = pow(bn.MONT_R, -1, bn.P)
    return [v * rinv % bn.P for v in limbs_to_ints(arr)]


def g1_pack(points) -> np.ndarray:
    """affine points (None = infinity) -> (n, 8) uint64, Montgomery"""
    flat = []
    for pt in points:
        flat += [0, 0] if pt is None else [pt[0], pt[1]]
    return fp_mont(flat).reshape(-1, 8)


def g1_unpack(arr) -> list:
    v = fp_unmont(np.asarray(arr).reshape(-1, 4))
    out = []
    for i in range(0, len(v), 2):
        out.append(None if v[i] == 0 and v[i + 1] == 0 else (v[i], v[i + 1]))
    return out


def g2_pack(points) -> np.ndarray:
    flat = []
    for pt in points:
        flat += [0, 0, 0, 0] if pt is None else [pt[0][0], pt[0][1], pt[1][0], pt[1][1]]
    return fp_mont(flat).reshape(-1, 16)


def g2_unpack(arr) -> list:
    v = fp_unmont(np.asarray(arr).reshape(-1, 4))
    out = []
    for i in range(0, len(v), 4):
        out.append(None if not any(v[i:i + 4]) else ((v[i], v[i + 1]), (v[i + 2], v[i + 3])))
    return out


def be32_array(vals) -> np.ndarray:
    return np.frombuffer(b"".join(int(v).to_bytes(32, "big") for v in vals), dtype=np.uint8).reshape(-1, 32).copy()


# ----------------------------------------------------------------------------- wrappers
def g1_fixed_base(scalars_plain: np.ndarray, threads=0) -> np.ndarray:
    s = np.ascontiguousarray(scalars_plain, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((s.shape[0], 8), dtype=np.uint64)
    lib().orc_g1_fixed_base(_p(s), C.c_size_t(s.shape[0]), _p(out), C.c_int(threads))
    return out


def g2_fixed_base(scalars_plain: np.ndarray, threads=0) -> np.ndarray:
    s = np.ascontiguousarray(scalars_plain, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((s.shape[0], 16), dtype=np.uint64)
    lib().orc_g2_fixed_base(_p(s), C.c_size_t(s.shape[0]), _p(out), C.c_int(threads))
    return out


def g1_msm(pts: np.ndarray, scalars_mont: np.ndarray, threads=0) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.uint64); sc = np.ascontiguousarray(scalars_mont, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    lib().orc_g1_msm(_p(pts), _p(sc), C.c_size_t(sc.size // 4), _p(out), C.c_int(threads))
    return out


def g2_msm(pts: np.ndarray, scalars_mont: np.ndarray, threads=0) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.uint64); sc = np.ascontiguousarray(scalars_mont, dtype=np.uint64)
    out = np.empty(16, dtype=np.uint64)
    lib().orc_g2_msm(_p(pts), _p(sc), C.c_size_t(sc.size // 4), _p(out), C.c_int(threads))
    return out


def ntt(data_mont: np.ndarray, logn: int, inverse: bool, dit: bool, coset: bool, threads=0) -> np.ndarray:
    a = np.array(data_mont, dtype=np.uint64, copy=True).reshape(-1, 4)
    assert a.shape[0] == 1 << logn
    lib().orc_ntt(_p(a), C.c_int(logn), C.c_int(inverse), C.c_int(dit), C.c_int(coset), C.c_int(threads))
    return a


def compute_h(a, b, c, logn: int, threads=0) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    c = np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((1 << logn, 4), dtype=np.uint64)
    lib().orc_compute_h(_p(a), _p(b), _p(c), C.c_size_t(a.shape[0]), C.c_int(logn), _p(out), C.c_int(threads))
    return out


def poseidon_set_out_lane(lane: int):
    lib().orc_poseidon_set_out_lane(C.c_int(lane))


def poseidon_constants(t: int):
    rp = C.c_int(0)
    lib().orc_poseidon_constants(C.c_int(t), None, None, C.byref(rp))
    rc = np.empty(((8 + rp.value) * t, 4), dtype=np.uint64)
    mds = np.empty((t * t, 4), dtype=np.uint64)
    lib().orc_poseidon_constants(C.c_int(t), _p(rc), _p(mds), C.byref(rp))
    return rc, mds, rp.value


def poseidon_hash(in_mont: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(in_mont, dtype=np.uint64).reshape(-1, 4)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_poseidon_hash(_p(a), C.c_size_t(a.shape[0]), _p(out))
    return out


def poseidon_node_batch(pairs_be: np.ndarray, threads=0) -> np.ndarray:
    p = np.ascontiguousarray(pairs_be, dtype=np.uint8).reshape(-1, 64)
    out = np.empty((p.shape[0], 32), dtype=np.uint8)
    lib().orc_poseidon_node_batch(_p(p), C.c_size_t(p.shape[0]), _p(out), C.c_int(threads))
    return out


def merkle_level_len(capacity: int, level: int) -> int:
    return lib().orc_merkle_level_len(capacity, level)


def merkle_build(leaves_be: np.ndarray, capacity: int, depth: int, nil_leaf: bytes, dirty: np.ndarray | None = None, threads=0):
    """returns (nodes[(total,32)], root bytes)"""
    leaves = np.ascontiguousarray(leaves_be, dtype=np.uint8).reshape(-1, 32)
    assert leaves.shape[0] == capacity
    total = lib().orc_merkle_nodes_total(capacity, depth)
    nodes = np.empty((total, 32), dtype=np.uint8)
    root = np.empty(32, dtype=np.uint8)
    nil = np.frombuffer(nil_leaf, dtype=np.uint8).copy()
    d = None if dirty is None else np.ascontiguousarray(dirty, dtype=np.uint64)
    lib().orc_merkle_build(_p(leaves), _p(d), C.c_size_t(capacity), C.c_int(depth), _p(nil), _p(nodes), _p(root), C.c_int(threads))
    return nodes, root.tobytes()


def merkle_proofs(leaves_be, nodes, capacity, depth, nil_leaf: bytes, keys, dirty=None) -> np.ndarray:
    leaves = np.ascontiguousarray(leaves_be, dtype=np.uint8)
    k = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.empty((k.size, depth, 32), dtype=np.uint8)
    nil = np.frombuffer(nil_leaf, dtype=np.uint8).copy()
    d = None if dirty is None else np.ascontiguousarray(dirty, dtype=np.uint64)
    lib().orc_merkle_proofs(_p(leaves), _p(d), _p(np.ascontiguousarray(nodes)), C.c_size_t(capacity), C.c_int(depth), _p(nil),
                            _p(k), C.c_size_t(k.size), _p(out))
    return out


def account_leaves(ids_be, totals_be, flat_assets, tier: int, threads=0) -> np.ndarray:
    ids = np.ascontiguousarray(ids_be, dtype=np.uint8).reshape(-1, 32)
    tot = np.ascontiguousarray(totals_be, dtype=np.uint8).reshape(-1, 96)
    fa = np.ascontiguousarray(flat_assets, dtype=np.uint64).reshape(ids.shape[0], tier * 6)
    out = np.empty((ids.shape[0], 32), dtype=np.uint8)
    lib().orc_account_leaves(_p(ids), _p(tot), _p(fa), C.c_size_t(ids.shape[0]), C.c_int(tier), _p(out), C.c_int(threads))
    return out


class OrcPk(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_a", "n_b", "n_k", "n_z", "n_ck")] + \
               [(n, C.c_void_p) for n in ("A", "B1", "K", "Z", "B2", "ck_basis", "ck_basis_exp_sigma",
                                          "alpha1", "beta1", "delta1", "beta2", "delta2")] + [("log_n", C.c_int)]


def groth16_prove(pkarr: dict, wa, wb, wk, committed, a, b, c, r: int, s: int, threads=0) -> bytes:
    """pkarr: dict of numpy arrays (A,B1,K,Z,B2,ck_basis,ck_basis_exp_sigma,alpha1,beta1,delta1,beta2,delta2) + log_n"""
    keep = {k: np.ascontiguousarray(v, dtype=np.uint64) for k, v in pkarr.items() if k != "log_n"}
    pk = OrcPk()
    pk.n_a, pk.n_b, pk.n_k, pk.n_z, pk.n_ck = (keep["A"].size // 8, keep["B1"].size // 8, keep["K"].size // 8,
                                                keep["Z"].size // 8, keep["ck_basis"].size // 8)
    for k in ("A", "B1", "K", "Z", "B2", "ck_basis", "ck_basis_exp_sigma", "alpha1", "beta1", "delta1", "beta2", "delta2"):
        setattr(pk, k, keep[k].ctypes.data)
    pk.log_n = pkarr["log_n"]
    arrs = [np.ascontiguousarray(x, dtype=np.uint64) for x in (wa, wb, wk, committed, a, b, c)]
    rs = ints_to_limbs([r % bn.R, s % bn.R])
    out = np.empty(388, dtype=np.uint8)
    rc = lib().orc_groth16_prove(C.byref(pk), *[_p(x) for x in arrs], C.c_size_t(arrs[4].size // 4),
                                 _p(rs[0]), _p(rs[1]), _p(out), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_groth16_prove failed: {rc}")
    return out.tobytes()


def fr_index_sums(v: np.ndarray, threads=0):
    """(sum_i v_i, sum_i i*v_i) mod r with the limb patterns taken as integers"""
    a = np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4)
    s = np.zeros(4, dtype=np.uint64); t = np.zeros(4, dtype=np.uint64)
    lib().orc_fr_index_sums(_p(a), C.c_size_t(a.shape[0]), _p(s), _p(t), C.c_int(threads))
    return limbs_to_ints(s)[0], limbs_to_ints(t)[0]


def eval_barycentric(evals_mont: np.ndarray, logn: int, x0: int, threads=0) -> int:
    a = np.ascontiguousarray(evals_mont, dtype=np.uint64).reshape(-1, 4)
    x = fr_mont([x0]); out = np.zeros(4, dtype=np.uint64)
    lib().orc_eval_barycentric(_p(a), C.c_size_t(a.shape[0]), C.c_int(logn), _p(x), _p(out), C.c_int(threads))
    return fr_unmont(out)[0]


def poly_eval_bitrev(coef_mont: np.ndarray, logn: int, x0: int, threads=0) -> int:
    a = np.ascontiguousarray(coef_mont, dtype=np.uint64).reshape(-1, 4)
    assert a.shape[0] == 1 << logn
    x = fr_mont([x0]); out = np.zeros(4, dtype=np.uint64)
    lib().orc_poly_eval_bitrev(_p(a), C.c_int(logn), _p(x), _p(out), C.c_int(threads))
    return fr_unmont(out)[0]
