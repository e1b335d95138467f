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


# ----------------------------------------------------------------------------- r1cs.Solve over the flat program (orc_solver.c)
class OrcProgram(C.Structure):
    _fields_ = ([("n_wires", C.c_uint64), ("n_public", C.c_uint64), ("n_secret", C.c_uint64), ("n_constraints", C.c_uint64)] +
                [(f"{m}_{k}", C.c_void_p) for m in "lro" for k in ("row_ptr", "wire", "coeff")] +
                [("coeffs", C.c_void_p), ("n_coeffs", C.c_uint64), ("n_instr", C.c_uint64), ("instr_kind", C.c_void_p), ("instr_arg", C.c_void_p),
                 ("n_levels", C.c_uint64), ("level_ptr", C.c_void_p), ("level_instr", C.c_void_p), ("n_hints", C.c_uint64),
                 ("hint_fn", C.c_void_p), ("hint_param", C.c_void_p), ("hint_out_first", C.c_void_p), ("hint_n_out", C.c_void_p),
                 ("hint_in_ptr", C.c_void_p), ("hint_in_end", C.c_void_p), ("aux_row_ptr", C.c_void_p), ("aux_wire", C.c_void_p),
                 ("aux_coeff", C.c_void_p), ("table_ptr", C.c_void_p), ("private_committed", C.c_void_p), ("n_committed", C.c_uint64)])


COMMIT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)
SOLVE_ERRORS = {1: "more than one unsolved wire", 2: "division by zero", 3: "index outside a table", 4: "unknown hint", 5: "hint reads an unsolved wire",
                6: "commitment hint without callback", 7: "committed wire unsolved", 8: "wire never solved", 9: "constraint not satisfied"}


def program_struct(flat: dict):
    """(OrcProgram, keep-alive list) from circuit_synth.flatten() output (numpy arrays)"""
    keep = []

    def arr(x, dt):
        a = np.ascontiguousarray(np.asarray(x), dtype=dt)
        if a.size == 0:
            a = np.zeros(1, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    p = OrcProgram()
    p.n_wires, p.n_public, p.n_secret, p.n_constraints = flat["n_wires"], flat["n_public"], flat["n_secret"], flat["n_constraints"]
    for m in "lro":
        setattr(p, f"{m}_row_ptr", arr(flat[f"{m}_row_ptr"], np.uint64))
        setattr(p, f"{m}_wire", arr(flat[f"{m}_wire"], np.uint32)); setattr(p, f"{m}_coeff", arr(flat[f"{m}_coeff"], np.uint32))
    tab = fr_mont(flat["coeffs"]); keep.append(tab)
    p.coeffs, p.n_coeffs = tab.ctypes.data, len(flat["coeffs"])
    p.n_instr, p.instr_kind, p.instr_arg = flat["n_instr"], arr(flat["instr_kind"], np.uint8), arr(flat["instr_arg"], np.uint32)
    p.n_levels, p.level_ptr, p.level_instr = flat["n_levels"], arr(flat["level_ptr"], np.uint64), arr(flat["level_instr"], np.uint32)
    p.n_hints = flat["n_hints"]
    for k in ("hint_fn", "hint_param", "hint_out_first", "hint_n_out"):
        setattr(p, k, arr(flat[k], np.uint32))
    p.hint_in_ptr, p.hint_in_end = arr(flat["hint_in_ptr"], np.uint64), arr(flat["hint_in_end"], np.uint64)
    p.aux_row_ptr, p.aux_wire, p.aux_coeff = arr(flat["aux_row_ptr"], np.uint64), arr(flat["aux_wire"], np.uint32), arr(flat["aux_coeff"], np.uint32)
    p.table_ptr = arr(flat["table_ptr"], np.uint64)
    p.private_committed, p.n_committed = arr(flat["private_committed"], np.uint64), len(flat["private_committed"])
    return p, keep


def solve(flat: dict, inputs_mont: np.ndarray, commit_fn=None, threads=0):
    """r1cs.Solve on the CPU.  commit_fn(values_mont (n, 4) uint64) -> challenge as int.  Returns (wires, a, b, c) Montgomery arrays."""
    p, keep = program_struct(flat)
    w = np.zeros((flat["n_wires"], 4), dtype=np.uint64)
    abc = [np.zeros((flat["n_constraints"], 4), dtype=np.uint64) for _ in range(3)]
    ins = np.ascontiguousarray(inputs_mont, dtype=np.uint64)

    def cb(vals, n, out, _user):
        v = np.ctypeslib.as_array(C.cast(vals, C.POINTER(C.c_uint64)), shape=(max(n, 1), 4))[:n].copy()
        ch = fr_mont([commit_fn(v) % bn.R])
        C.memmove(out, ch.ctypes.data, 32)

    err_at = C.c_uint64(0)
    rc = lib().orc_solve(C.byref(p), _p(ins), _p(w), _p(abc[0]), _p(abc[1]), _p(abc[2]), COMMIT_FN(cb) if commit_fn else COMMIT_FN(), None,
                         C.byref(err_at), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_solve: {SOLVE_ERRORS.get(rc, rc)} at {err_at.value}")
    return w, abc[0], abc[1], abc[2]


def commitment_challenge(raw64: bytes) -> int:
    out = np.zeros(4, dtype=np.uint64)
    buf = np.frombuffer(raw64, dtype=np.uint8).copy()
    lib().orc_commitment_challenge(_p(buf), C.c_size_t(len(raw64)), _p(out))
    return fr_unmont(out)[0]


def groth16_prove_program(pkarr: dict, flat: dict, infinity_a, infinity_b, inputs_mont, r: int, s: int, threads=0):
    """the whole groth16.Prove on the CPU: solver + proof.  Returns (proof bytes, (solve seconds, prove seconds))."""
    keep = {k: np.ascontiguousarray(v, dtype=np.uint64) for k, v in pkarr.items() if k != "log_n"}
    pk = OrcPk()
    pk.n_a, pk.n_b, pk.n_k, pk.n_z, pk.n_ck = (keep["A"].size // 8, keep["B1"].size // 8, keep["K"].size // 8,
                                                keep["Z"].size // 8, keep["ck_basis"].size // 8)
    for k in ("A", "B1", "K", "Z", "B2", "ck_basis", "ck_basis_exp_sigma", "alpha1", "beta1", "delta1", "beta2", "delta2"):
        setattr(pk, k, keep[k].ctypes.data)
    pk.log_n = pkarr["log_n"]
    p, keep2 = program_struct(flat)
    ia = np.ascontiguousarray(infinity_a, dtype=np.uint8); ib = np.ascontiguousarray(infinity_b, dtype=np.uint8)
    ins = np.ascontiguousarray(inputs_mont, dtype=np.uint64)
    rs = ints_to_limbs([r % bn.R, s % bn.R])
    out = np.empty(388, dtype=np.uint8)
    secs = (C.c_double * 2)()
    err_at = C.c_uint64(0)
    rc = lib().orc_groth16_prove_program(C.byref(pk), C.byref(p), _p(ia), _p(ib), C.c_uint64(max(0, int(flat["commitment_index"]))), _p(ins),
                                         _p(rs[0]), _p(rs[1]), _p(out), secs, C.byref(err_at), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_groth16_prove_program: {SOLVE_ERRORS.get(rc, rc)} at {err_at.value}")
    return out.tobytes(), (secs[0], secs[1])
