"""CPU checks of the drop-in boundary: libzkpor_b200.so loads, exports every symbol include/zkpor_b200.h declares,
refuses to run without a GPU (no CPU fallback), and the host-side mirrors reproduce the reference's host logic."""
import os
import re

import numpy as np
import pytest

import merkle
import zkpor_b200 as zk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zkpor_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zkpor_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = zk.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/zkpor_b200.h but not exported"
    assert lib.zkpor_version().decode().startswith("zkpor_b200")
    assert [lib.zkpor_stage_name(i).decode() for i in range(3)] == ["h2d", "digits", "sort"]


def test_header_is_plain_c():
    """the boundary is a C ABI: the header must compile as C11 (cgo includes it as such)"""
    import subprocess
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "zkpor_b200.h")])


def test_null_arguments_of_the_newer_entry_points():
    lib = zk.lib()
    for call in (lambda: lib.zkpor_pairing_check(None, None, None, 0, None),
                 lambda: lib.zkpor_groth16_verify(None, None, None, 0, None, 0, None),
                 lambda: lib.zkpor_groth16_verify_batch(None, None, None, 0, 0, None, 0, 0, None, None),
                 lambda: lib.zkpor_r1cs_upload(None, 0, 0, None, None, None, None, 0, None),
                 lambda: lib.zkpor_r1cs_eval(None, None, None, None, None, None),
                 lambda: lib.zkpor_groth16_prove_wires(None, None, None, None, None, None, None, None),
                 lambda: lib.zkpor_msm_set_affine_rounds(None, 0)):
        assert call() == 1                              # ZKPOR_ERR_INVALID_ARG, never a crash
        assert lib.zkpor_last_error()


def test_null_arguments_of_the_round2_entry_points():
    """solver, sharded proof, containers, witness batches: a null handle is ZKPOR_ERR_INVALID_ARG with a message, never a crash"""
    lib = zk.lib()
    for call in (lambda: lib.zkpor_program_upload(None, None, None),
                 lambda: lib.zkpor_program_stats(None, None),
                 lambda: lib.zkpor_program_tail_info(None, None),
                 lambda: lib.zkpor_program_tail_wires(None, None, 0),
                 lambda: lib.zkpor_program_r1cs(None, None),
                 lambda: lib.zkpor_r1cs_solve(None, None, None, None, None, None, None, None, None),
                 lambda: lib.zkpor_groth16_prove_solve(None, None, None, None, None, None, None, None),
                 lambda: lib.zkpor_pk_upload_shard(None, None, None),
                 lambda: lib.zkpor_pk_shard_info(None, None),
                 lambda: lib.zkpor_proof_decode(None, None, 0, None, None, None),
                 lambda: lib.zkpor_proof_encode(None, None, 0, 1, None, None),
                 lambda: lib.zkpor_vk_decode(None, None, 0, None, None, 0, None, 0, None),
                 lambda: lib.zkpor_vk_encode(None, None, None, None, 0, None, 0, None),
                 lambda: lib.zkpor_pk_read(None, None, 0, None, None, None),
                 lambda: lib.zkpor_pk_write(None, None, 0, None, 0, None),
                 lambda: lib.zkpor_g1_encode_batch(None, None, 1, 1, None),
                 lambda: lib.zkpor_witness_batches(None, None, None, None, None, 0, 50, 1380, None, None, None),
                 lambda: lib.zkpor_tree_build_sharded(None, None),
                 lambda: lib.zkpor_tree_shard_range(None, None, None, None, None)):
        assert call() == 1
        assert lib.zkpor_last_error()
    assert lib.zkpor_program_free(None, None) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert zk.device_count() == 0
    with pytest.raises(zk.ZkporError, match="no CPU fallback"):
        zk.Context(0)


def test_null_arguments_are_reported_not_crashes():
    lib = zk.lib()
    assert lib.zkpor_ctx_create(0, None) == 1          # ZKPOR_ERR_INVALID_ARG
    assert b"null" in lib.zkpor_last_error()
    assert lib.zkpor_msm_g1(None, None, None, 0, 0, None) == 1
    assert lib.zkpor_ctx_destroy(None) == 0


def test_host_mirror_padding_matches_reference_rule():
    cases = [[], [(3, 1, 2, 3, 4, 5), (7, 9, 9, 9, 9, 9)], [(i, 1, 0, 0, 0, 0) for i in range(50)],
             [(i * 3, 5, 4, 3, 2, 1) for i in range(51)], [(499, 1, 1, 1, 1, 1)]]
    for assets in cases:
        assert zk.padding_account_assets(assets).tolist() == merkle.padding_account_assets(assets)
    assert zk.assets_count_tier(50) == 50 and zk.assets_count_tier(51) == 500
    with pytest.raises(ValueError):
        zk.assets_count_tier(501)


def test_sum_partials_host_side():
    """zkpor_g1_sum_partials is host arithmetic (the tail of the multi-GPU combine): identity + infinity handling."""
    import orc
    import bn254 as bn
    pts = orc.g1_fixed_base(orc.ints_to_limbs([5, 7, 11]))
    one = orc.fp_mont([1])[0]
    parts = np.zeros((4, 16), dtype=np.uint64)
    for i in range(3):
        parts[i, :8] = pts[i]; parts[i, 8:12] = one; parts[i, 12:16] = one   # XYZZ of an affine point: ZZ = ZZZ = 1
    # parts[3] stays all-zero: ZZ = 0 encodes infinity
    assert orc.g1_unpack(zk.g1_sum_partials(parts))[0] == bn.pt_mul(bn.G1_GEN, 23)
    pts2 = orc.g2_fixed_base(orc.ints_to_limbs([5, 7]))
    parts2 = np.zeros((2, 32), dtype=np.uint64)
    for i in range(2):
        parts2[i, :16] = pts2[i]; parts2[i, 16:20] = one; parts2[i, 24:28] = one
    assert orc.g2_unpack(zk.g2_sum_partials(parts2))[0] == bn.pt_mul(bn.G2_GEN, 12, bn.FP2)


def test_cpp_host_mirror_compiles_links_and_guards():
    """include/zkpor_b200.hpp (the compiled-language host layer) against the shared library."""
    import subprocess
    pkg = os.path.join(ROOT, "zkmerkle-proof-of-solvency_b200")
    exe = os.path.join(pkg, "_build", "host_mirror_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                           "-L", pkg, "-lzkpor_b200", "-Wl,-rpath," + pkg, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
