"""ORACLE (test infrastructure, NOT product code) -- gnark's R1CS solver restated over the flat program arrays.

Reference: groth16.Prove's first step, r1cs.Solve (src/prover/prover/prover.go:269 -> gnark constraint/bn254/solver.go +
constraint/bn254/system.go, out of tree, bnb-chain/gnark v0.10.1-0.20240910145009-4b5261061f04), with the hints the circuit uses:
the user hint IntegerDivision (circuit/utils.go:103-110, registered at prover.go:68), gnark's std hints (bits.NBits, InvZero,
rangecheck decomposition, logderivlookup lookup, logderivarg multiplicity count) and the BSB22 commitment placeholder that Prove
overrides.  "Parity unpinned" against gnark itself (no Go toolchain); pinned by construction: every constraint
of the solved system is checked (L.w * R.w == O.w), and the proof made from the solution verifies (tests/test_oracle_solver.py).

The walk is gnark's: levels in order; inside a level every instruction independently; an R1C instruction looks for its ONE
unsolved wire at run time (solved flags), divides when that wire sits in L or R; a hint instruction evaluates its input
expressions and writes its consecutive output wires.
"""
from __future__ import annotations

import numpy as np

from bn254 import R

INS_R1C, INS_HINT = 0, 1
H_DIVMOD, H_NBITS, H_INVZERO, H_DECOMPOSE, H_LOOKUP, H_CMP, H_COUNT, H_COMMIT = 1, 2, 3, 4, 5, 6, 7, 8


class Unsatisfied(Exception):
    pass


def limbs_to_int(row) -> int:
    return int(row[0]) | int(row[1]) << 64 | int(row[2]) << 128 | int(row[3]) << 192


def solve_program(flat, inputs, commit_fn=None):
    """inputs: (n_public - 1 + n_secret) canonical ints.  commit_fn(values) -> challenge (int) implements the overridden
    BSB22 hint: Pedersen-commit the private committed wires, hash the commitment to the field.
    Returns (wires, a, b, c, info); info["committed"] = the committed values in key order."""
    nw = flat["n_wires"]
    coeffs = flat["coeffs"]
    w = [None] * nw
    w[0] = 1
    assert len(inputs) == flat["n_public"] - 1 + flat["n_secret"]
    for i, v in enumerate(inputs):
        w[1 + i] = int(v) % R
    mats = [(flat[k + "_row_ptr"], flat[k + "_wire"], flat[k + "_coeff"]) for k in "lro"]
    aux = (flat["aux_row_ptr"], flat["aux_wire"], flat["aux_coeff"])
    n_rows = flat["n_constraints"]
    a, b, c = [0] * n_rows, [0] * n_rows, [0] * n_rows
    info = {"committed": None}

    def eval_row(mat, row):
        """(sum of solved terms, [(coeff, wire)] of unsolved terms)"""
        ptr, wi, ci = mat
        acc, unk = 0, []
        for e in range(int(ptr[row]), int(ptr[row + 1])):
            wid, cf = int(wi[e]), coeffs[int(ci[e])]
            if w[wid] is None:
                unk.append((cf, wid))
            else:
                acc += cf * w[wid]
        return acc % R, unk

    def eval_aux(row):
        acc, unk = eval_row(aux, row)
        if unk:
            raise Unsatisfied(f"hint input row {row} reads an unsolved wire")
        return acc

    level_ptr, level_instr = flat["level_ptr"], flat["level_instr"]
    for lvl in range(flat["n_levels"]):
        writes = []
        for pos in range(int(level_ptr[lvl]), int(level_ptr[lvl + 1])):
            ins = int(level_instr[pos])
            arg = int(flat["instr_arg"][ins])
            if int(flat["instr_kind"][ins]) == INS_R1C:
                (av, ua), (bv, ub), (cv, uc) = (eval_row(m, arg) for m in mats)
                n_unk = len(ua) + len(ub) + len(uc)
                if n_unk > 1:
                    raise Unsatisfied(f"constraint {arg}: {n_unk} unsolved wires")
                if uc:
                    (cf, wid), = uc
                    val = (av * bv - cv) * pow(cf, -1, R) % R
                    cv = av * bv % R
                elif ua:
                    (cf, wid), = ua
                    if bv == 0:
                        raise Unsatisfied(f"constraint {arg}: division by zero")
                    val = (cv * pow(bv, -1, R) - av) * pow(cf, -1, R) % R
                    av = cv * pow(bv, -1, R) % R
                elif ub:
                    (cf, wid), = ub
                    if av == 0:
                        raise Unsatisfied(f"constraint {arg}: division by zero")
                    val = (cv * pow(av, -1, R) - bv) * pow(cf, -1, R) % R
                    bv = cv * pow(av, -1, R) % R
                else:
                    wid = None
                    if av * bv % R != cv:
                        raise Unsatisfied(f"constraint {arg} is not satisfied")
                if wid is not None:
                    writes.append((wid, val))
                a[arg], b[arg], c[arg] = av, bv, cv
            else:
                fn, param = int(flat["hint_fn"][arg]), int(flat["hint_param"][arg])
                first, n_out = int(flat["hint_out_first"][arg]), int(flat["hint_n_out"][arg])
                r0, r1 = int(flat["hint_in_ptr"][arg]), int(flat["hint_in_end"][arg])
                if fn == H_COUNT:
                    cnt = [0] * n_out
                    for row in range(r0, r1):
                        q = eval_aux(row)
                        if q >= n_out:
                            raise Unsatisfied(f"query {q} outside a table of {n_out} entries")
                        cnt[q] += 1
                    outs = cnt
                elif fn == H_COMMIT:
                    vals = [w[int(i)] for i in flat["private_committed"]]
                    assert all(v is not None for v in vals), "commitment hint fired before its wires were solved"
                    info["committed"] = vals
                    outs = [commit_fn(vals) % R]
                else:
                    ins_v = [eval_aux(row) for row in range(r0, r1)]
                    if fn == H_DIVMOD:                      # circuit/utils.go:103-110: out[0].DivMod(in[0], in[1], out[1])
                        if ins_v[1] == 0:
                            raise Unsatisfied("IntegerDivision by zero")
                        outs = list(divmod(ins_v[0], ins_v[1]))
                    elif fn == H_NBITS:
                        outs = [(ins_v[0] >> i) & 1 for i in range(n_out)]
                    elif fn == H_INVZERO:
                        outs = [pow(ins_v[0], -1, R) if ins_v[0] else 0]
                    elif fn == H_DECOMPOSE:
                        outs = [(ins_v[0] >> (param * i)) & ((1 << param) - 1) for i in range(n_out)]
                    elif fn == H_LOOKUP:
                        t0, t1 = int(flat["table_ptr"][param]), int(flat["table_ptr"][param + 1])
                        outs = []
                        for q in ins_v:
                            if q >= t1 - t0:
                                raise Unsatisfied(f"lookup index {q} outside table {param}")
                            outs.append(eval_aux(t0 + q))
                    elif fn == H_CMP:
                        outs = [(R - 1) if ins_v[0] < ins_v[1] else (0 if ins_v[0] == ins_v[1] else 1)]
                    else:
                        raise ValueError(f"unknown hint {fn}")
                assert len(outs) == n_out
                for k, v in enumerate(outs):
                    writes.append((first + k, v % R))
        for wid, val in writes:          # a level's results become visible to the next level only
            assert w[wid] is None, f"wire {wid} solved twice"
            w[wid] = val
    missing = [i for i, v in enumerate(w) if v is None]
    if missing:
        raise Unsatisfied(f"{len(missing)} wires never solved, first {missing[0]}")
    return w, a, b, c, info


def to_r1cs(flat):
    """the flat program as the oracle's groth16.R1CS (for Setup / Prove / check_in_exponent)"""
    import groth16 as g
    coeffs = flat["coeffs"]

    def rows(k):
        ptr, wi, ci = flat[k + "_row_ptr"], flat[k + "_wire"], flat[k + "_coeff"]
        return [[(coeffs[int(ci[e])], int(wi[e])) for e in range(int(ptr[r]), int(ptr[r + 1]))] for r in range(flat["n_constraints"])]

    cs = g.R1CS(flat["n_public"], flat["n_secret"], flat["n_wires"] - flat["n_public"] - flat["n_secret"], rows("l"), rows("r"), rows("o"),
                commitment_index=int(flat["commitment_index"]), private_committed=[int(x) for x in flat["private_committed"]])
    return cs
