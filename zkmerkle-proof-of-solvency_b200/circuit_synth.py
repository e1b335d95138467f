"""Synthetic BatchCreateUser-shaped constraint systems in the flat-array form `zkpor_program_desc` takes.

Why this exists.  The reference compiles circuit/batch_create_user_circuit.go with gnark's frontend (Go, not available
here), and gnark's solver then walks the compiled system: `Instructions` (blueprint id + calldata), `Levels`
(instruction ids that may run in parallel), a coefficient table and the hint functions (constraint/bn254/solver.go, out
of tree; reached from groth16.Prove, src/prover/prover/prover.go:269; hint registration prover.go:68).  The cgo shim
exports exactly those arrays ("solver contract"); this module produces arrays of the same form for a
circuit built from the same gadget set the reference circuit uses (SURVEY.md App. A "solver-visible primitive set"):

  api.Mul / Add / Sub / Select / IsZero / ToBinary / AssertIsBoolean / AssertIsEqual   circuit/utils.go:12-25,129,143
  IntegerDivision hint (DivMod by 100) + range checks                                   circuit/utils.go:103-110,166-177
  rangecheck.Check(v, {8,16,64,128})  -> limb decomposition hint + log-derivative argument
  logderivlookup.Table Insert / Lookup -> lookup hint + multiplicity hint + log-derivative argument
  the BSB22 commitment placeholder hint (challenge = hash(Pedersen commitment))          batch_create_user_circuit.go:110,112
  poseidon.Poseidon (x^5, R_F = 8, 12 inputs per permutation, lane 0 chains)            circuit/utils.go:19,48

It is a workload generator and test fixture, not part of the proving path; it does no field arithmetic beyond building
coefficients and does not import anything under oracle/.

Structure.  A circuit is a sequence of SECTIONS; a section is a template recorded once and instantiated `count` times
(the reference circuit is exactly that: one block per user, one per CEX asset, plus global parts), optionally chained
(copy k reads the carry wires of copy k-1: the 10 000-element CEX commitment is a chain of 834 permutations).  Templates
are recorded with Python ints; `flatten()` expands them with array broadcasting only (numpy here, torch on the GPU for
the 2^26 bench), so a 65 M-constraint system takes seconds to materialise.

Wire order follows gnark: [ONE, public inputs | secret inputs | internal wires].
"""
from __future__ import annotations

import numpy as np

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617

# instruction kinds / hint functions (mirrored in include/zkpor_b200.h)
INS_R1C, INS_HINT = 0, 1
H_DIVMOD, H_NBITS, H_INVZERO, H_DECOMPOSE, H_LOOKUP, H_CMP, H_COUNT, H_COMMIT = 1, 2, 3, 4, 5, 6, 7, 8
HINT_NAMES = {H_DIVMOD: "IntegerDivision", H_NBITS: "bits.NBits", H_INVZERO: "InvZero", H_DECOMPOSE: "rangecheck.Decompose",
              H_LOOKUP: "logderivlookup.Lookup", H_CMP: "CmpNOp", H_COUNT: "logderivarg.count", H_COMMIT: "Bsb22CommitmentComputePlaceholder"}

SP_PUBLIC, SP_SECRET, SP_INTERNAL = 0, 1, 2
MODE_SAME, MODE_PREV, MODE_FIXED = 0, 1, 2
RANGE_LIMB_BITS = 16                      # rangecheck's limb width: table 0 .. 2^16-1
LE_CHUNK = 256                            # terms per partial sum when a long sum is materialised (gnark compresses long expressions too)


def ref(space, mode, sec, fcopy, local):
    return (((space * 4 + mode) * 4096 + sec) << 48) | (fcopy << 24) | local


def unref(code):
    hi = code >> 48
    return (hi // 4096) // 4, (hi // 4096) % 4, hi % 4096, (code >> 24) & 0xFFFFFF, code & 0xFFFFFF


ONE = ref(SP_PUBLIC, MODE_FIXED, 0, 0, 0)


class LE:
    """linear expression: {wire reference: coefficient mod r}"""
    __slots__ = ("t",)

    def __init__(self, t=None):
        self.t = t if t is not None else {}

    @staticmethod
    def const(k):
        k %= R
        return LE({ONE: k} if k else {})

    def copy(self):
        return LE(dict(self.t))

    def is_const(self):
        return all(k == ONE for k in self.t)

    def const_value(self):
        return self.t.get(ONE, 0)

    def __add__(self, o):
        o = as_le(o)
        a, b = (self.t, o.t) if len(self.t) >= len(o.t) else (o.t, self.t)
        t = dict(a)
        for k, v in b.items():
            nv = (t.get(k, 0) + v) % R
            if nv:
                t[k] = nv
            else:
                t.pop(k, None)
        return LE(t)

    __radd__ = __add__

    def __neg__(self):
        return LE({k: R - v for k, v in self.t.items()})

    def __sub__(self, o):
        return self + (-as_le(o))

    def __rsub__(self, o):
        return as_le(o) - self

    def __mul__(self, k):
        assert isinstance(k, int)
        k %= R
        if k == 0:
            return LE()
        if k == 1:
            return self
        return LE({w: v * k % R for w, v in self.t.items()})

    __rmul__ = __mul__


def as_le(x):
    return x if isinstance(x, LE) else LE.const(x)


def le_sum(xs):
    t = {}
    for x in xs:
        for k, v in as_le(x).t.items():
            t[k] = t.get(k, 0) + v
    return LE({k: v % R for k, v in t.items() if v % R})


class Section:
    def __init__(self, idx, name, count, chain):
        self.idx, self.name, self.count, self.chain = idx, name, count, chain
        self.n_secret = 0
        self.n_int = 0
        self.secret_specs = []          # (kind, param)
        self.rows = []                  # (L, R, O) dicts
        self.instr = []                 # (kind, local arg, level)
        self.hints = []                 # (fn, param, [LE], out_first_local, n_out)
        self.wire_level = {}            # local internal wire -> level (absolute; relative to the copy for chain sections)
        self.queries = []               # range-table queries (LE)
        self.lookups = {}               # table id -> [(index LE, result LE)]
        self.committed = []             # local internal wires that go into the commitment
        self.carry_init = {}            # chain: local wire -> LE used instead of the previous copy's wire in copy 0 (constants / fixed refs)
        self.ext_level = -1             # chain: highest level among the wires read from outside the copy
        self.max_rel = -1
        self.base0, self.step = 0, 0    # chain: level of (copy k, rel) = base0 + k*step + rel


class Table:
    def __init__(self, tid, entries):
        self.id, self.entries = tid, entries


class CircuitBuilder:
    """records sections; `poseidon_constants(t) -> (round_constants[(8+rp)*t], mds[t][t], rp)` supplies the permutation
    parameters (the product library's own table in bench.py, the oracle's in the parity tests)."""

    def __init__(self, n_public_inputs=1, poseidon_constants=None, limb_bits=RANGE_LIMB_BITS):
        self.limb_bits = limb_bits      # rangecheck limb width (16 in gnark for these sizes; smaller tables for small tests)
        self.sections = []
        self.n_public = 1 + n_public_inputs
        self.poseidon_constants = poseidon_constants
        self.tables = []
        self.cur = None
        self.commit_wire = None
        self.coeff_ids = {}
        self.coeffs = []
        self.special = []               # (fn, section idx, hint local idx) of COUNT / COMMIT hints
        self.count_specs = {}           # (section idx, hint idx) -> ("range", None) | ("table", tid)
        for v in (0, 1, R - 1):
            self.coeff_id(v)
        self.section("header", 1)

    # ------------------------------------------------------------------ bookkeeping
    def coeff_id(self, v):
        i = self.coeff_ids.get(v)
        if i is None:
            i = self.coeff_ids[v] = len(self.coeffs)
            self.coeffs.append(v)
        return i

    def section(self, name, count=1, chain=False):
        s = Section(len(self.sections), name, count, chain)
        if self.cur is not None:
            self._close(self.cur)
        self.sections.append(s)
        self.cur = s
        return s

    def _close(self, s):
        if s.chain:
            s.base0 = s.ext_level + 1
            s.step = s.max_rel + 1

    def public(self, i):
        """i-th public input (wire 1 + i)"""
        assert 0 <= i < self.n_public - 1
        return LE({ref(SP_PUBLIC, MODE_FIXED, 0, 0, 1 + i): 1})

    def secret(self, kind="field", param=0):
        """a secret input of the current section (one per copy).  kind/param describe how satisfying values are drawn:
        ("bool",), ("uint", bits), ("below", n), ("field",), ("fixed", value), ("copy", offset): value = copy index + offset"""
        s = self.cur
        s.secret_specs.append((kind, param))
        s.n_secret += 1
        return LE({ref(SP_SECRET, MODE_SAME if s.count > 1 else MODE_FIXED, s.idx, 0, s.n_secret - 1): 1})

    def _new_wire(self):
        s = self.cur
        s.n_int += 1
        return ref(SP_INTERNAL, MODE_SAME if s.count > 1 else MODE_FIXED, s.idx, 0, s.n_int - 1)

    def at(self, le, copy):
        """the same expression read from a fixed copy of the (repeated, earlier) section it lives in"""
        out = {}
        for code, v in le.t.items():
            sp, mode, sec, _, loc = unref(code)
            if mode == MODE_SAME:
                assert copy < self.sections[sec].count
                code = ref(sp, MODE_FIXED, sec, copy, loc)
            out[code] = v
        return LE(out)

    def prev(self, le, init):
        """inside a chain section: the value `le` had in the previous copy (`init` in copy 0; a constant or fixed wires)"""
        s = self.cur
        assert s.chain and len(le.t) == 1
        (code, v), = le.t.items()
        sp, mode, sec, _, loc = unref(code)
        assert v == 1 and sp == SP_INTERNAL and sec == s.idx and mode == MODE_SAME
        s.carry_init[loc] = as_le(init)
        return LE({ref(SP_INTERNAL, MODE_PREV, s.idx, 0, loc): 1})

    def level_of(self, code):
        sp, mode, sec, fcopy, loc = unref(code)
        if sp != SP_INTERNAL:
            return -1
        s, cur = self.sections[sec], self.cur
        if sec == cur.idx and mode == MODE_SAME or (sec == cur.idx and cur.count == 1):
            return cur.wire_level[loc]
        if mode == MODE_PREV:
            return -1                   # relative to the copy; the chain step accounts for it
        assert sec < cur.idx or (sec == cur.idx and mode == MODE_FIXED), "forward reference"
        lvl = s.wire_level[loc]
        if s.chain:
            lvl = s.base0 + (fcopy if mode == MODE_FIXED else 0) * s.step + lvl
            assert mode == MODE_FIXED, "same-copy reference into a chain section"
        elif mode == MODE_SAME:
            assert s.count == cur.count, "same-copy reference between sections of different counts"
        if cur.chain:
            cur.ext_level = max(cur.ext_level, lvl)
            return -1
        return lvl

    def _instr_level(self, les, skip=None):
        lvl = -1
        for le in les:
            for code in le.t:
                if code != skip:
                    lvl = max(lvl, self.level_of(code))
        lvl += 1
        self.cur.max_rel = max(self.cur.max_rel, lvl)
        return lvl

    def r1c(self, L, Rr, O, defines=None):
        """one constraint L * R = O; `defines` = the wire this instruction solves for (None: an assertion)"""
        s = self.cur
        L, Rr, O = as_le(L), as_le(Rr), as_le(O)
        lvl = self._instr_level((L, Rr, O), skip=defines)
        if defines is not None:
            s.wire_level[unref(defines)[4]] = lvl
        s.rows.append((L.t, Rr.t, O.t))
        s.instr.append((INS_R1C, len(s.rows) - 1, lvl))

    def hint(self, fn, n_out, inputs, param=0, after_level=-1):
        """`after_level`: the hint also reads wires solved at that level (a lookup reads its table's entries)"""
        s = self.cur
        inputs = [as_le(x) for x in inputs]
        lvl = self._instr_level(inputs)
        if s.chain:
            s.ext_level = max(s.ext_level, after_level)
        else:
            lvl = max(lvl, after_level + 1)
            s.max_rel = max(s.max_rel, lvl)
        outs = [self._new_wire() for _ in range(n_out)]
        for w in outs:
            s.wire_level[unref(w)[4]] = lvl
        s.hints.append((fn, param, inputs, unref(outs[0])[4] if outs else 0, n_out))
        s.instr.append((INS_HINT, len(s.hints) - 1, lvl))
        return [LE({w: 1}) for w in outs]

    # ------------------------------------------------------------------ api.* gadgets
    def mul(self, a, b):
        a, b = as_le(a), as_le(b)
        if a.is_const():
            return b * a.const_value()
        if b.is_const():
            return a * b.const_value()
        w = self._new_wire()
        self.r1c(a, b, LE({w: 1}), defines=w)
        return LE({w: 1})

    def materialise(self, a):
        """a new wire equal to the expression (one constraint) -- what gnark does when an expression grows too long"""
        w = self._new_wire()
        self.r1c(a, LE.const(1), LE({w: 1}), defines=w)
        return LE({w: 1})

    def long_sum(self, xs):
        """sum of many terms as a chain of partial sums of LE_CHUNK terms"""
        acc = LE()
        for i in range(0, len(xs), LE_CHUNK):
            acc = le_sum([acc] + list(xs[i:i + LE_CHUNK]))
            if i + LE_CHUNK < len(xs):
                acc = self.materialise(acc)
        return acc

    def assert_eq(self, a, b):
        self.r1c(a, LE.const(1), b)

    def assert_bool(self, a):
        a = as_le(a)
        self.r1c(a, 1 - a, LE())

    def select(self, c, x, y):
        x, y = as_le(x), as_le(y)
        w = self._new_wire()
        self.r1c(c, x - y, LE({w: 1}) - y, defines=w)
        return LE({w: 1})

    def inverse(self, a, num=1):
        """DivUnchecked(num, a): the unknown sits in L, the solver divides"""
        w = self._new_wire()
        self.r1c(LE({w: 1}), a, as_le(num), defines=w)
        return LE({w: 1})

    def is_zero(self, a):
        a = as_le(a)
        m, = self.hint(H_INVZERO, 1, [a])
        w = self._new_wire()
        self.r1c(a, m, 1 - LE({w: 1}), defines=w)      # w = 1 - a*m : coefficient -1 on the unknown
        self.r1c(a, LE({w: 1}), LE())
        return LE({w: 1})

    def to_binary(self, a, n):
        bits = self.hint(H_NBITS, n, [a])
        for b in bits:
            self.assert_bool(b)
        self.assert_eq(le_sum(b * (1 << i) for i, b in enumerate(bits)), a)
        return bits

    def range_check(self, v, bits):
        """rangecheck.Check: limbs of RANGE_LIMB_BITS bits, every limb (and the scaled top limb when bits is not a multiple of the
        limb width) becomes a query of the range table"""
        s, lb = self.cur, self.limb_bits
        n_limbs = (bits + lb - 1) // lb
        limbs = self.hint(H_DECOMPOSE, n_limbs, [v], lb)
        self.assert_eq(le_sum(l * (1 << (lb * i)) for i, l in enumerate(limbs)), v)
        for l in limbs:
            s.queries.append(l)
            s.committed.append(unref(next(iter(l.t)))[4])
        rem = bits % lb
        if rem:
            s.queries.append(limbs[-1] * (1 << (lb - rem)))

    def divmod_const(self, x, d):
        """checkAndGetIntegerDivisionRes (circuit/utils.go:166-177)"""
        q, rem = self.hint(H_DIVMOD, 2, [x, LE.const(d)])
        self.range_check(q, 128)
        self.range_check(rem, 8)
        c, = self.hint(H_CMP, 1, [rem, LE.const(d)])
        self.assert_eq(c, LE.const(-1))
        self.assert_eq(q * d + rem, x)
        return q

    def new_table(self, entries):
        assert self.cur.count == 1
        t = Table(len(self.tables), [as_le(e) for e in entries])
        t.level = max([self.level_of(code) for e in t.entries for code in e.t] + [-1])
        self.tables.append(t)
        return t

    def lookup(self, table, idx):
        s = self.cur
        res, = self.hint(H_LOOKUP, 1, [idx], table.id, after_level=table.level)
        s.lookups.setdefault(table.id, []).append((as_le(idx), res))
        s.committed.append(unref(next(iter(res.t)))[4])
        return res

    # ------------------------------------------------------------------ Poseidon gadget
    def permute(self, state):
        t = len(state)
        rc, mds, rp = self.poseidon_constants(t)
        half = 4
        k = 0
        for rnd in range(8 + rp):
            state = [x + rc[k + i] for i, x in enumerate(state)]
            k += t
            full = rnd < half or rnd >= half + rp
            for i in range(t if full else 1):
                x = state[i]
                x2 = self.mul(x, x)
                x4 = self.mul(x2, x2)
                state[i] = self.mul(x4, x)
            state = [le_sum(state[j] * mds[i][j] for j in range(t)) for i in range(t)]
        return state

    def poseidon(self, inputs, out_lane=1):
        """poseidon.Poseidon(api, inputs...): 12 inputs per permutation, lane 0 chains, width rem+1 for the tail"""
        st0, i, n, st = LE(), 0, len(inputs), None
        while n - i >= 12:
            st = self.permute([st0] + [as_le(x) for x in inputs[i:i + 12]])
            st0 = st[0]
            i += 12
        if i < n:
            st = self.permute([st0] + [as_le(x) for x in inputs[i:]])
        return st[out_lane]

    # ------------------------------------------------------------------ the commitment and the log-derivative arguments
    def finalize_lookups(self):
        """What gnark defers to the end of Compile: multiplicity hints, ONE commitment over every table entry, query and
        multiplicity, and one log-derivative argument per table.  Opens its own sections."""
        secs = [s for s in self.sections if s.queries or s.lookups]
        range_bits = self.limb_bits
        # -- multiplicities
        cnt = self.section("multiplicities", 1)
        m_range = None
        if any(s.queries for s in secs):
            m_range = self.hint(H_COUNT, 1 << range_bits, [], 1 << range_bits)
            self.count_specs[(cnt.idx, len(cnt.hints) - 1)] = ("range", None)
            cnt.committed += [unref(next(iter(m.t)))[4] for m in m_range]
        m_tab = {}
        for t in self.tables:
            if any(t.id in s.lookups for s in secs):
                m_tab[t.id] = self.hint(H_COUNT, len(t.entries), [], len(t.entries))
                self.count_specs[(cnt.idx, len(cnt.hints) - 1)] = ("table", t.id)
                cnt.committed += [unref(next(iter(m.t)))[4] for m in m_tab[t.id]]
        # the COUNT hints read every query: their level is above every section that queries
        top = max([lv for s in secs for lv in ([s.base0 + s.count * s.step] if s.chain else s.wire_level.values())] + [-1]) + 1
        for i, (kind, arg, lvl) in enumerate(cnt.instr):
            cnt.instr[i] = (kind, arg, top)
        for k in cnt.wire_level:
            cnt.wire_level[k] = top
        # -- commitment: the placeholder hint's output is the challenge wire
        self.section("commitment", 1)
        c, = self.hint(H_COMMIT, 1, [LE({ref(SP_INTERNAL, MODE_FIXED, cnt.idx, 0, cnt.n_int - 1): 1})] if cnt.n_int else [])
        self.commit_wire = next(iter(c.t))
        c2 = self.mul(c, c)                     # second challenge: folds (index, value) rows of the lookup tables
        # -- table sides
        self.section("table_sides", 1)
        sides = []
        if m_range is not None:
            sides.append(self.long_sum([self.inverse(c - i, m_range[i]) for i in range(1 << range_bits)]))
        for t in self.tables:
            if t.id in m_tab:
                rows = [self.inverse(c - i - self.mul(c2, e), m_tab[t.id][i]) for i, e in enumerate(t.entries)]
                sides.append(self.long_sum(rows))
        # -- query sides: one repeated section per section that queries, same count, same-copy references
        totals = []
        for s in secs:
            q = self.section(s.name + "_queries", s.count)
            terms = [self.inverse(c - qe) for qe in s.queries]
            for tid, pairs in sorted(s.lookups.items()):
                terms += [self.inverse(c - idx - self.mul(c2, res)) for idx, res in pairs]
            part = self.materialise(self.long_sum(terms))
            totals.append((q, part))
        # -- the check: sum over the table sides = sum over every copy's partial sum
        self.section("logderiv_check", 1)
        rhs = []
        for q, part in totals:
            rhs += [self.at(part, k) for k in range(q.count)] if q.count > 1 else [part]
        self.assert_eq(le_sum(sides), self.long_sum(rhs))

    # ------------------------------------------------------------------ flat arrays
    def flatten(self, xp=np, device=None):
        """-> dict of flat arrays (zkpor_program_desc fields).  xp = numpy, or torch with `device` for the full-size bench."""
        self._close(self.cur)
        F = _Xp(xp, device)
        secs = self.sections
        n_pub = self.n_public
        sec_base = np.zeros(len(secs) + 1, dtype=np.int64)
        int_base = np.zeros(len(secs) + 1, dtype=np.int64)
        for s in secs:
            sec_base[s.idx + 1] = sec_base[s.idx] + s.n_secret * s.count
            int_base[s.idx + 1] = int_base[s.idx] + s.n_int * s.count
        n_secret = int(sec_base[-1])
        n_wires = n_pub + n_secret + int(int_base[-1])

        def resolve(code):
            """(base id for copy 0, stride per copy, copy-0 override or -1) of a reference seen from a section template"""
            sp, mode, sec, fcopy, loc = unref(code)
            s = secs[sec]
            if sp == SP_PUBLIC:
                return loc, 0
            base, n = (n_pub + sec_base[sec], s.n_secret) if sp == SP_SECRET else (n_pub + n_secret + int_base[sec], s.n_int)
            if mode == MODE_FIXED:
                return int(base + fcopy * n + loc), 0
            if mode == MODE_SAME:
                return int(base + loc), n
            return int(base + loc - n), n            # MODE_PREV

        out = {"n_wires": n_wires, "n_public": n_pub, "n_secret": n_secret}
        mats = {m: {"ptr": [], "wire": [], "coeff": []} for m in "LRO"}
        aux = {"ptr": [], "wire": [], "coeff": []}
        instr_kind, instr_arg, instr_level = [], [], []
        hint_fn, hint_param, hint_out, hint_nout, hint_inptr = [], [], [], [], []
        committed = []
        row_base = hint_base = aux_row_base = 0
        nnz_base = {m: 0 for m in "LRO"}
        aux_nnz_base = 0
        table_ptr = [0]
        commit_abs = None
        deferred_counts = []

        def csr_template(les):
            """list of term dicts -> (row_ptr[r+1], base[nnz], stride[nnz], coeff[nnz], prev_fix {pos: LE})"""
            ptr, base, stride, coeff, prevpos = [0], [], [], [], []
            for terms in les:
                for code, v in terms.items():
                    b, st = resolve(code)
                    if unref(code)[1] == MODE_PREV:
                        prevpos.append((len(base), unref(code)[4], v))
                    base.append(b); stride.append(st); coeff.append(self.coeff_id(v))
                ptr.append(len(base))
            return (np.asarray(ptr, dtype=np.int64), np.asarray(base, dtype=np.int64), np.asarray(stride, dtype=np.int64),
                    np.asarray(coeff, dtype=np.int64), prevpos)

        def expand(tpl, count, nnz0, sec):
            """instantiate a CSR template `count` times -> (row starts [count*r], wire ids, coeff ids)"""
            ptr, base, stride, coeff, prevpos = tpl
            r, nnz = len(ptr) - 1, len(base)
            if count == 1 and not prevpos:
                return F.arr(ptr[:-1] + nnz0), F.arr(base), F.arr(coeff), nnz
            if prevpos:
                # copy 0 of a chain reads `carry_init` instead of the previous copy: only single-wire / constant initial values are
                # supported in place (same number of terms), which is all the sponge chains need
                base0 = base.copy(); coeff0 = coeff.copy()
                for pos, loc, v in prevpos:
                    init = sec.carry_init[loc]
                    assert len(init.t) <= 1, "chain initial value must be a constant or a single wire"
                    if init.t:
                        (icode, iv), = init.t.items()
                        base0[pos] = resolve(icode)[0]; coeff0[pos] = self.coeff_id(iv * v % R)
                    else:
                        base0[pos] = 0; coeff0[pos] = self.coeff_id(0)
                copies = F.arange(count)
                wires = F.arr(base)[None, :] + copies[:, None] * F.arr(stride)[None, :]
                wires[0, :] = F.arr(base0)
                coeffs = F.tile_rows(F.arr(coeff), count)
                coeffs[0, :] = F.arr(coeff0)
            else:
                copies = F.arange(count)
                wires = F.arr(base)[None, :] + copies[:, None] * F.arr(stride)[None, :]
                coeffs = F.tile_rows(F.arr(coeff), count)
            starts = F.arr(ptr[:-1])[None, :] + (copies * nnz)[:, None] + nnz0
            return starts.reshape(-1), wires.reshape(-1), coeffs.reshape(-1), nnz * count

        for s in secs:
            U = s.count
            copies = F.arange(U)
            # -- constraint rows
            for m, k in (("L", 0), ("R", 1), ("O", 2)):
                tpl = csr_template([row[k] for row in s.rows])
                st, w, cf, used = expand(tpl, U, nnz_base[m], s)
                mats[m]["ptr"].append(st); mats[m]["wire"].append(w); mats[m]["coeff"].append(cf)
                nnz_base[m] += used
            # -- hints and their inputs (rows of the auxiliary matrix)
            n_h = len(s.hints)
            in_counts = np.asarray([len(h[2]) for h in s.hints], dtype=np.int64)
            in_ptr_t = np.concatenate([[0], np.cumsum(in_counts)]).astype(np.int64)
            rows_per_copy = int(in_ptr_t[-1])
            tpl = csr_template([le.t for h in s.hints for le in h[2]])
            st, w, cf, used = expand(tpl, U, aux_nnz_base, s)
            aux["ptr"].append(st); aux["wire"].append(w); aux["coeff"].append(cf)
            aux_nnz_base += used
            if n_h:
                hint_fn.append(F.tile(F.arr(np.asarray([h[0] for h in s.hints], dtype=np.int64)), U))
                hint_param.append(F.tile(F.arr(np.asarray([h[1] for h in s.hints], dtype=np.int64)), U))
                hint_nout.append(F.tile(F.arr(np.asarray([h[4] for h in s.hints], dtype=np.int64)), U))
                ob = n_pub + n_secret + int(int_base[s.idx])
                hint_out.append((F.arr(np.asarray([h[3] for h in s.hints], dtype=np.int64))[None, :] + ob + copies[:, None] * s.n_int).reshape(-1))
                hint_inptr.append((F.arr(in_ptr_t[:-1])[None, :] + aux_row_base + copies[:, None] * rows_per_copy).reshape(-1))
            for hi, h in enumerate(s.hints):
                if h[0] == H_COUNT:
                    deferred_counts.append((hint_base + hi, self.count_specs[(s.idx, hi)]))
                if h[0] == H_COMMIT:
                    commit_abs = n_pub + n_secret + int(int_base[s.idx]) + h[3]
            # -- instructions
            if s.instr:
                kinds = np.asarray([i[0] for i in s.instr], dtype=np.int64)
                args = np.asarray([i[1] for i in s.instr], dtype=np.int64)
                lvls = np.asarray([i[2] for i in s.instr], dtype=np.int64)
                argbase = np.where(kinds == INS_R1C, row_base, hint_base)
                argstride = np.where(kinds == INS_R1C, len(s.rows), n_h)
                instr_kind.append(F.tile(F.arr(kinds), U))
                instr_arg.append((F.arr(args + argbase)[None, :] + copies[:, None] * F.arr(argstride)[None, :]).reshape(-1))
                lv = F.arr(lvls)[None, :] + (copies[:, None] * s.step + s.base0 if s.chain else 0)
                instr_level.append((lv + F.zeros((U, 1))).reshape(-1))
            if s.committed:
                cb_ = n_pub + n_secret + int(int_base[s.idx])
                committed.append((F.arr(np.asarray(sorted(set(s.committed)), dtype=np.int64))[None, :] + cb_ + copies[:, None] * s.n_int).reshape(-1))
            row_base += len(s.rows) * U
            hint_base += n_h * U
            aux_row_base += rows_per_copy * U

        # -- lookup tables: entry rows of the auxiliary matrix
        for t in self.tables:
            tpl = csr_template([e.t for e in t.entries])
            st, w, cf, used = expand(tpl, 1, aux_nnz_base, None)
            aux["ptr"].append(st); aux["wire"].append(w); aux["coeff"].append(cf)
            aux_nnz_base += used
            table_ptr.append(table_ptr[-1] + len(t.entries))
        table_row0 = aux_row_base
        aux_row_base += table_ptr[-1]

        # -- the COUNT hints' inputs: every query of every copy, as single rows appended to the auxiliary matrix
        count_in = {}
        for hid, (kind, tid) in deferred_counts:
            first_row = aux_row_base
            for s in secs:
                les = s.queries if kind == "range" else [idx for idx, _ in s.lookups.get(tid, [])]
                if not les:
                    continue
                tpl = csr_template([le.t for le in les])
                st, w, cf, used = expand(tpl, s.count, aux_nnz_base, s)
                aux["ptr"].append(st); aux["wire"].append(w); aux["coeff"].append(cf)
                aux_nnz_base += used
                aux_row_base += len(les) * s.count
            count_in[hid] = (first_row, aux_row_base)

        def cat(parts, dtype):
            parts = [p for p in parts if p is not None and len(p)]
            return F.cat(parts, dtype) if parts else F.empty(dtype)

        for m, key in (("L", "l"), ("R", "r"), ("O", "o")):
            out[key + "_row_ptr"] = F.cat(mats[m]["ptr"] + [F.arr(np.asarray([nnz_base[m]], dtype=np.int64))], "u64")
            out[key + "_wire"] = cat(mats[m]["wire"], "u32"); out[key + "_coeff"] = cat(mats[m]["coeff"], "u32")
        out["aux_row_ptr"] = F.cat(aux["ptr"] + [F.arr(np.asarray([aux_nnz_base], dtype=np.int64))], "u64")
        out["aux_wire"] = cat(aux["wire"], "u32"); out["aux_coeff"] = cat(aux["coeff"], "u32")
        out["n_constraints"] = row_base
        out["n_hints"] = hint_base
        hin = cat(hint_inptr, "i64")
        hfn = cat(hint_fn, "u32")
        # hint h reads aux rows [hint_in_ptr[h], hint_in_end[h]); COUNT hints point at their deferred rows
        hcnt = cat([F.tile(F.arr(np.asarray([len(h[2]) for h in s.hints], dtype=np.int64)), s.count) for s in secs if s.hints], "i64")
        hend = hin + hcnt
        for hid, (r0, r1) in count_in.items():
            hin[hid] = r0; hend[hid] = r1
        out["hint_fn"] = hfn; out["hint_param"] = cat(hint_param, "u32"); out["hint_out_first"] = cat(hint_out, "u32")
        out["hint_n_out"] = cat(hint_nout, "u32")
        out["hint_in_ptr"] = F.astype(hin, "u64"); out["hint_in_end"] = F.astype(hend, "u64")
        out["table_ptr"] = np.asarray(table_ptr, dtype=np.uint64) + np.uint64(table_row0)
        out["n_tables"] = len(self.tables)
        # -- levels: instruction ids grouped by level (stable)
        kind_all = cat(instr_kind, "u8"); arg_all = cat(instr_arg, "u32"); lvl_all = cat(instr_level, "i64")
        order, level_ptr = F.group_by_level(lvl_all)
        out["instr_kind"] = kind_all; out["instr_arg"] = arg_all
        out["level_instr"] = F.astype(order, "u32"); out["level_ptr"] = level_ptr
        out["n_instr"] = int(len(kind_all)); out["n_levels"] = int(len(level_ptr) - 1)
        out["coeffs"] = list(self.coeffs)
        cm = cat(committed, "i64")
        out["private_committed"] = F.sort(cm) if len(cm) else cm
        out["commitment_index"] = commit_abs if commit_abs is not None else -1
        out["secret_layout"] = [(int(n_pub + sec_base[s.idx]), s.n_secret, s.count, list(s.secret_specs)) for s in secs if s.n_secret]
        return out


class _Xp:
    """the handful of array operations flatten() needs, over numpy or torch"""

    def __init__(self, xp, device):
        self.xp, self.dev, self.is_np = xp, device, xp is np
        self.dt = {"u8": "uint8", "u32": "uint32", "u64": "uint64", "i64": "int64"}

    def arr(self, a):
        return a if self.is_np else self.xp.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def arange(self, n):
        return np.arange(n, dtype=np.int64) if self.is_np else self.xp.arange(n, dtype=self.xp.int64, device=self.dev)

    def zeros(self, shape):
        return np.zeros(shape, dtype=np.int64) if self.is_np else self.xp.zeros(shape, dtype=self.xp.int64, device=self.dev)

    def tile(self, a, n):
        return np.tile(a, n) if self.is_np else a.repeat(n)

    def tile_rows(self, a, n):
        return np.tile(a[None, :], (n, 1)) if self.is_np else a[None, :].repeat(n, 1)

    def empty(self, dtype):
        return self.astype(self.zeros((0,)), dtype)

    def astype(self, a, dtype):
        if self.is_np:
            return a.astype(self.dt[dtype])
        # torch has no unsigned 32/64-bit arithmetic: values are non-negative, so int32 / int64 share the bit patterns
        t = {"u8": self.xp.uint8, "u32": self.xp.int32, "u64": self.xp.int64, "i64": self.xp.int64}[dtype]
        return a.to(t)

    def cat(self, parts, dtype):
        a = np.concatenate([np.asarray(p).reshape(-1) for p in parts]) if self.is_np else self.xp.cat([p.reshape(-1) for p in parts])
        return self.astype(a, dtype)

    def sort(self, a):
        return np.sort(a) if self.is_np else self.xp.sort(a).values

    def group_by_level(self, lvl):
        if self.is_np:
            order = np.argsort(lvl, kind="stable")
            counts = np.bincount(lvl.astype(np.int64), minlength=int(lvl.max()) + 1 if len(lvl) else 0)
            return order, np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        order = self.xp.sort(lvl, stable=True).indices
        counts = self.xp.bincount(lvl, minlength=int(lvl.max().item()) + 1)
        ptr = self.xp.cat([self.xp.zeros(1, dtype=self.xp.int64, device=self.dev), self.xp.cumsum(counts, 0)])
        return order, ptr.cpu().numpy().astype(np.uint64)


# ---------------------------------------------------------------------------------------------------------------- inputs
def _splitmix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def draw_inputs(flat, seed, public_values=None):
    """A satisfying assignment of the inputs: (n_public - 1 + n_secret) canonical values as uint64[.., 4] little-endian limbs.
    Every gadget of the builder is satisfiable for any value its input kind allows, so no native evaluation is needed."""
    n_pub, n_sec = flat["n_public"], flat["n_secret"]
    out = np.zeros((n_pub - 1 + n_sec, 4), dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(n_pub - 1):
            v = (public_values[i] if public_values is not None else int(_splitmix(np.uint64(seed * 1000003 + i))) | (1 << 200)) % R
            out[i] = [(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]
        for first, n_s, count, specs in flat["secret_layout"]:
            base = first - 1                    # position in the input vector (the ONE wire is not an input)
            copies = np.arange(count, dtype=np.uint64)
            for j, (kind, param) in enumerate(specs):
                pos = base + copies.astype(np.int64) * n_s + j
                ctr = (np.uint64(seed) << np.uint64(40)) + (pos.astype(np.uint64) << np.uint64(2))
                h0 = _splitmix(ctr)
                if kind == "bool":
                    out[pos, 0] = h0 & np.uint64(1)
                elif kind == "uint":
                    bits = int(param)
                    limbs = [_splitmix(ctr + np.uint64(k)) for k in range(4)]
                    for k in range(4):
                        nb = min(64, max(0, bits - 64 * k))
                        if nb:
                            out[pos, k] = limbs[k] if nb == 64 else limbs[k] & np.uint64((1 << nb) - 1)
                elif kind == "below":
                    out[pos, 0] = h0 % np.uint64(param)
                elif kind == "fixed":
                    v = int(param) % R
                    out[pos] = [(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]
                elif kind == "copy":
                    out[pos, 0] = copies + np.uint64(param)
                else:                           # "field": 253 uniform bits (below r)
                    for k in range(4):
                        out[pos, k] = _splitmix(ctr + np.uint64(k))
                    out[pos, 3] &= np.uint64((1 << 61) - 1)
    return out


# ---------------------------------------------------------------------------------------------------------------- circuits
def batch_create_user_like(users, assets_per_user=4, cex_assets=6, tiers=3, merkle_depth=4, chain_perms=3, poseidon_constants=None,
                           limb_bits=RANGE_LIMB_BITS):
    """A circuit with the structure of BatchCreateUserCircuit.Define (circuit/batch_create_user_circuit.go:99-321), every size a
    parameter.  The reference tier-50 batch is users=1380, assets_per_user=50, cex_assets=500, tiers=12, merkle_depth=28,
    chain_perms=834.  Returns the builder (call .flatten())."""
    cb = CircuitBuilder(1, poseidon_constants, limb_bits)
    batch_commitment = cb.public(0)

    # ---- global part: CEX asset info, tier tables, collateral precomputation (batch_create_user_circuit.go:99-134)
    cb.section("cex_assets", cex_assets)
    price = cb.secret("uint", 32)
    tot_eq, tot_debt = cb.secret("uint", 60), cb.secret("uint", 60)
    for v in (price, tot_eq, tot_debt):
        cb.range_check(v, 64)
    bnd_prev, pre_prev = LE(), LE()
    tier_rows = []
    for j in range(tiers):
        dbnd = cb.secret("uint", 100)           # boundary increments keep the boundaries sorted
        ratio = cb.secret("below", 101)
        bnd = bnd_prev + dbnd
        cur = cb.divmod_const(cb.mul(dbnd, ratio), 100)      # generateRapidArithmeticForCollateral, circuit/utils.go:83-101
        pre = pre_prev + cur
        cb.range_check(pre, 128); cb.range_check(ratio, 8); cb.range_check(bnd, 128)
        tier_rows.append((cb.materialise(bnd), ratio, cb.materialise(pre)))
        bnd_prev, pre_prev = bnd, pre
    asset_sec = cb.cur
    packed = cb.materialise(tot_eq * (1 << 128) + tot_debt * (1 << 64) + price)

    # ---- tables (constructLoanTierRatiosLookupTable ..., circuit/utils.go:179-225): 3 dummy entries + 3 per tier, per asset
    cb.section("tables", 1)
    entries, prices = [], []
    for a in range(cex_assets):
        entries += [LE(), LE(), LE()]
        for bnd, ratio, pre in tier_rows:
            entries += [cb.at(bnd, a), cb.at(ratio, a), cb.at(pre, a)]
        prices.append(cb.at(price, a))
    tier_table = cb.new_table(entries)
    price_table = cb.new_table(prices)
    per_asset = 3 * (tiers + 1)

    # ---- the CEX assets commitment: a sponge over packed values, serial (witness.go:159-166; circuit :129)
    absorb = [cb.at(packed, a % cex_assets) + a for a in range(12)]
    cb.section("cex_commitment", chain_perms, chain=True)
    lane0 = LE({cb._new_wire(): 1})              # carry wire, defined below by the last constraint of the copy
    carry_in = cb.prev(lane0, 0)
    st = cb.permute([carry_in] + absorb)
    w0 = next(iter(lane0.t))
    cb.r1c(st[0], LE.const(1), lane0, defines=w0)
    chain_sec = cb.cur

    # ---- one block per user (batch_create_user_circuit.go:140-273)
    cb.section("users", users)
    index = cb.secret("copy", 0)
    id_hash = cb.secret("field")
    helper = cb.to_binary(index, merkle_depth)                       # accountIdToMerkleHelper, circuit/utils.go:23-26
    flat_assets, total_eq, total_debt, total_col = [], LE(), LE(), LE()
    for a in range(assets_per_user):
        aidx = cb.secret("below", cex_assets)
        eq, debt = cb.secret("uint", 50), cb.secret("uint", 40)
        col = [cb.secret("uint", 30) for _ in range(3)]
        for v in [eq, debt] + col:
            cb.range_check(v, 64)
        p = cb.lookup(price_table, aidx)
        total_eq += cb.mul(eq, p); total_debt += cb.mul(debt, p)
        for kind, c_ in enumerate(col):                               # getAndCheckTierRatiosQueryResults, circuit/utils.go:112-164
            tier_i = cb.secret("below", tiers)
            flag = cb.secret("bool")
            cb.assert_bool(flag)
            cv = cb.mul(c_, p)
            z = cb.is_zero(cv)
            start = aidx * per_asset + tier_i * 3
            res = [cb.lookup(tier_table, start + k) for k in range(6)]
            diff = cb.select(flag, cv + res[3], res[3] + cv)          # both branches in range for any input
            cb.range_check(cb.select(z, 0, diff), 128)
            q = cb.divmod_const(cb.mul(cv + res[0], res[4]), 100)
            total_col += cb.select(cb.is_zero(flag), res[2] + q, res[5])
        flat_assets += [aidx, eq, debt] + col
    packed_assets = [flat_assets[3 * i] * (1 << 128) + flat_assets[3 * i + 1] * (1 << 64) + flat_assets[3 * i + 2] for i in range(len(flat_assets) // 3)]
    commitment = cb.poseidon(packed_assets)                           # computeUserAssetsCommitment, circuit/utils.go:28-49
    node = cb.poseidon([id_hash, total_eq, total_debt, total_col, commitment])
    for lvl in range(merkle_depth):                                   # verifyMerkleProof, circuit/utils.go:12-21
        sib = cb.secret("field")
        d1 = cb.select(helper[lvl], sib, node)
        d2 = cb.select(helper[lvl], node, sib)
        node = cb.poseidon([d1, d2])
    user_root = cb.materialise(node)
    user_sec = cb.cur

    # ---- tail: batch commitment over the last user's root, the chain's end and the public input
    cb.section("tail", 1)
    last = LE({ref(SP_INTERNAL, MODE_FIXED, chain_sec.idx, chain_perms - 1, unref(w0)[4]): 1})
    roots = cb.long_sum([cb.at(user_root, k) for k in range(users)])
    cb.poseidon([roots, last, batch_commitment, cb.at(packed, 0), 7])
    cb.finalize_lookups()
    return cb
