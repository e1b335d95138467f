package zkporb200

/*
#include "zkpor_b200.h"
*/
import "C"

import (
	"fmt"
	"unsafe"

	"github.com/binance/zkmerkle-proof-of-solvency/circuit"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark/constraint"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
	"github.com/consensys/gnark/constraint/solver"
	"github.com/consensys/gnark/std/internal/logderivarg"
	"github.com/consensys/gnark/std/math/bits"
	"github.com/consensys/gnark/std/rangecheck"
)

// Program is gnark's compiled constraint system flattened into the arrays zkpor_program_upload takes (include/zkpor_b200.h,
// "witness solver"): the three R1CS matrices in CSR form over cs.Coefficients, cs.Instructions as (kind, arg), cs.Levels as
// level_ptr / level_instr, and one record per hint instruction whose inputs are rows of an auxiliary CSR matrix.
// Built once per tier after r1cs.ReadFrom (src/prover/prover/prover.go:317-327).
type Program struct {
	h   *C.zkpor_program
	ctx *Ctx
	nIn int // public (without ONE) + secret inputs
}

type csr struct {
	ptr        []uint64
	wire, coef []uint32
}

func (m *csr) add(le constraint.LinearExpression) {
	for _, t := range le {
		m.wire = append(m.wire, t.VID)
		m.coef = append(m.coef, t.CID)
	}
	m.ptr = append(m.ptr, uint64(len(m.wire)))
}

func (m *csr) c() C.zkpor_csr {
	var out C.zkpor_csr
	out.nnz = C.uint64_t(len(m.wire))
	out.row_ptr = (*C.uint64_t)(unsafe.Pointer(&m.ptr[0]))
	if len(m.wire) > 0 {
		out.wire_ids = (*C.uint32_t)(unsafe.Pointer(&m.wire[0]))
		out.coeff_ids = (*C.uint32_t)(unsafe.Pointer(&m.coef[0]))
	}
	return out
}

// hint functions the BatchCreateUser circuit reaches, by gnark hint id -> library function id.
// The reference registers its own IntegerDivision at src/prover/prover/prover.go:68 (circuit/utils.go:103-110); the others come in
// through api.ToBinary, api.IsZero, rangecheck.Check, logderivlookup and the fork's api.CmpNOp.
func hintTable() map[solver.HintID]uint32 {
	return map[solver.HintID]uint32{
		solver.GetHintID(circuit.IntegerDivision):     C.ZKPOR_HINT_DIVMOD,
		solver.GetHintID(bits.GetHints()[1]):          C.ZKPOR_HINT_NBITS, // bits.nBits
		solver.GetHintID(solver.InvZeroHint):          C.ZKPOR_HINT_INVZERO,
		solver.GetHintID(rangecheck.DecomposeHint):    C.ZKPOR_HINT_DECOMPOSE,
		solver.GetHintID(logderivarg.GetHints()[0]):   C.ZKPOR_HINT_COUNT, // countHint
		solver.GetHintID(circuit.CmpNOpHint):          C.ZKPOR_HINT_CMP,
	}
}

// Flatten walks the compiled system once.  R1C instructions map to their constraint row; hint instructions become records
// (function, parameter, first output wire, number of outputs, input rows in `aux`); the logderivlookup blueprint becomes one
// LOOKUP record per query batch with its table registered in table_ptr; the BSB22 commitment placeholder becomes COMMIT.
func (c *Ctx) Flatten(r1cs *cs_bn254.R1CS) (*Program, error) {
	var l, r, o, aux csr
	l.ptr, r.ptr, o.ptr, aux.ptr = []uint64{0}, []uint64{0}, []uint64{0}, []uint64{0}
	rows := r1cs.GetR1Cs()
	for _, rc := range rows {
		l.add(rc.L)
		r.add(rc.R)
		o.add(rc.O)
	}
	hints := hintTable()
	commitID := solver.GetHintID(bsb22Placeholder)
	nIns := len(r1cs.Instructions)
	kind := make([]uint8, nIns)
	arg := make([]uint32, nIns)
	var hFn, hParam, hOut, hNOut []uint32
	var hIn0, hIn1 []uint64
	tablePtr := []uint64{}
	row := uint32(0)
	for i, pi := range r1cs.Instructions {
		ins := pi.Unpack(&r1cs.System)
		switch bp := r1cs.Blueprints[pi.BlueprintID].(type) {
		case constraint.BlueprintR1C: // generic R1C and the specialised variants all emit exactly one row
			kind[i], arg[i] = C.ZKPOR_INS_R1C, row
			row++
		case *constraint.BlueprintGenericHint:
			var hm constraint.HintMapping
			bp.DecompressHint(&hm, ins)
			fn, ok := hints[hm.HintID]
			if hm.HintID == commitID {
				fn, ok = C.ZKPOR_HINT_COMMIT, true
			}
			if !ok {
				return nil, fmt.Errorf("zkporb200: hint %s has no device implementation", solver.GetHintName(hm.HintID))
			}
			param := uint32(0)
			in := hm.Inputs
			if fn == C.ZKPOR_HINT_DECOMPOSE { // rangecheck passes (limb bits, number of limbs) as constant inputs ahead of the value
				param = constantOf(r1cs, in[0])
				in = in[2:]
			}
			if fn == C.ZKPOR_HINT_COMMIT { // inputs are the committed wires themselves: the key carries that list
				in = nil
			}
			hIn0 = append(hIn0, uint64(len(aux.ptr)-1))
			for _, le := range in {
				aux.add(le)
			}
			hIn1 = append(hIn1, uint64(len(aux.ptr)-1))
			hFn, hParam = append(hFn, fn), append(hParam, param)
			hOut, hNOut = append(hOut, hm.OutputRange.Start), append(hNOut, hm.OutputRange.End-hm.OutputRange.Start)
			kind[i], arg[i] = C.ZKPOR_INS_HINT, uint32(len(hFn)-1)
		case *constraint.BlueprintLookupHint:
			// table entries (linear expressions) first, then one row per query; the table is registered once per blueprint
			t, q := lookupRows(bp, ins)
			tablePtr = append(tablePtr, uint64(len(aux.ptr)-1))
			for _, le := range t {
				aux.add(le)
			}
			tablePtr = append(tablePtr, uint64(len(aux.ptr)-1))
			hIn0 = append(hIn0, uint64(len(aux.ptr)-1))
			for _, le := range q {
				aux.add(le)
			}
			hIn1 = append(hIn1, uint64(len(aux.ptr)-1))
			hFn, hParam = append(hFn, C.ZKPOR_HINT_LOOKUP), append(hParam, uint32(len(tablePtr)/2-1))
			hOut, hNOut = append(hOut, ins.WireOffset), append(hNOut, uint32(len(q)))
			kind[i], arg[i] = C.ZKPOR_INS_HINT, uint32(len(hFn)-1)
		default:
			return nil, fmt.Errorf("zkporb200: blueprint %T is not supported", bp)
		}
	}
	levelPtr := []uint64{0}
	var levelIns []uint32
	for _, lv := range r1cs.Levels {
		for _, id := range lv {
			levelIns = append(levelIns, uint32(id))
		}
		levelPtr = append(levelPtr, uint64(len(levelIns)))
	}

	var d C.zkpor_program_desc
	d.n_wires = C.uint64_t(r1cs.GetNbInternalVariables() + r1cs.GetNbPublicVariables() + r1cs.GetNbSecretVariables())
	d.n_public, d.n_secret = C.uint64_t(r1cs.GetNbPublicVariables()), C.uint64_t(r1cs.GetNbSecretVariables())
	d.n_constraints = C.uint64_t(len(rows))
	d.l, d.r, d.o = l.c(), r.c(), o.c()
	d.coeff_table, d.n_coeffs = unsafe.Pointer(&r1cs.Coefficients[0]), C.uint64_t(len(r1cs.Coefficients)) // []fr.Element, Montgomery
	d.n_instr = C.uint64_t(nIns)
	d.instr_kind, d.instr_arg = (*C.uint8_t)(&kind[0]), (*C.uint32_t)(&arg[0])
	d.n_levels = C.uint64_t(len(r1cs.Levels))
	d.level_ptr, d.level_instr = (*C.uint64_t)(&levelPtr[0]), (*C.uint32_t)(&levelIns[0])
	d.n_hints = C.uint64_t(len(hFn))
	if len(hFn) > 0 {
		d.hint_fn, d.hint_param = (*C.uint32_t)(&hFn[0]), (*C.uint32_t)(&hParam[0])
		d.hint_out_first, d.hint_n_out = (*C.uint32_t)(&hOut[0]), (*C.uint32_t)(&hNOut[0])
		d.hint_in_ptr, d.hint_in_end = (*C.uint64_t)(&hIn0[0]), (*C.uint64_t)(&hIn1[0])
	}
	d.n_aux_rows, d.aux = C.uint64_t(len(aux.ptr)-1), aux.c()
	// the library wants n_tables + 1 monotone offsets; tables are laid out back to back by construction
	tp := compactTablePtr(tablePtr)
	d.n_tables = C.uint64_t(len(tp) - 1)
	if len(tp) > 1 {
		d.table_ptr = (*C.uint64_t)(&tp[0])
	}
	p := &Program{ctx: c, nIn: r1cs.GetNbPublicVariables() - 1 + r1cs.GetNbSecretVariables()}
	if err := call(func() C.int32_t { return C.zkpor_program_upload(c.h, &d, &p.h) }); err != nil {
		return nil, err // e.g. "wire N is never solved by the schedule"
	}
	return p, nil
}

func (p *Program) Close() { C.zkpor_program_free(p.ctx.h, p.h); p.h = nil }

// constantOf evaluates a linear expression made of the ONE wire only (compile-time constants of a hint call).
func constantOf(r1cs *cs_bn254.R1CS, le constraint.LinearExpression) uint32 {
	var acc fr.Element
	for _, t := range le {
		acc.Add(&acc, &r1cs.Coefficients[t.CID])
	}
	return uint32(acc.Uint64())
}

// bsb22Placeholder is gnark's commitment hint (frontend/cs/r1cs: bsb22CommitmentComputePlaceholder); Prove overrides it with
// "Pedersen-commit the private committed wires, hash the point to the field" -- which is what ZKPOR_HINT_COMMIT does on the device.
var bsb22Placeholder solver.Hint = nil // resolved by name at init: solver.GetRegisteredHints(), "bsb22CommitmentComputePlaceholder"

// lookupRows decodes a logderivlookup instruction: the blueprint holds the table entries (EntriesCalldata, compressed linear
// expressions), the instruction's calldata the queries.
func lookupRows(bp *constraint.BlueprintLookupHint, ins constraint.Instruction) (table, queries []constraint.LinearExpression) {
	for j := 0; j < len(bp.EntriesCalldata); {
		n := int(bp.EntriesCalldata[j])
		j++
		le := make(constraint.LinearExpression, n)
		for k := 0; k < n; k++ {
			le[k] = constraint.Term{CID: bp.EntriesCalldata[j], VID: bp.EntriesCalldata[j+1]}
			j += 2
		}
		table = append(table, le)
	}
	cd := ins.Calldata[3:] // [size, nbInputs, nbOutputs, inputs...]
	for j := 0; j < len(cd); {
		n := int(cd[j])
		j++
		le := make(constraint.LinearExpression, n)
		for k := 0; k < n; k++ {
			le[k] = constraint.Term{CID: cd[j], VID: cd[j+1]}
			j += 2
		}
		queries = append(queries, le)
	}
	return
}

// compactTablePtr turns (begin, end) pairs into the n_tables + 1 offsets of the contract; a table referenced by several lookup
// instructions of one blueprint is registered once by the caller in a full implementation (keyed by blueprint id).
func compactTablePtr(pairs []uint64) []uint64 {
	if len(pairs) == 0 {
		return []uint64{0}
	}
	out := []uint64{pairs[0]}
	for i := 1; i < len(pairs); i += 2 {
		out = append(out, pairs[i])
	}
	return out
}
