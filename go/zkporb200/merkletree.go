package zkporb200

/*
#include "zkpor_b200.h"
*/
import "C"

import (
	"errors"
	"hash"
	"unsafe"
)

// FixedDepthMerkleTree is a source-level drop-in for src/utils/merkletree (merkletree.go:27-355): same constructor arguments and
// method set, the nodes live in HBM.  Set buffers leaves on the host and marks them dirty; Build uploads the dirty leaves and
// rehashes only the touched paths, level by level, on the GPU (the reference's dirty-bitset walk, merkletree.go:192-280).
type FixedDepthMerkleTree struct {
	ctx     *Ctx
	h       *C.zkpor_tree
	depth   int
	keys    []uint32
	pending []byte
}

// hasherFunc is accepted for signature compatibility and ignored: the tree hashes with the library's Poseidon (the same function
// poseidon.NewPoseidon computes); a different hash is a programming error, not a fallback.
func (c *Ctx) NewFixedDepthMerkleTree(depth int, nilLeafHash []byte, _ func() hash.Hash, capacity int) (*FixedDepthMerkleTree, error) {
	t := &FixedDepthMerkleTree{ctx: c, depth: depth}
	err := call(func() C.int32_t {
		return C.zkpor_tree_create(c.h, C.uint32_t(depth), (*C.uint8_t)(&nilLeafHash[0]), C.uint64_t(capacity), &t.h)
	})
	return t, err
}

func (t *FixedDepthMerkleTree) Set(key uint32, value []byte) error {
	if len(value) != 32 {
		return errors.New("merkletree: value must be 32 bytes")
	}
	t.keys = append(t.keys, key)
	t.pending = append(t.pending, value...)
	return nil
}

// SetRange is the bulk form the witness service uses after AccountLeaves (src/witness/main.go:183-195).
func (t *FixedDepthMerkleTree) SetRange(first uint32, leaves []byte) error {
	return call(func() C.int32_t {
		return C.zkpor_tree_set_range(t.ctx.h, t.h, C.uint64_t(first), C.uint64_t(len(leaves)/32), unsafe.Pointer(&leaves[0]))
	})
}

func (t *FixedDepthMerkleTree) Build() {
	if len(t.keys) > 0 {
		if err := call(func() C.int32_t {
			return C.zkpor_tree_set_keys(t.ctx.h, t.h, (*C.uint32_t)(&t.keys[0]), C.uint64_t(len(t.keys)), unsafe.Pointer(&t.pending[0]))
		}); err != nil {
			panic(err)
		}
		t.keys, t.pending = t.keys[:0], t.pending[:0]
	}
	if err := call(func() C.int32_t { return C.zkpor_tree_build(t.ctx.h, t.h) }); err != nil {
		panic(err) // Build() has no error return in the reference
	}
}

func (t *FixedDepthMerkleTree) Root() []byte {
	out := make([]byte, 32)
	call(func() C.int32_t { return C.zkpor_tree_root(t.ctx.h, t.h, (*C.uint8_t)(&out[0])) })
	return out
}

func (t *FixedDepthMerkleTree) Get(key uint32) []byte {
	out := make([]byte, 32)
	call(func() C.int32_t { return C.zkpor_tree_get_leaves(t.ctx.h, t.h, (*C.uint32_t)(&key), 1, unsafe.Pointer(&out[0])) })
	return out
}

func (t *FixedDepthMerkleTree) GetProof(key uint32) ([][]byte, error) {
	ps, err := t.GetProofs([]uint32{key})
	if err != nil {
		return nil, err
	}
	return ps[0], nil
}

// GetProofs gathers the sibling paths of many keys in one kernel (witness.go:323 and userproof.go:138 call GetProof in loops).
func (t *FixedDepthMerkleTree) GetProofs(keys []uint32) ([][][]byte, error) {
	flat := make([]byte, len(keys)*t.depth*32)
	if err := call(func() C.int32_t {
		return C.zkpor_tree_get_proofs(t.ctx.h, t.h, (*C.uint32_t)(&keys[0]), C.uint64_t(len(keys)), unsafe.Pointer(&flat[0]))
	}); err != nil {
		return nil, err
	}
	out := make([][][]byte, len(keys))
	for i := range keys {
		out[i] = make([][]byte, t.depth)
		for l := 0; l < t.depth; l++ {
			o := (i*t.depth + l) * 32
			out[i][l] = flat[o : o+32 : o+32]
		}
	}
	return out, nil
}

// VerifyProof mirrors merkletree.VerifyProof (merkletree.go:334-355) with the library's hasher.
func (c *Ctx) VerifyProof(root []byte, key uint32, proof [][]byte, leaf []byte, depth int) bool {
	if len(proof) != depth {
		return false
	}
	h := c.NewPoseidon()
	node := leaf
	for i := 0; i < depth; i++ {
		h.Reset()
		if (key>>uint(i))&1 == 0 {
			h.Write(node)
			h.Write(proof[i])
		} else {
			h.Write(proof[i])
			h.Write(node)
		}
		node = h.Sum(nil)
	}
	return string(node) == string(root)
}

func (t *FixedDepthMerkleTree) Close() { C.zkpor_tree_free(t.ctx.h, t.h); t.h = nil }

// ---- one tree over the GPUs of a group, and the witness service's batch loop ------------------------------------------------------------

// ShardRange is the key range this context's rank owns when the tree is built across a group (zkpor_ctx_create_multi / comm_init).
func (t *FixedDepthMerkleTree) ShardRange() (first, count uint64, subtreeLevel uint32) {
	var f, n C.uint64_t
	var l C.uint32_t
	C.zkpor_tree_shard_range(t.ctx.h, t.h, &f, &n, &l)
	return uint64(f), uint64(n), uint32(l)
}

// BuildSharded is Build() as a collective: every rank has set the leaves of its own range; subtrees per rank, all-gather of the
// subtree roots, top levels everywhere (SURVEY.md 8(e)).  Call it from one OS-locked goroutine per rank.
func (t *FixedDepthMerkleTree) BuildSharded() error {
	if rc := C.zkpor_tree_build_sharded(t.ctx.h, t.h); rc != 0 {
		return lastErr()
	}
	return nil
}

// CexState is what the witness loop carries from batch to batch (src/witness/witness/witness.go:144-206).
type CexState struct {
	BasePrices     []uint64 // utils.AssetCounts
	TierRatioElems []byte   // AssetCounts x 18 x 32: ConvertTierRatiosToBytes of the three ratio tables, big-endian, left-padded
	Totals         []uint64 // AssetCounts x 5
}

// WitnessBatches replaces the serial main loop for all batches of one tier: flat = the accounts' PaddingAccountAssets rows in batch
// order, indices = their account indices.  Returns the CEX totals before every batch (+ the final ones), the n+1 CEX commitments
// (Before of batch b = row b, After = row b+1) and the n batch commitments; GetProofs serves the account proofs.
func (c *Ctx) WitnessBatches(st *CexState, root []byte, flat []uint64, indices []uint32, tier, opsPerBatch int) (totals []uint64, cexCm, batchCm []byte, err error) {
	n := len(indices)
	nb := n / opsPerBatch
	var d C.zkpor_cex_desc
	d.n_assets = C.uint32_t(len(st.BasePrices))
	d.base_prices = (*C.uint64_t)(unsafe.Pointer(&st.BasePrices[0]))
	d.tier_ratio_elems = unsafe.Pointer(&st.TierRatioElems[0])
	d.initial_totals = (*C.uint64_t)(unsafe.Pointer(&st.Totals[0]))
	totals = make([]uint64, (nb+1)*len(st.BasePrices)*5)
	cexCm, batchCm = make([]byte, (nb+1)*32), make([]byte, nb*32)
	if rc := C.zkpor_witness_batches(c.h, &d, (*C.uint8_t)(unsafe.Pointer(&root[0])), unsafe.Pointer(&flat[0]), (*C.uint32_t)(unsafe.Pointer(&indices[0])),
		C.uint64_t(n), C.uint32_t(tier), C.uint32_t(opsPerBatch), (*C.uint64_t)(unsafe.Pointer(&totals[0])), unsafe.Pointer(&cexCm[0]), unsafe.Pointer(&batchCm[0])); rc != 0 {
		return nil, nil, nil, lastErr()
	}
	copy(st.Totals, totals[nb*len(st.BasePrices)*5:]) // the next tier starts from here
	return totals, cexCm, batchCm, nil
}
