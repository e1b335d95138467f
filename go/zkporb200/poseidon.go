package zkporb200

/*
#include "zkpor_b200.h"
*/
import "C"

import (
	"hash"
	"unsafe"
)

// Hasher is poseidon.NewPoseidon() of the bnb-chain gnark-crypto fork as the reference uses it: a hash.Hash whose Write appends
// 32-byte big-endian field elements and whose Sum(b) APPENDS the 32-byte digest to b (src/utils/account_tree.go:19,27;
// merkletree.go:251-259 relies on Sum(buf[off:off]) writing in place).  One GPU round trip per Sum: a drop-in for correctness;
// the batch calls (HashBatch, AccountLeaves, the Merkle tree below) are what the witness service should use for throughput.
type Hasher struct {
	ctx *Ctx
	buf []byte
}

var _ hash.Hash = (*Hasher)(nil)

func (c *Ctx) NewPoseidon() hash.Hash { return &Hasher{ctx: c} }

func (h *Hasher) Write(p []byte) (int, error) { h.buf = append(h.buf, p...); return len(p), nil }
func (h *Hasher) Reset()                      { h.buf = h.buf[:0] }
func (h *Hasher) Size() int                   { return 32 }
func (h *Hasher) BlockSize() int              { return 32 }

func (h *Hasher) Sum(b []byte) []byte {
	in := h.buf
	if len(in) == 0 {
		in = make([]byte, 32) // the fork hashes []byte{0} for empty input (witness/main.go:181 relies on it)
	}
	if r := len(in) % 32; r != 0 { // a short trailing chunk is one element, left-padded
		pad := make([]byte, 32-r)
		in = append(append(in[:len(in)-r:len(in)-r], pad...), in[len(in)-r:]...)
	}
	var out [32]byte
	if err := call(func() C.int32_t {
		return C.zkpor_poseidon_hash_batch(h.ctx.h, unsafe.Pointer(&in[0]), C.uint32_t(len(in)/32), 1, unsafe.Pointer(&out[0]))
	}); err != nil {
		panic(err) // hash.Hash cannot return an error; no GPU is a deployment fault, never a silent CPU fallback
	}
	return append(b, out[:]...)
}

// HashBatch: count independent hashes of nIn elements each (poseidon.PoseidonBytes over many inputs at once).
func (c *Ctx) HashBatch(in []byte, nIn, count int) ([]byte, error) {
	out := make([]byte, 32*count)
	err := call(func() C.int32_t {
		return C.zkpor_poseidon_hash_batch(c.h, unsafe.Pointer(&in[0]), C.uint32_t(nIn), C.uint64_t(count), unsafe.Pointer(&out[0]))
	})
	return out, err
}

// AccountLeaves is utils.AccountInfoToHash (src/utils/utils.go:744-750) for all accounts of one asset tier; flat comes from
// utils.PaddingAccountAssets (host logic, src/utils/utils.go:147-186) laid out as tier*6 uint64 per account.
func (c *Ctx) AccountLeaves(idsBE, totalsBE []byte, flat []uint64, n, tier int) ([]byte, error) {
	out := make([]byte, 32*n)
	err := call(func() C.int32_t {
		return C.zkpor_account_leaves(c.h, unsafe.Pointer(&idsBE[0]), unsafe.Pointer(&totalsBE[0]), unsafe.Pointer(&flat[0]), C.uint64_t(n),
			C.uint32_t(tier), unsafe.Pointer(&out[0]))
	})
	return out, err
}
