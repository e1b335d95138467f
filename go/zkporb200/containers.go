// Byte containers: proofs, verifying keys and proving keys in gnark's own file formats, read straight into HBM.
//
// Replaces (reference file:line):
//   pk.UnsafeReadFrom(f)                    src/prover/prover/prover.go:342-346   -> ReadProvingKey (zkpor_pk_read)
//   vk.ReadFrom(f)                          src/prover/prover/prover.go:358-362, src/verifier/main.go:33-34 -> ReadVerifyingKey
//   proof.ReadFrom(bytes.NewBuffer(b))      src/verifier/main.go:208-216          -> DecodeProof
//   pk.WriteTo / vk.WriteTo                 src/keygen/main.go:46-62              -> (*DeviceKey).WriteTo, EncodeVerifyingKey
// The r1cs file stays with gnark (cs.ReadFrom, prover.go:317-327): program.go flattens the parsed system.
//
// Source only: this image has no Go toolchain (see go/README.md).
package zkporb200

/*
#include "zkpor_b200.h"
*/
import "C"

import (
	"os"
	"syscall"
	"unsafe"

	"github.com/consensys/gnark/constraint"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
)

// ReadProvingKey maps the key file and lets the library decode every point array on the GPU (one square root per compressed point)
// directly into the resident key: the 12 GB file never becomes 26 GB of Go heap.
func (c *Ctx) ReadProvingKey(path string, r1cs *cs_bn254.R1CS) (*DeviceKey, error) {
	f, err := os.Open(path)
	if err != nil {
		return nil, err
	}
	defer f.Close()
	st, err := f.Stat()
	if err != nil {
		return nil, err
	}
	data, err := syscall.Mmap(int(f.Fd()), 0, int(st.Size()), syscall.PROT_READ, syscall.MAP_SHARED)
	if err != nil {
		return nil, err
	}
	defer syscall.Munmap(data)

	var info C.zkpor_pk_cs_info
	info.n_public = C.uint64_t(r1cs.GetNbPublicVariables())
	commitments := r1cs.CommitmentInfo.(constraint.Groth16Commitments)
	var committed []uint64
	if len(commitments) == 1 {
		for _, w := range commitments[0].PrivateCommitted {
			committed = append(committed, uint64(w))
		}
		info.commitment_index = C.uint64_t(commitments[0].CommitmentIndex)
		info.n_committed = C.uint64_t(len(committed))
		if len(committed) > 0 {
			info.private_committed = (*C.uint64_t)(unsafe.Pointer(&committed[0]))
		}
	}
	var h *C.zkpor_pk
	var used C.uint64_t
	if rc := C.zkpor_pk_read(c.h, (*C.uint8_t)(unsafe.Pointer(&data[0])), C.uint64_t(len(data)), &info, &h, &used); rc != 0 {
		return nil, lastErr()
	}
	return &DeviceKey{h: h}, nil
}

// WriteTo is pk.WriteTo (raw = false) / pk.WriteRawTo from the resident key.
func (k *DeviceKey) WriteTo(c *Ctx, raw bool) ([]byte, error) {
	var n C.uint64_t
	r := C.int32_t(0)
	if raw {
		r = 1
	}
	if rc := C.zkpor_pk_write(c.h, k.h, r, nil, 0, &n); rc != 0 {
		return nil, lastErr()
	}
	out := make([]byte, int(n))
	if rc := C.zkpor_pk_write(c.h, k.h, r, (*C.uint8_t)(unsafe.Pointer(&out[0])), n, &n); rc != 0 {
		return nil, lastErr()
	}
	return out, nil
}

// DecodeProof is proof.ReadFrom: compressed or raw bytes in, the raw layout Verify / VerifyBatch take out.
func (c *Ctx) DecodeProof(b []byte) ([]byte, error) {
	out := make([]byte, 260+64*17)
	n := C.uint32_t(len(out))
	if rc := C.zkpor_proof_decode(c.h, (*C.uint8_t)(unsafe.Pointer(&b[0])), C.uint64_t(len(b)), (*C.uint8_t)(unsafe.Pointer(&out[0])), &n, nil); rc != 0 {
		return nil, lastErr()
	}
	return out[:n], nil
}

// VerifyingKeyFile is vk.ReadFrom: the points come back as gnark memory images (G1Affine / G2Affine), ready for zkpor_vk_desc.
type VerifyingKeyFile struct {
	Host            C.zkpor_vk_host
	K               []byte   // n_k x 64
	PublicCommitted []uint64 // vk.PublicAndCommitmentCommitted[0]
}

func (c *Ctx) ReadVerifyingKey(b []byte) (*VerifyingKeyFile, error) {
	vk := &VerifyingKeyFile{K: make([]byte, 64*64), PublicCommitted: make([]uint64, 64)}
	if rc := C.zkpor_vk_decode(c.h, (*C.uint8_t)(unsafe.Pointer(&b[0])), C.uint64_t(len(b)), &vk.Host, unsafe.Pointer(&vk.K[0]), 64,
		(*C.uint64_t)(unsafe.Pointer(&vk.PublicCommitted[0])), 64, nil); rc != 0 {
		return nil, lastErr()
	}
	vk.K = vk.K[:64*int(vk.Host.n_k)]
	vk.PublicCommitted = vk.PublicCommitted[:int(vk.Host.n_public_committed)]
	return vk, nil
}
