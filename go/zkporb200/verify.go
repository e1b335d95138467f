package zkporb200

/*
#include "zkpor_b200.h"
*/
import "C"

import (
	"bytes"
	"crypto/rand"
	"errors"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
)

var ErrPairingCheckFailed = errors.New("pairing doesn't match")

type vkKeep struct{ pc []uint64 }

func describeVK(vk *groth16_bn254.VerifyingKey, keep *vkKeep) C.zkpor_vk_desc {
	var d C.zkpor_vk_desc
	d.g1_alpha = unsafe.Pointer(&vk.G1.Alpha)
	d.g2_beta, d.g2_gamma, d.g2_delta = unsafe.Pointer(&vk.G2.Beta), unsafe.Pointer(&vk.G2.Gamma), unsafe.Pointer(&vk.G2.Delta)
	d.g1_k, d.n_k = unsafe.Pointer(&vk.G1.K[0]), C.uint64_t(len(vk.G1.K))
	d.n_commitments = C.uint64_t(len(vk.PublicAndCommitmentCommitted))
	if d.n_commitments == 1 {
		for _, w := range vk.PublicAndCommitmentCommitted[0] { // empty for the reference circuits
			keep.pc = append(keep.pc, uint64(w))
		}
		if len(keep.pc) > 0 {
			d.public_committed = (*C.uint64_t)(unsafe.Pointer(&keep.pc[0]))
		}
		d.n_public_committed = C.uint64_t(len(keep.pc))
		d.g2_ped_g, d.g2_ped_g_root_sigma_neg = unsafe.Pointer(&vk.CommitmentKey.G), unsafe.Pointer(&vk.CommitmentKey.GRootSigmaNeg)
	}
	return d
}

// Verify has groth16.Verify's contract (src/prover/prover/prover.go:276, src/verifier/main.go:284): nil = valid.
func (c *Ctx) Verify(proof *groth16_bn254.Proof, vk *groth16_bn254.VerifyingKey, pub fr.Vector) error {
	var keep vkKeep
	d := describeVK(vk, &keep)
	var raw bytes.Buffer
	proof.WriteRawTo(&raw) // the bytes prover.go:201 stores
	var ok C.int32_t
	if err := call(func() C.int32_t {
		return C.zkpor_groth16_verify(c.h, &d, (*C.uint8_t)(&raw.Bytes()[0]), C.uint32_t(raw.Len()), unsafe.Pointer(&pub[0]), C.uint64_t(len(pub)), &ok)
	}); err != nil {
		return err // malformed proof
	}
	if ok == 0 {
		return ErrPairingCheckFailed
	}
	return nil
}

// VerifyBatch checks every batch proof of a snapshot (the verifier's loop, src/verifier/main.go:176-302) with ONE pairing product;
// when it fails the caller falls back to Verify per proof to name the offender.
func (c *Ctx) VerifyBatch(proofs []*groth16_bn254.Proof, vk *groth16_bn254.VerifyingKey, pubs []fr.Vector) (bool, error) {
	var keep vkKeep
	d := describeVK(vk, &keep)
	var raw bytes.Buffer
	for _, p := range proofs {
		p.WriteRawTo(&raw)
	}
	stride := raw.Len() / len(proofs)
	flat := make(fr.Vector, 0, len(pubs)*len(pubs[0]))
	for _, p := range pubs {
		flat = append(flat, p...)
	}
	var seed [32]byte
	rand.Read(seed[:])
	var ok C.int32_t
	err := call(func() C.int32_t {
		return C.zkpor_groth16_verify_batch(c.h, &d, (*C.uint8_t)(&raw.Bytes()[0]), C.uint32_t(stride), C.uint64_t(stride), unsafe.Pointer(&flat[0]),
			C.uint64_t(len(pubs[0])), C.uint64_t(len(proofs)), (*C.uint8_t)(&seed[0]), &ok)
	})
	return ok == 1, err
}
