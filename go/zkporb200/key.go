package zkporb200

/*
#include "zkpor_b200.h"
*/
import "C"

import (
	"math/bits"
	"unsafe"

	"github.com/consensys/gnark/backend/groth16"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	"github.com/consensys/gnark/constraint"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
)

// DeviceKey is the proving key resident in HBM: filled once per asset tier right after pk.UnsafeReadFrom
// (src/prover/prover/prover.go:342-346); gnark's Go copy can be dropped afterwards (12 GB of host memory per tier).
type DeviceKey struct {
	h   *C.zkpor_pk
	ctx *Ctx
}

func describe(pk *groth16_bn254.ProvingKey, r1cs *cs_bn254.R1CS, keep *[]uint64) C.zkpor_pk_desc {
	var d C.zkpor_pk_desc
	d.log_n = C.uint32_t(bits.TrailingZeros64(pk.Domain.Cardinality))
	d.n_wires, d.n_public = C.uint64_t(len(pk.InfinityA)), C.uint64_t(r1cs.GetNbPublicVariables())
	d.n_a, d.n_b, d.n_k, d.n_z = C.uint64_t(len(pk.G1.A)), C.uint64_t(len(pk.G1.B)), C.uint64_t(len(pk.G1.K)), C.uint64_t(len(pk.G1.Z))
	// bn254.G1Affine in memory = X || Y, 4 little-endian Montgomery words each: exactly the layout the library reads
	d.g1_a, d.g1_b = unsafe.Pointer(&pk.G1.A[0]), unsafe.Pointer(&pk.G1.B[0])
	d.g1_k, d.g1_z, d.g2_b = unsafe.Pointer(&pk.G1.K[0]), unsafe.Pointer(&pk.G1.Z[0]), unsafe.Pointer(&pk.G2.B[0])
	d.g1_alpha, d.g1_beta, d.g1_delta = unsafe.Pointer(&pk.G1.Alpha), unsafe.Pointer(&pk.G1.Beta), unsafe.Pointer(&pk.G1.Delta)
	d.g2_beta, d.g2_delta = unsafe.Pointer(&pk.G2.Beta), unsafe.Pointer(&pk.G2.Delta)
	d.infinity_a = (*C.uint8_t)(unsafe.Pointer(&pk.InfinityA[0])) // []bool is one byte per element
	d.infinity_b = (*C.uint8_t)(unsafe.Pointer(&pk.InfinityB[0]))
	if info := r1cs.CommitmentInfo.(constraint.Groth16Commitments); len(info) == 1 {
		ck := pk.CommitmentKeys[0]
		d.n_committed = C.uint64_t(len(ck.Basis))
		d.ck_basis, d.ck_basis_exp_sigma = unsafe.Pointer(&ck.Basis[0]), unsafe.Pointer(&ck.BasisExpSigma[0])
		*keep = make([]uint64, len(info[0].PrivateCommitted))
		for i, w := range info[0].PrivateCommitted {
			(*keep)[i] = uint64(w)
		}
		d.private_committed = (*C.uint64_t)(unsafe.Pointer(&(*keep)[0]))
		d.commitment_index = C.uint64_t(info[0].CommitmentIndex)
	}
	return d
}

// UploadKey copies the whole key to c's GPU; on a context of a group (NewGroup) it copies this rank's chunks only.
func (c *Ctx) UploadKey(pk groth16.ProvingKey, r1cs constraint.ConstraintSystem, shard bool) (*DeviceKey, error) {
	var keep []uint64
	d := describe(pk.(*groth16_bn254.ProvingKey), r1cs.(*cs_bn254.R1CS), &keep)
	k := &DeviceKey{ctx: c}
	err := call(func() C.int32_t {
		if shard {
			return C.zkpor_pk_upload_shard(c.h, &d, &k.h)
		}
		return C.zkpor_pk_upload(c.h, &d, &k.h)
	})
	if err != nil {
		return nil, err
	}
	return k, nil
}

func (k *DeviceKey) Close() { C.zkpor_pk_free(k.ctx.h, k.h); k.h = nil }
