package zkporb200

/*
#include "zkpor_b200.h"
*/
import "C"

import (
	"unsafe"

	curve "github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
)

// The building blocks of groth16.Setup (src/keygen/main.go:42) on the device.  The toxic waste is drawn on the Go side
// (crypto/rand) exactly as gnark does and never leaves the process; the composition -- Lagrange basis at tau, per-wire A/B/C
// sums over the R1CS columns, the K and Z scalars, then fixed-base multiplications -- is the one zkpor_b200.groth16_setup (Python)
// runs in the parity test tests/test_gpu_setup.py.

// BatchScalarMultiplicationG1 replaces curve.BatchScalarMultiplicationG1(&base, scalars).
func (c *Ctx) BatchScalarMultiplicationG1(base *curve.G1Affine, scalars []fr.Element) ([]curve.G1Affine, error) {
	out := make([]curve.G1Affine, len(scalars))
	err := call(func() C.int32_t {
		return C.zkpor_g1_fixed_base_batch(c.h, unsafe.Pointer(base), unsafe.Pointer(&scalars[0]), C.uint64_t(len(scalars)), C.ZKPOR_SCALARS_MONT,
			unsafe.Pointer(&out[0]))
	})
	return out, err
}

// BatchScalarMultiplicationG2 replaces curve.BatchScalarMultiplicationG2(&base, scalars).
func (c *Ctx) BatchScalarMultiplicationG2(base *curve.G2Affine, scalars []fr.Element) ([]curve.G2Affine, error) {
	out := make([]curve.G2Affine, len(scalars))
	err := call(func() C.int32_t {
		return C.zkpor_g2_fixed_base_batch(c.h, unsafe.Pointer(base), unsafe.Pointer(&scalars[0]), C.uint64_t(len(scalars)), C.ZKPOR_SCALARS_MONT,
			unsafe.Pointer(&out[0]))
	})
	return out, err
}
