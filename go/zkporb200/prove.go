package zkporb200

/*
#include "zkpor_b200.h"
*/
import "C"

import (
	"bytes"
	"runtime"
	"sync"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	"github.com/consensys/gnark/backend/witness"
)

// Prove has the contract of groth16.Prove(r1cs, pk, fullWitness) -- src/prover/prover/prover.go:269 -- with the key and the
// compiled system already resident on the GPU: the witness solver (hints included, the BSB22 commitment mid-solve), computeH and
// the five multi-scalar multiplications all run on the device; 66 MB of inputs go in, 388 proof bytes come out.
func (c *Ctx) Prove(prog *Program, key *DeviceKey, full witness.Witness) (*groth16_bn254.Proof, error) {
	w := full.Vector().(fr.Vector) // public (without ONE) then secret, Montgomery fr.Elements: the library's `inputs`
	var r, s fr.Element
	if _, err := r.SetRandom(); err != nil { // gnark draws the blinding scalars the same way
		return nil, err
	}
	if _, err := s.SetRandom(); err != nil {
		return nil, err
	}
	rb, sb := r.Bytes(), s.Bytes()
	var raw [512]byte
	var n C.uint32_t
	err := call(func() C.int32_t {
		return C.zkpor_groth16_prove_solve(c.h, key.h, prog.h, unsafe.Pointer(&w[0]), (*C.uint8_t)(&rb[0]), (*C.uint8_t)(&sb[0]),
			(*C.uint8_t)(&raw[0]), &n)
	})
	if err != nil {
		return nil, err // an unsatisfied constraint surfaces here, as with gnark (the reference propagates it: prover.go:195-199)
	}
	proof := new(groth16_bn254.Proof)
	_, err = proof.ReadFrom(bytes.NewReader(raw[:n])) // WriteRawTo layout; gnark's decoder accepts raw points
	return proof, err
}

// ProveSharded splits ONE proof across the contexts of a group (NewGroup): progs[i] / keys[i] live on ctxs[i] (keys uploaded with
// shard = true).  The call is collective -- one OS-locked goroutine per GPU -- and every rank returns the same bytes.
func ProveSharded(ctxs []*Ctx, progs []*Program, keys []*DeviceKey, full witness.Witness) (*groth16_bn254.Proof, error) {
	w := full.Vector().(fr.Vector)
	var r, s fr.Element
	r.SetRandom()
	s.SetRandom()
	rb, sb := r.Bytes(), s.Bytes()
	raws := make([][512]byte, len(ctxs))
	lens := make([]C.uint32_t, len(ctxs))
	errs := make([]error, len(ctxs))
	var wg sync.WaitGroup
	for i := range ctxs {
		wg.Add(1)
		go func(i int) {
			defer wg.Done()
			runtime.LockOSThread()
			defer runtime.UnlockOSThread()
			if rc := C.zkpor_groth16_prove_solve(ctxs[i].h, keys[i].h, progs[i].h, unsafe.Pointer(&w[0]), (*C.uint8_t)(&rb[0]), (*C.uint8_t)(&sb[0]),
				(*C.uint8_t)(&raws[i][0]), &lens[i]); rc != C.ZKPOR_OK {
				errs[i] = lastErr() // a failing rank releases its peers inside the library; nobody is left waiting
			}
		}(i)
	}
	wg.Wait()
	for _, e := range errs {
		if e != nil {
			return nil, e
		}
	}
	proof := new(groth16_bn254.Proof)
	_, err := proof.ReadFrom(bytes.NewReader(raws[0][:lens[0]]))
	return proof, err
}
