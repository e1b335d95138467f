// Package zkporb200 binds libzkpor_b200 (CUDA sm_100a kernels behind include/zkpor_b200.h) into the reference's Go services.
// UNCOMPILED in this repository: the image has no Go toolchain (see go/README.md).
package zkporb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../zkmerkle-proof-of-solvency_b200 -lzkpor_b200
#include <stdlib.h>
#include "zkpor_b200.h"
*/
import "C"

import (
	"errors"
	"runtime"
)

// Ctx is one GPU (one CUDA stream).  A Ctx is not re-entrant: one call in flight, like the reference's single prover loop
// (src/prover/prover/prover.go:141-247).  There is no CPU fallback: without a B200 NewCtx fails.
type Ctx struct{ h *C.zkpor_ctx }

func lastErr() error { return errors.New("zkpor_b200: " + C.GoString(C.zkpor_last_error())) }

// every library call reports its error through a thread-local slot, so the calling goroutine must not migrate between the call and
// zkpor_last_error: call() pins it for the duration.
func call(f func() C.int32_t) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != C.ZKPOR_OK {
		return lastErr()
	}
	return nil
}

func NewCtx(device int) (*Ctx, error) {
	c := &Ctx{}
	if err := call(func() C.int32_t { return C.zkpor_ctx_create(C.int32_t(device), &c.h) }); err != nil {
		return nil, err
	}
	return c, nil
}

// NewGroup returns len(devices) contexts joined in one in-process group: one proof is then split across them (ProveSharded).
func NewGroup(devices []int) ([]*Ctx, error) {
	ids := make([]C.int32_t, len(devices))
	for i, d := range devices {
		ids[i] = C.int32_t(d)
	}
	hs := make([]*C.zkpor_ctx, len(devices))
	if err := call(func() C.int32_t { return C.zkpor_ctx_create_multi(&ids[0], C.int32_t(len(ids)), &hs[0]) }); err != nil {
		return nil, err
	}
	out := make([]*Ctx, len(hs))
	for i := range hs {
		out[i] = &Ctx{hs[i]}
	}
	return out, nil
}

func (c *Ctx) Close() { C.zkpor_ctx_destroy(c.h); c.h = nil }
