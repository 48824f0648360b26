// Additive cgo binding a maintainer would drop next to AL/operator/mina/mina.go: route all Mina items of a
// batch through ONE call instead of one goroutine + one cgo call per proof
// (AL/operator/pkg/operator.go:448-465).  UNCOMPILED here (no Go toolchain in this image).
// The existing VerifyMinaState in mina.go needs no change at all: same header, same library name.
package mina

/*
#cgo linux LDFLAGS: ${SRCDIR}/lib/libmina_state_verifier_ffi.so -ldl -lrt -lm -Wl,--allow-multiple-definition
#include <stdlib.h>
#include "lib/mina_verifier.h"
*/
import "C"
import "unsafe"

// VerifyMinaStateBatch returns one accept bit per (proof, pubInput) pair.
func VerifyMinaStateBatch(proofs [][]byte, pubInputs [][]byte) []bool {
	n := len(proofs)
	out := make([]bool, n)
	if n == 0 {
		return out
	}
	ptrSize := C.size_t(unsafe.Sizeof(uintptr(0)))
	lenSize := C.size_t(unsafe.Sizeof(C.size_t(0)))
	pp := (*[1 << 28]*C.uchar)(C.malloc(C.size_t(n) * ptrSize))
	qp := (*[1 << 28]*C.uchar)(C.malloc(C.size_t(n) * ptrSize))
	pl := (*[1 << 28]C.size_t)(C.malloc(C.size_t(n) * lenSize))
	ql := (*[1 << 28]C.size_t)(C.malloc(C.size_t(n) * lenSize))
	defer C.free(unsafe.Pointer(pp))
	defer C.free(unsafe.Pointer(qp))
	defer C.free(unsafe.Pointer(pl))
	defer C.free(unsafe.Pointer(ql))
	var pins []unsafe.Pointer
	for i := 0; i < n; i++ {
		p, q := C.CBytes(proofs[i]), C.CBytes(pubInputs[i])
		pins = append(pins, p, q)
		pp[i], qp[i] = (*C.uchar)(p), (*C.uchar)(q)
		pl[i], ql[i] = C.size_t(len(proofs[i])), C.size_t(len(pubInputs[i]))
	}
	defer func() {
		for _, p := range pins {
			C.free(p)
		}
	}()
	accept := make([]C.uint8_t, n)
	rc := C.verify_mina_state_batch_ffi(C.size_t(n), &pp[0], &pl[0], &qp[0], &ql[0], &accept[0])
	for i := 0; i < n && rc == 0; i++ {
		out[i] = accept[i] == 1
	}
	return out
}
