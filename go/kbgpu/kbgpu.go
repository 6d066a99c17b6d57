// Package kbgpu binds libkbgpu.so (include/kbgpu.h) into kube-batch through cgo.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain.  The file is the binding a
// kube-batch maintainer adds under pkg/scheduler/kbgpu; everything below the C ABI is exercised through
// the same ABI from Python/ctypes (tests/) and C (tests/test_abi.py).
package kbgpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../kube_batch_b200 -lkbgpu
#include <stdlib.h>
#include "kbgpu.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// Engine wraps one kb_engine (one CUDA device, one session in flight — like the single runOnce goroutine).
type Engine struct{ h *C.struct_kb_engine }

func NewEngine(device int) (*Engine, error) {
	opts := C.kb_engine_opts{abi_version: C.KB_ABI_VERSION, device: C.int32_t(device), rank: 0, world_size: 1}
	var h *C.struct_kb_engine
	if rc := C.kb_engine_create(&opts, &h); rc != 0 {
		return nil, fmt.Errorf("kb_engine_create: %s (%s)", C.GoString(C.kb_last_error(nil)), C.GoString(C.kb_status_str(rc)))
	}
	return &Engine{h: h}, nil
}

func (e *Engine) Close() { C.kb_engine_destroy(e.h) }

func (e *Engine) err(what string, rc C.int) error {
	return fmt.Errorf("%s: %s (%s)", what, C.GoString(C.kb_last_error(e.h)), C.GoString(C.kb_status_str(rc)))
}

// Load hands the flattened snapshot + tiers to the device.  kb_snapshot / kb_plugin_conf and every array they point to
// live in C memory for the call (cgo pointer rules: C may not be handed Go memory that contains pointers);
// kb_session_load copies everything it needs before it returns, so the arena is freed right after.
func (e *Engine) Load(f *Flat) error {
	if int(f.T) != len(f.Tasks) || int(f.N) != len(f.NodeNames) {
		return fmt.Errorf("kbgpu: inconsistent Flat (T=%d, %d tasks; N=%d, %d nodes)", f.T, len(f.Tasks), f.N, len(f.NodeNames))
	}
	var a arena
	defer a.free()
	snap := (*C.kb_snapshot)(C.calloc(1, C.size_t(unsafe.Sizeof(C.kb_snapshot{}))))
	defer C.free(unsafe.Pointer(snap))
	f.cSnapshot(snap, &a)
	conf := (*C.kb_plugin_conf)(C.calloc(1, C.size_t(unsafe.Sizeof(C.kb_plugin_conf{}))))
	defer C.free(unsafe.Pointer(conf))
	conf.n_tiers = C.uint32_t(f.nTiers)
	conf.tiers = f.cTiers
	if rc := C.kb_session_load(e.h, snap, conf); rc != 0 {
		if rc == C.KB_E_UNSUPPORTED_FEATURE || rc == C.KB_E_UNSUPPORTED_PLUGIN {
			return &ErrUnsupported{C.GoString(C.kb_last_error(e.h))}
		}
		return e.err("kb_session_load", rc)
	}
	return nil
}

// Decision mirrors kb_decision.
type Decision struct {
	Node         int32
	Kind         uint8 // 0 none, 1 allocated, 2 pipelined, 3 skipped (BestEffort)
	Dispatched   bool
	Step         uint32
	DispatchStep uint32
}

// Allocate runs the whole allocate cycle on the GPU and returns one decision per flattened task.
func (e *Engine) Allocate(nTasks int) ([]Decision, error) { return e.action(nTasks, false) }

// Backfill runs backfillAction.Execute (backfill.go:40-71) on the loaded session's current device state.
func (e *Engine) Backfill(nTasks int) ([]Decision, error) { return e.action(nTasks, true) }

func (e *Engine) action(nTasks int, backfill bool) ([]Decision, error) {
	raw := make([]C.kb_decision, nTasks+1)
	var st C.kb_stats
	if backfill {
		if rc := C.kb_backfill(e.h, (*C.kb_decision)(unsafe.Pointer(&raw[0])), &st); rc != 0 {
			return nil, e.err("kb_backfill", rc)
		}
	} else if rc := C.kb_allocate(e.h, (*C.kb_decision)(unsafe.Pointer(&raw[0])), &st); rc != 0 {
		return nil, e.err("kb_allocate", rc)
	}
	out := make([]Decision, nTasks)
	for i := range out {
		out[i] = Decision{Node: int32(raw[i].node), Kind: uint8(raw[i].kind), Dispatched: raw[i].dispatched != 0,
			Step: uint32(raw[i].step), DispatchStep: uint32(raw[i].dispatch_step)}
	}
	return out, nil
}

// Bind is one entry of the bind fan-out list.
type Bind struct {
	Task uint32 // index into Flat.Tasks
	Node int32  // index into Flat.NodeNames
}

// BindList returns the (task, node) pairs that reach cache.Bind in the cycle just run, in ssn.dispatch order
// (framework/session.go:277-314), compacted and sorted on the device.  A batched Binder (one API round trip for the whole list
// instead of cache.Bind's goroutine per task, cache/cache.go:491-535) consumes it; the per-task replay through ssn.Allocate
// keeps the session state, the list replaces only the fan-out.
func (e *Engine) BindList(nTasks int) ([]Bind, error) {
	tasks := make([]C.uint32_t, nTasks+1)
	nodes := make([]C.int32_t, nTasks+1)
	var n C.uint32_t
	if rc := C.kb_bind_list(e.h, &tasks[0], &nodes[0], &n); rc != 0 {
		return nil, e.err("kb_bind_list", rc)
	}
	out := make([]Bind, int(n))
	for i := range out {
		out[i] = Bind{Task: uint32(tasks[i]), Node: int32(nodes[i])}
	}
	return out, nil
}
