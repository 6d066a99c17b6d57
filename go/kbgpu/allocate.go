package kbgpu

import (
	"sort"

	"github.com/golang/glog"

	"github.com/kubernetes-sigs/kube-batch/pkg/scheduler/framework"
)

// allocateAction is a drop-in for pkg/scheduler/actions/allocate: same Name(), same framework.Action
// interface (framework/interface.go:20-32), registered with framework.RegisterAction in actions/factory.go.
//
// Fallback is the ORIGINAL action (actions/allocate.New()), still linked in.  libkbgpu itself has no CPU path and never
// will; but one tenant pod with inter-pod affinity, a custom plugin or a CUDA error must not stop scheduling for the
// whole cluster, so the shim hands exactly those cycles to the reference implementation and says so in the log.
type allocateAction struct {
	engine   *Engine
	Fallback framework.Action
}

func New(e *Engine, fallback framework.Action) *allocateAction { return &allocateAction{engine: e, Fallback: fallback} }
func (alloc *allocateAction) Name() string                      { return "allocate" }
func (alloc *allocateAction) Initialize()                       {}
func (alloc *allocateAction) UnInitialize()                     {}

// Execute replaces the queue->job->task loop of allocate.go:43-194: flatten, one kb_allocate, replay.
func (alloc *allocateAction) Execute(ssn *framework.Session) {
	if err := run(alloc.engine, ssn, false); err != nil {
		fallBack(alloc.Fallback, ssn, err)
	}
}

func fallBack(fb framework.Action, ssn *framework.Session, err error) {
	if fb == nil {
		glog.Errorf("kbgpu: %v; no fallback action configured: this cycle schedules nothing", err)
		return
	}
	glog.Warningf("kbgpu: %v; running the Go %s action for this cycle", err, fb.Name())
	fb.Execute(ssn)
}

// run is one action on the GPU; every action flattens the session as it is NOW, like the reference runs its
// actions one after the other on the same *framework.Session (scheduler.go:88-92).  Nothing has been changed in
// the session when an error is returned, so the caller can still run the reference action.
func run(engine *Engine, ssn *framework.Session, backfill bool) error {
	flat, err := Flatten(ssn) // canonical orders + label/taint/port atom interning; see flatten.go
	if err != nil {
		return err
	}
	defer flat.Free()
	if err := engine.Load(flat); err != nil {
		return err
	}
	var dec []Decision
	if backfill {
		dec, err = engine.Backfill(len(flat.Tasks))
	} else {
		dec, err = engine.Allocate(len(flat.Tasks))
	}
	if err != nil {
		return err
	}
	// Replay in the order the reference would have made the calls, through the UNCHANGED session methods, so
	// event handlers (drf / proportion), gang dispatch, cache.Bind, metrics and status updates behave as today
	// (framework/session.go:194-314).  ssn.Allocate itself dispatches when ssn.JobReady — the Dispatched bit of
	// the decision is only used to cross-check.
	order := make([]int, 0, len(dec))
	for i, d := range dec {
		if d.Kind == 1 || d.Kind == 2 {
			order = append(order, i)
		}
	}
	sort.Slice(order, func(a, b int) bool { return dec[order[a]].Step < dec[order[b]].Step })
	for _, i := range order {
		task := flat.Tasks[i]
		node := flat.NodeNames[dec[i].Node]
		var err error
		if dec[i].Kind == 1 {
			err = ssn.Allocate(task, node)
		} else {
			err = ssn.Pipeline(task, node)
		}
		if err != nil {
			glog.Errorf("kbgpu: replay of task %v on %v failed: %v", task.UID, node, err)
		}
	}
	return nil
}
