package kbgpu

import (
	"sort"

	"github.com/golang/glog"

	"github.com/kubernetes-sigs/kube-batch/pkg/scheduler/api"
	"github.com/kubernetes-sigs/kube-batch/pkg/scheduler/framework"
)

// allocateAction is a drop-in for pkg/scheduler/actions/allocate: same Name(), same framework.Action
// interface (framework/interface.go:20-32), registered with framework.RegisterAction in actions/factory.go.
type allocateAction struct{ engine *Engine }

func New(e *Engine) *allocateAction              { return &allocateAction{engine: e} }
func (alloc *allocateAction) Name() string       { return "allocate" }
func (alloc *allocateAction) Initialize()        {}
func (alloc *allocateAction) UnInitialize()      {}

// Execute replaces the queue->job->task loop of allocate.go:43-194: flatten, one kb_allocate, replay.
func (alloc *allocateAction) Execute(ssn *framework.Session) { run(alloc.engine, ssn, false) }

// run is one action on the GPU; every action flattens the session as it is NOW, like the reference runs its
// actions one after the other on the same *framework.Session (scheduler.go:88-92).
func run(engine *Engine, ssn *framework.Session, backfill bool) {
	flat, tiers, err := Flatten(ssn) // canonical orders + label/taint/port atom interning; see flatten.go
	if err != nil {
		// e.g. a non built-in plugin registered a PredicateFn: no CPU fallback — skip the cycle loudly.
		glog.Errorf("kbgpu: session cannot be flattened: %v", err)
		return
	}
	if err := engine.Load(flat, tiers); err != nil {
		glog.Errorf("kbgpu: %v", err)
		return
	}
	var dec []Decision
	if backfill {
		dec, err = engine.Backfill(len(flat.Tasks))
	} else {
		dec, err = engine.Allocate(len(flat.Tasks))
	}
	if err != nil {
		glog.Errorf("kbgpu: %v", err)
		return
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
		task := flat.Tasks[i] // *api.TaskInfo
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
	_ = api.Pending
}
