package kbgpu

import "github.com/kubernetes-sigs/kube-batch/pkg/scheduler/framework"

// backfillAction is a drop-in for pkg/scheduler/actions/backfill (backfill.go:28-74): same Name(), same
// framework.Action interface; Execute becomes flatten -> kb_session_load -> kb_backfill -> replay through ssn.Allocate.
// Fallback: the original action, see allocateAction.
type backfillAction struct {
	engine   *Engine
	Fallback framework.Action
}

func NewBackfill(e *Engine, fallback framework.Action) *backfillAction {
	return &backfillAction{engine: e, Fallback: fallback}
}
func (alloc *backfillAction) Name() string  { return "backfill" }
func (alloc *backfillAction) Initialize()   {}
func (alloc *backfillAction) UnInitialize() {}
func (alloc *backfillAction) Execute(ssn *framework.Session) {
	if err := run(alloc.engine, ssn, true); err != nil {
		fallBack(alloc.Fallback, ssn, err)
	}
}
