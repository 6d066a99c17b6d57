package kbgpu

import "github.com/kubernetes-sigs/kube-batch/pkg/scheduler/framework"

// backfillAction is a drop-in for pkg/scheduler/actions/backfill (backfill.go:28-74): same Name(), same
// framework.Action interface; Execute becomes flatten -> kb_session_load -> kb_backfill -> replay through ssn.Allocate.
type backfillAction struct{ engine *Engine }

func NewBackfill(e *Engine) *backfillAction        { return &backfillAction{engine: e} }
func (alloc *backfillAction) Name() string         { return "backfill" }
func (alloc *backfillAction) Initialize()          {}
func (alloc *backfillAction) UnInitialize()        {}
func (alloc *backfillAction) Execute(ssn *framework.Session) { run(alloc.engine, ssn, true) }
