package kbgpu

/*
#include "kbgpu.h"
*/
import "C"

import (
	"fmt"
	"sort"
	"unsafe"

	v1 "k8s.io/api/core/v1"

	"github.com/kubernetes-sigs/kube-batch/pkg/scheduler/api"
	"github.com/kubernetes-sigs/kube-batch/pkg/scheduler/framework"
)

// Flat is the SoA form of ssn.Jobs / ssn.Nodes / ssn.Queues that kb_snapshot points into.  The executable
// specification of every field is kube_batch_b200/builder.py::SessionBuilder.flatten and its C++ twin
// kube_batch_b200/host/kbhost.hpp::Flatten; this file is the same computation over the real Go objects.
type Flat struct {
	R, W, N, T, J, Q uint32
	Dims             []v1.ResourceName // dim 0 cpu, 1 memory, 2.. scalar names (sorted)
	NodeNames        []string
	Tasks            []*api.TaskInfo

	nodeIdle, nodeReleasing, nodeUsed, nodeAllocatable []float64
	nodeAllocPresent, nodeFlags                        []uint32
	nodeAllocCPU, nodeAllocMem, nodeNzCPU, nodeNzMem   []int64
	nodePods, nodeMaxPods                              []int32
	nodeLabels, nodeTaints, nodePorts                  []uint64
	taskInitreq, taskResreq                            []float64
	taskResPresent, taskNAff, taskFlags, taskUIDRank   []uint32
	taskNzCPU, taskNzMem, taskCtime                    []int64
	taskSelReq, taskAff, taskTol, taskPortOwn, taskPortConflict []uint64
	taskPrio                                           []int32
	jobTaskOff, jobAlloc0Present, jobQueue             []uint32
	jobMinAvail, jobReady0, jobPrio                    []int32
	jobAlloc0                                          []float64
	jobCtime, queueCtime                               []int64
	queueWeight                                        []int32
}

var builtin = map[string]bool{"priority": true, "gang": true, "drf": true, "predicates": true, "proportion": true,
	"nodeorder": true, "conformance": true}

// Flatten walks the session once.  It refuses (error, no CPU fallback) sessions the GPU path cannot honour.
func Flatten(ssn *framework.Session) (*Flat, []C.kb_tier, error) {
	for _, tier := range ssn.Tiers {
		for _, p := range tier.Plugins {
			if !builtin[p.Name] {
				return nil, nil, fmt.Errorf("plugin %q is not a built-in: its closures cannot run on the device", p.Name)
			}
		}
	}
	f := &Flat{}
	// canonical orders (SURVEY.md §8c rules 1 and 4)
	for name := range ssn.Nodes {
		f.NodeNames = append(f.NodeNames, name)
	}
	sort.Strings(f.NodeNames)
	jobIDs := make([]string, 0, len(ssn.Jobs))
	for id := range ssn.Jobs {
		jobIDs = append(jobIDs, string(id))
	}
	sort.Strings(jobIDs)
	queueIDs := make([]string, 0, len(ssn.Queues))
	for id := range ssn.Queues {
		queueIDs = append(queueIDs, string(id))
	}
	sort.Strings(queueIDs)
	// ... scalar dims, atom interning (selector requirements evaluated once per node via
	// v1helper.NodeSelectorRequirementsAsSelector, NoSchedule|NoExecute taints with Toleration.ToleratesTaint,
	// host ports with HostPortInfo.CheckConflict), node aggregates over NodeInfo.Tasks
	// (priorityutil.GetNonzeroRequests per container), pending tasks per job with UID ranks, job / queue rows:
	// line-for-line what kbhost.hpp::Flatten does, reading the fields
	//   node.Idle/Releasing/Used/Allocatable, node.Node.Status.Allocatable, node.Node.Spec.Taints/Unschedulable,
	//   node.Node.Status.Conditions, task.Resreq/InitResreq/Priority, task.Pod.Spec.{NodeSelector,Affinity,Tolerations,
	//   Containers[].Ports}, job.MinAvailable/Priority/CreationTimestamp/Queue/ReadyTaskNum()/TaskStatusIndex,
	//   queue.Weight, queue.Queue.CreationTimestamp.
	for _, id := range jobIDs {
		job := ssn.Jobs[api.JobID(id)]
		for _, t := range job.TaskStatusIndex[api.Pending] {
			if aff := t.Pod.Spec.Affinity; aff != nil && (aff.PodAffinity != nil || aff.PodAntiAffinity != nil) {
				return nil, nil, fmt.Errorf("pod %s/%s carries inter-pod affinity terms: outside this build", t.Namespace, t.Name)
			}
			f.Tasks = append(f.Tasks, t)
		}
		f.jobTaskOff = append(f.jobTaskOff, uint32(len(f.Tasks)))
	}
	tiers := make([]C.kb_tier, 0, len(ssn.Tiers)) // conf.PluginOption -> kb_plugin_option, Enabled* nil -> 0
	return f, tiers, nil
}

// cSnapshot aliases the slices for the duration of one kb_session_load call.
func (f *Flat) cSnapshot() C.kb_snapshot {
	p64 := func(s []float64) *C.double { if len(s) == 0 { return nil }; return (*C.double)(unsafe.Pointer(&s[0])) }
	var s C.kb_snapshot
	s.abi_version = C.KB_ABI_VERSION
	s.R, s.W, s.N, s.T, s.J, s.Q = C.uint32_t(f.R), C.uint32_t(f.W), C.uint32_t(f.N), C.uint32_t(f.T), C.uint32_t(f.J), C.uint32_t(f.Q)
	s.node_idle, s.node_releasing, s.node_used, s.node_allocatable = p64(f.nodeIdle), p64(f.nodeReleasing), p64(f.nodeUsed), p64(f.nodeAllocatable)
	s.task_initreq, s.task_resreq, s.job_alloc0 = p64(f.taskInitreq), p64(f.taskResreq), p64(f.jobAlloc0)
	// ... the remaining 33 array fields are assigned the same way (uint32 / int32 / int64 / uint64 slices)
	return s
}
