/*
 * kb_oracle.h — C API of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a C++17 restatement of kube-batch's allocate hot path
 * (pkg/scheduler/actions/allocate/allocate.go:43-194 and everything it calls).  The reference
 * is pure Go and there is no Go toolchain in this image, so the reference itself cannot be
 * compiled or run here; this restatement is pinned against the reference's own golden
 * vectors (tests/test_oracle_golden.py: allocate_test.go cases, resource_info_test.go,
 * node_info_test.go known answers) — everything those tests do not pin (heap behaviour with
 * stale keys, the rand tie-break, plugin arithmetic) is "parity unpinned" by the reference
 * and defined by the deterministic rules of SURVEY.md §8c.
 *
 * Beyond the allocate path it also restates backfill (kb_backfill's oracle) and, ORACLE ONLY so far, NodeAffinityPriority
 * with preferred terms, reclaim and preempt (kbo_cycle) — pinned on preempt_test.go / reclaim_test.go.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  libkbgpu.so never links or calls it.
 */
#ifndef KB_ORACLE_H_
#define KB_ORACLE_H_

#include "../include/kbgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* mode: how the per-(task,node) work is organised (decisions are identical in both). */
#define KBO_MODE_OPTIMISED 0 /* "mode B": cached node aggregates                                         */
#define KBO_MODE_FAITHFUL  1 /* "mode A": reproduces the reference's cost pattern — rebuild the k8s
                                NodeInfo aggregate from the node's pod list per pair
                                (plugins/predicates/predicates.go:124, util/scheduler_helper.go:219-230)
                                and scan every allocated task of every job for the anti-affinity
                                check (plugins/util/util.go:62-85), deep-copying session-placed pods */

typedef struct kbo_opts {
  int32_t mode;          /* KBO_MODE_*                                                            */
  int32_t threads;       /* workers of the PredicateNodes / PrioritizeNodes sweeps; the reference
                            uses 16 (util/scheduler_helper.go:84,137); <=1 = serial                */
  int64_t max_tasks;     /* >0: stop after this many tasks were popped (bounded timing sample)    */
  double  max_seconds;   /* >0: stop once this much wall time has elapsed (bounded timing sample)  */
  int32_t actions;       /* bit 0 allocate (default when 0), bit 1 backfill afterwards on the same session
                            ("allocate, backfill" = the default action list, pkg/scheduler/util.go:31-42)  */
  int32_t warm_tasks;    /* timing samples only: the first warm_tasks tasks run with cached aggregates (mode B), then the
                            session switches to `mode`, the clock starts and max_tasks / max_seconds count from there — a
                            sample of the reference's cost pattern at a later point of the cycle (more allocated pods)   */
} kbo_opts;
#define KBO_ACTION_ALLOCATE 1
#define KBO_ACTION_BACKFILL 2

/* ---- the other two actions of the reference, ORACLE ONLY so far: the engine has no kb_reclaim /
 * kb_preempt yet.  They need what the flattened snapshot only carries as aggregates: the Running tasks, one by one. ---- */
typedef struct kbo_running {
  uint32_t n;                  /* Running tasks (api.Running) that sit on a node of the snapshot                         */
  const uint32_t* node;        /* [n] node index                                                                          */
  const uint32_t* job;         /* [n] job index (the job's job_ready0 / job_alloc0 already count the task)                */
  const double*   resreq;      /* [R][n] TaskInfo.Resreq                                                                  */
  const uint32_t* res_present; /* [n] scalar presence                                                                     */
  const int32_t*  prio;        /* [n] TaskInfo.Priority                                                                   */
  const int64_t*  ctime;       /* [n] Pod.CreationTimestamp                                                               */
  const uint32_t* uid_rank;    /* [n] rank of TaskInfo.UID among the running tasks                                        */
  const uint32_t* flags;       /* [n] bit 0: system-critical priority class or kube-system namespace (conformance.go:45-53) */
} kbo_running;

#define KBO_ACT_RECLAIM  0 /* actions/reclaim/reclaim.go   */
#define KBO_ACT_ALLOCATE 1 /* actions/allocate/allocate.go */
#define KBO_ACT_BACKFILL 2 /* actions/backfill/backfill.go */
#define KBO_ACT_PREEMPT  3 /* actions/preempt/preempt.go   */

typedef struct kbo_result {
  uint64_t pairs_logical;    /* sum over processed tasks of N                                      */
  uint32_t tasks_processed;
  uint32_t tasks_allocated;
  uint32_t tasks_pipelined;
  uint32_t jobs_ready;
  uint32_t visits;
  uint32_t truncated;        /* 1 if max_tasks / max_seconds stopped the cycle early               */
  uint32_t evictions;        /* cache.Evict calls (util.FakeEvictor.Evicts)                        */
  uint32_t timed_tasks;      /* tasks processed after the warm-up (== tasks_processed when warm_tasks == 0)     */
  double   seconds;          /* wall time of Execute (after the warm-up)                           */
} kbo_result;

/* allocateAction.Execute on the flattened snapshot.  out: T decisions.  Optional final state outputs
 * (any may be NULL): node_idle/releasing/used [R][N], node_pods [N], node_nz_cpu/mem [N],
 * node_ports [W][N], job_share [J], job_ready [J], queue_share [Q], queue_deserved/allocated [R][Q]. */
int kbo_allocate(const kb_snapshot* snap, const kb_plugin_conf* conf, const kbo_opts* opts,
                 kb_decision* out, kbo_result* res,
                 double* node_idle, double* node_releasing, double* node_used, int32_t* node_pods,
                 int64_t* node_nz_cpu, int64_t* node_nz_mem, uint64_t* node_ports,
                 double* job_share, int32_t* job_ready, double* queue_share,
                 double* queue_deserved, double* queue_allocated);

/* One scheduling cycle: the actions of `actions[0..n_actions)` in that order on ONE session (scheduler.go:88-101; the
 * shipped configuration is "reclaim, allocate, backfill, preempt", config/kube-batch-conf.yaml:1).  `running` may be NULL
 * (then reclaim / preempt find no victims).  evicted / evict_order [running->n]: cache.Evict calls in call order.
 * The other outputs are those of kbo_allocate. */
int kbo_cycle(const kb_snapshot* snap, const kbo_running* running, const kb_plugin_conf* conf, const kbo_opts* opts,
              const uint8_t* actions, uint32_t n_actions,
              kb_decision* out, uint8_t* evicted, uint32_t* evict_order, kbo_result* res,
              double* node_idle, double* node_releasing, double* node_used, int32_t* node_pods,
              int64_t* node_nz_cpu, int64_t* node_nz_mem, uint64_t* node_ports,
              double* job_share, int32_t* job_ready, double* queue_share,
              double* queue_deserved, double* queue_allocated);

/* predicateFn + PrioritizeNodes of one task against the snapshot's initial node state. */
int kbo_predicate_score(const kb_snapshot* snap, const kb_plugin_conf* conf, uint32_t task,
                        uint8_t* fit /*[N]*/, double* score /*[N]*/);

/* ---- unit-level entry points used to pin the restatement against the reference's golden vectors ---- */
/* api.Resource algebra (api/resource_info.go).  v[R] dense, present = scalar-map key mask (0 <=> nil map). */
int  kbo_res_less_equal(uint32_t R, const double* l, uint32_t lp, const double* r, uint32_t rp);
int  kbo_res_less(uint32_t R, const double* l, uint32_t lp, const double* r, uint32_t rp);
int  kbo_res_is_empty(uint32_t R, const double* l, uint32_t lp);
/* returns 0 ok, -1 if the reference would panic (Sub on insufficient resource, resource_info.go:158) */
int  kbo_res_sub(uint32_t R, double* l, uint32_t* lp, const double* r, uint32_t rp);
void kbo_res_add(uint32_t R, double* l, uint32_t* lp, const double* r, uint32_t rp);
void kbo_res_set_max(uint32_t R, double* l, uint32_t* lp, const double* r, uint32_t rp);
void kbo_res_fit_delta(uint32_t R, double* l, uint32_t* lp, const double* r, uint32_t rp);
/* vendored k8s priorities */
int64_t kbo_least_requested(int64_t req_cpu, int64_t alloc_cpu, int64_t req_mem, int64_t alloc_mem);
int64_t kbo_most_requested(int64_t req_cpu, int64_t alloc_cpu, int64_t req_mem, int64_t alloc_mem);
int64_t kbo_balanced(int64_t req_cpu, int64_t alloc_cpu, int64_t req_mem, int64_t alloc_mem);
/* container/heap restated (util/priority_queue.go over Go's container/heap): pushes keys[0..n) in order
 * with `less` = integer <, then pops all into out. */
void kbo_heap_sort(const int64_t* keys, uint32_t n, int64_t* out);
/* helpers.Share (api/helpers/helpers.go:47-60) */
double kbo_share(double l, double r);
/* util.SelectBestNode (util/scheduler_helper.go:188-208) under the deterministic rule "first max": index of the pick */
uint32_t kbo_select_best_node(const double* scores, uint32_t n);
/* framework.Arguments.GetInt (framework/arguments.go:29-46): `value` NULL = key absent; returns the resulting *ptr */
int kbo_arguments_get_int(const char* value, int base);

/* ---- inter-pod (anti)affinity on the RAW objects (labels, namespaces, selectors, terms) — the oracle does NOT read
 * kb_snapshot.pod_affinity (the flattener's aggregated form, which the engine consumes); it walks the pods like the reference:
 * predicate InterPodAffinityMatches (vendor/.../predicates/predicates.go:1261-1572, slow path) over util.PodLister
 * (plugins/util/util.go:37-85), priority CalculateInterPodAffinityPriority (vendor/.../priorities/interpod_affinity.go:99-235)
 * over nodeInfo.Pods() of the feasible nodes.  Strings are interned to ints by the caller (equal strings <=> equal ints).
 * Pods [0, T) are the snapshot's pending tasks (same index), pods [T, P) the existing pods. ---- */
typedef struct kbo_pod_objects {
  uint32_t P, T;
  const int32_t* pod_ns;          /* [P]                                                          */
  const uint32_t* lab_off;        /* [P+1] -> lab_key / lab_val                                    */
  const int32_t* lab_key; const int32_t* lab_val;
  const uint8_t* has_aff;         /* [P] Spec.Affinity.PodAffinity != nil                          */
  const uint8_t* has_anti;        /* [P] Spec.Affinity.PodAntiAffinity != nil                      */
  const uint32_t* term_off;       /* [P+1] -> terms                                                */
  const int32_t* term_kind;       /* 0 required affinity, 1 required anti-affinity, 2 preferred affinity, 3 preferred anti-affinity */
  const int32_t* term_weight;     /* preferred terms                                               */
  const int32_t* term_topo;       /* topology key id, -1 = ""                                      */
  const uint8_t* term_nil;        /* LabelSelector == nil -> labels.Nothing()                      */
  const uint32_t* term_ns_off;    /* [terms+1] -> term_ns (empty = the owner's namespace)          */
  const int32_t* term_ns;
  const uint32_t* term_req_off;   /* [terms+1] -> requirements                                     */
  const int32_t* req_key; const int32_t* req_op;   /* 0 In, 1 NotIn, 2 Exists, 3 DoesNotExist     */
  const uint32_t* req_val_off;    /* [reqs+1] -> req_val                                           */
  const int32_t* req_val;
  /* existing pods (index p - T) */
  const int32_t* pod_node;        /* node index                                                    */
  const uint8_t* pod_listed;      /* AllocatedStatus task of a session job: util.PodLister lists it */
  const uint8_t* pod_in_tasks;    /* in NodeInfo.Tasks of its node                                 */
  const uint8_t* pod_unbound;     /* Spec.NodeName == ""                                           */
  uint32_t n_topo;
  const int32_t* node_topo;       /* [n_topo][N] value id of the node's label, -1 absent           */
} kbo_pod_objects;
/* Deep-copied; used by the kbo_allocate / kbo_cycle / kbo_predicate_score calls of this thread until cleared with NULL. */
void kbo_set_pod_objects(const kbo_pod_objects* po, uint32_t N);

const char* kbo_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
