// kb_oracle.cpp — CPU ORACLE: a C++17 restatement of kube-batch's allocate hot path.
//
// TEST INFRASTRUCTURE ONLY (see kb_oracle.h).  Structure deliberately follows the Go code —
// Resource / NodeInfo / JobInfo / TaskInfo structs, a Session with per-plugin function
// registries dispatched in tier order, plugins that register closures in OnSessionOpen, Go's
// container/heap restated — so that each function can be checked against the file:line it cites.
// Paths are relative to /root/reference/pkg/scheduler unless they start with vendor/.
//
// Deterministic rules replacing the reference's randomness:
//  (1) nodes in ascending Name order (= snapshot index)          replaces allocate.go:71 map order
//  (2) PredicateNodes preserves that order                        replaces scheduler_helper.go:79-81
//  (3) tie-break = first max in that order                        replaces rand.Intn scheduler_helper.go:190
//  (4) ssn.Jobs iterated in ascending JobID (= snapshot index)    replaces allocate.go:50 map order
//  (5) container/heap restated bit-exactly                        util/priority_queue.go
//  (6) proportion iterates queues in ascending QueueID            replaces proportion.go:104,123 map order
// PARITY UNPINNED by the reference's own tests: heap with stale keys, tie-break, all plugin arithmetic.

#include "kb_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------------------------
// api/resource_info.go
// ---------------------------------------------------------------------------------------------
struct Resource {                       // resource_info.go:28-38
  double v[KB_MAX_R];                   // [0]=MilliCPU [1]=Memory [2..]=ScalarResources (0 when key absent)
  uint32_t present = 0;                 // scalar-map key set; present==0 <=> ScalarResources == nil
  int MaxTaskNum = 0;
  Resource() { for (double& x : v) x = 0; }
};

const double minMilliCPU = 10;                       // resource_info.go:68
const double minMilliScalarResources = 10;           // :69
const double minMemory = 10 * 1024 * 1024;           // :70

struct Algebra {
  uint32_t R;
  explicit Algebra(uint32_t r) : R(r) {}

  // resource_info.go:93-105
  bool IsEmpty(const Resource& r) const {
    if (!(r.v[0] < minMilliCPU && r.v[1] < minMemory)) return false;
    for (uint32_t k = 2; k < R; ++k)
      if ((r.present >> k) & 1u)
        if (r.v[k] >= minMilliScalarResources) return false;
    return true;
  }
  // resource_info.go:128-140
  void Add(Resource& r, const Resource& rr) const {
    r.v[0] += rr.v[0];
    r.v[1] += rr.v[1];
    for (uint32_t k = 2; k < R; ++k)
      if ((rr.present >> k) & 1u) { r.present |= 1u << k; r.v[k] += rr.v[k]; }
  }
  // resource_info.go:268-302
  static bool lessEqualFunc(double l, double r, double diff) { return l < r || std::fabs(l - r) < diff; }
  bool LessEqual(const Resource& r, const Resource& rr) const {
    if (!lessEqualFunc(r.v[0], rr.v[0], minMilliCPU)) return false;
    if (!lessEqualFunc(r.v[1], rr.v[1], minMemory)) return false;
    if (r.present == 0) return true;
    for (uint32_t k = 2; k < R; ++k) {
      if (!((r.present >> k) & 1u)) continue;
      double rQuant = r.v[k];
      if (rQuant <= minMilliScalarResources) continue;
      if (rr.present == 0) return false;
      double rrQuant = ((rr.present >> k) & 1u) ? rr.v[k] : 0.0;
      if (!lessEqualFunc(rQuant, rrQuant, minMilliScalarResources)) return false;
    }
    return true;
  }
  // resource_info.go:143-160; returns false where the reference panics
  bool Sub(Resource& r, const Resource& rr) const {
    if (!LessEqual(rr, r)) return false;
    r.v[0] -= rr.v[0];
    r.v[1] -= rr.v[1];
    for (uint32_t k = 2; k < R; ++k) {
      if (!((rr.present >> k) & 1u)) continue;
      if (r.present == 0) return true;             // `return r` inside the loop
      r.present |= 1u << k;                        // map index assignment creates the key
      r.v[k] -= rr.v[k];
    }
    return true;
  }
  // resource_info.go:163-188
  void SetMaxResource(Resource& r, const Resource& rr) const {
    if (rr.v[0] > r.v[0]) r.v[0] = rr.v[0];
    if (rr.v[1] > r.v[1]) r.v[1] = rr.v[1];
    for (uint32_t k = 2; k < R; ++k) {
      if (!((rr.present >> k) & 1u)) continue;
      if (r.present == 0) {
        for (uint32_t j = 2; j < R; ++j)
          if ((rr.present >> j) & 1u) { r.present |= 1u << j; r.v[j] = rr.v[j]; }
        return;
      }
      double cur = ((r.present >> k) & 1u) ? r.v[k] : 0.0;
      if (rr.v[k] > cur) { r.present |= 1u << k; r.v[k] = rr.v[k]; }
    }
  }
  // resource_info.go:194-214
  void FitDelta(Resource& r, const Resource& rr) const {
    if (rr.v[0] > 0) r.v[0] -= rr.v[0] + minMilliCPU;
    if (rr.v[1] > 0) r.v[1] -= rr.v[1] + minMemory;
    for (uint32_t k = 2; k < R; ++k) {
      if (!((rr.present >> k) & 1u)) continue;
      if (rr.v[k] > 0) { r.present |= 1u << k; r.v[k] -= rr.v[k] + minMilliScalarResources; }
    }
  }
  // resource_info.go:217-224
  void Multi(Resource& r, double ratio) const {
    r.v[0] = r.v[0] * ratio;
    r.v[1] = r.v[1] * ratio;
    for (uint32_t k = 2; k < R; ++k)
      if ((r.present >> k) & 1u) r.v[k] = r.v[k] * ratio;
  }
  // resource_info.go:227-265
  bool Less(const Resource& r, const Resource& rr) const {
    if (!(r.v[0] < rr.v[0])) return false;
    if (!(r.v[1] < rr.v[1])) return false;
    if (r.present == 0) {
      if (rr.present != 0)
        for (uint32_t k = 2; k < R; ++k)
          if (((rr.present >> k) & 1u) && rr.v[k] <= minMilliScalarResources) return false;
      return true;
    }
    if (rr.present == 0) return false;
    for (uint32_t k = 2; k < R; ++k) {
      if (!((r.present >> k) & 1u)) continue;
      double rrQuant = ((rr.present >> k) & 1u) ? rr.v[k] : 0.0;
      if (!(r.v[k] < rrQuant)) return false;
    }
    return true;
  }
  // resource_info.go:305-337
  void Diff(const Resource& r, const Resource& rr, Resource& inc, Resource& dec) const {
    inc = Resource(); dec = Resource();
    if (r.v[0] > rr.v[0]) inc.v[0] += r.v[0] - rr.v[0]; else dec.v[0] += rr.v[0] - r.v[0];
    if (r.v[1] > rr.v[1]) inc.v[1] += r.v[1] - rr.v[1]; else dec.v[1] += rr.v[1] - r.v[1];
    for (uint32_t k = 2; k < R; ++k) {
      if (!((r.present >> k) & 1u)) continue;
      double rrQuant = ((rr.present >> k) & 1u) ? rr.v[k] : 0.0;
      if (r.v[k] > rrQuant) { inc.present |= 1u << k; inc.v[k] += r.v[k] - rrQuant; }
      else { dec.present |= 1u << k; dec.v[k] += rrQuant - r.v[k]; }
    }
  }
  // api/helpers/helpers.go:28-44
  Resource Min(const Resource& l, const Resource& r) const {
    Resource res;
    res.v[0] = std::fmin(l.v[0], r.v[0]);
    res.v[1] = std::fmin(l.v[1], r.v[1]);
    if (l.present == 0 || r.present == 0) return res;
    for (uint32_t k = 2; k < R; ++k)
      if ((l.present >> k) & 1u) {
        res.present |= 1u << k;
        res.v[k] = std::fmin(l.v[k], ((r.present >> k) & 1u) ? r.v[k] : 0.0);
      }
    return res;
  }
};

// api/helpers/helpers.go:47-60
double Share(double l, double r) {
  double share;
  if (r == 0) { if (l == 0) share = 0; else share = 1; }
  else share = l / r;
  return share;
}

// ---------------------------------------------------------------------------------------------
// api/types.go:20-54, api/helpers.go:64-71
// ---------------------------------------------------------------------------------------------
enum TaskStatus { Pending = 1, Allocated = 2, Pipelined = 4, Binding = 8, Bound = 16, Running = 32,
                  Releasing = 64, Succeeded = 128, Failed = 256, Unknown = 512 };
bool AllocatedStatus(int s) { return s == Bound || s == Binding || s == Running || s == Allocated; }

struct TaskInfo {                       // api/job_info.go:36-54
  uint32_t idx = 0;                     // snapshot index == UID identity
  uint32_t Job = 0;
  Resource Resreq, InitResreq;
  int NodeName = -1;
  int Status = Pending;
  int32_t Priority = 0;
  int64_t ctime = 0;
  uint32_t uid_rank = 0;
  // "Pod" fields the predicates / priorities read
  int64_t nz_cpu = 0, nz_mem = 0;
  uint64_t sel_req[KB_MAX_W], tol[KB_MAX_W], port_own[KB_MAX_W], port_conflict[KB_MAX_W];
  uint64_t aff[KB_MAX_AFF_TERMS][KB_MAX_W];
  uint32_t n_aff = 0, flags = 0;
  // nodeAffinity.preferredDuringSchedulingIgnoredDuringExecution: requirement atoms + weight per term
  uint64_t pref[KB_MAX_PREF_TERMS][KB_MAX_W];
  int32_t pref_w[KB_MAX_PREF_TERMS];
  uint32_t n_pref = 0;
  // bookkeeping for the decision output
  uint32_t step = 0xFFFFFFFFu, dispatch_step = 0xFFFFFFFFu;
  bool dispatched = false;
  // running tasks (kbo_running): conformance-critical pod, eviction record (cache.Evict == FakeEvictor here)
  bool critical = false;
  bool evicted = false;
  uint32_t evict_order = 0xFFFFFFFFu;
};

struct PodStub {                        // what NewNodeInfo(node.Pods()...) re-aggregates per pod in mode A
  int64_t nz_cpu, nz_mem;
  uint64_t ports[KB_MAX_W];
  char payload[96];                     // stands in for the v1.Pod fields calculateResource walks
};

struct NodeInfo {                       // api/node_info.go:28-47
  uint32_t idx = 0;
  Resource Releasing, Idle, Used, Allocatable;
  std::map<uint32_t, TaskInfo> Tasks;   // session-added clones (node_info.go:186), key = task idx
  // k8s-side facts of v1.Node
  int64_t alloc_cpu = 0, alloc_mem = 0;
  int32_t max_pods = 0;
  uint32_t flags = 0;
  uint64_t labels[KB_MAX_W], taints[KB_MAX_W];
  // aggregate over ALL tasks on the node (pre-existing + session-added) — cached for mode B,
  // rebuilt per pair in mode A from `existing` + Tasks
  int32_t pods = 0;
  int64_t nz_cpu = 0, nz_mem = 0;
  uint64_t ports[KB_MAX_W];
  std::vector<PodStub> existing;        // mode A only
};

struct JobInfo {                        // api/job_info.go:127-154
  uint32_t idx = 0;
  uint32_t Queue = 0;
  int32_t Priority = 0;
  int32_t MinAvailable = 0;
  int64_t ctime = 0;
  std::map<int, std::set<uint32_t>> TaskStatusIndex;   // status -> task idx
  int32_t ready0 = 0;                   // tasks already Binding/Bound/Running/Succeeded at session open
  Resource Allocated;
};

struct QueueInfo { uint32_t idx = 0; int32_t Weight = 0; int64_t ctime = 0; };

// ---------------------------------------------------------------------------------------------
// util/priority_queue.go over Go's container/heap (go1.13 src/container/heap/heap.go), restated
// ---------------------------------------------------------------------------------------------
template <class T>
struct PriorityQueue {
  std::vector<T> items;
  std::function<bool(const T&, const T&)> lessFn;
  explicit PriorityQueue(std::function<bool(const T&, const T&)> f = nullptr) : lessFn(std::move(f)) {}
  bool less(int i, int j) const { return lessFn(items[i], items[j]); }
  void up(int j) {
    for (;;) {
      int i = (j - 1) / 2;  // parent
      if (i == j || !less(j, i)) break;
      std::swap(items[i], items[j]);
      j = i;
    }
  }
  bool down(int i0, int n) {
    int i = i0;
    for (;;) {
      int j1 = 2 * i + 1;
      if (j1 >= n || j1 < 0) break;
      int j = j1;
      int j2 = j1 + 1;
      if (j2 < n && less(j2, j1)) j = j2;
      if (!less(j, i)) break;
      std::swap(items[i], items[j]);
      i = j;
    }
    return i > i0;
  }
  void Push(const T& x) { items.push_back(x); up((int)items.size() - 1); }
  T Pop() {
    int n = (int)items.size() - 1;
    std::swap(items[0], items[n]);
    down(0, n);
    T it = items.back();
    items.pop_back();
    return it;
  }
  bool Empty() const { return items.empty(); }
  int Len() const { return (int)items.size(); }
};

// ---------------------------------------------------------------------------------------------
// a tiny worker pool standing in for workqueue.ParallelizeUntil(ctx, 16, n, fn)
// (vendor/k8s.io/client-go/util/workqueue/parallelizer.go:30-63)
// ---------------------------------------------------------------------------------------------
class Pool {
 public:
  explicit Pool(int workers) : n_(workers) {
    for (int i = 1; i < n_; ++i) th_.emplace_back([this] { loop(); });
  }
  ~Pool() {
    stop_.store(true);
    gen_.fetch_add(1);
    for (auto& t : th_) t.join();
  }
  void parallelize(int pieces, const std::function<void(int)>& fn) {
    if (n_ <= 1 || pieces < 64) { for (int i = 0; i < pieces; ++i) fn(i); return; }
    fn_ = &fn; pieces_ = pieces; next_.store(0); done_.store(0);
    gen_.fetch_add(1, std::memory_order_release);
    work();
    while (done_.load(std::memory_order_acquire) < n_ - 1) std::this_thread::yield();
  }
 private:
  void work() {
    const int chunk = 64;
    for (;;) {
      int lo = next_.fetch_add(chunk);
      if (lo >= pieces_) break;
      int hi = std::min(pieces_, lo + chunk);
      for (int i = lo; i < hi; ++i) (*fn_)(i);
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      uint64_t g;
      int spins = 0;
      while ((g = gen_.load(std::memory_order_acquire)) == seen) {
        if (++spins > 2000) { std::this_thread::yield(); spins = 0; }
      }
      seen = g;
      if (stop_.load()) return;
      work();
      done_.fetch_add(1, std::memory_order_release);
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> next_{0}, done_{0};
  std::atomic<bool> stop_{false};
  const std::function<void(int)>* fn_ = nullptr;
  int pieces_ = 0;
};

// ---------------------------------------------------------------------------------------------
// conf (conf/scheduler_conf.go:33-56) and framework.Arguments (framework/arguments.go:26-60)
// ---------------------------------------------------------------------------------------------
struct PluginOption {
  std::string Name;
  bool EnabledJobOrder = false, EnabledJobReady = false, EnabledJobPipelined = false, EnabledTaskOrder = false,
       EnabledPreemptable = false, EnabledReclaimable = false, EnabledQueueOrder = false, EnabledPredicate = false,
       EnabledNodeOrder = false;
  std::map<std::string, std::string> Arguments;
};
struct Tier { std::vector<PluginOption> Plugins; };

void GetInt(const std::map<std::string, std::string>& a, int* ptr, const std::string& key) {  // arguments.go:29-46
  auto it = a.find(key);
  if (it == a.end() || it->second.empty()) return;
  char* end = nullptr;
  long v = std::strtol(it->second.c_str(), &end, 10);
  if (end == it->second.c_str() || *end != '\0') return;  // strconv.Atoi error -> keep default
  *ptr = (int)v;
}
void GetBool(const std::map<std::string, std::string>& a, bool* ptr, const std::string& key) {  // arguments.go:49-66
  auto it = a.find(key);
  if (it == a.end() || it->second.empty()) return;
  const std::string& s = it->second;  // strconv.ParseBool
  if (s == "1" || s == "t" || s == "T" || s == "TRUE" || s == "true" || s == "True") *ptr = true;
  else if (s == "0" || s == "f" || s == "F" || s == "FALSE" || s == "false" || s == "False") *ptr = false;
}

// ---------------------------------------------------------------------------------------------
// vendor/k8s.io/kubernetes/pkg/scheduler/algorithm/priorities
// ---------------------------------------------------------------------------------------------
const int64_t MaxPriority = 10;  // vendor/.../scheduler/api/types.go

int64_t leastRequestedScore(int64_t requested, int64_t capacity) {  // least_requested.go:49-58
  if (capacity == 0) return 0;
  if (requested > capacity) return 0;
  return ((capacity - requested) * MaxPriority) / capacity;
}
int64_t leastResourceScorer(int64_t req_cpu, int64_t alloc_cpu, int64_t req_mem, int64_t alloc_mem) {  // :36-44
  int64_t nodeScore = 0, weightSum = 0;
  nodeScore += leastRequestedScore(req_cpu, alloc_cpu) * 1; weightSum += 1;
  nodeScore += leastRequestedScore(req_mem, alloc_mem) * 1; weightSum += 1;
  return nodeScore / weightSum;
}
int64_t mostRequestedScore(int64_t requested, int64_t capacity) {  // most_requested.go:52-61
  if (capacity == 0) return 0;
  if (requested > capacity) return 0;
  return (requested * MaxPriority) / capacity;
}
int64_t mostResourceScorer(int64_t req_cpu, int64_t alloc_cpu, int64_t req_mem, int64_t alloc_mem) {  // :34-43
  int64_t nodeScore = 0, weightSum = 0;
  nodeScore += mostRequestedScore(req_cpu, alloc_cpu) * 1; weightSum += 1;
  nodeScore += mostRequestedScore(req_mem, alloc_mem) * 1; weightSum += 1;
  return nodeScore / weightSum;
}
double fractionOfCapacity(int64_t requested, int64_t capacity) {  // balanced_resource_allocation.go:74-79
  if (capacity == 0) return 1;
  return (double)requested / (double)capacity;
}
int64_t balancedResourceScorer(int64_t req_cpu, int64_t alloc_cpu, int64_t req_mem, int64_t alloc_mem) {  // :42-72
  double cpuFraction = fractionOfCapacity(req_cpu, alloc_cpu);
  double memoryFraction = fractionOfCapacity(req_mem, alloc_mem);
  if (cpuFraction >= 1 || memoryFraction >= 1) return 0;
  // BalanceAttachedNodeVolumes feature gate is off by default -> two-fraction branch
  double diff = std::fabs(cpuFraction - memoryFraction);
  return (int64_t)((1 - diff) * (double)MaxPriority);
}

// ---------------------------------------------------------------------------------------------
// framework.Session (framework/session.go:37-61) + dispatchers (framework/session_plugins.go)
// ---------------------------------------------------------------------------------------------
struct PriorityConfig {                 // vendor/.../priorities/types.go:46-54 (Map [+ Reduce] configs on this path)
  std::string Name;
  std::function<int64_t(const TaskInfo&, const NodeInfo&, int64_t nz_cpu, int64_t nz_mem)> Map;
  int Weight = 0;
  std::function<void(std::vector<int64_t>&)> Reduce;   // over the FEASIBLE nodes' Map results (scheduler_helper.go:139-156)
};

// priorities.NormalizeReduce(maxPriority, reverse) (vendor/.../priorities/reduce.go:28-63)
inline std::function<void(std::vector<int64_t>&)> NormalizeReduce(int64_t maxPriority, bool reverse) {
  return [=](std::vector<int64_t>& result) {
    int64_t maxCount = 0;
    for (int64_t v : result) if (v > maxCount) maxCount = v;
    if (maxCount == 0) {
      if (reverse) for (int64_t& v : result) v = maxPriority;
      return;
    }
    for (int64_t& v : result) {
      int64_t score = maxPriority * v / maxCount;
      if (reverse) score = maxPriority - score;
      v = score;
    }
  };
}

// ---------------------------------------------------------------------------------------------
// inter-pod (anti)affinity on the raw pod objects (kbo_pod_objects)
// ---------------------------------------------------------------------------------------------
struct LabelReq { int key, op; std::vector<int> vals; };
struct AffTerm { int kind, weight, topo; bool nil; std::vector<int> ns; std::vector<LabelReq> reqs; };
struct PodObj { int ns = 0; std::vector<std::pair<int, int>> labels; bool has_aff = false, has_anti = false; std::vector<AffTerm> terms; };
struct ExistingPod { uint32_t pod; int node; bool listed, in_tasks, unbound; };
struct PodWorld {
  uint32_t T = 0, N = 0, n_topo = 0;
  std::vector<PodObj> pods;
  std::vector<ExistingPod> existing;
  std::vector<int32_t> node_topo;      // [n_topo][N]
  std::vector<std::vector<uint32_t>> existing_on;   // per node: indices into `existing` with in_tasks
  int topo(int key, uint32_t node) const { return key < 0 ? -1 : node_topo[(size_t)key * N + node]; }
  // labels.Requirement.Matches (vendor/k8s.io/apimachinery/pkg/labels/selector.go:193-230)
  static bool req_matches(const LabelReq& r, const PodObj& p) {
    const int* val = nullptr;
    for (auto& kv : p.labels) if (kv.first == r.key) { val = &kv.second; break; }
    switch (r.op) {
      case 0: if (!val) return false; for (int v : r.vals) if (v == *val) return true; return false;        // In
      case 1: if (!val) return true; for (int v : r.vals) if (v == *val) return false; return true;         // NotIn
      case 2: return val != nullptr;                                                                        // Exists
      default: return val == nullptr;                                                                       // DoesNotExist
    }
  }
  // priorityutil.PodMatchesTermsNamespaceAndSelector (vendor/.../priorities/util/topologies.go:38-49) with
  // GetNamespacesFromPodAffinityTerm (:25-36) and metav1.LabelSelectorAsSelector (nil -> Nothing, empty -> Everything)
  bool matches(const PodObj& owner, const AffTerm& t, const PodObj& pod) const {
    bool ns_ok = false;
    if (t.ns.empty()) ns_ok = pod.ns == owner.ns;
    else for (int n : t.ns) if (n == pod.ns) { ns_ok = true; break; }
    if (!ns_ok) return false;
    if (t.nil) return false;
    for (auto& r : t.reqs) if (!req_matches(r, pod)) return false;
    return true;
  }
  // NodesHaveSameTopologyKey (topologies.go:51-70)
  bool same_topology(uint32_t a, uint32_t b, int key) const {
    if (key < 0) return false;
    const int va = topo(key, a), vb = topo(key, b);
    return va >= 0 && vb >= 0 && va == vb;
  }
};
thread_local std::shared_ptr<PodWorld> g_pod_world;

struct Session;
struct Plugin {
  virtual ~Plugin() = default;
  virtual std::string Name() const = 0;
  virtual void OnSessionOpen(Session* ssn) = 0;
};

struct Session {
  Algebra A;
  uint32_t R, W;
  std::vector<JobInfo> Jobs;           // index = ascending JobID
  std::vector<NodeInfo> Nodes;         // index = ascending Name
  std::vector<QueueInfo> Queues;       // index = ascending QueueID
  std::vector<TaskInfo> Tasks;         // the TaskInfo objects the job maps point at
  std::vector<Tier> Tiers;
  std::map<std::string, std::unique_ptr<Plugin>> plugins;

  std::map<std::string, std::function<int(const JobInfo&, const JobInfo&)>> jobOrderFns;
  std::map<std::string, std::function<int(const QueueInfo&, const QueueInfo&)>> queueOrderFns;
  std::map<std::string, std::function<int(const TaskInfo&, const TaskInfo&)>> taskOrderFns;
  std::map<std::string, std::function<bool(const TaskInfo&, const NodeInfo&, int32_t pods, const uint64_t* ports)>> predicateFns;
  std::map<std::string, std::function<bool(const QueueInfo&)>> overusedFns;
  std::map<std::string, std::function<bool(const JobInfo&)>> jobReadyFns;
  std::map<std::string, std::vector<PriorityConfig>> nodePrioritizers;
  std::vector<std::function<void(const TaskInfo&)>> allocateHandlers;   // EventHandler.AllocateFunc
  std::vector<std::function<void(const TaskInfo&)>> deallocateHandlers; // EventHandler.DeallocateFunc
  // victims functions: (preemptor / reclaimer, candidate task ids) -> victim task ids, in candidate order
  using VictimFn = std::function<std::vector<uint32_t>(const TaskInfo&, const std::vector<uint32_t>&)>;
  std::map<std::string, VictimFn> preemptableFns, reclaimableFns;
  std::map<std::string, std::function<bool(const JobInfo&)>> jobPipelinedFns;
  uint32_t n_evicted = 0;               // cache.Evict calls (util.FakeEvictor.Evicts)
  uint32_t n_allocated_nowhere = 0;     // tasks ssn.Allocate left Allocated on no node (see Allocate): the predicates plugin fails from then on

  // resolved once after OnSessionOpen: the (tier, plugin) walk of PredicateFn with its map lookups hoisted
  std::vector<const std::function<bool(const TaskInfo&, const NodeInfo&, int32_t, const uint64_t*)>*> resolvedPredicates;
  bool predicates_plugin_enabled = false;

  int mode = KBO_MODE_OPTIMISED;
  uint32_t step_counter = 0;
  uint32_t n_allocated = 0, n_pipelined = 0;
  // mode A: node of every pre-existing AllocatedStatus task per job (what PodLister.FilteredList walks)
  std::vector<std::vector<uint32_t>> placeholder_alloc;
  std::unordered_map<std::string, uint32_t> nodeByName;   // ssn.Nodes map[string]*NodeInfo, mode A only
  std::vector<std::string> nodeNames;

  std::shared_ptr<PodWorld> pw;         // raw pod objects: inter-pod (anti)affinity is evaluated iff present

  // util.PodLister.FilteredList (plugins/util/util.go:62-85): every AllocatedStatus task of every session job, with the node
  // TaskInfo.NodeName names.  nodeInfo.Filter passes everything here (a listed pod is in its node's NodeInfo).  fn(pod, node);
  // returns false if fn did, or if a listed task has no node: a task ssn.Allocate left Allocated although node.AddTask refused it
  // (session.go:241-262) keeps NodeName "" and CachedNodeInfo.GetNodeInfo("") is an error that fails the predicate (:1386-1393).
  template <class F>
  bool for_each_listed_pod(F fn) const {
    for (const ExistingPod& e : pw->existing) if (e.listed) if (!fn(pw->pods[e.pod], (uint32_t)e.node)) return false;
    for (const JobInfo& job : Jobs)
      for (auto& kv : job.TaskStatusIndex) {
        if (!AllocatedStatus(kv.first)) continue;
        for (uint32_t id : kv.second) {
          if (id >= pw->T) continue;                       // Running tasks of kbo_running are the `existing` pods above
          const TaskInfo& t = Tasks[id];
          if (t.NodeName < 0) return false;
          if (!fn(pw->pods[id], (uint32_t)t.NodeName)) return false;
        }
      }
    return true;
  }

  // PodAffinityChecker.InterPodAffinityMatches (vendor/.../predicates/predicates.go:1261-1288), meta == nil
  bool InterPodAffinityMatches(const TaskInfo& task, const NodeInfo& node) const {
    const PodWorld& W_ = *pw;
    const PodObj& pod = W_.pods[task.idx];
    const uint32_t n = node.idx;
    // satisfiesExistingPodsAntiAffinity (:1400-1439): topology pairs of the existing pods' anti-affinity terms that match `pod`
    // (getMatchingAntiAffinityTopologyPairsOfPod :1354-1376), then any label of the node among them rejects it
    bool bad = false;
    bool ok = for_each_listed_pod([&](const PodObj& e, uint32_t en) {
      if (!e.has_anti) return true;
      for (const AffTerm& t : e.terms) {
        if (t.kind != 1) continue;
        if (!W_.matches(e, t, pod)) continue;
        const int v = W_.topo(t.topo, en);
        if (v >= 0 && W_.topo(t.topo, n) == v) bad = true;          // the node carries the pair (key, value)
      }
      return true;
    });
    if (!ok || bad) return false;
    if (!pod.has_aff && !pod.has_anti) return true;                   // :1274-1277
    // satisfiesPodsAffinityAntiAffinity, slow path (:1516-1562)
    std::vector<const AffTerm*> affinityTerms, antiAffinityTerms;
    for (const AffTerm& t : pod.terms) { if (t.kind == 0) affinityTerms.push_back(&t); else if (t.kind == 1) antiAffinityTerms.push_back(&t); }
    // podMatchesPodAffinityTerms (:1296-1320): (matches all terms + topologies, matches all term properties)
    auto podMatches = [&](const PodObj& target, uint32_t tn, const std::vector<const AffTerm*>& terms, bool* props) {
      *props = false;
      for (const AffTerm* t : terms) if (!W_.matches(pod, *t, target)) return false;   // podMatchesAllAffinityTermProperties
      *props = true;
      for (const AffTerm* t : terms) if (!W_.same_topology(n, tn, t->topo)) return false;
      return true;
    };
    bool matchFound = false, termsSelectorMatchFound = false, rejected = false;
    ok = for_each_listed_pod([&](const PodObj& target, uint32_t tn) {
      if (!matchFound && !affinityTerms.empty()) {
        bool props = false;
        const bool m = podMatches(target, tn, affinityTerms, &props);
        if (props) termsSelectorMatchFound = true;
        if (m) matchFound = true;
      }
      if (!antiAffinityTerms.empty()) {
        bool props = false;
        if (podMatches(target, tn, antiAffinityTerms, &props)) { rejected = true; return false; }
      }
      return true;
    });
    if (rejected) return false;
    if (!ok) return false;
    if (!matchFound && !affinityTerms.empty()) {
      if (termsSelectorMatchFound) return false;
      // targetPodMatchesAffinityOfPod(pod, pod) (metadata.go:767-778): the first pod of a self-affine series
      if (!pod.has_aff) return false;
      for (const AffTerm* t : affinityTerms) if (!W_.matches(pod, *t, pod)) return false;
    }
    return true;
  }

  // InterPodAffinity.CalculateInterPodAffinityPriority (vendor/.../priorities/interpod_affinity.go:99-235) over the feasible
  // nodes; hardPodAffinityWeight = v1.DefaultHardPodAffinitySymmetricWeight = 1 (nodeorder.go:159).  Returns the scores 0..10.
  std::vector<int64_t> InterPodAffinityPriority(const TaskInfo& task, const std::vector<uint32_t>& nodes) const {
    const PodWorld& W_ = *pw;
    const PodObj& pod = W_.pods[task.idx];
    std::vector<int64_t> counts(nodes.size(), 0);
    // cachedNodeInfo.GetNodeInfo (plugins/nodeorder/nodeorder.go:49-63) for a pod whose Spec.NodeName is "": the first node of
    // the session holding ANY pod with an empty Spec.NodeName (a Go map walk there; ascending node order here)
    int first_unbound = -1;
    for (const NodeInfo& ni : Nodes) {
      bool any = false;
      for (auto& kv : ni.Tasks) if (kv.first < W_.T) { any = true; break; }      // placed this session: the pod object is unchanged
      if (!any) for (uint32_t ei : W_.existing_on[ni.idx]) if (W_.existing[ei].unbound) { any = true; break; }
      if (any) { first_unbound = (int)ni.idx; break; }
    }
    auto processTerm = [&](const AffTerm& term, const PodObj& defining, const PodObj& toCheck, uint32_t fixedNode, int64_t weight) {
      if (!W_.matches(defining, term, toCheck)) return;
      for (size_t i = 0; i < nodes.size(); ++i) if (W_.same_topology(nodes[i], fixedNode, term.topo)) counts[i] += weight;
    };
    auto processPod = [&](const PodObj& existing, bool unbound, uint32_t host) {
      const uint32_t existingPodNode = unbound ? (uint32_t)first_unbound : host;
      if (pod.has_aff) for (const AffTerm& t : pod.terms) if (t.kind == 2) processTerm(t, pod, existing, existingPodNode, t.weight);
      if (pod.has_anti) for (const AffTerm& t : pod.terms) if (t.kind == 3) processTerm(t, pod, existing, existingPodNode, -(int64_t)t.weight);
      if (existing.has_aff) {
        for (const AffTerm& t : existing.terms) if (t.kind == 0) processTerm(t, existing, pod, existingPodNode, 1);
        for (const AffTerm& t : existing.terms) if (t.kind == 2) processTerm(t, existing, pod, existingPodNode, t.weight);
      }
      if (existing.has_anti) for (const AffTerm& t : existing.terms) if (t.kind == 3) processTerm(t, existing, pod, existingPodNode, -(int64_t)t.weight);
    };
    const bool all = pod.has_aff || pod.has_anti;             // else only nodeInfo.PodsWithAffinity()
    for (uint32_t m : nodes) {                                 // nodeNameToInfo holds the feasible nodes only (scheduler_helper.go:219-230)
      for (uint32_t ei : W_.existing_on[m]) {
        const ExistingPod& e = W_.existing[ei];
        const PodObj& x = W_.pods[e.pod];
        if (all || x.has_aff || x.has_anti) processPod(x, e.unbound, m);
      }
      for (auto& kv : Nodes[m].Tasks) {
        if (kv.first >= W_.T) continue;
        const PodObj& x = W_.pods[kv.first];
        if (all || x.has_aff || x.has_anti) processPod(x, true, m);
      }
    }
    int64_t maxCount = 0, minCount = 0;
    for (int64_t c : counts) { if (c > maxCount) maxCount = c; if (c < minCount) minCount = c; }
    std::vector<int64_t> result(nodes.size(), 0);
    if (maxCount - minCount > 0)
      for (size_t i = 0; i < nodes.size(); ++i) {
        const double fScore = 10.0 * ((double)(counts[i] - minCount) / (double)(maxCount - minCount));
        result[i] = (int64_t)fScore;
      }
    return result;
  }

  Session(uint32_t r, uint32_t w) : A(r), R(r), W(w) {}

  // api/job_info.go:383-393
  int32_t ReadyTaskNum(const JobInfo& ji) const {
    int occupid = ji.ready0;
    for (auto& kv : ji.TaskStatusIndex)
      if (AllocatedStatus(kv.first) || kv.first == Succeeded) occupid += (int)kv.second.size();
    return occupid;
  }
  bool Ready(const JobInfo& ji) const { return ReadyTaskNum(ji) >= ji.MinAvailable; }  // job_info.go:423-427

  // session_plugins.go:165-179 — NB: ignores Enabled* flags
  bool Overused(const QueueInfo& queue) const {
    for (auto& tier : Tiers)
      for (auto& plugin : tier.Plugins) {
        auto it = overusedFns.find(plugin.Name);
        if (it == overusedFns.end()) continue;
        if (it->second(queue)) return true;
      }
    return false;
  }
  // session_plugins.go:182-200
  bool JobReady(const JobInfo& job) const {
    for (auto& tier : Tiers)
      for (auto& plugin : tier.Plugins) {
        if (!plugin.EnabledJobReady) continue;
        auto it = jobReadyFns.find(plugin.Name);
        if (it == jobReadyFns.end()) continue;
        if (!it->second(job)) return false;
      }
    return true;
  }
  // session_plugins.go:243-267
  bool JobOrderFn(const JobInfo& l, const JobInfo& r) const {
    for (auto& tier : Tiers)
      for (auto& plugin : tier.Plugins) {
        if (!plugin.EnabledJobOrder) continue;
        auto it = jobOrderFns.find(plugin.Name);
        if (it == jobOrderFns.end()) continue;
        int j = it->second(l, r);
        if (j != 0) return j < 0;
      }
    if (l.ctime == r.ctime) return l.idx < r.idx;   // UID order == index order
    return l.ctime < r.ctime;
  }
  // session_plugins.go:270-295
  bool QueueOrderFn(const QueueInfo& l, const QueueInfo& r) const {
    for (auto& tier : Tiers)
      for (auto& plugin : tier.Plugins) {
        if (!plugin.EnabledQueueOrder) continue;
        auto it = queueOrderFns.find(plugin.Name);
        if (it == queueOrderFns.end()) continue;
        int j = it->second(l, r);
        if (j != 0) return j < 0;
      }
    if (l.ctime == r.ctime) return l.idx < r.idx;
    return l.ctime < r.ctime;
  }
  // session_plugins.go:298-331
  bool TaskOrderFn(const TaskInfo& l, const TaskInfo& r) const {
    for (auto& tier : Tiers)
      for (auto& plugin : tier.Plugins) {
        if (!plugin.EnabledTaskOrder) continue;
        auto it = taskOrderFns.find(plugin.Name);
        if (it == taskOrderFns.end()) continue;
        int j = it->second(l, r);
        if (j != 0) return j < 0;
      }
    if (l.ctime == r.ctime) return l.uid_rank < r.uid_rank;
    return l.ctime < r.ctime;
  }
  // session_plugins.go:334-351 — the tier/plugin walk is resolved once (resolvePredicates) instead of per pair
  void resolvePredicates() {
    resolvedPredicates.clear();
    for (auto& tier : Tiers)
      for (auto& plugin : tier.Plugins) {
        if (!plugin.EnabledPredicate) continue;
        auto it = predicateFns.find(plugin.Name);
        if (it == predicateFns.end()) continue;
        resolvedPredicates.push_back(&it->second);
        if (plugin.Name == "predicates") predicates_plugin_enabled = true;
      }
  }
  bool PredicateFn(const TaskInfo& task, const NodeInfo& node, int32_t pods, const uint64_t* ports) const {
    for (auto* fn : resolvedPredicates)
      if (!(*fn)(task, node, pods, ports)) return false;
    return true;
  }
  // session_plugins.go:354-369
  std::vector<PriorityConfig> NodePrioritizers() const {
    std::vector<PriorityConfig> priorityConfigs;
    for (auto& tier : Tiers)
      for (auto& plugin : tier.Plugins) {
        if (!plugin.EnabledNodeOrder) continue;
        auto it = nodePrioritizers.find(plugin.Name);
        if (it == nodePrioritizers.end()) continue;
        priorityConfigs.insert(priorityConfigs.end(), it->second.begin(), it->second.end());
      }
    return priorityConfigs;
  }

  // api/job_info.go:247-264 (status index + Allocated bookkeeping; TotalRequest is unused on this path)
  void UpdateTaskStatus(JobInfo& job, TaskInfo& task, int status) {
    auto it = job.TaskStatusIndex.find(task.Status);
    if (it != job.TaskStatusIndex.end()) {
      if (AllocatedStatus(task.Status)) A.Sub(job.Allocated, task.Resreq);
      it->second.erase(task.idx);
      if (it->second.empty()) job.TaskStatusIndex.erase(it);
    }
    task.Status = status;
    job.TaskStatusIndex[status].insert(task.idx);
    if (AllocatedStatus(status)) A.Add(job.Allocated, task.Resreq);
  }

  // api/node_info.go:172-212
  bool AddTask(NodeInfo& ni, TaskInfo& task) {
    if (task.NodeName >= 0 && task.NodeName != (int)ni.idx) return false;   // :173-176 "task already on different node"
    if (ni.Tasks.count(task.idx)) return false;
    TaskInfo ti = task;  // clone
    switch (ti.Status) {
      case Releasing:
        if (!A.LessEqual(ti.Resreq, ni.Idle)) return false;
        A.Sub(ni.Idle, ti.Resreq);
        A.Add(ni.Releasing, ti.Resreq);
        break;
      case Pipelined:
        if (!A.Sub(ni.Releasing, ti.Resreq)) { g_err = "panic: Resource is not sufficient to do operation (Releasing.Sub)"; return false; }
        break;
      default:
        if (!A.LessEqual(ti.Resreq, ni.Idle)) return false;   // allocateIdleResource, node_info.go:161-167
        A.Sub(ni.Idle, ti.Resreq);
        break;
    }
    A.Add(ni.Used, ti.Resreq);
    task.NodeName = (int)ni.idx;
    ti.NodeName = (int)ni.idx;
    ni.Tasks[task.idx] = ti;
    // the aggregate k8s NewNodeInfo(node.Pods()...) would compute now includes this pod
    ni.pods += 1;
    ni.nz_cpu += task.nz_cpu;
    ni.nz_mem += task.nz_mem;
    for (uint32_t w = 0; w < W; ++w) ni.ports[w] |= task.port_own[w];
    return true;
  }

  // framework/session.go:290-314 (cache.Bind is the FakeBinder here: record the bind)
  void dispatch(JobInfo& job, TaskInfo& task, uint32_t trigger_step) {
    task.dispatched = true;
    task.dispatch_step = trigger_step;
    UpdateTaskStatus(job, task, Binding);
  }

  // framework/session.go:235-288
  bool Allocate(TaskInfo& task, NodeInfo& node) {
    JobInfo& job = Jobs[task.Job];
    UpdateTaskStatus(job, task, Allocated);
    if (!AddTask(node, task)) {
      // the status changed before node.AddTask refused (session.go:241-262): the task stays Allocated with NodeName "".
      // util.PodLister lists it from now on, and CachedNodeInfo.GetNodeInfo("") (plugins/util/util.go:93-100) is an error that is
      // not apierrors.IsNotFound: getMatchingAntiAffinityTopologyPairsOfPods (vendor/.../predicates.go:1381-1393) returns it and
      // InterPodAffinityMatches fails — for EVERY later (pod, node) pair of the session, affinity terms or not.
      if (task.NodeName < 0) ++n_allocated_nowhere;
      return false;
    }
    task.step = step_counter++;
    ++n_allocated;
    for (auto& eh : allocateHandlers) eh(task);
    if (JobReady(job)) {
      auto it = job.TaskStatusIndex.find(Allocated);
      if (it != job.TaskStatusIndex.end()) {
        std::vector<uint32_t> ids(it->second.begin(), it->second.end());
        for (uint32_t id : ids) dispatch(job, Tasks[id], task.step);
      }
    }
    return true;
  }
  // framework/session.go:194-232
  bool Pipeline(TaskInfo& task, NodeInfo& node) {
    JobInfo& job = Jobs[task.Job];
    UpdateTaskStatus(job, task, Pipelined);
    if (!AddTask(node, task)) return false;
    task.step = step_counter++;
    ++n_pipelined;
    for (auto& eh : allocateHandlers) eh(task);
    return true;
  }

  // api/job_info.go:396-405, :430-434
  int32_t WaitingTaskNum(const JobInfo& ji) const {
    auto it = ji.TaskStatusIndex.find(Pipelined);
    return it == ji.TaskStatusIndex.end() ? 0 : (int32_t)it->second.size();
  }
  bool JobInfoPipelined(const JobInfo& ji) const { return WaitingTaskNum(ji) + ReadyTaskNum(ji) >= ji.MinAvailable; }
  // session_plugins.go:203-221
  bool JobPipelined(const JobInfo& job) const {
    for (auto& tier : Tiers)
      for (auto& plugin : tier.Plugins) {
        if (!plugin.EnabledJobPipelined) continue;
        auto it = jobPipelinedFns.find(plugin.Name);
        if (it == jobPipelinedFns.end()) continue;
        if (!it->second(job)) return false;
      }
    return true;
  }
  // session_plugins.go:80-162 (Reclaimable and Preemptable are the same walk over different registries).  Go's nil slice
  // and "empty" coincide here (every victims slice is built by append from nil), and `init` survives across tiers.
  std::vector<uint32_t> victims_of(bool reclaim, const TaskInfo& evictor, const std::vector<uint32_t>& evictees) const {
    std::vector<uint32_t> victims;
    bool init = false;
    const auto& fns = reclaim ? reclaimableFns : preemptableFns;
    for (auto& tier : Tiers) {
      for (auto& plugin : tier.Plugins) {
        if (!(reclaim ? plugin.EnabledReclaimable : plugin.EnabledPreemptable)) continue;
        auto it = fns.find(plugin.Name);
        if (it == fns.end()) continue;
        std::vector<uint32_t> candidates = it->second(evictor, evictees);
        if (!init) { victims = candidates; init = true; }
        else {
          std::vector<uint32_t> intersection;
          for (uint32_t v : victims) for (uint32_t c : candidates) if (v == c) intersection.push_back(v);
          victims = intersection;
        }
      }
      if (!victims.empty()) return victims;      // "Plugins in this tier made decision if victims is not nil"
    }
    return victims;
  }
  std::vector<uint32_t> Reclaimable(const TaskInfo& r, const std::vector<uint32_t>& c) const { return victims_of(true, r, c); }
  std::vector<uint32_t> Preemptable(const TaskInfo& r, const std::vector<uint32_t>& c) const { return victims_of(false, r, c); }

  // api/node_info.go:214-243; the node holds its own clone, whose status decides what is given back
  bool RemoveTask(NodeInfo& ni, const TaskInfo& ti) {
    auto it = ni.Tasks.find(ti.idx);
    if (it == ni.Tasks.end()) return false;
    const TaskInfo& task = it->second;
    switch (task.Status) {
      case Releasing:
        if (!A.Sub(ni.Releasing, task.Resreq)) { g_err = "panic: Resource is not sufficient to do operation (Releasing.Sub in RemoveTask)"; return false; }
        A.Add(ni.Idle, task.Resreq);
        break;
      case Pipelined:
        A.Add(ni.Releasing, task.Resreq);
        break;
      default:
        A.Add(ni.Idle, task.Resreq);
        break;
    }
    if (!A.Sub(ni.Used, task.Resreq)) { g_err = "panic: Resource is not sufficient to do operation (Used.Sub in RemoveTask)"; return false; }
    ni.pods -= 1; ni.nz_cpu -= task.nz_cpu; ni.nz_mem -= task.nz_mem;
    // host ports of the removed pod: recomputed from the remaining tasks would need the pre-existing pods' ports, which the
    // snapshot only carries as an aggregate; tasks that move through RemoveTask here (evicted Running pods keep their node,
    // un-pipelined preemptors) give their ports back only in the unpipeline case:
    if (task.Status == Pipelined) for (uint32_t w = 0; w < W; ++w) ni.ports[w] &= ~task.port_own[w];
    ni.Tasks.erase(it);
    return true;
  }
  // api/node_info.go:245-259
  bool UpdateTask(NodeInfo& ni, TaskInfo& ti) {
    if (!RemoveTask(ni, ti)) return false;
    const int keep = ti.NodeName;
    if (!AddTask(ni, ti)) { g_err = "glog.Fatalf: Failed to add Task to Node during task update"; ti.NodeName = keep; return false; }
    return true;
  }
  // framework/session.go:317-353 (Session.Evict: cache.Evict first) and framework/statement.go:36-64 (Statement.Evict: cache.Evict
  // only at Commit).  `commit_now` selects between the two.
  bool EvictTask(TaskInfo& reclaimee, bool commit_now) {
    if (commit_now) record_evict(reclaimee);
    JobInfo& job = Jobs[reclaimee.Job];
    UpdateTaskStatus(job, reclaimee, Releasing);
    if (reclaimee.NodeName >= 0) UpdateTask(Nodes[reclaimee.NodeName], reclaimee);
    for (auto& eh : deallocateHandlers) eh(reclaimee);
    return true;
  }
  void record_evict(TaskInfo& t) { t.evicted = true; t.evict_order = n_evicted++; }
};

// framework/statement.go
struct Statement {
  Session& ssn;
  struct Op { bool evict; uint32_t task; };
  std::vector<Op> operations;
  explicit Statement(Session& s) : ssn(s) {}
  void Evict(TaskInfo& reclaimee) {                                       // :36-64
    ssn.EvictTask(reclaimee, false);
    operations.push_back({true, reclaimee.idx});
  }
  void Pipeline(TaskInfo& task, NodeInfo& node) {                         // :110-148 (no volume / NodeName handling here)
    JobInfo& job = ssn.Jobs[task.Job];
    ssn.UpdateTaskStatus(job, task, Pipelined);
    if (ssn.AddTask(node, task)) { task.step = ssn.step_counter++; ++ssn.n_pipelined; }
    for (auto& eh : ssn.allocateHandlers) eh(task);
    operations.push_back({false, task.idx});
  }
  void unevict(TaskInfo& reclaimee) {                                     // :78-107
    JobInfo& job = ssn.Jobs[reclaimee.Job];
    ssn.UpdateTaskStatus(job, reclaimee, Running);
    if (reclaimee.NodeName >= 0) ssn.UpdateTask(ssn.Nodes[reclaimee.NodeName], reclaimee);
    for (auto& eh : ssn.allocateHandlers) eh(reclaimee);
  }
  void unpipeline(TaskInfo& task) {                                       // :153-188
    JobInfo& job = ssn.Jobs[task.Job];
    ssn.UpdateTaskStatus(job, task, Pending);
    if (task.NodeName >= 0) {
      if (ssn.RemoveTask(ssn.Nodes[task.NodeName], task)) { --ssn.n_pipelined; task.step = 0xFFFFFFFFu; }
    }
    for (auto& eh : ssn.deallocateHandlers) eh(task);
  }
  void Discard() {                                                        // :191-203
    for (size_t i = operations.size(); i-- > 0;) {
      TaskInfo& t = ssn.Tasks[operations[i].task];
      if (operations[i].evict) unevict(t); else unpipeline(t);
    }
    operations.clear();
  }
  void Commit() {                                                         // :206-217 (cache.Evict never fails here)
    for (auto& op : operations) if (op.evict) ssn.record_evict(ssn.Tasks[op.task]);
    operations.clear();
  }
};

// ---------------------------------------------------------------------------------------------
// plugins/priority/priority.go:39-101
// ---------------------------------------------------------------------------------------------
struct priorityPlugin : Plugin {
  std::string Name() const override { return "priority"; }
  void OnSessionOpen(Session* ssn) override {
    ssn->taskOrderFns[Name()] = [](const TaskInfo& lv, const TaskInfo& rv) {   // :40-56
      if (lv.Priority == rv.Priority) return 0;
      if (lv.Priority > rv.Priority) return -1;
      return 1;
    };
    ssn->jobOrderFns[Name()] = [](const JobInfo& lv, const JobInfo& rv) {      // :61-77
      if (lv.Priority > rv.Priority) return -1;
      if (lv.Priority < rv.Priority) return 1;
      return 0;
    };
    ssn->preemptableFns[Name()] = [ssn](const TaskInfo& preemptor, const std::vector<uint32_t>& preemptees) {   // :81-100
      const JobInfo& preemptorJob = ssn->Jobs[preemptor.Job];
      std::vector<uint32_t> victims;
      for (uint32_t id : preemptees)
        if (!(ssn->Jobs[ssn->Tasks[id].Job].Priority >= preemptorJob.Priority)) victims.push_back(id);
      return victims;
    };
  }
};

// ---------------------------------------------------------------------------------------------
// plugins/gang/gang.go:47-130
// ---------------------------------------------------------------------------------------------
struct gangPlugin : Plugin {
  std::string Name() const override { return "gang"; }
  void OnSessionOpen(Session* ssn) override {
    // validJobFn (:48-69) is registered but dead at this commit (session.go:89-108 runs before Tiers is set).
    ssn->jobOrderFns[Name()] = [ssn](const JobInfo& lv, const JobInfo& rv) {    // :96-119
      bool lReady = ssn->Ready(lv), rReady = ssn->Ready(rv);
      if (lReady && rReady) return 0;
      if (lReady) return 1;
      if (rReady) return -1;
      return 0;
    };
    ssn->jobReadyFns[Name()] = [ssn](const JobInfo& ji) { return ssn->Ready(ji); };  // :122-125
    auto preemptableFn = [ssn](const TaskInfo&, const std::vector<uint32_t>& preemptees) {                // :70-90
      std::vector<uint32_t> victims;
      for (uint32_t id : preemptees) {
        const JobInfo& job = ssn->Jobs[ssn->Tasks[id].Job];
        const int32_t occupid = ssn->ReadyTaskNum(job);
        const bool preemptable = job.MinAvailable <= occupid - 1 || job.MinAvailable == 1;
        if (preemptable) victims.push_back(id);
      }
      return victims;
    };
    ssn->reclaimableFns[Name()] = preemptableFn;                                                           // :93
    ssn->preemptableFns[Name()] = preemptableFn;                                                           // :94
    ssn->jobPipelinedFns[Name()] = [ssn](const JobInfo& ji) { return ssn->JobInfoPipelined(ji); };        // :126-129
  }
};

// ---------------------------------------------------------------------------------------------
// plugins/drf/drf.go:60-171
// ---------------------------------------------------------------------------------------------
struct drfPlugin : Plugin {
  Resource totalResource;
  struct drfAttr { double share = 0; Resource allocated; };
  std::vector<drfAttr> jobOpts;
  std::string Name() const override { return "drf"; }
  double calculateShare(const Session* ssn, const Resource& allocated, const Resource& total) const {  // :161-171
    double res = 0;
    for (uint32_t k = 0; k < ssn->R; ++k) {
      if (k >= 2 && !((total.present >> k) & 1u)) continue;   // total.ResourceNames()
      double a = (k < 2 || ((allocated.present >> k) & 1u)) ? allocated.v[k] : 0.0;  // Resource.Get
      double share = Share(a, total.v[k]);
      if (share > res) res = share;
    }
    return res;
  }
  void OnSessionOpen(Session* ssn) override {
    for (auto& n : ssn->Nodes) ssn->A.Add(totalResource, n.Allocatable);        // :62-64
    jobOpts.resize(ssn->Jobs.size());
    for (auto& job : ssn->Jobs) {                                                // :66-83
      drfAttr& attr = jobOpts[job.idx];
      attr.allocated = job.Allocated;   // = sum Resreq over AllocatedStatus tasks at open (snapshot's job_alloc0)
      attr.share = calculateShare(ssn, attr.allocated, totalResource);
    }
    ssn->jobOrderFns[Name()] = [this](const JobInfo& lv, const JobInfo& rv) {   // :114-130
      if (jobOpts[lv.idx].share == jobOpts[rv.idx].share) return 0;
      if (jobOpts[lv.idx].share < jobOpts[rv.idx].share) return -1;
      return 1;
    };
    ssn->allocateHandlers.push_back([this, ssn](const TaskInfo& task) {          // :136-144
      drfAttr& attr = jobOpts[task.Job];
      ssn->A.Add(attr.allocated, task.Resreq);
      attr.share = calculateShare(ssn, attr.allocated, totalResource);
    });
    ssn->deallocateHandlers.push_back([this, ssn](const TaskInfo& task) {        // :145-153
      drfAttr& attr = jobOpts[task.Job];
      if (!ssn->A.Sub(attr.allocated, task.Resreq)) g_err = "panic: Resource is not sufficient to do operation (drf DeallocateFunc)";
      attr.share = calculateShare(ssn, attr.allocated, totalResource);
    });
    ssn->preemptableFns[Name()] = [this, ssn](const TaskInfo& preemptor, const std::vector<uint32_t>& preemptees) {   // :84-110
      std::vector<uint32_t> victims;
      Resource lalloc = jobOpts[preemptor.Job].allocated;
      ssn->A.Add(lalloc, preemptor.Resreq);
      const double ls = calculateShare(ssn, lalloc, totalResource);
      std::map<uint32_t, Resource> allocations;
      for (uint32_t id : preemptees) {
        const TaskInfo& preemptee = ssn->Tasks[id];
        if (!allocations.count(preemptee.Job)) allocations[preemptee.Job] = jobOpts[preemptee.Job].allocated;
        Resource& ralloc = allocations[preemptee.Job];
        if (!ssn->A.Sub(ralloc, preemptee.Resreq)) { g_err = "panic: Resource is not sufficient to do operation (drf preemptableFn)"; break; }
        const double rs = calculateShare(ssn, ralloc, totalResource);
        if (ls < rs || std::fabs(ls - rs) <= 0.000001) victims.push_back(id);     // shareDelta, drf.go:31
      }
      return victims;
    };
  }
};

// ---------------------------------------------------------------------------------------------
// plugins/proportion/proportion.go:58-253
// ---------------------------------------------------------------------------------------------
struct proportionPlugin : Plugin {
  Resource totalResource;
  struct queueAttr { uint32_t queueID; int32_t weight = 0; double share = 0; Resource deserved, allocated, request; bool used = false; };
  std::vector<queueAttr> queueOpts;   // index = queue idx; `used` <=> key exists in the Go map
  std::string Name() const override { return "proportion"; }
  void updateShare(const Session* ssn, queueAttr& attr) const {                  // :241-253
    double res = 0;
    for (uint32_t k = 0; k < ssn->R; ++k) {
      if (k >= 2 && !((attr.deserved.present >> k) & 1u)) continue;              // deserved.ResourceNames()
      double a = (k < 2 || ((attr.allocated.present >> k) & 1u)) ? attr.allocated.v[k] : 0.0;
      double share = Share(a, attr.deserved.v[k]);
      if (share > res) res = share;
    }
    attr.share = res;
  }
  void OnSessionOpen(Session* ssn) override {
    const Algebra& A = ssn->A;
    for (auto& n : ssn->Nodes) A.Add(totalResource, n.Allocatable);              // :60-62
    queueOpts.resize(ssn->Queues.size());
    for (auto& job : ssn->Jobs) {                                                // :67-98
      queueAttr& attr = queueOpts[job.Queue];
      if (!attr.used) { attr.used = true; attr.queueID = job.Queue; attr.weight = ssn->Queues[job.Queue].Weight; }
      // AllocatedStatus tasks at open: only their sum survives flattening (job_alloc0)
      A.Add(attr.allocated, job.Allocated);
      A.Add(attr.request, job.Allocated);
      auto it = job.TaskStatusIndex.find(Pending);
      if (it != job.TaskStatusIndex.end())
        for (uint32_t id : it->second) A.Add(attr.request, ssn->Tasks[id].Resreq);
    }
    Resource remaining = totalResource;                                          // :100
    std::set<uint32_t> meet;
    for (;;) {
      int32_t totalWeight = 0;
      for (auto& attr : queueOpts) { if (!attr.used || meet.count(attr.queueID)) continue; totalWeight += attr.weight; }
      if (totalWeight == 0) break;
      Resource increasedDeserved, decreasedDeserved;
      for (auto& attr : queueOpts) {
        if (!attr.used || meet.count(attr.queueID)) continue;
        Resource oldDeserved = attr.deserved;
        Resource part = remaining;
        A.Multi(part, (double)attr.weight / (double)totalWeight);
        A.Add(attr.deserved, part);
        if (A.Less(attr.request, attr.deserved)) {
          attr.deserved = A.Min(attr.deserved, attr.request);
          meet.insert(attr.queueID);
        }
        updateShare(ssn, attr);
        Resource increased, decreased;
        A.Diff(attr.deserved, oldDeserved, increased, decreased);
        A.Add(increasedDeserved, increased);
        A.Add(decreasedDeserved, decreased);
      }
      if (!A.Sub(remaining, increasedDeserved)) { g_err = "panic: proportion remaining.Sub"; break; }
      A.Add(remaining, decreasedDeserved);
      if (A.IsEmpty(remaining)) break;
    }
    ssn->queueOrderFns[Name()] = [this](const QueueInfo& lv, const QueueInfo& rv) {   // :156-169
      if (queueOpts[lv.idx].share == queueOpts[rv.idx].share) return 0;
      if (queueOpts[lv.idx].share < queueOpts[rv.idx].share) return -1;
      return 1;
    };
    ssn->overusedFns[Name()] = [this, ssn](const QueueInfo& queue) {              // :198-209
      queueAttr& attr = queueOpts[queue.idx];
      return ssn->A.LessEqual(attr.deserved, attr.allocated);
    };
    ssn->allocateHandlers.push_back([this, ssn](const TaskInfo& task) {           // :213-222
      queueAttr& attr = queueOpts[ssn->Jobs[task.Job].Queue];
      ssn->A.Add(attr.allocated, task.Resreq);
      updateShare(ssn, attr);
    });
    ssn->deallocateHandlers.push_back([this, ssn](const TaskInfo& task) {         // :223-232
      queueAttr& attr = queueOpts[ssn->Jobs[task.Job].Queue];
      if (!ssn->A.Sub(attr.allocated, task.Resreq)) g_err = "panic: Resource is not sufficient to do operation (proportion DeallocateFunc)";
      updateShare(ssn, attr);
    });
    ssn->reclaimableFns[Name()] = [this, ssn](const TaskInfo&, const std::vector<uint32_t>& reclaimees) {   // :171-196
      std::vector<uint32_t> victims;
      std::map<uint32_t, Resource> allocations;
      for (uint32_t id : reclaimees) {
        const TaskInfo& reclaimee = ssn->Tasks[id];
        const uint32_t q = ssn->Jobs[reclaimee.Job].Queue;
        queueAttr& attr = queueOpts[q];
        if (!allocations.count(q)) allocations[q] = attr.allocated;
        Resource& allocated = allocations[q];
        if (ssn->A.Less(allocated, reclaimee.Resreq)) continue;                     // "not enough resource"
        if (!ssn->A.Sub(allocated, reclaimee.Resreq)) { g_err = "panic: Resource is not sufficient to do operation (proportion reclaimableFn)"; break; }
        if (ssn->A.LessEqual(attr.deserved, allocated)) victims.push_back(id);
      }
      return victims;
    };
  }
};

// ---------------------------------------------------------------------------------------------
// plugins/predicates/predicates.go:112-266 + vendored predicates
// ---------------------------------------------------------------------------------------------
struct predicatesPlugin : Plugin {
  std::map<std::string, std::string> args;
  std::string Name() const override { return "predicates"; }
  void OnSessionOpen(Session* ssn) override {
    bool memoryPressureEnable = false, diskPressureEnable = false, pidPressureEnable = false;   // :88-104
    GetBool(args, &memoryPressureEnable, "predicate.MemoryPressureEnable");
    GetBool(args, &diskPressureEnable, "predicate.DiskPressureEnable");
    GetBool(args, &pidPressureEnable, "predicate.PIDPressureEnable");
    const uint32_t W = ssn->W;
    ssn->predicateFns[Name()] = [=](const TaskInfo& task, const NodeInfo& node, int32_t pods, const uint64_t* ports) {
      // :127  node.Allocatable.MaxTaskNum <= len(nodeInfo.Pods())
      if (node.max_pods <= pods) return false;
      // InterPodAffinityMatches (:1261-1288, the plugin's last step) returns an error once a listed pod has no node: see Session::Allocate
      if (ssn->n_allocated_nowhere) return false;
      // CheckNodeConditionPredicate, vendor/.../predicates.go:1675-1698 (incl. Spec.Unschedulable)
      if (node.flags & (KB_NODE_NOT_READY | KB_NODE_NET_UNAVAILABLE | KB_NODE_UNSCHEDULABLE)) return false;
      // CheckNodeUnschedulablePredicate :1576-1593 — subsumed: an unschedulable node already failed above
      // PodMatchNodeSelector :973 -> PodMatchesNodeSelectorAndAffinityTerms :927-970
      for (uint32_t w = 0; w < W; ++w)
        if ((node.labels[w] & task.sel_req[w]) != task.sel_req[w]) return false;
      if (task.n_aff > 0) {   // required node affinity: OR over terms of AND over requirements
        bool any = false;
        for (uint32_t t = 0; t < task.n_aff && !any; ++t) {
          bool all = true;
          for (uint32_t w = 0; w < W; ++w)
            if ((node.labels[w] & task.aff[t][w]) != task.aff[t][w]) { all = false; break; }
          any = all;
        }
        if (!any) return false;
      }
      // PodFitsHostPorts :1153-1173, nodeinfo/host_ports.go:96-125
      for (uint32_t w = 0; w < W; ++w)
        if (ports[w] & task.port_conflict[w]) return false;
      // PodToleratesNodeTaints :1596-1624 (NoSchedule/NoExecute taints only)
      for (uint32_t w = 0; w < W; ++w)
        if (node.taints[w] & ~task.tol[w]) return false;
      if (memoryPressureEnable)   // :1633-1650 — only BestEffort pods are refused
        if ((task.flags & KB_TASK_BEST_EFFORT_QOS) && (node.flags & KB_NODE_MEM_PRESSURE)) return false;
      if (diskPressureEnable && (node.flags & KB_NODE_DISK_PRESSURE)) return false;    // :1654-1660
      if (pidPressureEnable && (node.flags & KB_NODE_PID_PRESSURE)) return false;      // :1664-1671
      // InterPodAffinityMatches :1261-1288 — identically true when no pod of the session carries (anti)affinity terms
      // (no raw pod objects were handed over: kbo_set_pod_objects)
      if (ssn->pw) return ssn->InterPodAffinityMatches(task, node);
      return true;
    };
  }
};

// ---------------------------------------------------------------------------------------------
// plugins/nodeorder/nodeorder.go:107-169
// ---------------------------------------------------------------------------------------------
struct nodeOrderPlugin : Plugin {
  std::map<std::string, std::string> args;
  std::string Name() const override { return "nodeorder"; }
  void OnSessionOpen(Session* ssn) override {
    int leastReqWeight = 1, mostReqWeight = 0, nodeAffinityWeight = 1, podAffinityWeight = 1, balancedResourceWeight = 1;
    GetInt(args, &leastReqWeight, "leastrequested.weight");
    GetInt(args, &mostReqWeight, "mostrequested.weight");
    GetInt(args, &nodeAffinityWeight, "nodeaffinity.weight");
    GetInt(args, &podAffinityWeight, "podaffinity.weight");
    GetInt(args, &balancedResourceWeight, "balancedresource.weight");
    // req_r = nodeInfo.NonZeroRequest().r + calculatePodResourceRequest(pod, r)  (resource_allocation.go:100-142)
    std::vector<PriorityConfig> cfgs;
    cfgs.push_back({"LeastRequestedPriority",
                    [](const TaskInfo& t, const NodeInfo& n, int64_t nzc, int64_t nzm) {
                      return leastResourceScorer(nzc + t.nz_cpu, n.alloc_cpu, nzm + t.nz_mem, n.alloc_mem); },
                    leastReqWeight, nullptr});
    cfgs.push_back({"MostRequestedPriority",
                    [](const TaskInfo& t, const NodeInfo& n, int64_t nzc, int64_t nzm) {
                      return mostResourceScorer(nzc + t.nz_cpu, n.alloc_cpu, nzm + t.nz_mem, n.alloc_mem); },
                    mostReqWeight, nullptr});
    // NodeAffinityPriority: CalculateNodeAffinityPriorityMap (node_affinity.go:34-77): count = sum of the weights of the
    // preferred terms whose selector matches the node's labels (weight 0 terms are skipped, an empty term matches every
    // node); CalculateNodeAffinityPriorityReduce = NormalizeReduce(10, false) over the feasible nodes.
    cfgs.push_back({"NodeAffinityPriority",
                    [](const TaskInfo& t, const NodeInfo& n, int64_t, int64_t) {
                      int64_t count = 0;
                      for (uint32_t p = 0; p < t.n_pref; ++p) {
                        if (t.pref_w[p] == 0) continue;
                        bool match = true;
                        for (uint32_t w = 0; w < KB_MAX_W; ++w) if ((n.labels[w] & t.pref[p][w]) != t.pref[p][w]) { match = false; break; }
                        if (match) count += t.pref_w[p];
                      }
                      return count; },
                    nodeAffinityWeight, NormalizeReduce(10, false)});
    // InterPodAffinityPriority (interpod_affinity.go:99-235): no pod (anti)affinity terms in the snapshot (tasks carrying
    // them are refused) -> every count is 0 -> score 0.
    cfgs.push_back({"InterPodAffinityPriority", [](const TaskInfo&, const NodeInfo&, int64_t, int64_t) { return (int64_t)0; }, podAffinityWeight, nullptr});
    cfgs.push_back({"BalancedResourceAllocation",
                    [](const TaskInfo& t, const NodeInfo& n, int64_t nzc, int64_t nzm) {
                      return balancedResourceScorer(nzc + t.nz_cpu, n.alloc_cpu, nzm + t.nz_mem, n.alloc_mem); },
                    balancedResourceWeight, nullptr});
    ssn->nodePrioritizers[Name()] = cfgs;
  }
};

struct conformancePlugin : Plugin {   // plugins/conformance/conformance.go:41-63 — evictable filter only
  std::string Name() const override { return "conformance"; }
  void OnSessionOpen(Session* ssn) override {
    auto evictableFn = [ssn](const TaskInfo&, const std::vector<uint32_t>& evictees) {
      std::vector<uint32_t> victims;
      for (uint32_t id : evictees) if (!ssn->Tasks[id].critical) victims.push_back(id);   // system-critical class / kube-system
      return victims;
    };
    ssn->preemptableFns[Name()] = evictableFn;
    ssn->reclaimableFns[Name()] = evictableFn;
  }
};

// plugins/factory.go:31-42
std::unique_ptr<Plugin> GetPluginBuilder(const PluginOption& opt) {
  if (opt.Name == "priority") return std::make_unique<priorityPlugin>();
  if (opt.Name == "gang") return std::make_unique<gangPlugin>();
  if (opt.Name == "drf") return std::make_unique<drfPlugin>();
  if (opt.Name == "proportion") return std::make_unique<proportionPlugin>();
  if (opt.Name == "predicates") { auto p = std::make_unique<predicatesPlugin>(); p->args = opt.Arguments; return p; }
  if (opt.Name == "nodeorder") { auto p = std::make_unique<nodeOrderPlugin>(); p->args = opt.Arguments; return p; }
  if (opt.Name == "conformance") return std::make_unique<conformancePlugin>();
  return nullptr;
}

// ---------------------------------------------------------------------------------------------
// snapshot -> Session   (what cache.Snapshot() + framework.OpenSession hand to the action)
// ---------------------------------------------------------------------------------------------
int check_snapshot(const kb_snapshot* s) {
  if (!s) { g_err = "snapshot is NULL"; return KB_E_BADARG; }
  if (s->abi_version != KB_ABI_VERSION) { g_err = "abi_version mismatch"; return KB_E_BADARG; }
  if (s->R < 2 || s->R > KB_MAX_R || s->W < 1 || s->W > KB_MAX_W) { g_err = "R or W out of range"; return KB_E_BADARG; }
  if (s->Q > KB_MAX_Q) { g_err = "too many queues"; return KB_E_BADARG; }
  return KB_OK;
}

int build_session(const kb_snapshot* s, const kb_plugin_conf* conf, int mode, std::unique_ptr<Session>& out, const kbo_running* run = nullptr) {
  int rc = check_snapshot(s);
  if (rc) return rc;
  auto ssn = std::make_unique<Session>(s->R, s->W);
  ssn->mode = mode;
  const uint32_t R = s->R, W = s->W, N = s->N, T = s->T, J = s->J, Q = s->Q;

  ssn->Nodes.resize(N);
  for (uint32_t n = 0; n < N; ++n) {
    NodeInfo& ni = ssn->Nodes[n];
    ni.idx = n;
    for (uint32_t r = 0; r < R; ++r) {
      ni.Idle.v[r] = s->node_idle[(size_t)r * N + n];
      ni.Releasing.v[r] = s->node_releasing[(size_t)r * N + n];
      ni.Used.v[r] = s->node_used[(size_t)r * N + n];
      ni.Allocatable.v[r] = s->node_allocatable[(size_t)r * N + n];
    }
    uint32_t p = s->node_alloc_present[n] & ~3u;
    // Idle starts as NewResource(Allocatable) and Used/Releasing accumulate task Resreq; for the
    // decisions only Allocatable's key set is observable (drf / proportion totals) — see DESIGN.md.
    ni.Idle.present = p; ni.Releasing.present = p; ni.Used.present = p; ni.Allocatable.present = p;
    for (uint32_t r = 2; r < R; ++r) if (!((p >> r) & 1u)) {
      // a scalar the node does not list can still be non-zero in Used/Releasing via pod requests
      if (ni.Idle.v[r] != 0) ni.Idle.present |= 1u << r;
      if (ni.Releasing.v[r] != 0) ni.Releasing.present |= 1u << r;
      if (ni.Used.v[r] != 0) ni.Used.present |= 1u << r;
    }
    ni.Allocatable.MaxTaskNum = s->node_max_pods[n];
    ni.alloc_cpu = s->node_alloc_cpu[n]; ni.alloc_mem = s->node_alloc_mem[n];
    ni.max_pods = s->node_max_pods[n];
    ni.flags = s->node_flags[n];
    ni.pods = s->node_pods[n];
    ni.nz_cpu = s->node_nz_cpu[n]; ni.nz_mem = s->node_nz_mem[n];
    for (uint32_t w = 0; w < KB_MAX_W; ++w) { ni.labels[w] = ni.taints[w] = ni.ports[w] = 0; }
    for (uint32_t w = 0; w < W; ++w) {
      ni.labels[w] = s->node_labels[(size_t)w * N + n];
      ni.taints[w] = s->node_taints[(size_t)w * N + n];
      ni.ports[w] = s->node_ports[(size_t)w * N + n];
    }
    if (mode == KBO_MODE_FAITHFUL) {
      // synthesize the node's pre-existing pod list: node_pods stubs whose aggregate equals the snapshot's
      int32_t k = ni.pods;
      ni.existing.resize((size_t)std::max(k, 0));
      for (int32_t i = 0; i < k; ++i) {
        PodStub& ps = ni.existing[i];
        std::memset(&ps, 0, sizeof(ps));
        ps.nz_cpu = ni.nz_cpu / k + (i == 0 ? ni.nz_cpu % k : 0);
        ps.nz_mem = ni.nz_mem / k + (i == 0 ? ni.nz_mem % k : 0);
        if (i == 0) for (uint32_t w = 0; w < W; ++w) ps.ports[w] = ni.ports[w];
      }
    }
  }

  ssn->Queues.resize(Q);
  for (uint32_t q = 0; q < Q; ++q) { ssn->Queues[q].idx = q; ssn->Queues[q].Weight = s->queue_weight[q]; ssn->Queues[q].ctime = s->queue_ctime[q]; }

  ssn->Tasks.resize(T);
  ssn->Jobs.resize(J);
  for (uint32_t j = 0; j < J; ++j) {
    JobInfo& ji = ssn->Jobs[j];
    ji.idx = j;
    ji.Queue = s->job_queue[j];
    if (ji.Queue >= Q) { g_err = "job_queue out of range (cache.Snapshot drops such jobs, cache.go:652-656)"; return KB_E_BADARG; }
    ji.Priority = s->job_prio[j];
    ji.MinAvailable = s->job_min_avail[j];
    ji.ctime = s->job_ctime[j];
    ji.ready0 = s->job_ready0[j];
    for (uint32_t r = 0; r < R; ++r) ji.Allocated.v[r] = s->job_alloc0[(size_t)r * J + j];
    ji.Allocated.present = s->job_alloc0_present[j] & ~3u;
    if (s->job_task_off[j] > s->job_task_off[j + 1] || s->job_task_off[j + 1] > T) { g_err = "job_task_off not monotone"; return KB_E_BADARG; }
    for (uint32_t t = s->job_task_off[j]; t < s->job_task_off[j + 1]; ++t) {
      TaskInfo& ti = ssn->Tasks[t];
      ti.idx = t; ti.Job = j;
      for (uint32_t r = 0; r < R; ++r) {
        ti.Resreq.v[r] = s->task_resreq[(size_t)r * T + t];
        ti.InitResreq.v[r] = s->task_initreq[(size_t)r * T + t];
        if (ti.Resreq.v[r] > ti.InitResreq.v[r]) { g_err = "task_resreq > task_initreq (violates api/pod_info.go:53-73)"; return KB_E_BADARG; }
      }
      ti.Resreq.present = s->task_res_present[t] & ~3u;
      ti.InitResreq.present = ti.Resreq.present;
      for (uint32_t r = 2; r < R; ++r) if (ti.InitResreq.v[r] != 0) ti.InitResreq.present |= 1u << r;
      ti.Priority = s->task_prio[t]; ti.ctime = s->task_ctime[t]; ti.uid_rank = s->task_uid_rank[t];
      ti.nz_cpu = s->task_nz_cpu[t]; ti.nz_mem = s->task_nz_mem[t];
      ti.flags = s->task_flags[t];
      if ((ti.flags & KB_TASK_HAS_POD_AFFINITY) && !g_pod_world) { g_err = "inter-pod affinity terms without the raw pod objects (kbo_set_pod_objects)"; return KB_E_UNSUPPORTED_FEATURE; }
      ti.n_pref = 0;
      for (auto& pw : ti.pref_w) pw = 0;
      for (auto& pr : ti.pref) for (uint32_t w = 0; w < KB_MAX_W; ++w) pr[w] = 0;
      if (ti.flags & KB_TASK_HAS_PREFERRED_NODE_AFFINITY) {
        if (!s->task_n_pref_terms || !s->task_pref_terms || !s->task_pref_weights) { g_err = "KB_TASK_HAS_PREFERRED_NODE_AFFINITY without task_pref_* arrays"; return KB_E_BADARG; }
        ti.n_pref = s->task_n_pref_terms[t];
        if (ti.n_pref > KB_MAX_PREF_TERMS) { g_err = "task_n_pref_terms > KB_MAX_PREF_TERMS"; return KB_E_BADARG; }
        for (uint32_t p = 0; p < ti.n_pref; ++p) {
          ti.pref_w[p] = s->task_pref_weights[(size_t)p * T + t];
          for (uint32_t w = 0; w < W; ++w) ti.pref[p][w] = s->task_pref_terms[((size_t)p * W + w) * T + t];
        }
      }
      ti.n_aff = s->task_n_aff_terms[t];
      if (ti.n_aff > KB_MAX_AFF_TERMS) { g_err = "task_n_aff_terms > KB_MAX_AFF_TERMS"; return KB_E_BADARG; }
      for (uint32_t w = 0; w < KB_MAX_W; ++w) { ti.sel_req[w] = ti.tol[w] = ti.port_own[w] = ti.port_conflict[w] = 0; for (auto& a : ti.aff) a[w] = 0; }
      for (uint32_t w = 0; w < W; ++w) {
        ti.sel_req[w] = s->task_sel_req[(size_t)w * T + t];
        ti.tol[w] = s->task_tol[(size_t)w * T + t];
        ti.port_own[w] = s->task_port_own[(size_t)w * T + t];
        ti.port_conflict[w] = s->task_port_conflict[(size_t)w * T + t];
        for (uint32_t a = 0; a < KB_MAX_AFF_TERMS; ++a) ti.aff[a][w] = s->task_aff_terms[((size_t)a * W + w) * T + t];
      }
      ti.Status = Pending;
      ji.TaskStatusIndex[Pending].insert(t);
    }
  }

  // Running tasks, one by one (reclaim / preempt walk node.Tasks): the node and job aggregates of the snapshot already count
  // them, so they only join job.TaskStatusIndex[Running] (moving out of the ready0 aggregate) and node.Tasks.
  if (run && run->n) {
    ssn->Tasks.resize((size_t)T + run->n);
    for (uint32_t i = 0; i < run->n; ++i) {
      TaskInfo& ti = ssn->Tasks[(size_t)T + i];
      ti.idx = T + i;
      if (run->job[i] >= J || run->node[i] >= N) { g_err = "kbo_running: job / node index out of range"; return KB_E_BADARG; }
      ti.Job = run->job[i];
      for (uint32_t r = 0; r < R; ++r) ti.Resreq.v[r] = ti.InitResreq.v[r] = run->resreq[(size_t)r * run->n + i];
      ti.Resreq.present = ti.InitResreq.present = run->res_present[i] & ~3u;
      ti.Priority = run->prio[i]; ti.ctime = run->ctime[i]; ti.uid_rank = run->uid_rank[i];
      ti.critical = (run->flags[i] & 1u) != 0;
      ti.n_aff = ti.n_pref = 0;
      for (uint32_t w = 0; w < KB_MAX_W; ++w) { ti.sel_req[w] = ti.tol[w] = ti.port_own[w] = ti.port_conflict[w] = 0; for (auto& a : ti.aff) a[w] = 0; for (auto& a : ti.pref) a[w] = 0; }
      ti.Status = Running;
      ti.NodeName = (int)run->node[i];
      JobInfo& ji = ssn->Jobs[ti.Job];
      if (ji.ready0 <= 0) { g_err = "kbo_running: job_ready0 does not cover the job's running tasks"; return KB_E_BADARG; }
      ji.ready0 -= 1;
      ji.TaskStatusIndex[Running].insert(ti.idx);
      ssn->Nodes[run->node[i]].Tasks[ti.idx] = ti;
    }
  }

  // framework.OpenSession (framework/framework.go:30-52)
  if (conf)
    for (uint32_t ti = 0; ti < conf->n_tiers; ++ti) {
      Tier tier;
      for (uint32_t pi = 0; pi < conf->tiers[ti].n_plugins; ++pi) {
        const kb_plugin_option& o = conf->tiers[ti].plugins[pi];
        PluginOption po;
        po.Name = o.name ? o.name : "";
        po.EnabledJobOrder = o.enabled_job_order; po.EnabledJobReady = o.enabled_job_ready;
        po.EnabledJobPipelined = o.enabled_job_pipelined; po.EnabledTaskOrder = o.enabled_task_order;
        po.EnabledPreemptable = o.enabled_preemptable; po.EnabledReclaimable = o.enabled_reclaimable;
        po.EnabledQueueOrder = o.enabled_queue_order; po.EnabledPredicate = o.enabled_predicate;
        po.EnabledNodeOrder = o.enabled_node_order;
        for (uint32_t a = 0; a < o.n_args; ++a) po.Arguments[o.arg_keys[a]] = o.arg_values[a];
        tier.Plugins.push_back(po);
      }
      ssn->Tiers.push_back(tier);
    }
  for (auto& tier : ssn->Tiers)
    for (auto& po : tier.Plugins) {
      auto pb = GetPluginBuilder(po);
      if (!pb) { g_err = "Failed to get plugin " + po.Name; return KB_E_UNSUPPORTED_PLUGIN; }
      ssn->plugins[pb->Name()] = std::move(pb);   // later option of the same name overrides, like the Go map
    }
  if (s->flags & KB_SNAPSHOT_LISTED_POD_WITHOUT_NODE) ssn->n_allocated_nowhere = 1;   // PodLister -> GetNodeInfo error for every pair (see Session::Allocate)
  if (g_pod_world) {
    if (g_pod_world->T != T || g_pod_world->N != N) { g_err = "kbo_set_pod_objects: pod objects do not belong to this snapshot"; return KB_E_BADARG; }
    ssn->pw = g_pod_world;
  }
  for (auto& kv : ssn->plugins) kv.second->OnSessionOpen(ssn.get());
  ssn->resolvePredicates();
  out = std::move(ssn);
  return KB_OK;
}

// ---------------------------------------------------------------------------------------------
// the per-pair work, in the two cost modes
// ---------------------------------------------------------------------------------------------
struct PairAgg { int32_t pods; int64_t nz_cpu, nz_mem; uint64_t ports[KB_MAX_W]; };

// mode A: schedulernodeinfo.NewNodeInfo(node.Pods()...) (vendor/.../nodeinfo/node_info.go:268-282, 502-524)
PairAgg rebuild_node_aggregate(const Session& ssn, const NodeInfo& node) {
  PairAgg a; a.pods = 0; a.nz_cpu = 0; a.nz_mem = 0;
  for (uint32_t w = 0; w < KB_MAX_W; ++w) a.ports[w] = 0;
  std::vector<const void*> pods;   // n.pods = append(n.pods, pod)
  for (const PodStub& p : node.existing) {
    a.nz_cpu += p.nz_cpu; a.nz_mem += p.nz_mem;
    for (uint32_t w = 0; w < ssn.W; ++w) a.ports[w] |= p.ports[w];
    pods.push_back(&p);
  }
  for (auto& kv : node.Tasks) {
    a.nz_cpu += kv.second.nz_cpu; a.nz_mem += kv.second.nz_mem;
    for (uint32_t w = 0; w < ssn.W; ++w) a.ports[w] |= kv.second.port_own[w];
    pods.push_back(&kv.second);
  }
  a.pods = (int32_t)pods.size();
  return a;
}

// mode A: satisfiesExistingPodsAntiAffinity slow path (vendor/.../predicates.go:1400-1439) ->
// PodLister.FilteredList over every allocated task of every job (plugins/util/util.go:62-85),
// pod.DeepCopy() for session-placed tasks, then per pod GetNodeInfo + affinity==nil test (:1376-1396).
// Cost model is deliberately CONSERVATIVE (cheaper than the Go original): no DeepCopy of the v1.Pod, an
// int compare instead of the NodeName string compare; only the ssn.Nodes map[string] lookup is kept.
uint64_t scan_all_allocated_pods(const Session& ssn, const NodeInfo& node) {
  uint64_t touched = 0;
  for (const JobInfo& job : ssn.Jobs) {
    for (uint32_t pn : ssn.placeholder_alloc[job.idx]) {            // tasks Bound/Running at session open
      if (pn == node.idx) continue;                                 // nodeInfo.Filter (pod is in nodeInfo)
      auto it = ssn.nodeByName.find(ssn.nodeNames[pn]);             // c.info.GetNodeInfo(existingPod.Spec.NodeName)
      touched += it->second;                                        // affinity == nil -> no topology pairs
    }
    for (auto& kv : job.TaskStatusIndex) {
      if (!AllocatedStatus(kv.first)) continue;
      for (uint32_t id : kv.second) {                               // session-placed tasks
        const TaskInfo& t = ssn.Tasks[id];
        if (t.NodeName == (int)node.idx && !node.Tasks.count(id)) continue;
        auto it = ssn.nodeByName.find(ssn.nodeNames[(uint32_t)t.NodeName]);
        touched += it->second;
      }
    }
  }
  return touched;
}

struct Executor {
  Session& ssn;
  Pool pool;
  std::vector<uint8_t> fitv;
  std::vector<double> scorev;
  std::atomic<uint64_t> sink{0};
  Executor(Session& s, int threads) : ssn(s), pool(std::max(1, threads)), fitv(s.Nodes.size()), scorev(s.Nodes.size()) {}

  // allocate.go:73-87 local predicateFn (resource fit, then ssn.PredicateFn)
  bool predicateFn(const TaskInfo& task, const NodeInfo& node) {
    if (!ssn.A.LessEqual(task.InitResreq, node.Idle) && !ssn.A.LessEqual(task.InitResreq, node.Releasing)) return false;
    if (ssn.mode == KBO_MODE_FAITHFUL && ssn.predicates_plugin_enabled) {
      PairAgg a = rebuild_node_aggregate(ssn, node);
      bool ok = ssn.PredicateFn(task, node, a.pods, a.ports);
      if (ok) sink.fetch_add(scan_all_allocated_pods(ssn, node), std::memory_order_relaxed);
      return ok;
    }
    return ssn.PredicateFn(task, node, node.pods, node.ports);
  }

  // util.PredicateNodes (util/scheduler_helper.go:63-86) with rule (2): order preserved
  std::vector<uint32_t> PredicateNodes(const TaskInfo& task) {
    const int n = (int)ssn.Nodes.size();
    pool.parallelize(n, [&](int i) { fitv[i] = predicateFn(task, ssn.Nodes[i]) ? 1 : 0; });
    std::vector<uint32_t> out;
    for (int i = 0; i < n; ++i) if (fitv[i]) out.push_back((uint32_t)i);
    return out;
  }

  // util.PrioritizeNodes (util/scheduler_helper.go:89-171)
  void PrioritizeNodes(const TaskInfo& task, const std::vector<uint32_t>& nodes, const std::vector<PriorityConfig>& cfgs,
                       std::vector<double>& result) {
    result.assign(nodes.size(), 0.0);
    pool.parallelize((int)nodes.size(), [&](int i) {
      const NodeInfo& node = ssn.Nodes[nodes[i]];
      int64_t nzc = node.nz_cpu, nzm = node.nz_mem;
      if (ssn.mode == KBO_MODE_FAITHFUL) {   // generateNodeMapAndSlice rebuilds NodeInfo for every feasible node (:219-230)
        PairAgg a = rebuild_node_aggregate(ssn, node);
        nzc = a.nz_cpu; nzm = a.nz_mem;
      }
      double score = 0;
      for (auto& c : cfgs) if (!c.Reduce) score += (double)(c.Map(task, node, nzc, nzm) * (int64_t)c.Weight);   // :162-168
      result[i] = score;
    });
    // configs with a Reduce step: Map over the feasible nodes, Reduce the whole list, then the weighted sum (:139-168).
    // Without preferred terms every count is 0 and NormalizeReduce leaves zeros: skip the pass.
    for (auto& c : cfgs) {
      if (!c.Reduce || task.n_pref == 0) continue;
      std::vector<int64_t> col(nodes.size());
      for (size_t i = 0; i < nodes.size(); ++i) col[i] = c.Map(task, ssn.Nodes[nodes[i]], 0, 0);
      c.Reduce(col);
      for (size_t i = 0; i < nodes.size(); ++i) result[i] += (double)(col[i] * (int64_t)c.Weight);
    }
    // the one `Function` config (scheduler_helper.go:112-124): InterPodAffinityPriority over the whole feasible list
    if (ssn.pw)
      for (auto& c : cfgs) {
        if (c.Name != "InterPodAffinityPriority") continue;
        const std::vector<int64_t> col = ssn.InterPodAffinityPriority(task, nodes);
        for (size_t i = 0; i < nodes.size(); ++i) result[i] += (double)(col[i] * (int64_t)c.Weight);
      }
  }

  // util.SelectBestNode (util/scheduler_helper.go:188-208) with rule (3): first max
  static size_t SelectBestNode(const std::vector<double>& priorityList) {
    size_t best = 0;
    double maxScore = priorityList[0];
    for (size_t i = 0; i < priorityList.size(); ++i)
      if (priorityList[i] > maxScore) { maxScore = priorityList[i]; best = i; }
    return best;
  }
};

// actions/allocate/allocate.go:43-194
void Execute(Session& ssn, const kbo_opts& opts, kbo_result& res) {
  auto t0 = std::chrono::steady_clock::now();
  const int64_t warm = opts.warm_tasks > 0 ? opts.warm_tasks : 0;     // timing samples: see kbo_opts.warm_tasks
  const int timed_mode = ssn.mode;
  bool warming = warm > 0;
  if (warming) ssn.mode = KBO_MODE_OPTIMISED;
  Executor ex(ssn, opts.threads);
  PriorityQueue<uint32_t> queues([&](const uint32_t& l, const uint32_t& r) { return ssn.QueueOrderFn(ssn.Queues[l], ssn.Queues[r]); });
  std::map<uint32_t, PriorityQueue<uint32_t>> jobsMap;
  auto jobLess = [&](const uint32_t& l, const uint32_t& r) { return ssn.JobOrderFn(ssn.Jobs[l], ssn.Jobs[r]); };
  auto taskLess = [&](const uint32_t& l, const uint32_t& r) { return ssn.TaskOrderFn(ssn.Tasks[l], ssn.Tasks[r]); };

  for (auto& job : ssn.Jobs) {                         // :50-65 (ascending JobID)
    queues.Push(job.Queue);                            // job.Queue validated at load
    if (!jobsMap.count(job.Queue)) jobsMap.emplace(job.Queue, PriorityQueue<uint32_t>(jobLess));
    jobsMap.at(job.Queue).Push(job.idx);
  }
  std::map<uint32_t, PriorityQueue<uint32_t>> pendingTasks;   // :69
  std::vector<PriorityConfig> cfgs;
  std::vector<double> priorityList;
  bool stop = false;

  for (;;) {                                           // :89
    if (queues.Empty() || stop) break;
    uint32_t q = queues.Pop();
    if (ssn.Overused(ssn.Queues[q])) continue;         // :95-98
    auto jit = jobsMap.find(q);
    if (jit == jobsMap.end() || jit->second.Empty()) continue;   // :104-107
    PriorityQueue<uint32_t>& jobs = jit->second;
    uint32_t j = jobs.Pop();                           // :109
    JobInfo& job = ssn.Jobs[j];
    ++res.visits;
    if (!pendingTasks.count(j)) {                      // :110-125
      PriorityQueue<uint32_t> tasks(taskLess);
      auto pit = job.TaskStatusIndex.find(Pending);
      if (pit != job.TaskStatusIndex.end())
        for (uint32_t id : pit->second) {
          if (ssn.A.IsEmpty(ssn.Tasks[id].Resreq)) continue;   // BestEffort skipped
          tasks.Push(id);
        }
      pendingTasks.emplace(j, std::move(tasks));
    }
    PriorityQueue<uint32_t>& tasks = pendingTasks.at(j);

    while (!tasks.Empty()) {                           // :129
      if (warming && (int64_t)res.tasks_processed >= warm) { warming = false; ssn.mode = timed_mode; t0 = std::chrono::steady_clock::now(); }
      if (!warming && opts.max_tasks > 0 && (int64_t)res.tasks_processed - warm >= opts.max_tasks) { stop = true; res.truncated = 1; break; }
      if (!warming && opts.max_seconds > 0) {
        double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (el >= opts.max_seconds) { stop = true; res.truncated = 1; break; }
      }
      uint32_t tid = tasks.Pop();
      TaskInfo& task = ssn.Tasks[tid];
      ++res.tasks_processed;
      res.pairs_logical += ssn.Nodes.size();
      std::vector<uint32_t> predicateNodes = ex.PredicateNodes(task);          // :143
      if (predicateNodes.empty()) break;                                        // :144-148
      cfgs = ssn.NodePrioritizers();                                            // :150
      ex.PrioritizeNodes(task, predicateNodes, cfgs, priorityList);
      NodeInfo& node = ssn.Nodes[predicateNodes[Executor::SelectBestNode(priorityList)]];   // :156-157
      if (ssn.A.LessEqual(task.InitResreq, node.Idle)) {                        // :160
        ssn.Allocate(task, node);
      } else {
        // :168-170 NodesFitDelta bookkeeping feeds only the unschedulable message (job_info.go FitError)
        if (ssn.A.LessEqual(task.InitResreq, node.Releasing)) {                 // :175
          ssn.Pipeline(task, node);
        }
      }
      if (ssn.JobReady(job) && !tasks.Empty()) {                                // :185-188
        jobs.Push(j);
        break;
      }
    }
    queues.Push(q);                                                             // :192
  }
  res.tasks_allocated = ssn.n_allocated;
  res.tasks_pipelined = ssn.n_pipelined;
  res.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  res.timed_tasks = warming ? 0u : (uint32_t)((int64_t)res.tasks_processed - warm);
  ssn.mode = timed_mode;
}

// actions/backfill/backfill.go:40-71.  Deterministic rules: jobs in ascending JobID, a job's Pending tasks in ascending
// TaskInfo.UID, nodes in ascending Name (all three are Go map iterations in the reference).
void ExecuteBackfill(Session& ssn, kbo_result& res) {
  for (auto& job : ssn.Jobs) {                                              // :45
    std::vector<uint32_t> pend;
    auto pit = job.TaskStatusIndex.find(Pending);
    if (pit != job.TaskStatusIndex.end()) pend.assign(pit->second.begin(), pit->second.end());
    std::sort(pend.begin(), pend.end(), [&](uint32_t a, uint32_t b) { return ssn.Tasks[a].uid_rank < ssn.Tasks[b].uid_rank; });
    for (uint32_t tid : pend) {                                             // :46
      TaskInfo& task = ssn.Tasks[tid];
      if (!ssn.A.IsEmpty(task.InitResreq)) continue;                        // :47 (else branch :66-68 is a TODO)
      ++res.tasks_processed;
      res.pairs_logical += ssn.Nodes.size();
      for (auto& node : ssn.Nodes) {                                        // :50
        if (!ssn.PredicateFn(task, node, node.pods, node.ports)) continue;  // :53-57 — ssn.PredicateFn only, no resource fit
        if (!ssn.Allocate(task, node)) continue;                            // :60-63 (AddTask: Resreq <= Idle, node_info.go:161-167)
        break;                                                              // :64
      }
    }
  }
  res.tasks_allocated = ssn.n_allocated;
  res.tasks_pipelined = ssn.n_pipelined;
}

// Running tasks of a node that `filter` accepts, in ascending TaskInfo.UID (node.Tasks is a Go map in the reference)
template <class F>
std::vector<uint32_t> node_running_tasks(Session& ssn, const NodeInfo& node, F filter) {
  std::vector<uint32_t> ids;
  for (auto& kv : node.Tasks) if (kv.second.Status == Running && filter(kv.second)) ids.push_back(kv.first);   // the NODE's clone decides
  std::sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return ssn.Tasks[a].uid_rank < ssn.Tasks[b].uid_rank; });
  return ids;
}

// actions/reclaim/reclaim.go:40-193.  Deterministic rules: ssn.Jobs in ascending JobID, ssn.Nodes in ascending Name,
// node.Tasks in ascending UID.
void ExecuteReclaim(Session& ssn, kbo_result& res) {
  PriorityQueue<uint32_t> queues([&](const uint32_t& l, const uint32_t& r) { return ssn.QueueOrderFn(ssn.Queues[l], ssn.Queues[r]); });
  std::set<uint32_t> queueMap;
  std::map<uint32_t, PriorityQueue<uint32_t>> preemptorsMap, preemptorTasks;
  auto jobLess = [&](const uint32_t& l, const uint32_t& r) { return ssn.JobOrderFn(ssn.Jobs[l], ssn.Jobs[r]); };
  auto taskLess = [&](const uint32_t& l, const uint32_t& r) { return ssn.TaskOrderFn(ssn.Tasks[l], ssn.Tasks[r]); };
  for (auto& job : ssn.Jobs) {                                            // :53-81
    if (!queueMap.count(job.Queue)) { queueMap.insert(job.Queue); queues.Push(job.Queue); }
    auto pit = job.TaskStatusIndex.find(Pending);
    if (pit != job.TaskStatusIndex.end() && !pit->second.empty()) {
      if (!preemptorsMap.count(job.Queue)) preemptorsMap.emplace(job.Queue, PriorityQueue<uint32_t>(jobLess));
      preemptorsMap.at(job.Queue).Push(job.idx);
      PriorityQueue<uint32_t> tasks(taskLess);
      for (uint32_t id : pit->second) tasks.Push(id);
      preemptorTasks.emplace(job.idx, std::move(tasks));
    }
  }
  for (;;) {                                                              // :83
    if (queues.Empty()) break;
    const uint32_t q = queues.Pop();
    if (ssn.Overused(ssn.Queues[q])) continue;                            // :93-96
    auto jit = preemptorsMap.find(q);
    if (jit == preemptorsMap.end() || jit->second.Empty()) continue;      // :99-103
    const uint32_t j = jit->second.Pop();
    auto tit = preemptorTasks.find(j);
    if (tit == preemptorTasks.end() || tit->second.Empty()) continue;     // :106-110
    TaskInfo& task = ssn.Tasks[tit->second.Pop()];
    const JobInfo& job = ssn.Jobs[j];
    ++res.tasks_processed;
    res.pairs_logical += ssn.Nodes.size();
    bool assigned = false;
    for (auto& n : ssn.Nodes) {                                           // :113
      if (!ssn.PredicateFn(task, n, n.pods, n.ports)) continue;           // :115-117
      const Resource resreq = task.InitResreq;
      Resource reclaimed;
      std::vector<uint32_t> reclaimees = node_running_tasks(ssn, n, [&](const TaskInfo& t) { return ssn.Jobs[t.Job].Queue != job.Queue; });   // :125-138
      std::vector<uint32_t> victims = ssn.Reclaimable(task, reclaimees);
      if (victims.empty()) continue;                                      // :141-144
      Resource allRes;
      for (uint32_t v : victims) ssn.A.Add(allRes, ssn.Tasks[v].Resreq);
      if (!ssn.A.LessEqual(resreq, allRes)) continue;                     // :147-154
      for (uint32_t v : victims) {                                        // :157-170
        TaskInfo& reclaimee = ssn.Tasks[v];
        ssn.EvictTask(reclaimee, true);
        ssn.A.Add(reclaimed, reclaimee.Resreq);
        if (ssn.A.LessEqual(resreq, reclaimed)) break;
      }
      if (ssn.A.LessEqual(task.InitResreq, reclaimed)) {                  // :175-185
        ssn.Pipeline(task, n);
        assigned = true;
        break;
      }
    }
    if (assigned) queues.Push(q);                                         // :188-190
  }
  res.tasks_pipelined = ssn.n_pipelined;
}

// actions/preempt/preempt.go:171-252
bool preempt_one(Session& ssn, Executor& ex, Statement& stmt, TaskInfo& preemptor, const std::function<bool(const TaskInfo&)>& filter,
                 kbo_result& res) {
  bool assigned = false;
  ++res.tasks_processed;
  res.pairs_logical += ssn.Nodes.size();
  std::vector<uint32_t> predicateNodes;                                   // util.PredicateNodes(preemptor, allNodes, ssn.PredicateFn)
  for (auto& n : ssn.Nodes) if (ssn.PredicateFn(preemptor, n, n.pods, n.ports)) predicateNodes.push_back(n.idx);
  std::vector<double> priorityList;
  ex.PrioritizeNodes(preemptor, predicateNodes, ssn.NodePrioritizers(), priorityList);
  // util.SortNodes: sort.Sort(sort.Reverse(priorityList)) — not stable in Go; deterministic rule: score descending, then node order
  std::vector<size_t> order(predicateNodes.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return priorityList[a] > priorityList[b]; });
  for (size_t oi : order) {
    NodeInfo& node = ssn.Nodes[predicateNodes[oi]];
    Resource preempted;
    const Resource resreq = preemptor.InitResreq;
    std::vector<uint32_t> preemptees = node_running_tasks(ssn, node, filter);      // :195-201 (filter sees the node's clone)
    std::vector<uint32_t> victims = ssn.Preemptable(preemptor, preemptees);
    {                                                                                // validateVictims :254-270
      if (victims.empty()) continue;
      Resource allRes;
      for (uint32_t v : victims) ssn.A.Add(allRes, ssn.Tasks[v].Resreq);
      if (!ssn.A.LessEqual(resreq, allRes)) continue;
    }
    PriorityQueue<uint32_t> victimsQueue([&](const uint32_t& l, const uint32_t& r) { return !ssn.TaskOrderFn(ssn.Tasks[l], ssn.Tasks[r]); });   // :210-215
    for (uint32_t v : victims) victimsQueue.Push(v);
    while (!victimsQueue.Empty()) {                                       // :217-231: lowest priority first
      TaskInfo& preemptee = ssn.Tasks[victimsQueue.Pop()];
      stmt.Evict(preemptee);
      ssn.A.Add(preempted, preemptee.Resreq);
      if (ssn.A.LessEqual(resreq, preempted)) break;
    }
    if (ssn.A.LessEqual(preemptor.InitResreq, preempted)) {               // :237-247
      stmt.Pipeline(preemptor, node);
      assigned = true;
      break;
    }
  }
  return assigned;
}

// actions/preempt/preempt.go:43-167.  Deterministic rules as above; `queues` (a Go map) in ascending QueueID.
void ExecutePreempt(Session& ssn, const kbo_opts& opts, kbo_result& res) {
  Executor ex(ssn, opts.threads);
  auto jobLess = [&](const uint32_t& l, const uint32_t& r) { return ssn.JobOrderFn(ssn.Jobs[l], ssn.Jobs[r]); };
  auto taskLess = [&](const uint32_t& l, const uint32_t& r) { return ssn.TaskOrderFn(ssn.Tasks[l], ssn.Tasks[r]); };
  std::map<uint32_t, PriorityQueue<uint32_t>> preemptorsMap, preemptorTasks;
  std::vector<uint32_t> underRequest;
  std::set<uint32_t> queues;
  for (auto& job : ssn.Jobs) {                                            // :54-75
    queues.insert(job.Queue);
    auto pit = job.TaskStatusIndex.find(Pending);
    if (pit != job.TaskStatusIndex.end() && !pit->second.empty()) {
      if (!preemptorsMap.count(job.Queue)) preemptorsMap.emplace(job.Queue, PriorityQueue<uint32_t>(jobLess));
      preemptorsMap.at(job.Queue).Push(job.idx);
      underRequest.push_back(job.idx);
      PriorityQueue<uint32_t> tasks(taskLess);
      for (uint32_t id : pit->second) tasks.Push(id);
      preemptorTasks.emplace(job.idx, std::move(tasks));
    }
  }
  for (uint32_t q : queues) {                                             // :78 Preemption between Jobs within Queue
    for (;;) {
      auto pit = preemptorsMap.find(q);
      if (pit == preemptorsMap.end() || pit->second.Empty()) break;       // :83-86
      const uint32_t pj = pit->second.Pop();
      JobInfo& preemptorJob = ssn.Jobs[pj];
      Statement stmt(ssn);
      bool assigned = false;
      for (;;) {
        PriorityQueue<uint32_t>& tasks = preemptorTasks.at(pj);
        if (tasks.Empty()) break;                                         // :95-99
        TaskInfo& preemptor = ssn.Tasks[tasks.Pop()];
        if (preempt_one(ssn, ex, stmt, preemptor, [&](const TaskInfo& task) {   // :103-116
              if (task.Status != Running) return false;
              return ssn.Jobs[task.Job].Queue == preemptorJob.Queue && preemptor.Job != task.Job;
            }, res)) assigned = true;
        if (ssn.JobPipelined(preemptorJob)) { stmt.Commit(); break; }     // :121-124
      }
      if (!ssn.JobPipelined(preemptorJob)) { stmt.Discard(); continue; }  // :128-131
      if (assigned) pit->second.Push(pj);                                 // :133-135
    }
    for (uint32_t uj : underRequest) {                                    // :139 Preemption between Task within Job
      for (;;) {
        auto tit = preemptorTasks.find(uj);
        if (tit == preemptorTasks.end() || tit->second.Empty()) break;
        TaskInfo& preemptor = ssn.Tasks[tit->second.Pop()];
        Statement stmt(ssn);
        const bool assigned = preempt_one(ssn, ex, stmt, preemptor, [&](const TaskInfo& task) {   // :152-159
          if (task.Status != Running) return false;
          return preemptor.Job == task.Job;
        }, res);
        stmt.Commit();
        if (!assigned) break;                                             // :163-165
      }
    }
  }
  res.tasks_pipelined = ssn.n_pipelined;
}

Resource mkres(uint32_t R, const double* v, uint32_t present) {
  Resource r;
  for (uint32_t k = 0; k < R && k < KB_MAX_R; ++k) r.v[k] = v[k];
  r.present = present & ~3u;
  return r;
}
void putres(uint32_t R, const Resource& r, double* v, uint32_t* present) {
  for (uint32_t k = 0; k < R && k < KB_MAX_R; ++k) v[k] = r.v[k];
  if (present) *present = r.present;
}

}  // namespace

extern "C" {

const char* kbo_last_error(void) { return g_err.c_str(); }

void kbo_set_pod_objects(const kbo_pod_objects* po, uint32_t N) {
  if (!po) { g_pod_world.reset(); return; }
  auto w = std::make_shared<PodWorld>();
  w->T = po->T; w->N = N; w->n_topo = po->n_topo;
  w->pods.resize(po->P);
  for (uint32_t p = 0; p < po->P; ++p) {
    PodObj& o = w->pods[p];
    o.ns = po->pod_ns[p]; o.has_aff = po->has_aff[p] != 0; o.has_anti = po->has_anti[p] != 0;
    for (uint32_t i = po->lab_off[p]; i < po->lab_off[p + 1]; ++i) o.labels.push_back({po->lab_key[i], po->lab_val[i]});
    for (uint32_t t = po->term_off[p]; t < po->term_off[p + 1]; ++t) {
      AffTerm a;
      a.kind = po->term_kind[t]; a.weight = po->term_weight[t]; a.topo = po->term_topo[t]; a.nil = po->term_nil[t] != 0;
      for (uint32_t i = po->term_ns_off[t]; i < po->term_ns_off[t + 1]; ++i) a.ns.push_back(po->term_ns[i]);
      for (uint32_t r = po->term_req_off[t]; r < po->term_req_off[t + 1]; ++r) {
        LabelReq q; q.key = po->req_key[r]; q.op = po->req_op[r];
        for (uint32_t i = po->req_val_off[r]; i < po->req_val_off[r + 1]; ++i) q.vals.push_back(po->req_val[i]);
        a.reqs.push_back(q);
      }
      o.terms.push_back(a);
    }
  }
  w->existing_on.resize(N);
  for (uint32_t p = po->T; p < po->P; ++p) {
    const uint32_t e = p - po->T;
    ExistingPod x{p, po->pod_node[e], po->pod_listed[e] != 0, po->pod_in_tasks[e] != 0, po->pod_unbound[e] != 0};
    if (x.in_tasks && x.node >= 0 && (uint32_t)x.node < N) w->existing_on[(uint32_t)x.node].push_back((uint32_t)w->existing.size());
    w->existing.push_back(x);
  }
  w->node_topo.assign(po->node_topo, po->node_topo + (size_t)std::max(1u, po->n_topo) * std::max(1u, N));
  g_pod_world = w;
}

int kbo_cycle(const kb_snapshot* snap, const kbo_running* running, const kb_plugin_conf* conf, const kbo_opts* opts_in,
              const uint8_t* action_list, uint32_t n_actions,
              kb_decision* out, uint8_t* evicted, uint32_t* evict_order, kbo_result* res_out,
              double* node_idle, double* node_releasing, double* node_used, int32_t* node_pods,
              int64_t* node_nz_cpu, int64_t* node_nz_mem, uint64_t* node_ports,
              double* job_share, int32_t* job_ready, double* queue_share,
              double* queue_deserved, double* queue_allocated) {
  g_err.clear();
  kbo_opts opts{};
  if (opts_in) opts = *opts_in;
  std::unique_ptr<Session> ssn;
  int rc = build_session(snap, conf, opts.mode, ssn, running);
  if (rc) return rc;
  if (opts.mode == KBO_MODE_FAITHFUL) {
    const uint32_t Nn = (uint32_t)ssn->Nodes.size();
    ssn->nodeNames.resize(Nn);
    for (uint32_t n = 0; n < Nn; ++n) {
      char buf[32]; std::snprintf(buf, sizeof buf, "node-%06u", n);
      ssn->nodeNames[n] = buf; ssn->nodeByName[buf] = n;
    }
    ssn->placeholder_alloc.resize(ssn->Jobs.size());
    uint32_t k = 0;
    for (auto& j : ssn->Jobs)
      for (int32_t i = 0; i < j.ready0 && Nn > 0; ++i) ssn->placeholder_alloc[j.idx].push_back((uint32_t)(((uint64_t)(k++) * 2654435761ull) % Nn));
  }
  kbo_result res{};
  bool backfill_ran = false;
  for (uint32_t i = 0; i < n_actions; ++i) {
    switch (action_list[i]) {
      case KBO_ACT_RECLAIM: ExecuteReclaim(*ssn, res); break;
      case KBO_ACT_ALLOCATE: Execute(*ssn, opts, res); break;
      case KBO_ACT_BACKFILL: ExecuteBackfill(*ssn, res); backfill_ran = true; break;
      case KBO_ACT_PREEMPT: ExecutePreempt(*ssn, opts, res); break;
      default: g_err = "unknown action"; return KB_E_BADARG;
    }
  }
  res.tasks_allocated = ssn->n_allocated;
  res.tasks_pipelined = ssn->n_pipelined;
  res.evictions = ssn->n_evicted;
  if (running) for (uint32_t i = 0; i < running->n; ++i) {
    const TaskInfo& ti = ssn->Tasks[(size_t)snap->T + i];
    if (evicted) evicted[i] = ti.evicted ? 1 : 0;
    if (evict_order) evict_order[i] = ti.evict_order;
  }
  // bench metric (BASELINE.json): PodGroups that received a placement in this cycle and are JobReady at its end
  res.jobs_ready = 0;
  for (auto& job : ssn->Jobs) {
    bool placed = false;
    for (uint32_t t = snap->job_task_off[job.idx]; t < snap->job_task_off[job.idx + 1] && !placed; ++t)
      placed = ssn->Tasks[t].step != 0xFFFFFFFFu;
    if (placed && ssn->JobReady(job)) ++res.jobs_ready;
  }
  const uint32_t R = snap->R, W = snap->W, N = snap->N, T = snap->T, J = snap->J, Q = snap->Q;
  if (out)
    for (uint32_t t = 0; t < T; ++t) {
      const TaskInfo& ti = ssn->Tasks[t];
      kb_decision d{};
      d.node = (ti.Status == Pending) ? -1 : ti.NodeName;     // an un-pipelined preemptor keeps a stale NodeName (statement.go:153-188)
      d.step = ti.step;
      d.dispatch_step = ti.dispatch_step;
      d.dispatched = ti.dispatched ? 1 : 0;
      const bool bf = backfill_ran;
      if (ti.Status == Pending) d.kind = ssn->A.IsEmpty(ti.Resreq) ? ((bf && ssn->A.IsEmpty(ti.InitResreq)) ? KB_KIND_NONE : KB_KIND_SKIPPED) : KB_KIND_NONE;
      else if (ti.Status == Pipelined) d.kind = KB_KIND_PIPELINED;
      else d.kind = KB_KIND_ALLOCATED;
      out[t] = d;
    }
  for (uint32_t n = 0; n < N; ++n) {
    const NodeInfo& ni = ssn->Nodes[n];
    for (uint32_t r = 0; r < R; ++r) {
      if (node_idle) node_idle[(size_t)r * N + n] = ni.Idle.v[r];
      if (node_releasing) node_releasing[(size_t)r * N + n] = ni.Releasing.v[r];
      if (node_used) node_used[(size_t)r * N + n] = ni.Used.v[r];
    }
    if (node_pods) node_pods[n] = ni.pods;
    if (node_nz_cpu) node_nz_cpu[n] = ni.nz_cpu;
    if (node_nz_mem) node_nz_mem[n] = ni.nz_mem;
    if (node_ports) for (uint32_t w = 0; w < W; ++w) node_ports[(size_t)w * N + n] = ni.ports[w];
  }
  auto dit = ssn->plugins.find("drf");
  for (uint32_t j = 0; j < J; ++j) {
    if (job_share) job_share[j] = dit != ssn->plugins.end() ? static_cast<drfPlugin*>(dit->second.get())->jobOpts[j].share : 0.0;
    if (job_ready) job_ready[j] = ssn->ReadyTaskNum(ssn->Jobs[j]);
  }
  auto pit = ssn->plugins.find("proportion");
  for (uint32_t q = 0; q < Q; ++q) {
    proportionPlugin* pp = pit != ssn->plugins.end() ? static_cast<proportionPlugin*>(pit->second.get()) : nullptr;
    if (queue_share) queue_share[q] = pp ? pp->queueOpts[q].share : 0.0;
    for (uint32_t r = 0; r < R; ++r) {
      if (queue_deserved) queue_deserved[(size_t)r * Q + q] = pp ? pp->queueOpts[q].deserved.v[r] : 0.0;
      if (queue_allocated) queue_allocated[(size_t)r * Q + q] = pp ? pp->queueOpts[q].allocated.v[r] : 0.0;
    }
  }
  if (res_out) *res_out = res;
  if (!g_err.empty()) return KB_E_STATE;
  return KB_OK;
}

int kbo_allocate(const kb_snapshot* snap, const kb_plugin_conf* conf, const kbo_opts* opts_in,
                 kb_decision* out, kbo_result* res_out,
                 double* node_idle, double* node_releasing, double* node_used, int32_t* node_pods,
                 int64_t* node_nz_cpu, int64_t* node_nz_mem, uint64_t* node_ports,
                 double* job_share, int32_t* job_ready, double* queue_share,
                 double* queue_deserved, double* queue_allocated) {
  const int actions = (opts_in && opts_in->actions) ? opts_in->actions : KBO_ACTION_ALLOCATE;
  uint8_t list[2]; uint32_t n = 0;
  if (actions & KBO_ACTION_ALLOCATE) list[n++] = KBO_ACT_ALLOCATE;
  if (actions & KBO_ACTION_BACKFILL) list[n++] = KBO_ACT_BACKFILL;
  return kbo_cycle(snap, nullptr, conf, opts_in, list, n, out, nullptr, nullptr, res_out, node_idle, node_releasing, node_used, node_pods,
                   node_nz_cpu, node_nz_mem, node_ports, job_share, job_ready, queue_share, queue_deserved, queue_allocated);
}

int kbo_predicate_score(const kb_snapshot* snap, const kb_plugin_conf* conf, uint32_t task,
                        uint8_t* fit, double* score) {
  g_err.clear();
  std::unique_ptr<Session> ssn;
  int rc = build_session(snap, conf, KBO_MODE_OPTIMISED, ssn);
  if (rc) return rc;
  if (task >= snap->T) { g_err = "task out of range"; return KB_E_BADARG; }
  Executor ex(*ssn, 1);
  const TaskInfo& t = ssn->Tasks[task];
  std::vector<uint32_t> nodes = ex.PredicateNodes(t);
  std::vector<double> pl;
  ex.PrioritizeNodes(t, nodes, ssn->NodePrioritizers(), pl);
  for (uint32_t n = 0; n < snap->N; ++n) { if (fit) fit[n] = 0; if (score) score[n] = 0; }
  for (size_t i = 0; i < nodes.size(); ++i) { if (fit) fit[nodes[i]] = 1; if (score) score[nodes[i]] = pl[i]; }
  return KB_OK;
}

int kbo_res_less_equal(uint32_t R, const double* l, uint32_t lp, const double* r, uint32_t rp) {
  return Algebra(R).LessEqual(mkres(R, l, lp), mkres(R, r, rp)) ? 1 : 0;
}
int kbo_res_less(uint32_t R, const double* l, uint32_t lp, const double* r, uint32_t rp) {
  return Algebra(R).Less(mkres(R, l, lp), mkres(R, r, rp)) ? 1 : 0;
}
int kbo_res_is_empty(uint32_t R, const double* l, uint32_t lp) { return Algebra(R).IsEmpty(mkres(R, l, lp)) ? 1 : 0; }
int kbo_res_sub(uint32_t R, double* l, uint32_t* lp, const double* r, uint32_t rp) {
  Resource a = mkres(R, l, *lp);
  if (!Algebra(R).Sub(a, mkres(R, r, rp))) return -1;
  putres(R, a, l, lp);
  return 0;
}
void kbo_res_add(uint32_t R, double* l, uint32_t* lp, const double* r, uint32_t rp) {
  Resource a = mkres(R, l, *lp);
  Algebra(R).Add(a, mkres(R, r, rp));
  putres(R, a, l, lp);
}
void kbo_res_set_max(uint32_t R, double* l, uint32_t* lp, const double* r, uint32_t rp) {
  Resource a = mkres(R, l, *lp);
  Algebra(R).SetMaxResource(a, mkres(R, r, rp));
  putres(R, a, l, lp);
}
void kbo_res_fit_delta(uint32_t R, double* l, uint32_t* lp, const double* r, uint32_t rp) {
  Resource a = mkres(R, l, *lp);
  Algebra(R).FitDelta(a, mkres(R, r, rp));
  putres(R, a, l, lp);
}
int64_t kbo_least_requested(int64_t rc, int64_t ac, int64_t rm, int64_t am) { return leastResourceScorer(rc, ac, rm, am); }
int64_t kbo_most_requested(int64_t rc, int64_t ac, int64_t rm, int64_t am) { return mostResourceScorer(rc, ac, rm, am); }
int64_t kbo_balanced(int64_t rc, int64_t ac, int64_t rm, int64_t am) { return balancedResourceScorer(rc, ac, rm, am); }
void kbo_heap_sort(const int64_t* keys, uint32_t n, int64_t* out) {
  PriorityQueue<int64_t> pq([](const int64_t& a, const int64_t& b) { return a < b; });
  for (uint32_t i = 0; i < n; ++i) pq.Push(keys[i]);
  for (uint32_t i = 0; i < n; ++i) out[i] = pq.Pop();
}
uint32_t kbo_select_best_node(const double* scores, uint32_t n) {
  std::vector<double> v(scores, scores + n);
  return (uint32_t)Executor::SelectBestNode(v);
}
int kbo_arguments_get_int(const char* value, int base) {
  std::map<std::string, std::string> a;
  if (value) a["intkey"] = value;
  int v = base;
  GetInt(a, &v, "intkey");
  return v;
}
double kbo_share(double l, double r) { return Share(l, r); }

}  // extern "C"
