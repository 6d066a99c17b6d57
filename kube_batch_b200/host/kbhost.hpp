// kbhost.hpp — C++17 host-side mirror of the kube-batch interface around the allocate path, sitting ABOVE the
// C ABI of libkbgpu.so (include/kbgpu.h).  The reference is Go and no Go toolchain exists in this image, so this
// header plays the role the Go shim (go/kbgpu) plays inside kube-batch: same names, same argument meaning, same
// error behaviour, so tests can be written like pkg/scheduler/actions/allocate/allocate_test.go.
//
//   kb::api        Resource, TaskInfo, NodeInfo, JobInfo, QueueInfo, TaskStatus      (pkg/scheduler/api)
//   kb::conf       PluginOption, Tier                                                 (pkg/scheduler/conf)
//   kb::cache      SchedulerCache{AddNode,AddPod,AddPodGroup,AddQueue,Snapshot}, Binder (pkg/scheduler/cache)
//   kb::framework  Session, Plugin, Action, Arguments, RegisterPluginBuilder, OpenSession, CloseSession
//   kb::plugins    the built-in plugin builders by name                                (pkg/scheduler/plugins/factory.go)
//   kb::actions::allocate::New()->Execute(ssn)                                         (pkg/scheduler/actions/allocate)
//
// What differs from the reference, by construction: plugin callbacks are arbitrary Go closures there; a GPU engine
// can only honour the BUILT-IN ones.  Session::Add*Fn therefore registers a *descriptor* (which built-in function,
// from which plugin); Execute refuses loudly (std::runtime_error) when a tier names a plugin that registered a
// foreign function.  There is NO CPU fallback: Execute without a CUDA device throws.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kbgpu.h"

namespace kb {

// ------------------------------------------------------------------------------------------------- api
namespace api {

using ResourceList = std::map<std::string, double>;   // v1.ResourceList: "cpu" cores, "memory" bytes, "pods", scalars in units

struct Resource {                                     // api/resource_info.go:28-38
  double MilliCPU = 0, Memory = 0;
  std::map<std::string, double> ScalarResources;      // empty == nil
  int MaxTaskNum = 0;
  static Resource New(const ResourceList& rl) {       // NewResource, resource_info.go:73-90
    Resource r;
    for (auto& kv : rl) {
      if (kv.first == "cpu") r.MilliCPU += std::ceil(kv.second * 1000.0 - 1e-9);      // Quantity.MilliValue rounds up
      else if (kv.first == "memory") r.Memory += kv.second;
      else if (kv.first == "pods") r.MaxTaskNum += (int)kv.second;
      else r.ScalarResources[kv.first] += std::ceil(kv.second * 1000.0 - 1e-9);
    }
    return r;
  }
  Resource& Add(const Resource& rr) { MilliCPU += rr.MilliCPU; Memory += rr.Memory; for (auto& kv : rr.ScalarResources) ScalarResources[kv.first] += kv.second; return *this; }
  Resource& SubUnchecked(const Resource& rr) { MilliCPU -= rr.MilliCPU; Memory -= rr.Memory; for (auto& kv : rr.ScalarResources) ScalarResources[kv.first] -= kv.second; return *this; }
  // Resource.LessEqual (resource_info.go:268-302): epsilons 10 m / 10 Mi / 10; scalars <= 10 are skipped; a nil map on the
  // right makes any larger scalar fail
  bool LessEqual(const Resource& rr) const {
    auto le = [](double l, double r, double diff) { return l < r || std::fabs(l - r) < diff; };
    if (!le(MilliCPU, rr.MilliCPU, 10.0)) return false;
    if (!le(Memory, rr.Memory, 10.0 * 1024 * 1024)) return false;
    for (auto& kv : ScalarResources) {
      if (kv.second <= 10.0) continue;
      if (rr.ScalarResources.empty()) return false;
      auto it = rr.ScalarResources.find(kv.first);
      if (!le(kv.second, it == rr.ScalarResources.end() ? 0.0 : it->second, 10.0)) return false;
    }
    return true;
  }
  void SetMaxResource(const Resource& rr) { MilliCPU = std::max(MilliCPU, rr.MilliCPU); Memory = std::max(Memory, rr.Memory);
    for (auto& kv : rr.ScalarResources) { auto& x = ScalarResources[kv.first]; x = std::max(x, kv.second); } }
};

enum TaskStatus { Pending = 1, Allocated = 2, Pipelined = 4, Binding = 8, Bound = 16, Running = 32, Releasing = 64,
                  Succeeded = 128, Failed = 256, Unknown = 512 };                     // api/types.go:20-54
inline bool AllocatedStatus(int s) { return s == Bound || s == Binding || s == Running || s == Allocated; }   // api/helpers.go:64-71

struct Pod {                                          // the v1.Pod fields the path reads (util/test_utils.go:66-93 BuildPod)
  std::string Namespace, Name, UID, NodeName, Phase = "Pending", GroupName;
  ResourceList Requests;                              // one container
  std::vector<ResourceList> InitRequests;
  std::map<std::string, std::string> Labels, NodeSelector;
  struct Toleration { std::string Key, Operator, Value, Effect; };
  std::vector<Toleration> Tolerations;
  struct HostPort { std::string HostIP, Protocol; int Port; };
  std::vector<HostPort> HostPorts;
  int32_t Priority = 1;
  int64_t CreationTimestamp = 0;
  bool Deleting = false;
};

struct Node {                                         // v1.Node fields the path reads (util/test_utils.go:52-63 BuildNode)
  std::string Name;
  ResourceList Allocatable;
  std::map<std::string, std::string> Labels;
  struct Taint { std::string Key, Value, Effect; };
  std::vector<Taint> Taints;
  bool Unschedulable = false, NotReady = false, NetworkUnavailable = false, MemoryPressure = false, DiskPressure = false, PIDPressure = false;
};

struct TaskInfo {                                     // api/job_info.go:36-54
  std::string UID, Job, Name, Namespace, NodeName;
  Resource Resreq, InitResreq;
  int Status = Pending;
  int32_t Priority = 1;
  std::shared_ptr<Pod> pod;
};

struct NodeInfo {                                     // api/node_info.go:28-47
  std::string Name;
  std::shared_ptr<Node> node;
  Resource Releasing, Idle, Used, Allocatable;
  std::map<std::string, std::shared_ptr<TaskInfo>> Tasks;
  // api/node_info.go:172-212
  void AddTask(const std::shared_ptr<TaskInfo>& task) {
    if (Tasks.count(task->Namespace + "/" + task->Name)) throw std::runtime_error("task already on node " + Name);
    switch (task->Status) {
      case Releasing_: Idle.SubUnchecked(task->Resreq); Releasing.Add(task->Resreq); break;
      case Pipelined: Releasing.SubUnchecked(task->Resreq); break;
      default: Idle.SubUnchecked(task->Resreq); break;
    }
    Used.Add(task->Resreq);
    task->NodeName = Name;
    Tasks[task->Namespace + "/" + task->Name] = task;
  }
  static constexpr int Releasing_ = api::Releasing;
};

struct JobInfo {                                      // api/job_info.go:127-154
  std::string UID, Name, Namespace, Queue;
  int32_t Priority = 0, MinAvailable = 0;
  int64_t CreationTimestamp = 0;
  std::map<std::string, std::shared_ptr<TaskInfo>> Tasks;
  std::map<int, std::map<std::string, std::shared_ptr<TaskInfo>>> TaskStatusIndex;
  Resource Allocated;
  void AddTaskInfo(const std::shared_ptr<TaskInfo>& t) { Tasks[t->UID] = t; TaskStatusIndex[t->Status][t->UID] = t; if (AllocatedStatus(t->Status)) Allocated.Add(t->Resreq); }
  void UpdateTaskStatus(const std::shared_ptr<TaskInfo>& t, int status) {   // job_info.go:247-264
    auto it = TaskStatusIndex.find(t->Status);
    if (it != TaskStatusIndex.end()) { if (AllocatedStatus(t->Status)) Allocated.SubUnchecked(t->Resreq); it->second.erase(t->UID); if (it->second.empty()) TaskStatusIndex.erase(it); }
    t->Status = status;
    TaskStatusIndex[status][t->UID] = t;
    if (AllocatedStatus(status)) Allocated.Add(t->Resreq);
  }
  int32_t ReadyTaskNum() const {                      // job_info.go:383-393
    int n = 0;
    for (auto& kv : TaskStatusIndex) if (AllocatedStatus(kv.first) || kv.first == Succeeded) n += (int)kv.second.size();
    return n;
  }
  bool Ready() const { return ReadyTaskNum() >= MinAvailable; }   // job_info.go:423-427
};

struct QueueInfo { std::string UID, Name; int32_t Weight = 1; int64_t CreationTimestamp = 0; };   // api/queue_info.go:74-81

}  // namespace api

// ------------------------------------------------------------------------------------------------ conf
namespace conf {
struct PluginOption {                                 // conf/scheduler_conf.go:33-56; -1 = nil *bool
  std::string Name;
  int EnabledJobOrder = -1, EnabledJobReady = -1, EnabledJobPipelined = -1, EnabledTaskOrder = -1, EnabledPreemptable = -1,
      EnabledReclaimable = -1, EnabledQueueOrder = -1, EnabledPredicate = -1, EnabledNodeOrder = -1;
  std::map<std::string, std::string> Arguments;
};
struct Tier { std::vector<PluginOption> Plugins; };
inline void ApplyPluginConfDefaults(PluginOption& o) {           // plugins/defaults.go:22-52: nil -> true
  for (int* p : {&o.EnabledJobOrder, &o.EnabledJobReady, &o.EnabledJobPipelined, &o.EnabledTaskOrder, &o.EnabledPreemptable,
                 &o.EnabledReclaimable, &o.EnabledQueueOrder, &o.EnabledPredicate, &o.EnabledNodeOrder}) if (*p < 0) *p = 1;
}
}  // namespace conf

// ----------------------------------------------------------------------------------------------- cache
namespace cache {
struct Binder { virtual ~Binder() = default; virtual void Bind(const api::TaskInfo& task, const std::string& hostname) = 0; };
struct FakeBinder : Binder {                          // util/test_utils.go:95-112
  std::map<std::string, std::string> Binds;
  void Bind(const api::TaskInfo& task, const std::string& hostname) override { Binds[task.Namespace + "/" + task.Name] = hostname; }
};
struct PodGroup { std::string Namespace, Name, Queue; int32_t MinMember = 0; int32_t Priority = 0; int64_t CreationTimestamp = 0; };
struct ClusterInfo {                                  // api/cluster_info.go:22-26
  std::map<std::string, std::shared_ptr<api::JobInfo>> Jobs;
  std::map<std::string, std::shared_ptr<api::NodeInfo>> Nodes;
  std::map<std::string, std::shared_ptr<api::QueueInfo>> Queues;
};
struct SchedulerCache {                               // the literal cache of allocate_test.go:154-177
  std::map<std::string, std::shared_ptr<api::NodeInfo>> Nodes;
  std::map<std::string, std::shared_ptr<api::JobInfo>> Jobs;
  std::map<std::string, std::shared_ptr<api::QueueInfo>> Queues;
  std::shared_ptr<Binder> binder;
  void AddNode(const api::Node& n) {
    auto ni = std::make_shared<api::NodeInfo>();
    ni->Name = n.Name; ni->node = std::make_shared<api::Node>(n);
    ni->Allocatable = api::Resource::New(n.Allocatable); ni->Idle = ni->Allocatable; ni->Idle.MaxTaskNum = ni->Allocatable.MaxTaskNum;
    Nodes[n.Name] = ni;
  }
  void AddQueue(const std::string& name, int32_t weight) { auto q = std::make_shared<api::QueueInfo>(); q->UID = q->Name = name; q->Weight = weight; Queues[name] = q; }
  std::shared_ptr<api::JobInfo> getOrCreateJob(const std::string& ns, const std::string& group) {
    const std::string id = ns + "/" + group;
    auto it = Jobs.find(id);
    if (it != Jobs.end()) return it->second;
    auto j = std::make_shared<api::JobInfo>(); j->UID = id; j->Namespace = ns; j->Name = group; Jobs[id] = j; return j;
  }
  void AddPodGroup(const PodGroup& pg) {
    auto j = getOrCreateJob(pg.Namespace, pg.Name);
    j->Queue = pg.Queue; j->MinAvailable = pg.MinMember; j->Priority = pg.Priority; j->CreationTimestamp = pg.CreationTimestamp;
    has_spec.insert(j->UID);
  }
  static int taskStatus(const api::Pod& p) {          // api/helpers.go:38-62
    if (p.Phase == "Running") return p.Deleting ? api::Releasing : api::Running;
    if (p.Phase == "Pending") { if (p.Deleting) return api::Releasing; return p.NodeName.empty() ? api::Pending : api::Bound; }
    if (p.Phase == "Succeeded") return api::Succeeded;
    if (p.Phase == "Failed") return api::Failed;
    return api::Unknown;
  }
  void AddPod(const api::Pod& pod) {
    auto t = std::make_shared<api::TaskInfo>();
    t->pod = std::make_shared<api::Pod>(pod);
    t->UID = pod.UID.empty() ? pod.Namespace + "-" + pod.Name : pod.UID;
    t->Name = pod.Name; t->Namespace = pod.Namespace; t->NodeName = pod.NodeName; t->Priority = pod.Priority;
    t->Resreq = api::Resource::New(pod.Requests);
    t->InitResreq = t->Resreq;
    for (auto& ir : pod.InitRequests) t->InitResreq.SetMaxResource(api::Resource::New(ir));   // api/pod_info.go:53-73
    t->Status = taskStatus(pod);
    auto j = getOrCreateJob(pod.Namespace, pod.GroupName);
    t->Job = j->UID;
    j->AddTaskInfo(t);
    // cache.addTask (cache/event_handlers.go): the job gets the task, then node.AddTask — which refuses a task that does not
    // fit into Idle (allocateIdleResource, node_info.go:161-167; a Pipelined task takes Releasing instead), so a node is never
    // over-committed in the snapshot
    if (!pod.NodeName.empty() && Nodes.count(pod.NodeName) && t->Status != api::Pending) {
      auto& n = *Nodes[pod.NodeName];
      if (t->Status == api::Pipelined || t->Resreq.LessEqual(n.Idle)) n.AddTask(t);
    }
  }
  ClusterInfo Snapshot() const {                      // cache/cache.go:627-683 (objects are shared, not deep-copied)
    ClusterInfo ci;
    ci.Nodes = Nodes; ci.Queues = Queues;
    for (auto& kv : Jobs) { if (!has_spec.count(kv.first)) continue; if (!Queues.count(kv.second->Queue)) continue; ci.Jobs[kv.first] = kv.second; }
    return ci;
  }
  std::set<std::string> has_spec;
};
}  // namespace cache

// ------------------------------------------------------------------------------------------- framework
namespace framework {

using Arguments = std::map<std::string, std::string>;                 // framework/arguments.go:26
class Session;

// Which built-in function a plugin registers.  A value of Foreign marks a callback the GPU path cannot honour.
enum class Builtin { PriorityTaskOrder, PriorityJobOrder, GangJobOrder, GangJobReady, GangJobPipelined, DrfJobOrder, DrfPreemptable,
                     ProportionQueueOrder, ProportionOverused, ProportionReclaimable, Predicates, NodeOrder, Foreign };

class Plugin {                                        // framework/interface.go:34-41
 public:
  virtual ~Plugin() = default;
  virtual std::string Name() const = 0;
  virtual void OnSessionOpen(Session* ssn) = 0;
  virtual void OnSessionClose(Session*) {}
};
class Action {                                        // framework/interface.go:20-32
 public:
  virtual ~Action() = default;
  virtual std::string Name() const = 0;
  virtual void Initialize() {}
  virtual void Execute(Session* ssn) = 0;
  virtual void UnInitialize() {}
};

using PluginBuilder = std::function<std::unique_ptr<Plugin>(const Arguments&)>;
inline std::map<std::string, PluginBuilder>& pluginBuilders() { static std::map<std::string, PluginBuilder> m; return m; }
inline void RegisterPluginBuilder(const std::string& name, PluginBuilder pb) { pluginBuilders()[name] = std::move(pb); }   // framework/plugins.go:30
inline void CleanupPluginBuilders() { pluginBuilders().clear(); }

class Session {                                       // framework/session.go:37-61
 public:
  std::string UID = "session";
  std::map<std::string, std::shared_ptr<api::JobInfo>> Jobs;
  std::map<std::string, std::shared_ptr<api::NodeInfo>> Nodes;
  std::map<std::string, std::shared_ptr<api::QueueInfo>> Queues;
  std::vector<conf::Tier> Tiers;
  cache::SchedulerCache* cache = nullptr;
  std::map<std::string, std::unique_ptr<Plugin>> plugins;
  std::map<std::string, std::vector<Builtin>> registered;              // plugin name -> what it registered
  std::map<std::string, Arguments> pluginArguments;

  // session_plugins.go:25-77 — the registration surface (source-compatible names)
  void AddJobOrderFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddQueueOrderFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddTaskOrderFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddPreemptableFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddReclaimableFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddJobReadyFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddJobPipelinedFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddPredicateFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddNodePrioritizers(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddNodeOrderFn(const std::string& name, Builtin fn) { AddNodePrioritizers(name, fn); }   // alias (SURVEY: API-name drift)
  void AddOverusedFn(const std::string& name, Builtin fn) { registered[name].push_back(fn); }
  void AddJobValidFn(const std::string&, Builtin) {}                   // dead at this commit (session.go:89-108)

  // session_plugins.go:182-200 — JobReady = AND of enabled JobReadyFns; only gang registers one
  bool JobReady(const api::JobInfo& job) const {
    for (auto& tier : Tiers) for (auto& p : tier.Plugins) {
      if (p.EnabledJobReady != 1) continue;
      auto it = registered.find(p.Name);
      if (it == registered.end()) continue;
      for (Builtin b : it->second) if (b == Builtin::GangJobReady && !job.Ready()) return false;
    }
    return true;
  }
  // framework/session.go:235-288
  void Allocate(const std::shared_ptr<api::TaskInfo>& task, const std::string& hostname) {
    auto& job = Jobs.at(task->Job);
    job->UpdateTaskStatus(task, api::Allocated);
    Nodes.at(hostname)->AddTask(task);
    if (JobReady(*job)) {
      auto it = job->TaskStatusIndex.find(api::Allocated);
      if (it != job->TaskStatusIndex.end()) {
        std::vector<std::shared_ptr<api::TaskInfo>> ts;
        for (auto& kv : it->second) ts.push_back(kv.second);
        for (auto& t : ts) dispatch(t);
      }
    }
  }
  // framework/session.go:194-232
  void Pipeline(const std::shared_ptr<api::TaskInfo>& task, const std::string& hostname) {
    Jobs.at(task->Job)->UpdateTaskStatus(task, api::Pipelined);
    Nodes.at(hostname)->AddTask(task);
  }
  // framework/session.go:290-314
  void dispatch(const std::shared_ptr<api::TaskInfo>& task) {
    if (cache && cache->binder) cache->binder->Bind(*task, task->NodeName);
    Jobs.at(task->Job)->UpdateTaskStatus(task, api::Binding);
  }
};

// framework/framework.go:30-52
inline std::unique_ptr<Session> OpenSession(cache::SchedulerCache* cache, const std::vector<conf::Tier>& tiers) {
  auto ssn = std::make_unique<Session>();
  ssn->cache = cache;
  auto snap = cache->Snapshot();
  ssn->Jobs = snap.Jobs; ssn->Nodes = snap.Nodes; ssn->Queues = snap.Queues;
  ssn->Tiers = tiers;
  for (auto& tier : tiers) for (auto& po : tier.Plugins) {
    auto it = pluginBuilders().find(po.Name);
    if (it == pluginBuilders().end()) continue;        // "Failed to get plugin" is only logged in the reference
    auto pl = it->second(po.Arguments);
    ssn->pluginArguments[pl->Name()] = po.Arguments;
    ssn->plugins[pl->Name()] = std::move(pl);
  }
  for (auto& kv : ssn->plugins) kv.second->OnSessionOpen(ssn.get());
  return ssn;
}
inline void CloseSession(Session* ssn) { for (auto& kv : ssn->plugins) kv.second->OnSessionClose(ssn); }

}  // namespace framework

// --------------------------------------------------------------------------------------------- plugins
namespace plugins {
using framework::Arguments; using framework::Builtin; using framework::Plugin; using framework::Session;
#define KB_PLUGIN(cls, pname, body) struct cls : Plugin { Arguments args; explicit cls(const Arguments& a) : args(a) {} \
  std::string Name() const override { return pname; } void OnSessionOpen(Session* ssn) override body }
KB_PLUGIN(priorityPlugin, "priority", { ssn->AddTaskOrderFn(Name(), Builtin::PriorityTaskOrder); ssn->AddJobOrderFn(Name(), Builtin::PriorityJobOrder); });
KB_PLUGIN(gangPlugin, "gang", { ssn->AddJobOrderFn(Name(), Builtin::GangJobOrder); ssn->AddJobReadyFn(Name(), Builtin::GangJobReady); ssn->AddJobPipelinedFn(Name(), Builtin::GangJobPipelined); });
KB_PLUGIN(drfPlugin, "drf", { ssn->AddPreemptableFn(Name(), Builtin::DrfPreemptable); ssn->AddJobOrderFn(Name(), Builtin::DrfJobOrder); });
KB_PLUGIN(proportionPlugin, "proportion", { ssn->AddQueueOrderFn(Name(), Builtin::ProportionQueueOrder); ssn->AddReclaimableFn(Name(), Builtin::ProportionReclaimable); ssn->AddOverusedFn(Name(), Builtin::ProportionOverused); });
KB_PLUGIN(predicatesPlugin, "predicates", { ssn->AddPredicateFn(Name(), Builtin::Predicates); });
KB_PLUGIN(nodeOrderPlugin, "nodeorder", { ssn->AddNodePrioritizers(Name(), Builtin::NodeOrder); });
KB_PLUGIN(conformancePlugin, "conformance", { (void)ssn; });
#undef KB_PLUGIN
inline void RegisterBuiltins() {                       // plugins/factory.go:31-42
  using framework::RegisterPluginBuilder;
  RegisterPluginBuilder("priority", [](const Arguments& a) { return std::make_unique<priorityPlugin>(a); });
  RegisterPluginBuilder("gang", [](const Arguments& a) { return std::make_unique<gangPlugin>(a); });
  RegisterPluginBuilder("drf", [](const Arguments& a) { return std::make_unique<drfPlugin>(a); });
  RegisterPluginBuilder("proportion", [](const Arguments& a) { return std::make_unique<proportionPlugin>(a); });
  RegisterPluginBuilder("predicates", [](const Arguments& a) { return std::make_unique<predicatesPlugin>(a); });
  RegisterPluginBuilder("nodeorder", [](const Arguments& a) { return std::make_unique<nodeOrderPlugin>(a); });
  RegisterPluginBuilder("conformance", [](const Arguments& a) { return std::make_unique<conformancePlugin>(a); });
}
}  // namespace plugins

// ------------------------------------------------------------------------------------- actions/allocate
namespace actions { namespace allocate {

// Flattened session: the SoA arrays of kb_snapshot plus the back-references needed to replay decisions.
struct Flat {
  uint32_t R = 2, W = 1, N = 0, T = 0, J = 0, Q = 0;
  std::vector<std::string> dims, nodeNames;
  std::vector<std::shared_ptr<api::TaskInfo>> tasks;
  std::vector<double> node_idle, node_releasing, node_used, node_allocatable, task_initreq, task_resreq, job_alloc0;
  std::vector<uint32_t> node_alloc_present, node_flags, task_res_present, task_n_aff, task_flags, task_uid_rank, job_task_off, job_alloc0_present, job_queue;
  std::vector<int64_t> node_alloc_cpu, node_alloc_mem, node_nz_cpu, node_nz_mem, task_nz_cpu, task_nz_mem, task_ctime, job_ctime, queue_ctime;
  std::vector<int32_t> node_pods, node_max_pods, task_prio, job_min_avail, job_ready0, job_prio, queue_weight;
  std::vector<uint64_t> node_labels, node_taints, node_ports, task_sel_req, task_aff, task_tol, task_port_own, task_port_conflict;
};

inline void resourceVec(const api::Resource& r, const std::vector<std::string>& dims, double* out, size_t stride, uint32_t* present) {
  out[0] = r.MilliCPU; out[stride] = r.Memory;
  uint32_t p = 0;
  for (auto& kv : r.ScalarResources) { size_t k = std::find(dims.begin(), dims.end(), kv.first) - dims.begin(); out[k * stride] = kv.second; p |= 1u << k; }
  if (present) *present = p;
}

// What the Go shim's Flatten(ssn) does (INTEGRATION.md §"What Flatten must compute"); mirrors builder.py::flatten.
inline Flat Flatten(const framework::Session& ssn) {
  Flat f;
  std::set<std::string> sc;
  for (auto& kv : ssn.Nodes) for (auto& s : kv.second->Allocatable.ScalarResources) sc.insert(s.first);
  for (auto& jk : ssn.Jobs) for (auto& tk : jk.second->Tasks) { for (auto& s : tk.second->Resreq.ScalarResources) sc.insert(s.first); for (auto& s : tk.second->InitResreq.ScalarResources) sc.insert(s.first); }
  f.dims = {"cpu", "memory"}; for (auto& s : sc) f.dims.push_back(s);
  f.R = (uint32_t)f.dims.size();
  if (f.R > KB_MAX_R) throw std::runtime_error("too many scalar resources for KB_MAX_R");
  std::vector<std::shared_ptr<api::NodeInfo>> nodes; for (auto& kv : ssn.Nodes) nodes.push_back(kv.second);       // std::map: ascending Name
  std::vector<std::shared_ptr<api::QueueInfo>> queues; for (auto& kv : ssn.Queues) queues.push_back(kv.second);
  std::vector<std::shared_ptr<api::JobInfo>> jobs; for (auto& kv : ssn.Jobs) jobs.push_back(kv.second);          // ascending JobID
  std::map<std::string, uint32_t> qidx; for (uint32_t i = 0; i < queues.size(); ++i) qidx[queues[i]->UID] = i;
  f.N = (uint32_t)nodes.size(); f.J = (uint32_t)jobs.size(); f.Q = (uint32_t)queues.size();
  // pending tasks grouped by job; UID ranks
  std::vector<std::string> uids;
  f.job_task_off.assign(f.J + 1, 0);
  for (uint32_t j = 0; j < f.J; ++j) {
    auto it = jobs[j]->TaskStatusIndex.find(api::Pending);
    if (it != jobs[j]->TaskStatusIndex.end()) for (auto& kv : it->second) { f.tasks.push_back(kv.second); uids.push_back(kv.second->UID); }
    f.job_task_off[j + 1] = (uint32_t)f.tasks.size();
  }
  f.T = (uint32_t)f.tasks.size();
  std::sort(uids.begin(), uids.end());
  // atoms: selector requirements (nodeSelector pairs), NoSchedule/NoExecute taints, host ports
  std::map<std::pair<std::string, std::string>, uint32_t> labelAtoms;
  std::map<std::tuple<std::string, std::string, std::string>, uint32_t> taintAtoms;
  std::map<std::tuple<std::string, std::string, int>, uint32_t> portAtoms;
  auto san = [](const api::Pod::HostPort& h) { return std::make_tuple(h.HostIP.empty() ? std::string("0.0.0.0") : h.HostIP, h.Protocol.empty() ? std::string("TCP") : h.Protocol, h.Port); };
  for (auto& t : f.tasks) for (auto& kv : t->pod->NodeSelector) labelAtoms.emplace(std::make_pair(kv.first, kv.second), (uint32_t)labelAtoms.size());
  for (auto& n : nodes) for (auto& t : n->node->Taints) if (t.Effect == "NoSchedule" || t.Effect == "NoExecute") taintAtoms.emplace(std::make_tuple(t.Key, t.Value, t.Effect), (uint32_t)taintAtoms.size());
  for (auto& jk : ssn.Jobs) for (auto& tk : jk.second->Tasks) for (auto& h : tk.second->pod->HostPorts) if (h.Port > 0) portAtoms.emplace(san(h), (uint32_t)portAtoms.size());
  size_t need = std::max({labelAtoms.size(), taintAtoms.size(), portAtoms.size(), (size_t)1});
  f.W = (uint32_t)((need + 63) / 64);
  if (f.W > KB_MAX_W) throw std::runtime_error("too many atoms for KB_MAX_W");
  const uint32_t R = f.R, W = f.W, N = f.N, T = f.T, J = f.J, Q = f.Q;
  auto setbit = [](std::vector<uint64_t>& a, size_t stride, size_t i, uint32_t atom) { a[(atom / 64) * stride + i] |= 1ull << (atom % 64); };
  // nodes
  f.node_idle.assign((size_t)R * N, 0); f.node_releasing = f.node_used = f.node_allocatable = f.node_idle;
  f.node_alloc_present.assign(N, 0); f.node_flags.assign(N, 0); f.node_alloc_cpu.assign(N, 0); f.node_alloc_mem = f.node_nz_cpu = f.node_nz_mem = f.node_alloc_cpu;
  f.node_pods.assign(N, 0); f.node_max_pods.assign(N, 0);
  f.node_labels.assign((size_t)W * std::max(1u, N), 0); f.node_taints = f.node_ports = f.node_labels;
  for (uint32_t i = 0; i < N; ++i) {
    auto& n = *nodes[i];
    f.nodeNames.push_back(n.Name);
    resourceVec(n.Idle, f.dims, &f.node_idle[i], N, nullptr); resourceVec(n.Releasing, f.dims, &f.node_releasing[i], N, nullptr);
    resourceVec(n.Used, f.dims, &f.node_used[i], N, nullptr); resourceVec(n.Allocatable, f.dims, &f.node_allocatable[i], N, &f.node_alloc_present[i]);
    f.node_alloc_cpu[i] = (int64_t)n.Allocatable.MilliCPU; f.node_alloc_mem[i] = (int64_t)n.Allocatable.Memory;
    f.node_max_pods[i] = n.Allocatable.MaxTaskNum; f.node_pods[i] = (int32_t)n.Tasks.size();
    uint32_t fl = 0;
    fl |= n.node->NotReady ? KB_NODE_NOT_READY : 0u;
    fl |= n.node->NetworkUnavailable ? KB_NODE_NET_UNAVAILABLE : 0u;
    fl |= n.node->Unschedulable ? KB_NODE_UNSCHEDULABLE : 0u;
    fl |= n.node->MemoryPressure ? KB_NODE_MEM_PRESSURE : 0u;
    fl |= n.node->DiskPressure ? KB_NODE_DISK_PRESSURE : 0u;
    fl |= n.node->PIDPressure ? KB_NODE_PID_PRESSURE : 0u;
    f.node_flags[i] = fl;
    for (auto& kv : n.Tasks) {                         // nonzero requests + used ports of every task on the node
      auto& p = *kv.second->pod;
      f.node_nz_cpu[i] += p.Requests.count("cpu") ? (int64_t)std::ceil(p.Requests.at("cpu") * 1000.0 - 1e-9) : 100;           // non_zero.go:36
      f.node_nz_mem[i] += p.Requests.count("memory") ? (int64_t)p.Requests.at("memory") : 200ll * 1024 * 1024;             // non_zero.go:38
      for (auto& h : p.HostPorts) if (h.Port > 0) setbit(f.node_ports, N, i, portAtoms.at(san(h)));
    }
    for (auto& la : labelAtoms) { auto it = n.node->Labels.find(la.first.first); if (it != n.node->Labels.end() && it->second == la.first.second) setbit(f.node_labels, N, i, la.second); }
    for (auto& t : n.node->Taints) { auto it = taintAtoms.find(std::make_tuple(t.Key, t.Value, t.Effect)); if (it != taintAtoms.end()) setbit(f.node_taints, N, i, it->second); }
  }
  // tasks
  f.task_initreq.assign((size_t)R * std::max(1u, T), 0); f.task_resreq = f.task_initreq;
  f.task_res_present.assign(std::max(1u, T), 0); f.task_n_aff = f.task_flags = f.task_uid_rank = f.task_res_present;
  f.task_nz_cpu.assign(std::max(1u, T), 0); f.task_nz_mem = f.task_ctime = f.task_nz_cpu; f.task_prio.assign(std::max(1u, T), 0);
  f.task_sel_req.assign((size_t)W * std::max(1u, T), 0); f.task_tol = f.task_port_own = f.task_port_conflict = f.task_sel_req;
  f.task_aff.assign((size_t)KB_MAX_AFF_TERMS * W * std::max(1u, T), 0);
  for (uint32_t t = 0; t < T; ++t) {
    auto& ti = *f.tasks[t]; auto& p = *ti.pod;
    resourceVec(ti.Resreq, f.dims, &f.task_resreq[t], T, &f.task_res_present[t]); resourceVec(ti.InitResreq, f.dims, &f.task_initreq[t], T, nullptr);
    f.task_nz_cpu[t] = p.Requests.count("cpu") ? (int64_t)std::ceil(p.Requests.at("cpu") * 1000.0 - 1e-9) : 100;
    f.task_nz_mem[t] = p.Requests.count("memory") ? (int64_t)p.Requests.at("memory") : 200ll * 1024 * 1024;
    for (auto& kv : p.NodeSelector) setbit(f.task_sel_req, T, t, labelAtoms.at({kv.first, kv.second}));
    for (auto& ta : taintAtoms) for (auto& tol : p.Tolerations) {                // Toleration.ToleratesTaint, toleration.go:37-56
      if (!tol.Effect.empty() && tol.Effect != std::get<2>(ta.first)) continue;
      if (!tol.Key.empty() && tol.Key != std::get<0>(ta.first)) continue;
      bool ok = tol.Operator == "Exists" || ((tol.Operator.empty() || tol.Operator == "Equal") && tol.Value == std::get<1>(ta.first));
      if (ok) setbit(f.task_tol, T, t, ta.second);
    }
    for (auto& h : p.HostPorts) { if (h.Port <= 0) continue; auto me = san(h); setbit(f.task_port_own, T, t, portAtoms.at(me));
      for (auto& pa : portAtoms) {                                                // HostPortInfo.CheckConflict, host_ports.go:96-125
        if (std::get<1>(pa.first) != std::get<1>(me) || std::get<2>(pa.first) != std::get<2>(me)) continue;
        if (std::get<0>(pa.first) == std::get<0>(me) || std::get<0>(pa.first) == "0.0.0.0" || std::get<0>(me) == "0.0.0.0") setbit(f.task_port_conflict, T, t, pa.second);
      } }
    if (p.Requests.empty()) f.task_flags[t] |= KB_TASK_BEST_EFFORT_QOS;
    f.task_prio[t] = ti.Priority; f.task_ctime[t] = p.CreationTimestamp;
    f.task_uid_rank[t] = (uint32_t)(std::lower_bound(uids.begin(), uids.end(), ti.UID) - uids.begin());
  }
  // jobs / queues
  f.job_min_avail.assign(std::max(1u, J), 0); f.job_ready0 = f.job_prio = f.job_min_avail; f.job_alloc0.assign((size_t)R * std::max(1u, J), 0);
  f.job_alloc0_present.assign(std::max(1u, J), 0); f.job_queue = f.job_alloc0_present; f.job_ctime.assign(std::max(1u, J), 0);
  for (uint32_t j = 0; j < J; ++j) {
    auto& job = *jobs[j];
    f.job_min_avail[j] = job.MinAvailable; f.job_ready0[j] = job.ReadyTaskNum(); f.job_prio[j] = job.Priority; f.job_ctime[j] = job.CreationTimestamp;
    f.job_queue[j] = qidx.at(job.Queue);
    resourceVec(job.Allocated, f.dims, &f.job_alloc0[j], J, &f.job_alloc0_present[j]);
  }
  f.queue_weight.assign(std::max(1u, Q), 0); f.queue_ctime.assign(std::max(1u, Q), 0);
  for (uint32_t q = 0; q < Q; ++q) { f.queue_weight[q] = queues[q]->Weight; f.queue_ctime[q] = queues[q]->CreationTimestamp; }
  return f;
}

// One action on the GPU: flatten the session as it is NOW -> kb_session_load -> kb_allocate | kb_backfill -> replay the
// decisions through the session's own Allocate / Pipeline.  Every action flattens afresh, so "allocate, backfill" (the
// default list, pkg/scheduler/util.go:31-42) is two such calls, exactly like the reference runs two Execute()s on one
// session.  (A shim that wants to save the second flatten + upload can call kb_allocate and kb_backfill back to back on
// one loaded session instead: see INTEGRATION.md.)
inline void ExecuteOnGpu(framework::Session* ssn, const bool backfill) {
  {
    // tiers: only built-in registrations can be honoured
    std::vector<kb_plugin_option> opts; std::vector<kb_tier> tiers; std::vector<std::vector<const char*>> keys, vals;
    size_t total = 0; for (auto& t : ssn->Tiers) total += t.Plugins.size();
    opts.reserve(total); keys.reserve(total); vals.reserve(total);
    for (auto& t : ssn->Tiers) {
      kb_tier ct{}; ct.n_plugins = (uint32_t)t.Plugins.size(); ct.plugins = opts.data() + opts.size();
      for (auto& p : t.Plugins) {
        auto reg = ssn->registered.find(p.Name);
        if (reg != ssn->registered.end()) for (auto b : reg->second) if (b == framework::Builtin::Foreign)
          throw std::runtime_error("plugin " + p.Name + " registered a non built-in function: the GPU path cannot honour it (no CPU fallback)");
        kb_plugin_option o{}; o.name = p.Name.c_str();
        o.enabled_job_order = p.EnabledJobOrder == 1; o.enabled_job_ready = p.EnabledJobReady == 1; o.enabled_job_pipelined = p.EnabledJobPipelined == 1;
        o.enabled_task_order = p.EnabledTaskOrder == 1; o.enabled_preemptable = p.EnabledPreemptable == 1; o.enabled_reclaimable = p.EnabledReclaimable == 1;
        o.enabled_queue_order = p.EnabledQueueOrder == 1; o.enabled_predicate = p.EnabledPredicate == 1; o.enabled_node_order = p.EnabledNodeOrder == 1;
        keys.emplace_back(); vals.emplace_back();
        for (auto& kv : p.Arguments) { keys.back().push_back(kv.first.c_str()); vals.back().push_back(kv.second.c_str()); }
        o.n_args = (uint32_t)keys.back().size(); o.arg_keys = keys.back().data(); o.arg_values = vals.back().data();
        opts.push_back(o);
      }
      tiers.push_back(ct);
    }
    kb_plugin_conf conf{(uint32_t)tiers.size(), tiers.data()};
    Flat f = Flatten(*ssn);
    kb_snapshot s{};
    s.abi_version = KB_ABI_VERSION; s.R = f.R; s.W = f.W; s.N = f.N; s.T = f.T; s.J = f.J; s.Q = f.Q;
    s.node_idle = f.node_idle.data(); s.node_releasing = f.node_releasing.data(); s.node_used = f.node_used.data(); s.node_allocatable = f.node_allocatable.data();
    s.node_alloc_present = f.node_alloc_present.data(); s.node_alloc_cpu = f.node_alloc_cpu.data(); s.node_alloc_mem = f.node_alloc_mem.data();
    s.node_nz_cpu = f.node_nz_cpu.data(); s.node_nz_mem = f.node_nz_mem.data(); s.node_pods = f.node_pods.data(); s.node_max_pods = f.node_max_pods.data();
    s.node_flags = f.node_flags.data(); s.node_labels = f.node_labels.data(); s.node_taints = f.node_taints.data(); s.node_ports = f.node_ports.data();
    s.task_initreq = f.task_initreq.data(); s.task_resreq = f.task_resreq.data(); s.task_res_present = f.task_res_present.data();
    s.task_nz_cpu = f.task_nz_cpu.data(); s.task_nz_mem = f.task_nz_mem.data(); s.task_sel_req = f.task_sel_req.data(); s.task_aff_terms = f.task_aff.data();
    s.task_n_aff_terms = f.task_n_aff.data(); s.task_tol = f.task_tol.data(); s.task_port_own = f.task_port_own.data(); s.task_port_conflict = f.task_port_conflict.data();
    s.task_flags = f.task_flags.data(); s.task_prio = f.task_prio.data(); s.task_ctime = f.task_ctime.data(); s.task_uid_rank = f.task_uid_rank.data();
    s.job_task_off = f.job_task_off.data(); s.job_min_avail = f.job_min_avail.data(); s.job_ready0 = f.job_ready0.data(); s.job_alloc0 = f.job_alloc0.data();
    s.job_alloc0_present = f.job_alloc0_present.data(); s.job_queue = f.job_queue.data(); s.job_prio = f.job_prio.data(); s.job_ctime = f.job_ctime.data();
    s.queue_weight = f.queue_weight.data(); s.queue_ctime = f.queue_ctime.data();

    kb_engine_opts eo{}; eo.abi_version = KB_ABI_VERSION; eo.device = 0; eo.rank = 0; eo.world_size = 1;
    kb_engine* eng = nullptr;
    int rc = kb_engine_create(&eo, &eng);
    if (rc != KB_OK) throw std::runtime_error(std::string("kb_engine_create: ") + kb_last_error(nullptr) + " [" + kb_status_str(rc) + "]");
    struct Guard { kb_engine* e; ~Guard() { kb_engine_destroy(e); } } guard{eng};
    rc = kb_session_load(eng, &s, &conf);
    if (rc != KB_OK) throw std::runtime_error(std::string("kb_session_load: ") + kb_last_error(eng) + " [" + kb_status_str(rc) + "]");
    std::vector<kb_decision> dec(std::max(1u, f.T));
    kb_stats st{};
    rc = backfill ? kb_backfill(eng, dec.data(), &st) : kb_allocate(eng, dec.data(), &st);
    if (rc != KB_OK) throw std::runtime_error(std::string(backfill ? "kb_backfill: " : "kb_allocate: ") + kb_last_error(eng) + " [" + kb_status_str(rc) + "]");
    // replay in call order through the session's own Allocate / Pipeline (gang dispatch, Binder.Bind happen there)
    std::vector<uint32_t> order;
    for (uint32_t t = 0; t < f.T; ++t) if (dec[t].kind == KB_KIND_ALLOCATED || dec[t].kind == KB_KIND_PIPELINED) order.push_back(t);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return dec[a].step < dec[b].step; });
    for (uint32_t t : order) {
      if (dec[t].kind == KB_KIND_ALLOCATED) ssn->Allocate(f.tasks[t], f.nodeNames[dec[t].node]);
      else ssn->Pipeline(f.tasks[t], f.nodeNames[dec[t].node]);
    }
  }
}

class allocateAction : public framework::Action {
 public:
  std::string Name() const override { return "allocate"; }               // allocate.go:38
  // actions/allocate/allocate.go:43-194, on the GPU
  void Execute(framework::Session* ssn) override { ExecuteOnGpu(ssn, false); }
};
inline std::unique_ptr<framework::Action> New() { return std::make_unique<allocateAction>(); }

}}  // namespace actions::allocate

namespace actions { namespace backfill {
class backfillAction : public framework::Action {
 public:
  std::string Name() const override { return "backfill"; }               // backfill.go:35
  // actions/backfill/backfill.go:40-71, on the GPU
  void Execute(framework::Session* ssn) override { allocate::ExecuteOnGpu(ssn, true); }
};
inline std::unique_ptr<framework::Action> New() { return std::make_unique<backfillAction>(); }
}}  // namespace actions::backfill

}  // namespace kb
