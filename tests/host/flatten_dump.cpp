// flatten_dump.cpp — dumps what the C++ host mirror's Flatten(ssn) (kbhost.hpp; = the Go shim's Flatten of INTEGRATION.md)
// computes for a hand-built session, as JSON.  tests/test_host_cpp.py builds the SAME objects with kube_batch_b200/builder.py
// (the executable specification of the flattening) and compares every numeric array plus the three bitmask RELATIONS
// (selector match, taint tolerated, host-port conflict) per (task, node) — atom numbering may legitimately differ.
// No GPU involved: Flatten is host code.
#include <cstdio>
#include <iostream>

#include "../../kube_batch_b200/host/kbhost.hpp"

using namespace kb;

template <typename T> static void arr(const char* name, const std::vector<T>& v, size_t n, bool last = false) {
  std::cout << "\"" << name << "\": [";
  for (size_t i = 0; i < n && i < v.size(); ++i) { if (i) std::cout << ","; std::cout << (long double)v[i]; }
  std::cout << "]" << (last ? "" : ",") << "\n";
}

int main() {
  std::cout.precision(21);
  const double G = 1e9;
  cache::SchedulerCache sc;
  sc.binder = std::make_shared<cache::FakeBinder>();
  {
    api::Node n; n.Name = "n0"; n.Allocatable = {{"cpu", 8}, {"memory", 32 * G}, {"pods", 10}, {"nvidia.com/gpu", 4}}; n.Labels = {{"zone", "a"}}; sc.AddNode(n);
    n = api::Node(); n.Name = "n1"; n.Allocatable = {{"cpu", 16}, {"memory", 64 * G}, {"pods", 20}}; n.Labels = {{"zone", "b"}};
    n.Taints = {{"dedicated", "batch", "NoSchedule"}, {"soft", "x", "PreferNoSchedule"}}; sc.AddNode(n);
    n = api::Node(); n.Name = "n2"; n.Allocatable = {{"cpu", 4}, {"memory", 8 * G}, {"pods", 5}, {"nvidia.com/gpu", 2}}; n.Labels = {{"zone", "b"}};
    n.Unschedulable = true; n.MemoryPressure = true; sc.AddNode(n);
  }
  sc.AddQueue("q1", 1); sc.AddQueue("q2", 3);
  sc.AddPodGroup({"ns", "old", "q1", 1}); sc.AddPodGroup({"ns", "pgA", "q1", 2}); sc.AddPodGroup({"ns", "pgB", "q2", 1, 7, 5});
  auto pod = [](const char* name, const char* node, const char* phase, api::ResourceList req, const char* group, int64_t ctime) {
    api::Pod p; p.Namespace = "ns"; p.Name = name; p.UID = std::string("ns-") + name; p.NodeName = node; p.Phase = phase;
    p.Requests = std::move(req); p.GroupName = group; p.CreationTimestamp = ctime; return p;
  };
  {
    api::Pod p = pod("r0", "n0", "Running", {{"cpu", 2}, {"memory", 4 * G}}, "old", 1); p.HostPorts = {{"", "TCP", 8080}}; sc.AddPod(p);
    p = pod("r1", "n1", "Running", {{"cpu", 3}, {"memory", 6 * G}}, "old", 2); p.Deleting = true; sc.AddPod(p);
    p = pod("r2", "n2", "Running", {{"cpu", 6}, {"memory", 1 * G}}, "old", 3); sc.AddPod(p);      // 6 cpu on a 4-cpu node: node.AddTask refuses it
    p = pod("a0", "", "Pending", {{"cpu", 1}, {"memory", 1 * G}}, "pgA", 10); p.NodeSelector = {{"zone", "b"}};
    p.Tolerations = {{"dedicated", "Equal", "batch", "NoSchedule"}}; sc.AddPod(p);
    p = pod("a1", "", "Pending", {{"cpu", 1}, {"memory", 1 * G}}, "pgA", 11); p.NodeSelector = {{"zone", "b"}}; p.HostPorts = {{"", "TCP", 8080}}; sc.AddPod(p);
    p = pod("b0", "", "Pending", {}, "pgB", 12); sc.AddPod(p);
    p = pod("b1", "", "Pending", {{"cpu", 0.5}, {"nvidia.com/gpu", 1}}, "pgB", 13); p.HostPorts = {{"10.0.0.1", "TCP", 8080}, {"", "UDP", 53}}; p.Priority = 5;
    p.Tolerations = {{"", "Exists", "", ""}}; sc.AddPod(p);
  }
  plugins::RegisterBuiltins();
  conf::PluginOption gang; gang.Name = "gang"; gang.EnabledJobReady = 1;
  auto ssn = framework::OpenSession(&sc, {conf::Tier{{gang}}});
  actions::allocate::Flat f = actions::allocate::Flatten(*ssn);
  const uint32_t R = f.R, W = f.W, N = f.N, T = f.T, J = f.J, Q = f.Q;
  std::cout << "{\"R\":" << R << ",\"W\":" << W << ",\"N\":" << N << ",\"T\":" << T << ",\"J\":" << J << ",\"Q\":" << Q << ",\n";
  std::cout << "\"dims\": ["; for (size_t i = 0; i < f.dims.size(); ++i) std::cout << (i ? "," : "") << "\"" << f.dims[i] << "\""; std::cout << "],\n";
  std::cout << "\"nodes\": ["; for (size_t i = 0; i < f.nodeNames.size(); ++i) std::cout << (i ? "," : "") << "\"" << f.nodeNames[i] << "\""; std::cout << "],\n";
  std::cout << "\"tasks\": ["; for (size_t i = 0; i < f.tasks.size(); ++i) std::cout << (i ? "," : "") << "\"" << f.tasks[i]->Namespace << "/" << f.tasks[i]->Name << "\""; std::cout << "],\n";
  arr("node_idle", f.node_idle, (size_t)R * N); arr("node_releasing", f.node_releasing, (size_t)R * N); arr("node_used", f.node_used, (size_t)R * N);
  arr("node_allocatable", f.node_allocatable, (size_t)R * N); arr("node_alloc_present", f.node_alloc_present, N);
  arr("node_alloc_cpu", f.node_alloc_cpu, N); arr("node_alloc_mem", f.node_alloc_mem, N); arr("node_nz_cpu", f.node_nz_cpu, N); arr("node_nz_mem", f.node_nz_mem, N);
  arr("node_pods", f.node_pods, N); arr("node_max_pods", f.node_max_pods, N); arr("node_flags", f.node_flags, N);
  arr("task_initreq", f.task_initreq, (size_t)R * T); arr("task_resreq", f.task_resreq, (size_t)R * T); arr("task_res_present", f.task_res_present, T);
  arr("task_nz_cpu", f.task_nz_cpu, T); arr("task_nz_mem", f.task_nz_mem, T); arr("task_flags", f.task_flags, T); arr("task_prio", f.task_prio, T);
  arr("task_ctime", f.task_ctime, T); arr("task_uid_rank", f.task_uid_rank, T);
  arr("job_task_off", f.job_task_off, J + 1); arr("job_min_avail", f.job_min_avail, J); arr("job_ready0", f.job_ready0, J); arr("job_alloc0", f.job_alloc0, (size_t)R * J);
  arr("job_alloc0_present", f.job_alloc0_present, J); arr("job_queue", f.job_queue, J); arr("job_prio", f.job_prio, J); arr("job_ctime", f.job_ctime, J);
  arr("queue_weight", f.queue_weight, Q); arr("queue_ctime", f.queue_ctime, Q);
  // relations per (task, node)
  std::vector<int> sel((size_t)T * N), tol((size_t)T * N), conflict((size_t)T * N);
  for (uint32_t t = 0; t < T; ++t) for (uint32_t n = 0; n < N; ++n) {
    bool s = true, to = true, c = false;
    for (uint32_t w = 0; w < W; ++w) {
      s = s && (f.task_sel_req[(size_t)w * T + t] & ~f.node_labels[(size_t)w * N + n]) == 0;
      to = to && (f.node_taints[(size_t)w * N + n] & ~f.task_tol[(size_t)w * T + t]) == 0;
      c = c || (f.task_port_conflict[(size_t)w * T + t] & f.node_ports[(size_t)w * N + n]) != 0;
    }
    sel[(size_t)t * N + n] = s; tol[(size_t)t * N + n] = to; conflict[(size_t)t * N + n] = c;
  }
  arr("rel_selector", sel, sel.size()); arr("rel_tolerated", tol, tol.size()); arr("rel_port_conflict", conflict, conflict.size(), true);
  std::cout << "}\n";
  framework::CloseSession(ssn.get());
  return 0;
}
