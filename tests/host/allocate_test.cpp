// allocate_test.cpp — the reference's action test, restated against the C++ host mirror + libkbgpu.so.
// Mirrors /root/reference/pkg/scheduler/actions/allocate/allocate_test.go:38-212 (same objects, same tiers,
// same expected bind maps), plus a gang + predicates session in the style of BASELINE config 1.
// Exit code: 0 all cases passed, 1 a case failed, 3 no CUDA device (the engine has no CPU fallback).
#include <cstdio>
#include <cstring>
#include <iostream>

#include "../../kube_batch_b200/host/kbhost.hpp"

using namespace kb;

static api::ResourceList BuildResourceList(double cpu, double memory) { return {{"cpu", cpu}, {"memory", memory}, {"nvidia.com/gpu", 0}}; }   // util/test_utils.go:34-40
static api::Node BuildNode(const std::string& name, api::ResourceList alloc) { api::Node n; n.Name = name; n.Allocatable = std::move(alloc); return n; }
static api::Pod BuildPod(const std::string& ns, const std::string& name, const std::string& node, const std::string& phase,
                         api::ResourceList req, const std::string& group) {
  api::Pod p; p.Namespace = ns; p.Name = name; p.UID = ns + "-" + name; p.NodeName = node; p.Phase = phase; p.Requests = std::move(req); p.GroupName = group; return p;
}
static conf::PluginOption Opt(const std::string& name) { conf::PluginOption o; o.Name = name; return o; }

struct Case {
  std::string name;
  std::vector<cache::PodGroup> podGroups;
  std::vector<api::Pod> pods;
  std::vector<api::Node> nodes;
  std::vector<std::pair<std::string, int>> queues;
  std::vector<conf::Tier> tiers;
  std::map<std::string, std::string> expected;
  std::vector<std::string> actions = {"allocate"};
};

int main() {
  const double G = 1e9, Gi = 1024.0 * 1024.0 * 1024.0;
  plugins::RegisterBuiltins();                          // framework.RegisterPluginBuilder("drf", drf.New) ... (allocate_test.go:39-41)

  conf::PluginOption drf = Opt("drf"); drf.EnabledPreemptable = 1; drf.EnabledJobOrder = 1;                    // allocate_test.go:184-188
  conf::PluginOption prop = Opt("proportion"); prop.EnabledQueueOrder = 1; prop.EnabledReclaimable = 1;         // :189-193
  std::vector<conf::Tier> refTiers = {conf::Tier{{drf, prop}}};

  std::vector<Case> tests;
  tests.push_back({"one Job with two Pods on one node",                                                         // :51-85
                   {{"c1", "pg1", "c1"}},
                   {BuildPod("c1", "p1", "", "Pending", BuildResourceList(1, 1 * G), "pg1"), BuildPod("c1", "p2", "", "Pending", BuildResourceList(1, 1 * G), "pg1")},
                   {BuildNode("n1", BuildResourceList(2, 4 * Gi))}, {{"c1", 1}}, refTiers,
                   {{"c1/p1", "n1"}, {"c1/p2", "n1"}}});
  tests.push_back({"two Jobs on one node",                                                                      // :86-144
                   {{"c1", "pg1", "c1"}, {"c2", "pg2", "c2"}},
                   {BuildPod("c1", "p1", "", "Pending", BuildResourceList(1, 1 * G), "pg1"), BuildPod("c1", "p2", "", "Pending", BuildResourceList(1, 1 * G), "pg1"),
                    BuildPod("c2", "p1", "", "Pending", BuildResourceList(1, 1 * G), "pg2"), BuildPod("c2", "p2", "", "Pending", BuildResourceList(1, 1 * G), "pg2")},
                   {BuildNode("n1", BuildResourceList(2, 4 * G))}, {{"c1", 1}, {"c2", 1}}, refTiers,
                   {{"c2/p1", "n1"}, {"c1/p1", "n1"}}});
  {
    // gang + predicates: pgA (minMember 2) fits and is dispatched; pgB (minMember 3) gets only 2 of 3 tasks ->
    // they hold node.Idle but are never bound (no rollback in allocate)
    conf::PluginOption gang = Opt("gang"); gang.EnabledJobOrder = 1; gang.EnabledJobReady = 1;
    conf::PluginOption pred = Opt("predicates"); pred.EnabledPredicate = 1;
    api::ResourceList node = BuildResourceList(4, 16 * G); node["pods"] = 110;
    std::vector<api::Pod> pods;
    for (int i = 0; i < 2; ++i) pods.push_back(BuildPod("ns", "a" + std::to_string(i), "", "Pending", BuildResourceList(1, 1 * G), "pgA"));
    for (int i = 0; i < 3; ++i) pods.push_back(BuildPod("ns", "b" + std::to_string(i), "", "Pending", BuildResourceList(1, 1 * G), "pgB"));
    tests.push_back({"gang: dispatch only when minMember is reached", {{"ns", "pgA", "q", 2}, {"ns", "pgB", "q", 3}}, pods,
                     {BuildNode("n1", node)}, {{"q", 1}}, {conf::Tier{{gang}}, conf::Tier{{pred}}},
                     {{"ns/a0", "n1"}, {"ns/a1", "n1"}}});
  }

  {
    // "allocate, backfill" (the default action list, pkg/scheduler/util.go:31-42): the gang of 3 has two regular pods and
    // one best-effort pod; allocate leaves the job one short, backfill (backfill.go:40-71) places the best-effort pod and the
    // session then dispatches all three.  The selector-less best-effort pod of pgBE goes to the first node by name.
    conf::PluginOption gang = Opt("gang"); gang.EnabledJobOrder = 1; gang.EnabledJobReady = 1;
    conf::PluginOption pred = Opt("predicates"); pred.EnabledPredicate = 1;
    conf::PluginOption nodeorder = Opt("nodeorder"); nodeorder.EnabledNodeOrder = 1;
    api::ResourceList node = BuildResourceList(4, 16 * G); node["pods"] = 110;
    std::vector<api::Pod> pods;
    pods.push_back(BuildPod("ns", "g0", "", "Pending", BuildResourceList(1, 1 * G), "pgG"));
    pods.push_back(BuildPod("ns", "g1", "", "Pending", BuildResourceList(1, 1 * G), "pgG"));
    pods.push_back(BuildPod("ns", "g2-be", "", "Pending", {}, "pgG"));
    pods.push_back(BuildPod("ns", "solo-be", "", "Pending", {}, "pgBE"));
    Case c{"allocate, backfill: a best-effort pod completes the gang", {{"ns", "pgG", "q", 3}, {"ns", "pgBE", "q", 1}}, pods,
           {BuildNode("n1", node), BuildNode("n2", node)}, {{"q", 1}}, {conf::Tier{{gang}}, conf::Tier{{pred, nodeorder}}},
           {{"ns/g0", "n1"}, {"ns/g1", "n2"}, {"ns/g2-be", "n1"}, {"ns/solo-be", "n1"}}};
    c.actions = {"allocate", "backfill"};
    tests.push_back(c);
  }

  int failed = 0;
  for (size_t i = 0; i < tests.size(); ++i) {
    auto& test = tests[i];
    auto binder = std::make_shared<cache::FakeBinder>();
    cache::SchedulerCache schedulerCache;
    schedulerCache.binder = binder;
    for (auto& node : test.nodes) schedulerCache.AddNode(node);
    for (auto& pod : test.pods) schedulerCache.AddPod(pod);
    for (auto& ss : test.podGroups) schedulerCache.AddPodGroup(ss);
    for (auto& q : test.queues) schedulerCache.AddQueue(q.first, q.second);
    auto ssn = framework::OpenSession(&schedulerCache, test.tiers);
    try {
      for (auto& name : test.actions) {                                   // scheduler.go:88-92: every action on the same session
        auto action = name == "backfill" ? actions::backfill::New() : actions::allocate::New();
        action->Execute(ssn.get());
      }
    } catch (const std::exception& e) {
      std::cerr << "case " << i << " (" << test.name << "): Execute failed loudly: " << e.what() << "\n";
      framework::CloseSession(ssn.get());
      return strstr(e.what(), "KB_E_CUDA") ? 3 : 1;
    }
    framework::CloseSession(ssn.get());
    if (test.expected != binder->Binds) {
      ++failed;
      std::cerr << "case " << i << " (" << test.name << "): expected {";
      for (auto& kv : test.expected) std::cerr << kv.first << ":" << kv.second << " ";
      std::cerr << "} got {";
      for (auto& kv : binder->Binds) std::cerr << kv.first << ":" << kv.second << " ";
      std::cerr << "}\n";
    } else {
      std::cout << "ok   case " << i << " (" << test.name << ")\n";
    }
  }
  return failed ? 1 : 0;
}
