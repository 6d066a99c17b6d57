"""The reference's own action-test sessions (allocate_test.go:51-144) as builder calls."""
from kube_batch_b200 import builder as B
from kube_batch_b200.snapshot import PluginConf, PluginOption


def tiers_allocate_test():      # allocate_test.go:180-195
    return PluginConf([[PluginOption("drf", enabled_preemptable=True, enabled_job_order=True),
                        PluginOption("proportion", enabled_queue_order=True, enabled_reclaimable=True)]])


def allocate_test_case1():
    b = B.SessionBuilder()
    b.add_pod_group(B.PodGroup("c1", "pg1", "c1"))
    b.add_pod(B.build_pod("c1", "p1", "", "Pending", B.build_resource_list("1", "1G"), "pg1"))
    b.add_pod(B.build_pod("c1", "p2", "", "Pending", B.build_resource_list("1", "1G"), "pg1"))
    b.add_node(B.build_node("n1", B.build_resource_list("2", "4Gi")))
    b.add_queue(B.Queue("c1", 1))
    return b.flatten()


def allocate_test_case2():
    b = B.SessionBuilder()
    b.add_pod_group(B.PodGroup("c1", "pg1", "c1"))
    b.add_pod_group(B.PodGroup("c2", "pg2", "c2"))
    for ns, pg in (("c1", "pg1"), ("c2", "pg2")):
        b.add_pod(B.build_pod(ns, "p1", "", "Pending", B.build_resource_list("1", "1G"), pg))
        b.add_pod(B.build_pod(ns, "p2", "", "Pending", B.build_resource_list("1", "1G"), pg))
    b.add_node(B.build_node("n1", B.build_resource_list("2", "4G")))
    b.add_queue(B.Queue("c1", 1))
    b.add_queue(B.Queue("c2", 1))
    return b.flatten()
