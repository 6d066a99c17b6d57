"""torchrun --nproc-per-node N tests/multigpu_check.py  — sharded-node-axis parity on real GPUs (NCCL)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
from kube_batch_b200 import engine, synth
from kube_batch_b200.snapshot import PluginConf
from oracle import kbo
import util

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid_t = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid_t.copy_(torch.frombuffer(bytearray(engine.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(uid_t, 0)
eng = engine.Engine(device=local, rank=rank, world_size=world, nccl_unique_id=bytes(uid_t.cpu().numpy().tobytes()))
cases = {
    "c2": lambda: synth.make("c2"),
    "multi-tile/multi-queue": lambda: (synth.generate(synth.SynthSpec("mr", tasks=700, jobs=70, nodes=1100, queues=3, hetero_job_frac=0.3,
                                                                      prio_levels=2, min_member_frac=0.5, seed=4242)), PluginConf.default()),
    "few-tiles": lambda: (synth.random_session(5, tasks=80, jobs=9, nodes=40, queues=2), PluginConf.default()),
}
if len(sys.argv) > 1 and sys.argv[1] == "c3":
    cases["c3"] = lambda: synth.make("c3")
ok = True
for name, mk in cases.items():
    snap, conf = mk()
    eng.load(snap, conf)
    r = eng.allocate()
    ref = kbo.allocate(snap, conf)
    try:
        util.assert_same_decisions(ref.decisions, r.decisions, f"{name} rank{rank}")
        util.assert_same_state(ref, eng.node_state(), eng.order_state(), f"{name} rank{rank}")
        msg = "OK"
    except AssertionError as e:
        msg = "FAIL " + str(e)[:300]; ok = False
    t = [eng.allocate().stats.gpu_ms for _ in range(3)]
    print(f"[rank {rank}/{world}] {name}: {msg}; gpu_ms {min(t):.2f}; launches {r.stats.kernel_launches} scans {r.stats.scans} exchange_mode {r.stats.exchange_mode}", flush=True)
dist.barrier()
eng.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
