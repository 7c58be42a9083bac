"""Host-side cost of a slab step: torchrun --nproc-per-node 2 tools/slab_host_time.py [n]  (small n => GPU time is negligible)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from smokephysai_b200.slab import SlabNavierStokes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
halo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
slab = SlabNavierStokes((n, n), 0.01, 0.001, torch.device("cuda", local), rank=rank, world=world, jacobi_iters=20, sweeps_per_launch=10,
                        halo=halo or None)
slab.add_sources([(n // 2, n // 2, 8, 1.5)])
for _ in range(5):
    slab.step()
torch.cuda.synchronize()
dist.barrier()
for label, fn in (("whole step", slab.step),
                  ("exchange(u,v,d,p) only", lambda: slab.exchanger.exchange(slab.geom, slab.exchange_list(("u", "v", "d", "p")))),
                  ("exchange(u,v,d) only", lambda: slab.exchanger.exchange(slab.geom, slab.exchange_list(("u", "v", "d")))),
                  ("exchange(p) only", lambda: slab.exchanger.exchange(slab.geom, slab.exchange_list(("p",)))),
                  ("compute phases only", lambda: [arg() for kind, arg in slab.step_plan() if kind == "c"])):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if rank == 0:
        print("%-24s host enqueue %7.1f us/call, with drain %7.1f us/call" % (label, 1e6 * (t1 - t0) / 50, 1e6 * (t2 - t0) / 50), flush=True)
dist.destroy_process_group()
