"""Compute-only time of ONE interior slab of an 8192 x 8192 grid on one GPU (no neighbours, ghost rows left as they are): what a
rank of the c4 bench would need per step if the halo exchange were free.   python tools/slab_alone.py [world ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smokephysai_b200 import _lib
from smokephysai_b200.slab import SlabNavierStokes

n, K, T = 8192, 20, 10
for world in [int(x) for x in sys.argv[1:]] or [8, 4, 2]:
    slab = SlabNavierStokes((n, n), 0.01, 0.001, "cuda", rank=min(1, world - 1), world=world, jacobi_iters=K, sweeps_per_launch=T,
                            halo=K + 4, exchanger=False)
    slab.add_sources([(n // 2, slab.geom.A + slab.geom.hl // 2, 40, 1.5), (n // 3, slab.geom.A + 30, 25, 1.0)])
    for _ in range(5):
        slab._c_step()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            slab._c_step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
    own = slab.geom.own_hi - slab.geom.own_lo
    _lib.profile_begin(max_records=4096)
    for _ in range(10):
        slab._c_step()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    print("   per launch-serialised phase, us per step: " + ", ".join("%s %.1f" % (k, v[0] * 100) for k, v in sorted(prof.items()) if v[1] > 0), flush=True)
    print("world %d: slab of %d stored rows (%d owned): %.1f us per step compute-only -> %.1f G cell-steps/s over %d GPUs if the exchange were free"
          % (world, slab.geom.hl, own, best, n * n / best / 1e3, world), flush=True)
    del slab
    torch.cuda.empty_cache()
