"""Time each phase kernel alone on a big single grid: python tools/bench_phases.py [n] [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smokephysai_b200 import NavierStokesSimulator, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ns = NavierStokesSimulator((n, n), 0.01, 0.001, "cuda", jacobi_iters=K)
rng = np.random.default_rng(0)
ems = [(int(bx * 64 + rng.integers(8, 56)), int(by * 64 + rng.integers(8, 56)), 8, float(rng.uniform(0.5, 2.0)))
       for by in range(n // 64) for bx in range(n // 64)]
ns.add_sources([ems])
for _ in range(3):
    ns.step()
torch.cuda.synchronize()
_lib.profile_begin(4096)
for _ in range(5):
    ns.step()
torch.cuda.synchronize()
prof = _lib.profile_end()
cells = n * n
for k, (ms, cnt) in prof.items():
    if cnt:
        print("%-20s %8.1f us per step (%d launches)  %6.1f G cells/s" % (k, 1e3 * ms / 5, cnt // 5, cells / (ms / 5) / 1e6))
