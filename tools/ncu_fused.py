"""Small driver for ncu captures of k_step_fused: python tools/ncu_fused.py [batch] [K] [nsteps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smokephysai_b200 import NavierStokesSimulator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ns = NavierStokesSimulator((128, 128), 0.01, 0.001, "cuda", jacobi_iters=K, batch=B, step_kernel="fused")
rng = np.random.default_rng(0)
ems = []
for b in range(B):
    ems.append([(int(rng.integers(20, 108)), int(rng.integers(20, 108)), 8, float(rng.uniform(0.5, 2.0))) for _ in range(2)])
ns.add_sources(ems)
frames = torch.empty(B, n, 128, 128, device="cuda")
for _ in range(3):
    ns.run_steps(n, out=frames)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ns.run_steps(n, out=frames)
e1.record()
torch.cuda.synchronize()
print("batch %d K %d nsteps %d: %.1f us per launch, %.2f us per simulation-step per SM-round" % (
    B, K, n, 1e3 * e0.elapsed_time(e1), 1e3 * e0.elapsed_time(e1) / n / ((B + 147) // 148)))
