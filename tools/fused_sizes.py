"""k_step_fused on grids below 128 x 128 (the generic instantiation): python tools/fused_sizes.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smokephysai_b200 import NavierStokesSimulator

for (h, w) in ((128, 128), (127, 127), (128, 96), (96, 96), (64, 64), (32, 32)):
    for kernel in ("fused", "phases"):
        B, K, n = 148, 40, 10
        ns = NavierStokesSimulator((h, w), 0.01, 0.001, "cuda", jacobi_iters=K, batch=B, step_kernel=kernel)
        rng = np.random.default_rng(0)
        ns.add_sources([[(int(rng.integers(8, w - 8)), int(rng.integers(8, h - 8)), 6, 1.5)] for _ in range(B)])
        for _ in range(2):
            ns.run_steps(n, return_frames=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ns.run_steps(n, return_frames=False)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / n
        print("%3d x %3d %-6s %7.1f us per step of %d simulations, %6.2f G cell-steps/s" % (h, w, kernel, us, B, B * h * w / us / 1e3))
