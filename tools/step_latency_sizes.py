"""GPU time per step of one simulation (batch 1) on the fused and the phase-per-kernel path, n-step calls:
    python tools/step_latency_sizes.py            (SMK_PDL=0 for the phase path without programmatic dependent launch)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smokephysai_b200 import NavierStokesSimulator

for size, K in ((128, 20), (128, 40), (96, 20), (64, 20)):
    for batch in (1, 8, 32):
        row = []
        for kernel in ("fused", "phases"):
            ns = NavierStokesSimulator((size, size), device="cuda", jacobi_iters=K, batch=batch, step_kernel=kernel)
            ns.add_sources([[(size // 2, size // 2, 8, 1.5)]] * batch)
            for n in (1, 20):
                for _ in range(3):
                    ns.run_steps(n)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 20
                e0.record()
                for _ in range(reps):
                    ns.run_steps(n)
                e1.record()
                torch.cuda.synchronize()
                row.append("%s n=%-2d %6.1f us/step" % (kernel, n, 1e3 * e0.elapsed_time(e1) / reps / n))
        print("%dx%d K=%d batch %-2d: %s" % (size, size, K, batch, " | ".join(row)), flush=True)
