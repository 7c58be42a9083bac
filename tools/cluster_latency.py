"""us per step of run_steps(20) on 128 x 128, K = 40, for a few batch sizes: one CTA per simulation against clusters of 2 / 4 CTAs
and against the phase kernels.   python tools/cluster_latency.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smokephysai_b200 import NavierStokesSimulator, _lib

K, T = int(os.environ.get("K", "40")), 20
print("batch  " + "  ".join("%10s" % m for m in ("phases", "fused-1", "cluster-2", "cluster-4")))
for B in (1, 4, 16, 32, 48, 64, 74, 128):
    row = []
    for kernel, nc in (("phases", None), ("fused", 0), ("fused", 2), ("fused", 4)):
        if nc is None:
            os.environ.pop("SMK_FUSED_CLUSTER", None)
        else:
            os.environ["SMK_FUSED_CLUSTER"] = str(nc)
        os.environ["SMK_FUSED_SLICE"] = "0"
        _lib.reload_env()
        ns = NavierStokesSimulator((128, 128), 0.01, 0.001, "cuda", jacobi_iters=K, batch=B, step_kernel=kernel)
        ns.add_smoke_source(64, 64, 8, 1.5)
        frames = torch.empty(B, T, 128, 128, device="cuda")
        for _ in range(3):
            ns.run_steps(T, out=frames)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            ns.run_steps(T, out=frames)
        e1.record()
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) * 1e3 / (reps * T))
        del ns
    print("%5d  " % B + "  ".join("%10.2f" % x for x in row), flush=True)
