"""Per-call latency of the scalar API (batch = 1, the reference's own usage): python tools/step_latency.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smokephysai_b200 import SmokeSimulator

for kernel in ("fused", "phases"):
    sim = SmokeSimulator((128, 128), device="cuda", step_kernel=kernel)
    sim.add_incense_source([(64, 100), (32, 110), (96, 110)], [1.5, 1.0, 1.0])
    for _ in range(20):
        sim.simulate_step()
    torch.cuda.synchronize()
    n = 500
    t0 = time.perf_counter()
    for _ in range(n):
        f = sim.simulate_step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ns = sim.ns_solver
    t3 = time.perf_counter()
    for _ in range(n):
        ns.step()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    fr = torch.empty(1, 128, 128, device="cuda")
    t5 = time.perf_counter()
    for _ in range(n):
        ns.step_into(fr)
    torch.cuda.synchronize()
    t6 = time.perf_counter()
    print("%-7s simulate_step: host %.1f us/call, with drain %.1f us/call | ns.step() %.1f us | ns.step_into() %.1f us" % (
        kernel, 1e6 * (t1 - t0) / n, 1e6 * (t2 - t0) / n, 1e6 * (t4 - t3) / n, 1e6 * (t6 - t5) / n))
