"""Sweep (tile, T) of smk_jacobi on a grid and print cell-sweeps/s; run on the GPU box:
    python tools/tune_jacobi.py 1024 1024 100
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smokephysai_b200 import NavierStokesSimulator, _lib  # noqa: E402

h, w, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ns = NavierStokesSimulator((h, w), device="cuda", jacobi_iters=K, batch=batch)
ns._field("div").copy_(torch.randn(ns._field("div").shape, device="cuda"))
st = ns._state
flag = C.c_int32(0)
for tile in (1, 2, 3):
    os.environ["SMK_JACOBI_TILE"] = str(tile)
    for T in (2, 4, 6, 8, 10, 12, 16, 20, 24):
        if tile != 2 and T > 12:
            continue
        def run():
            _lib.call("smk_jacobi", C.byref(ns._grid), st.div, st.p[0], st.p[1], K, T, C.byref(flag), ns._stream())
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("tile %d T %2d: %8.1f us per %d sweeps, %7.1f G cell-sweeps/s" % (tile, T, ms * 1e3, K, batch * h * w * K / ms / 1e6), flush=True)
