"""Time smk_jacobi alone: python tools/bench_jacobi.py H W K [batch] [T]  (env SMK_JACOBI_PACKED / SMK_JACOBI_TILE apply)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smokephysai_b200 import NavierStokesSimulator, _lib  # noqa: E402

h, w, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
T = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ns = NavierStokesSimulator((h, w), device="cuda", jacobi_iters=K, batch=batch)
ns._field("div").copy_(torch.randn(ns._field("div").shape, device="cuda"))
st = ns._state
flag = C.c_int32(0)


def run():
    _lib.call("smk_jacobi", C.byref(ns._grid), st.div, st.p[0], st.p[1], K, T, C.byref(flag), ns._stream())


for _ in range(5):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 50
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("packed=%s %dx%d batch %d K %d T %d: %8.1f us, %7.1f G cell-sweeps/s" % (
    os.environ.get("SMK_JACOBI_PACKED", "0"), h, w, batch, K, T, ms * 1e3, batch * h * w * K / ms / 1e6), flush=True)
