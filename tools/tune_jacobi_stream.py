"""Time smk_jacobi with the two-CTAs-per-SM kernel and the streaming kernel (both staging engines); run on the GPU box:
    python tools/tune_jacobi_stream.py 8192 8192 20
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
quick = os.environ.get("QUICK") == "1"
for name, tile, stream, Ts in ((("stream tma", 2, 2, (10,)), ("stream tma x2", 1, 2, (10,)), ("64x128 x2", 1, 0, (10,))) if quick else
                               (("64x128 x2", 1, 0, (10,)), ("128x128", 2, 0, (10,)), ("stream ldgsts", 2, 1, (10,)),
                                ("stream tma", 2, 2, (5, 8, 10, 12)), ("stream tma x2", 1, 2, (5, 8, 10)), ("stream ldgsts x2", 1, 1, (10,)))):
    os.environ["SMK_JACOBI_TILE"] = str(tile)
    os.environ["SMK_JACOBI_STREAM"] = str(stream)
    _lib.reload_env()                      # the switches are read once per process
    for T in Ts:
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
        print("%dx%d K %d  %-16s T %2d: %8.1f us, %7.1f G cell-sweeps/s" % (h, w, K, name, T, ms * 1e3, batch * h * w * K / ms / 1e6), flush=True)
