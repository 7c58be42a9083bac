"""dataset generation (data_loader.py:37-99) -- the unmodified reference's own loop against the
batched GPU back-end.  python tools/bench_dataset.py [reference samples] [our samples]"""
import importlib.util
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from smokephysai_b200 import SmokeSimulator

n_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_ours = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
out = {"grid": [128, 128], "sequence_length": 20}

ref_root = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(ref_root, "src", "utils")):
    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_root, "src", *rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    for pkg in ("refpkg", "refpkg.physics", "refpkg.utils"):
        m = types.ModuleType(pkg); m.__path__ = [os.path.join(ref_root, "src", *pkg.split(".")[1:])]; sys.modules[pkg] = m
    load("refpkg.physics.fractal_generator", ("physics", "fractal_generator.py"))
    load("refpkg.physics.navier_stokes", ("physics", "navier_stokes.py"))
    load("refpkg.physics.smoke_simulator", ("physics", "smoke_simulator.py"))
    dl = load("refpkg.utils.data_loader", ("utils", "data_loader.py"))
    for dev in ("cpu", "cuda"):
        np.random.seed(1)
        t0 = time.perf_counter()
        dl.SyntheticSmokeDataset(num_samples=n_ref, grid_size=(128, 128), sequence_length=20, device=dev)
        if dev == "cuda":
            torch.cuda.synchronize()
        out["reference_%s_s_per_sample" % dev] = (time.perf_counter() - t0) / n_ref

for batch in (64, 256):
    sim = SmokeSimulator((128, 128), device="cuda", batch=batch)
    np.random.seed(1)
    sim.generate_dataset(batch, 20)                 # warm up (fractal field, buffers)
    sim.history = []
    np.random.seed(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    data = sim.generate_dataset(n_ours, 20)
    torch.cuda.synchronize()
    out["ours_batch%d_s_per_sample" % batch] = (time.perf_counter() - t0) / n_ours
print(json.dumps(out))
