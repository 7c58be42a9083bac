"""chaos features on the device (csrc/features.cu), the batched dataset back-end
(SmokeSimulator.generate_dataset == data_loader.py:37-99) and the launcher's import merge."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from helpers import assert_same, nerr

pytestmark = pytest.mark.gpu

from smokephysai_b200 import SmokeSimulator, chaos  # noqa: E402


def ref_box_counts(frame):
    """smoke_simulator.py:89-115 restated with numpy (the reference's python double loop, any() per box)."""
    binary = frame > frame.mean(dtype=np.float64).astype(np.float32)
    h, w = binary.shape
    out = []
    for s in chaos.SCALES:
        bh, bw = h // s, w // s
        out.append(int(binary[:bh * s, :bw * s].reshape(bh, s, bw, s).any(axis=(1, 3)).sum()) if bh and bw else 0)
    return out


@pytest.mark.parametrize("h,w", [(128, 128), (100, 75), (33, 260), (512, 384), (16, 16)])
def test_frame_counts_vs_host(h, w):
    rng = np.random.default_rng(h + w)
    n = 5
    frames = (rng.random((n, h, w)) ** 3 * 1.6).astype(np.float32)           # values in [0, 1.6): some outside the histogram range
    frames[0, :, : w // 2] = 0.0
    frames[1] = 1.0                                                            # right edge of the last bin
    pitch = (w + 3) & ~3
    dev = torch.zeros(n, h, pitch, device="cuda")
    dev[:, :, :w] = torch.from_numpy(frames).cuda()
    box, hist = chaos.frame_counts(dev, w)
    box, hist = box.cpu().numpy(), hist.cpu().numpy()
    for k in range(n):
        assert list(box[k]) == ref_box_counts(frames[k]), "frame %d box counts" % k
        want = torch.histogram(torch.from_numpy(frames[k]).flatten(), bins=256, range=(0, 1)).hist.numpy()
        assert np.array_equal(hist[k], want.astype(np.int64)), "frame %d histogram" % k
    d = chaos.frame_distances(dev, w)
    want = np.array([np.linalg.norm((frames[k + 1] - frames[k]).astype(np.float64)) for k in range(n - 1)])
    assert np.allclose(d, want, rtol=2e-7)


def reference_style_loop(sim, num_samples, T):
    """The reference's per-sample loop (data_loader.py:37-99) through the scalar class surface."""
    data = []
    h, w = sim.ns_solver.h, sim.ns_solver.w
    for _ in range(num_samples):
        sim.ns_solver.setup_grid()
        n_src = np.random.randint(1, 4)
        pos, inten = [], []
        for _ in range(n_src):
            x = np.random.randint(20, w - 20); y = np.random.randint(20, h - 20)
            inten.append(np.random.uniform(0.5, 2.0)); pos.append((x, y))
        sim.add_incense_source(pos, inten)
        seq, feats = [], []
        for t in range(T):
            seq.append(sim.simulate_step().clone())
            if t >= 10:
                f = sim.get_chaos_features()
                if f:
                    feats.append(f)
        avg = {k: np.mean([f[k] for f in feats]) for k in ("lyapunov_exponent", "fractal_dimension", "entropy")}
        data.append({"sequence": torch.stack(seq), "chaos_features": avg, "source_config": {"positions": pos, "intensities": inten}})
    return data


@pytest.mark.parametrize("batch", [4, 3])
def test_generate_dataset_matches_per_sample_loop(batch):
    N, T = 7, 20
    np.random.seed(123)
    want = reference_style_loop(SmokeSimulator((128, 128), device="cuda"), N, T)
    np.random.seed(123)
    bat = SmokeSimulator((128, 128), device="cuda", batch=batch)
    got = bat.generate_dataset(N, T)
    assert len(got) == N
    for k in range(N):
        assert got[k]["source_config"] == want[k]["source_config"]
        assert_same(got[k]["sequence"].cpu().numpy(), want[k]["sequence"].cpu().numpy(), "sequence %d" % k)
        for name in ("lyapunov_exponent", "fractal_dimension", "entropy"):
            assert abs(got[k]["chaos_features"][name] - want[k]["chaos_features"][name]) < 1e-9, (k, name)
    assert len(bat.history) == 100


def test_generate_dataset_vs_unmodified_reference():
    """The reference's own SyntheticSmokeDataset (installed copy in baseline/_ref, CPU device) against the batched
    GPU back-end with the same numpy seed: same emitters, frames within the expf/sinf tolerance, features close."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "src", "utils")):
        pytest.skip("reference not installed in baseline/_ref")
    import importlib.util

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_root, "src", *rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    # build a private package tree "refpkg" = the reference's src, so its relative imports resolve to ITS physics
    import types
    for pkg in ("refpkg", "refpkg.physics", "refpkg.utils"):
        m = types.ModuleType(pkg); m.__path__ = [os.path.join(ref_root, "src", *pkg.split(".")[1:])]; sys.modules[pkg] = m
    load("refpkg.physics.fractal_generator", ("physics", "fractal_generator.py"))
    load("refpkg.physics.navier_stokes", ("physics", "navier_stokes.py"))
    load("refpkg.physics.smoke_simulator", ("physics", "smoke_simulator.py"))
    dl = load("refpkg.utils.data_loader", ("utils", "data_loader.py"))
    N = 2
    np.random.seed(7)
    ref = dl.SyntheticSmokeDataset(num_samples=N, grid_size=(128, 128), sequence_length=20, device="cpu").data
    np.random.seed(7)
    got = SmokeSimulator((128, 128), device="cuda", batch=N).generate_dataset(N, 20)
    for k in range(N):
        assert got[k]["source_config"]["positions"] == ref[k]["source_config"]["positions"]
        assert nerr(got[k]["sequence"].cpu().numpy(), ref[k]["sequence"].numpy()) < 1e-6
        for name, tol in (("lyapunov_exponent", 1e-4), ("fractal_dimension", 1e-9), ("entropy", 1e-4)):
            assert abs(got[k]["chaos_features"][name] - ref[k]["chaos_features"][name]) < tol, (k, name)


def test_launcher_import_merge():
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "src", "models")):
        pytest.skip("reference not installed in baseline/_ref")
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); from smokephysai_b200.run import merge_src, patch_batched_generation; "
            "merge_src(%r); patch_batched_generation(8); "
            "from src.utils.data_loader import SyntheticSmokeDataset; import numpy as np; np.random.seed(1); "
            "ds = SyntheticSmokeDataset(num_samples=5, grid_size=(128, 128), device='cuda'); it = ds[0]; "
            "from src.physics.smoke_simulator import SmokeSimulator as S; "
            "print(len(ds), tuple(it['input'].shape), tuple(it['sequence'].shape), S.__module__)" % (ROOT, ref_root))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("5 (1, 128, 128) (20, 128, 128) smokephysai_b200.smoke_simulator")
