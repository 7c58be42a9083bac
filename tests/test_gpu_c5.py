"""BASELINE config 5 / north_star: "the simulator keeps its Python class/method surface so train.py, inference.py and
benchmark.py use it unchanged".  The reference's own scripts, byte for byte as installed in baseline/_ref by
baseline/install_ref.py, run to completion on this simulator through the launcher (smokephysai_b200.run).

Also settles the DataLoader question of SURVEY.md s7 on the same box: create_data_loaders (data_loader.py:126-182) asks
for num_workers=os.cpu_count() and pin_memory=True over CUDA-tensor samples whenever the device is a GPU.  The test runs
the UNMODIFIED reference with ITS OWN simulator first and records what happens; the launcher's --dataloader-workers 0 is
the documented remedy either way (a launcher flag, not an edit of the script).
"""
import glob
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

TINY = """
data: {grid_size: [128, 128], sequence_length: 20, num_train: 8, num_val: 4, cache_dir: "./cache"}
model: {input_dim: 128, hidden_dim: 512, num_layers: 6, num_heads: 8, output_channels: 64, chaos_strength: 0.1}
physics: {conservation_weight: 1.0, continuity_weight: 1.0, energy_weight: 0.5}
training: {batch_size: 4, num_epochs: 1, learning_rate: 0.001, weight_decay: 0.01}
simulation: {dt: 0.01, viscosity: 0.001, grid_size: [128, 128]}
"""


def run(cmd, cwd, env=None, timeout=900):
    e = dict(os.environ)
    e["PYTHONDONTWRITEBYTECODE"] = "1"
    e.update(env or {})
    p = subprocess.run(cmd, cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    return p.returncode, p.stdout.decode(errors="replace")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not os.path.exists(os.path.join(REF, "train.py")):
        pytest.skip("baseline/_ref holds no reference scripts (run baseline/install_ref.py where /root/reference exists)")
    d = tmp_path_factory.mktemp("c5")
    (d / "tiny.yaml").write_text(TINY)
    return str(d)


LAUNCH = [sys.executable, "-m", "smokephysai_b200.run", "--reference", REF]


def test_train_then_benchmark_then_inference_unchanged_on_this_simulator(workdir):
    env = {"PYTHONPATH": ROOT}
    # train.py:182-280 -- data generation by the GPU simulator (batched back-end), one epoch, best checkpoint
    rc, out = run(LAUNCH + ["--batched-generation", "8", "--dataloader-workers", "0", os.path.join(REF, "train.py"), "--config", "tiny.yaml"],
                  workdir, env)
    assert rc == 0, out[-3000:]
    assert "Training completed!" in out
    ckpts = glob.glob(os.path.join(workdir, "experiments", "*", "best_model.pth"))
    assert len(ckpts) == 1, out[-2000:]
    assert os.path.exists(os.path.join(workdir, "cache", "train_data.pkl"))       # data_loader.py:33-34 pickled our frames
    import pickle
    with open(os.path.join(workdir, "cache", "train_data.pkl"), "rb") as f:
        data = pickle.load(f)
    assert len(data) == 8 and tuple(data[0]["sequence"].shape) == (20, 128, 128)
    assert set(data[0]["chaos_features"]) == {"lyapunov_exponent", "fractal_dimension", "entropy"}

    # benchmark.py:236-278 -- passes device='cpu' to the dataset (:257-261): results come back as CPU tensors, the step runs on the GPU
    rc, out = run(LAUNCH + [os.path.join(REF, "benchmark.py"), "--config", "tiny.yaml", "--checkpoint", ckpts[0], "--num_samples", "4"],
                  workdir, env)
    assert rc == 0, out[-3000:]
    assert "SmokePhysAI" in out and "Farneback" in out and "Lucas-Kanade" in out

    # inference.py:111-145 -- the per-step facade path (simulate_step x 20 on fixed emitters); matplotlib is not in this image
    rc, out = run(LAUNCH + ["--stub-plotting", os.path.join(REF, "inference.py"), "--config", "tiny.yaml", "--checkpoint", ckpts[0]],
                  workdir, env)
    assert rc == 0, out[-3000:]
    assert "Visualization results have been saved" in out


def test_train_per_sample_generation_path(workdir, tmp_path):
    """The same script without --batched-generation: SyntheticSmokeDataset's own Python loop (data_loader.py:37-99) drives
    setup_grid / add_incense_source / simulate_step / get_chaos_features 20 x N times on this simulator."""
    (tmp_path / "tiny.yaml").write_text(TINY.replace("num_train: 8", "num_train: 3").replace("num_val: 4", "num_val: 2"))
    rc, out = run(LAUNCH + ["--dataloader-workers", "0", os.path.join(REF, "train.py"), "--config", "tiny.yaml"], str(tmp_path), {"PYTHONPATH": ROOT})
    assert rc == 0, out[-3000:]
    assert "Training completed!" in out


def test_unmodified_reference_dataloader_hazard_is_its_own(workdir, tmp_path, record_property):
    """The reference alone (its simulator, its DataLoader settings) on this GPU: documents whether the forked-worker /
    pin_memory failure of create_data_loaders pre-exists.  Whatever the outcome, it is recorded; the assertion is only that
    a failure, if any, is the DataLoader one and not something about the simulator."""
    (tmp_path / "tiny.yaml").write_text(TINY.replace("num_train: 8", "num_train: 2").replace("num_val: 4", "num_val: 2").replace("batch_size: 4", "batch_size: 2"))
    rc, out = run([sys.executable, os.path.join(REF, "train.py"), "--config", "tiny.yaml"], str(tmp_path), {"PYTHONPATH": REF}, timeout=1200)
    record_property("unmodified_reference_train_rc", rc)
    tail = out[-1500:]
    print("unmodified reference train.py on cuda: rc=%d\n%s" % (rc, tail))
    with open(os.path.join(ROOT, "gpurun_out", "c5_unmodified_reference.log") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else os.devnull, "w") as f:
        f.write("rc=%d\n%s" % (rc, out[-6000:]))
    if rc != 0:
        assert ("forked subprocess" in out or "pin" in out.lower() or "DataLoader worker" in out or "CUDA" in out), tail
