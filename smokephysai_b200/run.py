"""Run the reference's own, unmodified scripts (train.py, inference.py, benchmark.py) against this simulator.

    python -m smokephysai_b200.run --reference /path/to/SmokePhysAI [--batched-generation N] [--dataloader-workers N]
                                   [--stub-plotting] train.py --config config/config.yaml

The reference hard-wires `from src.physics.smoke_simulator import SmokeSimulator` (inference.py:13,
src/utils/data_loader.py:39).  This launcher makes the `src` package a merge of this repo's `src/` (which only
holds `physics`, re-exporting smokephysai_b200) and the reference's `src/` (models, utils, evaluation):
`src.physics.*` resolves here, everything else resolves to the reference, and the script runs under runpy
with its own argv.  --batched-generation N additionally routes SyntheticSmokeDataset._generate_synthetic_data
(data_loader.py:37-99) through SmokeSimulator.generate_dataset with N simulations per launch.
"""
import argparse
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def merge_src(reference_root):
    """Make `src` = this repo's src/ first, then the reference's src/ (a directory that contains src/)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import src
    ref_src = os.path.join(os.path.abspath(reference_root), "src")
    if not os.path.isdir(ref_src):
        raise SystemExit("no src/ package under %s" % reference_root)
    if ref_src not in src.__path__:
        src.__path__.append(ref_src)
    import src.physics.smoke_simulator as ours
    if not ours.__file__.startswith(ROOT):
        raise SystemExit("src.physics resolved to %s, not to this repo" % ours.__file__)
    return src


def patch_batched_generation(batch):
    """SyntheticSmokeDataset._generate_synthetic_data -> one batched GPU generation (same sample dict format)."""
    import src.utils.data_loader as dl
    from smokephysai_b200 import SmokeSimulator

    def _generate(self):
        sim = SmokeSimulator(self.grid_size, device=self.device, batch=min(int(batch), max(1, self.num_samples)))
        return sim.generate_dataset(self.num_samples, self.sequence_length)
    dl.SyntheticSmokeDataset._generate_synthetic_data = _generate


def patch_dataloader_workers(num_workers):
    """create_data_loaders (data_loader.py:126-182) builds its DataLoaders with num_workers=os.cpu_count() and pin_memory=True
    whenever the dataset device is not 'cpu' -- over samples that are CUDA tensors.  Forked workers cannot touch CUDA tensors
    ("Cannot re-initialize CUDA in forked subprocess") and pinning only applies to host tensors, so the UNMODIFIED reference fails
    in its first training batch on any GPU, with its own simulator too (tests/test_gpu_c5.py shows it).  This launcher flag, not
    an edit of the script, overrides the two arguments where the reference looks DataLoader up."""
    import src.utils.data_loader as dl
    base = dl.DataLoader

    class DataLoader(base):
        def __init__(self, dataset, *args, **kwargs):
            kwargs["num_workers"] = int(num_workers)
            if int(num_workers) == 0:
                kwargs["pin_memory"] = False
            super().__init__(dataset, *args, **kwargs)
    dl.DataLoader = DataLoader


def stub_plotting():
    """inference.py and src/utils/visualization.py import matplotlib / seaborn at module level; where they are not installed
    (this image), stand-ins that accept every call let the script run its simulation and model inference and skip the PNGs."""
    import importlib.util
    import types
    from unittest import mock
    if importlib.util.find_spec("matplotlib") is None:
        plt = mock.MagicMock(name="matplotlib.pyplot")
        plt.subplots.side_effect = lambda *a, **k: (mock.MagicMock(name="figure"), mock.MagicMock(name="axes"))
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if importlib.util.find_spec("seaborn") is None:
        sys.modules["seaborn"] = mock.MagicMock(name="seaborn")


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="checkout (or install directory) of the reference: the directory that contains src/")
    ap.add_argument("--batched-generation", type=int, default=0, metavar="N")
    ap.add_argument("--dataloader-workers", type=int, default=None, metavar="N",
                    help="force num_workers=N (and, for 0, pin_memory=False) on the DataLoaders create_data_loaders builds")
    ap.add_argument("--stub-plotting", action="store_true", help="stand-ins for matplotlib / seaborn when they are not installed")
    ap.add_argument("script")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    merge_src(a.reference)
    if a.stub_plotting:
        stub_plotting()
    if a.batched_generation:
        patch_batched_generation(a.batched_generation)
    if a.dataloader_workers is not None:
        patch_dataloader_workers(a.dataloader_workers)
    script = a.script if os.path.exists(a.script) else os.path.join(a.reference, a.script)
    sys.argv = [script] + a.args
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
