"""Put the UNMODIFIED reference where the GPU box can see it: baseline/_ref (git-ignored, shipped by gpurun).

    python baseline/install_ref.py [--reference /root/reference]

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the reference>` installs its one
   package, `src` (models, utils, evaluation, physics), from a copy under /tmp because the build writes into the source tree and
   /root/reference is read-only;
2. the reference's scripts (train.py, inference.py, benchmark.py) and config/config.yaml -- which setup.py does not install --
   are copied next to it, so tests/test_gpu_c5.py can run them unchanged through smokephysai_b200.run on the GPU box,
   where /root/reference does not exist.
Nothing under baseline/_ref is product code or tracked; it is the reference arm's install.  No-op when the reference is absent.
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SCRIPTS = ("train.py", "inference.py", "benchmark.py")


def install(reference="/root/reference", force=False):
    if not os.path.isdir(os.path.join(reference, "src")):
        return False
    if force or not os.path.isdir(os.path.join(DEST, "src", "physics")):
        with tempfile.TemporaryDirectory() as tmp:
            work = os.path.join(tmp, "reference")
            shutil.copytree(reference, work, ignore=shutil.ignore_patterns("inference_output", "__pycache__"))
            cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
                   "--find-links", "/opt/wheelhouse", "--target", DEST, "--upgrade", work]
            subprocess.check_call(cmd)
    os.makedirs(os.path.join(DEST, "config"), exist_ok=True)
    for name in SCRIPTS:
        shutil.copyfile(os.path.join(reference, name), os.path.join(DEST, name))
    shutil.copyfile(os.path.join(reference, "config", "config.yaml"), os.path.join(DEST, "config", "config.yaml"))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print("installed" if install(a.reference, a.force) else "reference not found at %s: nothing to do" % a.reference)
