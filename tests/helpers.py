"""Shared helpers for the parity tests."""
import contextlib
import hashlib
import os

import numpy as np


def sha(a):
    """sha256 of the fp32 bytes with -0.0 canonicalised to +0.0 (numeric bit pin, as make_golden.py)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    a = a + np.float32(0.0)
    return hashlib.sha256(a.tobytes()).hexdigest()


def nerr(ours, ref):
    """Normalised max-abs error max|a-b| / max|b| (SURVEY.md s8c tolerance definition)."""
    ours = np.asarray(ours, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).max()
    return float(np.abs(ours - ref).max() / (den if den > 0 else 1.0))


def assert_same(ours, ref, what=""):
    """Bit-level parity: numeric equality everywhere (treats -0.0 == +0.0)."""
    ours = np.asarray(ours)
    ref = np.asarray(ref)
    assert ours.shape == ref.shape, "%s shape %s vs %s" % (what, ours.shape, ref.shape)
    if not np.array_equal(ours, ref):
        bad = np.argwhere(ours != ref)
        k = tuple(bad[0])
        raise AssertionError("%s: %d/%d elements differ, first at %s: %r vs %r (nerr %.3e)" % (
            what, len(bad), ours.size, k, ours[k], ref[k], nerr(ours, ref)))


def emitters_for_sequence(s, h=128, w=128):
    """Emitter law of data_loader.py:49-58 with the fixed seeds of SURVEY.md s8d (1234 + sequence index)."""
    rng = np.random.default_rng(1234 + s)
    n = int(rng.integers(1, 4))
    out = []
    for _ in range(n):
        x = int(rng.integers(20, w - 20))
        y = int(rng.integers(20, h - 20))
        inten = float(rng.uniform(0.5, 2.0))
        out.append((x, y, 8, inten))
    return out


@contextlib.contextmanager
def smk_env(**switches):
    """Set SMK_* switches for the body and make the library re-read them (it reads the environment once per process:
    include/smoke_b200.h smk_reload_env); restored and re-read on exit.  smk_env(SMK_FUSED_SLICE=3)."""
    from smokephysai_b200 import _lib
    saved = {k: os.environ.get(k) for k in switches}
    for k, v in switches.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    _lib.reload_env()
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        _lib.reload_env()
