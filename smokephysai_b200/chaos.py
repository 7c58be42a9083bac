"""Chaos features of SmokeSimulator (reference: src/physics/smoke_simulator.py:47-140) from the device
kernels in csrc/features.cu: box counts, [0,1] histogram and consecutive-frame distances for many frames per
launch; the few scalars left (polyfit over 5 points, entropy of 256 counts, mean of 18 log-ratios) are
finished on the host exactly as the reference does."""
import numpy as np
import torch

from . import _lib

SCALES = (2, 4, 8, 16, 32)           # smoke_simulator.py:98
NBINS = 256                          # :134
_edges = {}


def _edges_for(dev):
    """torch.linspace(0, 1, 257): the bin edges torch.histogram builds on the host (:134)."""
    if dev not in _edges:
        _edges[dev] = torch.linspace(0, 1, NBINS + 1, dtype=torch.float32, device="cpu").to(dev)
    return _edges[dev]


def _stream(dev):
    return _lib.stream_on(dev)


def frame_counts(frames, w):
    """frames: contiguous fp32 [n, h, pitch] device tensor -> (box_counts [n, 5], hist [n, 256]) int32 on the device."""
    n, h, pitch = frames.shape
    dev = frames.device
    box = torch.empty(n, len(SCALES), dtype=torch.int32, device=dev)
    hist = torch.empty(n, NBINS, dtype=torch.int32, device=dev)
    with _lib.on_device(dev):
        _lib.call("smk_frame_features", frames.data_ptr(), h * pitch, n, h, w, pitch, _edges_for(dev).data_ptr(), NBINS, 0.0, 1.0,
                  box.data_ptr(), hist.data_ptr(), None, _stream(dev))
    return box, hist


def frame_distances(frames, w):
    """||frames[k+1] - frames[k]||_2 for consecutive frames, float64 numpy [n-1] (values rounded through fp32 like torch.norm)."""
    n, h, pitch = frames.shape
    if n < 2:
        return np.zeros(0)
    out = torch.empty(n - 1, dtype=torch.float64, device=frames.device)
    with _lib.on_device(frames.device):
        _lib.call("smk_frame_distances", frames.data_ptr(), h * pitch, n, h, w, pitch, out.data_ptr(), _stream(frames.device))
    return out.sqrt().float().cpu().numpy().astype(np.float64)


def fractal_dimension_from_counts(counts):
    """abs(slope) of log(count + 1) against log(scale) (smoke_simulator.py:117-124)."""
    log_scales = np.log(SCALES)
    log_counts = np.log(np.asarray(counts, dtype=np.int64) + 1)
    return abs(np.polyfit(log_scales, log_counts, 1)[0])


def entropy_from_hist(hist_row):
    """-sum(p * log2(p + 1e-8)) of the normalised histogram (smoke_simulator.py:135-140), evaluated on the host in fp32 like the reference."""
    h = torch.as_tensor(hist_row).cpu().float()
    probs = h / h.sum()
    return (-torch.sum(probs * torch.log2(probs + 1e-8))).item()


def lyapunov_from_distances(distances):
    """max(0, mean(diff(log(d + 1e-8)))) over the given consecutive-frame distances (smoke_simulator.py:81-87)."""
    d = np.asarray(distances, dtype=np.float64)
    if len(d) > 1:
        return max(0, np.mean(np.diff(np.log(d + 1e-8))))
    return 0.0
