"""SmokeSimulator -- the reference's simulator facade (src/physics/smoke_simulator.py:8-140) on the
sm_100a solver: same constructor, attributes (ns_solver, fractal_gen, device, history, max_history) and
methods (add_incense_source, simulate_step, get_chaos_features, compute_*).

simulate_step() is ONE pass of the CUDA step with the fractal multiply fused into the returned copy
(the reference: ns.step() + ~960 eager launches of fractal recompute).
Keyword-only extensions: jacobi_iters, batch, sweeps_per_launch (forwarded to NavierStokesSimulator), and
generate_sequences() -- the batched back-end of data_loader.py:37-99.
"""
import numpy as np
import torch
import torch.nn as nn

from .fractal_generator import FractalGenerator
from .navier_stokes import NavierStokesSimulator


class SmokeSimulator(nn.Module):
    """Complete smoke physics simulation system (reference: smoke_simulator.py:8)."""

    def __init__(self, grid_size=(128, 128), dt=0.01, viscosity=0.001, device="cuda", *,
                 jacobi_iters=20, batch=1, sweeps_per_launch=0):
        super().__init__()
        self.ns_solver = NavierStokesSimulator(grid_size, dt, viscosity, device, jacobi_iters=jacobi_iters,
                                               batch=batch, sweeps_per_launch=sweeps_per_launch)
        self.fractal_gen = FractalGenerator(device)
        self.device = device
        self.history = []          # smoke_simulator.py:23
        self.max_history = 100     # :24

    def add_incense_source(self, positions, intensities):
        """Add incense smoke sources: radius is fixed to 8 (smoke_simulator.py:26-29)."""
        for (x, y), intensity in zip(positions, intensities):
            self.ns_solver.add_smoke_source(x, y, radius=8, intensity=intensity)

    def simulate_step(self, add_fractal=True):
        """One simulation step; returns the (fractal-scaled) density frame (smoke_simulator.py:31-45)."""
        ns = self.ns_solver
        fmul = self.fractal_gen.multiplier((ns.h, ns.w), 0.05) if add_fractal else None     # intensity=0.05 (:38)
        density = ns.step(fmul=fmul)
        self.history.append(density.clone())                                               # :41
        if len(self.history) > self.max_history:
            self.history.pop(0)                                                            # :42-43
        return density

    # -------------------------------------------------------------- batched generation (data_loader.py:37-99)
    def generate_sequences(self, emitters, sequence_length=20, add_fractal=True, to_host=False):
        """Reset, splat one emitter list per simulation, run sequence_length steps.

        emitters[b] = [((x, y), intensity), ...] for simulation b (len == batch).  Returns frames
        [batch, sequence_length, h, w] -- each [b] is what the reference's per-sample loop stacks into
        sample['sequence'] (data_loader.py:66-68, :91).  History/chaos features are not touched.

        to_host=True hands the frames back in pinned host memory (the dataset is pickled from the host,
        data_loader.py:33-34): the frames of step t are copied device->host on a second stream while step
        t+1 computes, from a time-major [sequence_length, batch, h, w] device buffer so every copy is one
        contiguous block.  The returned tensor is the [batch, sequence_length, h, w] view of the pinned
        time-major buffer, which this simulator owns and reuses on the next call."""
        ns = self.ns_solver
        L = ns._layout
        B, T = ns.batch, int(sequence_length)
        ns.setup_grid()
        src, off, _ = ns.upload_sources([[(x, y, 8, inten) for (x, y), inten in lst] for lst in emitters], pin=to_host)
        ns.splat_uploaded(src, off)
        fmul = self.fractal_gen.multiplier((ns.h, ns.w), 0.05) if add_fractal else None
        if not to_host:
            fr = ns.run_steps(T, fmul=fmul)
            return fr if B > 1 else fr.unsqueeze(0)
        key = (T,)
        if getattr(self, "_gen_key", None) != key:
            self._gen_dev = torch.empty(T, B, L.h, L.pitch_c, dtype=torch.float32, device=ns._cuda)
            self._gen_host = torch.empty(T, B, L.h, L.pitch_c, dtype=torch.float32).pin_memory()
            self._gen_copy_stream = torch.cuda.Stream(device=ns._cuda)
            self._gen_key = key
        dev, host, cs = self._gen_dev, self._gen_host, self._gen_copy_stream
        main = torch.cuda.current_stream(ns._cuda)
        cs.wait_stream(main)                      # the previous call's consumer is done with the buffers
        for t in range(T):
            ns.step_into(dev[t], fmul=fmul)
            ev = torch.cuda.Event()
            ev.record(main)
            cs.wait_event(ev)
            with torch.cuda.stream(cs):
                host[t].copy_(dev[t], non_blocking=True)
        cs.synchronize()
        return host.permute(1, 0, 2, 3)[..., :L.w]

    # -------------------------------------------------------------- chaos features (smoke_simulator.py:47-140)
    def get_chaos_features(self):
        if len(self.history) < 10:
            return {}
        return {
            "lyapunov_exponent": self.compute_lyapunov_exponent(),
            "fractal_dimension": self.compute_fractal_dimension(),
            "entropy": self.compute_entropy(),
        }

    def compute_lyapunov_exponent(self):
        """Mean log-growth of the distance between consecutive frames over the last 20 (smoke_simulator.py:67-87)."""
        if len(self.history) < 20:
            return 0.0
        states = torch.stack(self.history[-20:])
        diffs = states[1:] - states[:-1]
        distances = torch.linalg.vector_norm(diffs.reshape(diffs.shape[0], -1), dim=1).cpu().numpy().astype(np.float64)
        if len(distances) > 1:
            log_distances = np.log(distances + 1e-8)
            return max(0, np.mean(np.diff(log_distances)))
        return 0.0

    def compute_fractal_dimension(self):
        """Box-counting dimension of the above-mean mask at scales 2..32 (smoke_simulator.py:89-124)."""
        if not self.history:
            return 0.0
        current = self.history[-1]
        binary = current > current.mean()
        scales = [2, 4, 8, 16, 32]
        counts = []
        h, w = binary.shape[-2:]
        for s in scales:
            bh, bw = h // s, w // s
            if bh == 0 or bw == 0:
                counts.append(0)
                continue
            boxes = binary[..., :bh * s, :bw * s].reshape(bh, s, bw, s)
            counts.append(int(boxes.any(dim=3).any(dim=1).sum().item()))
        log_scales = np.log(scales)
        log_counts = np.log(np.array(counts) + 1)
        return abs(np.polyfit(log_scales, log_counts, 1)[0])

    def compute_entropy(self):
        """Shannon entropy of the 256-bin histogram of the last frame over [0, 1] (smoke_simulator.py:126-140)."""
        if not self.history:
            return 0.0
        cur = self.history[-1].detach().cpu()
        hist = torch.histogram(cur.flatten(), bins=256, range=(0, 1))
        probs = hist.hist.float() / hist.hist.sum()
        return (-torch.sum(probs * torch.log2(probs + 1e-8))).item()

    def forward(self, add_fractal=True):
        return self.simulate_step(add_fractal)
