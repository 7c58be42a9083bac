"""FractalGenerator -- the reference's static fractal multiplier (src/physics/fractal_generator.py:5-62)
evaluated by one sm_100a kernel (csrc/stencil.cu: k_fractal_fields).

The reference recomputes the field from scratch on every simulate_step although it depends only on the
grid shape (960 eager launches and 200 host syncs per step); here it is computed once per
(shape, scale, iterations) and cached.  The torch.linspace grids are made on the host CPU with torch,
exactly as the reference makes them: ATen's CPU linspace is vectorised and ISA-dependent in the last ulp,
and the Mandelbrot escape count is sensitive to that ulp at a handful of pixels.
"""
import torch
import torch.nn as nn

from . import _lib
from .navier_stokes import _round4, resolve_devices


class FractalGenerator(nn.Module):
    """Fractal geometry generator (reference: fractal_generator.py:5)."""

    def __init__(self, device="cuda"):
        super().__init__()
        self.device = device
        self._cuda, self._out_device = resolve_devices(device)
        self._cache = {}

    def _stream(self):
        return _lib.stream_on(self._cuda)

    @_lib.scoped
    def _fields(self, shape, scale=10.0, iterations=100, intensity=0.0, want=("perlin",)):
        """Launch k_fractal_fields for grid shape (h, w); outputs are laid out (w, h) like the reference's meshgrid('ij')."""
        h, w = int(shape[0]), int(shape[1])
        key = (h, w, float(scale), int(iterations), float(intensity), tuple(want))
        if key in self._cache:
            return self._cache[key]
        pitch = _round4(h)
        dev = self._cuda
        px = torch.linspace(0, scale, w, device="cpu").to(dev)          # fractal_generator.py:17
        py = torch.linspace(0, scale, h, device="cpu").to(dev)          # :18
        mx = torch.linspace(-2.5, 1.5, w, device="cpu").to(dev)         # :38
        my = torch.linspace(-1.5, 1.5, h, device="cpu").to(dev)         # :39
        outs = {k: torch.zeros(w, pitch, dtype=torch.float32, device=dev) for k in want}
        ptr = lambda k: outs[k].data_ptr() if k in outs else None
        _lib.call("smk_fractal_fields", ptr("perlin"), ptr("mandel"), ptr("mul"), w, h, pitch, float(intensity),
                  int(iterations), px.data_ptr(), py.data_ptr(), mx.data_ptr(), my.data_ptr(), self._stream())
        self._cache[key] = outs
        return outs

    def _ret(self, t, h):
        t = t[:, :h]
        return t.clone() if self._out_device.type == "cuda" else t.to(self._out_device)

    def generate_perlin_noise(self, shape, scale=10.0):
        """Six sin*cos octaves normalised to [0, 1]; shape (w, h) as in the reference (fractal_generator.py:12-31)."""
        return self._ret(self._fields(shape, scale=scale, want=("perlin",))["perlin"], int(shape[0]))

    def generate_mandelbrot_field(self, shape, iterations=100):
        """Mandelbrot escape count / iterations (fractal_generator.py:33-51)."""
        return self._ret(self._fields(shape, iterations=iterations, want=("mandel",))["mandel"], int(shape[0]))

    def multiplier(self, shape, intensity):
        """intensity * (0.7*perlin + 0.3*mandelbrot) as a padded [n, pitch] device tensor (square grids), cached.

        This is what NavierStokesSimulator.step(fmul=...) fuses into the returned frame."""
        h, w = int(shape[0]), int(shape[1])
        if h != w:
            # the reference builds (w, h)-shaped fields and fails to broadcast them against an (h, w) frame
            raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (%d) at non-singleton "
                               "dimension 1 (fractal field is laid out (w, h); only square grids are supported, "
                               "as in the reference)" % (w, h))
        return self._fields(shape, intensity=intensity, want=("mul",))["mul"]

    @_lib.scoped
    def apply_fractal_perturbation(self, field, intensity=0.1):
        """field + intensity*F*field (fractal_generator.py:53-62)."""
        field_t = torch.as_tensor(field)
        h, w = field_t.shape[-2:]
        mul = self.multiplier((h, w), intensity)
        lead = field_t.shape[:-2]
        pitch = _round4(w)
        buf = torch.zeros((int(torch.Size(lead).numel()) if lead else 1, h, pitch), dtype=torch.float32, device=self._cuda)
        buf[:, :, :w] = field_t.to(device=self._cuda, dtype=torch.float32).reshape(-1, h, w)
        out = torch.empty_like(buf)
        _lib.call("smk_apply_mul", buf.data_ptr(), mul.data_ptr(), out.data_ptr(), h, w, pitch, buf.shape[0], h * pitch,
                  self._stream())
        out = out[:, :, :w].reshape(*lead, h, w).contiguous()
        return out if self._out_device.type == "cuda" else out.to(self._out_device)
