"""NavierStokesSimulator -- the reference's grid fluid solver surface (src/physics/navier_stokes.py:6-173)
on hand-written sm_100a kernels behind libsmoke_sm100.so.

Same constructor, attributes (grid_size, dt, viscosity, device, h, w, u, v, p, density, boundary) and
methods (setup_grid, add_smoke_source, diffusion_step, advection_step, interpolate_velocity_u/v,
bilinear_interpolate, pressure_projection, step) as the reference class, so callers written against
`src.physics.navier_stokes` run unchanged.  Keyword-only extensions (reference defaults):
`jacobi_iters=20` (the literal at navier_stokes.py:139), `batch=1` (independent simulations; fields gain
a leading dimension when batch > 1), `sweeps_per_launch=0` (temporal-blocking depth, 0 = library default),
`step_kernel="auto"` ("fused": whole simulation on one SM, all steps of a call in one launch, grids up to
128 x 128; "phases": one kernel per phase; "auto": fused when the grid qualifies and the call is run_steps(n >= 2)
or carries at least 32 simulations -- bit-identical results either way).

There is no CPU compute path.  `device='cpu'` (benchmark.py:260 passes it) only means "hand results back
as CPU tensors": the step always runs on the current CUDA device.

Host PyTorch is used for device memory and streams only; all arithmetic on the path is in csrc/*.cu.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import Grid, Params, Source, State


SOURCE_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("radius", "<i4"), ("intensity", "<f4")])     # smk_source_t


def _round4(n):
    return (int(n) + 3) & ~3


class FieldLayout:
    """Where each field lives inside the single fp32 arena (pure host arithmetic, testable without a GPU).

    Every field is [batch][rows][pitch] with pitch = cols rounded up to 4 elements so each row starts
    16-byte aligned (float4 access); u, v, density and p have two copies (the step ping-pongs them).
    """
    FIELDS = ("u0", "u1", "v0", "v1", "d0", "d1", "p0", "p1", "div", "boundary")

    def __init__(self, h, w, batch=1, row0=0, gh=0):
        self.h, self.w, self.batch = int(h), int(w), int(batch)
        self.row0, self.gh = int(row0), int(gh)            # row-slab placement inside a gh-row grid (0 = not a slab)
        if self.h < 1 or self.w < 1 or self.batch < 1:
            raise ValueError("grid_size and batch must be positive, got %r x %r, batch %r" % (h, w, batch))
        self.pitch_u = _round4(self.w)
        self.pitch_v = _round4(self.w + 1)
        self.pitch_c = _round4(self.w)
        self.stride_u = (self.h + 1) * self.pitch_u
        self.stride_v = self.h * self.pitch_v
        self.stride_c = self.h * self.pitch_c
        self.offset = {}
        off = 0
        for name in self.FIELDS:
            self.offset[name] = off
            off += self.batch * self.stride_of(name)
        self.total = off

    def kind(self, name):
        return name[0] if name[0] in "uv" else "c"

    def stride_of(self, name):
        return {"u": self.stride_u, "v": self.stride_v, "c": self.stride_c}[self.kind(name)]

    def shape_of(self, name):
        """(rows, cols, pitch) of a field."""
        k = self.kind(name)
        if k == "u":
            return self.h + 1, self.w, self.pitch_u
        if k == "v":
            return self.h, self.w + 1, self.pitch_v
        return self.h, self.w, self.pitch_c

    def grid_struct(self):
        return Grid(self.h, self.w, self.batch, self.pitch_u, self.pitch_v, self.pitch_c,
                    self.stride_u, self.stride_v, self.stride_c, self.row0, self.gh)


def resolve_devices(device):
    """-> (compute cuda device, device results are returned on).  Raises when no GPU is visible."""
    out = torch.device(device) if not isinstance(device, torch.device) else device
    if not torch.cuda.is_available():
        raise RuntimeError("smokephysai_b200 needs a CUDA device (B200, sm_100a): the smoke step has no CPU fallback")
    if out.type == "cuda":
        idx = out.index if out.index is not None else torch.cuda.current_device()
        return torch.device("cuda", idx), torch.device("cuda", idx)
    return torch.device("cuda", torch.cuda.current_device()), out


class NavierStokesSimulator(nn.Module):
    """Simplified Navier-Stokes smoke solver (reference: navier_stokes.py:6)."""

    def __init__(self, grid_size=(128, 128), dt=0.01, viscosity=0.001, device="cuda", *,
                 jacobi_iters=20, batch=1, sweeps_per_launch=0, step_kernel="auto", _slab=None):
        super().__init__()
        self.grid_size = grid_size
        self.dt = dt
        self.viscosity = viscosity
        self.device = device
        self.h, self.w = grid_size
        self.jacobi_iters = int(jacobi_iters)
        self.batch = int(batch)
        self.sweeps_per_launch = int(sweeps_per_launch)
        if step_kernel not in _lib.STEP_KERNELS:
            raise ValueError("step_kernel must be one of %s, got %r" % (sorted(_lib.STEP_KERNELS), step_kernel))
        self.step_kernel = step_kernel
        self._cuda, self._out_device = resolve_devices(device)
        self._lib = _lib.load()
        row0, gh = _slab if _slab is not None else (0, 0)   # (global row of local row 0, global rows): slab.py
        self._layout = FieldLayout(self.h, self.w, self.batch, row0, gh)
        self._grid = self._layout.grid_struct()
        self._arena = None
        self._state = State()
        self._mirrors = {}
        self.setup_grid()

    # ------------------------------------------------------------------ state (navier_stokes.py:24-35)
    def setup_grid(self):
        """Reset u, v, p, density (and the unused boundary mask) to zero; drops the pressure warm start."""
        L = self._layout
        if self._arena is None:
            self._arena = torch.zeros(L.total, dtype=torch.float32, device=self._cuda)
            base = self._arena.data_ptr()
            st = self._state
            for k, (a, b) in {"u": ("u0", "u1"), "v": ("v0", "v1"), "d": ("d0", "d1"), "p": ("p0", "p1")}.items():
                arr = getattr(st, k)
                arr[0] = base + 4 * L.offset[a]
                arr[1] = base + 4 * L.offset[b]
            st.div = base + 4 * L.offset["div"]
        else:
            self._mirrors.clear()
            self._arena.zero_()
        self._state.cur_u = self._state.cur_v = self._state.cur_d = self._state.cur_p = 0

    def _field(self, name):
        """Strided view [batch, rows, cols] (or [rows, cols] when batch == 1) of a field in the arena."""
        L = self._layout
        rows, cols, pitch = L.shape_of(name)
        off = L.offset[name]
        v = self._arena[off: off + L.batch * rows * pitch].view(L.batch, rows, pitch)[:, :, :cols]
        return v[0] if L.batch == 1 else v

    def _live(self, k):
        cur = getattr(self._state, "cur_" + k)
        return self._field("%s%d" % (k, cur))

    def _get(self, k):
        v = self._live(k)
        if self._out_device.type == "cuda":
            return v
        # device='cpu' (benchmark.py:260): the caller gets a host copy.  The reference hands out its LIVE tensor, so
        # in-place edits (`sim.density[mask] += x`) must not be lost: remember the copy with its version counter and
        # write it back before the next launch if it was modified (_flush_mirrors).
        self._flush_mirrors()               # an earlier copy of this field may carry edits that are not on the device yet
        t = self._live(k).to(self._out_device)
        self._mirrors[k] = (t, t._version)
        return t

    def _flush_mirrors(self):
        """Write host copies that were edited in place since they were handed out back to the device state."""
        if not self._mirrors:
            return
        for k, (t, ver) in self._mirrors.items():
            if t._version != ver:
                self._live(k).copy_(t)
        self._mirrors.clear()

    def _set(self, k, value):
        mirror = self._mirrors.get(k)
        if mirror is not None and value is mirror[0]:        # `sim.density *= c` on the host copy re-assigns it
            self._mirrors.pop(k)
            self._live(k).copy_(value)
            return
        self._mirrors.pop(k, None)
        dst = self._live(k)
        if not torch.is_tensor(value):
            value = torch.as_tensor(value, dtype=torch.float32)
        if value.device == dst.device and value.data_ptr() == dst.data_ptr() and value.stride() == dst.stride():
            return                                      # `sim.density *= c` re-assigns the view to itself
        if tuple(value.shape) != tuple(dst.shape):
            raise ValueError("cannot assign a tensor of shape %s to a field of shape %s" % (tuple(value.shape), tuple(dst.shape)))
        dst.copy_(value)

    u = property(lambda self: self._get("u"), lambda self, t: self._set("u", t), doc="x-velocity, [h+1, w]")
    v = property(lambda self: self._get("v"), lambda self, t: self._set("v", t), doc="y-velocity, [h, w+1]")
    p = property(lambda self: self._get("p"), lambda self, t: self._set("p", t), doc="pressure, [h, w]")
    density = property(lambda self: self._get("d"), lambda self, t: self._set("d", t), doc="smoke density, [h, w]")

    @property
    def boundary(self):
        b = self._field("boundary")
        return b if self._out_device.type == "cuda" else b.to(self._out_device)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        """The stream to launch on.  Also the one place every launch passes through: host copies edited in place are
        written back here, and the library's (statically linked) CUDA runtime is pointed at this simulator's device."""
        self._flush_mirrors()
        return _lib.stream_on(self._cuda)

    def _params(self):
        dt, nu = float(self.dt), float(self.viscosity)
        # c = float32(dt*viscosity): the Python-double product of navier_stokes.py:72; density uses viscosity*0.1 (:160)
        return Params(dt, dt * nu, dt * (nu * 0.1), 0.995, int(self.jacobi_iters), int(self.sweeps_per_launch),
                      _lib.STEP_KERNELS[self.step_kernel])

    def step_is_fused(self, nsteps=1):
        """True when a call of nsteps steps (step(): 1, run_steps(n): n) takes the single-launch on-chip path."""
        flag = C.c_int32(0)
        prm = self._params()
        _lib.call("smk_step_is_fused", C.byref(self._grid), C.byref(prm), int(nsteps), C.byref(flag))
        return bool(flag.value)

    def _to_out(self, t):
        return t if self._out_device.type == "cuda" else t.to(self._out_device)

    def _stage(self, t):
        """Any [.., rows, cols] tensor -> zero-padded contiguous cuda fp32 [B, rows, pitch]; returns (buf, B, rows, cols, pitch)."""
        t = torch.as_tensor(t)
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if t.dim() != 3:
            raise ValueError("expected a 2-D field (or a [batch, rows, cols] stack), got shape %s" % (tuple(t.shape),))
        B, rows, cols = t.shape
        pitch = _round4(cols)
        buf = torch.zeros(B, rows, pitch, dtype=torch.float32, device=self._cuda)
        buf[:, :, :cols] = t.to(device=self._cuda, dtype=torch.float32)
        return buf, B, rows, cols, pitch

    def _unstage(self, buf, cols, like):
        out = buf[:, :, :cols]
        if torch.as_tensor(like).dim() == 2:
            out = out[0]
        return self._to_out(out.contiguous())

    # ------------------------------------------------------------------ a2 (navier_stokes.py:37-48)
    def add_smoke_source(self, x, y, radius=10, intensity=1.0, *, batch_index=None):
        """Add a Gaussian smoke source centred on column x, row y (to every simulation unless batch_index is given)."""
        targets = range(self.batch) if batch_index is None else [int(batch_index)]
        per_sim = [[] for _ in range(self.batch)]
        for b in targets:
            per_sim[b].append((int(x), int(y), int(radius), float(intensity)))
        self.add_sources(per_sim)

    def upload_sources(self, per_sim, pin=False):
        """Pack per-simulation emitter lists into (sources, offsets) device tensors for splat_uploaded().

        per_sim[b] is the ordered list [(x, y, radius, intensity), ...] of simulation b.  The records travel
        as one H2D copy of 16 B per emitter plus 4 B per simulation."""
        if len(per_sim) != self.batch:
            raise ValueError("need one emitter list per simulation (%d), got %d" % (self.batch, len(per_sim)))
        flat, offs = [], [0]
        for lst in per_sim:
            flat.extend(lst)
            offs.append(len(flat))
        rec = np.zeros(max(len(flat), 1), dtype=SOURCE_DTYPE)          # smk_source_t records, 16 B each
        if flat:
            arr = np.asarray(flat, dtype=np.float64).reshape(len(flat), 4)
            rec["x"], rec["y"], rec["radius"] = arr[:, 0].astype(np.int32), arr[:, 1].astype(np.int32), arr[:, 2].astype(np.int32)
            rec["intensity"] = arr[:, 3].astype(np.float32)
        src_h = torch.from_numpy(rec.view(np.uint8))
        off_h = torch.tensor(offs, dtype=torch.int32)
        if pin:
            # one cached pinned staging buffer (cudaHostAlloc per call costs more than the copy); the previous
            # call's H2D copy has long been consumed by the splat that followed it on the same stream
            need = src_h.numel() + 4 * off_h.numel()
            stage = getattr(self, "_pinned_sources", None)
            if stage is None or stage.numel() < need:
                stage = torch.empty(max(need, 4096), dtype=torch.uint8).pin_memory()
                self._pinned_sources = stage
            else:
                torch.cuda.current_stream(self._cuda).synchronize()
            stage[:src_h.numel()].copy_(src_h)
            stage[src_h.numel():need].view(torch.int32).copy_(off_h)
            dev = stage[:need].to(self._cuda, non_blocking=True)
            return dev[:src_h.numel()], dev[src_h.numel():need].view(torch.int32), need
        return src_h.to(self._cuda), off_h.to(self._cuda), src_h.numel() + 4 * off_h.numel()

    def splat_uploaded(self, src, off):
        _lib.call("smk_splat_sources", C.byref(self._grid), self._ptr("d"), src.data_ptr(), off.data_ptr(), self._stream())

    def add_sources(self, per_sim):
        """Batched splat: per_sim[b] is the ordered emitter list [(x, y, radius, intensity), ...] of simulation b."""
        if not any(len(l) for l in per_sim):
            if len(per_sim) != self.batch:
                raise ValueError("need one emitter list per simulation (%d), got %d" % (self.batch, len(per_sim)))
            return
        src, off, _ = self.upload_sources(per_sim)
        self.splat_uploaded(src, off)

    # ------------------------------------------------------------------ a4 (navier_stokes.py:50-72)
    def diffusion_step(self, field, viscosity):
        buf, B, rows, cols, pitch = self._stage(field)
        out = torch.zeros_like(buf)
        _lib.call("smk_diffuse", buf.data_ptr(), out.data_ptr(), rows, cols, pitch, B, rows * pitch,
                  float(self.dt) * float(viscosity), self._stream())
        return self._unstage(out, cols, field)

    # ------------------------------------------------------------------ a10 (navier_stokes.py:74-95)
    def advection_step(self, field, u, v):
        fb, B, rows, cols, pitch = self._stage(field)
        ub, Bu, hu, wu, pu = self._stage(u)
        vb, Bv, hv, wv, pv = self._stage(v)
        if Bu != B or Bv != B:
            raise ValueError("field, u and v must have the same batch size")
        # the staggered pair defines the cell grid: u [h+1, w], v [h, w+1] (navier_stokes.py:27-28).  The reference clamps
        # to whatever shapes it is given; the kernels index u and v by that grid, so anything else is refused, not guessed.
        if hu != hv + 1 or wv != wu + 1:
            raise ValueError("advection_step needs u of shape [h+1, w] and v of shape [h, w+1]; got u %s, v %s"
                             % ((hu, wu), (hv, wv)))
        if rows > hv + 1 or cols > wu + 1:
            raise ValueError("advection_step: a %d x %d field does not fit the %d x %d cell grid of u, v" % (rows, cols, hv, wu))
        g = Grid(hv, wu, B, pu, pv, _round4(wu), hu * pu, hv * pv, hv * _round4(wu))
        out = torch.zeros_like(fb)
        _lib.call("smk_advect", C.byref(g), fb.data_ptr(), out.data_ptr(), rows, cols, pitch, rows * pitch,
                  ub.data_ptr(), vb.data_ptr(), float(self.dt), 1.0, None, 0, None, self._stream())
        return self._unstage(out, cols, field)

    # ------------------------------------------------------------------ a8, a9 (navier_stokes.py:97-131)
    def _bilerp(self, field, y, x, mode):
        fb, B, rows, cols, pitch = self._stage(field)
        if B != 1:
            raise ValueError("bilinear_interpolate takes one 2-D field")
        y = torch.as_tensor(y)
        x = torch.as_tensor(x)
        shape = torch.broadcast_shapes(y.shape, x.shape)
        yc = y.to(device=self._cuda, dtype=torch.float32).expand(shape).contiguous()
        xc = x.to(device=self._cuda, dtype=torch.float32).expand(shape).contiguous()
        out = torch.empty(shape, dtype=torch.float32, device=self._cuda)
        _lib.call("smk_bilerp", fb.data_ptr(), rows, cols, pitch, yc.data_ptr(), xc.data_ptr(), out.data_ptr(),
                  out.numel(), mode, self._stream())
        return self._to_out(out)

    def interpolate_velocity_u(self, u, y, x):
        return self._bilerp(u, y, x, 1)

    def interpolate_velocity_v(self, v, y, x):
        return self._bilerp(v, y, x, 2)

    def bilinear_interpolate(self, field, y, x):
        return self._bilerp(field, y, x, 0)

    # ------------------------------------------------------------------ a5-a7 (navier_stokes.py:133-149)
    def _ptr(self, k, cur=None):
        cur = getattr(self._state, "cur_" + k) if cur is None else cur
        return getattr(self._state, k)[cur]

    def pressure_projection(self):
        """Divergence, jacobi_iters Jacobi sweeps (warm-started p, zero ring), gradient subtract -- in place."""
        st, s, g = self._state, self._stream(), C.byref(self._grid)
        _lib.call("smk_divergence", g, self._ptr("u"), self._ptr("v"), st.div, float(self.dt), s)
        flag = C.c_int32(0)
        _lib.call("smk_jacobi", g, st.div, self._ptr("p"), self._ptr("p", st.cur_p ^ 1), int(self.jacobi_iters),
                  int(self.sweeps_per_launch), C.byref(flag), s)
        st.cur_p ^= flag.value
        _lib.call("smk_project", g, self._ptr("p"), self._ptr("u"), self._ptr("v"), float(self.dt), s)

    # ------------------------------------------------------------------ a3-a11 (navier_stokes.py:151-173)
    def _new_frames(self, nsteps=None):
        L = self._layout
        shape = (L.batch, L.h, L.pitch_c) if nsteps is None else (L.batch, nsteps, L.h, L.pitch_c)
        return torch.empty(shape, dtype=torch.float32, device=self._cuda)

    def _finish_frames(self, fr):
        L = self._layout
        if L.pitch_c != L.w:
            fr = fr[..., :L.w].contiguous()
        if L.batch == 1:
            fr = fr[0]
        return self._to_out(fr)

    def step(self, fmul=None):
        """One time step; returns a copy of the new density ([h, w], or [batch, h, w]).

        fmul (optional, [h, pitch_c] device tensor) fuses the fractal multiply of SmokeSimulator.simulate_step
        into the returned copy: frame = d + fmul*d (fractal_generator.py:62); solver state is not perturbed.
        """
        fr = self._new_frames()
        prm = self._params()
        _lib.call("smk_step", C.byref(self._grid), C.byref(self._state), C.byref(prm), fr.data_ptr(),
                  self._layout.h * self._layout.pitch_c, fmul.data_ptr() if fmul is not None else None, self._stream())
        return self._finish_frames(fr)

    def step_into(self, frame, fmul=None):
        """One time step whose returned copy is written into `frame`, a contiguous fp32 [batch, h, pitch_c]
        device tensor owned by the caller (no allocation, no sync)."""
        L = self._layout
        if (tuple(frame.shape) != (L.batch, L.h, L.pitch_c) or not frame.is_contiguous()
                or frame.dtype != torch.float32 or frame.device != self._cuda):
            raise ValueError("frame must be a contiguous fp32 [%d, %d, %d] tensor on %s" % (L.batch, L.h, L.pitch_c, self._cuda))
        prm = self._params()
        _lib.call("smk_step", C.byref(self._grid), C.byref(self._state), C.byref(prm), frame.data_ptr(),
                  L.h * L.pitch_c, fmul.data_ptr() if fmul is not None else None, self._stream())

    def run_steps(self, nsteps, fmul=None, return_frames=True, out=None):
        """nsteps consecutive steps enqueued back to back; returns frames [batch, nsteps, h, w] ([nsteps, h, w] if batch == 1).

        out: optional preallocated [batch, nsteps, h, pitch_c] fp32 device buffer to write the frames into."""
        L = self._layout
        if out is not None and (tuple(out.shape) != (L.batch, nsteps, L.h, L.pitch_c) or not out.is_contiguous()
                                or out.dtype != torch.float32 or out.device != self._cuda):
            raise ValueError("out must be a contiguous fp32 [%d, %d, %d, %d] tensor on %s" % (L.batch, nsteps, L.h, L.pitch_c, self._cuda))
        fr = out if out is not None else (self._new_frames(nsteps) if return_frames else None)
        prm = self._params()
        _lib.call("smk_run_steps", C.byref(self._grid), C.byref(self._state), C.byref(prm), int(nsteps),
                  fr.data_ptr() if fr is not None else None, L.h * L.pitch_c, nsteps * L.h * L.pitch_c,
                  fmul.data_ptr() if fmul is not None else None, self._stream())
        return self._finish_frames(fr) if fr is not None else None

    def run_steps_time_major(self, nsteps, out, fmul=None):
        """nsteps consecutive steps whose frames go to `out`, a contiguous fp32 [nsteps, batch, h, pitch_c] device
        buffer owned by the caller (time-major: the frames of one step are one contiguous block)."""
        L = self._layout
        if (tuple(out.shape) != (nsteps, L.batch, L.h, L.pitch_c) or not out.is_contiguous()
                or out.dtype != torch.float32 or out.device != self._cuda):
            raise ValueError("out must be a contiguous fp32 [%d, %d, %d, %d] tensor on %s" % (nsteps, L.batch, L.h, L.pitch_c, self._cuda))
        prm = self._params()
        _lib.call("smk_run_steps", C.byref(self._grid), C.byref(self._state), C.byref(prm), int(nsteps), out.data_ptr(),
                  L.batch * L.h * L.pitch_c, L.h * L.pitch_c, fmul.data_ptr() if fmul is not None else None, self._stream())

    def divergence_norms(self):
        """(max|div|, ||div||_2) of the un-normalised divergence of the live (u, v), per simulation: [batch, 2] on the host."""
        out = torch.zeros(self.batch, 2, dtype=torch.float32, device=self._cuda)
        _lib.call("smk_div_norms", C.byref(self._grid), self._ptr("u"), self._ptr("v"), out.data_ptr(), self._stream())
        out = out.cpu().double()
        out[:, 1] = out[:, 1].sqrt()
        return out

    def jacobi_residual_norms(self):
        """(max|p' - p|, ||p' - p||_2) per simulation, p' = one more Jacobi sweep (navier_stokes.py:139-145) over the live
        pressure and the divergence of the last projection: [batch, 2] on the host.  Meaningful right after
        pressure_projection() or a step() of the phase-per-kernel path (the fused kernel keeps the divergence in registers)."""
        out = torch.zeros(self.batch, 2, dtype=torch.float32, device=self._cuda)
        _lib.call("smk_jacobi_residual", C.byref(self._grid), self._state.div, self._ptr("p"), out.data_ptr(), self._stream())
        out = out.cpu().double()
        out[:, 1] = out[:, 1].sqrt()
        return out

    def forward(self):
        return self.step()


# every method that launches runs with the simulator's device current and restores the caller's afterwards
for _name in ("setup_grid", "add_sources", "splat_uploaded", "upload_sources", "diffusion_step", "advection_step", "_bilerp",
              "pressure_projection", "step", "step_into", "run_steps", "run_steps_time_major", "divergence_norms",
              "jacobi_residual_norms", "step_is_fused"):
    setattr(NavierStokesSimulator, _name, _lib.scoped(getattr(NavierStokesSimulator, _name)))
del _name
