"""ctypes/numpy binding of oracle/libsmoke_oracle.so (the C restatement of the reference step).

TEST INFRASTRUCTURE ONLY: the checker for the CUDA path, never the product and never a fallback.
Each wrapper cites the reference function it restates (paths relative to the reference root).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsmoke_oracle.so")
_lib = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc, -ffp-contract=off)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "smoke_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsmoke_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        i, f, l, v = C.c_int, C.c_float, C.c_long, C.c_void_p
        sig = {
            "smk_oracle_bilerp": [f32p, i, i, f32p, f32p, f32p, l],
            "smk_oracle_interp_u": [f32p, i, i, i, i, f32p],
            "smk_oracle_interp_v": [f32p, i, i, i, i, f32p],
            "smk_oracle_advect": [f32p, f32p, i, i, f32p, f32p, i, i, f],
            "smk_oracle_advect_slab": [f32p, f32p, i, i, f32p, f32p, i, i, f, i, i],
            "smk_oracle_diffuse": [f32p, f32p, i, i, f],
            "smk_oracle_buoyancy": [f32p, f32p, i, i, f],
            "smk_oracle_divergence": [f32p, f32p, f32p, i, i, f],
            "smk_oracle_jacobi": [f32p, f32p, f32p, i, i, i],
            "smk_oracle_grad_subtract": [f32p, f32p, f32p, i, i, f],
            "smk_oracle_div_norms": [f32p, f32p, i, i, f64p, f64p],
            "smk_oracle_step": [f32p, f32p, f32p, f32p, i, i, f, f, f, i, f32p, v, v],
            "smk_oracle_run_batch": [f32p, f32p, f32p, f32p, i, i, i, f, f, f, i, i, v, v, i],
            "smk_oracle_splat": [f32p, i, i, i, i, i, f],
            "smk_oracle_perlin": [f32p, i, i, f32p, f32p],
            "smk_oracle_mandelbrot": [f32p, i, i, f32p, f32p, i, i],
            "smk_oracle_fractal_mul": [f32p, i, f, f32p, f32p, f32p, i],
            "smk_oracle_apply_mul": [f32p, f32p, f32p, l],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = None
        L.smk_oracle_work_floats.argtypes = [i, i]
        L.smk_oracle_work_floats.restype = l
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def coef(dt, viscosity):
    """c = float32(dt*viscosity): the Python-double product at navier_stokes.py:72."""
    return np.float32(float(dt) * float(viscosity))


def bilinear_interpolate(field, y, x):
    """navier_stokes.py:111-131"""
    field, y, x = _c(field), _c(y), _c(x)
    out = np.empty_like(x)
    lib().smk_oracle_bilerp(field, field.shape[0], field.shape[1], y.ravel(), x.ravel(), out.reshape(-1), x.size)
    return out


def diffusion_step(field, dt, viscosity):
    """navier_stokes.py:50-72"""
    field = _c(field)
    out = np.empty_like(field)
    lib().smk_oracle_diffuse(field, out, field.shape[0], field.shape[1], coef(dt, viscosity))
    return out


def advection_step(field, u, v, dt):
    """navier_stokes.py:74-95"""
    field, u, v = _c(field), _c(u), _c(v)
    h, w = v.shape[0], u.shape[1]
    out = np.empty_like(field)
    lib().smk_oracle_advect(field, out, field.shape[0], field.shape[1], u, v, h, w, np.float32(dt))
    return out


def advection_step_slab(field, u, v, dt, row0, gh):
    """Extension (no reference counterpart): advection of a row slab whose local row 0 is global row row0 of a
    gh-row grid; equals advection_step of the whole grid on the rows the slab holds exactly."""
    field, u, v = _c(field), _c(u), _c(v)
    h, w = v.shape[0], u.shape[1]
    out = np.empty_like(field)
    lib().smk_oracle_advect_slab(field, out, field.shape[0], field.shape[1], u, v, h, w, np.float32(dt), int(row0), int(gh))
    return out


def buoyancy(v, d, dt):
    """navier_stokes.py:154-155 (returns the new v)"""
    v, d = _c(v).copy(), _c(d)
    lib().smk_oracle_buoyancy(v, d, d.shape[0], d.shape[1], np.float32(dt))
    return v


def grad_subtract(u, v, p, dt):
    """navier_stokes.py:148-149 (returns new u, v)"""
    u, v, p = _c(u).copy(), _c(v).copy(), _c(p)
    lib().smk_oracle_grad_subtract(u, v, p, p.shape[0], p.shape[1], np.float32(dt))
    return u, v


def interpolate_velocity(u, v, rows, cols):
    """navier_stokes.py:97-109 on the integer grid rows x cols"""
    u, v = _c(u), _c(v)
    h, w = v.shape[0], u.shape[1]
    ou = np.empty((rows, cols), np.float32)
    ov = np.empty((rows, cols), np.float32)
    lib().smk_oracle_interp_u(u, h, w, rows, cols, ou)
    lib().smk_oracle_interp_v(v, h, w, rows, cols, ov)
    return ou, ov


def divergence(u, v, dt):
    """navier_stokes.py:136"""
    u, v = _c(u), _c(v)
    h, w = v.shape[0], u.shape[1]
    out = np.empty((h, w), np.float32)
    lib().smk_oracle_divergence(u, v, out, h, w, np.float32(dt))
    return out


def jacobi(p, div, K):
    """navier_stokes.py:139-145"""
    p = _c(p).copy()
    div = _c(div)
    tmp = np.empty_like(p)
    lib().smk_oracle_jacobi(p, div, tmp, p.shape[0], p.shape[1], int(K))
    return p


def pressure_projection(u, v, p, dt, K=20):
    """navier_stokes.py:133-149 -> (u, v, p)"""
    u, v = _c(u).copy(), _c(v).copy()
    div = divergence(u, v, dt)
    p = jacobi(p, div, K)
    lib().smk_oracle_grad_subtract(u, v, p, p.shape[0], p.shape[1], np.float32(dt))
    return u, v, p


def div_norms(u, v):
    u, v = _c(u), _c(v)
    h, w = v.shape[0], u.shape[1]
    mx, l2 = np.zeros(1), np.zeros(1)
    lib().smk_oracle_div_norms(u, v, h, w, mx, l2)
    return float(mx[0]), float(l2[0])


def splat(density, x, y, radius, intensity):
    """navier_stokes.py:37-48 (in place)"""
    assert density.dtype == np.float32 and density.flags.c_contiguous
    lib().smk_oracle_splat(density, density.shape[0], density.shape[1], int(x), int(y), int(radius), np.float32(intensity))
    return density


def host_linspaces(n):
    """The three torch.linspace grids of fractal_generator.py:17-18,:38-39, made by torch on the host CPU
    exactly as the reference makes them (ATen's vectorised linspace is ISA-dependent in the last ulp)."""
    import torch
    return tuple(torch.linspace(a, b, n, device="cpu").numpy().copy()
                 for a, b in ((0.0, 10.0), (-2.5, 1.5), (-1.5, 1.5)))


def perlin(n, px=None):
    """fractal_generator.py:12-31 (square n x n)"""
    px = _c(host_linspaces(n)[0] if px is None else px)
    out = np.empty((n, n), np.float32)
    lib().smk_oracle_perlin(out, n, n, px, px)
    return out


def mandelbrot_count(n, mx=None, my=None, iters=100, tail_mod=0):
    """fractal_generator.py:33-51 escape counts (before the /iterations)"""
    if mx is None:
        _, mx, my = host_linspaces(n)
    out = np.empty((n, n), np.float32)
    lib().smk_oracle_mandelbrot(out, n, n, _c(mx), _c(my), iters, tail_mod)
    return out


def fractal_mul(n, intensity=0.05, grids=None, tail_mod=0):
    """intensity * (0.7*perlin + 0.3*mandelbrot): fractal_generator.py:55-62"""
    px, mx, my = host_linspaces(n) if grids is None else grids
    out = np.empty((n, n), np.float32)
    lib().smk_oracle_fractal_mul(out, n, np.float32(intensity), _c(px), _c(mx), _c(my), tail_mod)
    return out


def apply_mul(field, mul):
    field, mul = _c(field), _c(mul)
    out = np.empty_like(field)
    lib().smk_oracle_apply_mul(field.reshape(-1), mul.reshape(-1), out.reshape(-1), field.size)
    return out


class OracleSolver:
    """Dense-array twin of NavierStokesSimulator (navier_stokes.py:6-173) with a jacobi_iters knob."""

    def __init__(self, grid_size=(128, 128), dt=0.01, viscosity=0.001, jacobi_iters=20):
        self.h, self.w = int(grid_size[0]), int(grid_size[1])
        self.dt, self.viscosity, self.jacobi_iters = dt, viscosity, int(jacobi_iters)
        self.setup_grid()

    def setup_grid(self):
        h, w = self.h, self.w
        self.u = np.zeros((h + 1, w), np.float32)
        self.v = np.zeros((h, w + 1), np.float32)
        self.p = np.zeros((h, w), np.float32)
        self.density = np.zeros((h, w), np.float32)
        self._work = np.empty(lib().smk_oracle_work_floats(h, w), np.float32)
        self.last_div_norms = None

    def add_smoke_source(self, x, y, radius=10, intensity=1.0):
        splat(self.density, x, y, radius, intensity)

    def step(self, want_norms=False):
        for k in ("u", "v", "p", "density"):
            setattr(self, k, _c(getattr(self, k)))
        frame = np.empty((self.h, self.w), np.float32)
        norms = np.zeros(2)
        lib().smk_oracle_step(self.u, self.v, self.p, self.density, self.h, self.w,
                              np.float32(self.dt), coef(self.dt, self.viscosity),
                              np.float32(float(self.dt) * (float(self.viscosity) * 0.1)),
                              self.jacobi_iters, self._work, frame.ctypes.data,
                              norms.ctypes.data if want_norms else None)
        if want_norms:
            self.last_div_norms = (float(norms[0]), float(norms[1]))
        return frame


def run_batch(u, v, p, d, dt, viscosity, K, nsteps, fmul=None, want_frames=True, nthreads=1):
    """Advance B independent states [B,...] in place; returns frames [B,nsteps,h,w] (x(1+mul) if fmul)."""
    B, h, w = d.shape
    for a in (u, v, p, d):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    frames = np.empty((B, nsteps, h, w), np.float32) if want_frames else None
    if fmul is not None:
        fmul = _c(fmul)
    lib().smk_oracle_run_batch(u, v, p, d, B, h, w, np.float32(dt), coef(dt, viscosity),
                               np.float32(float(dt) * (float(viscosity) * 0.1)), int(K), int(nsteps),
                               fmul.ctypes.data if fmul is not None else None,
                               frames.ctypes.data if frames is not None else None, int(nthreads))
    return frames
