#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (the reference mount does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports /root/reference/src/physics on CPU (single thread, fp32) and records
inputs/outputs of every function on the hot path.
The reference has no tests and no golden vectors of its own, so
these files are what pins the oracle (oracle/smoke_oracle.c) and, through it,
the CUDA path.  Nothing here is imported by the product.

The only edit applied to the reference is the Jacobi sweep count: the literal
``range(20)`` at navier_stokes.py:139 is substituted with ``range(K)`` in a
subclass built from ``inspect.getsource`` so K=40/100 (BASELINE.json configs)
have a reference answer too.
"""
import hashlib
import inspect
import os
import re
import sys
import textwrap

import numpy as np
import torch

REF = os.environ.get("SMOKE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from src.physics.navier_stokes import NavierStokesSimulator  # noqa: E402
from src.physics.smoke_simulator import SmokeSimulator  # noqa: E402
from src.physics.fractal_generator import FractalGenerator  # noqa: E402

torch.set_num_threads(1)


def with_iters(K):
    """Reference solver class whose Jacobi loop runs K sweeps (navier_stokes.py:139)."""
    if K == 20:
        return NavierStokesSimulator
    src = textwrap.dedent(inspect.getsource(NavierStokesSimulator.pressure_projection))
    src, n = re.subn(r"range\(20\)", "range(%d)" % K, src)
    assert n == 1
    ns = {"torch": torch}
    exec(src, ns)
    return type("NavierStokesSimulatorK%d" % K, (NavierStokesSimulator,),
                {"pressure_projection": ns["pressure_projection"]})


def sha(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    a = a + np.float32(0.0)  # canonicalise -0.0 -> +0.0 so the hash is a numeric pin
    return hashlib.sha256(a.tobytes()).hexdigest()


def state(ns):
    return {k: getattr(ns, k).detach().cpu().numpy().copy() for k in ("u", "v", "p", "density")}


def div_after_projection(ns):
    """Run one step but sample the divergence right after pressure_projection (:163)."""
    box = {}
    orig = ns.pressure_projection

    def wrapped():
        orig()
        d = (ns.u[1:, :] - ns.u[:-1, :] + ns.v[:, 1:] - ns.v[:, :-1])
        box["max"] = float(d.abs().max())
        box["l2"] = float(torch.linalg.norm(d.double()))
    ns.pressure_projection = wrapped
    out = ns.step()
    ns.pressure_projection = orig
    return out, box


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("%-28s %8.1f KiB" % (name, os.path.getsize(path) / 1024))


# --------------------------------------------------------------------------- C1
def scenario_c1():
    """inference.py:40-41 emitters on the config.yaml grid; solver state only."""
    out = {}
    for K in (20, 40, 100):
        ns = with_iters(K)((128, 128), 0.01, 0.001, "cpu")
        for (x, y), inten in zip([(64, 64), (32, 32), (96, 96)], [1.5, 1.0, 0.8]):
            ns.add_smoke_source(x, y, radius=8, intensity=inten)
        if K == 20:
            out["density0"] = ns.density.numpy().copy()
        stats = []
        hashes = []
        for t in range(1, 21):
            _, box = div_after_projection(ns)
            if t in (1, 5, 10, 20):
                s = state(ns)
                stats.append([t, s["density"].astype(np.float64).sum(), s["density"].max(),
                              np.linalg.norm(s["u"].astype(np.float64)),
                              np.linalg.norm(s["v"].astype(np.float64)),
                              np.linalg.norm(s["p"].astype(np.float64)), box["max"], box["l2"]])
                hashes.append([sha(s[k]) for k in ("u", "v", "p", "density")])
        out["stats_K%d" % K] = np.array(stats, dtype=np.float64)
        out["sha_K%d" % K] = np.array(hashes)
        if K == 20:
            for k, a in state(ns).items():
                out["final_" + k] = a
    save("scenario_c1.npz", **out)


# ------------------------------------------------------------ small random grids
def small_random():
    """Random states with |dt*vel| of several cells so backtraces cross cells and hit the clamps."""
    out = {}
    cases = [("a", 24, 24, 20, 0.01, 0.001, 300.0), ("b", 17, 29, 20, 0.01, 0.001, 400.0),
             ("c", 40, 33, 7, 0.05, 0.02, 60.0), ("d", 8, 8, 40, 0.01, 0.001, 100.0),
             ("e", 3, 5, 20, 0.01, 0.001, 100.0), ("f", 132, 36, 3, 0.02, 0.003, 200.0)]
    names = []
    for tag, h, w, K, dt, nu, vel in cases:
        g = torch.Generator().manual_seed(ord(tag) * 1000 + h * 131 + w)
        ns = with_iters(K)((h, w), dt, nu, "cpu")
        ns.u = (torch.rand(h + 1, w, generator=g) - 0.5) * 2 * vel
        ns.v = (torch.rand(h, w + 1, generator=g) - 0.5) * 2 * vel
        ns.p = torch.randn(h, w, generator=g)
        ns.density = torch.rand(h, w, generator=g)
        out["%s_meta" % tag] = np.array([h, w, K, dt, nu], dtype=np.float64)
        for k, a in state(ns).items():
            out["%s_0_%s" % (tag, k)] = a
        for t in (1, 2, 3):
            ret = ns.step()
            for k, a in state(ns).items():
                out["%s_%d_%s" % (tag, t, k)] = a
            assert np.array_equal(ret.numpy(), ns.density.numpy())
        names.append(tag)
    out["cases"] = np.array(names)
    save("small_random.npz", **out)


# ---------------------------------------------------------------- unit vectors
def units():
    out = {}
    g = torch.Generator().manual_seed(7)
    ns = NavierStokesSimulator((20, 28), 0.01, 0.001, "cpu")
    h, w = 20, 28
    # a4 diffusion_step on each field shape, two coefficients
    for nm, shape, visc in (("u", (h + 1, w), 0.001), ("v", (h, w + 1), 0.3), ("d", (h, w), 0.001 * 0.1)):
        f = torch.randn(*shape, generator=g)
        out["diff_%s_in" % nm] = f.numpy().copy()
        out["diff_%s_visc" % nm] = np.float64(visc)
        out["diff_%s_out" % nm] = ns.diffusion_step(f, visc).numpy()
    # a8 bilinear_interpolate: in-range, on-node, last-index and clamped coordinates
    f = torch.randn(h, w, generator=g)
    y = torch.rand(64, generator=g) * (h - 1)
    x = torch.rand(64, generator=g) * (w - 1)
    y[:8] = torch.tensor([0., h - 1., 3., 3.5, h - 1., 0., h - 2., h - 1.5])
    x[:8] = torch.tensor([0., w - 1., w - 1., 4., 0., w - 1., w - 2., w - 1.25])
    out["bil_f"], out["bil_y"], out["bil_x"] = f.numpy().copy(), y.numpy().copy(), x.numpy().copy()
    out["bil_out"] = ns.bilinear_interpolate(f, y, x).numpy()
    # a9/a10 advection_step on each field shape
    u = (torch.rand(h + 1, w, generator=g) - 0.5) * 500
    v = (torch.rand(h, w + 1, generator=g) - 0.5) * 500
    d = torch.rand(h, w, generator=g)
    out["adv_u"], out["adv_v"], out["adv_d"] = u.numpy().copy(), v.numpy().copy(), d.numpy().copy()
    out["adv_u_out"] = ns.advection_step(u, u, v).numpy()
    out["adv_v_out"] = ns.advection_step(v, u, v).numpy()
    out["adv_d_out"] = ns.advection_step(d, u, v).numpy()
    Yd, Xd = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    out["interp_u_out"] = ns.interpolate_velocity_u(u, Yd, Xd).numpy()
    out["interp_v_out"] = ns.interpolate_velocity_v(v, Yd, Xd).numpy()
    # a5-a7 pressure_projection, K = 20 and 7, warm-started random p with non-zero ring
    for K in (20, 7):
        s = with_iters(K)((h, w), 0.01, 0.001, "cpu")
        s.u = torch.randn(h + 1, w, generator=g)
        s.v = torch.randn(h, w + 1, generator=g)
        s.p = torch.randn(h, w, generator=g)
        for k in ("u", "v", "p"):
            out["proj%d_%s_in" % (K, k)] = getattr(s, k).numpy().copy()
        out["proj%d_div" % K] = ((s.u[1:, :] - s.u[:-1, :] + s.v[:, 1:] - s.v[:, :-1]) / s.dt).numpy()
        s.pressure_projection()
        for k in ("u", "v", "p"):
            out["proj%d_%s_out" % (K, k)] = getattr(s, k).numpy().copy()
    # a2 add_smoke_source: default radius, facade radius, overlapping, clipped by the border
    s = NavierStokesSimulator((48, 40), 0.01, 0.001, "cpu")
    src = [(20, 24, 10, 1.0), (22, 25, 8, 1.5), (2, 45, 8, 0.7), (39, 0, 8, 2.0)]
    for x_, y_, r_, i_ in src:
        s.add_smoke_source(x_, y_, radius=r_, intensity=i_)
    out["splat_src"] = np.array(src, dtype=np.float64)
    out["splat_out"] = s.density.numpy().copy()
    save("units.npz", **out)


# --------------------------------------------------------------------- facade
def facade():
    out = {}
    fg = FractalGenerator("cpu")
    for n in (16, 32, 64, 128, 200):
        # the torch.linspace grids themselves (ATen's vectorised CPU linspace is ISA-dependent in the last ulp)
        out["lin_p_%d" % n] = torch.linspace(0, 10.0, n).numpy()
        out["lin_mx_%d" % n] = torch.linspace(-2.5, 1.5, n).numpy()
        out["lin_my_%d" % n] = torch.linspace(-1.5, 1.5, n).numpy()
        out["perlin_%d" % n] = fg.generate_perlin_noise((n, n)).numpy()
        out["mandel_count_%d" % n] = np.rint(fg.generate_mandelbrot_field((n, n)).numpy() * 100).astype(np.int16)
        out["mandel_%d" % n] = fg.generate_mandelbrot_field((n, n)).numpy()
        out["pert_%d" % n] = fg.apply_fractal_perturbation(torch.ones(n, n), 0.05).numpy()
    # simulate_step sequence of inference.py:35-50 (frames are the fractal-scaled copies)
    sim = SmokeSimulator((128, 128), 0.01, 0.001, "cpu")
    sim.ns_solver.setup_grid()
    sim.add_incense_source([(64, 64), (32, 32), (96, 96)], [1.5, 1.0, 0.8])
    frames = [sim.simulate_step().numpy().copy() for _ in range(20)]
    out["frames_sha"] = np.array([sha(f) for f in frames])
    out["frame_1"], out["frame_20"] = frames[0], frames[-1]
    out["frames_sum"] = np.array([f.astype(np.float64).sum() for f in frames])
    out["frames_max"] = np.array([f.max() for f in frames], dtype=np.float64)
    out["nofractal_21"] = sim.simulate_step(add_fractal=False).numpy().copy()
    feats = sim.get_chaos_features()
    out["chaos"] = np.array([feats["lyapunov_exponent"], feats["fractal_dimension"], feats["entropy"]], dtype=np.float64)
    out["history_len"] = np.int64(len(sim.history))
    save("facade.npz", **out)


# -------------------------------------------- batch / K extension (configs 2, 3)
def batch_k40():
    """Config-2 shaped case at reduced size: 6 independent 128x128 sequences, K=40, seeds 1234+s
    (emitter law of data_loader.py:49-58), 5 steps. Only hashes + one full sequence are stored."""
    out = {}
    K, B, steps = 40, 6, 5
    cls = with_iters(K)
    hashes, srcs, offs = [], [], [0]
    for s in range(B):
        rng = np.random.default_rng(1234 + s)
        n = int(rng.integers(1, 4))
        ns = cls((128, 128), 0.01, 0.001, "cpu")
        for _ in range(n):
            x = int(rng.integers(20, 128 - 20)); y = int(rng.integers(20, 128 - 20))
            inten = float(rng.uniform(0.5, 2.0))
            ns.add_smoke_source(x, y, radius=8, intensity=inten)
            srcs.append([s, x, y, 8, inten])
        offs.append(len(srcs))
        out["density0_%d" % s] = ns.density.numpy().copy()
        for t in range(steps):
            ns.step()
        hashes.append([sha(getattr(ns, k).numpy()) for k in ("u", "v", "p", "density")])
        if s == 3:
            for k, a in state(ns).items():
                out["final3_" + k] = a
    out["sources"] = np.array(srcs, dtype=np.float64)
    out["offsets"] = np.array(offs)
    out["sha"] = np.array(hashes)
    out["meta"] = np.array([K, B, steps])
    save("batch_k40.npz", **out)


def grid_k100():
    """Config-3 shaped case at reduced size: one 256x256 grid, one emitter per 64x64 block, K=100, 3 steps."""
    out = {}
    n, K = 256, 100
    ns = with_iters(K)((n, n), 0.01, 0.001, "cpu")
    rng = np.random.default_rng(99)
    srcs = []
    for by in range(n // 64):
        for bx in range(n // 64):
            x = int(bx * 64 + rng.integers(8, 56)); y = int(by * 64 + rng.integers(8, 56))
            inten = float(rng.uniform(0.5, 2.0))
            ns.add_smoke_source(x, y, radius=8, intensity=inten)
            srcs.append([0, x, y, 8, inten])
    out["sources"] = np.array(srcs, dtype=np.float64)
    out["density0"] = ns.density.numpy().copy()
    hashes = []
    for t in range(3):
        ns.step()
        hashes.append([sha(getattr(ns, k).numpy()) for k in ("u", "v", "p", "density")])
    out["sha"] = np.array(hashes)
    out["final_p"] = ns.p.numpy().copy()
    out["final_v"] = ns.v.numpy().copy()
    save("grid_k100.npz", **out)


if __name__ == "__main__":
    print("reference:", REF, "torch", torch.__version__, "numpy", np.__version__)
    scenario_c1()
    small_random()
    units()
    facade()
    batch_k40()
    grid_k100()
