"""Parity of the CUDA path (through the Python class surface and the C ABI beneath it) against the
golden vectors generated from the reference and against the live CPU oracle.

Bar: the whole step is +,-,*,/ ,floor, clamp and gathers, kept in the reference's
association with no FMA, so every field must be BIT-EQUAL (numeric ==) to the reference's CPU result.
The only tolerances are the two transcendental spots: expf in the emitter splat (<= 5e-7 normalised)
and sinf/cosf in the Perlin octaves (<= 5e-7 absolute).  north_star's 1e-4 is therefore met with
four orders of margin; the tests assert the tighter bound.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import assert_same, emitters_for_sequence, nerr, sha, smk_env

pytestmark = pytest.mark.gpu

from smokephysai_b200 import FractalGenerator, NavierStokesSimulator, SmokeSimulator, _lib  # noqa: E402


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def N(t):
    return t.detach().cpu().numpy()


def make(h, w, dt=0.01, nu=0.001, K=20, **kw):
    return NavierStokesSimulator((h, w), dt, nu, "cuda", jacobi_iters=K, **kw)


# ------------------------------------------------------------------------------------------- unit hooks
def test_units_diffusion(golden):
    g = golden("units")
    ns = make(20, 28)
    for nm in ("u", "v", "d"):
        out = ns.diffusion_step(T(g["diff_%s_in" % nm]), float(g["diff_%s_visc" % nm]))
        assert_same(N(out), g["diff_%s_out" % nm], "diffusion " + nm)


def test_units_bilerp_and_interp(golden):
    g = golden("units")
    ns = make(20, 28)
    assert_same(N(ns.bilinear_interpolate(T(g["bil_f"]), T(g["bil_y"]), T(g["bil_x"]))), g["bil_out"], "bilerp")
    Y, X = torch.meshgrid(torch.arange(20, dtype=torch.float32), torch.arange(28, dtype=torch.float32), indexing="ij")
    assert_same(N(ns.interpolate_velocity_u(T(g["adv_u"]), Y, X)), g["interp_u_out"], "interp u")
    assert_same(N(ns.interpolate_velocity_v(T(g["adv_v"]), Y, X)), g["interp_v_out"], "interp v")


def test_units_advection(golden):
    g = golden("units")
    ns = make(20, 28)
    u, v, d = T(g["adv_u"]), T(g["adv_v"]), T(g["adv_d"])
    assert_same(N(ns.advection_step(u, u, v)), g["adv_u_out"], "advect u")
    assert_same(N(ns.advection_step(v, u, v)), g["adv_v_out"], "advect v")
    assert_same(N(ns.advection_step(d, u, v)), g["adv_d_out"], "advect d")


@pytest.mark.parametrize("K", [20, 7])
def test_units_projection(golden, K):
    g = golden("units")
    ns = make(20, 28, K=K)
    for k in ("u", "v", "p"):
        setattr(ns, k, T(g["proj%d_%s_in" % (K, k)]))
    ns.pressure_projection()
    for k in ("p", "u", "v"):
        assert_same(N(getattr(ns, k)), g["proj%d_%s_out" % (K, k)], "projection " + k)


def test_units_splat(golden):
    g = golden("units")
    ns = make(48, 40)
    for x, y, r, i in g["splat_src"]:
        ns.add_smoke_source(int(x), int(y), radius=int(r), intensity=float(i))
    d = N(ns.density)
    assert np.array_equal(d != 0, g["splat_out"] != 0)
    assert nerr(d, g["splat_out"]) < 5e-7


@pytest.mark.parametrize("h,w,B,n", [(200, 333, 1, 700), (130, 260, 3, 300), (64, 128, 2, 40), (1024, 1024, 1, 2000)])
def test_splat_big_tile_kernel_equals_the_small_tile_one(h, w, B, n):
    """k_splat_big (a CTA owns 128 x 64 cells: big grids with thousands of emitters) against k_splat (32 x 8) bit for bit, and against
    the oracle's add_smoke_source: overlapping emitters (the sum order matters), emitters across tile edges and outside the grid,
    lists longer than one 256-entry cull pass, a simulation with no emitter, a non-zero starting density."""
    rng = np.random.default_rng(h + n)
    per_sim = []
    for b in range(B):
        m = 0 if (B > 1 and b == 1) else n
        per_sim.append([(int(rng.integers(-5, w + 5)), int(rng.integers(-5, h + 5)), int(rng.integers(1, 14)), float(rng.uniform(0.2, 2.0)))
                        for _ in range(m)])
    d0 = rng.uniform(0, 1, (B, h, w)).astype(np.float32)
    got = {}
    for big in (0, 1):
        with smk_env(SMK_SPLAT_BIG=big):
            ns = make(h, w, batch=B)
            ns.density = T(d0 if B > 1 else d0[0])
            ns.add_sources(per_sim)
            got[big] = N(ns.density).reshape(B, h, w)
    assert np.array_equal(got[0], got[1])
    if h * w <= 64 * 1024:
        for b in range(B):
            want = d0[b].copy()
            for x, y, r, i in per_sim[b]:
                want = oracle.splat(want, x, y, r, i)
            assert nerr(got[1][b], want) < 5e-7


# ------------------------------------------------------------------------------------- whole-step parity
def test_small_random_steps(golden):
    g = golden("small_random")
    for tag in g["cases"]:
        h, w, K, dt, nu = g["%s_meta" % tag]
        ns = make(int(h), int(w), float(dt), float(nu), int(K))
        for k in ("u", "v", "p", "density"):
            setattr(ns, k, T(g["%s_0_%s" % (tag, k)]))
        for t in (1, 2, 3):
            frame = ns.step()
            for k in ("u", "v", "p", "density"):
                assert_same(N(getattr(ns, k)), g["%s_%d_%s" % (tag, t, k)], "case %s step %d %s" % (tag, t, k))
            assert_same(N(frame), N(ns.density), "returned frame")
            assert frame.data_ptr() != ns.density.data_ptr()


@pytest.mark.parametrize("K", [20, 40, 100])
def test_scenario_c1(golden, K):
    """inference.py:40-41 emitters on the config.yaml grid (BASELINE config 1)."""
    g = golden("scenario_c1")
    ns = make(128, 128, K=K)
    ns.density = T(g["density0"])
    rows = {int(r[0]): (r, hs) for r, hs in zip(g["stats_K%d" % K], g["sha_K%d" % K])}
    for t in range(1, 21):
        ns.step()
        if t in rows:
            r, hs = rows[t]
            got = [sha(N(getattr(ns, k))) for k in ("u", "v", "p", "density")]
            assert got == list(hs), "K=%d t=%d sha mismatch" % (K, t)
    if K == 20:
        for k in ("u", "v", "p", "density"):
            assert_same(N(getattr(ns, k)), g["final_" + k], k)


def test_post_projection_divergence_norms(golden):
    """smk_div_norms after pressure_projection matches the reference's residual (north_star: matching
    post-projection divergence residual)."""
    g = golden("scenario_c1")
    ns = make(128, 128)
    ns.density = T(g["density0"])
    row = g["stats_K20"][0]        # t = 1
    st = ns._state
    prm = ns._params()
    grid = C.byref(ns._grid)
    s = ns._stream()
    _lib.call("smk_forces_diffuse_div", grid, st.u[0], st.v[0], st.d[0], st.u[1], st.v[1], st.d[1], st.div,
              prm.dt, prm.c_uv, prm.c_d, s)
    st.cur_u = st.cur_v = st.cur_d = 1
    flag = C.c_int32(0)
    _lib.call("smk_jacobi", grid, st.div, st.p[0], st.p[1], 20, 0, C.byref(flag), s)
    st.cur_p = flag.value
    _lib.call("smk_project", grid, st.p[st.cur_p], st.u[1], st.v[1], prm.dt, s)
    nrm = ns.divergence_norms()[0]
    assert abs(float(nrm[0]) - row[6]) <= 1e-6 * row[6]
    assert abs(float(nrm[1]) - row[7]) <= 1e-4 * row[7]      # fp32 atomics sum vs float64 reference sum


def test_batch_k40(golden):
    """BASELINE config 2 at reduced size: independent sequences, K=40, batched in one state."""
    g = golden("batch_k40")
    K, B, steps = (int(x) for x in g["meta"])
    ns = make(128, 128, K=K, batch=B)
    ns.density = T(np.stack([g["density0_%d" % s] for s in range(B)]))
    frames = ns.run_steps(steps)
    assert tuple(frames.shape) == (B, steps, 128, 128)
    for s in range(B):
        got = [sha(N(getattr(ns, k)[s])) for k in ("u", "v", "p", "density")]
        assert got == list(g["sha"][s]), "sequence %d" % s
    for k in ("u", "v", "p", "density"):
        assert_same(N(getattr(ns, k)[3]), g["final3_" + k], k)
    assert_same(N(frames[:, -1]), N(ns.density), "last frame")


def test_batch_own_splat_matches_golden_support(golden):
    g = golden("batch_k40")
    B = int(g["meta"][1])
    ns = make(128, 128, K=40, batch=B)
    per = [[] for _ in range(B)]
    for s, x, y, r, i in g["sources"]:
        per[int(s)].append((int(x), int(y), int(r), float(i)))
    ns.add_sources(per)
    d = N(ns.density)
    for s in range(B):
        assert per[s] == [tuple(e) for e in emitters_for_sequence(s)]
        assert nerr(d[s], g["density0_%d" % s]) < 5e-7


@pytest.mark.parametrize("T_", [0, 1, 5, 8, 12, 24])
def test_grid_k100_tiled_jacobi(golden, T_):
    """BASELINE config 3 at reduced size (256x256, K=100): exercises the overlapped-tile temporal blocking."""
    g = golden("grid_k100")
    ns = make(256, 256, K=100, sweeps_per_launch=T_)
    ns.density = T(g["density0"])
    for t in range(3):
        ns.step()
        got = [sha(N(getattr(ns, k))) for k in ("u", "v", "p", "density")]
        assert got == list(g["sha"][t]), "T=%d t=%d" % (T_, t)
    assert_same(N(ns.p), g["final_p"], "p")


# ----------------------------------------------------------------- live oracle: ragged / odd / edge sizes
@pytest.mark.parametrize("h,w,K,T_", [
    (1, 1, 3, 0), (2, 2, 5, 0), (3, 3, 4, 0), (5, 131, 9, 4), (131, 5, 9, 4), (129, 129, 20, 8),
    (130, 260, 17, 5), (300, 200, 33, 8), (257, 383, 20, 12), (64, 64, 20, 0), (33, 128, 20, 0), (100, 127, 40, 0),
])
def test_step_vs_oracle_ragged(h, w, K, T_):
    rng = np.random.default_rng(h * 1000 + w)
    ref = oracle.OracleSolver((h, w), 0.02, 0.01, K)
    ref.u = ((rng.random((h + 1, w)) - 0.5) * 300).astype(np.float32)
    ref.v = ((rng.random((h, w + 1)) - 0.5) * 300).astype(np.float32)
    ref.p = rng.standard_normal((h, w)).astype(np.float32)
    ref.density = rng.random((h, w)).astype(np.float32)
    ns = make(h, w, 0.02, 0.01, K, sweeps_per_launch=T_)
    for k in ("u", "v", "p", "density"):
        setattr(ns, k, T(getattr(ref, k)))
    for t in range(2):
        ref.step()
        ns.step()
        for k in ("u", "v", "p", "density"):
            assert_same(N(getattr(ns, k)), getattr(ref, k), "%dx%d K=%d T=%d step %d %s" % (h, w, K, T_, t, k))


@pytest.mark.parametrize("h,w,K,T_", [(128, 128, 40, 0), (200, 136, 10, 3), (512, 640, 24, 8)])
def test_jacobi_abi_vs_oracle(h, w, K, T_):
    """smk_jacobi directly through the C ABI on random p (non-zero ring) and div."""
    rng = np.random.default_rng(5)
    p = rng.standard_normal((h, w)).astype(np.float32)
    div = rng.standard_normal((h, w)).astype(np.float32)
    ns = make(h, w, K=K, sweeps_per_launch=T_)
    ns.p = T(p)
    ns._field("div").copy_(T(div))
    st = ns._state
    flag = C.c_int32(0)
    _lib.call("smk_jacobi", C.byref(ns._grid), st.div, st.p[0], st.p[1], K, T_, C.byref(flag), ns._stream())
    st.cur_p = flag.value
    assert_same(N(ns.p), oracle.jacobi(p, div, K), "jacobi")


def test_full_size_c3_vs_oracle():
    """BASELINE config 3 at full size: one 1024x1024 grid, K=100, one emitter per 64x64 block; 2 steps vs the oracle."""
    n, K = 1024, 100
    rng = np.random.default_rng(3)
    ref = oracle.OracleSolver((n, n), 0.01, 0.001, K)
    ns = make(n, n, K=K)
    src = []
    for by in range(n // 64):
        for bx in range(n // 64):
            src.append((int(bx * 64 + rng.integers(8, 56)), int(by * 64 + rng.integers(8, 56)), 8, float(rng.uniform(0.5, 2.0))))
    ns.add_sources([src])
    ref.density = N(ns.density).copy()          # isolate expf: same initial density on both sides
    for t in range(2):
        ref.step()
        ns.step()
    for k in ("u", "v", "p", "density"):
        assert_same(N(getattr(ns, k)), getattr(ref, k), "1024^2 K=100 " + k)


def test_full_size_c2_properties():
    """BASELINE config 2 at full size (256 x 128^2, K=40): batch result == each sequence run alone, and
    sequence 0..3 == oracle after 20 steps."""
    B, K, steps = 256, 40, 20
    ns = make(128, 128, K=K, batch=B)
    ems = [emitters_for_sequence(s) for s in range(B)]
    ns.add_sources(ems)
    d0 = N(ns.density).copy()
    frames = ns.run_steps(steps)
    u, v, p, d = (N(getattr(ns, k)) for k in ("u", "v", "p", "density"))
    # vs oracle on a few sequences (same initial density)
    sel = [0, 1, 2, 3, 100, 255]
    ou = np.zeros((len(sel), 129, 128), np.float32); ov = np.zeros((len(sel), 128, 129), np.float32)
    op = np.zeros((len(sel), 128, 128), np.float32); od = d0[sel].copy()
    ofr = oracle.run_batch(ou, ov, op, od, 0.01, 0.001, K, steps, nthreads=4)
    for n_, s in enumerate(sel):
        assert_same(u[s], ou[n_], "u seq %d" % s); assert_same(v[s], ov[n_], "v seq %d" % s)
        assert_same(p[s], op[n_], "p seq %d" % s); assert_same(d[s], od[n_], "d seq %d" % s)
        assert_same(N(frames[s]), ofr[n_], "frames seq %d" % s)
    # batch == alone
    one = make(128, 128, K=K)
    one.density = T(d0[77])
    fr1 = one.run_steps(steps)
    assert_same(N(fr1), N(frames[77]), "sequence 77 alone vs in batch")
    # edge quirk at scale: last row / col of density are exactly zero, pressure ring is zero
    assert not d[:, -1, :].any() and not d[:, :, -1].any()
    assert not p[:, 0, :].any() and not p[:, -1, :].any() and not p[:, :, 0].any() and not p[:, :, -1].any()


# --------------------------------------------------------------------------------------------- facade
@pytest.mark.parametrize("n", [16, 32, 64, 128, 200])
def test_fractal_fields(golden, n):
    g = golden("facade")
    fg = FractalGenerator("cuda")
    # pin the kernel with the golden linspace grids (ISA-independent), then check the host-made grids agree here
    pitch = (n + 3) & ~3
    outs = [torch.zeros(n, pitch, device="cuda") for _ in range(3)]
    grids = [T(g["lin_p_%d" % n]), T(g["lin_mx_%d" % n]), T(g["lin_my_%d" % n])]
    _lib.call("smk_fractal_fields", outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), n, n, pitch, 0.05, 100,
              grids[0].data_ptr(), grids[0].data_ptr(), grids[1].data_ptr(), grids[2].data_ptr(), fg._stream())
    perlin, mandel, mul = (N(o[:, :n]) for o in outs)
    assert_same(np.rint(mandel * 100), g["mandel_count_%d" % n].astype(np.float32), "mandelbrot escape counts")
    assert_same(mandel, g["mandel_%d" % n], "mandelbrot field")
    assert np.abs(perlin - g["perlin_%d" % n]).max() < 5e-7
    assert np.abs((1.0 + mul) - g["pert_%d" % n]).max() < 3e-7
    # class surface
    assert np.abs(N(fg.generate_perlin_noise((n, n))) - g["perlin_%d" % n]).max() < 2e-6
    assert np.abs(N(fg.apply_fractal_perturbation(torch.ones(n, n), 0.05)) - g["pert_%d" % n]).max() < 2e-6
    m = N(fg.generate_mandelbrot_field((n, n)))
    assert (np.rint(m * 100) != g["mandel_count_%d" % n]).sum() <= 2     # host linspace ulps may move a razor-edge pixel


def test_fractal_nonsquare_raises_like_reference():
    sim = SmokeSimulator((96, 160), device="cuda")
    sim.add_incense_source([(50, 40)], [1.0])
    sim.simulate_step(add_fractal=False)            # the solver itself handles non-square grids
    with pytest.raises(RuntimeError):
        sim.simulate_step()                          # fractal field is (w, h): the reference fails to broadcast


def test_facade_sequence(golden):
    """inference.py:35-50: 20 simulate_step() frames, then add_fractal=False, then chaos features."""
    g, c1 = golden("facade"), golden("scenario_c1")
    sim = SmokeSimulator((128, 128), 0.01, 0.001, "cuda")
    sim.ns_solver.setup_grid()
    sim.add_incense_source([(64, 64), (32, 32), (96, 96)], [1.5, 1.0, 0.8])
    assert nerr(N(sim.ns_solver.density), c1["density0"]) < 5e-7
    sim.ns_solver.density = T(c1["density0"])
    frames = [N(sim.simulate_step()) for _ in range(20)]
    assert nerr(frames[0], g["frame_1"]) < 3e-7 and nerr(frames[-1], g["frame_20"]) < 3e-7
    for t in range(20):
        assert abs(float(frames[t].astype(np.float64).sum()) - g["frames_sum"][t]) < 1e-4
    assert_same(N(sim.ns_solver.density), c1["final_density"], "solver state is not perturbed by the fractal multiply")
    assert_same(N(sim.simulate_step(add_fractal=False)), g["nofractal_21"], "add_fractal=False")
    assert len(sim.history) == int(g["history_len"]) == 21
    feats = sim.get_chaos_features()
    ref = g["chaos"]
    assert abs(feats["lyapunov_exponent"] - ref[0]) < 1e-6
    assert abs(feats["fractal_dimension"] - ref[1]) < 1e-9
    assert abs(feats["entropy"] - ref[2]) < 1e-5


def test_history_ring_and_reset():
    sim = SmokeSimulator((32, 32), device="cuda")
    sim.max_history = 5
    sim.add_incense_source([(16, 16)], [1.0])
    for _ in range(8):
        sim.simulate_step()
    assert len(sim.history) == 5
    sim.ns_solver.setup_grid()
    assert len(sim.history) == 5            # setup_grid() does not clear history
    assert not N(sim.ns_solver.density).any() and not N(sim.ns_solver.p).any()
    assert sim.get_chaos_features() == {}


def test_generate_sequences_matches_loop():
    """Batched back-end of data_loader.py:37-99 == the reference's per-sample loop through the scalar API."""
    B, L = 5, 6
    ems = [[((x, y), i) for x, y, _, i in emitters_for_sequence(s)] for s in range(B)]
    bat = SmokeSimulator((128, 128), device="cuda", batch=B)
    fr = N(bat.generate_sequences(ems, L))
    for _ in range(2):                                   # the pinned, double-streamed path reuses its buffers
        fr_h = bat.generate_sequences(ems, L, to_host=True)
        assert fr_h.device.type == "cpu" and fr_h.is_pinned() and tuple(fr_h.shape) == (B, L, 128, 128)
        assert_same(fr_h.numpy(), fr, "to_host frames")
    one = SmokeSimulator((128, 128), device="cuda")
    for s in range(B):
        one.ns_solver.setup_grid()
        one.add_incense_source([e[0] for e in ems[s]], [e[1] for e in ems[s]])
        seq = np.stack([N(one.simulate_step()) for _ in range(L)])
        assert_same(fr[s], seq, "sequence %d" % s)


def test_device_cpu_returns_cpu_tensors():
    sim = SmokeSimulator((32, 32), device="cpu")         # benchmark.py:260
    sim.add_incense_source([(16, 16)], [1.0])
    f = sim.simulate_step()
    assert f.device.type == "cpu" and sim.ns_solver.density.device.type == "cpu"
    gsim = SmokeSimulator((32, 32), device="cuda")
    gsim.add_incense_source([(16, 16)], [1.0])
    assert_same(N(gsim.simulate_step()), N(f), "cpu-returning and cuda-returning simulators agree")


def test_inplace_field_edits_and_pickle():
    import pickle
    ns = make(16, 16)
    ns.density[4:8, 4:8] += 1.0
    ns.density *= 0.5
    assert float(ns.density.sum()) == 8.0
    fr = ns.step()
    assert torch.equal(pickle.loads(pickle.dumps(fr.cpu())), fr.cpu())
    with pytest.raises(ValueError):
        ns.u = torch.zeros(3, 3)


def test_abi_rejects_bad_arguments():
    ns = make(16, 16)
    lib = _lib.load()
    g = ns._grid
    st = ns._state
    flag = C.c_int32(0)
    assert lib.smk_jacobi(C.byref(g), st.div, st.p[0], st.p[0], 4, 0, C.byref(flag), None) == -1
    assert b"differ" in lib.smk_last_error_string()
    assert lib.smk_jacobi(None, st.div, st.p[0], st.p[1], 4, 0, C.byref(flag), None) == -1
    assert lib.smk_project(C.byref(g), st.p[0] + 4, st.u[0], st.v[0], 0.01, None) == -1      # misaligned
    bad = type(g)(16, 16, 1, 15, 20, 16, 0, 0, 0)
    assert lib.smk_divergence(C.byref(bad), st.u[0], st.v[0], st.div, 0.01, None) == -1
    with pytest.raises(_lib.SmokeLibraryError):
        _lib.call("smk_diffuse", st.u[0], st.u[0], 4, 4, 4, 1, 16, 0.1, None)


@pytest.mark.parametrize("tile", [1, 2, 3, "stream1", "stream2", "half-stream1", "half-stream2"])
@pytest.mark.parametrize("h,w,K,T_", [(300, 200, 33, 8), (130, 520, 20, 10), (1030, 260, 24, 12), (700, 900, 7, 1), (129, 131, 5, 5),
                                      (1500, 1900, 20, 10)])
def test_jacobi_every_tile_shape_vs_oracle(tile, h, w, K, T_):
    """The tiled Jacobi picks its CTA tile by grid size (128 x 128 one CTA per SM, 64 x 128 two CTAs per SM for grids of
    many CTA waves); every shape must give the oracle's pressure bit for bit.  SMK_JACOBI_TILE forces the shape."""
    stream = int(tile[-1]) if isinstance(tile, str) else 0
    rng = np.random.default_rng((9 if stream else tile) * 100 + h)
    p = rng.standard_normal((h, w)).astype(np.float32)
    div = rng.standard_normal((h, w)).astype(np.float32)
    ns = make(h, w, K=K, sweeps_per_launch=T_)
    ns.p = T(p)
    ns._field("div").copy_(T(div))
    st = ns._state
    flag = C.c_int32(0)
    # "stream1/2": the persistent kernel that prefetches the next 128 x 128 tile into shared memory with 16-byte cp.async (1) or
    # with two TMA tensor loads completed on an mbarrier (2); 1500 x 1900 gives its CTAs two tiles each, the others one or none
    # "half-stream1/2": the same kernel on 64 x 128 tiles, two persistent CTAs per SM
    with smk_env(SMK_JACOBI_TILE=("1" if tile.startswith("half") else "2") if stream else str(tile), SMK_JACOBI_STREAM=stream):
        _lib.call("smk_jacobi", C.byref(ns._grid), st.div, st.p[0], st.p[1], K, T_, C.byref(flag), ns._stream())
    st.cur_p = flag.value
    assert_same(N(ns.p), oracle.jacobi(p, div, K), "jacobi tile %s" % tile)


@pytest.mark.parametrize("stream", [1, 2])
def test_jacobi_stream_batched_vs_oracle(stream):
    """The streaming kernel's tile index runs over (batch, tile row, tile column): 40 simulations of 300 x 260 are 360 tiles,
    so every CTA walks across simulation boundaries."""
    h, w, K, B = 300, 260, 12, 40
    rng = np.random.default_rng(77 + stream)
    p = rng.standard_normal((B, h, w)).astype(np.float32)
    div = rng.standard_normal((B, h, w)).astype(np.float32)
    ns = make(h, w, K=K, sweeps_per_launch=6, batch=B)
    ns.p = T(p)
    ns._field("div").copy_(T(div))
    st = ns._state
    flag = C.c_int32(0)
    with smk_env(SMK_JACOBI_TILE=2, SMK_JACOBI_STREAM=stream):
        _lib.call("smk_jacobi", C.byref(ns._grid), st.div, st.p[0], st.p[1], K, 6, C.byref(flag), ns._stream())
    st.cur_p = flag.value
    got = N(ns.p)
    for b in range(B):
        assert_same(got[b], oracle.jacobi(p[b], div[b], K), "jacobi stream %d, simulation %d" % (stream, b))


@pytest.fixture
def force_advect_kernel():
    """SMK_ADVECT_TILED = 1 / 0 forces the shared-memory tiled / the direct advection kernel (the library picks by field
    size otherwise, so small test grids would never reach the tiled one)."""
    old = os.environ.get("SMK_ADVECT_TILED")

    def set_(v):
        os.environ["SMK_ADVECT_TILED"] = str(v)
        _lib.reload_env()
    yield set_
    os.environ.pop("SMK_ADVECT_TMA", None)
    if old is None:
        os.environ.pop("SMK_ADVECT_TILED", None)
    else:
        os.environ["SMK_ADVECT_TILED"] = old
    _lib.reload_env()


@pytest.mark.parametrize("tiled", [0, 1, 2])
@pytest.mark.parametrize("h,w,K", [(1, 1, 2), (3, 3, 2), (5, 131, 3), (131, 5, 3), (129, 129, 4), (130, 260, 4), (300, 200, 3),
                                   (257, 383, 3), (16, 128, 2), (17, 132, 2), (40, 300, 2), (64, 1000, 2)])
def test_both_advection_kernels_vs_oracle(force_advect_kernel, tiled, h, w, K):
    """Phase-per-kernel step with each advection kernel forced, on ragged grids with back-traces of up to 3 cells (inside
    the tiled kernel's staged window) and of up to 12 cells (its global fallback), against the oracle, bit for bit.
    tiled: 0 the direct kernel, 1 the tiled one (interior tiles staged by TMA box loads), 2 the tiled one with cp.async staging
    everywhere (SMK_ADVECT_TMA=0)."""
    force_advect_kernel(min(tiled, 1))
    if tiled == 2:
        os.environ["SMK_ADVECT_TMA"] = "0"
        _lib.reload_env()
    for vel in (300.0, 1200.0):
        rng = np.random.default_rng(h * 1000 + w + int(vel))
        ref = oracle.OracleSolver((h, w), 0.02, 0.01, K)
        ref.u = ((rng.random((h + 1, w)) - 0.5) * vel).astype(np.float32)
        ref.v = ((rng.random((h, w + 1)) - 0.5) * vel).astype(np.float32)
        ref.p = rng.standard_normal((h, w)).astype(np.float32)
        ref.density = rng.random((h, w)).astype(np.float32)
        ns = make(h, w, 0.02, 0.01, K, step_kernel="phases")
        for k in ("u", "v", "p", "density"):
            setattr(ns, k, T(getattr(ref, k)))
        fmul = torch.rand(h, ns._layout.pitch_c, device="cuda") * 0.05
        for t in range(2):
            fr_ref = ref.step()
            fr = ns.step(fmul=fmul)
            for k in ("u", "v", "p", "density"):
                assert_same(N(getattr(ns, k)), getattr(ref, k), "%dx%d vel %g tiled %d step %d %s" % (h, w, vel, tiled, t, k))
            want = fr_ref + N(fmul)[:, :w] * fr_ref
            assert_same(N(fr), want, "frame with fractal multiplier")


@pytest.mark.parametrize("tiled", [0, 1])
def test_both_advection_kernels_in_slabs(force_advect_kernel, tiled):
    """Slab variant of both advection kernels (global coordinates, local memory, overflow guard): three slabs == whole."""
    from smokephysai_b200.slab import LocalGroup
    force_advect_kernel(tiled)
    H, W, K = 150, 260, 6
    rng = np.random.default_rng(9)
    st0 = {"u": ((rng.random((H + 1, W)) - 0.5) * 200).astype(np.float32), "v": ((rng.random((H, W + 1)) - 0.5) * 200).astype(np.float32),
           "p": rng.standard_normal((H, W)).astype(np.float32), "d": rng.random((H, W)).astype(np.float32)}
    whole = make(H, W, 0.02, 0.01, K, step_kernel="phases")
    grp = LocalGroup((H, W), 0.02, 0.01, "cuda", world=3, jacobi_iters=K, sweeps_per_launch=3)
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        setattr(whole, name, T(st0[k]))
        grp.scatter(k, st0[k])
    for _ in range(3):
        whole.step()
        grp.step()
    grp.check()
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        assert_same(N(grp.gather(k))[:, :st0[k].shape[1]], N(getattr(whole, name)), "%s tiled %d" % (k, tiled))


@pytest.mark.parametrize("bulk", [0, 1, 2])
@pytest.mark.parametrize("h,w", [(100, 400), (64, 300), (35, 260), (130, 1000), (48, 392)])
def test_both_staging_paths_of_forces_diffuse_div_vs_oracle(bulk, h, w):
    """k_forces_diffuse_div stages interior tiles of big grids with bulk cp.async copies (SMK_FDD_BULK=1) or three TMA box loads
    (=2, the default there); 0 is the LDG -> STS path of small grids and edge tiles.  All three, on grids that have interior AND
    edge tiles, against the oracle, bit for bit."""
    with smk_env(SMK_FDD_BULK=bulk):
        rng = np.random.default_rng(h + w + bulk)
        ref = oracle.OracleSolver((h, w), 0.02, 0.01, 5)
        ref.u = ((rng.random((h + 1, w)) - 0.5) * 40).astype(np.float32)
        ref.v = ((rng.random((h, w + 1)) - 0.5) * 40).astype(np.float32)
        ref.p = rng.standard_normal((h, w)).astype(np.float32)
        ref.density = rng.random((h, w)).astype(np.float32)
        ns = make(h, w, 0.02, 0.01, 5, step_kernel="phases")
        for k in ("u", "v", "p", "density"):
            setattr(ns, k, T(getattr(ref, k)))
        for t in range(2):
            ref.step()
            ns.step()
            for k in ("u", "v", "p", "density"):
                assert_same(N(getattr(ns, k)), getattr(ref, k), "%dx%d bulk %d step %d %s" % (h, w, bulk, t, k))


def test_jacobi_residual_norms_vs_oracle():
    """north_star: warp-shuffle reductions for the residual norm.  smk_jacobi_residual = max and L2 of p' - p, p' one more
    sweep of navier_stokes.py:139-145 over the pressure and divergence the projection just used."""
    h, w, K = 96, 160, 20
    rng = np.random.default_rng(5)
    u = ((rng.random((h + 1, w)) - 0.5) * 2).astype(np.float32)
    v = ((rng.random((h, w + 1)) - 0.5) * 2).astype(np.float32)
    p0 = rng.standard_normal((h, w)).astype(np.float32)
    ns = make(h, w, K=K)
    ns.u, ns.v, ns.p = T(u), T(v), T(p0)
    ns.pressure_projection()
    got = ns.jacobi_residual_norms()[0].numpy()
    div = oracle.divergence(u, v, 0.01)
    pK = oracle.jacobi(p0, div, K)
    assert_same(N(ns.p), pK, "p after K sweeps")
    r = (oracle.jacobi(pK, div, 1) - pK).astype(np.float32)            # fp32 difference, as the kernel forms it
    assert got[0] == np.abs(r).max()                                   # the max is exact
    l2 = np.sqrt((r.astype(np.float64) ** 2).sum())
    assert abs(got[1] - l2) <= 1e-5 * l2                               # fp32 partial sums in a different order: 1e-5 relative
    # more sweeps bring the iteration closer to its fixed point
    ns2 = make(h, w, K=4 * K)
    ns2.u, ns2.v, ns2.p = T(u), T(v), T(p0)
    ns2.pressure_projection()
    assert ns2.jacobi_residual_norms()[0, 1] < got[1]


def test_advection_step_refuses_shapes_it_would_read_out_of_bounds():
    """the hook builds the cell grid from v's rows and u's columns; anything inconsistent is refused."""
    ns = make(16, 20)
    u, v = torch.zeros(17, 20), torch.zeros(16, 21)
    ns.advection_step(torch.zeros(17, 20), u, v)                       # u-shaped, v-shaped and cell-centred fields are fine
    ns.advection_step(torch.zeros(16, 21), u, v)
    with pytest.raises(ValueError):
        ns.advection_step(torch.zeros(16, 20), torch.zeros(16, 20), v)   # u without its staggered row
    with pytest.raises(ValueError):
        ns.advection_step(torch.zeros(16, 20), u, torch.zeros(16, 20))   # v without its staggered column
    with pytest.raises(ValueError):
        ns.advection_step(torch.zeros(18, 20), u, v)                     # a field taller than u
    g = _lib.Grid(16, 20, 1, 20, 24, 20, 17 * 20, 16 * 24, 16 * 20, 0, 0)
    buf = torch.zeros(4096, device="cuda")
    rc = _lib.load().smk_advect(C.byref(g), buf.data_ptr(), buf.data_ptr() + 8192, 18, 20, 20, 0, buf.data_ptr(), buf.data_ptr(),
                                0.01, 1.0, None, 0, None, None)
    assert rc == -1 and b"does not fit" in _lib.load().smk_last_error_string()


def test_cpu_output_device_writes_in_place_edits_through():
    """with device='cpu' (benchmark.py:260) the field properties hand out host copies; in-place edits of those
    copies -- which work on the reference's live tensors -- must reach the device state before the next launch."""
    a = NavierStokesSimulator((32, 32), 0.01, 0.001, "cpu")
    b = NavierStokesSimulator((32, 32), 0.01, 0.001, "cuda")
    a.add_smoke_source(16, 16, 6, 1.0)
    b.add_smoke_source(16, 16, 6, 1.0)
    d = a.density
    assert d.device.type == "cpu"
    d[4:8, 4:8] += 2.0                                                 # in place on the host copy
    b.density[4:8, 4:8] += 2.0                                         # in place on the live view
    a.u[3, :] = 0.5
    b.u[3, :] = 0.5
    dens = a.density
    dens *= 0.5                                                        # augmented assignment on a fetched copy
    b.density *= 0.5
    fa, fb = a.step(), b.step()
    assert fa.device.type == "cpu"
    assert_same(N(fa), N(fb), "frame after in-place host edits")
    for k in ("u", "v", "p", "density"):
        assert_same(N(getattr(a, k)), N(getattr(b, k)), k)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_simulator_on_another_device_leaves_the_current_device_alone():
    """a simulator on cuda:1 must not change the process's current device (the library's own CUDA runtime and
    torch's share the driver's current context)."""
    torch.cuda.set_device(0)
    sim = SmokeSimulator((64, 64), device="cuda:1")
    sim.add_incense_source([(32, 32)], [1.0])
    f = sim.simulate_step()
    assert torch.cuda.current_device() == 0 and f.device.index == 1
    x = torch.ones(4, device="cuda")                                   # default-device allocation still lands on cuda:0
    assert x.device.index == 0
    ref = SmokeSimulator((64, 64), device="cuda:0")
    ref.add_incense_source([(32, 32)], [1.0])
    assert_same(N(f), N(ref.simulate_step()), "cuda:1 vs cuda:0")


@pytest.mark.parametrize("vel", [40.0, 600.0])
@pytest.mark.parametrize("h,w", [(100, 400), (64, 300), (35, 260), (130, 1000), (48, 392), (300, 140), (129, 129)])
def test_gradient_subtract_fused_into_the_tiled_advection_vs_oracle(h, w, vel):
    """Big grids run k_project inside the u and v advections (k_advect_tiled<.., 1 / 2>: u, v projected in shared memory from
    a staged pressure window).  SMK_PROJECT_FUSED=1 forces that path on small grids with interior AND edge tiles; vel = 600
    makes back-traces of up to 12 cells, which leave the staged window and take the project-on-the-fly global path."""
    rng = np.random.default_rng(h * 7 + w + int(vel))
    ref = oracle.OracleSolver((h, w), 0.02, 0.01, 6)
    ref.u = ((rng.random((h + 1, w)) - 0.5) * vel).astype(np.float32)
    ref.v = ((rng.random((h, w + 1)) - 0.5) * vel).astype(np.float32)
    ref.p = (rng.standard_normal((h, w)) * 50).astype(np.float32)
    ref.density = rng.random((h, w)).astype(np.float32)
    with smk_env(SMK_PROJECT_FUSED=1):
        ns = make(h, w, 0.02, 0.01, 6, step_kernel="phases")
        for k in ("u", "v", "p", "density"):
            setattr(ns, k, T(getattr(ref, k)))
        n0 = _lib.launch_count()
        for t in range(2):
            ref.step()
            ns.step()
            for k in ("u", "v", "p", "density"):
                assert_same(N(getattr(ns, k)), getattr(ref, k), "%dx%d vel %g step %d %s" % (h, w, vel, t, k))
        assert _lib.launch_count() - n0 == 2 * (1 + 1 + 3)          # no k_project launch: fdd, one Jacobi launch, three advections
