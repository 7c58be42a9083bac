"""GPU parity of the fused whole-step kernel (csrc/fused.cu: the simulation on one SM, n steps per launch)
against the CPU oracle and against the phase-per-kernel path, through the same class surface / C ABI.

Tolerance: none -- every field must be numerically equal (fp32, no FMA, same association).
"""
import numpy as np
import pytest
import torch

import oracle
from helpers import assert_same, emitters_for_sequence, smk_env

pytestmark = pytest.mark.gpu

from smokephysai_b200 import NavierStokesSimulator, SmokeSimulator, _lib  # noqa: E402

FIELDS = ("u", "v", "p", "density")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def N(t):
    return t.detach().cpu().numpy()


def make(h, w, dt=0.01, nu=0.001, K=20, **kw):
    return NavierStokesSimulator((h, w), dt, nu, "cuda", jacobi_iters=K, **kw)


def random_state(h, w, seed, vel=300.0):
    rng = np.random.default_rng(seed)
    return {
        "u": ((rng.random((h + 1, w)) - 0.5) * vel).astype(np.float32),
        "v": ((rng.random((h, w + 1)) - 0.5) * vel).astype(np.float32),
        "p": rng.standard_normal((h, w)).astype(np.float32),
        "density": rng.random((h, w)).astype(np.float32),
    }


def test_dispatch_rules():
    """auto on 128 x 128: up to 33 simulations run on clusters of four CTAs each (k_step_cluster) whatever the call; more than
    that need enough simulations for one-simulation-per-SM execution to win -- 12 for a multi-step call, 48 for a single step;
    smaller grids from 96 x 96 need 96; everything else runs on the phase kernels spread over all SMs."""
    assert make(128, 128).step_is_fused() and make(128, 128).step_is_fused(nsteps=20)       # one simulation: a cluster of four SMs
    assert make(128, 128, batch=8).step_is_fused(nsteps=20) and make(128, 128, batch=33).step_is_fused()
    assert make(128, 128, batch=34).step_is_fused(nsteps=2) and not make(128, 128, batch=34).step_is_fused()
    assert make(128, 128, batch=48).step_is_fused() and not make(128, 128, batch=47).step_is_fused()
    with smk_env(SMK_FUSED_CLUSTER=0):                     # without the cluster kernel: the round-1 rules
        assert not make(128, 128).step_is_fused() and not make(128, 128, batch=8).step_is_fused(nsteps=20)
        assert make(128, 128, batch=12).step_is_fused(nsteps=2) and not make(128, 128, batch=32).step_is_fused()
    assert not make(96, 40, batch=256).step_is_fused(20) and not make(96, 96, batch=32).step_is_fused(20)
    assert make(96, 96, batch=96).step_is_fused(20) and make(100, 120, batch=148).step_is_fused()
    assert not make(100, 120, batch=3).step_is_fused()
    assert make(128, 128, step_kernel="fused").step_is_fused() and make(2, 2, step_kernel="fused").step_is_fused()
    assert not make(128, 128, batch=64, step_kernel="phases").step_is_fused(20)
    for big in ((129, 128), (128, 132), (1, 64)):
        assert not make(*big, batch=64).step_is_fused(20)
    with pytest.raises(_lib.SmokeLibraryError, match="128 x 128"):
        make(256, 256, step_kernel="fused").step()
    with pytest.raises(ValueError):
        make(64, 64, step_kernel="tensor-cores")
    # the dispatch is observable in the launch count: 6 phase kernels per step against 1 fused launch per call
    ns = make(100, 128)
    n0 = _lib.launch_count(); ns.step(); n1 = _lib.launch_count(); ns.run_steps(3); n2 = _lib.launch_count()
    assert (n1 - n0, n2 - n1) == (6, 18)
    ns = make(128, 128, batch=40)
    n0 = _lib.launch_count(); ns.step(); n1 = _lib.launch_count(); ns.run_steps(3); n2 = _lib.launch_count()
    assert (n1 - n0, n2 - n1) == (6, 1)
    ns = make(128, 128, batch=16)
    n0 = _lib.launch_count(); ns.step(); n1 = _lib.launch_count(); ns.run_steps(3); n2 = _lib.launch_count()
    assert (n1 - n0, n2 - n1) == (1, 1)


@pytest.mark.parametrize("h,w,K", [
    (128, 128, 20), (128, 128, 0), (128, 128, 1), (128, 128, 7), (2, 2, 5), (3, 3, 4), (5, 128, 9), (128, 5, 9), (127, 127, 12),
    (128, 96, 10), (96, 128, 10), (128, 127, 6), (127, 128, 6), (64, 64, 20), (33, 100, 11), (100, 33, 11), (8, 8, 3), (9, 4, 3),
])
def test_fused_step_vs_oracle_and_phases(h, w, K):
    """Random fields with back-traces of several cells (|dt*vel| up to 3) so gathers cross warps and hit the clamps."""
    st = random_state(h, w, h * 1000 + w)
    ref = oracle.OracleSolver((h, w), 0.02, 0.01, K)
    fz = make(h, w, 0.02, 0.01, K, step_kernel="fused")
    ph = make(h, w, 0.02, 0.01, K, step_kernel="phases")
    assert fz.step_is_fused() and not ph.step_is_fused()
    for k in FIELDS:
        setattr(ref, k, st[k].copy())
        setattr(fz, k, T(st[k]))
        setattr(ph, k, T(st[k]))
    for t in range(3):
        fr_ref = ref.step()
        fr_fz = fz.step()
        fr_ph = ph.step()
        for k in FIELDS:
            assert_same(N(getattr(fz, k)), getattr(ref, k), "%dx%d K=%d step %d fused vs oracle %s" % (h, w, K, t, k))
            assert_same(N(getattr(fz, k)), N(getattr(ph, k)), "%dx%d K=%d step %d fused vs phases %s" % (h, w, K, t, k))
        assert_same(N(fr_fz), fr_ref, "returned frame vs oracle")
        assert_same(N(fr_fz), N(fr_ph), "returned frame vs phases")


@pytest.mark.parametrize("h,w", [(128, 128), (60, 128), (128, 52)])
def test_fused_multi_step_launch_equals_single_steps(h, w):
    """run_steps(n) is ONE launch that keeps the state on chip; it must equal n single-step launches, the phase
    path and the oracle -- frames, final fields and the padding columns of the arena."""
    B, n, K = 3, 7, 15
    sims = {k: make(h, w, 0.02, 0.005, K, batch=B, step_kernel=k) for k in ("fused", "phases")}
    one = make(h, w, 0.02, 0.005, K, batch=B, step_kernel="fused")
    states = [random_state(h, w, 77 + b, vel=120.0) for b in range(B)]
    for s in list(sims.values()) + [one]:
        for k in FIELDS:
            setattr(s, k, T(np.stack([st[k] for st in states])))
    fmul = torch.rand(h, sims["fused"]._layout.pitch_c, device="cuda") * 0.05
    n0 = _lib.launch_count()
    fr_fz = sims["fused"].run_steps(n, fmul=fmul)
    assert _lib.launch_count() - n0 == 1
    fr_ph = sims["phases"].run_steps(n, fmul=fmul)
    fr_one = torch.stack([one.step(fmul=fmul) for _ in range(n)], dim=1)
    assert_same(N(fr_fz), N(fr_ph), "frames fused vs phases")
    assert_same(N(fr_fz), N(fr_one), "frames one launch vs n launches")
    for k in FIELDS:
        assert_same(N(getattr(sims["fused"], k)), N(getattr(sims["phases"], k)), k + " fused vs phases")
        assert_same(N(getattr(sims["fused"], k)), N(getattr(one, k)), k + " one launch vs n launches")
    # the oracle on simulation 1 (no fractal multiplier on its frames: compare fields only)
    ref = oracle.OracleSolver((h, w), 0.02, 0.005, K)
    for k in FIELDS:
        setattr(ref, k, states[1][k].copy())
    for _ in range(n):
        ref.step()
    for k in FIELDS:
        assert_same(N(getattr(sims["fused"], k))[1], getattr(ref, k), k + " vs oracle")
    # padding columns of the live arena fields stay zero (the float4 kernels rely on it)
    fz = sims["fused"]
    L = fz._layout
    for name in ("u", "v", "d", "p"):
        cur = getattr(fz._state, "cur_" + name)
        rows, cols, pitch = L.shape_of("%s%d" % (name, cur))
        off = L.offset["%s%d" % (name, cur)]
        full = fz._arena[off: off + L.batch * rows * pitch].view(L.batch, rows, pitch)
        assert not full[:, :, cols:].any(), "padding of %s" % name


@pytest.mark.parametrize("h,w,B,n,seg", [(128, 128, 5, 7, 3), (128, 128, 5, 7, 9), (128, 128, 4, 6, 1), (60, 128, 6, 5, 4), (128, 52, 3, 8, 5),
                                         (128, 128, 7, 4, 28)])
def test_fused_time_sliced_schedule_equals_one_cta_per_simulation(h, w, B, n, seg):
    """The time-sliced schedule of k_step_fused (a CTA runs `seg` consecutive simulation-steps of the simulation-major
    line and hands a simulation cut by its piece boundary over to the next CTA through global memory) must give the
    frames and fields of the one-CTA-per-simulation schedule bit for bit.  SMK_FUSED_SLICE forces the piece length:
    seg < n chains a simulation through several CTAs, seg > n packs several simulations into one CTA."""
    K = 12
    states = [random_state(h, w, 500 + b, vel=150.0) for b in range(B)]
    out = {}
    for mode in ("0", str(seg)):
        with smk_env(SMK_FUSED_SLICE=mode):
            ns = make(h, w, 0.02, 0.004, K, batch=B, step_kernel="fused")
            for k in FIELDS:
                setattr(ns, k, T(np.stack([st[k] for st in states])))
            fmul = torch.linspace(0.0, 0.05, h * ns._layout.pitch_c, device="cuda").view(h, ns._layout.pitch_c)
            n0 = _lib.launch_count()
            frames = ns.run_steps(n, fmul=fmul)
            assert _lib.launch_count() - n0 == 1
            out[mode] = (N(frames), {k: N(getattr(ns, k)) for k in FIELDS})
    assert_same(out[str(seg)][0], out["0"][0], "frames, sliced vs classic")
    for k in FIELDS:
        assert_same(out[str(seg)][1][k], out["0"][1][k], k + ", sliced vs classic")
    ref = oracle.OracleSolver((h, w), 0.02, 0.004, K)
    for k in FIELDS:
        setattr(ref, k, states[B - 1][k].copy())
    for _ in range(n):
        ref.step()
    for k in FIELDS:
        assert_same(out[str(seg)][1][k][B - 1], getattr(ref, k), k + " vs oracle")


def test_fused_time_major_frames_and_generate_sequences():
    B, L = 6, 9
    ems = [[((x, y), i) for x, y, _, i in emitters_for_sequence(s)] for s in range(B)]
    fz = SmokeSimulator((128, 128), device="cuda", batch=B, jacobi_iters=40, step_kernel="fused")
    ph = SmokeSimulator((128, 128), device="cuda", batch=B, jacobi_iters=40, step_kernel="phases")
    a = N(fz.generate_sequences(ems, L))
    b = N(ph.generate_sequences(ems, L))
    assert_same(a, b, "generate_sequences fused vs phases")
    for spc in (None, 1, 4, 9, 20):
        h_ = fz.generate_sequences(ems, L, to_host=True, steps_per_copy=spc)
        assert h_.is_pinned() and tuple(h_.shape) == (B, L, 128, 128)
        assert_same(h_.numpy(), a, "to_host, steps_per_copy=%r" % spc)


def test_fused_c2_full_size_vs_oracle():
    """BASELINE config 2 at full size: 256 x 128^2, K=40, 20 steps in ONE launch; six sequences against the oracle."""
    B, K, steps = 256, 40, 20
    ns = make(128, 128, K=K, batch=B, step_kernel="fused")
    ns.add_sources([emitters_for_sequence(s) for s in range(B)])
    d0 = N(ns.density).copy()
    frames = ns.run_steps(steps)
    u, v, p, d = (N(getattr(ns, k)) for k in FIELDS)
    sel = [0, 1, 147, 148, 200, 255]
    ou = np.zeros((len(sel), 129, 128), np.float32); ov = np.zeros((len(sel), 128, 129), np.float32)
    op = np.zeros((len(sel), 128, 128), np.float32); od = d0[sel].copy()
    ofr = oracle.run_batch(ou, ov, op, od, 0.01, 0.001, K, steps, nthreads=4)
    for n_, s in enumerate(sel):
        assert_same(u[s], ou[n_], "u seq %d" % s); assert_same(v[s], ov[n_], "v seq %d" % s)
        assert_same(p[s], op[n_], "p seq %d" % s); assert_same(d[s], od[n_], "d seq %d" % s)
        assert_same(N(frames[s]), ofr[n_], "frames seq %d" % s)
    assert not d[:, -1, :].any() and not d[:, :, -1].any()


@pytest.mark.parametrize("scale", [1e-44, 1e-41, 1e-38, 1e-33, 1e33])
def test_fused_divergence_extreme_magnitudes(scale):
    """The /dt of the divergence (navier_stokes.py:136) on denormal and huge numerators: the fused kernel routes those
    warps through an fp64 division, which must round exactly like the IEEE fp32 division of the oracle / reference."""
    h = w = 128
    rng = np.random.default_rng(int(-np.log10(scale)) + 100)
    st = {
        "u": (rng.standard_normal((h + 1, w)) * scale).astype(np.float32),
        "v": (rng.standard_normal((h, w + 1)) * scale).astype(np.float32),
        "p": np.zeros((h, w), np.float32),
        "density": np.zeros((h, w), np.float32),
    }
    # a patch of ordinary values so that warps mix ordinary and extreme lanes
    st["u"][40:60, 30:90] = rng.standard_normal((20, 60)).astype(np.float32)
    ref = oracle.OracleSolver((h, w), 0.01, 0.001, 3)
    fz = make(h, w, K=3, step_kernel="fused")
    ph = make(h, w, K=3, step_kernel="phases")
    for k in FIELDS:
        setattr(ref, k, st[k].copy())
        setattr(fz, k, T(st[k]))
        setattr(ph, k, T(st[k]))
    for t in range(2):
        ref.step()
        fz.step()
        ph.step()
        for k in FIELDS:
            a, b, c = N(getattr(fz, k)), getattr(ref, k), N(getattr(ph, k))
            assert np.array_equal(a, b, equal_nan=True), "scale %g step %d %s fused vs oracle" % (scale, t, k)
            assert np.array_equal(c, b, equal_nan=True), "scale %g step %d %s phases vs oracle" % (scale, t, k)


def test_fused_time_sliced_hand_over_with_a_co_resident_kernel():
    """The hand-over of the time-sliced schedule may only ever wait for a CTA that has already started: items are claimed
    by ticket at CTA entry (fused.cu), so the guarantee does not depend on the order the hardware dispatches the grid in.
    A long-running kernel on a second stream takes SMs away while run_steps executes (CTAs then start late and out of
    their planned order); frames and fields must still equal the one-CTA-per-simulation schedule bit for bit."""
    h = w = 128
    B, n, K = 40, 6, 8
    states = [random_state(h, w, 900 + b, vel=100.0) for b in range(B)]
    out = {}
    side = torch.cuda.Stream()
    for mode in ("0", "2"):
        with smk_env(SMK_FUSED_SLICE=mode):
            ns = make(h, w, 0.02, 0.004, K, batch=B, step_kernel="fused")
            for k in FIELDS:
                setattr(ns, k, T(np.stack([st[k] for st in states])))
            torch.cuda.synchronize()
            if mode != "0":
                with torch.cuda.stream(side):
                    torch.cuda._sleep(int(2e8))                       # ~0.1 s spin kernel, co-resident with the step kernel
                    x = torch.randn(4096, 4096, device="cuda")
                    for _ in range(8):
                        x = (x @ x).clamp_(-1, 1)                    # and a stream of full-grid kernels competing for SMs
            frames = ns.run_steps(n)
            torch.cuda.synchronize()
            out[mode] = (N(frames), {k: N(getattr(ns, k)) for k in FIELDS})
    assert_same(out["2"][0], out["0"][0], "frames, sliced under contention vs classic")
    for k in FIELDS:
        assert_same(out["2"][1][k], out["0"][1][k], k + ", sliced under contention vs classic")


@pytest.mark.timeout(120)
@pytest.mark.parametrize("nc", [2, 4])
@pytest.mark.parametrize("B,n,K,vel", [(1, 3, 20, 40.0), (3, 4, 7, 900.0), (5, 2, 40, 150.0)])
def test_cluster_kernel_equals_oracle_and_one_cta_kernel(nc, B, n, K, vel):
    """k_step_cluster: one 128 x 128 simulation on a cluster of 2 or 4 CTAs (rows split over the CTAs, halo rows and the Jacobi
    boundary rows through distributed shared memory).  Frames and fields must equal the one-CTA-per-simulation kernel and the
    oracle bit for bit; vel = 900 with dt = 0.02 back-traces up to 9 cells, past the two halo rows a CTA keeps, so those
    gathers read the owner CTA's shared memory; K = 7 ends on the odd-sweep tail."""
    h = w = 128
    states = [random_state(h, w, 700 + 10 * nc + b, vel=vel) for b in range(B)]
    out = {}
    for mode in ("0", str(nc)):
        with smk_env(SMK_FUSED_CLUSTER=mode, SMK_FUSED_SLICE=0):
            ns = make(h, w, 0.02, 0.004, K, batch=B, step_kernel="fused")
            for k in FIELDS:
                setattr(ns, k, T(np.stack([st[k] for st in states])) if B > 1 else T(states[0][k]))
            fmul = torch.linspace(0.0, 0.05, h * ns._layout.pitch_c, device="cuda").view(h, ns._layout.pitch_c)
            n0 = _lib.launch_count()
            frames = ns.run_steps(n, fmul=fmul)
            assert _lib.launch_count() - n0 == 1
            torch.cuda.synchronize()
            out[mode] = (N(frames), {k: N(getattr(ns, k)) for k in FIELDS})
    assert_same(out[str(nc)][0], out["0"][0], "frames, cluster of %d vs one CTA" % nc)
    for k in FIELDS:
        assert_same(out[str(nc)][1][k], out["0"][1][k], k + ", cluster of %d vs one CTA" % nc)
    ref = oracle.OracleSolver((h, w), 0.02, 0.004, K)
    for k in FIELDS:
        setattr(ref, k, states[B - 1][k].copy())
    for _ in range(n):
        ref.step()
    for k in FIELDS:
        got = out[str(nc)][1][k]
        assert_same(got[B - 1] if B > 1 else got, getattr(ref, k), k + " vs oracle")
