#!/usr/bin/env python
"""bench.py -- throughput of the grid smoke step (BASELINE.json metric: grid cell-updates/s and % of the
HBM roofline) on N B200s, with the CPU restatement of the reference timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

Workload (default c2 = BASELINE.json configs[1], the one the metric is quoted on): batched dataset
generation, 256 independent 128x128 sequences per GPU, 40 Jacobi sweeps per time step, 20 time steps per
sequence, emitters drawn by the law of data_loader.py:49-58 with seed 1234 + global sequence index.
One bench "step" = one pass of the hot path over the rank's batch: reset the state, splat the emitters,
run the 20 time steps and write the 20 returned (fractal-scaled) frames of every sequence.
Sequences are independent, so ranks share nothing on the data path (no collective; weak scaling:
256 sequences per GPU).  c3 = one 1024x1024 grid, 100 sweeps per step (single-GPU roofline case).

Printed JSON (one line, rank 0): see the contract in the task statement.  `value` = cell-steps/s of the
whole job with the emitter records already on the device and frames left in HBM; `e2e` = the same through
SmokeSimulator.generate_sequences() from host emitter lists (pinned H2D) to frames in pinned host memory
(D2H), copies inside the timed region.  `roofline` is for the kernel with the largest share of the step,
from CUDA events the library records around every launch (smk_profile_begin/end) in a second pass over
the same steps; `cpu_baseline` is the oracle (a C port of the reference step, NOT the product) on the
host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "grid_cell_updates_per_sec"
UNIT = "cell-steps/s"

WORKLOADS = {
    # name: (h, w, sequences per GPU, jacobi sweeps, time steps per bench step)
    "c2": dict(h=128, w=128, batch=256, K=40, tsteps=20,
               name="c2: batched dataset generation, 256 independent 128x128 sequences per GPU, 40 Jacobi sweeps/step, "
                    "20 steps/sequence (BASELINE.json configs[1])"),
    "c3": dict(h=1024, w=1024, batch=1, K=100, tsteps=20,
               name="c3: single 1024x1024 grid, 100 Jacobi sweeps/step, 20 steps (BASELINE.json configs[2])"),
    "c4": dict(h=8192, w=8192, batch=1, K=20, tsteps=5,
               name="c4: single 8192x8192 grid, 20 Jacobi sweeps/step, 5 steps per bench step, row slabs over the GPUs "
                    "(BASELINE.json configs[3]; strong scaling)"),
}

# algorithmic bytes per cell per launch of each kernel family (fp32, each array touched once)
def phase_bytes_per_cell(K, sweeps_in_launch, steps_in_launch=1):
    return {
        "step_fused": (100 + 12 * K) * steps_in_launch,   # k_step_fused: whole steps, every phase of SURVEY s8d
        "forces_diffuse_div": 24 + 12,      # R(u,v,d) W(u,v,d)  +  divergence R(u,v) W(div), fused in one kernel
        "jacobi": 12 * sweeps_in_launch,    # per sweep R(p,div) W(p)
        "project": 20,                      # R(p,u,v) W(u,v)
        "advect_u": 12, "advect_v": 12,     # R(u,v) W(field)
        "advect_d": 16 + 4,                 # R(u,v,d) W(d) + the returned copy
        "splat": 8,
    }


def emitters_for_sequence(s, h, w):
    """data_loader.py:49-58 with fixed seeds; larger grids get one emitter per 64x64 block."""
    rng = np.random.default_rng(1234 + s)
    out = []
    if h <= 128 and w <= 128:
        n = int(rng.integers(1, 4))
        for _ in range(n):
            x = int(rng.integers(20, w - 20)); y = int(rng.integers(20, h - 20))
            out.append((x, y, 8, float(rng.uniform(0.5, 2.0))))
    else:
        for by in range(h // 64):
            for bx in range(w // 64):
                out.append((int(bx * 64 + rng.integers(8, 56)), int(by * 64 + rng.integers(8, 56)), 8,
                            float(rng.uniform(0.5, 2.0))))
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(kernel, workload):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(workload, {}).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------ CPU (oracle) arm
def cpu_run(wl, nseq, nthreads, seq0=0):
    """Time the oracle (C port of the reference step) on nseq sequences x tsteps steps with nthreads threads."""
    import oracle
    h, w, K, T = wl["h"], wl["w"], wl["K"], wl["tsteps"]
    u = np.zeros((nseq, h + 1, w), np.float32); v = np.zeros((nseq, h, w + 1), np.float32)
    p = np.zeros((nseq, h, w), np.float32); d = np.zeros((nseq, h, w), np.float32)
    for s in range(nseq):
        for x, y, r, i in emitters_for_sequence(seq0 + s, h, w):
            y0, y1, x0, x1 = max(0, y - r), min(h, y + r + 1), max(0, x - r), min(w, x + r + 1)
            sub = np.ascontiguousarray(d[s, y0:y1, x0:x1])          # the splat only touches the emitter's bounding box
            oracle.splat(sub, x - x0, y - y0, r, i)
            d[s, y0:y1, x0:x1] = sub
    fmul = oracle.fractal_mul(h, 0.05) if h == w else None
    t0 = time.perf_counter()
    oracle.run_batch(u, v, p, d, 0.01, 0.001, K, T, fmul=fmul, want_frames=True, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return nseq * h * w * T / dt, dt


def cpu_sample_size(wl):
    """Sequences in the bounded CPU sample: the whole c2 batch (about 15 core-seconds), 2 steps' worth for c3."""
    return wl["batch"]


def reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    nseq = cpu_sample_size(wl)
    wl_cpu = dict(wl)
    if wl["h"] >= 1024:
        wl_cpu["tsteps"] = 2 if wl["h"] < 4096 else 1
    for _ in range(args.warmup):
        cpu_run(wl_cpu, nseq, cores)
    t_tot, cells = 0.0, 0
    for _ in range(args.steps):
        thr, dt = cpu_run(wl_cpu, nseq, cores)
        t_tot += dt
        cells += nseq * wl["h"] * wl["w"] * wl_cpu["tsteps"]
    val = cells / t_tot
    sample = "%d sequences x %d steps of %dx%d, K=%d per bench step" % (nseq, wl_cpu["tsteps"], wl["h"], wl["w"], wl["K"])
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "host": "rank 0 only; oracle/ C port of the reference step (the reference is pure "
                   "Python/PyTorch and publishes no number for this path), pthreads over sequences"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": min(cores, nseq), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


def reference_torch_numbers(dev):
    """Informational: the UNMODIFIED reference NavierStokesSimulator (installed in git-ignored baseline/_ref by
    `pip install --no-deps --target baseline/_ref <copy of /root/reference>`) stepped on the host CPU and, eagerly,
    on this GPU.  128x128, its hard-coded 20 Jacobi sweeps (navier_stokes.py:139), one sequence.  None if absent."""
    path = os.path.join(ROOT, "baseline", "_ref", "src", "physics", "navier_stokes.py")
    if not os.path.exists(path):
        return None
    try:
        import importlib.util
        import torch
        spec = importlib.util.spec_from_file_location("_reference_navier_stokes", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out = {"config": "unmodified reference NavierStokesSimulator.step(), 128x128, K=20 (its literal), 1 sequence"}
        for name, device, nsteps in (("cpu", "cpu", 10), ("cuda_eager", str(dev), 20)):
            sim = mod.NavierStokesSimulator((128, 128), 0.01, 0.001, device)
            sim.add_smoke_source(64, 64, radius=8, intensity=1.5)
            for _ in range(3):
                sim.step()
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(nsteps):
                sim.step()
            if device != "cpu":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out[name] = {"cell_steps_per_s": 128 * 128 * nsteps / dt, "ms_per_step": 1e3 * dt / nsteps,
                         "threads": torch.get_num_threads() if device == "cpu" else None}
        return out
    except Exception as e:           # informational only
        return {"error": repr(e)[:200]}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweeps-per-launch", type=int, default=0)
    ap.add_argument("--halo", type=int, default=0, help="c4: ghost rows per slab side (0: K + 4, one halo exchange per step; "
                    "sweeps-per-launch + 4 is the minimum and exchanges p after every Jacobi launch)")
    ap.add_argument("--step-kernel", default="auto", choices=["auto", "phases", "fused"],
                    help="auto: whole simulation on one SM for grids <= 128x128 (k_step_fused), else one kernel per phase")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return reference_arm(args, wl)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from smokephysai_b200 import SmokeSimulator, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the smoke step has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    h, w, B, K, T = wl["h"], wl["w"], wl["batch"], wl["K"], wl["tsteps"]
    slab_mode = args.workload == "c4"
    if not slab_mode:
        sim = SmokeSimulator((h, w), 0.01, 0.001, dev, jacobi_iters=K, batch=B, sweeps_per_launch=args.sweeps_per_launch,
                             step_kernel=args.step_kernel)
        ns = sim.ns_solver
        L = ns._layout
        ems = [emitters_for_sequence(rank * B + s, h, w) for s in range(B)]
        fmul = sim.fractal_gen.multiplier((h, w), 0.05)
        frames = torch.empty(B, T, h, L.pitch_c, dtype=torch.float32, device=dev)
        src, off, h2d_bytes = ns.upload_sources(ems)
        cells_per_step_rank = B * h * w * T
        cells_per_launch = B * h * w
        working_set = ns._arena.numel() * 4 + frames.numel() * 4
        emitter_lists = [[((x, y), i) for x, y, _, i in lst] for lst in ems]

        replayed = [0, 0]

        def device_step(eager=False):
            ns.setup_grid()
            ns.splat_uploaded(src, off)
            ns.run_steps(T, fmul=fmul, out=frames)

        def e2e_step():
            return sim.generate_sequences(emitter_lists, T, to_host=True)      # returns after the last D2H copy

        d2h_bytes = T * B * h * L.pitch_c * 4
        e2e_api = "SmokeSimulator.generate_sequences(host emitter lists, to_host=True) -> pinned host frames"
        scaling = "weak"
        total_cells_per_step = world * cells_per_step_rank
        parallelism = "sequences sharded over %d GPU(s), no data-path collective" % world
    else:
        from smokephysai_b200.slab import SlabNavierStokes, sweep_split
        Tj = args.sweeps_per_launch or 10
        slab = SlabNavierStokes((h, w), 0.01, 0.001, dev, rank=rank, world=world, jacobi_iters=K, sweeps_per_launch=Tj,
                                halo=args.halo or (K + 4 if world > 1 else None))
        ems = emitters_for_sequence(0, h, w)
        slab.add_sources(ems)
        cells_per_step_rank = (slab.geom.R1 - slab.geom.R0) * w * T
        cells_per_launch = slab.geom.hl * w
        working_set = slab.local._arena.numel() * 4
        h2d_bytes = 16 * len(ems) + 8
        host_d = torch.empty(slab.geom.R1 - slab.geom.R0, w, dtype=torch.float32).pin_memory()

        slab.step()                                   # eager once: NCCL sets up its P2P connections on first use
        graph_note = "eager launches"
        graph = None
        replayed = [0, 0]                             # [graph replays, library launches recorded in the graph]
        if os.environ.get("SMK_BENCH_GRAPH", "0") == "1":   # opt-in: capturing NCCL P2P hung on this stack
            try:
                n_cap = 1 if len(sweep_split(K, Tj)) % 2 == 0 else 2
                if T % n_cap == 0:
                    c0 = _lib.launch_count()
                    graph = slab.capture(n_cap)
                    replayed[1] = _lib.launch_count() - c0
                    graph_note = "CUDA graph of %d step(s) (kernels + NCCL send/recv), replayed" % n_cap
            except Exception as e:                    # capture is an optimisation of the host path only
                graph = None
                graph_note = "eager launches (graph capture failed: %s)" % repr(e)[:120]

        def device_step(eager=False):
            if graph is not None and not eager:
                for _ in range(T // n_cap):
                    graph.replay()
                replayed[0] += T // n_cap
            else:
                for _ in range(T):
                    slab.step()

        def e2e_step():
            slab.setup_grid()
            slab.add_sources(ems)
            for _ in range(T):
                slab.step()
            host_d.copy_(slab.owned("d"), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return host_d.unsqueeze(0)

        d2h_bytes = host_d.numel() * 4
        e2e_api = "SlabNavierStokes: setup_grid + add_sources(host list) + %d x step() + owned density rows -> pinned host" % T
        scaling = "strong"
        total_cells_per_step = h * w * T
        if world == 1:
            parallelism = "single GPU, undecomposed; " + graph_note
        elif slab.single_exchange:
            parallelism = ("row slabs over %d GPU(s), halo %d rows (>= K + 4): ONE NCCL send/recv group per step (u, v, density, p), "
                           "issued from C (smk_nccl_exchange); %s" % (world, slab.halo, graph_note))
        else:
            parallelism = ("row slabs over %d GPU(s), halo %d rows, NCCL send/recv (smk_nccl_exchange) of p after every launch of <= %d "
                           "fused sweeps and of u,v,density once per step; %s" % (world, slab.halo, Tj, graph_note))

    # ---- device-resident throughput -----------------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = _lib.launch_count()
    r0 = replayed[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - n0 + (replayed[0] - r0) * replayed[1]      # kernels inside replayed graphs count too
    value = total_cells_per_step * args.steps / (ms * 1e-3)

    # ---- per-kernel shares (second pass over the same steps, events around every launch) ------------------
    barrier()
    _lib.profile_begin(max_records=launches + 64)
    for _ in range(args.steps):
        device_step(eager=True)          # per-launch events cannot be recorded inside a replayed graph
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    # clocks / throttle reasons were sampled from before the timed region to here: the timed steps and the same
    # steps once more for the per-kernel events, i.e. only while the GPU runs the measured kernels
    clk = clocks.stop() if rank == 0 else None

    # ---- end to end through the public API: host emitter lists -> results in pinned host memory ------------
    for _ in range(2):
        host_frames = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_frames = e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = total_cells_per_step * args.steps / e2e_s
    checksum = float(host_frames[:, -1].double().sum())
    if slab_mode:
        slab.check()

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peak, peak_src = peaks()
    dom = max((k for k in prof if prof[k][1] > 0), key=lambda k: prof[k][0])
    dom_ms, dom_n = prof[dom]
    total_prof_ms = sum(v[0] for v in prof.values())
    sweeps_in_launch = K * T * args.steps / dom_n if dom == "jacobi" else 0
    steps_in_launch = T * args.steps / dom_n if dom == "step_fused" else 1
    bpc = phase_bytes_per_cell(K, sweeps_in_launch, steps_in_launch)
    alg_bytes_per_launch = bpc[dom] * cells_per_launch
    achieved = alg_bytes_per_launch / (dom_ms / dom_n * 1e-3) / 1e9
    kernel_name = {"jacobi": "k_jacobi", "forces_diffuse_div": "k_forces_diffuse_div", "project": "k_project",
                   "advect_u": "k_advect", "advect_v": "k_advect", "advect_d": "k_advect", "splat": "k_splat",
                   "step_fused": "k_step_fused"}.get(dom, dom)
    if dom == "jacobi":
        note = ("achieved > peak is possible: %d sweeps are fused per launch with the pressure tile on-chip, so real DRAM "
                "traffic is below the algorithmic 12 B/cell-sweep (see `traffic`)" % round(sweeps_in_launch))
    elif dom == "step_fused":
        note = ("achieved > peak is possible: %d whole steps run per launch with u, v, density in shared memory and the pressure "
                "in registers, so real DRAM traffic is one state read + write per launch and 4 B/cell-step of frames, far below "
                "the algorithmic %d B/cell-step (see `traffic`); the kernel is bound by FP32 issue, not HBM"
                % (round(steps_in_launch), 100 + 12 * K))
    else:
        note = ""
    roofline = {
        "bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": committed_traffic(kernel_name, args.workload), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": dom_ms / dom_n,
        "launches_timed": dom_n, "share_of_step": dom_ms / total_prof_ms,
        "note": note,
    }
    step_bytes = 100 + 12 * K
    # fp32 adds / multiplies / divides of one cell-step, counted from the reference's expressions (no FMA by the
    # bit-parity contract, so one instruction = one flop): buoyancy 3, diffusion 3 x 7, divergence 4, Jacobi 5 K,
    # gradient subtract 6, advection 3 x 27, decay 1, fractal multiply 2.  Clamps, floors and index math not counted.
    step_flops = 118 + 5 * K
    sm_count, sm_mhz = torch.cuda.get_device_properties(dev).multi_processor_count, (clk or {}).get("sm_mhz") or 1965.0
    fp32_peak = sm_count * 128 * sm_mhz * 1e6 / 1e12            # 128 fp32 lanes per SM, one add or multiply per lane per clock
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "grid": [h, w], "sequences_per_gpu": B, "jacobi_iters": K, "time_steps_per_bench_step": T,
                   "parallelism": parallelism, "step_kernel": ("fused" if (not slab_mode and ns.step_is_fused(T)) else "phases"),
                   "l2": "no flush: working set %.0f MB per GPU (fields + frames) > 126 MB L2" % (working_set / 1e6)},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_s / args.steps, "api": e2e_api,
                "last_frame_checksum": checksum},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_step": {"algorithmic_bytes_per_cell_step": step_bytes, "achieved": value / world * step_bytes / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": value / world * step_bytes / 1e9 / peak, "per": "GPU (owned cells only; ghost-row recompute not counted)"},
        "roofline_fp32": {"bound": "fp32 issue (what actually bounds the on-chip kernels; the HBM roofline above is the contract's)",
                          "flops_per_cell_step": step_flops, "achieved": value / world * step_flops / 1e12, "peak": fp32_peak,
                          "unit": "Tflop/s per GPU, non-FMA fp32", "frac": value / world * step_flops / 1e12 / fp32_peak,
                          "peak_source": "%d SMs x 128 lanes x %.0f MHz (tools/micro/fp32_pipes.cu measures 118 of 128 results/clk/SM)" % (sm_count, sm_mhz)},
        "phases_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if v[1]},
    }
    if not args.no_cpu_baseline:
        out["reference_torch"] = reference_torch_numbers(dev)
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        nseq = cpu_sample_size(wl)
        wl_cpu = dict(wl)
        if h >= 1024:
            wl_cpu["tsteps"] = 2 if h < 4096 else 1
        cpu_run(wl_cpu, min(nseq, cores), cores)                 # warm the threads / page in
        cval, cdt = cpu_run(wl_cpu, nseq, cores)
        out["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": min(cores, nseq), "kind": "port", "seconds": cdt,
                               "sample": "%d sequences x %d steps of %dx%d, K=%d (oracle C port, pthreads over sequences)"
                                         % (nseq, wl_cpu["tsteps"], h, w, K)}
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
