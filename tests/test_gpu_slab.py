"""GPU tests of the row-slab decomposition: every slab of a grid on ONE GPU with in-process halo copies
(LocalGroup) must reproduce the undecomposed CUDA solver -- and therefore the reference -- bit for bit;
with two or more GPUs the same through NCCL send/recv."""
import os
import socket

import numpy as np
import pytest
import torch

import oracle
from helpers import assert_same
from test_slab_host import random_state, reference_run

pytestmark = pytest.mark.gpu
PeerSeqOffset = 64          # PeerExchanger.NCOUNTERS: the uint32 counters sit behind the mailbox regions

from smokephysai_b200 import NavierStokesSimulator, _lib  # noqa: E402
from smokephysai_b200.slab import LocalGroup, SlabNavierStokes  # noqa: E402


def N(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("world,K,T,H,W,halo", [
    (2, 20, 10, 96, 40, None), (3, 13, 4, 90, 133, None), (4, 24, 8, 300, 260, None), (8, 20, 10, 1024, 512, None),
    # deep halos (>= K + 4): one exchange per step, u, v, density and p together
    (2, 20, 10, 96, 40, 28), (3, 9, 4, 150, 133, 14), (8, 20, 10, 1024, 512, 26),
])
def test_local_group_matches_undecomposed(world, K, T, H, W, halo):
    dt, nu, steps = 0.02, 0.01, 3
    st0 = random_state(H, W, seed=world * 7 + K)
    whole = NavierStokesSimulator((H, W), dt, nu, "cuda", jacobi_iters=K)
    grp = LocalGroup((H, W), dt, nu, "cuda", world=world, jacobi_iters=K, sweeps_per_launch=T, halo=halo)
    assert all(s.single_exchange == (halo is not None) for s in grp.slabs)
    assert sum(1 for kind, _ in grp.slabs[0].step_plan() if kind == "x") == (1 if halo is not None else 1 + len(range(0, K, T)))
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        setattr(whole, name, torch.from_numpy(st0[k]).cuda())
        grp.scatter(k, st0[k])
    for _ in range(steps):
        whole.step()
        grp.step()
    grp.check()
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        cols = st0[k].shape[1]
        assert_same(N(grp.gather(k))[:, :cols], N(getattr(whole, name)), "%s world %d" % (k, world))
    if H <= 128:                                              # and against the oracle directly
        want = reference_run(st0, dt, nu, K, steps)
        assert_same(N(grp.gather("d"))[:, :W], want["d"], "density vs oracle")


def test_slab_sources_and_emitter_scenario():
    """Emitters given in global coordinates land on the right slab rows; 5 steps equal the whole-grid run."""
    H = W = 256
    whole = NavierStokesSimulator((H, W), device="cuda", jacobi_iters=20)
    grp = LocalGroup((H, W), device="cuda", world=4, jacobi_iters=20, sweeps_per_launch=10)
    for x, y, i in ((64, 60, 1.5), (130, 128, 1.0), (200, 190, 0.8), (30, 250, 1.2)):
        whole.add_smoke_source(x, y, radius=8, intensity=i)
        grp.add_smoke_source(x, y, radius=8, intensity=i)
    assert_same(N(grp.gather("d"))[:, :W], N(whole.density), "initial density")
    for _ in range(5):
        whole.step()
        grp.step()
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        assert_same(N(grp.gather(k))[:, :N(getattr(whole, name)).shape[1]], N(getattr(whole, name)), k)


def test_overflow_guard_trips_on_huge_velocity():
    H, W = 128, 64
    grp = LocalGroup((H, W), 0.02, 0.01, "cuda", world=2, jacobi_iters=4, sweeps_per_launch=4)      # halo 8
    st0 = random_state(H, W, seed=1, vel=4000.0)                                                       # |dt*v| up to 40 rows
    for k in ("u", "v", "p", "d"):
        grp.scatter(k, st0[k])
    grp.step()
    with pytest.raises(RuntimeError, match="halo"):
        grp.check()


def test_single_rank_slab_is_the_plain_solver():
    H, W = 64, 48
    st0 = random_state(H, W, seed=5)
    one = SlabNavierStokes((H, W), 0.02, 0.01, "cuda", rank=0, world=1, jacobi_iters=10)
    whole = NavierStokesSimulator((H, W), 0.02, 0.01, "cuda", jacobi_iters=10)
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        one.scatter(k, st0[k])
        setattr(whole, name, torch.from_numpy(st0[k]).cuda())
    for _ in range(2):
        one.step()
        whole.step()
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        assert_same(N(one.gather(k)), N(getattr(whole, name)), k)


# --------------------------------------------------------------------------- real NCCL, needs >= 2 GPUs
def _nccl_worker(rank, world, port, H, W, K, T, steps, out_dir, exch):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        st0 = random_state(H, W, seed=21)
        from smokephysai_b200.slab import DistExchanger, NcclExchanger, PeerExchanger
        halo = K + 6 if exch.startswith("peer-one-call") else None
        slab = SlabNavierStokes((H, W), 0.02, 0.01, "cuda:%d" % rank, rank=rank, world=world, jacobi_iters=K, sweeps_per_launch=T, halo=halo,
                                exchanger=DistExchanger() if exch == "torch" else None,
                                exchange="peer" if exch.startswith("peer") else "nccl")
        assert isinstance(slab.exchanger, {"torch": DistExchanger, "library": NcclExchanger}.get(exch, PeerExchanger))
        assert slab.single_exchange == exch.startswith("peer-one-call")
        for k in ("u", "v", "p", "d"):
            slab.scatter(k, st0[k])
        if exch == "peer-one-call-ahead":
            os.environ["SMK_ADVECT_TILED"] = "1"          # the banded density advection with the push in its middle
            _lib.reload_env()
            slab.run_steps(steps - 1)
            slab.step()                                   # a head push again after a run that ended without a tail push
        else:
            for _ in range(steps):
                slab.step()
        slab.check()
        full = {k: slab.gather(k).cpu().numpy() for k in ("u", "v", "p", "d")}
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), **full)
        dist.barrier()
        if hasattr(slab.exchanger, "close"):
            slab.exchanger.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(150)
@pytest.mark.parametrize("exch", ["library", "torch", "peer", "peer-one-call", "peer-one-call-ahead"])
def test_nccl_slabs_match_undecomposed(tmp_path, exch):
    """Real NVLink: halo exchange through the library's own NCCL communicator (smk_nccl_exchange), through torch.distributed
    P2P ops, and by direct peer stores into IPC-mapped mailboxes (smk_peer_push / smk_peer_unpack; "peer-one-call": deep halo,
    the whole step issued by smk_slab_step); every one must reproduce the undecomposed run bit for bit."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    H, W, K, T, steps = 512, 384, 20, 10, (5 if exch == "peer-one-call-ahead" else 3)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_nccl_worker, args=(world, port, H, W, K, T, steps, str(tmp_path), exch), nprocs=world, join=True)
    st0 = random_state(H, W, seed=21)
    whole = NavierStokesSimulator((H, W), 0.02, 0.01, "cuda", jacobi_iters=K)
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        setattr(whole, name, torch.from_numpy(st0[k]).cuda())
    for _ in range(steps):
        whole.step()
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        ref = N(getattr(whole, name))
        assert_same(got[k][:, :ref.shape[1]], ref, "%s over NCCL, world %d" % (k, world))


def test_full_size_c4_slabs_equal_undecomposed():
    """BASELINE config 4 at full size: one 8192 x 8192 grid, K = 20, one emitter per 64 x 64 block.  Eight row slabs
    (in-process halo copies, both exchange schedules) must equal the undecomposed run bit for bit after 2 steps."""
    n, K = 8192, 20
    rng = np.random.default_rng(4)
    src = [(int(bx * 64 + rng.integers(8, 56)), int(by * 64 + rng.integers(8, 56)), 8, float(rng.uniform(0.5, 2.0)))
           for by in range(n // 64) for bx in range(n // 64)]
    whole = NavierStokesSimulator((n, n), device="cuda", jacobi_iters=K)
    whole.add_sources([src])
    for _ in range(2):
        whole.step()
    want = {k: getattr(whole, name).clone() for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density"))}
    assert float(want["u"].abs().max()) > 0 and float(want["p"].abs().max()) > 0
    del whole
    torch.cuda.empty_cache()
    for halo in (None, K + 4):                        # per-launch p exchanges (halo T + 4) and the single-exchange plan
        grp = LocalGroup((n, n), device="cuda", world=8, jacobi_iters=K, sweeps_per_launch=10, halo=halo)
        for s in grp.slabs:
            s.add_sources(src)
        for _ in range(2):
            grp.step()
        grp.check()
        for k in ("u", "v", "p", "d"):
            got = grp.gather(k)
            assert torch.equal(got[:, :want[k].shape[1]], want[k]), "%s differs (halo %r)" % (k, halo)
        del grp
        torch.cuda.empty_cache()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world,H,W,K,T,halo", [(2, 200, 132, 8, 4, None), (3, 300, 260, 12, 6, 18), (4, 512, 384, 20, 10, None),
                                                (8, 1024, 200, 20, 10, 26)])
def test_peer_exchange_kernels_on_one_gpu(world, H, W, K, T, halo):
    """smk_peer_push / smk_peer_unpack with every slab in this process (the "remote" mailboxes are the other slabs' tensors):
    the mailbox addressing, the two-slot parity, the counters and the ghost-row offsets -- u's extra staggered row, v's wider
    pitch -- against the undecomposed run, bit for bit, over enough steps for both slots to be reused several times."""
    st0 = random_state(H, W, seed=33 + world)
    grp = LocalGroup((H, W), 0.02, 0.01, "cuda", world=world, jacobi_iters=K, sweeps_per_launch=T, halo=halo, peer=True)
    whole = NavierStokesSimulator((H, W), 0.02, 0.01, "cuda", jacobi_iters=K)
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        grp.scatter(k, st0[k])
        setattr(whole, name, torch.from_numpy(st0[k]).cuda())
    for _ in range(5):
        grp.step()
        whole.step()
    grp.check()
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        ref = N(getattr(whole, name))
        assert_same(N(grp.gather(k))[:, :ref.shape[1]], ref, "%s, peer kernels, world %d" % (k, world))
    # the counters moved in lockstep: every slab completed the same number of exchanges
    seqs = {int(s.exchanger.buf[-PeerSeqOffset:].view(torch.int32)[2].item()) for s in grp.slabs}
    assert len(seqs) == 1 and seqs.pop() > 0


@pytest.mark.timeout(120)
@pytest.mark.parametrize("tiled,side", [(0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("world,H,W,K,halo", [(2, 200, 132, 8, 14), (3, 300, 260, 12, 18), (4, 512, 384, 20, 26), (8, 1024, 200, 20, 26)])
def test_ghost_rows_pushed_ahead_from_the_density_advection(world, H, W, K, halo, tiled, side):
    """smk_slab_step with SMK_SLAB_PUSH_TAIL: the tile rows holding the rows a slab sends are advected first, the push for the NEXT
    step leaves between them and the rest of the density advection, and the next step starts with the unpack alone.  Tiled
    advection (bands, merged where a slab is only a few tiles tall) and the direct kernel (push after the whole launch), over enough
    steps for both mailbox slots to be reused, against the undecomposed run bit for bit; then two more plain steps on the same
    slabs (a head push after a run that ended without a tail push).  side: the push on the library's high-priority side stream
    (default) or on the caller's stream (SMK_PUSH_STREAM=0)."""
    from helpers import smk_env
    st0 = random_state(H, W, seed=70 + world)
    with smk_env(SMK_ADVECT_TILED=tiled, SMK_PUSH_STREAM=side):
        grp = LocalGroup((H, W), 0.02, 0.01, "cuda", world=world, jacobi_iters=K, sweeps_per_launch=max(K // 2, 1), halo=halo, peer=True)
        assert all(s.single_exchange for s in grp.slabs)
        whole = NavierStokesSimulator((H, W), 0.02, 0.01, "cuda", jacobi_iters=K, step_kernel="phases")
        for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
            grp.scatter(k, st0[k])
            setattr(whole, name, torch.from_numpy(st0[k]).cuda())
        _lib.profile_begin(max_records=4096)
        grp.run_steps(5)
        torch.cuda.synchronize()
        prof = _lib.profile_end()
        for _ in range(5):
            whole.step()
        grp.check()
        for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
            ref = N(getattr(whole, name))
            assert_same(N(grp.gather(k))[:, :ref.shape[1]], ref, "%s, pushed ahead, world %d, tiled %d" % (k, world, tiled))
        # one push and one unpack per slab and step; with the tiled kernel the density advection of a push-ahead step is
        # bands + rest (2 or 3 launches) instead of 1
        assert prof["halo"][1] == 5 * world and prof["halo_unpack"][1] == 5 * world
        assert prof["advect_d"][1] == 5 * world if not tiled else prof["advect_d"][1] >= 4 * world * 2 + world
        grp.run_steps(2)
        for _ in range(2):
            whole.step()
        for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
            ref = N(getattr(whole, name))
            assert_same(N(grp.gather(k))[:, :ref.shape[1]], ref, "%s, second run, world %d, tiled %d" % (k, world, tiled))


@pytest.mark.timeout(120)
@pytest.mark.parametrize("tiled", [0, 1])
def test_pushed_ahead_steps_replay_from_a_cuda_graph(tiled):
    """run_steps inside a stream capture: the side stream of the tail push forks from and joins the capturing stream, the kernels
    read the exchange number from device memory, programmatic dependent launch is off inside the capture.  Four captured steps
    (the ping-pong state returns to the same buffers), replayed twice, against the undecomposed run."""
    from helpers import smk_env
    world, H, W, K, halo = 3, 300, 260, 12, 18
    st0 = random_state(H, W, seed=91)
    with smk_env(SMK_ADVECT_TILED=tiled):
        grp = LocalGroup((H, W), 0.02, 0.01, "cuda", world=world, jacobi_iters=K, sweeps_per_launch=6, halo=halo, peer=True)
        whole = NavierStokesSimulator((H, W), 0.02, 0.01, "cuda", jacobi_iters=K, step_kernel="phases")
        for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
            grp.scatter(k, st0[k])
            setattr(whole, name, torch.from_numpy(st0[k]).cuda())
        grp.run_steps(2)                                  # eager first: kernels loaded, side stream and events created
        torch.cuda.synchronize()
        before = [(s.local._state.cur_u, s.local._state.cur_v, s.local._state.cur_d, s.local._state.cur_p) for s in grp.slabs]
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            grp.run_steps(4)
        assert before == [(s.local._state.cur_u, s.local._state.cur_v, s.local._state.cur_d, s.local._state.cur_p) for s in grp.slabs]
        graph.replay()
        graph.replay()
        torch.cuda.synchronize()
        for _ in range(2 + 8):
            whole.step()
        grp.check()
        for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
            ref = N(getattr(whole, name))
            assert_same(N(grp.gather(k))[:, :ref.shape[1]], ref, "%s after two replays of a four-step graph, tiled %d" % (k, tiled))


def test_fields_cannot_change_while_a_push_is_in_flight():
    a = SlabNavierStokes((300, 260), 0.02, 0.01, "cuda", rank=0, world=1, jacobi_iters=12, sweeps_per_launch=6)
    a._pushed_ahead = True
    for call in (lambda: a.add_smoke_source(10, 10), lambda: a.add_sources([(10, 10, 3, 1.0)]), a.setup_grid,
                 lambda: a.scatter("u", np.zeros((301, 260), np.float32))):
        with pytest.raises(RuntimeError, match="push_ahead"):
            call()


def test_slab_step_in_one_c_call_equals_the_phase_calls():
    """smk_slab_step (one C call per step) against the per-phase entry points on an undecomposed grid."""
    H, W, K = 300, 260, 12
    st0 = random_state(H, W, seed=8)
    a = SlabNavierStokes((H, W), 0.02, 0.01, "cuda", rank=0, world=1, jacobi_iters=K, sweeps_per_launch=6)
    b = NavierStokesSimulator((H, W), 0.02, 0.01, "cuda", jacobi_iters=K, sweeps_per_launch=6, step_kernel="phases")
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        a.scatter(k, st0[k])
        setattr(b, name, torch.from_numpy(st0[k]).cuda())
    n0 = _lib.launch_count()
    for _ in range(3):
        a.step()
    assert _lib.launch_count() - n0 == 3 * (1 + 2 + 1 + 3)
    for _ in range(3):
        b.step()
    for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
        ref = N(getattr(b, name))
        assert_same(N(a.owned(k))[:, :ref.shape[1]], ref, k)


@pytest.mark.parametrize("world,H,W,K,T,halo", [(2, 200, 132, 8, 4, None), (3, 300, 260, 12, 6, 18), (4, 512, 384, 20, 10, 26)])
def test_slabs_with_the_fused_gradient_subtract(world, H, W, K, T, halo):
    """The slab step with k_project fused into the tiled u / v advections (forced on these small grids) against the undecomposed
    run with k_project as its own kernel."""
    from helpers import smk_env
    st0 = random_state(H, W, seed=5 + world)
    with smk_env(SMK_PROJECT_FUSED=0):
        whole = NavierStokesSimulator((H, W), 0.02, 0.01, "cuda", jacobi_iters=K, step_kernel="phases")
        for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
            setattr(whole, name, torch.from_numpy(st0[k]).cuda())
        for _ in range(3):
            whole.step()
    with smk_env(SMK_PROJECT_FUSED=1):
        grp = LocalGroup((H, W), 0.02, 0.01, "cuda", world=world, jacobi_iters=K, sweeps_per_launch=T, halo=halo)
        for k in ("u", "v", "p", "d"):
            grp.scatter(k, st0[k])
        for _ in range(3):
            grp.step()
        grp.check()
        for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("d", "density")):
            ref = N(getattr(whole, name))
            assert_same(N(grp.gather(k))[:, :ref.shape[1]], ref, "%s, fused gradient subtract in slabs, world %d" % (k, world))
