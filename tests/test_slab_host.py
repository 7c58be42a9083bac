"""CPU tests of the row-slab decomposition (smokephysai_b200/slab.py): the exchange plan, the halo-depth
arithmetic, and a world_size-2 run over torch.distributed/gloo with the oracle as the compute backend
(tests/slab_oracle.py), bit-compared with the undecomposed oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from slab_oracle import OracleSlab
from smokephysai_b200.slab import DistExchanger, SlabGeometry, local_exchange, sweep_split


def random_state(H, W, seed, vel=300.0):
    rng = np.random.default_rng(seed)
    return {"u": ((rng.random((H + 1, W)) - 0.5) * vel).astype(np.float32),
            "v": ((rng.random((H, W + 1)) - 0.5) * vel).astype(np.float32),
            "p": rng.standard_normal((H, W)).astype(np.float32),
            "d": rng.random((H, W)).astype(np.float32)}


def reference_run(state, dt, nu, K, steps):
    H, W = state["d"].shape
    ref = oracle.OracleSolver((H, W), dt, nu, K)
    ref.u, ref.v, ref.p, ref.density = (state[k].copy() for k in ("u", "v", "p", "d"))
    for _ in range(steps):
        ref.step()
    return {"u": ref.u, "v": ref.v, "p": ref.p, "d": ref.density}


def test_sweep_split():
    assert sweep_split(100, 12) == [12, 11, 11, 11, 11, 11, 11, 11, 11]
    assert sweep_split(20, 10) == [10, 10] and sweep_split(20, 24) == [20] and sweep_split(7, 1) == [1] * 7
    assert sum(sweep_split(37, 5)) == 37 and max(sweep_split(37, 5)) <= 5


@pytest.mark.parametrize("H,world,halo", [(64, 2, 6), (100, 3, 9), (8192, 8, 14), (130, 4, 14), (97, 5, 3)])
def test_geometry_plan_is_consistent(H, world, halo):
    geoms = [SlabGeometry(H, 32, world, r, halo) for r in range(world)]
    assert geoms[0].R0 == 0 and geoms[-1].R1 == H
    for a, b in zip(geoms, geoms[1:]):
        assert a.R1 == b.R0
    for kind in ("u", "c"):
        covered = np.zeros(H + (1 if kind == "u" else 0), int)
        for g in geoms:
            lo, hi = g.owned_rows(kind)
            covered[g.A + lo: g.A + hi] += 1
            sends, recvs = g.blocks(kind)
            for peer, a, n in recvs:                 # every receive has the matching send on the peer, same size,
                ps, _ = geoms[peer].blocks(kind)     # of the same GLOBAL rows, all owned by the sender
                b, m = next((b, m) for p, b, m in ps if p == g.rank)
                assert m == n and geoms[peer].A + b == g.A + a
                plo, phi = geoms[peer].owned_rows(kind)
                assert plo <= b and b + m <= phi + (1 if kind == "u" else 0)
                # received rows are ghost rows: outside what this rank owns
                assert a + n <= lo or a >= hi
        assert (covered == 1).all()


def test_geometry_rejects_thin_slabs():
    with pytest.raises(ValueError):
        SlabGeometry(64, 32, 8, 0, 8)


@pytest.mark.parametrize("world,K,T,H,W,halo", [
    (2, 20, 10, 96, 40, None), (3, 13, 4, 90, 33, None), (4, 9, 1, 64, 24, None),
    # deep halos (>= K + 4): ONE exchange per step, u, v, density and p together
    (2, 20, 10, 96, 40, 24), (3, 9, 4, 90, 33, 13), (2, 6, 6, 64, 24, 12),
])
def test_local_emulation_matches_undecomposed_oracle(world, K, T, H, W, halo):
    """All slabs in one process (in-process halo copies): owned rows == the whole-grid oracle, bit for bit."""
    dt, nu, steps = 0.02, 0.01, 3
    st0 = random_state(H, W, seed=world * 100 + K)
    want = reference_run(st0, dt, nu, K, steps)
    slabs = [OracleSlab((H, W), dt, nu, r, world, K, T, halo=halo) for r in range(world)]
    n_exchanges = sum(1 for kind, _ in slabs[0].step_plan() if kind == "x")
    assert n_exchanges == (1 if halo is not None else 1 + len(sweep_split(K, T)))
    for s in slabs:
        for k in ("u", "v", "p", "d"):
            s.scatter(k, st0[k])
    for _ in range(steps):
        plans = [s.step_plan() for s in slabs]
        for i in range(len(plans[0])):
            if plans[0][i][0] == "x":
                local_exchange([s.geom for s in slabs], [s.exchange_list(plans[0][i][1]) for s in slabs])
            else:
                for p in plans:
                    p[i][1]()
    for k in ("u", "v", "p", "d"):
        got = np.concatenate([s.owned(k) for s in slabs], axis=0)
        assert np.array_equal(got, want[k]), "%s differs (world %d, K %d, T %d)" % (k, world, K, T)


def test_too_shallow_halo_is_not_exact():
    """Sanity of the erosion argument: with a halo below T + 3 the decomposed run must not be trusted
    (and SlabNavierStokes refuses it); here it really does differ."""
    H, W, K, T, dt, nu = 64, 24, 12, 12, 0.02, 0.01
    st0 = random_state(H, W, seed=3)
    want = reference_run(st0, dt, nu, K, 2)
    slabs = [OracleSlab((H, W), dt, nu, r, 2, K, T, halo=5) for r in range(2)]
    for s in slabs:
        for k in ("u", "v", "p", "d"):
            s.scatter(k, st0[k])
    for _ in range(2):
        plans = [s.step_plan() for s in slabs]
        for i in range(len(plans[0])):
            if plans[0][i][0] == "x":
                local_exchange([s.geom for s in slabs], [s.exchange_list(plans[0][i][1]) for s in slabs])
            else:
                for p in plans:
                    p[i][1]()
    got = np.concatenate([s.owned("p") for s in slabs], axis=0)
    assert not np.array_equal(got, want["p"])


# ------------------------------------------------------------------------------- world_size 2 over gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, H, W, K, T, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dt, nu = 0.02, 0.01
        st0 = random_state(H, W, seed=11)
        slab = OracleSlab((H, W), dt, nu, rank, world, K, T)
        for k in ("u", "v", "p", "d"):
            slab.scatter(k, st0[k])
        ex = DistExchanger()
        for _ in range(steps):
            for kind, arg in slab.step_plan():
                if kind == "x":
                    ex.exchange(slab.geom, slab.exchange_list(arg))
                else:
                    arg()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{k: slab.owned(k) for k in ("u", "v", "p", "d")})
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world2_gloo_slabs_match_undecomposed_oracle(tmp_path):
    H, W, K, T, steps, world = 80, 36, 20, 10, 3, 2
    mp.spawn(_gloo_worker, args=(world, _free_port(), H, W, K, T, steps, str(tmp_path)), nprocs=world, join=True)
    want = reference_run(random_state(H, W, seed=11), 0.02, 0.01, K, steps)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for k in ("u", "v", "p", "d"):
        got = np.concatenate([p[k] for p in parts], axis=0)
        assert np.array_equal(got, want[k]), k


@pytest.mark.parametrize("H,W,world,halo", [(64, 40, 2, 6), (100, 33, 3, 9), (256, 130, 4, 14), (1024, 20, 8, 24)])
def test_peer_mailbox_arithmetic_equals_the_plain_halo_copy(H, W, world, halo):
    """The offsets / counts PeerExchanger hands to smk_peer_push / smk_peer_unpack (slab.peer_links), driven through a numpy
    model of the mailboxes (two parity slots per neighbour, four field regions per slot, a rank writes into the region its
    neighbour reserves for the OPPOSITE side), must move exactly the rows local_exchange moves -- u's extra staggered row and
    v's wider pitch included -- over several exchanges so that both parity slots are reused."""
    from smokephysai_b200.navier_stokes import FieldLayout
    from smokephysai_b200.slab import FIELD_ORDER, peer_links
    geoms = [SlabGeometry(H, W, world, r, halo) for r in range(world)]
    lay = [FieldLayout(g.hl, W) for g in geoms]
    pitch = {"u": lay[0].pitch_u, "v": lay[0].pitch_v, "d": lay[0].pitch_c, "p": lay[0].pitch_c}
    field_stride = (halo + 1) * max(pitch.values())
    links = [peer_links(g, pitch) for g in geoms]
    rng = np.random.default_rng(world)
    rows = lambda g, n: g.hl + (1 if n == "u" else 0)
    mail = [np.full((2, 2, 4, field_stride), np.nan, np.float32) for _ in range(world)]       # [side][slot][field][elements]
    for seq in range(5):
        flat = [{n: rng.standard_normal(rows(g, n) * pitch[n]).astype(np.float32) for n in FIELD_ORDER} for g in geoms]
        want = [{n: torch.from_numpy(f[n].copy()).view(rows(g, n), pitch[n]) for n in FIELD_ORDER} for g, f in zip(geoms, flat)]
        local_exchange(geoms, [[(w_[n], "u" if n == "u" else "c") for n in FIELD_ORDER] for w_ in want])
        for r in range(world):                                                                # push
            for side, link in links[r].items():
                peer = r - 1 if side == 0 else r + 1
                for f, n in enumerate(FIELD_ORDER):
                    so, sc, _, _ = link[n]
                    assert sc <= field_stride and so % 4 == 0 and sc % 4 == 0
                    mail[peer][1 - side, seq & 1, f, :sc] = flat[r][n][so:so + sc]
        for r in range(world):                                                                # unpack
            for side, link in links[r].items():
                for f, n in enumerate(FIELD_ORDER):
                    _, _, ro, rc = link[n]
                    flat[r][n][ro:ro + rc] = mail[r][side, seq & 1, f, :rc]
        for r, g in enumerate(geoms):
            for n in FIELD_ORDER:
                assert np.array_equal(flat[r][n].reshape(rows(g, n), pitch[n]), want[r][n].numpy()), (seq, r, n)
