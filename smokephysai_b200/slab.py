"""Row-slab decomposition of one large grid over several GPUs (BASELINE config 4:
"single 8192x8192 grid slab-decomposed across 2/4/8 B200 with per-sweep NCCL halo exchange over NVLink").

The reference has nothing like this (single process, single device); it is an extension behind the same
solver semantics, and its contract is: the slab run is BIT-IDENTICAL to the undecomposed run.

How.  Rank r owns the cell rows [R0, R1) of the H x W grid and stores [A, B) = [R0 - halo, R1 + halo)
clipped to the grid ("ghost rows").  Every phase of the step is a local stencil with fixed arithmetic, so
evaluating it redundantly on the ghost rows gives exactly the values the neighbour computes; each phase
only erodes the band of exact rows by its stencil radius.  A halo exchange (rows are contiguous in memory:
one send and one receive of `rows x pitch` floats per neighbour and field) refreshes the ghosts:

    exchange(u, v, density)                               once per step
    forces + diffusion + divergence                       erodes 1 row (2 for div)
    for each launch of t <= T fused Jacobi sweeps:        erodes t rows of p
        exchange(p)                                       -- "every sweep" when T = 1; deep halos when T > 1
    gradient subtract, advect u, v, density               erode 1 + (reach + 1) rows, reach = |dt * velocity|

so halo >= T + 4 rows keeps every owned row exact as long as the back-trace reaches at most halo - 4 rows;
the advect kernel checks that on the device (smk_slab_check_t) and `check()` raises if it was violated.
With a halo of at least K + 4 rows (K = jacobi_iters) the p ghosts survive all K sweeps of a step, and the plan
collapses to ONE exchange per step -- u, v, density and p in a single NCCL group -- at the price of K + 4 instead
of T + 4 redundant rows per side: at 8 slabs of 8192^2 that trades 2 x 14 -> 2 x 24 ghost rows (of 1024) for two
fewer latency-bound exchange phases per step (`halo=` picks; bench.py --halo).
Only the advection needs to know where the slab sits (absolute fp32 coordinates, global edge tests:
smk_grid_t.row0 / gh); all other kernels run on the slab as if it were a small grid, because a slab edge
that is not a grid edge only produces garbage in ghost rows that are already written off.

Exchangers: `NcclExchanger` (the default on GPUs: the library's own NCCL communicator, one C call per exchange
phase: ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd over NVLink), `DistExchanger` (torch.distributed P2P ops:
gloo in the CPU tests, NCCL if asked for) and `LocalGroup` (all slabs in one process on one GPU: the way to test the decomposition without a multi-GPU box).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import SlabCheck


class SlabGeometry:
    """Pure host arithmetic of the decomposition (testable without a GPU)."""

    def __init__(self, H, W, world, rank, halo):
        self.H, self.W, self.world, self.rank, self.halo = int(H), int(W), int(world), int(rank), int(halo)
        if not (0 <= self.rank < self.world):
            raise ValueError("rank %d outside world of %d" % (rank, world))
        self.splits = self.split(self.H, self.world)
        self.R0, self.R1 = self.splits[self.rank]
        if self.world > 1 and min(b - a for a, b in self.splits) < self.halo + 1:
            raise ValueError("slabs of %d rows are too thin for a halo of %d rows (need >= halo + 1)"
                             % (min(b - a for a, b in self.splits), self.halo))
        self.A, self.B = self.stored(self.rank)
        self.hl = self.B - self.A                    # local cell rows (u has hl + 1)
        self.top_edge, self.bottom_edge = self.A == 0, self.B == self.H
        self.own_lo, self.own_hi = self.R0 - self.A, self.R1 - self.A      # owned rows, local indices

    @staticmethod
    def split(H, world):
        base, rem = divmod(H, world)
        out, r0 = [], 0
        for r in range(world):
            n = base + (1 if r < rem else 0)
            out.append((r0, r0 + n))
            r0 += n
        return out

    def stored(self, rank):
        r0, r1 = self.splits[rank]
        return max(0, r0 - self.halo), min(self.H, r1 + self.halo)

    def rows(self, kind):
        """Local rows of a field: 'u' is staggered along y (hl + 1 rows), everything else has hl."""
        return self.hl + 1 if kind == "u" else self.hl

    def blocks(self, kind):
        """-> (sends, recvs), lists of (peer, first local row, number of rows) in a fixed order
        (upper neighbour first), identical on both sides of every pair."""
        sends, recvs = [], []
        extra = 1 if kind == "u" else 0
        r = self.rank
        if r > 0:                                            # upper neighbour r-1
            n = self.R0 - self.A
            recvs.append((r - 1, 0, n))
            b_up = self.stored(r - 1)[1]                     # it stores [.., b_up): its lower ghost is [R0, b_up) (+1 row of u)
            sends.append((r - 1, self.R0 - self.A, b_up - self.R0 + extra))
        if r < self.world - 1:                               # lower neighbour r+1
            recvs.append((r + 1, self.R1 - self.A, self.B - self.R1 + extra))
            a_dn = self.stored(r + 1)[0]                     # its upper ghost is [a_dn, R1)
            sends.append((r + 1, a_dn - self.A, self.R1 - a_dn))
        return sends, recvs

    def owned_rows(self, kind):
        """Local [lo, hi) of the rows this rank owns (the last rank also owns u's row H)."""
        hi = self.own_hi + (1 if kind == "u" and self.rank == self.world - 1 else 0)
        return self.own_lo, hi


def sweep_split(K, T):
    """K sweeps over ceil(K/T) launches, as evenly as possible (same rule as the library's tiled runs)."""
    if K <= 0:
        return []
    nl = (K + T - 1) // T
    out, left = [], K
    for l in range(nl):
        t = (left + (nl - l) - 1) // (nl - l)
        out.append(t)
        left -= t
    return out


def needs_single_exchange(world, halo, K):
    """Deep halo: the pressure ghosts stay exact through all K sweeps of a step, so one exchange per step suffices."""
    return world > 1 and halo >= int(K) + 4


def build_step_plan(world, K, T, single_exchange, fdd, jacobi, project_advect):
    """The step as a list of ("x", field names) halo exchanges and ("c", callable) compute phases.  Shared by
    SlabNavierStokes and the CPU mirror the tests drive, so the exchange schedule itself is what the CPU tests
    check against the undecomposed run."""
    plan = []
    if world > 1:
        plan.append(("x", ("u", "v", "d", "p") if single_exchange else ("u", "v", "d")))
    plan.append(("c", fdd))
    for t in sweep_split(K, T):
        plan.append(("c", (lambda t=t: jacobi(t))))
        if world > 1 and not single_exchange:
            plan.append(("x", ("p",)))
    plan.append(("c", project_advect))
    return plan


class DistExchanger:
    """Halo exchange over torch.distributed point-to-point ops (NCCL send/recv on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        self.group = group

    def exchange(self, geom, tensors):
        """tensors: list of (contiguous [rows, pitch] tensor, kind)."""
        import torch.distributed as dist
        ops = []
        for t, kind in tensors:
            sends, recvs = geom.blocks(kind)
            for peer, a, n in recvs:
                ops.append(dist.P2POp(dist.irecv, t[a:a + n], peer, self.group))
            for peer, a, n in sends:
                ops.append(dist.P2POp(dist.isend, t[a:a + n], peer, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


class NcclExchanger:
    """Halo exchange through the library's own NCCL communicator (include/smoke_b200.h: smk_nccl_*): all sends and
    receives of an exchange phase are one ncclGroup issued by ONE C call on the current stream.  The same traffic
    through torch.distributed's P2P ops costs ~25 us of host time per op (~250 us per step for the three phases),
    which bound the 8-GPU step (tools/slab_host_time.py).  The 128-byte NCCL id travels over the existing
    torch.distributed group; construction is collective (every rank of the group must build one)."""

    description = "NCCL send/recv over NVLink, one ncclGroup per exchange issued from C (smk_nccl_exchange)"

    def __init__(self, device, group=None):
        import torch.distributed as dist
        self.device = torch.device(device)
        ver = C.c_int32(0)
        _lib.call("smk_nccl_load", None, C.byref(ver))
        self.nccl_version = ver.value
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        on_gpu = dist.get_backend(group) == "nccl"
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            _lib.call("smk_nccl_unique_id", idbuf.data_ptr())
        t = idbuf.to(self.device) if on_gpu else idbuf
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idbuf = t.cpu().contiguous()
        self.comm = C.c_void_p()
        _lib.call("smk_set_device", self.device.index)
        _lib.call("smk_nccl_comm_init", idbuf.data_ptr(), self.rank, self.world, C.byref(self.comm))
        self._plans = {}

    def _plan(self, geom, tensors):
        key = tuple((t.data_ptr(), kind) for t, kind in tensors)
        plan = self._plans.get(key)
        if plan is None:
            blocks = []
            for t, kind in tensors:
                if not t.is_contiguous() or t.dtype != torch.float32:
                    raise ValueError("halo exchange needs contiguous fp32 [rows, pitch] tensors")
                pitch = t.shape[1]
                sends, recvs = geom.blocks(kind)
                for peer, a, n in recvs:
                    blocks.append(_lib.HaloBlock(t.data_ptr() + 4 * a * pitch, n * pitch, peer, 0))
                for peer, a, n in sends:
                    blocks.append(_lib.HaloBlock(t.data_ptr() + 4 * a * pitch, n * pitch, peer, 1))
            plan = ((_lib.HaloBlock * max(len(blocks), 1))(*blocks), len(blocks))
            self._plans[key] = plan
        return plan

    def exchange(self, geom, tensors):
        arr, n = self._plan(geom, tensors)
        _lib.call("smk_nccl_exchange", self.comm, arr, n, torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "comm", None) is not None and self.comm:
            torch.cuda.synchronize(self.device)
            _lib.call("smk_nccl_comm_destroy", self.comm)
            self.comm = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:                                # interpreter shutdown may already have torn torch / the library down
            self.close()
        except Exception:
            pass


FIELD_ORDER = ("u", "v", "d", "p")          # smk_peer_comm_t field order


def peer_links(geom, pitch):
    """Pure host arithmetic of the peer exchange: {side: {field: (send_off, send_count, recv_off, recv_count)}} in fp32
    elements relative to the base of each field, side 0 = upper neighbour (rank - 1), 1 = lower (rank + 1); pitch maps the
    field names u / v / d / p to their row pitch.  What one rank sends on a side is what its neighbour receives on the
    opposite side (tests/test_slab_host.py drives a numpy mailbox with these numbers against the plain halo copy)."""
    out = {}
    for side, peer in ((0, geom.rank - 1), (1, geom.rank + 1)):
        if not (0 <= peer < geom.world):
            continue
        link = {}
        for name in FIELD_ORDER:
            sends, recvs = geom.blocks("u" if name == "u" else "c")
            so = sc = ro = rc = 0
            for p_, a, n in sends:
                if p_ == peer:
                    so, sc = a * pitch[name], n * pitch[name]
            for p_, a, n in recvs:
                if p_ == peer:
                    ro, rc = a * pitch[name], n * pitch[name]
            link[name] = (so, sc, ro, rc)
        out[side] = link
    return out


class PeerExchanger:
    """Halo exchange by direct peer stores over NVLink (include/smoke_b200.h: smk_peer_*, csrc/peer_halo.cu).

    Every rank owns a mailbox tensor -- per neighbour two slots (exchange number & 1) of four field regions, plus the
    arrival counters -- exports it once with CUDA IPC (the 64-byte handle travels over the existing torch.distributed
    group) and maps its neighbours'.  An exchange is two small kernels on the caller's stream: push (boundary rows ->
    the neighbours' mailboxes, then st.release.sys on their counters) and unpack (ld.acquire.sys on the own counters,
    mailbox -> ghost rows); no NCCL proxy, no host round trip, capturable in a CUDA graph.  With a halo of at least
    K + 4 rows SlabNavierStokes.step() hands the whole step, exchange included, to ONE C call (smk_slab_step)."""
    description = ("direct peer stores into the neighbour's mailbox over NVLink (CUDA IPC mapping, st.release.sys / ld.acquire.sys "
                   "counters), two kernels per exchange (smk_peer_push / smk_peer_unpack)")
    NCOUNTERS = 64                          # uint32 slots behind the mailbox: [0..1] arrival counters, [2..3] seq

    def __init__(self, device, geom, layout, group=None, _bases=None):
        self.device = torch.device(device)
        self.geom, self.group = geom, group
        self.pitch = {"u": layout.pitch_u, "v": layout.pitch_v, "d": layout.pitch_c, "p": layout.pitch_c}
        # all ranks use the same region size so that a rank can address its neighbour's mailbox without asking
        self.field_stride = (geom.halo + 1) * max(self.pitch.values())
        self.parity_stride = 4 * self.field_stride
        self.region = 2 * self.parity_stride                # floats per neighbour: two slots
        self.buf = torch.zeros(2 * self.region + self.NCOUNTERS, dtype=torch.float32, device=self.device)
        self._opened = []
        self.comm = None
        if _bases is not None:                              # all slabs in one process (tests, LocalGroup): plain pointers
            self.wire(_bases)
        elif geom.world > 1:
            self._wire_ipc()

    # -- addresses inside a mailbox tensor (the same on every rank) ---------------------------------------------------
    def _region(self, base, side):
        return base + 4 * side * self.region

    def _counter(self, base, k):
        return base + 4 * (2 * self.region + k)

    def _wire_ipc(self):
        """Export this rank's mailbox, map the neighbours'.  Collective: every rank goes through the same all_gather / all_reduce
        whatever fails locally, and either every rank ends up wired or every rank raises SmokeLibraryError (so that a caller
        can fall back to NCCL on all ranks together instead of leaving some of them waiting in a collective)."""
        import torch.distributed as dist
        world = dist.get_world_size(self.group)
        on_gpu = dist.get_backend(self.group) == "nccl"
        handle = torch.zeros(72, dtype=torch.uint8)
        err = None
        try:
            _lib.call("smk_set_device", self.device.index)     # the library's own CUDA runtime must be on this rank's device
            off = C.c_int64(0)
            _lib.call("smk_ipc_export", self.buf.data_ptr(), handle.data_ptr(), C.byref(off))
            handle[64:72] = torch.tensor(list(int(off.value).to_bytes(8, "little", signed=True)), dtype=torch.uint8)
        except _lib.SmokeLibraryError as e:
            err = e
        mine = handle.to(self.device) if on_gpu else handle
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=self.group)
        bases = {}
        if err is None:
            try:
                for peer in (self.geom.rank - 1, self.geom.rank + 1):
                    if 0 <= peer < world:
                        h = every[peer].cpu().contiguous()
                        poff = int.from_bytes(bytes(h[64:72].tolist()), "little", signed=True)
                        ptr = C.c_void_p()
                        _lib.call("smk_ipc_open", h.data_ptr(), poff, C.byref(ptr))
                        self._opened.append((ptr.value, poff))
                        bases[peer] = ptr.value
            except _lib.SmokeLibraryError as e:
                err = e
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device if on_gpu else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)     # also orders "every mapping exists" before the first push
        if not int(ok.item()):
            self.close()
            raise _lib.SmokeLibraryError("peer halo exchange unavailable: %s" % (err if err is not None else "another rank could not map its neighbours"))
        self.wire(bases)

    def wire(self, bases):
        """bases: {neighbour rank: address of its mailbox tensor as seen from this process}."""
        g, mine = self.geom, self.buf.data_ptr()
        comm = _lib.PeerComm()
        comm.field_stride, comm.parity_stride = self.field_stride, self.parity_stride
        comm.seq = self._counter(mine, 2)
        for side, peer in ((0, g.rank - 1), (1, g.rank + 1)):
            link = comm.link[side]
            if peer not in bases:
                continue
            # this rank is the LOWER neighbour of rank - 1 (its side 1) and the UPPER neighbour of rank + 1 (its side 0)
            link.remote_mailbox = self._region(bases[peer], 1 - side)
            link.remote_flag = self._counter(bases[peer], 1 - side)
            link.local_mailbox = self._region(mine, side)
            link.local_flag = self._counter(mine, side)
            numbers = peer_links(g, self.pitch)[side]
            for f, name in enumerate(FIELD_ORDER):
                link.send_off[f], link.send_count[f], link.recv_off[f], link.recv_count[f] = numbers[name]
        self.comm = comm

    def _bases(self, named):
        arr = (C.c_void_p * 4)()
        for t, name in named:
            arr[FIELD_ORDER.index(name)] = t.data_ptr()
        return arr

    def push(self, named):
        _lib.call("smk_peer_push", C.byref(self.comm), self._bases(named), torch.cuda.current_stream(self.device).cuda_stream)

    def unpack(self, named):
        _lib.call("smk_peer_unpack", C.byref(self.comm), self._bases(named), torch.cuda.current_stream(self.device).cuda_stream)

    def exchange_named(self, named):
        """named: list of (contiguous [rows, pitch] tensor, field name in u / v / d / p)."""
        arr = self._bases(named)
        s = torch.cuda.current_stream(self.device).cuda_stream
        _lib.call("smk_peer_push", C.byref(self.comm), arr, s)
        _lib.call("smk_peer_unpack", C.byref(self.comm), arr, s)

    def close(self):
        if self._opened:
            torch.cuda.synchronize(self.device)
            for ptr, off in self._opened:
                _lib.call("smk_ipc_close", ptr, off)
            self._opened = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def peer_exchange_possible(device, group=None):
    """True when this rank can map both neighbours' memory: one process per GPU on one node, peer access between the
    GPUs of adjacent ranks (every rank must agree, so the answer is all-reduced)."""
    import torch.distributed as dist
    dev = torch.device(device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ok = 1 if torch.cuda.device_count() >= world else 0
    if ok:
        for peer in (rank - 1, rank + 1):
            if 0 <= peer < world and not torch.cuda.can_device_access_peer(dev.index, (dev.index - rank + peer) % torch.cuda.device_count()):
                ok = 0
    t = torch.tensor([ok], dtype=torch.int32, device=dev if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def default_exchanger(device, exchange="auto"):
    """NcclExchanger when the default process group runs NCCL on GPUs, else torch.distributed P2P (gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_backend() == "nccl" and torch.device(device).type == "cuda":
        if exchange == "peer" or (exchange == "auto" and peer_exchange_possible(device)):
            return "peer"                   # built by SlabNavierStokes once the geometry and layout exist
        return NcclExchanger(device)
    if exchange == "peer":
        raise RuntimeError("exchange='peer' needs one process per GPU under an NCCL process group")
    return DistExchanger()


def local_exchange(geoms, tensor_lists):
    """In-process exchange between all slabs: tensor_lists[r] = list of (tensor, kind) of rank r."""
    for r, geom in enumerate(geoms):
        for k, (t, kind) in enumerate(tensor_lists[r]):
            _, recvs = geom.blocks(kind)
            for peer, a, n in recvs:
                sends, _ = geoms[peer].blocks(kind)
                b, m = next((b, m) for p, b, m in sends if p == r)
                assert m == n, "halo plan mismatch between rank %d and %d" % (r, peer)
                t[a:a + n].copy_(tensor_lists[peer][k][0][b:b + m])


class SlabNavierStokes:
    """One rank's slab of a NavierStokesSimulator (navier_stokes.py:6-173 semantics on the global grid)."""

    def __init__(self, grid_size, dt=0.01, viscosity=0.001, device="cuda", *, rank, world, jacobi_iters=20,
                 sweeps_per_launch=10, halo=None, exchanger=None, exchange="auto"):
        from .navier_stokes import NavierStokesSimulator
        H, W = int(grid_size[0]), int(grid_size[1])
        self.T = max(1, int(sweeps_per_launch))
        self.halo = int(halo) if halo is not None else self.T + 4
        if world > 1 and self.halo < self.T + 4:
            # T rows eroded by the fused sweeps of a launch, 2 by diffusion + divergence, 1 by the gradient subtract and 1 for
            # the second row of a sub-cell bilinear back-trace; deeper back-traces need halo - 4 >= their reach (check())
            raise ValueError("halo of %d rows is too shallow for %d fused sweeps per launch (need >= T + 4)" % (self.halo, self.T))
        # deep halo: the pressure ghosts stay exact through all K sweeps, one exchange per step (u, v, density, p)
        self.single_exchange = needs_single_exchange(world, self.halo, jacobi_iters)
        self.geom = SlabGeometry(H, W, world, rank, self.halo if world > 1 else 0)
        self.grid_size, self.dt, self.viscosity = (H, W), dt, viscosity
        self.jacobi_iters = int(jacobi_iters)
        self.rank, self.world = int(rank), int(world)
        self.local = NavierStokesSimulator((self.geom.hl, W), dt, viscosity, device, jacobi_iters=jacobi_iters,
                                           sweeps_per_launch=self.T, _slab=(self.geom.A, H))
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl', got %r" % (exchange,))
        self.exchanger = exchanger if exchanger is not None else (default_exchanger(self.local._cuda, exchange) if world > 1 else None)
        if isinstance(self.exchanger, str) and self.exchanger == "peer":
            try:
                self.exchanger = PeerExchanger(self.local._cuda, self.geom, self.local._layout)
            except _lib.SmokeLibraryError:          # raised on every rank together (PeerExchanger._wire_ipc)
                if exchange == "peer":
                    raise
                self.exchanger = NcclExchanger(self.local._cuda)
        self._overflow = torch.zeros(1, dtype=torch.int32, device=self.local._cuda)
        self.steps_done = 0
        self._pushed_ahead = False          # the last step already pushed the ghost rows of the next one (step(push_ahead=True))

    # ------------------------------------------------------------------ fields
    def full(self, name):
        """Contiguous [local rows, pitch] tensor of the live copy of u / v / d / p (ghost rows and padding included)."""
        ns, L = self.local, self.local._layout
        k = {"density": "d"}.get(name, name)
        cur = getattr(ns._state, "cur_" + k)
        fname = "%s%d" % (k, cur)
        rows, _, pitch = L.shape_of(fname)
        off = L.offset[fname]
        return ns._arena[off: off + rows * pitch].view(rows, pitch)

    def _kind(self, name):
        return "u" if name == "u" else "c"

    def owned(self, name):
        """View of the rows this rank owns, [rows, cols]."""
        lo, hi = self.geom.owned_rows(self._kind(name))
        cols = {"u": self.geom.W, "v": self.geom.W + 1}.get(name, self.geom.W)
        return self.full(name)[lo:hi, :cols]

    def _no_push_in_flight(self, what):
        if self._pushed_ahead:
            raise RuntimeError("%s after step(push_ahead=True): the boundary rows of the next step are already on their way to the "
                               "neighbours; finish with a plain step() first (run_steps does)" % what)

    def scatter(self, name, global_field):
        """Fill the stored rows (owned + ghosts) of a field from the global [rows, cols] array."""
        self._no_push_in_flight("scatter")
        g = torch.as_tensor(global_field, dtype=torch.float32)
        rows = self.geom.rows(self._kind(name))
        dst = self.full(name)
        dst[:, :g.shape[1]].copy_(g[self.geom.A: self.geom.A + rows].to(dst.device))

    def add_smoke_source(self, x, y, radius=10, intensity=1.0):
        """navier_stokes.py:37-48 with the centre given in global coordinates."""
        self._no_push_in_flight("add_smoke_source")
        self.local.add_smoke_source(x, int(y) - self.geom.A, radius, intensity)

    def add_sources(self, sources):
        """Ordered emitter list [(x, y, radius, intensity), ...] in global coordinates, one batched splat."""
        self._no_push_in_flight("add_sources")
        self.local.add_sources([[(x, int(y) - self.geom.A, r, i) for x, y, r, i in sources]])

    def setup_grid(self):
        self._no_push_in_flight("setup_grid")
        self.local.setup_grid()
        self._overflow.zero_()

    # ------------------------------------------------------------------ the step, phase by phase
    def _g(self):
        return C.byref(self.local._grid)

    def _fdd(self):
        ns = self.local
        st, prm = ns._state, ns._params()
        cu, cv, cd = st.cur_u, st.cur_v, st.cur_d
        _lib.call("smk_forces_diffuse_div", self._g(), st.u[cu], st.v[cv], st.d[cd], st.u[cu ^ 1], st.v[cv ^ 1], st.d[cd ^ 1],
                  st.div, prm.dt, prm.c_uv, prm.c_d, ns._stream())
        st.cur_u, st.cur_v, st.cur_d = cu ^ 1, cv ^ 1, cd ^ 1

    def _jacobi(self, t):
        ns = self.local
        st = ns._state
        flag = C.c_int32(0)
        _lib.call("smk_jacobi", self._g(), st.div, st.p[st.cur_p], st.p[st.cur_p ^ 1], int(t), int(t), C.byref(flag), ns._stream())
        st.cur_p ^= flag.value

    def _check(self, rows):
        g = self.geom
        if self.world == 1:
            return None
        # rows of the advected field / velocities that are no longer exact next to a cut: 2 when p was refreshed
        # just before the gradient subtract, K + 2 when the step's only exchange was at its start
        m = self.jacobi_iters + 2 if self.single_exchange else 2
        lo = 0 if g.top_edge else m
        hi = rows if g.bottom_edge else rows - m
        return SlabCheck(g.own_lo, min(g.own_hi + 1, rows), lo, hi, self._overflow.data_ptr())

    def _project_advect(self):
        """Gradient subtract + the three advections + decay in one C call (smk_project_advect: navier_stokes.py:148-149, :166-171)."""
        ns, g = self.local, self.geom
        prm = ns._params()
        chk = [self._check(rows) for rows in (g.hl + 1, g.hl, g.hl)]
        ref = [C.byref(c) if c is not None else None for c in chk]
        _lib.call("smk_project_advect", self._g(), C.byref(ns._state), C.byref(prm), ref[0], ref[1], ref[2], ns._stream())
        self.steps_done += 1

    def step_plan(self):
        """The step as a list of ("x", field names) halo exchanges and ("c", callable) compute phases."""
        return build_step_plan(self.world, self.jacobi_iters, self.T, self.single_exchange, self._fdd, self._jacobi, self._project_advect)

    def exchange_list(self, names):
        return [(self.full(n), self._kind(n)) for n in names]

    def exchange(self, names):
        """Refresh the ghost rows of the named fields (u / v / d / p) from both neighbours."""
        if isinstance(self.exchanger, PeerExchanger):
            self.exchanger.exchange_named([(self.full(n), n) for n in names])
        else:
            self.exchanger.exchange(self.geom, self.exchange_list(names))

    PUSH_HEAD, PUSH_TAIL = 1, 2             # SMK_SLAB_PUSH_* of include/smoke_b200.h

    def _c_step(self, flags=1):
        """The whole step in one C call (smk_slab_step): peer exchange, forces / diffusion / divergence, Jacobi launches,
        gradient subtract, the three advections.  Same kernels, same order as the plan below."""
        ns, g = self.local, self.geom
        prm = ns._params()
        chk = [self._check(rows) for rows in (g.hl + 1, g.hl, g.hl)]
        ref = [C.byref(c) if c is not None else None for c in chk]
        comm = C.byref(self.exchanger.comm) if isinstance(self.exchanger, PeerExchanger) else None
        _lib.call("smk_slab_step", self._g(), C.byref(ns._state), C.byref(prm), comm, int(flags), ref[0], ref[1], ref[2], ns._stream())
        self.steps_done += 1

    def step(self, push_ahead=False):
        """One time step of this rank's slab (all ranks must call it together, with the same arguments).

        push_ahead=True (peer-store exchange with one exchange per step; ignored otherwise) promises that the next thing to happen
        to this slab is another step(): the boundary rows that step needs are then sent from the middle of this step's density
        advection (the rows to send are advected first), and the transfer overlaps the rest of the launch instead of heading the
        next step.  run_steps(n) does this for all but the last step."""
        if self.world == 1 or (self.single_exchange and self.exchanger and not getattr(self, "_phase_by_phase", False)):
            # one exchange per step: the compute part is ONE C call (smk_slab_step), which also issues the exchange when it
            # is the peer-store one; an NCCL / torch.distributed exchange is issued from here first
            peer = self.world > 1 and isinstance(self.exchanger, PeerExchanger)
            if self.world > 1 and not peer:
                self.exchange(("u", "v", "d", "p"))
            flags = 0
            if peer:
                flags = (0 if self._pushed_ahead else self.PUSH_HEAD) | (self.PUSH_TAIL if push_ahead else 0)
                self._pushed_ahead = bool(push_ahead)
            return self._c_step(flags)
        for kind, arg in self.step_plan():
            if kind == "x":
                self.exchange(arg)
            else:
                arg()

    def run_steps(self, nsteps):
        """nsteps consecutive steps; every step but the last sends the next step's ghost rows ahead (step(push_ahead=True))."""
        for k in range(int(nsteps)):
            self.step(push_ahead=k + 1 < int(nsteps))

    def exchange_description(self):
        """One line for logs / bench output: how the ghost rows travel."""
        if self.world == 1 or not self.exchanger:
            return "no exchange"
        plan = ("ONE exchange per step (u, v, density, p)" if self.single_exchange
                else "u, v, density once per step and p after every launch of <= %d fused sweeps" % self.T)
        return "%s: %s" % (getattr(self.exchanger, "description", type(self.exchanger).__name__), plan)

    def capture(self, nsteps=1):
        """Record `nsteps` consecutive steps -- kernels and NCCL halo exchanges -- into a CUDA graph and return it
        (`graph.replay()` then advances the slab by nsteps; all ranks must replay together).  At eight slabs the
        step is a few hundred microseconds of kernels, less than what Python and the NCCL host path spend
        enqueueing its ~10 launches and 3 grouped send/recvs; a replayed graph removes that host time.  One eager
        step must have run before (NCCL builds its P2P connections on first use, which cannot be captured), and
        the ping-pong state has to come back to the same buffers after nsteps (it does after every step when the
        number of Jacobi launches per step is even, else after every second step)."""
        ns = self.local
        st = ns._state
        before = (st.cur_u, st.cur_v, st.cur_d, st.cur_p)
        torch.cuda.synchronize(ns._cuda)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.run_steps(nsteps)
        after = (st.cur_u, st.cur_v, st.cur_d, st.cur_p)
        if after != before:
            raise RuntimeError("the ping-pong state does not return to the same buffers after %d step(s): capture an even number" % nsteps)
        return graph

    def check(self):
        """Raise if an advection back-trace ever left the rows this slab holds exactly (synchronises)."""
        if int(self._overflow.item()) != 0:
            reach = self.halo - (self.jacobi_iters + 4 if self.single_exchange else 4)
            raise RuntimeError("slab halo of %d rows is too shallow for the velocities reached (|dt*v| > %d rows): "
                               "results near the slab boundary are not exact; use a deeper halo" % (self.halo, reach))

    def gather(self, name):
        """All ranks: the global field assembled from every rank's owned rows (torch.distributed all_gather)."""
        import torch.distributed as dist
        self.check()                       # rows next to a cut are only exact while no back-trace left the halo
        mine = self.owned(name).contiguous()
        if self.world == 1:
            return mine.clone()
        kind = self._kind(name)
        sizes = []
        for r, (r0, r1) in enumerate(self.geom.splits):
            sizes.append(r1 - r0 + (1 if kind == "u" and r == self.world - 1 else 0))
        mx = max(sizes)
        pad = torch.zeros(mx, mine.shape[1], dtype=mine.dtype, device=mine.device)
        pad[:mine.shape[0]] = mine
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad)
        return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)


class LocalGroup:
    """All `world` slabs of one grid in this process on one GPU, stepped in lockstep with in-process halo
    copies: the single-GPU emulation of the multi-GPU run (tests; bit-identical to the undecomposed solver)."""

    def __init__(self, grid_size, dt=0.01, viscosity=0.001, device="cuda", *, world, jacobi_iters=20, sweeps_per_launch=10, halo=None,
                 peer=False):
        self.slabs = [SlabNavierStokes(grid_size, dt, viscosity, device, rank=r, world=world, jacobi_iters=jacobi_iters,
                                       sweeps_per_launch=sweeps_per_launch, halo=halo, exchanger=False) for r in range(world)]
        self.world = world
        self.peer = bool(peer)
        if self.peer:
            # the kernels of the peer exchange with "remote" pointers that are plain pointers into the other slabs' mailboxes:
            # pushes of all slabs first, then the unpacks (one stream: an unpack waits for pushes that must already be queued)
            for s in self.slabs:
                s.exchanger = PeerExchanger(s.local._cuda, s.geom, s.local._layout, _bases={})
                s._phase_by_phase = True
            bases = {r: s.exchanger.buf.data_ptr() for r, s in enumerate(self.slabs)}
            for s in self.slabs:
                s.exchanger.wire({p: bases[p] for p in (s.rank - 1, s.rank + 1) if 0 <= p < world})

    def scatter(self, name, global_field):
        for s in self.slabs:
            s.scatter(name, global_field)

    def add_smoke_source(self, x, y, radius=10, intensity=1.0):
        for s in self.slabs:
            s.add_smoke_source(x, y, radius, intensity)

    def step(self):
        plans = [s.step_plan() for s in self.slabs]
        for k in range(len(plans[0])):
            kind = plans[0][k][0]
            if kind == "x":
                names = plans[0][k][1]
                if self.peer:
                    for s in self.slabs:
                        s.exchanger.push([(s.full(n), n) for n in names])
                    for s in self.slabs:
                        s.exchanger.unpack([(s.full(n), n) for n in names])
                else:
                    local_exchange([s.geom for s in self.slabs], [s.exchange_list(names) for s in self.slabs])
            else:
                for p in plans:
                    p[k][1]()

    def run_steps(self, nsteps):
        """nsteps steps through smk_slab_step with the ghost rows of every step but the first pushed ahead from the previous step's
        density advection (SlabNavierStokes.run_steps on one GPU).  Everything runs on one stream, so a kernel that waits for a push
        must be queued after it: the first step's pushes of ALL slabs are issued up front, and every step's tail pushes are queued
        before the next step's unpacks because the slabs are walked step by step."""
        if not (self.peer and all(s.single_exchange for s in self.slabs)):
            for _ in range(int(nsteps)):
                self.step()
            return
        names = ("u", "v", "d", "p")
        for s in self.slabs:
            s.exchanger.push([(s.full(n), n) for n in names])
        for k in range(int(nsteps)):
            for s in self.slabs:
                s._c_step(SlabNavierStokes.PUSH_TAIL if k + 1 < int(nsteps) else 0)

    def gather(self, name):
        self.check()
        return torch.cat([s.owned(name) for s in self.slabs], dim=0)

    def check(self):
        for s in self.slabs:
            s.check()
