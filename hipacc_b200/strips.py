"""Row-strip sharding of one image across the GPUs of a box.

One process per GPU (torch.distributed).  Rank i owns rows [H*i/N, H*(i+1)/N) of the global image
and stores them with R ghost rows above / below (R = vertical radius of the local operator, or the
summed radii of a fused pipeline).  Ghost rows hold REAL neighbour data, so the kernels apply the
vertical boundary mode only at the global top / bottom edge (hb_view.ghost_top / ghost_bottom) and
the sharded result is identical to the single-GPU result.

Exchange = one nearest-neighbour send/recv pair per direction (R * stride * sizeof(T) bytes, e.g.
64 KiB for a 32768-wide uchar image at R = 2) -- latency-bound, batched into a single NCCL group.
Global reductions combine per-GPU partials with one all-reduce of three scalars.

The reference has no multi-device code at all; this module is the
B200-native addition.  Works on any torch.distributed backend (NCCL on GPUs, gloo in CPU tests).
"""
from dataclasses import dataclass

from . import _abi as A


@dataclass
class StripPlan:
    width: int
    height: int        # global rows
    world: int
    rank: int
    radius: int        # ghost rows requested per side
    boundary: int = A.CLAMP

    @property
    def y0(self):
        return self.height * self.rank // self.world

    @property
    def y1(self):
        return self.height * (self.rank + 1) // self.world

    @property
    def rows(self):
        return self.y1 - self.y0

    def _has_up(self):
        return self.rank > 0 or (self.boundary == A.REPEAT and self.world > 1)

    def _has_down(self):
        return self.rank < self.world - 1 or (self.boundary == A.REPEAT and self.world > 1)

    @property
    def ghost_top(self):
        return self.radius if self._has_up() else 0

    @property
    def ghost_bottom(self):
        return self.radius if self._has_down() else 0

    @property
    def buffer_rows(self):
        return self.ghost_top + self.rows + self.ghost_bottom

    def roi(self):
        """(w, h, ox, oy) of the owned rows inside the strip buffer"""
        return (self.width, self.rows, 0, self.ghost_top)

    def ghost(self):
        return (self.ghost_top, self.ghost_bottom)

    def up(self):
        return (self.rank - 1) % self.world

    def down(self):
        return (self.rank + 1) % self.world

    def validate(self):
        assert self.world >= 1 and 0 <= self.rank < self.world
        assert self.height // self.world >= self.radius, "strips thinner than the halo: shard fewer ways"


def owned(buf, plan):
    """The rows this rank owns (a view into the strip buffer)."""
    return buf[plan.ghost_top:plan.ghost_top + plan.rows]


def exchange_halos(buf, plan, group=None):
    """Fill the ghost rows of `buf` (shape [plan.buffer_rows, stride]) from the neighbouring ranks.

    `buf` must be the full-stride buffer so that a block of rows is one contiguous message.  All
    sends / receives of one exchange are issued as a single batch (one NCCL group launch)."""
    import torch.distributed as dist
    if plan.world == 1 or plan.radius == 0:
        return
    R, gt, rows = plan.radius, plan.ghost_top, plan.rows
    ops = []
    if plan._has_up():     # my first R owned rows -> upper neighbour's bottom ghost; its last R rows -> my top ghost
        ops.append(dist.P2POp(dist.isend, buf[gt:gt + R], plan.up(), group))
        ops.append(dist.P2POp(dist.irecv, buf[0:gt], plan.up(), group))
    if plan._has_down():
        ops.append(dist.P2POp(dist.isend, buf[gt + rows - R:gt + rows], plan.down(), group))
        ops.append(dist.P2POp(dist.irecv, buf[gt + rows:gt + rows + plan.ghost_bottom], plan.down(), group))
    if plan.world == 2 and plan.boundary == A.REPEAT:
        # both neighbours are the same peer and messages of one pair match in order: on both ranks
        # send top rows, send bottom rows, then receive the peer's top rows (my bottom ghost) and its
        # bottom rows (my top ghost)
        ops = [ops[0], ops[2], ops[3], ops[1]]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class P2PHalo:
    """Peer-to-peer halo exchange of one strip buffer over NVLink (hb_halo_exchange): the ranks swap CUDA IPC handles
    of their buffers and control blocks once (host side, through torch.distributed), afterwards every exchange is ONE
    kernel launch per rank that pushes the edge rows into the neighbours' ghost rows and waits on device-side flags.
    `buf` must come from hipacc_b200.alloc_image (a whole CUDA allocation) with shape [plan.buffer_rows, stride]."""

    def __init__(self, hb, buf, plan, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.hb, self.buf, self.plan = hb, buf, plan
        self.L = hb.lib()
        self.desc = None
        if plan.world == 1 or plan.radius == 0:
            return
        L = self.L
        mem, cmem = A.hb_ipc_mem(), A.hb_ipc_mem()
        self.ctrl = C.c_void_p()
        es = buf.element_size()
        # every rank reaches every collective below even if a local CUDA call fails: errors travel with the
        # gathered objects and all ranks raise together (no rank is left waiting in a collective)
        err = None
        try:
            hb._check(L.hb_halo_ctrl_create(C.byref(self.ctrl)), "hb_halo_ctrl_create")
            hb._check(L.hb_ipc_export(C.c_void_p(buf.data_ptr()), C.byref(mem)), "hb_ipc_export(buffer)")
            hb._check(L.hb_ipc_export(self.ctrl, C.byref(cmem)), "hb_ipc_export(control block)")
        except Exception as e:  # noqa: BLE001
            err = f"rank {plan.rank}: {e}"
        mine = {"mem": bytes(mem.handle), "ctrl": bytes(cmem.handle), "gt": plan.ghost_top, "rows": plan.rows,
                "pitch": buf.stride(0) * es, "err": err}
        everyone = [None] * plan.world
        dist.all_gather_object(everyone, mine, group=group)
        errs = [o["err"] for o in everyone if o["err"]]
        if errs:
            raise RuntimeError("P2PHalo: CUDA IPC export failed: " + "; ".join(errs))
        self._peers = {}

        def open_peer(r):
            if r not in self._peers:
                pm, pc = A.hb_ipc_mem(), A.hb_ipc_mem()
                C.memmove(pm.handle, everyone[r]["mem"], 64)
                C.memmove(pc.handle, everyone[r]["ctrl"], 64)
                pb, pk = C.c_void_p(), C.c_void_p()
                hb._check(L.hb_ipc_open(C.byref(pm), C.byref(pb)), "hb_ipc_open(buffer)")
                hb._check(L.hb_ipc_open(C.byref(pc), C.byref(pk)), "hb_ipc_open(control block)")
                self._peers[r] = (pb, pk)
            return self._peers[r]

        d = A.hb_halo_desc()
        d.buf, d.pitch_bytes, d.row_bytes = buf.data_ptr(), buf.stride(0) * es, plan.width * es
        d.ghost_top, d.rows, d.radius = plan.ghost_top, plan.rows, plan.radius
        d.ctrl = self.ctrl
        try:
            if plan._has_up():
                pb, pk = open_peer(plan.up())
                info = everyone[plan.up()]
                d.up_buf, d.up_ctrl, d.up_pitch_bytes, d.up_ghost_top, d.up_rows = pb, pk, info["pitch"], info["gt"], info["rows"]
            if plan._has_down():
                pb, pk = open_peer(plan.down())
                info = everyone[plan.down()]
                d.down_buf, d.down_ctrl, d.down_pitch_bytes, d.down_ghost_top = pb, pk, info["pitch"], info["gt"]
        except Exception as e:  # noqa: BLE001
            err = f"rank {plan.rank}: {e}"
        status = [None] * plan.world
        dist.all_gather_object(status, err, group=group)   # also the barrier: every rank has mapped its neighbours before the first push
        errs = [e for e in status if e]
        if errs:
            raise RuntimeError("P2PHalo: mapping a neighbour's memory failed: " + "; ".join(errs))
        self.desc = d

    def exchange(self, stream=None):
        import ctypes as C
        if self.desc is None:
            return
        self.hb._check(self.L.hb_halo_exchange(C.byref(self.desc), self.hb.stream_ptr(stream)), "hb_halo_exchange")

    def close(self):
        """Unmap the neighbours' memory and free the control block (after the last exchange has completed on every
        rank: the neighbours write into this control block)."""
        if self.desc is None:
            return
        for pb, pk in self._peers.values():
            self.L.hb_ipc_close(pb)
            self.L.hb_ipc_close(pk)
        self._peers = {}
        self.L.hb_halo_ctrl_destroy(self.ctrl)
        self.desc = None

    def check(self):
        """raise if an exchange on this control block ever timed out (the block is poisoned afterwards)"""
        n, timed_out = self.status()
        if timed_out or n < 0:
            raise RuntimeError(f"rank {self.plan.rank}: a halo exchange timed out waiting for a neighbour; ghost rows are stale")

    @staticmethod
    def exchange_batch(halos, stream=None):
        """several strip buffers in ONE launch (hb_halo_exchange_batch, one CTA per buffer)"""
        import ctypes as C
        live = [h for h in halos if h.desc is not None]
        if not live:
            return
        arr = (C.POINTER(A.hb_halo_desc) * len(live))(*[C.pointer(h.desc) for h in live])
        live[0].hb._check(live[0].L.hb_halo_exchange_batch(arr, len(live), live[0].hb.stream_ptr(stream)), "hb_halo_exchange_batch")

    def status(self):
        """(exchanges completed, timed_out flag) read back from the control block (synchronising)"""
        import ctypes as C
        if self.desc is None:
            return 0, 0
        n, t = C.c_int(), C.c_int()
        self.hb._check(self.L.hb_halo_status(self.ctrl, C.byref(n), C.byref(t)), "hb_halo_status")
        return n.value, t.value


def allgather_minmaxsum(partials, group=None):
    """Combine the per-rank device partials of hb_reduce_minmaxsum_f32_async ({float min, float max, double sum} =
    16 bytes, passed as a float32[4] CUDA tensor) with ONE collective: all-gather the 16-byte records, fold locally in
    rank order (deterministic).  Returns a float64[3] CUDA tensor (min, max, sum); no host synchronisation."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        allp = partials.view(1, 4)
    else:
        allp = torch.empty((world, 4), dtype=torch.float32, device=partials.device)
        dist.all_gather_into_tensor(allp, partials.view(1, 4), group=group)
    sums = allp[:, 2:4].contiguous().view(torch.float64)
    return torch.stack([allp[:, 0].min().double(), allp[:, 1].max().double(), sums.sum()])


def allreduce_minmaxsum(mn, mx, sm, group=None, device=None):
    """Combine per-rank (min, max, sum) partials: one all-reduce per op over a scalar each
    (MIN / MAX exact, SUM in float64)."""
    import torch
    import torch.distributed as dist
    t_mn = torch.tensor([mn], dtype=torch.float32, device=device)
    t_mx = torch.tensor([mx], dtype=torch.float32, device=device)
    t_sm = torch.tensor([sm], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t_mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(t_mx, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(t_sm, op=dist.ReduceOp.SUM, group=group)
    return float(t_mn.item()), float(t_mx.item()), float(t_sm.item())


# ----------------------------------------------------------------------------- sharded pyramids
class StripPyramid:
    """This rank's row strips of every level of an image pyramid (level l = (width >> l) x (height >> l)).

    Every level is cut at the same relative rows (rank * H_l / world), which requires the strip boundaries of level
    0 to be multiples of 2^(depth-1): then a strip's local row parity equals the global one and the NN / LF
    mappings of the level transitions are unchanged.  Each level buffer carries `radius` ghost rows per interior
    side: the fused down kernel needs mask/2 + 2 rows of the finer level, the up kernel one row of the coarser."""

    def __init__(self, width, height, depth, world, rank, radius, device, stride_align=64, hb=None):
        import torch
        assert height % (world << (depth - 1)) == 0 and width % (1 << (depth - 1)) == 0, \
            "sharded pyramid: level-0 strips must be multiples of 2^(depth-1) rows and the width of 2^(depth-1)"
        self.depth, self.world, self.rank = depth, world, rank
        self.plans = [StripPlan(width >> l, height >> l, world, rank, radius, A.CLAMP) for l in range(depth)]
        for p in self.plans:
            p.validate()
        self.bufs, self.halos = [], None
        for p in self.plans:
            stride = (p.width + stride_align - 1) // stride_align * stride_align
            if hb is None:
                self.bufs.append(torch.zeros((p.buffer_rows, stride), dtype=torch.float32, device=device))
            else:   # library allocations: CUDA IPC (peer-to-peer halo exchange) needs whole allocations
                self.bufs.append(hb.alloc_image(A.F32, stride, p.buffer_rows, device=device, align_bytes=4 * stride_align).zero_())

    def enable_p2p(self, hb, group=None):
        """map the neighbours' level buffers (CUDA IPC): exchanges become single kernel launches (P2PHalo)"""
        self.halos = [P2PHalo(hb, b, p, group) for b, p in zip(self.bufs, self.plans)]

    def exchange(self, l, stream=None, group=None):
        if self.halos is not None:
            self.halos[l].exchange(stream)
        else:
            exchange_many([(self.bufs[l], self.plans[l])], group)

    def owned(self, l):
        return owned(self.bufs[l], self.plans[l])[:, :self.plans[l].width]

    def strip(self, l):
        """(tensor, roi, ghost) of level l: the owned rows with their ghost rows declared"""
        p = self.plans[l]
        return (self.bufs[l][:, :p.width], p.roi(), p.ghost())

    def region(self, l):
        """(tensor, roi, no ghosts): the owned rows only (outputs)"""
        p = self.plans[l]
        return (self.bufs[l][:, :p.width], p.roi(), (0, 0))


def exchange_many(pairs, group=None):
    """One batched halo exchange for several (buffer, plan) pairs (a single NCCL group launch)."""
    import torch.distributed as dist
    ops = []
    for buf, plan in pairs:
        if plan.world == 1 or plan.radius == 0:
            continue
        R, gt, rows = plan.radius, plan.ghost_top, plan.rows
        if plan._has_up():
            ops.append(dist.P2POp(dist.isend, buf[gt:gt + R], plan.up(), group))
            ops.append(dist.P2POp(dist.irecv, buf[0:gt], plan.up(), group))
        if plan._has_down():
            ops.append(dist.P2POp(dist.isend, buf[gt + rows - R:gt + rows], plan.down(), group))
            ops.append(dist.P2POp(dist.irecv, buf[gt + rows:gt + rows + plan.ghost_bottom], plan.down(), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def pyramid_down_step(hb, pg, pl, l, mask, stream=None):
    """level l-1 -> l on this rank's strips (ghost rows of gaus(l-1) must be current)"""
    hb.pyr_down(pg.strip(l - 1), pg.region(l), mask, lap_fine=pl.region(l - 1), stream=stream)


def pyramid_up_step(hb, pg, pl, l, stream=None):
    """level l+1 -> l on this rank's strips (ghost rows of gaus(l+1) and lap(l+1) must be current)"""
    hb.pyr_up(pg.strip(l + 1), pl.strip(l + 1), pg.region(l), pl.region(l), stream=stream)


def pyramid_traverse_strips(hb, pg, pl, mask, stream=None, group=None):
    """The Gaussian / Laplacian pyramid traversal (Gaussian_Laplacian_Pyramid/src/main.cpp:199-248) on row strips:
    one halo exchange (radius rows to each neighbour) before every level transition, kernels unchanged otherwise;
    results are bit-identical to the unsharded traversal.  Exchanges go peer to peer (one kernel launch each) when
    the pyramids were set up with enable_p2p(), through NCCL send/recv otherwise."""
    for l in range(1, pg.depth):
        pg.exchange(l - 1, stream, group)
        pyramid_down_step(hb, pg, pl, l, mask, stream)
    for l in range(pg.depth - 2, -1, -1):
        if pg.halos is not None and pl.halos is not None:
            P2PHalo.exchange_batch([pg.halos[l + 1], pl.halos[l + 1]], stream)   # both planes in one launch
        else:
            pg.exchange(l + 1, stream, group)
            pl.exchange(l + 1, stream, group)
        pyramid_up_step(hb, pg, pl, l, stream)


# ----------------------------------------------------------------------------- sharded pyramids, second design
class P2PGather:
    """All-gather of row strips over NVLink peer memory (hb_allgather_rows): every rank keeps the FULL image `buf`
    (shape [height, stride], from hipacc_b200.alloc_image), owns rows [row0, row0 + rows) of it and pushes them into
    every peer's copy with ONE kernel launch (one CTA per peer, device-side flags, CUDA-graph replayable)."""

    def __init__(self, hb, buf, row0, rows, row_elems, world, rank, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.hb, self.L, self.desc, self.rank = hb, hb.lib(), None, rank
        if world == 1:
            return
        assert world - 1 <= A.HB_MAX_PEERS
        L = self.L
        mem, cmem = A.hb_ipc_mem(), A.hb_ipc_mem()
        self.ctrl = C.c_void_p()
        es = buf.element_size()
        err = None
        try:
            hb._check(L.hb_halo_ctrl_create(C.byref(self.ctrl)), "hb_halo_ctrl_create")
            hb._check(L.hb_ipc_export(C.c_void_p(buf.data_ptr()), C.byref(mem)), "hb_ipc_export(buffer)")
            hb._check(L.hb_ipc_export(self.ctrl, C.byref(cmem)), "hb_ipc_export(control block)")
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        mine = {"mem": bytes(mem.handle), "ctrl": bytes(cmem.handle), "pitch": buf.stride(0) * es, "rows": buf.shape[0], "err": err}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        errs = [o["err"] for o in everyone if o["err"]]
        if errs:
            raise RuntimeError("P2PGather: CUDA IPC export failed: " + "; ".join(errs))
        assert all(o["pitch"] == mine["pitch"] and o["rows"] == mine["rows"] for o in everyone), "P2PGather: the copies must share one layout"
        d = A.hb_gather_desc()
        d.buf, d.pitch_bytes, d.row_bytes = buf.data_ptr(), buf.stride(0) * es, row_elems * es
        d.row0, d.rows, d.ctrl, d.my_slot, d.n_peers = row0, rows, self.ctrl, rank, world - 1
        self._peers = []
        try:
            # peer order starts at my lower neighbour: the ranks do not all push to rank 0 first
            for j, r in enumerate([(rank + 1 + k) % world for k in range(world - 1)]):
                pm, pc = A.hb_ipc_mem(), A.hb_ipc_mem()
                C.memmove(pm.handle, everyone[r]["mem"], 64)
                C.memmove(pc.handle, everyone[r]["ctrl"], 64)
                pb, pk = C.c_void_p(), C.c_void_p()
                hb._check(L.hb_ipc_open(C.byref(pm), C.byref(pb)), "hb_ipc_open(buffer)")
                hb._check(L.hb_ipc_open(C.byref(pc), C.byref(pk)), "hb_ipc_open(control block)")
                self._peers.append((pb, pk))
                d.peer_buf[j], d.peer_ctrl[j], d.peer_slot[j] = pb.value, pk.value, r
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        status = [None] * world
        dist.all_gather_object(status, err, group=group)
        errs = [e for e in status if e]
        if errs:
            raise RuntimeError("P2PGather: mapping a peer's memory failed: " + "; ".join(errs))
        self.desc = d

    def gather(self, stream=None):
        import ctypes as C
        if self.desc is not None:
            self.hb._check(self.L.hb_allgather_rows(C.byref(self.desc), self.hb.stream_ptr(stream)), "hb_allgather_rows")

    def check(self):
        import ctypes as C
        if self.desc is None:
            return
        n, t = C.c_int(), C.c_int()
        self.hb._check(self.L.hb_halo_status(self.ctrl, C.byref(n), C.byref(t)), "hb_halo_status")
        if t.value or n.value < 0:
            raise RuntimeError(f"rank {self.rank}: an all-gather timed out waiting for a peer")


class PyramidShardPlan:
    """Host logic of the sharded pyramid traversal with ONE halo exchange and ONE all-gather (SURVEY.md 8e "Pyramid").

    Levels 0 .. G-1 are cut into row strips (every level at the same relative rows); levels >= G are small and every
    rank keeps them in full.  Instead of exchanging halos before every level transition, a rank RECOMPUTES the few rows
    of its neighbours it will need later ("extension" rows): the level-0 strip is loaded with E0 ghost rows per interior
    side (one peer-to-peer exchange), every down step l-1 -> l then produces its strip of level l plus e[l] extension
    rows per side from data it already holds, level G is all-gathered (each rank publishes its own rows), the coarse
    levels are traversed redundantly on every rank, and the way up needs no communication at all: the up step towards
    level l writes f[l] extension rows so that the next one finds its coarse ghost row locally.  Every pixel is computed
    by the same kernels in the same order as in the unsharded traversal, so the results are bit-identical.

    Ghost-row arithmetic (K = mask/2 + 2 rows is what the fused down kernel reads beyond its fine region):
        f[0] = 0, f[l] = 2 (1 <= l < G)            fine-region extension of the up step that writes level l
        e[G] = f[G-1] / 2 + d                        lap(G-1) must cover the up step's fine region
        e[l] = max(2 e[l+1] + K, f[l-1] / 2 + d)     level l must hold what the next down step reads
                                                     (d = 1 with split_dog: the separate DoG kernel reads one more coarse row)
        V[l] = 2 e[l+1] + K                          valid rows needed beyond the strip at level l;  E0 = V[0]
    """

    def __init__(self, width, height, depth, world, rank, mask_size, gather_level=None, max_gather_pixels=1 << 20, split_dog=False):
        assert depth >= 1 and world >= 1 and 0 <= rank < world
        # split_dog: the DifferenceOfGaussian of every sharded level runs as its own kernel (hb_pyr_dog) on a second
        # stream, in the shadow of the latency-bound chain blur+subsample -> ... -> all-gather -> coarse levels.  The
        # separate DoG reads one coarse row beyond its region, so every extension is one row larger (d = 1).
        self.split_dog = bool(split_dog)
        d_ = 1 if split_dog else 0
        assert height % (world << (depth - 1)) == 0 and width % (1 << (depth - 1)) == 0, \
            "sharded pyramid: level-0 strips must be multiples of 2^(depth-1) rows and the width of 2^(depth-1)"
        self.width, self.height, self.depth, self.world, self.rank, self.mask_size = width, height, depth, world, rank, mask_size
        self.K = mask_size // 2 + 2
        if gather_level is None:   # the finest level that is small enough to replicate
            gather_level = next((l for l in range(1, depth) if (width >> l) * (height >> l) <= max_gather_pixels), depth - 1)
        self.G = G = max(1, min(gather_level, depth - 1)) if depth > 1 else 1
        self.top, self.bot = (1 if rank > 0 else 0), (1 if rank < world - 1 else 0)
        f = [0] + [2] * max(G - 1, 0)
        e = [0] * (G + 2)
        if depth > 1:
            e[G] = f[G - 1] // 2 + d_
            for l in range(G - 1, 0, -1):
                e[l] = max(2 * e[l + 1] + self.K, f[l - 1] // 2 + d_)
        self.f, self.e = f, e
        self.V = [2 * e[l + 1] + self.K for l in range(G)] if depth > 1 else [0]
        self.E0 = self.V[0] if world > 1 else 0
        for l in range(min(G, depth)):
            assert world == 1 or self.V[l] <= self.rows(l), \
                f"level {l}: strips of {self.rows(l)} rows are thinner than the {self.V[l]} extension rows -- shard fewer ways or gather at a finer level"

    # ---- geometry (global rows of level l)
    def y0(self, l):
        return (self.height >> l) * self.rank // self.world

    def y1(self, l):
        return (self.height >> l) * (self.rank + 1) // self.world

    def rows(self, l):
        return self.y1(l) - self.y0(l)

    def sharded(self, l):
        return l < self.G and self.world > 1

    def buffer_span(self, l):
        """global rows [a, b) level l's buffer holds on this rank"""
        if not self.sharded(l):
            return 0, self.height >> l
        return self.y0(l) - self.V[l] * self.top, self.y1(l) + self.V[l] * self.bot

    def span(self, l, ext):
        """global rows of this rank's strip of level l extended by `ext` rows per interior side"""
        if self.world == 1:
            return 0, self.height >> l
        return self.y0(l) - ext * self.top, self.y1(l) + ext * self.bot

    def view_args(self, l, rows, ghost_wanted):
        """(roi, ghost) in buffer coordinates for global rows `rows` = (r0, r1) of level l; ghosts are what the buffer
        really holds beyond the region, capped at `ghost_wanted` (0 at the global image edge)"""
        a, b = self.buffer_span(l)
        r0, r1 = rows
        assert a <= r0 < r1 <= b, (l, rows, (a, b))
        gt, gb = min(r0 - a, ghost_wanted), min(b - r1, ghost_wanted)
        return (self.width >> l, r1 - r0, 0, r0 - a), (gt, gb)


class ShardedPyramid:
    """This rank's buffers of a Gaussian and a Laplacian pyramid under a PyramidShardPlan, plus the traversal.
    `transport` supplies the two communication steps: exchange0(stream) fills the E0 ghost rows of gaus(0),
    gather(stream) publishes this rank's rows of gaus(G) into every rank's full copy.  With hb + torch.distributed
    they are the peer-to-peer kernels (enable_p2p); tests emulate all ranks in one process with device copies."""

    def __init__(self, plan, device, hb=None, stride_align=64):
        import torch
        self.plan, self.hb_mod = plan, hb
        self.gaus, self.lap = [], []
        for l in range(plan.depth):
            a, b = plan.buffer_span(l)
            w = plan.width >> l
            stride = (w + stride_align - 1) // stride_align * stride_align
            for dst in (self.gaus, self.lap):
                if hb is None:
                    dst.append(torch.zeros((b - a, stride), dtype=torch.float32, device=device))
                else:   # whole CUDA allocations: exportable through CUDA IPC
                    dst.append(hb.alloc_image(A.F32, stride, b - a, device=device, align_bytes=4 * stride_align).zero_())
        self.halo0 = self.gatherG = None

    def owned(self, pyr, l):
        """this rank's rows of level l (levels >= G: its share of the replicated image)"""
        p = self.plan
        a, _ = p.buffer_span(l)
        return pyr[l][p.y0(l) - a:p.y1(l) - a, :p.width >> l]

    def strip0_plan(self):
        p = self.plan
        return StripPlan(p.width, p.height, p.world, p.rank, p.E0, A.CLAMP)

    def enable_p2p(self, hb, group=None):
        p = self.plan
        if p.world == 1:
            return
        self.halo0 = P2PHalo(hb, self.gaus[0], self.strip0_plan(), group)
        if p.G < p.depth:
            self.gatherG = P2PGather(hb, self.gaus[p.G], p.y0(p.G), p.rows(p.G), p.width >> p.G, p.world, p.rank, group)

    def _t(self, pyr, l, rows, ghost):
        roi, g = self.plan.view_args(l, rows, ghost)
        return (pyr[l][:, :self.plan.width >> l], roi, g)

    def traverse(self, hb, mask, stream=None, overlap=True, marks=None):
        """Gaussian_Laplacian_Pyramid/src/main.cpp:199-248 on this rank's strips: 1 halo exchange + 1 all-gather.
        overlap: the level-0 halo exchange runs on a side stream while the first down step works on the rows that
        need no ghost rows; its two edge bands follow the join (fork / join events, capturable into a CUDA graph)."""
        import torch
        mark = (lambda i, s_: hb.timestamp(marks, i, s_)) if marks is not None else (lambda i, s_: None)
        p = self.plan
        st = torch.cuda.current_stream() if stream is None else stream
        mark(0, st)
        split = p.split_dog and self.halo0 is not None
        if self.halo0 is None:
            self.down_sharded(hb, mask, st)
        elif not overlap or not self._can_split0():
            self.halo0.exchange(st)
            mark(1, st)
            self._down_region(hb, mask, 1, p.span(1, p.e[1]), st, dog=not split)
            mark(2, st); mark(3, st)
        else:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.gaus[0].device)
                self._ev = (torch.cuda.Event(), torch.cuda.Event())
            self._ev[0].record(st)
            self._side.wait_event(self._ev[0])
            self.halo0.exchange(self._side)
            mark(1, self._side)
            # the two edge bands follow the exchange on the side stream and overlap the tail of the interior band
            self._down0(hb, mask, self._side, "edges", dog=not split)
            mark(3, self._side)
            self._ev[1].record(self._side)
            self._down0(hb, mask, st, "interior", dog=not split)
            mark(2, st)
            st.wait_event(self._ev[1])
        if self.halo0 is not None:
            if split:
                # critical chain on `st`: blur+subsample only.  The DoG of level l-1 needs gaus(l) (just produced) and runs
                # on the DoG stream; the up step that consumes lap(l-1) waits for it (coarse_and_up).
                if self._dog_stream is None:
                    self._dog_stream = torch.cuda.Stream(device=self.gaus[0].device)
                    self._dog_ev = [torch.cuda.Event() for _ in range(2 * p.depth)]
                for l in range(1, min(p.G, p.depth - 1) + 1):
                    if l > 1:
                        self._down_region(hb, mask, l, p.span(l, p.e[l]), st, dog=False)
                    self._dog_ev[l].record(st)
                    self._dog_stream.wait_event(self._dog_ev[l])
                    c_rows = p.span(l, p.e[l] - 1)
                    f_rows = (2 * c_rows[0], 2 * c_rows[1])
                    hb.pyr_dog(self._t(self.gaus, l - 1, f_rows, 0), self._t(self.gaus, l, c_rows, 1), self._t(self.lap, l - 1, f_rows, 0), stream=self._dog_stream)
                    self._dog_ev[p.depth + l].record(self._dog_stream)
            else:
                self.down_sharded(hb, mask, st, first=2)
        mark(4, st)
        if self.gatherG is not None:
            self.gatherG.gather(st)
        mark(5, st)
        self.coarse_and_up(hb, mask, st, marks=marks, wait_dog=split)
        mark(8, st)

    _dog_stream = None
    _dog_ev = None

    MARK_NAMES = ["start", "exchange0 done", "down L0 interior done", "down L0 edges done", "down L1..G done", "gather done",
                  "replicated levels done", "up to level 1 done", "end"]

    _side = None
    _ev = None
    SPLIT_ROWS = 2   # coarse rows next to a strip edge whose fine footprint (+ K ghost rows) reaches into the ghost rows

    def _can_split0(self):
        p = self.plan
        return p.world > 1 and p.depth > 1 and p.rows(1) > 4 * self.SPLIT_ROWS + 64

    def _down_region(self, hb, mask, l, c_rows, stream, dog=True):
        p = self.plan
        f_rows = (2 * c_rows[0], 2 * c_rows[1])
        hb.pyr_down(self._t(self.gaus, l - 1, f_rows, p.K), self._t(self.gaus, l, c_rows, 0), mask,
                    lap_fine=self._t(self.lap, l - 1, f_rows, 0) if dog else None, stream=stream)

    def _down0(self, hb, mask, stream, which, dog=True):
        """the first down step (level 0 -> 1) in three row bands: the interior band reads only rows this rank owns"""
        p = self.plan
        full = p.span(1, p.e[1])
        s = self.SPLIT_ROWS
        lo = p.y0(1) + s if p.top else full[0]      # at the global image edge there is no ghost band
        hi = p.y1(1) - s if p.bot else full[1]
        if which == "interior":
            self._down_region(hb, mask, 1, (lo, hi), stream, dog)
        else:
            if p.top:
                self._down_region(hb, mask, 1, (full[0], lo), stream, dog)
            if p.bot:
                self._down_region(hb, mask, 1, (hi, full[1]), stream, dog)

    def down_sharded(self, hb, mask, stream=None, first=1):
        """way down through the sharded levels (needs the E0 ghost rows of gaus(0)); ends with this rank's rows of
        gaus(G) (+ e[G] extension rows) in its full copy of level G"""
        p = self.plan
        if p.world == 1:
            return
        for l in range(first, min(p.G, p.depth - 1) + 1):
            self._down_region(hb, mask, l, p.span(l, p.e[l]), stream)

    def coarse_and_up(self, hb, mask, stream=None, fuse_coarse=False, marks=None, wait_dog=False):
        """the replicated coarse levels (needs the gathered gaus(G)) and the whole way up: no communication"""
        p = self.plan
        first = 1 if p.world == 1 else p.G + 1
        top_up = p.depth - 2
        if p.world > 1 and p.depth - p.G >= 2 and fuse_coarse:
            # the replicated levels G .. depth-1 in ONE cooperative launch (hb_pyr_traverse_coarse)
            lv = lambda pyr: [pyr[l][:, :p.width >> l] for l in range(p.G, p.depth)]  # noqa: E731
            if hb.pyr_traverse_coarse(lv(self.gaus), lv(self.lap), mask, stream=stream):
                first, top_up = p.depth, p.G - 1
        for l in range(first, p.depth):
            w0, w1 = p.width >> (l - 1), p.width >> l
            hb.pyr_down(self.gaus[l - 1][:, :w0], self.gaus[l][:, :w1], mask, lap_fine=self.lap[l - 1][:, :w0], stream=stream)
        for l in range(top_up, -1, -1):
            if marks is not None and l == p.G - 1:
                hb.timestamp(marks, 6, stream)
            if marks is not None and l == 0:
                hb.timestamp(marks, 7, stream)
            if l < p.G and p.world > 1:
                if wait_dog:   # lap(l) comes from the DoG stream
                    stream.wait_event(self._dog_ev[p.depth + l + 1])
                f_rows = p.span(l, p.f[l])
                c_rows = (f_rows[0] // 2, f_rows[1] // 2)
                hb.pyr_up(self._t(self.gaus, l + 1, c_rows, 1), self._t(self.lap, l + 1, c_rows, 1),
                          self._t(self.gaus, l, f_rows, 0), self._t(self.lap, l, f_rows, 0), stream=stream)
            else:
                w0, w1 = p.width >> l, p.width >> (l + 1)
                hb.pyr_up(self.gaus[l + 1][:, :w1], self.lap[l + 1][:, :w1], self.gaus[l][:, :w0], self.lap[l][:, :w0], stream=stream)
