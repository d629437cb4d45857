"""Multi-GPU sharding of the far field: one process per GPU, far-field tiles
distributed over the ranks, ONE all-gather of the finished power tiles at the end.

A *tile* is a slab of far-field rows (a contiguous range of ux) of one batch item
(wavelength / polarisation / source).  Tiles are independent -- the reference's own RAM chunk
loop (nearfield_farfield.py:45-66) already computes disjoint slabs of the far field
separately -- so the only exchange step on the path is assembling the result.  With B items
and G ranks every item is cut into S = G / gcd(B, G) slabs, giving B*S tiles, B*S/G per rank:
whole items per rank when G divides B (no aperture replication), finer slabs otherwise.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the plumbing; the per-tile compute
is a FarfieldPlan restricted to its rows.
"""
import math
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Tile:
    item: int       # batch item
    slab: int       # slab number within the item
    row0: int       # first far-field row (ux index) of the tile
    row1: int       # one past the last row


def tile_schedule(n_items, n_rows, world):
    """Tiles of every rank: list (len world) of lists of Tile.  Requires n_rows % S == 0."""
    assert n_items >= 1 and world >= 1
    slabs = world // math.gcd(n_items, world)
    if n_rows % slabs:
        raise ValueError("far-field rows (%d) must divide into %d equal slabs" % (n_rows, slabs))
    per = n_rows // slabs
    tiles = [Tile(i, s, s * per, (s + 1) * per) for i in range(n_items) for s in range(slabs)]
    per_rank = len(tiles) // world
    assert per_rank * world == len(tiles)
    return [tiles[r * per_rank:(r + 1) * per_rank] for r in range(world)]


def gather_tiles(local, n_items, n_rows, world, group=None, out=None, async_op=False):
    """All-gather the per-rank tile stacks into the full far field on every rank.

    local : (tiles_per_rank, rows_per_tile, n_cols) tensor of this rank's tiles in schedule order.
    Returns (n_items, n_rows, n_cols) [, work handle when async_op].  Tiles are ordered item-major,
    slab-minor and ranks own consecutive tiles, so the gathered buffer IS the result: no reshuffle
    after the collective.
    """
    t, rows, cols = local.shape
    if out is None:
        out = torch.empty((world * t, rows, cols), dtype=local.dtype, device=local.device)
    work = None
    if world == 1:
        out.copy_(local)
    else:
        work = dist.all_gather_into_tensor(out, local.contiguous(), group=group, async_op=async_op)
    res = out.view(n_items, n_rows, cols)
    return (res, work) if async_op else res


class ShardedFarfield:
    """Far field of a batch of apertures computed across `world` ranks.

    make_plan(item, row0, row1) must return a FarfieldPlan-like object whose ``run(fields)``
    yields (P tile (row1-row0, Ky) device tensor, total_P scalar tensor); ``fields_of(item)``
    the four device fields of an item this rank owns.
    """

    def __init__(self, n_items, n_rows, make_plan, rank=None, world=None, group=None, tail_priority=0, gather="auto"):
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank
        self.group = group
        self.n_items, self.n_rows = n_items, n_rows
        self.schedule = tile_schedule(n_items, n_rows, self.world)
        self.tiles = self.schedule[self.rank]
        self.plans = [make_plan(t.item, t.row0, t.row1) for t in self.tiles]
        # double-buffered tile stacks / gather targets so that the all-gather of step i can overlap the
        # kernels of step i+1 (overlap=True): a buffer is reused only after its collective has completed
        self._local = [None, None]
        self._out = [None, None]
        self._work = [None, None]
        self._flip = 0
        # overlap=True pipelines the tiles over two streams: the HBM-bound pass over each aperture runs on the
        # caller's stream, everything after it (column pass, power epilogue, tile copy, the collective) on a side
        # stream, so the tail of tile k executes under the aperture pass of tile k+1
        self._side = None
        self._tail_priority = tail_priority             # side stream priority (-1 = high; measured: no effect on B200)
        self._tail_done = [None] * len(self.tiles)     # per plan: its buffers are free again (side-stream event)
        self._side_done = None
        # gather="p2p": the finished tiles are exchanged by PULLING them from the peers' buffers over NVLink with the
        # copy engines (torch symmetric memory: peer-mapped buffers + signal-pad barriers), no SM-resident collective
        # kernel competing with the persistent row pass; "nccl": all_gather_into_tensor; "auto": p2p when it can be
        # set up (CUDA, world > 1), else nccl
        assert gather in ("auto", "nccl", "p2p")
        self._gather = gather
        self._symm = None                               # per flip buffer: (handle, [peer tile-stack tensors])
        self._comm = None
        self._gather_done = [None, None]

    def _setup_p2p(self, shape, dtype, device):
        """Allocate the two local tile stacks in symmetric memory and map the peers' (collective call)."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        symm = []
        for b in (0, 1):
            t = symm_mem.empty(shape, dtype=dtype, device=device)
            hdl = symm_mem.rendezvous(t, group)
            peers = [hdl.get_buffer(r, shape, dtype) for r in range(self.world)]
            self._local[b] = t
            self._out[b] = torch.empty((self.world * shape[0],) + tuple(shape[1:]), dtype=dtype, device=device)
            symm.append((hdl, peers))
        self._symm = symm
        self._comm = torch.cuda.Stream(device=device)

    def _alloc(self, b, P):
        """Tile stack + gather target of flip buffer b (first use)."""
        shape = (len(self.tiles),) + tuple(P.shape)
        if self._gather != "nccl" and self.world > 1 and P.is_cuda and self._symm is None:
            try:
                self._setup_p2p(shape, P.dtype, P.device)
                self._gather = "p2p"
                return
            except Exception as e:                      # no symmetric memory here: the NCCL collective does the job
                if self._gather == "p2p":
                    raise
                import sys
                print("metalens_b200.sharding: peer-to-peer gather unavailable (%s: %s), using NCCL all-gather"
                      % (type(e).__name__, str(e)[:200]), file=sys.stderr)
                self._gather = "nccl"
                self._symm = None
        if self._local[b] is None:
            self._local[b] = torch.empty(shape, dtype=P.dtype, device=P.device)
            self._out[b] = torch.empty((self.world * shape[0],) + shape[1:], dtype=P.dtype, device=P.device)

    def _gather_p2p(self, b, producer_stream):
        """Pull every rank's tile stack into out[b] on the communication stream (copy engines), bracketed by the two
        barriers of torch's low-contention all-gather: all stacks ready before anyone pulls, all pulls done before
        anyone overwrites its stack.  Returns the (n_items, n_rows, Ky) view of out[b]; complete after finish()."""
        hdl, peers = self._symm[b]
        ready = torch.cuda.Event()
        ready.record(producer_stream)
        t = len(self.tiles)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ready)
            hdl.barrier(channel=b)
            for step in range(self.world):
                r = (self.rank - step) % self.world
                self._out[b][r * t:(r + 1) * t].copy_(peers[r], non_blocking=True)
            hdl.barrier(channel=b)
            done = torch.cuda.Event()
            done.record(self._comm)
        self._gather_done[b] = done
        return self._out[b].view(self.n_items, self.n_rows, self._out[b].shape[-1])

    @property
    def items_needed(self):
        return sorted({t.item for t in self.tiles})

    def run(self, fields_of, runner=None, overlap=False):
        """Compute this rank's tiles, then the single all-gather.  Returns
        (P (n_items, n_rows, Ky) on every rank, partial total_P per local tile).
        `runner(plan, fields)` defaults to ``plan.run(fields)``.  With overlap=True the collective is
        asynchronous (the result is complete after ``finish()`` or a device synchronize) and runs
        concurrently with the next call's kernels."""
        b = self._flip
        self._flip ^= 1
        if overlap and runner is None and all(hasattr(p, "run_split") for p in self.plans) \
                and self.plans and self.plans[0].P.is_cuda:
            return self._run_pipelined(fields_of, b)
        if self._work[b] is not None:            # the collective that last used this buffer pair
            self._work[b].wait()
            self._work[b] = None
        totals = []
        for k, (tile, plan) in enumerate(zip(self.tiles, self.plans)):
            P, total = plan.run(fields_of(tile.item)) if runner is None else runner(plan, fields_of(tile.item))
            if self._local[b] is None:
                self._alloc(b, P)
            if k == 0 and self._gather_done[b] is not None:      # peers may still be pulling this stack (p2p, two steps ago)
                torch.cuda.current_stream().wait_event(self._gather_done[b])
            self._local[b][k].copy_(P)
            totals.append(total)
        if self._gather == "p2p" and self.world > 1 and self._symm is not None:
            res = self._gather_p2p(b, torch.cuda.current_stream())
            if not overlap:
                torch.cuda.current_stream().wait_event(self._gather_done[b])
        elif overlap and self.world > 1:
            res, self._work[b] = gather_tiles(self._local[b], self.n_items, self.n_rows, self.world, self.group,
                                              out=self._out[b], async_op=True)
        else:
            res = gather_tiles(self._local[b], self.n_items, self.n_rows, self.world, self.group, out=self._out[b])
        return res, totals

    def _run_pipelined(self, fields_of, b):
        """overlap=True on CUDA: two-stream software pipeline over the local tiles (see __init__)."""
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.plans[0].P.device, priority=self._tail_priority)
        side = self._side
        totals = []
        for k, (tile, plan) in enumerate(zip(self.tiles, self.plans)):
            first, second = plan.run_split(fields_of(tile.item))
            if self._tail_done[k] is not None:           # the previous tail of this plan still reads its buffers
                main.wait_event(self._tail_done[k])
            first()
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                if k == 0 and self._work[b] is not None:     # the collective that last used this buffer pair
                    self._work[b].wait()
                    self._work[b] = None
                P, total = second()
                if self._local[b] is None:
                    self._alloc(b, P)
                if k == 0 and self._gather_done[b] is not None:  # peers may still be pulling this stack (p2p)
                    side.wait_event(self._gather_done[b])
                self._local[b][k].copy_(P)
                totals.append(total)
                done = torch.cuda.Event()
                done.record(side)
                self._tail_done[k] = done
        if self._gather == "p2p" and self.world > 1 and self._symm is not None:
            res = self._gather_p2p(b, side)
            with torch.cuda.stream(side):
                self._side_done = torch.cuda.Event()
                self._side_done.record(side)
            return res, totals
        with torch.cuda.stream(side):
            if self.world > 1:
                res, self._work[b] = gather_tiles(self._local[b], self.n_items, self.n_rows, self.world, self.group,
                                                  out=self._out[b], async_op=True)
            else:
                res = gather_tiles(self._local[b], self.n_items, self.n_rows, self.world, self.group, out=self._out[b])
            self._side_done = torch.cuda.Event()
            self._side_done.record(side)
        return res, totals

    def capture(self, fields_of):
        """Capture one pipelined step (every local tile, both streams, the tile copies) into a CUDA graph and
        return it; ``replay()`` then re-runs the step on whatever is in the same field buffers with a single
        launch.  Single-rank only (world == 1): with more ranks the collective stays an eager NCCL call."""
        if self.world != 1:
            raise ValueError("capture() is for world == 1; use run(overlap=True) across ranks")
        self.run(fields_of, overlap=True)                # warm-up outside capture (lazy allocations, attributes)
        self.finish()
        torch.cuda.synchronize()
        self._tail_done = [None] * len(self.tiles)       # no dependencies on events recorded outside the capture
        self._flip = 0
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._graph_result = self.run(fields_of, overlap=True)
            self.finish()                                # joins the side stream back into the capturing stream
        self._tail_done = [None] * len(self.tiles)
        self._graph = g
        return g

    def replay(self):
        """Replay the captured step; returns (P (n_items, n_rows, Ky), totals) like run()."""
        self._graph.replay()
        return self._graph_result

    def finish(self):
        """Wait for outstanding asynchronous all-gathers (and, in pipelined mode, make the caller's stream
        wait for the side stream)."""
        for b in (0, 1):
            if self._work[b] is not None:
                if self._side is not None:
                    with torch.cuda.stream(self._side):
                        self._work[b].wait()
                    self._side_done = torch.cuda.Event()
                    self._side_done.record(self._side)
                else:
                    self._work[b].wait()
                self._work[b] = None
        for b in (0, 1):
            if self._gather_done[b] is not None:
                torch.cuda.current_stream().wait_event(self._gather_done[b])
        if self._side_done is not None:
            torch.cuda.current_stream().wait_event(self._side_done)
            self._side_done = None
