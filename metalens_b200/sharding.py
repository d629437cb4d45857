"""Multi-GPU sharding of the far field: one process per GPU, far-field tiles
distributed over the ranks, ONE all-gather of the finished power tiles at the end.

A *tile* is a slab of far-field rows (a contiguous range of ux) of one batch item
(wavelength / polarisation / source).  Tiles are independent -- the reference's own RAM chunk
loop (nearfield_farfield.py:45-66) already computes disjoint slabs of the far field
separately -- so the only exchange step on the path is assembling the result.  With B items
and G ranks every item is cut into S = G / gcd(B, G) slabs, giving B*S tiles, B*S/G per rank:
whole items per rank when G divides B (no aperture replication), finer slabs otherwise.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the plumbing; the per-tile compute
is a FarfieldPlan restricted to its rows.  On GPUs the exchange itself is a kernel of this package: every rank
PUSHES its finished tiles into the peers' result buffers over NVLink (mlb_peer_allgather, csrc/peer.cu), or --
without peer mappings -- calls the C-ABI's NCCL wrapper (mlb_allgather_P).  One aperture spread over the ranks
(BASELINE config 4) is metalens_b200/slab.py.
"""
import math
from dataclasses import dataclass

import torch
import torch.distributed as dist


def _lib_error(msg):
    from ._lib import MetalensB200Error
    return MetalensB200Error(msg)


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

    def __init__(self, n_items, n_rows, make_plan, rank=None, world=None, group=None, tail_priority=0, gather="auto",
                 push_ctas=0):
        self._push_ctas = push_ctas                     # CTAs of the push kernel (0 = library default)
        self._pushed = False
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
        self._tail_done = {}                            # per plan object: its buffers are free again (side-stream event)
        self._side_done = None
        # How the finished tiles travel (CUDA, world > 1):
        #   "p2p" (default of "auto"): the peers' stacks are PULLED over NVLink by the copy engines (torch symmetric
        #           memory: peer-mapped stacks + two signal-pad barriers per step) -- no SM time next to the persistent,
        #           HBM-bound row pass.  Measured best for this overlapped exchange: 2 GPUs 0.358 ms/step vs 0.386 pushed;
        #   "push": mlb_peer_allgather -- CTAs of this library store the rank's stack into every peer's result buffer
        #           (flag epochs instead of barriers).  The mechanism of the one-aperture path (slab.py), where the
        #           exchange is on the critical path anyway; here its CTAs take issue slots from the row pass;
        #   "nccl": mlb_allgather_P (the C-ABI's NCCL wrapper) on a communication stream.
        # CPU tensors (gloo tests) always use torch's all_gather_into_tensor.
        assert gather in ("auto", "nccl", "p2p", "push")
        self._gather = gather
        self._symm = None                               # p2p, per flip buffer: (handle, [peer tile-stack tensors])
        self._chan = None                               # push: PeerChannel
        self._out_ptrs = [None, None]                   # push: peer addresses of the two result buffers
        self._nccl = None
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

    def _setup_push(self, shape, dtype, device):
        """Both result buffers in peer-mapped memory; a rank's tile stack IS its section of its own result buffer."""
        from .peer import PeerChannel, SymmetricPeers
        peers = SymmetricPeers(self.group, device)
        t = shape[0]
        nbytes = self.world * int(torch.tensor([], dtype=dtype).element_size()) * t * shape[1] * shape[2]
        for b in (0, 1):
            buf, ptrs = peers.alloc("tiles%d" % b, nbytes)
            self._out[b] = buf[:nbytes].view(dtype).view((self.world * t,) + tuple(shape[1:]))
            self._local[b] = self._out[b][self.rank * t:(self.rank + 1) * t]
            self._out_ptrs[b] = ptrs
        self._chan = PeerChannel(peers, "tiles.chan")
        self._stack_bytes = nbytes // self.world
        self._comm = torch.cuda.Stream(device=device)
        peers.sync()

    def _alloc(self, b, P):
        """Tile stack + gather target of flip buffer b (first use).  The exchange mechanism is agreed on by ALL ranks
        (a rank that cannot map peer memory makes everyone use NCCL: mixed protocols would hang)."""
        shape = (len(self.tiles),) + tuple(P.shape)
        if self.world > 1 and P.is_cuda and self._chan is None and self._symm is None and self._nccl is None:
            mode, err = self._gather, None
            if mode in ("auto", "push", "p2p"):
                try:
                    if mode == "push":
                        if shape[1] * shape[2] * P.element_size() % 16:
                            raise ValueError("tile bytes not a multiple of 16")
                        self._setup_push(shape, P.dtype, P.device)
                    else:
                        self._setup_p2p(shape, P.dtype, P.device)
                        mode = "p2p"
                except Exception as e:
                    err = e
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=P.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                if self._gather in ("push", "p2p"):
                    raise _lib_error("peer-memory tile exchange unavailable on some rank: %r" % (err,))
                import sys
                if err is not None:
                    print("metalens_b200.sharding: peer-memory exchange unavailable (%s: %s), using NCCL"
                          % (type(err).__name__, str(err)[:200]), file=sys.stderr)
                mode = "nccl"
                self._symm = self._chan = None
                self._local, self._out = [None, None], [None, None]
            self._gather = mode
            if mode == "nccl":
                from .peer import NcclComm
                self._nccl = NcclComm(self.group)
                self._comm = torch.cuda.Stream(device=P.device)
            if mode != "nccl":
                return
        if self._local[b] is None:
            self._local[b] = torch.empty(shape, dtype=P.dtype, device=P.device)
            self._out[b] = torch.empty((self.world * shape[0],) + shape[1:], dtype=P.dtype, device=P.device)

    def _bind(self, plan, b, k):
        """The plan writes tile k of flip buffer b straight into its slot of the tile stack (no copy)."""
        if self._local[b] is not None and hasattr(plan, "bind_output"):
            slot = self._local[b][k]
            if slot.is_cuda and tuple(slot.shape) == (plan.Kx, plan.Ky) and slot.dtype == plan.p_dtype:
                plan.bind_output(slot)

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

    def _gather_push(self, b, producer_stream):
        """mlb_peer_allgather on the communication stream, behind the producer's last tile copy: store this rank's
        stack into every peer's out[b].  Its own stream, because the kernel first waits for the peers to enter the
        same gather: that wait must not hold up the tails of the next step's tiles.  Nothing else waits here;
        finish() acquires the peers' completion flags."""
        ready = torch.cuda.Event()
        ready.record(producer_stream)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ready)
            self._chan.allgather(self._local[b].data_ptr(), self._stack_bytes, 1, self._stack_bytes, self._out_ptrs[b],
                                 self._stack_bytes, self.rank * self._stack_bytes, n_ctas=self._push_ctas)
            done = torch.cuda.Event()
            done.record(self._comm)
        self._gather_done[b] = done                     # the stack may be overwritten once its push has run
        self._pushed = True
        return self._out[b].view(self.n_items, self.n_rows, self._out[b].shape[-1])

    def _gather_nccl(self, b, producer_stream):
        ready = torch.cuda.Event()
        ready.record(producer_stream)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ready)
            self._nccl.allgather_P(self._local[b], self._out[b])
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
        `runner(plan, fields)` defaults to ``plan.run(fields)``.  With overlap=True the exchange is
        asynchronous (the result is complete after ``finish()``) and runs concurrently with the next call's
        kernels; the result buffers are double-buffered, so a result stays valid until the second-next run()."""
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
            if k == 0 and self._gather_done[b] is not None:      # peers may still be pulling this stack (two steps ago)
                torch.cuda.current_stream().wait_event(self._gather_done[b])
            self._bind(plan, b, k)
            P, total = plan.run(fields_of(tile.item)) if runner is None else runner(plan, fields_of(tile.item))
            if self._local[b] is None:
                self._alloc(b, P)
            if P.data_ptr() != self._local[b][k].data_ptr():
                self._local[b][k].copy_(P)
            totals.append(total)
        cuda = self.world > 1 and self._local[b].is_cuda
        if cuda and self._gather == "push":
            res = self._gather_push(b, torch.cuda.current_stream())
            if not overlap:
                torch.cuda.current_stream().wait_event(self._gather_done[b])
                self._chan.wait()
        elif cuda and self._gather in ("p2p", "nccl"):
            cur = torch.cuda.current_stream()
            res = self._gather_p2p(b, cur) if self._gather == "p2p" else self._gather_nccl(b, cur)
            if not overlap:
                cur.wait_event(self._gather_done[b])
        elif overlap and self.world > 1:
            res, self._work[b] = gather_tiles(self._local[b], self.n_items, self.n_rows, self.world, self.group,
                                              out=self._out[b], async_op=True)
        else:
            res = gather_tiles(self._local[b], self.n_items, self.n_rows, self.world, self.group, out=self._out[b])
        return res, totals

    def _run_pipelined(self, fields_of, b, exchange=True):
        """overlap=True on CUDA: two-stream software pipeline over the local tiles (see __init__).
        exchange=False stops after the tile copies (the captured part of a step, see capture())."""
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.plans[0].P.device, priority=self._tail_priority)
        side = self._side
        totals = []
        for k, (tile, plan) in enumerate(zip(self.tiles, self.plans)):
            self._bind(plan, b, k)
            first, second = plan.run_split(fields_of(tile.item))
            prev = self._tail_done.get(id(plan))
            if prev is not None:                         # the previous tail of this plan still reads its buffers
                main.wait_event(prev)
            first()
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                if k == 0 and self._work[b] is not None:     # the collective that last used this buffer pair
                    self._work[b].wait()
                    self._work[b] = None
                if k == 0 and self._gather_done[b] is not None:  # peers may still be reading this stack (two steps ago)
                    side.wait_event(self._gather_done[b])
                P, total = second()
                if self._local[b] is None:
                    self._alloc(b, P)
                if P.data_ptr() != self._local[b][k].data_ptr():
                    self._local[b][k].copy_(P)
                totals.append(total)
                done = torch.cuda.Event()
                done.record(side)
                self._tail_done[id(plan)] = done
        cuda = self.world > 1
        with torch.cuda.stream(side):
            if not exchange:
                res = self._out[b].view(self.n_items, self.n_rows, self._out[b].shape[-1])
                if self.world == 1:
                    gather_tiles(self._local[b], self.n_items, self.n_rows, 1, self.group, out=self._out[b])
            elif cuda and self._gather == "push":
                res = self._gather_push(b, side)
            elif cuda and self._gather == "p2p":
                res = self._gather_p2p(b, side)
            elif cuda and self._gather == "nccl":
                res = self._gather_nccl(b, side)
            elif self.world > 1:
                res, self._work[b] = gather_tiles(self._local[b], self.n_items, self.n_rows, self.world, self.group,
                                                  out=self._out[b], async_op=True)
            else:
                res = gather_tiles(self._local[b], self.n_items, self.n_rows, self.world, self.group, out=self._out[b])
            self._side_done = torch.cuda.Event()
            self._side_done.record(side)
        return res, totals

    def capture(self, fields_of):
        """Capture the pipelined step (every local tile, both streams, the tile copies) into CUDA graphs, one per
        flip buffer; ``replay()`` then re-runs the step on whatever is in the same field buffers with a single graph
        launch.  Across ranks (push exchange only) the exchange stays OUTSIDE the graphs: replay() launches the push
        kernel on the side stream behind the graph, so it still runs under the next step's graph."""
        if self.world != 1 and self._gather not in ("auto", "push"):
            raise ValueError("capture() across ranks needs the push exchange")
        for _ in range(2):                               # warm-up outside capture (lazy allocations, both flip buffers)
            self.run(fields_of, overlap=True)
        self.finish()
        torch.cuda.synchronize()
        if self.world != 1 and self._gather != "push":
            raise ValueError("capture() across ranks needs the push exchange (got %s)" % self._gather)
        self._gather_done = [None, None]                 # recorded outside the capture: nothing to wait for after the sync
        self._graphs, self._graph_results = [], []
        for b in (0, 1):
            self._tail_done = {}                         # no dependencies on events recorded outside the capture
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                res = self._run_pipelined(fields_of, b, exchange=False)
                torch.cuda.current_stream().wait_event(self._side_done)   # join the side stream back
                self._side_done = None
            self._graphs.append(g)
            self._graph_results.append(res)
        self._tail_done = {}
        self._flip = 0
        return self._graphs

    def replay(self):
        """Replay the captured step; returns (P (n_items, n_rows, Ky), totals) like run(overlap=True)."""
        b = self._flip
        self._flip ^= 1
        main = torch.cuda.current_stream()
        if self._gather_done[b] is not None:             # the push that last read this tile stack
            main.wait_event(self._gather_done[b])
        self._graphs[b].replay()
        if self.world > 1:
            self._gather_push(b, main)
        return self._graph_results[b]

    def finish(self):
        """Wait for outstanding asynchronous exchanges (and, in pipelined mode, make the caller's stream
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
        if self._pushed:                                 # the peers' pushes of the latest step have landed here
            self._chan.wait()

    def check(self):
        """Raise if a peer exchange ever timed out (synchronises)."""
        if self._chan is not None:
            self._chan.check()
