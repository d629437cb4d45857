"""Peer-mapped device memory and the exchange kernels of csrc/peer.cu.

Two providers of buffers that every rank can address:

* :class:`SymmetricPeers` -- one process per GPU (``torch.distributed``): buffers come from torch symmetric
  memory, whose rendezvous maps every rank's buffer into this process (NVLink 5 / NVSwitch peer access).  torch
  only allocates and maps; every byte that crosses a link is moved by a kernel of ``libmetalens_b200.so``.
* :class:`VirtualPeers` -- G "virtual ranks" on ONE device, each on its own stream, their buffers plain tensors
  of that device.  The kernels cannot tell the difference (a peer pointer is a pointer), so the whole exchange
  protocol -- scatter stores, flag barrier, pushed all-gather -- is exercised on a single-GPU box.

A :class:`PeerChannel` holds the flag block (peer-visible) and the epoch state (local) of one stream of exchanges
and wraps ``mlb_peer_barrier`` / ``mlb_peer_allgather`` / ``mlb_peer_wait``.
"""
import ctypes as C

import torch

from . import _lib


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SymmetricPeers:
    """Rank-local view of peer-mapped memory across a ``torch.distributed`` group (collective calls)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._keep = []

    def alloc(self, name, nbytes):
        """-> (local uint8 tensor of nbytes, zeroed; [address of rank p's buffer in this process])."""
        import torch.distributed._symmetric_memory as symm_mem
        nbytes = (int(nbytes) + 15) // 16 * 16
        t = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        t.zero_()
        hdl = symm_mem.rendezvous(t, self.group)
        try:
            ptrs = [int(p) for p in hdl.buffer_ptrs]
        except AttributeError:          # older handle API: map each peer's buffer as a tensor and take its address
            ptrs = [hdl.get_buffer(r, (nbytes,), torch.uint8).data_ptr() for r in range(self.world)]
        assert ptrs[self.rank] == t.data_ptr()
        self._keep.append((t, hdl))
        return t, ptrs

    def sync(self):
        """All ranks' allocations are zeroed and mapped before anyone stores into them."""
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(self.group)


class VirtualPeers:
    """`world` virtual ranks on one device.  ``view(rank)`` gives what SymmetricPeers gives a real rank."""

    def __init__(self, world, device=None):
        self.world = int(world)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._bufs = {}

    def _alloc(self, name, nbytes):
        if name not in self._bufs:
            nbytes = (int(nbytes) + 15) // 16 * 16
            self._bufs[name] = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
        bufs = self._bufs[name]
        assert bufs[0].numel() >= nbytes
        return bufs

    def view(self, rank):
        outer = self

        class _View:
            world = outer.world
            device = outer.device
            virtual = True           # every rank's waiting kernels share ONE GPU: exchanges must use few CTAs

            def __init__(self):
                self.rank = rank

            def alloc(self, name, nbytes):
                bufs = outer._alloc(name, nbytes)
                return bufs[rank], [b.data_ptr() for b in bufs]

            def sync(self):
                torch.cuda.synchronize()
        return _View()


class PeerChannel:
    """Flag block + epoch state of one stream of exchanges between the ranks of `peers`."""

    def __init__(self, peers, name="chan"):
        self.lib = _lib.load()
        self.peers = peers
        self.rank, self.world = peers.rank, peers.world
        if self.world > self.lib.mlb_peer_flag_words() // 3:
            raise _lib.MetalensB200Error("at most %d peers" % (self.lib.mlb_peer_flag_words() // 3))
        self.flags, self.flag_ptrs = peers.alloc(name + ".flags", 4 * self.lib.mlb_peer_flag_words())
        self.state = torch.zeros(self.lib.mlb_peer_state_words(), dtype=torch.int32, device=peers.device)
        self._pflags, self._k = _lib.ptr_array(self.flag_ptrs)

    def barrier(self):
        """Device-side barrier across the ranks on the current stream (no-op for one rank)."""
        if self.world == 1:
            return
        _lib.check(self.lib.mlb_peer_barrier(self._pflags, self.rank, self.world, self.state.data_ptr(), _stream_ptr()),
                   "mlb_peer_barrier")

    def allgather(self, src_ptr, src_pitch, rows, row_bytes, dst_ptrs, dst_pitch, dst_offset, aux=None, n_ctas=0):
        """Push the local block into every rank's destination (mlb_peer_allgather) on the current stream.
        aux = (src tensor float64, [dst addresses], offset in doubles, count) travels along."""
        pd, keep = _lib.ptr_array(dst_ptrs)
        if aux is None:
            a_src, a_dst, a_off, a_n, keep2 = None, None, 0, 0, None
        else:
            a_src, a_ptrs, a_off, a_n = aux
            a_dst, keep2 = _lib.ptr_array(a_ptrs)
        _lib.check(self.lib.mlb_peer_allgather(src_ptr, src_pitch, rows, row_bytes, pd, dst_pitch, dst_offset,
                                               a_src, a_dst, a_off, a_n, self._pflags, self.rank, self.world,
                                               self.state.data_ptr(), n_ctas, _stream_ptr()), "mlb_peer_allgather")

    def wait(self):
        """The peers' pushes of the latest allgather have landed here (device-side, current stream)."""
        _lib.check(self.lib.mlb_peer_wait(self.flags.data_ptr(), self.world, self.state.data_ptr(), _stream_ptr()),
                   "mlb_peer_wait")

    def wait_sum(self, block_sums, n, scale, out):
        """wait() and out[0] = scale * sum(block_sums[:n]) in one launch (same fixed order as mlb_sum_f64)."""
        _lib.check(self.lib.mlb_peer_wait_sum(self.flags.data_ptr(), self.world, self.state.data_ptr(), block_sums.data_ptr(),
                                              n, scale, out.data_ptr(), _stream_ptr()), "mlb_peer_wait_sum")

    def check(self):
        """Raise if any wait of this channel ever timed out (synchronises the device)."""
        if int(self.state[3].item()) != 0:
            raise _lib.MetalensB200Error("peer exchange timed out waiting for another rank")


class NcclComm:
    """NCCL communicator of the C-ABI (mlb_comm_*): the exchange steps for callers without peer mappings.  The
    128-byte unique id travels through the launcher's own channel (here: torch.distributed's store)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        box = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            _lib.check(self.lib.mlb_comm_unique_id(buf), "mlb_comm_unique_id")
            box[0] = buf.raw
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        handle = C.c_void_p()
        _lib.check(self.lib.mlb_comm_init(self.rank, self.world, box[0], C.byref(handle)), "mlb_comm_init")
        self.handle = handle

    def allgather_P(self, send, recv):
        assert send.dtype == torch.float32 and recv.dtype == torch.float32 and send.is_contiguous() and recv.is_contiguous()
        assert recv.numel() == self.world * send.numel()
        _lib.check(self.lib.mlb_allgather_P(self.handle, send.data_ptr(), recv.data_ptr(), send.numel(), _stream_ptr()),
                   "mlb_allgather_P")

    def allgather_fields(self, send, recv):
        assert send.dtype == torch.complex64 and recv.dtype == torch.complex64 and send.is_contiguous() and recv.is_contiguous()
        assert recv.numel() == self.world * send.numel()
        _lib.check(self.lib.mlb_allgather_fields(self.handle, send.data_ptr(), recv.data_ptr(), send.numel(),
                                                 _stream_ptr()), "mlb_allgather_fields")

    def allreduce_scalar(self, t):
        assert t.dtype == torch.float64 and t.is_contiguous()
        _lib.check(self.lib.mlb_allreduce_scalar(self.handle, t.data_ptr(), t.data_ptr(), t.numel(), _stream_ptr()),
                   "mlb_allreduce_scalar")
        return t

    def destroy(self):
        if self.handle:
            self.lib.mlb_comm_destroy(self.handle)
            self.handle = None
