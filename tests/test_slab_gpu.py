"""GPU: ONE aperture over several ranks (BASELINE config 4 shape, metalens_b200/slab.py) and the peer exchange
kernels (csrc/peer.cu) through the C-ABI.

The exchange kernels only see pointers, so G "virtual ranks" on one device (each on its own stream, buffers of
the same device standing in for peer-mapped ones) run the complete protocol -- scattering row pass, flag barrier,
column pass, pushed all-gather -- on a single-GPU box; with >= 2 devices the same test also runs as real ranks
over NVLink (torch symmetric memory supplies the mappings).  Bar: P and total_P BIT-identical to the single-GPU
FarfieldPlan, which test_farfield_gpu.py pins to the oracle / reference fixtures.
"""
import os
import socket

import numpy as np
import pytest

import apertures

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
WL = 532e-9
NG = apertures.N_GLASS[532]


def _same(a, b):
    return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())


def _single_gpu(fields, M, stride):
    """Single-GPU far field with the same kernels the ranks run: the TMA-fed radix-4 row pass (without a fold the
    single-GPU default would be the radix-16 register kernel, equal only to rounding) + fused column/power pass."""
    from metalens_b200 import _lib
    from metalens_b200.farfield import FarfieldPlan
    lib = _lib.load()
    d = WL / 2.2
    plan = FarfieldPlan((M, M), d, d, WL, NG, stride=stride, method="fft", fuse_power="always")
    assert plan.fused
    old = lib.mlb_get_option(b"rows_engine")
    lib.mlb_set_option(b"rows_engine", 0)
    try:
        P, total = plan.run(fields)
    finally:
        lib.mlb_set_option(b"rows_engine", old)
    return P.clone(), total.clone(), plan


@pytest.mark.parametrize("M,stride,world", [(1024, 4, 1), (1024, 4, 2), (1024, 4, 4), (2048, 2, 4), (2048, 8, 2),
                                            (1024, 1, 4), (4096, 2, 8)])
def test_virtual_ranks_equal_single_gpu(M, stride, world):
    from metalens_b200.peer import VirtualPeers
    from metalens_b200.slab import SlabFarfield
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 7 + world, WL, NG, rotate=bool(world & 2))
    full = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
    P_ref, total_ref, _ = _single_gpu(full, M, stride)
    d = WL / 2.2
    vp = VirtualPeers(world)
    slabs = [SlabFarfield((M, M), d, d, WL, NG, stride, vp.view(r)) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    local = []
    for s in slabs:
        idx = torch.from_numpy(s.x_rows).cuda()
        local.append([f.index_select(0, idx).contiguous() for f in full])
    covered = np.sort(np.concatenate([s.x_rows for s in slabs]))
    assert np.array_equal(covered, np.arange(M))                       # every aperture row on exactly one rank
    for r, s in enumerate(slabs):
        s.warm(local[r])
    torch.cuda.synchronize()
    for rep in range(3):                                               # epochs advance; buffers are reused
        for r, s in enumerate(slabs):
            with torch.cuda.stream(streams[r]):
                s.run(local[r], wait=False)
        for r, s in enumerate(slabs):
            with torch.cuda.stream(streams[r]):
                s.finish()
        torch.cuda.synchronize()
        for s in slabs:
            s.chan.check()
            assert _same(s.P, P_ref), (rep, s.rank)
            if s.exact_total:
                assert float(s.total) == float(total_ref)
            else:
                assert abs(float(s.total) - float(total_ref)) <= 1e-12 * abs(float(total_ref))


def test_slab_rejects_unsupported_shapes():
    from metalens_b200.peer import VirtualPeers
    from metalens_b200.slab import SlabFarfield
    d = WL / 2.2
    with pytest.raises(ValueError):
        SlabFarfield((900, 900), d, d, WL, NG, 1, VirtualPeers(2).view(0))        # not a power of two
    with pytest.raises(ValueError):
        SlabFarfield((1024, 1024), d, d, WL, NG, 4, VirtualPeers(3).view(0))      # 3 ranks


def test_pushed_allgather_epochs():
    """mlb_peer_allgather / mlb_peer_wait / mlb_peer_barrier alone: 4 virtual ranks, 2-D blocks with a pitch, an aux
    block, five epochs with fresh data, one rank deliberately late."""
    from metalens_b200.peer import PeerChannel, VirtualPeers
    world, rows, cols = 4, 37, 64                                      # 64 floats = 256 bytes per row segment
    vp = VirtualPeers(world)
    views = [vp.view(r) for r in range(world)]
    chans = [PeerChannel(v, "t") for v in views]
    dst = [v.alloc("dst", rows * cols * world * 4) for v in views]
    aux = [v.alloc("aux", 8 * 3 * world) for v in views]
    streams = [torch.cuda.Stream() for _ in range(world)]
    late = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    late.normal_()                     # first launches load their kernels: none may happen while ranks wait for each other
    torch.cuda.synchronize()
    for epoch in range(5):
        srcs = []
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                if r == 2:
                    late.normal_()                                      # this rank arrives late at the gather
                full = dst[r][0][:rows * cols * world * 4].view(torch.float32).view(rows, cols * world)
                mine = torch.full((rows, cols), float(10 * epoch + r), device="cuda") + \
                    torch.arange(cols, device="cuda")[None, :] / 1000.0
                full[:, r * cols:(r + 1) * cols].copy_(mine)
                a = torch.tensor([epoch, r, 7.0], dtype=torch.float64, device="cuda")
                aux[r][0][:8 * 3 * world].view(torch.float64)[3 * r:3 * r + 3].copy_(a)
                srcs.append((mine, a))
                chans[r].barrier()
                chans[r].allgather(full.data_ptr() + 4 * r * cols, 4 * cols * world, rows, 4 * cols,
                                   dst[r][1], 4 * cols * world, 4 * r * cols,
                                   aux=(aux[r][0].data_ptr() + 8 * 3 * r, aux[r][1], 3 * r, 3))
                chans[r].wait()
        torch.cuda.synchronize()
        for r in range(world):
            chans[r].check()
            full = dst[r][0][:rows * cols * world * 4].view(torch.float32).view(rows, cols * world)
            av = aux[r][0][:8 * 3 * world].view(torch.float64)
            for p in range(world):
                assert torch.equal(full[:, p * cols:(p + 1) * cols], srcs[p][0]), (epoch, r, p)
                assert torch.equal(av[3 * p:3 * p + 3], srcs[p][1])


def test_assemble_slab_rows_equal_full_assembly(golden_dir):
    """Hot path B on a rank's rows only == those rows of the full assembly (bit for bit); partial incident powers
    add up to the full lens."""
    import synth_lens
    from metalens_b200 import grating, lens_center
    from metalens_b200.design import make_design
    from metalens_b200.nearfield import NearfieldPlan
    from metalens_b200.peer import VirtualPeers
    from metalens_b200.slab import SlabFarfield, assemble_slab
    wl = 580e-9
    M = 1024
    R = M * (wl / 2.2) / 2
    f = R / np.tan(np.radians(40.0))
    spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 42.0, 650e-9, 1.1)], source_distance=f, radius=R * 0.999)
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periph, center, _ = make_design(collections, f, spec["radius"], hgs)
    plan = NearfieldPlan(wl, periph, center, hgs)
    x = np.linspace(-R, R, M)
    full, power = plan.run(0.0, 0.0, -f, "y", x, x)
    full = full.clone()
    world = 4
    vp = VirtualPeers(world)
    d = float(x[1] - x[0])
    parts = 0.0
    for r in range(world):
        slab = SlabFarfield((M, M), d, d, wl, plan.n_glass, 4, vp.view(r))
        loc, p = assemble_slab(plan, slab, (0.0, 0.0, -f), "y", x, x)
        idx = torch.from_numpy(slab.x_rows).cuda()
        assert torch.equal(loc.view(torch.float32), full.index_select(1, idx).contiguous().view(torch.float32))
        parts += float(p)
    assert abs(parts - float(power)) <= 1e-12 * abs(float(power))


# ----------------------------------------------------------------------------- real ranks (>= 2 GPUs)
def _worker(rank, world, port, M, stride, q):
    import torch.distributed as dist
    from metalens_b200.peer import NcclComm, SymmetricPeers
    from metalens_b200.slab import SlabFarfield
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 11, WL, NG)
        full = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
        P_ref, total_ref, _ = _single_gpu(full, M, stride)
        d = WL / 2.2
        slab = SlabFarfield((M, M), d, d, WL, NG, stride, SymmetricPeers())
        idx = torch.from_numpy(slab.x_rows).cuda()
        local = [f.index_select(0, idx).contiguous() for f in full]
        ok = True
        for rep in range(3):
            P, total = slab.run(local)
            torch.cuda.synchronize()
            ok = ok and _same(P, P_ref) and float(total) == float(total_ref)
        slab.chan.check()
        # the C-ABI's NCCL wrappers
        comm = NcclComm()
        t = torch.tensor([1.0 + rank, 2.0], dtype=torch.float64, device="cuda")
        comm.allreduce_scalar(t)
        send = torch.full((5,), float(rank), dtype=torch.float32, device="cuda")
        recv = torch.empty(5 * world, dtype=torch.float32, device="cuda")
        comm.allgather_P(send, recv)
        sc = torch.full((3,), complex(rank, -rank), dtype=torch.complex64, device="cuda")
        rc = torch.empty(3 * world, dtype=torch.complex64, device="cuda")
        comm.allgather_fields(sc, rc)
        torch.cuda.synchronize()
        ok = ok and t.tolist() == [sum(1.0 + r for r in range(world)), 2.0 * world]
        ok = ok and recv.view(world, 5)[:, 0].tolist() == [float(r) for r in range(world)]
        ok = ok and rc.view(world, 3)[:, 0].tolist() == [complex(r, -r) for r in range(world)]
        comm.destroy()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M,stride", [(1024, 4), (2048, 2)])
def test_real_ranks_equal_single_gpu(M, stride):
    world = 2 if torch.cuda.device_count() < 4 else 4
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the virtual-rank test covers the protocol on one)")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, stride, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
