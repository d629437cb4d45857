"""GPU (needs >= 2 devices, skipped otherwise): far-field tiles sharded over 2 ranks equal the single-GPU far
field bit for bit -- whole items per rank (fft) and row slabs of one item (fold), with both exchange
mechanisms (peer-to-peer pulls through symmetric memory, NCCL all-gather)."""
import os
import socket

import numpy as np
import pytest

import apertures

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
WL = 532e-9
NG = apertures.N_GLASS[532]
M = 256


def _fields(item):
    return apertures.focusing_lens(M, 50 + item, WL, NG, rotate=bool(item % 2))


def _worker(rank, world, port, n_items, method, gather, q):
    import torch.distributed as dist
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.sharding import ShardedFarfield
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        d = WL / 2.2
        K = M // 4

        def make_plan(item, r0, r1):
            rows = None if (r0, r1) == (0, K) else (r0, r1)
            return FarfieldPlan((M, M), d, d, WL, NG, stride=4, method=method, rows=rows)
        graph = gather == "graph"
        gather = "push" if graph else gather
        sh = ShardedFarfield(n_items, K, make_plan, gather=gather)
        dev = {i: [torch.from_numpy(a).cuda() for a in _fields(i)[:4]] for i in sh.items_needed}
        P, _ = sh.run(lambda item: dev[item])
        torch.cuda.synchronize()
        P = P.clone()
        # overlapped (asynchronous, double-buffered) gathers give the same result
        if graph:
            sh.capture(lambda item: dev[item])
            outs = [sh.replay()[0] for _ in range(5)]
        else:
            outs = [sh.run(lambda item: dev[item], overlap=True)[0] for _ in range(3)]
        sh.finish()
        torch.cuda.synchronize()
        sh.check()
        assert all(bool(((o == P) | (torch.isnan(o) & torch.isnan(P))).all()) for o in outs[-2:])
        ref = []
        for i in range(n_items):
            plan = FarfieldPlan((M, M), d, d, WL, NG, stride=4, method=method)
            ref.append(plan.run([torch.from_numpy(a).cuda() for a in _fields(i)[:4]])[0].clone())
        ref = torch.stack(ref)
        same = bool(((P == ref) | (torch.isnan(P) & torch.isnan(ref))).all())
        q.put((rank, same and (gather == "auto" or sh._gather == gather)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items,method,gather", [(2, "fft", "push"), (4, "fold", "nccl"), (1, "fold", "push"),
                                                   (3, "dense", "auto"), (2, "fft", "p2p"), (6, "fft", "graph")])
def test_two_rank_gather_equals_single_gpu(n_items, method, gather):
    """gather: 'push' = mlb_peer_allgather (every rank stores its tiles into the peers' buffers over NVLink), 'p2p' =
    tiles pulled by the copy engines, 'nccl' = mlb_allgather_P (the C-ABI's NCCL wrapper); 'graph' = push with the
    step replayed as CUDA graphs"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, method, gather, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
