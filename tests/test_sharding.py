"""CPU: far-field tile sharding -- schedule properties and the N > 1 gather path on
the gloo backend with world_size 2 and 4 (no GPU, no kernels: tiles are filled with a known pattern)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metalens_b200.sharding import ShardedFarfield, gather_tiles, tile_schedule


def test_schedule_properties():
    for n_items, rows, world in ((3, 1024, 1), (3, 1024, 2), (3, 1024, 8), (24, 512, 8), (2, 512, 4), (6, 96, 4)):
        sched = tile_schedule(n_items, rows, world)
        assert len(sched) == world and len({len(s) for s in sched}) == 1            # equal work per rank
        flat = [t for s in sched for t in s]
        cover = np.zeros((n_items, rows), int)
        for t in flat:
            cover[t.item, t.row0:t.row1] += 1
        assert (cover == 1).all()                                                    # disjoint, complete
        assert flat == sorted(flat, key=lambda t: (t.item, t.row0))                  # gather order = result order
    assert all(len({t.item for t in s}) == 3 for s in tile_schedule(24, 512, 8))     # whole items when G | B
    with pytest.raises(ValueError):
        tile_schedule(3, 1000, 16)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class _FakePlan:
    """Stands in for FarfieldPlan on CPU: P[i, j] = 1000*item + row + j/1000."""
    def __init__(self, item, row0, row1, cols):
        r = torch.arange(row0, row1, dtype=torch.float32)[:, None]
        c = torch.arange(cols, dtype=torch.float32)[None, :]
        self.P = 1000.0 * item + r + c / 1000.0

    def run(self, fields):
        assert fields == "fields"
        return self.P, torch.tensor([float(self.P.sum())], dtype=torch.float64)


def _worker(rank, world, port, n_items, rows, cols, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = ShardedFarfield(n_items, rows, lambda i, a, b: _FakePlan(i, a, b, cols))
        owned = sh.items_needed
        P, totals = sh.run(lambda item: "fields" if item in owned else None)
        r = torch.arange(rows, dtype=torch.float32)[None, :, None]
        c = torch.arange(cols, dtype=torch.float32)[None, None, :]
        expect = 1000.0 * torch.arange(n_items, dtype=torch.float32)[:, None, None] + r + c / 1000.0
        q.put((rank, bool(torch.equal(P, expect)), len(totals)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 3), (2, 4), (4, 2)])
def test_gather_on_gloo(world, n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, 16, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world)) and all(r[1] for r in res)


def test_gather_single_rank():
    local = torch.arange(2 * 4 * 3, dtype=torch.float32).view(2, 4, 3)
    assert torch.equal(gather_tiles(local, 1, 8, 1), local.view(1, 8, 3))
