"""CPU, world_size 2 and 4 on gloo: the one-aperture-over-ranks decomposition of metalens_b200/slab.py with REAL
processes and a real exchange -- every rank folds and row-transforms only its aperture rows (numpy stands in for the
kernels), the column slabs travel with all_to_all, every rank column-transforms its slab, the P slabs are all-gathered.
The assembled map must equal fftshift(fft2(fftshift(J)))[::s, ::s] (nearfield_farfield.py:18-20, :68) on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metalens_b200.slab import slab_geometry


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, M, s, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        J = rng.standard_normal((M, M)) + 1j * rng.standard_normal((M, M))      # every rank could build any row; it uses its own
        g = slab_geometry(M, M, s, s, rank, world)
        K, n, cpr = g["K1"], g["rows_per_rank"], g["cols_per_rank"]
        local = J[g["x_rows"]]                                                   # the rows this rank "assembles"
        p = np.arange(K)
        rows = np.zeros((n, K), complex)
        for r in range(n):
            folded = np.zeros(K, complex)
            for t1 in range(s):
                for t2 in range(s):
                    folded += local[r + t1 * n][((p - g["roll_c"]) % K) + t2 * K]
            rows[r] = np.roll(np.fft.fft(folded), g["out_roll_rows"])
        # all-to-all: column slab `peer` of my rows goes to rank `peer` (what mlb_fft_rows_scatter stores over NVLink)
        send = torch.from_numpy(np.ascontiguousarray(np.stack([rows[:, c * cpr:(c + 1) * cpr] for c in range(world)])))
        send = torch.view_as_real(send).contiguous()
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        recv = torch.view_as_complex(recv).numpy()                               # [src rank][its rows][my columns]
        W = np.zeros((K, cpr), complex)
        for src in range(world):
            gs = slab_geometry(M, M, s, s, src, world)
            for r in range(n):
                R = (gs["out_row0"] + (r // gs["row_block"]) * gs["row_stride"] + r % gs["row_block"]) % K
                W[R] = recv[src, r]
        F_slab = np.roll(np.fft.fft(W, axis=0), g["out_roll_cols"], axis=0)
        P_slab = torch.from_numpy(np.abs(F_slab) ** 2)
        gathered = [torch.empty_like(P_slab) for _ in range(world)]
        dist.all_gather(gathered, P_slab)                                        # the ONE all-gather at the end
        P = np.concatenate([t.numpy() for t in gathered], axis=1)
        ref = np.abs(np.fft.fftshift(np.fft.fft2(np.fft.fftshift(J)))[::s, ::s]) ** 2
        q.put((rank, bool(np.allclose(P, ref, rtol=1e-10, atol=1e-12 * ref.max()))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,M,s", [(2, 64, 4), (4, 128, 2), (2, 96, 1)])
def test_slab_decomposition_across_processes(world, M, s):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, s, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world)) and all(r[1] for r in res)
