"""CPU: placement of pinned host buffers next to a GPU (metalens_b200/hostmem.py) -- cpulist parsing, the sysfs lookup
on a fake tree, affinity set for the allocation only and restored afterwards, and every failure mode falling back to
a plain allocation."""
import os

import pytest

from metalens_b200 import hostmem


def test_parse_cpulist():
    assert hostmem.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem.parse_cpulist("5") == {5}
    assert hostmem.parse_cpulist("") == set() and hostmem.parse_cpulist("\n") == set()
    assert hostmem.parse_cpulist("0-x") == set() and hostmem.parse_cpulist("a,b") == set()


def test_sysfs_lookup(tmp_path):
    d = tmp_path / "0000:1b:00.0"
    d.mkdir()
    (d / "local_cpulist").write_text("0-1,6\n")
    assert hostmem.gpu_local_cpus(0, sysfs=str(tmp_path), address="0000:1b:00.0") == {0, 1, 6}
    assert hostmem.gpu_local_cpus(0, sysfs=str(tmp_path), address="0000:ff:00.0") is None      # no such device
    (d / "local_cpulist").write_text("\n")
    assert hostmem.gpu_local_cpus(0, sysfs=str(tmp_path), address="0000:1b:00.0") is None      # topology not exported
    assert hostmem.gpu_local_cpus(0) is None or isinstance(hostmem.gpu_local_cpus(0), set)     # no CUDA here: None


def test_affinity_is_scoped_and_restored():
    allowed = os.sched_getaffinity(0)
    if len(allowed) < 2:
        pytest.skip("one CPU")
    some = set(sorted(allowed)[:1])
    with hostmem.near_gpu(0, cpus=some | {10 ** 6}) as used:            # CPUs outside the cpuset are ignored
        assert used == some and os.sched_getaffinity(0) == some
    assert os.sched_getaffinity(0) == allowed
    with hostmem.near_gpu(0, cpus=allowed) as used:                      # the whole cpuset: nothing to do
        assert used is None and os.sched_getaffinity(0) == allowed
    with hostmem.near_gpu(0, cpus={10 ** 6}) as used:                    # no local CPU in the cpuset: nothing to do
        assert used is None and os.sched_getaffinity(0) == allowed
    with pytest.raises(RuntimeError):
        with hostmem.near_gpu(0, cpus=some):
            raise RuntimeError("allocation failed")
    assert os.sched_getaffinity(0) == allowed                            # restored on errors too


def test_pinned_empty_falls_back(monkeypatch):
    import torch
    allowed = os.sched_getaffinity(0)
    seen = {}

    def alloc():
        seen["cpus"] = os.sched_getaffinity(0)
        return torch.empty((2, 3), dtype=torch.complex64)
    # unknown topology (no CUDA device here / no sysfs entry): plain allocation on the unchanged cpuset
    t = hostmem.pinned_empty((2, 3), torch.complex64, 0, _alloc=alloc)
    assert tuple(t.shape) == (2, 3) and seen["cpus"] == allowed
    # known topology: the allocation runs on the local CPUs only
    first = set(sorted(allowed)[:1])
    monkeypatch.setattr(hostmem, "gpu_local_cpus", lambda idx: first)
    hostmem.pinned_empty((2, 3), torch.complex64, 0, _alloc=alloc)
    assert seen["cpus"] == (first if len(allowed) > 1 else allowed) and os.sched_getaffinity(0) == allowed
    # device_index None without CUDA: current_device() raises -> plain allocation
    hostmem.pinned_empty((2, 3), torch.complex64, None, _alloc=alloc)
    assert seen["cpus"] == allowed
