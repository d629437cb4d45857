"""CPU: the reference arm of bench.py (--impl reference) on a shrunken workload -- the JSON line carries the keys the driver
reads, names the same `config` as the GPU arm, honours --steps / --warmup, and runs the UNMODIFIED reference when
baseline/_ref is installed (the oracle port otherwise)."""
import argparse
import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line(monkeypatch):
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    small = dict(M=128, stride=4, items=[(450e-9, 1.466, False), (532e-9, 1.4607, False), (635e-9, 1.457, True)],
                 name="cfg3 (shrunken for the test)")
    monkeypatch.setitem(bench.WORKLOADS, "cfg3", small)
    lines = []
    monkeypatch.setattr(bench, "emit", lambda line: lines.append(json.loads(json.dumps(line))))
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(workload="cfg3", steps=3, warmup=2, gpus=1, ref_budget=60.0)
    bench.reference_arm(args)
    assert len(lines) == 1
    d = lines[0]
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 3 and d["warmup"] == 2 and d["higher_is_better"] is True
    assert d["config"] == bench.common_config(small, 128, 32, 3)              # the same dict the GPU arm prints
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]
    from baseline import install_ref
    assert cb["kind"] == ("reference" if install_ref.load() is not None else "port")
    assert abs(d["value"] - 3 * 32 * 32 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    # ranks other than 0 do no work and print nothing
    monkeypatch.setenv("RANK", "1")
    bench.reference_arm(args)
    assert len(lines) == 1
