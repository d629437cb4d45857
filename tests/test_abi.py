"""CPU: the C-ABI library loads and exports every symbol include/metalens_b200.h declares
(no compute calls without a GPU), and the ctypes table covers the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "metalens_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlb_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from metalens_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n


def test_ctypes_table_matches_header():
    from metalens_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.mlb_version() == 100
    assert lib.mlb_launch_count() >= 0


def test_descriptor_structs_match_the_library():
    """The ctypes mirrors of mlb_table_pack / mlb_lens_desc have the size the library was compiled with."""
    from metalens_b200 import _lib, nearfield
    out = (ctypes.c_int * 2)()
    assert _lib.load().mlb_struct_sizes(out) == 0
    assert out[0] == ctypes.sizeof(nearfield._PackC) and out[1] == ctypes.sizeof(nearfield._LensC)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from metalens_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MetalensB200Error):
        _lib.load()


def test_no_cuda_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from metalens_b200 import farfield, MetalensB200Error
    with pytest.raises(MetalensB200Error):
        farfield.FarfieldPlan((8, 8), 1e-7, 1e-7, 532e-9, 1.46)
    # the three reference-signature drop-ins: valid arguments, no device -> the package's own error, never a CPU result
    import numpy as np
    import synth_lens
    from metalens_b200 import design, grating, lens_center, nearfield
    x = np.arange(16) * 200e-9
    F = np.zeros((16, 16), dtype=complex)
    with pytest.raises(MetalensB200Error):
        farfield.farfield_from_nearfield(F, F, F, F, x, x, 532e-9, 1.46)
    with pytest.raises(MetalensB200Error):
        farfield.farfield_from_fields(F, F, F, F, x, x, 532e-9, 1.46)
    spec = synth_lens.SMALL_LENS
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periphery, center, _ = design.make_design(collections, spec["source_distance"], spec["radius"], hgs)
    g = np.linspace(-spec["radius"], spec["radius"], 96)
    for fn in (nearfield.build_nearfield, nearfield.build_nearfield_big):
        with pytest.raises(MetalensB200Error):
            fn(0.0, 0.0, -spec["source_distance"], "x", 580e-9, periphery, center, hgs, x_pts=g, y_pts=g)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: the product package must not reference it."""
    pkg = os.path.join(ROOT, "metalens_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_plain_c_client_links_and_runs(tmp_path):
    """The boundary is a C ABI: a C99 program (examples/c_abi_host_only.c) compiles against include/metalens_b200.h with
    gcc -pedantic, links against the shared library and runs its host-only calls (no GPU needed)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "metalens_b200")
    if not os.path.exists(os.path.join(libdir, "libmetalens_b200.so")):
        pytest.skip("library not built")
    exe = str(tmp_path / "c_abi_host_only")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "c_abi_host_only.c"), "-L", libdir, "-lmetalens_b200",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "3375 = 15 15 15" in out and "C-ABI" in out
