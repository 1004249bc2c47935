"""The C-ABI library loads on a machine without a GPU and exports every symbol the header declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from critic2_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "critic2_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(c2g_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = capi.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/critic2_gpu.h but not exported"
    for s in capi.EXPORTS:
        assert s in syms


def test_no_torch_types_in_header():
    src = open(os.path.join(ROOT, "include", "critic2_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in src.lower() and "at::" not in src and "std::" not in src


def test_init_fails_loudly_without_gpu():
    """No CPU fallback: creating a context without a usable CUDA device must raise."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.C2GError):
        capi.Context(0)


def test_slab_bounds_cover_the_grid():
    for n3 in (4, 7, 64, 90, 1000, 1024):
        for G in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(G):
                a, b = capi.slab_bounds(n3, G, r)
                assert a == prev and b >= a
                assert a % 4 == 0 or a == n3
                prev = b
            assert prev == n3


def test_product_path_does_not_import_oracle():
    """The shipped package must not route through the CPU oracle."""
    for root, _, files in os.walk(os.path.join(ROOT, "critic2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_host_mirror_library_exports_the_reference_driver_names():
    """critic2_b200/libcritic2_host.so: the C++ mirror of bader_integrate / yt_integrate / intgrid_fields / nciplot
    (csrc/host/critic2_host.hpp) must build and link against the C ABI only."""
    import subprocess
    lib = os.path.join(ROOT, "critic2_b200", "libcritic2_host.so")
    assert os.path.exists(lib), "run __graft_entry__.build()"
    syms = subprocess.run(["nm", "-DC", "--defined-only", lib], capture_output=True, text=True).stdout
    for name in ("c2h::bader_integrate", "c2h::yt_integrate", "c2h::intgrid_fields", "c2h::yt_weights", "c2h::nci_rdg",
                 "c2h::nci_rdg_fourier", "c2h::grid_fft", "c2h::grid_read_text", "c2h::grid_write_text",
                 "c2h::ferror", "c2h::gpu_init", "c2h::system::identify_atom", "c2h::system::are_lclose"):
        assert name in syms, name
    und = subprocess.run(["nm", "-DC", "--undefined-only", lib], capture_output=True, text=True).stdout
    assert "c2g_bader_assign" in und and "orc_" not in und
