"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, and refuses to run without a GPU instead of falling back to the CPU."""
import ctypes
import os

import pytest


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.capi.lib()
    names = pkg.capi.exported_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.slam_b200_version() == 100


def test_no_cpu_fallback_without_a_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.SlamB200Error):
        pkg.Context(0)


def test_product_does_not_link_or_import_the_oracle(pkg):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pdir = os.path.join(root, "opendlv-logic-cfsd18-sensation-slam_b200")
    for dirpath, _, files in os.walk(pdir):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
                assert "orc_" not in txt, f
