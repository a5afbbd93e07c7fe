"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/voxgraph_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from voxgraph_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from voxgraph_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "voxgraph_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vgx_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    from voxgraph_b200 import _lib
    # field counts / sizes of the plain-C structs crossing the boundary
    assert C.sizeof(_lib.TsdfConfig) == 14 * 4
    assert C.sizeof(_lib.TsdfStats) == 4 * 8
    assert C.sizeof(_lib.RegConfig) == 24
    assert C.sizeof(_lib.SolverOptions) == 8 + 10 * 8 + 8
    assert C.sizeof(_lib.SolverSummary) == 16 + 3 * 8


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.vgx_device_count() == 0
    assert lib.vgx_ctx_create(0, C.byref(h)) == -2   # VGX_ERR_CUDA
    from voxgraph_b200 import api
    with pytest.raises(api.VgxError):
        api.Context(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under voxgraph_b200/ may reference it."""
    pkg = os.path.join(ROOT, "voxgraph_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "vg_oracle" not in txt and "import oracle" not in txt and \
                    "from oracle" not in txt, os.path.join(dirpath, f)
