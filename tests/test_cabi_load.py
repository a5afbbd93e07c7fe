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


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors vs the C compiler's view of include/voxgraph_b200.h (size and every field offset)."""
    import subprocess
    from voxgraph_b200 import _lib
    structs = {"vgx_tsdf_config": _lib.TsdfConfig, "vgx_tsdf_stats": _lib.TsdfStats,
               "vgx_reg_config": _lib.RegConfig, "vgx_solver_options": _lib.SolverOptions,
               "vgx_solver_summary": _lib.SolverSummary, "vgx_registration_filter": _lib.RegistrationFilter,
               "vgx_esdf_config": _lib.EsdfConfig}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "voxgraph_b200.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE,
                                                 text=True).stdout.strip().splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)
    assert C.sizeof(_lib.TsdfConfig) == 14 * 4 and C.sizeof(_lib.TsdfStats) == 5 * 8


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
