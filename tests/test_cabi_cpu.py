"""CPU-side checks of the boundary: the library loads, exports every symbol of include/halva_dpa.h, and the
host-side argument validation answers without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from halva_b200 import _lib, build
    build.build()
    return _lib.load()


def test_exports_match_header():
    lib = _lib()
    hdr = open(os.path.join(ROOT, "include", "halva_dpa.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(dpa_\w+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    from halva_b200._lib import EXPORTS
    assert declared == set(EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    from halva_b200._lib import ABI_VERSION
    assert lib.dpa_version() == ABI_VERSION == 4


def test_ctypes_structs_mirror_the_header():
    from halva_b200._lib import Args, Shape, Workspace
    lib = _lib()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    assert lib.dpa_struct_sizes(C.byref(a), C.byref(b), C.byref(c)) == 0
    assert (a.value, b.value, c.value) == (C.sizeof(Shape), C.sizeof(Workspace), C.sizeof(Args))


def test_workspace_query_and_shape_validation():
    from halva_b200._lib import Shape, Workspace
    lib = _lib()
    ws = Workspace()
    ok = Shape(B=16, T=2048, B_ref=16, T_ref=2048, H=4096, V=32000, V_local=32000, v_offset=0, dtype=0, index_dtype=0,
               flags=3, chunk_rows=0)
    assert lib.dpa_query_workspace(C.byref(ok), C.byref(ws)) == 0
    assert ws.state_bytes > 0 and ws.fwd_bytes > 2 * 16 * 2047 * 4096 * 2 and ws.bwd_bytes > 16384 * 32000 * 2
    bad = Shape(B=1, T=16, B_ref=1, T_ref=16, H=100, V=512, V_local=512, v_offset=0, dtype=0, index_dtype=0, flags=0, chunk_rows=0)
    assert lib.dpa_query_workspace(C.byref(bad), C.byref(ws)) == -1
    assert b"H=100" in lib.dpa_last_error()
    bad = Shape(B=1, T=16, B_ref=1, T_ref=16, H=128, V=512, V_local=512, v_offset=0, dtype=3, index_dtype=0, flags=0, chunk_rows=0)
    assert lib.dpa_query_workspace(C.byref(bad), C.byref(ws)) == -2


def test_python_op_refuses_cpu_tensors():
    import torch
    from halva_b200 import DPAError, dpa_loss
    h = torch.zeros(2, 8, 64, dtype=torch.bfloat16)
    y = torch.zeros(2, 8, dtype=torch.int64)
    with pytest.raises(DPAError):
        dpa_loss(h, y, y, h[:1], y[:1], torch.zeros(64, 64, dtype=torch.bfloat16), h_ref_frozen=h[:1], alpha=0.4)


def test_missing_library_fails_loudly():
    """No CPU / PyTorch fallback: without the built library the product path raises at load time."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['HALVA_DPA_LIB'] = '/nonexistent/libhalva_dpa.so'\n"
            "from halva_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.DPAError as e:\n    print('RAISED', 'no CPU or PyTorch fallback' in str(e))\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.stdout.strip() == "RAISED True", (out.stdout, out.stderr[-500:])
