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
    assert lib.dpa_version() == ABI_VERSION == 5


def test_ctypes_structs_mirror_the_header():
    from halva_b200._lib import Args, Shape, Workspace
    lib = _lib()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    assert lib.dpa_struct_sizes(C.byref(a), C.byref(b), C.byref(c)) == 0
    assert (a.value, b.value, c.value) == (C.sizeof(Shape), C.sizeof(Workspace), C.sizeof(Args))


def test_host_batch_structs_mirror_the_header():
    """dpa_host_batch_t / dpa_stage_info_t / dpa_span_t: field order and sizes as declared in the header."""
    from halva_b200._lib import HostBatch, Span, StageInfo
    assert C.sizeof(Span) == 12 and C.sizeof(HostBatch) == 9 * 4 + 4 + 12 * 8 and C.sizeof(StageInfo) == 8 + 4 * 4 + 2 * 4
    hdr = open(os.path.join(ROOT, "include", "halva_dpa.h")).read()
    body = hdr[hdr.index("typedef struct {\n  int32_t B, T, B_ref, T_ref, H;"):hdr.index("} dpa_host_batch_t;")]
    ptrs = re.findall(r"(?:const )?void\* (\w+)", body)
    assert ptrs == [n for n, _ in HostBatch._fields_[9:]], ptrs
    lib = _lib()
    assert lib.dpa_stage_host_batch(None, None, None) == -7          # argument validation answers without a GPU


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


def test_workspace_query_fp32_inputs_and_flag_conflicts():
    """fp32 hidden states / weights (dtype 1): split-bf16 operands triple the packed buffers and add the split weights;
    the flags that describe the bf16 execution are refused with them."""
    from halva_b200._lib import FLAG_FUSED_GRAD, FLAG_NEED_DW, FLAG_ROUND_LOGITS, Shape, Workspace
    lib = _lib()
    kw = dict(B=1, T=1024, B_ref=1, T_ref=1024, H=4096, V=32000, V_local=32000, v_offset=0, index_dtype=0, chunk_rows=0)
    w16, w32 = Workspace(), Workspace()
    assert lib.dpa_query_workspace(C.byref(Shape(dtype=0, flags=FLAG_NEED_DW, **kw)), C.byref(w16)) == 0
    assert lib.dpa_query_workspace(C.byref(Shape(dtype=1, flags=FLAG_NEED_DW, **kw)), C.byref(w32)) == 0
    assert w32.state_bytes == w16.state_bytes
    assert w32.fwd_bytes >= w16.fwd_bytes + 2 * 3 * 32000 * 4096 * 2          # W and W_ref as three bf16 parts each
    assert w32.bwd_bytes > w16.bwd_bytes
    wf = Workspace()
    assert lib.dpa_query_workspace(C.byref(Shape(dtype=1, flags=FLAG_NEED_DW | FLAG_FUSED_GRAD, **kw)), C.byref(wf)) == 0
    assert wf.fused_bytes > 3 * 32000 * 4096 * 2                              # fused fp32 pass: split weights + kept fp32 logits
    for bad in (FLAG_ROUND_LOGITS, FLAG_FUSED_GRAD | FLAG_ROUND_LOGITS):
        assert lib.dpa_query_workspace(C.byref(Shape(dtype=1, flags=bad, **kw)), C.byref(w32)) == -7
        assert b"fp32 inputs" in lib.dpa_last_error()


def test_workspace_follows_the_chunk_and_unit_rules(monkeypatch):
    """Host-side layout rules, visible through dpa_query_workspace without a GPU: the default row chunk doubles for batches of
    more than 131072 positions (halva_b200._lib.default_chunk_rows states the same rule), and a small vocabulary shard takes
    single-tile work units (one partial-statistics record per row and 256 columns instead of per 1280)."""
    from halva_b200._lib import FLAG_FUSED_GRAD, FLAG_NEED_DW, FLAG_ROUND_LOGITS, Shape, Workspace, default_chunk_rows
    lib = _lib()
    monkeypatch.delenv("DPA_TPU", raising=False)
    fl = FLAG_FUSED_GRAD | FLAG_NEED_DW | FLAG_ROUND_LOGITS

    def fused(chunk_rows=0, **over):
        kw = dict(B=16, T=2048, B_ref=16, T_ref=2048, H=4096, V=32000, V_local=32000, v_offset=0, dtype=0, index_dtype=0, flags=fl)
        kw.update(over)
        ws = Workspace()
        assert lib.dpa_query_workspace(C.byref(Shape(chunk_rows=chunk_rows, **kw)), C.byref(ws)) == 0, lib.dpa_last_error()
        return ws
    # cfg-2 (98 K positions): 32768-row chunks; cfg-4 (196 K positions): 65536
    assert default_chunk_rows(32 * 2047, 16 * 2047) == 32768 and default_chunk_rows(32 * 4095, 16 * 4095) == 65536
    assert fused().fused_bytes == fused(32768).fused_bytes < fused(65536).fused_bytes
    big = dict(T=4096, T_ref=4096, H=5120)
    assert fused(**big).fused_bytes == fused(65536, **big).fused_bytes > fused(32768, **big).fused_bytes
    # the boundary itself: 4 x 32768 positions (padded to 256 per segment) still take the small chunk
    edge = dict(B=16, T=2049, B_ref=32, T_ref=2049)                    # 32 * 2048 + 32 * 2048 = 131072
    assert default_chunk_rows(32 * 2048, 32 * 2048) == 32768 and fused(**edge).fused_bytes == fused(32768, **edge).fused_bytes
    # statistics partials of a V / 8 shard: 16 column groups of one tile (forced groups of five: 4)
    shard = dict(H=5120, V=32000, V_local=4000, v_offset=4000, flags=FLAG_NEED_DW)
    one = fused(**shard).fwd_bytes
    monkeypatch.setenv("DPA_TPU", "5")
    five = fused(**shard).fwd_bytes
    monkeypatch.setenv("DPA_TPU", "1")
    assert fused(**shard).fwd_bytes == one > five
    cap_pair, cap_ref = -(-32 * 2047 // 256) * 256, -(-16 * 2047 // 256) * 256
    assert one - five == (16 - 4) * (cap_pair * 16 + cap_ref * 16 + cap_ref * 4)
    # the full vocabulary keeps groups of five: 25 groups
    monkeypatch.delenv("DPA_TPU")
    full = fused(flags=FLAG_NEED_DW).fwd_bytes
    monkeypatch.setenv("DPA_TPU", "5")
    assert fused(flags=FLAG_NEED_DW).fwd_bytes == full


def test_python_op_argument_validation_without_gpu():
    """The two forms of the frozen branch are exclusive, and ref_logits belongs to the chunked bf16 pass."""
    import torch
    from halva_b200 import dpa_loss
    h = torch.zeros(2, 8, 64, dtype=torch.bfloat16)
    y = torch.zeros(2, 8, dtype=torch.int64)
    W = torch.zeros(64, 64, dtype=torch.bfloat16)
    z = torch.zeros(1, 8, 64, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="exactly one"):
        dpa_loss(h, y, y, h[:1], y[:1], W, alpha=0.4)
    with pytest.raises(ValueError, match="exactly one"):
        dpa_loss(h, y, y, h[:1], y[:1], W, h_ref_frozen=h[:1], ref_logits=z, alpha=0.4)
    with pytest.raises(ValueError, match="ref_logits needs"):
        dpa_loss(h, y, y, h[:1], y[:1], W, ref_logits=z, alpha=0.4, round_logits=False)
    with pytest.raises(ValueError, match="ref_logits must be"):
        dpa_loss(h, y, y, h[:1], y[:1], W, ref_logits=z[:, :4], alpha=0.4)
    for bad in (0.0, float("nan"), float("inf")):
        with pytest.raises(ValueError, match="grad_scale_hint"):
            dpa_loss(h, y, y, h[:1], y[:1], W, h_ref_frozen=h[:1], alpha=0.4, grad_scale_hint=bad)
    with pytest.raises(TypeError):                        # mixed precisions
        dpa_loss(h.float(), y, y, h[:1], y[:1], W, h_ref_frozen=h[:1], alpha=0.4)
    with pytest.raises(ValueError, match="fp32 inputs"):
        dpa_loss(h.float(), y, y, h[:1].float(), y[:1], W.float(), h_ref_frozen=h[:1].float(), alpha=0.4, norm_weight=W[0])


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


def test_gemm_kernels_are_tcgen05_tma_in_sass():
    """Static check of the built library (cuobjdump, no GPU): every GEMM kernel of the path issues tcgen05.mma (UTCHMMA; the
    CTA-pair configurations as .2CTA), TMA tensor loads (UTMALDG) and tcgen05.ld (LDTM) -- and none falls back to the legacy
    mma.sync path (HMMA).  /opt/skills/guides/B200_PROFILING.md names these mnemonics as the proof."""
    import shutil
    import subprocess
    _lib()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("no cuobjdump here")
    lib = os.path.join(ROOT, "halva_b200", "libhalva_dpa.so")
    sass = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, timeout=300).stdout
    blocks = [b for b in sass.split("Function : ")[1:] if "dpa_gemm" in b.split()[0]]
    assert len(blocks) >= 20, len(blocks)                      # every (tile, operand majors, cta_group, epilogue) instantiation
    pairs = 0
    for b in blocks:
        name = b.split()[0]
        assert re.search(r"\bUTCHMMA\b", b) and re.search(r"\bUTMALDG\b", b) and re.search(r"\bLDTM\b", b), name
        assert not re.search(r"\bHMMA\b", b), name
        pairs += bool(re.search(r"UTCHMMA\.2CTA", b))
    assert pairs >= len(blocks) // 2                           # the default configurations are CTA pairs (cta_group::2)
