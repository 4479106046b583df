"""ctypes binding of libhalva_dpa.so (include/halva_dpa.h).  No CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HALVA_DPA_LIB") or os.path.join(HERE, "libhalva_dpa.so")   # override: A/B builds of the kernels

DPA_BF16, DPA_F32 = 0, 1
DPA_I64, DPA_I32 = 0, 1
FLAG_ROUND_LOGITS, FLAG_NEED_DW, FLAG_DENSE, FLAG_FUSED_GRAD, FLAG_FUSED_NORM, FLAG_LOSS_ONLY, FLAG_GRAD_HINT = 1, 2, 4, 8, 16, 32, 64
ABI_VERSION = 5
STATUS_BAD_SIGN, STATUS_K_MISMATCH, STATUS_K_ZERO, STATUS_BAD_LABEL = 1, 2, 4, 8
EXPORTS = ["dpa_version", "dpa_last_error", "dpa_query_workspace", "dpa_struct_sizes", "dpa_forward", "dpa_backward",
           "dpa_forward_stats", "dpa_forward_finish", "dpa_row_stats_view", "dpa_debug_gemm",
           "dpa_profile_enable", "dpa_profile_collect", "dpa_profile_kind_name", "dpa_set_cta_group",
           "dpa_comm_unique_id", "dpa_comm_init", "dpa_comm_destroy", "dpa_comm_world", "dpa_tmap_cache_stats", "dpa_plan_host_spans", "dpa_stage_host_batch", "dpa_debug_max_clusters"]


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "T", "B_ref", "T_ref", "H", "V", "V_local", "v_offset", "dtype",
                                         "index_dtype", "flags", "chunk_rows")]


class Workspace(C.Structure):
    _fields_ = [("state_bytes", C.c_size_t), ("fwd_bytes", C.c_size_t), ("bwd_bytes", C.c_size_t), ("fused_bytes", C.c_size_t)]


class Args(C.Structure):
    _fields_ = [("shape", Shape), ("alpha", C.c_float),
                ("h_pair", C.c_void_p), ("labels_pair", C.c_void_p), ("signs_pair", C.c_void_p),
                ("h_ref_pol", C.c_void_p), ("h_ref_frz", C.c_void_p), ("labels_ref", C.c_void_p),
                ("W", C.c_void_p), ("W_ref", C.c_void_p),
                ("state", C.c_void_p), ("state_bytes", C.c_size_t),
                ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t),
                ("losses", C.c_void_p), ("status", C.c_void_p),
                ("d_loss", C.c_void_p), ("dH_pair", C.c_void_p), ("dH_ref", C.c_void_p), ("dW", C.c_void_p),
                ("norm_w", C.c_void_p), ("norm_w_ref", C.c_void_p), ("d_norm_w", C.c_void_p), ("norm_eps", C.c_float),
                ("ref_logits", C.c_void_p), ("ref_logits_ld", C.c_int64), ("ref_logits_dtype", C.c_int32)]


class Span(C.Structure):
    _fields_ = [("seq", C.c_int32), ("t0", C.c_int32), ("t1", C.c_int32)]


class HostBatch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "T", "B_ref", "T_ref", "H", "dtype", "index_dtype", "dense", "merge_gap")] + \
               [(n, C.c_void_p) for n in ("h_pair", "labels_pair", "signs_pair", "h_ref_pol", "h_ref_frz", "labels_ref",
                                          "d_h_pair", "d_labels_pair", "d_signs_pair", "d_h_ref_pol", "d_h_ref_frz", "d_labels_ref")]


class StageInfo(C.Structure):
    _fields_ = [("bytes", C.c_int64), ("spans_pair", C.c_int32), ("spans_ref", C.c_int32), ("rows_pair", C.c_int32),
                ("rows_ref", C.c_int32), ("plan_us", C.c_float), ("issue_us", C.c_float)]


class DPAError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DPAError(f"{LIB_PATH} is missing: build it with `python -m halva_b200.build` "
                       "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the DPA-loss path.")
    lib = C.CDLL(LIB_PATH)
    lib.dpa_version.restype = C.c_int
    if lib.dpa_version() != ABI_VERSION:
        raise DPAError(f"{LIB_PATH} has ABI version {lib.dpa_version()}, this package needs {ABI_VERSION}: rebuild it "
                       "(`python -m halva_b200.build --force`)")
    lib.dpa_last_error.restype = C.c_char_p
    sz = [C.c_int(), C.c_int(), C.c_int()]
    lib.dpa_struct_sizes(*[C.byref(x) for x in sz])
    if [x.value for x in sz] != [C.sizeof(Shape), C.sizeof(Workspace), C.sizeof(Args)]:
        raise DPAError(f"{LIB_PATH}: struct sizes {[x.value for x in sz]} differ from this package's ctypes mirror")
    lib.dpa_query_workspace.argtypes = [C.POINTER(Shape), C.POINTER(Workspace)]
    for name in ("dpa_forward", "dpa_backward", "dpa_forward_stats"):
        getattr(lib, name).argtypes = [C.POINTER(Args), C.c_void_p]
    lib.dpa_forward_finish.argtypes = [C.POINTER(Args), C.c_void_p, C.c_int, C.c_void_p]
    lib.dpa_row_stats_view.argtypes = [C.POINTER(Args), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.dpa_debug_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p]
    lib.dpa_plan_host_spans.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Span), C.c_int]
    lib.dpa_stage_host_batch.argtypes = [C.POINTER(HostBatch), C.c_void_p, C.POINTER(StageInfo)]
    lib.dpa_tmap_cache_stats.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.dpa_profile_enable.argtypes = [C.c_int]
    lib.dpa_set_cta_group.argtypes = [C.c_int]
    lib.dpa_comm_unique_id.argtypes = [C.c_void_p]
    lib.dpa_comm_init.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.dpa_comm_world.restype = C.c_int
    lib.dpa_profile_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    lib.dpa_profile_kind_name.argtypes = [C.c_int]
    lib.dpa_profile_kind_name.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().dpa_last_error().decode(errors="replace")
        raise DPAError(f"{what} failed ({rc}): {msg}")


DEFAULT_CHUNK_ROWS = 32768


def default_chunk_rows(n_pos_pair: int = 0, n_pos_ref: int = 0) -> int:
    """Rows per chunk of the packed row space when dpa_shape_t::chunk_rows is 0 (include/halva_dpa.h) for a bf16 batch with
    `n_pos_pair` = 2B(T-1) and `n_pos_ref` = B_ref(T_ref-1) predicting positions: twice the default beyond four default
    chunks of positions (the rule of make_layout in csrc/dpa_api.cu)."""
    pad = lambda n: -(-n // 256) * 256
    return DEFAULT_CHUNK_ROWS * (2 if pad(n_pos_pair) + pad(n_pos_ref) > 4 * DEFAULT_CHUNK_ROWS else 1)


def set_cta_group(cg: int):
    """1 = one CTA per 128 x 256 tile, 2 = CTA pairs on 256 x 256 tiles (tcgen05 cta_group::2; the default)."""
    check(load().dpa_set_cta_group(int(cg)), "dpa_set_cta_group")


def profile_enable(on: bool):
    load().dpa_profile_enable(1 if on else 0)


def profile_collect():
    """-> {kernel class: (milliseconds, kernel launches)} accumulated since the last collect (synchronises)."""
    lib = load()
    ms = (C.c_double * 16)()
    ln = (C.c_int * 16)()
    n = lib.dpa_profile_collect(ms, ln, 16)
    return {lib.dpa_profile_kind_name(k).decode(): (ms[k], ln[k]) for k in range(n)}
