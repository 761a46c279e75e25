"""ctypes binding of ``libsepref_b200.so`` (declared in ``include/sepref.h``).

The library is the product; there is deliberately no Python/CPU fallback: if it is missing or cannot be
loaded, importing callers get an ImportError that says how to build it (``python __graft_entry__.py build``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsepref_b200.so")

OPT_GEMM_PATH = 1
OPT_DEBUG_SYNC = 2
OPT_PROFILE = 3
OPT_CLUSTER = 4
OPT_HOST_CHUNK = 5
OPT_GCFN_WIDE = 6
OPT_RAW_F16 = 7
OPT_GCFN_PAIR = 8
OPT_CUDA_GRAPH = 9
OPT_GCFN_TRIO = 10
OPT_CLA_FUSED = 11
OPT_GCFN_TM = 12


class SeprefConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("feat", "heads", "num_stages", "num_spks", "cla_kernel", "down_kernel",
                                         "maxlen", "per_stage_split")]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    # SEPREF_LIB selects another build of the same library (kernel-variant A/B runs, tools/variants.py)
    path = os.environ.get("SEPREF_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python __graft_entry__.py build` "
                          "(nvcc, sm_100a). sepreformer_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(path)
    vp, cp, i, sz = C.c_void_p, C.c_char_p, C.c_int, C.c_size_t
    fp = C.c_void_p          # device/host float pointers travel as integers
    L.sepref_last_error.restype = cp
    L.sepref_version.restype = cp
    L.sepref_create.argtypes = [C.POINTER(SeprefConfig), i, C.POINTER(vp)]
    L.sepref_destroy.argtypes = [vp]
    L.sepref_destroy.restype = None
    L.sepref_set_option.argtypes = [vp, i, i]
    L.sepref_set_param.argtypes = [vp, cp, fp, C.POINTER(C.c_int64), i]
    L.sepref_missing_params.argtypes = [vp, C.POINTER(cp)]
    L.sepref_finalize.argtypes = [vp]
    L.sepref_padded_frames.argtypes = [vp, i]
    L.sepref_workspace_bytes.argtypes = [vp, i, i]
    L.sepref_workspace_bytes.restype = sz
    L.sepref_separator_forward.argtypes = [vp, fp, i, i, fp, C.POINTER(fp), vp, sz, vp]
    L.sepref_separator_forward_host.argtypes = [vp, fp, i, i, fp, C.POINTER(fp), vp]
    L.sepref_separator_submit_host.argtypes = [vp, i, fp, i, i, fp, C.POINTER(fp)]
    L.sepref_separator_wait_host.argtypes = [vp, i]
    L.sepref_last_launch_count.argtypes = [vp]
    L.sepref_model_frames.argtypes = [vp, i]
    L.sepref_model_output_samples.argtypes = [vp, i]
    L.sepref_model_workspace_bytes.argtypes = [vp, i, i]
    L.sepref_model_workspace_bytes.restype = sz
    L.sepref_model_forward.argtypes = [vp, fp, i, i, fp, C.POINTER(fp), vp, sz, vp]
    L.sepref_model_submit_host.argtypes = [vp, i, fp, i, i, fp]
    L.sepref_model_wait_host.argtypes = [vp, i]
    L.sepref_pit_sisnri.argtypes = [vp, fp, fp, fp, i, i, i, C.c_double, fp, vp]
    L.sepref_f16_fallback_count.argtypes = [vp]
    L.sepref_range_rerun_count.argtypes = [vp]
    L.sepref_range_rerun_count.restype = C.c_longlong
    L.sepref_graph_replay_count.argtypes = [vp]
    L.sepref_profile_report.argtypes = [vp, C.c_char_p, sz]
    L.sepref_block_workspace_bytes.argtypes = [vp, i, i]
    L.sepref_block_workspace_bytes.restype = sz
    blk = [vp, cp, fp, i, i, fp, vp, sz, vp]
    for name in ("gcfn", "cla", "local_block", "spk_attention", "spk_split"):
        getattr(L, f"sepref_{name}_forward").argtypes = blk
    L.sepref_ega_forward.argtypes = [vp, cp, fp, i, i, i, fp, vp, sz, vp]
    L.sepref_global_block_forward.argtypes = [vp, cp, fp, i, i, i, fp, vp, sz, vp]
    L.sepref_debug_gcfn_h.argtypes = [vp, cp, fp, i, i, fp, fp, vp]
    L.sepref_debug_gcfn_timeline.argtypes = [vp, cp, fp, i, i, fp, vp, vp]
    L.sepref_debug_tok_timeline.argtypes = [vp, cp, vp]
    L.sepref_down_conv_forward.argtypes = [vp, cp, fp, i, i, fp, vp]
    L.sepref_fusion_forward.argtypes = [vp, cp, fp, fp, i, i, fp, vp, sz, vp]
    _lib = L
    return L


EXPORTS = [
    "sepref_last_error", "sepref_version", "sepref_create", "sepref_destroy", "sepref_set_option",
    "sepref_set_param", "sepref_missing_params", "sepref_finalize", "sepref_padded_frames",
    "sepref_workspace_bytes", "sepref_separator_forward", "sepref_separator_forward_host",
    "sepref_separator_submit_host", "sepref_separator_wait_host",
    "sepref_last_launch_count", "sepref_f16_fallback_count", "sepref_range_rerun_count", "sepref_graph_replay_count", "sepref_model_frames", "sepref_model_output_samples",
    "sepref_model_workspace_bytes", "sepref_model_forward", "sepref_model_submit_host", "sepref_model_wait_host", "sepref_pit_sisnri", "sepref_profile_report", "sepref_block_workspace_bytes", "sepref_gcfn_forward", "sepref_debug_gcfn_h", "sepref_debug_gcfn_timeline", "sepref_debug_tok_timeline", "sepref_cla_forward",
    "sepref_ega_forward", "sepref_global_block_forward", "sepref_local_block_forward",
    "sepref_spk_attention_forward", "sepref_down_conv_forward", "sepref_spk_split_forward",
    "sepref_fusion_forward",
]


def check(rc: int, what: str = "sepref") -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().sepref_last_error().decode(errors='replace')}")
